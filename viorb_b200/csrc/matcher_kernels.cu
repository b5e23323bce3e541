/*
 * matcher_kernels.cu -- hand-written sm_100a kernels for the Hamming matchers of ORBmatcher
 * (reference: src/ORBmatcher.cc) and Frame::ComputeStereoMatches (src/Frame.cc:646-820).
 *
 * 256-bit Hamming distance = 8 x (XOR + POPC) on 32-bit words (DescriptorDistance :1648-1664 computes the
 * same value with a SWAR bit hack).  No tensor cores: this is integer/popc-pipe work.
 */
#include "matcher_kernels.cuh"

namespace {

__device__ __forceinline__ int hamming256(const uint4& a0, const uint4& a1, const uint4& b0, const uint4& b1) {
    return __popc(a0.x ^ b0.x) + __popc(a0.y ^ b0.y) + __popc(a0.z ^ b0.z) + __popc(a0.w ^ b0.w) +
           __popc(a1.x ^ b1.x) + __popc(a1.y ^ b1.y) + __popc(a1.z ^ b1.z) + __popc(a1.w ^ b1.w);
}

/* ------------------------------------------------------------------------------------------------
 * Brute-force top-2 (config 5).  Query-stationary: each thread keeps one query descriptor in 8
 * registers and its running (d1,i1,d2,i2); the CTA streams a slice of the map through shared memory
 * in tiles that are read back as warp-wide broadcasts (2 x LDS.128 per map descriptor per warp).
 * grid = (map slices, query tiles).  Each (slice, query) pair emits a partial record; a second
 * kernel merges the slices with the total order (distance, index), which is exactly what the
 * reference's sequential strict-< scan (src/ORBmatcher.cc:201-226) produces.
 * ---------------------------------------------------------------------------------------------- */
#define T2_THREADS 128
#define T2_TILE 256          /* map descriptors per shared-memory tile (8 KB) */

__global__ void __launch_bounds__(T2_THREADS) hamming_top2_kernel(const uint4* __restrict__ q, int Q,
                                                                  const uint4* __restrict__ map, long long M,
                                                                  long long indexBase, long long sliceLen,
                                                                  viorb_top2* __restrict__ partials) {
    __shared__ uint4 tile[2][T2_TILE * 2];
    const int tid = threadIdx.x;
    const int qi = blockIdx.y * T2_THREADS + tid;
    const long long m0 = (long long)blockIdx.x * sliceLen;
    const long long m1 = min(m0 + sliceLen, M);
    uint4 qa = make_uint4(0, 0, 0, 0), qb = qa;
    if (qi < Q) { qa = q[2 * qi]; qb = q[2 * qi + 1]; }
    int d1 = 256, d2 = 256, i1 = -1, i2 = -1;
    const long long ntiles = (m1 - m0 + T2_TILE - 1) / T2_TILE;
    /* prologue: stage tile 0 */
    auto stage = [&](int buf, long long t) {
        const long long base = m0 + t * T2_TILE;
        const long long lim = (m1 - base) * 2;      /* uint4 elements available */
#pragma unroll
        for (int k = 0; k < (T2_TILE * 2) / T2_THREADS; k++) {
            const int e = tid + k * T2_THREADS;
            uint4 v = make_uint4(0, 0, 0, 0);
            if (e < lim) v = __ldg(&map[base * 2 + e]);
            tile[buf][e] = v;
        }
    };
    if (ntiles > 0) stage(0, 0);
    __syncthreads();
    for (long long t = 0; t < ntiles; t++) {
        const int buf = (int)(t & 1);
        if (t + 1 < ntiles) stage(buf ^ 1, t + 1);
        const long long base = m0 + t * T2_TILE;
        const int cnt = (int)min((long long)T2_TILE, m1 - base);
        const int gbase = (int)(indexBase + base);
#pragma unroll 4
        for (int j = 0; j < cnt; j++) {
            const uint4 ma = tile[buf][2 * j], mb = tile[buf][2 * j + 1];
            const int d = hamming256(qa, qb, ma, mb);
            if (d < d2) {
                if (d < d1) { d2 = d1; i2 = i1; d1 = d; i1 = gbase + j; }
                else { d2 = d; i2 = gbase + j; }
            }
        }
        __syncthreads();
    }
    if (qi < Q) {
        viorb_top2 r;
        r.d1 = d1; r.i1 = i1; r.d2 = d2; r.i2 = i2;
        partials[(size_t)blockIdx.x * Q + qi] = r;
    }
}

/* merge of per-part records: two smallest (d, i) pairs under lexicographic order */
__global__ void top2_merge_kernel(const viorb_top2* __restrict__ parts, int nparts, int Q, viorb_top2* __restrict__ out) {
    const int qi = blockIdx.x * blockDim.x + threadIdx.x;
    if (qi >= Q) return;
    int d1 = 256, d2 = 256, i1 = 0x7fffffff, i2 = 0x7fffffff;
    for (int p = 0; p < nparts; p++) {
        const viorb_top2 r = parts[(size_t)p * Q + qi];
#pragma unroll
        for (int h = 0; h < 2; h++) {
            const int d = h ? r.d2 : r.d1, i = h ? r.i2 : r.i1;
            if (i < 0) continue;
            if (d < d1 || (d == d1 && i < i1)) { d2 = d1; i2 = i1; d1 = d; i1 = i; }
            else if (d < d2 || (d == d2 && i < i2)) { d2 = d; i2 = i; }
        }
    }
    viorb_top2 r;
    r.d1 = d1; r.i1 = i1 == 0x7fffffff ? -1 : i1;
    r.d2 = d2; r.i2 = i2 == 0x7fffffff ? -1 : i2;
    out[qi] = r;
}

__global__ void descriptor_distance_kernel(const uint4* __restrict__ a, const uint4* __restrict__ b, int n,
                                           int32_t* __restrict__ dist) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    dist[i] = hamming256(a[2 * i], a[2 * i + 1], b[2 * i], b[2 * i + 1]);
}

}  // namespace

int viorb_top2_slices(int Q, int64_t M, int sms) {
    const int qtiles = (Q + T2_THREADS - 1) / T2_THREADS;
    int target = (sms * 8 + qtiles - 1) / qtiles;           /* ~8 CTAs per SM in total */
    int64_t maxSlices = (M + T2_TILE - 1) / T2_TILE;
    if (maxSlices < 1) maxSlices = 1;
    if (target > maxSlices) target = (int)maxSlices;
    if (target < 1) target = 1;
    return target;
}

int viorb_launch_hamming_top2(const uint8_t* d_q, int Q, const uint8_t* d_map, int64_t M, int64_t indexBase,
                              viorb_top2* d_partials, int nslices, viorb_top2* d_out, cudaStream_t s) {
    if (Q <= 0) return 0;
    long long sliceLen = (M + nslices - 1) / nslices;
    sliceLen = (sliceLen + T2_TILE - 1) / T2_TILE * T2_TILE;
    if (sliceLen < T2_TILE) sliceLen = T2_TILE;
    dim3 grid(nslices, (Q + T2_THREADS - 1) / T2_THREADS);
    hamming_top2_kernel<<<grid, T2_THREADS, 0, s>>>((const uint4*)d_q, Q, (const uint4*)d_map, M, indexBase, sliceLen,
                                                    d_partials);
    top2_merge_kernel<<<(Q + 127) / 128, 128, 0, s>>>(d_partials, nslices, Q, d_out);
    return 2;
}

int viorb_launch_top2_merge(const viorb_top2* d_parts, int nparts, int Q, viorb_top2* d_out, cudaStream_t s) {
    if (Q <= 0) return 0;
    top2_merge_kernel<<<(Q + 127) / 128, 128, 0, s>>>(d_parts, nparts, Q, d_out);
    return 1;
}

int viorb_launch_descriptor_distance(const uint8_t* d_a, const uint8_t* d_b, int n, int32_t* d_dist, cudaStream_t s) {
    if (n <= 0) return 0;
    descriptor_distance_kernel<<<(n + 255) / 256, 256, 0, s>>>((const uint4*)d_a, (const uint4*)d_b, n, d_dist);
    return 1;
}
