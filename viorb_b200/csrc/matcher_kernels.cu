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

/* carry-save adder on bit planes: a + b + c = sum + 2 * carry   (two LOP3) */
__device__ __forceinline__ void csa(unsigned a, unsigned b, unsigned c, unsigned& sum, unsigned& carry) {
    sum = a ^ b ^ c;
    carry = (a & b) | (a & c) | (b & c);
}

/* The same 256-bit Hamming distance with Harley-Seal compression: three carry-save adders fold the eight
 * XOR words into 2 words of weight 1 and 3 of weight 2, so 5 POPC (quarter-rate pipe) replace 8 at the price
 * of 6 LOP3 on the full-rate ALU pipe.  Exact: popc(x)+popc(y)+popc(z) = popc(sum) + 2 popc(carry). */
__device__ __forceinline__ int hamming256_hs(const uint4& a0, const uint4& a1, const uint4& b0, const uint4& b1) {
    unsigned s0, c0, s1, c1, s2, c2;
    csa(a0.x ^ b0.x, a0.y ^ b0.y, a0.z ^ b0.z, s0, c0);
    csa(a0.w ^ b0.w, a1.x ^ b1.x, a1.y ^ b1.y, s1, c1);
    csa(s0, s1, a1.z ^ b1.z, s2, c2);
    return __popc(s2) + __popc(a1.w ^ b1.w) + 2 * (__popc(c0) + __popc(c1) + __popc(c2));
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
            const int d = hamming256_hs(qa, qb, ma, mb);
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

/* ------------------------------------------------------------------------------------------------
 * Small query counts (Q <= 8): map-stationary streaming scan.  Every thread reads whole map descriptors with
 * two coalesced LDG.128 (the map is read exactly once: this is the HBM-bound regime), compares them against
 * the QT queries broadcast from shared memory and keeps a private running top-2 per query; lanes see
 * increasing indices, so strict < keeps the lowest index and the final merge is by (distance, index).
 * ---------------------------------------------------------------------------------------------- */
#define SQ_THREADS 256

template <int QT>
__global__ void __launch_bounds__(SQ_THREADS) hamming_top2_smallq_kernel(const uint4* __restrict__ q, int Q,
                                                                        const uint4* __restrict__ map, long long M,
                                                                        long long indexBase,
                                                                        viorb_top2* __restrict__ partials) {
    __shared__ uint4 sq[QT * 2];
    __shared__ viorb_top2 red[SQ_THREADS / 32][QT];
    const int tid = threadIdx.x;
    if (tid < QT * 2) sq[tid] = (tid >> 1) < Q ? q[tid] : make_uint4(0, 0, 0, 0);
    __syncthreads();
    int d1[QT], d2[QT], i1[QT], i2[QT];
#pragma unroll
    for (int k = 0; k < QT; k++) { d1[k] = 256; d2[k] = 256; i1[k] = 0x7fffffff; i2[k] = 0x7fffffff; }
    const long long stride = (long long)gridDim.x * SQ_THREADS;
    long long m = (long long)blockIdx.x * SQ_THREADS + tid;
    /* two descriptors in flight per thread */
    for (; m + stride < M; m += 2 * stride) {
        const uint4 a0 = __ldcs(&map[2 * m]), a1 = __ldcs(&map[2 * m + 1]);
        const uint4 b0 = __ldcs(&map[2 * (m + stride)]), b1 = __ldcs(&map[2 * (m + stride) + 1]);
        const int ga = (int)(indexBase + m), gb = (int)(indexBase + m + stride);
#pragma unroll
        for (int k = 0; k < QT; k++) {
            const int da = hamming256_hs(sq[2 * k], sq[2 * k + 1], a0, a1);
            if (da < d2[k]) {
                if (da < d1[k]) { d2[k] = d1[k]; i2[k] = i1[k]; d1[k] = da; i1[k] = ga; }
                else { d2[k] = da; i2[k] = ga; }
            }
            const int db = hamming256_hs(sq[2 * k], sq[2 * k + 1], b0, b1);
            if (db < d2[k]) {
                if (db < d1[k]) { d2[k] = d1[k]; i2[k] = i1[k]; d1[k] = db; i1[k] = gb; }
                else { d2[k] = db; i2[k] = gb; }
            }
        }
    }
    for (; m < M; m += stride) {
        const uint4 a0 = __ldcs(&map[2 * m]), a1 = __ldcs(&map[2 * m + 1]);
        const int ga = (int)(indexBase + m);
#pragma unroll
        for (int k = 0; k < QT; k++) {
            const int da = hamming256_hs(sq[2 * k], sq[2 * k + 1], a0, a1);
            if (da < d2[k]) {
                if (da < d1[k]) { d2[k] = d1[k]; i2[k] = i1[k]; d1[k] = da; i1[k] = ga; }
                else { d2[k] = da; i2[k] = ga; }
            }
        }
    }
    /* merge the private records: warp shuffles, then across warps through shared memory */
    const int lane = tid & 31, warp = tid >> 5;
#pragma unroll
    for (int k = 0; k < QT; k++) {
        unsigned long long k1 = ((unsigned long long)d1[k] << 32) | (unsigned)i1[k];
        unsigned long long k2 = ((unsigned long long)d2[k] << 32) | (unsigned)i2[k];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const unsigned long long b1 = __shfl_xor_sync(0xffffffffu, k1, o), b2 = __shfl_xor_sync(0xffffffffu, k2, o);
            const unsigned long long lo = k1 < b1 ? k1 : b1, hi = k1 < b1 ? b1 : k1, s2 = k2 < b2 ? k2 : b2;
            k1 = lo;
            k2 = hi < s2 ? hi : s2;
        }
        if (lane == 0) {
            viorb_top2 r;
            r.d1 = (int)(k1 >> 32); r.i1 = (int)(k1 & 0xffffffffu);
            r.d2 = (int)(k2 >> 32); r.i2 = (int)(k2 & 0xffffffffu);
            red[warp][k] = r;
        }
    }
    __syncthreads();
    if (tid < Q && tid < QT) {
        int e1 = 256, e2 = 256, j1 = 0x7fffffff, j2 = 0x7fffffff;
        for (int w = 0; w < SQ_THREADS / 32; w++) {
            const viorb_top2 r = red[w][tid];
#pragma unroll
            for (int h = 0; h < 2; h++) {
                const int d = h ? r.d2 : r.d1, i = h ? r.i2 : r.i1;
                if (d < e1 || (d == e1 && i < j1)) { e2 = e1; j2 = j1; e1 = d; j1 = i; }
                else if (d < e2 || (d == e2 && i < j2)) { e2 = d; j2 = i; }
            }
        }
        viorb_top2 r;
        r.d1 = e1; r.i1 = j1 == 0x7fffffff ? -1 : j1;
        r.d2 = e2; r.i2 = j2 == 0x7fffffff ? -1 : j2;
        partials[(size_t)blockIdx.x * Q + tid] = r;
    }
}

/* merge of per-part records: the two smallest (d, i) pairs under lexicographic order; one warp per query,
 * lanes stride over the parts, then a shuffle two-min */
__global__ void __launch_bounds__(128) top2_merge_kernel(const viorb_top2* __restrict__ parts, int nparts, int Q,
                                                         viorb_top2* __restrict__ out) {
    const int lane = threadIdx.x & 31;
    const int qi = blockIdx.x * 4 + (threadIdx.x >> 5);
    if (qi >= Q) return;
    const unsigned long long NONE = (256ull << 32) | 0x7fffffffull;
    unsigned long long k1 = NONE, k2 = NONE;
    for (int p = lane; p < nparts; p += 32) {
        const int4 r = __ldg(reinterpret_cast<const int4*>(&parts[(size_t)p * Q + qi]));
        const unsigned long long a = r.y >= 0 ? ((unsigned long long)r.x << 32) | (unsigned)r.y : NONE;
        const unsigned long long b = r.w >= 0 ? ((unsigned long long)r.z << 32) | (unsigned)r.w : NONE;
        if (a < k1) { k2 = k1; k1 = a; } else if (a < k2) k2 = a;
        if (b < k1) { k2 = k1; k1 = b; } else if (b < k2) k2 = b;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const unsigned long long b1 = __shfl_xor_sync(0xffffffffu, k1, o), b2 = __shfl_xor_sync(0xffffffffu, k2, o);
        const unsigned long long lo = k1 < b1 ? k1 : b1, hi = k1 < b1 ? b1 : k1, s2 = k2 < b2 ? k2 : b2;
        k1 = lo;
        k2 = hi < s2 ? hi : s2;
    }
    if (lane == 0) {
        viorb_top2 r;
        r.d1 = (int)(k1 >> 32); r.i1 = k1 == NONE ? -1 : (int)(k1 & 0xffffffffu);
        r.d2 = (int)(k2 >> 32); r.i2 = k2 == NONE ? -1 : (int)(k2 & 0xffffffffu);
        if (k1 == NONE) r.d1 = 256;
        if (k2 == NONE) r.d2 = 256;
        out[qi] = r;
    }
}

__global__ void descriptor_distance_kernel(const uint4* __restrict__ a, const uint4* __restrict__ b, int n,
                                           int32_t* __restrict__ dist) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    dist[i] = hamming256(a[2 * i], a[2 * i + 1], b[2 * i], b[2 * i + 1]);
}

}  // namespace

int viorb_top2_slices(int Q, int64_t M, int sms) {
    if (Q <= 8) {                                     /* streaming kernel: a few CTAs per SM, at most one per 256 descriptors */
        int64_t n = (M + SQ_THREADS - 1) / SQ_THREADS;
        if (n > (int64_t)sms * 8) n = (int64_t)sms * 8;
        return n < 1 ? 1 : (int)n;
    }
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
    if (Q <= 8) {
        /* streaming scan: nslices CTAs grid-stride over the map */
        const uint4* q4 = (const uint4*)d_q;
        const uint4* m4 = (const uint4*)d_map;
        if (Q == 1) hamming_top2_smallq_kernel<1><<<nslices, SQ_THREADS, 0, s>>>(q4, Q, m4, M, indexBase, d_partials);
        else if (Q == 2) hamming_top2_smallq_kernel<2><<<nslices, SQ_THREADS, 0, s>>>(q4, Q, m4, M, indexBase, d_partials);
        else if (Q <= 4) hamming_top2_smallq_kernel<4><<<nslices, SQ_THREADS, 0, s>>>(q4, Q, m4, M, indexBase, d_partials);
        else hamming_top2_smallq_kernel<8><<<nslices, SQ_THREADS, 0, s>>>(q4, Q, m4, M, indexBase, d_partials);
        top2_merge_kernel<<<(Q + 3) / 4, 128, 0, s>>>(d_partials, nslices, Q, d_out);
        return 2;
    }
    long long sliceLen = (M + nslices - 1) / nslices;
    sliceLen = (sliceLen + T2_TILE - 1) / T2_TILE * T2_TILE;
    if (sliceLen < T2_TILE) sliceLen = T2_TILE;
    dim3 grid(nslices, (Q + T2_THREADS - 1) / T2_THREADS);
    hamming_top2_kernel<<<grid, T2_THREADS, 0, s>>>((const uint4*)d_q, Q, (const uint4*)d_map, M, indexBase, sliceLen,
                                                    d_partials);
    top2_merge_kernel<<<(Q + 3) / 4, 128, 0, s>>>(d_partials, nslices, Q, d_out);
    return 2;
}

int viorb_launch_top2_merge(const viorb_top2* d_parts, int nparts, int Q, viorb_top2* d_out, cudaStream_t s) {
    if (Q <= 0) return 0;
    top2_merge_kernel<<<(Q + 3) / 4, 128, 0, s>>>(d_parts, nparts, Q, d_out);
    return 1;
}

int viorb_launch_descriptor_distance(const uint8_t* d_a, const uint8_t* d_b, int n, int32_t* d_dist, cudaStream_t s) {
    if (n <= 0) return 0;
    descriptor_distance_kernel<<<(n + 255) / 256, 256, 0, s>>>((const uint4*)d_a, (const uint4*)d_b, n, d_dist);
    return 1;
}
