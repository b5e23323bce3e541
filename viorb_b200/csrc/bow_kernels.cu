/*
 * bow_kernels.cu -- DBoW2 vocabulary-tree transform of ORB descriptors on sm_100a:
 *
 *   Frame::ComputeBoW / KeyFrame::ComputeBoW                 src/Frame.cc:575-582, src/KeyFrame.cc:350-359
 *   TemplatedVocabulary::transform(features, v, fv, levelsup) Thirdparty/DBoW2/DBoW2/TemplatedVocabulary.h:1138-1204
 *   TemplatedVocabulary::transform(feature, id, w, nid, ...)  :1230-1272   (tree descent)
 *   FORB::distance                                            Thirdparty/DBoW2/DBoW2/FORB.cpp:81-101
 *   BowVector::addWeight / addIfNotExist / normalize          Thirdparty/DBoW2/DBoW2/BowVector.cpp:34-86
 *   FeatureVector::addFeature                                 Thirdparty/DBoW2/DBoW2/FeatureVector.cpp:31-45
 *
 * bow_descend_kernel: one warp per descriptor; the children of the current node are stored contiguously
 * (32 B each), lane c takes child c, the warp minimum of (distance, child order) is the reference's
 * "first child with the least distance" (strict <, :1252).  bow_reduce_kernel (one CTA): the std::map semantics
 * of BowVector / FeatureVector become a shared-memory bitonic sort of (id, feature index) keys; every word's
 * weight is accumulated in feature order with sequential double adds and the L1/L2 norm is summed in ascending
 * word order by one thread, so every double equals the reference's bit for bit.  Integer/popc work, no tensor cores.
 */
#include <cuda_runtime.h>

#include <algorithm>
#include <cstring>
#include <new>
#include <vector>

#include "viorb_gpu.h"
#include "viorb_internal.h"

namespace {

struct VocabDev {
    const uint8_t* cdesc;     /* descriptors of all non-root nodes, children of a node contiguous */
    const int* childPtr;      /* [nnodes+1] first child slot of node i */
    const int* childIdx;      /* [nnodes-1] node id in each child slot */
    const double* weight;     /* [nnodes] */
    const int* wordId;        /* [nnodes] word id of a leaf, -1 otherwise */
};

__device__ __forceinline__ unsigned long long warp_min_key(unsigned long long v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const unsigned long long t = __shfl_xor_sync(0xffffffffu, v, o);
        v = t < v ? t : v;
    }
    return v;
}

__global__ void __launch_bounds__(256) bow_descend_kernel(VocabDev v, const uint8_t* __restrict__ desc, int n, int nidLevel,
                                                          int* __restrict__ word, int* __restrict__ nid,
                                                          double* __restrict__ w) {
    const int f = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (f >= n) return;
    const uint4 q0 = __ldg(reinterpret_cast<const uint4*>(desc + (size_t)f * 32)), q1 = __ldg(reinterpret_cast<const uint4*>(desc + (size_t)f * 32) + 1);
    int node = 0, level = 0, nidv = 0;                       /* nid_level <= 0: root (:1240) */
    int beg = v.childPtr[0], cnt = v.childPtr[1] - beg;
    while (cnt > 0) {                                        /* do { } while(!isLeaf()): the root of a non-empty tree has children */
        ++level;
        unsigned long long best = ~0ull;
        for (int c = lane; c < cnt; c += 32) {
            const uint4* d = reinterpret_cast<const uint4*>(v.cdesc + (size_t)(beg + c) * 32);
            const uint4 a = __ldg(d), b = __ldg(d + 1);
            const unsigned dist = __popc(a.x ^ q0.x) + __popc(a.y ^ q0.y) + __popc(a.z ^ q0.z) + __popc(a.w ^ q0.w) +
                                  __popc(b.x ^ q1.x) + __popc(b.y ^ q1.y) + __popc(b.z ^ q1.z) + __popc(b.w ^ q1.w);
            const unsigned long long key = ((unsigned long long)dist << 32) | (unsigned)c;
            best = key < best ? key : best;
        }
        best = warp_min_key(best);
        node = v.childIdx[beg + (int)(best & 0xffffffffu)];
        if (level == nidLevel) nidv = node;
        beg = v.childPtr[node];
        cnt = v.childPtr[node + 1] - beg;
    }
    if (lane == 0) {
        word[f] = v.wordId[node];
        w[f] = v.weight[node];
        nid[f] = nidv;
    }
}

/* ---- one-CTA map building -------------------------------------------------------------------- */
#define BOW_THREADS 1024

__device__ void bitonic_sort_u64(unsigned long long* keys, int n2) {
    for (int k = 2; k <= n2; k <<= 1)
        for (int j = k >> 1; j > 0; j >>= 1) {
            __syncthreads();
            for (int i = threadIdx.x; i < n2; i += blockDim.x) {
                const int ixj = i ^ j;
                if (ixj > i) {
                    const unsigned long long a = keys[i], b = keys[ixj];
                    const bool up = (i & k) == 0;
                    if ((a > b) == up) { keys[i] = b; keys[ixj] = a; }
                }
            }
        }
    __syncthreads();
}

/* exclusive rank of the set flags among items [0, n): every thread owns a contiguous chunk; returns the total */
__device__ int block_rank_heads(const unsigned long long* keys, int n, int chunk, int* rankOfChunk, int* scratch) {
    const int t = threadIdx.x, b = t * chunk, e = min(b + chunk, n);
    int c = 0;
    for (int i = b; i < e; i++) c += (i == 0 || (keys[i] >> 32) != (keys[i - 1] >> 32)) ? 1 : 0;
    scratch[t] = c;
    __syncthreads();
    for (int o = 1; o < BOW_THREADS; o <<= 1) {
        const int vprev = t >= o ? scratch[t - o] : 0;
        __syncthreads();
        scratch[t] += vprev;
        __syncthreads();
    }
    rankOfChunk[t] = scratch[t] - c;
    const int total = scratch[BOW_THREADS - 1];
    __syncthreads();
    return total;
}

__global__ void __launch_bounds__(BOW_THREADS) bow_reduce_kernel(int n, int n2, const int* __restrict__ word,
                                                                 const int* __restrict__ nid, const double* __restrict__ w,
                                                                 int weighting, int mustNormalize, int normL2,
                                                                 int* __restrict__ bowIds, double* __restrict__ bowVals,
                                                                 int* __restrict__ fvNode, int* __restrict__ fvPtr,
                                                                 int* __restrict__ fvIdx, int* __restrict__ counts) {
    extern __shared__ unsigned long long keys[];
    __shared__ int scratch[BOW_THREADS], rankOfChunk[BOW_THREADS];
    __shared__ int nvalidS;
    __shared__ double normS;
    const int t = threadIdx.x;
    const int chunk = (n2 + BOW_THREADS - 1) / BOW_THREADS;
    if (t == 0) nvalidS = 0;
    __syncthreads();
    /* ---- BowVector: key = word id | feature index; stopped words (w <= 0, :1170) sort to the end ---- */
    int myValid = 0;
    for (int i = t; i < n2; i += BOW_THREADS) {
        const bool ok = i < n && w[i] > 0;
        keys[i] = ok ? ((unsigned long long)(unsigned)word[i] << 32) | (unsigned)i : ~0ull;
        myValid += ok;
    }
    if (myValid) atomicAdd(&nvalidS, myValid);
    bitonic_sort_u64(keys, n2);
    const int nvalid = nvalidS;
    const int nbow = nvalid ? block_rank_heads(keys, nvalid, chunk, rankOfChunk, scratch) : 0;
    {
        const int b = t * chunk, e = min(b + chunk, nvalid);
        int r = rankOfChunk[t];
        for (int i = b; i < e; i++) {
            if (i != 0 && (keys[i] >> 32) == (keys[i - 1] >> 32)) continue;
            const unsigned id = (unsigned)(keys[i] >> 32);
            const double wi = w[(unsigned)keys[i]];             /* the leaf's weight: the same for every feature of the word */
            double val = wi;
            if (weighting == 0 || weighting == 1)                /* TF_IDF / TF: addWeight accumulates in feature order */
                for (int j = i + 1; j < nvalid && (unsigned)(keys[j] >> 32) == id; j++) val = __dadd_rn(val, wi);
            bowIds[r] = (int)id;
            bowVals[r] = val;
            r++;
        }
    }
    __syncthreads();
    if (nbow > 0) {
        if (mustNormalize) {                                     /* BowVector::normalize, ascending word id */
            if (t == 0) {
                double norm = 0.0;
                if (!normL2) for (int i = 0; i < nbow; i++) norm = __dadd_rn(norm, fabs(bowVals[i]));
                else {
                    for (int i = 0; i < nbow; i++) norm = __dadd_rn(norm, __dmul_rn(bowVals[i], bowVals[i]));
                    norm = sqrt(norm);
                }
                normS = norm;
            }
            __syncthreads();
            const double norm = normS;
            if (norm > 0.0) for (int i = t; i < nbow; i += BOW_THREADS) bowVals[i] = bowVals[i] / norm;
        } else if (weighting == 0 || weighting == 1) {           /* :1177-1183 */
            const double nd = (double)nbow;
            for (int i = t; i < nbow; i += BOW_THREADS) bowVals[i] = bowVals[i] / nd;
        }
    }
    __syncthreads();
    /* ---- FeatureVector: key = node id | feature index ---- */
    for (int i = t; i < n2; i += BOW_THREADS) {
        const bool ok = i < n && w[i] > 0;
        keys[i] = ok ? ((unsigned long long)(unsigned)nid[i] << 32) | (unsigned)i : ~0ull;
    }
    bitonic_sort_u64(keys, n2);
    const int nfv = nvalid ? block_rank_heads(keys, nvalid, chunk, rankOfChunk, scratch) : 0;
    {
        const int b = t * chunk, e = min(b + chunk, nvalid);
        int r = rankOfChunk[t];
        for (int i = b; i < e; i++) {
            fvIdx[i] = (int)(unsigned)keys[i];
            if (i != 0 && (keys[i] >> 32) == (keys[i - 1] >> 32)) continue;
            fvNode[r] = (int)(keys[i] >> 32);
            fvPtr[r] = i;
            r++;
        }
    }
    if (t == 0) {
        fvPtr[nfv] = nvalid;
        counts[0] = nbow;
        counts[1] = nfv;
        counts[2] = nvalid;
    }
}

size_t pad(size_t b) { return (b + 255) & ~(size_t)255; }

}  // namespace

struct viorb_vocabulary {
    viorb_ctx* ctx = nullptr;
    uint8_t* mem = nullptr;
    VocabDev dev = {};
    int k = 0, L = 0, weighting = 0, scoring = 0, nnodes = 0, nwords = 0;
};

extern "C" {

int viorb_vocabulary_create(viorb_ctx* c, int k, int L, int weighting, int scoring, int nnodes, const int32_t* parent,
                            const uint8_t* node_desc, const double* node_weight, viorb_vocabulary** out) {
    if (!c || !out || nnodes < 1 || !parent || !node_desc || !node_weight || weighting < 0 || weighting > 3 || scoring < 0 ||
        scoring > 5)
        return viorb_fail(VIORB_ERR_INVALID, "bad argument");
    /* children lists in node-id order (loadFromTextFile pushes children as it reads the nodes, :1392-1407) */
    std::vector<int> childPtr((size_t)nnodes + 1, 0), childIdx((size_t)std::max(nnodes - 1, 1)), wordId((size_t)nnodes, -1);
    for (int i = 1; i < nnodes; i++) {
        if (parent[i] < 0 || parent[i] >= i) return viorb_fail(VIORB_ERR_INVALID, "node %d: parent %d must precede it", i, parent[i]);
        childPtr[parent[i] + 1]++;
    }
    for (int i = 0; i < nnodes; i++) childPtr[i + 1] += childPtr[i];
    std::vector<int> fill(childPtr.begin(), childPtr.end() - 1);
    std::vector<uint8_t> cdesc((size_t)std::max(nnodes - 1, 1) * 32);
    for (int i = 1; i < nnodes; i++) {
        const int slot = fill[parent[i]]++;
        childIdx[slot] = i;
        memcpy(&cdesc[(size_t)slot * 32], node_desc + (size_t)i * 32, 32);
    }
    int nwords = 0;
    for (int i = 1; i < nnodes; i++)
        if (childPtr[i + 1] == childPtr[i]) wordId[i] = nwords++;        /* leaves get word ids in file order (:1418-1425) */
    int rc;
    if ((rc = viorb_ctx_bind(c))) return rc;
    viorb_vocabulary* v = new (std::nothrow) viorb_vocabulary;
    if (!v) return viorb_fail(VIORB_ERR_INVALID, "out of memory");
    v->ctx = c; v->k = k; v->L = L; v->weighting = weighting; v->scoring = scoring; v->nnodes = nnodes; v->nwords = nwords;
    const size_t nc = (size_t)std::max(nnodes - 1, 1);
    const size_t bytes = pad(nc * 32) + pad(((size_t)nnodes + 1) * 4) + pad(nc * 4) + pad((size_t)nnodes * 8) + pad((size_t)nnodes * 4);
    cudaError_t e = cudaMalloc((void**)&v->mem, bytes);
    if (e != cudaSuccess) { delete v; return viorb_fail(VIORB_ERR_CUDA, "cudaMalloc(%zu) failed: %s", bytes, cudaGetErrorString(e)); }
    uint8_t* p = v->mem;
    uint8_t* d_cdesc = p; p += pad(nc * 32);
    int* d_ptr = (int*)p; p += pad(((size_t)nnodes + 1) * 4);
    int* d_idx = (int*)p; p += pad(nc * 4);
    double* d_w = (double*)p; p += pad((size_t)nnodes * 8);
    int* d_wid = (int*)p;
    cudaStream_t s = viorb_ctx_stream(c);
    if (cudaMemcpyAsync(d_cdesc, cdesc.data(), nc * 32, cudaMemcpyHostToDevice, s) != cudaSuccess ||
        cudaMemcpyAsync(d_ptr, childPtr.data(), ((size_t)nnodes + 1) * 4, cudaMemcpyHostToDevice, s) != cudaSuccess ||
        cudaMemcpyAsync(d_idx, childIdx.data(), nc * 4, cudaMemcpyHostToDevice, s) != cudaSuccess ||
        cudaMemcpyAsync(d_w, node_weight, (size_t)nnodes * 8, cudaMemcpyHostToDevice, s) != cudaSuccess ||
        cudaMemcpyAsync(d_wid, wordId.data(), (size_t)nnodes * 4, cudaMemcpyHostToDevice, s) != cudaSuccess ||
        cudaStreamSynchronize(s) != cudaSuccess) {
        cudaFree(v->mem);
        delete v;
        return viorb_fail(VIORB_ERR_CUDA, "vocabulary upload failed: %s", cudaGetErrorString(cudaGetLastError()));
    }
    v->dev.cdesc = d_cdesc; v->dev.childPtr = d_ptr; v->dev.childIdx = d_idx; v->dev.weight = d_w; v->dev.wordId = d_wid;
    *out = v;
    return VIORB_OK;
}

int viorb_vocabulary_destroy(viorb_vocabulary* v) {
    if (!v) return VIORB_OK;
    cudaFree(v->mem);
    delete v;
    return VIORB_OK;
}

int viorb_vocabulary_info(const viorb_vocabulary* v, int* nnodes, int* nwords) {
    if (!v) return viorb_fail(VIORB_ERR_INVALID, "bad argument");
    if (nnodes) *nnodes = v->nnodes;
    if (nwords) *nwords = v->nwords;
    return VIORB_OK;
}

int viorb_bow_transform(viorb_vocabulary* v, const uint8_t* desc, int n, int levelsup, int32_t* bow_ids, double* bow_values,
                        int* nbow, int32_t* fv_node, int32_t* fv_ptr, int32_t* fv_idx, int* nfv, int32_t* word_of,
                        int32_t* node_of) {
    if (!v || n < 0 || !nbow || !nfv || (n > 0 && (!desc || !bow_ids || !bow_values || !fv_node || !fv_ptr || !fv_idx)))
        return viorb_fail(VIORB_ERR_INVALID, "bad argument");
    *nbow = 0; *nfv = 0;
    if (n == 0 || v->nnodes < 2) {                     /* empty(): transform clears v and fv and returns (:1145-1148) */
        if (fv_ptr) fv_ptr[0] = 0;
        return VIORB_OK;
    }
    int n2 = 1;
    while (n2 < n) n2 <<= 1;
    if ((size_t)n2 * 8 > 200 * 1024) return viorb_fail(VIORB_ERR_UNSUPPORTED, "more than 25600 features per transform");
    viorb_ctx* c = v->ctx;
    int rc;
    if ((rc = viorb_ctx_bind(c))) return rc;
    cudaStream_t s = viorb_ctx_stream(c);
    const size_t bytes = pad((size_t)n * 32) + 6 * pad((size_t)(n + 1) * 4) + 2 * pad((size_t)n * 8) + pad(64) + 4096;
    uint8_t* base = nullptr;
    if ((rc = viorb_ctx_scratch(c, bytes, &base))) return rc;
    uint8_t* p = base;
    uint8_t* d_desc = p; p += pad((size_t)n * 32);
    int* d_word = (int*)p; p += pad((size_t)(n + 1) * 4);
    int* d_nid = (int*)p; p += pad((size_t)(n + 1) * 4);
    int* d_bowIds = (int*)p; p += pad((size_t)(n + 1) * 4);
    int* d_fvNode = (int*)p; p += pad((size_t)(n + 1) * 4);
    int* d_fvPtr = (int*)p; p += pad((size_t)(n + 1) * 4);
    int* d_fvIdx = (int*)p; p += pad((size_t)(n + 1) * 4);
    double* d_w = (double*)p; p += pad((size_t)n * 8);
    double* d_bowVals = (double*)p; p += pad((size_t)n * 8);
    int* d_counts = (int*)p;
    VCU(cudaMemcpyAsync(d_desc, desc, (size_t)n * 32, cudaMemcpyHostToDevice, s));
    bow_descend_kernel<<<(n + 7) / 8, 256, 0, s>>>(v->dev, d_desc, n, v->L - levelsup, d_word, d_nid, d_w);
    /* mustNormalize / norm of the scoring object (ScoringObject.h:74-90): all but DOT_PRODUCT normalise, L2 only for L2_NORM */
    const int must = v->scoring != 5, normL2 = v->scoring == 1;
    static bool attr = false;
    if (!attr) {
        VCU(cudaFuncSetAttribute(bow_reduce_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
        attr = true;
    }
    bow_reduce_kernel<<<1, BOW_THREADS, (size_t)n2 * 8, s>>>(n, n2, d_word, d_nid, d_w, v->weighting, must, normL2, d_bowIds,
                                                           d_bowVals, d_fvNode, d_fvPtr, d_fvIdx, d_counts);
    viorb_ctx_add_launches(c, 2);
    VCU(cudaGetLastError());
    int counts[3] = {0, 0, 0};
    VCU(cudaMemcpyAsync(counts, d_counts, 12, cudaMemcpyDeviceToHost, s));
    VCU(cudaStreamSynchronize(s));
    VCU(cudaMemcpyAsync(bow_ids, d_bowIds, (size_t)counts[0] * 4, cudaMemcpyDeviceToHost, s));
    VCU(cudaMemcpyAsync(bow_values, d_bowVals, (size_t)counts[0] * 8, cudaMemcpyDeviceToHost, s));
    VCU(cudaMemcpyAsync(fv_node, d_fvNode, (size_t)counts[1] * 4, cudaMemcpyDeviceToHost, s));
    VCU(cudaMemcpyAsync(fv_ptr, d_fvPtr, ((size_t)counts[1] + 1) * 4, cudaMemcpyDeviceToHost, s));
    VCU(cudaMemcpyAsync(fv_idx, d_fvIdx, (size_t)counts[2] * 4, cudaMemcpyDeviceToHost, s));
    if (word_of) VCU(cudaMemcpyAsync(word_of, d_word, (size_t)n * 4, cudaMemcpyDeviceToHost, s));
    if (node_of) VCU(cudaMemcpyAsync(node_of, d_nid, (size_t)n * 4, cudaMemcpyDeviceToHost, s));
    VCU(cudaStreamSynchronize(s));
    *nbow = counts[0];
    *nfv = counts[1];
    return VIORB_OK;
}

}  // extern "C"
