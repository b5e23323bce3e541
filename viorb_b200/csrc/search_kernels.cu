/*
 * search_kernels.cu -- sm_100a kernels for the small windowed matchers of the hot path:
 *
 *   stereo_match_kernel (+ rank / filter)   Frame::ComputeStereoMatches            src/Frame.cc:646-820
 *   grid_build_kernel                       Frame::AssignFeaturesToGrid / PosInGrid src/Frame.cc:410-425,562-572
 *   features_in_area_kernel                 Frame::GetFeaturesInArea                src/Frame.cc:507-560
 *   search_local_kernel                     ORBmatcher::SearchByProjection(Frame&, vector<MapPoint*>&, th)   :45-129
 *   search_frame_kernel                     ORBmatcher::SearchByProjection(Frame&, const Frame&, th, bMono)  :1328-1471
 *   triangulation_kernel (+ finalize)       ORBmatcher::SearchForTriangulation                               :657-823
 *
 * The projection searches carry a sequential dependence in the reference (a keypoint claimed by an
 * earlier map point is skipped by later ones, ORBmatcher.cc:87-89,123).  Query i's result depends only
 * on the results of queries j < i, so the kernels iterate the whole query set in parallel to the
 * unique fixed point of that triangular system: after pass k every query whose dependence chain is
 * shorter than k holds its sequential answer; the loop ends when a pass changes nothing (typically
 * 2-3 passes).  All of it is integer / popc work with a few single-rounded float gates; no tensor cores.
 */
#include <limits.h>

#include <algorithm>

#include "matcher_kernels.cuh"

namespace {

#define TH_HIGH 100
#define TH_LOW 50
#define HISTO_LENGTH 30
#define GRID_COLS 64
#define GRID_ROWS 48

__device__ __forceinline__ int hamming_rows(const uint8_t* a, const uint8_t* b) {
    const uint4 a0 = reinterpret_cast<const uint4*>(a)[0], a1 = reinterpret_cast<const uint4*>(a)[1];
    const uint4 b0 = reinterpret_cast<const uint4*>(b)[0], b1 = reinterpret_cast<const uint4*>(b)[1];
    return __popc(a0.x ^ b0.x) + __popc(a0.y ^ b0.y) + __popc(a0.z ^ b0.z) + __popc(a0.w ^ b0.w) +
           __popc(a1.x ^ b1.x) + __popc(a1.y ^ b1.y) + __popc(a1.z ^ b1.z) + __popc(a1.w ^ b1.w);
}

__device__ __forceinline__ unsigned long long warp_min_u64(unsigned long long v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const unsigned long long t = __shfl_xor_sync(0xffffffffu, v, o);
        v = t < v ? t : v;
    }
    return v;
}

/* two smallest keys over the warp; every lane contributes its own (k1 <= k2) */
__device__ __forceinline__ void warp_two_min(unsigned long long& k1, unsigned long long& k2) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const unsigned long long b1 = __shfl_xor_sync(0xffffffffu, k1, o);
        const unsigned long long b2 = __shfl_xor_sync(0xffffffffu, k2, o);
        const unsigned long long lo = k1 < b1 ? k1 : b1, hi = k1 < b1 ? b1 : k1;
        const unsigned long long s2 = k2 < b2 ? k2 : b2;
        k1 = lo;
        k2 = hi < s2 ? hi : s2;
    }
}

/* ------------------------------------------------------------------------------------------------
 * Frame::ComputeStereoMatches, one warp per left keypoint.
 * ---------------------------------------------------------------------------------------------- */
__global__ void __launch_bounds__(128) stereo_match_kernel(const __grid_constant__ StereoParams p,
                                                           const viorb_keypoint* __restrict__ kl,
                                                           const uint8_t* __restrict__ dl,
                                                           const viorb_keypoint* __restrict__ kr,
                                                           const uint8_t* __restrict__ dr, float* __restrict__ uRight,
                                                           float* __restrict__ depth, int* __restrict__ sad,
                                                           int* __restrict__ nMatched) {
    const int lane = threadIdx.x & 31;
    const int iL = blockIdx.x * 4 + (threadIdx.x >> 5);
    if (iL >= p.nl) return;
    const viorb_keypoint kpL = kl[iL];
    const int levelL = kpL.octave;
    const float vL = kpL.y, uL = kpL.x;
    if (lane == 0) { uRight[iL] = -1.0f; depth[iL] = -1.0f; sad[iL] = -1; }
    const float minZ = p.mb;
    const float minD = 0;
    const float maxD = __fdiv_rn(p.mbf, minZ);
    const float minU = __fsub_rn(uL, maxD), maxU = __fsub_rn(uL, minD);
    if (maxU < 0) return;                                          /* :699-700 */
    const int yi = (int)vL;                                        /* vRowIndices[vL] :691 */
    unsigned long long best = ~0ull;
    const uint8_t* dL = dl + (size_t)iL * 32;
    for (int iR = lane; iR < p.nr; iR += 32) {
        const viorb_keypoint kpR = kr[iR];
        const float r = __fmul_rn(2.0f, p.scale[kpR.octave]);      /* :667 */
        const int maxr = (int)ceilf(__fadd_rn(kpR.y, r)), minr = (int)floorf(__fsub_rn(kpR.y, r));
        if (yi < minr || yi > maxr) continue;                      /* row table :668-672 */
        if (kpR.octave < levelL - 1 || kpR.octave > levelL + 1) continue;
        const float uR = kpR.x;
        if (uR >= minU && uR <= maxU) {
            const int dist = hamming_rows(dL, dr + (size_t)iR * 32);
            if (dist < TH_HIGH) {                                  /* bestDist starts at TH_HIGH, strict < */
                const unsigned long long key = ((unsigned long long)dist << 32) | (unsigned)iR;
                best = key < best ? key : best;
            }
        }
    }
    best = warp_min_u64(best);
    if (best == ~0ull) return;
    const int bestDist = (int)(best >> 32), bestIdxR = (int)(best & 0xffffffffu);
    const int thOrbDist = (TH_HIGH + TH_LOW) / 2;
    if (!(bestDist < thOrbDist)) return;
    /* sub-pixel match by correlation (:732-803) */
    const float uR0 = kr[bestIdxR].x;
    const float scaleFactor = p.invScale[levelL];
    const float scaleduL = roundf(__fmul_rn(kpL.x, scaleFactor));
    const float scaledvL = roundf(__fmul_rn(kpL.y, scaleFactor));
    const float scaleduR0 = roundf(__fmul_rn(uR0, scaleFactor));
    const int w = 5, L = 5;
    const StereoLevel& lv = p.lv[levelL];
    const int cvL = (int)scaledvL, cuL = (int)scaleduL, cuR = (int)scaleduR0;
    if (cvL - w < 0 || cvL + w + 1 > lv.h || cuL - w < 0 || cuL + w + 1 > lv.w) return;   /* cv::Mat range (C.5) */
    const float iniu = __fadd_rn(scaleduR0, (float)(L - w));
    const float endu = __fadd_rn(scaleduR0, (float)(L + w + 1));
    if (iniu < 0 || endu >= lv.w) return;                          /* :753-756 */
    if (cuR - L - w < 0 || cuR + L + w + 1 > lv.w) return;         /* colRange would throw */
    int il[4], py[4], px[4];
    const int cL = lv.roiL[(size_t)cvL * lv.stepL + cuL];
#pragma unroll
    for (int k = 0; k < 4; k++) {
        const int pidx = lane + 32 * k;
        py[k] = pidx / 11; px[k] = pidx - py[k] * 11;
        il[k] = pidx < 121 ? (int)lv.roiL[(size_t)(cvL - w + py[k]) * lv.stepL + (cuL - w + px[k])] - cL : 0;
    }
    int bestSad = INT_MAX, bestinc = 0;
    int d[11];
#pragma unroll
    for (int inc = -L; inc <= L; inc++) {
        const int cR = lv.roiR[(size_t)cvL * lv.stepR + (cuR + inc)];
        int s = 0;
#pragma unroll
        for (int k = 0; k < 4; k++) {
            if (lane + 32 * k < 121) {
                const int ir = (int)lv.roiR[(size_t)(cvL - w + py[k]) * lv.stepR + (cuR + inc - w + px[k])] - cR;
                s += abs(il[k] - ir);
            }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
        d[inc + L] = s;
        if (s < bestSad) { bestSad = s; bestinc = inc; }
    }
    if (bestinc == -L || bestinc == L) return;
    float dist1 = 0, dist2 = 0, dist3 = 0;
#pragma unroll
    for (int i = 1; i < 10; i++)
        if (i == bestinc + L) { dist1 = (float)d[i - 1]; dist2 = (float)d[i]; dist3 = (float)d[i + 1]; }
    const float deltaR = __fdiv_rn(__fsub_rn(dist1, dist3),
                                   __fmul_rn(2.0f, __fsub_rn(__fadd_rn(dist1, dist3), __fmul_rn(2.0f, dist2))));
    if (deltaR < -1 || deltaR > 1) return;
    float bestuR = __fmul_rn(p.scale[levelL], __fadd_rn(__fadd_rn(scaleduR0, (float)bestinc), deltaR));
    float disparity = __fsub_rn(uL, bestuR);
    if (disparity >= minD && disparity < maxD) {
        if (disparity <= 0) {
            disparity = 0.01f;
            bestuR = (float)((double)uL - 0.01);
        }
        if (lane == 0) {
            depth[iL] = __fdiv_rn(p.mbf, disparity);
            uRight[iL] = bestuR;
            sad[iL] = bestSad;
            atomicAdd(nMatched, 1);
        }
    }
}

/* median of the accepted SAD values: the element of rank size/2 in the (sad, iL) order (:806-807) */
__global__ void stereo_rank_kernel(const int* __restrict__ sad, int nl, const int* __restrict__ nMatched,
                                   int* __restrict__ median) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nl) return;
    const int s = sad[i];
    if (s < 0) return;
    int rank = 0;
    for (int j = 0; j < nl; j++) {
        const int t = sad[j];
        rank += (t >= 0) && (t < s || (t == s && j < i));
    }
    if (rank == *nMatched / 2) *median = s;
}

__global__ void stereo_filter_kernel(const int* __restrict__ sad, int nl, const int* __restrict__ median,
                                     float* __restrict__ uRight, float* __restrict__ depth) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nl) return;
    const int s = sad[i];
    if (s < 0) return;
    const float thDist = __fmul_rn(__fmul_rn(1.5f, 1.4f), (float)*median);      /* :808 */
    if (!((float)s < thDist)) { uRight[i] = -1; depth[i] = -1; }
}

/* ------------------------------------------------------------------------------------------------
 * 64x48 grid: AssignFeaturesToGrid.  Single CTA; slot order inside a cell = keypoint index order.
 * ---------------------------------------------------------------------------------------------- */
__global__ void __launch_bounds__(1024) grid_build_kernel(const viorb_keypoint* __restrict__ kps, int n, float minX,
                                                          float minY, float invW, float invH, int* __restrict__ cellOf,
                                                          int* __restrict__ cellStart, int* __restrict__ cellItems) {
    __shared__ int cnt[GRID_COLS * GRID_ROWS + 1];
    __shared__ int part[1024];
    const int tid = threadIdx.x;
    const int NCELL = GRID_COLS * GRID_ROWS;
    for (int c = tid; c <= NCELL; c += 1024) cnt[c] = 0;
    __syncthreads();
    for (int i = tid; i < n; i += 1024) {
        /* PosInGrid :562-572 */
        const int posX = (int)roundf(__fmul_rn(__fsub_rn(kps[i].x, minX), invW));
        const int posY = (int)roundf(__fmul_rn(__fsub_rn(kps[i].y, minY), invH));
        int c = -1;
        if (!(posX < 0 || posX >= GRID_COLS || posY < 0 || posY >= GRID_ROWS)) {
            c = posX * GRID_ROWS + posY;
            atomicAdd(&cnt[c], 1);
        }
        cellOf[i] = c;
    }
    __syncthreads();
    /* exclusive scan of 3072 counters: 3 per thread */
    const int per = (NCELL + 1023) / 1024;
    int s = 0;
    for (int k = 0; k < per; k++) {
        const int c = tid * per + k;
        if (c < NCELL) s += cnt[c];
    }
    part[tid] = s;
    __syncthreads();
    for (int o = 1; o < 1024; o <<= 1) {
        const int t = tid >= o ? part[tid - o] : 0;
        __syncthreads();
        part[tid] += t;
        __syncthreads();
    }
    int run = part[tid] - s;
    for (int k = 0; k < per; k++) {
        const int c = tid * per + k;
        if (c < NCELL) {
            const int t = cnt[c];
            cellStart[c] = run;
            cnt[c] = run;
            run += t;
        }
    }
    if (tid == 1023) cellStart[NCELL] = part[1023];
    __syncthreads();
    /* slot order inside a cell = keypoint index order (mGrid[x][y].push_back(i), :419-423): blocks of 1024 keypoints in
     * index order, the 32 warps of a block take turns, and inside a warp a keypoint's slot is the cell's running count
     * plus the number of lower lanes that fall into the same cell */
    const int lane = tid & 31, warp = tid >> 5;
    for (int base = 0; base < n; base += 1024) {
        const int i = base + tid;
        const int c = i < n ? cellOf[i] : -1;
        const unsigned same = __match_any_sync(0xffffffffu, c);
        const int rank = __popc(same & ((1u << lane) - 1));
        for (int w = 0; w < 32; w++) {
            if (warp == w) {
                const int start = c >= 0 ? cnt[c] : 0;
                __syncwarp();
                if (c >= 0) {
                    cellItems[start + rank] = i;
                    if (rank == 0) cnt[c] = start + __popc(same);
                }
            }
            __syncthreads();
        }
    }
}

/* GetFeaturesInArea enumeration (:507-560): calls f(pos, idx) for every keypoint that passes the level and
 * window gates, in parallel over the lanes of a warp; `pos` is the rank of the keypoint in the reference's
 * sequential enumeration (cells ix-major, then iy, then slot), so ties can be broken exactly. */
template <typename F>
__device__ __forceinline__ void for_features_in_area(const FrameIndexDev& fi, float x, float y, float r, int minLevel,
                                                     int maxLevel, int lane, F f) {
    const int nMinCellX = max(0, (int)floorf(__fmul_rn(__fsub_rn(__fsub_rn(x, fi.minX), r), fi.invW)));
    if (nMinCellX >= GRID_COLS) return;
    const int nMaxCellX = min(GRID_COLS - 1, (int)ceilf(__fmul_rn(__fadd_rn(__fsub_rn(x, fi.minX), r), fi.invW)));
    if (nMaxCellX < 0) return;
    const int nMinCellY = max(0, (int)floorf(__fmul_rn(__fsub_rn(__fsub_rn(y, fi.minY), r), fi.invH)));
    if (nMinCellY >= GRID_ROWS) return;
    const int nMaxCellY = min(GRID_ROWS - 1, (int)ceilf(__fmul_rn(__fadd_rn(__fsub_rn(y, fi.minY), r), fi.invH)));
    if (nMaxCellY < 0) return;
    const bool bCheckLevels = (minLevel > 0) || (maxLevel >= 0);
    int pos = 0;
    for (int ix = nMinCellX; ix <= nMaxCellX; ix++) {
        /* cells (ix, nMinCellY..nMaxCellY) are contiguous in the CSR */
        const int beg = fi.cellStart[ix * GRID_ROWS + nMinCellY], end = fi.cellStart[ix * GRID_ROWS + nMaxCellY + 1];
        for (int j = beg + lane; j < end; j += 32) {
            const int idx = fi.cellItems[j];
            const viorb_keypoint kp = fi.kps[idx];
            if (bCheckLevels) {
                if (kp.octave < minLevel) continue;
                if (maxLevel >= 0 && kp.octave > maxLevel) continue;
            }
            const float distx = __fsub_rn(kp.x, x), disty = __fsub_rn(kp.y, y);
            if (fabsf(distx) < r && fabsf(disty) < r) f(pos + (j - beg), idx, kp.octave);
        }
        pos += end - beg;
    }
}

__global__ void features_in_area_kernel(FrameIndexDev fi, float x, float y, float r, int minLevel, int maxLevel,
                                        unsigned long long* __restrict__ keys, int* __restrict__ count) {
    /* single warp: emit (pos, idx) pairs; the host sorts by pos (test / debugging entry point) */
    const int lane = threadIdx.x;
    for_features_in_area(fi, x, y, r, minLevel, maxLevel, lane, [&](int pos, int idx, int) {
        const int k = atomicAdd(count, 1);
        keys[k] = ((unsigned long long)pos << 32) | (unsigned)idx;
    });
}

/* ------------------------------------------------------------------------------------------------
 * The projection searches (ORBmatcher::SearchByProjection x4) in two steps:
 *
 *  (1) candidate lists -- every query (map point) in parallel, one warp each, many CTAs: enumerate the grid window,
 *      apply every gate that does not depend on other queries (level range, window, stereo gate, keypoints that
 *      already hold an observed map point on entry), compute the Hamming distances and keep the LIST_T best
 *      candidates sorted by (distance, enumeration rank) plus the total number of passing candidates;
 *  (2) ordered resolve -- one CTA.  In the reference a keypoint claimed by an earlier map point is skipped by later
 *      ones (:87-89, :123), so query i depends on the queries j < i only: all queries are re-evaluated in parallel
 *      (one thread each, walking its short list) until nothing changes -- the unique fixed point of that triangular
 *      system, reached after (longest dependence chain + 1) sweeps, typically 2-3.  A query whose list runs out while
 *      more candidates exist (more than LIST_T, with the best ones claimed) falls back to a full enumeration by a warp.
 * ---------------------------------------------------------------------------------------------- */
#define LIST_T 8
#define LIST_BUF 160

struct LocalArgs {
    const float *projX, *projY, *projXR, *viewCos;
    const int* predLevel;
    const uint8_t* valid;
    const int* nobs;
    const uint8_t* mpDesc;
    int nmp;
    float th, nnratio;
};

struct FrameArgs {
    const float *u, *v, *invz, *lastAngle;
    const int* lastOctave;
    const uint8_t* valid;
    const int* nobs;
    const uint8_t* mpDesc;
    int nlast;
    float th, mbf;
    int mode, checkOri, thHigh;
};

/* the window of one query: centre, radius, level range, stereo prediction */
struct Window {
    bool valid;
    float x, y, rad, ur;
    int minL, maxL;
    bool stereoGate;
};

__device__ __forceinline__ Window make_window(const FrameIndexDev& fi, const LocalArgs& a, int i) {
    Window w;
    w.valid = a.valid[i] != 0;
    if (!w.valid) return w;
    const int lvl = a.predLevel[i];
    float r = (double)a.viewCos[i] > 0.998 ? 2.5f : 4.0f;                      /* RadiusByViewingCos :131-137 */
    if (a.th != 1.0f) r = __fmul_rn(r, a.th);
    w.rad = __fmul_rn(r, fi.scale[lvl]);
    w.x = a.projX[i]; w.y = a.projY[i]; w.ur = a.projXR[i];
    w.minL = lvl - 1; w.maxL = lvl;                                            /* :68-69 */
    w.stereoGate = true;
    return w;
}

__device__ __forceinline__ Window make_window(const FrameIndexDev& fi, const FrameArgs& a, int i) {
    Window w;
    const float u = a.u[i], v = a.v[i];
    w.valid = a.valid[i] && !(u < fi.minX || u > fi.maxX) && !(v < fi.minY || v > fi.maxY);
    if (!w.valid) return w;
    const int oct = a.lastOctave[i];
    w.rad = __fmul_rn(a.th, fi.scale[oct]);
    const int lm = a.mode & 7;
    if (lm == 1) { w.minL = oct; w.maxL = -1; }                                /* forward  :1385-1386 */
    else if (lm == 2) { w.minL = 0; w.maxL = oct; }                            /* backward :1387-1388 */
    else if (lm == 3) { w.minL = oct - 1; w.maxL = oct; }                      /* Sim3 overload :375-379 */
    else { w.minL = oct - 1; w.maxL = oct + 1; }
    w.stereoGate = !(a.mode & 8);                                              /* the KeyFrame overloads have no uRight gate */
    w.x = u; w.y = v;
    w.ur = __fsub_rn(u, __fmul_rn(a.mbf, a.invz[i]));
    return w;
}

/* key of one candidate, or ~0 when a gate rejects it.  LOCAL keys carry the octave (the ratio test of :118-124 compares
 * the levels of best and second best): dist:16 | rank:20 | octave:8 | index:20; FRAME keys: dist:16 | rank:24 | index:24 */
template <bool LOCAL>
__device__ __forceinline__ unsigned long long candidate_key(const FrameIndexDev& fi, const Window& w, const uint8_t* dMP,
                                                            int pos, int idx, int oct) {
    const float uR = fi.uRight[idx];
    if (w.stereoGate && uR > 0) {
        const float er = fabsf(__fsub_rn(w.ur, uR));
        if (er > w.rad) return ~0ull;                                           /* :91-96, :1408-1414 */
    }
    const unsigned long long dist = (unsigned long long)hamming_rows(dMP, fi.desc + (size_t)idx * 32);
    if (LOCAL) return (dist << 48) | ((unsigned long long)pos << 28) | ((unsigned long long)oct << 20) | (unsigned)idx;
    return (dist << 48) | ((unsigned long long)pos << 24) | (unsigned)idx;
}
template <bool LOCAL>
__device__ __forceinline__ int key_index(unsigned long long k) { return LOCAL ? (int)(k & 0xfffff) : (int)(k & 0xffffff); }

template <bool LOCAL, typename ARGS>
__global__ void __launch_bounds__(128) candidate_lists_kernel(FrameIndexDev fi, ARGS a, int nq, const int* __restrict__ obs0,
                                                              unsigned long long* __restrict__ listKeys, int* __restrict__ listCount) {
    __shared__ unsigned long long buf[4][LIST_BUF];
    __shared__ int cnt[4];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const int i = blockIdx.x * 4 + wib;
    if (i >= nq) return;
    const Window w = make_window(fi, a, i);
    if (lane == 0) cnt[wib] = 0;
    __syncwarp();
    int total = 0;
    if (w.valid) {
        const uint8_t* dMP = a.mpDesc + (size_t)i * 32;
        for_features_in_area(fi, w.x, w.y, w.rad, w.minL, w.maxL, lane, [&](int pos, int idx, int oct) {
            if (obs0[idx] > 0) return;                                          /* :87-89 on entry */
            const unsigned long long key = candidate_key<LOCAL>(fi, w, dMP, pos, idx, oct);
            if (key == ~0ull) return;
            const int slot = atomicAdd(&cnt[wib], 1);
            if (slot < LIST_BUF) buf[wib][slot] = key;
        });
        __syncwarp();
        total = cnt[wib];
    }
    if (total > LIST_BUF) {                                                     /* a window this crowded is resolved by full scans */
        if (lane == 0) listCount[i] = -1;
        return;
    }
    /* the LIST_T smallest keys in ascending order: keys are distinct (they contain the keypoint index) */
    unsigned long long last = 0;
    const int keep = min(total, LIST_T);
    for (int t = 0; t < keep; t++) {
        unsigned long long m = ~0ull;
        for (int j = lane; j < total; j += 32) {
            const unsigned long long k = buf[wib][j];
            if ((t == 0 || k > last) && k < m) m = k;
        }
        m = warp_min_u64(m);
        if (lane == 0) listKeys[(size_t)i * LIST_T + t] = m;
        last = m;
    }
    if (lane == 0) listCount[i] = total;
}

/* full evaluation of one query by a warp under the current claims (the slow path of the resolve step) */
template <bool LOCAL>
__device__ __forceinline__ void scan_window(const FrameIndexDev& fi, const Window& w, const uint8_t* dMP, int i, int lane,
                                            const int* __restrict__ obs0, const int* __restrict__ minClaim,
                                            unsigned long long& k1, unsigned long long& k2) {
    k1 = ~0ull; k2 = ~0ull;
    if (!w.valid) return;
    for_features_in_area(fi, w.x, w.y, w.rad, w.minL, w.maxL, lane, [&](int pos, int idx, int oct) {
        if (obs0[idx] > 0 || minClaim[idx] < i) return;                         /* :87-89 (+ earlier matches :123) */
        const unsigned long long key = candidate_key<LOCAL>(fi, w, dMP, pos, idx, oct);
        if (key < k1) { k2 = k1; k1 = key; }
        else if (key < k2) k2 = key;
    });
    warp_two_min(k1, k2);
}

/* accept / reject of the two best candidates: the local-map overload applies the ratio test between candidates of the
 * same level (:118-124), the others a plain threshold (:1421, :1551, :389) */
__device__ __forceinline__ int decide(const LocalArgs& a, unsigned long long k1, unsigned long long k2) {
    const float nnratio = a.nnratio;
    if (k1 == ~0ull) return -1;
    const int bestDist = (int)(k1 >> 48), bestLevel = (int)((k1 >> 20) & 0xff);
    const int bestDist2 = k2 != ~0ull ? (int)(k2 >> 48) : 256;
    const int bestLevel2 = k2 != ~0ull ? (int)((k2 >> 20) & 0xff) : -1;
    if (bestDist > TH_HIGH) return -1;
    if (bestLevel == bestLevel2 && (float)bestDist > __fmul_rn(nnratio, (float)bestDist2)) return -1;
    return (int)(k1 & 0xfffff);
}
__device__ __forceinline__ int decide(const FrameArgs& a, unsigned long long k1, unsigned long long) {
    return (k1 != ~0ull && (int)(k1 >> 48) <= a.thHigh) ? (int)(k1 & 0xffffff) : -1;
}
/* the rotation-consistency histogram exists in the top-1 overloads only (:1432-1468) */
__device__ __forceinline__ bool wants_histogram(const LocalArgs&) { return false; }
__device__ __forceinline__ bool wants_histogram(const FrameArgs& a) { return a.checkOri != 0; }
__device__ __forceinline__ float query_angle(const LocalArgs&, int) { return 0.f; }
__device__ __forceinline__ float query_angle(const FrameArgs& a, int i) { return a.lastAngle[i]; }

/* ComputeThreeMaxima :1602-1643 on bin counts */
__device__ void three_maxima(const int* histo, int L, int& ind1, int& ind2, int& ind3) {
    int max1 = 0, max2 = 0, max3 = 0;
    ind1 = ind2 = ind3 = -1;
    for (int i = 0; i < L; i++) {
        const int s = histo[i];
        if (s > max1) { max3 = max2; max2 = max1; max1 = s; ind3 = ind2; ind2 = ind1; ind1 = i; }
        else if (s > max2) { max3 = max2; max2 = s; ind3 = ind2; ind2 = i; }
        else if (s > max3) { max3 = s; ind3 = i; }
    }
    if ((float)max2 < __fmul_rn(0.1f, (float)max1)) { ind2 = -1; ind3 = -1; }
    else if ((float)max3 < __fmul_rn(0.1f, (float)max1)) { ind3 = -1; }
}

__device__ __forceinline__ int rot_bin(float a1, float a2) {
    /* :1434-1439: rot = a1 - a2; if (rot < 0) rot += 360; bin = round(rot * (1/30)) ; bin == 30 -> 0 */
    float rot = __fsub_rn(a1, a2);
    if (rot < 0.0f) rot = __fadd_rn(rot, 360.0f);
    int bin = (int)roundf(__fmul_rn(rot, 1.0f / HISTO_LENGTH));
    if (bin == HISTO_LENGTH) bin = 0;
    return bin;
}

/* the ordered resolve, one CTA.  LOCAL: SearchByProjection(Frame&, const vector<MapPoint*>&, th) (:45-129);
 * otherwise the three overloads that share the top-1 loop (:1378-1468, :1473-1600, :290-403). */
template <bool LOCAL, typename ARGS>
__global__ void __launch_bounds__(1024) resolve_kernel(FrameIndexDev fi, ARGS a, int nq, const int* __restrict__ obs0,
                                                       const unsigned long long* __restrict__ listKeys,
                                                       const int* __restrict__ listCount, int* __restrict__ obs,
                                                       int* __restrict__ claim, int* __restrict__ minClaim, int* __restrict__ slowQueue,
                                                       int* __restrict__ match, int* __restrict__ nmatches) {
    __shared__ int changed, total, nslow;
    __shared__ int hist[HISTO_LENGTH];
    __shared__ int keep[3];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarps = blockDim.x >> 5;
    for (int i = tid; i < nq; i += blockDim.x) claim[i] = -1;
    for (int k = tid; k < fi.n; k += blockDim.x) { minClaim[k] = INT_MAX; match[k] = -1; }
    if (tid < HISTO_LENGTH) hist[tid] = 0;
    if (tid == 0) total = 0;
    __syncthreads();
    for (int iter = 0; iter <= nq; iter++) {
        if (tid == 0) { changed = 0; nslow = 0; }
        __syncthreads();
        /* fast path: one thread per query walks its sorted list and skips what earlier queries hold */
        for (int i = tid; i < nq; i += blockDim.x) {
            const int cnt = listCount[i];
            if (cnt == 0) continue;                                             /* no candidate at all: claim stays -1 */
            bool slow = cnt < 0;
            if (!slow) {
                const int len = min(cnt, LIST_T);
                unsigned long long k1 = ~0ull, k2 = ~0ull;
                for (int t = 0; t < len; t++) {
                    const unsigned long long k = listKeys[(size_t)i * LIST_T + t];
                    if (minClaim[key_index<LOCAL>(k)] < i) continue;
                    if (k1 == ~0ull) { k1 = k; if (!LOCAL) break; }
                    else { k2 = k; break; }
                }
                /* the list answers the query unless it ran out before `more` candidates were seen */
                const bool exhausted = cnt > LIST_T && (LOCAL ? k2 == ~0ull : k1 == ~0ull);
                if (exhausted) slow = true;
                else {
                    const int result = decide(a, k1, k2);
                    if (result != claim[i]) { claim[i] = result; changed = 1; }
                }
            }
            if (slow) slowQueue[atomicAdd(&nslow, 1)] = i;
        }
        __syncthreads();
        for (int s = warp; s < nslow; s += nwarps) {
            const int i = slowQueue[s];
            const Window w = make_window(fi, a, i);
            unsigned long long k1, k2;
            scan_window<LOCAL>(fi, w, a.mpDesc + (size_t)i * 32, i, lane, obs0, minClaim, k1, k2);
            const int result = decide(a, k1, k2);
            if (lane == 0 && result != claim[i]) { claim[i] = result; changed = 1; }
        }
        __syncthreads();
        if (!changed) break;
        for (int k = tid; k < fi.n; k += blockDim.x) minClaim[k] = INT_MAX;
        __syncthreads();
        for (int i = tid; i < nq; i += blockDim.x)
            if (claim[i] >= 0 && a.nobs[i] > 0) atomicMin(&minClaim[claim[i]], i);
        __syncthreads();
    }
    /* the last claimant of a keypoint owns it (F.mvpMapPoints[bestIdx]=pMP overwrites) */
    const bool checkOri = wants_histogram(a);
    for (int i = tid; i < nq; i += blockDim.x)
        if (claim[i] >= 0) {
            atomicMax(&match[claim[i]], i);
            atomicAdd(&total, 1);
            if (checkOri) atomicAdd(&hist[rot_bin(query_angle(a, i), fi.kps[claim[i]].angle)], 1);
        }
    __syncthreads();
    if (checkOri) {
        if (tid == 0) three_maxima(hist, HISTO_LENGTH, keep[0], keep[1], keep[2]);
        __syncthreads();
        /* every claim in a rejected bin nulls its keypoint and decrements the count (:1455-1465) */
        for (int i = tid; i < nq; i += blockDim.x)
            if (claim[i] >= 0) {
                const int b = rot_bin(query_angle(a, i), fi.kps[claim[i]].angle);
                if (b != keep[0] && b != keep[1] && b != keep[2]) { match[claim[i]] = -2; atomicSub(&total, 1); }
            }
        __syncthreads();
    }
    for (int k = tid; k < fi.n; k += blockDim.x) {
        const int m = match[k];
        if (m >= 0) obs[k] = a.nobs[m];
        else if (m == -2) obs[k] = 0;                  /* the caller sees -2: matched, then removed (:1455-1465) */
        else obs[k] = obs0[k];
    }
    if (tid == 0) *nmatches = total;
}

/* ------------------------------------------------------------------------------------------------
 * ORBmatcher::SearchForTriangulation (:657-823): one warp per keypoint entry of KF1's feature vector
 * ---------------------------------------------------------------------------------------------- */
struct TriArgs {
    const viorb_keypoint *k1, *k2;
    const uint8_t *d1, *d2;
    const float *ur1, *ur2;
    const uint8_t *mp1, *mp2;
    const int *nodeId1, *nodePtr1, *idx1, *nodeId2, *nodePtr2, *idx2;
    int n1, n2, nn1, nn2, nentries1;
    float F12[9];
    float ex, ey;
    float scale2[12], sigma2[12];
    int onlyStereo, checkOri;
};

__device__ __forceinline__ bool check_dist_epipolar(const viorb_keypoint& kp1, const viorb_keypoint& kp2, const float* F,
                                                    const float* sigma2) {
    /* :140-157, left-to-right single-rounded float arithmetic */
    const float a = __fadd_rn(__fadd_rn(__fmul_rn(kp1.x, F[0]), __fmul_rn(kp1.y, F[3])), F[6]);
    const float b = __fadd_rn(__fadd_rn(__fmul_rn(kp1.x, F[1]), __fmul_rn(kp1.y, F[4])), F[7]);
    const float c = __fadd_rn(__fadd_rn(__fmul_rn(kp1.x, F[2]), __fmul_rn(kp1.y, F[5])), F[8]);
    const float num = __fadd_rn(__fadd_rn(__fmul_rn(a, kp2.x), __fmul_rn(b, kp2.y)), c);
    const float den = __fadd_rn(__fmul_rn(a, a), __fmul_rn(b, b));
    if (den == 0) return false;
    const float dsqr = __fdiv_rn(__fmul_rn(num, num), den);
    return (double)dsqr < __dmul_rn(3.84, (double)sigma2[kp2.octave]);
}

__global__ void __launch_bounds__(128) triangulation_kernel(const __grid_constant__ TriArgs a, int* __restrict__ matches12) {
    const int lane = threadIdx.x & 31;
    const int e = blockIdx.x * 4 + (threadIdx.x >> 5);
    if (e >= a.nentries1) return;
    /* node f1 of entry e: last f with nodePtr1[f] <= e */
    int lo = 0, hi = a.nn1 - 1;
    while (lo < hi) {
        const int mid = (lo + hi + 1) >> 1;
        if (a.nodePtr1[mid] <= e) lo = mid; else hi = mid - 1;
    }
    const int node = a.nodeId1[lo];
    /* same node in KF2 (the merge walk of :691-789 visits exactly the common node ids) */
    int l2 = 0, h2 = a.nn2 - 1, f2 = -1;
    while (l2 <= h2) {
        const int mid = (l2 + h2) >> 1;
        const int v = a.nodeId2[mid];
        if (v == node) { f2 = mid; break; }
        if (v < node) l2 = mid + 1; else h2 = mid - 1;
    }
    if (f2 < 0) return;
    const int idx1 = a.idx1[e];
    if (a.mp1[idx1]) return;                                    /* already a MapPoint :704-705 */
    const bool bStereo1 = a.ur1[idx1] >= 0;
    if (a.onlyStereo && !bStereo1) return;
    const viorb_keypoint kp1 = a.k1[idx1];
    const uint8_t* d1 = a.d1 + (size_t)idx1 * 32;
    const int beg = a.nodePtr2[f2], end = a.nodePtr2[f2 + 1];
    /* best = min distance, LAST candidate wins ties (dist > bestDist -> continue, :738) */
    unsigned long long best = ~0ull;
    for (int j = beg + lane; j < end; j += 32) {
        const int idx2 = a.idx2[j];
        if (a.mp2[idx2]) continue;
        const bool bStereo2 = a.ur2[idx2] >= 0;
        if (a.onlyStereo && !bStereo2) continue;
        const int dist = hamming_rows(d1, a.d2 + (size_t)idx2 * 32);
        if (dist > TH_LOW) continue;
        const viorb_keypoint kp2 = a.k2[idx2];
        if (!bStereo1 && !bStereo2) {
            const float distex = __fsub_rn(a.ex, kp2.x), distey = __fsub_rn(a.ey, kp2.y);
            if (__fadd_rn(__fmul_rn(distex, distex), __fmul_rn(distey, distey)) < __fmul_rn(100.f, a.scale2[kp2.octave])) continue;
        }
        if (!check_dist_epipolar(kp1, kp2, a.F12, a.sigma2)) continue;
        const unsigned long long key = ((unsigned long long)dist << 48) | ((unsigned long long)(0xffffff - (j - beg)) << 24) |
                                       (unsigned)idx2;
        best = key < best ? key : best;
    }
    best = warp_min_u64(best);
    if (lane == 0 && best != ~0ull) matches12[idx1] = (int)(best & 0xffffff);
}

__global__ void __launch_bounds__(1024) triangulation_finalize_kernel(const viorb_keypoint* __restrict__ k1,
                                                                      const viorb_keypoint* __restrict__ k2, int n1,
                                                                      int checkOri, int* __restrict__ matches12,
                                                                      int* __restrict__ nmatches) {
    __shared__ int hist[HISTO_LENGTH];
    __shared__ int keep[3];
    __shared__ int total;
    const int tid = threadIdx.x;
    if (tid < HISTO_LENGTH) hist[tid] = 0;
    if (tid == 0) total = 0;
    __syncthreads();
    for (int i = tid; i < n1; i += blockDim.x)
        if (matches12[i] >= 0) {
            atomicAdd(&total, 1);
            if (checkOri) atomicAdd(&hist[rot_bin(k1[i].angle, k2[matches12[i]].angle)], 1);
        }
    __syncthreads();
    if (checkOri) {
        if (tid == 0) three_maxima(hist, HISTO_LENGTH, keep[0], keep[1], keep[2]);
        __syncthreads();
        for (int i = tid; i < n1; i += blockDim.x)
            if (matches12[i] >= 0) {
                const int b = rot_bin(k1[i].angle, k2[matches12[i]].angle);
                if (b != keep[0] && b != keep[1] && b != keep[2]) { matches12[i] = -1; atomicSub(&total, 1); }
            }
        __syncthreads();
    }
    if (tid == 0) *nmatches = total;
}

/* ------------------------------------------------------------------------------------------------
 * MapPoint::ComputeDistinctiveDescriptors (src/MapPoint.cc:249-314), batched: one CTA per map point, one
 * warp per row of the point's N x N Hamming matrix.  The matrix is never stored: the median of row i,
 * vDists[(int)(0.5*(N-1))] of the sorted row (:298-299), is the smallest value whose cumulative count in
 * the row's 257-bin histogram reaches rank+1 (distances are integers in [0, 256]; the diagonal counts as 0,
 * :286).  The first row with the least median wins (strict <, :301-305).
 * ---------------------------------------------------------------------------------------------- */
__global__ void __launch_bounds__(128) distinctive_kernel(const uint8_t* __restrict__ desc, const int* __restrict__ ptr,
                                                          int* __restrict__ best, int* __restrict__ bestMedian) {
    __shared__ unsigned hist[4][288];
    __shared__ unsigned long long red[4];
    const int mp = blockIdx.x, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int beg = ptr[mp], N = ptr[mp + 1] - beg;
    if (N <= 0) {
        if (threadIdx.x == 0) { best[mp] = -1; if (bestMedian) bestMedian[mp] = INT_MAX; }
        return;
    }
    const uint8_t* d0 = desc + (size_t)beg * 32;
    const unsigned rank = (unsigned)(int)(0.5 * (N - 1)) + 1u;
    unsigned long long mine = ~0ull;
    for (int i = warp; i < N; i += 4) {
#pragma unroll
        for (int b = 0; b < 9; b++) hist[warp][lane * 9 + b] = 0;
        __syncwarp();
        for (int j = lane; j < N; j += 32)
            atomicAdd(&hist[warp][i == j ? 0 : hamming_rows(d0 + (size_t)i * 32, d0 + (size_t)j * 32)], 1u);
        __syncwarp();
        unsigned cnt[9], sum = 0;
#pragma unroll
        for (int b = 0; b < 9; b++) { cnt[b] = hist[warp][lane * 9 + b]; sum += cnt[b]; }
        unsigned incl = sum;                                   /* inclusive scan of the per-lane bin sums */
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const unsigned t = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += t;
        }
        const unsigned hit = __ballot_sync(0xffffffffu, incl >= rank);
        const int src = __ffs(hit) - 1;                        /* hit != 0: the last lane's incl == N >= rank */
        int median = 0;
        if (lane == src) {
            unsigned c = incl - sum;
#pragma unroll
            for (int b = 0; b < 9; b++) {
                c += cnt[b];
                if (c >= rank) { median = lane * 9 + b; break; }
            }
        }
        median = __shfl_sync(0xffffffffu, median, src);
        const unsigned long long key = ((unsigned long long)(unsigned)median << 32) | (unsigned)i;
        mine = key < mine ? key : mine;
        __syncwarp();
    }
    if (lane == 0) red[warp] = mine;
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned long long b = red[0];
        for (int w = 1; w < 4; w++) b = red[w] < b ? red[w] : b;
        best[mp] = (int)(b & 0xffffffffu);
        if (bestMedian) bestMedian[mp] = (int)(b >> 32);
    }
}

/* ------------------------------------------------------------------------------------------------
 * ORBmatcher::SearchByBoW(KeyFrame*, Frame&, vector<MapPoint*>&)          (:159-288)  mode 0
 * ORBmatcher::SearchByBoW(KeyFrame*, KeyFrame*, vector<MapPoint*>&)       (:522-655)  mode 1
 * The merge walk over the two FeatureVectors visits exactly the common node ids; a feature belongs to one node,
 * so the "already matched" state (vpMapPointMatches / vbMatched2) of a node's candidates is only touched by the
 * queries of the same node: one warp owns a node pair, runs its queries in list order (the reference's order)
 * and scans the candidates with its lanes (top-2 by warp reduction; ties keep the first list position).
 * ---------------------------------------------------------------------------------------------- */
struct BowArgs {
    const uint8_t *d1, *d2;
    const uint8_t *valid1, *valid2;      /* valid2 may be NULL (mode 0: every frame keypoint is a candidate) */
    const int *nodeId1, *nodePtr1, *idx1, *nodeId2, *nodePtr2, *idx2;
    int nn1, nn2, mode;
    float nnratio;
};

__global__ void __launch_bounds__(128) search_bow_kernel(const __grid_constant__ BowArgs a, int* __restrict__ taken,
                                                         int* __restrict__ match) {
    const int lane = threadIdx.x & 31;
    const int f1 = blockIdx.x * 4 + (threadIdx.x >> 5);
    if (f1 >= a.nn1) return;
    const int node = a.nodeId1[f1];
    int l2 = 0, h2 = a.nn2 - 1, f2 = -1;
    while (l2 <= h2) {
        const int mid = (l2 + h2) >> 1;
        const int v = a.nodeId2[mid];
        if (v == node) { f2 = mid; break; }
        if (v < node) l2 = mid + 1; else h2 = mid - 1;
    }
    if (f2 < 0) return;
    const int beg2 = a.nodePtr2[f2], end2 = a.nodePtr2[f2 + 1];
    for (int e = a.nodePtr1[f1]; e < a.nodePtr1[f1 + 1]; e++) {
        const int i1 = a.idx1[e];
        if (!a.valid1[i1]) continue;                        /* !pMP || pMP->isBad() */
        const uint8_t* d1 = a.d1 + (size_t)i1 * 32;
        unsigned long long k1 = ~0ull, k2 = ~0ull;
        for (int j = beg2 + lane; j < end2; j += 32) {
            const int i2 = a.idx2[j];
            if (taken[i2]) continue;
            if (a.valid2 && !a.valid2[i2]) continue;
            const int dist = hamming_rows(d1, a.d2 + (size_t)i2 * 32);
            if (dist >= 256) continue;                       /* bestDist1 = bestDist2 = 256, strict < */
            const unsigned long long key = ((unsigned long long)dist << 48) | ((unsigned long long)(j - beg2) << 24) | (unsigned)i2;
            if (key < k1) { k2 = k1; k1 = key; }
            else if (key < k2) k2 = key;
        }
        warp_two_min(k1, k2);
        if (k1 == ~0ull) continue;
        const int best1 = (int)(k1 >> 48), best2 = k2 == ~0ull ? 256 : (int)(k2 >> 48), bestIdx = (int)(k1 & 0xffffff);
        const bool ok = (a.mode == 0 ? best1 <= TH_LOW : best1 < TH_LOW) && (float)best1 < __fmul_rn(a.nnratio, (float)best2);
        if (ok && lane == 0) {
            taken[bestIdx] = 1;
            if (a.mode == 0) match[bestIdx] = i1; else match[i1] = bestIdx;
        }
        __syncwarp();
    }
}

/* rotation-histogram filter (:246-265, :633-652) and match count.  match[i] = j pairs keypoint i of set A with
 * keypoint j of set B; rot = angle(first) - angle(second) where first is A when aFirst, else B. */
__global__ void __launch_bounds__(1024) rotation_finalize_kernel(const viorb_keypoint* __restrict__ kA,
                                                                 const viorb_keypoint* __restrict__ kB, int nA, int aFirst,
                                                                 int checkOri, int* __restrict__ match,
                                                                 int* __restrict__ nmatches) {
    __shared__ int hist[HISTO_LENGTH];
    __shared__ int keep[3];
    __shared__ int total;
    const int tid = threadIdx.x;
    if (tid < HISTO_LENGTH) hist[tid] = 0;
    if (tid == 0) total = 0;
    __syncthreads();
    for (int i = tid; i < nA; i += blockDim.x)
        if (match[i] >= 0) {
            atomicAdd(&total, 1);
            if (checkOri) atomicAdd(&hist[aFirst ? rot_bin(kA[i].angle, kB[match[i]].angle) : rot_bin(kB[match[i]].angle, kA[i].angle)], 1);
        }
    __syncthreads();
    if (checkOri) {
        if (tid == 0) three_maxima(hist, HISTO_LENGTH, keep[0], keep[1], keep[2]);
        __syncthreads();
        for (int i = tid; i < nA; i += blockDim.x)
            if (match[i] >= 0) {
                const int b = aFirst ? rot_bin(kA[i].angle, kB[match[i]].angle) : rot_bin(kB[match[i]].angle, kA[i].angle);
                if (b != keep[0] && b != keep[1] && b != keep[2]) { match[i] = -1; atomicSub(&total, 1); }
            }
        __syncthreads();
    }
    if (tid == 0) *nmatches = total;
}

/* ------------------------------------------------------------------------------------------------
 * ORBmatcher::SearchForInitialization                                     (:405-520)
 * Phase 1 (parallel, one warp per level-0 keypoint of F1): the candidates GetFeaturesInArea returns in F2 and
 * their distances, as keys dist | enumeration rank | index.  Phase 2 (one warp, F1 order): the loop-carried
 * part -- vMatchedDistance / vnMatches21 let a later keypoint steal an earlier one's match (:444-470) -- runs
 * over the precomputed keys only.
 * ---------------------------------------------------------------------------------------------- */
__global__ void __launch_bounds__(128) init_candidates_kernel(FrameIndexDev f2, const viorb_keypoint* __restrict__ k1,
                                                              const uint8_t* __restrict__ d1, int n1,
                                                              const float* __restrict__ prev, float window,
                                                              unsigned long long* __restrict__ entries, long long cap,
                                                              int* __restrict__ start, int* __restrict__ count,
                                                              unsigned long long* __restrict__ cursor, int* __restrict__ overflow) {
    __shared__ int slot[4];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int i1 = blockIdx.x * 4 + warp;
    if (i1 >= n1) return;
    const viorb_keypoint kp1 = k1[i1];
    int cnt = 0;
    const int level1 = kp1.octave;
    const float x = prev[2 * i1], y = prev[2 * i1 + 1];
    if (level1 <= 0) for_features_in_area(f2, x, y, window, level1, level1, lane, [&](int, int, int) { cnt++; });
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
    unsigned long long base = 0;
    if (lane == 0) {
        if (cnt) base = atomicAdd(cursor, (unsigned long long)cnt);
        if ((long long)(base + cnt) > cap) { atomicExch(overflow, 1); }
        start[i1] = (int)base;
        count[i1] = (long long)(base + cnt) > cap ? 0 : cnt;
        slot[warp] = 0;
    }
    base = __shfl_sync(0xffffffffu, base, 0);
    __syncwarp();
    if (cnt == 0 || (long long)(base + cnt) > cap) return;
    const uint8_t* dq = d1 + (size_t)i1 * 32;
    for_features_in_area(f2, x, y, window, level1, level1, lane, [&](int pos, int idx, int) {
        const int dist = hamming_rows(dq, f2.desc + (size_t)idx * 32);
        entries[base + atomicAdd(&slot[warp], 1)] = ((unsigned long long)dist << 48) | ((unsigned long long)pos << 24) | (unsigned)idx;
    });
}

__global__ void __launch_bounds__(32) init_match_kernel(const viorb_keypoint* __restrict__ k1, const viorb_keypoint* __restrict__ k2,
                                                        int n1, int n2, const unsigned long long* __restrict__ entries,
                                                        const int* __restrict__ start, const int* __restrict__ count,
                                                        float nnratio, int checkOri, int* __restrict__ matchedDist,
                                                        int* __restrict__ matches21, int* __restrict__ binOf,
                                                        int* __restrict__ matches12, float* __restrict__ prev,
                                                        int* __restrict__ nmatches) {
    __shared__ int hist[HISTO_LENGTH];
    __shared__ int keep[3];
    const int lane = threadIdx.x;
    if (lane < HISTO_LENGTH) hist[lane] = 0;
    for (int i = lane; i < n2; i += 32) { matchedDist[i] = INT_MAX; matches21[i] = -1; }
    for (int i = lane; i < n1; i += 32) { matches12[i] = -1; binOf[i] = -1; }
    __syncwarp();
    int total = 0;                                          /* kept by lane 0 */
    for (int i1 = 0; i1 < n1; i1++) {
        const int c = count[i1];
        if (c == 0) continue;
        const unsigned long long* e = entries + start[i1];
        unsigned long long a = ~0ull, b = ~0ull;
        for (int j = lane; j < c; j += 32) {
            const unsigned long long key = e[j];
            if (matchedDist[(int)(key & 0xffffff)] <= (int)(key >> 48)) continue;       /* :444-445 */
            if (key < a) { b = a; a = key; }
            else if (key < b) b = key;
        }
        warp_two_min(a, b);
        if (a == ~0ull) continue;
        const int bestDist = (int)(a >> 48), bestIdx2 = (int)(a & 0xffffff);
        const float second = b == ~0ull ? (float)INT_MAX : (float)(int)(b >> 48);
        if (bestDist <= TH_LOW && (float)bestDist < __fmul_rn(second, nnratio)) {
            if (lane == 0) {
                if (matches21[bestIdx2] >= 0) { matches12[matches21[bestIdx2]] = -1; total--; }
                matches12[i1] = bestIdx2;
                matches21[bestIdx2] = i1;
                matchedDist[bestIdx2] = bestDist;
                total++;
                if (checkOri) {
                    const int bin = rot_bin(k1[i1].angle, k2[bestIdx2].angle);
                    hist[bin]++;                             /* rotHist keeps stolen entries too (:462-471) */
                    binOf[i1] = bin;
                }
            }
            __syncwarp();
        }
    }
    __syncwarp();
    if (checkOri) {
        if (lane == 0) three_maxima(hist, HISTO_LENGTH, keep[0], keep[1], keep[2]);
        __syncwarp();
        int removed = 0;
        for (int i = lane; i < n1; i += 32) {
            const int bin = binOf[i];
            if (bin >= 0 && bin != keep[0] && bin != keep[1] && bin != keep[2] && matches12[i] >= 0) { matches12[i] = -1; removed++; }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) removed += __shfl_xor_sync(0xffffffffu, removed, o);
        total -= removed;
    }
    __syncwarp();
    for (int i = lane; i < n1; i += 32)                     /* update prev matched (:513-516) */
        if (matches12[i] >= 0) { prev[2 * i] = k2[matches12[i]].x; prev[2 * i + 1] = k2[matches12[i]].y; }
    if (lane == 0) *nmatches = total;
}

/* ------------------------------------------------------------------------------------------------
 * Frame::UndistortKeyPoints (src/Frame.cc:584-614) and Frame::ComputeImageBounds (:616-645):
 * cv::undistortPoints(pts, pts, mK, mDistCoef, cv::Mat(), mK) -- OpenCV's cvUndistortPoints: normalise with the
 * inverse intrinsics, five iterations of x = (x0 - delta(x)) * icdist(x) in double precision, re-project with
 * P = K.  Every double operation is rounded separately (no FMA), in the order of the OpenCV expression.
 * ---------------------------------------------------------------------------------------------- */
__device__ __forceinline__ void undistort_point(const UndistortParams& p, float uf, float vf, float& xo, float& yo) {
    const double u = (double)uf, v = (double)vf;
    double x = __dmul_rn(__dsub_rn(u, p.cx), p.ifx), y = __dmul_rn(__dsub_rn(v, p.cy), p.ify);
    const double x0 = x, y0 = y;
    const double* k = p.k;
    for (int j = 0; j < 5; j++) {
        const double r2 = __dadd_rn(__dmul_rn(x, x), __dmul_rn(y, y));
        const double num = __dadd_rn(1.0, __dmul_rn(__dadd_rn(__dmul_rn(__dadd_rn(__dmul_rn(k[7], r2), k[6]), r2), k[5]), r2));
        const double den = __dadd_rn(1.0, __dmul_rn(__dadd_rn(__dmul_rn(__dadd_rn(__dmul_rn(k[4], r2), k[1]), r2), k[0]), r2));
        const double icdist = __ddiv_rn(num, den);
        if (icdist < 0) {                       /* OpenCV >= 3.4: give up and return the normalised input point */
            x = __dmul_rn(__dsub_rn(u, p.cx), p.ifx);
            y = __dmul_rn(__dsub_rn(v, p.cy), p.ify);
            break;
        }
        const double twox = __dmul_rn(2.0, x), twoy = __dmul_rn(2.0, y);
        double dX = __dmul_rn(__dmul_rn(__dmul_rn(2.0, k[2]), x), y);
        dX = __dadd_rn(dX, __dmul_rn(k[3], __dadd_rn(r2, __dmul_rn(twox, x))));
        dX = __dadd_rn(dX, __dmul_rn(k[8], r2));
        dX = __dadd_rn(dX, __dmul_rn(__dmul_rn(k[9], r2), r2));
        double dY = __dmul_rn(k[2], __dadd_rn(r2, __dmul_rn(twoy, y)));
        dY = __dadd_rn(dY, __dmul_rn(__dmul_rn(__dmul_rn(2.0, k[3]), x), y));
        dY = __dadd_rn(dY, __dmul_rn(k[10], r2));
        dY = __dadd_rn(dY, __dmul_rn(__dmul_rn(k[11], r2), r2));
        x = __dmul_rn(__dsub_rn(x0, dX), icdist);
        y = __dmul_rn(__dsub_rn(y0, dY), icdist);
    }
    /* RR = P * R = K:  xx = fx*x + 0*y + cx, yy = 0*x + fy*y + cy, ww = 1 / (0*x + 0*y + 1) */
    const double xx = __dadd_rn(__dadd_rn(__dmul_rn(p.fx, x), __dmul_rn(0.0, y)), p.cx);
    const double yy = __dadd_rn(__dadd_rn(__dmul_rn(0.0, x), __dmul_rn(p.fy, y)), p.cy);
    const double ww = __ddiv_rn(1.0, __dadd_rn(__dadd_rn(__dmul_rn(0.0, x), __dmul_rn(0.0, y)), 1.0));
    xo = (float)__dmul_rn(xx, ww);
    yo = (float)__dmul_rn(yy, ww);
}

__global__ void __launch_bounds__(128) undistort_kernel(const __grid_constant__ UndistortParams p,
                                                        const viorb_keypoint* __restrict__ in, int n,
                                                        viorb_keypoint* __restrict__ out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    viorb_keypoint kp = in[i];
    if (p.active) undistort_point(p, kp.x, kp.y, kp.x, kp.y);       /* else mvKeysUn = mvKeys (:586-590) */
    out[i] = kp;
}

__global__ void image_bounds_kernel(const __grid_constant__ UndistortParams p, int cols, int rows, float* __restrict__ bounds) {
    /* :616-645: corners (0,0) (cols,0) (0,rows) (cols,rows) */
    if (threadIdx.x != 0) return;
    if (!p.active) { bounds[0] = 0.0f; bounds[1] = (float)cols; bounds[2] = 0.0f; bounds[3] = (float)rows; return; }
    float x[4], y[4];
    undistort_point(p, 0.0f, 0.0f, x[0], y[0]);
    undistort_point(p, (float)cols, 0.0f, x[1], y[1]);
    undistort_point(p, 0.0f, (float)rows, x[2], y[2]);
    undistort_point(p, (float)cols, (float)rows, x[3], y[3]);
    bounds[0] = fminf(x[0], x[2]);      /* mnMinX */
    bounds[1] = fmaxf(x[1], x[3]);      /* mnMaxX */
    bounds[2] = fminf(y[0], y[1]);      /* mnMinY */
    bounds[3] = fmaxf(y[2], y[3]);      /* mnMaxY */
}

/* ------------------------------------------------------------------------------------------------
 * Independent windowed top-1 search of projected map points in a KeyFrame: the inner loops of
 *   ORBmatcher::SearchBySim3 (:1102-1326, both directions, TH_HIGH),
 *   ORBmatcher::Fuse(KeyFrame*, const vector<MapPoint*>&, th) (:825-976, chi-square reprojection gates, TH_LOW),
 *   ORBmatcher::Fuse(KeyFrame*, cv::Mat Scw, ...) (:978-1100, TH_LOW).
 * KeyFrame::GetFeaturesInArea (src/KeyFrame.cc:906-945) enumerates like Frame's; the level filter
 * [nPredictedLevel-1, nPredictedLevel] is applied inside the enumeration (same set, same order).  No query depends
 * on another: one warp per query over the whole grid.
 * ---------------------------------------------------------------------------------------------- */
struct WindowArgs {
    const float *u, *v, *ur;         /* ur == NULL: no reprojection-error gates */
    const int* level;
    const uint8_t* valid;
    const uint8_t* desc;
    int n, thDist;
    float th;
    float invSigma2[12];
};

__global__ void __launch_bounds__(128) search_window_kernel(FrameIndexDev fi, const __grid_constant__ WindowArgs a,
                                                            int* __restrict__ bestIdx, int* __restrict__ bestDist) {
    const int lane = threadIdx.x & 31;
    const int i = blockIdx.x * 4 + (threadIdx.x >> 5);
    if (i >= a.n) return;
    unsigned long long k1 = ~0ull;
    if (a.valid[i]) {
        const float u = a.u[i], v = a.v[i];
        const int lvl = a.level[i];
        const float radius = __fmul_rn(a.th, fi.scale[lvl]);
        const float ur = a.ur ? a.ur[i] : 0.f;
        const uint8_t* dMP = a.desc + (size_t)i * 32;
        for_features_in_area(fi, u, v, radius, lvl - 1, lvl, lane, [&](int pos, int idx, int octave) {
            if (a.ur) {                                               /* Fuse :901-925 */
                const viorb_keypoint kp = fi.kps[idx];
                const float kpr = fi.uRight[idx];
                const float ex = __fsub_rn(u, kp.x), ey = __fsub_rn(v, kp.y);
                float e2 = __fadd_rn(__fmul_rn(ex, ex), __fmul_rn(ey, ey));
                if (kpr >= 0) {
                    const float er = __fsub_rn(ur, kpr);
                    e2 = __fadd_rn(e2, __fmul_rn(er, er));
                    if ((double)__fmul_rn(e2, a.invSigma2[octave]) > 7.8) return;
                } else {
                    if ((double)__fmul_rn(e2, a.invSigma2[octave]) > 5.99) return;
                }
            }
            const int dist = hamming_rows(dMP, fi.desc + (size_t)idx * 32);
            const unsigned long long key = ((unsigned long long)dist << 48) | ((unsigned long long)pos << 24) | (unsigned)idx;
            k1 = key < k1 ? key : k1;
        });
        k1 = warp_min_u64(k1);
    }
    if (lane == 0) {
        const bool ok = k1 != ~0ull && (int)(k1 >> 48) <= a.thDist && (int)(k1 >> 48) < 256;
        bestIdx[i] = ok ? (int)(k1 & 0xffffff) : -1;
        if (bestDist) bestDist[i] = ok ? (int)(k1 >> 48) : INT_MAX;
    }
}

/* SearchBySim3 "check agreement" (:1305-1320) */
__global__ void __launch_bounds__(256) sim3_agree_kernel(const int* __restrict__ m1, const int* __restrict__ m2, int n1,
                                                         int* __restrict__ match12, int* __restrict__ nfound) {
    const int i1 = blockIdx.x * blockDim.x + threadIdx.x;
    bool ok = false;
    int idx2 = -1;
    if (i1 < n1) {
        idx2 = m1[i1];
        ok = idx2 >= 0 && m2[idx2] == i1;
        match12[i1] = ok ? idx2 : -1;
    }
    const unsigned m = __ballot_sync(0xffffffffu, ok);
    if ((threadIdx.x & 31) == 0 && m) atomicAdd(nfound, __popc(m));
}

}  // namespace

/* ------------------------------------------------------------------------------------------------ launchers */
int viorb_launch_stereo(const StereoParams& p, const viorb_keypoint* d_kl, const uint8_t* d_dl,
                        const viorb_keypoint* d_kr, const uint8_t* d_dr, int* d_scratch /* [2]: nMatched, median */,
                        float* d_uRight, float* d_depth, int* d_sad, cudaStream_t s) {
    if (p.nl <= 0) return 0;
    cudaMemsetAsync(d_scratch, 0, 2 * sizeof(int), s);
    stereo_match_kernel<<<(p.nl + 3) / 4, 128, 0, s>>>(p, d_kl, d_dl, d_kr, d_dr, d_uRight, d_depth, d_sad, d_scratch);
    stereo_rank_kernel<<<(p.nl + 127) / 128, 128, 0, s>>>(d_sad, p.nl, d_scratch, d_scratch + 1);
    stereo_filter_kernel<<<(p.nl + 127) / 128, 128, 0, s>>>(d_sad, p.nl, d_scratch + 1, d_uRight, d_depth);
    return 3;
}

int viorb_launch_grid_build(const viorb_keypoint* d_kps, int n, float minX, float minY, float invW, float invH,
                            int* d_cellOf, int* d_cellStart, int* d_cellItems, cudaStream_t s) {
    grid_build_kernel<<<1, 1024, 0, s>>>(d_kps, n, minX, minY, invW, invH, d_cellOf, d_cellStart, d_cellItems);
    return 1;
}

int viorb_launch_features_in_area(const FrameIndexDev& fi, float x, float y, float r, int minLevel, int maxLevel,
                                  unsigned long long* d_keys, int* d_count, cudaStream_t s) {
    cudaMemsetAsync(d_count, 0, sizeof(int), s);
    features_in_area_kernel<<<1, 32, 0, s>>>(fi, x, y, r, minLevel, maxLevel, d_keys, d_count);
    return 1;
}

int viorb_search_scratch_ints(int nq, int nf) {
    /* list keys (2 ints each) + list counts + claim + slow queue per query, minClaim per frame keypoint */
    return nq * (2 * LIST_T + 3) + nf + 64;
}

int viorb_launch_search_local(const FrameIndexDev& fi, const float* projX, const float* projY, const float* projXR,
                              const int* predLevel, const float* viewCos, const uint8_t* valid, const int* nobs,
                              const uint8_t* mpDesc, int nmp, float th, float nnratio, const int* d_obs0, int* d_obs,
                              int* d_scratch, int* d_match, int* d_nmatches, cudaStream_t s) {
    LocalArgs a;
    a.projX = projX; a.projY = projY; a.projXR = projXR; a.viewCos = viewCos; a.predLevel = predLevel;
    a.valid = valid; a.nobs = nobs; a.mpDesc = mpDesc; a.nmp = nmp; a.th = th; a.nnratio = nnratio;
    const int nq = nmp > 0 ? nmp : 0;
    unsigned long long* keys = reinterpret_cast<unsigned long long*>(d_scratch);
    int* count = d_scratch + (size_t)2 * LIST_T * nq;
    int* claim = count + nq;
    int* slow = claim + nq;
    int* minClaim = slow + nq;
    int launches = 1;
    if (nq > 0) {
        candidate_lists_kernel<true, LocalArgs><<<(nq + 3) / 4, 128, 0, s>>>(fi, a, nq, d_obs0, keys, count);
        launches++;
    }
    resolve_kernel<true, LocalArgs><<<1, 1024, 0, s>>>(fi, a, nq, d_obs0, keys, count, d_obs, claim, minClaim, slow, d_match, d_nmatches);
    return launches;
}

int viorb_launch_search_frame(const FrameIndexDev& fi, const float* u, const float* v, const float* invz,
                              const int* lastOctave, const float* lastAngle, const uint8_t* valid, const int* nobs,
                              const uint8_t* mpDesc, int nlast, float th, float mbf, int mode, int checkOri, int thHigh,
                              const int* d_obs0, int* d_obs, int* d_scratch, int* d_match, int* d_nmatches, cudaStream_t s) {
    FrameArgs a;
    a.u = u; a.v = v; a.invz = invz; a.lastAngle = lastAngle; a.lastOctave = lastOctave; a.valid = valid; a.nobs = nobs;
    a.mpDesc = mpDesc; a.nlast = nlast; a.th = th; a.mbf = mbf; a.mode = mode; a.checkOri = checkOri; a.thHigh = thHigh;
    const int nq = nlast > 0 ? nlast : 0;
    unsigned long long* keys = reinterpret_cast<unsigned long long*>(d_scratch);
    int* count = d_scratch + (size_t)2 * LIST_T * nq;
    int* claim = count + nq;
    int* slow = claim + nq;
    int* minClaim = slow + nq;
    int launches = 1;
    if (nq > 0) {
        candidate_lists_kernel<false, FrameArgs><<<(nq + 3) / 4, 128, 0, s>>>(fi, a, nq, d_obs0, keys, count);
        launches++;
    }
    resolve_kernel<false, FrameArgs><<<1, 1024, 0, s>>>(fi, a, nq, d_obs0, keys, count, d_obs, claim, minClaim, slow, d_match, d_nmatches);
    return launches;
}

int viorb_launch_triangulation(const viorb_keypoint* k1, const uint8_t* d1, const float* ur1, const uint8_t* mp1, int n1,
                               const viorb_keypoint* k2, const uint8_t* d2, const float* ur2, const uint8_t* mp2, int n2,
                               const int* nodeId1, const int* nodePtr1, const int* idx1, int nn1, int nentries1,
                               const int* nodeId2, const int* nodePtr2, const int* idx2, int nn2, const float* F12,
                               float ex, float ey, const float* scale2, const float* sigma2, int nlevels, int onlyStereo,
                               int checkOri, int* d_matches12, int* d_nmatches, cudaStream_t s) {
    TriArgs a;
    a.k1 = k1; a.k2 = k2; a.d1 = d1; a.d2 = d2; a.ur1 = ur1; a.ur2 = ur2; a.mp1 = mp1; a.mp2 = mp2;
    a.nodeId1 = nodeId1; a.nodePtr1 = nodePtr1; a.idx1 = idx1; a.nodeId2 = nodeId2; a.nodePtr2 = nodePtr2; a.idx2 = idx2;
    a.n1 = n1; a.n2 = n2; a.nn1 = nn1; a.nn2 = nn2; a.nentries1 = nentries1;
    for (int i = 0; i < 9; i++) a.F12[i] = F12[i];
    a.ex = ex; a.ey = ey;
    for (int i = 0; i < 12; i++) { a.scale2[i] = i < nlevels ? scale2[i] : 0.f; a.sigma2[i] = i < nlevels ? sigma2[i] : 0.f; }
    a.onlyStereo = onlyStereo; a.checkOri = checkOri;
    cudaMemsetAsync(d_matches12, 0xff, (size_t)n1 * sizeof(int), s);
    int launches = 0;
    if (nentries1 > 0 && nn1 > 0 && nn2 > 0) {
        triangulation_kernel<<<(nentries1 + 3) / 4, 128, 0, s>>>(a, d_matches12);
        launches++;
    }
    triangulation_finalize_kernel<<<1, 1024, 0, s>>>(k1, k2, n1, checkOri, d_matches12, d_nmatches);
    return launches + 1;
}

int viorb_launch_distinctive(const uint8_t* d_desc, const int* d_ptr, int nmp, int* d_best, int* d_bestMedian,
                             cudaStream_t s) {
    if (nmp <= 0) return 0;
    distinctive_kernel<<<nmp, 128, 0, s>>>(d_desc, d_ptr, d_best, d_bestMedian);
    return 1;
}

int viorb_launch_search_bow(int mode, const viorb_keypoint* k1, const uint8_t* d1, const uint8_t* valid1, int n1,
                            const viorb_keypoint* k2, const uint8_t* d2, const uint8_t* valid2, int n2, const int* nodeId1,
                            const int* nodePtr1, const int* idx1, int nn1, const int* nodeId2, const int* nodePtr2,
                            const int* idx2, int nn2, float nnratio, int checkOri, int* d_taken, int* d_match,
                            int* d_nmatches, cudaStream_t s) {
    BowArgs a;
    a.d1 = d1; a.d2 = d2; a.valid1 = valid1; a.valid2 = valid2;
    a.nodeId1 = nodeId1; a.nodePtr1 = nodePtr1; a.idx1 = idx1; a.nodeId2 = nodeId2; a.nodePtr2 = nodePtr2; a.idx2 = idx2;
    a.nn1 = nn1; a.nn2 = nn2; a.mode = mode; a.nnratio = nnratio;
    const int nOut = mode == 0 ? n2 : n1;
    cudaMemsetAsync(d_match, 0xff, (size_t)std::max(nOut, 1) * sizeof(int), s);
    cudaMemsetAsync(d_taken, 0, (size_t)std::max(n2, 1) * sizeof(int), s);
    int launches = 0;
    if (nn1 > 0 && nn2 > 0) {
        search_bow_kernel<<<(nn1 + 3) / 4, 128, 0, s>>>(a, d_taken, d_match);
        launches++;
    }
    /* mode 0: match is indexed by the frame keypoint, rot = kpKF.angle - F.mvKeys[idx].angle (:229);
     * mode 1: indexed by KF1's keypoint, rot = vKeysUn1[idx1].angle - vKeysUn2[bestIdx2].angle (:616) */
    if (mode == 0) rotation_finalize_kernel<<<1, 1024, 0, s>>>(k2, k1, n2, 0, checkOri, d_match, d_nmatches);
    else rotation_finalize_kernel<<<1, 1024, 0, s>>>(k1, k2, n1, 1, checkOri, d_match, d_nmatches);
    return launches + 1;
}

int viorb_launch_search_init(const FrameIndexDev& f2, const viorb_keypoint* k1, const uint8_t* d1, int n1, float* d_prev,
                             float window, float nnratio, int checkOri, unsigned long long* d_entries, long long cap,
                             int* d_start, int* d_count, unsigned long long* d_cursor, int* d_overflow, int* d_matchedDist,
                             int* d_matches21, int* d_binOf, int* d_matches12, int* d_nmatches, cudaStream_t s) {
    cudaMemsetAsync(d_cursor, 0, sizeof(unsigned long long), s);
    cudaMemsetAsync(d_overflow, 0, sizeof(int), s);
    int launches = 0;
    if (n1 > 0) {
        init_candidates_kernel<<<(n1 + 3) / 4, 128, 0, s>>>(f2, k1, d1, n1, d_prev, window, d_entries, cap, d_start, d_count,
                                                           d_cursor, d_overflow);
        launches++;
    }
    init_match_kernel<<<1, 32, 0, s>>>(k1, f2.kps, n1, f2.n, d_entries, d_start, d_count, nnratio, checkOri, d_matchedDist,
                                       d_matches21, d_binOf, d_matches12, d_prev, d_nmatches);
    return launches + 1;
}

int viorb_launch_undistort(const UndistortParams& p, const viorb_keypoint* d_in, int n, viorb_keypoint* d_out, cudaStream_t s) {
    if (n <= 0) return 0;
    undistort_kernel<<<(n + 127) / 128, 128, 0, s>>>(p, d_in, n, d_out);
    return 1;
}

int viorb_launch_image_bounds(const UndistortParams& p, int cols, int rows, float* d_bounds, cudaStream_t s) {
    image_bounds_kernel<<<1, 32, 0, s>>>(p, cols, rows, d_bounds);
    return 1;
}

int viorb_launch_search_window(const FrameIndexDev& fi, const float* u, const float* v, const float* ur, const int* level,
                               const uint8_t* valid, const uint8_t* desc, int n, float th, int thDist, const float* invSigma2,
                               int nlevels, int* d_bestIdx, int* d_bestDist, cudaStream_t s) {
    if (n <= 0) return 0;
    WindowArgs a;
    a.u = u; a.v = v; a.ur = ur; a.level = level; a.valid = valid; a.desc = desc; a.n = n; a.thDist = thDist; a.th = th;
    for (int i = 0; i < 12; i++) a.invSigma2[i] = (invSigma2 && i < nlevels) ? invSigma2[i] : 0.f;
    search_window_kernel<<<(n + 3) / 4, 128, 0, s>>>(fi, a, d_bestIdx, d_bestDist);
    return 1;
}

int viorb_launch_sim3_agree(const int* d_m1, const int* d_m2, int n1, int* d_match12, int* d_nfound, cudaStream_t s) {
    cudaMemsetAsync(d_nfound, 0, sizeof(int), s);
    if (n1 <= 0) return 0;
    sim3_agree_kernel<<<(n1 + 255) / 256, 256, 0, s>>>(d_m1, d_m2, n1, d_match12, d_nfound);
    return 1;
}
