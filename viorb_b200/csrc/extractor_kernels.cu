/*
 * extractor_kernels.cu -- hand-written sm_100a kernels for ORBextractor::operator()
 * (reference: src/ORBextractor.cc:1043-1105).  One kernel per stage of the hot path:
 *
 *   pyr_level0_kernel / pyr_resize_kernel   ComputePyramid            :1107-1132  (+ cv::resize, copyMakeBorder)
 *   fast_cells_kernel                       ComputeKeyPointsOctTree   :765-829    (+ cv::FAST x2 per cell)
 *   octree_kernel                           DistributeOctTree         :539-763
 *   blur_levels_kernel                      GaussianBlur :1085-1086 of whole levels               (batches)
 *   describe_blurred_kernel                 IC_Angle :77-104, computeOrbDescriptor :108-147       (batches)
 *   orient_describe_kernel                  IC_Angle, GaussianBlur, computeOrbDescriptor fused per keypoint (single frames)
 *
 * All integer stages are bit-exact restatements; float steps use round-to-nearest single ops without
 * FMA contraction (the file is compiled with -fmad=false and uses __f*_rn where order matters).
 * Nothing here is a dense contraction: no tensor cores, by design.
 */
#include <mutex>
#include <type_traits>

#include "extractor_kernels.cuh"

#include <cuda.h>
#include <stdlib.h>
#include <string.h>

#include "viorb_orb_pattern.h"

namespace {

__device__ __forceinline__ int reflect101(int i, int n) {
    /* BORDER_REFLECT_101 for |overshoot| < n (border 19 << n) */
    if (i < 0) i = -i;
    if (i >= n) i = 2 * n - 2 - i;
    return i;
}

/* ------------------------------------------------------------------------------------------------
 * ComputePyramid level 0: copyMakeBorder(image, temp, 19.., BORDER_REFLECT_101)   (:1127-1128)
 * One CTA = 16 stored rows of one frame.  Interior 16-byte vectors of a stored row are straight 128-bit
 * copies (any source alignment: five 32-bit loads and a byte funnel when the source is not 16-byte aligned);
 * the few vectors per row that hold reflected border pixels or alignment padding are gathered byte by byte
 * in a second, densely packed loop so that the slow path does not diverge the copy warps.
 * ---------------------------------------------------------------------------------------------- */
#define L0_ROWS 16

/* Programmatic dependent launch (sm_90+): a kernel launched with the programmatic-stream-serialization attribute may start
 * while its predecessor in the stream is still running; it runs its prologue (table loads, barrier set-up, zeroing of its
 * own shared memory) and blocks in pdl_wait() until the predecessor has completed and flushed.  Every kernel of a pass
 * calls pdl_trigger() first, which lets its successor be scheduled as soon as all of this grid's CTAs are resident.  Nothing
 * a predecessor writes may be touched, and no global memory written, before pdl_wait().  Both are no-ops in a kernel that
 * was launched without the attribute. */
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

__device__ __forceinline__ uint4 load16_any(const uint8_t* p) {
    const uintptr_t a = reinterpret_cast<uintptr_t>(p);
    const int sh = (int)(a & 3);
    const uint32_t* w = reinterpret_cast<const uint32_t*>(a - sh);
    const uint32_t v0 = __ldg(w), v1 = __ldg(w + 1), v2 = __ldg(w + 2), v3 = __ldg(w + 3);
    const uint32_t v4 = sh ? __ldg(w + 4) : 0u;        /* same 4-byte word as byte 15: never past the needed data */
    const unsigned sel = 0x3210u + 0x1111u * (unsigned)sh;
    return make_uint4(__byte_perm(v0, v1, sel), __byte_perm(v1, v2, sel), __byte_perm(v2, v3, sel), __byte_perm(v3, v4, sel));
}

__device__ __forceinline__ void pyr_level0_tile(const FrameGeom& g, const uint8_t* __restrict__ images, size_t inStep,
                                                size_t frameStride, uint8_t* __restrict__ pyr, int aligned, int rowBlock,
                                                int frame) {
    const LevelGeom& L = g.lv[0];
    const int row0 = rowBlock * L0_ROWS;                        /* first stored row of this tile */
    const int nrows = min(L0_ROWS, L.h + 2 * VIORB_EDGE - row0);
    const int V = L.step >> 4;                                   /* 16-byte vectors per stored row */
    /* interior vectors: x0 = 16*vi - 32 >= 0 and x0 + 16 <= w */
    const int viLo = VIORB_ROI_X0 / 16, viHi = (L.w + VIORB_ROI_X0) / 16 - 1;
    const int nInt = max(viHi - viLo + 1, 0), nBor = V - nInt;
    const uint8_t* img = images + (size_t)frame * frameStride;
    uint8_t* dst0 = pyr + (size_t)frame * g.pyrFrameBytes + L.pyrOff;
    if (nInt > 0) {
        const unsigned inv = 0xffffffffu / (unsigned)nInt + 1u;     /* i / nInt == umulhi(i, inv) for i < 2^16 */
        for (int i = threadIdx.x; i < nrows * nInt; i += blockDim.x) {
            const int r = (int)__umulhi((unsigned)i, inv), vi = viLo + i - r * nInt;
            const uint8_t* src = img + (size_t)reflect101(row0 + r - VIORB_EDGE, L.h) * inStep + (vi * 16 - VIORB_ROI_X0);
            const uint4 v = aligned ? __ldg(reinterpret_cast<const uint4*>(src)) : load16_any(src);
            reinterpret_cast<uint4*>(dst0 + (size_t)(row0 + r) * L.step)[vi] = v;
        }
    }
    const unsigned invB = 0xffffffffu / (unsigned)max(nBor, 1) + 1u;
    for (int i = threadIdx.x; i < nrows * nBor; i += blockDim.x) {
        const int r = (int)__umulhi((unsigned)i, invB), k = i - r * nBor;
        const int vi = (nInt > 0 && k >= viLo) ? k + nInt : k;
        const uint8_t* src = img + (size_t)reflect101(row0 + r - VIORB_EDGE, L.h) * inStep;
        const int x0 = vi * 16 - VIORB_ROI_X0;
        uint32_t wds[4];
#pragma unroll
        for (int q = 0; q < 4; q++) {
            uint32_t word = 0;
#pragma unroll
            for (int j = 0; j < 4; j++) {
                const int x = x0 + 4 * q + j;
                uint32_t v = 0;
                if (x >= -VIORB_EDGE && x < L.w + VIORB_EDGE) v = src[reflect101(x, L.w)];
                word |= v << (8 * j);
            }
            wds[q] = word;
        }
        reinterpret_cast<uint4*>(dst0 + (size_t)(row0 + r) * L.step)[vi] = make_uint4(wds[0], wds[1], wds[2], wds[3]);
    }
}

__global__ void __launch_bounds__(256) pyr_level0_kernel(const __grid_constant__ FrameGeom g,
                                                         const uint8_t* __restrict__ images, size_t inStep,
                                                         size_t frameStride, uint8_t* __restrict__ pyr, int aligned) {
    pdl_trigger();
    pyr_level0_tile(g, images, inStep, frameStride, pyr, aligned, blockIdx.x, blockIdx.y);
}

/* ------------------------------------------------------------------------------------------------
 * ComputePyramid level l > 0: resize(level l-1 ROI -> level l ROI, INTER_LINEAR) then
 * copyMakeBorder(..., BORDER_REFLECT_101 + BORDER_ISOLATED)                     (:1120-1123)
 * cv::resize 8-bit fixed point: T = S[sx]*a0 + S[sx+1]*a1 ; D = (((b0*(T0>>4))>>16) + ((b1*(T1>>4))>>16) + 2) >> 2
 * The border pixels are produced by evaluating the same expression at the reflected ROI coordinate.
 *
 * One CTA = a 128 x 64 tile of the stored (padded) level: the source rectangle it needs (about 1.2x larger)
 * is staged in shared memory with 16-byte loads.  A thread owns four adjacent output columns and walks 16
 * rows.  Horizontal pass of a source row: the <= 8 source bytes the four columns touch are fetched as three
 * aligned words and normalised with two PRMTs; each column's pair (S[sx], S[sx+1]) is one PRMT and its T one
 * IDP.2A with the packed (a0, a1).  Consecutive output rows share a source row, so the lower row's T values
 * stay in registers and become the upper row of the next output row.  Vertical pass: (b * T) >> 16 is one
 * IMAD.HI with b pre-shifted by 16.
 * ---------------------------------------------------------------------------------------------- */
#define RZ_TW 128               /* tile width in stored bytes */
#define RZ_TH 64                /* tile height in stored rows */
#define RZ_WROWS 16             /* rows per warp */
#define RZ_SSTRIDE 224          /* staged source row stride (bytes): 128*1.5 + 1 + 15, rounded up to 16 */
#define RZ_SROWS 100            /* staged source rows: 64*1.5 + 2, padded (scaleFactor <= 1.5) */

/* one 128 x (4*WROWS) tile (bx, by) of level `level` of frame `frame`; 128 threads; src / rowInfo are the CTA's shared
 * buffers.  Everything that depends only on the geometry comes from two host-built tables (c_api.cu build_geometry):
 *   col[3 * word]   per stored word of a level row (4 output columns): {coef[4]} {selP[4]} {lo, hi, okMask, -}
 *                   coef = a0 | a1 << 16, lo = leftmost source byte of the four columns (ROI x of level l-1), hi = last
 *                   source byte, selP = PRMT selector picking (S[sx], S[sx+1]) of a column out of the 8-byte window at lo
 *   row[stored row] {sy | sy+1 << 16 (clamped source rows), b0 << 16, b1 << 16, -}
 * so a tile's setup is five 16-byte loads and four warp reductions (the source rectangle) per thread. */
template <int WROWS>      /* rows per warp: tile height = 4 * WROWS (64 for batches, 16 for the per-frame latency path) */
__device__ __forceinline__ void pyr_resize_tile(const FrameGeom& g, int level, const ResizeTables& t, uint8_t* __restrict__ pyr,
                                                int bx, int by, int frame, uint8_t* src, uint4* rowInfo) {
    constexpr int TH = 4 * WROWS;
    const LevelGeom& L = g.lv[level];
    const LevelGeom& P = g.lv[level - 1];
    const int tid = threadIdx.x, lane = tid & 31;
    uint8_t* base = pyr + (size_t)frame * g.pyrFrameBytes;
    const int stepWords = L.step >> 2, nstored = L.h + 2 * VIORB_EDGE;
    const int wi = bx * (RZ_TW / 4) + lane;                /* stored word of this thread */
    const int row0 = by * TH, r0 = (tid >> 5) * WROWS;     /* first stored row of the tile / of this warp inside it */
    const int rlast = min(TH, nstored - row0) - 1;         /* last tile row that exists */
    const uint4* cd = t.col + 3 * (size_t)(L.xtab + min(wi, stepWords - 1));
    const uint4 coef = __ldg(cd), selP = __ldg(cd + 1), cinfo = __ldg(cd + 2);
    /* row descriptors of the whole tile (every warp reduces the same values: no shared round trip) */
    uint4 rd[(TH + 31) / 32];
    int symin = 0x7fffffff, symax = 0;
#pragma unroll
    for (int k = 0; k < (TH + 31) / 32; k++) {
        const int r = min(TH >= 32 ? lane + 32 * k : (lane & (TH - 1)), rlast);
        rd[k] = __ldg(t.row + L.ytab + row0 + r);
        symin = min(symin, (int)(rd[k].x & 0xffffu));
        symax = max(symax, (int)(rd[k].x >> 16));
    }
    const int sx0 = __reduce_min_sync(0xffffffffu, (int)cinfo.x) & ~15;      /* x origin aligned down to 16 bytes */
    const int sx1 = __reduce_max_sync(0xffffffffu, (int)cinfo.y);
    const int sy0 = __reduce_min_sync(0xffffffffu, symin), sy1 = __reduce_max_sync(0xffffffffu, symax);
    const int nvec = ((sx1 - sx0) >> 4) + 1;                         /* 16-byte vectors per staged row (<= 14) */
    const int nrow = sy1 - sy0 + 1;
    pdl_wait();                                    /* level l-1 is complete from here on */
    {
        const uint8_t* sroi = base + P.pyrOff + (size_t)VIORB_EDGE * P.step + VIORB_ROI_X0;   /* 16-byte aligned */
        const int v = tid & 15;
        if (v < nvec) {
            const uint8_t* gp = sroi + (size_t)(sy0 + (tid >> 4)) * P.step + sx0 + 16 * v;
            unsigned sp = (unsigned)__cvta_generic_to_shared(&src[(tid >> 4) * RZ_SSTRIDE + 16 * v]);
            for (int r = tid >> 4; r < nrow; r += 8, gp += 8 * (size_t)P.step, sp += 8 * RZ_SSTRIDE)
                asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(sp), "l"(gp) : "memory");
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
    }
    if (tid < 32) {
#pragma unroll
        for (int k = 0; k < (TH + 31) / 32; k++) {
            const int r = lane + 32 * k;
            if (r < TH) {
                const unsigned o0 = ((rd[k].x & 0xffffu) - (unsigned)sy0) * RZ_SSTRIDE, o1 = ((rd[k].x >> 16) - (unsigned)sy0) * RZ_SSTRIDE;
                rowInfo[r] = make_uint4(o0 | (o1 << 16), rd[k].y, rd[k].z, 0u);
            }
        }
    }
    const int lo = (int)cinfo.x - sx0;
    const unsigned selN = 0x3210u + 0x1111u * (unsigned)(lo & 3);  /* normalise: window byte 0 = staged byte lo */
    const unsigned okMask = cinfo.z;
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    __syncthreads();
    if (wi >= stepWords) return;
    const unsigned* srcw = reinterpret_cast<const unsigned*>(src) + (lo >> 2);
    /* horizontal pass of the staged row at byte offset `off`: T[j] = (S[sx]*a0 + S[sx+1]*a1) >> 4 */
    auto hpass = [&](unsigned off, unsigned (&T)[4]) {
        const unsigned* w = srcw + (off >> 2);
        const unsigned W0 = __byte_perm(w[0], w[1], selN), W1 = __byte_perm(w[1], w[2], selN);
        T[0] = __dp2a_lo(coef.x, __byte_perm(W0, W1, selP.x), 0u) >> 4;
        T[1] = __dp2a_lo(coef.y, __byte_perm(W0, W1, selP.y), 0u) >> 4;
        T[2] = __dp2a_lo(coef.z, __byte_perm(W0, W1, selP.z), 0u) >> 4;
        T[3] = __dp2a_lo(coef.w, __byte_perm(W0, W1, selP.w), 0u) >> 4;
    };
    /* vertical pass + packing of four pixels: s = ((b0*T0)>>16) + ((b1*T1)>>16) + 2 <= 1022, result s >> 2.  The two
     * high products are plain IMAD.HI (no 64-bit addend to set up); 64*s is formed with two IMADs so that (s >> 2) is
     * byte 1 of the word and three PRMTs gather the four result bytes (the arithmetic stays on the FMA pipe, the ALU
     * pipe -- the busier one in this kernel -- only does the byte picking) */
    auto vrow = [&](const uint4& ri, const unsigned (&U)[4], const unsigned (&D)[4]) -> uint32_t {
        unsigned sj[4];
#pragma unroll
        for (int j = 0; j < 4; j++) sj[j] = __umulhi(ri.y, U[j]) * 64u + (__umulhi(ri.z, D[j]) * 64u + 128u);
        return __byte_perm(__byte_perm(sj[0], sj[1], 0x0051), __byte_perm(sj[2], sj[3], 0x0051), 0x5410) & okMask;
    };
    /* Consecutive output rows mostly share a source row (sy advances by 1 or 2 per output row): the lower row's T of
     * one output row is the upper row's T of the next.  Two rows per iteration with the two register sets swapping
     * roles, so the reuse costs no register moves. */
    unsigned A[4] = {0, 0, 0, 0}, B[4] = {0, 0, 0, 0};
    unsigned inA = 0xffffffffu;                       /* staged offset whose T is held in A */
    uint32_t* out = reinterpret_cast<uint32_t*>(base + L.pyrOff + (size_t)(row0 + r0) * L.step) + wi;
#pragma unroll 1
    for (int r = r0; r < r0 + WROWS; r += 2, out += 2 * stepWords) {
        if (r > rlast) break;
        const uint4 ri0 = rowInfo[r];
        const unsigned o0 = ri0.x & 0xffffu, o1 = ri0.x >> 16;
        if (o0 != inA) hpass(o0, A);
        hpass(o1, B);
        out[0] = vrow(ri0, A, B);
        if (r + 1 > rlast) break;
        const uint4 ri1 = rowInfo[r + 1];
        const unsigned q0 = ri1.x & 0xffffu, q1 = ri1.x >> 16;
        if (q0 != o1) hpass(q0, B);
        hpass(q1, A);
        out[stepWords] = vrow(ri1, B, A);
        inA = q1;
    }
}

template <int WROWS>
__global__ void __launch_bounds__(128) pyr_resize_kernel(const __grid_constant__ FrameGeom g, int level,
                                                         ResizeTables t, uint8_t* __restrict__ pyr) {
    __shared__ __align__(16) uint8_t src[(4 * WROWS * 3 / 2 + 4) * RZ_SSTRIDE];
    __shared__ uint4 rowInfo[4 * WROWS];      /* {staged byte offset of row sy | of row sy+1 << 16, b0 << 16, b1 << 16, -} */
    pdl_trigger();
    pyr_resize_tile<WROWS>(g, level, t, pyr, blockIdx.x, blockIdx.y, blockIdx.z, src, rowInfo);
}

/* ------------------------------------------------------------------------------------------------
 * FAST-9/16 per cell with the iniThFAST -> minThFAST retry (:789-829, cv::FAST TYPE_9_16 + NMS).
 *
 * One CTA per cell.  The cell sub-image [iniX, maxX) x [iniY, maxY) is staged in shared memory with its
 * detection window (inset 3: cv::FAST never tests the 3-pixel rim) on a 4-byte boundary.  Each thread
 * scores FOUR horizontally adjacent pixels at once: the 16 ring samples are fetched as 32-bit words,
 * split into two s16x2 registers (pixels 0/2 and 1/3) and the threshold-independent corner score
 *     S = max( max_arc min_{9} d , max_arc min_{9} -d ) - 1          (cornerScore<16>)
 * is evaluated with the packed 3-input DPX min/max (VIMNMX3.S16x2): min9 = min3(min3,min3,min3).
 * A pixel is a corner at threshold t iff S >= t.  cv::FAST's 3x3 non-max suppression keeps a corner iff
 * its score is strictly greater than all 8 neighbours, where non-corners and pixels outside the window
 * count 0 -- i.e. "S_p >= t and S_p > S_n for all window neighbours", so one local-max map serves both
 * thresholds and the per-cell retry (:812) only re-filters by minThFAST.
 * ---------------------------------------------------------------------------------------------- */
#define FAST_GROUP 4            /* horizontally adjacent cells per CTA */
#ifndef FAST_MIN_CTAS
#define FAST_MIN_CTAS 9
#endif
#define FAST_ROWS VIORB_FAST_TILE_ROWS   /* cell sub-image rows  (hCell + 6 <= 66) */
#define FAST_MAXQ 45            /* quads per window row: 4 cells x wCell (<= 45 when nCols >= 2; one cell of <= 59 otherwise) */
#define FAST_TW (VIORB_FAST_TILE_BYTES / 4)   /* tile row stride in words = the TMA box width: 1 lead word + 45 quads + 1 tail word */
#define FAST_STRIP 8           /* rows per phase-0 task (a vertical strip of one quad column) */
#define FAST_SCW 49             /* score row stride in words (odd): 1 zero word + 45 quads + 1 zero word, padded */

__device__ __forceinline__ unsigned funnel_bytes(unsigned lo, unsigned hi, int sh) {
    /* bytes sh..sh+3 of the 8-byte little-endian sequence lo|hi (sh in 0..3) */
    return __byte_perm(lo, hi, 0x3210u + 0x1111u * (unsigned)sh);
}

/* shared-memory atomic add issued as written: the callers already elect one lane per warp, and the compiler's own
 * warp-aggregation wrapper (VOTEU/FLO/UPOPC around ATOMS) would only add instructions */
__device__ __forceinline__ int smem_add(int* p, int v) {
    int old;
    asm volatile("atom.shared.add.u32 %0, [%1], %2;" : "=r"(old) : "r"((unsigned)__cvta_generic_to_shared(p)), "r"(v) : "memory");
    return old;
}

/* ---- TMA (cp.async.bulk.tensor) + mbarrier, raw PTX ---- */
__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long* bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");      /* the init must be visible to the async proxy */
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long* bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, unsigned parity) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "MBAR_WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@!p bra MBAR_WAIT_%=;\n\t}"
        ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
/* one 3-D box {x, y, z} of the tensor described by `map` into shared memory; completes `bar` with the box bytes */
__device__ __forceinline__ void tma_load_3d(void* dst, const void* map, int x, int y, int z, unsigned long long* bar) {
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
        ::"r"(smem_u32(dst)), "l"(map), "r"(x), "r"(y), "r"(z), "r"(smem_u32(bar)) : "memory");
}

/* Four adjacent pixels x+dx .. x+dx+3 of the window row `r` (r = word holding window bytes 4q .. 4q+3 when the
 * shift is 0).  The TMA box starts on a 16-byte boundary of the stored row, so the detection window starts at
 * byte SH (0..3) of a tile word; S = SH + dx is a compile-time constant and the access folds to one aligned
 * word (S % 4 == 0) or two words and one PRMT. */
template <int S>
__device__ __forceinline__ unsigned fast_ld4(const unsigned* r) {
    constexpr int w = S >= 0 ? S / 4 : -((3 - S) / 4);
    constexpr int b = S - 4 * w;
    if (b == 0) return r[w];
    return funnel_bytes(r[w], r[w + 1], b);
}

/* ring sample k of four adjacent pixels as two s16x2 registers (A = pixels 0,2; B = pixels 1,3), raw grey
 * values; ring offsets are compile-time constants. */
#define FAST_RING_LIST(OP)                                                                            \
    OP(0, 3, 0) OP(1, 3, 1) OP(2, 2, 2) OP(3, 1, 3) OP(4, 0, 3) OP(5, -1, 3) OP(6, -2, 2) OP(7, -3, 1)    \
    OP(8, -3, 0) OP(9, -3, -1) OP(10, -2, -2) OP(11, -1, -3) OP(12, 0, -3) OP(13, 1, -3) OP(14, 2, -2) OP(15, 3, -1)

template <int SH>
__device__ __forceinline__ void fast_load_ring(const unsigned* row, unsigned (&rA)[16], unsigned (&rB)[16]) {
#define RING(k, dy, dx)                                                                                \
    {                                                                                                  \
        const unsigned w_ = fast_ld4<SH + (dx)>(row + (dy) * FAST_TW);                                 \
        rA[k] = __byte_perm(w_, 0, 0x4240);                                                            \
        rB[k] = __byte_perm(w_, 0, 0x4341);                                                            \
    }
    FAST_RING_LIST(RING)
#undef RING
}

/* cornerScore<16> of two pixels (s16 lanes) from the raw ring samples r[k] and the centre value v:
 *   S = max( max_arc min_9 (r - v), max_arc min_9 (v - r) ) - 1 = max(a - v, v - b) - 1
 * with a = max_arc min_9 r, b = min_arc max_9 r  (the centre is constant over the ring, so the 9-window
 * min/max trees run on the raw samples: min9 = min3(min3, min3, min3) with VIMNMX3.S16x2).
 * Returns relu(S - (minTh - 1)): zero iff the pixel is not a corner at minTh, order preserving otherwise. */
__device__ __forceinline__ unsigned fast_score_s16x2(const unsigned (&r)[16], unsigned v, unsigned negBias) {
    unsigned mn3[16], mx3[16];
#pragma unroll
    for (int k = 0; k < 16; k++) {
        mn3[k] = __vimin3_s16x2(r[k], r[(k + 1) & 15], r[(k + 2) & 15]);
        mx3[k] = __vimax3_s16x2(r[k], r[(k + 1) & 15], r[(k + 2) & 15]);
    }
    unsigned mn9[16], mx9[16];
#pragma unroll
    for (int k = 0; k < 16; k++) {
        mn9[k] = __vimin3_s16x2(mn3[k], mn3[(k + 3) & 15], mn3[(k + 6) & 15]);
        mx9[k] = __vimax3_s16x2(mx3[k], mx3[(k + 3) & 15], mx3[(k + 6) & 15]);
    }
    unsigned a = __vimax3_s16x2(mn9[0], mn9[1], mn9[2]), b = __vimin3_s16x2(mx9[0], mx9[1], mx9[2]);
#pragma unroll
    for (int k = 3; k < 15; k += 2) {
        a = __vimax3_s16x2(a, mn9[k], mn9[k + 1]);
        b = __vimin3_s16x2(b, mx9[k], mx9[k + 1]);
    }
    a = __vmaxs2(a, mn9[15]);
    b = __vmins2(b, mx9[15]);
    const unsigned best = __vmaxs2(__vsub2(a, v), __vsub2(v, b));          /* max(a - v, v - b) = S + 1 */
    return __viaddmax_s16x2_relu(best, negBias, 0u);                         /* relu(S + 1 - minTh) */
}

template <int SH>
__global__ void __launch_bounds__(128, FAST_MIN_CTAS) fast_cells_kernel(const __grid_constant__ FrameGeom g,
                                                         const __grid_constant__ TmaMaps maps,
                                                         const int4* __restrict__ groups,
                                                         uint32_t* __restrict__ cand, int* __restrict__ candCount,
                                                         int* __restrict__ status) {
    /* dynamic shared memory, sized by the host for this geometry (FrameGeom::fast*):
     *   [ tile: fastTileRows x 208 B | work0: fastMaxWork u16, later the staged candidate records | pad ]  = fastPixBytes
     *   [ sc: (fastTileRows - 4) x 196 B ] [ work: fastMaxWork u16 ] */
    extern __shared__ __align__(128) unsigned char raw[];
    unsigned* sc = reinterpret_cast<unsigned*>(raw + g.fastPixBytes);
    unsigned short* work = reinterpret_cast<unsigned short*>(raw + g.fastPixBytes + (g.fastTileRows - 4) * FAST_SCW * 4);   /* quads that survive the high-speed test */
    __shared__ __align__(8) unsigned long long bar;   /* mbarrier of the TMA tile load */
    __shared__ int nwork, nwork0, nout, nqLo, nqHi;
    __shared__ unsigned char qlist[64];               /* retry pass: quad columns under work (first warp's at 0, second's at 32) */
    __shared__ int cellCnt[FAST_GROUP];               /* per cell: local maxima found by the pass */
    unsigned* tile = reinterpret_cast<unsigned*>(raw);
    unsigned short* work0 = reinterpret_cast<unsigned short*>(raw + g.fastTileRows * FAST_TW * 4);   /* non-flat quads */
    uint32_t* stage = reinterpret_cast<uint32_t*>(raw + g.fastTileRows * FAST_TW * 4);   /* the pass's candidate records (work0 is dead by then; the tile stays for the retry pass) */
    const int frame = blockIdx.y;
    /* A CTA owns up to FAST_GROUP horizontally adjacent cells of one cell row.  Everything that depends only on the
     * geometry comes precomputed from the host (c_api.cu build_geometry), two 16-byte words per group:
     *   {level | cells << 8 | wCell << 16, ww | wh << 16, NQ | w0 << 8 | boxH << 16, 2^32 / NQ}
     *   {boxX, first stored row of the box, candBase, recBase = (3 + j0 * wCell) | (3 + i * hCell) << 12}
     * ww x wh = detection window of the group (the cells tile it exactly: cell j detects x in [j*wCell, (j+1)*wCell), the
     * last one is clipped at maxBorderX, :798-806; skipped cells, :795-796, :804-805, and empty windows are not listed). */
    const int4 ga = __ldg(&groups[2 * blockIdx.x]), gb = __ldg(&groups[2 * blockIdx.x + 1]);
    const int l = ga.x & 0xff, ncell = (ga.x >> 8) & 0xff, wC = ga.x >> 16;
    const int ww = ga.y & 0xffff, wh = ga.y >> 16;
    const int NQ = ga.z & 0xff, w0 = (ga.z >> 8) & 0xff, boxH = ga.z >> 16;      /* quads per window row, first window word of a tile row */
    const unsigned invNQ = (unsigned)ga.w;                                       /* t / NQ == umulhi(t, invNQ) for t < 2^16 */
    const int tid = threadIdx.x, lane = tid & 31;
    /* The tile first: one TMA box per CTA, requested before the rest of the set-up.  The box starts at the 16-byte boundary
     * at or below window x = -3 of the stored row (TMA needs a 16-byte aligned start), so window pixel x of row y sits at
     * tile byte 4*w0 + SH + x of row y, SH = (stored byte of window x=0) & 3 -- the same for all groups of a launch (the
     * host sorts the groups by SH; with 4-cell groups every window starts at byte 51 + 4k*wCell, SH = 3).  Bytes of the box
     * outside the stored level read as 0 and are never used. */
    pdl_trigger();
    if (tid == 0) {
        mbar_init(&bar, 1);
        pdl_wait();                           /* the pyramid is complete from here on */
        mbar_expect_tx(&bar, (unsigned)(VIORB_FAST_TILE_BYTES * boxH));
        tma_load_3d(tile, &maps.fast[l], gb.x, gb.y, frame, &bar);
    }
    const LevelGeom& L = g.lv[l];
    const unsigned lt = (1u << lane) - 1;
    uint32_t* out = cand + (size_t)frame * g.candPerFrame + gb.z;
    const uint32_t recBase = (uint32_t)gb.w;
    pdl_wait();                               /* every thread: the candidate pool is written below */

    /* Two passes, like the reference (:808-816): pass 0 runs cv::FAST at iniThFAST on all cells of the group; a cell
     * that returns nothing is run again at minThFAST in pass 1 (about one cell in ten, so most CTAs stop after
     * pass 0 and never pay for the weak-corner work).  Both passes use the same code with a different threshold. */
    unsigned retry = (1u << ncell) - 1;       /* cells the pass works on */
#pragma unroll 1
    for (int pass = 0; pass < 2; pass++) {
        const int th = pass == 0 ? g.iniTh : g.minTh;
        if (tid == 0) { nwork = 0; nwork0 = 0; nout = 0; }
        {   /* zero the score rows with 16-byte stores (the array starts on a 128-byte boundary; a few words past the
             * last needed row stay inside the (fastTileRows - 4)-row array) */
            uint4* sc4 = reinterpret_cast<uint4*>(sc);
            const int n4 = min(((wh + 2) * FAST_SCW + 3) >> 2, ((g.fastTileRows - 4) * FAST_SCW) >> 2);
            for (int i = tid; i < n4; i += blockDim.x) sc4[i] = make_uint4(0u, 0u, 0u, 0u);
            for (int i = 4 * n4 + tid; i < (wh + 2) * FAST_SCW; i += blockDim.x) sc[i] = 0;
        }
        if (tid < FAST_GROUP) cellCnt[tid] = 0;
        if (pass && tid < 64) {
            /* retry pass: the pre-test only visits the quad columns that touch a cell under retry (one cell in ten is, so
             * the dense task grid of pass 0 would run at a quarter of its lanes) -- warps 0 and 1 list them */
            const int q = tid;
            bool on = false;
            if (q < NQ) {
                const int x = 4 * q, x3 = min(x + 3, ww - 1);
                const int c0 = (x >= wC) + (x >= 2 * wC) + (x >= 3 * wC), c3 = (x3 >= wC) + (x3 >= 2 * wC) + (x3 >= 3 * wC);
                on = (((retry >> c0) | (retry >> c3)) & 1u) != 0;
            }
            const unsigned m = __ballot_sync(0xffffffffu, on);
            if (tid < 32) nqLo = __popc(m);
            else nqHi = __popc(m);
            if (on) qlist[(tid < 32 ? 0 : 32) + __popc(m & lt)] = (unsigned char)q;
        }
        __syncthreads();
        if (pass == 0) mbar_wait(&bar, 0u);        /* the tile has landed (it stays intact for the retry pass) */

        /* phase 0 -- byte-domain pre-test on the four compass samples (k = 0, 4, 8, 12) of every quad: every 9-arc
         * of the ring contains two ADJACENT compass samples, so a pixel can only be a corner at th if
         * |ring - centre| > th at two adjacent compass points.  flag byte bit 7 = (|d| > th); carries between
         * bytes can only add false positives.  Surviving quads are compacted so the later phases run on dense warps.
         * A task is a vertical strip of FAST_STRIP rows of one quad column: the centre words of its FAST_STRIP + 6 tile
         * rows are loaded once, the vertical difference |C[j] - C[j+3]| serves row j as its south and row j+3 as its
         * north sample, and only the east / west samples are fetched per row.  One shared atomic per warp and strip. */
        {
            const unsigned addc = (unsigned)(127 - th) * 0x01010101u;
            const int nq0 = pass ? nqLo : NQ, nqe = pass ? nq0 + nqHi : NQ;        /* quad columns this pass visits */
            const unsigned invNQe = pass ? 0xffffffffu / (unsigned)nqe + 1u : invNQ;
            const int nstrip = (wh + FAST_STRIP - 1) / FAST_STRIP, ntask0 = nqe * nstrip;
            for (int t0 = 0; t0 < ntask0; t0 += blockDim.x) {
                const int t = t0 + tid;
                unsigned keepMask = 0;                  /* bit r: window row ys + r of this quad column survives */
                int q = 0, ys = 0;
                const bool active = t < ntask0;
                if (active) {
                    const int sidx = nqe > 1 ? (int)__umulhi((unsigned)t, invNQe) : t;
                    q = t - sidx * nqe;
                    if (pass) q = qlist[q < nq0 ? q : 32 + q - nq0];
                    ys = sidx * FAST_STRIP;
                    {
                        /* tile row ys + j holds window row ys + j - 3; rows past the tile read the lists that follow it in
                         * shared memory and only reach rows >= wh, which are masked below */
                        const unsigned* col = &tile[ys * FAST_TW + w0 + q];
                        unsigned C[FAST_STRIP + 6], fV[FAST_STRIP + 3];
#pragma unroll
                        for (int j = 0; j < FAST_STRIP + 6; j++) C[j] = fast_ld4<SH>(col + j * FAST_TW);
#pragma unroll
                        for (int j = 0; j < FAST_STRIP + 3; j++) {
                            const unsigned d = __vabsdiffu4(C[j], C[j + 3]);
                            fV[j] = (d + addc) | d;
                        }
#pragma unroll
                        for (int r = 0; r < FAST_STRIP; r++) {
                            const unsigned* row = col + (r + 3) * FAST_TW;
                            const unsigned dE = __vabsdiffu4(fast_ld4<SH + 3>(row), C[r + 3]);
                            const unsigned dW = __vabsdiffu4(fast_ld4<SH - 3>(row), C[r + 3]);
                            const unsigned fEW = (dE + addc) | dE | (dW + addc) | dW;
                            /* (f0&f4)|(f4&f8)|(f8&f12)|(f12&f0) = (f0|f8) & (f4|f12); the row's bit without a predicate */
                            keepMask += min(((fV[r] | fV[r + 3]) & fEW) & 0x80808080u, 1u) << r;
                        }
                        const int rowsLeft = wh - ys;                 /* rows past the window do not exist */
                        if (rowsLeft < FAST_STRIP) keepMask &= (1u << rowsLeft) - 1u;
                    }
                }
                /* compaction: warp scan of the per-thread counts (five shuffles), one shared atomic per warp, each thread
                 * appends its rows */
                const int n = __popc(keepMask);
                int incl = n;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                    const int up = __shfl_up_sync(0xffffffffu, incl, o);
                    if (lane >= o) incl += up;
                }
                const int total = __shfl_sync(0xffffffffu, incl, 31);
                if (total == 0) continue;
                int basePos = 0;
                if (lane == 31) basePos = smem_add(&nwork0, total);
                int pos = __shfl_sync(0xffffffffu, basePos, 31) + incl - n;
                const int tb = (ys << 8) | q;                /* list entry of window row ys: q | y << 8 */
#pragma unroll
                for (int r = 0; r < FAST_STRIP; r++) {
                    if ((keepMask >> r) & 1u) work0[pos++] = (unsigned short)(tb + (r << 8));
                }
            }
        }
        __syncthreads();

        /* phase 1 -- segment test without signs on the non-flat quads, still in the byte domain (four pixels per
         * operation, no splitting into 16-bit lanes): a corner at th has 9 contiguous ring samples that all differ from the
         * centre by more than th (all brighter or all darker; ignoring which only admits a few more quads: 11.9 % of all
         * quads pass on the benchmark frames, 11.0 % pass the signed opposite-pair test used before, 9.5 % hold a corner).
         * f[k] bit 7 = |ring k - centre| > th; 9 in a row = and3 of and3s, any start = or over the 16 arcs. */
        const unsigned thP = (unsigned)th * 0x00010001u;
        {
            const unsigned addc = (unsigned)(127 - th) * 0x01010101u;
            const int n0 = nwork0;
            for (int i0 = 0; i0 < n0; i0 += blockDim.x) {
                const int i = i0 + tid;
                bool keep = false;
                int t = 0;
                if (i < n0) {
                    t = work0[i];
                    const int y = t >> 8, q = t & 0xff;
                    const unsigned* row = &tile[(y + 3) * FAST_TW + w0 + q];
                    const unsigned cw4 = fast_ld4<SH>(row);
                    unsigned f[16];
#define RING(k, dy, dx)                                                                               \
                    {                                                                                  \
                        const unsigned d_ = __vabsdiffu4(fast_ld4<SH + (dx)>(row + (dy) * FAST_TW), cw4); \
                        f[k] = (d_ + addc) | d_;                                                       \
                    }
                    FAST_RING_LIST(RING)
#undef RING
                    unsigned a3[16];
#pragma unroll
                    for (int k = 0; k < 16; k++) a3[k] = f[k] & f[(k + 1) & 15] & f[(k + 2) & 15];
                    unsigned a9[16];
#pragma unroll
                    for (int k = 0; k < 16; k++) a9[k] = a3[k] & a3[(k + 3) & 15] & a3[(k + 6) & 15];
                    const unsigned o0 = a9[0] | a9[1] | a9[2], o1 = a9[3] | a9[4] | a9[5], o2 = a9[6] | a9[7] | a9[8];
                    const unsigned o3 = a9[9] | a9[10] | a9[11], o4 = a9[12] | a9[13] | a9[14];
                    keep = (((o0 | o1 | o2) | (o3 | o4 | a9[15])) & 0x80808080u) != 0;
                }
                const unsigned m = __ballot_sync(0xffffffffu, keep);
                int basePos = 0;
                if (lane == 0 && m) basePos = smem_add(&nwork, __popc(m));
                basePos = __shfl_sync(0xffffffffu, basePos, 0);
                if (keep) work[basePos + __popc(m & lt)] = (unsigned short)t;
            }
        }
        __syncthreads();

        /* phase 2 -- exact cornerScore on the surviving quads only, densely packed over the CTA.
         * sc holds relu(S - (th - 1)): 0 = not a corner at th */
        const unsigned negBias = __vneg2(thP);
        const int nw = nwork;
        for (int i = tid; i < nw; i += blockDim.x) {
            const int t = work[i];
            const int y = t >> 8, q = t & 0xff;
            const unsigned* row = &tile[(y + 3) * FAST_TW + w0 + q];
            unsigned rA[16], rB[16];
            fast_load_ring<SH>(row, rA, rB);
            const unsigned cw4 = fast_ld4<SH>(row);
            const unsigned sA = fast_score_s16x2(rA, __byte_perm(cw4, 0, 0x4240), negBias);
            const unsigned sB = fast_score_s16x2(rB, __byte_perm(cw4, 0, 0x4341), negBias);
            unsigned word = sA | (sB << 8);                       /* bytes = pixels x, x+1, x+2, x+3 */
            const int x = q * 4;
            if (x + 4 > ww) word &= 0xffffffffu >> (8 * (x + 4 - ww));   /* beyond the window: not a corner */
            sc[(y + 1) * FAST_SCW + q + 1] = word;
        }
        __syncthreads();

        /* 3x3 non-max suppression and output, one thread per scored quad.  cv::FAST runs on the cell's own sub-image, so
         * neighbours in another cell (or outside the window) count as 0.  The nine score words around the quad are split
         * into u16 lanes (pixels 0,2 / 1,3) and reduced with the packed 3-input max: first over the three rows (per
         * column), then over the left and right neighbour columns, which are the same words shifted by one pixel.  A
         * pixel survives iff its score exceeds all eight neighbours (scores are > 0 exactly for the corners at th, and
         * every local maximum of the cells this pass works on is a keypoint of its cell, :808-816).  The records
         * are staged in shared memory. */
        for (int i0 = 0; i0 < nw; i0 += blockDim.x) {
            const int i = i0 + tid;
            unsigned lm = 0;                 /* bit k: pixel k of the quad is a local maximum of a cell under work */
            unsigned cw = 0;
            int x0 = 0, y = 0, cg0 = 0, jb = 8;
            if (i < nw) {
                const int t = work[i];
                y = t >> 8;
                const int q = t & 0xff;
                const unsigned* c = &sc[(y + 1) * FAST_SCW + q + 1];
                cw = c[0];
                if (cw) {
                    x0 = 4 * q;
                    cg0 = (x0 >= wC) + (x0 >= 2 * wC) + (x0 >= 3 * wC);
                    jb = wC - (x0 - cg0 * wC);       /* pixel jb of the quad is the first column of the next cell (jb >= 1) */
                    const unsigned ul = c[-FAST_SCW - 1], uc = c[-FAST_SCW], ur = c[-FAST_SCW + 1];
                    const unsigned ml = c[-1], mr = c[1];
                    const unsigned dl = c[FAST_SCW - 1], dc = c[FAST_SCW], dr = c[FAST_SCW + 1];
                    /* vertical maxima per column, u16 lanes: A = pixels 0,2, B = pixels 1,3 of a word */
                    const unsigned lB = __vimax3_u16x2(__byte_perm(ul, 0, 0x4341), __byte_perm(ml, 0, 0x4341), __byte_perm(dl, 0, 0x4341));
                    const unsigned rA = __vimax3_u16x2(__byte_perm(ur, 0, 0x4240), __byte_perm(mr, 0, 0x4240), __byte_perm(dr, 0, 0x4240));
                    const unsigned cA = __byte_perm(cw, 0, 0x4240), cB = __byte_perm(cw, 0, 0x4341);
                    const unsigned udA = __vmaxu2(__byte_perm(uc, 0, 0x4240), __byte_perm(dc, 0, 0x4240));
                    const unsigned udB = __vmaxu2(__byte_perm(uc, 0, 0x4341), __byte_perm(dc, 0, 0x4341));
                    const unsigned vA = __vmaxu2(udA, cA), vB = __vmaxu2(udB, cB);      /* the column itself, three rows */
                    /* neighbour columns: left of pixels (0,2) = (pixel 3 of the left word, pixel 1), right of (1,3) = (pixel 2,
                     * pixel 0 of the right word); left of (1,3) = (0,2) and right of (0,2) = (1,3) of the same word */
                    unsigned leftA = __byte_perm(lB, vB, 0x5432), rightB = __byte_perm(vA, rA, 0x5432);
                    unsigned leftB = vA, rightA = vB;
                    /* cell borders: the first column of a cell has no left neighbours, the last one no right neighbours */
                    if (x0 - cg0 * wC == 0) leftA &= 0xffff0000u;
                    if (jb <= 4) {
                        if (jb == 1) { leftB &= 0xffff0000u; rightA &= 0xffff0000u; }
                        else if (jb == 2) { leftA &= 0x0000ffffu; rightB &= 0xffff0000u; }
                        else if (jb == 3) { leftB &= 0x0000ffffu; rightA &= 0x0000ffffu; }
                        else rightB &= 0x0000ffffu;
                    }
                    const unsigned nA = __vimax3_u16x2(leftA, rightA, udA), nB = __vimax3_u16x2(leftB, rightB, udB);
                    /* score > neighbours per lane: bit 15 of (score + 0x7fff - max) (all values <= 255: no carries between lanes) */
                    const unsigned gA = (cA + 0x7fff7fffu - nA) & 0x80008000u, gB = (cB + 0x7fff7fffu - nB) & 0x80008000u;
                    lm = ((gA >> 15) & 1u) | ((gB >> 14) & 2u) | ((gA >> 29) & 4u) | ((gB >> 28) & 8u);
                    /* only the cells this pass works on */
                    const unsigned inLo = (retry >> cg0) & 1u, inHi = (retry >> min(cg0 + 1, FAST_GROUP - 1)) & 1u;
                    const unsigned loMask = jb >= 4 ? 0xfu : ((1u << jb) - 1u);
                    lm &= (inLo ? loMask : 0u) | (inHi ? (0xfu & ~loMask) : 0u);
                }
            }
            /* compaction as in phase 0: warp scan of the per-thread counts, one shared atomic per warp */
            const int n = __popc(lm);
            int incl = n;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int up = __shfl_up_sync(0xffffffffu, incl, o);
                if (lane >= o) incl += up;
            }
            const int total = __shfl_sync(0xffffffffu, incl, 31);
            if (total == 0) continue;
            int base = 0;
            if (lane == 31) base = smem_add(&nout, total);
            int pos = __shfl_sync(0xffffffffu, base, 31) + incl - n;
            if (lm) {
                /* window coordinates: cell-relative (x+3, y+3) + (j*wCell, i*hCell)   (:820-825); score byte + th - 1 */
                const uint32_t rec = recBase + (uint32_t)x0 + ((uint32_t)y << 12) + ((uint32_t)(th - 1) << 24);
#pragma unroll
                for (int k = 0; k < 4; k++)
                    if ((lm >> k) & 1u) stage[pos++] = rec + (uint32_t)k + (((cw >> (8 * k)) & 0xffu) << 24);
                const unsigned lo = lm & (jb >= 4 ? 0xfu : ((1u << jb) - 1u));
                if (lo) smem_add(&cellCnt[cg0], __popc(lo));
                if (lm & ~lo) smem_add(&cellCnt[cg0 + 1], __popc(lm & ~lo));
            }
        }
        __syncthreads();
        /* warp 0 appends the CTA's records to the (frame, level) candidate pool: one global atomic per CTA and pass; the
         * other warps go on (the staging area is the dead tile, which warp 0 itself reloads at the top of a retry pass) */
        if (tid < 32) {
            const int n = nout;
            if (n) {
                int b0 = 0;
                if (lane == 0) {
                    b0 = atomicAdd(candCount + frame * g.nlevels + l, n);
                    if (b0 + n > L.candCap) atomicOr(status, VIORB_DEV_CAND_OVERFLOW);
                }
                b0 = __shfl_sync(0xffffffffu, b0, 0);
                for (int i = lane; i < n; i += 32)
                    if (b0 + i < L.candCap) out[b0 + i] = stage[i];
            }
        }
        /* cells that found nothing are run again at minThFAST */
        unsigned again = 0;
        for (int c = 0; c < ncell; c++) again |= (cellCnt[c] == 0 ? 1u : 0u) << c;
        again &= retry;
        if (pass == 1 || again == 0 || g.minTh >= g.iniTh) break;
        retry = again;
        __syncthreads();                 /* the lists and the score rows are rewritten by the next pass */
    }
}

/* ------------------------------------------------------------------------------------------------
 * DistributeOctTree (:539-763) as a depth-synchronous quadtree refinement, one CTA per (frame, level).
 *
 * The reference's std::list is always ordered by node creation time, newest first (children are
 * push_front'ed, roots push_back'ed), so the list can be kept as an array in that order and rebuilt
 * per pass:   new list = [children, newest first] ++ [surviving old nodes in old order].
 * Normal passes expand every node holding > 1 key in list order; once  size + 3*nToExpand > N  the
 * reference expands in (size desc, pointer desc) order and stops as soon as size >= N; the pointer
 * is replaced by the creation order (newest first == list order), the documented tie convention.
 * Keys never move: each key only tracks the list position of its node (nodeOf).
 * ---------------------------------------------------------------------------------------------- */
#define OCT_THREADS 256           /* upper bound; batches launch 128 threads (more CTAs per SM overlap the barrier waits), single
                                     frames 256 (shorter critical path of the level-0 CTA) */

struct OctSmem {
    short4* rectA; short4* rectB;   /* x0,y0,x1,y1 */
    int* cntA; int* cntB;
    int* cc;        /* [4*NC] child key counts */
    int* erank;     /* [NC] expansion rank or -1 */
    int* aux;       /* [NC] scan / new position */
    unsigned* srt;  /* [pow2(NC)] sort keys */
    unsigned short* childPos;   /* [4*NC] */
};

__device__ __forceinline__ int next_pow2(int v) {
    int p = 1;
    while (p < v) p <<= 1;
    return p;
}

/* exclusive scan of data[0..n) in place; returns the total.  All threads of the block must call. */
__device__ int block_exscan(int* data, int n, int* warpTmp /* 32 ints shared */) {
    const int tid = threadIdx.x, nt = blockDim.x;
    const int per = (n + nt - 1) / nt;
    const int beg = min(tid * per, n), end = min(beg + per, n);
    int sum = 0;
    for (int i = beg; i < end; i++) sum += data[i];
    /* block scan of per-thread sums */
    int v = sum;
    const int lane = tid & 31, wid = tid >> 5;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        int t = __shfl_up_sync(0xffffffffu, v, o);
        if (lane >= o) v += t;
    }
    if (lane == 31) warpTmp[wid] = v;
    __syncthreads();
    if (wid == 0) {
        int w = lane < (nt >> 5) ? warpTmp[lane] : 0;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            int t = __shfl_up_sync(0xffffffffu, w, o);
            if (lane >= o) w += t;
        }
        warpTmp[lane] = w;    /* inclusive over warps */
    }
    __syncthreads();
    const int total = warpTmp[(nt >> 5) - 1];
    int run = v - sum + (wid > 0 ? warpTmp[wid - 1] : 0);
    __syncthreads();
    for (int i = beg; i < end; i++) {
        int t = data[i];
        data[i] = run;
        run += t;
    }
    __syncthreads();
    return total;
}

__device__ void block_bitonic_sort(unsigned* a, int n /* power of two */) {
    for (int k = 2; k <= n; k <<= 1) {
        for (int j = k >> 1; j > 0; j >>= 1) {
            for (int i = threadIdx.x; i < n; i += blockDim.x) {
                const int ixj = i ^ j;
                if (ixj > i) {
                    const unsigned x = a[i], y = a[ixj];
                    const bool up = (i & k) == 0;
                    if ((x > y) == up) { a[i] = y; a[ixj] = x; }
                }
            }
            __syncthreads();
        }
    }
}

__device__ __forceinline__ int quadrant(int kx, int ky, short4 r) {
    /* DivideNode :481-526: halfX = ceil((UR.x-UL.x)/2.f); key goes left iff pt.x < UL.x+halfX */
    const int mx = r.x + ((r.z - r.x + 1) >> 1);
    const int my = r.y + ((r.w - r.y + 1) >> 1);
    return (kx < mx ? 0 : 1) + (ky < my ? 0 : 2);
}

__global__ void __launch_bounds__(OCT_THREADS) octree_kernel(const __grid_constant__ FrameGeom g,
                                                             const uint32_t* __restrict__ cand,
                                                             const int* __restrict__ candCount,
                                                             uint16_t* __restrict__ nodeOfAll,
                                                             uint32_t* __restrict__ sel, int* __restrict__ selCount,
                                                             int* __restrict__ status, int NC) {
    extern __shared__ __align__(16) unsigned char smemRaw[];
    __shared__ int warpTmp[32];
    __shared__ int sh[8];
    const int level = blockIdx.x, frame = blockIdx.y;
    const LevelGeom& L = g.lv[level];
    const int tid = threadIdx.x, nt = blockDim.x;
    pdl_trigger();
    pdl_wait();                               /* the candidate pools are complete from here on */
    const int n = min(candCount[frame * g.nlevels + level], L.candCap);
    const uint32_t* keys = cand + (size_t)frame * g.candPerFrame + L.candBase;
    uint16_t* nodeOf = nodeOfAll + (size_t)frame * g.candPerFrame + L.candBase;
    uint32_t* out = sel + (size_t)frame * g.selPerFrame + L.selBase;
    const int N = L.quota;
    if (n == 0) {
        if (tid == 0) selCount[frame * g.nlevels + level] = 0;
        return;
    }
    /* carve shared memory */
    OctSmem S;
    {
        unsigned char* p = smemRaw;
        S.rectA = (short4*)p; p += sizeof(short4) * NC;
        S.rectB = (short4*)p; p += sizeof(short4) * NC;
        S.cntA = (int*)p; p += sizeof(int) * NC;
        S.cntB = (int*)p; p += sizeof(int) * NC;
        S.cc = (int*)p; p += sizeof(int) * 4 * NC;
        S.erank = (int*)p; p += sizeof(int) * NC;
        S.aux = (int*)p; p += sizeof(int) * NC;
        S.srt = (unsigned*)p; p += sizeof(unsigned) * next_pow2(NC);
        S.childPos = (unsigned short*)p;
    }
    short4* rect = S.rectA; short4* rectN = S.rectB;
    int* cnt = S.cntA; int* cntN = S.cntB;

    /* ---- roots (:543-585) ---- */
    const int W = (L.w - VIORB_FAST_BORDER) - VIORB_FAST_BORDER, H = (L.h - VIORB_FAST_BORDER) - VIORB_FAST_BORDER;
    const int nIni = L.nIni;
    const float hX = __fdiv_rn((float)W, (float)nIni);
    for (int i = tid; i < nIni; i += nt) {
        rect[i] = make_short4((short)(int)__fmul_rn(hX, (float)i), 0, (short)(int)__fmul_rn(hX, (float)(i + 1)), (short)H);
        cnt[i] = 0;
    }
    __syncthreads();
    for (int k = tid; k < n; k += nt) {
        const int kx = keys[k] & 0xfff;
        int r = (int)__fdiv_rn((float)kx, hX);
        r = min(r, nIni - 1);
        nodeOf[k] = (uint16_t)r;
        atomicAdd(&cnt[r], 1);
    }
    __syncthreads();
    /* drop empty roots, keep order */
    for (int i = tid; i < nIni; i += nt) S.aux[i] = cnt[i] > 0 ? 1 : 0;
    __syncthreads();
    int Sn = block_exscan(S.aux, nIni, warpTmp);
    for (int i = tid; i < nIni; i += nt)
        if (cnt[i] > 0) { rectN[S.aux[i]] = rect[i]; cntN[S.aux[i]] = cnt[i]; }
    __syncthreads();
    for (int k = tid; k < n; k += nt) nodeOf[k] = (uint16_t)S.aux[nodeOf[k]];
    { short4* t = rect; rect = rectN; rectN = t; int* u = cnt; cnt = cntN; cntN = u; }
    __syncthreads();

    bool careful = false;
    for (;;) {
        const int prev = Sn;
        /* number of expandable nodes */
        int myMulti = 0;
        for (int i = tid; i < Sn; i += nt) myMulti += cnt[i] > 1;
        const int anyMulti = __syncthreads_or(myMulti);
        if (!anyMulti) break;                       /* size == prevSize -> finish (:669) */
        /* child key counts of every expandable node */
        for (int i = tid; i < 4 * Sn; i += nt) S.cc[i] = 0;
        __syncthreads();
        for (int k = tid; k < n; k += nt) {
            const int i = nodeOf[k];
            if (cnt[i] > 1) {
                const uint32_t key = keys[k];
                atomicAdd(&S.cc[4 * i + quadrant(key & 0xfff, (key >> 12) & 0xfff, rect[i])], 1);
            }
        }
        __syncthreads();
        int E;     /* number of nodes expanded this pass */
        if (!careful) {
            /* list order (:594-665) */
            for (int i = tid; i < Sn; i += nt) S.aux[i] = cnt[i] > 1 ? 1 : 0;
            __syncthreads();
            E = block_exscan(S.aux, Sn, warpTmp);
            for (int i = tid; i < Sn; i += nt) S.erank[i] = cnt[i] > 1 ? S.aux[i] : -1;
            __syncthreads();
        } else {
            /* (size desc, newest first) order with early stop at N (:673-738) */
            for (int i = tid; i < Sn; i += nt) S.aux[i] = cnt[i] > 1 ? 1 : 0;
            __syncthreads();
            const int nM = block_exscan(S.aux, Sn, warpTmp);
            const int P2 = next_pow2(nM);
            for (int i = tid; i < P2; i += nt) S.srt[i] = 0xffffffffu;
            __syncthreads();
            for (int i = tid; i < Sn; i += nt)
                if (cnt[i] > 1) S.srt[S.aux[i]] = ((unsigned)(0xfffff - cnt[i]) << 12) | (unsigned)i;   /* cnt < 2^20, i < 4096 */
            __syncthreads();
            block_bitonic_sort(S.srt, P2);
            /* gains in sorted order */
            for (int r = tid; r < nM; r += nt) {
                const int i = S.srt[r] & 0xfff;
                const int ne = (S.cc[4 * i] > 0) + (S.cc[4 * i + 1] > 0) + (S.cc[4 * i + 2] > 0) + (S.cc[4 * i + 3] > 0);
                S.aux[r] = ne - 1;
            }
            for (int i = tid; i < Sn; i += nt) S.erank[i] = -1;
            if (tid == 0) sh[0] = nM;
            __syncthreads();
            block_exscan(S.aux, nM, warpTmp);       /* aux[r] = gain before r */
            for (int r = tid; r < nM; r += nt) {
                const int i = S.srt[r] & 0xfff;
                const int ne = (S.cc[4 * i] > 0) + (S.cc[4 * i + 1] > 0) + (S.cc[4 * i + 2] > 0) + (S.cc[4 * i + 3] > 0);
                const int before = prev + S.aux[r], after = before + ne - 1;
                if (before < N && after >= N) sh[0] = r + 1;      /* first rank reaching N: expand r, then break */
            }
            __syncthreads();
            E = sh[0];
            for (int r = tid; r < E; r += nt) S.erank[S.srt[r] & 0xfff] = r;
            __syncthreads();
        }
        /* creation ranks of the children: expansion order, then n1..n4 */
        for (int i = tid; i < Sn; i += nt) {
            const int e = S.erank[i];
            if (e >= 0) S.srt[e] = (unsigned)i;       /* node expanded at rank e */
        }
        __syncthreads();
        for (int e = tid; e < E; e += nt) {
            const int i = S.srt[e];
            S.aux[e] = (S.cc[4 * i] > 0) + (S.cc[4 * i + 1] > 0) + (S.cc[4 * i + 2] > 0) + (S.cc[4 * i + 3] > 0);
        }
        __syncthreads();
        const int C = block_exscan(S.aux, E, warpTmp);      /* aux[e] = children created before rank e */
        const int SnNew = C + (Sn - E);
        if (SnNew > NC || SnNew > L.selCap) {
            if (tid == 0) { atomicOr(status, VIORB_DEV_NODE_OVERFLOW); selCount[frame * g.nlevels + level] = 0; }
            return;
        }
        /* children */
        for (int e = tid; e < E; e += nt) {
            const int i = S.srt[e];
            const short4 r = rect[i];
            const int mx = r.x + ((r.z - r.x + 1) >> 1), my = r.y + ((r.w - r.y + 1) >> 1);
            int j = S.aux[e];
#pragma unroll
            for (int q = 0; q < 4; q++) {
                const int c = S.cc[4 * i + q];
                if (c > 0) {
                    const int pos = C - 1 - j;
                    j++;
                    short4 cr;
                    cr.x = (q & 1) ? (short)mx : r.x;
                    cr.z = (q & 1) ? r.z : (short)mx;
                    cr.y = (q & 2) ? (short)my : r.y;
                    cr.w = (q & 2) ? r.w : (short)my;
                    rectN[pos] = cr;
                    cntN[pos] = c;
                    S.childPos[4 * i + q] = (unsigned short)pos;
                }
            }
        }
        __syncthreads();
        /* survivors keep their relative order behind the new children */
        for (int i = tid; i < Sn; i += nt) S.aux[i] = S.erank[i] < 0 ? 1 : 0;
        __syncthreads();
        block_exscan(S.aux, Sn, warpTmp);
        for (int i = tid; i < Sn; i += nt)
            if (S.erank[i] < 0) {
                const int pos = C + S.aux[i];
                rectN[pos] = rect[i];
                cntN[pos] = cnt[i];
                S.aux[i] = pos;
            }
        __syncthreads();
        for (int k = tid; k < n; k += nt) {
            const int i = nodeOf[k];
            if (S.erank[i] >= 0) {
                const uint32_t key = keys[k];
                nodeOf[k] = S.childPos[4 * i + quadrant(key & 0xfff, (key >> 12) & 0xfff, rect[i])];
            } else {
                nodeOf[k] = (uint16_t)S.aux[i];
            }
        }
        { short4* t = rect; rect = rectN; rectN = t; int* u = cnt; cnt = cntN; cntN = u; }
        Sn = SnNew;
        __syncthreads();
        if (Sn >= N || Sn == prev) break;           /* (:669, :734) */
        if (!careful) {
            int m = 0;
            for (int i = tid; i < Sn; i += nt) m += cnt[i] > 1;
            /* nToExpand = children created this pass that hold > 1 key == all multi-key nodes now */
            for (int o = 16; o > 0; o >>= 1) m += __shfl_xor_sync(0xffffffffu, m, o);
            if ((tid & 31) == 0) warpTmp[tid >> 5] = m;
            __syncthreads();
            int nToExpand = 0;
            for (int w = 0; w < (nt >> 5); w++) nToExpand += warpTmp[w];
            __syncthreads();
            if (Sn + 3 * nToExpand > N) careful = true;     /* (:673) */
        }
    }

    /* ---- retain the best key per node (:741-760): max response, first in candidate order on ties.
     * candidate order = (cell row, cell col, y, x)  (:789-829) ---- */
    int* best = S.cc;           /* reuse */
    unsigned* tie = (unsigned*)S.erank;
    for (int i = tid; i < Sn; i += nt) { best[i] = 0; tie[i] = 0xffffffffu; }
    __syncthreads();
    for (int k = tid; k < n; k += nt) atomicMax(&best[nodeOf[k]], (int)(keys[k] >> 24));
    __syncthreads();
    for (int k = tid; k < n; k += nt) {
        const uint32_t key = keys[k];
        const int i = nodeOf[k];
        if ((int)(key >> 24) == best[i]) {
            const int x = (key & 0xfff) - 3, y = ((key >> 12) & 0xfff) - 3;
            const int cy = y / L.hCell, cx = x / L.wCell;
            const unsigned ord = ((unsigned)cy << 20) | ((unsigned)cx << 12) | ((unsigned)(y - cy * L.hCell) << 6) |
                                 (unsigned)(x - cx * L.wCell);
            atomicMin(&tie[i], ord);
        }
    }
    __syncthreads();
    for (int k = tid; k < n; k += nt) {
        const uint32_t key = keys[k];
        const int i = nodeOf[k];
        if ((int)(key >> 24) == best[i]) {
            const int x = (key & 0xfff) - 3, y = ((key >> 12) & 0xfff) - 3;
            const int cy = y / L.hCell, cx = x / L.wCell;
            const unsigned ord = ((unsigned)cy << 20) | ((unsigned)cx << 12) | ((unsigned)(y - cy * L.hCell) << 6) |
                                 (unsigned)(x - cx * L.wCell);
            if (ord == tie[i]) {
                /* level coordinates: += minBorder (:840-841) */
                const uint32_t X = (key & 0xfff) + VIORB_FAST_BORDER, Y = ((key >> 12) & 0xfff) + VIORB_FAST_BORDER;
                out[i] = X | (Y << 12) | (key & 0xff000000u);
            }
        }
    }
    if (tid == 0) selCount[frame * g.nlevels + level] = Sn;
}

/* ------------------------------------------------------------------------------------------------
 * IC_Angle (:77-104) + GaussianBlur 7x7 sigma 2 (:1086) + computeOrbDescriptor (:108-147), fused:
 * one warp per selected keypoint.  The 43x43 neighbourhood (radius 18 pattern reach + 3 blur taps)
 * is staged in shared memory from the padded pyramid (whose REFLECT_101 border equals the blur's own
 * border rule), blurred there with the OpenCV >= 3.4 fixed-point taps [18 34 48 56 48 34 18]/256,
 * and sampled at the 512 steered pattern points.  The blurred level image never exists in HBM.
 * ---------------------------------------------------------------------------------------------- */
#define DESC_WARPS 4
#define DESC_KPW 8            /* most keypoints per warp (batch launches), fused kernel */
#define DESC2_KPW 16          /* the same for describe_blurred_kernel (<= 32: lane j keeps keypoint j) */
#define PR 21                 /* patch radius */
#define PROWS 43              /* patch rows / columns */
#define PWORDS 13             /* patch row stride in 32-bit words (odd: conflict-free row pairs; 12 words used) */
#define BW 37                 /* blurred width (radius 18) */
#define HT_WORDS 25           /* words per column of the transposed horizontal-pass buffer (>= 23; = 1 mod 8 so the
                                 transposed stores of 4-column quads fall into distinct banks) */
#define VT_STRIDE 44          /* bytes per column of the transposed blurred patch (37 rows + pad; 11 words: odd) */

__device__ __align__(16) const float d_pattern[1024] = VIORB_ORB_PATTERN_INIT;   /* as floats: x*b + y*a needs no I2F */
__constant__ int c_umax[16] = {15, 15, 15, 15, 14, 14, 14, 13, 13, 12, 11, 10, 9, 8, 6, 3};

/* shifts that give 0 for amounts >= 32 (PTX shl/shr clamp; C++ << is undefined there) */
__device__ __forceinline__ unsigned shl_clamp(unsigned x, int n) {
    unsigned r;
    asm("shl.b32 %0, %1, %2;" : "=r"(r) : "r"(x), "r"(n));
    return r;
}
__device__ __forceinline__ unsigned shr_clamp(unsigned x, int n) {
    unsigned r;
    asm("shr.u32 %0, %1, %2;" : "=r"(r) : "r"(x), "r"(n));
    return r;
}
/* four unsigned bytes x four signed bytes + c */
__device__ __forceinline__ int dp4a_u8s8(unsigned a, int b, int c) {
    int r;
    asm("dp4a.u32.s32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(c));
    return r;
}

/* cv::fastAtan2 (degrees), OpenCV core mathfuncs_core atan_f32 polynomial, no FMA */
__device__ __forceinline__ float fast_atan2_deg(float y, float x) {
    const float scale = (float)(180.0 / 3.14159265358979323846);
    const float p1 = __fmul_rn(0.9997878412794807f, scale);
    const float p3 = __fmul_rn(-0.3258083974640975f, scale);
    const float p5 = __fmul_rn(0.1555786518463281f, scale);
    const float p7 = __fmul_rn(-0.04432655554792128f, scale);
    const float eps = (float)2.2204460492503131e-16;
    const float ax = fabsf(x), ay = fabsf(y);
    float a, c, c2;
    if (ax >= ay) {
        c = __fdiv_rn(ay, __fadd_rn(ax, eps));
        c2 = __fmul_rn(c, c);
        a = __fmul_rn(__fadd_rn(__fmul_rn(__fadd_rn(__fmul_rn(__fadd_rn(__fmul_rn(p7, c2), p5), c2), p3), c2), p1), c);
    } else {
        c = __fdiv_rn(ax, __fadd_rn(ay, eps));
        c2 = __fmul_rn(c, c);
        a = __fsub_rn(90.f, __fmul_rn(__fadd_rn(__fmul_rn(__fadd_rn(__fmul_rn(__fadd_rn(__fmul_rn(p7, c2), p5), c2), p3), c2), p1), c));
    }
    if (x < 0) a = __fsub_rn(180.f, a);
    if (y < 0) a = __fsub_rn(360.f, a);
    return a;
}

/* Horizontal blur pass of orient_describe_kernel for a patch whose rows sit SH bytes into their first word.
 * The 7-tap window of output column 4j+i starts at byte i+SH of raw word j: instead of shifting the pixels into place
 * the tap words are shifted (compile-time constants per SH), two or three IDP.4A per output and no PRMT. */
/* GV selects the 8-bit Gaussian kernel of cv::GaussianBlur(7x7, sigma 2): 0 = OpenCV >= 3.4 (fixed point
 * [18,34,48,56,48,34,18]/256), 1 = OpenCV 2.4, the version the reference pins (round(getGaussianKernel * 256) =
 * [18,34,49,55,49,34,18], sum 257; the result saturates at 255).  viorb_extractor_set_gaussian. */
__host__ __device__ constexpr int blur_tap(int gv, int p) {
    const int k[2][7] = {{18, 34, 48, 56, 48, 34, 18}, {18, 34, 49, 55, 49, 34, 18}};
    return k[gv][p];
}
__host__ __device__ constexpr unsigned blur_tap_word(int gv, int o, int w) {
    unsigned r = 0;
    for (int b = 0; b < 4; b++) {
        const int p = 4 * w + b - o;
        if (p >= 0 && p < 7) r |= (unsigned)blur_tap(gv, p) << (8 * b);
    }
    return r;
}
template <int GV, int O>
__device__ __forceinline__ unsigned blur_hsum(unsigned r0, unsigned r1, unsigned r2, unsigned r3) {
    unsigned acc = 0;
    if constexpr (blur_tap_word(GV, O, 3) != 0) acc = __dp4a(r3, blur_tap_word(GV, O, 3), acc);
    if constexpr (blur_tap_word(GV, O, 2) != 0) acc = __dp4a(r2, blur_tap_word(GV, O, 2), acc);
    if constexpr (blur_tap_word(GV, O, 1) != 0) acc = __dp4a(r1, blur_tap_word(GV, O, 1), acc);
    if constexpr (blur_tap_word(GV, O, 0) != 0) acc = __dp4a(r0, blur_tap_word(GV, O, 0), acc);
    return acc;
}
/* Shared-memory banks (the L1 data pipe is this kernel's busiest unit): an instruction covers columns 0..31
 * (8 quads) of four row pairs whose patch rows start 8 banks apart -- {b, b+4, b+8, b+12}, and {16, 20, 18},
 * {17, 21, 19} for the rest (row pair rp starts at bank 26*rp mod 32) -- so the 32 loaded words fall into 32
 * banks; row pair rp is stored at pos(rp) = 6*(rp&3) + (rp>>2), which makes the four row pairs of an
 * instruction neighbours in Hw and the transposed stores conflict-free as well.  Columns 32..36 are a seventh
 * instruction with one row pair per lane. */
template <int GV, int SH>
__device__ __forceinline__ void blur_h_pass(const unsigned* __restrict__ P, unsigned* __restrict__ Hw, int lane) {
    const int q = lane >> 3, j = lane & 7;
#pragma unroll 1
    for (int it = 0; it < 6; it++) {
        int rp;
        if (it < 4) rp = it + 4 * q;
        else rp = (q == 3) ? -1 : 12 + it + (q == 1 ? 4 : (q == 2 ? 2 : 0));
        if (rp < 0) continue;
        const unsigned* p0 = &P[(2 * rp) * PWORDS + j];
        const unsigned* p1 = rp < 21 ? p0 + PWORDS : p0;     /* row 43 does not exist: its sums are never used */
        const unsigned a0 = p0[0], a1 = p0[1], a2 = p0[2], a3 = p0[3], b0 = p1[0], b1 = p1[1], b2 = p1[2], b3 = p1[3];
        unsigned* dst = &Hw[(4 * j) * HT_WORDS + 6 * (rp & 3) + (rp >> 2)];
        dst[0 * HT_WORDS] = __byte_perm(blur_hsum<GV, SH + 0>(a0, a1, a2, a3), blur_hsum<GV, SH + 0>(b0, b1, b2, b3), 0x5410);
        dst[1 * HT_WORDS] = __byte_perm(blur_hsum<GV, SH + 1>(a0, a1, a2, a3), blur_hsum<GV, SH + 1>(b0, b1, b2, b3), 0x5410);
        dst[2 * HT_WORDS] = __byte_perm(blur_hsum<GV, SH + 2>(a0, a1, a2, a3), blur_hsum<GV, SH + 2>(b0, b1, b2, b3), 0x5410);
        dst[3 * HT_WORDS] = __byte_perm(blur_hsum<GV, SH + 3>(a0, a1, a2, a3), blur_hsum<GV, SH + 3>(b0, b1, b2, b3), 0x5410);
    }
    if (lane < 22) {                                           /* columns 32..36 of row pair `lane` */
        const int rp = lane;
        const unsigned* p0 = &P[(2 * rp) * PWORDS + 8];
        const unsigned* p1 = rp < 21 ? p0 + PWORDS : p0;
        const unsigned a0 = p0[0], a1 = p0[1], a2 = p0[2], a3 = p0[3], b0 = p1[0], b1 = p1[1], b2 = p1[2], b3 = p1[3];
        unsigned* dst = &Hw[32 * HT_WORDS + 6 * (rp & 3) + (rp >> 2)];
        dst[0 * HT_WORDS] = __byte_perm(blur_hsum<GV, SH + 0>(a0, a1, a2, a3), blur_hsum<GV, SH + 0>(b0, b1, b2, b3), 0x5410);
        dst[1 * HT_WORDS] = __byte_perm(blur_hsum<GV, SH + 1>(a0, a1, a2, a3), blur_hsum<GV, SH + 1>(b0, b1, b2, b3), 0x5410);
        dst[2 * HT_WORDS] = __byte_perm(blur_hsum<GV, SH + 2>(a0, a1, a2, a3), blur_hsum<GV, SH + 2>(b0, b1, b2, b3), 0x5410);
        dst[3 * HT_WORDS] = __byte_perm(blur_hsum<GV, SH + 3>(a0, a1, a2, a3), blur_hsum<GV, SH + 3>(b0, b1, b2, b3), 0x5410);
        dst[4 * HT_WORDS] = __byte_perm(blur_hsum<GV, SH + 4>(a0, a1, a2, a3), blur_hsum<GV, SH + 4>(b0, b1, b2, b3), 0x5410);
    }
}

/* glibc flt-32 sincosf (double polynomial, generic reduction), every op rounded separately.
 * Valid for |y| < 120; the extractor feeds angles in [0, 2*pi].   */
__device__ __forceinline__ void sincosf_glibc(float y, float* sinp, float* cosp) {
    const double hpi_inv = 0x1.45F306DC9C883p+23, hpi = 0x1.921FB54442D18p0;
    const double C0 = 0x1p0, C1 = -0x1.ffffffd0c621cp-2, C2 = 0x1.55553e1068f19p-5, C3 = -0x1.6c087e89a359dp-10,
                 C4 = 0x1.99343027bf8c3p-16;
    const double S1 = -0x1.555545995a603p-3, S2 = 0x1.1107605230bc4p-7, S3 = -0x1.994eb3774cf24p-13;
    double x = (double)y;
    const unsigned top = (__float_as_uint(y) >> 20) & 0x7ff;
    int n = 0;
    double sgnc = 1.0;     /* cosine polynomial sign (table entry 1 negates it) */
    double xs;
    double x2;
    if (top < 0x3f4u) {            /* abstop12(pi/4 = 0x1.921FB6p-1f) */
        if (top < 0x398u) {        /* abstop12(0x1p-12f) */
            *sinp = y;
            *cosp = 1.0f;
            return;
        }
        x2 = __dmul_rn(x, x);
        xs = x;
    } else {
        const double r = __dmul_rn(x, hpi_inv);
        n = ((int)r + 0x800000) >> 24;
        x = __dsub_rn(x, __dmul_rn((double)n, hpi));
        const double s = (n & 3) == 1 || (n & 3) == 2 ? -1.0 : 1.0;
        if (n & 2) sgnc = -1.0;
        x2 = __dmul_rn(x, x);
        xs = __dmul_rn(x, s);
    }
    const double c0 = C0 * sgnc, c1 = C1 * sgnc, c2k = C2 * sgnc, c3 = C3 * sgnc, c4 = C4 * sgnc;
    const double x4 = __dmul_rn(x2, x2);
    const double x3 = __dmul_rn(x2, xs);
    const double cc2 = __dadd_rn(c3, __dmul_rn(x2, c4));
    const double s1 = __dadd_rn(S2, __dmul_rn(x2, S3));
    const double cc1 = __dadd_rn(c0, __dmul_rn(x2, c1));
    const double x5 = __dmul_rn(x3, x2);
    const double x6 = __dmul_rn(x4, x2);
    const double s = __dadd_rn(xs, __dmul_rn(x3, S1));
    const double c = __dadd_rn(cc1, __dmul_rn(x4, c2k));
    const float sv = (float)__dadd_rn(s, __dmul_rn(x5, s1));
    const float cv = (float)__dadd_rn(c, __dmul_rn(x6, cc2));
    if (n & 1) { *cosp = sv; *sinp = cv; }
    else { *sinp = sv; *cosp = cv; }
}

template <int GV>
__global__ void __launch_bounds__(DESC_WARPS * 32, 6) orient_describe_kernel(const __grid_constant__ FrameGeom g,
                                                                          const uint8_t* __restrict__ pyr,
                                                                          const uint32_t* __restrict__ sel,
                                                                          const int* __restrict__ selCount,
                                                                          viorb_keypoint* __restrict__ kps,
                                                                          uint8_t* __restrict__ desc, int cap,
                                                                          int32_t* __restrict__ counts,
                                                                          int* __restrict__ status, int kpw) {
    /* per warp: patch (43 x 12 words; later reused for the blurred 37x37 bytes) + transposed H-pass buffer */
    __shared__ __align__(16) unsigned patchW[DESC_WARPS][PROWS * PWORDS];
    __shared__ __align__(16) unsigned hbW[DESC_WARPS][BW * HT_WORDS];
    const int frame = blockIdx.y;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    pdl_trigger();
    /* this lane's 8 binary tests (16 sampling points, 32 floats), kept in registers for all keypoints of the warp:
     * read per keypoint they were 4 KB through L1 -- twice the patch */
    float4 pat[8];
#pragma unroll
    for (int k = 0; k < 8; k++) pat[k] = __ldg(reinterpret_cast<const float4*>(d_pattern) + 8 * lane + k);
    pdl_wait();                               /* the selected keypoints (and the pyramid) are complete from here on */
    /* lane l holds [lo, hi) of level l in the concatenated per-level lists (:1076-1103) */
    int myLo, myHi, total;
    {
        const int c = lane < g.nlevels ? selCount[frame * g.nlevels + lane] : 0;
        int incl = c;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int t = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += t;
        }
        myHi = incl;
        myLo = incl - c;
        total = __shfl_sync(0xffffffffu, incl, 31);
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        counts[frame] = min(total, cap);
        if (total > cap) atomicOr(status, VIORB_DEV_OUT_OVERFLOW);
    }
    unsigned* P = patchW[warp];
    unsigned* Hw = hbW[warp];
    for (int it = 0; it < kpw; it++) {
    const int slot = (it * gridDim.x + blockIdx.x) * DESC_WARPS + warp;
    if (slot >= total || slot >= cap) break;
    const int level = __ffs(__ballot_sync(0xffffffffu, slot >= myLo && slot < myHi)) - 1;
    const int idx = slot - __shfl_sync(0xffffffffu, myLo, level);
    const LevelGeom& L = g.lv[level];
    const uint32_t key = sel[(size_t)frame * g.selPerFrame + L.selBase + idx];
    const int kx = key & 0xfff, ky = (key >> 12) & 0xfff, score = key >> 24;
    const uint8_t* roi = pyr + (size_t)frame * g.pyrFrameBytes + L.pyrOff + (size_t)VIORB_EDGE * L.step + VIORB_ROI_X0;
    /* stage the 43x43 neighbourhood with cp.async (global -> shared without a register round trip): patch row r =
     * the 13 aligned words that hold level pixels (kx-21 .. kx+21, ky-21+r), i.e. patch byte (r, c) sits at byte
     * sh + c of the row, sh = (kx-21) & 3.  Half a warp copies one row; the consumers below realign by sh. */
    const int sh = (kx - PR) & 3;
    {
        const int gx0 = kx - PR;
        const int w = lane & 15, rsub = lane >> 4;
        const unsigned* p = reinterpret_cast<const unsigned*>(roi + (ptrdiff_t)(ky - PR + rsub) * L.step + (gx0 - sh)) + w;
        const ptrdiff_t twoRows = (ptrdiff_t)(L.step >> 1);           /* 2 * step bytes in words */
        unsigned dst = (unsigned)__cvta_generic_to_shared(P + rsub * PWORDS + w);
        const bool okW = w < 13;
#pragma unroll
        for (int k = 0; k < 22; k++) {
            if (okW && (k < 21 || rsub == 0))                           /* row 2k + rsub < 43 */
                asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(dst), "l"(p) : "memory");
            p += twoRows;
            dst += 2 * PWORDS * 4;
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
    }
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    __syncwarp();
    /* IC_Angle: lane v+15 sums row v of the circular patch */
    int m10 = 0, m01 = 0;
    if (lane < 31) {
        const int v = lane - 15;
        const int d = c_umax[v < 0 ? -v : v];
        /* the row as nine aligned words (bytes 4..39 of the patch row: u = -17..18); bytes outside the row's
         * half-width d are masked off, then one IDP.4A per word and moment (weights u as signed bytes) */
        const unsigned* rw = P + (PR + v) * PWORDS;
        unsigned sacc = 0;
        unsigned nxt = rw[1];
#pragma unroll
        for (int w = 1; w <= 9; w++) {
            const int u0 = 4 * w - PR;
            const unsigned cur = nxt;
            nxt = rw[w + 1];
            unsigned x = funnel_bytes(cur, nxt, sh);
            if (u0 + 3 < 0) x &= shl_clamp(0xffffffffu, 8 * max(-d - u0, 0));
            else if (u0 > 0) x &= shr_clamp(0xffffffffu, 8 * max(u0 + 3 - d, 0));
            const int wts = (u0 & 0xff) | (((u0 + 1) & 0xff) << 8) | (((u0 + 2) & 0xff) << 16) | (((u0 + 3) & 0xff) << 24);
            m10 = dp4a_u8s8(x, wts, m10);
            sacc = __dp4a(x, 0x01010101u, sacc);
        }
        m01 = v * (int)sacc;
    }
    m10 = __reduce_add_sync(0xffffffffu, m10);          /* REDUX.SUM: one instruction per moment instead of five shuffles */
    m01 = __reduce_add_sync(0xffffffffu, m01);
    const float angle = fast_atan2_deg((float)m01, (float)m10);
    /* horizontal 7-tap pass: a task = 2 rows x 4 columns.  The 16-bit sums (exact: the taps sum to 256) of
     * vertically adjacent rows are packed into one word and stored transposed (Hw[c][pos(rp)]) so the vertical
     * pass can use IDP.2A on row pairs.  One instantiation per byte offset of the patch rows (warp-uniform). */
    switch (sh) {
        case 0: blur_h_pass<GV, 0>(P, Hw, lane); break;
        case 1: blur_h_pass<GV, 1>(P, Hw, lane); break;
        case 2: blur_h_pass<GV, 2>(P, Hw, lane); break;
        default: blur_h_pass<GV, 3>(P, Hw, lane); break;
    }
    __syncwarp();
    /* vertical pass: a task = 8 output rows (segment seg) of one column, read as 7 words (row pairs 4*seg + k at
     * pos = seg + {0, 6, 12, 18, 1, 7, 13}); even and odd rows use the tap pairs shifted by one.
     * out = (sum + 32768) >> 16 (GaussianBlur's rounding); the blurred patch is stored transposed as well
     * (Vt[c][r], 11-word column stride) so a task writes two words.  Tasks are ordered segment-major: the lanes
     * of an instruction hold neighbouring columns, whose words are 25 (loads) and 11 (stores) banks apart. */
    uint8_t* Vt = reinterpret_cast<uint8_t*>(P);       /* the patch is dead now */
    {
        constexpr unsigned K0 = blur_tap(GV, 0), K1 = blur_tap(GV, 1), K2 = blur_tap(GV, 2), K3 = blur_tap(GV, 3);
        constexpr unsigned K4 = blur_tap(GV, 4), K5 = blur_tap(GV, 5), K6 = blur_tap(GV, 6);
        const unsigned E01 = K0 | (K1 << 8), E23 = K2 | (K3 << 8), E45 = K4 | (K5 << 8), E6 = K6;              /* even row */
        const unsigned O0 = K0 << 8, O12 = K1 | (K2 << 8), O34 = K3 | (K4 << 8), O56 = K5 | (K6 << 8);          /* odd row */
        for (int t = lane; t < BW * 5; t += 32) {
            const int seg = t / BW, c = t - seg * BW;        /* rows 8*seg .. 8*seg+7 */
            const unsigned* hcol = &Hw[c * HT_WORDS + seg];
            unsigned w[7];
#pragma unroll
            for (int k = 0; k < 7; k++) w[k] = hcol[6 * (k & 3) + (k >> 2)];
            unsigned pr[4];          /* rows 2k, 2k+1: the sums are < 2^24, so (sum >> 16) is byte 2 -- picked by PRMT */
#pragma unroll
            for (int k = 0; k < 4; k++) {
                unsigned ve = __dp2a_lo(w[k], E01, __dp2a_lo(w[k + 1], E23, __dp2a_lo(w[k + 2], E45, __dp2a_lo(w[k + 3], E6, 32768u))));
                unsigned vo = __dp2a_lo(w[k], O0, __dp2a_lo(w[k + 1], O12, __dp2a_lo(w[k + 2], O34, __dp2a_lo(w[k + 3], O56, 32768u))));
                if constexpr (GV == 1) {              /* taps sum to 257: a saturated neighbourhood reaches 257 -> saturate_cast<uchar> */
                    ve = min(ve, 0x00ffffffu);
                    vo = min(vo, 0x00ffffffu);
                }
                pr[k] = __byte_perm(ve, vo, 0x0062);
            }
            unsigned* dst = reinterpret_cast<unsigned*>(Vt + c * VT_STRIDE + 8 * seg);
            dst[0] = __byte_perm(pr[0], pr[1], 0x5410);
            dst[1] = __byte_perm(pr[2], pr[3], 0x5410);
        }
    }
    __syncwarp();
    /* steered BRIEF: lane i produces descriptor byte i (:123-144) */
    const float factorPI = (float)(3.14159265358979323846 / 180.f);
    float a, b;
    sincosf_glibc(__fmul_rn(angle, factorPI), &b, &a);
    const uint8_t* centre = &Vt[18 * VT_STRIDE + 18];
    unsigned val = 0;
#pragma unroll
    for (int k = 0; k < 8; k++) {
        const float x0 = pat[k].x, y0 = pat[k].y, x1 = pat[k].z, y1 = pat[k].w;
        const int r0 = __float2int_rn(__fadd_rn(__fmul_rn(x0, b), __fmul_rn(y0, a)));
        const int c0 = __float2int_rn(__fsub_rn(__fmul_rn(x0, a), __fmul_rn(y0, b)));
        const int r1 = __float2int_rn(__fadd_rn(__fmul_rn(x1, b), __fmul_rn(y1, a)));
        const int c1 = __float2int_rn(__fsub_rn(__fmul_rn(x1, a), __fmul_rn(y1, b)));
        const int t0 = centre[c0 * VT_STRIDE + r0], t1 = centre[c1 * VT_STRIDE + r1];
        val |= (unsigned)(t0 < t1) << k;
    }
    desc[((size_t)frame * cap + slot) * 32 + lane] = (uint8_t)val;
    if (lane == 0) {
        viorb_keypoint kp;
        /* pt *= mvScaleFactor[level] for level > 0 (:1094-1101) */
        kp.x = level ? __fmul_rn((float)kx, L.scale) : (float)kx;
        kp.y = level ? __fmul_rn((float)ky, L.scale) : (float)ky;
        kp.size = (float)L.patchSize;
        kp.angle = angle;
        kp.response = (float)score;
        kp.octave = level;
        kp.class_id = -1;
        kps[(size_t)frame * cap + slot] = kp;
    }
    __syncwarp();           /* the next keypoint's copies overwrite the blurred patch */
    }
}


/* ------------------------------------------------------------------------------------------------
 * GaussianBlur 7x7 sigma 2 of every pyramid level (:1085-1086), dense.  A batch holds about as many keypoint
 * neighbourhoods (1 000 x 43 x 43 px per EuRoC frame) as pyramid pixels (1.12 M) and the neighbourhoods overlap -- 3x on
 * the small levels -- so blurring the levels once costs less than half the instructions of blurring per keypoint, and
 * the work is regular: a thread owns an 8-pixel column strip of a band of output rows and walks down it.
 *   horizontal: the row's 4 words (one LDG.64 + two LDG.32) -> eight 16-bit sums, 20 IDP.4A with shifted tap words
 *   vertical:   the sums of rows (2m, 2m+1) packed into one word per column; an output row needs four such words
 *               (IDP.2A x4); the window of four pair rows lives in registers and rotates (loop unrolled by four)
 * The fixed-point blur rounds once at the end ((sum + 32768) >> 16), so it is exact in any evaluation order.  The
 * source is the padded pyramid, whose REFLECT_101 border equals the blur's own border rule; only ROI pixels are
 * written (same layout as the pyramid; a keypoint is >= 19 px from the level's edge and the pattern reaches 18).
 * ---------------------------------------------------------------------------------------------- */
#define BLUR_THREADS 128

struct BlurRow { unsigned wm; uint2 w01; unsigned w2; };     /* the 16 bytes around an 8-pixel strip of one row */

__device__ __forceinline__ BlurRow blur_load_row(const uint8_t* __restrict__ p) {
    BlurRow r;
    r.wm = __ldg(reinterpret_cast<const unsigned*>(p - 4));
    r.w01 = __ldg(reinterpret_cast<const uint2*>(p));
    r.w2 = __ldg(reinterpret_cast<const unsigned*>(p + 8));
    return r;
}

template <int GV>
__device__ __forceinline__ void blur_hrow8(const BlurRow& r, unsigned (&h)[8]) {
    h[0] = blur_hsum<GV, 1>(r.wm, r.w01.x, r.w01.y, r.w2);
    h[1] = blur_hsum<GV, 2>(r.wm, r.w01.x, r.w01.y, r.w2);
    h[2] = blur_hsum<GV, 3>(r.wm, r.w01.x, r.w01.y, r.w2);
    h[3] = blur_hsum<GV, 4>(r.wm, r.w01.x, r.w01.y, r.w2);
    h[4] = blur_hsum<GV, 1>(r.w01.x, r.w01.y, r.w2, 0u);
    h[5] = blur_hsum<GV, 2>(r.w01.x, r.w01.y, r.w2, 0u);
    h[6] = blur_hsum<GV, 3>(r.w01.x, r.w01.y, r.w2, 0u);
    h[7] = blur_hsum<GV, 4>(r.w01.x, r.w01.y, r.w2, 0u);
}

/* rows (2m, 2m+1) -> one word per column: low half = the sums of row 2m, high half = row 2m+1 */
template <int GV>
__device__ __forceinline__ void blur_pair8(const BlurRow& r0, const BlurRow& r1, unsigned (&P)[8]) {
    unsigned h0[8], h1[8];
    blur_hrow8<GV>(r0, h0);
    blur_hrow8<GV>(r1, h1);
#pragma unroll
    for (int c = 0; c < 8; c++) P[c] = __byte_perm(h0[c], h1[c], 0x5410);
}

template <int GV>
__global__ void __launch_bounds__(BLUR_THREADS) blur_levels_kernel(const __grid_constant__ FrameGeom g,
                                                                   const uint8_t* __restrict__ pyr,
                                                                   uint8_t* __restrict__ blur, int frame0) {
    pdl_trigger();
    const int frame = frame0 + blockIdx.y;
    int t = blockIdx.x * BLUR_THREADS + threadIdx.x;
    if (t >= g.blurTasks) return;
    int l = 0;
    while (l + 1 < g.nlevels && t >= g.lv[l + 1].blurBase) l++;
    const LevelGeom& L = g.lv[l];
    t -= L.blurBase;
    const int band = t / L.blurStrips, strip = t - band * L.blurStrips;
    const int y0 = band * L.blurRows;
    const size_t step = (size_t)L.step;
    const size_t off = (size_t)frame * g.pyrFrameBytes + L.pyrOff + (size_t)(VIORB_EDGE + y0) * step + VIORB_ROI_X0 + 8 * strip;
    const uint8_t* src = pyr + off - 3 * step;        /* input row y0 - 3 */
    uint8_t* dst = blur + off;
    const int nrows = min(L.blurRows, L.h - y0);                /* output rows of this band that exist */
    constexpr unsigned K0 = blur_tap(GV, 0), K1 = blur_tap(GV, 1), K2 = blur_tap(GV, 2), K3 = blur_tap(GV, 3);
    constexpr unsigned K4 = blur_tap(GV, 4), K5 = blur_tap(GV, 5), K6 = blur_tap(GV, 6);
    constexpr unsigned E01 = K0 | (K1 << 8), E23 = K2 | (K3 << 8), E45 = K4 | (K5 << 8), E6 = K6;              /* even output row */
    constexpr unsigned O0 = K0 << 8, O12 = K1 | (K2 << 8), O34 = K3 | (K4 << 8), O56 = K5 | (K6 << 8);          /* odd output row */
    pdl_wait();                               /* the pyramid is complete from here on */
    unsigned P[4][8];
    BlurRow n0, n1;                           /* the next row pair, loaded one step ahead of its use */
    {
        const BlurRow a0 = blur_load_row(src), a1 = blur_load_row(src + step);
        const BlurRow b0 = blur_load_row(src + 2 * step), b1 = blur_load_row(src + 3 * step);
        const BlurRow c0 = blur_load_row(src + 4 * step), c1 = blur_load_row(src + 5 * step);
        n0 = blur_load_row(src + 6 * step);
        n1 = blur_load_row(src + 7 * step);
        blur_pair8<GV>(a0, a1, P[0]);
        blur_pair8<GV>(b0, b1, P[1]);
        blur_pair8<GV>(c0, c1, P[2]);
    }
    src += 8 * step;
    /* step k: input rows 6+2k, 7+2k (loaded during step k-1) complete the window of output rows 2k, 2k+1; the rows of
     * step k+1 are requested before the vertical pass (they lie inside the stored level: at most 12 rows below the ROI).
     * The kernel runs at the rate of the integer pipe (IDP.4A / IDP.2A issue at 64 lanes per clock and SM): 6.5 IDP per pixel */
#pragma unroll 1
    for (int r = 0; r < nrows; r += 8) {
#pragma unroll
        for (int j = 0; j < 4; j++) {
            if (r + 2 * j < nrows) {
                blur_pair8<GV>(n0, n1, P[(j + 3) & 3]);
                n0 = blur_load_row(src);
                n1 = blur_load_row(src + step);
                unsigned e[8], o[8];
#pragma unroll
                for (int c = 0; c < 8; c++) {
                    const unsigned a0 = P[j & 3][c], a1 = P[(j + 1) & 3][c], a2 = P[(j + 2) & 3][c], a3 = P[(j + 3) & 3][c];
                    e[c] = __dp2a_lo(a0, E01, __dp2a_lo(a1, E23, __dp2a_lo(a2, E45, __dp2a_lo(a3, E6, 32768u))));
                    o[c] = __dp2a_lo(a0, O0, __dp2a_lo(a1, O12, __dp2a_lo(a2, O34, __dp2a_lo(a3, O56, 32768u))));
                    if constexpr (GV == 1) {      /* taps sum to 257: saturate_cast<uchar> */
                        e[c] = min(e[c], 0x00ffffffu);
                        o[c] = min(o[c], 0x00ffffffu);
                    }
                }
                /* (sum >> 16) is byte 2 of each result */
                const uint2 we = make_uint2(__byte_perm(__byte_perm(e[0], e[1], 0x0062), __byte_perm(e[2], e[3], 0x0062), 0x5410),
                                            __byte_perm(__byte_perm(e[4], e[5], 0x0062), __byte_perm(e[6], e[7], 0x0062), 0x5410));
                *reinterpret_cast<uint2*>(dst) = we;
                if (r + 2 * j + 1 < nrows) {
                    const uint2 wo = make_uint2(__byte_perm(__byte_perm(o[0], o[1], 0x0062), __byte_perm(o[2], o[3], 0x0062), 0x5410),
                                                __byte_perm(__byte_perm(o[4], o[5], 0x0062), __byte_perm(o[6], o[7], 0x0062), 0x5410));
                    *reinterpret_cast<uint2*>(dst + step) = wo;
                }
                src += 2 * step;
                dst += 2 * step;
            }
        }
    }
}

/* ------------------------------------------------------------------------------------------------
 * IC_Angle (:77-104) + computeOrbDescriptor (:108-147) on the blurred levels: one warp per selected keypoint.
 * The circular patch of the level (31 rows x 48 bytes) and the 37 x 37 blurred neighbourhood (37 rows x 64 bytes) are
 * staged with 16-byte cp.async from the 16-byte aligned rows at or left of the keypoint's window; the consumers add the
 * byte offset.  Orientation: lane v+15 sums row v (three conflict-free LDS.128; masked IDP.4A per word).
 * ---------------------------------------------------------------------------------------------- */
#define D2_RAW_STRIDE 48          /* bytes per staged row of the level patch (31 + <= 15 alignment bytes; 3 x 16 B: LDS.128 of eight rows hit eight bank groups) */
#define D2_BLUR_STRIDE 64         /* bytes per staged row of the blurred patch (37 + <= 15 alignment bytes) */
#define D2_RAW_BYTES (31 * D2_RAW_STRIDE)
#define D2_BLUR_BYTES (37 * D2_BLUR_STRIDE)

/* cvRound of a float with |x| < 2^22 on the FMA pipe: adding 1.5 * 2^23 leaves the integer (round to nearest even, like
 * cvRound's lrint) in the low mantissa bits; the caller subtracts D2_RND_BITS -- F2I would go through the quarter-rate XU pipe */
#define D2_RND_MAGIC 12582912.0f
#define D2_RND_BITS 0x4B400000

__global__ void __launch_bounds__(DESC_WARPS * 32, 7) describe_blurred_kernel(const __grid_constant__ FrameGeom g,
                                                                          const uint8_t* __restrict__ pyr,
                                                                          const uint8_t* __restrict__ blur,
                                                                          const uint32_t* __restrict__ sel,
                                                                          const int* __restrict__ selCount,
                                                                          viorb_keypoint* __restrict__ kps,
                                                                          uint8_t* __restrict__ desc, int cap,
                                                                          int32_t* __restrict__ counts,
                                                                          int* __restrict__ status, int kpw, int frame0) {
    /* A warp owns up to kpw (<= 32) keypoints and works in two sweeps: (A) the intensity-centroid moments of every
     * keypoint, then atan2 and sincos of all of them at once -- lane j does the scalar math of keypoint j, which would
     * otherwise be a 150-instruction dependent chain run by 32 identical lanes per keypoint; (B) the steered tests.
     * Both sweeps stage through two buffers per warp so that the copies of keypoint j+1 and j+2 are in flight while
     * keypoint j is being worked on. */
    __shared__ __align__(16) uint8_t rawS[DESC_WARPS][2][D2_RAW_BYTES];
    __shared__ __align__(16) uint8_t blurS[DESC_WARPS][2][D2_BLUR_BYTES];
    const int frame = frame0 + blockIdx.y;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    pdl_trigger();
    float4 pat[8];      /* this lane's 8 binary tests, kept for all keypoints of the warp */
#pragma unroll
    for (int k = 0; k < 8; k++) pat[k] = __ldg(reinterpret_cast<const float4*>(d_pattern) + 8 * lane + k);
    const int rawRow = lane / 3, rawCh = lane - 3 * rawRow;       /* this lane's (row, chunk) of a level-patch copy step */
    const unsigned rawDst = (unsigned)__cvta_generic_to_shared(rawS[warp][0]) + rawRow * D2_RAW_STRIDE + 16 * rawCh;
    const unsigned blurDst = (unsigned)__cvta_generic_to_shared(blurS[warp][0]) + (lane >> 2) * D2_BLUR_STRIDE + 16 * (lane & 3);
    unsigned icMask[8];       /* IC_Angle: bytes of this lane's patch row (u = 4j-15 .. 4j-12 in word j) inside the circle */
    {
        const int v = lane - 15, d = c_umax[min(v < 0 ? -v : v, 15)];
#pragma unroll
        for (int j = 0; j < 8; j++)
            icMask[j] = shl_clamp(0xffffffffu, 8 * max(15 - d - 4 * j, 0)) & shr_clamp(0xffffffffu, 8 * max(4 * j + 3 - 15 - d, 0));
    }
    pdl_wait();                               /* the selected keypoints and the blurred levels are complete from here on */
    /* lane l holds [lo, hi) of level l in the concatenated per-level lists (:1076-1103) */
    int myHi, total;
    {
        const int c = lane < g.nlevels ? selCount[frame * g.nlevels + lane] : 0;
        int incl = c;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int t = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += t;
        }
        myHi = incl;
        total = __shfl_sync(0xffffffffu, incl, 31);
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        counts[frame] = min(total, cap);
        if (total > cap) atomicOr(status, VIORB_DEV_OUT_OVERFLOW);
    }
    /* the warp's keypoints: lane j keeps keypoint j -- its key and level, the row stride of the level and the offsets
     * (inside the frame's pyramid block, < 2 GiB) of the 16-byte aligned top-left corners of its two staged windows */
    uint32_t myKey = 0;
    int myLevel = 0, myStep = 0;
    unsigned myRawOff = 0, myBlurOff = 0;
    int nkp;
    {
        const int slot = (lane * gridDim.x + blockIdx.x) * DESC_WARPS + warp;
        const bool valid = lane < kpw && slot < total && slot < cap;
        nkp = __popc(__ballot_sync(0xffffffffu, valid));          /* slots grow with the lane: the valid lanes are a prefix */
        int lo = 0;
        for (int l = 0; l + 1 < g.nlevels; l++) {
            const int hi = __shfl_sync(0xffffffffu, myHi, l);
            if (slot >= hi) { myLevel = l + 1; lo = hi; }
        }
        if (valid) {
            const LevelGeom& L = g.lv[myLevel];
            myKey = sel[(size_t)frame * g.selPerFrame + L.selBase + (slot - lo)];
            const int kx = myKey & 0xfff, ky = (myKey >> 12) & 0xfff;
            myStep = L.step;
            myRawOff = (unsigned)L.pyrOff + (unsigned)(VIORB_EDGE + ky - 15) * (unsigned)L.step + VIORB_ROI_X0 + ((kx - 15) & ~15);
            myBlurOff = (unsigned)L.pyrOff + (unsigned)(VIORB_EDGE + ky - 18) * (unsigned)L.step + VIORB_ROI_X0 + ((kx - 18) & ~15);
        }
    }
    const uint8_t* pyrF = pyr + (size_t)frame * g.pyrFrameBytes;
    const uint8_t* blurF = blur + (size_t)frame * g.pyrFrameBytes;
    /* sixteen-byte copies of keypoint j into buffer j & 1, lane -> (row, chunk) fixed: ten rows x 3 chunks of the level
     * patch (4 instructions) / eight rows x 4 chunks of the blurred patch (5 instructions) per step; the lanes of a row
     * read neighbouring chunks, so each 32-byte sector is fetched once.  One cp.async group per call (an empty one past
     * the last keypoint keeps the group count uniform). */
    auto stage_raw = [&](int j) {
        if (j < nkp) {
            const unsigned st = (unsigned)__shfl_sync(0xffffffffu, myStep, j);
            const uint8_t* src = pyrF + (__shfl_sync(0xffffffffu, myRawOff, j) + (unsigned)rawRow * st + 16u * (unsigned)rawCh);
            const size_t rstep = (size_t)(10u * st);
            unsigned d = rawDst + (j & 1) * D2_RAW_BYTES;
#pragma unroll
            for (int k = 0; k < 4; k++) {
                if (lane < 30 && (k < 3 || lane < 3))
                    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d), "l"(src) : "memory");
                src += rstep;
                d += 10 * D2_RAW_STRIDE;
            }
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
    };
    /* the fourth chunk of a blurred row is only needed (and only inside the stored row) when the window starts at
     * byte >= 12 of its first chunk */
    auto stage_blur = [&](int j) {
        if (j < nkp) {
            const unsigned st = (unsigned)__shfl_sync(0xffffffffu, myStep, j);
            const unsigned kx = __shfl_sync(0xffffffffu, myKey, j) & 0xfff;
            const uint8_t* src = blurF + (__shfl_sync(0xffffffffu, myBlurOff, j) + (unsigned)(lane >> 2) * st + 16u * (unsigned)(lane & 3));
            const size_t bstep = (size_t)(8u * st);
            const bool chOk = (lane & 3) < 3 || ((kx - 18) & 15) >= 12;
            unsigned d = blurDst + (j & 1) * D2_BLUR_BYTES;
#pragma unroll
            for (int k = 0; k < 5; k++) {
                if (chOk && (k < 4 || lane < 20))
                    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d), "l"(src) : "memory");
                src += bstep;
                d += 8 * D2_BLUR_STRIDE;
            }
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
    };
    stage_raw(0);
    stage_raw(1);
    stage_blur(0);
    stage_blur(1);
    /* ---- sweep A: IC_Angle moments (:77-104).  Lane v+15 sums row v of the circular patch: three LDS.128, the eight
     * words that hold u = -15..16 funnelled into place (the word offset selects one of four instantiations, the byte
     * offset is a PRMT selector), masked to the row's half-width and one IDP.4A per word and moment. */
    int myM01 = 0, myM10 = 0;
    for (int it = 0; it < nkp; it++) {
        /* groups so far: raw 0, raw 1, blur 0, blur 1, raw 2 .. raw it+1: all but the latest min(it, 1) + 2... are needed;
         * waiting for everything except the most recent group is enough once it >= 1, and in iteration 0 three groups may stay */
        if (it == 0) asm volatile("cp.async.wait_group 3;" ::: "memory");
        else asm volatile("cp.async.wait_group 1;" ::: "memory");
        __syncwarp();
        const int roff = ((__shfl_sync(0xffffffffu, myKey, it) & 0xfff) - 15) & 15;      /* staged bytes roff .. roff+30 hold x = kx-15 .. kx+15 */
        const uint8_t* R = rawS[warp][it & 1];
        int m10 = 0, m01 = 0;
        if (lane < 31) {
            const uint4* rw = reinterpret_cast<const uint4*>(R + lane * D2_RAW_STRIDE);
            const uint4 q0 = rw[0], q1 = rw[1], q2 = rw[2];
            const unsigned w[12] = {q0.x, q0.y, q0.z, q0.w, q1.x, q1.y, q1.z, q1.w, q2.x, q2.y, q2.z, q2.w};
            const unsigned selN = 0x3210u + 0x1111u * (unsigned)(roff & 3);
            unsigned s1 = 0;
            auto rows = [&](auto O) {
                constexpr int o = decltype(O)::value;
#pragma unroll
                for (int j = 0; j < 8; j++) {
                    const unsigned x = __byte_perm(w[o + j], w[o + j + 1], selN) & icMask[j];
                    const int u0 = 4 * j - 15;
                    const int wts = (u0 & 0xff) | (((u0 + 1) & 0xff) << 8) | (((u0 + 2) & 0xff) << 16) | (((u0 + 3) & 0xff) << 24);
                    m10 = dp4a_u8s8(x, wts, m10);
                    s1 = __dp4a(x, 0x01010101u, s1);
                }
            };
            switch (roff >> 2) {
                case 0: rows(std::integral_constant<int, 0>()); break;
                case 1: rows(std::integral_constant<int, 1>()); break;
                case 2: rows(std::integral_constant<int, 2>()); break;
                default: rows(std::integral_constant<int, 3>()); break;
            }
            m01 = (lane - 15) * (int)s1;
        }
        m10 = __reduce_add_sync(0xffffffffu, m10);          /* REDUX.SUM (also orders the reads before the next copies) */
        m01 = __reduce_add_sync(0xffffffffu, m01);
        if (lane == it) { myM10 = m10; myM01 = m01; }
        stage_raw(it + 2);
    }
    /* orientation and steering of all the warp's keypoints at once: lane j = keypoint j (:103, :112-113) */
    const float myAngle = fast_atan2_deg((float)myM01, (float)myM10);
    float myA, myB;
    sincosf_glibc(__fmul_rn(myAngle, (float)(3.14159265358979323846 / 180.f)), &myB, &myA);
    /* ---- sweep B: steered BRIEF, lane i produces descriptor byte i (:123-144) */
    for (int it = 0; it < nkp; it++) {
        asm volatile("cp.async.wait_group 1;" ::: "memory");       /* blur it is complete (blur it+1 may be in flight) */
        __syncwarp();
        const int slot = (it * gridDim.x + blockIdx.x) * DESC_WARPS + warp;
        const uint32_t key = __shfl_sync(0xffffffffu, myKey, it);
        const int level = __shfl_sync(0xffffffffu, myLevel, it);
        const float angle = __shfl_sync(0xffffffffu, myAngle, it);
        const float a = __shfl_sync(0xffffffffu, myA, it), b = __shfl_sync(0xffffffffu, myB, it);
        const LevelGeom& L = g.lv[level];
        const int kx = key & 0xfff, ky = (key >> 12) & 0xfff, score = key >> 24;
        const int boff = (kx - 18) & 15;      /* staged bytes boff .. boff+36 hold x = kx-18 .. kx+18 */
        /* shared address = centre + r * 64 + c with r, c still biased by D2_RND_BITS: the bias goes into the base (32-bit
         * wrap-around arithmetic) */
        const unsigned centre = (unsigned)__cvta_generic_to_shared(blurS[warp][it & 1]) + 18 * D2_BLUR_STRIDE + boff + 18 -
                                (unsigned)D2_RND_BITS * (D2_BLUR_STRIDE + 1);
        auto tap = [&](float x, float y) -> unsigned {
            const unsigned r = __float_as_uint(__fadd_rn(__fadd_rn(__fmul_rn(x, b), __fmul_rn(y, a)), D2_RND_MAGIC));
            const unsigned c = __float_as_uint(__fadd_rn(__fsub_rn(__fmul_rn(x, a), __fmul_rn(y, b)), D2_RND_MAGIC));
            unsigned v;
            asm volatile("ld.shared.u8 %0, [%1];" : "=r"(v) : "r"(r * D2_BLUR_STRIDE + c + centre));
            return v;
        };
        unsigned val = 0;
#pragma unroll
        for (int k = 0; k < 8; k++) val |= (unsigned)(tap(pat[k].x, pat[k].y) < tap(pat[k].z, pat[k].w)) << k;
        desc[((size_t)frame * cap + slot) * 32 + lane] = (uint8_t)val;
        if (lane == 0) {
            viorb_keypoint kp;
            kp.x = level ? __fmul_rn((float)kx, L.scale) : (float)kx;       /* pt *= mvScaleFactor[level] (:1094-1101) */
            kp.y = level ? __fmul_rn((float)ky, L.scale) : (float)ky;
            kp.size = (float)L.patchSize;
            kp.angle = angle;
            kp.response = (float)score;
            kp.octave = level;
            kp.class_id = -1;
            kps[(size_t)frame * cap + slot] = kp;
        }
        __syncwarp();           /* the copies of keypoint it + 2 overwrite this buffer */
        stage_blur(it + 2);
    }
}

}  // namespace

/* ------------------------------------------------------------------------------------------------ launchers */
/* kernel launch with or without the programmatic-dependent-launch attribute (see pdl_wait).  The launchers take a `pdl`
 * mask: VIORB_PDL_INNER = between the kernels of one stage, VIORB_PDL_EDGE = on the first kernel of a stage (off while the
 * per-stage timers are on: an event is recorded between the stages then). */
template <typename... KArgs, typename... Args>
static void launch_k(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t s, bool pdl, Args&&... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = s;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at;
    cfg.numAttrs = pdl ? 1 : 0;
    cudaLaunchKernelEx(&cfg, kern, static_cast<KArgs>(args)...);
}

int viorb_launch_pyramid(const FrameGeom& g, const ResizeTables& t, const uint8_t* d_images, size_t step,
                         size_t frameStride, int F, const ExtractBuffers& b, cudaStream_t s, int pdl) {
    const bool inner = (pdl & VIORB_PDL_INNER) != 0;
    const int aligned = ((uintptr_t)d_images % 16 == 0) && (step % 16 == 0) && (frameStride % 16 == 0);
    int launches = 0;
    for (int l = 0; l < g.nlevels; l++) {
        const LevelGeom& L = g.lv[l];
        if (l == 0) {
            dim3 grid((L.h + 2 * VIORB_EDGE + L0_ROWS - 1) / L0_ROWS, F);
            /* the first kernel of a pass follows a memset or another pass: ordinary stream order */
            launch_k(pyr_level0_kernel, grid, dim3(256), 0, s, false, g, d_images, step, frameStride, b.pyr, aligned);
        } else {
            if (F <= 8) {        /* few frames: 128 x 16 tiles, four times the CTAs, a quarter of the per-tile latency */
                dim3 grid((L.step + RZ_TW - 1) / RZ_TW, (L.h + 2 * VIORB_EDGE + 15) / 16, F);
                launch_k(pyr_resize_kernel<4>, grid, dim3(128), 0, s, inner, g, l, t, b.pyr);
            } else {
                /* 128 x 64 tiles on every level.  Measured alternatives for the levels above the third, whose launches hold
                 * less than three waves of CTAs and lose 20-47 % of their threads to ragged tiles: 128 x 32 and 128 x 16 tiles
                 * (more CTAs, but each repeats the set-up: 4.9 -> 4.9 / 5.2 ms per 4096 frames) and a tile-free kernel with one
                 * thread per stored word and 16-row band reading the source rows straight from L1/L2 (no staging, no ragged
                 * tiles, a third fewer instructions, but two dependent global loads per row: 5.6 ms) */
                dim3 grid((L.step + RZ_TW - 1) / RZ_TW, (L.h + 2 * VIORB_EDGE + RZ_TH - 1) / RZ_TH, F);
                launch_k(pyr_resize_kernel<RZ_WROWS>, grid, dim3(128), 0, s, inner, g, l, t, b.pyr);
            }
        }
        launches++;
    }
    return launches;
}

size_t viorb_fast_smem_bytes(const FrameGeom& g) {
    return (size_t)g.fastPixBytes + (size_t)(g.fastTileRows - 4) * FAST_SCW * 4 + (size_t)g.fastMaxWork * 2;
}

int viorb_fast_prepare(const FrameGeom& g) {
    static int current[64] = {0};        /* raise-only, see viorb_octree_prepare */
    static std::mutex guard;             /* extractors are created from several threads (stereo: Frame.cc:258-261) */
    std::lock_guard<std::mutex> lock(guard);
    int dev = 0;
    cudaGetDevice(&dev);
    const int smem = (int)viorb_fast_smem_bytes(g);
    if (dev < 0 || dev >= 64 || smem <= current[dev] || smem <= 48 * 1024) return (int)cudaSuccess;
    current[dev] = smem;
    cudaError_t e = cudaFuncSetAttribute(fast_cells_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(fast_cells_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(fast_cells_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(fast_cells_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    return (int)e;
}

int viorb_launch_fast(const FrameGeom& g, const TmaMaps& maps, const int4* d_groups, const int* classStart, int F,
                      const ExtractBuffers& b, cudaStream_t s, int pdl) {
    int launches = 0;
    const size_t smem = viorb_fast_smem_bytes(g);
    for (int sh = 0; sh < 4; sh++) {
        const int n = classStart[sh + 1] - classStart[sh];
        if (n <= 0) continue;
        const bool first = launches == 0;
        dim3 grid(n, F);
        const int4* grp = d_groups + 2 * (size_t)classStart[sh];       /* two 16-byte words per group */
        switch (sh) {
            case 0: launch_k(fast_cells_kernel<0>, grid, dim3(128), smem, s, (pdl & (first ? VIORB_PDL_EDGE : VIORB_PDL_INNER)) != 0, g, maps, grp, b.cand, b.candCount, b.status); break;
            case 1: launch_k(fast_cells_kernel<1>, grid, dim3(128), smem, s, (pdl & (first ? VIORB_PDL_EDGE : VIORB_PDL_INNER)) != 0, g, maps, grp, b.cand, b.candCount, b.status); break;
            case 2: launch_k(fast_cells_kernel<2>, grid, dim3(128), smem, s, (pdl & (first ? VIORB_PDL_EDGE : VIORB_PDL_INNER)) != 0, g, maps, grp, b.cand, b.candCount, b.status); break;
            default: launch_k(fast_cells_kernel<3>, grid, dim3(128), smem, s, (pdl & (first ? VIORB_PDL_EDGE : VIORB_PDL_INNER)) != 0, g, maps, grp, b.cand, b.candCount, b.status); break;
        }
        launches++;
    }
    return launches;
}

/* CUtensorMap of every stored pyramid level {step bytes, h+38 rows, F frames}, box = the FAST tile of the level.
 * cuTensorMapEncodeTiled is a driver entry point; it is resolved through the runtime so that the library does
 * not link libcuda (the build box has no driver). */
int viorb_encode_tma_maps(const FrameGeom& g, uint8_t* d_pyr, int F, TmaMaps* out) {
    typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                 const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                 CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
    static EncodeFn encode = nullptr;
    if (!encode) {
        void* fn = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres) != cudaSuccess || !fn) return -1;
        encode = (EncodeFn)fn;
    }
    static_assert(sizeof(CUtensorMap) == 128, "CUtensorMap size");
    for (int l = 0; l < g.nlevels; l++) {
        const LevelGeom& L = g.lv[l];
        const cuuint64_t dims[3] = {(cuuint64_t)L.step, (cuuint64_t)(L.h + 2 * VIORB_EDGE), (cuuint64_t)F};
        const cuuint64_t strides[2] = {(cuuint64_t)L.step, (cuuint64_t)g.pyrFrameBytes};
        int boxH = L.hCell + 6 < g.fastTileRows ? L.hCell + 6 : g.fastTileRows;
        if (boxH < 1) boxH = 1;
        const cuuint32_t box[3] = {VIORB_FAST_TILE_BYTES, (cuuint32_t)boxH, 1};
        const cuuint32_t estr[3] = {1, 1, 1};
        CUtensorMap m;
        const CUresult r = encode(&m, CU_TENSOR_MAP_DATA_TYPE_UINT8, 3, d_pyr + L.pyrOff, dims, strides, box, estr,
                                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) return -2;
        memcpy(out->fast[l], &m, 128);
    }
    return 0;
}

size_t viorb_octree_smem_bytes(int NC) {
    int p2 = 1;
    while (p2 < NC) p2 <<= 1;
    return (size_t)NC * (2 * sizeof(short4) + 2 * sizeof(int) + 4 * sizeof(int) + 2 * sizeof(int) + 4 * sizeof(unsigned short)) +
           (size_t)p2 * sizeof(unsigned);
}

/* The opt-in limit is a property of the kernel (per device), shared by every extractor of the process: it is only
 * ever raised, so an extractor with a small quota cannot lower it under one with a large quota. */
int viorb_octree_prepare(int NC) {
    static int current[64] = {0};
    static std::mutex guard;
    std::lock_guard<std::mutex> lock(guard);
    int dev = 0;
    cudaGetDevice(&dev);
    const int want = (int)viorb_octree_smem_bytes(NC);
    if (dev < 0 || dev >= 64 || want <= current[dev]) return (int)cudaSuccess;
    const cudaError_t e = cudaFuncSetAttribute(octree_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, want);
    if (e == cudaSuccess) current[dev] = want;
    return (int)e;
}

int viorb_launch_octree(const FrameGeom& g, int F, const ExtractBuffers& b, int nodeCap, cudaStream_t s, int pdl) {
    dim3 grid(g.nlevels, F);
    launch_k(octree_kernel, grid, dim3(F <= 8 ? OCT_THREADS : OCT_THREADS / 2), viorb_octree_smem_bytes(nodeCap), s, (pdl & VIORB_PDL_EDGE) != 0,
             g, b.cand, b.candCount, b.nodeOf, b.sel, b.selCount, b.status, nodeCap);
    return 1;
}

/* parity hook: the steering pair (b, a) of :112-113 for the consecutive float bit patterns firstBits, firstBits+1, ...
 * taken as keypoint angles in degrees -- lets the tests sweep every angle the extractor can produce */
__global__ void __launch_bounds__(256) steering_sweep_kernel(unsigned firstBits, long long n, float* __restrict__ sinOut,
                                                             float* __restrict__ cosOut) {
    const float factorPI = (float)(3.14159265358979323846 / 180.f);
    for (long long i = blockIdx.x * 256ll + threadIdx.x; i < n; i += gridDim.x * 256ll) {
        float a, b;
        sincosf_glibc(__fmul_rn(__uint_as_float(firstBits + (unsigned)i), factorPI), &b, &a);
        sinOut[i] = b;
        cosOut[i] = a;
    }
}

/* parity hook: the orientation of IC_Angle (:103) for given integer moments */
__global__ void __launch_bounds__(256) orientation_sweep_kernel(const int* __restrict__ m01, const int* __restrict__ m10,
                                                                long long n, float* __restrict__ deg) {
    for (long long i = blockIdx.x * 256ll + threadIdx.x; i < n; i += gridDim.x * 256ll)
        deg[i] = fast_atan2_deg((float)m01[i], (float)m10[i]);
}

int viorb_launch_orientation_sweep(const int* d_m01, const int* d_m10, long long n, float* d_deg, int sms, cudaStream_t s) {
    orientation_sweep_kernel<<<sms * 8, 256, 0, s>>>(d_m01, d_m10, n, d_deg);
    return 1;
}

int viorb_launch_steering_sweep(unsigned firstBits, long long n, float* d_sin, float* d_cos, int sms, cudaStream_t s) {
    steering_sweep_kernel<<<sms * 8, 256, 0, s>>>(firstBits, n, d_sin, d_cos);
    return 1;
}

int viorb_launch_describe(const FrameGeom& g, int frame0, int F, const ExtractBuffers& b, viorb_keypoint* d_kps, uint8_t* d_desc,
                          int cap, int32_t* d_counts, cudaStream_t s, int pdl) {
    const bool edge = (pdl & VIORB_PDL_EDGE) != 0;
    const int slots = g.selPerFrame < cap ? g.selPerFrame : cap;
    /* keypoints per warp: batches amortise the warp's pattern registers over up to DESC_KPW keypoints, as long as the
     * grid still holds two waves of CTAs (148 SMs x 7); a few frames keep one keypoint per warp (shortest latency) */
    int kpw = (int)(((long long)F * slots) / (DESC_WARPS * 148 * 14));
    const int kmax = b.blur ? DESC2_KPW : DESC_KPW;         /* (8 keypoints per warp: 2.5 % slower describe stage; 32: no different from 16) */
    kpw = kpw < 1 ? 1 : (kpw > kmax ? kmax : kpw);
    dim3 grid((slots + DESC_WARPS * kpw - 1) / (DESC_WARPS * kpw), F);
    if (grid.x == 0) grid.x = 1;
    if (b.blur) {
        launch_k(describe_blurred_kernel, grid, dim3(DESC_WARPS * 32), 0, s, edge, g, b.pyr, b.blur, b.sel, b.selCount, d_kps, d_desc, cap, d_counts, b.status, kpw, frame0);
        return 1;
    }
    /* the fused kernel indexes frames by blockIdx.y: whole passes only */
    if (g.gaussVariant)
        launch_k(orient_describe_kernel<1>, grid, dim3(DESC_WARPS * 32), 0, s, edge, g, b.pyr, b.sel, b.selCount, d_kps, d_desc, cap, d_counts, b.status, kpw);
    else
        launch_k(orient_describe_kernel<0>, grid, dim3(DESC_WARPS * 32), 0, s, edge, g, b.pyr, b.sel, b.selCount, d_kps, d_desc, cap, d_counts, b.status, kpw);
    return 1;
}

/* GaussianBlur of all levels of frames [frame0, frame0 + F) (no-op when the lane has no blur buffer: the fused kernel
 * blurs per keypoint) */
int viorb_launch_blur(const FrameGeom& g, int frame0, int F, const ExtractBuffers& b, cudaStream_t s, int pdl) {
    if (!b.blur) return 0;
    const bool edge = (pdl & VIORB_PDL_EDGE) != 0;
    dim3 grid((g.blurTasks + BLUR_THREADS - 1) / BLUR_THREADS, F);
    if (g.gaussVariant)
        launch_k(blur_levels_kernel<1>, grid, dim3(BLUR_THREADS), 0, s, edge, g, (const uint8_t*)b.pyr, b.blur, frame0);
    else
        launch_k(blur_levels_kernel<0>, grid, dim3(BLUR_THREADS), 0, s, edge, g, (const uint8_t*)b.pyr, b.blur, frame0);
    return 1;
}
