/*
 * c_api.cu -- the extern "C" boundary of libviorb_b200.so (include/viorb_gpu.h): contexts, device
 * workspaces, host<->device staging and kernel orchestration.  No compute happens on the host; if
 * CUDA is unavailable every entry point fails with VIORB_ERR_CUDA (there is no CPU fallback).
 */
#include <cuda_runtime.h>

#include <algorithm>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>
#include <vector>

#include "extractor_kernels.cuh"
#include "matcher_kernels.cuh"
#include "viorb_gpu.h"
#include "viorb_internal.h"

namespace {

thread_local char g_err[512] = "";

int fail(int code, const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    return code;
}

#define CU(call)                                                                                   \
    do {                                                                                           \
        cudaError_t e_ = (call);                                                                   \
        if (e_ != cudaSuccess)                                                                     \
            return fail(VIORB_ERR_CUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); \
    } while (0)

inline int cvRoundF(float v) { return (int)lrintf(v); }

template <typename T>
struct DevBuf {
    T* p = nullptr;
    size_t n = 0;
    int ensure(size_t count) {
        if (count <= n) return VIORB_OK;
        if (p) cudaFree(p);
        p = nullptr; n = 0;
        cudaError_t e = cudaMalloc((void**)&p, count * sizeof(T));
        if (e != cudaSuccess) return fail(VIORB_ERR_CUDA, "cudaMalloc(%zu) failed: %s", count * sizeof(T), cudaGetErrorString(e));
        n = count;
        return VIORB_OK;
    }
    void release() { if (p) cudaFree(p); p = nullptr; n = 0; }
};

}  // namespace

struct viorb_ctx {
    int device = 0;
    cudaStream_t stream = nullptr;
    bool ownStream = false;
    cudaStream_t h2d = nullptr, d2h = nullptr;
    int sms = 148;
    int64_t launches = 0;
    /* matcher scratch */
    DevBuf<uint8_t> mq, mmap;
    DevBuf<viorb_top2> mparts, mout;
    DevBuf<uint8_t> scratchA, scratchB;
    DevBuf<int32_t> scratchI;
    DevBuf<uint8_t> arena;
    /* pinned host mirrors of the arena: a matcher call packs all its inputs into one block and moves it with ONE copy.
     * slot 0 serves the searches (they end in a stream synchronisation), slot 1 the frame-index builds (which do not);
     * the event records when the last upload from a slot has been consumed */
    uint8_t* stage[2] = {nullptr, nullptr};
    size_t stageBytes[2] = {0, 0};
    cudaEvent_t stageDone[2] = {nullptr, nullptr};
    bool stagePending[2] = {false, false};
    /* device blocks of destroyed frame indices, reused by the next index of the context (stream ordered: no cudaMalloc /
     * cudaFree and no synchronisation on the per-frame path) */
    std::vector<std::pair<size_t, uint8_t*> > pool;
};

#define VIORB_MAX_LANES 4
struct viorb_extractor {
    viorb_ctx* ctx = nullptr;
    int nfeatures = 0, nlevels = 0, iniTh = 0, minTh = 0;
    double scaleFactor = 1.2;
    std::vector<float> scale, invScale, sigma2, invSigma2;
    std::vector<int> quota;
    int chunk = 128, candDiv = 16;
    int gaussVariant = 0;        /* viorb_extractor_set_gaussian */
    int copyMode = 0;            /* viorb_extractor_set_copy_mode: 0 = input and output copies on two streams, 1 = on one */
    bool chunkUser = false;      /* viorb_extractor_configure chose the pass size */
    /* GaussianBlur of whole levels (blur_levels_kernel + describe_blurred_kernel) or per keypoint inside
     * orient_describe_kernel.  Measured cost per frame in thread instructions: 12.9 per pyramid pixel + 16.5 k per keypoint
     * against 46.4 k per keypoint, so whole levels win while the pyramid has fewer than about 2300 pixels per feature
     * (EuRoC 1100, KITTI 720, 1080p 1300: +9 %, +21 %, +4 % frames/s; 4K 5100: -7 %); single-frame calls keep the fused
     * kernel (one launch less on the latency path).  VIORB_DESCRIBE=fused|dense overrides. */
    int describeMode = 0;        /* 0 = by the rule above, 1 = always fused, 2 = always whole levels */
    bool denseBlur = false;      /* decision for the current geometry (build_geometry) */
    /* geometry for the current image size */
    int rows = 0, cols = 0;
    FrameGeom geom;
    int nodeCap = 0;
    DevBuf<uint4> tabCol, tabRow;  /* resize tables (ResizeTables) */
    DevBuf<int4> groups;           /* FAST cell groups: {level, cell row, first cell, cells} */
    int ngroups = 0;
    int groupClass[5] = {0, 0, 0, 0, 0};   /* groups sorted by tile byte shift: class sh = [groupClass[sh], groupClass[sh+1]) */
    ResizeTables tables;
    /* pass workspaces: consecutive passes rotate over `nlanes` lanes (own stream + buffers) so that the
     * latency-bound kernels of one pass overlap with those of the next */
    struct Lane {
        cudaStream_t stream = nullptr;
        cudaEvent_t evDone = nullptr;
        int allocFrames = 0;
        ExtractBuffers buf = {};
        DevBuf<uint8_t> pyr, blur;
        DevBuf<uint32_t> cand, sel;
        DevBuf<int> counters;      /* candCount | selCount | status */
        DevBuf<uint16_t> nodeOf;
        TmaMaps maps;              /* TMA descriptors of this lane's pyramid buffer */
    } lanes[VIORB_MAX_LANES];
    int nlanes = 4;
    cudaEvent_t evFork = nullptr;
    int* hostStatus = nullptr;     /* pinned copy of the device status word (single-pass host path) */
    uint8_t* pyrHost = nullptr;    /* pinned staging of one frame's pyramid block (viorb_extractor_pyramid_download_all) */
    size_t pyrHostBytes = 0;
    /* CUDA graph of one single-frame pass (memset + 12 kernels) on lane 0: the per-frame call replays it instead of
     * issuing 13 launches; rebuilt when the geometry, the buffers or the capacity change */
    cudaGraphExec_t frameGraph = nullptr;
    const void* graphKey[4] = {nullptr, nullptr, nullptr, nullptr};
    int graphCap = 0, graphGen = -1, graphLaunches = 0, geomGen = 0;
    ExtractBuffers buf = {};       /* buffers of the most recent pass (resident pyramids, debug views) */
    /* staging for host-buffer entry points */
    /* host-buffer path: VIORB_SLOTS staging slots, slot s runs on lane s (H2D(k+2) || H2D(k+1) || compute(k) || D2H(k-1)) */
    DevBuf<uint8_t> in[4];
    DevBuf<viorb_keypoint> okps[4];
    DevBuf<uint8_t> odesc[4];
    DevBuf<int32_t> ocnt[4];
    cudaEvent_t evIn[4] = {nullptr, nullptr, nullptr, nullptr}, evDone[4] = {nullptr, nullptr, nullptr, nullptr},
                evOut[4] = {nullptr, nullptr, nullptr, nullptr};
    int nslots = 4;
    int residentFirst = 0, residentCount = 0;
    int lastOverflow = 0;
    bool profiling = false;
    std::vector<cudaEvent_t> profEvents;      /* 5 per pass: start, pyramid, fast, octree, describe */
    std::vector<cudaEvent_t> profPool;
};

namespace {

int ctx_bind(viorb_ctx* c) {
    CU(cudaSetDevice(c->device));
    return VIORB_OK;
}

/* ORBextractor::ORBextractor tables, src/ORBextractor.cc:410-446 */
void build_tables(viorb_extractor* e) {
    const int nl = e->nlevels;
    e->scale.assign(nl, 1.0f); e->sigma2.assign(nl, 1.0f);
    for (int i = 1; i < nl; i++) {
        e->scale[i] = (float)(e->scale[i - 1] * e->scaleFactor);
        e->sigma2[i] = e->scale[i] * e->scale[i];
    }
    e->invScale.resize(nl); e->invSigma2.resize(nl);
    for (int i = 0; i < nl; i++) {
        e->invScale[i] = 1.0f / e->scale[i];
        e->invSigma2[i] = 1.0f / e->sigma2[i];
    }
    e->quota.resize(nl);
    float factor = (float)(1.0f / e->scaleFactor);
    float nDesired = e->nfeatures * (1 - factor) / (1 - (float)pow((double)factor, (double)nl));
    int sum = 0;
    for (int l = 0; l < nl - 1; l++) {
        e->quota[l] = cvRoundF(nDesired);
        sum += e->quota[l];
        nDesired *= factor;
    }
    e->quota[nl - 1] = std::max(e->nfeatures - sum, 0);
}

/* frames per device pass when the caller did not choose: about 69 Mpixel of input per pass, at most 128 frames.  Larger
 * passes mean more CTAs per launch (full waves, shorter tails); L2 residency of a pass does not matter -- the kernels are
 * bound by integer issue (tools/shape_bench.py: 1080p 12 frames/pass 30.7 k, 48 frames/pass 32.4 k frames/s; 4K 3 -> 7.1 k,
 * 8 -> 8.8 k) */
int default_pass_frames(const viorb_extractor* e, int rows, int cols) {
    if (e->chunkUser) return e->chunk;
    const long long px = (long long)rows * cols;
    return (int)std::max<long long>(1, std::min<long long>(128, (192LL * 752 * 480 + px / 2) / px));
}

/* geometry + cv::resize coefficient tables for a given image size */
int build_geometry(viorb_extractor* e, int rows, int cols) {
    if (e->rows == rows && e->cols == cols) return VIORB_OK;
    /* the cached size is only valid while e->geom is complete: a failing rebuild must not leave the old size behind */
    e->rows = e->cols = 0;
    FrameGeom& g = e->geom;
    memset(&g, 0, sizeof(g));
    g.nlevels = e->nlevels; g.rows = rows; g.cols = cols; g.iniTh = e->iniTh; g.minTh = e->minTh;
    g.gaussVariant = e->gaussVariant;
    size_t pyrOff = 0;
    int cellBase = 0, candBase = 0, selBase = 0, xtab = 0, ytab = 0, nodeCap = 0, blurTasks = 0;
    for (int l = 0; l < e->nlevels; l++) {
        LevelGeom& L = g.lv[l];
        L.w = cvRoundF((float)cols * e->invScale[l]);       /* :1112 */
        L.h = cvRoundF((float)rows * e->invScale[l]);
        if (L.w > 4095 + 2 * VIORB_FAST_BORDER || L.h > 4095 + 2 * VIORB_FAST_BORDER)
            return fail(VIORB_ERR_UNSUPPORTED, "level %d is %dx%d: coordinates above 4095 are not supported", l, L.w, L.h);
        const int W = L.w - 2 * VIORB_FAST_BORDER, H = L.h - 2 * VIORB_FAST_BORDER;
        if (W < 30 || H < 30)
            return fail(VIORB_ERR_UNSUPPORTED, "level %d (%dx%d) is smaller than one FAST cell; use fewer levels", l, L.w, L.h);
        L.step = (VIORB_ROI_X0 + L.w + VIORB_EDGE + 15) / 16 * 16;
        L.pyrOff = (int)pyrOff;
        pyrOff += (size_t)L.step * (L.h + 2 * VIORB_EDGE);
        pyrOff = (pyrOff + 255) / 256 * 256;
        /* :779-787 */
        const float width = (float)W, height = (float)H;
        L.nCols = (int)(width / 30.f);
        L.nRows = (int)(height / 30.f);
        L.wCell = (int)ceilf(width / L.nCols);
        L.hCell = (int)ceilf(height / L.nRows);
        L.cellBase = cellBase;
        cellBase += L.nCols * L.nRows;
        L.quota = e->quota[l];
        L.nIni = (int)roundf((float)W / (float)H);          /* :543 */
        if (L.nIni < 1) return fail(VIORB_ERR_UNSUPPORTED, "portrait aspect ratio %dx%d gives zero quadtree roots", L.w, L.h);
        long cc = (long)L.w * L.h / e->candDiv;
        L.candCap = (int)std::min<long>(std::max<long>(cc, 1024), (1 << 20) - 1);
        L.candBase = candBase;
        candBase += L.candCap;
        L.selCap = std::max(L.quota + 4, 4 * L.nIni + 4);
        L.selBase = selBase;
        selBase += L.selCap;
        nodeCap = std::max(nodeCap, L.selCap);
        L.scale = e->scale[l];
        L.patchSize = (int)(31 * e->scale[l]);               /* :837 */
        L.xtab = xtab; L.ytab = ytab;
        xtab += L.step / 4; ytab += L.h + 2 * VIORB_EDGE;
        /* dense blur: bands of at most 48 output rows (the 6 extra input rows of a band cost 1/8), equal per level */
        const int bandMax = 48;            /* 24 .. 96 measure within 2 % of each other */
        const int bands = (L.h + bandMax - 1) / bandMax;
        L.blurRows = ((L.h + bands - 1) / bands + 7) / 8 * 8;
        L.blurStrips = (L.w + 7) / 8;
        L.blurBase = blurTasks;
        blurTasks += L.blurStrips * ((L.h + L.blurRows - 1) / L.blurRows);
    }
    g.blurTasks = blurTasks;
    {
        long long px = 0;
        for (int l = 0; l < e->nlevels; l++) px += (long long)g.lv[l].w * g.lv[l].h;
        e->denseBlur = e->describeMode == VIORB_DESCRIBE_LEVELS || (e->describeMode == VIORB_DESCRIBE_AUTO && px < 2300LL * e->nfeatures);
    }
    if (pyrOff >= (1ull << 31)) return fail(VIORB_ERR_UNSUPPORTED, "pyramid larger than 2 GiB per frame");
    g.pyrFrameBytes = pyrOff;
    g.cellsPerFrame = cellBase; g.candPerFrame = candBase; g.selPerFrame = selBase;
    if (nodeCap > 4096 || viorb_octree_smem_bytes(nodeCap) > 200 * 1024)
        return fail(VIORB_ERR_UNSUPPORTED, "per-level feature quota %d too large for the quadtree kernel", nodeCap);
    e->nodeCap = nodeCap;
    /* cv::resize INTER_LINEAR coefficient tables (OpenCV resize.cpp), level l from level l-1, expanded per stored word /
     * stored row of the padded level (ResizeTables, extractor_kernels.cuh) */
    std::vector<uint4> tcol(3 * (size_t)xtab), trow((size_t)ytab);
    auto reflect = [](int p, int n) { return p < 0 ? -p : (p >= n ? 2 * n - 2 - p : p); };
    for (int l = 1; l < e->nlevels; l++) {
        const LevelGeom& L = g.lv[l];
        const LevelGeom& P = g.lv[l - 1];
        const double inv_sx = (double)L.w / P.w, inv_sy = (double)L.h / P.h;
        const double sx_ = 1. / inv_sx, sy_ = 1. / inv_sy;
        std::vector<int> xs(L.w);
        std::vector<uint32_t> xc(L.w);
        for (int dx = 0; dx < L.w; dx++) {
            float fx = (float)((dx + 0.5) * sx_ - 0.5);
            int sx = (int)floorf(fx);
            fx -= sx;
            if (sx < 0) { fx = 0; sx = 0; }
            if (sx >= P.w - 1) { fx = 0; sx = P.w - 1; }
            xs[dx] = sx;
            xc[dx] = (uint32_t)(uint16_t)cvRoundF((1.f - fx) * 2048) | ((uint32_t)(uint16_t)cvRoundF(fx * 2048) << 16);
        }
        for (int wi = 0; wi < L.step / 4; wi++) {
            int rel[4];
            uint32_t coef[4], ok = 0;
            for (int j = 0; j < 4; j++) {
                const int x = wi * 4 + j - VIORB_ROI_X0;
                /* columns in the alignment padding borrow the nearest stored column so the 8-byte window stays valid */
                const int dx = reflect(std::min(std::max(x, -VIORB_EDGE), L.w + VIORB_EDGE - 1), L.w);
                rel[j] = xs[dx];
                coef[j] = xc[dx];
                if (x >= -VIORB_EDGE && x < L.w + VIORB_EDGE) ok |= 0xffu << (8 * j);
            }
            const int lo = std::min(std::min(rel[0], rel[1]), std::min(rel[2], rel[3]));
            const int hi = std::max(std::max(rel[0], rel[1]), std::max(rel[2], rel[3])) + 1;
            if (hi - lo > 7) return fail(VIORB_ERR_UNSUPPORTED, "scale factor too large for the resize kernel's 8-byte window");
            uint32_t sel[4];
            for (int j = 0; j < 4; j++) sel[j] = (uint32_t)(rel[j] - lo) | ((uint32_t)(rel[j] - lo + 1) << 4);
            uint4* c = &tcol[3 * ((size_t)L.xtab + wi)];
            c[0] = make_uint4(coef[0], coef[1], coef[2], coef[3]);
            c[1] = make_uint4(sel[0], sel[1], sel[2], sel[3]);
            c[2] = make_uint4((uint32_t)lo, (uint32_t)hi, ok, 0u);
        }
        for (int r = 0; r < L.h + 2 * VIORB_EDGE; r++) {
            const int dy = reflect(r - VIORB_EDGE, L.h);
            float fy = (float)((dy + 0.5) * sy_ - 0.5);
            int sy = (int)floorf(fy);
            fy -= sy;
            sy = std::max(sy, 0);
            const uint32_t b0 = (uint32_t)(uint16_t)cvRoundF((1.f - fy) * 2048), b1 = (uint32_t)(uint16_t)cvRoundF(fy * 2048);
            const uint32_t s0 = (uint32_t)std::min(sy, P.h - 1), s1 = (uint32_t)std::min(sy + 1, P.h - 1);
            trow[(size_t)L.ytab + r] = make_uint4(s0 | (s1 << 16), b0 << 16, b1 << 16, 0u);
        }
    }
    /* FAST cell groups: the cells the reference visits (:789-806), VIORB_FAST_GROUP neighbours per CTA */
    std::vector<int4> groups;
    int fastRows = 8, fastWork = 64, fastStage = 0;
    for (int l = 0; l < e->nlevels; l++) {
        const LevelGeom& L = g.lv[l];
        const int maxBorderX = L.w - VIORB_FAST_BORDER, maxBorderY = L.h - VIORB_FAST_BORDER;
        for (int i = 0; i < L.nRows; i++) {
            if (VIORB_FAST_BORDER + i * L.hCell >= maxBorderY - 3) continue;
            int nvalid = 0;
            while (nvalid < L.nCols && VIORB_FAST_BORDER + nvalid * L.wCell < maxBorderX - 6) nvalid++;
            /* cells per CTA: as many as fit the kernel's tile (45 quads wide) and work lists (2048 quads) */
            int G = VIORB_FAST_GROUP;
            while (G > 1 && (G * L.wCell > 180 || ((G * L.wCell + 3) / 4) * std::min(L.hCell, 59) > 2048)) G--;
            for (int j = 0; j < nvalid; j += G) {
                const int n = std::min(G, nvalid - j);
                /* the group's detection window (fast_cells_kernel): ww x wh pixels, NQ quads per row */
                const int iniX = VIORB_FAST_BORDER + j * L.wCell, iniY = VIORB_FAST_BORDER + i * L.hCell;
                const int ww = std::min(iniX + n * L.wCell + 6, maxBorderX) - iniX - 6;
                const int wh = std::min(iniY + L.hCell + 6, maxBorderY) - iniY - 6;
                if (ww > 0 && wh > 0) {
                    groups.push_back(make_int4(l, i, j, n));
                    fastWork = std::max(fastWork, ((ww + 3) / 4) * wh);
                    /* staged candidate records (4 bytes each, in the work0 region): at most every other pixel per row and
                     * column of a cell is a 3x3 local maximum */
                    fastStage = std::max(fastStage, 4 * n * ((L.wCell + 1) / 2) * ((wh + 1) / 2));
                }
            }
        }
        fastRows = std::max(fastRows, std::min(L.hCell + 6, VIORB_FAST_TILE_ROWS));
    }
    g.fastTileRows = fastRows;
    g.fastMaxWork = (fastWork + 63) & ~63;
    g.fastPixBytes = (fastRows * VIORB_FAST_TILE_BYTES + std::max(g.fastMaxWork * 2, fastStage) + 127) & ~127;
    if (viorb_fast_prepare(g) != 0) return fail(VIORB_ERR_CUDA, "FAST kernel attribute: %s", cudaGetErrorString(cudaGetLastError()));
    {   /* sort by the byte shift of the group's window inside a 4-byte word of the stored row (fast_cells_kernel<SH>) */
        auto shiftOf = [&](const int4& gr) { return (VIORB_ROI_X0 + VIORB_FAST_BORDER + gr.z * g.lv[gr.x].wCell + 3) & 3; };
        std::stable_sort(groups.begin(), groups.end(), [&](const int4& a, const int4& b) { return shiftOf(a) < shiftOf(b); });
        int cnt[4] = {0, 0, 0, 0};
        for (const int4& gr : groups) cnt[shiftOf(gr)]++;
        e->groupClass[0] = 0;
        for (int i = 0; i < 4; i++) e->groupClass[i + 1] = e->groupClass[i] + cnt[i];
    }
    e->ngroups = (int)groups.size();
    /* what the kernel needs per group, precomputed (two int4 per group, fast_cells_kernel) */
    std::vector<int4> gdesc;
    gdesc.reserve(2 * groups.size());
    for (const int4& gr : groups) {
        const LevelGeom& L = g.lv[gr.x];
        const int i = gr.y, j0 = gr.z, n = gr.w;
        const int maxBorderX = L.w - VIORB_FAST_BORDER, maxBorderY = L.h - VIORB_FAST_BORDER;
        const int iniX = VIORB_FAST_BORDER + j0 * L.wCell, iniY = VIORB_FAST_BORDER + i * L.hCell;
        const int ww = std::min(iniX + n * L.wCell + 6, maxBorderX) - iniX - 6;
        const int wh = std::min(iniY + L.hCell + 6, maxBorderY) - iniY - 6;
        const int NQ = (ww + 3) >> 2;
        const int gstart = VIORB_ROI_X0 + iniX + 3, boxX = (gstart - 3) & ~15, w0 = (gstart - boxX) >> 2;
        const int boxH = std::min(L.hCell + 6, g.fastTileRows);
        if (L.wCell > 255 || NQ > 255 || ww > 65535 || wh > 32767 || boxH > 32767)
            return fail(VIORB_ERR_UNSUPPORTED, "FAST cell %dx%d outside the kernel's envelope", L.wCell, L.hCell);
        gdesc.push_back(make_int4(gr.x | (n << 8) | (L.wCell << 16), ww | (wh << 16), NQ | (w0 << 8) | (boxH << 16),
                                  (int)(0xffffffffu / (unsigned)NQ + 1u)));
        gdesc.push_back(make_int4(boxX, VIORB_EDGE + iniY, L.candBase, (3 + j0 * L.wCell) | ((3 + i * L.hCell) << 12)));
    }
    int rc;
    if ((rc = e->groups.ensure(gdesc.size() + 1))) return rc;
    CU(cudaMemcpyAsync(e->groups.p, gdesc.data(), gdesc.size() * sizeof(int4), cudaMemcpyHostToDevice, e->ctx->stream));
    if ((rc = e->tabCol.ensure(tcol.size() + 1))) return rc;
    if ((rc = e->tabRow.ensure(trow.size() + 1))) return rc;
    CU(cudaMemcpyAsync(e->tabCol.p, tcol.data(), tcol.size() * sizeof(uint4), cudaMemcpyHostToDevice, e->ctx->stream));
    CU(cudaMemcpyAsync(e->tabRow.p, trow.data(), trow.size() * sizeof(uint4), cudaMemcpyHostToDevice, e->ctx->stream));
    CU(cudaStreamSynchronize(e->ctx->stream));
    e->tables.col = e->tabCol.p;
    e->tables.row = e->tabRow.p;
    cudaError_t ce = (cudaError_t)viorb_octree_prepare(nodeCap);
    if (ce != cudaSuccess) return fail(VIORB_ERR_CUDA, "octree kernel attribute: %s", cudaGetErrorString(ce));
    e->rows = rows; e->cols = cols;
    e->geomGen++;                  /* invalidates the captured single-frame graph */
    for (int i = 0; i < VIORB_MAX_LANES; i++) e->lanes[i].allocFrames = 0;
    return VIORB_OK;
}

int ensure_workspace(viorb_extractor* e, int F) {
    const FrameGeom& g = e->geom;
    for (int li = 0; li < e->nlanes; li++) {
        viorb_extractor::Lane& ln = e->lanes[li];
        if (F <= ln.allocFrames) continue;
        int rc;
        if ((rc = ln.pyr.ensure((size_t)F * g.pyrFrameBytes))) return rc;
        if (e->denseBlur && (rc = ln.blur.ensure((size_t)F * g.pyrFrameBytes + 64))) return rc;
        if ((rc = ln.cand.ensure((size_t)F * g.candPerFrame))) return rc;
        if ((rc = ln.nodeOf.ensure((size_t)F * g.candPerFrame))) return rc;
        if ((rc = ln.sel.ensure((size_t)F * g.selPerFrame))) return rc;
        if ((rc = ln.counters.ensure((size_t)F * g.nlevels * 2 + 4))) return rc;
        CU(cudaMemsetAsync(ln.counters.p, 0, ln.counters.n * sizeof(int), e->ctx->stream));
        ln.buf.pyr = ln.pyr.p;
        ln.buf.blur = e->denseBlur ? ln.blur.p : nullptr;
        if (viorb_encode_tma_maps(g, ln.pyr.p, F, &ln.maps)) return fail(VIORB_ERR_CUDA, "cuTensorMapEncodeTiled failed");
        ln.buf.cand = ln.cand.p;
        ln.buf.nodeOf = ln.nodeOf.p;
        ln.buf.sel = ln.sel.p;
        ln.buf.candCount = ln.counters.p;
        ln.buf.selCount = ln.counters.p + (size_t)F * g.nlevels;
        ln.buf.status = ln.counters.p + (size_t)F * g.nlevels * 2;
        ln.allocFrames = F;
    }
    CU(cudaStreamSynchronize(e->ctx->stream));
    e->buf = e->lanes[0].buf;
    return VIORB_OK;
}

/* one device pass over F frames already resident on the device */
int run_pass(viorb_extractor* e, int lane, const uint8_t* d_images, size_t step, size_t frameStride, int F,
             viorb_keypoint* d_kps, uint8_t* d_desc, int cap, int32_t* d_counts) {
    viorb_ctx* c = e->ctx;
    const FrameGeom& g = e->geom;
    viorb_extractor::Lane& ln = e->lanes[lane];
    cudaStream_t st = ln.stream;
    CU(cudaMemsetAsync(ln.buf.candCount, 0, (size_t)F * g.nlevels * sizeof(int), st));
    cudaEvent_t ev[5] = {nullptr, nullptr, nullptr, nullptr, nullptr};
    if (e->profiling)
        for (int i = 0; i < 5; i++) {
            if (e->profPool.empty()) CU(cudaEventCreate(&ev[i]));
            else { ev[i] = e->profPool.back(); e->profPool.pop_back(); }
            e->profEvents.push_back(ev[i]);
        }
    if (e->profiling) CU(cudaEventRecord(ev[0], st));
    /* programmatic dependent launch along the pass's kernel chain.  Default: between the pyramid levels (single-lane pyramid
     * stage 5.4 -> 4.9 ms per 4096 frames); on the stage edges it changes nothing measurable once four lanes overlap, and the
     * waiting CTAs of the next stage would hold shared memory the other lanes could use.  VIORB_PDL=0..3 for A/B runs. */
    static const int pdlEnv = [] { const char* v = getenv("VIORB_PDL"); return v ? atoi(v) : VIORB_PDL_INNER; }();
    const int pdl = e->profiling ? (pdlEnv & VIORB_PDL_INNER) : pdlEnv;
    c->launches += viorb_launch_pyramid(g, e->tables, d_images, step, frameStride, F, ln.buf, st, pdl);
    if (e->profiling) CU(cudaEventRecord(ev[1], st));
    c->launches += viorb_launch_fast(g, ln.maps, e->groups.p, e->groupClass, F, ln.buf, st, pdl);
    if (e->profiling) CU(cudaEventRecord(ev[2], st));
    c->launches += viorb_launch_octree(g, F, ln.buf, e->nodeCap, st, pdl);
    if (e->profiling) CU(cudaEventRecord(ev[3], st));
    /* (sub-batches of 16..64 frames, to keep the blurred levels in the L2 until they are sampled, were measured and are
     * slower than one launch per pass: 6.2 ms -> 7.2 ms (64) / 8.9 ms (32) per 4096 frames) */
    ExtractBuffers db = ln.buf;
    if (F <= 8 && e->describeMode != VIORB_DESCRIBE_LEVELS) db.blur = nullptr;        /* latency path: fused kernel */
    c->launches += viorb_launch_blur(g, 0, F, db, st, pdl);
    c->launches += viorb_launch_describe(g, 0, F, db, d_kps, d_desc, cap, d_counts, st, pdl);
    if (e->profiling) CU(cudaEventRecord(ev[4], st));
    CU(cudaGetLastError());
    e->buf = ln.buf;
    return VIORB_OK;
}

/* lanes start after everything already queued on the context stream ... */
int lanes_fork(viorb_extractor* e) {
    CU(cudaEventRecord(e->evFork, e->ctx->stream));
    for (int li = 0; li < e->nlanes; li++) CU(cudaStreamWaitEvent(e->lanes[li].stream, e->evFork, 0));
    return VIORB_OK;
}

/* ... and the context stream continues after both lanes have drained */
int lanes_join(viorb_extractor* e) {
    for (int li = 0; li < e->nlanes; li++) {
        CU(cudaEventRecord(e->lanes[li].evDone, e->lanes[li].stream));
        CU(cudaStreamWaitEvent(e->ctx->stream, e->lanes[li].evDone, 0));
    }
    return VIORB_OK;
}

int check_status(viorb_extractor* e) {
    int st = 0;
    for (int li = 0; li < e->nlanes; li++) {
        int s1 = 0;
        if (!e->lanes[li].buf.status) continue;
        CU(cudaStreamSynchronize(e->lanes[li].stream));
        CU(cudaMemcpyAsync(&s1, e->lanes[li].buf.status, sizeof(int), cudaMemcpyDeviceToHost, e->ctx->stream));
        CU(cudaStreamSynchronize(e->ctx->stream));
        if (s1) cudaMemsetAsync(e->lanes[li].buf.status, 0, sizeof(int), e->ctx->stream);
        st |= s1;
    }
    CU(cudaStreamSynchronize(e->ctx->stream));
    e->lastOverflow = st;
    if (st) {
        if (st & VIORB_DEV_CAND_OVERFLOW)
            return fail(VIORB_ERR_CAPACITY, "FAST candidate pool overflow (raise it with viorb_extractor_configure cand_div)");
        if (st & VIORB_DEV_OUT_OVERFLOW) return fail(VIORB_ERR_CAPACITY, "more keypoints than the caller's capacity");
        return fail(VIORB_ERR_CAPACITY, "quadtree node pool overflow (device status %d)", st);
    }
    return VIORB_OK;
}

}  // namespace

extern "C" {

const char* viorb_last_error(void) { return g_err; }

int viorb_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) return 0;
    return n;
}

int viorb_ctx_create(int device, void* stream, viorb_ctx** out) {
    if (!out) return fail(VIORB_ERR_INVALID, "out is NULL");
    *out = nullptr;
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n == 0)
        return fail(VIORB_ERR_CUDA, "no CUDA device: %s (libviorb_b200 has no CPU fallback)", cudaGetErrorString(e));
    if (device < 0 || device >= n) return fail(VIORB_ERR_INVALID, "device %d out of range (%d devices)", device, n);
    CU(cudaSetDevice(device));
    cudaDeviceProp prop;
    CU(cudaGetDeviceProperties(&prop, device));
    if (prop.major < 10)
        return fail(VIORB_ERR_CUDA, "device %d is sm_%d%d; this library is built for sm_100a only", device, prop.major, prop.minor);
    viorb_ctx* c = new (std::nothrow) viorb_ctx();
    if (!c) return fail(VIORB_ERR_INVALID, "out of host memory");
    c->device = device;
    c->sms = prop.multiProcessorCount;
    if (stream) { c->stream = (cudaStream_t)stream; c->ownStream = false; }
    else {
        if (cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking) != cudaSuccess) { delete c; return fail(VIORB_ERR_CUDA, "cudaStreamCreate failed"); }
        c->ownStream = true;
    }
    if (cudaStreamCreateWithFlags(&c->h2d, cudaStreamNonBlocking) != cudaSuccess ||
        cudaStreamCreateWithFlags(&c->d2h, cudaStreamNonBlocking) != cudaSuccess) {
        delete c;
        return fail(VIORB_ERR_CUDA, "cudaStreamCreate failed");
    }
    *out = c;
    return VIORB_OK;
}

int viorb_ctx_destroy(viorb_ctx* c) {
    if (!c) return VIORB_OK;
    cudaSetDevice(c->device);
    cudaStreamSynchronize(c->stream);
    c->mq.release(); c->mmap.release(); c->mparts.release(); c->mout.release();
    c->scratchA.release(); c->scratchB.release(); c->scratchI.release(); c->arena.release();
    for (int i = 0; i < 2; i++) {
        if (c->stage[i]) cudaFreeHost(c->stage[i]);
        if (c->stageDone[i]) cudaEventDestroy(c->stageDone[i]);
    }
    for (size_t i = 0; i < c->pool.size(); i++) cudaFree(c->pool[i].second);
    if (c->ownStream) cudaStreamDestroy(c->stream);
    cudaStreamDestroy(c->h2d);
    cudaStreamDestroy(c->d2h);
    delete c;
    return VIORB_OK;
}

int viorb_ctx_synchronize(viorb_ctx* c) {
    if (!c) return fail(VIORB_ERR_INVALID, "ctx is NULL");
    CU(cudaSetDevice(c->device));
    CU(cudaStreamSynchronize(c->stream));
    return VIORB_OK;
}

int64_t viorb_ctx_launch_count(const viorb_ctx* c) { return c ? c->launches : 0; }

int viorb_host_alloc(size_t bytes, void** out) {
    if (!out) return fail(VIORB_ERR_INVALID, "out is NULL");
    CU(cudaHostAlloc(out, bytes, cudaHostAllocDefault));
    return VIORB_OK;
}

int viorb_host_free(void* p) {
    if (p) CU(cudaFreeHost(p));
    return VIORB_OK;
}

/* ---------------------------------------------------------------------------------------------- extractor */
int viorb_extractor_create(viorb_ctx* ctx, int nfeatures, float scale_factor, int nlevels, int ini, int mn,
                           viorb_extractor** out) {
    if (!ctx || !out) return fail(VIORB_ERR_INVALID, "NULL argument");
    *out = nullptr;
    if (nfeatures <= 0 || nlevels < 1 || nlevels > VIORB_MAX_LEVELS || !(scale_factor > 1.0f) || ini < mn || mn < 1 || ini > 127)      /* the byte-domain pre-test of the FAST kernel needs thresholds <= 127 */
        return fail(VIORB_ERR_INVALID, "bad ORB parameters (nfeatures %d, scale %f, levels %d, FAST %d/%d)", nfeatures,
                    scale_factor, nlevels, ini, mn);
    if (scale_factor > 1.5f)
        return fail(VIORB_ERR_UNSUPPORTED, "scaleFactor > 1.5 is outside the pyramid kernel's tile envelope "
                                           "(and 2.0 would take cv::resize's INTER_AREA path); shipped configs use 1.2");
    viorb_extractor* e = new (std::nothrow) viorb_extractor();
    if (!e) return fail(VIORB_ERR_INVALID, "out of host memory");
    e->ctx = ctx;
    e->nfeatures = nfeatures; e->nlevels = nlevels; e->iniTh = ini; e->minTh = mn;
    e->scaleFactor = scale_factor;
    build_tables(e);
    if (ctx_bind(ctx)) { delete e; return VIORB_ERR_CUDA; }
    cudaEventCreateWithFlags(&e->evFork, cudaEventDisableTiming);
    cudaHostAlloc((void**)&e->hostStatus, 64, cudaHostAllocDefault);
    if (getenv("VIORB_LANES")) e->nlanes = std::min(VIORB_MAX_LANES, std::max(2, atoi(getenv("VIORB_LANES"))));   /* the host-buffer path pairs its staging slots with the first lanes */
    for (int i = 0; i < VIORB_MAX_LANES; i++) {
        cudaStreamCreateWithFlags(&e->lanes[i].stream, cudaStreamNonBlocking);
        cudaEventCreateWithFlags(&e->lanes[i].evDone, cudaEventDisableTiming);
    }
    if (const char* v = getenv("VIORB_DESCRIBE"))
        e->describeMode = !strcmp(v, "fused") ? VIORB_DESCRIBE_FUSED : (!strcmp(v, "dense") ? VIORB_DESCRIBE_LEVELS : VIORB_DESCRIBE_AUTO);
    if (getenv("VIORB_SLOTS")) e->nslots = std::min(4, std::max(2, atoi(getenv("VIORB_SLOTS"))));
    e->nlanes = std::max(e->nlanes, e->nslots);
    for (int i = 0; i < 4; i++) {
        cudaEventCreateWithFlags(&e->evIn[i], cudaEventDisableTiming);
        cudaEventCreateWithFlags(&e->evDone[i], cudaEventDisableTiming);
        cudaEventCreateWithFlags(&e->evOut[i], cudaEventDisableTiming);
    }
    *out = e;
    return VIORB_OK;
}

int viorb_extractor_destroy(viorb_extractor* e) {
    if (!e) return VIORB_OK;
    cudaSetDevice(e->ctx->device);
    cudaStreamSynchronize(e->ctx->stream);
    cudaStreamSynchronize(e->ctx->h2d);
    cudaStreamSynchronize(e->ctx->d2h);
    e->tabCol.release(); e->tabRow.release(); e->groups.release();
    for (int i = 0; i < VIORB_MAX_LANES; i++) {
        viorb_extractor::Lane& ln = e->lanes[i];
        if (ln.stream) { cudaStreamSynchronize(ln.stream); cudaStreamDestroy(ln.stream); }
        if (ln.evDone) cudaEventDestroy(ln.evDone);
        ln.pyr.release(); ln.blur.release(); ln.cand.release(); ln.sel.release(); ln.counters.release(); ln.nodeOf.release();
    }
    if (e->evFork) cudaEventDestroy(e->evFork);
    if (e->hostStatus) cudaFreeHost(e->hostStatus);
    if (e->pyrHost) cudaFreeHost(e->pyrHost);
    if (e->frameGraph) cudaGraphExecDestroy(e->frameGraph);
    for (int i = 0; i < 4; i++) {
        e->in[i].release(); e->okps[i].release(); e->odesc[i].release(); e->ocnt[i].release();
        if (e->evIn[i]) cudaEventDestroy(e->evIn[i]);
        if (e->evDone[i]) cudaEventDestroy(e->evDone[i]);
        if (e->evOut[i]) cudaEventDestroy(e->evOut[i]);
    }
    delete e;
    return VIORB_OK;
}

int viorb_extractor_set_gaussian(viorb_extractor* e, int opencv_variant) {
    if (!e) return fail(VIORB_ERR_INVALID, "extractor is NULL");
    if (opencv_variant != VIORB_GAUSSIAN_OPENCV4 && opencv_variant != VIORB_GAUSSIAN_OPENCV24)
        return fail(VIORB_ERR_INVALID, "unknown Gaussian variant %d", opencv_variant);
    e->gaussVariant = opencv_variant;
    e->geom.gaussVariant = opencv_variant;
    e->geomGen++;                  /* the captured single-frame graph holds the kernel choice */
    return VIORB_OK;
}

int viorb_extractor_set_describe_mode(viorb_extractor* e, int mode) {
    if (!e) return fail(VIORB_ERR_INVALID, "extractor is NULL");
    if (mode != VIORB_DESCRIBE_AUTO && mode != VIORB_DESCRIBE_FUSED && mode != VIORB_DESCRIBE_LEVELS)
        return fail(VIORB_ERR_INVALID, "unknown describe mode %d", mode);
    if (mode != e->describeMode) {
        e->describeMode = mode;
        e->rows = e->cols = 0;         /* the workspace (blurred levels) and the captured graph follow the mode */
    }
    return VIORB_OK;
}

int viorb_extractor_set_copy_mode(viorb_extractor* e, int mode) {
    if (!e) return fail(VIORB_ERR_INVALID, "extractor is NULL");
    if (mode != VIORB_COPY_DUPLEX && mode != VIORB_COPY_SERIAL) return fail(VIORB_ERR_INVALID, "unknown copy mode %d", mode);
    e->copyMode = mode;
    return VIORB_OK;
}

int viorb_extractor_configure(viorb_extractor* e, int chunk_frames, int cand_div) {
    if (!e) return fail(VIORB_ERR_INVALID, "extractor is NULL");
    if (chunk_frames > 0) { e->chunk = chunk_frames; e->chunkUser = true; }
    if (cand_div > 0 && cand_div != e->candDiv) { e->candDiv = cand_div; e->rows = e->cols = 0; }
    return VIORB_OK;
}

}  // extern "C" (reopened below)

int viorb_fail(int code, const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    return code;
}
cudaStream_t viorb_ctx_stream(viorb_ctx* c) { return c->stream; }
int viorb_ctx_device(const viorb_ctx* c) { return c->device; }
int viorb_ctx_bind(viorb_ctx* c) { return ctx_bind(c); }
int viorb_ctx_scratch(viorb_ctx* c, size_t bytes, uint8_t** out) {
    int rc = c->arena.ensure(bytes);
    if (rc) return rc;
    *out = c->arena.p;
    return VIORB_OK;
}
void viorb_ctx_add_launches(viorb_ctx* c, int n) { c->launches += n; }

/* pinned staging block `slot` of at least `bytes`; waits until the previous upload from it has left the host */
int viorb_ctx_stage(viorb_ctx* c, int slot, size_t bytes, uint8_t** out) {
    if (c->stagePending[slot]) {
        CU(cudaEventSynchronize(c->stageDone[slot]));
        c->stagePending[slot] = false;
    }
    if (bytes > c->stageBytes[slot]) {
        if (c->stage[slot]) cudaFreeHost(c->stage[slot]);
        c->stage[slot] = nullptr; c->stageBytes[slot] = 0;
        const size_t want = std::max<size_t>(bytes + bytes / 2, 1 << 20);
        CU(cudaHostAlloc((void**)&c->stage[slot], want, cudaHostAllocDefault));
        c->stageBytes[slot] = want;
    }
    if (!c->stageDone[slot]) CU(cudaEventCreateWithFlags(&c->stageDone[slot], cudaEventDisableTiming));
    *out = c->stage[slot];
    return VIORB_OK;
}
int viorb_ctx_stage_mark(viorb_ctx* c, int slot) {
    CU(cudaEventRecord(c->stageDone[slot], c->stream));
    c->stagePending[slot] = true;
    return VIORB_OK;
}

/* device block of at least `bytes` from the context's pool (or a fresh allocation); *got = its real size */
int viorb_ctx_block_get(viorb_ctx* c, size_t bytes, uint8_t** out, size_t* got) {
    int best = -1;
    for (size_t i = 0; i < c->pool.size(); i++)
        if (c->pool[i].first >= bytes && (best < 0 || c->pool[i].first < c->pool[best].first)) best = (int)i;
    if (best >= 0) {
        *out = c->pool[best].second; *got = c->pool[best].first;
        c->pool.erase(c->pool.begin() + best);
        return VIORB_OK;
    }
    const size_t want = (bytes + (bytes >> 2) + 65535) & ~(size_t)65535;
    CU(cudaMalloc((void**)out, want));
    *got = want;
    return VIORB_OK;
}
void viorb_ctx_block_put(viorb_ctx* c, uint8_t* p, size_t bytes) {
    if (!p) return;
    if (c->pool.size() >= 16) {                    /* keep the pool small: drop the smallest block */
        size_t k = 0;
        for (size_t i = 1; i < c->pool.size(); i++) if (c->pool[i].first < c->pool[k].first) k = i;
        cudaFree(c->pool[k].second);
        c->pool.erase(c->pool.begin() + k);
    }
    c->pool.push_back(std::make_pair(bytes, p));
}
viorb_ctx* viorb_extractor_ctx(viorb_extractor* e) { return e->ctx; }

extern "C" {

int viorb_extractor_profile(viorb_extractor* e, int enable) {
    if (!e) return fail(VIORB_ERR_INVALID, "extractor is NULL");
    e->profiling = enable != 0;
    return VIORB_OK;
}

int viorb_extractor_stage_ms(viorb_extractor* e, float ms[4], int* passes) {
    if (!e || !ms) return fail(VIORB_ERR_INVALID, "NULL argument");
    int rc;
    if ((rc = ctx_bind(e->ctx))) return rc;
    CU(cudaStreamSynchronize(e->ctx->stream));
    for (int i = 0; i < 4; i++) ms[i] = 0.f;
    const int np = (int)e->profEvents.size() / 5;
    for (int p = 0; p < np; p++)
        for (int i = 0; i < 4; i++) {
            float t = 0.f;
            CU(cudaEventElapsedTime(&t, e->profEvents[5 * p + i], e->profEvents[5 * p + i + 1]));
            ms[i] += t;
        }
    for (cudaEvent_t ev : e->profEvents) e->profPool.push_back(ev);
    e->profEvents.clear();
    if (passes) *passes = np;
    return VIORB_OK;
}

int viorb_extractor_tables(const viorb_extractor* e, int* nlevels, float* scale, float* inv_scale, float* sigma2,
                           float* inv_sigma2, int* fpl) {
    if (!e) return fail(VIORB_ERR_INVALID, "extractor is NULL");
    if (nlevels) *nlevels = e->nlevels;
    for (int i = 0; i < e->nlevels; i++) {
        if (scale) scale[i] = e->scale[i];
        if (inv_scale) inv_scale[i] = e->invScale[i];
        if (sigma2) sigma2[i] = e->sigma2[i];
        if (inv_sigma2) inv_sigma2[i] = e->invSigma2[i];
        if (fpl) fpl[i] = e->quota[i];
    }
    return VIORB_OK;
}

int viorb_extract_batch_device(viorb_extractor* e, const uint8_t* d_images, int B, int rows, int cols, size_t step,
                               size_t frame_stride, viorb_keypoint* d_kps, uint8_t* d_desc, int cap, int32_t* d_counts) {
    if (!e || !d_images || !d_kps || !d_desc || !d_counts) return fail(VIORB_ERR_INVALID, "NULL argument");
    if (B <= 0 || rows <= 0 || cols <= 0 || cap <= 0 || step < (size_t)cols) return fail(VIORB_ERR_INVALID, "bad shape");
    int rc;
    if ((rc = ctx_bind(e->ctx))) return rc;
    if ((rc = build_geometry(e, rows, cols))) return rc;
    const int F = std::min(default_pass_frames(e, rows, cols), B);
    if ((rc = ensure_workspace(e, F))) return rc;
    if ((rc = lanes_fork(e))) return rc;
    int k = 0;
    for (int b0 = 0; b0 < B; b0 += F, k++) {
        const int f = std::min(F, B - b0);
        if ((rc = run_pass(e, e->profiling ? 0 : k % e->nlanes, d_images + (size_t)b0 * frame_stride, step, frame_stride, f,
                           d_kps + (size_t)b0 * cap, d_desc + (size_t)b0 * cap * 32, cap, d_counts + b0)))
            return rc;
        e->residentFirst = b0; e->residentCount = f;
    }
    return lanes_join(e);
}

int viorb_extractor_check(viorb_extractor* e) {
    if (!e) return fail(VIORB_ERR_INVALID, "extractor is NULL");
    int rc;
    if ((rc = ctx_bind(e->ctx))) return rc;
    CU(cudaStreamSynchronize(e->ctx->stream));
    return check_status(e);
}

static int extract_batch_once(viorb_extractor* e, const uint8_t* images, int B, int rows, int cols, size_t step,
                              size_t frame_stride, viorb_keypoint* kps, uint8_t* desc, int cap, int32_t* counts);

int viorb_extract_batch(viorb_extractor* e, const uint8_t* images, int B, int rows, int cols, size_t step,
                        size_t frame_stride, viorb_keypoint* kps, uint8_t* desc, int cap, int32_t* counts) {
    if (!e) return fail(VIORB_ERR_INVALID, "NULL argument");
    for (;;) {
        const int rc = extract_batch_once(e, images, B, rows, cols, step, frame_stride, kps, desc, cap, counts);
        /* a corner-dense frame overflowed the candidate pool: grow it and run the batch again */
        if (rc == VIORB_ERR_CAPACITY && e->lastOverflow == VIORB_DEV_CAND_OVERFLOW && e->candDiv > 4) {
            e->candDiv /= 2;
            e->rows = e->cols = 0;
            continue;
        }
        return rc;
    }
}

static int extract_batch_once(viorb_extractor* e, const uint8_t* images, int B, int rows, int cols, size_t step,
                              size_t frame_stride, viorb_keypoint* kps, uint8_t* desc, int cap, int32_t* counts) {
    if (!e || !kps || !desc || !counts) return fail(VIORB_ERR_INVALID, "NULL argument");
    if (B <= 0 || !images || rows <= 0 || cols <= 0) {           /* empty input: silent return (:1046-1047) */
        for (int b = 0; b < B; b++) counts[b] = 0;
        return VIORB_OK;
    }
    if (cap <= 0 || step < (size_t)cols) return fail(VIORB_ERR_INVALID, "bad shape");
    int rc;
    viorb_ctx* c = e->ctx;
    if ((rc = ctx_bind(c))) return rc;
    if ((rc = build_geometry(e, rows, cols))) return rc;
    /* frames per pass of the copy/compute pipeline: about a twelfth of the call, so that filling and draining the
     * pipeline (first copy in, last pass, last copy out) stays small for short batches; 32..128 (measured:
     * tools/small_batch_sweep.sh) */
    int passFrames = default_pass_frames(e, rows, cols);
    if (!e->chunkUser) {
        /* this path is bound by the input copies (55 GB/s of PCIe carry 150 k EuRoC frames/s, the kernels do 220 k), so what
         * counts is how soon the first pass can start and how short the last one is: three eighths of the device-resident
         * pass size (48 EuRoC frames; 4096 frames per call: 128 -> 138.6 k, 96 -> 144.8 k, 64 -> 145.0 k, 48 -> 146.2 k,
         * 32 -> 141.4 k frames/s, the copies alone allow 150.8 k) */
        passFrames = std::max(1, (passFrames * 3 + 4) / 8);
        passFrames = std::min(passFrames, std::max(std::min(32, passFrames), (B / 12 + 15) / 16 * 16));
    }
    const int F = std::min(passFrames, B);
    if ((rc = ensure_workspace(e, F))) return rc;
    const size_t inFrame = (size_t)rows * cols;        /* device copy is packed */
    const int nslots = B <= F ? 1 : std::min(e->nslots, (B + F - 1) / F);
    for (int s = 0; s < nslots; s++) {
        if ((rc = e->in[s].ensure((size_t)F * inFrame))) return rc;
        if ((rc = e->okps[s].ensure((size_t)F * cap))) return rc;
        if ((rc = e->odesc[s].ensure((size_t)F * cap * 32))) return rc;
        if ((rc = e->ocnt[s].ensure(F))) return rc;
    }
    CU(cudaStreamSynchronize(c->stream));
    if (B <= F && e->hostStatus && !e->profiling) {
        /* one pass (the per-frame call of Frame::ExtractORB): copy in, kernels, copy out and the device status word
         * on a single stream, one synchronisation -- no cross-stream events on the latency path */
        viorb_extractor::Lane& ln = e->lanes[0];
        cudaStream_t ls = ln.stream;
        if (step == (size_t)cols && frame_stride == inFrame) {
            CU(cudaMemcpyAsync(e->in[0].p, images, (size_t)B * inFrame, cudaMemcpyHostToDevice, ls));
        } else if (frame_stride == step * rows) {
            CU(cudaMemcpy2DAsync(e->in[0].p, cols, images, step, cols, (size_t)rows * B, cudaMemcpyHostToDevice, ls));
        } else {
            for (int i = 0; i < B; i++)
                CU(cudaMemcpy2DAsync(e->in[0].p + (size_t)i * inFrame, cols, images + (size_t)i * frame_stride, step, cols, rows,
                                     cudaMemcpyHostToDevice, ls));
        }
        if (B == 1 && !getenv("VIORB_NO_GRAPH")) {
            const void* key[4] = {e->in[0].p, e->okps[0].p, e->odesc[0].p, ln.buf.pyr};
            if (!e->frameGraph || memcmp(key, e->graphKey, sizeof(key)) != 0 || e->graphCap != cap || e->graphGen != e->geomGen) {
                if (e->frameGraph) { cudaGraphExecDestroy(e->frameGraph); e->frameGraph = nullptr; }
                cudaGraph_t graph = nullptr;
                CU(cudaStreamBeginCapture(ls, cudaStreamCaptureModeThreadLocal));
                const long long before = c->launches;
                rc = run_pass(e, 0, e->in[0].p, cols, inFrame, 1, e->okps[0].p, e->odesc[0].p, cap, e->ocnt[0].p);
                e->graphLaunches = (int)(c->launches - before);
                c->launches = before;
                cudaError_t ce = cudaStreamEndCapture(ls, &graph);
                if (rc) { if (graph) cudaGraphDestroy(graph); return rc; }
                if (ce != cudaSuccess) return fail(VIORB_ERR_CUDA, "graph capture failed: %s", cudaGetErrorString(ce));
                ce = cudaGraphInstantiate(&e->frameGraph, graph, 0);
                cudaGraphDestroy(graph);
                if (ce != cudaSuccess) { e->frameGraph = nullptr; return fail(VIORB_ERR_CUDA, "graph instantiate failed: %s", cudaGetErrorString(ce)); }
                memcpy(e->graphKey, key, sizeof(key));
                e->graphCap = cap; e->graphGen = e->geomGen;
            }
            CU(cudaGraphLaunch(e->frameGraph, ls));
            c->launches += e->graphLaunches;
            e->buf = ln.buf;
        } else if ((rc = run_pass(e, 0, e->in[0].p, cols, inFrame, B, e->okps[0].p, e->odesc[0].p, cap, e->ocnt[0].p))) {
            return rc;
        }
        CU(cudaMemcpyAsync(kps, e->okps[0].p, (size_t)B * cap * sizeof(viorb_keypoint), cudaMemcpyDeviceToHost, ls));
        CU(cudaMemcpyAsync(desc, e->odesc[0].p, (size_t)B * cap * 32, cudaMemcpyDeviceToHost, ls));
        CU(cudaMemcpyAsync(counts, e->ocnt[0].p, (size_t)B * sizeof(int32_t), cudaMemcpyDeviceToHost, ls));
        CU(cudaMemcpyAsync(e->hostStatus, ln.buf.status, sizeof(int), cudaMemcpyDeviceToHost, ls));
        CU(cudaStreamSynchronize(ls));
        e->residentFirst = 0; e->residentCount = B;
        if (*e->hostStatus == 0) { e->lastOverflow = 0; return VIORB_OK; }
        return check_status(e);          /* rare: reads, reports and clears the device status */
    }
    /* pipeline: H2D(chunks k+1, k+2) || compute(chunk k) || D2H(chunk k-1), `nslots` staging slots.
     * Copy mode 1 (viorb_extractor_set_copy_mode) puts the output copies on the input-copy stream, two passes behind the
     * input they follow, so the two directions never run at the same time: on hosts where concurrent device-to-host
     * writes slow the host-to-device reads down by more than they take on their own (measured on the 4- and 8-GPU boxes
     * of this pool with every GPU copying: 6.8 ms in alone, 9.8 ms with 19 % as many bytes going out beside it), the
     * serial order is the faster one. */
    const bool serial = e->copyMode == 1 && nslots >= 4;
    cudaStream_t outStream = serial ? c->h2d : c->d2h;
    auto copy_out = [&](int kk) -> int {
        const int ss = kk % nslots;
        const int bb = kk * F, ff = std::min(F, B - bb);
        CU(cudaStreamWaitEvent(outStream, e->evDone[ss], 0));
        CU(cudaMemcpyAsync(kps + (size_t)bb * cap, e->okps[ss].p, (size_t)ff * cap * sizeof(viorb_keypoint), cudaMemcpyDeviceToHost, outStream));
        CU(cudaMemcpyAsync(desc + (size_t)bb * cap * 32, e->odesc[ss].p, (size_t)ff * cap * 32, cudaMemcpyDeviceToHost, outStream));
        CU(cudaMemcpyAsync(counts + bb, e->ocnt[ss].p, (size_t)ff * sizeof(int32_t), cudaMemcpyDeviceToHost, outStream));
        CU(cudaEventRecord(e->evOut[ss], outStream));
        return VIORB_OK;
    };
    int k = 0;
    for (int b0 = 0; b0 < B; b0 += F, k++) {
        const int s = k % nslots;
        const int f = std::min(F, B - b0);
        if (k >= nslots) {
            CU(cudaStreamWaitEvent(c->h2d, e->evDone[s], 0));    /* slot input free once its compute finished */
        }
        if (step == (size_t)cols && frame_stride == inFrame) {       /* packed frames: one flat copy */
            CU(cudaMemcpyAsync(e->in[s].p, images + (size_t)b0 * frame_stride, (size_t)f * inFrame, cudaMemcpyHostToDevice, c->h2d));
        } else if (frame_stride == step * rows) {
            CU(cudaMemcpy2DAsync(e->in[s].p, cols, images + (size_t)b0 * frame_stride, step, cols, (size_t)rows * f,
                                 cudaMemcpyHostToDevice, c->h2d));
        } else {      /* frames are not back to back: one copy per frame */
            for (int i = 0; i < f; i++)
                CU(cudaMemcpy2DAsync(e->in[s].p + (size_t)i * inFrame, cols, images + (size_t)(b0 + i) * frame_stride, step,
                                     cols, rows, cudaMemcpyHostToDevice, c->h2d));
        }
        CU(cudaEventRecord(e->evIn[s], c->h2d));
        if (serial && k >= 2 && (rc = copy_out(k - 2))) return rc;      /* behind the input of pass k on the same stream */
        cudaStream_t ls = e->lanes[s].stream;
        CU(cudaStreamWaitEvent(ls, e->evIn[s], 0));
        if (k >= nslots) CU(cudaStreamWaitEvent(ls, e->evOut[s], 0));     /* slot outputs free once copied out */
        if ((rc = run_pass(e, s, e->in[s].p, cols, inFrame, f, e->okps[s].p, e->odesc[s].p, cap, e->ocnt[s].p))) return rc;
        CU(cudaEventRecord(e->evDone[s], ls));
        if (!serial && (rc = copy_out(k))) return rc;
        e->residentFirst = b0; e->residentCount = f;
    }
    if (serial)
        for (int kk = std::max(k - 2, 0); kk < k; kk++)
            if ((rc = copy_out(kk))) return rc;
    CU(cudaStreamSynchronize(outStream));
    return check_status(e);
}

int viorb_extract(viorb_extractor* e, const uint8_t* image, int rows, int cols, size_t step, viorb_keypoint* kps,
                  uint8_t* desc, int cap, int* n) {
    if (!n) return fail(VIORB_ERR_INVALID, "n is NULL");
    *n = 0;
    if (!image || rows <= 0 || cols <= 0) return VIORB_OK;      /* empty image: silent return (:1046-1047) */
    int32_t cnt = 0;
    int rc = viorb_extract_batch(e, image, 1, rows, cols, step, step * rows, kps, desc, cap, &cnt);
    if (rc) return rc;
    *n = cnt;
    return VIORB_OK;
}

int viorb_extractor_pyramid_info(const viorb_extractor* e, int level, int* w, int* h) {
    if (!e || level < 0 || level >= e->nlevels || !e->rows) return fail(VIORB_ERR_INVALID, "no pyramid (level %d)", level);
    if (w) *w = e->geom.lv[level].w;
    if (h) *h = e->geom.lv[level].h;
    return VIORB_OK;
}

int viorb_extractor_resident(const viorb_extractor* e, int* first, int* count) {
    if (!e) return fail(VIORB_ERR_INVALID, "extractor is NULL");
    if (first) *first = e->residentFirst;
    if (count) *count = e->residentCount;
    return VIORB_OK;
}

int viorb_extractor_pyramid_device(const viorb_extractor* e, int frame, int level, const uint8_t** d_roi, size_t* d_step) {
    if (!e || level < 0 || level >= e->nlevels || !e->rows || frame < 0 || frame >= e->residentCount)
        return fail(VIORB_ERR_INVALID, "frame %d / level %d not resident", frame, level);
    const LevelGeom& L = e->geom.lv[level];
    if (d_roi) *d_roi = e->buf.pyr + (size_t)frame * e->geom.pyrFrameBytes + L.pyrOff + (size_t)VIORB_EDGE * L.step + VIORB_ROI_X0;
    if (d_step) *d_step = L.step;
    return VIORB_OK;
}

int viorb_extractor_pyramid_download(viorb_extractor* e, int frame, int level, uint8_t* dst, size_t dst_step) {
    if (!e || !dst || level < 0 || level >= e->nlevels || !e->rows || frame < 0 || frame >= e->residentCount)
        return fail(VIORB_ERR_INVALID, "frame %d / level %d not resident", frame, level);
    int rc;
    if ((rc = ctx_bind(e->ctx))) return rc;
    const LevelGeom& L = e->geom.lv[level];
    const uint8_t* src = e->buf.pyr + (size_t)frame * e->geom.pyrFrameBytes + L.pyrOff + (VIORB_ROI_X0 - VIORB_EDGE);
    CU(cudaMemcpy2DAsync(dst, dst_step, src, L.step, L.w + 2 * VIORB_EDGE, L.h + 2 * VIORB_EDGE, cudaMemcpyDeviceToHost,
                         e->ctx->stream));
    CU(cudaStreamSynchronize(e->ctx->stream));
    return VIORB_OK;
}

int viorb_extractor_pyramid_download_all(viorb_extractor* e, int frame, uint8_t* const* dst, const size_t* dst_step, int nlevels) {
    if (!e || !dst || !dst_step || nlevels != e->nlevels || !e->rows || frame < 0 || frame >= e->residentCount)
        return fail(VIORB_ERR_INVALID, "frame %d not resident or bad level count", frame);
    int rc;
    if ((rc = ctx_bind(e->ctx))) return rc;
    const size_t bytes = (size_t)e->geom.pyrFrameBytes;
    if (bytes > e->pyrHostBytes) {
        if (e->pyrHost) cudaFreeHost(e->pyrHost);
        e->pyrHost = nullptr; e->pyrHostBytes = 0;
        CU(cudaHostAlloc((void**)&e->pyrHost, bytes, cudaHostAllocDefault));
        e->pyrHostBytes = bytes;
    }
    CU(cudaMemcpyAsync(e->pyrHost, e->buf.pyr + (size_t)frame * bytes, bytes, cudaMemcpyDeviceToHost, e->ctx->stream));
    CU(cudaStreamSynchronize(e->ctx->stream));
    for (int l = 0; l < nlevels; l++) {
        const LevelGeom& L = e->geom.lv[l];
        if (!dst[l]) continue;
        const uint8_t* src = e->pyrHost + L.pyrOff + (VIORB_ROI_X0 - VIORB_EDGE);
        const size_t wbytes = (size_t)L.w + 2 * VIORB_EDGE;
        for (int r = 0; r < L.h + 2 * VIORB_EDGE; r++) memcpy(dst[l] + (size_t)r * dst_step[l], src + (size_t)r * L.step, wbytes);
    }
    return VIORB_OK;
}

static int download_packed(viorb_extractor* e, const uint32_t* d_src, const int* d_count, int capEntries, int32_t* xys,
                           int cap, int* n) {
    int cnt = 0;
    CU(cudaMemcpyAsync(&cnt, d_count, sizeof(int), cudaMemcpyDeviceToHost, e->ctx->stream));
    CU(cudaStreamSynchronize(e->ctx->stream));
    cnt = std::min(cnt, capEntries);
    std::vector<uint32_t> tmp(std::max(cnt, 1));
    CU(cudaMemcpyAsync(tmp.data(), d_src, (size_t)cnt * 4, cudaMemcpyDeviceToHost, e->ctx->stream));
    CU(cudaStreamSynchronize(e->ctx->stream));
    for (int i = 0; i < cnt && i < cap; i++) {
        xys[3 * i] = tmp[i] & 0xfff;
        xys[3 * i + 1] = (tmp[i] >> 12) & 0xfff;
        xys[3 * i + 2] = tmp[i] >> 24;
    }
    *n = cnt;
    return VIORB_OK;
}

int viorb_extractor_debug_candidates(viorb_extractor* e, int frame, int level, int32_t* xys, int cap, int* n) {
    if (!e || !n || level < 0 || level >= e->nlevels || !e->rows || frame < 0 || frame >= e->residentCount)
        return fail(VIORB_ERR_INVALID, "frame %d / level %d not resident", frame, level);
    int rc;
    if ((rc = ctx_bind(e->ctx))) return rc;
    const LevelGeom& L = e->geom.lv[level];
    rc = download_packed(e, e->buf.cand + (size_t)frame * e->geom.candPerFrame + L.candBase,
                         e->buf.candCount + frame * e->nlevels + level, L.candCap, xys, cap, n);
    if (rc) return rc;
    for (int i = 0; i < *n && i < cap; i++) { xys[3 * i] += VIORB_FAST_BORDER; xys[3 * i + 1] += VIORB_FAST_BORDER; }
    return VIORB_OK;
}

int viorb_extractor_debug_selected(viorb_extractor* e, int frame, int level, int32_t* xys, int cap, int* n) {
    if (!e || !n || level < 0 || level >= e->nlevels || !e->rows || frame < 0 || frame >= e->residentCount)
        return fail(VIORB_ERR_INVALID, "frame %d / level %d not resident", frame, level);
    int rc;
    if ((rc = ctx_bind(e->ctx))) return rc;
    const LevelGeom& L = e->geom.lv[level];
    return download_packed(e, e->buf.sel + (size_t)frame * e->geom.selPerFrame + L.selBase,
                           e->buf.selCount + frame * e->nlevels + level, L.selCap, xys, cap, n);
}

/* ---------------------------------------------------------------------------------------------- matcher */
int viorb_debug_orientation(viorb_ctx* c, const int32_t* m01, const int32_t* m10, int64_t n, float* deg) {
    if (!c || !m01 || !m10 || !deg || n < 0 || n > (1ll << 28)) return fail(VIORB_ERR_INVALID, "bad argument");
    if (n == 0) return VIORB_OK;
    int rc;
    if ((rc = ctx_bind(c))) return rc;
    if ((rc = c->scratchA.ensure((size_t)n * 8)) || (rc = c->scratchB.ensure((size_t)n * 4))) return rc;
    int* d = reinterpret_cast<int*>(c->scratchA.p);
    float* o = reinterpret_cast<float*>(c->scratchB.p);
    CU(cudaMemcpyAsync(d, m01, (size_t)n * 4, cudaMemcpyHostToDevice, c->stream));
    CU(cudaMemcpyAsync(d + n, m10, (size_t)n * 4, cudaMemcpyHostToDevice, c->stream));
    c->launches += viorb_launch_orientation_sweep(d, d + n, n, o, c->sms, c->stream);
    CU(cudaMemcpyAsync(deg, o, (size_t)n * 4, cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    return VIORB_OK;
}

int viorb_debug_steering(viorb_ctx* c, uint32_t first_bits, int64_t n, float* sin_out, float* cos_out) {
    if (!c || !sin_out || !cos_out || n < 0 || n > (1ll << 28)) return fail(VIORB_ERR_INVALID, "bad argument");
    if (n == 0) return VIORB_OK;
    int rc;
    if ((rc = ctx_bind(c))) return rc;
    if ((rc = c->scratchA.ensure((size_t)n * 8))) return rc;
    float* d = reinterpret_cast<float*>(c->scratchA.p);
    c->launches += viorb_launch_steering_sweep(first_bits, n, d, d + n, c->sms, c->stream);
    CU(cudaMemcpyAsync(sin_out, d, (size_t)n * 4, cudaMemcpyDeviceToHost, c->stream));
    CU(cudaMemcpyAsync(cos_out, d + n, (size_t)n * 4, cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    return VIORB_OK;
}

int viorb_descriptor_distance(viorb_ctx* c, const uint8_t* a, const uint8_t* b, int n, int32_t* dist) {
    if (!c || !a || !b || !dist || n < 0) return fail(VIORB_ERR_INVALID, "bad argument");
    if (n == 0) return VIORB_OK;
    int rc;
    if ((rc = ctx_bind(c))) return rc;
    if ((rc = c->scratchA.ensure((size_t)n * 32)) || (rc = c->scratchB.ensure((size_t)n * 32)) || (rc = c->scratchI.ensure(n))) return rc;
    CU(cudaMemcpyAsync(c->scratchA.p, a, (size_t)n * 32, cudaMemcpyHostToDevice, c->stream));
    CU(cudaMemcpyAsync(c->scratchB.p, b, (size_t)n * 32, cudaMemcpyHostToDevice, c->stream));
    c->launches += viorb_launch_descriptor_distance(c->scratchA.p, c->scratchB.p, n, c->scratchI.p, c->stream);
    CU(cudaMemcpyAsync(dist, c->scratchI.p, (size_t)n * 4, cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    return VIORB_OK;
}

int viorb_hamming_top2_device(viorb_ctx* c, const uint8_t* d_q, int Q, const uint8_t* d_map, int64_t M, int64_t base,
                              viorb_top2* d_out) {
    if (!c || !d_q || !d_out || Q < 0 || M < 0 || (M > 0 && !d_map)) return fail(VIORB_ERR_INVALID, "bad argument");
    if (base + M >= (1ll << 31)) return fail(VIORB_ERR_UNSUPPORTED, "map indices must fit in int32");
    if (Q == 0) return VIORB_OK;
    int rc;
    if ((rc = ctx_bind(c))) return rc;
    const int ns = viorb_top2_slices(Q, M, c->sms);
    if ((rc = c->mparts.ensure((size_t)ns * Q))) return rc;
    c->launches += viorb_launch_hamming_top2(d_q, Q, d_map, M, base, c->mparts.p, ns, d_out, c->stream);
    CU(cudaGetLastError());
    return VIORB_OK;
}

int viorb_hamming_top2(viorb_ctx* c, const uint8_t* q, int Q, const uint8_t* map, int64_t M, int64_t base, viorb_top2* out) {
    if (!c || !q || !out || Q < 0 || M < 0 || (M > 0 && !map)) return fail(VIORB_ERR_INVALID, "bad argument");
    if (Q == 0) return VIORB_OK;
    int rc;
    if ((rc = ctx_bind(c))) return rc;
    if ((rc = c->mq.ensure((size_t)Q * 32)) || (rc = c->mmap.ensure((size_t)std::max<int64_t>(M, 1) * 32)) || (rc = c->mout.ensure(Q))) return rc;
    CU(cudaMemcpyAsync(c->mq.p, q, (size_t)Q * 32, cudaMemcpyHostToDevice, c->stream));
    if (M) CU(cudaMemcpyAsync(c->mmap.p, map, (size_t)M * 32, cudaMemcpyHostToDevice, c->stream));
    if ((rc = viorb_hamming_top2_device(c, c->mq.p, Q, c->mmap.p, M, base, c->mout.p))) return rc;
    CU(cudaMemcpyAsync(out, c->mout.p, (size_t)Q * sizeof(viorb_top2), cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    return VIORB_OK;
}

int viorb_top2_merge_device(viorb_ctx* c, const viorb_top2* d_parts, int nparts, int Q, viorb_top2* d_out) {
    if (!c || !d_parts || !d_out || nparts < 1 || Q < 0) return fail(VIORB_ERR_INVALID, "bad argument");
    int rc;
    if ((rc = ctx_bind(c))) return rc;
    c->launches += viorb_launch_top2_merge(d_parts, nparts, Q, d_out, c->stream);
    CU(cudaGetLastError());
    return VIORB_OK;
}

}  // extern "C"
