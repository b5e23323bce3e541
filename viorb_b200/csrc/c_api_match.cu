/*
 * c_api_match.cu -- extern "C" entry points of the windowed matchers and the stereo matcher
 * (include/viorb_gpu.h): host<->device marshalling around search_kernels.cu.  No host compute.
 */
#include <cuda_runtime.h>

#include <algorithm>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <new>
#include <vector>

#include "matcher_kernels.cuh"
#include "viorb_gpu.h"
#include "viorb_internal.h"

/* launchers of search_kernels.cu */
int viorb_launch_stereo(const StereoParams& p, const viorb_keypoint* d_kl, const uint8_t* d_dl,
                        const viorb_keypoint* d_kr, const uint8_t* d_dr, int* d_scratch, float* d_uRight,
                        float* d_depth, int* d_sad, cudaStream_t s);
int viorb_launch_grid_build(const viorb_keypoint* d_kps, int n, float minX, float minY, float invW, float invH,
                            int* d_cellOf, int* d_cellStart, int* d_cellItems, cudaStream_t s);
int viorb_launch_features_in_area(const FrameIndexDev& fi, float x, float y, float r, int minLevel, int maxLevel,
                                  unsigned long long* d_keys, int* d_count, cudaStream_t s);
int viorb_launch_search_local(const FrameIndexDev& fi, const float* projX, const float* projY, const float* projXR,
                              const int* predLevel, const float* viewCos, const uint8_t* valid, const int* nobs,
                              const uint8_t* mpDesc, int nmp, float th, float nnratio, const int* d_obs0, int* d_obs,
                              int* d_scratch, int* d_match, int* d_nmatches, cudaStream_t s);
int viorb_search_scratch_ints(int nq, int nf);
int viorb_launch_search_frame(const FrameIndexDev& fi, const float* u, const float* v, const float* invz,
                              const int* lastOctave, const float* lastAngle, const uint8_t* valid, const int* nobs,
                              const uint8_t* mpDesc, int nlast, float th, float mbf, int mode, int checkOri, int thHigh,
                              const int* d_obs0, int* d_obs, int* d_scratch, int* d_match, int* d_nmatches, cudaStream_t s);
int viorb_launch_triangulation(const viorb_keypoint* k1, const uint8_t* d1, const float* ur1, const uint8_t* mp1, int n1,
                               const viorb_keypoint* k2, const uint8_t* d2, const float* ur2, const uint8_t* mp2, int n2,
                               const int* nodeId1, const int* nodePtr1, const int* idx1, int nn1, int nentries1,
                               const int* nodeId2, const int* nodePtr2, const int* idx2, int nn2, const float* F12,
                               float ex, float ey, const float* scale2, const float* sigma2, int nlevels, int onlyStereo,
                               int checkOri, int* d_matches12, int* d_nmatches, cudaStream_t s);

int viorb_launch_distinctive(const uint8_t* d_desc, const int* d_ptr, int nmp, int* d_best, int* d_bestMedian,
                             cudaStream_t s);

int viorb_launch_search_bow(int mode, const viorb_keypoint* k1, const uint8_t* d1, const uint8_t* valid1, int n1,
                            const viorb_keypoint* k2, const uint8_t* d2, const uint8_t* valid2, int n2, const int* nodeId1,
                            const int* nodePtr1, const int* idx1, int nn1, const int* nodeId2, const int* nodePtr2,
                            const int* idx2, int nn2, float nnratio, int checkOri, int* d_taken, int* d_match,
                            int* d_nmatches, cudaStream_t s);
int viorb_launch_search_init(const FrameIndexDev& f2, const viorb_keypoint* k1, const uint8_t* d1, int n1, float* d_prev,
                             float window, float nnratio, int checkOri, unsigned long long* d_entries, long long cap,
                             int* d_start, int* d_count, unsigned long long* d_cursor, int* d_overflow, int* d_matchedDist,
                             int* d_matches21, int* d_binOf, int* d_matches12, int* d_nmatches, cudaStream_t s);

int viorb_launch_search_window(const FrameIndexDev& fi, const float* u, const float* v, const float* ur, const int* level,
                               const uint8_t* valid, const uint8_t* desc, int n, float th, int thDist, const float* invSigma2,
                               int nlevels, int* d_bestIdx, int* d_bestDist, cudaStream_t s);
int viorb_launch_sim3_agree(const int* d_m1, const int* d_m2, int n1, int* d_match12, int* d_nfound, cudaStream_t s);

namespace {

/* Bump allocator over one device block with a pinned host mirror of the same layout.  A call take()s all its arrays,
 * put()s its inputs into the mirror, flush()es them with ONE host-to-device copy, launches its kernels, registers its
 * results with get() and finish()es: ONE device-to-host copy into the mirror, one stream synchronisation, then plain
 * memcpy into the caller's buffers.  (A dozen small pageable copies per search cost more than the search itself.) */
struct Arena {
    uint8_t* base = nullptr;
    uint8_t* host = nullptr;
    size_t cap = 0, hostCap = 0, off = 0;
    int slot = 0;
    bool overflow = false;
    size_t dirtyLo = (size_t)-1, dirtyHi = 0;
    struct Pending { void* user; size_t off, bytes; };
    Pending outs[12];
    int nouts = 0;
    /* returns NULL (and latches `overflow`) when the request does not fit: callers check once after their takes */
    template <typename T>
    T* take(size_t n) {
        off = (off + 255) & ~(size_t)255;
        if (off + n * sizeof(T) > cap) { overflow = true; return nullptr; }
        T* p = reinterpret_cast<T*>(base + off);
        off += n * sizeof(T);
        return p;
    }
    bool ok() const { return !overflow; }
    template <typename T>
    void put(T* dst, const T* src, size_t n) {
        if (n == 0 || !dst || !src) return;
        const size_t o = (size_t)(reinterpret_cast<uint8_t*>(dst) - base), bytes = n * sizeof(T);
        if (o + bytes > hostCap) { overflow = true; return; }
        memcpy(host + o, src, bytes);
        dirtyLo = std::min(dirtyLo, o);
        dirtyHi = std::max(dirtyHi, o + bytes);
    }
    int flush(viorb_ctx* c) {
        if (overflow) return viorb_fail(VIORB_ERR_CAPACITY, "scratch arena under-sized (%zu B)", cap);
        if (dirtyHi > dirtyLo) {
            VCU(cudaMemcpyAsync(base + dirtyLo, host + dirtyLo, dirtyHi - dirtyLo, cudaMemcpyHostToDevice, viorb_ctx_stream(c)));
            const int rc = viorb_ctx_stage_mark(c, slot);
            if (rc) return rc;
        }
        dirtyLo = (size_t)-1; dirtyHi = 0;
        return VIORB_OK;
    }
    void get(void* user, const void* dsrc, size_t bytes) {
        if (bytes == 0 || !user) return;
        const size_t o = (size_t)(reinterpret_cast<const uint8_t*>(dsrc) - base);
        if (nouts >= 12 || o + bytes > hostCap) { overflow = true; return; }
        outs[nouts].user = user; outs[nouts].off = o; outs[nouts].bytes = bytes;
        nouts++;
    }
    int finish(viorb_ctx* c) {
        if (overflow) return viorb_fail(VIORB_ERR_CAPACITY, "scratch arena under-sized (%zu B)", cap);
        cudaStream_t s = viorb_ctx_stream(c);
        VCU(cudaGetLastError());
        size_t lo = (size_t)-1, hi = 0, sum = 0;
        for (int i = 0; i < nouts; i++) { lo = std::min(lo, outs[i].off); hi = std::max(hi, outs[i].off + outs[i].bytes); sum += outs[i].bytes; }
        if (nouts > 0) {
            if (hi - lo <= 2 * sum + 65536) {
                VCU(cudaMemcpyAsync(host + lo, base + lo, hi - lo, cudaMemcpyDeviceToHost, s));
            } else {
                for (int i = 0; i < nouts; i++)
                    VCU(cudaMemcpyAsync(host + outs[i].off, base + outs[i].off, outs[i].bytes, cudaMemcpyDeviceToHost, s));
            }
        }
        VCU(cudaStreamSynchronize(s));
        for (int i = 0; i < nouts; i++) memcpy(outs[i].user, host + outs[i].off, outs[i].bytes);
        nouts = 0;
        return VIORB_OK;
    }
};
#define ARENA_CHECK(a) do { if (!(a).ok()) return viorb_fail(VIORB_ERR_CAPACITY, "%s: scratch arena under-sized (%zu B)", __func__, (a).cap); } while (0)

size_t pad(size_t b) { return (b + 255) & ~(size_t)255; }

/* the context's scratch arena, sized by the caller: every take() is bounds-checked against `bytes`.  The first `mirror`
 * bytes (all of it by default) are mirrored in pinned host memory: whatever a call put()s or get()s must lie below that. */
int arena_open(viorb_ctx* c, size_t bytes, Arena* a, size_t mirror = 0) {
    int rc = viorb_ctx_scratch(c, bytes, &a->base);
    if (rc) return rc;
    if (mirror == 0 || mirror > bytes) mirror = bytes;
    if ((rc = viorb_ctx_stage(c, 0, mirror, &a->host))) return rc;
    a->cap = bytes; a->hostCap = mirror; a->off = 0; a->overflow = false; a->slot = 0;
    return VIORB_OK;
}

}  // namespace

struct viorb_frame_index {
    viorb_ctx* ctx = nullptr;
    uint8_t* mem = nullptr;        /* block of the context's pool (viorb_ctx_block_get) */
    size_t memBytes = 0;
    FrameIndexDev dev;
    int* cellOf = nullptr;
    int n = 0;
};

extern "C" {

int viorb_stereo_match(viorb_extractor* left, int frame_l, viorb_extractor* right, int frame_r,
                       const viorb_keypoint* kps_l, const uint8_t* desc_l, int nl, const viorb_keypoint* kps_r,
                       const uint8_t* desc_r, int nr, float mbf, float mb, float* u_right, float* depth) {
    if (!left || !right || nl < 0 || nr < 0 || (nl > 0 && (!kps_l || !desc_l || !u_right || !depth)) ||
        (nr > 0 && (!kps_r || !desc_r)))
        return viorb_fail(VIORB_ERR_INVALID, "bad argument");
    if (nl == 0) return VIORB_OK;
    viorb_ctx* c = viorb_extractor_ctx(left);
    if (viorb_extractor_ctx(right) != c && viorb_ctx_device(viorb_extractor_ctx(right)) != viorb_ctx_device(c))
        return viorb_fail(VIORB_ERR_INVALID, "left and right extractors live on different devices");
    int rc;
    if ((rc = viorb_ctx_bind(c))) return rc;
    StereoParams p;
    memset(&p, 0, sizeof(p));
    int nlevels = 0;
    float scale[16], inv[16];
    if ((rc = viorb_extractor_tables(left, &nlevels, scale, inv, nullptr, nullptr, nullptr))) return rc;
    p.nlevels = nlevels; p.nl = nl; p.nr = nr; p.mbf = mbf; p.mb = mb;
    for (int l = 0; l < nlevels; l++) {
        int w, h, w2, h2;
        const uint8_t *rl, *rr;
        size_t sl, sr;
        if ((rc = viorb_extractor_pyramid_info(left, l, &w, &h))) return rc;
        if ((rc = viorb_extractor_pyramid_info(right, l, &w2, &h2))) return rc;
        if (w != w2 || h != h2) return viorb_fail(VIORB_ERR_INVALID, "left/right pyramids differ in size");
        if ((rc = viorb_extractor_pyramid_device(left, frame_l, l, &rl, &sl))) return rc;
        if ((rc = viorb_extractor_pyramid_device(right, frame_r, l, &rr, &sr))) return rc;
        p.scale[l] = scale[l]; p.invScale[l] = inv[l];
        p.lv[l].roiL = rl; p.lv[l].roiR = rr; p.lv[l].w = w; p.lv[l].h = h; p.lv[l].stepL = (int)sl; p.lv[l].stepR = (int)sr;
    }
    p.nRows = p.lv[0].h;
    /* the extractors may run on other streams of the same device: make their work visible */
    if ((rc = viorb_ctx_synchronize(viorb_extractor_ctx(right)))) return rc;
    if ((rc = viorb_ctx_synchronize(c))) return rc;
    const size_t bytes = pad((size_t)nl * sizeof(viorb_keypoint)) + pad((size_t)nl * 32) + pad((size_t)std::max(nr, 1) * sizeof(viorb_keypoint)) +
                         pad((size_t)std::max(nr, 1) * 32) + 3 * pad((size_t)nl * 4) + pad(64) + 4096;
    Arena a;
    if ((rc = arena_open(c, bytes, &a))) return rc;
    viorb_keypoint* dkl = a.take<viorb_keypoint>(nl);
    uint8_t* ddl = a.take<uint8_t>((size_t)nl * 32);
    viorb_keypoint* dkr = a.take<viorb_keypoint>(std::max(nr, 1));
    uint8_t* ddr = a.take<uint8_t>((size_t)std::max(nr, 1) * 32);
    float* dur = a.take<float>(nl);
    float* ddp = a.take<float>(nl);
    int* dsad = a.take<int>(nl);
    int* dscr = a.take<int>(16);
    ARENA_CHECK(a);
    a.put(dkl, kps_l, nl); a.put(ddl, desc_l, (size_t)nl * 32);
    a.put(dkr, kps_r, nr); a.put(ddr, desc_r, (size_t)nr * 32);
    if ((rc = a.flush(c))) return rc;
    viorb_ctx_add_launches(c, viorb_launch_stereo(p, dkl, ddl, dkr, ddr, dscr, dur, ddp, dsad, viorb_ctx_stream(c)));
    a.get(u_right, dur, (size_t)nl * 4);
    a.get(depth, ddp, (size_t)nl * 4);
    return a.finish(c);
}

static int make_undistort_params(float fx, float fy, float cx, float cy, const float* dist, int ndist, UndistortParams* p) {
    if (ndist != 0 && ndist != 4 && ndist != 5 && ndist != 8 && ndist != 12) return viorb_fail(VIORB_ERR_UNSUPPORTED, "distortion model with %d coefficients", ndist);
    if (ndist > 0 && !dist) return viorb_fail(VIORB_ERR_INVALID, "bad argument");
    memset(p, 0, sizeof(*p));
    p->fx = fx; p->fy = fy; p->cx = cx; p->cy = cy;
    p->ifx = 1.0 / (double)fx; p->ify = 1.0 / (double)fy;
    for (int i = 0; i < ndist; i++) p->k[i] = (double)dist[i];
    p->active = ndist > 0 && dist[0] != 0.0f;            /* src/Frame.cc:586 */
    return VIORB_OK;
}

/* shared by the two constructors: kps are already undistorted (und == NULL) or are undistorted on the device */
static int frame_index_build(viorb_ctx* c, const viorb_keypoint* kps, const uint8_t* desc, const float* u_right, int n,
                             const UndistortParams* und, int cols, int rows, const float* bounds_in,
                             const float* scale_factors, int nlevels, viorb_frame_index** out, bool deviceInputs = false) {
    if (n >= (1 << 20)) return viorb_fail(VIORB_ERR_UNSUPPORTED, "more than 2^20 keypoints in one frame");
    *out = nullptr;
    int rc;
    if ((rc = viorb_ctx_bind(c))) return rc;
    viorb_frame_index* fi = new (std::nothrow) viorb_frame_index();
    if (!fi) return viorb_fail(VIORB_ERR_INVALID, "out of host memory");
    const int nn = std::max(n, 1);
    /* inputs first (one staged copy), then what the kernels produce */
    const size_t inBytes = pad((size_t)nn * sizeof(viorb_keypoint)) + pad((size_t)nn * 32) + pad((size_t)nn * 4) + 1024;
    const size_t bytes = inBytes + pad((size_t)nn * sizeof(viorb_keypoint)) + pad((size_t)nn * 4) * 2 + pad((64 * 48 + 1) * 4) + pad(64) + 4096;
    /* the block comes from the context's pool and goes back to it (viorb_frame_index_destroy): the per-frame path
     * performs no cudaMalloc / cudaFree and, for undistorted keypoints, no synchronisation at all */
    if ((rc = viorb_ctx_block_get(c, bytes, &fi->mem, &fi->memBytes))) { delete fi; return rc; }
    auto fail_index = [&](int code) { viorb_ctx_block_put(c, fi->mem, fi->memBytes); delete fi; return code; };
    Arena a;
    a.base = fi->mem; a.cap = bytes; a.slot = 1;
    if (!deviceInputs) {
        if ((rc = viorb_ctx_stage(c, 1, inBytes, &a.host))) return fail_index(rc);
        a.hostCap = inBytes;
    }
    viorb_keypoint* din = a.take<viorb_keypoint>(nn);          /* keypoints as given: undistorted already, or raw */
    uint8_t* dd = a.take<uint8_t>((size_t)nn * 32);
    float* dur = a.take<float>(nn);
    viorb_keypoint* dund = a.take<viorb_keypoint>(nn);         /* output of the undistortion kernel */
    int* cellOf = a.take<int>(nn);
    int* cellItems = a.take<int>(nn);
    int* cellStart = a.take<int>(64 * 48 + 1);
    float* dbounds = a.take<float>(4);
    if (!a.ok()) return fail_index(viorb_fail(VIORB_ERR_CAPACITY, "frame index arena under-sized"));
    viorb_keypoint* dk = und ? dund : din;
    fi->ctx = c; fi->n = n; fi->cellOf = cellOf;
    cudaStream_t s = viorb_ctx_stream(c);
    if (deviceInputs) {
        /* keypoints and descriptors are already on the device (viorb_extract_batch_device outputs): device-to-device */
        cudaError_t ce = cudaSuccess;
        if (n > 0) ce = cudaMemcpyAsync(din, kps, (size_t)n * sizeof(viorb_keypoint), cudaMemcpyDeviceToDevice, s);
        if (ce == cudaSuccess && n > 0) ce = cudaMemcpyAsync(dd, desc, (size_t)n * 32, cudaMemcpyDeviceToDevice, s);
        if (ce == cudaSuccess && n > 0) {
            if (u_right) ce = cudaMemcpyAsync(dur, u_right, (size_t)n * 4, cudaMemcpyDeviceToDevice, s);
            else ce = cudaMemsetAsync(dur, 0xbf, (size_t)n * 4, s);              /* 0xbfbfbfbf = -1.49: "no right match" (< 0) */
        }
        if (ce != cudaSuccess) return fail_index(viorb_fail(VIORB_ERR_CUDA, "device copy: %s", cudaGetErrorString(ce)));
    } else {
        a.put(din, kps, n);
        a.put(dd, desc, (size_t)n * 32);
        if (u_right) a.put(dur, u_right, n);
        else if (n > 0) {
            float* h = reinterpret_cast<float*>(a.host + (reinterpret_cast<uint8_t*>(dur) - a.base));
            for (int i = 0; i < n; i++) h[i] = -1.0f;
            a.dirtyLo = std::min(a.dirtyLo, (size_t)(reinterpret_cast<uint8_t*>(dur) - a.base));
            a.dirtyHi = std::max(a.dirtyHi, (size_t)(reinterpret_cast<uint8_t*>(dur) - a.base) + (size_t)n * 4);
        }
        if ((rc = a.flush(c))) return fail_index(rc);
    }
    float b[4] = {0, 0, 0, 0};
    if (und) {
        /* Frame::UndistortKeyPoints + ComputeImageBounds on the device; only the four bounds come back */
        viorb_ctx_add_launches(c, viorb_launch_undistort(*und, din, n, dund, s));
        viorb_ctx_add_launches(c, viorb_launch_image_bounds(*und, cols, rows, dbounds, s));
        cudaError_t e = cudaMemcpyAsync(b, dbounds, sizeof(b), cudaMemcpyDeviceToHost, s);
        if (e == cudaSuccess) e = cudaStreamSynchronize(s);
        if (e != cudaSuccess) return fail_index(viorb_fail(VIORB_ERR_CUDA, "undistort: %s", cudaGetErrorString(e)));
    } else {
        memcpy(b, bounds_in, sizeof(b));
    }
    FrameIndexDev& d = fi->dev;
    memset(&d, 0, sizeof(d));
    d.kps = dk; d.desc = dd; d.uRight = dur; d.cellStart = cellStart; d.cellItems = cellItems; d.n = n;
    d.minX = b[0]; d.maxX = b[1]; d.minY = b[2]; d.maxY = b[3];
    d.invW = 64.0f / (d.maxX - d.minX);         /* mfGridElementWidthInv, src/Frame.cc:181 */
    d.invH = 48.0f / (d.maxY - d.minY);
    d.nlevels = nlevels;
    for (int l = 0; l < nlevels; l++) d.scale[l] = scale_factors[l];
    viorb_ctx_add_launches(c, viorb_launch_grid_build(dk, n, d.minX, d.minY, d.invW, d.invH, cellOf, cellStart, cellItems, s));
    const cudaError_t e = cudaGetLastError();      /* the index is used by later calls on the same stream: no wait here */
    if (e != cudaSuccess) return fail_index(viorb_fail(VIORB_ERR_CUDA, "grid build: %s", cudaGetErrorString(e)));
    *out = fi;
    return VIORB_OK;
}

int viorb_frame_index_create(viorb_ctx* c, const viorb_keypoint* kps_un, const uint8_t* desc, const float* u_right, int n,
                             float min_x, float max_x, float min_y, float max_y, const float* scale_factors, int nlevels,
                             viorb_frame_index** out) {
    if (!c || !out || n < 0 || (n > 0 && (!kps_un || !desc)) || !scale_factors || nlevels < 1 || nlevels > 12)
        return viorb_fail(VIORB_ERR_INVALID, "bad argument");
    const float b[4] = {min_x, max_x, min_y, max_y};
    return frame_index_build(c, kps_un, desc, u_right, n, nullptr, 0, 0, b, scale_factors, nlevels, out);
}

int viorb_frame_index_create_distorted(viorb_ctx* c, const viorb_keypoint* kps, const uint8_t* desc, const float* u_right, int n,
                                       float fx, float fy, float cx, float cy, const float* dist_coef, int ndist, int cols,
                                       int rows, const float* scale_factors, int nlevels, viorb_frame_index** out) {
    if (!c || !out || n < 0 || (n > 0 && (!kps || !desc)) || !scale_factors || nlevels < 1 || nlevels > 12 || cols <= 0 || rows <= 0)
        return viorb_fail(VIORB_ERR_INVALID, "bad argument");
    UndistortParams p;
    int rc;
    if ((rc = make_undistort_params(fx, fy, cx, cy, dist_coef, ndist, &p))) return rc;
    return frame_index_build(c, kps, desc, u_right, n, &p, cols, rows, nullptr, scale_factors, nlevels, out);
}

int viorb_frame_index_create_device(viorb_ctx* c, const viorb_keypoint* d_kps, const uint8_t* d_desc, const float* d_u_right, int n,
                                    float fx, float fy, float cx, float cy, const float* dist_coef, int ndist, int cols, int rows,
                                    const float* scale_factors, int nlevels, viorb_frame_index** out) {
    if (!c || !out || n < 0 || (n > 0 && (!d_kps || !d_desc)) || !scale_factors || nlevels < 1 || nlevels > 12 || cols <= 0 || rows <= 0)
        return viorb_fail(VIORB_ERR_INVALID, "bad argument");
    UndistortParams p;
    int rc;
    if ((rc = make_undistort_params(fx, fy, cx, cy, dist_coef, ndist, &p))) return rc;
    return frame_index_build(c, d_kps, d_desc, d_u_right, n, &p, cols, rows, nullptr, scale_factors, nlevels, out, true);
}

int viorb_frame_index_keys(viorb_frame_index* fi, viorb_keypoint* kps_un, float bounds[4]) {
    if (!fi) return viorb_fail(VIORB_ERR_INVALID, "bad argument");
    int rc;
    if ((rc = viorb_ctx_bind(fi->ctx))) return rc;
    if (kps_un && fi->n > 0) {
        VCU(cudaMemcpyAsync(kps_un, fi->dev.kps, (size_t)fi->n * sizeof(viorb_keypoint), cudaMemcpyDeviceToHost, viorb_ctx_stream(fi->ctx)));
        VCU(cudaStreamSynchronize(viorb_ctx_stream(fi->ctx)));
    }
    if (bounds) { bounds[0] = fi->dev.minX; bounds[1] = fi->dev.maxX; bounds[2] = fi->dev.minY; bounds[3] = fi->dev.maxY; }
    return VIORB_OK;
}

int viorb_frame_index_grid(viorb_frame_index* fi, int32_t* cell_start, int32_t* cell_items) {
    if (!fi || !cell_start) return viorb_fail(VIORB_ERR_INVALID, "bad argument");
    int rc;
    if ((rc = viorb_ctx_bind(fi->ctx))) return rc;
    cudaStream_t s = viorb_ctx_stream(fi->ctx);
    VCU(cudaMemcpyAsync(cell_start, fi->dev.cellStart, (64 * 48 + 1) * sizeof(int32_t), cudaMemcpyDeviceToHost, s));
    VCU(cudaStreamSynchronize(s));
    const int total = cell_start[64 * 48];
    if (total > 0) {
        if (!cell_items) return viorb_fail(VIORB_ERR_INVALID, "bad argument");
        VCU(cudaMemcpyAsync(cell_items, fi->dev.cellItems, (size_t)total * sizeof(int32_t), cudaMemcpyDeviceToHost, s));
        VCU(cudaStreamSynchronize(s));
    }
    return VIORB_OK;
}

int viorb_undistort_keypoints(viorb_ctx* c, const viorb_keypoint* kps, int n, float fx, float fy, float cx, float cy,
                              const float* dist_coef, int ndist, viorb_keypoint* kps_un) {
    if (!c || n < 0 || (n > 0 && (!kps || !kps_un))) return viorb_fail(VIORB_ERR_INVALID, "bad argument");
    UndistortParams p;
    int rc;
    if ((rc = make_undistort_params(fx, fy, cx, cy, dist_coef, ndist, &p))) return rc;
    if (n == 0) return VIORB_OK;
    if ((rc = viorb_ctx_bind(c))) return rc;
    Arena a;
    if ((rc = arena_open(c, 2 * pad((size_t)n * sizeof(viorb_keypoint)) + 1024, &a))) return rc;
    viorb_keypoint* din = a.take<viorb_keypoint>(n);
    viorb_keypoint* dout = a.take<viorb_keypoint>(n);
    a.put(din, kps, n);
    if ((rc = a.flush(c))) return rc;
    viorb_ctx_add_launches(c, viorb_launch_undistort(p, din, n, dout, viorb_ctx_stream(c)));
    a.get(kps_un, dout, (size_t)n * sizeof(viorb_keypoint));
    return a.finish(c);
}

int viorb_compute_image_bounds(viorb_ctx* c, int cols, int rows, float fx, float fy, float cx, float cy, const float* dist_coef,
                               int ndist, float bounds[4]) {
    if (!c || !bounds || cols <= 0 || rows <= 0) return viorb_fail(VIORB_ERR_INVALID, "bad argument");
    UndistortParams p;
    int rc;
    if ((rc = make_undistort_params(fx, fy, cx, cy, dist_coef, ndist, &p))) return rc;
    if ((rc = viorb_ctx_bind(c))) return rc;
    Arena a;
    if ((rc = arena_open(c, 1024, &a))) return rc;
    float* db = a.take<float>(4);
    viorb_ctx_add_launches(c, viorb_launch_image_bounds(p, cols, rows, db, viorb_ctx_stream(c)));
    a.get(bounds, db, 16);
    return a.finish(c);
}

int viorb_frame_index_destroy(viorb_frame_index* fi) {
    if (!fi) return VIORB_OK;
    /* the block returns to the context's pool; work still queued on the context's stream that reads it is ordered
     * before whatever the next owner enqueues on the same stream, so no synchronisation is needed */
    viorb_ctx_block_put(fi->ctx, fi->mem, fi->memBytes);
    delete fi;
    return VIORB_OK;
}

int viorb_frame_features_in_area(viorb_frame_index* fi, float x, float y, float r, int min_level, int max_level,
                                 int32_t* out, int cap, int* n) {
    if (!fi || !n || (cap > 0 && !out)) return viorb_fail(VIORB_ERR_INVALID, "bad argument");
    viorb_ctx* c = fi->ctx;
    int rc;
    if ((rc = viorb_ctx_bind(c))) return rc;
    const size_t bytes = pad((size_t)std::max(fi->n, 1) * 8) + pad(64) + 1024;
    Arena a;
    if ((rc = arena_open(c, bytes, &a))) return rc;
    unsigned long long* keys = a.take<unsigned long long>(std::max(fi->n, 1));
    int* cnt = a.take<int>(4);
    viorb_ctx_add_launches(c, viorb_launch_features_in_area(fi->dev, x, y, r, min_level, max_level, keys, cnt, viorb_ctx_stream(c)));
    int h = 0;
    a.get(&h, cnt, 4);
    if ((rc = a.finish(c))) return rc;
    std::vector<unsigned long long> k(std::max(h, 1));
    VCU(cudaMemcpy(k.data(), keys, (size_t)h * 8, cudaMemcpyDeviceToHost));
    std::sort(k.begin(), k.begin() + h);          /* presentation order only: enumeration rank is part of the key */
    for (int i = 0; i < h && i < cap; i++) out[i] = (int32_t)(k[i] & 0xffffffffu);
    *n = h;
    return VIORB_OK;
}

static int check_levels(const int32_t* level, const uint8_t* valid, int n, int nlevels) {
    for (int i = 0; i < n; i++)
        if (valid[i] && (level[i] < 0 || level[i] >= nlevels)) return viorb_fail(VIORB_ERR_INVALID, "query %d: level %d out of range", i, level[i]);
    return VIORB_OK;
}

/* inputs first (one staged host-to-device copy), then the outputs (one copy back), then device-only scratch */
static size_t search_bytes(int nf, int nq) {
    const size_t f = (size_t)std::max(nf, 1), q = (size_t)std::max(nq, 1);
    return pad(f * 4) * 3 + 6 * pad(q * 4) + pad(q) + pad(q * 32) + pad(64) + pad((size_t)viorb_search_scratch_ints((int)q, (int)f) * 4) + 8192;
}

int viorb_search_by_projection_local(viorb_frame_index* fi, int32_t* frame_mp_obs, const float* proj_x, const float* proj_y,
                                     const float* proj_xr, const int32_t* pred_level, const float* view_cos,
                                     const uint8_t* valid, const int32_t* nobs, const uint8_t* mp_desc, int nmp, float th,
                                     float nnratio, int32_t* match, int* nmatches) {
    if (!fi || !frame_mp_obs || !match || !nmatches || nmp < 0 ||
        (nmp > 0 && (!proj_x || !proj_y || !proj_xr || !pred_level || !view_cos || !valid || !nobs || !mp_desc)))
        return viorb_fail(VIORB_ERR_INVALID, "bad argument");
    if (nmp >= (1 << 20) || fi->n >= (1 << 20)) return viorb_fail(VIORB_ERR_UNSUPPORTED, "more than 2^20 map points or keypoints");
    viorb_ctx* c = fi->ctx;
    int rc;
    if ((rc = viorb_ctx_bind(c))) return rc;
    if ((rc = check_levels(pred_level, valid, nmp, fi->dev.nlevels))) return rc;
    const int nf = fi->n, nq = std::max(nmp, 1), mf = std::max(nf, 1);
    Arena a;
    if ((rc = arena_open(c, search_bytes(nf, nmp), &a))) return rc;
    int* dobs0 = a.take<int>(mf);
    float* dpx = a.take<float>(nq); float* dpy = a.take<float>(nq); float* dpr = a.take<float>(nq);
    float* dvc = a.take<float>(nq); int* dlv = a.take<int>(nq); int* dno = a.take<int>(nq);
    uint8_t* dva = a.take<uint8_t>(nq); uint8_t* dde = a.take<uint8_t>((size_t)nq * 32);
    int* dmatch = a.take<int>(mf);
    int* dobs = a.take<int>(mf);
    int* dn = a.take<int>(4);
    int* dscratch = a.take<int>((size_t)viorb_search_scratch_ints(nq, mf));
    ARENA_CHECK(a);
    a.put(dobs0, frame_mp_obs, nf);
    a.put(dpx, proj_x, nmp); a.put(dpy, proj_y, nmp); a.put(dpr, proj_xr, nmp); a.put(dvc, view_cos, nmp);
    a.put(dlv, pred_level, nmp); a.put(dno, nobs, nmp); a.put(dva, valid, nmp); a.put(dde, mp_desc, (size_t)nmp * 32);
    if ((rc = a.flush(c))) return rc;
    viorb_ctx_add_launches(c, viorb_launch_search_local(fi->dev, dpx, dpy, dpr, dlv, dvc, dva, dno, dde, nmp, th, nnratio, dobs0, dobs,
                                                        dscratch, dmatch, dn, viorb_ctx_stream(c)));
    a.get(match, dmatch, (size_t)nf * 4);
    a.get(frame_mp_obs, dobs, (size_t)nf * 4);
    a.get(nmatches, dn, 4);
    return a.finish(c);
}

int viorb_search_by_projection_frame(viorb_frame_index* fi, int32_t* frame_mp_obs, const float* u, const float* v,
                                     const float* invz, const int32_t* last_octave, const float* last_angle,
                                     const uint8_t* valid, const int32_t* nobs, const uint8_t* mp_desc, int nlast, float th,
                                     float mbf, int mode, int check_orientation, int th_high, int32_t* match, int* nmatches) {
    if (!fi || !frame_mp_obs || !match || !nmatches || nlast < 0 || mode < 0 || (mode & 7) > 3 || mode > 15 ||
        (nlast > 0 && (!u || !v || !invz || !last_octave || !last_angle || !valid || !nobs || !mp_desc)))
        return viorb_fail(VIORB_ERR_INVALID, "bad argument");
    if (nlast >= (1 << 24) || fi->n >= (1 << 24)) return viorb_fail(VIORB_ERR_UNSUPPORTED, "more than 2^24 map points or keypoints");
    viorb_ctx* c = fi->ctx;
    int rc;
    if ((rc = viorb_ctx_bind(c))) return rc;
    if ((rc = check_levels(last_octave, valid, nlast, fi->dev.nlevels))) return rc;
    const int nf = fi->n, nq = std::max(nlast, 1), mf = std::max(nf, 1);
    Arena a;
    if ((rc = arena_open(c, search_bytes(nf, nlast), &a))) return rc;
    int* dobs0 = a.take<int>(mf);
    float* du = a.take<float>(nq); float* dv = a.take<float>(nq); float* dz = a.take<float>(nq); float* dan = a.take<float>(nq);
    int* doc = a.take<int>(nq); int* dno = a.take<int>(nq);
    uint8_t* dva = a.take<uint8_t>(nq); uint8_t* dde = a.take<uint8_t>((size_t)nq * 32);
    int* dmatch = a.take<int>(mf);
    int* dobs = a.take<int>(mf);
    int* dn = a.take<int>(4);
    int* dscratch = a.take<int>((size_t)viorb_search_scratch_ints(nq, mf));
    ARENA_CHECK(a);
    a.put(dobs0, frame_mp_obs, nf);
    a.put(du, u, nlast); a.put(dv, v, nlast); a.put(dz, invz, nlast); a.put(dan, last_angle, nlast);
    a.put(doc, last_octave, nlast); a.put(dno, nobs, nlast); a.put(dva, valid, nlast); a.put(dde, mp_desc, (size_t)nlast * 32);
    if ((rc = a.flush(c))) return rc;
    viorb_ctx_add_launches(c, viorb_launch_search_frame(fi->dev, du, dv, dz, doc, dan, dva, dno, dde, nlast, th, mbf, mode,
                                                        check_orientation, th_high, dobs0, dobs, dscratch, dmatch, dn,
                                                        viorb_ctx_stream(c)));
    a.get(match, dmatch, (size_t)nf * 4);
    a.get(frame_mp_obs, dobs, (size_t)nf * 4);
    a.get(nmatches, dn, 4);
    return a.finish(c);
}

int viorb_search_for_triangulation(viorb_ctx* c, const viorb_keypoint* k1, const uint8_t* d1, const float* ur1,
                                   const uint8_t* has_mp1, int n1, const viorb_keypoint* k2, const uint8_t* d2,
                                   const float* ur2, const uint8_t* has_mp2, int n2, const int32_t* node_id1,
                                   const int32_t* node_ptr1, const int32_t* idx1, int nn1, const int32_t* node_id2,
                                   const int32_t* node_ptr2, const int32_t* idx2, int nn2, const float* F12, float ex, float ey,
                                   const float* scale_factors2, const float* level_sigma2_2, int nlevels, int only_stereo,
                                   int check_orientation, int32_t* matches12, int* nmatches) {
    if (!c || n1 < 0 || n2 < 0 || nn1 < 0 || nn2 < 0 || !F12 || !scale_factors2 || !level_sigma2_2 || !matches12 || !nmatches ||
        nlevels < 1 || nlevels > 12 || (n1 > 0 && (!k1 || !d1 || !ur1 || !has_mp1)) || (n2 > 0 && (!k2 || !d2 || !ur2 || !has_mp2)) ||
        (nn1 > 0 && (!node_id1 || !node_ptr1 || !idx1)) || (nn2 > 0 && (!node_id2 || !node_ptr2 || !idx2)))
        return viorb_fail(VIORB_ERR_INVALID, "bad argument");
    if (n2 >= (1 << 24)) return viorb_fail(VIORB_ERR_UNSUPPORTED, "more than 2^24 keypoints");
    int rc;
    if ((rc = viorb_ctx_bind(c))) return rc;
    const int e1 = nn1 > 0 ? node_ptr1[nn1] : 0, e2 = nn2 > 0 ? node_ptr2[nn2] : 0;
    const int m1 = std::max(n1, 1), m2 = std::max(n2, 1);
    const size_t bytes = pad((size_t)m1 * 28) + pad((size_t)m1 * 32) + pad((size_t)m1 * 4) * 2 + pad(m1) + pad((size_t)m2 * 28) +
                         pad((size_t)m2 * 32) + pad((size_t)m2 * 4) + pad(m2) + pad((size_t)(nn1 + 2) * 4) * 2 + pad((size_t)(e1 + 1) * 4) +
                         pad((size_t)(nn2 + 2) * 4) * 2 + pad((size_t)(e2 + 1) * 4) + pad(64) + 16384;
    Arena a;
    if ((rc = arena_open(c, bytes, &a))) return rc;
    viorb_keypoint* dk1 = a.take<viorb_keypoint>(m1); uint8_t* dd1 = a.take<uint8_t>((size_t)m1 * 32);
    float* du1 = a.take<float>(m1); uint8_t* dm1 = a.take<uint8_t>(m1); int* dmatch = a.take<int>(m1);
    viorb_keypoint* dk2 = a.take<viorb_keypoint>(m2); uint8_t* dd2 = a.take<uint8_t>((size_t)m2 * 32);
    float* du2 = a.take<float>(m2); uint8_t* dm2 = a.take<uint8_t>(m2);
    int* dni1 = a.take<int>(nn1 + 1); int* dnp1 = a.take<int>(nn1 + 2); int* di1 = a.take<int>(e1 + 1);
    int* dni2 = a.take<int>(nn2 + 1); int* dnp2 = a.take<int>(nn2 + 2); int* di2 = a.take<int>(e2 + 1);
    int* dn = a.take<int>(4);
    a.put(dk1, k1, n1);
    a.put(dd1, d1, (size_t)n1 * 32);
    a.put(du1, ur1, n1);
    a.put(dm1, has_mp1, n1);
    a.put(dk2, k2, n2);
    a.put(dd2, d2, (size_t)n2 * 32);
    a.put(du2, ur2, n2);
    a.put(dm2, has_mp2, n2);
    a.put(dni1, node_id1, nn1);
    a.put(dnp1, node_ptr1, nn1 ? nn1 + 1 : 0);
    a.put(di1, idx1, e1);
    a.put(dni2, node_id2, nn2);
    a.put(dnp2, node_ptr2, nn2 ? nn2 + 1 : 0);
    a.put(di2, idx2, e2);
    if ((rc = a.flush(c))) return rc;
    viorb_ctx_add_launches(c, viorb_launch_triangulation(dk1, dd1, du1, dm1, n1, dk2, dd2, du2, dm2, n2, dni1, dnp1, di1, nn1, e1,
                                                         dni2, dnp2, di2, nn2, F12, ex, ey, scale_factors2, level_sigma2_2, nlevels,
                                                         only_stereo, check_orientation, dmatch, dn, viorb_ctx_stream(c)));
    a.get(matches12, dmatch, (size_t)n1 * 4);
    a.get(nmatches, dn, 4);
    return a.finish(c);
}

int viorb_distinctive_descriptors(viorb_ctx* c, const uint8_t* obs_desc, const int32_t* obs_ptr, int nmp, int32_t* best,
                                  int32_t* best_median) {
    if (!c || nmp < 0 || (nmp > 0 && (!obs_ptr || !best))) return viorb_fail(VIORB_ERR_INVALID, "bad argument");
    if (nmp == 0) return VIORB_OK;
    if (obs_ptr[0] != 0) return viorb_fail(VIORB_ERR_INVALID, "obs_ptr[0] must be 0");
    for (int i = 0; i < nmp; i++)
        if (obs_ptr[i + 1] < obs_ptr[i]) return viorb_fail(VIORB_ERR_INVALID, "obs_ptr must be non-decreasing");
    const size_t total = (size_t)obs_ptr[nmp];
    if (total > 0 && !obs_desc) return viorb_fail(VIORB_ERR_INVALID, "bad argument");
    int rc;
    if ((rc = viorb_ctx_bind(c))) return rc;
    const size_t bytes = pad(std::max<size_t>(total, 1) * 32) + pad((size_t)(nmp + 1) * 4) + 2 * pad((size_t)nmp * 4) + 2048;
    Arena a;
    if ((rc = arena_open(c, bytes, &a))) return rc;
    uint8_t* dd = a.take<uint8_t>(std::max<size_t>(total, 1) * 32);
    int* dp = a.take<int>(nmp + 1);
    int* db = a.take<int>(nmp);
    int* dm = a.take<int>(nmp);
    a.put(dd, obs_desc, total * 32);
    a.put(dp, obs_ptr, (size_t)nmp + 1);
    if ((rc = a.flush(c))) return rc;
    viorb_ctx_add_launches(c, viorb_launch_distinctive(dd, dp, nmp, db, dm, viorb_ctx_stream(c)));
    a.get(best, db, (size_t)nmp * 4);
    if (best_median) a.get(best_median, dm, (size_t)nmp * 4);
    return a.finish(c);
}

int viorb_search_by_bow(viorb_ctx* c, int mode, const viorb_keypoint* k1, const uint8_t* d1, const uint8_t* valid1, int n1,
                        const viorb_keypoint* k2, const uint8_t* d2, const uint8_t* valid2, int n2, const int32_t* node_id1,
                        const int32_t* node_ptr1, const int32_t* idx1, int nn1, const int32_t* node_id2,
                        const int32_t* node_ptr2, const int32_t* idx2, int nn2, float nnratio, int check_orientation,
                        int32_t* match, int* nmatches) {
    if (!c || (mode != 0 && mode != 1) || n1 < 0 || n2 < 0 || nn1 < 0 || nn2 < 0 || !match || !nmatches ||
        (n1 > 0 && (!k1 || !d1 || !valid1)) || (n2 > 0 && (!k2 || !d2)) || (mode == 1 && n2 > 0 && !valid2) ||
        (nn1 > 0 && (!node_id1 || !node_ptr1 || !idx1)) || (nn2 > 0 && (!node_id2 || !node_ptr2 || !idx2)))
        return viorb_fail(VIORB_ERR_INVALID, "bad argument");
    if (n2 >= (1 << 24)) return viorb_fail(VIORB_ERR_UNSUPPORTED, "more than 2^24 keypoints");
    int rc;
    if ((rc = viorb_ctx_bind(c))) return rc;
    const int e1 = nn1 > 0 ? node_ptr1[nn1] : 0, e2 = nn2 > 0 ? node_ptr2[nn2] : 0;
    const int m1 = std::max(n1, 1), m2 = std::max(n2, 1), mo = std::max(m1, m2);
    const size_t bytes = pad((size_t)m1 * 28) + pad((size_t)m1 * 32) + pad(m1) + pad((size_t)m2 * 28) + pad((size_t)m2 * 32) + pad(m2) +
                         pad((size_t)m2 * 4) + pad((size_t)mo * 4) + pad((size_t)(nn1 + 2) * 4) * 2 + pad((size_t)(e1 + 1) * 4) +
                         pad((size_t)(nn2 + 2) * 4) * 2 + pad((size_t)(e2 + 1) * 4) + pad(64) + 16384;
    Arena a;
    if ((rc = arena_open(c, bytes, &a))) return rc;
    viorb_keypoint* dk1 = a.take<viorb_keypoint>(m1); uint8_t* dd1 = a.take<uint8_t>((size_t)m1 * 32); uint8_t* dv1 = a.take<uint8_t>(m1);
    viorb_keypoint* dk2 = a.take<viorb_keypoint>(m2); uint8_t* dd2 = a.take<uint8_t>((size_t)m2 * 32); uint8_t* dv2 = a.take<uint8_t>(m2);
    int* dtaken = a.take<int>(m2); int* dmatch = a.take<int>(mo);
    int* dni1 = a.take<int>(nn1 + 1); int* dnp1 = a.take<int>(nn1 + 2); int* di1 = a.take<int>(e1 + 1);
    int* dni2 = a.take<int>(nn2 + 1); int* dnp2 = a.take<int>(nn2 + 2); int* di2 = a.take<int>(e2 + 1);
    int* dn = a.take<int>(4);
    a.put(dk1, k1, n1);
    a.put(dd1, d1, (size_t)n1 * 32);
    a.put(dv1, valid1, n1);
    a.put(dk2, k2, n2);
    a.put(dd2, d2, (size_t)n2 * 32);
    if (valid2) a.put(dv2, valid2, n2);
    a.put(dni1, node_id1, nn1);
    a.put(dnp1, node_ptr1, nn1 ? nn1 + 1 : 0);
    a.put(di1, idx1, e1);
    a.put(dni2, node_id2, nn2);
    a.put(dnp2, node_ptr2, nn2 ? nn2 + 1 : 0);
    a.put(di2, idx2, e2);
    if ((rc = a.flush(c))) return rc;
    viorb_ctx_add_launches(c, viorb_launch_search_bow(mode, dk1, dd1, dv1, n1, dk2, dd2, valid2 ? dv2 : nullptr, n2, dni1, dnp1, di1,
                                                      nn1, dni2, dnp2, di2, nn2, nnratio, check_orientation, dtaken, dmatch, dn,
                                                      viorb_ctx_stream(c)));
    a.get(match, dmatch, (size_t)(mode == 0 ? n2 : n1) * 4);
    a.get(nmatches, dn, 4);
    return a.finish(c);
}

int viorb_search_for_initialization(viorb_frame_index* f2, const viorb_keypoint* k1_un, const uint8_t* d1, int n1,
                                    float* prev_matched, int window_size, float nnratio, int check_orientation,
                                    int32_t* matches12, int* nmatches) {
    if (!f2 || n1 < 0 || !matches12 || !nmatches || (n1 > 0 && (!k1_un || !d1 || !prev_matched)))
        return viorb_fail(VIORB_ERR_INVALID, "bad argument");
    viorb_ctx* c = f2->ctx;
    int rc;
    if ((rc = viorb_ctx_bind(c))) return rc;
    const int n2 = f2->n;
    if (n2 >= (1 << 24)) return viorb_fail(VIORB_ERR_UNSUPPORTED, "more than 2^24 keypoints");
    /* candidate pool: only level-0 keypoints of F1 search (:424-426) and only level-0 keypoints of F2 are returned */
    long long q0 = 0;
    for (int i = 0; i < n1; i++) q0 += k1_un[i].octave <= 0;
    const long long cap = std::max<long long>(q0 * std::max(n2, 1), 1);
    if (cap > (1ll << 28)) return viorb_fail(VIORB_ERR_UNSUPPORTED, "candidate pool of %lld entries", cap);
    const int m1 = std::max(n1, 1), m2 = std::max(n2, 1);
    const size_t bytes = pad((size_t)m1 * 28) + pad((size_t)m1 * 32) + pad((size_t)m1 * 8) + pad((size_t)cap * 8) + 4 * pad((size_t)m1 * 4) +
                         2 * pad((size_t)m2 * 4) + 3 * pad(64) + 16384;
    Arena a;
    /* the candidate pool is device-only scratch: it goes last and stays outside of the pinned mirror */
    if ((rc = arena_open(c, bytes, &a, bytes - pad((size_t)cap * 8)))) return rc;
    viorb_keypoint* dk1 = a.take<viorb_keypoint>(m1); uint8_t* dd1 = a.take<uint8_t>((size_t)m1 * 32);
    float* dprev = a.take<float>((size_t)m1 * 2);
    int* dm12 = a.take<int>(m1);
    int* dovf = a.take<int>(4); int* dn = a.take<int>(4);
    int* dstart = a.take<int>(m1); int* dcount = a.take<int>(m1); int* dbin = a.take<int>(m1);
    int* dmd = a.take<int>(m2); int* dm21 = a.take<int>(m2);
    unsigned long long* dcur = a.take<unsigned long long>(2);
    unsigned long long* dent = a.take<unsigned long long>((size_t)cap);
    ARENA_CHECK(a);
    a.put(dk1, k1_un, n1);
    a.put(dd1, d1, (size_t)n1 * 32);
    a.put(dprev, prev_matched, (size_t)n1 * 2);
    if ((rc = a.flush(c))) return rc;
    viorb_ctx_add_launches(c, viorb_launch_search_init(f2->dev, dk1, dd1, n1, dprev, (float)window_size, nnratio, check_orientation, dent,
                                                       cap, dstart, dcount, dcur, dovf, dmd, dm21, dbin, dm12, dn, viorb_ctx_stream(c)));
    int ovf = 0;
    a.get(matches12, dm12, (size_t)n1 * 4);
    a.get(prev_matched, dprev, (size_t)n1 * 8);
    a.get(nmatches, dn, 4);
    a.get(&ovf, dovf, 4);
    if ((rc = a.finish(c))) return rc;
    if (ovf) return viorb_fail(VIORB_ERR_CAPACITY, "candidate pool overflow");
    return VIORB_OK;
}

/* uploads one direction's queries into the arena and launches the windowed top-1 search */
static int window_search_stage(viorb_ctx* c, Arena& a, viorb_frame_index* kf, const float* u, const float* v, const float* ur,
                                const int32_t* level, const uint8_t* valid, const uint8_t* desc, int n, float th, int thDist,
                                const float* invSigma2, int** d_best, int** d_dist) {
    const int m = std::max(n, 1);
    float* du = a.take<float>(m); float* dv = a.take<float>(m); float* dr = a.take<float>(m);
    int* dl = a.take<int>(m); uint8_t* dva = a.take<uint8_t>(m); uint8_t* dde = a.take<uint8_t>((size_t)m * 32);
    *d_best = a.take<int>(m);
    *d_dist = a.take<int>(m);
    int rc;
    a.put(du, u, n);
    a.put(dv, v, n);
    if (ur) a.put(dr, ur, n);
    a.put(dl, level, n);
    a.put(dva, valid, n);
    a.put(dde, desc, (size_t)n * 32);
    if ((rc = a.flush(c))) return rc;
    viorb_ctx_add_launches(c, viorb_launch_search_window(kf->dev, du, dv, ur ? dr : nullptr, dl, dva, dde, n, th, thDist, invSigma2,
                                                         kf->dev.nlevels, *d_best, *d_dist, viorb_ctx_stream(c)));
    return VIORB_OK;
}

static size_t window_search_bytes(int n) {
    const size_t m = (size_t)std::max(n, 1);
    return 6 * pad(m * 4) + pad(m) + pad(m * 32) + 4096;
}

int viorb_search_window_top1(viorb_frame_index* kf, const float* u, const float* v, const float* ur, const int32_t* pred_level,
                             const uint8_t* valid, const uint8_t* mp_desc, int n, float th, int th_dist,
                             const float* inv_level_sigma2, int32_t* best_idx, int32_t* best_dist) {
    if (!kf || n < 0 || !best_idx || (n > 0 && (!u || !v || !pred_level || !valid || !mp_desc)) || (ur && !inv_level_sigma2))
        return viorb_fail(VIORB_ERR_INVALID, "bad argument");
    if (kf->n >= (1 << 24)) return viorb_fail(VIORB_ERR_UNSUPPORTED, "more than 2^24 keypoints");
    if (n == 0) return VIORB_OK;
    viorb_ctx* c = kf->ctx;
    int rc;
    if ((rc = check_levels(pred_level, valid, n, kf->dev.nlevels))) return rc;
    if ((rc = viorb_ctx_bind(c))) return rc;
    Arena a;
    if ((rc = arena_open(c, window_search_bytes(n), &a))) return rc;
    int *db = nullptr, *dd = nullptr;
    if ((rc = window_search_stage(c, a, kf, u, v, ur, pred_level, valid, mp_desc, n, th, th_dist, inv_level_sigma2, &db, &dd))) return rc;
    a.get(best_idx, db, (size_t)n * 4);
    if (best_dist) a.get(best_dist, dd, (size_t)n * 4);
    return a.finish(c);
}

int viorb_search_by_sim3(viorb_frame_index* kf1, viorb_frame_index* kf2, const float* u12, const float* v12,
                         const int32_t* level12, const uint8_t* valid12, const uint8_t* mp_desc1, const float* u21,
                         const float* v21, const int32_t* level21, const uint8_t* valid21, const uint8_t* mp_desc2, float th,
                         int32_t* match12, int* nfound) {
    if (!kf1 || !kf2 || !match12 || !nfound || kf1->ctx != kf2->ctx) return viorb_fail(VIORB_ERR_INVALID, "bad argument");
    const int n1 = kf1->n, n2 = kf2->n;
    if ((n1 > 0 && (!u12 || !v12 || !level12 || !valid12 || !mp_desc1)) || (n2 > 0 && (!u21 || !v21 || !level21 || !valid21 || !mp_desc2)))
        return viorb_fail(VIORB_ERR_INVALID, "bad argument");
    if (n1 >= (1 << 24) || n2 >= (1 << 24)) return viorb_fail(VIORB_ERR_UNSUPPORTED, "more than 2^24 keypoints");
    viorb_ctx* c = kf1->ctx;
    int rc;
    if ((rc = check_levels(level12, valid12, n1, kf2->dev.nlevels)) || (rc = check_levels(level21, valid21, n2, kf1->dev.nlevels))) return rc;
    if ((rc = viorb_ctx_bind(c))) return rc;
    Arena a;
    if ((rc = arena_open(c, window_search_bytes(n1) + window_search_bytes(n2) + pad((size_t)std::max(n1, 1) * 4) + 1024, &a))) return rc;
    int *dm1 = nullptr, *dm2 = nullptr, *dd = nullptr;
    /* KF1's map points searched in KF2 (:1149-1226), KF2's in KF1 (:1228-1303), TH_HIGH = 100 */
    if ((rc = window_search_stage(c, a, kf2, u12, v12, nullptr, level12, valid12, mp_desc1, n1, th, 100, nullptr, &dm1, &dd))) return rc;
    if ((rc = window_search_stage(c, a, kf1, u21, v21, nullptr, level21, valid21, mp_desc2, n2, th, 100, nullptr, &dm2, &dd))) return rc;
    int* dmatch = a.take<int>(std::max(n1, 1));
    int* dn = a.take<int>(4);
    viorb_ctx_add_launches(c, viorb_launch_sim3_agree(dm1, dm2, n1, dmatch, dn, viorb_ctx_stream(c)));
    a.get(match12, dmatch, (size_t)n1 * 4);
    a.get(nfound, dn, 4);
    return a.finish(c);
}

}  // extern "C"
