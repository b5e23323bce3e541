/*
 * orb_oracle.cpp -- CPU oracle for the ORB extractor.  TEST INFRASTRUCTURE ONLY (see orb_oracle.h).
 *
 * Restates, function by function, /root/reference/src/ORBextractor.cc and the OpenCV primitives
 * it imports (OpenCV is not vendored by the reference; arithmetic = OpenCV 4.13 as verified
 * against python cv2 in tests/test_oracle_vs_cv2.py).  Build with -ffp-contract=off: the
 * reference binary contains no FMA.
 *
 * Conventions where the reference is under-determined:
 *   - octree tie rule: std::sort on pair<int,ExtractorNode*> (ORBextractor.cc:684) orders equal
 *     sizes by heap address.  We replace the pointer by the node's creation sequence number,
 *     i.e. the behaviour of the reference under a monotonically growing allocator.
 *   - sincosf: glibc 2.39 flt-32 algorithm evaluated in double without FMA (orc_sincosf).
 */
#include "orb_oracle.h"

#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <list>
#include <thread>
#include <utility>
#include <vector>

#include "../include/viorb_orb_pattern.h"

namespace {

const int PATCH_SIZE = 31;       /* ORBextractor.cc:72 */
const int HALF_PATCH_SIZE = 15;  /* :73 */
const int EDGE_THRESHOLD = 19;   /* :74 */

const int8_t kPattern[1024] = VIORB_ORB_PATTERN_INIT;

inline int cvRoundF(float v) { return (int)lrintf(v); }   /* round-half-to-even (default mode) */
inline int cvRoundD(double v) { return (int)lrint(v); }
inline int cvFloorF(float v) { return (int)floorf(v); }

inline int reflect101(int i, int n) {
    if (n == 1) return 0;
    while (i < 0 || i >= n) {
        if (i < 0) i = -i;
        else i = 2 * n - 2 - i;
    }
    return i;
}

double now_s() {
    using namespace std::chrono;
    return duration<double>(steady_clock::now().time_since_epoch()).count();
}

}  // namespace

extern "C" int orc_cv_round_f(float v) { return cvRoundF(v); }

/* ------------------------------------------------------------------------------------------------
 * cv::resize(..., INTER_LINEAR) for CV_8UC1 (OpenCV imgproc resize.cpp: resizeGeneric_ with
 * HResizeLinear<uchar,int,short,2048> + VResizeLinear fixed point).  Called at ORBextractor.cc:1120.
 * ---------------------------------------------------------------------------------------------- */
extern "C" void orc_resize_linear_u8(const uint8_t* src, int sw, int sh, size_t sstep,
                                     uint8_t* dst, int dw, int dh, size_t dstep) {
    const double inv_scale_x = (double)dw / sw, inv_scale_y = (double)dh / sh;
    const double scale_x = 1. / inv_scale_x, scale_y = 1. / inv_scale_y;
    std::vector<int> xofs(dw), yofs(dh);
    std::vector<short> ialpha(dw * 2), ibeta(dh * 2);
    for (int dx = 0; dx < dw; dx++) {
        float fx = (float)((dx + 0.5) * scale_x - 0.5);
        int sx = cvFloorF(fx);
        fx -= sx;
        if (sx < 0) { fx = 0; sx = 0; }
        if (sx >= sw - 1) { fx = 0; sx = sw - 1; }
        xofs[dx] = sx;
        ialpha[dx * 2] = (short)cvRoundF((1.f - fx) * 2048);
        ialpha[dx * 2 + 1] = (short)cvRoundF(fx * 2048);
    }
    for (int dy = 0; dy < dh; dy++) {
        float fy = (float)((dy + 0.5) * scale_y - 0.5);
        int sy = cvFloorF(fy);
        fy -= sy;
        yofs[dy] = sy;
        ibeta[dy * 2] = (short)cvRoundF((1.f - fy) * 2048);
        ibeta[dy * 2 + 1] = (short)cvRoundF(fy * 2048);
    }
    std::vector<int> row0(dw), row1(dw);
    for (int dy = 0; dy < dh; dy++) {
        int sy0 = std::min(std::max(yofs[dy], 0), sh - 1);
        int sy1 = std::min(std::max(yofs[dy] + 1, 0), sh - 1);
        const uint8_t* S0 = src + (size_t)sy0 * sstep;
        const uint8_t* S1 = src + (size_t)sy1 * sstep;
        for (int dx = 0; dx < dw; dx++) {
            int sx = xofs[dx];
            int sx1 = std::min(sx + 1, sw - 1);
            int a0 = ialpha[dx * 2], a1 = ialpha[dx * 2 + 1];
            row0[dx] = S0[sx] * a0 + S0[sx1] * a1;
            row1[dx] = S1[sx] * a0 + S1[sx1] * a1;
        }
        int b0 = ibeta[dy * 2], b1 = ibeta[dy * 2 + 1];
        uint8_t* D = dst + (size_t)dy * dstep;
        for (int dx = 0; dx < dw; dx++)
            D[dx] = (uint8_t)((((b0 * (row0[dx] >> 4)) >> 16) + ((b1 * (row1[dx] >> 4)) >> 16) + 2) >> 2);
    }
}

/* copyMakeBorder(..., BORDER_REFLECT_101) -- ORBextractor.cc:1122-1128.  dst is the padded buffer
 * origin ((w+2b) x (h+2b)); src may alias the interior of dst. */
extern "C" void orc_border_reflect101_u8(const uint8_t* src, int w, int h, size_t sstep,
                                         uint8_t* dst, size_t dstep, int border) {
    std::vector<uint8_t> tmp((size_t)w * h);
    for (int y = 0; y < h; y++) memcpy(&tmp[(size_t)y * w], src + (size_t)y * sstep, w);
    for (int y = -border; y < h + border; y++) {
        int sy = reflect101(y, h);
        uint8_t* D = dst + (size_t)(y + border) * dstep;
        for (int x = -border; x < w + border; x++) D[x + border] = tmp[(size_t)sy * w + reflect101(x, w)];
    }
}

/* ------------------------------------------------------------------------------------------------
 * cv::FAST(img, kps, threshold, nonmaxSuppression) TYPE_9_16 (OpenCV features2d fast.cpp FAST_t<16>
 * and fast_score.cpp cornerScore<16>).  Called at ORBextractor.cc:809,814.
 * ---------------------------------------------------------------------------------------------- */
namespace {
const int kRingX[16] = {0, 1, 2, 3, 3, 3, 2, 1, 0, -1, -2, -3, -3, -3, -2, -1};
const int kRingY[16] = {3, 3, 2, 1, 0, -1, -2, -3, -3, -3, -2, -1, 0, 1, 2, 3};

int corner_score16(const uint8_t* p, const int* pixel, int threshold) {
    const int K = 8, N = K * 3 + 1;
    int v = p[0];
    short d[N];
    for (int k = 0; k < N; k++) d[k] = (short)(v - p[pixel[k]]);
    int a0 = threshold;
    for (int k = 0; k < 16; k += 2) {
        int a = std::min((int)d[k + 1], (int)d[k + 2]);
        a = std::min(a, (int)d[k + 3]);
        if (a <= a0) continue;
        a = std::min(a, (int)d[k + 4]);
        a = std::min(a, (int)d[k + 5]);
        a = std::min(a, (int)d[k + 6]);
        a = std::min(a, (int)d[k + 7]);
        a = std::min(a, (int)d[k + 8]);
        a0 = std::max(a0, std::min(a, (int)d[k]));
        a0 = std::max(a0, std::min(a, (int)d[k + 9]));
    }
    int b0 = -a0;
    for (int k = 0; k < 16; k += 2) {
        int b = std::max((int)d[k + 1], (int)d[k + 2]);
        b = std::max(b, (int)d[k + 3]);
        b = std::max(b, (int)d[k + 4]);
        b = std::max(b, (int)d[k + 5]);
        if (b >= b0) continue;
        b = std::max(b, (int)d[k + 6]);
        b = std::max(b, (int)d[k + 7]);
        b = std::max(b, (int)d[k + 8]);
        b0 = std::min(b0, std::max(b, (int)d[k]));
        b0 = std::min(b0, std::max(b, (int)d[k + 9]));
    }
    return -b0 - 1;
}
}  // namespace

extern "C" int orc_fast9_16(const uint8_t* img, int w, int h, size_t step, int threshold, int nms,
                            orc_corner* out, int cap) {
    const int K = 8, N = 25;
    int pixel[25];
    for (int k = 0; k < 16; k++) pixel[k] = kRingX[k] + kRingY[k] * (int)step;
    for (int k = 16; k < N; k++) pixel[k] = pixel[k - 16];
    threshold = std::min(std::max(threshold, 0), 255);
    uint8_t tab[512];
    for (int i = -255; i <= 255; i++) tab[i + 255] = (uint8_t)(i < -threshold ? 1 : i > threshold ? 2 : 0);

    std::vector<uint8_t> bufmem((size_t)w * 3, 0);
    uint8_t* buf[3] = {&bufmem[0], &bufmem[w], &bufmem[2 * (size_t)w]};
    std::vector<int> cpmem((size_t)(w + 1) * 3, 0);
    int* cpbuf[3] = {&cpmem[0], &cpmem[w + 1], &cpmem[2 * (size_t)(w + 1)]};
    int nout = 0;

    for (int i = 3; i < h - 2; i++) {
        const uint8_t* ptr = img + (size_t)i * step + 3;
        uint8_t* curr = buf[(i - 3) % 3];
        int* cornerpos = cpbuf[(i - 3) % 3];
        memset(curr, 0, w);
        int ncorners = 0;
        if (i < h - 3) {
            for (int j = 3; j < w - 3; j++, ptr++) {
                int v = ptr[0];
                const uint8_t* t = &tab[0] - v + 255;
                int d = t[ptr[pixel[0]]] | t[ptr[pixel[8]]];
                if (d == 0) continue;
                d &= t[ptr[pixel[2]]] | t[ptr[pixel[10]]];
                d &= t[ptr[pixel[4]]] | t[ptr[pixel[12]]];
                d &= t[ptr[pixel[6]]] | t[ptr[pixel[14]]];
                if (d == 0) continue;
                d &= t[ptr[pixel[1]]] | t[ptr[pixel[9]]];
                d &= t[ptr[pixel[3]]] | t[ptr[pixel[11]]];
                d &= t[ptr[pixel[5]]] | t[ptr[pixel[13]]];
                d &= t[ptr[pixel[7]]] | t[ptr[pixel[15]]];
                bool is_corner = false;
                if (d & 1) {
                    int vt = v - threshold, count = 0;
                    for (int k = 0; k < N; k++) {
                        int x = ptr[pixel[k]];
                        if (x < vt) {
                            if (++count > K) { is_corner = true; break; }
                        } else count = 0;
                    }
                }
                if (!is_corner && (d & 2)) {
                    int vt = v + threshold, count = 0;
                    for (int k = 0; k < N; k++) {
                        int x = ptr[pixel[k]];
                        if (x > vt) {
                            if (++count > K) { is_corner = true; break; }
                        } else count = 0;
                    }
                }
                if (is_corner) {
                    cornerpos[ncorners++] = j;
                    if (nms) curr[j] = (uint8_t)corner_score16(ptr, pixel, threshold);
                    else {
                        if (nout < cap) { out[nout].x = j; out[nout].y = i; out[nout].score = 0; }
                        nout++;
                    }
                }
            }
        }
        cornerpos[w] = ncorners;  /* count kept in the spare last slot of the row buffer */
        if (!nms || i == 3) continue;
        const uint8_t* prev = buf[(i - 4 + 3) % 3];
        const uint8_t* pprev = buf[(i - 5 + 3) % 3];
        const int* cp = cpbuf[(i - 4 + 3) % 3];
        int nc = cp[w];
        for (int k = 0; k < nc; k++) {
            int j = cp[k];
            int score = prev[j];
            if (score > prev[j + 1] && score > prev[j - 1] && score > pprev[j - 1] && score > pprev[j] &&
                score > pprev[j + 1] && score > curr[j - 1] && score > curr[j] && score > curr[j + 1]) {
                if (nout < cap) { out[nout].x = j; out[nout].y = i - 1; out[nout].score = score; }
                nout++;
            }
        }
    }
    return nout;
}

/* ------------------------------------------------------------------------------------------------
 * cv::GaussianBlur(src, dst, Size(7,7), 2, 2, BORDER_REFLECT_101) for CV_8UC1, OpenCV >= 3.4
 * fixed-point path (smooth.dispatch: ufixedpoint16 taps [18,34,48,56,48,34,18]/256).
 * Called at ORBextractor.cc:1086 on a clone of the level ROI (border reflects at the ROI edge).
 * ---------------------------------------------------------------------------------------------- */
/* variant 0: OpenCV >= 3.4 (the 4.13 in this image; verified on cv2).  variant 1: OpenCV 2.4.x, the version the reference
 * pins (CMakeLists.txt:31): getGaussianKernel(7, 2) in float, scaled by 256 and rounded (filter.cpp createSeparableLinearFilter,
 * bits = 8) = [18,34,49,55,49,34,18] (sum 257), row filter 8u->32s, column filter FixedPtCastEx (v + 2^15) >> 16.  The 2.4
 * taps are restated from its published source; no 2.4 binary is runnable here. */
extern "C" void orc_gaussian7_u8_variant(const uint8_t* src, int w, int h, size_t sstep, uint8_t* dst, size_t dstep, int variant) {
    static const int taps[2][7] = {{18, 34, 48, 56, 48, 34, 18}, {18, 34, 49, 55, 49, 34, 18}};
    const int* k = taps[variant ? 1 : 0];
    std::vector<uint16_t> t((size_t)w * h);
    std::vector<uint8_t> row((size_t)w + 6);
    for (int y = 0; y < h; y++) {
        const uint8_t* S = src + (size_t)y * sstep;
        for (int x = -3; x < w + 3; x++) row[x + 3] = S[reflect101(x, w)];
        uint16_t* T = &t[(size_t)y * w];
        const uint8_t* r = row.data();
        for (int x = 0; x < w; x++)
            T[x] = (uint16_t)(k[0] * r[x] + k[1] * r[x + 1] + k[2] * r[x + 2] + k[3] * r[x + 3] + k[4] * r[x + 4] +
                              k[5] * r[x + 5] + k[6] * r[x + 6]);             /* <= 255*257 = 65535 */
    }
    for (int y = 0; y < h; y++) {
        uint8_t* D = dst + (size_t)y * dstep;
        const uint16_t* T[7];
        for (int j = 0; j < 7; j++) T[j] = &t[(size_t)reflect101(y + j - 3, h) * w];
        for (int x = 0; x < w; x++) {
            uint32_t acc = (uint32_t)k[0] * T[0][x] + (uint32_t)k[1] * T[1][x] + (uint32_t)k[2] * T[2][x] +
                           (uint32_t)k[3] * T[3][x] + (uint32_t)k[4] * T[4][x] + (uint32_t)k[5] * T[5][x] +
                           (uint32_t)k[6] * T[6][x];
            const uint32_t v = (acc + 32768u) >> 16;
            D[x] = (uint8_t)(v > 255u ? 255u : v);                              /* 2.4 taps can reach 256 on saturated areas */
        }
    }
}
extern "C" void orc_gaussian7_u8(const uint8_t* src, int w, int h, size_t sstep, uint8_t* dst, size_t dstep) {
    orc_gaussian7_u8_variant(src, w, h, sstep, dst, dstep, 0);
}

/* cv::fastAtan2 (OpenCV core mathfuncs_core: atan_f32 polynomial, degrees).  ORBextractor.cc:103 */
extern "C" float orc_fast_atan2(float y, float x) {
    const float scale = (float)(180.0 / 3.14159265358979323846);
    const float p1 = 0.9997878412794807f * scale;
    const float p3 = -0.3258083974640975f * scale;
    const float p5 = 0.1555786518463281f * scale;
    const float p7 = -0.04432655554792128f * scale;
    const float eps = (float)2.2204460492503131e-16;
    float ax = std::fabs(x), ay = std::fabs(y);
    float a, c, c2;
    if (ax >= ay) {
        c = ay / (ax + eps);
        c2 = c * c;
        a = (((p7 * c2 + p5) * c2 + p3) * c2 + p1) * c;
    } else {
        c = ax / (ay + eps);
        c2 = c * c;
        a = 90.f - (((p7 * c2 + p5) * c2 + p3) * c2 + p1) * c;
    }
    if (x < 0) a = 180.f - a;
    if (y < 0) a = 360.f - a;
    return a;
}

/* sincosf: glibc >= 2.28 sysdeps/ieee754/flt-32/s_sincosf.c (ARM optimized-routines algorithm),
 * generic (non-TOINT_INTRINSICS) reduction, evaluated without FMA.  Only the ranges the extractor
 * can produce (|x| < 120) are implemented; larger arguments fall back to libm.  ORBextractor.cc:113 */
namespace {
struct SinCosTab {
    double sign[4], hpi_inv, hpi, c0, c1, c2, c3, c4, s1, s2, s3;
};
const SinCosTab kSinCos[2] = {
    {{1.0, -1.0, -1.0, 1.0}, 0x1.45F306DC9C883p+23, 0x1.921FB54442D18p0, 0x1p0, -0x1.ffffffd0c621cp-2,
     0x1.55553e1068f19p-5, -0x1.6c087e89a359dp-10, 0x1.99343027bf8c3p-16, -0x1.555545995a603p-3,
     0x1.1107605230bc4p-7, -0x1.994eb3774cf24p-13},
    {{1.0, -1.0, -1.0, 1.0}, 0x1.45F306DC9C883p+23, 0x1.921FB54442D18p0, -0x1p0, 0x1.ffffffd0c621cp-2,
     -0x1.55553e1068f19p-5, 0x1.6c087e89a359dp-10, -0x1.99343027bf8c3p-16, -0x1.555545995a603p-3,
     0x1.1107605230bc4p-7, -0x1.994eb3774cf24p-13}};

inline uint32_t abstop12(float x) {
    uint32_t u;
    memcpy(&u, &x, 4);
    return (u >> 20) & 0x7ff;
}

inline void sincosf_poly(double x, double x2, const SinCosTab* p, int n, float* sinp, float* cosp) {
    double x3, x4, x5, x6, s, c, c1, c2, s1;
    x4 = x2 * x2;
    x3 = x2 * x;
    c2 = p->c3 + x2 * p->c4;
    s1 = p->s2 + x2 * p->s3;
    float* tmp = (n & 1 ? cosp : sinp);
    cosp = (n & 1 ? sinp : cosp);
    sinp = tmp;
    c1 = p->c0 + x2 * p->c1;
    x5 = x3 * x2;
    x6 = x4 * x2;
    s = x + x3 * p->s1;
    c = c1 + x4 * p->c2;
    *sinp = (float)(s + x5 * s1);
    *cosp = (float)(c + x6 * c2);
}
}  // namespace

extern "C" void orc_sincosf(float y, float* sinp, float* cosp) {
    double x = y;
    const SinCosTab* p = &kSinCos[0];
    if (abstop12(y) < abstop12(0x1.921FB6p-1f)) {
        double x2 = x * x;
        if (abstop12(y) < abstop12(0x1p-12f)) {
            *sinp = y;
            *cosp = 1.0f;
            return;
        }
        sincosf_poly(x, x2, p, 0, sinp, cosp);
    } else if (abstop12(y) < abstop12(120.0f)) {
        double r = x * p->hpi_inv;
        int n = ((int32_t)r + 0x800000) >> 24;
        x = x - n * p->hpi;
        double s = p->sign[n & 3];
        if (n & 2) p = &kSinCos[1];
        sincosf_poly(x * s, x * x, p, n, sinp, cosp);
    } else {
        sincosf(y, sinp, cosp);
    }
}

/* IC_Angle's fastAtan2((float)m_01, (float)m_10) (ORBextractor.cc:103) over arrays of integer moments */
extern "C" void orc_orientation_sweep(const int32_t* m01, const int32_t* m10, int64_t n, float* deg, int threads) {
    if (threads < 1) threads = 1;
    std::vector<std::thread> pool;
    for (int t = 0; t < threads; t++)
        pool.emplace_back([=]() {
            for (int64_t i = n * t / threads; i < n * (t + 1) / threads; i++) deg[i] = orc_fast_atan2((float)m01[i], (float)m10[i]);
        });
    for (auto& th : pool) th.join();
}

/* Sweep of the steering of computeOrbDescriptor (ORBextractor.cc:107-113) over the n consecutive float bit patterns
 * first_bits, first_bits+1, ... taken as keypoint angles in degrees.  which = 0: this image's libm sincosf (what the
 * reference binary calls); which = 1: the restatement above.  Outputs may be null; the return value is
 * the number of inputs on which restatement and libm differ in either output. */
extern "C" int64_t orc_steering_sweep(uint32_t first_bits, int64_t n, int which, float* sin_out, float* cos_out, int threads) {
    if (threads < 1) threads = 1;
    std::vector<int64_t> bad((size_t)threads, 0);
    std::vector<std::thread> pool;
    const float factorPI = (float)(3.14159265358979323846 / 180.f);
    for (int t = 0; t < threads; t++)
        pool.emplace_back([&, t]() {
            const int64_t lo = n * t / threads, hi = n * (t + 1) / threads;
            for (int64_t i = lo; i < hi; i++) {
                const uint32_t u = first_bits + (uint32_t)i;
                float deg, sl, cl, sr, cr;
                memcpy(&deg, &u, 4);
                const float ang = deg * factorPI;
                sincosf(ang, &sl, &cl);
                orc_sincosf(ang, &sr, &cr);
                if (memcmp(&sl, &sr, 4) || memcmp(&cl, &cr, 4)) bad[(size_t)t]++;
                if (sin_out) sin_out[i] = which ? sr : sl;
                if (cos_out) cos_out[i] = which ? cr : cl;
            }
        });
    for (auto& th : pool) th.join();
    int64_t total = 0;
    for (int64_t b : bad) total += b;
    return total;
}

/* ------------------------------------------------------------------------------------------------
 * ORBextractor
 * ---------------------------------------------------------------------------------------------- */
namespace {

struct Image {
    std::vector<uint8_t> buf;
    int w = 0, h = 0;   /* ROI size */
    size_t step = 0;    /* padded row stride */
    uint8_t* roi() { return &buf[(size_t)EDGE_THRESHOLD * step + EDGE_THRESHOLD]; }
    const uint8_t* roi() const { return &buf[(size_t)EDGE_THRESHOLD * step + EDGE_THRESHOLD]; }
};

struct Key {            /* minimal cv::KeyPoint during detection */
    float x, y, response;
    int seq;            /* position in vToDistributeKeys (for stand-alone octree tests) */
};

/* ORBextractor.h:35-50 */
struct Node {
    std::vector<Key> vKeys;
    int ULx, ULy, URx, URy, BLx, BLy, BRx, BRy;
    std::list<Node>::iterator lit;
    bool bNoMore = false;
    long seq = 0;       /* creation sequence number: stands in for the heap address (tie rule) */
};

/* ExtractorNode::DivideNode, ORBextractor.cc:481-537 */
void divide_node(const Node& n, Node& n1, Node& n2, Node& n3, Node& n4) {
    const int halfX = (int)ceilf(static_cast<float>(n.URx - n.ULx) / 2);
    const int halfY = (int)ceilf(static_cast<float>(n.BRy - n.ULy) / 2);
    n1.ULx = n.ULx; n1.ULy = n.ULy;
    n1.URx = n.ULx + halfX; n1.URy = n.ULy;
    n1.BLx = n.ULx; n1.BLy = n.ULy + halfY;
    n1.BRx = n.ULx + halfX; n1.BRy = n.ULy + halfY;
    n2.ULx = n1.URx; n2.ULy = n1.URy;
    n2.URx = n.URx; n2.URy = n.URy;
    n2.BLx = n1.BRx; n2.BLy = n1.BRy;
    n2.BRx = n.URx; n2.BRy = n.ULy + halfY;
    n3.ULx = n1.BLx; n3.ULy = n1.BLy;
    n3.URx = n1.BRx; n3.URy = n1.BRy;
    n3.BLx = n.BLx; n3.BLy = n.BLy;
    n3.BRx = n1.BRx; n3.BRy = n.BLy;
    n4.ULx = n3.URx; n4.ULy = n3.URy;
    n4.URx = n2.BRx; n4.URy = n2.BRy;
    n4.BLx = n3.BRx; n4.BLy = n3.BRy;
    n4.BRx = n.BRx; n4.BRy = n.BRy;
    for (size_t i = 0; i < n.vKeys.size(); i++) {
        const Key& kp = n.vKeys[i];
        if (kp.x < n1.URx) {
            if (kp.y < n1.BRy) n1.vKeys.push_back(kp);
            else n3.vKeys.push_back(kp);
        } else if (kp.y < n1.BRy) n2.vKeys.push_back(kp);
        else n4.vKeys.push_back(kp);
    }
    if (n1.vKeys.size() == 1) n1.bNoMore = true;
    if (n2.vKeys.size() == 1) n2.bNoMore = true;
    if (n3.vKeys.size() == 1) n3.bNoMore = true;
    if (n4.vKeys.size() == 1) n4.bNoMore = true;
}

/* ORBextractor::DistributeOctTree, ORBextractor.cc:539-763 (literal list-based restatement) */
std::vector<Key> distribute_octree(const std::vector<Key>& vToDistributeKeys, int minX, int maxX, int minY,
                                   int maxY, int N) {
    const int nIni = (int)roundf(static_cast<float>(maxX - minX) / (maxY - minY));
    const float hX = static_cast<float>(maxX - minX) / nIni;
    std::list<Node> lNodes;
    std::vector<Node*> vpIniNodes(nIni);
    long seq = 0;
    for (int i = 0; i < nIni; i++) {
        Node ni;
        ni.ULx = (int)(hX * static_cast<float>(i)); ni.ULy = 0;
        ni.URx = (int)(hX * static_cast<float>(i + 1)); ni.URy = 0;
        ni.BLx = ni.ULx; ni.BLy = maxY - minY;
        ni.BRx = ni.URx; ni.BRy = maxY - minY;
        ni.seq = seq++;
        lNodes.push_back(ni);
        vpIniNodes[i] = &lNodes.back();
    }
    for (size_t i = 0; i < vToDistributeKeys.size(); i++) {
        const Key& kp = vToDistributeKeys[i];
        vpIniNodes[(int)(kp.x / hX)]->vKeys.push_back(kp);
    }
    std::list<Node>::iterator lit = lNodes.begin();
    while (lit != lNodes.end()) {
        if (lit->vKeys.size() == 1) { lit->bNoMore = true; lit++; }
        else if (lit->vKeys.empty()) lit = lNodes.erase(lit);
        else lit++;
    }
    bool bFinish = false;
    std::vector<std::pair<int, Node*> > vSizeAndPointerToNode;

    auto push_children = [&](Node& n1, Node& n2, Node& n3, Node& n4, int* nToExpand) {
        Node* ch[4] = {&n1, &n2, &n3, &n4};
        for (int c = 0; c < 4; c++) {
            if (ch[c]->vKeys.size() > 0) {
                ch[c]->seq = seq++;
                lNodes.push_front(*ch[c]);
                if (ch[c]->vKeys.size() > 1) {
                    if (nToExpand) (*nToExpand)++;
                    vSizeAndPointerToNode.push_back(std::make_pair((int)ch[c]->vKeys.size(), &lNodes.front()));
                    lNodes.front().lit = lNodes.begin();
                }
            }
        }
    };
    /* (size, pointer) ordering with pointer := creation sequence */
    auto size_ptr_less = [](const std::pair<int, Node*>& a, const std::pair<int, Node*>& b) {
        if (a.first != b.first) return a.first < b.first;
        return a.second->seq < b.second->seq;
    };

    while (!bFinish) {
        int prevSize = (int)lNodes.size();
        lit = lNodes.begin();
        int nToExpand = 0;
        vSizeAndPointerToNode.clear();
        while (lit != lNodes.end()) {
            if (lit->bNoMore) { lit++; continue; }
            Node n1, n2, n3, n4;
            divide_node(*lit, n1, n2, n3, n4);
            push_children(n1, n2, n3, n4, &nToExpand);
            lit = lNodes.erase(lit);
        }
        if ((int)lNodes.size() >= N || (int)lNodes.size() == prevSize) {
            bFinish = true;
        } else if (((int)lNodes.size() + nToExpand * 3) > N) {
            while (!bFinish) {
                prevSize = (int)lNodes.size();
                std::vector<std::pair<int, Node*> > vPrev = vSizeAndPointerToNode;
                vSizeAndPointerToNode.clear();
                std::sort(vPrev.begin(), vPrev.end(), size_ptr_less);
                for (int j = (int)vPrev.size() - 1; j >= 0; j--) {
                    Node n1, n2, n3, n4;
                    divide_node(*vPrev[j].second, n1, n2, n3, n4);
                    push_children(n1, n2, n3, n4, nullptr);
                    lNodes.erase(vPrev[j].second->lit);
                    if ((int)lNodes.size() >= N) break;
                }
                if ((int)lNodes.size() >= N || (int)lNodes.size() == prevSize) bFinish = true;
            }
        }
    }
    std::vector<Key> vResultKeys;
    for (lit = lNodes.begin(); lit != lNodes.end(); lit++) {
        std::vector<Key>& vNodeKeys = lit->vKeys;
        Key* pKP = &vNodeKeys[0];
        float maxResponse = pKP->response;
        for (size_t k = 1; k < vNodeKeys.size(); k++) {
            if (vNodeKeys[k].response > maxResponse) {
                pKP = &vNodeKeys[k];
                maxResponse = vNodeKeys[k].response;
            }
        }
        vResultKeys.push_back(*pKP);
    }
    return vResultKeys;
}

}  // namespace

struct orc_extractor {
    int nfeatures, nlevels, iniThFAST, minThFAST;
    double scaleFactor;
    std::vector<float> mvScaleFactor, mvInvScaleFactor, mvLevelSigma2, mvInvLevelSigma2;
    std::vector<int> mnFeaturesPerLevel, umax;
    std::vector<Image> pyramid, blurred;
    std::vector<std::vector<Key> > candidates;
    std::vector<std::vector<orc_keypoint> > levelKeys;
    double stage[6];
    int retried;
    int gaussVariant = 0;        /* 0 = OpenCV >= 3.4 taps, 1 = OpenCV 2.4 taps (orc_gaussian7_u8_variant) */
};

extern "C" orc_extractor* orc_extractor_create(int nfeatures, float scale_factor, int nlevels, int ini, int mn) {
    /* ORBextractor::ORBextractor, ORBextractor.cc:410-470 */
    orc_extractor* e = new orc_extractor();
    e->nfeatures = nfeatures; e->nlevels = nlevels; e->iniThFAST = ini; e->minThFAST = mn;
    e->scaleFactor = scale_factor;
    e->mvScaleFactor.resize(nlevels); e->mvLevelSigma2.resize(nlevels);
    e->mvScaleFactor[0] = 1.0f; e->mvLevelSigma2[0] = 1.0f;
    for (int i = 1; i < nlevels; i++) {
        e->mvScaleFactor[i] = (float)(e->mvScaleFactor[i - 1] * e->scaleFactor);
        e->mvLevelSigma2[i] = e->mvScaleFactor[i] * e->mvScaleFactor[i];
    }
    e->mvInvScaleFactor.resize(nlevels); e->mvInvLevelSigma2.resize(nlevels);
    for (int i = 0; i < nlevels; i++) {
        e->mvInvScaleFactor[i] = 1.0f / e->mvScaleFactor[i];
        e->mvInvLevelSigma2[i] = 1.0f / e->mvLevelSigma2[i];
    }
    e->mnFeaturesPerLevel.resize(nlevels);
    float factor = (float)(1.0f / e->scaleFactor);
    float nDesired = nfeatures * (1 - factor) / (1 - (float)pow((double)factor, (double)nlevels));
    int sumFeatures = 0;
    for (int level = 0; level < nlevels - 1; level++) {
        e->mnFeaturesPerLevel[level] = cvRoundF(nDesired);
        sumFeatures += e->mnFeaturesPerLevel[level];
        nDesired *= factor;
    }
    e->mnFeaturesPerLevel[nlevels - 1] = std::max(nfeatures - sumFeatures, 0);

    e->umax.resize(HALF_PATCH_SIZE + 1);
    int v, v0, vmax = cvFloorF(HALF_PATCH_SIZE * sqrtf(2.f) / 2 + 1);
    int vmin = (int)ceilf(HALF_PATCH_SIZE * sqrtf(2.f) / 2);
    const double hp2 = HALF_PATCH_SIZE * HALF_PATCH_SIZE;
    for (v = 0; v <= vmax; ++v) e->umax[v] = cvRoundD(sqrt(hp2 - v * v));
    for (v = HALF_PATCH_SIZE, v0 = 0; v >= vmin; --v) {
        while (e->umax[v0] == e->umax[v0 + 1]) ++v0;
        e->umax[v] = v0;
        ++v0;
    }
    e->pyramid.resize(nlevels); e->blurred.resize(nlevels);
    e->candidates.resize(nlevels); e->levelKeys.resize(nlevels);
    memset(e->stage, 0, sizeof(e->stage));
    e->retried = 0;
    return e;
}

extern "C" void orc_extractor_destroy(orc_extractor* e) { delete e; }

namespace {

/* ORBextractor::ComputePyramid, ORBextractor.cc:1107-1132 */
void compute_pyramid(orc_extractor* e, const uint8_t* img, int rows, int cols, size_t step) {
    for (int level = 0; level < e->nlevels; ++level) {
        float scale = e->mvInvScaleFactor[level];
        int w = cvRoundF((float)cols * scale), h = cvRoundF((float)rows * scale);
        Image& im = e->pyramid[level];
        im.w = w; im.h = h; im.step = (size_t)w + 2 * EDGE_THRESHOLD;
        im.buf.assign(im.step * (h + 2 * EDGE_THRESHOLD), 0);
        if (level != 0) {
            const Image& pv = e->pyramid[level - 1];
            orc_resize_linear_u8(pv.roi(), pv.w, pv.h, pv.step, im.roi(), w, h, im.step);
            orc_border_reflect101_u8(im.roi(), w, h, im.step, &im.buf[0], im.step, EDGE_THRESHOLD);
        } else {
            orc_border_reflect101_u8(img, cols, rows, step, &im.buf[0], im.step, EDGE_THRESHOLD);
        }
    }
}

/* IC_Angle, ORBextractor.cc:77-104 */
float ic_angle(const Image& im, float px, float py, const std::vector<int>& u_max) {
    int m_01 = 0, m_10 = 0;
    const int step = (int)im.step;
    const uint8_t* center = im.roi() + (size_t)cvRoundF(py) * im.step + cvRoundF(px);
    for (int u = -HALF_PATCH_SIZE; u <= HALF_PATCH_SIZE; ++u) m_10 += u * center[u];
    for (int v = 1; v <= HALF_PATCH_SIZE; ++v) {
        int v_sum = 0;
        int d = u_max[v];
        for (int u = -d; u <= d; ++u) {
            int val_plus = center[u + v * step], val_minus = center[u - v * step];
            v_sum += (val_plus - val_minus);
            m_10 += u * (val_plus + val_minus);
        }
        m_01 += v * v_sum;
    }
    return orc_fast_atan2((float)m_01, (float)m_10);
}

/* computeOrbDescriptor, ORBextractor.cc:108-147; img = blurred level (own buffer, same geometry) */
void compute_orb_descriptor(const orc_keypoint& kpt, const Image& img, uint8_t* desc) {
    const float factorPI = (float)(3.1415926535897932384626433832795 / 180.f);
    float angle = (float)kpt.angle * factorPI;
    float a, b;
    orc_sincosf(angle, &b, &a);
    const uint8_t* center = img.roi() + (size_t)cvRoundF(kpt.y) * img.step + cvRoundF(kpt.x);
    const int step = (int)img.step;
    const int8_t* pat = kPattern;
    for (int i = 0; i < 32; ++i, pat += 32) {
        int val = 0;
        for (int k = 0; k < 8; k++) {
            int x0 = pat[4 * k], y0 = pat[4 * k + 1], x1 = pat[4 * k + 2], y1 = pat[4 * k + 3];
            int t0 = center[cvRoundF(x0 * b + y0 * a) * step + cvRoundF(x0 * a - y0 * b)];
            int t1 = center[cvRoundF(x1 * b + y1 * a) * step + cvRoundF(x1 * a - y1 * b)];
            val |= (t0 < t1) << k;
        }
        desc[i] = (uint8_t)val;
    }
}

/* ORBextractor::ComputeKeyPointsOctTree, ORBextractor.cc:765-853 */
void compute_keypoints_octree(orc_extractor* e) {
    const float W = 30;
    double t_fast = 0, t_oct = 0;
    e->retried = 0;
    std::vector<orc_corner> cell(4096);
    for (int level = 0; level < e->nlevels; ++level) {
        double t0 = now_s();
        const Image& im = e->pyramid[level];
        const int minBorderX = EDGE_THRESHOLD - 3;
        const int minBorderY = minBorderX;
        const int maxBorderX = im.w - EDGE_THRESHOLD + 3;
        const int maxBorderY = im.h - EDGE_THRESHOLD + 3;
        std::vector<Key>& vToDistributeKeys = e->candidates[level];
        vToDistributeKeys.clear();
        const float width = (float)(maxBorderX - minBorderX);
        const float height = (float)(maxBorderY - minBorderY);
        const int nCols = (int)(width / W);
        const int nRows = (int)(height / W);
        const int wCell = (int)ceilf(width / nCols);
        const int hCell = (int)ceilf(height / nRows);
        for (int i = 0; i < nRows; i++) {
            const float iniY = (float)(minBorderY + i * hCell);
            float maxY = iniY + hCell + 6;
            if (iniY >= maxBorderY - 3) continue;
            if (maxY > maxBorderY) maxY = (float)maxBorderY;
            for (int j = 0; j < nCols; j++) {
                const float iniX = (float)(minBorderX + j * wCell);
                float maxX = iniX + wCell + 6;
                if (iniX >= maxBorderX - 6) continue;
                if (maxX > maxBorderX) maxX = (float)maxBorderX;
                const int x0 = (int)iniX, y0 = (int)iniY, cw = (int)maxX - x0, chh = (int)maxY - y0;
                const uint8_t* sub = im.roi() + (size_t)y0 * im.step + x0;
                int n = orc_fast9_16(sub, cw, chh, im.step, e->iniThFAST, 1, cell.data(), (int)cell.size());
                if (n == 0) {
                    n = orc_fast9_16(sub, cw, chh, im.step, e->minThFAST, 1, cell.data(), (int)cell.size());
                    e->retried++;
                }
                for (int k = 0; k < n && k < (int)cell.size(); k++) {
                    Key kp;
                    kp.x = (float)cell[k].x + j * wCell;
                    kp.y = (float)cell[k].y + i * hCell;
                    kp.response = (float)cell[k].score;
                    kp.seq = (int)vToDistributeKeys.size();
                    vToDistributeKeys.push_back(kp);
                }
            }
        }
        double t1 = now_s();
        std::vector<Key> sel;
        if (!vToDistributeKeys.empty())
            sel = distribute_octree(vToDistributeKeys, minBorderX, maxBorderX, minBorderY, maxBorderY,
                                    e->mnFeaturesPerLevel[level]);
        const int scaledPatchSize = (int)(PATCH_SIZE * e->mvScaleFactor[level]);
        std::vector<orc_keypoint>& out = e->levelKeys[level];
        out.clear();
        for (size_t k = 0; k < sel.size(); k++) {
            orc_keypoint kp;
            kp.x = sel[k].x + minBorderX;
            kp.y = sel[k].y + minBorderY;
            kp.size = (float)scaledPatchSize;
            kp.angle = -1;
            kp.response = sel[k].response;
            kp.octave = level;
            kp.class_id = -1;
            out.push_back(kp);
        }
        double t2 = now_s();
        t_fast += t1 - t0;
        t_oct += t2 - t1;
    }
    double t3 = now_s();
    for (int level = 0; level < e->nlevels; ++level)
        for (auto& kp : e->levelKeys[level]) kp.angle = ic_angle(e->pyramid[level], kp.x, kp.y, e->umax);
    e->stage[1] = t_fast; e->stage[2] = t_oct; e->stage[3] = now_s() - t3;
}

}  // namespace

/* ORBextractor::operator(), ORBextractor.cc:1043-1105 */
extern "C" int orc_extract(orc_extractor* e, const uint8_t* img, int rows, int cols, size_t step,
                           orc_keypoint* kps, uint8_t* desc, int cap) {
    if (!img || rows <= 0 || cols <= 0) return 0;
    double t0 = now_s();
    compute_pyramid(e, img, rows, cols, step);
    e->stage[0] = now_s() - t0;
    compute_keypoints_octree(e);
    int offset = 0;
    double t_blur = 0, t_desc = 0;
    for (int level = 0; level < e->nlevels; ++level) {
        std::vector<orc_keypoint>& keypoints = e->levelKeys[level];
        Image& bl = e->blurred[level];
        const Image& im = e->pyramid[level];
        if (keypoints.empty()) { bl.buf.clear(); bl.w = bl.h = 0; continue; }
        double t1 = now_s();
        bl.w = im.w; bl.h = im.h; bl.step = im.step;
        bl.buf.assign(im.buf.size(), 0);
        orc_gaussian7_u8_variant(im.roi(), im.w, im.h, im.step, bl.roi(), bl.step, e->gaussVariant);
        double t2 = now_s();
        for (size_t i = 0; i < keypoints.size(); i++) {
            if (offset + (int)i < cap) compute_orb_descriptor(keypoints[i], bl, desc + (size_t)(offset + i) * 32);
        }
        float scale = e->mvScaleFactor[level];
        for (size_t i = 0; i < keypoints.size(); i++) {
            if (offset + (int)i >= cap) break;
            orc_keypoint kp = keypoints[i];
            if (level != 0) { kp.x *= scale; kp.y *= scale; }
            kps[offset + i] = kp;
        }
        offset += (int)keypoints.size();
        t_blur += t2 - t1;
        t_desc += now_s() - t2;
    }
    e->stage[4] = t_blur; e->stage[5] = t_desc;
    return offset;
}

extern "C" void orc_extractor_set_gaussian(orc_extractor* e, int variant) { e->gaussVariant = variant ? 1 : 0; }
extern "C" int orc_extractor_levels(const orc_extractor* e) { return e->nlevels; }
extern "C" int orc_extractor_quota(const orc_extractor* e, int l) { return e->mnFeaturesPerLevel[l]; }
extern "C" float orc_extractor_scale(const orc_extractor* e, int l) { return e->mvScaleFactor[l]; }
extern "C" int orc_extractor_umax(const orc_extractor* e, int v) { return e->umax[v]; }
extern "C" const uint8_t* orc_extractor_pyramid(const orc_extractor* e, int l, int* w, int* h, size_t* step) {
    const Image& im = e->pyramid[l];
    *w = im.w; *h = im.h; *step = im.step;
    return im.buf.empty() ? nullptr : im.buf.data();
}
extern "C" const uint8_t* orc_extractor_blurred(const orc_extractor* e, int l, int* w, int* h, size_t* step) {
    const Image& im = e->blurred[l];
    *w = im.w; *h = im.h; *step = im.step;
    return im.buf.empty() ? nullptr : im.buf.data();
}
extern "C" int orc_extractor_candidates(const orc_extractor* e, int l, orc_corner* out, int cap) {
    const std::vector<Key>& c = e->candidates[l];
    for (size_t i = 0; i < c.size() && (int)i < cap; i++) {
        out[i].x = (int)c[i].x; out[i].y = (int)c[i].y; out[i].score = (int)c[i].response;
    }
    return (int)c.size();
}
extern "C" int orc_extractor_level_keypoints(const orc_extractor* e, int l, orc_keypoint* out, int cap) {
    const std::vector<orc_keypoint>& k = e->levelKeys[l];
    for (size_t i = 0; i < k.size() && (int)i < cap; i++) out[i] = k[i];
    return (int)k.size();
}
extern "C" void orc_extractor_stage_seconds(const orc_extractor* e, double out[6]) {
    for (int i = 0; i < 6; i++) out[i] = e->stage[i];
}
extern "C" int orc_extractor_retried_cells(const orc_extractor* e) { return e->retried; }

extern "C" int orc_distribute_octree(const orc_corner* cand, int n, int minX, int maxX, int minY, int maxY,
                                     int N, int32_t* out_index, int cap) {
    std::vector<Key> keys(n);
    for (int i = 0; i < n; i++) {
        keys[i].x = (float)cand[i].x; keys[i].y = (float)cand[i].y;
        keys[i].response = (float)cand[i].score; keys[i].seq = i;
    }
    if (n == 0) return 0;
    std::vector<Key> sel = distribute_octree(keys, minX, maxX, minY, maxY, N);
    for (size_t i = 0; i < sel.size() && (int)i < cap; i++) out_index[i] = sel[i].seq;
    return (int)sel.size();
}

extern "C" int orc_num_threads(void) { return (int)std::thread::hardware_concurrency(); }
