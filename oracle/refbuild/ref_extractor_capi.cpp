/*
 * ref_extractor_capi.cpp -- flat C entry points around the REFERENCE's own ORB_SLAM2::ORBextractor, compiled unmodified from
 * /root/reference/src/ORBextractor.cc against the stand-in OpenCV headers of oracle/refbuild/minicv (see the Makefile).
 * TEST INFRASTRUCTURE ONLY.  Signatures mirror the orc_* functions of oracle/orb_oracle.h so the same test code drives
 * either the restatement (liborb_oracle.so) or the reference (oracle/_ref/libviorb_ref.so).
 *
 * The namespace is renamed at compile time (-DORB_SLAM2=ORB_SLAM2_ref) so that the reference's classes can live next to
 * the product's drop-in classes of the same name in one test process.
 *
 * DistributeOctTree sorts pair<int, ExtractorNode*> (src/ORBextractor.cc:684): equal sizes are ordered by heap address, so
 * the reference's result depends on the allocator.  ref_set_allocator() selects what operator new hands out while an
 * extraction runs: 0 = the process allocator (glibc malloc), 1 = a bump arena with ascending addresses (a later allocation
 * has a higher address: the convention the oracle and the CUDA path implement), 2 = a bump arena with descending addresses
 * (the opposite tie order, to measure how much of the output the pointer tie-break decides).
 */
#include <opencv2/core/core.hpp>

#include <cstdint>
#include <cstdlib>
#include <new>
#include <vector>

#include "ORBextractor.h"
#include "../orb_oracle.h"

#include "minicv_hooks.h"

/* ---------------------------------------------------------------------------------- allocator modes */
#include <sys/mman.h>

#include <atomic>
namespace {
/* A block serves one extraction call.  Memory handed out is never reused while the block lives, so objects that outlive
 * the call (mvImagePyramid) stay valid; the block is dropped once the same extractor has finished its NEXT call. */
struct Block {
    char* base = nullptr;
    size_t cap = 0, lo = 0, hi = 0;
};
const int kMaxBlocks = 256;
std::atomic<char*> g_block_base[kMaxBlocks];
std::atomic<size_t> g_block_cap[kMaxBlocks];
std::atomic<long> g_overflows{0};
int g_allocator_mode = 1;

struct ThreadArena {
    Block* cur = nullptr;
    int mode = 0;
    bool active = false;
};
thread_local ThreadArena g_arena;

bool in_any_block(void* p) {
    for (int i = 0; i < kMaxBlocks; i++) {
        char* b = g_block_base[i].load(std::memory_order_acquire);
        if (b && (char*)p >= b && (char*)p < b + g_block_cap[i].load(std::memory_order_relaxed)) return true;
    }
    return false;
}
Block* block_create(size_t bytes) {
    Block* blk = (Block*)std::malloc(sizeof(Block));
    /* address space only: pages are committed as the bump pointer touches them */
    void* m = mmap(nullptr, bytes, PROT_READ | PROT_WRITE, MAP_PRIVATE | MAP_ANONYMOUS | MAP_NORESERVE, -1, 0);
    if (m == MAP_FAILED) std::abort();
    blk->base = (char*)m;
    blk->cap = bytes; blk->lo = 0; blk->hi = bytes;
    for (int i = 0; i < kMaxBlocks; i++) {
        char* expect = nullptr;
        if (g_block_base[i].load() == nullptr) {
            g_block_cap[i].store(bytes);
            if (g_block_base[i].compare_exchange_strong(expect, blk->base)) return blk;
        }
    }
    std::abort();
}
void block_destroy(Block* blk) {
    if (!blk) return;
    for (int i = 0; i < kMaxBlocks; i++)
        if (g_block_base[i].load() == blk->base) g_block_base[i].store(nullptr);
    munmap(blk->base, blk->cap);
    std::free(blk);
}
void* arena_alloc(size_t n) {
    ThreadArena& a = g_arena;
    Block* b = a.cur;
    n = (n + 15) & ~(size_t)15;
    if (b->lo + n + 16 > b->hi) { g_overflows++; return nullptr; }
    if (a.mode == 1) { void* p = b->base + b->lo; b->lo += n; return p; }
    b->hi -= n;
    return b->base + b->hi;
}
/* allocations made between construction and destruction come from `blk` (or from malloc when the mode is 0) */
struct ArenaScope {
    ArenaScope(Block* blk) {
        g_arena.mode = g_allocator_mode;
        g_arena.cur = blk;
        g_arena.active = g_arena.mode != 0 && blk != nullptr;
    }
    ~ArenaScope() { g_arena.active = false; g_arena.cur = nullptr; }
};
struct ArenaPause {          /* results that must outlive the block are built while the arena is paused */
    bool was;
    ArenaPause() : was(g_arena.active) { g_arena.active = false; }
    ~ArenaPause() { g_arena.active = was; }
};
}  // namespace

/* replaced for this shared object only (linked with -Bsymbolic; the symbols are not exported) */
void* operator new(size_t n) {
    if (g_arena.active) {
        void* p = arena_alloc(n);
        if (p) return p;
    }
    void* p = std::malloc(n ? n : 1);
    if (!p) throw std::bad_alloc();
    return p;
}
void* operator new[](size_t n) { return operator new(n); }
void operator delete(void* p) noexcept { if (p && !in_any_block(p)) std::free(p); }
void operator delete[](void* p) noexcept { operator delete(p); }
void operator delete(void* p, size_t) noexcept { operator delete(p); }
void operator delete[](void* p, size_t) noexcept { operator delete(p); }

namespace {
/* access to the protected members (tables and the per-stage methods); adds no behaviour */
class Probe : public ORB_SLAM2::ORBextractor {
public:
    using ORBextractor::ORBextractor;
    using ORBextractor::ComputeKeyPointsOctTree;
    using ORBextractor::ComputePyramid;
    using ORBextractor::DistributeOctTree;
    using ORBextractor::mnFeaturesPerLevel;
    using ORBextractor::mvScaleFactor;
    using ORBextractor::umax;
    using ORBextractor::nlevels;
};

struct RefExtractor {
    Probe* ex = nullptr;
    std::vector<std::vector<cv::KeyPoint> > levelKeys;      /* per level, level coordinates, list order */
    std::vector<std::vector<orc_corner> > candidates;       /* per level, window coordinates, the order of vToDistributeKeys */
    int retried = 0;
    Block* cur = nullptr;                                   /* holds mvImagePyramid of the last call */
    size_t blockBytes = (size_t)16 << 30;
};
}  // namespace

extern "C" {

/* 0 = process allocator, 1 = ascending addresses (default), 2 = descending addresses */
void ref_set_allocator(int mode) { g_allocator_mode = mode; }
long ref_arena_overflows(void) { return g_overflows.load(); }
void ref_set_gaussian_variant(int variant) { cv::minicv_set_gaussian_variant(variant); }

RefExtractor* ref_extractor_create(int nfeatures, float scale_factor, int nlevels, int ini_th_fast, int min_th_fast) {
    RefExtractor* r = new RefExtractor();
    r->ex = new Probe(nfeatures, scale_factor, nlevels, ini_th_fast, min_th_fast);
    return r;
}
void ref_extractor_destroy(RefExtractor* r) {
    if (!r) return;
    delete r->ex;
    block_destroy(r->cur);
    delete r;
}

/* ORBextractor::operator() (src/ORBextractor.cc:1043-1105) */
int ref_extract(RefExtractor* r, const uint8_t* img, int rows, int cols, size_t step, orc_keypoint* kps, uint8_t* desc, int cap) {
    Block* blk = g_allocator_mode ? block_create(r->blockBytes) : nullptr;
    int n = 0;
    {
        ArenaScope scope(blk);
        cv::Mat image(rows, cols, CV_8UC1, (void*)img, step);
        std::vector<cv::KeyPoint> keys;
        cv::Mat d;
        (*r->ex)(image, cv::Mat(), keys, d);
        n = (int)keys.size();
        for (int i = 0; i < n && i < cap; i++) {
            memcpy(&kps[i], &keys[i], sizeof(orc_keypoint));
            memcpy(desc + (size_t)i * 32, d.ptr(i), 32);
        }
        /* per-level keypoints before the final scaling: the reference's own stage method on the pyramid it just built */
        std::vector<std::vector<cv::KeyPoint> > all;
        std::vector<MinicvFastCall> calls;
        {
            ArenaPause pause;
            calls.reserve(1 << 16);
        }
        cv::minicv_set_fast_log(&calls);
        r->ex->ComputeKeyPointsOctTree(all);
        cv::minicv_set_fast_log(nullptr);
        ArenaPause pause;
        std::vector<std::vector<cv::KeyPoint> > out(all.size());
        for (size_t l = 0; l < all.size(); l++) out[l].assign(all[l].begin(), all[l].end());
        r->levelKeys.swap(out);
        /* vToDistributeKeys of every level (:778-829): a cell contributes the result of its last FAST call (the retry at
         * minThFAST happens only when the first came back empty), shifted by the cell origin inside the level's window */
        std::vector<std::vector<orc_corner> > cand(all.size());
        int retried = 0;
        for (size_t ci = 0; ci < calls.size(); ci++) {
            const MinicvFastCall& c = calls[ci];
            const bool hasRetry = ci + 1 < calls.size() && calls[ci + 1].origin == c.origin && c.corners.empty();
            if (hasRetry) { retried++; continue; }
            for (int l = 0; l < (int)all.size(); l++) {
                const cv::Mat& m = r->ex->mvImagePyramid[l];
                const ptrdiff_t off = c.origin - m.data;
                if (off < 0 || off >= (ptrdiff_t)(m.step * (size_t)m.rows)) continue;
                const int iniY = (int)(off / (ptrdiff_t)m.step), iniX = (int)(off % (ptrdiff_t)m.step);
                for (size_t k = 0; k < c.corners.size(); k++) {
                    orc_corner o = c.corners[k];
                    o.x += iniX - 16; o.y += iniY - 16;             /* minBorderX = minBorderY = EDGE_THRESHOLD - 3 */
                    cand[l].push_back(o);
                }
                break;
            }
        }
        r->candidates.swap(cand);
        r->retried = retried;
    }
    /* the previous call's block only held the previous pyramid, which operator() has replaced */
    block_destroy(r->cur);
    r->cur = blk;
    return n;
}

/* ORBextractor::operator() and nothing else, on the process allocator: the call bench.py --impl reference times
 * (ref_extract above runs the detection stage a second time to expose the per-level intermediates to the tests) */
int ref_extract_plain(RefExtractor* r, const uint8_t* img, int rows, int cols, size_t step, orc_keypoint* kps, uint8_t* desc, int cap) {
    cv::Mat image(rows, cols, CV_8UC1, (void*)img, step);
    std::vector<cv::KeyPoint> keys;
    cv::Mat d;
    (*r->ex)(image, cv::Mat(), keys, d);
    const int n = (int)keys.size();
    for (int i = 0; i < n && i < cap; i++) {
        memcpy(&kps[i], &keys[i], sizeof(orc_keypoint));
        memcpy(desc + (size_t)i * 32, d.ptr(i), 32);
    }
    return n;
}

int ref_extractor_levels(const RefExtractor* r) { return r->ex->nlevels; }
int ref_extractor_quota(const RefExtractor* r, int level) { return r->ex->mnFeaturesPerLevel[level]; }
float ref_extractor_scale(const RefExtractor* r, int level) { return r->ex->mvScaleFactor[level]; }
int ref_extractor_umax(const RefExtractor* r, int v) { return r->ex->umax[v]; }

/* mvImagePyramid[level]: returns the origin of the padded buffer (border 19) like orc_extractor_pyramid */
const uint8_t* ref_extractor_pyramid(const RefExtractor* r, int level, int* w, int* h, size_t* step) {
    const cv::Mat& m = r->ex->mvImagePyramid[level];
    *w = m.cols; *h = m.rows; *step = m.step;
    if (m.empty()) return nullptr;
    return m.data - 19 * (size_t)m.step - 19;
}

int ref_extractor_candidates(const RefExtractor* r, int level, orc_corner* out, int cap) {
    const std::vector<orc_corner>& c = r->candidates[level];
    for (size_t i = 0; i < c.size() && (int)i < cap; i++) out[i] = c[i];
    return (int)c.size();
}
int ref_extractor_retried_cells(const RefExtractor* r) { return r->retried; }

int ref_extractor_level_keypoints(const RefExtractor* r, int level, orc_keypoint* out, int cap) {
    const std::vector<cv::KeyPoint>& k = r->levelKeys[level];
    for (size_t i = 0; i < k.size() && (int)i < cap; i++) memcpy(&out[i], &k[i], sizeof(orc_keypoint));
    return (int)k.size();
}

/* ORBextractor::DistributeOctTree (src/ORBextractor.cc:539-763) on caller-supplied candidates; out_index = position of each
 * selected keypoint in the candidate array (class_id carries it through the reference code untouched) */
int ref_distribute_octree(const orc_corner* cand, int n, int minX, int maxX, int minY, int maxY, int N, int32_t* out_index, int cap) {
    Probe ex(1000, 1.2f, 8, 20, 7);
    Block* blk = g_allocator_mode ? block_create((size_t)4 << 30) : nullptr;
    std::vector<int> idx;
    {
        ArenaScope scope(blk);
        std::vector<cv::KeyPoint> keys(n);
        for (int i = 0; i < n; i++) keys[i] = cv::KeyPoint((float)cand[i].x, (float)cand[i].y, 7.f, -1, (float)cand[i].score, 0, i);
        std::vector<cv::KeyPoint> sel = ex.DistributeOctTree(keys, minX, maxX, minY, maxY, N, 0);
        ArenaPause pause;
        idx.reserve(sel.size());
        for (size_t i = 0; i < sel.size(); i++) idx.push_back(sel[i].class_id);
    }
    block_destroy(blk);
    for (size_t i = 0; i < idx.size() && (int)i < cap; i++) out_index[i] = idx[i];
    return (int)idx.size();
}

}  // extern "C"
