/*
 * extractor_kernels.cuh -- shared declarations of the sm_100a ORB extraction kernels.
 *
 * Device data layout (one "pass" = F frames processed together; F = 128 by default: full waves of CTAs per
 * launch; the 229 MB working set of such a pass does not fit the 126 MB L2 and does not need to -- DESIGN.md):
 *
 *   pyramid   [F][pyrFrameBytes]   per frame, all levels back to back.  A level is stored as
 *                                  (h+38) rows of `step` bytes; the ROI (x=0) starts at byte 32 of a
 *                                  row so that ROI rows are 16-byte aligned; the 19-pixel
 *                                  REFLECT_101 border of the reference's mvImagePyramid occupies
 *                                  bytes [13,32) and [32+w, 32+w+19) of the row.
 *   blur      [F][pyrFrameBytes]   the 7x7 Gaussian blur of every level (the reference's workingMat, :1085-1086), same
 *                                  layout; only ROI pixels are written
 *   cand      [F][candPerFrame]    FAST candidates, packed x:12 | y:12 | score:8 (window coordinates,
 *                                  i.e. level coordinates minus 16), appended with atomics
 *   candCount [F][nlevels]
 *   nodeOf    [F][candPerFrame]    u16 quadtree node of each candidate (octree scratch)
 *   sel       [F][selPerFrame]     selected keypoints per level in reference list order, packed
 *                                  x:12 | y:12 | score:8 in LEVEL coordinates
 *   selCount  [F][nlevels]
 */
#ifndef VIORB_EXTRACTOR_KERNELS_CUH
#define VIORB_EXTRACTOR_KERNELS_CUH

#include <cuda_runtime.h>
#include <stdint.h>

#include "viorb_gpu.h"

#define VIORB_MAX_LEVELS 12
#define VIORB_EDGE 19          /* EDGE_THRESHOLD, src/ORBextractor.cc:74 */
#define VIORB_ROI_X0 32        /* byte offset of ROI x=0 inside a stored row */
#define VIORB_FAST_BORDER 16   /* EDGE_THRESHOLD-3, src/ORBextractor.cc:773 */

struct LevelGeom {
    int w, h;             /* ROI size */
    int step;             /* stored row stride in bytes (multiple of 16) */
    int pyrOff;           /* byte offset of the stored level (row -19) inside a frame's pyramid block */
    int nCols, nRows, wCell, hCell;   /* FAST cell grid, src/ORBextractor.cc:784-787 */
    int cellBase;         /* first cell id of this level inside a frame */
    int quota;            /* mnFeaturesPerLevel */
    int candCap, candBase;
    int selCap, selBase;
    int nIni;             /* root nodes of the quadtree, src/ORBextractor.cc:543 */
    int patchSize;        /* (int)(PATCH_SIZE * mvScaleFactor[level]) */
    float scale;          /* mvScaleFactor[level] */
    int xtab, ytab;       /* offsets into the resize tables: first stored word / first stored row of this level */
    /* blur_levels_kernel: a thread blurs an 8-pixel column strip of a band of blurRows (multiple of 8) output rows;
     * thread blurBase + band * blurStrips + strip of a frame */
    int blurStrips, blurRows, blurBase;
};

struct FrameGeom {
    int nlevels, rows, cols;
    int iniTh, minTh;
    int gaussVariant;        /* GaussianBlur taps: 0 = OpenCV >= 3.4, 1 = OpenCV 2.4 (viorb_extractor_set_gaussian) */
    int cellsPerFrame, candPerFrame, selPerFrame;
    /* fast_cells_kernel shared-memory layout for this geometry: tile rows (max hCell + 6), quads per CTA (max NQ * wh),
     * bytes of the tile + work0 region */
    int fastTileRows, fastMaxWork, fastPixBytes;
    int blurTasks;           /* threads of blur_levels_kernel per frame */
    unsigned long long pyrFrameBytes;
    LevelGeom lv[VIORB_MAX_LEVELS];
};

/* per-level bilinear tables (cv::resize fixed point, 11 fractional bits), laid out for pyr_resize_kernel:
 *   col[3 * (L.xtab + word)]  word = stored 32-bit word of a level row (4 output columns, border and padding included):
 *                             {a0 | a1 << 16 per column} {PRMT selector per column} {lo, hi, okMask, 0}
 *                             lo / hi = first / last source byte (ROI x of level l-1) the four columns read; the selector
 *                             picks (S[sx], S[sx+1]) out of the 8 bytes starting at lo; okMask = 0xff per stored column
 *                             (columns of the alignment padding are written as 0)
 *   row[L.ytab + stored row]  {sy | (sy+1) << 16 (source rows, clamped to the level), b0 << 16, b1 << 16, 0}
 * Border columns / rows carry the coefficients of their REFLECT_101 source coordinate. */
struct ResizeTables {
    const uint4* col;
    const uint4* row;
};

enum {
    VIORB_DEV_CAND_OVERFLOW = 1,
    VIORB_DEV_SEL_OVERFLOW = 2,
    VIORB_DEV_NODE_OVERFLOW = 4,
    VIORB_DEV_OUT_OVERFLOW = 8
};

struct ExtractBuffers {
    uint8_t* pyr;
    uint8_t* blur;        /* GaussianBlur of every level, same layout as pyr (ROI pixels only); NULL: blur per keypoint */
    uint32_t* cand;
    int* candCount;
    uint16_t* nodeOf;
    uint32_t* sel;
    int* selCount;
    int* status;          /* device status word (bit mask above) */
};

/* TMA descriptors (CUtensorMap, 128 bytes each) of the stored pyramid levels of one lane: a level is the 3-D
 * tensor {step bytes, h+38 rows, F frames}; fast_cells_kernel fetches its cell strip with one
 * cp.async.bulk.tensor per CTA (box = VIORB_FAST_TILE_BYTES x (hCell+6) x 1, out-of-range bytes read as 0). */
struct TmaMaps {
    alignas(64) unsigned char fast[VIORB_MAX_LEVELS][128];
};
#define VIORB_FAST_TILE_BYTES 208   /* box width: <= 15 alignment bytes + 3 rim + 180 window + 3 rim, padded to 16 bytes */
#define VIORB_FAST_TILE_ROWS 68     /* box height bound: hCell + 6 */
/* encodes the maps for a pyramid buffer holding F frames (host, driver entry point cuTensorMapEncodeTiled) */
int viorb_encode_tma_maps(const FrameGeom& g, uint8_t* d_pyr, int F, TmaMaps* out);

/* launchers (extractor_kernels.cu); every launcher returns the number of kernel launches issued.  `pdl`: which launches carry
 * the programmatic-dependent-launch attribute (the kernels of a pass form one dependency chain; with the attribute a kernel's
 * CTAs are scheduled, and run their set-up, while the tail of its predecessor is still executing) */
#define VIORB_PDL_INNER 1       /* between the kernels of one stage (pyramid levels, FAST shift classes) */
#define VIORB_PDL_EDGE 2        /* on the first kernel of FAST / quadtree / describe */
int viorb_launch_pyramid(const FrameGeom& g, const ResizeTables& t, const uint8_t* d_images, size_t step,
                         size_t frameStride, int F, const ExtractBuffers& b, cudaStream_t s, int pdl);
/* d_groups is sorted by the byte shift SH = (stored byte of the group's window x=0) & 3; class sh owns
 * [classStart[sh], classStart[sh+1]) and is one launch of the kernel instantiated for that shift */
int viorb_launch_fast(const FrameGeom& g, const TmaMaps& maps, const int4* d_groups, const int* classStart, int F,
                      const ExtractBuffers& b, cudaStream_t s, int pdl);
size_t viorb_fast_smem_bytes(const FrameGeom& g);
int viorb_fast_prepare(const FrameGeom& g);   /* opt-in dynamic shared memory; returns cudaError */
#define VIORB_FAST_GROUP 4      /* horizontally adjacent FAST cells per CTA (extractor_kernels.cu FAST_GROUP) */
int viorb_launch_octree(const FrameGeom& g, int F, const ExtractBuffers& b, int nodeCap, cudaStream_t s, int pdl);
int viorb_launch_orientation_sweep(const int* d_m01, const int* d_m10, long long n, float* d_deg, int sms, cudaStream_t s);
int viorb_launch_steering_sweep(unsigned firstBits, long long n, float* d_sin, float* d_cos, int sms, cudaStream_t s);
/* describe / blur work on frames [frame0, frame0 + F) of the pass (sub-batches keep the blurred levels in the L2) */
int viorb_launch_describe(const FrameGeom& g, int frame0, int F, const ExtractBuffers& b, viorb_keypoint* d_kps,
                          uint8_t* d_desc, int cap, int32_t* d_counts, cudaStream_t s, int pdl);
int viorb_launch_blur(const FrameGeom& g, int frame0, int F, const ExtractBuffers& b, cudaStream_t s, int pdl);
size_t viorb_octree_smem_bytes(int nodeCap);
int viorb_octree_prepare(int nodeCap);   /* opt-in dynamic shared memory; returns cudaError */

#endif
