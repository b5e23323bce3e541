/*
 * matcher_kernels.cuh -- launchers of the sm_100a Hamming-matching kernels (matcher_kernels.cu).
 */
#ifndef VIORB_MATCHER_KERNELS_CUH
#define VIORB_MATCHER_KERNELS_CUH

#include <cuda_runtime.h>
#include <stdint.h>

#include "viorb_gpu.h"

/* brute-force top-2: returns launches issued; d_partials must hold viorb_top2_partial_count(Q,M)*Q records */
int viorb_top2_slices(int Q, int64_t M, int sms);
int viorb_launch_hamming_top2(const uint8_t* d_q, int Q, const uint8_t* d_map, int64_t M, int64_t indexBase,
                              viorb_top2* d_partials, int nslices, viorb_top2* d_out, cudaStream_t s);
int viorb_launch_top2_merge(const viorb_top2* d_parts, int nparts, int Q, viorb_top2* d_out, cudaStream_t s);
int viorb_launch_descriptor_distance(const uint8_t* d_a, const uint8_t* d_b, int n, int32_t* d_dist, cudaStream_t s);

/* Frame::ComputeStereoMatches (src/Frame.cc:646-820) */
struct StereoLevel {
    const uint8_t* roiL;
    const uint8_t* roiR;
    int w, h, stepL, stepR;
};
struct StereoParams {
    int nlevels, nl, nr, nRows;
    float mbf, mb;
    float scale[12], invScale[12];
    StereoLevel lv[12];
};
int viorb_launch_stereo(const StereoParams& p, const viorb_keypoint* d_kl, const uint8_t* d_dl,
                        const viorb_keypoint* d_kr, const uint8_t* d_dr, int* d_rowStart, int* d_rowItems,
                        int* d_scratch, float* d_uRight, float* d_depth, int* d_sad, cudaStream_t s);

/* 64x48 frame grid + windowed searches */
struct FrameIndexDev {
    const viorb_keypoint* kps;
    const uint8_t* desc;
    const float* uRight;
    const int* cellStart;     /* [64*48+1] CSR, cell id = ix*48 + iy */
    const int* cellItems;
    int n;
    float minX, maxX, minY, maxY, invW, invH;
    float scale[12];
    int nlevels;
};


/* Frame::UndistortKeyPoints / ComputeImageBounds (src/Frame.cc:584-645): cv::undistortPoints(mat, mat, mK, mDistCoef,
 * cv::Mat(), mK) in double precision, five fixed-point iterations */
struct UndistortParams {
    double fx, fy, cx, cy, ifx, ify;
    double k[12];            /* k1 k2 p1 p2 k3 k4 k5 k6 s1 s2 s3 s4 (OpenCV order), zero padded */
    int active;              /* mDistCoef.at<float>(0) != 0.0 (:586) */
};
int viorb_launch_undistort(const UndistortParams& p, const viorb_keypoint* d_in, int n, viorb_keypoint* d_out, cudaStream_t s);
int viorb_launch_image_bounds(const UndistortParams& p, int cols, int rows, float* d_bounds, cudaStream_t s);

#endif
