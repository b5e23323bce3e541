/*
 * orb_oracle.h -- C ABI of the CPU oracle.  TEST INFRASTRUCTURE ONLY.
 *
 * This library is a dependency-free CPU restatement of the ORB front-end of sta105/VIORB
 * (ORBextractor / ORBmatcher / Frame::ComputeStereoMatches) and of the OpenCV primitives those
 * files import.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
 * reference legs may load it.  The product path (libviorb_b200.so) never links or calls it.
 *
 * Parity status: the reference ships NO golden vectors for this path and
 * cannot be compiled here (no OpenCV C++ headers), so parity is pinned on the *third-party
 * arithmetic* instead: every primitive below is checked bit-for-bit against OpenCV 4.13 (python
 * cv2) live in tests/test_oracle_vs_cv2.py and against committed fixtures in tests/golden/.
 * Where the reference itself is under-determined the oracle fixes a convention (DESIGN.md
 * "Conventions"): octree tie rule, sincosf = glibc flt-32 algorithm without FMA, OpenCV >= 3.4
 * Gaussian taps, no FMA contraction anywhere.
 */
#ifndef ORB_ORACLE_H
#define ORB_ORACLE_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* cv::KeyPoint-compatible POD (28 bytes). */
typedef struct {
    float x, y, size, angle, response;
    int32_t octave, class_id;
} orc_keypoint;

/* FAST candidate (cell-relative or window-relative coordinates, integer valued). */
typedef struct {
    int32_t x, y, score;
} orc_corner;

/* ---- OpenCV primitives (Appendix A of SURVEY.md) ---- */
void orc_resize_linear_u8(const uint8_t* src, int sw, int sh, size_t sstep,
                          uint8_t* dst, int dw, int dh, size_t dstep);
void orc_border_reflect101_u8(const uint8_t* src, int w, int h, size_t sstep,
                              uint8_t* dst, size_t dstep, int border);
int  orc_fast9_16(const uint8_t* img, int w, int h, size_t step, int threshold, int nms,
                  orc_corner* out, int cap);
void orc_gaussian7_u8(const uint8_t* src, int w, int h, size_t sstep, uint8_t* dst, size_t dstep);
/* variant 0 = OpenCV >= 3.4 taps [18,34,48,56,48,34,18], 1 = OpenCV 2.4 taps [18,34,49,55,49,34,18] */
void orc_gaussian7_u8_variant(const uint8_t* src, int w, int h, size_t sstep, uint8_t* dst, size_t dstep, int variant);
float orc_fast_atan2(float y, float x);
void orc_sincosf(float x, float* s, float* c);
void orc_orientation_sweep(const int32_t* m01, const int32_t* m10, int64_t n, float* deg, int threads);
int64_t orc_steering_sweep(uint32_t first_bits, int64_t n, int which, float* sin_out, float* cos_out, int threads);
int  orc_cv_round_f(float v);

/* ---- ORBextractor ---- */
typedef struct orc_extractor orc_extractor;
orc_extractor* orc_extractor_create(int nfeatures, float scale_factor, int nlevels,
                                    int ini_th_fast, int min_th_fast);
void orc_extractor_destroy(orc_extractor*);
/* GaussianBlur taps of ORBextractor.cc:1086: 0 = OpenCV >= 3.4 (default), 1 = OpenCV 2.4 */
void orc_extractor_set_gaussian(orc_extractor*, int variant);
/* runs operator(); returns number of keypoints (<= cap written), or <0 on error */
int orc_extract(orc_extractor*, const uint8_t* img, int rows, int cols, size_t step,
                orc_keypoint* kps, uint8_t* desc, int cap);
/* tables / intermediate state of the last orc_extract call */
int   orc_extractor_levels(const orc_extractor*);
int   orc_extractor_quota(const orc_extractor*, int level);
float orc_extractor_scale(const orc_extractor*, int level);
int   orc_extractor_umax(const orc_extractor*, int v);
/* padded pyramid level: pointer to the padded buffer (border 19), ROI size returned */
const uint8_t* orc_extractor_pyramid(const orc_extractor*, int level, int* w, int* h, size_t* step);
const uint8_t* orc_extractor_blurred(const orc_extractor*, int level, int* w, int* h, size_t* step);
/* FAST candidates of a level, window coordinates (before adding minBorder), in reference order */
int orc_extractor_candidates(const orc_extractor*, int level, orc_corner* out, int cap);
/* selected keypoints of a level in *level* coordinates, in reference list order */
int orc_extractor_level_keypoints(const orc_extractor*, int level, orc_keypoint* out, int cap);
/* per-stage wall time of the last call, seconds: pyramid, fast, octree, orient, blur, desc */
void orc_extractor_stage_seconds(const orc_extractor*, double out[6]);
/* number of cell retries (minThFAST) of the last call */
int orc_extractor_retried_cells(const orc_extractor*);

/* stand-alone DistributeOctTree on window-coordinate candidates (testing the octree in isolation) */
int orc_distribute_octree(const orc_corner* cand, int n, int minX, int maxX, int minY, int maxY,
                          int N, int32_t* out_index, int cap);

/* ---- matcher ---- */
int orc_descriptor_distance(const uint8_t* a, const uint8_t* b);           /* bit-hack, ORBmatcher.cc:1648 */
int orc_descriptor_distance_popcnt(const uint8_t* a, const uint8_t* b);    /* same value, hw popcount */

/* top-2 record */
typedef struct {
    int32_t d1, i1, d2, i2;
} orc_top2;
/* brute force: sequential strict-< scan, semantics of ORBmatcher.cc:201-226 */
void orc_hamming_top2(const uint8_t* q, int Q, const uint8_t* map, int64_t M, int64_t index_base,
                      orc_top2* out, int use_popcnt, int nthreads);
void orc_top2_merge(const orc_top2* parts, int nparts, int Q, orc_top2* out);

/* Frame::ComputeStereoMatches (Frame.cc:646-820).  Pyramid pointers are ROI origins. */
typedef struct {
    const uint8_t* data;
    int32_t w, h;
    size_t step;
} orc_image;
int orc_stereo_match(const orc_keypoint* kl, const uint8_t* dl, int nl,
                     const orc_keypoint* kr, const uint8_t* dr, int nr,
                     const orc_image* pyr_l, const orc_image* pyr_r, int nlevels,
                     const float* scale_factors, const float* inv_scale_factors,
                     float mbf, float mb, float* u_right, float* depth,
                     int32_t* best_dist_out /* optional, nl */, int32_t* best_idx_out /* optional */);

/* 64x48 frame grid (Frame.cc:410-425,507-572) + windowed searches */
typedef struct orc_grid orc_grid;
orc_grid* orc_grid_create(const orc_keypoint* kps_un, int n, float minx, float maxx, float miny, float maxy);
void orc_grid_destroy(orc_grid*);
int orc_grid_features_in_area(const orc_grid*, float x, float y, float r, int min_level, int max_level,
                              int32_t* out, int cap);

/* ORBmatcher::SearchByProjection(Frame&, vector<MapPoint*>&, th)  (ORBmatcher.cc:45-129)
 * inputs are the per-map-point fields the reference reads; frame_mp_obs[i] > 0 means the frame
 * keypoint i already holds a map point with observations.  match_out[i] = map point index or -1
 * (pre-filled by the caller with the existing assignment is NOT required: entries are only written
 * on a match).  Returns nmatches. */
int orc_search_by_projection_local(const orc_grid* grid, const orc_keypoint* kps_un, const uint8_t* fdesc,
                                   const float* fu_right, int32_t* frame_mp_obs, int nf,
                                   const float* scale_factors,
                                   const float* proj_x, const float* proj_y, const float* proj_xr,
                                   const int32_t* pred_level, const float* view_cos,
                                   const uint8_t* valid, const int32_t* mp_nobs, const uint8_t* mpdesc, int nmp,
                                   float th, float nnratio, int32_t* match_out);

/* ORBmatcher::SearchByProjection(Frame& Cur, const Frame& Last, th, bMono) (ORBmatcher.cc:1328-1471)
 * after projection: u,v,invz per last-frame point are given (valid[i]=0 => skipped).  mode: 0 = +-1
 * level window, 1 = forward (>= octave), 2 = backward (0..octave). */
int orc_search_by_projection_frame(const orc_grid* grid, const orc_keypoint* kps_un, const uint8_t* fdesc,
                                   const float* fu_right, int32_t* frame_mp_obs, int nf,
                                   const float* scale_factors,
                                   const float* u, const float* v, const float* invz,
                                   const int32_t* last_octave, const float* last_angle,
                                   const uint8_t* valid, const int32_t* mp_nobs, const uint8_t* mpdesc, int nlast,
                                   float th, float mbf, int mode, int check_ori, int th_high,
                                   int32_t* match_out /* nf: index into last or -1 */);

/* ORBmatcher::SearchForTriangulation (ORBmatcher.cc:657-823) on flattened DBoW2 feature vectors:
 * node_id{1,2} ascending node ids, node_ptr{1,2} CSR offsets (nn+1) into idx{1,2}. */
int orc_search_for_triangulation(const orc_keypoint* k1, const uint8_t* d1, const float* ur1,
                                 const uint8_t* has_mp1, int n1,
                                 const orc_keypoint* k2, const uint8_t* d2, const float* ur2,
                                 const uint8_t* has_mp2, int n2,
                                 const int32_t* node_id1, const int32_t* node_ptr1, const int32_t* idx1, int nn1,
                                 const int32_t* node_id2, const int32_t* node_ptr2, const int32_t* idx2, int nn2,
                                 const float* F12 /* 9, row-major */, float ex, float ey,
                                 const float* scale_factors2, const float* level_sigma2_2,
                                 int only_stereo, int check_ori, int32_t* matches12 /* n1 */);

/* synthetic-input independent helpers */
/* MapPoint::ComputeDistinctiveDescriptors, MapPoint.cc:249-314 (one point) */
int orc_distinctive_descriptor(const uint8_t* desc, int N, int* median_out);
/* ORBmatcher::SearchByBoW x2 (ORBmatcher.cc:159-288, :522-655) and SearchForInitialization (:405-520) */
int orc_search_by_bow(int mode, const orc_keypoint* k1, const uint8_t* d1, const uint8_t* valid1, int n1,
                      const orc_keypoint* k2, const uint8_t* d2, const uint8_t* valid2, int n2,
                      const int32_t* node_id1, const int32_t* node_ptr1, const int32_t* idx1v, int nn1,
                      const int32_t* node_id2, const int32_t* node_ptr2, const int32_t* idx2v, int nn2,
                      float mfNNratio, int check_ori, int32_t* match);
int orc_search_for_initialization(const orc_grid* grid2, const orc_keypoint* k2, const uint8_t* d2, int n2,
                                  const orc_keypoint* k1, const uint8_t* d1, int n1, float* prev,
                                  int windowSize, float mfNNratio, int check_ori, int32_t* vnMatches12);
/* search loops of Fuse x2 (ORBmatcher.cc:825-1100) and SearchBySim3 (:1102-1326) */
void orc_search_window_top1(const orc_grid* grid, const orc_keypoint* kps_un, const uint8_t* kdesc, const float* mvuRight,
                            const float* scale_factors, const float* u, const float* v, const float* ur,
                            const int32_t* pred_level, const uint8_t* valid, const uint8_t* mp_desc, int n, float th,
                            int th_dist, const float* inv_level_sigma2, int32_t* best_idx, int32_t* best_dist);
int orc_search_by_sim3(const orc_grid* g1, const orc_keypoint* k1, const uint8_t* d1, const float* sf1, int n1,
                       const orc_grid* g2, const orc_keypoint* k2, const uint8_t* d2, const float* sf2, int n2,
                       const float* u12, const float* v12, const int32_t* level12, const uint8_t* valid12,
                       const uint8_t* mp_desc1, const float* u21, const float* v21, const int32_t* level21,
                       const uint8_t* valid21, const uint8_t* mp_desc2, float th, int32_t* match12);
/* Frame::UndistortKeyPoints / ComputeImageBounds (Frame.cc:584-645) */
void orc_undistort_keypoints(const orc_keypoint* kps, int n, float fx, float fy, float cx, float cy,
                             const float* dist, int ndist, orc_keypoint* out);
void orc_compute_image_bounds(int cols, int rows, float fx, float fy, float cx, float cy, const float* dist,
                              int ndist, float* bounds);
/* DBoW2 vocabulary transform (bow_oracle.cpp) */
typedef struct orc_vocabulary orc_vocabulary;
orc_vocabulary* orc_vocabulary_create(int k, int L, int weighting, int scoring, int nnodes, const int32_t* parent,
                                      const uint8_t* desc, const double* weight);
void orc_vocabulary_destroy(orc_vocabulary* v);
int orc_bow_transform(const orc_vocabulary* voc, const uint8_t* desc, int n, int levelsup, int32_t* bow_ids,
                      double* bow_values, int32_t* fv_node, int32_t* fv_ptr, int32_t* fv_idx, int* nfv,
                      int32_t* word_of, int32_t* node_of);
int orc_num_threads(void);

#ifdef __cplusplus
}
#endif
#endif
