/*
 * viorb_gpu.h -- C ABI of libviorb_b200.so: the B200-native (sm_100a) ORB feature front-end for VIORB.
 *
 * This is the drop-in boundary.  The reference (sta105/VIORB, an ORB-SLAM2 fork) has no FFI layer; its
 * boundary is the C++ ABI of ORBextractor / ORBmatcher / Frame::ComputeStereoMatches.  Each entry point
 * below cites the reference interface it replaces; the C++ shims in viorb_b200/host/ re-create those
 * classes on top of this ABI.
 *
 * Conventions: plain pointers and sizes only; every function returns VIORB_OK (0) or a negative
 * viorb_status and never throws; outputs are caller allocated; a context owns one CUDA stream and must
 * not be entered concurrently (use one context per calling thread, as the reference uses one
 * ORBextractor instance per thread -- src/Frame.cc:258-261).  There is NO CPU fallback: every entry
 * point fails with VIORB_ERR_CUDA when no sm_100 device is usable.
 */
#ifndef VIORB_GPU_H
#define VIORB_GPU_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef enum {
    VIORB_OK = 0,
    VIORB_ERR_INVALID = -1,       /* bad argument */
    VIORB_ERR_CUDA = -2,          /* CUDA runtime / launch failure, or no device */
    VIORB_ERR_CAPACITY = -3,      /* caller buffer (cap) or an internal candidate pool too small */
    VIORB_ERR_UNSUPPORTED = -4    /* parameters outside the supported envelope */
} viorb_status;

/* Layout-identical to cv::KeyPoint (28 bytes): pt.x, pt.y, size, angle, response, octave, class_id. */
typedef struct {
    float x, y, size, angle, response;
    int32_t octave, class_id;
} viorb_keypoint;

/* Best and second-best Hamming match of one query: distances in [0,256], map indices (-1 = none).
 * Semantics of the sequential strict-< scan in src/ORBmatcher.cc:201-226: ties keep the lowest index. */
typedef struct {
    int32_t d1, i1, d2, i2;
} viorb_top2;

typedef struct viorb_ctx viorb_ctx;
typedef struct viorb_extractor viorb_extractor;

/* ---- context ---------------------------------------------------------------------------------- */
/* stream: a cudaStream_t owned by the caller (e.g. torch's current stream), or NULL to create one. */
int viorb_ctx_create(int device, void* stream, viorb_ctx** out);
int viorb_ctx_destroy(viorb_ctx* ctx);
int viorb_ctx_synchronize(viorb_ctx* ctx);
const char* viorb_last_error(void);                  /* thread-local message of the last failure */
int viorb_device_count(void);
/* number of kernel launches issued through this context so far (bench.py "gpu_launches") */
int64_t viorb_ctx_launch_count(const viorb_ctx* ctx);
/* pinned host memory for fast host<->device staging of batches */
int viorb_host_alloc(size_t bytes, void** out);
int viorb_host_free(void* p);

/* ---- ORBextractor ------------------------------------------------------------------------------
 * replaces ORB_SLAM2::ORBextractor::ORBextractor(int nfeatures, float scaleFactor, int nlevels,
 *          int iniThFAST, int minThFAST)                         include/ORBextractor.h:52-53,
 *                                                                 src/ORBextractor.cc:410-470        */
int viorb_extractor_create(viorb_ctx* ctx, int nfeatures, float scale_factor, int nlevels,
                           int ini_th_fast, int min_th_fast, viorb_extractor** out);
int viorb_extractor_destroy(viorb_extractor* ex);
/* optional tuning: frames processed per device pass (default 128 EuRoC-size frames = 229 MB of pyramid + lists per pass, 412 MB with the blurred levels,
 * more than the 126 MB L2: the kernels are bound by integer issue, not by DRAM, so the pass is sized for full waves of
 * CTAs per launch rather than for L2 residency -- measured 2.1x the compulsory DRAM bytes at 5 % of the HBM peak; the
 * host-buffer batch call, which is bound by its input copies, uses 48, down to 32 for short batches, so that the first
 * pass starts early and the last one is short) and the
 * candidate pool per level as a fraction 1/div of the level's pixel count. */
int viorb_extractor_configure(viorb_extractor* ex, int chunk_frames, int cand_div);
/* How viorb_extract_batch schedules its host<->device copies.  VIORB_COPY_DUPLEX (default): input and output copies on two
 * streams, both directions of the link busy at once -- the faster order on a host that sustains both (one GPU of this pool:
 * 27.3 ms for a 4096-frame step against 26.7 + 5.0 ms one after the other).  VIORB_COPY_SERIAL: the output copies queue behind
 * the input copies on one stream, for hosts where concurrent device-to-host writes slow the host-to-device reads down by
 * more than they take on their own (all four GPUs of a 4-GPU box of this pool copying at once: 9.8 ms duplex, 6.8 + 1.3 ms
 * serial).  bench.py measures both and picks; DESIGN.md section 6. */
enum { VIORB_COPY_DUPLEX = 0, VIORB_COPY_SERIAL = 1 };
int viorb_extractor_set_copy_mode(viorb_extractor* ex, int mode);
/* Which OpenCV the GaussianBlur(workingMat, ..., Size(7,7), 2, 2, BORDER_REFLECT_101) of src/ORBextractor.cc:1086 is:
 * the 8-bit kernel differs between releases and the reference does not vendor OpenCV.
 *   VIORB_GAUSSIAN_OPENCV4  (default) OpenCV >= 3.4 fixed point [18,34,48,56,48,34,18]/256 -- bit-equal to cv2 4.13;
 *   VIORB_GAUSSIAN_OPENCV24 OpenCV 2.4.x, the version the reference pins (CMakeLists.txt:31, README.md:58) and its
 *                           shipped libORB_SLAM2.so links: [18,34,49,55,49,34,18]/256, saturating.
 * Keypoints are unaffected (the blur only feeds the descriptor tests); descriptors differ in a few bits. */
enum { VIORB_GAUSSIAN_OPENCV4 = 0, VIORB_GAUSSIAN_OPENCV24 = 1 };
int viorb_extractor_set_gaussian(viorb_extractor* ex, int opencv_variant);
/* Where the GaussianBlur of src/ORBextractor.cc:1085-1086 is evaluated (results are identical -- the fixed-point blur is
 * exact in any order):
 *   VIORB_DESCRIBE_AUTO   (default) whole levels when the pyramid has fewer than about 2300 pixels per feature and the
 *                         pass holds more than 8 frames, per keypoint otherwise;
 *   VIORB_DESCRIBE_FUSED  on the 43 x 43 neighbourhood of every selected keypoint, inside the descriptor kernel;
 *   VIORB_DESCRIBE_LEVELS every level once (the reference's own order: workingMat, then computeDescriptors). */
enum { VIORB_DESCRIBE_AUTO = 0, VIORB_DESCRIBE_FUSED = 1, VIORB_DESCRIBE_LEVELS = 2 };
int viorb_extractor_set_describe_mode(viorb_extractor* ex, int mode);

/* per-stage device timing (CUDA events on the context stream around each stage of every pass):
 * ms[0..3] = pyramid, FAST, quadtree, orient+describe, summed over `passes` passes since the last query. */
int viorb_extractor_profile(viorb_extractor* ex, int enable);
int viorb_extractor_stage_ms(viorb_extractor* ex, float ms[4], int* passes);

/* GetLevels/GetScaleFactors/GetInverseScaleFactors/GetScaleSigmaSquares/GetInverseScaleSigmaSquares
 * (include/ORBextractor.h:63-83) + mnFeaturesPerLevel; any pointer may be NULL. */
int viorb_extractor_tables(const viorb_extractor* ex, int* nlevels, float* scale, float* inv_scale,
                           float* sigma2, float* inv_sigma2, int* features_per_level);

/* replaces ORBextractor::operator()(InputArray image, InputArray mask, vector<KeyPoint>&, OutputArray)
 *          include/ORBextractor.h:58-60, src/ORBextractor.cc:1043-1105.
 * image: CV_8UC1, rows x cols, row stride `step` bytes, HOST memory (mask is ignored by the reference).
 * kps/desc: host, capacity `cap` keypoints / cap*32 bytes; *n = keypoints found.
 * An empty image (NULL / 0x0) returns VIORB_OK with *n = 0 (src/ORBextractor.cc:1046-1047).        */
int viorb_extract(viorb_extractor* ex, const uint8_t* image, int rows, int cols, size_t step,
                  viorb_keypoint* kps, uint8_t* desc, int cap, int* n);

/* The same operator over a batch of B equally sized frames (frame b at images + b*frame_stride),
 * HOST buffers (pinned memory recommended): kps[b*cap ...], desc[b*cap*32 ...], counts[b].
 * Host<->device copies are part of the call.                                                       */
int viorb_extract_batch(viorb_extractor* ex, const uint8_t* images, int B, int rows, int cols,
                        size_t step, size_t frame_stride, viorb_keypoint* kps, uint8_t* desc, int cap,
                        int32_t* counts);

/* Batch over DEVICE-resident frames and outputs; asynchronous on the context stream.  Errors detected
 * on the device (pool overflow) are reported by the next viorb_extractor_check(). */
int viorb_extract_batch_device(viorb_extractor* ex, const uint8_t* d_images, int B, int rows, int cols,
                               size_t step, size_t frame_stride, viorb_keypoint* d_kps, uint8_t* d_desc,
                               int cap, int32_t* d_counts);
int viorb_extractor_check(viorb_extractor* ex);      /* synchronises, returns deferred device status */

/* ORBextractor::mvImagePyramid (include/ORBextractor.h:85; read by src/Frame.cc:653,743,755,760).
 * The pyramids of the frames of the most recent device pass stay resident; `frame` indexes into it
 * (0 for viorb_extract).  Geometry: ROI w x h; the padded image is (w+38) x (h+38) (EDGE_THRESHOLD 19). */
int viorb_extractor_pyramid_info(const viorb_extractor* ex, int level, int* w, int* h);
int viorb_extractor_pyramid_download(viorb_extractor* ex, int frame, int level, uint8_t* dst_padded,
                                     size_t dst_step);
/* all levels of one frame in one call: a single device-to-host copy of the frame's pyramid block into a pinned staging
 * buffer, then row copies into dst_padded[l] ((w_l+38) x (h_l+38), row stride dst_step[l]) -- what a caller that reads the
 * whole mvImagePyramid after operator() wants (eight synchronous pageable copies otherwise) */
int viorb_extractor_pyramid_download_all(viorb_extractor* ex, int frame, uint8_t* const* dst_padded,
                                         const size_t* dst_step, int nlevels);
/* device view of the ROI origin of (frame, level) and its row stride (for device-side consumers) */
int viorb_extractor_pyramid_device(const viorb_extractor* ex, int frame, int level,
                                   const uint8_t** d_roi, size_t* d_step);
/* how many frames of the last call are still resident, and the index of the first of them */
int viorb_extractor_resident(const viorb_extractor* ex, int* first_frame, int* nframes);

/* parity hooks for the stage tests: FAST candidates (x, y in level coordinates, score) of (frame, level)
 * in no particular order, and the quadtree-selected keypoints of a level in reference list order. */
int viorb_extractor_debug_candidates(viorb_extractor* ex, int frame, int level, int32_t* xys, int cap, int* n);
int viorb_extractor_debug_selected(viorb_extractor* ex, int frame, int level, int32_t* xys, int cap, int* n);
/* parity hook for IC_Angle's cv::fastAtan2((float)m_01, (float)m_10) (src/ORBextractor.cc:103): degrees in [0, 360)
 * for n pairs of integer patch moments (n <= 2^28 per call); host buffers. */
int viorb_debug_orientation(viorb_ctx* ctx, const int32_t* m01, const int32_t* m10, int64_t n, float* deg);
/* parity hook for the steering of computeOrbDescriptor (src/ORBextractor.cc:107-113): b = sinf, a = cosf of
 * angle_deg * (float)(CV_PI/180.f) for the n consecutive float bit patterns first_bits, first_bits+1, ... taken as
 * keypoint angles in degrees (n <= 2^28 per call); host outputs. */
int viorb_debug_steering(viorb_ctx* ctx, uint32_t first_bits, int64_t n, float* sin_out, float* cos_out);

/* ---- ORBmatcher --------------------------------------------------------------------------------
 * replaces static int ORBmatcher::DescriptorDistance(const cv::Mat&, const cv::Mat&)
 *          include/ORBmatcher.h:47, src/ORBmatcher.cc:1648-1664 -- n pairs at once, host buffers.     */
int viorb_descriptor_distance(viorb_ctx* ctx, const uint8_t* a, const uint8_t* b, int n, int32_t* dist);

/* Brute-force top-2 Hamming search of Q query descriptors against M map descriptors (32 B each);
 * inner loop semantics of src/ORBmatcher.cc:201-226 (SearchByBoW) / :76-115 (SearchByProjection):
 * best1 = best2 = 256, strict <, first index wins.  index_base is added to the reported indices
 * (map shard offset for multi-GPU sharding).  Host buffers; copies are part of the call.           */
int viorb_hamming_top2(viorb_ctx* ctx, const uint8_t* queries, int Q, const uint8_t* map, int64_t M,
                       int64_t index_base, viorb_top2* out);
/* device-resident variant, asynchronous on the context stream */
int viorb_hamming_top2_device(viorb_ctx* ctx, const uint8_t* d_queries, int Q, const uint8_t* d_map,
                              int64_t M, int64_t index_base, viorb_top2* d_out);
/* deterministic merge of `nparts` per-shard records per query ([part][Q] layout, device memory), used
 * after the all-gather of the multi-GPU search; identical to one scan over the concatenated map.   */
int viorb_top2_merge_device(viorb_ctx* ctx, const viorb_top2* d_parts, int nparts, int Q, viorb_top2* d_out);

/* replaces void Frame::ComputeStereoMatches()           include/Frame.h:127, src/Frame.cc:646-820.
 * Uses the pyramids resident in the two extractors (frame index `frame_l`/`frame_r` of their last pass)
 * exactly as the reference reads mpORBextractorLeft/Right->mvImagePyramid.  Host buffers.          */
int viorb_stereo_match(viorb_extractor* left, int frame_l, viorb_extractor* right, int frame_r,
                       const viorb_keypoint* kps_l, const uint8_t* desc_l, int nl,
                       const viorb_keypoint* kps_r, const uint8_t* desc_r, int nr,
                       float mbf, float mb, float* u_right, float* depth);

/* Windowed projection searches.  The caller flattens the fields the reference reads from Frame /
 * MapPoint / KeyFrame into arrays (INTEGRATION.md shows the marshalling); results are identical to the
 * sequential loops including the "keypoint already taken" dependence.
 *
 * viorb_frame_index: device-side copy of a Frame's matching state: undistorted keypoints, descriptors,
 * mvuRight and the 64x48 grid of Frame::AssignFeaturesToGrid (src/Frame.cc:410-425).               */
typedef struct viorb_frame_index viorb_frame_index;
int viorb_frame_index_create(viorb_ctx* ctx, const viorb_keypoint* kps_un, const uint8_t* desc,
                             const float* u_right /* may be NULL: all -1 */, int n,
                             float min_x, float max_x, float min_y, float max_y,
                             const float* scale_factors, int nlevels, viorb_frame_index** out);
/* The same index from the RAW keypoints of ORBextractor::operator(): Frame::UndistortKeyPoints (src/Frame.cc:584-614,
 * cv::undistortPoints(mat, mat, mK, mDistCoef, cv::Mat(), mK)), Frame::ComputeImageBounds (:616-645) and
 * AssignFeaturesToGrid run on the device; keypoints do not return to the host between extraction and matching
 * (SURVEY.md 8(f) F3).  dist_coef = mDistCoef (k1 k2 p1 p2 [k3 [k4 k5 k6 [s1 s2 s3 s4]]]), ndist in {0,4,5,8,12};
 * cols, rows = image size for the bounds.                                                                  */
int viorb_frame_index_create_distorted(viorb_ctx* ctx, const viorb_keypoint* kps, const uint8_t* desc,
                                       const float* u_right, int n, float fx, float fy, float cx, float cy,
                                       const float* dist_coef, int ndist, int cols, int rows,
                                       const float* scale_factors, int nlevels, viorb_frame_index** out);
/* The same from DEVICE-resident raw keypoints / descriptors (one frame's slice of the viorb_extract_batch_device
 * outputs): extraction -> undistortion -> grid -> matching without the keypoints ever visiting the host.
 * d_u_right may be NULL (monocular: all -1).                                                                     */
int viorb_frame_index_create_device(viorb_ctx* ctx, const viorb_keypoint* d_kps, const uint8_t* d_desc,
                                    const float* d_u_right, int n, float fx, float fy, float cx, float cy,
                                    const float* dist_coef, int ndist, int cols, int rows,
                                    const float* scale_factors, int nlevels, viorb_frame_index** out);
/* mvKeysUn (may be NULL) and {mnMinX, mnMaxX, mnMinY, mnMaxY} (may be NULL) of an index */
int viorb_frame_index_keys(viorb_frame_index* fi, viorb_keypoint* kps_un, float bounds[4]);
/* Frame::mGrid (include/Frame.h:191) as built by AssignFeaturesToGrid (src/Frame.cc:410-425): CSR over the 64 x 48
 * cells, cell id = ix * 48 + iy; cell_start has 64*48 + 1 entries, cell_items (capacity n) lists the keypoint indices
 * of every cell in ascending order (the reference pushes them in keypoint order).                               */
int viorb_frame_index_grid(viorb_frame_index* fi, int32_t* cell_start, int32_t* cell_items);
/* stand-alone forms of the two Frame members, host buffers */
int viorb_undistort_keypoints(viorb_ctx* ctx, const viorb_keypoint* kps, int n, float fx, float fy, float cx, float cy,
                              const float* dist_coef, int ndist, viorb_keypoint* kps_un);
int viorb_compute_image_bounds(viorb_ctx* ctx, int cols, int rows, float fx, float fy, float cx, float cy,
                               const float* dist_coef, int ndist, float bounds[4]);
int viorb_frame_index_destroy(viorb_frame_index* fi);
/* Frame::GetFeaturesInArea (src/Frame.cc:507-560), reference enumeration order */
int viorb_frame_features_in_area(viorb_frame_index* fi, float x, float y, float r, int min_level,
                                 int max_level, int32_t* out, int cap, int* n);

/* replaces int ORBmatcher::SearchByProjection(Frame& F, const vector<MapPoint*>&, const float th)
 *          include/ORBmatcher.h:52, src/ORBmatcher.cc:45-129.
 * per map point i: valid = mbTrackInView && !isBad(); proj_* = mTrackProjX/Y/XR; pred_level =
 * mnTrackScaleLevel; view_cos = mTrackViewCos; nobs = Observations(); desc = GetDescriptor().
 * frame_mp_obs[k] (in/out, nf) = Observations() of the map point frame keypoint k holds, 0 if none.
 * match[k] (out, nf) = index of the map point assigned to keypoint k by this call, or -1.          */
int viorb_search_by_projection_local(viorb_frame_index* fi, int32_t* frame_mp_obs,
                                     const float* proj_x, const float* proj_y, const float* proj_xr,
                                     const int32_t* pred_level, const float* view_cos, const uint8_t* valid,
                                     const int32_t* nobs, const uint8_t* mp_desc, int nmp,
                                     float th, float nnratio, int32_t* match, int* nmatches);

/* replaces int ORBmatcher::SearchByProjection(Frame& Cur, const Frame& Last, const float th, bool bMono)
 *          include/ORBmatcher.h:56, src/ORBmatcher.cc:1328-1471  (and, with mode 0 and th_high = ORBdist,
 *          the relocalisation overload :1473-1600 whose search loop is the same).
 * u, v, invz = projection of each last-frame map point into the current frame (:1359-1376);
 * mode & 7 = level window: 0: [o-1,o+1]; 1 (forward): >= o; 2 (backward): [0,o] (:1385-1390); 3: [o-1,o].
 * mode & 8 = no stereo (uRight) gate.  The same search loop serves the other two overloads:
 *   SearchByProjection(Frame&, KeyFrame*, const set<MapPoint*>&, th, ORBdist)        :1473-1600 (relocalisation):
 *     mode 0|8, th_high = ORBdist, last_octave = PredictScale(dist3D), frame_mp_obs[k] = (mvpMapPoints[k] != NULL),
 *     nobs = 1 for every point (any assigned keypoint is skipped, :1541-1542);
 *   SearchByProjection(KeyFrame*, cv::Mat Scw, vpPoints, vpMatched, th)             :290-403 (loop closing):
 *     mode 3|8, th_high = TH_LOW, no orientation check, the frame index built from the KeyFrame's keypoints,
 *     frame_mp_obs[k] = (vpMatched[k] != NULL), nobs = 1.
 * match[k] = index of the point assigned to keypoint k, -1 = untouched, -2 = assigned and then removed by the
 * rotation-consistency check: the reference leaves such a slot NULL (:1460, :1586) even if it held a map point
 * without observations before the call, so the caller must clear it.                                  */
int viorb_search_by_projection_frame(viorb_frame_index* fi, int32_t* frame_mp_obs,
                                     const float* u, const float* v, const float* invz,
                                     const int32_t* last_octave, const float* last_angle,
                                     const uint8_t* valid, const int32_t* nobs, const uint8_t* mp_desc,
                                     int nlast, float th, float mbf, int mode, int check_orientation,
                                     int th_high, int32_t* match, int* nmatches);

/* replaces int ORBmatcher::SearchForTriangulation(KeyFrame*, KeyFrame*, cv::Mat F12,
 *          vector<pair<size_t,size_t>>&, const bool bOnlyStereo)
 *          include/ORBmatcher.h:71-72, src/ORBmatcher.cc:657-823.
 * DBoW2 feature vectors are passed flattened: node ids ascending, CSR offsets, keypoint indices.
 * matches12[i] (n1) = index in KF2 matched to keypoint i of KF1, or -1.                            */
int viorb_search_for_triangulation(viorb_ctx* ctx,
                                   const viorb_keypoint* k1, const uint8_t* d1, const float* ur1,
                                   const uint8_t* has_mp1, int n1,
                                   const viorb_keypoint* k2, const uint8_t* d2, const float* ur2,
                                   const uint8_t* has_mp2, int n2,
                                   const int32_t* node_id1, const int32_t* node_ptr1, const int32_t* idx1, int nn1,
                                   const int32_t* node_id2, const int32_t* node_ptr2, const int32_t* idx2, int nn2,
                                   const float* F12, float ex, float ey,
                                   const float* scale_factors2, const float* level_sigma2_2, int nlevels,
                                   int only_stereo, int check_orientation,
                                   int32_t* matches12, int* nmatches);

/* replaces void MapPoint::ComputeDistinctiveDescriptors()      include/MapPoint.h:71, src/MapPoint.cc:249-314,
 * batched over the nmp map points a keyframe insertion touches (callers src/LocalMapping.cc:1176,1475,1559).
 * obs_desc: the descriptors of each point's non-bad observations in std::map<KeyFrame*,size_t> iteration order
 * (:269-275), concatenated; point p owns rows [obs_ptr[p], obs_ptr[p+1]) (obs_ptr[0] = 0, nmp+1 entries).
 * best[p] = BestIdx within the point's rows (the row to clone into mDescriptor), -1 for a point without
 * observations (the reference returns early, :262-263,277-278); best_median (may be NULL) = BestMedian.  */
int viorb_distinctive_descriptors(viorb_ctx* ctx, const uint8_t* obs_desc, const int32_t* obs_ptr, int nmp,
                                  int32_t* best, int32_t* best_median);

/* replaces int ORBmatcher::SearchByBoW(KeyFrame* pKF, Frame& F, vector<MapPoint*>& vpMapPointMatches)
 *          include/ORBmatcher.h:61, src/ORBmatcher.cc:159-288                                        (mode 0)
 *     and  int ORBmatcher::SearchByBoW(KeyFrame* pKF1, KeyFrame* pKF2, vector<MapPoint*>& vpMatches12)
 *          include/ORBmatcher.h:62, src/ORBmatcher.cc:522-655                                        (mode 1).
 * Set 1 is pKF / pKF1 (k1 = mvKeysUn, valid1[i] = its MapPoint i exists and !isBad()); set 2 is F (k2 = F.mvKeys,
 * valid2 = NULL) or pKF2 (k2 = mvKeysUn, valid2 as valid1).  Feature vectors flattened as for
 * viorb_search_for_triangulation.  mode 0: match[k] (n2 entries) = index in pKF of the map point given to frame
 * keypoint k, or -1; mode 1: match[i] (n1 entries) = keypoint of pKF2 whose map point is matched to keypoint i.  */
int viorb_search_by_bow(viorb_ctx* ctx, int mode, const viorb_keypoint* k1, const uint8_t* d1, const uint8_t* valid1, int n1,
                        const viorb_keypoint* k2, const uint8_t* d2, const uint8_t* valid2, int n2,
                        const int32_t* node_id1, const int32_t* node_ptr1, const int32_t* idx1, int nn1,
                        const int32_t* node_id2, const int32_t* node_ptr2, const int32_t* idx2, int nn2,
                        float nnratio, int check_orientation, int32_t* match, int* nmatches);

/* replaces int ORBmatcher::SearchForInitialization(Frame& F1, Frame& F2, vector<cv::Point2f>& vbPrevMatched,
 *          vector<int>& vnMatches12, int windowSize)       include/ORBmatcher.h:66, src/ORBmatcher.cc:405-520.
 * f2 = frame index of F2 (mvKeysUn, descriptors, grid); k1_un / d1 = F1.mvKeysUn / mDescriptors;
 * prev_matched (in/out, n1 x 2 floats) = vbPrevMatched; matches12 (out, n1) = vnMatches12.              */
int viorb_search_for_initialization(viorb_frame_index* f2, const viorb_keypoint* k1_un, const uint8_t* d1, int n1,
                                    float* prev_matched, int window_size, float nnratio, int check_orientation,
                                    int32_t* matches12, int* nmatches);

/* Independent windowed top-1 search of projected map points in a KeyFrame: the search loop shared by
 *   int ORBmatcher::Fuse(KeyFrame* pKF, const vector<MapPoint*>& vpMapPoints, const float th)
 *       include/ORBmatcher.h:81, src/ORBmatcher.cc:825-976 -- pass ur = u - bf*invz and inv_level_sigma2 =
 *       pKF->mvInvLevelSigma2 (chi-square gates 7.8 / 5.99, :901-925), th_dist = TH_LOW;
 *   int ORBmatcher::Fuse(KeyFrame* pKF, cv::Mat Scw, const vector<MapPoint*>&, float th, vector<MapPoint*>&)
 *       include/ORBmatcher.h:84, src/ORBmatcher.cc:978-1100 -- ur = NULL, th_dist = TH_LOW.
 * kf = frame index of the KeyFrame (mvKeysUn, mDescriptors, mvuRight, grid).  Per map point i: valid[i] = it passed the
 * reference's projection gates (:846-881 / :1001-1041, evaluated by the caller), u/v = projection, pred_level =
 * PredictScale(dist3D, pKF).  best_idx[i] = keypoint with the least distance <= th_dist (first in GetFeaturesInArea
 * order on ties) or -1; the caller replays the Replace / AddObservation bookkeeping (:945-967) in order.          */
int viorb_search_window_top1(viorb_frame_index* kf, const float* u, const float* v, const float* ur,
                             const int32_t* pred_level, const uint8_t* valid, const uint8_t* mp_desc, int n, float th,
                             int th_dist, const float* inv_level_sigma2, int32_t* best_idx, int32_t* best_dist);

/* replaces int ORBmatcher::SearchBySim3(KeyFrame* pKF1, KeyFrame* pKF2, vector<MapPoint*>& vpMatches12, const float& s12,
 *          const cv::Mat& R12, const cv::Mat& t12, const float th)   include/ORBmatcher.h:76-77, src/ORBmatcher.cc:1102-1326.
 * kf1 / kf2 = frame indices of the two key frames.  Arrays *12 have kf1->n entries (map point of keypoint i1 of KF1
 * projected into KF2: valid12[i1] = has a good map point, not already matched, passed the gates of :1155-1191);
 * arrays *21 the reverse (:1234-1270).  match12[i1] = idx2 where both directions agree (:1305-1320), else -1.     */
int viorb_search_by_sim3(viorb_frame_index* kf1, viorb_frame_index* kf2, const float* u12, const float* v12,
                         const int32_t* level12, const uint8_t* valid12, const uint8_t* mp_desc1, const float* u21,
                         const float* v21, const int32_t* level21, const uint8_t* valid21, const uint8_t* mp_desc2,
                         float th, int32_t* match12, int* nfound);

/* ---- DBoW2 vocabulary transform ---------------------------------------------------
 * replaces ORBVocabulary (= DBoW2::TemplatedVocabulary<FORB::TDescriptor, FORB>) as used by
 *          Frame::ComputeBoW src/Frame.cc:575-582 and KeyFrame::ComputeBoW src/KeyFrame.cc:350-359:
 *          mpORBvocabulary->transform(vCurrentDesc, mBowVec, mFeatVec, 4)
 *          Thirdparty/DBoW2/DBoW2/TemplatedVocabulary.h:1138-1204 (vector transform), :1230-1272 (tree descent).
 * The tree is passed as loadFromTextFile reads it (:1351-1437): node 0 is the root; node i > 0 has parent[i] < i,
 * a 32-byte descriptor and a weight; children keep node-id order; a node without children is a leaf (isLeaf())
 * and leaves get word ids 0, 1, ... in node order.  weighting: 0 TF_IDF, 1 TF, 2 IDF, 3 BINARY; scoring: 0 L1_NORM,
 * 1 L2_NORM, 2 CHI_SQUARE, 3 KL, 4 BHATTACHARYYA, 5 DOT_PRODUCT (BowVector.h:36-53) -- it selects the normalisation
 * (ScoringObject.h:74-90).                                                                                      */
typedef struct viorb_vocabulary viorb_vocabulary;
int viorb_vocabulary_create(viorb_ctx* ctx, int k, int L, int weighting, int scoring, int nnodes, const int32_t* parent,
                            const uint8_t* node_desc, const double* node_weight, viorb_vocabulary** out);
int viorb_vocabulary_destroy(viorb_vocabulary* voc);
int viorb_vocabulary_info(const viorb_vocabulary* voc, int* nnodes, int* nwords);
/* transform(features, v, fv, levelsup) of n descriptors (host, n x 32 bytes).  Outputs (host, caller allocated):
 *   BowVector    bow_ids[*nbow] ascending word ids, bow_values[*nbow] (capacity n each);
 *   FeatureVector in the flattened form viorb_search_for_triangulation takes: fv_node[*nfv] ascending node ids,
 *                fv_ptr[*nfv + 1] offsets into fv_idx, fv_idx[...] feature indices ascending within a node
 *                (capacities n, n + 1, n);
 *   word_of / node_of (may be NULL, n each): the word id and the node at level L - levelsup of every feature.
 * A feature that reaches a leaf above level L - levelsup reports node 0 (the reference leaves *nid unset there). */
int viorb_bow_transform(viorb_vocabulary* voc, const uint8_t* desc, int n, int levelsup, int32_t* bow_ids,
                        double* bow_values, int* nbow, int32_t* fv_node, int32_t* fv_ptr, int32_t* fv_idx, int* nfv,
                        int32_t* word_of, int32_t* node_of);

#ifdef __cplusplus
}
#endif
#endif /* VIORB_GPU_H */
