/* temporary stubs -- replaced by the windowed-search / stereo implementation */
#include "viorb_gpu.h"
extern "C" {
int viorb_stereo_match(viorb_extractor*, int, viorb_extractor*, int, const viorb_keypoint*, const uint8_t*, int,
                       const viorb_keypoint*, const uint8_t*, int, float, float, float*, float*) { return VIORB_ERR_UNSUPPORTED; }
int viorb_frame_index_create(viorb_ctx*, const viorb_keypoint*, const uint8_t*, const float*, int, float, float, float,
                             float, const float*, int, viorb_frame_index**) { return VIORB_ERR_UNSUPPORTED; }
int viorb_frame_index_destroy(viorb_frame_index*) { return VIORB_ERR_UNSUPPORTED; }
int viorb_frame_features_in_area(viorb_frame_index*, float, float, float, int, int, int32_t*, int, int*) { return VIORB_ERR_UNSUPPORTED; }
int viorb_search_by_projection_local(viorb_frame_index*, int32_t*, const float*, const float*, const float*, const int32_t*,
                                     const float*, const uint8_t*, const int32_t*, const uint8_t*, int, float, float,
                                     int32_t*, int*) { return VIORB_ERR_UNSUPPORTED; }
int viorb_search_by_projection_frame(viorb_frame_index*, int32_t*, const float*, const float*, const float*, const int32_t*,
                                     const float*, const uint8_t*, const int32_t*, const uint8_t*, int, float, float, int,
                                     int, int, int32_t*, int*) { return VIORB_ERR_UNSUPPORTED; }
int viorb_search_for_triangulation(viorb_ctx*, const viorb_keypoint*, const uint8_t*, const float*, const uint8_t*, int,
                                   const viorb_keypoint*, const uint8_t*, const float*, const uint8_t*, int,
                                   const int32_t*, const int32_t*, const int32_t*, int, const int32_t*, const int32_t*,
                                   const int32_t*, int, const float*, float, float, const float*, const float*, int, int,
                                   int, int32_t*, int*) { return VIORB_ERR_UNSUPPORTED; }
}
