/* minicv_hooks.h -- test hooks of the stand-in OpenCV (oracle/refbuild).  TEST INFRASTRUCTURE ONLY. */
#ifndef VIORB_MINICV_HOOKS_H
#define VIORB_MINICV_HOOKS_H
#include <cstddef>
#include <vector>

#include "../orb_oracle.h"

struct MinicvFastCall {
    const unsigned char* origin;      /* first pixel of the image cv::FAST was given */
    size_t step;
    int threshold;
    std::vector<orc_corner> corners;  /* what it returned, in order */
};
namespace cv {
void minicv_set_gaussian_variant(int v);
void minicv_set_fast_log(std::vector<MinicvFastCall>* log);
}
#endif
