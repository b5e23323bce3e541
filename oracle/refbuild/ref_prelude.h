/*
 * ref_prelude.h -- force-included (-include) ahead of the reference's matcher sources.  TEST INFRASTRUCTURE ONLY.
 *
 * The reference's include/ORBmatcher.h pulls in Frame.h / KeyFrame.h / MapPoint.h, which need Eigen, g2o and the IMU code to
 * compile.  The Makefile pre-defines their include guards (FRAME_H, KEYFRAME_H, MAPPOINT_H) and force-includes the stand-in
 * classes of viorb_b200/host/orbslam_compat.h instead (the same stand-ins the product shims and tests/cpp use), so
 * include/ORBmatcher.h and src/ORBmatcher.cc themselves compile unmodified.  DBoW2's BowVector / FeatureVector are the
 * reference's own (Thirdparty/DBoW2/DBoW2).
 */
#ifndef VIORB_REF_PRELUDE_H
#define VIORB_REF_PRELUDE_H
#include <list>
#include <mutex>
#include <set>
#include <vector>

#include "Thirdparty/DBoW2/DBoW2/BowVector.h"
#include "Thirdparty/DBoW2/DBoW2/FeatureVector.h"

/* the reference's headers rely on `using namespace std` having leaked from headers included earlier (Frame.h et al.) */
using namespace std;
#endif
