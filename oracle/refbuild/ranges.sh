#!/bin/sh
# oracle/refbuild/ranges.sh REF -- writes to stdout ONE translation unit made of member functions of Frame / KeyFrame /
# MapPoint, taken as line ranges straight from the reference's source files (which need Eigen, g2o and the IMU code to
# compile whole).  The Makefile pipes it into the compiler; nothing of the reference is written into the repository.
# Each range must still start with the expected signature, otherwise the build fails (the reference tree has changed).
# TEST INFRASTRUCTURE ONLY.
set -e
REF="$1"
emit() {   # file first last 'signature prefix'
    first_line=$(sed -n "${2}p" "$REF/$1")
    case "$first_line" in
        "$4"*) ;;
        *) echo "ranges.sh: $1:$2 does not start with '$4' (found: $first_line)" >&2; exit 1 ;;
    esac
    echo "#line $2 \"$REF/$1\""
    sed -n "${2},${3}p" "$REF/$1"
}
cat <<'HDR'
/* generated on the fly by oracle/refbuild/ranges.sh -- member functions of the reference, verbatim line ranges */
#include "ORBextractor.h"
#include "ORBmatcher.h"
#include <limits.h>
#include <thread>
namespace ORB_SLAM2
{
HDR
emit src/Frame.cc 410 425 'void Frame::AssignFeaturesToGrid()'
emit src/Frame.cc 449 505 'bool Frame::isInFrustum(MapPoint *pMP, float viewingCosLimit)'
emit src/Frame.cc 507 560 'vector<size_t> Frame::GetFeaturesInArea('
emit src/Frame.cc 562 572 'bool Frame::PosInGrid('
emit src/Frame.cc 584 614 'void Frame::UndistortKeyPoints()'
emit src/Frame.cc 616 644 'void Frame::ComputeImageBounds('
emit src/Frame.cc 646 821 'void Frame::ComputeStereoMatches()'
emit src/KeyFrame.cc 906 945 'vector<size_t> KeyFrame::GetFeaturesInArea('
emit src/MapPoint.cc 249 314 'void MapPoint::ComputeDistinctiveDescriptors()'
emit src/MapPoint.cc 380 390 'float MapPoint::GetMinDistanceInvariance()'
emit src/MapPoint.cc 392 424 'int MapPoint::PredictScale(const float &currentDist, KeyFrame* pKF)'
echo '}'
