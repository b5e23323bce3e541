/*
 * ORBextractor.h -- drop-in for ORB_SLAM2::ORBextractor (reference include/ORBextractor.h:45-111) backed by
 * libviorb_b200.so.  Same constructor, operator(), getters and public mvImagePyramid, so Frame and Tracking
 * compile and link against it unchanged; every stage runs as a hand-written sm_100a kernel behind
 * include/viorb_gpu.h.  Error behaviour: the reference returns silently on an empty image and asserts on a
 * non-CV_8UC1 image; device-side failures (which the reference cannot have) throw std::runtime_error.
 */
#ifndef ORBEXTRACTOR_H
#define ORBEXTRACTOR_H

#include <vector>

#include "cv_compat.h"

struct viorb_ctx;
struct viorb_extractor;

namespace ORB_SLAM2 {

class ORBextractor {
public:
    enum { HARRIS_SCORE = 0, FAST_SCORE = 1 };

    ORBextractor(int nfeatures, float scaleFactor, int nlevels, int iniThFAST, int minThFAST);
    ~ORBextractor();

    /* Compute the ORB features and descriptors on an image.  Mask is ignored (as in the reference). */
    void operator()(cv::InputArray image, cv::InputArray mask, std::vector<cv::KeyPoint>& keypoints,
                    cv::OutputArray descriptors);

    int inline GetLevels() { return nlevels; }
    float inline GetScaleFactor() { return (float)scaleFactor; }
    std::vector<float> inline GetScaleFactors() { return mvScaleFactor; }
    std::vector<float> inline GetInverseScaleFactors() { return mvInvScaleFactor; }
    std::vector<float> inline GetScaleSigmaSquares() { return mvLevelSigma2; }
    std::vector<float> inline GetInverseScaleSigmaSquares() { return mvInvLevelSigma2; }

    /* ROI views into padded (border 19) level images, as in the reference.  Downloaded after every call while
     * SetPyramidDownload(true) (default); the device copy stays resident either way and is what the GPU
     * ComputeStereoMatches reads. */
    std::vector<cv::Mat> mvImagePyramid;
    void SetPyramidDownload(bool on) { mbDownloadPyramid = on; }

    /* B200 extension: the same operator over a batch of equally sized frames (one device pass per 128 frames) */
    void ExtractBatch(const std::vector<cv::Mat>& images, std::vector<std::vector<cv::KeyPoint> >& keypoints,
                      std::vector<cv::Mat>& descriptors);

    viorb_extractor* Handle() { return mpHandle; }
    viorb_ctx* Context() { return mpCtx; }

protected:
    int nfeatures;
    double scaleFactor;
    int nlevels;
    int iniThFAST;
    int minThFAST;
    std::vector<int> mnFeaturesPerLevel;
    std::vector<float> mvScaleFactor, mvInvScaleFactor, mvLevelSigma2, mvInvLevelSigma2;

    viorb_ctx* mpCtx;
    viorb_extractor* mpHandle;
    bool mbDownloadPyramid;
    std::vector<cv::Mat> mvPaddedLevels;
};

}  // namespace ORB_SLAM2
#endif
