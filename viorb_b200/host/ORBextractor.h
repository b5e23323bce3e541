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

    /* mvImagePyramid[l]: ROI view into the padded (border 19) level image, as in the reference (include/ORBextractor.h:85).
     * The levels live on the device (that copy is what the GPU ComputeStereoMatches reads); the host copy is fetched
     * lazily, on the first mvImagePyramid[l] after a call, so callers that never look at it (monocular tracking) pay
     * nothing.  Source compatible with the reference's std::vector<cv::Mat> for what Frame.cc does with it:
     * operator[], size(). */
    class LazyPyramid {
    public:
        explicit LazyPyramid(ORBextractor* owner = nullptr) : mpOwner(owner) {}
        const cv::Mat& operator[](size_t level) const { mpOwner->SyncPyramid(); return mpOwner->mvRoiLevels[level]; }
        size_t size() const { return mpOwner->mvRoiLevels.size(); }
        void resize(size_t) {}
    private:
        ORBextractor* mpOwner;
    };
    LazyPyramid mvImagePyramid;
    /* true: fetch the host copy inside every operator() (the reference's timing behaviour); default false = lazy */
    void SetPyramidDownload(bool on) { mbDownloadPyramid = on; }
    void SyncPyramid();
    /* which OpenCV's GaussianBlur (:1086) to reproduce: 0 = OpenCV >= 3.4 (default), 1 = OpenCV 2.4, the version the
     * reference's CMakeLists.txt pins.  See viorb_extractor_set_gaussian. */
    void SetGaussianVariant(int opencvVariant);
    /* where the GaussianBlur is evaluated (0 automatic, 1 per keypoint, 2 whole levels): results are identical.  See
     * viorb_extractor_set_describe_mode. */
    void SetDescribeMode(int mode);

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
    bool mbDownloadPyramid, mbPyramidStale;
    std::vector<cv::Mat> mvPaddedLevels, mvRoiLevels;
    ORBextractor(const ORBextractor&);              /* owns device state: not copyable (the reference's is never copied) */
    ORBextractor& operator=(const ORBextractor&);
};

}  // namespace ORB_SLAM2
#endif
