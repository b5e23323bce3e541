/* ORBextractor.cc -- see ORBextractor.h.  Host-side marshalling only; no image arithmetic happens here. */
#include "ORBextractor.h"

#include <cassert>
#include <stdexcept>
#include <string>

#include "viorb_gpu.h"

namespace ORB_SLAM2 {

static void check(int rc, const char* what) {
    if (rc != VIORB_OK) throw std::runtime_error(std::string(what) + ": " + viorb_last_error());
}

ORBextractor::ORBextractor(int _nfeatures, float _scaleFactor, int _nlevels, int _iniThFAST, int _minThFAST)
    : mvImagePyramid(this), nfeatures(_nfeatures), scaleFactor(_scaleFactor), nlevels(_nlevels), iniThFAST(_iniThFAST), minThFAST(_minThFAST),
      mpCtx(nullptr), mpHandle(nullptr), mbDownloadPyramid(false), mbPyramidStale(false) {
    check(viorb_ctx_create(0, nullptr, &mpCtx), "viorb_ctx_create");
    check(viorb_extractor_create(mpCtx, nfeatures, _scaleFactor, nlevels, iniThFAST, minThFAST, &mpHandle), "viorb_extractor_create");
    mvScaleFactor.resize(nlevels); mvInvScaleFactor.resize(nlevels);
    mvLevelSigma2.resize(nlevels); mvInvLevelSigma2.resize(nlevels);
    mnFeaturesPerLevel.resize(nlevels);
    int n = 0;
    check(viorb_extractor_tables(mpHandle, &n, mvScaleFactor.data(), mvInvScaleFactor.data(), mvLevelSigma2.data(),
                                 mvInvLevelSigma2.data(), mnFeaturesPerLevel.data()), "viorb_extractor_tables");
    mvRoiLevels.resize(nlevels);
    mvPaddedLevels.resize(nlevels);
}

ORBextractor::~ORBextractor() {
    viorb_extractor_destroy(mpHandle);
    viorb_ctx_destroy(mpCtx);
}

void ORBextractor::operator()(cv::InputArray _image, cv::InputArray /*mask*/, std::vector<cv::KeyPoint>& _keypoints,
                              cv::OutputArray _descriptors) {
    if (_image.empty()) return;                                  /* reference :1046-1047 */
    const cv::Mat image = _image.getMat();
    assert(image.type() == CV_8UC1);                             /* reference :1050 */
    const int cap = nfeatures + 8 * nlevels + 64;                /* the quadtree may overshoot nfeatures by a few */
    _keypoints.resize(cap);
    cv::Mat desc(cap, 32, CV_8U);
    int n = 0;
    static_assert(sizeof(cv::KeyPoint) == sizeof(viorb_keypoint), "KeyPoint layout");
    check(viorb_extract(mpHandle, image.data, image.rows, image.cols, image.step,
                        reinterpret_cast<viorb_keypoint*>(_keypoints.data()), desc.data, cap, &n), "viorb_extract");
    _keypoints.resize(n);
    if (n == 0) _descriptors.release();                          /* reference :1064-1065 */
    else {
        _descriptors.create(n, 32, CV_8U);
        cv::Mat out = _descriptors.getMat();
        for (int i = 0; i < n; i++) memcpy(out.ptr<unsigned char>(i), desc.ptr<unsigned char>(i), 32);
    }
    mbPyramidStale = true;
    if (mbDownloadPyramid) SyncPyramid();
}

void ORBextractor::SetDescribeMode(int mode) {
    check(viorb_extractor_set_describe_mode(mpHandle, mode), "viorb_extractor_set_describe_mode");
}

void ORBextractor::SetGaussianVariant(int opencvVariant) {
    check(viorb_extractor_set_gaussian(mpHandle, opencvVariant), "viorb_extractor_set_gaussian");
}

/* host copy of the padded levels of the last call (reference ComputePyramid :1107-1132 leaves them in mvImagePyramid) */
void ORBextractor::SyncPyramid() {
    if (!mbPyramidStale) return;
    std::vector<uint8_t*> dst(nlevels);
    std::vector<size_t> steps(nlevels);
    std::vector<int> ws(nlevels), hs(nlevels);
    for (int l = 0; l < nlevels; l++) {
        check(viorb_extractor_pyramid_info(mpHandle, l, &ws[l], &hs[l]), "viorb_extractor_pyramid_info");
        mvPaddedLevels[l].create(hs[l] + 38, ws[l] + 38, CV_8U);
        dst[l] = mvPaddedLevels[l].data;
        steps[l] = mvPaddedLevels[l].step;
    }
    /* one device-to-host copy of the frame's pyramid block (pinned staging inside the library), then row copies */
    check(viorb_extractor_pyramid_download_all(mpHandle, 0, dst.data(), steps.data(), nlevels), "pyramid download");
    for (int l = 0; l < nlevels; l++) mvRoiLevels[l] = mvPaddedLevels[l](cv::Rect(19, 19, ws[l], hs[l]));
    mbPyramidStale = false;
}

void ORBextractor::ExtractBatch(const std::vector<cv::Mat>& images, std::vector<std::vector<cv::KeyPoint> >& keypoints,
                                std::vector<cv::Mat>& descriptors) {
    const int B = (int)images.size();
    keypoints.assign(B, std::vector<cv::KeyPoint>());
    descriptors.assign(B, cv::Mat());
    if (B == 0) return;
    const int rows = images[0].rows, cols = images[0].cols;
    const int cap = nfeatures + 8 * nlevels + 64;
    uint8_t* staging = nullptr;
    check(viorb_host_alloc((size_t)B * rows * cols, (void**)&staging), "viorb_host_alloc");
    for (int b = 0; b < B; b++) {
        assert(images[b].rows == rows && images[b].cols == cols && images[b].type() == CV_8UC1);
        for (int r = 0; r < rows; r++) memcpy(staging + ((size_t)b * rows + r) * cols, images[b].ptr<uint8_t>(r), cols);
    }
    std::vector<viorb_keypoint> kps((size_t)B * cap);
    std::vector<uint8_t> desc((size_t)B * cap * 32);
    std::vector<int32_t> counts(B);
    const int rc = viorb_extract_batch(mpHandle, staging, B, rows, cols, cols, (size_t)rows * cols, kps.data(), desc.data(), cap, counts.data());
    viorb_host_free(staging);
    check(rc, "viorb_extract_batch");
    for (int b = 0; b < B; b++) {
        const int n = counts[b];
        keypoints[b].resize(n);
        memcpy((void*)keypoints[b].data(), &kps[(size_t)b * cap], (size_t)n * sizeof(viorb_keypoint));
        if (n) {
            descriptors[b].create(n, 32, CV_8U);
            memcpy(descriptors[b].data, &desc[(size_t)b * cap * 32], (size_t)n * 32);
        }
    }
}

}  // namespace ORB_SLAM2
