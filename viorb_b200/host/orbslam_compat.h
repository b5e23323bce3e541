/*
 * orbslam_compat.h -- the members of Frame / KeyFrame / MapPoint that the ORB front-end reads or writes
 * (reference include/Frame.h, include/KeyFrame.h, include/MapPoint.h), as plain structs, for builds outside
 * the VIORB tree (tests, this repository).  Inside the VIORB tree define VIORB_USE_ORBSLAM_HEADERS and the
 * shims compile against the real classes: only public members and accessors with these names are used.
 */
#ifndef VIORB_ORBSLAM_COMPAT_H
#define VIORB_ORBSLAM_COMPAT_H

#ifdef VIORB_USE_ORBSLAM_HEADERS
#include "Frame.h"
#include "KeyFrame.h"
#include "MapPoint.h"
#else

#include <cmath>
#include <map>
#include <mutex>
#include <set>
#include <vector>

#include "cv_compat.h"

#ifndef VIORB_HAVE_DBOW2          /* defined when the real Thirdparty/DBoW2 headers are on the include path */
namespace DBoW2 {
typedef std::map<unsigned int, std::vector<unsigned int> > FeatureVector;   /* Thirdparty/DBoW2/DBoW2/FeatureVector.h */
typedef std::map<unsigned int, double> BowVector;                          /* Thirdparty/DBoW2/DBoW2/BowVector.h:56-57 */
}
#endif

#define FRAME_GRID_ROWS 48        /* include/Frame.h:41-42 */
#define FRAME_GRID_COLS 64

namespace ORB_SLAM2 {

class ORBextractor;
class KeyFrame;
class Frame;

/* the reference guards these members with std::mutex; the stand-ins stay copyable */
struct CompatMutex : std::mutex {
    CompatMutex() {}
    CompatMutex(const CompatMutex&) {}
    CompatMutex& operator=(const CompatMutex&) { return *this; }
};

class MapPoint {   /* include/MapPoint.h */
public:
    /* observation bookkeeping used by Fuse / SearchBySim3 (include/MapPoint.h:55-64), reduced to what the searches touch */
    std::map<KeyFrame*, size_t> mObservations;
    MapPoint* mpReplaced = nullptr;
    int descriptorEpoch = 0;                   /* bumped when Replace recomputes the distinctive descriptor */
    void AddObservation(KeyFrame* pKF, size_t idx) {
        if (mObservations.count(pKF)) return;
        mObservations[pKF] = idx;
        nObs++;
    }
    int GetIndexInKeyFrame(KeyFrame* pKF) { return mObservations.count(pKF) ? (int)mObservations[pKF] : -1; }
    bool IsInKeyFrame(KeyFrame* pKF) { return mObservations.count(pKF) != 0; }
    inline void Replace(MapPoint* pMP);        /* src/MapPoint.cc:181-225 */
    bool mbTrackInView = false;                /* :85-89 tracking fields written by Frame::isInFrustum */
    float mTrackProjX = 0, mTrackProjY = 0, mTrackProjXR = 0;
    int mnTrackScaleLevel = 0;
    float mTrackViewCos = 1;
    bool bad = false;
    int nObs = 1;
    cv::Mat descriptor;                        /* 1x32 CV_8U */
    cv::Mat worldPos;                          /* 3x1 CV_32F */
    cv::Mat normal;                            /* 3x1 CV_32F mean viewing direction */
    float mfMinDistance = 0, mfMaxDistance = 0;
    bool isBad() const { return bad; }
    int Observations() const { return nObs; }
    cv::Mat GetDescriptor() const { return descriptor; }
    cv::Mat GetWorldPos() const { return worldPos; }
    cv::Mat GetNormal() const { return normal; }
    CompatMutex mMutexPos, mMutexFeatures;
#ifdef VIORB_REF_MEMBERS        /* oracle/refbuild: the bodies are the reference's own source lines (src/MapPoint.cc) */
    float GetMinDistanceInvariance();
    float GetMaxDistanceInvariance();
    int PredictScale(const float& currentDist, KeyFrame* pKF);
    int PredictScale(const float& currentDist, Frame* pF);
    void ComputeDistinctiveDescriptors();       /* src/MapPoint.cc:249-314 */
    cv::Mat mDescriptor;                        /* written by ComputeDistinctiveDescriptors */
    bool mbBad = false;
#else
    float GetMinDistanceInvariance() const { return 0.8f * mfMinDistance; }     /* src/MapPoint.cc:380-384 */
    float GetMaxDistanceInvariance() const { return 1.2f * mfMaxDistance; }     /* :386-390 */
    template <class FrameOrKeyFrame>
    int PredictScale(const float& currentDist, FrameOrKeyFrame* pF) const {     /* :392-424 */
        const float ratio = mfMaxDistance / currentDist;
        int nScale = (int)std::ceil(std::log(ratio) / pF->mfLogScaleFactor);
        if (nScale < 0) nScale = 0;
        else if (nScale >= pF->mnScaleLevels) nScale = pF->mnScaleLevels - 1;
        return nScale;
    }
#endif
};

class Frame {      /* include/Frame.h */
public:
    int N = 0;
    std::vector<cv::KeyPoint> mvKeys, mvKeysRight, mvKeysUn;
    std::vector<float> mvuRight, mvDepth;
    cv::Mat mDescriptors, mDescriptorsRight;
    std::vector<MapPoint*> mvpMapPoints;
    DBoW2::FeatureVector mFeatVec;
    std::vector<bool> mvbOutlier;
    std::vector<float> mvScaleFactors, mvInvScaleFactors;
    float mnMinX = 0, mnMaxX = 0, mnMinY = 0, mnMaxY = 0;
    float fx = 0, fy = 0, cx = 0, cy = 0, mbf = 0, mb = 0;
    int mnScaleLevels = 8;
    float mfLogScaleFactor = 0.18232156f;      /* log(1.2) */
    cv::Mat mTcw;                              /* 4x4 CV_32F */
    ORBextractor *mpORBextractorLeft = nullptr, *mpORBextractorRight = nullptr;
    cv::Mat mK, mDistCoef;                     /* 3x3 and 4x1 (or 5x1) CV_32F */
    float mfGridElementWidthInv = 0, mfGridElementHeightInv = 0;
    std::vector<size_t> mGrid[64][48];         /* FRAME_GRID_COLS x FRAME_GRID_ROWS, include/Frame.h:41-42,191 */
    /* the members of the reference Frame that sit on the ORB path, each one call into libviorb_b200 (host/ORBmatcher.cc) */
    void ExtractORB(int flag, const cv::Mat& im);                      /* src/Frame.cc:427-433 */
    void UndistortKeyPoints();                                         /* :584-614 -> viorb_undistort_keypoints */
    void ComputeImageBounds(const cv::Mat& imLeft);                    /* :616-645 -> viorb_compute_image_bounds */
    void AssignFeaturesToGrid();                                       /* :410-425 -> viorb_frame_index_create + _grid */
    std::vector<size_t> GetFeaturesInArea(const float& x, const float& y, const float& r, const int minLevel = -1,
                                          const int maxLevel = -1) const;      /* :507-560 -> viorb_frame_features_in_area */
    void ComputeStereoMatches();               /* src/Frame.cc:646-820 -> viorb_stereo_match */
    bool PosInGrid(const cv::KeyPoint& kp, int& posX, int& posY);     /* :562-572 */
    bool isInFrustum(MapPoint* pMP, float viewingCosLimit);            /* :449-505, fills the mTrack* fields read by :45-129 */
    cv::Mat mRcw, mtcw, mOw;                   /* pose parts read by isInFrustum (Frame::UpdatePoseMatrices, :441-447) */
};

class KeyFrame {   /* include/KeyFrame.h */
public:
    int N = 0;
    std::vector<cv::KeyPoint> mvKeysUn;
    std::vector<float> mvuRight;
    cv::Mat mDescriptors;
    DBoW2::FeatureVector mFeatVec;
    std::vector<MapPoint*> mapPoints;
    std::vector<float> mvScaleFactors, mvLevelSigma2;
    float fx = 0, fy = 0, cx = 0, cy = 0;
    float mnMinX = 0, mnMaxX = 0, mnMinY = 0, mnMaxY = 0;
    int mnScaleLevels = 8;
    float mfLogScaleFactor = 0.18232156f;
    cv::Mat Rcw, tcw, Ow;                      /* 3x3, 3x1, 3x1 CV_32F */
    std::vector<float> mvInvLevelSigma2;
    float mbf = 0;
    void AddMapPoint(MapPoint* pMP, const size_t& idx) { mapPoints[idx] = pMP; }
    void ReplaceMapPointMatch(const size_t& idx, MapPoint* pMP) { mapPoints[idx] = pMP; }
    void EraseMapPointMatch(const size_t& idx) { mapPoints[idx] = nullptr; }
    std::set<MapPoint*> GetMapPoints() const {
        std::set<MapPoint*> s;
        for (size_t i = 0; i < mapPoints.size(); i++)
            if (mapPoints[i] && !mapPoints[i]->isBad()) s.insert(mapPoints[i]);
        return s;
    }
    MapPoint* GetMapPoint(size_t idx) const { return mapPoints[idx]; }
    std::vector<MapPoint*> GetMapPointMatches() const { return mapPoints; }
    bool IsInImage(const float& x, const float& y) const { return x >= mnMinX && x < mnMaxX && y >= mnMinY && y < mnMaxY; }
    cv::Mat GetCameraCenter() const { return Ow; }
    cv::Mat GetRotation() const { return Rcw; }
    bool mbBad = false;
    bool isBad() const { return mbBad; }
    cv::Mat GetTranslation() const { return tcw; }
    /* grid copied from the Frame the key frame was made of (src/KeyFrame.cc:301-306) and its window query (:906-945) */
    int mnGridCols = FRAME_GRID_COLS, mnGridRows = FRAME_GRID_ROWS;
    float mfGridElementWidthInv = 0, mfGridElementHeightInv = 0;
    std::vector<std::vector<std::vector<size_t> > > mGrid;
    std::vector<size_t> GetFeaturesInArea(const float& x, const float& y, const float& r) const;
};

inline void MapPoint::Replace(MapPoint* pMP) {
    if (pMP == this) return;
    std::map<KeyFrame*, size_t> obs = mObservations;
    mObservations.clear();
    bad = true;
    mpReplaced = pMP;
    for (std::map<KeyFrame*, size_t>::iterator mit = obs.begin(); mit != obs.end(); ++mit) {
        KeyFrame* pKF = mit->first;
        if (!pMP->IsInKeyFrame(pKF)) {
            pKF->ReplaceMapPointMatch(mit->second, pMP);
            pMP->AddObservation(pKF, mit->second);
        } else {
            pKF->EraseMapPointMatch(mit->second);
        }
    }
    pMP->descriptorEpoch++;                    /* pMP->ComputeDistinctiveDescriptors() */
}

}  // namespace ORB_SLAM2
#endif
#endif
