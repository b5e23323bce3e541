/*
 * ORBmatcher.h -- drop-in for ORB_SLAM2::ORBmatcher (reference include/ORBmatcher.h:37-102) for the searches
 * on the hot path: DescriptorDistance, the four SearchByProjection overloads, SearchForTriangulation, both
 * SearchByBoW overloads and SearchForInitialization.  Each call marshals the fields the reference reads into flat
 * arrays and runs the whole search as CUDA kernels behind include/viorb_gpu.h; results (matches, counts,
 * tie-breaks, the "already matched" dependence) are identical to the reference's sequential loops.
 * SearchBySim3 and the two Fuse overloads run their search loops on the GPU in one call and then replay the reference's
 * map-graph bookkeeping (Replace / AddObservation / AddMapPoint) on the host in the reference's order.
 */
#ifndef ORBMATCHER_H
#define ORBMATCHER_H

#include <set>
#include <utility>
#include <vector>

#include "cv_compat.h"
#include "orbslam_compat.h"

namespace ORB_SLAM2 {

class ORBmatcher {
public:
    ORBmatcher(float nnratio = 0.6, bool checkOri = true);

    /* Hamming distance between two 256-bit ORB descriptors (reference :1648-1664) */
    static int DescriptorDistance(const cv::Mat& a, const cv::Mat& b);
    /* B200 extension: n pairs in one device call (row i of a vs row i of b) */
    static std::vector<int> DescriptorDistances(const cv::Mat& a, const cv::Mat& b);

    /* Search matches between Frame keypoints and projected MapPoints; returns number of matches (:45-129) */
    int SearchByProjection(Frame& F, const std::vector<MapPoint*>& vpMapPoints, const float th = 3);

    /* Project MapPoints tracked in last frame into the current frame and search matches (:1328-1471) */
    int SearchByProjection(Frame& CurrentFrame, const Frame& LastFrame, const float th, const bool bMono);

    /* Project MapPoints seen in KeyFrame into the Frame and search matches; used in relocalisation (:1473-1600) */
    int SearchByProjection(Frame& CurrentFrame, KeyFrame* pKF, const std::set<MapPoint*>& sAlreadyFound, const float th,
                           const int ORBdist);

    /* Project MapPoints using a similarity transformation and search matches; used in loop detection (:290-403) */
    int SearchByProjection(KeyFrame* pKF, cv::Mat Scw, const std::vector<MapPoint*>& vpPoints,
                           std::vector<MapPoint*>& vpMatched, int th);

    /* Search matches between MapPoints in a KeyFrame and ORB in a Frame, brute force constrained to ORB of the same
     * vocabulary node; used in relocalisation and loop detection (:159-288) */
    int SearchByBoW(KeyFrame* pKF, Frame& F, std::vector<MapPoint*>& vpMapPointMatches);
    int SearchByBoW(KeyFrame* pKF1, KeyFrame* pKF2, std::vector<MapPoint*>& vpMatches12);              /* (:522-655) */

    /* Matching for the map initialisation, monocular case only (:405-520) */
    int SearchForInitialization(Frame& F1, Frame& F2, std::vector<cv::Point2f>& vbPrevMatched, std::vector<int>& vnMatches12,
                                int windowSize = 10);

    /* Search matches between MapPoints seen in KF1 and KF2 transforming by a Sim3 [s12*R12|t12] (:1102-1326) */
    int SearchBySim3(KeyFrame* pKF1, KeyFrame* pKF2, std::vector<MapPoint*>& vpMatches12, const float& s12, const cv::Mat& R12,
                     const cv::Mat& t12, const float th);

    /* Project MapPoints into KeyFrame and search for duplicated MapPoints (:825-976) */
    int Fuse(KeyFrame* pKF, const std::vector<MapPoint*>& vpMapPoints, const float th = 3.0);
    /* Project MapPoints into KeyFrame using a given Sim3 and search for duplicated MapPoints (:978-1100) */
    int Fuse(KeyFrame* pKF, cv::Mat Scw, const std::vector<MapPoint*>& vpPoints, float th, std::vector<MapPoint*>& vpReplacePoint);

    /* Matching to triangulate new MapPoints, epipolar constraint check (:657-823) */
    int SearchForTriangulation(KeyFrame* pKF1, KeyFrame* pKF2, cv::Mat F12,
                               std::vector<std::pair<size_t, size_t> >& vMatchedPairs, const bool bOnlyStereo);

public:
    static const int TH_LOW;
    static const int TH_HIGH;
    static const int HISTO_LENGTH;

protected:
    float mfNNratio;
    bool mbCheckOrientation;
};

}  // namespace ORB_SLAM2
#endif
