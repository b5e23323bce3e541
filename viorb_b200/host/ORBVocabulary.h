/*
 * ORBVocabulary.h -- drop-in for ORB_SLAM2::ORBVocabulary (reference include/ORBVocabulary.h:29-30, a typedef of
 * DBoW2::TemplatedVocabulary<DBoW2::FORB::TDescriptor, DBoW2::FORB>) for the two members the ORB front-end uses:
 *   loadFromTextFile(filename)                         Thirdparty/DBoW2/DBoW2/TemplatedVocabulary.h:1351-1437
 *   transform(features, BowVector&, FeatureVector&, levelsup)                                 :1138-1204
 * as called by Frame::ComputeBoW (src/Frame.cc:575-582) and KeyFrame::ComputeBoW (src/KeyFrame.cc:350-359).
 * The tree lives on the GPU (viorb_vocabulary_create); transform runs the descent and builds both maps there.
 */
#ifndef VIORB_ORBVOCABULARY_H
#define VIORB_ORBVOCABULARY_H

#include <string>
#include <vector>

#include "cv_compat.h"
#include "orbslam_compat.h"

struct viorb_vocabulary;

namespace ORB_SLAM2 {

class ORBVocabulary {
public:
    ORBVocabulary();
    ~ORBVocabulary();
    ORBVocabulary(const ORBVocabulary&) = delete;
    ORBVocabulary& operator=(const ORBVocabulary&) = delete;

    bool loadFromTextFile(const std::string& filename);
    bool empty() const { return nwords_ == 0; }
    unsigned int size() const { return (unsigned int)nwords_; }

    void transform(const std::vector<cv::Mat>& features, DBoW2::BowVector& v, DBoW2::FeatureVector& fv, int levelsup) const;

private:
    viorb_vocabulary* voc_;
    int k_, L_, nwords_;
};

}  // namespace ORB_SLAM2
#endif
