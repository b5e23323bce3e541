/* ORBVocabulary.cc -- see ORBVocabulary.h.  Host side: text parsing and std::map marshalling only. */
#include "ORBVocabulary.h"

#include <cstring>
#include <fstream>
#include <iostream>
#include <sstream>
#include <stdexcept>

#include "viorb_gpu.h"

namespace ORB_SLAM2 {

namespace {
viorb_ctx* vocab_ctx() {
    struct Holder {
        viorb_ctx* c = nullptr;
        ~Holder() { viorb_ctx_destroy(c); }
    };
    static thread_local Holder h;
    if (!h.c && viorb_ctx_create(0, nullptr, &h.c) != VIORB_OK) throw std::runtime_error(std::string("viorb_ctx_create: ") + viorb_last_error());
    return h.c;
}
}  // namespace

ORBVocabulary::ORBVocabulary() : voc_(nullptr), k_(0), L_(0), nwords_(0) {}
ORBVocabulary::~ORBVocabulary() { viorb_vocabulary_destroy(voc_); }

/* the text format of TemplatedVocabulary::saveToTextFile (:1441-1460): "k L scoring weighting" then one line per node
 * "parent isLeaf d0 .. d31 weight" */
bool ORBVocabulary::loadFromTextFile(const std::string& filename) {
    std::ifstream f(filename.c_str());
    if (!f.good()) return false;
    std::string s;
    std::getline(f, s);
    std::stringstream ss(s);
    int n1 = -1, n2 = -1;
    ss >> k_ >> L_ >> n1 >> n2;
    if (k_ < 0 || k_ > 20 || L_ < 1 || L_ > 10 || n1 < 0 || n1 > 5 || n2 < 0 || n2 > 3) {      /* :1372-1376 */
        std::cerr << "Vocabulary loading failure: This is not a correct text file!" << std::endl;
        return false;
    }
    std::vector<int32_t> parent(1, 0);
    std::vector<uint8_t> desc(32, 0);
    std::vector<double> weight(1, 0.0);
    while (std::getline(f, s)) {
        if (s.find_first_not_of(" \t\r\n") == std::string::npos) continue;      /* the trailing empty line */
        std::stringstream sn(s);
        int pid = 0, leaf = 0;
        sn >> pid >> leaf;
        uint8_t d[32];
        for (int i = 0; i < 32; i++) {
            int v = 0;
            sn >> v;
            d[i] = (uint8_t)v;
        }
        double w = 0;
        sn >> w;
        if (sn.fail()) return false;
        parent.push_back(pid);
        desc.insert(desc.end(), d, d + 32);
        weight.push_back(w);
    }
    viorb_vocabulary_destroy(voc_);
    voc_ = nullptr;
    if (viorb_vocabulary_create(vocab_ctx(), k_, L_, n2, n1, (int)parent.size(), parent.data(), desc.data(), weight.data(), &voc_) != VIORB_OK) {
        std::cerr << "Vocabulary loading failure: " << viorb_last_error() << std::endl;
        return false;
    }
    int nn = 0;
    viorb_vocabulary_info(voc_, &nn, &nwords_);
    return true;
}

void ORBVocabulary::transform(const std::vector<cv::Mat>& features, DBoW2::BowVector& v, DBoW2::FeatureVector& fv, int levelsup) const {
    v.clear();
    fv.clear();
    if (empty() || features.empty()) return;
    const int n = (int)features.size();
    std::vector<uint8_t> d((size_t)n * 32);
    for (int i = 0; i < n; i++) memcpy(&d[(size_t)i * 32], features[i].ptr<uint8_t>(0), 32);
    std::vector<int32_t> ids(n), fvn(n), fvp(n + 1), fvi(n);
    std::vector<double> vals(n);
    int nb = 0, nf = 0;
    if (viorb_bow_transform(voc_, d.data(), n, levelsup, ids.data(), vals.data(), &nb, fvn.data(), fvp.data(), fvi.data(), &nf, nullptr,
                            nullptr) != VIORB_OK)
        throw std::runtime_error(std::string("viorb_bow_transform: ") + viorb_last_error());
    DBoW2::BowVector::iterator vit = v.end();
    for (int i = 0; i < nb; i++) vit = v.insert(vit, DBoW2::BowVector::value_type((unsigned)ids[i], vals[i]));
    for (int j = 0; j < nf; j++) {
        std::vector<unsigned int>& list = fv[(unsigned)fvn[j]];
        list.assign(fvi.begin() + fvp[j], fvi.begin() + fvp[j + 1]);
    }
}

}  // namespace ORB_SLAM2
