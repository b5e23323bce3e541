/*
 * bow_oracle.cpp -- CPU oracle for the DBoW2 vocabulary transform.  TEST INFRASTRUCTURE ONLY (see orb_oracle.h).
 *
 * Restates /root/reference/Thirdparty/DBoW2/DBoW2/TemplatedVocabulary.h (loadFromTextFile :1351-1437, transform
 * :1138-1204 and :1230-1272), BowVector.cpp (addWeight :34-46, addIfNotExist :50-58, normalize :62-86),
 * FeatureVector.cpp (addFeature :31-45), FORB.cpp (distance :81-101) and ScoringObject.h (mustNormalize :74-90) on
 * plain arrays, with the same containers (std::map, std::vector) so that insertion and summation orders are the
 * reference's.  DBoW2 cannot be compiled here (its headers need OpenCV) and the ORB vocabulary file is not part of
 * the repository: parity unpinned for this row; the tests cross-check against an independent numpy statement.
 */
#include <cmath>
#include <cstring>
#include <map>
#include <vector>

#include "orb_oracle.h"

namespace {
struct Node {
    int parent = 0;
    std::vector<int> children;
    uint8_t descriptor[32];
    double weight = 0;
    int word_id = -1;
    bool isLeaf() const { return children.empty(); }
};
}  // namespace

struct orc_vocabulary {
    int k, L, weighting, scoring;
    std::vector<Node> nodes;
    std::vector<int> words;
};

extern "C" orc_vocabulary* orc_vocabulary_create(int k, int L, int weighting, int scoring, int nnodes, const int32_t* parent,
                                                 const uint8_t* desc, const double* weight) {
    orc_vocabulary* v = new orc_vocabulary;
    v->k = k; v->L = L; v->weighting = weighting; v->scoring = scoring;
    v->nodes.resize(nnodes);
    for (int nid = 1; nid < nnodes; nid++) {          /* :1392-1413 */
        v->nodes[nid].parent = parent[nid];
        v->nodes[parent[nid]].children.push_back(nid);
        memcpy(v->nodes[nid].descriptor, desc + (size_t)nid * 32, 32);
        v->nodes[nid].weight = weight[nid];
    }
    for (int nid = 1; nid < nnodes; nid++)            /* :1415-1425; a node is a word iff it is a leaf */
        if (v->nodes[nid].isLeaf()) {
            v->nodes[nid].word_id = (int)v->words.size();
            v->words.push_back(nid);
        }
    return v;
}

extern "C" void orc_vocabulary_destroy(orc_vocabulary* v) { delete v; }

/* transform(feature, word_id, weight, nid, levelsup), :1230-1272 */
static void transform_one(const orc_vocabulary* v, const uint8_t* feature, int& word_id, double& weight, int* nid, int levelsup) {
    const int nid_level = v->L - levelsup;
    if (nid_level <= 0 && nid != NULL) *nid = 0;
    int final_id = 0, current_level = 0;
    do {
        ++current_level;
        const std::vector<int>& nodes = v->nodes[final_id].children;
        final_id = nodes[0];
        double best_d = orc_descriptor_distance(feature, v->nodes[final_id].descriptor);
        for (size_t i = 1; i < nodes.size(); i++) {
            const int id = nodes[i];
            const double d = orc_descriptor_distance(feature, v->nodes[id].descriptor);
            if (d < best_d) {
                best_d = d;
                final_id = id;
            }
        }
        if (nid != NULL && current_level == nid_level) *nid = final_id;
    } while (!v->nodes[final_id].isLeaf());
    word_id = v->nodes[final_id].word_id;
    weight = v->nodes[final_id].weight;
}

/* transform(features, v, fv, levelsup), :1138-1204.  Returns the BowVector size; *nfv = FeatureVector size. */
extern "C" int orc_bow_transform(const orc_vocabulary* voc, const uint8_t* desc, int n, int levelsup, int32_t* bow_ids,
                                 double* bow_values, int32_t* fv_node, int32_t* fv_ptr, int32_t* fv_idx, int* nfv,
                                 int32_t* word_of, int32_t* node_of) {
    std::map<unsigned, double> v;
    std::map<unsigned, std::vector<unsigned> > fv;
    *nfv = 0;
    fv_ptr[0] = 0;
    if (voc->words.empty()) return 0;
    const bool must = voc->scoring != 5;               /* ScoringObject.h:74-90 */
    const bool l2 = voc->scoring == 1;
    for (int i = 0; i < n; i++) {
        int id = 0, nid = 0;                           /* *nid stays unset in the reference when a leaf sits above nid_level */
        double w = 0;
        transform_one(voc, desc + (size_t)i * 32, id, w, &nid, levelsup);
        if (word_of) word_of[i] = id;
        if (node_of) node_of[i] = nid;
        if (w > 0) {
            if (voc->weighting == 0 || voc->weighting == 1) {           /* addWeight */
                std::map<unsigned, double>::iterator vit = v.lower_bound(id);
                if (vit != v.end() && !(v.key_comp()(id, vit->first))) vit->second += w;
                else v.insert(vit, std::make_pair((unsigned)id, w));
            } else {                                                    /* addIfNotExist */
                std::map<unsigned, double>::iterator vit = v.lower_bound(id);
                if (vit == v.end() || v.key_comp()(id, vit->first)) v.insert(vit, std::make_pair((unsigned)id, w));
            }
            fv[nid].push_back(i);                                       /* addFeature */
        }
    }
    if ((voc->weighting == 0 || voc->weighting == 1) && !v.empty() && !must) {
        const double nd = v.size();
        for (std::map<unsigned, double>::iterator vit = v.begin(); vit != v.end(); vit++) vit->second /= nd;
    }
    if (must) {                                                         /* BowVector::normalize */
        double norm = 0.0;
        if (!l2) {
            for (std::map<unsigned, double>::iterator it = v.begin(); it != v.end(); ++it) norm += fabs(it->second);
        } else {
            for (std::map<unsigned, double>::iterator it = v.begin(); it != v.end(); ++it) norm += it->second * it->second;
            norm = sqrt(norm);
        }
        if (norm > 0.0)
            for (std::map<unsigned, double>::iterator it = v.begin(); it != v.end(); ++it) it->second /= norm;
    }
    int k = 0;
    for (std::map<unsigned, double>::iterator it = v.begin(); it != v.end(); ++it, ++k) {
        bow_ids[k] = (int)it->first;
        bow_values[k] = it->second;
    }
    int f = 0, pos = 0;
    for (std::map<unsigned, std::vector<unsigned> >::iterator it = fv.begin(); it != fv.end(); ++it, ++f) {
        fv_node[f] = (int)it->first;
        fv_ptr[f] = pos;
        for (size_t j = 0; j < it->second.size(); j++) fv_idx[pos++] = (int)it->second[j];
    }
    fv_ptr[f] = pos;
    *nfv = f;
    return k;
}
