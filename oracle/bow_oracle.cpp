// bow_oracle.cpp -- CPU oracle of DBoW2's image -> (BowVector, FeatureVector) conversion as Frame::ComputeBoW / KeyFrame::ComputeBoW use it
// (src/Frame.cc:393-408: mpORBvocabulary->transform(vCurrentDesc, mvBowVec[c], mvFeatVec[c], 4)).
//
// TEST INFRASTRUCTURE ONLY (see orb_oracle.h).  Restates, with the reference's own containers (std::map, std::vector):
//   TemplatedVocabulary::loadFromTextFile   Thirdparty/DBoW2/DBoW2/TemplatedVocabulary.h:1362-1447 (node table: parent, leaf flag, descriptor, weight)
//   TemplatedVocabulary::transform (image)  :1149-1227  (TF_IDF weighting, L1 normalisation: the ORB vocabulary's header "10 6 0 0")
//   TemplatedVocabulary::transform (feature):1249-1292  (tree descent, first child wins ties)
//   FORB::distance                          Thirdparty/DBoW2/DBoW2/FORB.cpp:82-102
//   BowVector::addWeight / normalize        Thirdparty/DBoW2/DBoW2/BowVector.cpp:36-88
//   FeatureVector::addFeature               Thirdparty/DBoW2/DBoW2/FeatureVector.cpp
//   L1Scoring::score                        Thirdparty/DBoW2/DBoW2/ScoringObject.cpp:23-68
// PARITY UNPINNED by the reference: the vocabulary file (ORBvoc.txt) is not in the repository and DBoW2 ships no test vectors.
#include <cmath>
#include <cstdint>
#include <cstring>
#include <map>
#include <vector>

#include "orb_oracle.h"

struct orc_vocab {
    int k, L;
    struct Node { int parent; std::vector<int> children; uint8_t desc[32]; double weight; int word_id; };
    std::vector<Node> nodes;
    int n_words;
};

namespace {
int forb_distance(const uint8_t* a, const uint8_t* b) {
    int dist = 0;
    for (int i = 0; i < 8; i++) {
        uint32_t pa, pb;
        std::memcpy(&pa, a + 4 * i, 4);
        std::memcpy(&pb, b + 4 * i, 4);
        unsigned int v = pa ^ pb;
        v = v - ((v >> 1) & 0x55555555);
        v = (v & 0x33333333) + ((v >> 2) & 0x33333333);
        dist += (((v + (v >> 4)) & 0xF0F0F0F) * 0x1010101) >> 24;
    }
    return dist;
}
}  // namespace

extern "C" {

// node 0 is the root (its row of parent / desc / weight is ignored); rows 1.. are the lines of the vocabulary text file in file order
orc_vocab* orc_vocab_create(int k, int L, int n_nodes, const int32_t* parent, const uint8_t* is_leaf, const uint8_t* desc, const double* weight) {
    orc_vocab* V = new orc_vocab;
    V->k = k; V->L = L; V->n_words = 0;
    V->nodes.resize(n_nodes);
    V->nodes[0].parent = -1; V->nodes[0].weight = 0; V->nodes[0].word_id = -1;
    for (int nid = 1; nid < n_nodes; nid++) {
        orc_vocab::Node& N = V->nodes[nid];
        N.parent = parent[nid];
        if (N.parent < 0 || N.parent >= nid) { delete V; return nullptr; }      // the file lists a parent before its children
        V->nodes[N.parent].children.push_back(nid);
        std::memcpy(N.desc, desc + 32 * (size_t)nid, 32);
        N.weight = weight[nid];
        N.word_id = is_leaf[nid] ? V->n_words++ : -1;
    }
    return V;
}
void orc_vocab_destroy(orc_vocab* V) { delete V; }
int orc_vocab_words(const orc_vocab* V) { return V->n_words; }

// transform(features, v, fv, levelsup).  word_id / node_id [n] per feature (for stage parity); BowVector as (ids ascending, values);
// FeatureVector as CSR (node ids ascending; fv_off [n_fv_nodes + 1]; fv_idx feature indices in insertion order).
int orc_vocab_transform(const orc_vocab* V, const uint8_t* desc, int n, int levelsup, int32_t* word_id, int32_t* node_id, int32_t* bow_ids,
                        double* bow_vals, int32_t* n_words, int32_t* fv_node, int32_t* fv_off, int32_t* fv_idx, int32_t* n_fv_nodes) {
    std::map<unsigned, double> v;
    std::map<unsigned, std::vector<unsigned>> fv;
    const int nid_level = V->L - levelsup;
    for (int i = 0; i < n; i++) {
        const uint8_t* feature = desc + 32 * (size_t)i;
        int nid = 0;                                     // (`if(nid_level <= 0) *nid = 0`; left at the leaf's ancestor otherwise)
        int final_id = 0, current_level = 0;
        do {
            ++current_level;
            const std::vector<int>& nodes = V->nodes[final_id].children;
            final_id = nodes[0];
            double best_d = forb_distance(feature, V->nodes[final_id].desc);
            for (size_t j = 1; j < nodes.size(); j++) {
                const int id = nodes[j];
                const double d = forb_distance(feature, V->nodes[id].desc);
                if (d < best_d) { best_d = d; final_id = id; }
            }
            if (current_level == nid_level) nid = final_id;
        } while (!V->nodes[final_id].children.empty());
        const int id = V->nodes[final_id].word_id;
        const double w = V->nodes[final_id].weight;
        word_id[i] = id; node_id[i] = nid;
        if (w > 0) {
            v[(unsigned)id] += w;                        // BowVector::addWeight
            fv[(unsigned)nid].push_back((unsigned)i);    // FeatureVector::addFeature
        }
    }
    double norm = 0.0;                                   // BowVector::normalize(L1)
    for (auto& e : v) norm += std::fabs(e.second);
    if (norm > 0.0) for (auto& e : v) e.second /= norm;
    int m = 0;
    for (auto& e : v) { bow_ids[m] = (int32_t)e.first; bow_vals[m] = e.second; m++; }
    *n_words = m;
    int f = 0, o = 0;
    fv_off[0] = 0;
    for (auto& e : fv) {
        fv_node[f] = (int32_t)e.first;
        for (unsigned i : e.second) fv_idx[o++] = (int32_t)i;
        fv_off[++f] = o;
    }
    *n_fv_nodes = f;
    return 0;
}

// L1Scoring::score  (ScoringObject.cpp:23-68) on two BowVectors given as ascending (id, value) arrays
double orc_bow_score_l1(const int32_t* ids1, const double* v1, int n1, const int32_t* ids2, const double* v2, int n2) {
    int a = 0, b = 0;
    double score = 0;
    while (a < n1 && b < n2) {
        if (ids1[a] == ids2[b]) { score += std::fabs(v1[a] - v2[b]) - std::fabs(v1[a]) - std::fabs(v2[b]); a++; b++; }
        else if (ids1[a] < ids2[b]) { while (a < n1 && ids1[a] < ids2[b]) a++; }
        else { while (b < n2 && ids2[b] < ids1[a]) b++; }
    }
    return -score / 2.0;
}

}  // extern "C"
