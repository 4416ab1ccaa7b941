// TEST INFRASTRUCTURE ONLY - never linked, imported or executed by the product (morb_slam_b200/, include/).
// CPU restatement of Frame::ComputeBoW (reference src/Frame.cc:822-827): DBoW2's
// TemplatedVocabulary::transform(features, BowVector, FeatureVector, levelsup)
// (Thirdparty/DBoW2/DBoW2/TemplatedVocabulary.h:1126-1200), the per-feature tree descent (:1217-1260),
// FORB::distance (FORB.cpp:81-101), BowVector::addWeight / addIfNotExist / normalize (BowVector.cpp:34-98) and
// FeatureVector::addFeature (FeatureVector.cpp:34-48), in the formulation of the CUDA kernels: descent per feature,
// then sort by (word, feature) / (node, feature) instead of std::map insertion. Pinned against the reference's own
// DBoW2 by tests/test_oracle_bow.py (oracle/_ref/libmorb_ref_bow.so).
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <vector>

#include "orb_oracle.h"

namespace {

struct Vocab {
  int k, L, scoring, weighting;
  std::vector<std::vector<int>> children;   // in file order (loadFromTextFile pushes the node id onto its parent, :1388)
  std::vector<uint8_t> desc;
  std::vector<double> weight;
  std::vector<uint32_t> word_id;            // 0 for nodes that are not flagged as leaves (Node(): word_id(0))
  int nwords;
};

inline int hamming256(const uint8_t* a, const uint8_t* b) {
  int d = 0;
  for (int i = 0; i < 32; i += 8) {
    uint64_t x, y;
    std::memcpy(&x, a + i, 8);
    std::memcpy(&y, b + i, 8);
    d += __builtin_popcountll(x ^ y);
  }
  return d;
}

}  // namespace

extern "C" {

void* oro_vocab_create(int k, int L, int scoring, int weighting, int n_nodes, const int32_t* parent, const uint8_t* is_leaf,
                       const uint8_t* desc, const double* weight) {
  Vocab* v = new Vocab;
  v->k = k; v->L = L; v->scoring = scoring; v->weighting = weighting;
  v->children.resize(n_nodes);
  v->desc.assign(desc, desc + (size_t)n_nodes * 32);
  v->weight.assign(weight, weight + n_nodes);
  v->word_id.assign(n_nodes, 0);
  v->nwords = 0;
  for (int i = 1; i < n_nodes; ++i) {
    v->children[parent[i]].push_back(i);
    if (is_leaf[i]) v->word_id[i] = v->nwords++;      // :1408-1415: words are numbered in file order
  }
  return v;
}

void oro_vocab_free(void* v) { delete (Vocab*)v; }

int oro_bow_transform(void* v_, const uint8_t* desc, int n, int levelsup, int cap, uint32_t* bow_word, double* bow_val, int* bow_n,
                      uint32_t* fv_node, int* fv_off, uint32_t* fv_feat, int* fv_n, int32_t* feat_word, int32_t* feat_node) {
  const Vocab& v = *(const Vocab*)v_;
  *bow_n = 0; *fv_n = 0; fv_off[0] = 0;
  if (v.nwords == 0) return 0;                              // empty() (:1133)
  std::vector<uint64_t> kw, kn;                             // (word << 32 | feature), (node << 32 | feature)
  std::vector<double> wv(n, 0.0);
  const int nid_level = v.L - levelsup;
  for (int f = 0; f < n; ++f) {
    const uint8_t* d = desc + 32 * (size_t)f;
    int final_id = 0, level = 0, nid = 0;                   // nid_level <= 0 -> root (:1228)
    do {                                                    // :1233-1256
      ++level;
      const std::vector<int>& ch = v.children[final_id];
      final_id = ch[0];
      int best = hamming256(d, &v.desc[32 * (size_t)final_id]);
      for (size_t c = 1; c < ch.size(); ++c) {
        const int dist = hamming256(d, &v.desc[32 * (size_t)ch[c]]);
        if (dist < best) { best = dist; final_id = ch[c]; }
      }
      if (level == nid_level) nid = final_id;
    } while (!v.children[final_id].empty());
    // a leaf above nid_level leaves *nid unset in the reference (uninitialised local, :1149): defined here as that leaf
    if (nid_level > 0 && level < nid_level) nid = final_id;
    const double w = v.weight[final_id];
    if (feat_word) feat_word[f] = w > 0 ? (int32_t)v.word_id[final_id] : -1;
    if (feat_node) feat_node[f] = w > 0 ? nid : -1;
    if (w > 0) {                                            // not stopped (:1157)
      kw.push_back(((uint64_t)v.word_id[final_id] << 32) | (uint32_t)f);
      kn.push_back(((uint64_t)(uint32_t)nid << 32) | (uint32_t)f);
      wv[f] = w;
    }
  }
  std::sort(kw.begin(), kw.end());
  std::sort(kn.begin(), kn.end());
  // BowVector: TF_IDF / TF add the weight once per feature in feature order (addWeight), IDF / BINARY keep the first
  int nb = 0;
  for (size_t i = 0; i < kw.size();) {
    const uint32_t word = (uint32_t)(kw[i] >> 32);
    double val = wv[(uint32_t)kw[i]];
    size_t j = i + 1;
    for (; j < kw.size() && (uint32_t)(kw[j] >> 32) == word; ++j)
      if (v.weighting == 0 || v.weighting == 1) val += wv[(uint32_t)kw[j]];
    if (nb >= cap) return -1;
    bow_word[nb] = word; bow_val[nb] = val; ++nb;
    i = j;
  }
  const bool must = v.scoring != 5;                         // every scoring but DOT_PRODUCT normalises (ScoringObject.h:74-89)
  const bool l2 = v.scoring == 1;
  if ((v.weighting == 0 || v.weighting == 1) && nb && !must) {   // :1163-1169
    const double nd = nb;
    for (int i = 0; i < nb; ++i) bow_val[i] /= nd;
  }
  if (must) {                                               // BowVector::normalize (:62-86)
    double norm = 0.0;
    if (!l2) for (int i = 0; i < nb; ++i) norm += std::fabs(bow_val[i]);
    else { for (int i = 0; i < nb; ++i) norm += bow_val[i] * bow_val[i]; norm = std::sqrt(norm); }
    if (norm > 0.0) for (int i = 0; i < nb; ++i) bow_val[i] /= norm;
  }
  *bow_n = nb;
  int nn = 0, o = 0;
  for (size_t i = 0; i < kn.size();) {
    const uint32_t node = (uint32_t)(kn[i] >> 32);
    if (nn >= cap) return -1;
    fv_node[nn] = node; fv_off[nn] = o; ++nn;
    for (; i < kn.size() && (uint32_t)(kn[i] >> 32) == node; ++i) fv_feat[o++] = (uint32_t)kn[i];
  }
  fv_off[nn] = o;
  *fv_n = nn;
  return 0;
}

}  // extern "C"
