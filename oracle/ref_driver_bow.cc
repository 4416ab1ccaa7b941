// TEST INFRASTRUCTURE ONLY - never linked, imported or executed by the product (morb_slam_b200/, include/).
// C driver around the UNMODIFIED DBoW2 of the reference (Thirdparty/DBoW2/DBoW2/{TemplatedVocabulary.h, FORB.cpp,
// BowVector.cpp, FeatureVector.cpp, ScoringObject.cpp}, DUtils/Random.cpp), compiled where it lies under /root/reference
// against the oracle's cv::Mat shim (+ oracle/shim_dbow: FileStorage / Boost declarations that are never exercised).
// Frame::ComputeBoW (src/Frame.cc:822-827) = Converter::toDescriptorVector (one 1 x 32 Mat per row) +
// ORBVocabulary::transform(vCurrentDesc, mBowVec, mFeatVec, 4); the vocabulary is loaded with loadFromTextFile
// (src/System.cc:132) from the ORBvoc.txt text format.
#include <cstdint>
#include <cstring>
#include <vector>

#include <opencv2/core/core.hpp>
#include "DBoW2/FORB.h"
#include "DBoW2/TemplatedVocabulary.h"

typedef DBoW2::TemplatedVocabulary<DBoW2::FORB::TDescriptor, DBoW2::FORB> ORBVocabulary;  // include/ORBVocabulary.h:30-31

extern "C" {

void* refb_vocab_load_text(const char* filename) {
  ORBVocabulary* v = new ORBVocabulary();
  if (!v->loadFromTextFile(filename)) { delete v; return nullptr; }
  return v;
}

void refb_vocab_free(void* v) { delete (ORBVocabulary*)v; }

// k, L, scoring, weighting, number of words
void refb_vocab_info(void* v_, int* info5) {
  ORBVocabulary* v = (ORBVocabulary*)v_;
  info5[0] = v->getBranchingFactor(); info5[1] = v->getDepthLevels(); info5[2] = (int)v->getScoringType();
  info5[3] = (int)v->getWeightingType(); info5[4] = (int)v->size();
}

// Frame::ComputeBoW on n descriptors (n x 32 bytes). Outputs in map order:
//   bow_word / bow_val [cap]: BowVector entries, *bow_n of them
//   fv_node [cap], fv_off [cap + 1], fv_feat [cap]: FeatureVector as CSR, *fv_n nodes
// Returns 0, or -1 when cap is too small.
int refb_transform(void* v_, const uint8_t* desc, int n, int levelsup, int cap, uint32_t* bow_word, double* bow_val, int* bow_n,
                   uint32_t* fv_node, int* fv_off, uint32_t* fv_feat, int* fv_n) {
  ORBVocabulary* v = (ORBVocabulary*)v_;
  std::vector<cv::Mat> vDesc;                     // Converter::toDescriptorVector (src/Converter.cc): one row per Mat
  vDesc.reserve(n);
  for (int i = 0; i < n; ++i) {
    cv::Mat m(1, 32, CV_8UC1);
    std::memcpy(m.data, desc + 32 * (size_t)i, 32);
    vDesc.push_back(m);
  }
  DBoW2::BowVector bv;
  DBoW2::FeatureVector fv;
  v->transform(vDesc, bv, fv, levelsup);
  if ((int)bv.size() > cap || (int)fv.size() > cap) return -1;
  int i = 0;
  for (DBoW2::BowVector::const_iterator it = bv.begin(); it != bv.end(); ++it, ++i) { bow_word[i] = it->first; bow_val[i] = it->second; }
  *bow_n = i;
  int j = 0, o = 0;
  for (DBoW2::FeatureVector::const_iterator it = fv.begin(); it != fv.end(); ++it, ++j) {
    fv_node[j] = it->first;
    fv_off[j] = o;
    for (size_t t = 0; t < it->second.size(); ++t) {
      if (o >= cap) return -1;
      fv_feat[o++] = it->second[t];
    }
  }
  fv_off[j] = o;
  *fv_n = j;
  return 0;
}

}  // extern "C"
