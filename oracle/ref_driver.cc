// TEST INFRASTRUCTURE ONLY (oracle/_ref). Not part of the product path.
//
// C entry points around the UNMODIFIED reference sources, compiled where they lie under
// /root/reference by oracle/Makefile (outputs only into oracle/_ref/):
//   * src/ORBextractor.cc (whole file) against the oracle's OpenCV shim,
//   * src/Frame.cc:889-1047 (Frame::ComputeStereoMatches), src/ORBmatcher.cc:34-36 (thresholds) and
//     src/ORBmatcher.cc:1880-1894 (ORBmatcher::DescriptorDistance), cut out by line range into
//     oracle/_ref/*.inc at build time and included into the stub Frame/ORBmatcher below.
// Nothing of the reference is copied into the repository.
#include <cstdint>
#include <cstring>
#include <list>
#include <vector>
#include <algorithm>
#include <utility>
#include <climits>
#include <cmath>

#include "ORBextractor.h"  // from /root/reference/include

using namespace std;

namespace ORB_SLAM3 {

// Derived class only to reach the protected stage functions (include/ORBextractor.h:78-88)
class RefExtractor : public ORBextractor {
 public:
  using ORBextractor::ORBextractor;
  void pyramid(cv::Mat im) { ComputePyramid(im); }
  void keypoints(std::vector<std::vector<cv::KeyPoint>>& all) { ComputeKeyPointsOctTree(all); }
  std::vector<cv::KeyPoint> distribute(const std::vector<cv::KeyPoint>& v, int minX, int maxX, int minY, int maxY,
                                       int N, int level) {
    return DistributeOctTree(v, minX, maxX, minY, maxY, N, level);
  }
  const std::vector<int>& featuresPerLevel() const { return mnFeaturesPerLevel; }
  const std::vector<int>& uMax() const { return umax; }
};

// ---- stub types for the line-range extraction of the stereo matcher -------------------------------
struct ORBmatcher {
  static const int TH_LOW;
  static const int TH_HIGH;
  static const int HISTO_LENGTH;
  static int DescriptorDistance(const cv::Mat& a, const cv::Mat& b);
};
#include "orbmatcher_consts.inc"  // src/ORBmatcher.cc:34-36

struct Frame {
  int N;
  std::vector<cv::KeyPoint> mvKeys, mvKeysRight;
  cv::Mat mDescriptors, mDescriptorsRight;
  std::vector<float> mvuRight, mvDepth;
  std::vector<float> mvScaleFactors, mvInvScaleFactors;
  ORBextractor *mpORBextractorLeft, *mpORBextractorRight;
  float mb, mbf;
  void ComputeStereoMatches();
};
#include "frame_stereo.inc"       // src/Frame.cc:889-1047
#include "orbmatcher_dist.inc"    // src/ORBmatcher.cc:1880-1894

}  // namespace ORB_SLAM3

using ORB_SLAM3::RefExtractor;

extern "C" {

void* ref_create(int nfeatures, float scale, int nlevels, int ini, int mn) {
  return new RefExtractor(nfeatures, scale, nlevels, ini, mn);
}
void ref_destroy(void* h) { delete (RefExtractor*)h; }

void ref_tables(void* h, float* scale, float* inv_scale, float* sigma2, float* inv_sigma2, int* nfeat, int* umax) {
  RefExtractor* e = (RefExtractor*)h;
  int n = e->GetLevels();
  std::vector<float> a = e->GetScaleFactors(), b = e->GetInverseScaleFactors(), c = e->GetScaleSigmaSquares(),
                     d = e->GetInverseScaleSigmaSquares();
  for (int i = 0; i < n; ++i) {
    scale[i] = a[i]; inv_scale[i] = b[i]; sigma2[i] = c[i]; inv_sigma2[i] = d[i];
    nfeat[i] = e->featuresPerLevel()[i];
  }
  for (int i = 0; i < 16; ++i) umax[i] = e->uMax()[i];
}

// full operator(); returns monoIndex (or -1), *n_out = number of keypoints
int ref_extract(void* h, const uint8_t* img, int w, int hh, int stride, int lap0, int lap1, void* kps_out,
                uint8_t* desc_out, int cap, int* n_out) {
  RefExtractor* e = (RefExtractor*)h;
  cv::Mat im = (img && w > 0 && hh > 0) ? cv::Mat(hh, w, CV_8UC1, (void*)img, (size_t)stride) : cv::Mat();
  std::vector<cv::KeyPoint> kps;
  cv::Mat desc;
  std::vector<int> lap = {lap0, lap1};
  int mono = (*e)(im, cv::Mat(), kps, desc, lap);
  if (mono < 0) { *n_out = 0; return mono; }
  int n = (int)kps.size();
  *n_out = n;
  if (n > cap) return -2;
  if (n) std::memcpy(kps_out, kps.data(), (size_t)n * sizeof(cv::KeyPoint));
  for (int i = 0; i < n; ++i) std::memcpy(desc_out + 32 * (size_t)i, desc.ptr(i), 32);
  return mono;
}

int ref_level_size(void* h, int level, int* w, int* hh) {
  RefExtractor* e = (RefExtractor*)h;
  *w = e->mvImagePyramid[level].cols; *hh = e->mvImagePyramid[level].rows;
  return 0;
}
int ref_get_level(void* h, int level, uint8_t* dst) {
  RefExtractor* e = (RefExtractor*)h;
  const cv::Mat& m = e->mvImagePyramid[level];
  for (int y = 0; y < m.rows; ++y) std::memcpy(dst + (size_t)y * m.cols, m.ptr(y), m.cols);
  return 0;
}

// ComputePyramid + ComputeKeyPointsOctTree: per-level keypoints in level coordinates, with angle
int ref_keypoints_per_level(void* h, const uint8_t* img, int w, int hh, int stride, void* kps_out, int cap,
                            int* counts) {
  RefExtractor* e = (RefExtractor*)h;
  cv::Mat im(hh, w, CV_8UC1, (void*)img, (size_t)stride);
  e->pyramid(im);
  std::vector<std::vector<cv::KeyPoint>> all;
  e->keypoints(all);
  int n = 0;
  for (size_t l = 0; l < all.size(); ++l) {
    counts[l] = (int)all[l].size();
    if (n + counts[l] > cap) return -2;
    if (counts[l]) std::memcpy((char*)kps_out + (size_t)n * sizeof(cv::KeyPoint), all[l].data(), counts[l] * sizeof(cv::KeyPoint));
    n += counts[l];
  }
  return n;
}

// DistributeOctTree on an arbitrary candidate list (fuzzing the quad-tree / std::sort behaviour)
int ref_distribute(void* h, const void* cands, int n, int minX, int maxX, int minY, int maxY, int N, int level,
                   void* out, int cap) {
  RefExtractor* e = (RefExtractor*)h;
  std::vector<cv::KeyPoint> v((const cv::KeyPoint*)cands, (const cv::KeyPoint*)cands + n);
  std::vector<cv::KeyPoint> r = e->distribute(v, minX, maxX, minY, maxY, N, level);
  if ((int)r.size() > cap) return -2;
  if (!r.empty()) std::memcpy(out, r.data(), r.size() * sizeof(cv::KeyPoint));
  return (int)r.size();
}

// the real libstdc++ std::sort with a comparator that, like compareNodes (src/ORBextractor.cc:525-538),
// looks at the key only; used to pin the oracle's introsort emulation
void ref_std_sort(uint32_t* keys, uint32_t* payload, int n) {
  std::vector<std::pair<uint32_t, uint32_t>> v(n);
  for (int i = 0; i < n; ++i) v[i] = std::make_pair(keys[i], payload[i]);
  std::sort(v.begin(), v.end(), [](std::pair<uint32_t, uint32_t>& a, std::pair<uint32_t, uint32_t>& b) { return a.first < b.first; });
  for (int i = 0; i < n; ++i) { keys[i] = v[i].first; payload[i] = v[i].second; }
}

int ref_descriptor_distance(const uint8_t* a, const uint8_t* b) {
  cv::Mat ma(1, 32, CV_8UC1, (void*)a), mb(1, 32, CV_8UC1, (void*)b);
  return ORB_SLAM3::ORBmatcher::DescriptorDistance(ma, mb);
}

// Frame::ComputeStereoMatches on the pyramids held by two extractor handles (their last call)
int ref_stereo(void* hl, void* hr, const void* kpsL, const uint8_t* descL, int nL, const void* kpsR,
               const uint8_t* descR, int nR, float mbf, float mb, float* uRight, float* depth) {
  RefExtractor* el = (RefExtractor*)hl;
  RefExtractor* er = (RefExtractor*)hr;
  ORB_SLAM3::Frame f;
  f.N = nL;
  f.mvKeys.assign((const cv::KeyPoint*)kpsL, (const cv::KeyPoint*)kpsL + nL);
  f.mvKeysRight.assign((const cv::KeyPoint*)kpsR, (const cv::KeyPoint*)kpsR + nR);
  f.mDescriptors = cv::Mat(std::max(nL, 1), 32, CV_8UC1);
  f.mDescriptorsRight = cv::Mat(std::max(nR, 1), 32, CV_8UC1);
  if (nL) std::memcpy(f.mDescriptors.data, descL, (size_t)nL * 32);
  if (nR) std::memcpy(f.mDescriptorsRight.data, descR, (size_t)nR * 32);
  f.mvScaleFactors = el->GetScaleFactors();
  f.mvInvScaleFactors = el->GetInverseScaleFactors();
  f.mpORBextractorLeft = el;
  f.mpORBextractorRight = er;
  f.mb = mb;
  f.mbf = mbf;
  f.ComputeStereoMatches();
  for (int i = 0; i < nL; ++i) { uRight[i] = f.mvuRight[i]; depth[i] = f.mvDepth[i]; }
  return 0;
}

}  // extern "C"
