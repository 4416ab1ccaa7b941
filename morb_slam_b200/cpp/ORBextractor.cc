// Host side of the drop-in ORBextractor: forwards to the C ABI of liborb_b200.so.
#include "ORBextractor.h"

#include <cstring>
#include <stdexcept>
#include <string>

#include "orb_b200.h"

namespace ORB_SLAM3 {

static_assert(sizeof(cv::KeyPoint) == sizeof(orb_keypoint), "cv::KeyPoint must be the 28-byte record of the C ABI");

static void Check(orb_handle* h, int st, const char* what) {
  if (st != ORB_OK)
    throw std::runtime_error(std::string(what) + ": " + orb_status_string(st) + " - " + (h ? orb_last_error(h) : ""));
}

ORBextractor::ORBextractor(int _nfeatures, float _scaleFactor, int _nlevels, int _iniThFAST, int _minThFAST)
    : nfeatures(_nfeatures), scaleFactor(_scaleFactor), nlevels(_nlevels), iniThFAST(_iniThFAST), minThFAST(_minThFAST),
      mpHandle(nullptr), mnMaxW(0), mnMaxH(0), mnDevice(0), mbDownloadPyramid(true), mpPinnedKeys(nullptr), mpPinnedDesc(nullptr), mpPinnedStereo(nullptr), mnLastN(0), mnLastMono(0) {
  mvImagePyramid.resize(nlevels);
  // the tables are filled HERE like in the reference (src/ORBextractor.cc:413-443): every Frame constructor copies the getters'
  // results before the first extraction (src/Frame.cc:181-187). Pure host arithmetic, no device needed.
  mvScaleFactor.resize(nlevels); mvInvScaleFactor.resize(nlevels); mvLevelSigma2.resize(nlevels);
  mvInvLevelSigma2.resize(nlevels); mnFeaturesPerLevel.resize(nlevels);
  orb_params p;
  p.nfeatures = nfeatures; p.scale_factor = (float)scaleFactor; p.nlevels = nlevels;
  p.ini_th_fast = iniThFAST; p.min_th_fast = minThFAST;
  Check(nullptr, orb_compute_tables(&p, mvScaleFactor.data(), mvInvScaleFactor.data(), mvLevelSigma2.data(), mvInvLevelSigma2.data(),
                                    mnFeaturesPerLevel.data()), "orb_compute_tables");
}

ORBextractor::~ORBextractor() {
  if (mpPinnedKeys) orb_host_free(mpPinnedKeys);
  if (mpPinnedDesc) orb_host_free(mpPinnedDesc);
  if (mpPinnedStereo) orb_host_free(mpPinnedStereo);
  if (mpHandle) orb_destroy(mpHandle);
}

// The reference constructor does not know the image size; the DEVICE handle (only that - the tables exist since the
// constructor) is created on the first call and re-created only if a larger image arrives.
void ORBextractor::EnsureHandle(int width, int height) {
  if (mpHandle && width <= mnMaxW && height <= mnMaxH) return;
  if (mpHandle) { orb_destroy(mpHandle); mpHandle = nullptr; }
  orb_params p;
  p.nfeatures = nfeatures; p.scale_factor = (float)scaleFactor; p.nlevels = nlevels;
  p.ini_th_fast = iniThFAST; p.min_th_fast = minThFAST;
  mnMaxW = width > mnMaxW ? width : mnMaxW;
  mnMaxH = height > mnMaxH ? height : mnMaxH;
  Check(nullptr, orb_create(&p, mnMaxW, mnMaxH, 1, mnDevice, &mpHandle), "orb_create");
  if (!mpPinnedKeys) {   // sized by the constructor arguments alone (nfeatures + 3 * nlevels records): survives a re-created handle
    const size_t cap = (size_t)orb_keypoint_capacity(mpHandle);
    Check(mpHandle, orb_host_alloc(&mpPinnedKeys, cap * sizeof(orb_keypoint)), "orb_host_alloc");
    Check(mpHandle, orb_host_alloc(&mpPinnedDesc, cap * 32), "orb_host_alloc");
  }
}

// operator() in two halves: Enqueue uploads the image and queues the whole extraction plus the downloads into the page-locked
// staging buffers on the extractor's stream (ORB_ASYNC), Collect waits for it and fills the caller's containers.
int ORBextractor::Enqueue(const cv::Mat& image, std::vector<int>& vLappingArea) {
  if (image.type() != CV_8UC1) throw std::runtime_error("ORBextractor: image must be CV_8UC1");  // assert at :1014
  EnsureHandle(image.cols, image.rows);
  const int cap = orb_keypoint_capacity(mpHandle);
  mnLastN = 0; mnLastMono = -1;
  Check(mpHandle, orb_extract_batch(mpHandle, image.data, 1, image.cols, image.rows, (size_t)image.step, (size_t)image.step * image.rows,
                                    vLappingArea[0], vLappingArea[1], static_cast<orb_keypoint*>(mpPinnedKeys),
                                    static_cast<unsigned char*>(mpPinnedDesc), cap, &mnLastN, &mnLastMono, ORB_ASYNC), "orb_extract_batch");
  return 0;
}

int ORBextractor::Collect(std::vector<cv::KeyPoint>& _keypoints, cv::OutputArray _descriptors) {
  Check(mpHandle, orb_sync(mpHandle), "orb_sync");
  const int n = mnLastN;
  const orb_keypoint* kps = static_cast<orb_keypoint*>(mpPinnedKeys);
  const unsigned char* desc = static_cast<unsigned char*>(mpPinnedDesc);
  _keypoints.resize(n);
  if (n) std::memcpy((void*)_keypoints.data(), kps, (size_t)n * sizeof(orb_keypoint));
  if (n == 0) {
    _descriptors.release();  // :1028-1029
  } else {
    _descriptors.create(n, 32, CV_8U);
    cv::Mat d = _descriptors.getMat();
    if (d.isContinuous()) std::memcpy(d.data, desc, (size_t)n * 32);
    else for (int i = 0; i < n; ++i) std::memcpy(d.ptr(i), desc + (size_t)i * 32, 32);
  }
  if (mbDownloadPyramid) {
    for (int l = 0; l < nlevels; ++l) {
      int w = 0, h = 0;
      Check(mpHandle, orb_pyramid_level_size(mpHandle, l, &w, &h), "orb_pyramid_level_size");
      mvImagePyramid[l].create(h, w, CV_8UC1);
      Check(mpHandle, orb_pyramid_level(mpHandle, 0, l, mvImagePyramid[l].data, (size_t)mvImagePyramid[l].step),
            "orb_pyramid_level");
    }
  }
  return mnLastMono;
}

int ORBextractor::operator()(cv::InputArray _image, cv::InputArray _mask, std::vector<cv::KeyPoint>& _keypoints,
                             cv::OutputArray _descriptors, std::vector<int>& vLappingArea) {
  (void)_mask;
  if (_image.empty()) return -1;  // src/ORBextractor.cc:1011
  Enqueue(_image.getMat(), vLappingArea);
  return Collect(_keypoints, _descriptors);
}

int ORBextractor::ExtractPair(ORBextractor* pLeft, ORBextractor* pRight, cv::InputArray imLeft, cv::InputArray imRight,
                              std::vector<cv::KeyPoint>& vKeysLeft, cv::OutputArray descLeft, std::vector<cv::KeyPoint>& vKeysRight,
                              cv::OutputArray descRight, std::vector<int>& vLappingLeft, std::vector<int>& vLappingRight, int* pMonoRight) {
  if (imLeft.empty() || imRight.empty()) {   // each operator() would return -1 for its own empty image
    const int ml = (*pLeft)(imLeft, cv::Mat(), vKeysLeft, descLeft, vLappingLeft);
    const int mr = (*pRight)(imRight, cv::Mat(), vKeysRight, descRight, vLappingRight);
    if (pMonoRight) *pMonoRight = mr;
    return ml;
  }
  pLeft->Enqueue(imLeft.getMat(), vLappingLeft);
  pRight->Enqueue(imRight.getMat(), vLappingRight);
  const int ml = pLeft->Collect(vKeysLeft, descLeft);
  const int mr = pRight->Collect(vKeysRight, descRight);
  if (pMonoRight) *pMonoRight = mr;
  return ml;
}

void ComputeStereoMatchesB200(ORBextractor* pLeft, ORBextractor* pRight, const std::vector<cv::KeyPoint>& vKeysLeft,
                              const cv::Mat& descLeft, const std::vector<cv::KeyPoint>& vKeysRight,
                              const cv::Mat& descRight, float mbf, float maxD, std::vector<float>& vuRight,
                              std::vector<float>& vDepth) {
  const int nL = (int)vKeysLeft.size(), nR = (int)vKeysRight.size();
  vuRight.assign(nL, -1.0f);
  vDepth.assign(nL, -1.0f);
  if (nL == 0 || nR == 0) return;
  // descriptors are N x 32 CV_8U; rows are contiguous for Mats created by operator()
  std::vector<unsigned char> dl((size_t)nL * 32), dr((size_t)nR * 32);
  for (int i = 0; i < nL; ++i) std::memcpy(&dl[(size_t)i * 32], descLeft.ptr(i), 32);
  for (int i = 0; i < nR; ++i) std::memcpy(&dr[(size_t)i * 32], descRight.ptr(i), 32);
  Check(pLeft->Handle(), orb_stereo_match(pLeft->Handle(), pRight->Handle(), (const orb_keypoint*)vKeysLeft.data(), dl.data(), nL,
                                          (const orb_keypoint*)vKeysRight.data(), dr.data(), nR, mbf, maxD, vuRight.data(),
                                          vDepth.data()), "orb_stereo_match");
}

void ComputeStereoMatchesB200(ORBextractor* pLeft, ORBextractor* pRight, int nLeft, float mbf, float maxD, std::vector<float>& vuRight,
                              std::vector<float>& vDepth) {
  vuRight.assign(nLeft, -1.0f);
  vDepth.assign(nLeft, -1.0f);
  if (nLeft == 0) return;
  // results land in page-locked staging (written by the last kernel itself for a single frame) and leave as two memcpys
  const int cap = orb_keypoint_capacity(pLeft->Handle());
  float* st = pLeft->StereoStaging();
  Check(pLeft->Handle(), orb_stereo_match_batch(pLeft->Handle(), pRight->Handle(), mbf, maxD, st, st + cap, cap, 0), "orb_stereo_match_batch");
  std::memcpy(vuRight.data(), st, (size_t)nLeft * sizeof(float));
  std::memcpy(vDepth.data(), st + cap, (size_t)nLeft * sizeof(float));
}

float* ORBextractor::StereoStaging() {
  if (!mpPinnedStereo) Check(mpHandle, orb_host_alloc(&mpPinnedStereo, (size_t)orb_keypoint_capacity(mpHandle) * 2 * sizeof(float)), "orb_host_alloc");
  return static_cast<float*>(mpPinnedStereo);
}

int DescriptorDistanceB200(const cv::Mat& a, const cv::Mat& b) { return orb_hamming_distance(a.ptr(), b.ptr()); }

}  // namespace ORB_SLAM3
