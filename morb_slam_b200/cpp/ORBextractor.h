// Drop-in replacement for the reference's include/ORBextractor.h (class ORB_SLAM3::ORBextractor,
// reference include/ORBextractor.h:44-105) backed by the B200 C ABI (include/orb_b200.h).
// Same constructor, operator(), getters and public mvImagePyramid, so Tracking.cc (:615-624, :1224-1233)
// and Frame.cc (:181-187, :535-537) compile and link against it unchanged.
//
// Compiles against OpenCV's core headers (cv::Mat, cv::KeyPoint, cv::InputArray/OutputArray only - no
// OpenCV algorithm is called).
#ifndef ORBEXTRACTOR_H
#define ORBEXTRACTOR_H

#include <opencv2/core/core.hpp>
#include <vector>

struct orb_handle;

namespace ORB_SLAM3 {

class ORBextractor {
 public:
  enum { HARRIS_SCORE = 0, FAST_SCORE = 1 };

  ORBextractor(int nfeatures, float scaleFactor, int nlevels, int iniThFAST, int minThFAST);
  ~ORBextractor();
  ORBextractor(const ORBextractor&) = delete;
  ORBextractor& operator=(const ORBextractor&) = delete;

  // Compute the ORB features and descriptors on an image (mask is ignored, as in the reference).
  // Returns monoIndex (src/ORBextractor.cc:1085), -1 for an empty image. Throws std::runtime_error
  // when the CUDA path fails (there is no CPU fallback).
  int operator()(cv::InputArray _image, cv::InputArray _mask, std::vector<cv::KeyPoint>& _keypoints,
                 cv::OutputArray _descriptors, std::vector<int>& vLappingArea);

  int inline GetLevels() { return nlevels; }
  float inline GetScaleFactor() { return (float)scaleFactor; }
  std::vector<float> inline GetScaleFactors() { return mvScaleFactor; }
  std::vector<float> inline GetInverseScaleFactors() { return mvInvScaleFactor; }
  std::vector<float> inline GetScaleSigmaSquares() { return mvLevelSigma2; }
  std::vector<float> inline GetInverseScaleSigmaSquares() { return mvInvLevelSigma2; }

  // Un-blurred pyramid of the last call (read by the reference's Frame::ComputeStereoMatches,
  // src/Frame.cc:895,974-994). Filled from the device after every call unless disabled.
  std::vector<cv::Mat> mvImagePyramid;

  // ---- additions (not in the reference) ----
  // Both images of a stereo frame in one go: what the Frame constructor does with two std::threads (src/Frame.cc:194-197:
  // thread threadLeft(&Frame::ExtractORB, this, 0, imLeft, ...), threadRight(..., 1, imRight, ...); join both) on the caller's
  // thread - both extractions are enqueued on their extractors' streams before the first synchronisation, so they overlap on the
  // device without host threads. Results as from the two operator() calls; returns the left monoIndex, *pMonoRight the right one.
  static int ExtractPair(ORBextractor* pLeft, ORBextractor* pRight, cv::InputArray imLeft, cv::InputArray imRight,
                         std::vector<cv::KeyPoint>& vKeysLeft, cv::OutputArray descLeft, std::vector<cv::KeyPoint>& vKeysRight,
                         cv::OutputArray descRight, std::vector<int>& vLappingLeft, std::vector<int>& vLappingRight, int* pMonoRight);

  // Skip the device->host copy of the pyramid when the stereo matcher below is used instead of the
  // reference's CPU ComputeStereoMatches.
  void SetDownloadPyramid(bool on) { mbDownloadPyramid = on; }
  void SetDevice(int device) { mnDevice = device; }
  orb_handle* Handle() { return mpHandle; }

 protected:
  void EnsureHandle(int width, int height);
  int Enqueue(const cv::Mat& image, std::vector<int>& vLappingArea);   // asynchronous half of operator()
  int Collect(std::vector<cv::KeyPoint>& _keypoints, cv::OutputArray _descriptors);   // synchronises, fills the containers

  int nfeatures;
  double scaleFactor;
  int nlevels;
  int iniThFAST;
  int minThFAST;

  std::vector<int> mnFeaturesPerLevel;
  std::vector<float> mvScaleFactor;
  std::vector<float> mvInvScaleFactor;
  std::vector<float> mvLevelSigma2;
  std::vector<float> mvInvLevelSigma2;

  orb_handle* mpHandle;
  int mnMaxW, mnMaxH, mnDevice;
  bool mbDownloadPyramid;
  // page-locked staging of the results (keypoints, descriptors), allocated with the handle: the device-to-host copies of a call
  // land here at full PCIe speed and leave as two memcpys into the caller's containers
  void* mpPinnedKeys;
  void* mpPinnedDesc;
 public:
  float* StereoStaging();   // page-locked 2 x capacity floats (mvuRight | mvDepth) of ComputeStereoMatchesB200, allocated on first use
 protected:
  void* mpPinnedStereo;
  int mnLastN, mnLastMono;   // results of the asynchronous call in flight (written by orb_sync)
};

// Frame::ComputeStereoMatches (src/Frame.cc:889-1047) on the two extractors' device pyramids.
// maxD = mbf / mb: the reference reads mb before assigning it (src/Frame.cc:915 vs :253).
void ComputeStereoMatchesB200(ORBextractor* pLeft, ORBextractor* pRight, const std::vector<cv::KeyPoint>& vKeysLeft,
                              const cv::Mat& descLeft, const std::vector<cv::KeyPoint>& vKeysRight,
                              const cv::Mat& descRight, float mbf, float maxD, std::vector<float>& vuRight,
                              std::vector<float>& vDepth);

// The same on the DEVICE-RESIDENT results of the two extractors' last operator() calls - the situation of the Frame constructor,
// where ComputeStereoMatches follows ExtractORB directly (src/Frame.cc:194-217): nothing is uploaded again, only mvuRight / mvDepth
// (nLeft entries each) come back.
void ComputeStereoMatchesB200(ORBextractor* pLeft, ORBextractor* pRight, int nLeft, float mbf, float maxD, std::vector<float>& vuRight,
                              std::vector<float>& vDepth);

// ORBmatcher::DescriptorDistance (src/ORBmatcher.cc:1880-1894)
int DescriptorDistanceB200(const cv::Mat& a, const cv::Mat& b);

}  // namespace ORB_SLAM3

#endif  // ORBEXTRACTOR_H
