// Test driver for the C++ drop-in ORBextractor (morb_slam_b200/cpp). Built against the oracle's
// minimal OpenCV type shim because the image has no OpenCV C++ headers; a deployment builds the same
// two files against real OpenCV. Usage:
//   dropin_driver <w> <h> <nfeatures> <lap0> <lap1> <left.raw> <right.raw|-> <out.bin> [mbf maxD]
// Output: int32 mono, int32 n, n x 28-byte keypoints, n x 32 descriptors, 8 x (w,h) level sizes +
// level bytes, then (if a right image is given) nL floats uRight, nL floats depth.
#include <algorithm>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <thread>
#include <vector>

#include "ORBextractor.h"

static std::vector<unsigned char> slurp(const char* path, size_t n) {
  std::vector<unsigned char> b(n);
  FILE* f = fopen(path, "rb");
  if (!f || fread(b.data(), 1, n, f) != n) { fprintf(stderr, "cannot read %s\n", path); exit(2); }
  fclose(f);
  return b;
}

int main(int argc, char** argv) {
  if (argc == 3 && std::string(argv[1]) == "--tables") {
    // the getters right after construction, before any extraction and without a device (the first Frame constructor reads them
    // like this, src/Frame.cc:181-187): one line per table, %.9g round-trips a float
    ORB_SLAM3::ORBextractor ex(atoi(argv[2]), 1.2f, 8, 20, 7);
    std::vector<float> t[4] = {ex.GetScaleFactors(), ex.GetInverseScaleFactors(), ex.GetScaleSigmaSquares(), ex.GetInverseScaleSigmaSquares()};
    printf("%d %.9g\n", ex.GetLevels(), ex.GetScaleFactor());
    for (int k = 0; k < 4; ++k) {
      for (size_t i = 0; i < t[k].size(); ++i) printf("%.9g ", t[k][i]);
      printf("\n");
    }
    return 0;
  }
  if (argc >= 8 && std::string(argv[1]) == "--latency") {
    // dropin_driver --latency <w> <h> <nfeatures> <left.raw> <right.raw> <reps> [mbf maxD]: one stereo pair the way the Frame
    // constructor runs it (src/Frame.cc:194-217) through the drop-in class: operator() left, operator() right, ComputeStereoMatches;
    //   seq      the three calls one after the other on the caller's thread
    //   threads  the two extractions on two std::threads like the reference, then the matcher
    //   pair     ORBextractor::ExtractPair: both extractions enqueued from the caller's thread before the first synchronisation
    // pageable cv::Mat in, std::vector<cv::KeyPoint> / cv::Mat out; the pyramid download is off (the device matcher needs none).
    // Prints median / p90 milliseconds per pair.
    const int w = atoi(argv[2]), h = atoi(argv[3]), nf = atoi(argv[4]), reps = atoi(argv[7]);
    const float mbf = argc > 8 ? (float)atof(argv[8]) : 47.9f, maxD = argc > 9 ? (float)atof(argv[9]) : 435.2f;
    std::vector<unsigned char> bl = slurp(argv[5], (size_t)w * h), br = slurp(argv[6], (size_t)w * h);
    cv::Mat imL(h, w, CV_8UC1, bl.data()), imR(h, w, CV_8UC1, br.data());
    ORB_SLAM3::ORBextractor exL(nf, 1.2f, 8, 20, 7), exR(nf, 1.2f, 8, 20, 7);
    exL.SetDownloadPyramid(false); exR.SetDownloadPyramid(false);
    std::vector<cv::KeyPoint> kL, kR;
    cv::Mat dL, dR;
    std::vector<int> lap = {0, 0};
    std::vector<float> uR, depth;
    for (int mode = 0; mode < 3; ++mode) {
      std::vector<double> ts;
      for (int r = 0; r < reps + 20; ++r) {
        const auto t0 = std::chrono::steady_clock::now();
        if (mode == 0) {
          exL(imL, cv::Mat(), kL, dL, lap);
          exR(imR, cv::Mat(), kR, dR, lap);
        } else if (mode == 1) {
          std::thread tl([&] { exL(imL, cv::Mat(), kL, dL, lap); });
          std::thread tr([&] { exR(imR, cv::Mat(), kR, dR, lap); });
          tl.join(); tr.join();
        } else {
          ORB_SLAM3::ORBextractor::ExtractPair(&exL, &exR, imL, imR, kL, dL, kR, dR, lap, lap, nullptr);
        }
        ORB_SLAM3::ComputeStereoMatchesB200(&exL, &exR, (int)kL.size(), mbf, maxD, uR, depth);
        const auto t1 = std::chrono::steady_clock::now();
        if (r >= 20) ts.push_back(std::chrono::duration<double, std::milli>(t1 - t0).count());
      }
      std::sort(ts.begin(), ts.end());
      int nm = 0;
      for (float u : uR) nm += u >= 0;
      printf("%s median %.3f ms p90 %.3f ms (K = %zu / %zu, %d stereo matches)\n", mode == 0 ? "seq    " : (mode == 1 ? "threads" : "pair   "), ts[ts.size() / 2],
             ts[ts.size() * 9 / 10], kL.size(), kR.size(), nm);
    }
    return 0;
  }
  if (argc < 9) return 1;
  const int w = atoi(argv[1]), h = atoi(argv[2]), nf = atoi(argv[3]), lap0 = atoi(argv[4]), lap1 = atoi(argv[5]);
  std::vector<unsigned char> bl = slurp(argv[6], (size_t)w * h);
  cv::Mat imL(h, w, CV_8UC1, bl.data());
  ORB_SLAM3::ORBextractor exL(nf, 1.2f, 8, 20, 7), exR(nf, 1.2f, 8, 20, 7);
  std::vector<cv::KeyPoint> kL, kR;
  cv::Mat dL, dR;
  std::vector<int> lap = {lap0, lap1};
  if ((int)exL.GetScaleFactors().size() != exL.GetLevels() || (int)exL.GetInverseScaleSigmaSquares().size() != exL.GetLevels()) {
    fprintf(stderr, "scale tables are empty before the first extraction\n");
    return 4;
  }
  // empty image must return -1 (reference behaviour)
  cv::Mat empty;
  if (exL(empty, cv::Mat(), kL, dL, lap) != -1) { fprintf(stderr, "empty image did not return -1\n"); return 3; }
  const int mono = exL(imL, cv::Mat(), kL, dL, lap);
  FILE* f = fopen(argv[8], "wb");
  int n = (int)kL.size();
  fwrite(&mono, 4, 1, f); fwrite(&n, 4, 1, f);
  fwrite(kL.data(), sizeof(cv::KeyPoint), n, f);
  for (int i = 0; i < n; ++i) fwrite(dL.ptr(i), 1, 32, f);
  for (int l = 0; l < exL.GetLevels(); ++l) {
    const cv::Mat& m = exL.mvImagePyramid[l];
    fwrite(&m.cols, 4, 1, f); fwrite(&m.rows, 4, 1, f);
    for (int y = 0; y < m.rows; ++y) fwrite(m.ptr(y), 1, m.cols, f);
  }
  if (argv[7][0] != '-' && argc >= 11) {
    std::vector<unsigned char> br = slurp(argv[7], (size_t)w * h);
    cv::Mat imR(h, w, CV_8UC1, br.data());
    exR.SetDownloadPyramid(false);
    exR(imR, cv::Mat(), kR, dR, lap);
    std::vector<float> uR, depth, uR2, depth2;
    // the overload on the device-resident results of the two calls above must give what the explicit one gives
    exL.SetDownloadPyramid(false);
    exL(imL, cv::Mat(), kL, dL, lap);
    ORB_SLAM3::ComputeStereoMatchesB200(&exL, &exR, (int)kL.size(), (float)atof(argv[9]), (float)atof(argv[10]), uR2, depth2);
    ORB_SLAM3::ComputeStereoMatchesB200(&exL, &exR, kL, dL, kR, dR, (float)atof(argv[9]), (float)atof(argv[10]), uR, depth);
    if (uR.size() != uR2.size() || memcmp(uR.data(), uR2.data(), uR.size() * 4) || memcmp(depth.data(), depth2.data(), depth.size() * 4)) {
      fprintf(stderr, "resident and explicit ComputeStereoMatchesB200 differ\n");
      return 5;
    }
    // ExtractPair must give what the two operator() calls gave
    std::vector<cv::KeyPoint> pL, pR;
    cv::Mat pdL, pdR;
    int monoR = 0;
    const int monoL = ORB_SLAM3::ORBextractor::ExtractPair(&exL, &exR, imL, imR, pL, pdL, pR, pdR, lap, lap, &monoR);
    bool same = monoL == mono && pL.size() == kL.size() && pR.size() == kR.size() &&
                !memcmp(pL.data(), kL.data(), kL.size() * sizeof(cv::KeyPoint)) && !memcmp(pR.data(), kR.data(), kR.size() * sizeof(cv::KeyPoint));
    for (size_t i = 0; same && i < kL.size(); ++i) same = !memcmp(pdL.ptr((int)i), dL.ptr((int)i), 32);
    for (size_t i = 0; same && i < kR.size(); ++i) same = !memcmp(pdR.ptr((int)i), dR.ptr((int)i), 32);
    if (!same) { fprintf(stderr, "ExtractPair differs from the two operator() calls\n"); return 6; }
    fwrite(uR.data(), 4, uR.size(), f);
    fwrite(depth.data(), 4, depth.size(), f);
  }
  fclose(f);
  std::vector<float> sf = exL.GetScaleFactors();
  printf("mono=%d n=%d levels=%d scale1=%.9g dist=%d\n", mono, n, exL.GetLevels(), sf[1],
         n > 1 ? ORB_SLAM3::DescriptorDistanceB200(dL.row(0), dL.row(1)) : -1);
  return 0;
}
