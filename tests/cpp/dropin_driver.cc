// Test driver for the C++ drop-in ORBextractor (morb_slam_b200/cpp). Built against the oracle's
// minimal OpenCV type shim because the image has no OpenCV C++ headers; a deployment builds the same
// two files against real OpenCV. Usage:
//   dropin_driver <w> <h> <nfeatures> <lap0> <lap1> <left.raw> <right.raw|-> <out.bin> [mbf maxD]
// Output: int32 mono, int32 n, n x 28-byte keypoints, n x 32 descriptors, 8 x (w,h) level sizes +
// level bytes, then (if a right image is given) nL floats uRight, nL floats depth.
#include <cstdio>
#include <cstdlib>
#include <string>
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
    std::vector<float> uR, depth;
    ORB_SLAM3::ComputeStereoMatchesB200(&exL, &exR, kL, dL, kR, dR, (float)atof(argv[9]), (float)atof(argv[10]), uR, depth);
    fwrite(uR.data(), 4, uR.size(), f);
    fwrite(depth.data(), 4, depth.size(), f);
  }
  fclose(f);
  std::vector<float> sf = exL.GetScaleFactors();
  printf("mono=%d n=%d levels=%d scale1=%.9g dist=%d\n", mono, n, exL.GetLevels(), sf[1],
         n > 1 ? ORB_SLAM3::DescriptorDistanceB200(dL.row(0), dL.row(1)) : -1);
  return 0;
}
