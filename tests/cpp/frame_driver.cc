// Test driver for the reference-typed helpers of morb_slam_b200/cpp/FrameB200.h (UndistortKeyPoints, AssignFeaturesToGrid,
// ComputeBoW with std::map types), built against the oracle's OpenCV type shim like dropin_driver. Usage:
//   frame_driver <w> <h> <nfeatures> <image.raw> <vocabulary.txt> <levelsup> <out.bin>
// Output: int32 n, n x 28-byte mvKeysUn, int32 nb, nb x (uint32 word, double value), int32 nn, then per node uint32 node,
// int32 count, count x uint32 feature.
#include <cstdio>
#include <cstdlib>
#include <map>
#include <vector>

#include "FrameB200.h"

#ifndef CV_32F
#define CV_32F 5
#endif

int main(int argc, char** argv) {
  if (argc < 8) return 1;
  const int w = atoi(argv[1]), h = atoi(argv[2]), nf = atoi(argv[3]), levelsup = atoi(argv[6]);
  std::vector<unsigned char> img((size_t)w * h);
  FILE* fi = fopen(argv[4], "rb");
  if (!fi || fread(img.data(), 1, img.size(), fi) != img.size()) return 2;
  fclose(fi);
  cv::Mat im(h, w, CV_8UC1, img.data());
  ORB_SLAM3::ORBextractor ex(nf, 1.2f, 8, 20, 7);
  ex.SetDownloadPyramid(false);
  std::vector<cv::KeyPoint> keys, keysUn;
  cv::Mat desc;
  std::vector<int> lap = {0, 0};
  ex(im, cv::Mat(), keys, desc, lap);
  // EuRoC cam0 (Examples/Monocular/EuRoC.yaml): toK(), mDistCoef (k1 k2 p1 p2), mK
  float k[9] = {458.654f, 0, 367.215f, 0, 457.296f, 248.375f, 0, 0, 1};
  float d[4] = {-0.28340811f, 0.07395907f, 0.00019359f, 1.76187114e-05f};
  cv::Mat K(3, 3, CV_32F, k), D(4, 1, CV_32F, d);
  ORB_SLAM3::UndistortKeyPointsB200(&ex, K, D, K, keys, keysUn);
  ORB_SLAM3::AssignFeaturesToGridB200(&ex, 0.f, 0.f, (float)w, (float)h, 64.f / w, 48.f / h);
  orb_vocab* voc = nullptr;
  if (orb_vocab_load_text(0, argv[5], &voc) != ORB_OK) return 3;
  std::map<unsigned int, double> bow;
  std::map<unsigned int, std::vector<unsigned int> > fv;
  ORB_SLAM3::ComputeBoWB200(&ex, voc, bow, fv, levelsup);
  orb_vocab_destroy(voc);
  // Frame::ComputeStereoFishEyeMatches through the reference-typed helper: a second extractor on the same image with the whole
  // width as lapping area; identical images and a sideways rig: every pair is parallel rays (parallax gate), so no match survives
  {
    ORB_SLAM3::ORBextractor exL(nf, 1.2f, 8, 20, 7), exR(nf, 1.2f, 8, 20, 7);
    exL.SetDownloadPyramid(false); exR.SetDownloadPyramid(false);
    std::vector<cv::KeyPoint> kl, kr;
    cv::Mat dl, dr;
    std::vector<int> lapAll = {0, w - 1};
    const int monoL = exL(im, cv::Mat(), kl, dl, lapAll), monoR = exR(im, cv::Mat(), kr, dr, lapAll);
    orb_kb8_rig rig = {{190.f, 190.f, w * 0.5f, h * 0.5f, 0, 0, 0, 0}, {190.f, 190.f, w * 0.5f, h * 0.5f, 0, 0, 0, 0}, 1e-6f, 1e-6f,
                       {1, 0, 0, 0, 1, 0, 0, 0, 1}, {0.1f, 0.f, 0.f}};
    std::vector<int> l2r, r2l;
    std::vector<float> depth;
    struct V3 { float v[3]; float& operator[](int i) { return v[i]; } };
    std::vector<V3> p3d;
    const int nm = ORB_SLAM3::ComputeStereoFishEyeMatchesB200(&exL, &exR, rig, (int)kl.size(), (int)kr.size(), l2r, r2l, depth, p3d);
    printf("fisheye: mono %d/%d, N %d/%d, matches %d\n", monoL, monoR, (int)kl.size(), (int)kr.size(), nm);
    if (nm != 0 || (int)l2r.size() != (int)kl.size()) return 4;
  }
  // ORBmatcher::SearchForInitialization through the reference-typed helper: the frame against itself as the initial frame, the search
  // windows moved off the keypoints by (3, -2) px; the result goes behind the other records of the output file
  std::vector<cv::Point2f> prevMatched(keysUn.size());
  for (size_t i = 0; i < keysUn.size(); ++i) { prevMatched[i].x = keysUn[i].pt.x + 3.f; prevMatched[i].y = keysUn[i].pt.y - 2.f; }
  std::vector<int> iniMatches;
  const int nIni = ORB_SLAM3::SearchForInitializationB200(&ex, keysUn, desc, prevMatched, iniMatches, 30, 0.9f, true);
  FILE* f = fopen(argv[7], "wb");
  int n = (int)keysUn.size();
  fwrite(&n, 4, 1, f);
  fwrite(keysUn.data(), sizeof(cv::KeyPoint), n, f);
  int nb = (int)bow.size();
  fwrite(&nb, 4, 1, f);
  for (std::map<unsigned int, double>::const_iterator it = bow.begin(); it != bow.end(); ++it) { fwrite(&it->first, 4, 1, f); fwrite(&it->second, 8, 1, f); }
  int nn = (int)fv.size();
  fwrite(&nn, 4, 1, f);
  for (std::map<unsigned int, std::vector<unsigned int> >::const_iterator it = fv.begin(); it != fv.end(); ++it) {
    int c = (int)it->second.size();
    fwrite(&it->first, 4, 1, f); fwrite(&c, 4, 1, f); fwrite(it->second.data(), 4, c, f);
  }
  fwrite(&nIni, 4, 1, f);
  fwrite(iniMatches.data(), 4, iniMatches.size(), f);
  fwrite(prevMatched.data(), 8, prevMatched.size(), f);
  fclose(f);
  printf("n=%d words=%d nodes=%d ini=%d\n", n, nb, nn, nIni);
  return 0;
}
