// TEST INFRASTRUCTURE ONLY (oracle/_ref). Not part of the product path.
//
// ORBmatcher::SearchForTriangulation between TWO-CAMERA keyframes (mpCamera2 != NULL: the fisheye rig), UNMODIFIED reference lines cut
// out by range at build time (oracle/Makefile) into oracle/_ref/*.inc and compiled inside the stub classes below:
//   * src/ORBmatcher.cc:821-1042                     ORBmatcher::SearchForTriangulation
//   * src/ORBmatcher.cc:35-37, 1844-1876, 1880-1894  thresholds, ComputeThreeMaxima, DescriptorDistance
//   * src/CameraModels/KannalaBrandt8.cpp:229-236    KannalaBrandt8::epipolarConstrain
//   * src/CameraModels/KannalaBrandt8.cpp:68-94, 111-147, 323-395, 415-428   project / unproject / TriangulateMatches / Triangulate
// Eigen and Sophus are not in this image: Eigen is oracle/shim_eigen/mini_eigen.h (see ref_driver_kb8.cc for what that leaves
// unpinned), and the pose products of :846-855 (Tll = T1w * Tw2, Tlr = T1w * Twr2, Trl = Tr1w * Tw2, Trr = Tr1w * Twr2) are host glue
// in the library's interface (the caller passes their rotation / translation), so the Sophus stand-in only tags the four poses the
// keyframes return and looks the four products up in a table the test fills. The epipole of :833-835 is computed but never read on
// this path (`!pKF1->mpCamera2` is false at :943). Nothing of the reference is copied into the repository.
#include <math.h>
#include <climits>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <cassert>
#include <vector>
#include <algorithm>

#include <opencv2/core/core.hpp>   // the oracle's shim
#include "mini_eigen.h"
#include "DBoW2/FeatureVector.h"   // the reference's Thirdparty/DBoW2 (Boost declarations: oracle/shim_dbow)

using namespace std;

namespace cv {
struct Point3f {
  float x, y, z;
  Point3f() : x(0), y(0), z(0) {}
  Point3f(float _x, float _y, float _z) : x(_x), y(_y), z(_z) {}
};
}  // namespace cv

namespace Sophus {
struct SE3f {
  Eigen::Matrix3f R;
  Eigen::Vector3f t;
  int tag = 0;   // 1: T1w, 2: Tr1w, 3: Tw2, 4: Twr2, 0: anything else
  SE3f() { R = Eigen::Matrix3f::Identity(); }
  Eigen::Matrix3f rotationMatrix() const { return R; }
  Eigen::Vector3f translation() const { return t; }
  Eigen::Vector3f operator*(const Eigen::Vector3f& v) const { return R * v + t; }
  SE3f operator*(const SE3f& o) const;
};
static SE3f g_prod[2][2];   // [T1w | Tr1w] x [Tw2 | Twr2] = Tll Tlr / Trl Trr
inline SE3f SE3f::operator*(const SE3f& o) const {
  if ((tag == 1 || tag == 2) && (o.tag == 3 || o.tag == 4)) return g_prod[tag - 1][o.tag - 3];
  return SE3f();
}
}  // namespace Sophus

namespace ORB_SLAM3 {
class GeometricCamera {
 public:
  virtual ~GeometricCamera() {}
  virtual Eigen::Vector2f project(const Eigen::Vector3f& v3D) = 0;
  virtual Eigen::Vector3f unprojectEig(const cv::Point2f& p2D) = 0;
  virtual bool epipolarConstrain(GeometricCamera* pCamera2, const cv::KeyPoint& kp1, const cv::KeyPoint& kp2, const Eigen::Matrix3f& R12,
                                 const Eigen::Vector3f& t12, const float sigmaLevel, const float unc) = 0;
};

class KannalaBrandt8 : public GeometricCamera {
 public:
  KannalaBrandt8(const float* p, float prec) : mvParameters(p, p + 8), precision(prec) {}
  Eigen::Vector2f project(const Eigen::Vector3f& v3D);
  Eigen::Vector3f unprojectEig(const cv::Point2f& p2D);
  cv::Point3f unproject(const cv::Point2f& p2D);
  bool epipolarConstrain(GeometricCamera* pCamera2, const cv::KeyPoint& kp1, const cv::KeyPoint& kp2, const Eigen::Matrix3f& R12,
                         const Eigen::Vector3f& t12, const float sigmaLevel, const float unc);
  float TriangulateMatches(GeometricCamera* pCamera2, const cv::KeyPoint& kp1, const cv::KeyPoint& kp2, const Eigen::Matrix3f& R12,
                           const Eigen::Vector3f& t12, const float sigmaLevel, const float unc, Eigen::Vector3f& p3D);
  void Triangulate(const cv::Point2f& p1, const cv::Point2f& p2, const Eigen::Matrix<float, 3, 4>& Tcw1,
                   const Eigen::Matrix<float, 3, 4>& Tcw2, Eigen::Vector3f& x3D);
  std::vector<float> mvParameters;
  const float precision;
};

#include "kb8_project.inc"
#include "kb8_unproject_eig.inc"
#include "kb8_unproject.inc"
#include "kb8_epipolar.inc"              // src/CameraModels/KannalaBrandt8.cpp:229-236
#include "kb8_triangulate_matches.inc"
#include "kb8_triangulate.inc"

struct MapPoint {};

struct KeyFrame {
  int N = 0, NLeft = -1;
  std::vector<cv::KeyPoint> mvKeys, mvKeysUn, mvKeysRight;
  std::vector<float> mvuRight, mvScaleFactors, mvLevelSigma2;
  cv::Mat mDescriptors;
  std::vector<MapPoint*> mvpMapPoints;
  DBoW2::FeatureVector mFeatVec;
  GeometricCamera* mpCamera = nullptr;
  GeometricCamera* mpCamera2 = nullptr;
  int first = 1;   // pKF1 hands out the tags 1 / 2, pKF2 the tags 3 / 4
  Sophus::SE3f tagged(int t) { Sophus::SE3f T; T.tag = t; return T; }
  Sophus::SE3f GetPose() { return tagged(first ? 1 : 0); }
  Sophus::SE3f GetPoseInverse() { return tagged(first ? 0 : 3); }
  Sophus::SE3f GetRightPose() { return tagged(first ? 2 : 0); }
  Sophus::SE3f GetRightPoseInverse() { return tagged(first ? 0 : 4); }
  Eigen::Vector3f GetCameraCenter() { return Eigen::Vector3f(0.f, 0.f, 1.f); }
  MapPoint* GetMapPoint(const size_t& idx) { return mvpMapPoints[idx]; }
};

struct ORBmatcher {
  static const int TH_LOW;
  static const int TH_HIGH;
  static const int HISTO_LENGTH;
  float mfNNratio;
  bool mbCheckOrientation;
  ORBmatcher(float nnratio, bool checkOri) : mfNNratio(nnratio), mbCheckOrientation(checkOri) {}
  static int DescriptorDistance(const cv::Mat& a, const cv::Mat& b);
  int SearchForTriangulation(KeyFrame* pKF1, KeyFrame* pKF2, std::vector<pair<size_t, size_t> >& vMatchedPairs, const bool bOnlyStereo,
                             const bool bCoarse = false);
  void ComputeThreeMaxima(std::vector<int>* histo, const int L, int& ind1, int& ind2, int& ind3);
};
#include "orbmatcher_consts3.inc"   // src/ORBmatcher.cc:35-37
#include "orbmatcher_sft.inc"       // src/ORBmatcher.cc:821-1042
#include "orbmatcher_max3.inc"      // src/ORBmatcher.cc:1844-1876
#include "orbmatcher_dist.inc"      // src/ORBmatcher.cc:1880-1894
}  // namespace ORB_SLAM3

using namespace ORB_SLAM3;

struct RigC {   // same layout as orb_kb8_rig (include/orb_b200.h)
  float cam1[8], cam2[8], prec1, prec2, R12[9], t12[3];
};

static void fill_fv(DBoW2::FeatureVector& fv, const uint32_t* node, const int* off, const uint32_t* feat, int nn) {
  for (int j = 0; j < nn; ++j)
    for (int t = off[j]; t < off[j + 1]; ++t) fv.addFeature(node[j], feat[t]);
}

static void fill_kf(KeyFrame& k, GeometricCamera* cl, GeometricCamera* cr, const cv::KeyPoint* kps, const uint8_t* desc, const uint8_t* has_mp,
                    int n, int nleft, const uint32_t* node, const int* off, const uint32_t* feat, int nn, const float* scale, const float* sigma2,
                    int nlevels, MapPoint* some) {
  k.N = n; k.NLeft = nleft;
  k.mvKeys.assign(kps, kps + nleft);
  k.mvKeysRight.assign(kps + nleft, kps + n);
  k.mvKeysUn = k.mvKeys;
  k.mvuRight.assign(n, -1.f);
  k.mDescriptors = cv::Mat(std::max(n, 1), 32, CV_8UC1);
  if (n) std::memcpy(k.mDescriptors.data, desc, (size_t)n * 32);
  k.mvpMapPoints.assign(n, (MapPoint*)nullptr);
  for (int i = 0; i < n; ++i) if (has_mp[i]) k.mvpMapPoints[i] = some;
  fill_fv(k.mFeatVec, node, off, feat, nn);
  k.mpCamera = cl; k.mpCamera2 = cr;
  k.mvScaleFactors.assign(scale, scale + nlevels);
  k.mvLevelSigma2.assign(sigma2, sigma2 + nlevels);
}

extern "C" {
// kps1 / kps2: the left keypoints followed by the right ones (n entries, the first nleft are left). rigs[4] = ll, lr, rl, rr:
// cam1 / cam2 of rigs[0] are the two LEFT cameras, of rigs[3] the two RIGHT cameras (pKF1->mpCamera2, pKF2->mpCamera2).
// match12[n1] = vMatches12 as the pair list reports it. Returns nmatches.
int refsft2_search(const void* kps1, const uint8_t* desc1, const uint8_t* has_mp1, int n1, int nleft1, const uint32_t* node1, const int* off1,
                   const uint32_t* feat1, int nn1, const void* kps2, const uint8_t* desc2, const uint8_t* has_mp2, int n2, int nleft2,
                   const uint32_t* node2, const int* off2, const uint32_t* feat2, int nn2, const float* scale, const float* sigma2, int nlevels,
                   const RigC* rigs, int only_stereo, int coarse, int check_orientation, int* match12) {
  KannalaBrandt8 l1(rigs[0].cam1, rigs[0].prec1), r1(rigs[3].cam1, rigs[3].prec1), l2(rigs[0].cam2, rigs[0].prec2), r2(rigs[3].cam2, rigs[3].prec2);
  for (int a = 0; a < 2; ++a)
    for (int b = 0; b < 2; ++b) {
      Sophus::SE3f& T = Sophus::g_prod[a][b];
      const RigC& r = rigs[2 * a + b];
      for (int i = 0; i < 9; ++i) T.R.d[i] = r.R12[i];
      for (int i = 0; i < 3; ++i) T.t.d[i] = r.t12[i];
      T.tag = 0;
    }
  MapPoint some;
  KeyFrame k1, k2;
  fill_kf(k1, &l1, &r1, (const cv::KeyPoint*)kps1, desc1, has_mp1, n1, nleft1, node1, off1, feat1, nn1, scale, sigma2, nlevels, &some);
  fill_kf(k2, &l2, &r2, (const cv::KeyPoint*)kps2, desc2, has_mp2, n2, nleft2, node2, off2, feat2, nn2, scale, sigma2, nlevels, &some);
  k1.first = 1; k2.first = 0;
  ORBmatcher m(0.6f, check_orientation != 0);
  std::vector<pair<size_t, size_t> > pairs;
  const int nm = m.SearchForTriangulation(&k1, &k2, pairs, only_stereo != 0, coarse != 0);
  for (int i = 0; i < n1; ++i) match12[i] = -1;
  for (size_t i = 0; i < pairs.size(); ++i) match12[pairs[i].first] = (int)pairs[i].second;
  return nm;
}
}
