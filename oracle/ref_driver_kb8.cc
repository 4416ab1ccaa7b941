// TEST INFRASTRUCTURE ONLY (oracle/_ref). Not part of the product path.
//
// C entry points around UNMODIFIED reference sources of the fisheye stereo triangulation, cut out by line
// range at build time (oracle/Makefile) into oracle/_ref/*.inc and compiled inside the stub class below:
//   * src/CameraModels/KannalaBrandt8.cpp:68-94    KannalaBrandt8::project(const Eigen::Vector3f&)
//   * src/CameraModels/KannalaBrandt8.cpp:111-114  unprojectEig
//   * src/CameraModels/KannalaBrandt8.cpp:116-147  unproject
//   * src/CameraModels/KannalaBrandt8.cpp:323-395  TriangulateMatches
//   * src/CameraModels/KannalaBrandt8.cpp:415-428  Triangulate
//   * src/Frame.cc:1244-1273                       the acceptance loop of Frame::ComputeStereoFishEyeMatches
// Eigen is not in this image: oracle/shim_eigen/mini_eigen.h supplies the expressions these functions use, with a one-sided Jacobi
// SVD in double behind Eigen::JacobiSVD (see that header for what this leaves unpinned). <math.h> is included next to <cmath>, so
// the unqualified cos(psi) / sin(psi) of project() resolve to the float overloads like in a translation unit that sees OpenCV's
// headers. Nothing of the reference is copied into the repository.
#include <math.h>
#include <cmath>
#include <cstdint>
#include <vector>

#include <opencv2/core/core.hpp>   // the oracle's shim
#include "mini_eigen.h"

namespace cv {
struct Point3f {
  float x, y, z;
  Point3f() : x(0), y(0), z(0) {}
  Point3f(float _x, float _y, float _z) : x(_x), y(_y), z(_z) {}
};
}  // namespace cv

namespace ORB_SLAM3 {
class GeometricCamera {
 public:
  virtual ~GeometricCamera() {}
  virtual Eigen::Vector2f project(const Eigen::Vector3f& v3D) = 0;
  virtual Eigen::Vector3f unprojectEig(const cv::Point2f& p2D) = 0;
};

class KannalaBrandt8 : public GeometricCamera {
 public:
  KannalaBrandt8(const float* p, float prec) : mvParameters(p, p + 8), precision(prec) {}
  Eigen::Vector2f project(const Eigen::Vector3f& v3D);
  Eigen::Vector3f unprojectEig(const cv::Point2f& p2D);
  cv::Point3f unproject(const cv::Point2f& p2D);
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
#include "kb8_triangulate_matches.inc"
#include "kb8_triangulate.inc"
}  // namespace ORB_SLAM3

// the members Frame::ComputeStereoFishEyeMatches touches from :1246 on
struct FrameStub {
  std::vector<cv::KeyPoint> mvKeys, mvKeysRight;
  std::vector<float> mvLevelSigma2;
  int monoLeft, monoRight;
  ORB_SLAM3::GeometricCamera *mpCamera, *mpCamera2;
  Eigen::Matrix3f mRlr;
  Eigen::Vector3f mtlr;
  std::vector<int> mvLeftToRightMatch, mvRightToLeftMatch;
  std::vector<float> mvDepth;
  std::vector<Eigen::Vector3f> mvStereo3Dpoints;
  void accept(std::vector<std::vector<cv::DMatch>>& matches) {
    using namespace ORB_SLAM3;
    using std::vector;
#include "frame_fisheye_accept.inc"
  }
};

extern "C" {
// ret[i] = TriangulateMatches(kp1[i], kp2[i]); p3d[3 i ..] the point (0 when rejected)
void ref_kb8_triangulate(const float* cam1, float prec1, const float* cam2, float prec2, const float* R12, const float* t12, const float* xy1,
                         const float* xy2, const float* s1, const float* s2, int n, float* ret, float* p3d) {
  ORB_SLAM3::KannalaBrandt8 c1(cam1, prec1), c2(cam2, prec2);
  Eigen::Matrix3f R;
  Eigen::Vector3f t;
  for (int i = 0; i < 9; ++i) R.d[i] = R12[i];
  for (int i = 0; i < 3; ++i) t.d[i] = t12[i];
  for (int i = 0; i < n; ++i) {
    cv::KeyPoint a(xy1[2 * i], xy1[2 * i + 1], 31.f), b(xy2[2 * i], xy2[2 * i + 1], 31.f);
    Eigen::Vector3f X;
    ret[i] = c1.TriangulateMatches(&c2, a, b, R, t, s1[i], s2[i], X);
    p3d[3 * i] = X[0]; p3d[3 * i + 1] = X[1]; p3d[3 * i + 2] = X[2];
  }
}
void ref_kb8_unproject(const float* cam, float prec, const float* xy, int n, float* rays) {
  ORB_SLAM3::KannalaBrandt8 c(cam, prec);
  for (int i = 0; i < n; ++i) {
    const cv::Point3f r = c.unproject(cv::Point2f(xy[2 * i], xy[2 * i + 1]));
    rays[3 * i] = r.x; rays[3 * i + 1] = r.y; rays[3 * i + 2] = r.z;
  }
}
void ref_kb8_project(const float* cam, const float* xyz, int n, float* uv) {
  ORB_SLAM3::KannalaBrandt8 c(cam, 1e-6f);
  for (int i = 0; i < n; ++i) {
    const Eigen::Vector2f r = c.project(Eigen::Vector3f(xyz[3 * i], xyz[3 * i + 1], xyz[3 * i + 2]));
    uv[2 * i] = r[0]; uv[2 * i + 1] = r[1];
  }
}
// the loop of Frame::ComputeStereoFishEyeMatches after knnMatch: knn_idx / knn_dist hold nq x 2 (trainIdx, distance; -1 = absent)
void ref_fisheye_accept(const float* cam1, float prec1, const float* cam2, float prec2, const float* R12, const float* t12,
                        const cv::KeyPoint* kL, int nL, int monoL, const cv::KeyPoint* kR, int nR, int monoR, const float* sigma2, int nlev,
                        const int* knn_idx, const int* knn_dist, int nq, int* l2r, int* r2l, float* depth, float* p3d) {
  ORB_SLAM3::KannalaBrandt8 c1(cam1, prec1), c2(cam2, prec2);
  FrameStub F;
  F.mvKeys.assign(kL, kL + nL);
  F.mvKeysRight.assign(kR, kR + nR);
  F.mvLevelSigma2.assign(sigma2, sigma2 + nlev);
  F.monoLeft = monoL; F.monoRight = monoR;
  F.mpCamera = &c1; F.mpCamera2 = &c2;
  for (int i = 0; i < 9; ++i) F.mRlr.d[i] = R12[i];
  for (int i = 0; i < 3; ++i) F.mtlr.d[i] = t12[i];
  F.mvLeftToRightMatch.assign(nL, -1);
  F.mvRightToLeftMatch.assign(nR, -1);
  F.mvDepth.assign(nL, -1.0f);
  F.mvStereo3Dpoints.assign(nL, Eigen::Vector3f());
  std::vector<std::vector<cv::DMatch>> matches(nq);
  for (int i = 0; i < nq; ++i)
    for (int k = 0; k < 2; ++k)
      if (knn_idx[2 * i + k] >= 0) matches[i].push_back(cv::DMatch(i, knn_idx[2 * i + k], (float)knn_dist[2 * i + k]));
  F.accept(matches);
  for (int i = 0; i < nL; ++i) {
    l2r[i] = F.mvLeftToRightMatch[i];
    depth[i] = F.mvDepth[i];
    for (int k = 0; k < 3; ++k) p3d[3 * i + k] = F.mvStereo3Dpoints[i][k];
  }
  for (int i = 0; i < nR; ++i) r2l[i] = F.mvRightToLeftMatch[i];
}
}
