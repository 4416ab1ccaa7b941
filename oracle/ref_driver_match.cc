// TEST INFRASTRUCTURE ONLY (oracle/_ref). Not part of the product path.
//
// C entry points around UNMODIFIED reference sources of the windowed matcher, cut out by
// line range at build time (oracle/Makefile) into oracle/_ref/*.inc and compiled inside the stub classes below:
//   * src/Frame.cc:501-528   Frame::AssignFeaturesToGrid
//   * src/Frame.cc:742-807   Frame::GetFeaturesInArea
//   * src/Frame.cc:809-820   Frame::PosInGrid
//   * src/ORBmatcher.cc:35-37        thresholds
//   * src/ORBmatcher.cc:1521-1733    ORBmatcher::SearchByProjection(Frame&, const Frame&, th, bMono)
//   * src/ORBmatcher.cc:42-216       ORBmatcher::SearchByProjection(Frame&, const vector<MapPoint*>&, th, bFarPoints, thFarPoints)
//                                    and RadiusByViewingCos
//   * src/ORBmatcher.cc:218-395      ORBmatcher::SearchByBoW(KeyFrame*, Frame&, vector<MapPoint*>&), with the reference's own
//                                    Thirdparty/DBoW2/DBoW2/FeatureVector.{h,cpp} (unmodified)
//   * src/ORBmatcher.cc:603-700      ORBmatcher::SearchForInitialization
//   * src/ORBmatcher.cc:1844-1876    ORBmatcher::ComputeThreeMaxima
//   * src/ORBmatcher.cc:1880-1894    ORBmatcher::DescriptorDistance
// Eigen and Sophus are not in this image. The stubs give the pose arithmetic pure-translation semantics
// (SE3f * v = v + t, exact for t = 0) and the camera an identity projection (u, v) = (x, y), so the test drives the
// code AFTER the projection step with exact inputs: a map point at (u, v, z) projects to (u, v) with
// invzc = (float)(1.0 / z). Everything after that line - bounds, window, level gates, uRight gate, greedy
// assignment with the Observations() lock, rotation histogram - is the reference's own code.
// Nothing of the reference is copied into the repository.
#include <cstdint>
#include <cstring>
#include <cassert>
#include <climits>
#include <cmath>
#include <vector>
#include <algorithm>

#include <opencv2/core/core.hpp>   // the oracle's shim
#include "DBoW2/FeatureVector.h"   // the reference's Thirdparty/DBoW2 (Boost declarations: oracle/shim_dbow)

using namespace std;

namespace Eigen {
struct Vector3f {
  float d[3];
  Vector3f() : d{0, 0, 0} {}
  Vector3f(float x, float y, float z) : d{x, y, z} {}
  float operator()(int i) const { return d[i]; }
  float& operator()(int i) { return d[i]; }
};
struct Vector2f {
  float d[2];
  Vector2f() : d{0, 0} {}
  Vector2f(float x, float y) : d{x, y} {}
  float operator()(int i) const { return d[i]; }
};
}  // namespace Eigen

namespace Sophus {
struct SE3f {  // pure translation
  Eigen::Vector3f t;
  SE3f inverse() const { SE3f r; r.t = Eigen::Vector3f(-t.d[0], -t.d[1], -t.d[2]); return r; }
  Eigen::Vector3f translation() const { return t; }
  Eigen::Vector3f operator*(const Eigen::Vector3f& v) const { return Eigen::Vector3f(v.d[0] + t.d[0], v.d[1] + t.d[1], v.d[2] + t.d[2]); }
};
}  // namespace Sophus

#define FRAME_GRID_ROWS 48   // include/Frame.h:44-45
#define FRAME_GRID_COLS 64

namespace ORB_SLAM3 {

struct MapPoint {
  Eigen::Vector3f pos;
  cv::Mat desc;
  int nobs = 0;
  Eigen::Vector3f GetWorldPos() { return pos; }
  cv::Mat GetDescriptor() { return desc; }
  int Observations() { return nobs; }
  // tracking members written by Frame::isInFrustum (include/MapPoint.h:171-179)
  float mTrackProjX = 0, mTrackProjY = 0, mTrackDepth = 0, mTrackDepthR = 0, mTrackProjXR = 0, mTrackProjYR = 0;
  bool mbTrackInView = false, mbTrackInViewR = false;
  int mnTrackScaleLevel = 0, mnTrackScaleLevelR = -1;
  float mTrackViewCos = 0, mTrackViewCosR = 0;
  bool bad = false;
  bool isBad() { return bad; }
};

struct GeometricCamera {
  Eigen::Vector2f project(const Eigen::Vector3f& v) { return Eigen::Vector2f(v.d[0], v.d[1]); }
};

struct KeyFrame {   // what SearchByBoW reads of it (include/KeyFrame.h)
  std::vector<MapPoint*> mvpMapPoints;
  std::vector<MapPoint*> GetMapPointMatches() { return mvpMapPoints; }
  DBoW2::FeatureVector mFeatVec;
  cv::Mat mDescriptors;
  std::vector<cv::KeyPoint> mvKeys, mvKeysUn, mvKeysRight;
  GeometricCamera* mpCamera2 = nullptr;
  int NLeft = -1;
};

struct Frame {
  DBoW2::FeatureVector mFeatVec;
  GeometricCamera* mpCamera2 = nullptr;
  int N = 0, Nleft = -1;
  std::vector<cv::KeyPoint> mvKeys, mvKeysUn, mvKeysRight;
  std::vector<MapPoint*> mvpMapPoints;
  std::vector<bool> mvbOutlier;
  cv::Mat mDescriptors;
  std::vector<float> mvuRight;
  std::vector<int> mvLeftToRightMatch, mvRightToLeftMatch;
  std::vector<float> mvScaleFactors;
  float mb = 0, mbf = 0;
  static float mnMinX, mnMaxX, mnMinY, mnMaxY;
  static float mfGridElementWidthInv, mfGridElementHeightInv;
  std::vector<std::size_t> mGrid[FRAME_GRID_COLS][FRAME_GRID_ROWS];
  std::vector<std::size_t> mGridRight[FRAME_GRID_COLS][FRAME_GRID_ROWS];
  GeometricCamera* mpCamera = nullptr;
  Sophus::SE3f mTcw, mTrl;
  Sophus::SE3f GetPose() const { return mTcw; }
  Sophus::SE3f GetRelativePoseTrl() { return mTrl; }
  void AssignFeaturesToGrid();
  vector<size_t> GetFeaturesInArea(const float& x, const float& y, const float& r, const int minLevel = -1,
                                   const int maxLevel = -1, const bool bRight = false) const;  // include/Frame.h:113
  bool PosInGrid(const cv::KeyPoint& kp, int& posX, int& posY);
};
float Frame::mnMinX, Frame::mnMaxX, Frame::mnMinY, Frame::mnMaxY;
float Frame::mfGridElementWidthInv, Frame::mfGridElementHeightInv;

#include "frame_grid_assign.inc"   // src/Frame.cc:501-528
#include "frame_grid_area.inc"     // src/Frame.cc:742-807
#include "frame_grid_pos.inc"      // src/Frame.cc:809-820

struct ORBmatcher {
  static const int TH_LOW;
  static const int TH_HIGH;
  static const int HISTO_LENGTH;
  float mfNNratio;
  bool mbCheckOrientation;
  ORBmatcher(float nnratio, bool checkOri) : mfNNratio(nnratio), mbCheckOrientation(checkOri) {}
  static int DescriptorDistance(const cv::Mat& a, const cv::Mat& b);
  int SearchByProjection(Frame& CurrentFrame, const Frame& LastFrame, const float th, const bool bMono);
  int SearchByProjection(Frame& F, const std::vector<MapPoint*>& vpMapPoints, const float th = 3, const bool bFarPoints = false,
                         const float thFarPoints = 50.0f);  // include/ORBmatcher.h:49-51
  float RadiusByViewingCos(const float& viewCos);
  int SearchByBoW(KeyFrame* pKF, Frame& F, std::vector<MapPoint*>& vpMapPointMatches);
  int SearchForInitialization(Frame& F1, Frame& F2, std::vector<cv::Point2f>& vbPrevMatched, std::vector<int>& vnMatches12,
                              int windowSize = 10);  // include/ORBmatcher.h:72-74
  void ComputeThreeMaxima(std::vector<int>* histo, const int L, int& ind1, int& ind2, int& ind3);
};
#include "orbmatcher_consts3.inc"  // src/ORBmatcher.cc:35-37
#include "orbmatcher_sbp.inc"      // src/ORBmatcher.cc:1521-1733
#include "orbmatcher_sbp_map.inc"  // src/ORBmatcher.cc:42-209
#include "orbmatcher_radius.inc"   // src/ORBmatcher.cc:211-216
#include "orbmatcher_sbow.inc"     // src/ORBmatcher.cc:218-395
#include "orbmatcher_sfi.inc"      // src/ORBmatcher.cc:603-700
#include "orbmatcher_max3.inc"     // src/ORBmatcher.cc:1844-1876
#include "orbmatcher_dist.inc"     // src/ORBmatcher.cc:1880-1894

}  // namespace ORB_SLAM3

using namespace ORB_SLAM3;

struct QueryC {   // same layout as orb_proj_query (include/orb_b200.h)
  float u, v, z, angle;
  int octave, flags;  // bit 0: map point present and not an outlier, bit 1: Observations() > 0
};

static void set_frame_statics(const float* gp) {
  Frame::mnMinX = gp[0]; Frame::mnMinY = gp[1]; Frame::mnMaxX = gp[2]; Frame::mnMaxY = gp[3];
  Frame::mfGridElementWidthInv = gp[4]; Frame::mfGridElementHeightInv = gp[5];
}

extern "C" {

// Frame::AssignFeaturesToGrid: CSR of the 64 x 48 cell lists in the reference's own cell order (ix major, iy)
int refm_assign_grid(const void* kps, int n, const float* gp, int* cell_off /*[3073]*/, int* idx /*[n]*/) {
  set_frame_statics(gp);
  Frame f;
  f.N = n;
  f.mvKeysUn.assign((const cv::KeyPoint*)kps, (const cv::KeyPoint*)kps + n);
  f.mvKeys = f.mvKeysUn;
  f.AssignFeaturesToGrid();
  int o = 0;
  for (int ix = 0; ix < FRAME_GRID_COLS; ++ix)
    for (int iy = 0; iy < FRAME_GRID_ROWS; ++iy) {
      cell_off[ix * FRAME_GRID_ROWS + iy] = o;
      for (size_t j = 0; j < f.mGrid[ix][iy].size(); ++j) idx[o++] = (int)f.mGrid[ix][iy][j];
    }
  cell_off[FRAME_GRID_COLS * FRAME_GRID_ROWS] = o;
  return o;
}

// Frame::GetFeaturesInArea on a frame built from the keypoints
int refm_features_in_area(const void* kps, int n, const float* gp, float x, float y, float r, int minLevel, int maxLevel,
                          int* out, int cap) {
  set_frame_statics(gp);
  Frame f;
  f.N = n;
  f.mvKeysUn.assign((const cv::KeyPoint*)kps, (const cv::KeyPoint*)kps + n);
  f.mvKeys = f.mvKeysUn;
  f.AssignFeaturesToGrid();
  vector<size_t> v = f.GetFeaturesInArea(x, y, r, minLevel, maxLevel);
  if ((int)v.size() > cap) return -2;
  for (size_t i = 0; i < v.size(); ++i) out[i] = (int)v[i];
  return (int)v.size();
}

// ORBmatcher::SearchByProjection(CurrentFrame, LastFrame, th, bMono), rectified / monocular frames (Nleft == -1).
// Current frame: keypoints, descriptors, uRight; last frame: one query per keypoint. tlc_z drives the
// forward / backward test (tlc(2) of the reference). match_out[i2] = index of the last-frame keypoint whose map
// point CurrentFrame.mvpMapPoints[i2] holds at the end, or -1. Returns nmatches.
int refm_search_by_projection(const void* kpsC, const uint8_t* descC, const float* uRightC, int nC, const float* scale, int nlevels,
                              const float* gp, float mb, float mbf, const QueryC* q, const uint8_t* qdesc, int nq, float th, int bMono,
                              float tlc_z, int check_orientation, int* match_out) {
  set_frame_statics(gp);
  GeometricCamera cam;
  Frame cur, last;
  cur.N = nC;
  cur.mvKeysUn.assign((const cv::KeyPoint*)kpsC, (const cv::KeyPoint*)kpsC + nC);
  cur.mvKeys = cur.mvKeysUn;
  cur.mvpMapPoints.assign(nC, (MapPoint*)nullptr);
  cur.mDescriptors = cv::Mat(std::max(nC, 1), 32, CV_8UC1);
  if (nC) std::memcpy(cur.mDescriptors.data, descC, (size_t)nC * 32);
  cur.mvuRight.assign(uRightC, uRightC + nC);
  cur.mvScaleFactors.assign(scale, scale + nlevels);
  cur.mb = mb; cur.mbf = mbf;
  cur.mpCamera = &cam;
  cur.AssignFeaturesToGrid();
  last.N = nq;
  last.mvKeys.resize(nq); last.mvKeysUn.resize(nq);
  last.mvpMapPoints.assign(nq, (MapPoint*)nullptr);
  last.mvbOutlier.assign(nq, false);
  last.mTcw.t = Eigen::Vector3f(0.f, 0.f, tlc_z);   // tlc = Tlw * twc = twc + tlw = (0, 0, tlc_z) for Tcw = identity
  std::vector<MapPoint> mps(nq);
  for (int i = 0; i < nq; ++i) {
    last.mvKeys[i].octave = q[i].octave; last.mvKeys[i].angle = q[i].angle;
    last.mvKeysUn[i] = last.mvKeys[i];
    if (q[i].flags & 1) {
      mps[i].pos = Eigen::Vector3f(q[i].u, q[i].v, q[i].z);
      mps[i].desc = cv::Mat(1, 32, CV_8UC1);
      std::memcpy(mps[i].desc.data, qdesc + 32 * (size_t)i, 32);
      mps[i].nobs = (q[i].flags & 2) ? 1 : 0;
      last.mvpMapPoints[i] = &mps[i];
    }
  }
  ORBmatcher m(0.9f, check_orientation != 0);
  const int nm = m.SearchByProjection(cur, last, th, bMono != 0);
  for (int i = 0; i < nC; ++i) match_out[i] = cur.mvpMapPoints[i] ? (int)(cur.mvpMapPoints[i] - mps.data()) : -1;
  return nm;
}

struct TrackQueryC {   // same layout as orb_track_query (include/orb_b200.h)
  float proj_x, proj_y, proj_xr, view_cos;
  int level, flags;      // bit 0: mbTrackInView && !isBad() && !(bFarPoints && mTrackDepth > thFarPoints); bit 1: Observations() > 0
};

// ORBmatcher::SearchByProjection(F, vpMapPoints, th, bFarPoints, thFarPoints) (local map -> frame), Nleft == -1.
// locked0[i2] != 0: F.mvpMapPoints[i2] already holds a map point with Observations() > 0 when the call starts.
// match_out[i2] = index of the map point the call assigned to keypoint i2, or -1. Returns nmatches.
int refm_search_local_points(const void* kpsC, const uint8_t* descC, const float* uRightC, const uint8_t* locked0, int nC, const float* scale,
                             int nlevels, const float* gp, const TrackQueryC* q, const uint8_t* qdesc, int nq, float th, float nnratio,
                             int* match_out) {
  set_frame_statics(gp);
  Frame f;
  f.N = nC;
  f.mvKeysUn.assign((const cv::KeyPoint*)kpsC, (const cv::KeyPoint*)kpsC + nC);
  f.mvKeys = f.mvKeysUn;
  MapPoint prior;      // stands for "some map point with observations" / "some map point without" already in the frame
  prior.nobs = 1;
  f.mvpMapPoints.assign(nC, (MapPoint*)nullptr);
  for (int i = 0; i < nC; ++i)
    if (locked0[i]) f.mvpMapPoints[i] = &prior;
  f.mDescriptors = cv::Mat(std::max(nC, 1), 32, CV_8UC1);
  if (nC) std::memcpy(f.mDescriptors.data, descC, (size_t)nC * 32);
  f.mvuRight.assign(uRightC, uRightC + nC);
  f.mvScaleFactors.assign(scale, scale + nlevels);
  f.AssignFeaturesToGrid();
  std::vector<MapPoint> mps(nq);
  std::vector<MapPoint*> vp(nq);
  for (int i = 0; i < nq; ++i) {
    mps[i].mbTrackInView = (q[i].flags & 1) != 0;
    mps[i].mTrackProjX = q[i].proj_x; mps[i].mTrackProjY = q[i].proj_y; mps[i].mTrackProjXR = q[i].proj_xr;
    mps[i].mTrackViewCos = q[i].view_cos;
    mps[i].mnTrackScaleLevel = q[i].level;
    mps[i].nobs = (q[i].flags & 2) ? 1 : 0;
    mps[i].desc = cv::Mat(1, 32, CV_8UC1);
    std::memcpy(mps[i].desc.data, qdesc + 32 * (size_t)i, 32);
    vp[i] = &mps[i];
  }
  ORBmatcher m(nnratio, true);
  const int nm = m.SearchByProjection(f, vp, th, false, 50.0f);
  for (int i = 0; i < nC; ++i) {
    MapPoint* p = f.mvpMapPoints[i];
    match_out[i] = (p && p != &prior) ? (int)(p - mps.data()) : -1;
  }
  return nm;
}

static void fill_fv(DBoW2::FeatureVector& fv, const uint32_t* node, const int* off, const uint32_t* feat, int nn) {
  for (int j = 0; j < nn; ++j)
    for (int t = off[j]; t < off[j + 1]; ++t) fv.addFeature(node[j], feat[t]);
}

// ORBmatcher::SearchByBoW(pKF, F, vpMapPointMatches) (src/ORBmatcher.cc:218-395), single camera (Nleft == -1, no mpCamera2).
// kf_flags[i] != 0: the keyframe's keypoint i holds a map point that is not bad. Feature vectors as CSR in map order.
// match_out[iF] = keyframe keypoint whose map point the frame keypoint iF received, or -1. Returns nmatches.
int refm_search_by_bow(const uint8_t* descKF, const float* angleKF, const uint8_t* kf_flags, int nKF, const uint32_t* kf_node,
                       const int* kf_off, const uint32_t* kf_feat, int kf_nn, const uint8_t* descF, const float* angleF, int nF,
                       const uint32_t* f_node, const int* f_off, const uint32_t* f_feat, int f_nn, float nnratio, int check_orientation,
                       int* match_out) {
  KeyFrame kf;
  std::vector<MapPoint> mps(std::max(nKF, 1));
  kf.mvpMapPoints.assign(nKF, (MapPoint*)nullptr);
  for (int i = 0; i < nKF; ++i)
    if (kf_flags[i]) kf.mvpMapPoints[i] = &mps[i];
  kf.mDescriptors = cv::Mat(std::max(nKF, 1), 32, CV_8UC1);
  if (nKF) std::memcpy(kf.mDescriptors.data, descKF, (size_t)nKF * 32);
  kf.mvKeysUn.resize(nKF);
  for (int i = 0; i < nKF; ++i) kf.mvKeysUn[i].angle = angleKF[i];
  kf.mvKeys = kf.mvKeysUn;
  fill_fv(kf.mFeatVec, kf_node, kf_off, kf_feat, kf_nn);
  Frame f;
  f.N = nF;
  f.mDescriptors = cv::Mat(std::max(nF, 1), 32, CV_8UC1);
  if (nF) std::memcpy(f.mDescriptors.data, descF, (size_t)nF * 32);
  f.mvKeys.resize(nF);
  for (int i = 0; i < nF; ++i) f.mvKeys[i].angle = angleF[i];
  f.mvKeysUn = f.mvKeys;
  fill_fv(f.mFeatVec, f_node, f_off, f_feat, f_nn);
  ORBmatcher m(nnratio, check_orientation != 0);
  std::vector<MapPoint*> matches;
  const int nm = m.SearchByBoW(&kf, f, matches);
  for (int i = 0; i < nF; ++i) match_out[i] = matches[i] ? (int)(matches[i] - mps.data()) : -1;
  return nm;
}

// ---- two-camera frames (Nleft != -1: the fisheye rig of TUM-VI) -----------------------------------------------
// A two-camera Frame holds N = Nleft + Nright keypoints: mvKeys (left), mvKeysRight, mDescriptors = the left rows followed by
// the right rows, mvpMapPoints over the combined index space, the second grid mGridRight (src/Frame.cc:510-526) and the stereo
// pairing mvLeftToRightMatch / mvRightToLeftMatch of ComputeStereoFishEyeMatches.
static void fill_two_camera_frame(Frame& f, const void* kpsL, const uint8_t* descL, int nL, const void* kpsR, const uint8_t* descR, int nR,
                                  const float* scale, int nlevels) {
  f.N = nL + nR;
  f.Nleft = nL;
  f.mvKeys.assign((const cv::KeyPoint*)kpsL, (const cv::KeyPoint*)kpsL + nL);
  f.mvKeysRight.assign((const cv::KeyPoint*)kpsR, (const cv::KeyPoint*)kpsR + nR);
  f.mvKeysUn = f.mvKeys;
  f.mvpMapPoints.assign(f.N, (MapPoint*)nullptr);
  f.mDescriptors = cv::Mat(std::max(f.N, 1), 32, CV_8UC1);
  if (nL) std::memcpy(f.mDescriptors.data, descL, (size_t)nL * 32);
  if (nR) std::memcpy(f.mDescriptors.data + (size_t)nL * 32, descR, (size_t)nR * 32);
  f.mvuRight.assign(f.N, -1.f);
  f.mvScaleFactors.assign(scale, scale + nlevels);
  f.AssignFeaturesToGrid();
}

// Frame::GetFeaturesInArea(x, y, r, minLevel, maxLevel, bRight) on a two-camera frame
int refm_features_in_area2(const void* kpsL, int nL, const void* kpsR, int nR, const float* gp, float x, float y, float r, int minLevel,
                           int maxLevel, int bRight, int* out, int cap) {
  set_frame_statics(gp);
  Frame f;
  const float one = 1.f;
  std::vector<uint8_t> dl((size_t)std::max(nL, 1) * 32), dr((size_t)std::max(nR, 1) * 32);
  fill_two_camera_frame(f, kpsL, dl.data(), nL, kpsR, dr.data(), nR, &one, 1);
  vector<size_t> v = f.GetFeaturesInArea(x, y, r, minLevel, maxLevel, bRight != 0);
  if ((int)v.size() > cap) return -2;
  for (size_t i = 0; i < v.size(); ++i) out[i] = (int)v[i];
  return (int)v.size();
}

// ORBmatcher::SearchByProjection(CurrentFrame, LastFrame, th, bMono) with a two-camera CurrentFrame (src/ORBmatcher.cc:1638-1707
// is the right-camera half). The stubs' pure-translation pose makes the right projection the left one shifted by trl = (trl_x,
// trl_y, 0): x3Dr = Trl * x3Dc. Queries as in refm_search_by_projection (octave / angle of the last-frame keypoint).
// match_out[N]: last-frame keypoint whose map point keypoint i2 of the combined index space holds at the end, or -1.
int refm_search_by_projection2(const void* kpsL, const uint8_t* descL, int nL, const void* kpsR, const uint8_t* descR, int nR,
                               const float* scale, int nlevels, const float* gp, float mb, float trl_x, float trl_y, const QueryC* q,
                               const uint8_t* qdesc, int nq, float th, int bMono, float tlc_z, int check_orientation, int* match_out) {
  set_frame_statics(gp);
  GeometricCamera cam;
  Frame cur, last;
  fill_two_camera_frame(cur, kpsL, descL, nL, kpsR, descR, nR, scale, nlevels);
  cur.mb = mb; cur.mbf = 0.f;
  cur.mpCamera = &cam;
  cur.mTrl.t = Eigen::Vector3f(trl_x, trl_y, 0.f);
  last.N = nq;
  last.mvKeys.resize(nq); last.mvKeysUn.resize(nq);
  last.mvpMapPoints.assign(nq, (MapPoint*)nullptr);
  last.mvbOutlier.assign(nq, false);
  last.mTcw.t = Eigen::Vector3f(0.f, 0.f, tlc_z);
  std::vector<MapPoint> mps(std::max(nq, 1));
  for (int i = 0; i < nq; ++i) {
    last.mvKeys[i].octave = q[i].octave; last.mvKeys[i].angle = q[i].angle;
    last.mvKeysUn[i] = last.mvKeys[i];
    if (q[i].flags & 1) {
      mps[i].pos = Eigen::Vector3f(q[i].u, q[i].v, q[i].z);
      mps[i].desc = cv::Mat(1, 32, CV_8UC1);
      std::memcpy(mps[i].desc.data, qdesc + 32 * (size_t)i, 32);
      mps[i].nobs = (q[i].flags & 2) ? 1 : 0;
      last.mvpMapPoints[i] = &mps[i];
    }
  }
  ORBmatcher m(0.9f, check_orientation != 0);
  const int nm = m.SearchByProjection(cur, last, th, bMono != 0);
  for (int i = 0; i < cur.N; ++i) match_out[i] = cur.mvpMapPoints[i] ? (int)(cur.mvpMapPoints[i] - mps.data()) : -1;
  return nm;
}

struct TrackQuery2C {   // same layout as orb_track_query2 (include/orb_b200.h)
  float proj_x, proj_y, view_cos;
  int level;
  float proj_xr, proj_yr, view_cos_r;
  int level_r;
  int flags, pad;        // bit 0: mbTrackInView, bit 1: Observations() > 0, bit 2: mbTrackInViewR (both after isBad / far-point filtering)
};

// ORBmatcher::SearchByProjection(F, vpMapPoints, th, bFarPoints, thFarPoints) with a two-camera F (src/ORBmatcher.cc:127-205 is
// the right-camera half). locked0[N]: keypoints of the combined index space that hold a map point with observations when the
// call starts; l2r[Nleft] / r2l[Nright] = mvLeftToRightMatch / mvRightToLeftMatch. match_out[N] as above.
int refm_search_local_points2(const void* kpsL, const uint8_t* descL, int nL, const void* kpsR, const uint8_t* descR, int nR,
                              const uint8_t* locked0, const int* l2r, const int* r2l, const float* scale, int nlevels, const float* gp,
                              const TrackQuery2C* q, const uint8_t* qdesc, int nq, float th, float nnratio, int* match_out) {
  set_frame_statics(gp);
  Frame f;
  fill_two_camera_frame(f, kpsL, descL, nL, kpsR, descR, nR, scale, nlevels);
  f.mvLeftToRightMatch.assign(l2r, l2r + nL);
  f.mvRightToLeftMatch.assign(r2l, r2l + nR);
  MapPoint prior;
  prior.nobs = 1;
  for (int i = 0; i < f.N; ++i)
    if (locked0[i]) f.mvpMapPoints[i] = &prior;
  std::vector<MapPoint> mps(std::max(nq, 1));
  std::vector<MapPoint*> vp(nq);
  for (int i = 0; i < nq; ++i) {
    mps[i].mbTrackInView = (q[i].flags & 1) != 0;
    mps[i].mbTrackInViewR = (q[i].flags & 4) != 0;
    mps[i].mTrackProjX = q[i].proj_x; mps[i].mTrackProjY = q[i].proj_y; mps[i].mTrackViewCos = q[i].view_cos;
    mps[i].mnTrackScaleLevel = q[i].level;
    mps[i].mTrackProjXR = q[i].proj_xr; mps[i].mTrackProjYR = q[i].proj_yr; mps[i].mTrackViewCosR = q[i].view_cos_r;
    mps[i].mnTrackScaleLevelR = q[i].level_r;
    mps[i].nobs = (q[i].flags & 2) ? 1 : 0;
    mps[i].desc = cv::Mat(1, 32, CV_8UC1);
    std::memcpy(mps[i].desc.data, qdesc + 32 * (size_t)i, 32);
    vp[i] = &mps[i];
  }
  ORBmatcher m(nnratio, true);
  const int nm = m.SearchByProjection(f, vp, th, false, 50.0f);
  for (int i = 0; i < f.N; ++i) {
    MapPoint* p = f.mvpMapPoints[i];
    match_out[i] = (p && p != &prior) ? (int)(p - mps.data()) : -1;
  }
  return nm;
}

// ORBmatcher::SearchByBoW(pKF, F, vpMapPointMatches) with a two-camera F (src/ORBmatcher.cc:218-395, Nleft != -1 branches :283-301,
// :331-352). The frame's descriptors / FeatureVector cover the left keypoints followed by the right ones.
// match_out[N] = keyframe keypoint whose map point feature i of the combined frame received, or -1. Returns nmatches.
int refm_search_by_bow2(const uint8_t* descKF, const float* angleKF, const uint8_t* kf_flags, int nKF, const uint32_t* kf_node,
                        const int* kf_off, const uint32_t* kf_feat, int kf_nn, const uint8_t* descL, const float* angleL, int nL,
                        const uint8_t* descR, const float* angleR, int nR, const uint32_t* f_node, const int* f_off, const uint32_t* f_feat,
                        int f_nn, float nnratio, int check_orientation, int* match_out) {
  GeometricCamera cam;
  KeyFrame kf;
  std::vector<MapPoint> mps(std::max(nKF, 1));
  kf.mvpMapPoints.assign(nKF, (MapPoint*)nullptr);
  for (int i = 0; i < nKF; ++i)
    if (kf_flags[i]) kf.mvpMapPoints[i] = &mps[i];
  kf.mDescriptors = cv::Mat(std::max(nKF, 1), 32, CV_8UC1);
  if (nKF) std::memcpy(kf.mDescriptors.data, descKF, (size_t)nKF * 32);
  kf.mvKeysUn.resize(nKF);
  for (int i = 0; i < nKF; ++i) kf.mvKeysUn[i].angle = angleKF[i];
  kf.mvKeys = kf.mvKeysUn;
  fill_fv(kf.mFeatVec, kf_node, kf_off, kf_feat, kf_nn);
  Frame f;
  f.N = nL + nR;
  f.Nleft = nL;
  f.mpCamera2 = &cam;
  f.mDescriptors = cv::Mat(std::max(f.N, 1), 32, CV_8UC1);
  if (nL) std::memcpy(f.mDescriptors.data, descL, (size_t)nL * 32);
  if (nR) std::memcpy(f.mDescriptors.data + (size_t)nL * 32, descR, (size_t)nR * 32);
  f.mvKeys.resize(nL); f.mvKeysRight.resize(nR);
  for (int i = 0; i < nL; ++i) f.mvKeys[i].angle = angleL[i];
  for (int i = 0; i < nR; ++i) f.mvKeysRight[i].angle = angleR[i];
  f.mvKeysUn = f.mvKeys;
  fill_fv(f.mFeatVec, f_node, f_off, f_feat, f_nn);
  ORBmatcher m(nnratio, check_orientation != 0);
  std::vector<MapPoint*> matches;
  const int nm = m.SearchByBoW(&kf, f, matches);
  for (int i = 0; i < f.N; ++i) match_out[i] = matches[i] ? (int)(matches[i] - mps.data()) : -1;
  return nm;
}

// ORBmatcher::SearchForInitialization(F1, F2, vbPrevMatched, vnMatches12, windowSize) (src/ORBmatcher.cc:603-700).
// prev[n1][2] = vbPrevMatched, updated in place like the reference's vector; matches12_out[n1] = vnMatches12. Returns nmatches.
int refm_search_for_initialization(const void* kps1, const uint8_t* desc1, int n1, const void* kps2, const uint8_t* desc2, int n2,
                                   const float* gp, float* prev, int window, float nnratio, int check_orientation, int* matches12_out) {
  set_frame_statics(gp);
  Frame f1, f2;
  f1.N = n1;
  f1.mvKeysUn.assign((const cv::KeyPoint*)kps1, (const cv::KeyPoint*)kps1 + n1);
  f1.mvKeys = f1.mvKeysUn;
  f1.mDescriptors = cv::Mat(std::max(n1, 1), 32, CV_8UC1);
  if (n1) std::memcpy(f1.mDescriptors.data, desc1, (size_t)n1 * 32);
  f2.N = n2;
  f2.mvKeysUn.assign((const cv::KeyPoint*)kps2, (const cv::KeyPoint*)kps2 + n2);
  f2.mvKeys = f2.mvKeysUn;
  f2.mDescriptors = cv::Mat(std::max(n2, 1), 32, CV_8UC1);
  if (n2) std::memcpy(f2.mDescriptors.data, desc2, (size_t)n2 * 32);
  f2.AssignFeaturesToGrid();
  std::vector<cv::Point2f> vprev(n1);
  for (int i = 0; i < n1; ++i) { vprev[i].x = prev[2 * i]; vprev[i].y = prev[2 * i + 1]; }
  std::vector<int> m12;
  ORBmatcher m(nnratio, check_orientation != 0);
  const int nm = m.SearchForInitialization(f1, f2, vprev, m12, window);
  for (int i = 0; i < n1; ++i) { matches12_out[i] = m12[i]; prev[2 * i] = vprev[i].x; prev[2 * i + 1] = vprev[i].y; }
  return nm;
}

}  // extern "C"
