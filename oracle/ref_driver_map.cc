// TEST INFRASTRUCTURE ONLY (oracle/_ref). Not part of the product path.
//
// C entry points around UNMODIFIED reference sources of the LocalMapping / Relocalization matchers (widening beyond SURVEY.md 8,
// VERDICT round 1 item 9), cut out by line range at build time (oracle/Makefile) into oracle/_ref/*.inc and compiled inside the
// stub classes below:
//   * src/ORBmatcher.cc:1044-1215   ORBmatcher::Fuse(KeyFrame*, const vector<MapPoint*>&, th, bRight)
//   * src/ORBmatcher.cc:1217-1322   ORBmatcher::Fuse(KeyFrame*, Sophus::Sim3f&, const vector<MapPoint*>&, th, vector<MapPoint*>&)
//   * src/ORBmatcher.cc:821-1042    ORBmatcher::SearchForTriangulation
//   * src/ORBmatcher.cc:1735-1842   ORBmatcher::SearchByProjection(Frame&, KeyFrame*, const set<MapPoint*>&, th, ORBdist)
//   * src/ORBmatcher.cc:397-494     ORBmatcher::SearchByProjection(KeyFrame*, Sim3f&, vpPoints, vpMatched, th, ratioHamming)
//   * src/ORBmatcher.cc:702-819     ORBmatcher::SearchByBoW(KeyFrame*, KeyFrame*, vpMatches12)
//   * src/ORBmatcher.cc:1323-1519   ORBmatcher::SearchBySim3
//   * src/KeyFrame.cc:729-778       KeyFrame::GetFeaturesInArea, KeyFrame::IsInImage
//   * src/CameraModels/Pinhole.cpp:125-138   the body of Pinhole::epipolarConstrain after the fundamental matrix
//   * src/MapPoint.cc:367-435       MapPoint::ComputeDistinctiveDescriptors
//   * src/Frame.cc:501-528, 742-820 Frame::AssignFeaturesToGrid / GetFeaturesInArea / PosInGrid
//   * src/ORBmatcher.cc:35-37, 1844-1876, 1880-1894   thresholds, ComputeThreeMaxima, DescriptorDistance
// Eigen and Sophus are not in this image. As in ref_driver_match.cc the stubs give the pose arithmetic pure-translation semantics
// (SE3f * v = v + t, exact for t = 0) and the camera an identity projection, so the tests drive the code AFTER the projection
// with exact inputs; PredictScale returns the level the test stored in the map point (the library takes the predicted level
// from its caller). The map surgery of Fuse (Replace / AddObservation / AddMapPoint) acts on a small model of the map:
//   MapPoint  : bad, nobs (observations elsewhere + in this keyframe), kf_idx (its keypoint in pKF or -1)
//   Replace(o): this' keypoint in pKF goes to o when o is not in pKF yet (ReplaceMapPointMatch + AddObservation), else it is
//               erased; o inherits this' other observations; this turns bad      (src/MapPoint.cc:257-313)
//   AddObservation(pKF, idx): kf_idx = idx, nobs += 2 for a stereo keypoint, else 1 (src/MapPoint.cc:141-167)
// oracle/oracle_map_py.py replays the same model from the library's search results.
// Nothing of the reference is copied into the repository.
#include <climits>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <cassert>
#include <map>
#include <mutex>
#include <set>
#include <tuple>
#include <vector>
#include <algorithm>

#include <opencv2/core/core.hpp>   // the oracle's shim
#include "DBoW2/FeatureVector.h"   // the reference's Thirdparty/DBoW2 (Boost declarations: oracle/shim_dbow)

using namespace std;               // like Thirdparty/DBoW2/DBoW2/TemplatedVocabulary.h, which every reference source includes

namespace Eigen {
struct Vector3f {
  float d[3];
  Vector3f() : d{0, 0, 0} {}
  Vector3f(float x, float y, float z) : d{x, y, z} {}
  float operator()(int i) const { return d[i]; }
  float& operator()(int i) { return d[i]; }
  Vector3f operator-(const Vector3f& o) const { return Vector3f(d[0] - o.d[0], d[1] - o.d[1], d[2] - o.d[2]); }
  Vector3f operator/(float s) const { return Vector3f(d[0] / s, d[1] / s, d[2] / s); }
  float dot(const Vector3f& o) const { return d[0] * o.d[0] + (d[1] * o.d[1] + d[2] * o.d[2]); }
  float norm() const { return std::sqrt(dot(*this)); }
};
struct Vector2f {
  float d[2];
  Vector2f() : d{0, 0} {}
  Vector2f(float x, float y) : d{x, y} {}
  float operator()(int i) const { return d[i]; }
};
struct Matrix3f {
  float m[9];   // row-major
  Matrix3f() : m{1, 0, 0, 0, 1, 0, 0, 0, 1} {}
  float operator()(int r, int c) const { return m[3 * r + c]; }
};
}  // namespace Eigen

namespace Sophus {
struct SE3f {  // pure translation
  Eigen::Vector3f t;
  SE3f() {}
  SE3f(const Eigen::Matrix3f&, const Eigen::Vector3f& tt) : t(tt) {}
  SE3f inverse() const { SE3f r; r.t = Eigen::Vector3f(-t.d[0], -t.d[1], -t.d[2]); return r; }
  Eigen::Vector3f translation() const { return t; }
  Eigen::Matrix3f rotationMatrix() const { return Eigen::Matrix3f(); }
  Eigen::Vector3f operator*(const Eigen::Vector3f& v) const { return Eigen::Vector3f(v.d[0] + t.d[0], v.d[1] + t.d[1], v.d[2] + t.d[2]); }
  SE3f operator*(const SE3f& o) const { SE3f r; r.t = Eigen::Vector3f(t.d[0] + o.t.d[0], t.d[1] + o.t.d[1], t.d[2] + o.t.d[2]); return r; }
};
struct Sim3f {  // unit scale, pure translation
  Eigen::Vector3f t;
  Eigen::Matrix3f rotationMatrix() const { return Eigen::Matrix3f(); }
  Eigen::Vector3f translation() const { return t; }
  float scale() const { return 1.f; }
  Sim3f inverse() const { Sim3f r; r.t = Eigen::Vector3f(-t.d[0], -t.d[1], -t.d[2]); return r; }
  Eigen::Vector3f operator*(const Eigen::Vector3f& v) const { return Eigen::Vector3f(v.d[0] + t.d[0], v.d[1] + t.d[1], v.d[2] + t.d[2]); }
};
}  // namespace Sophus

#define FRAME_GRID_ROWS 48   // include/Frame.h:44-45
#define FRAME_GRID_COLS 64

namespace ORB_SLAM3 {

struct KeyFrame;
struct Frame;

struct MapPoint {
  Eigen::Vector3f pos, normal;
  float min_dist = 0.f, max_dist = 1e30f;
  cv::Mat desc;
  int level = 0;           // what PredictScale returns (host glue in the library's interface)
  int nobs = 0;            // observations (other keyframes + this one)
  int kf_idx = -1;         // keypoint of pKF that observes this map point
  bool bad = false;
  KeyFrame* kf = nullptr;  // the one keyframe of the model
  int id = 0;              // candidates: index >= 0; the keyframe's own initial points: -2 - keypoint
  Eigen::Vector3f GetWorldPos() { return pos; }
  Eigen::Vector3f GetNormal() { return normal; }
  float GetMinDistanceInvariance() { return min_dist; }
  float GetMaxDistanceInvariance() { return max_dist; }
  cv::Mat GetDescriptor() { return desc; }
  int Observations() { return nobs; }
  bool isBad() { return bad; }
  bool IsInKeyFrame(KeyFrame*) { return kf_idx >= 0; }
  std::tuple<int, int> GetIndexInKeyFrame(KeyFrame*) { return std::make_tuple(kf_idx, -1); }
  int PredictScale(const float&, KeyFrame*) { return level; }
  int PredictScale(const float&, Frame*) { return level; }
  void AddObservation(KeyFrame* pKF, int idx);
  void Replace(MapPoint* pMP);
};

struct GeometricCamera {
  Eigen::Matrix3f F12;   // what Pinhole::epipolarConstrain computes at :120-123 (host glue in the library's interface)
  Eigen::Vector2f project(const Eigen::Vector3f& v) { return Eigen::Vector2f(v.d[0], v.d[1]); }
  bool epipolarConstrain(GeometricCamera* pCamera2, const cv::KeyPoint& kp1, const cv::KeyPoint& kp2, const Eigen::Matrix3f& R12,
                         const Eigen::Vector3f& t12, const float sigmaLevel, const float unc) {
#include "pinhole_epi.inc"   // src/CameraModels/Pinhole.cpp:125-138
  }
};

struct KeyFrame {
  int N = 0, NLeft = -1;
  std::vector<cv::KeyPoint> mvKeys, mvKeysUn, mvKeysRight;
  std::vector<float> mvuRight, mvScaleFactors, mvLevelSigma2, mvInvLevelSigma2;
  cv::Mat mDescriptors;
  std::vector<MapPoint*> mvpMapPoints;
  DBoW2::FeatureVector mFeatVec;
  GeometricCamera* mpCamera = nullptr;
  GeometricCamera* mpCamera2 = nullptr;
  float mbf = 0.f;
  float fx = 1.f, fy = 1.f, cx = 0.f, cy = 0.f;   // SearchBySim3 projects with these: u = fx * x / z + cx is exact for z = 1
  float mnMinX = 0, mnMinY = 0, mnMaxX = 0, mnMaxY = 0, mfGridElementWidthInv = 0, mfGridElementHeightInv = 0;
  int mnGridCols = FRAME_GRID_COLS, mnGridRows = FRAME_GRID_ROWS;
  std::vector<std::vector<std::vector<size_t> > > mGrid, mGridRight;
  Sophus::SE3f mTcw;
  Eigen::Vector3f mOw;
  Sophus::SE3f GetPose() { return mTcw; }
  Sophus::SE3f GetPoseInverse() { return mTcw.inverse(); }
  Sophus::SE3f GetRightPose() { return mTcw; }
  Sophus::SE3f GetRightPoseInverse() { return mTcw.inverse(); }
  Eigen::Vector3f GetCameraCenter() { return mOw; }
  Eigen::Vector3f GetRightCameraCenter() { return mOw; }
  MapPoint* GetMapPoint(const size_t& idx) { return mvpMapPoints[idx]; }
  void AddMapPoint(MapPoint* pMP, const size_t& idx) { mvpMapPoints[idx] = pMP; }
  std::vector<MapPoint*> GetMapPointMatches() { return mvpMapPoints; }
  std::set<MapPoint*> GetMapPoints() {
    std::set<MapPoint*> s;
    for (MapPoint* p : mvpMapPoints)
      if (p && !p->isBad()) s.insert(p);   // src/KeyFrame.cc:322-333
    return s;
  }
  bool isBad() { return false; }
  std::vector<size_t> GetFeaturesInArea(const float& x, const float& y, const float& r, const bool bRight = false) const;  // include/KeyFrame.h
  bool IsInImage(const float& x, const float& y) const;
};

// what Fuse did, in call order: (1, map point, keypoint) = AddObservation + AddMapPoint, (2, replaced, replacement) = Replace
static std::vector<int> g_events;
static bool g_nested = false;
void MapPoint::AddObservation(KeyFrame* pKF, int idx) {
  if (!g_nested) { g_events.push_back(1); g_events.push_back(id); g_events.push_back(idx); }
  kf_idx = idx;
  nobs += (!pKF->mpCamera2 && pKF->mvuRight[idx] >= 0) ? 2 : 1;
}
void MapPoint::Replace(MapPoint* pMP) {
  g_events.push_back(2); g_events.push_back(id); g_events.push_back(pMP->id);
  g_nested = true;
  if (kf_idx >= 0) {
    const int in_kf = (!kf->mpCamera2 && kf->mvuRight[kf_idx] >= 0) ? 2 : 1;
    nobs -= in_kf;
    if (!pMP->IsInKeyFrame(kf)) {
      kf->mvpMapPoints[kf_idx] = pMP;
      pMP->AddObservation(kf, kf_idx);
    } else {
      kf->mvpMapPoints[kf_idx] = nullptr;
    }
    kf_idx = -1;
  }
  pMP->nobs += nobs;   // the observations in other keyframes move over
  nobs = 0;
  bad = true;
  g_nested = false;
}

#include "keyframe_area.inc"   // src/KeyFrame.cc:729-778

struct Frame {
  DBoW2::FeatureVector mFeatVec;
  GeometricCamera* mpCamera2 = nullptr;
  int N = 0, Nleft = -1;
  std::vector<cv::KeyPoint> mvKeys, mvKeysUn, mvKeysRight;
  std::vector<MapPoint*> mvpMapPoints;
  cv::Mat mDescriptors;
  std::vector<float> mvuRight;
  std::vector<float> mvScaleFactors;
  static float mnMinX, mnMaxX, mnMinY, mnMaxY;
  static float mfGridElementWidthInv, mfGridElementHeightInv;
  std::vector<std::size_t> mGrid[FRAME_GRID_COLS][FRAME_GRID_ROWS];
  std::vector<std::size_t> mGridRight[FRAME_GRID_COLS][FRAME_GRID_ROWS];
  GeometricCamera* mpCamera = nullptr;
  Sophus::SE3f mTcw;
  Sophus::SE3f GetPose() const { return mTcw; }
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
  int Fuse(KeyFrame* pKF, const vector<MapPoint*>& vpMapPoints, const float th = 3.0, const bool bRight = false);   // include/ORBmatcher.h:93
  int Fuse(KeyFrame* pKF, Sophus::Sim3f& Scw, const std::vector<MapPoint*>& vpPoints, float th, vector<MapPoint*>& vpReplacePoint);
  int SearchForTriangulation(KeyFrame* pKF1, KeyFrame* pKF2, std::vector<pair<size_t, size_t> >& vMatchedPairs, const bool bOnlyStereo,
                             const bool bCoarse = false);
  int SearchByProjection(Frame& CurrentFrame, KeyFrame* pKF, const std::set<MapPoint*>& sAlreadyFound, const float th, const int ORBdist);
  int SearchByProjection(KeyFrame* pKF, Sophus::Sim3f& Scw, const std::vector<MapPoint*>& vpPoints, std::vector<MapPoint*>& vpMatched, int th,
                         float ratioHamming = 1.0);   // include/ORBmatcher.h:64-66
  int SearchBySim3(KeyFrame* pKF1, KeyFrame* pKF2, std::vector<MapPoint*>& vpMatches12, const Sophus::Sim3f& S12, const float th);
  int SearchByBoW(KeyFrame* pKF1, KeyFrame* pKF2, std::vector<MapPoint*>& vpMatches12);
  void ComputeThreeMaxima(std::vector<int>* histo, const int L, int& ind1, int& ind2, int& ind3);
};
#include "orbmatcher_consts3.inc"   // src/ORBmatcher.cc:35-37
#include "orbmatcher_sft.inc"       // src/ORBmatcher.cc:821-1042
#include "orbmatcher_fuse.inc"      // src/ORBmatcher.cc:1044-1215
#include "orbmatcher_fuse_sim3.inc" // src/ORBmatcher.cc:1217-1322
#include "orbmatcher_sbp_kf.inc"    // src/ORBmatcher.cc:1735-1842
#include "orbmatcher_sbp_sim3.inc"  // src/ORBmatcher.cc:397-494
#include "orbmatcher_sbow_kf.inc"   // src/ORBmatcher.cc:702-819
#include "orbmatcher_sbs3.inc"      // src/ORBmatcher.cc:1323-1519
#include "orbmatcher_max3.inc"      // src/ORBmatcher.cc:1844-1876
#include "orbmatcher_dist.inc"      // src/ORBmatcher.cc:1880-1894

// MapPoint::ComputeDistinctiveDescriptors works on the observation map of the real class: a second, minimal class of that name
// in its own namespace takes the method's lines
namespace distinctive {
struct KeyFrame {
  cv::Mat mDescriptors;
  bool bad = false;
  bool isBad() { return bad; }
};
using ORB_SLAM3::ORBmatcher;
struct MapPoint {
  std::map<KeyFrame*, std::tuple<int, int> > mObservations;
  std::mutex mMutexFeatures;
  bool mbBad = false;
  cv::Mat mDescriptor;
  void ComputeDistinctiveDescriptors();
};
#include "mappoint_distinctive.inc"   // src/MapPoint.cc:367-435
}  // namespace distinctive

}  // namespace ORB_SLAM3

using namespace ORB_SLAM3;

static void fill_keyframe(KeyFrame& kf, GeometricCamera* cam, const void* kps, const uint8_t* desc, const float* uright, int n, const float* gp,
                          const float* scale, const float* sigma2, int nlevels) {
  Frame::mnMinX = gp[0]; Frame::mnMinY = gp[1]; Frame::mnMaxX = gp[2]; Frame::mnMaxY = gp[3];
  Frame::mfGridElementWidthInv = gp[4]; Frame::mfGridElementHeightInv = gp[5];
  Frame f;   // the keyframe's grid is the grid of the Frame it was made from (src/KeyFrame.cc:105-118)
  f.N = n;
  f.mvKeysUn.assign((const cv::KeyPoint*)kps, (const cv::KeyPoint*)kps + n);
  f.mvKeys = f.mvKeysUn;
  f.AssignFeaturesToGrid();
  kf.N = n;
  kf.mvKeysUn = f.mvKeysUn;
  kf.mvKeys = f.mvKeys;
  kf.mvuRight.assign(n, -1.f);
  if (uright) kf.mvuRight.assign(uright, uright + n);
  kf.mDescriptors = cv::Mat(std::max(n, 1), 32, CV_8UC1);
  if (n) std::memcpy(kf.mDescriptors.data, desc, (size_t)n * 32);
  kf.mvpMapPoints.assign(n, (MapPoint*)nullptr);
  kf.mpCamera = cam;
  kf.mnMinX = gp[0]; kf.mnMinY = gp[1]; kf.mnMaxX = gp[2]; kf.mnMaxY = gp[3];
  kf.mfGridElementWidthInv = gp[4]; kf.mfGridElementHeightInv = gp[5];
  kf.mGrid.resize(kf.mnGridCols);
  for (int i = 0; i < kf.mnGridCols; ++i) {
    kf.mGrid[i].resize(kf.mnGridRows);
    for (int j = 0; j < kf.mnGridRows; ++j) kf.mGrid[i][j] = f.mGrid[i][j];
  }
  kf.mvScaleFactors.assign(scale, scale + nlevels);
  kf.mvLevelSigma2.assign(sigma2, sigma2 + nlevels);
  kf.mvInvLevelSigma2.resize(nlevels);
  for (int l = 0; l < nlevels; ++l) kf.mvInvLevelSigma2[l] = 1.0f / sigma2[l];   // src/ORBextractor.cc:424-427
}

struct FusePointC {   // one candidate map point of Fuse
  float x, y, z;        // world position = camera coordinates (identity pose): projects to (x, y), depth z
  float nx, ny, nz;     // GetNormal()
  float min_dist, max_dist;
  int level;            // PredictScale
  int nobs;             // Observations() elsewhere
  int flags;            // bit 0: NULL pointer, bit 1: isBad(), bit 2: duplicate of the previous entry (the same MapPoint object again)
};

extern "C" {

// ORBmatcher::Fuse(pKF, vpMapPoints, th, false) (sim3 == 0) or Fuse(pKF, Scw, vpPoints, th, vpReplacePoint) (sim3 != 0).
// kf_mp_nobs[i] >= 0: keypoint i of pKF already holds a map point with that many observations (kf_mp_bad[i]: it is bad).
// Outputs: events[3 * k] = what the call did in order: (1, candidate, keypoint) AddObservation + AddMapPoint, (2, replaced,
// replacement) Replace, with candidates as their index and the keyframe's own initial point of keypoint j as -2 - j;
// repl[nq] (sim3) = keypoint whose (own) map point vpReplacePoint[i] received, -1000 - c for candidate c, or -1; kf_final[n] = map point the keypoint holds at
// the end (-1 none); cand_bad / cand_nobs[nq] = isBad() / Observations() of the candidates afterwards. Returns nFused.
static int run_fuse(KeyFrame& kf, int n, const int* kf_mp_nobs, const uint8_t* kf_mp_bad, const FusePointC* pts, const uint8_t* pdesc, int nq,
                    float th, int sim3, bool right, int* events, int events_cap, int* n_events, int* repl, int* kf_final, int* cand_bad,
                    int* cand_nobs) {
  std::vector<MapPoint> own(std::max(n, 1));
  for (int i = 0; i < n; ++i) {
    own[i].id = -2 - i;
    if (kf_mp_nobs[i] >= 0) {
      own[i].kf = &kf; own[i].kf_idx = i; own[i].nobs = kf_mp_nobs[i]; own[i].bad = kf_mp_bad[i] != 0;
      kf.mvpMapPoints[i] = &own[i];
    }
  }
  std::vector<MapPoint> mps(std::max(nq, 1));
  std::vector<MapPoint*> vp(nq, (MapPoint*)nullptr);
  for (int i = 0; i < nq; ++i) {
    MapPoint& m = mps[i];
    m.kf = &kf; m.id = i;
    m.pos = Eigen::Vector3f(pts[i].x, pts[i].y, pts[i].z);
    m.normal = Eigen::Vector3f(pts[i].nx, pts[i].ny, pts[i].nz);
    m.min_dist = pts[i].min_dist; m.max_dist = pts[i].max_dist;
    m.level = pts[i].level; m.nobs = pts[i].nobs; m.bad = (pts[i].flags & 2) != 0;
    m.desc = cv::Mat(1, 32, CV_8UC1);
    std::memcpy(m.desc.data, pdesc + 32 * (size_t)i, 32);
    if (pts[i].flags & 1) vp[i] = nullptr;
    else if ((pts[i].flags & 4) && i > 0 && vp[i - 1]) vp[i] = vp[i - 1];
    else vp[i] = &m;
  }
  g_events.clear();
  g_nested = false;
  ORBmatcher m(0.6f, true);
  int nf;
  std::vector<MapPoint*> rp(nq, (MapPoint*)nullptr);
  if (sim3) {
    Sophus::Sim3f Scw;
    std::vector<MapPoint*> vq(nq);
    for (int i = 0; i < nq; ++i) vq[i] = vp[i] ? vp[i] : &mps[i];   // the Sim3 overload dereferences every pointer: no NULLs
    nf = m.Fuse(&kf, Scw, vq, th, rp);
  } else {
    nf = m.Fuse(&kf, vp, th, right);
  }
  *n_events = (int)g_events.size() / 3;
  if ((int)g_events.size() > events_cap) return -1000;
  for (size_t i = 0; i < g_events.size(); ++i) events[i] = g_events[i];
  for (int i = 0; i < nq; ++i) {
    repl[i] = rp[i] ? (rp[i]->id <= -2 ? -2 - rp[i]->id : -1000 - rp[i]->id) : -1;   // keypoint of the keyframe's own point, or -1000 - candidate
    cand_bad[i] = mps[i].bad; cand_nobs[i] = mps[i].nobs;
  }
  for (int i = 0; i < n; ++i) {
    MapPoint* p = kf.mvpMapPoints[i];
    kf_final[i] = p ? p->id : -1;
  }
  return nf;
}

int refmap_fuse(const void* kps, const uint8_t* desc, const float* uright, int n, const float* gp, const float* scale, const float* sigma2,
                int nlevels, float bf, const int* kf_mp_nobs, const uint8_t* kf_mp_bad, const FusePointC* pts, const uint8_t* pdesc, int nq,
                float th, int sim3, int* events, int events_cap, int* n_events, int* repl, int* kf_final, int* cand_bad, int* cand_nobs) {
  GeometricCamera cam;
  KeyFrame kf;
  fill_keyframe(kf, &cam, kps, desc, uright, n, gp, scale, sigma2, nlevels);
  kf.mbf = bf;
  return run_fuse(kf, n, kf_mp_nobs, kf_mp_bad, pts, pdesc, nq, th, sim3, false, events, events_cap, n_events, repl, kf_final, cand_bad, cand_nobs);
}

// ORBmatcher::Fuse(pKF, vpMapPoints, th, bRight = true) on a two-camera keyframe (NLeft = nL, mpCamera2 set): the window search runs on
// mGridRight / mvKeysRight, the match lands on the combined index NLeft + idx (:1173) and mvuRight is -1 everywhere. kf_mp_nobs /
// kf_mp_bad / kf_final cover the combined index space [0, nL + nR).
int refmap_fuse_right(const void* kpsL, const uint8_t* descL, int nL, const void* kpsR, const uint8_t* descR, int nR, const float* gp,
                      const float* scale, const float* sigma2, int nlevels, float bf, const int* kf_mp_nobs, const uint8_t* kf_mp_bad,
                      const FusePointC* pts, const uint8_t* pdesc, int nq, float th, int* events, int events_cap, int* n_events, int* repl,
                      int* kf_final, int* cand_bad, int* cand_nobs) {
  GeometricCamera cam, cam2;
  KeyFrame kf;
  fill_keyframe(kf, &cam, kpsL, descL, nullptr, nL, gp, scale, sigma2, nlevels);   // left half: mvKeys, mGrid
  Frame f;                                                                           // the right grid, as Frame::AssignFeaturesToGrid fills it
  f.N = nL + nR; f.Nleft = nL;
  f.mvKeys = kf.mvKeys;
  f.mvKeysRight.assign((const cv::KeyPoint*)kpsR, (const cv::KeyPoint*)kpsR + nR);
  f.AssignFeaturesToGrid();
  kf.N = nL + nR; kf.NLeft = nL;
  kf.mvKeysRight = f.mvKeysRight;
  kf.mpCamera2 = &cam2;
  kf.mvuRight.assign(kf.N, -1.f);
  kf.mDescriptors = cv::Mat(std::max(kf.N, 1), 32, CV_8UC1);
  if (nL) std::memcpy(kf.mDescriptors.data, descL, (size_t)nL * 32);
  if (nR) std::memcpy(kf.mDescriptors.data + (size_t)nL * 32, descR, (size_t)nR * 32);
  kf.mvpMapPoints.assign(kf.N, (MapPoint*)nullptr);
  kf.mGridRight.resize(kf.mnGridCols);
  for (int i = 0; i < kf.mnGridCols; ++i) {
    kf.mGridRight[i].resize(kf.mnGridRows);
    for (int j = 0; j < kf.mnGridRows; ++j) kf.mGridRight[i][j] = f.mGridRight[i][j];
  }
  kf.mbf = bf;
  return run_fuse(kf, kf.N, kf_mp_nobs, kf_mp_bad, pts, pdesc, nq, th, 0, true, events, events_cap, n_events, repl, kf_final, cand_bad, cand_nobs);
}

// KeyFrame::GetFeaturesInArea(x, y, r) on a keyframe built from the keypoints
int refmap_features_in_area(const void* kps, int n, const float* gp, float x, float y, float r, int* out, int cap) {
  GeometricCamera cam;
  KeyFrame kf;
  const float one = 1.f;
  std::vector<uint8_t> d((size_t)std::max(n, 1) * 32);
  fill_keyframe(kf, &cam, kps, d.data(), nullptr, n, gp, &one, &one, 1);
  vector<size_t> v = kf.GetFeaturesInArea(x, y, r);
  if ((int)v.size() > cap) return -2;
  for (size_t i = 0; i < v.size(); ++i) out[i] = (int)v[i];
  return (int)v.size();
}

static void fill_fv(DBoW2::FeatureVector& fv, const uint32_t* node, const int* off, const uint32_t* feat, int nn) {
  for (int j = 0; j < nn; ++j)
    for (int t = off[j]; t < off[j + 1]; ++t) fv.addFeature(node[j], feat[t]);
}

// ORBmatcher::SearchForTriangulation(pKF1, pKF2, vMatchedPairs, bOnlyStereo, bCoarse), single-camera keyframes. F12 row-major =
// the fundamental matrix of Pinhole::epipolarConstrain, (epx, epy) the epipole (the camera centre of pKF1 is (epx, epy, 1), the
// pose of pKF2 the identity). match12[n1] = vMatches12 rebuilt from vMatchedPairs. Returns nmatches.
int refmap_search_for_triangulation(const void* kps1, const uint8_t* desc1, const float* ur1, const uint8_t* has_mp1, int n1,
                                    const uint32_t* node1, const int* off1, const uint32_t* feat1, int nn1, const void* kps2,
                                    const uint8_t* desc2, const float* ur2, const uint8_t* has_mp2, int n2, const uint32_t* node2,
                                    const int* off2, const uint32_t* feat2, int nn2, const float* gp, const float* scale, const float* sigma2,
                                    int nlevels, const float* F12, float epx, float epy, int only_stereo, int coarse, int check_orientation,
                                    int* match12) {
  GeometricCamera cam1, cam2;
  std::memcpy(cam1.F12.m, F12, 36);
  KeyFrame k1, k2;
  fill_keyframe(k1, &cam1, kps1, desc1, ur1, n1, gp, scale, sigma2, nlevels);
  fill_keyframe(k2, &cam2, kps2, desc2, ur2, n2, gp, scale, sigma2, nlevels);
  MapPoint some;
  for (int i = 0; i < n1; ++i) if (has_mp1[i]) k1.mvpMapPoints[i] = &some;
  for (int i = 0; i < n2; ++i) if (has_mp2[i]) k2.mvpMapPoints[i] = &some;
  fill_fv(k1.mFeatVec, node1, off1, feat1, nn1);
  fill_fv(k2.mFeatVec, node2, off2, feat2, nn2);
  k1.mOw = Eigen::Vector3f(epx, epy, 1.f);
  ORBmatcher m(0.6f, check_orientation != 0);
  std::vector<pair<size_t, size_t> > pairs;
  const int nm = m.SearchForTriangulation(&k1, &k2, pairs, only_stereo != 0, coarse != 0);
  for (int i = 0; i < n1; ++i) match12[i] = -1;
  for (size_t i = 0; i < pairs.size(); ++i) match12[pairs[i].first] = (int)pairs[i].second;
  return nm;
}

struct QueryC {   // same layout as orb_proj_query (include/orb_b200.h)
  float u, v, z, angle;
  int octave, flags;
};

// ORBmatcher::SearchByProjection(CurrentFrame, pKF, sAlreadyFound, th, ORBdist). One query per keyframe map point: position
// (u, v, z) (identity pose and projection), angle of pKF->mvKeysUn[i], octave = what PredictScale returns, flags bit 0: the map
// point exists, bit 2: it is bad, bit 3: it is in sAlreadyFound. locked0[i2] != 0: CurrentFrame.mvpMapPoints[i2] holds a map
// point when the call starts. match_out[i2] = query whose map point keypoint i2 received, or -1. Returns nmatches.
int refmap_search_by_projection_kf(const void* kpsC, const uint8_t* descC, const uint8_t* locked0, int nC, const float* scale, int nlevels,
                                   const float* gp, const QueryC* q, const uint8_t* qdesc, int nq, float th, int orb_dist,
                                   int check_orientation, int* match_out) {
  Frame::mnMinX = gp[0]; Frame::mnMinY = gp[1]; Frame::mnMaxX = gp[2]; Frame::mnMaxY = gp[3];
  Frame::mfGridElementWidthInv = gp[4]; Frame::mfGridElementHeightInv = gp[5];
  GeometricCamera cam;
  Frame cur;
  cur.N = nC;
  cur.mvKeysUn.assign((const cv::KeyPoint*)kpsC, (const cv::KeyPoint*)kpsC + nC);
  cur.mvKeys = cur.mvKeysUn;
  MapPoint prior;
  cur.mvpMapPoints.assign(nC, (MapPoint*)nullptr);
  for (int i = 0; i < nC; ++i) if (locked0 && locked0[i]) cur.mvpMapPoints[i] = &prior;
  cur.mDescriptors = cv::Mat(std::max(nC, 1), 32, CV_8UC1);
  if (nC) std::memcpy(cur.mDescriptors.data, descC, (size_t)nC * 32);
  cur.mvScaleFactors.assign(scale, scale + nlevels);
  cur.mpCamera = &cam;
  cur.AssignFeaturesToGrid();
  KeyFrame kf;
  kf.N = nq;
  kf.mvKeysUn.resize(nq);
  kf.mvpMapPoints.assign(nq, (MapPoint*)nullptr);
  std::vector<MapPoint> mps(std::max(nq, 1));
  std::set<MapPoint*> found;
  for (int i = 0; i < nq; ++i) {
    kf.mvKeysUn[i].angle = q[i].angle;
    if (q[i].flags & 1) {
      mps[i].pos = Eigen::Vector3f(q[i].u, q[i].v, q[i].z);
      mps[i].level = q[i].octave;
      mps[i].bad = (q[i].flags & 4) != 0;
      mps[i].desc = cv::Mat(1, 32, CV_8UC1);
      std::memcpy(mps[i].desc.data, qdesc + 32 * (size_t)i, 32);
      kf.mvpMapPoints[i] = &mps[i];
      if (q[i].flags & 8) found.insert(&mps[i]);
    }
  }
  ORBmatcher m(0.9f, check_orientation != 0);
  const int nm = m.SearchByProjection(cur, &kf, found, th, orb_dist);
  for (int i = 0; i < nC; ++i) {
    MapPoint* p = cur.mvpMapPoints[i];
    match_out[i] = (p && p != &prior) ? (int)(p - mps.data()) : -1;
  }
  return nm;
}

// MapPoint::ComputeDistinctiveDescriptors on a map point observed by n keyframes, one descriptor each (leftIndex = 0,
// rightIndex = -1). The observation map is ordered by the keyframe pointers: the keyframes live in one array, so the order is the
// order of the rows. Returns the row whose descriptor the map point holds afterwards (first row with that content), -1 if none.
int refmap_distinctive(const uint8_t* desc, int n) {
  std::vector<distinctive::KeyFrame> kfs(std::max(n, 1));
  distinctive::MapPoint mp;
  for (int i = 0; i < n; ++i) {
    kfs[i].mDescriptors = cv::Mat(1, 32, CV_8UC1);
    std::memcpy(kfs[i].mDescriptors.data, desc + 32 * (size_t)i, 32);
    mp.mObservations[&kfs[i]] = std::make_tuple(0, -1);
  }
  mp.ComputeDistinctiveDescriptors();
  if (mp.mDescriptor.empty()) return -1;
  for (int i = 0; i < n; ++i)
    if (std::memcmp(mp.mDescriptor.data, desc + 32 * (size_t)i, 32) == 0) return i;
  return -2;
}

// ORBmatcher::SearchByProjection(pKF, Scw, vpPoints, vpMatched, th, ratioHamming) (:397-494). matched0[i] != 0: vpMatched[i] holds some
// other map point when the call starts; found_slot[q] >= 0: candidate q itself sits in vpMatched[found_slot[q]] (it is in
// spAlreadyFound). Queries: position (u, v, z), octave = PredictScale, flags bit 0 set, bit 2: bad. match_out[i] = candidate
// vpMatched[i] received during the call, or -1. Returns nmatches.
int refmap_search_by_projection_sim3(const void* kps, const uint8_t* desc, const uint8_t* matched0, int n, const float* gp, const float* scale,
                                     const float* sigma2, int nlevels, const QueryC* q, const uint8_t* qdesc, const int* found_slot, int nq, int th,
                                     float ratio, int* match_out) {
  GeometricCamera cam;
  KeyFrame kf;
  fill_keyframe(kf, &cam, kps, desc, nullptr, n, gp, scale, sigma2, nlevels);
  MapPoint prior;
  std::vector<MapPoint*> matched(n, (MapPoint*)nullptr);
  for (int i = 0; i < n; ++i) if (matched0[i]) matched[i] = &prior;
  std::vector<MapPoint> mps(std::max(nq, 1));
  std::vector<MapPoint*> vp(nq);
  for (int i = 0; i < nq; ++i) {
    mps[i].pos = Eigen::Vector3f(q[i].u, q[i].v, q[i].z);
    mps[i].normal = Eigen::Vector3f(q[i].u, q[i].v, q[i].z);   // seen head-on: PO . Pn = |PO|^2 >= 0.5 |PO| for |PO| >= 0.5
    mps[i].level = q[i].octave;
    mps[i].bad = (q[i].flags & 4) != 0;
    mps[i].desc = cv::Mat(1, 32, CV_8UC1);
    std::memcpy(mps[i].desc.data, qdesc + 32 * (size_t)i, 32);
    vp[i] = &mps[i];
    if (found_slot[i] >= 0 && found_slot[i] < n) matched[found_slot[i]] = &mps[i];
  }
  std::vector<MapPoint*> before = matched;
  Sophus::Sim3f Scw;
  ORBmatcher m(0.75f, true);
  const int nm = m.SearchByProjection(&kf, Scw, vp, matched, th, ratio);
  for (int i = 0; i < n; ++i) match_out[i] = (matched[i] && matched[i] != before[i]) ? (int)(matched[i] - mps.data()) : -1;
  return nm;
}

struct Sim3PointC {   // a map point of one keyframe: where it projects in the OTHER keyframe and what PredictScale returns there
  float u, v;
  int level;
  int flags;          // bit 0: the keypoint holds a map point, bit 1: it is bad
};

// ORBmatcher::SearchBySim3(pKF1, pKF2, vpMatches12, S12, th) (:1323-1519) with identity poses / S12 and fx = fy = 1, cx = cy = 0:
// the map point of keypoint i sits at (u, v, 1). init12[i] >= 0: vpMatches12[i] holds the map point of pKF2's keypoint init12[i]
// when the call starts. match12[i] = keypoint of pKF2 whose map point vpMatches12[i] holds at the end, or -1. Returns nFound.
int refmap_search_by_sim3(const void* kps1, const uint8_t* desc1, int n1, const Sim3PointC* p1, const uint8_t* pdesc1, const void* kps2,
                          const uint8_t* desc2, int n2, const Sim3PointC* p2, const uint8_t* pdesc2, const float* gp, const float* scale,
                          const float* sigma2, int nlevels, const int* init12, float th, int* match12) {
  GeometricCamera cam1, cam2;
  KeyFrame k1, k2;
  fill_keyframe(k1, &cam1, kps1, desc1, nullptr, n1, gp, scale, sigma2, nlevels);
  fill_keyframe(k2, &cam2, kps2, desc2, nullptr, n2, gp, scale, sigma2, nlevels);
  std::vector<MapPoint> m1(std::max(n1, 1)), m2(std::max(n2, 1));
  auto fill = [](std::vector<MapPoint>& mps, KeyFrame& kf, const Sim3PointC* p, const uint8_t* pd, int n) {
    for (int i = 0; i < n; ++i) {
      if (!(p[i].flags & 1)) continue;
      mps[i].pos = Eigen::Vector3f(p[i].u, p[i].v, 1.f);
      mps[i].level = p[i].level;
      mps[i].bad = (p[i].flags & 2) != 0;
      mps[i].kf = &kf; mps[i].kf_idx = i;
      mps[i].desc = cv::Mat(1, 32, CV_8UC1);
      std::memcpy(mps[i].desc.data, pd + 32 * (size_t)i, 32);
      kf.mvpMapPoints[i] = &mps[i];
    }
  };
  fill(m1, k1, p1, pdesc1, n1);
  fill(m2, k2, p2, pdesc2, n2);
  std::vector<MapPoint*> v12(n1, (MapPoint*)nullptr);
  for (int i = 0; i < n1; ++i)
    if (init12[i] >= 0 && init12[i] < n2 && k2.mvpMapPoints[init12[i]]) v12[i] = k2.mvpMapPoints[init12[i]];
  Sophus::Sim3f S12;
  ORBmatcher m(0.75f, true);
  const int nf = m.SearchBySim3(&k1, &k2, v12, S12, th);
  for (int i = 0; i < n1; ++i) match12[i] = v12[i] ? (int)(v12[i] - m2.data()) : -1;
  return nf;
}

// ORBmatcher::SearchByBoW(pKF1, pKF2, vpMatches12) (:702-819). mp_state[i]: 0 no map point, 1 a good one, 2 a bad one.
// match12[i] = keypoint of pKF2 whose map point vpMatches12[i] received, or -1. Returns nmatches.
int refmap_search_by_bow_kf(const void* kps1, const uint8_t* desc1, const uint8_t* mp_state1, int n1, const uint32_t* node1, const int* off1,
                            const uint32_t* feat1, int nn1, const void* kps2, const uint8_t* desc2, const uint8_t* mp_state2, int n2,
                            const uint32_t* node2, const int* off2, const uint32_t* feat2, int nn2, const float* gp, float nnratio,
                            int check_orientation, int* match12) {
  GeometricCamera cam1, cam2;
  const float one = 1.f;
  KeyFrame k1, k2;
  fill_keyframe(k1, &cam1, kps1, desc1, nullptr, n1, gp, &one, &one, 1);
  fill_keyframe(k2, &cam2, kps2, desc2, nullptr, n2, gp, &one, &one, 1);
  std::vector<MapPoint> m1(std::max(n1, 1)), m2(std::max(n2, 1));
  for (int i = 0; i < n1; ++i) if (mp_state1[i]) { m1[i].bad = mp_state1[i] == 2; k1.mvpMapPoints[i] = &m1[i]; }
  for (int i = 0; i < n2; ++i) if (mp_state2[i]) { m2[i].bad = mp_state2[i] == 2; k2.mvpMapPoints[i] = &m2[i]; }
  fill_fv(k1.mFeatVec, node1, off1, feat1, nn1);
  fill_fv(k2.mFeatVec, node2, off2, feat2, nn2);
  ORBmatcher m(nnratio, check_orientation != 0);
  std::vector<MapPoint*> v12;
  const int nm = m.SearchByBoW(&k1, &k2, v12);
  for (int i = 0; i < n1; ++i) match12[i] = v12[i] ? (int)(v12[i] - m2.data()) : -1;
  return nm;
}

}  // extern "C"
