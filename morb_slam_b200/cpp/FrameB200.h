// Reference-typed helpers for the Frame / ORBmatcher steps either side of the extractor, backed by the B200 C ABI
// (include/orb_b200.h). Header-only; every function names the reference code it stands for. They work on the
// device-resident results of the extractor's last call (batch 1), so nothing but the small per-keypoint outputs moves.
// Compiles against OpenCV's core headers (cv::Mat, cv::KeyPoint). BowVector / FeatureVector are template parameters so
// that this header does not depend on DBoW2 (any std::map<unsigned, double> / std::map<unsigned, std::vector<unsigned>>).
#ifndef FRAME_B200_H
#define FRAME_B200_H

#include <algorithm>
#include <stdexcept>
#include <string>
#include <vector>

#include "ORBextractor.h"
#include "orb_b200.h"

namespace ORB_SLAM3 {

inline void CheckB200(orb_handle* h, int st, const char* what) {
  if (st != ORB_OK) throw std::runtime_error(std::string(what) + ": " + orb_status_string(st) + " - " + (h ? orb_last_error(h) : ""));
}

// Frame::UndistortKeyPoints (src/Frame.cc:829-857). K = toK(), mK, distCoef: CV_32F, continuous.
inline void UndistortKeyPointsB200(ORBextractor* ex, const cv::Mat& K, const cv::Mat& distCoef, const cv::Mat& mK,
                                   const std::vector<cv::KeyPoint>& mvKeys, std::vector<cv::KeyPoint>& mvKeysUn) {
  orb_handle* h = ex->Handle();
  std::vector<orb_keypoint> un(orb_keypoint_capacity(h));
  CheckB200(h, orb_undistort_keypoints(h, (const float*)K.data, (const float*)distCoef.data, distCoef.rows * distCoef.cols, (const float*)mK.data,
                                       un.data(), (int)un.size(), 0), "orb_undistort_keypoints");
  mvKeysUn = mvKeys;                                     // kp = mvKeys[i] with the undistorted pt (:851-856)
  for (size_t i = 0; i < mvKeysUn.size(); ++i) { mvKeysUn[i].pt.x = un[i].x; mvKeysUn[i].pt.y = un[i].y; }
}

// Frame::AssignFeaturesToGrid (src/Frame.cc:501-528) on the resident mvKeysUn; the grid stays on the device for the searches.
inline void AssignFeaturesToGridB200(ORBextractor* ex, float mnMinX, float mnMinY, float mnMaxX, float mnMaxY, float mfGridElementWidthInv,
                                     float mfGridElementHeightInv) {
  const orb_grid_params gp = {mnMinX, mnMinY, mnMaxX, mnMaxY, mfGridElementWidthInv, mfGridElementHeightInv};
  CheckB200(ex->Handle(), orb_assign_features_to_grid(ex->Handle(), &gp, 0), "orb_assign_features_to_grid");
}

// ORBmatcher::SearchByProjection(CurrentFrame, LastFrame, th, bMono) (src/ORBmatcher.cc:1521-1733) after the projection:
// one query per last-frame keypoint. match[i2] = last-frame keypoint index or -1. Returns nmatches.
inline int SearchByProjectionB200(ORBextractor* ex, const std::vector<orb_proj_query>& queries, const cv::Mat& queryDescriptors, float th,
                                  bool bMono, float tlc_z, float mb, float mbf, bool checkOrientation, std::vector<int>& match) {
  orb_handle* h = ex->Handle();
  const int nq = (int)queries.size();
  match.assign(orb_keypoint_capacity(h), -1);
  int nmatches = 0;
  if (nq == 0) return 0;
  CheckB200(h, orb_search_by_projection(h, queries.data(), queryDescriptors.data, &nq, nq, th, bMono ? 1 : 0, &tlc_z, mb, mbf, checkOrientation ? 1 : 0,
                                        match.data(), &nmatches, 0), "orb_search_by_projection");
  return nmatches;
}

// ORBmatcher::SearchByProjection(F, vpMapPoints, th, bFarPoints, thFarPoints) (src/ORBmatcher.cc:42-209), the local-map search.
// locked[i2] != 0: F.mvpMapPoints[i2] already holds a map point with observations. match[i2] = index into the queries or -1.
inline int SearchLocalPointsB200(ORBextractor* ex, const std::vector<orb_track_query>& queries, const cv::Mat& queryDescriptors,
                                 const std::vector<unsigned char>& locked, float th, float nnratio, std::vector<int>& match) {
  orb_handle* h = ex->Handle();
  const int nq = (int)queries.size(), kcap = orb_keypoint_capacity(h);
  match.assign(kcap, -1);
  int nmatches = 0;
  if (nq == 0) return 0;
  std::vector<unsigned char> lk(kcap, 0);
  for (size_t i = 0; i < locked.size() && i < lk.size(); ++i) lk[i] = locked[i];
  CheckB200(h, orb_search_local_points(h, queries.data(), queryDescriptors.data, &nq, nq, lk.data(), th, nnratio, match.data(), &nmatches, 0),
            "orb_search_local_points");
  return nmatches;
}

// ORBmatcher::SearchForInitialization(F1, F2, vbPrevMatched, vnMatches12, windowSize) (src/ORBmatcher.cc:603-700, called by
// Tracking::MonocularInitialization with windowSize 100): F2 = the frame whose keypoints are resident on `exF2` with the grid built
// (AssignFeaturesToGridB200), F1 = the initial frame given by its mvKeysUn and mDescriptors. vbPrevMatched is updated in place like
// in the reference. Returns nmatches.
inline int SearchForInitializationB200(ORBextractor* exF2, const std::vector<cv::KeyPoint>& vKeysUn1, const cv::Mat& descriptors1,
                                       std::vector<cv::Point2f>& vbPrevMatched, std::vector<int>& vnMatches12, int windowSize, float nnratio,
                                       bool checkOrientation) {
  orb_handle* h = exF2->Handle();
  const int n1 = (int)vKeysUn1.size();
  vnMatches12.assign(n1, -1);
  if (n1 == 0) return 0;
  std::vector<orb_init_query> q(n1);
  for (int i = 0; i < n1; ++i) {
    q[i].x = vbPrevMatched[i].x; q[i].y = vbPrevMatched[i].y;
    q[i].angle = vKeysUn1[i].angle; q[i].octave = vKeysUn1[i].octave;
  }
  std::vector<float> prev((size_t)n1 * 2);
  int nmatches = 0;
  CheckB200(h, orb_search_for_initialization(h, q.data(), descriptors1.data, &n1, n1, windowSize, nnratio, checkOrientation ? 1 : 0,
                                             vnMatches12.data(), prev.data(), &nmatches, 0), "orb_search_for_initialization");
  for (int i = 0; i < n1; ++i) { vbPrevMatched[i].x = prev[2 * i]; vbPrevMatched[i].y = prev[2 * i + 1]; }
  return nmatches;
}

// Frame::ComputeBoW (src/Frame.cc:822-827): mpORBvocabulary->transform(vCurrentDesc, mBowVec, mFeatVec, 4) on the resident descriptors.
template <class BowVector, class FeatureVector>
inline void ComputeBoWB200(ORBextractor* ex, const orb_vocab* voc, BowVector& mBowVec, FeatureVector& mFeatVec, int levelsup = 4) {
  orb_handle* h = ex->Handle();
  const int kcap = orb_keypoint_capacity(h);
  int32_t nb = 0, nn = 0;
  std::vector<uint32_t> word(kcap), node(kcap), feat(kcap);
  std::vector<double> val(kcap);
  std::vector<int32_t> off(kcap + 1);
  orb_bow_out out = {&nb, word.data(), val.data(), &nn, node.data(), off.data(), feat.data(), nullptr, nullptr};
  CheckB200(h, orb_compute_bow(h, voc, levelsup, &out, 0), "orb_compute_bow");
  mBowVec.clear();
  mFeatVec.clear();
  for (int i = 0; i < nb; ++i) mBowVec.insert(mBowVec.end(), typename BowVector::value_type(word[i], val[i]));   // already in map order
  for (int j = 0; j < nn; ++j)
    mFeatVec.insert(mFeatVec.end(), typename FeatureVector::value_type(
                                        node[j], typename FeatureVector::mapped_type(feat.begin() + off[j], feat.begin() + off[j + 1])));
}

// Frame::ComputeStereoFishEyeMatches (src/Frame.cc:1222-1273) on the resident results of the two extractors' last calls: knnMatch(k = 2)
// of the lapping-area descriptors + Lowe's ratio + KannalaBrandt8::TriangulateMatches + depth > 0.0001f. Vec3 is any type with
// operator[] over three floats (Eigen::Vector3f). rig: see orb_kb8_rig (mpCamera / mpCamera2 parameters, mRlr, mtlr). Returns nMatches.
template <class Vec3>
inline int ComputeStereoFishEyeMatchesB200(ORBextractor* exLeft, ORBextractor* exRight, const orb_kb8_rig& rig, int Nleft, int Nright,
                                           std::vector<int>& mvLeftToRightMatch, std::vector<int>& mvRightToLeftMatch, std::vector<float>& mvDepth,
                                           std::vector<Vec3>& mvStereo3Dpoints) {
  orb_handle *hL = exLeft->Handle(), *hR = exRight->Handle();
  const int cap = std::max(orb_keypoint_capacity(hL), orb_keypoint_capacity(hR));
  std::vector<int32_t> l2r(cap, -1), r2l(cap, -1);
  std::vector<float> depth(cap, -1.0f), p3d((size_t)cap * 3, 0.0f);
  CheckB200(hL, orb_stereo_fisheye_match_batch(hL, hR, nullptr, nullptr, nullptr, 0, ORB_NO_OUTPUT | ORB_ASYNC), "orb_stereo_fisheye_match_batch");
  CheckB200(hL, orb_stereo_fisheye_triangulate_batch(hL, hR, &rig, l2r.data(), r2l.data(), depth.data(), p3d.data(), nullptr, cap, 0),
            "orb_stereo_fisheye_triangulate_batch");
  mvLeftToRightMatch.assign(l2r.begin(), l2r.begin() + Nleft);
  mvRightToLeftMatch.assign(r2l.begin(), r2l.begin() + Nright);
  mvDepth.assign(depth.begin(), depth.begin() + Nleft);
  mvStereo3Dpoints.resize(Nleft);
  int nMatches = 0;
  for (int i = 0; i < Nleft; ++i) {
    for (int k = 0; k < 3; ++k) mvStereo3Dpoints[i][k] = p3d[(size_t)3 * i + k];
    nMatches += l2r[i] >= 0;
  }
  return nMatches;
}

}  // namespace ORB_SLAM3

#endif  // FRAME_B200_H
