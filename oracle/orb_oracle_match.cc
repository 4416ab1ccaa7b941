// TEST INFRASTRUCTURE ONLY (oracle). CPU restatement of the windowed matcher:
//   Frame::AssignFeaturesToGrid / PosInGrid     src/Frame.cc:501-528, 809-820
//   Frame::GetFeaturesInArea                     src/Frame.cc:742-807
//   ORBmatcher::SearchByProjection(Frame &CurrentFrame, const Frame &LastFrame, th, bMono)
//                                                src/ORBmatcher.cc:1521-1733 (Nleft == -1: rectified stereo / mono),
//                                                from the projected point on (the projection itself is host glue)
//   ORBmatcher::ComputeThreeMaxima               src/ORBmatcher.cc:1844-1876
// Pinned against the reference's own code (oracle/_ref/libmorb_ref_match.so, built from /root/reference by line
// range) in tests/test_oracle_match.py. Plain sequential code in the reference's order of operations.
#include <cmath>
#include <cstdint>
#include <cstring>
#include <vector>

#include "orb_oracle.h"

namespace {

struct KeyPoint28 { float x, y, size, angle, response; int octave, class_id; };   // cv::KeyPoint
const int GRID_COLS = 64, GRID_ROWS = 48;                                         // include/Frame.h:44-45
const int TH_HIGH = 100, HISTO_LENGTH = 30;                                       // src/ORBmatcher.cc:34-36

struct Grid {
  std::vector<int> cell[GRID_COLS][GRID_ROWS];
};

// gp = {mnMinX, mnMinY, mnMaxX, mnMaxY, mfGridElementWidthInv, mfGridElementHeightInv}
void build_grid(const KeyPoint28* kp, int n, const float* gp, Grid& g) {
  for (int i = 0; i < n; ++i) {
    // PosInGrid (:809-820): round() of a float expression, half away from zero
    const int px = (int)std::round((kp[i].x - gp[0]) * gp[4]);
    const int py = (int)std::round((kp[i].y - gp[1]) * gp[5]);
    if (px < 0 || px >= GRID_COLS || py < 0 || py >= GRID_ROWS) continue;
    g.cell[px][py].push_back(i);
  }
}

// GetFeaturesInArea (:742-807), bRight = false
void features_in_area(const Grid& g, const KeyPoint28* kp, const float* gp, float x, float y, float r, int minLevel, int maxLevel,
                      std::vector<int>& out) {
  out.clear();
  const int nMinCellX = std::max(0, (int)std::floor((x - gp[0] - r) * gp[4]));
  if (nMinCellX >= GRID_COLS) return;
  const int nMaxCellX = std::min(GRID_COLS - 1, (int)std::ceil((x - gp[0] + r) * gp[4]));
  if (nMaxCellX < 0) return;
  const int nMinCellY = std::max(0, (int)std::floor((y - gp[1] - r) * gp[5]));
  if (nMinCellY >= GRID_ROWS) return;
  const int nMaxCellY = std::min(GRID_ROWS - 1, (int)std::ceil((y - gp[1] + r) * gp[5]));
  if (nMaxCellY < 0) return;
  const bool bCheckLevels = (minLevel > 0) || (maxLevel >= 0);
  for (int ix = nMinCellX; ix <= nMaxCellX; ix++)
    for (int iy = nMinCellY; iy <= nMaxCellY; iy++) {
      const std::vector<int>& c = g.cell[ix][iy];
      for (size_t j = 0; j < c.size(); ++j) {
        const KeyPoint28& k = kp[c[j]];
        if (bCheckLevels) {
          if (k.octave < minLevel) continue;
          if (maxLevel >= 0 && k.octave > maxLevel) continue;
        }
        const float distx = k.x - x, disty = k.y - y;
        if (std::fabs(distx) < r && std::fabs(disty) < r) out.push_back(c[j]);
      }
    }
}

int hamming256(const uint8_t* a, const uint8_t* b) {
  int d = 0;
  for (int i = 0; i < 8; ++i) {
    uint32_t x, y;
    std::memcpy(&x, a + 4 * i, 4);
    std::memcpy(&y, b + 4 * i, 4);
    d += __builtin_popcount(x ^ y);
  }
  return d;
}

struct Query { float u, v, z, angle; int octave, flags; };

}  // namespace

extern "C" {

int oro_assign_grid(const void* kps, int n, const float* gp, int* cell_off, int* idx) {
  Grid g;
  build_grid((const KeyPoint28*)kps, n, gp, g);
  int o = 0;
  for (int ix = 0; ix < GRID_COLS; ++ix)
    for (int iy = 0; iy < GRID_ROWS; ++iy) {
      cell_off[ix * GRID_ROWS + iy] = o;
      for (size_t j = 0; j < g.cell[ix][iy].size(); ++j) idx[o++] = g.cell[ix][iy][j];
    }
  cell_off[GRID_COLS * GRID_ROWS] = o;
  return o;
}

int oro_features_in_area(const void* kps, int n, const float* gp, float x, float y, float r, int minLevel, int maxLevel, int* out,
                         int cap) {
  Grid g;
  build_grid((const KeyPoint28*)kps, n, gp, g);
  std::vector<int> v;
  features_in_area(g, (const KeyPoint28*)kps, gp, x, y, r, minLevel, maxLevel, v);
  if ((int)v.size() > cap) return -2;
  for (size_t i = 0; i < v.size(); ++i) out[i] = v[i];
  return (int)v.size();
}

// src/ORBmatcher.cc:1521-1733 for Nleft == -1, from "x3Dc" on: query i = last-frame keypoint i whose map point
// projects to (u, v) with depth z (flags bit 0: map point present and not an outlier; bit 1: Observations() > 0).
int oro_search_by_projection(const void* kpsC_, const uint8_t* descC, const float* uRightC, int nC, const float* scale, int nlevels,
                             const float* gp, float mb, float mbf, const void* q_, const uint8_t* qdesc, int nq, float th, int bMono,
                             float tlc_z, int check_orientation, int* match_out) {
  const KeyPoint28* kpC = (const KeyPoint28*)kpsC_;
  const Query* q = (const Query*)q_;
  (void)nlevels;
  Grid g;
  build_grid(kpC, nC, gp, g);
  int nmatches = 0;
  std::vector<int> rotHist[HISTO_LENGTH];
  const float factor = 1.0f / HISTO_LENGTH;
  const bool bForward = tlc_z > mb && !bMono;        // :1537
  const bool bBackward = -tlc_z > mb && !bMono;      // :1538
  std::vector<int> assigned(nC, -1);                 // CurrentFrame.mvpMapPoints as query indices
  std::vector<int> cand;
  for (int i = 0; i < nq; ++i) {
    if (!(q[i].flags & 1)) continue;                 // :1541-1543
    const float invzc = (float)(1.0 / (double)q[i].z);  // :1550: const float invzc = 1.0 / x3Dc(2)
    if (invzc < 0) continue;
    const float u = q[i].u, v = q[i].v;
    if (u < gp[0] || u > gp[2]) continue;            // :1556-1559
    if (v < gp[1] || v > gp[3]) continue;
    const int nLastOctave = q[i].octave;
    const float radius = th * scale[nLastOctave];    // :1567
    if (bForward) features_in_area(g, kpC, gp, u, v, radius, nLastOctave, -1, cand);
    else if (bBackward) features_in_area(g, kpC, gp, u, v, radius, 0, nLastOctave, cand);
    else features_in_area(g, kpC, gp, u, v, radius, nLastOctave - 1, nLastOctave + 1, cand);
    if (cand.empty()) continue;
    int bestDist = 256, bestIdx2 = -1;
    for (size_t c = 0; c < cand.size(); ++c) {
      const int i2 = cand[c];
      if (assigned[i2] >= 0 && (q[assigned[i2]].flags & 2)) continue;   // :1592-1593 Observations() > 0
      if (uRightC[i2] > 0) {                                            // :1595-1599 (Nleft == -1)
        const float ur = u - mbf * invzc;
        const float er = std::fabs(ur - uRightC[i2]);
        if (er > radius) continue;
      }
      const int dist = hamming256(qdesc + 32 * (size_t)i, descC + 32 * (size_t)i2);
      if (dist < bestDist) { bestDist = dist; bestIdx2 = i2; }
    }
    if (bestDist <= TH_HIGH) {
      assigned[bestIdx2] = i;
      nmatches++;
      if (check_orientation) {
        float rot = q[i].angle - kpC[bestIdx2].angle;
        if (rot < 0.0) rot += 360.0f;
        int bin = (int)std::round(rot * factor);
        if (bin == HISTO_LENGTH) bin = 0;
        rotHist[bin].push_back(bestIdx2);
      }
    }
  }
  if (check_orientation) {
    // ComputeThreeMaxima (:1844-1876)
    int ind1 = -1, ind2 = -1, ind3 = -1, max1 = 0, max2 = 0, max3 = 0;
    for (int i = 0; i < HISTO_LENGTH; i++) {
      const int s = (int)rotHist[i].size();
      if (s > max1) { max3 = max2; max2 = max1; max1 = s; ind3 = ind2; ind2 = ind1; ind1 = i; }
      else if (s > max2) { max3 = max2; max2 = s; ind3 = ind2; ind2 = i; }
      else if (s > max3) { max3 = s; ind3 = i; }
    }
    if (max2 < 0.1f * (float)max1) { ind2 = -1; ind3 = -1; }
    else if (max3 < 0.1f * (float)max1) { ind3 = -1; }
    for (int i = 0; i < HISTO_LENGTH; i++)
      if (i != ind1 && i != ind2 && i != ind3)
        for (size_t j = 0; j < rotHist[i].size(); j++) { assigned[rotHist[i][j]] = -1; nmatches--; }
  }
  for (int i = 0; i < nC; ++i) match_out[i] = assigned[i];
  return nmatches;
}

// ORBmatcher::SearchByProjection(Frame &F, const vector<MapPoint*> &vpMapPoints, th, bFarPoints, thFarPoints)
// (src/ORBmatcher.cc:42-209) for Nleft == -1. q = {mTrackProjX, mTrackProjY, mTrackProjXR, mTrackViewCos, mnTrackScaleLevel,
// flags}: bit 0 = mbTrackInView && !isBad() && !(bFarPoints && mTrackDepth > thFarPoints), bit 1 = Observations() > 0.
int oro_search_local_points(const void* kpsC_, const uint8_t* descC, const float* uRightC, const uint8_t* locked0, int nC,
                            const float* scale, int nlevels, const float* gp, const void* q_, const uint8_t* qdesc, int nq, float th,
                            float nnratio, int* match_out) {
  struct TQ { float px, py, pxr, vcos; int level, flags; };
  const KeyPoint28* kpC = (const KeyPoint28*)kpsC_;
  const TQ* q = (const TQ*)q_;
  (void)nlevels;
  Grid g;
  build_grid(kpC, nC, gp, g);
  int nmatches = 0;
  const bool bFactor = th != 1.0;                       // :48
  std::vector<int> assigned(nC, -1);
  std::vector<uint8_t> locked(locked0, locked0 + nC);   // keypoint holds a map point with Observations() > 0
  std::vector<int> cand;
  for (int i = 0; i < nq; ++i) {
    if (!(q[i].flags & 1)) continue;                    // :52-56
    const int nPredictedLevel = q[i].level;
    float r = q[i].vcos > 0.998 ? 2.5f : 4.0f;          // RadiusByViewingCos (:211-216): float > double literal
    if (bFactor) r *= th;
    features_in_area(g, kpC, gp, q[i].px, q[i].py, r * scale[nPredictedLevel], nPredictedLevel - 1, nPredictedLevel, cand);
    if (cand.empty()) continue;
    int bestDist = 256, bestLevel = -1, bestDist2 = 256, bestLevel2 = -1, bestIdx = -1;
    for (size_t c = 0; c < cand.size(); ++c) {
      const int idx = cand[c];
      if (locked[idx]) continue;                        // :86-87
      if (uRightC[idx] > 0) {                           // :89-92
        const float er = std::fabs(q[i].pxr - uRightC[idx]);
        if (er > r * scale[nPredictedLevel]) continue;
      }
      const int dist = hamming256(qdesc + 32 * (size_t)i, descC + 32 * (size_t)idx);
      if (dist < bestDist) {
        bestDist2 = bestDist; bestDist = dist; bestLevel2 = bestLevel; bestLevel = kpC[idx].octave; bestIdx = idx;
      } else if (dist < bestDist2) {
        bestLevel2 = kpC[idx].octave; bestDist2 = dist;
      }
    }
    if (bestDist <= TH_HIGH) {
      if (bestLevel == bestLevel2 && bestDist > nnratio * bestDist2) continue;   // float * int -> float compare (:123)
      if (bestLevel != bestLevel2 || bestDist <= nnratio * bestDist2) {
        assigned[bestIdx] = i;
        locked[bestIdx] = (q[i].flags & 2) ? 1 : 0;
        nmatches++;
      }
    }
  }
  for (int i = 0; i < nC; ++i) match_out[i] = assigned[i];
  return nmatches;
}

// ComputeThreeMaxima (src/ORBmatcher.cc:1844-1876) on bin counts
static void three_maxima(const int* hist, int& ind1, int& ind2, int& ind3) {
  int max1 = 0, max2 = 0, max3 = 0;
  ind1 = ind2 = ind3 = -1;
  for (int i = 0; i < HISTO_LENGTH; i++) {
    const int s = hist[i];
    if (s > max1) { max3 = max2; max2 = max1; max1 = s; ind3 = ind2; ind2 = ind1; ind1 = i; }
    else if (s > max2) { max3 = max2; max2 = s; ind3 = ind2; ind2 = i; }
    else if (s > max3) { max3 = s; ind3 = i; }
  }
  if (max2 < 0.1f * (float)max1) { ind2 = -1; ind3 = -1; }
  else if (max3 < 0.1f * (float)max1) { ind3 = -1; }
}

// ORBmatcher::SearchByBoW(KeyFrame *pKF, Frame &F, vector<MapPoint*> &vpMapPointMatches) (src/ORBmatcher.cc:218-395), single
// camera. The nodes the two feature vectors share are independent (a frame keypoint belongs to one node), inside a node the
// keyframe keypoints are taken in list order and each one takes the best frame keypoint of the node that is still free:
// TH_LOW, ratio against the second best, rotation histogram over all nodes at the end.
int oro_search_by_bow(const uint8_t* descKF, const float* angleKF, const uint8_t* kf_flags, int nKF, const uint32_t* kf_node, const int* kf_off,
                      const uint32_t* kf_feat, int kf_nn, const uint8_t* descF, const float* angleF, int nF, const uint32_t* f_node,
                      const int* f_off, const uint32_t* f_feat, int f_nn, float nnratio, int check_orientation, int* match_out) {
  (void)nKF;
  const int TH_LOW = 50;
  std::vector<int> match(nF, -1);
  std::vector<std::pair<int, int>> recs;   // (frame keypoint, bin)
  int hist[HISTO_LENGTH] = {0};
  int nmatches = 0;
  const float factor = 1.0f / HISTO_LENGTH;
  int a = 0, b = 0;
  while (a < kf_nn && b < f_nn) {
    if (kf_node[a] < f_node[b]) { ++a; continue; }        // lower_bound steps (:383-386) visit the same pairs
    if (kf_node[a] > f_node[b]) { ++b; continue; }
    for (int t = kf_off[a]; t < kf_off[a + 1]; ++t) {
      const int iKF = (int)kf_feat[t];
      if (!kf_flags[iKF]) continue;                        // :246-250
      int bestDist1 = 256, bestIdxF = -1, bestDist2 = 256;
      for (int u = f_off[b]; u < f_off[b + 1]; ++u) {
        const int iF = (int)f_feat[u];
        if (match[iF] >= 0) continue;                      // :266
        const int dist = hamming256(descKF + 32 * (size_t)iKF, descF + 32 * (size_t)iF);
        if (dist < bestDist1) { bestDist2 = bestDist1; bestDist1 = dist; bestIdxF = iF; }
        else if (dist < bestDist2) bestDist2 = dist;
      }
      if (bestDist1 <= TH_LOW && (float)bestDist1 < nnratio * (float)bestDist2) {   // :305-307
        match[bestIdxF] = iKF;
        if (check_orientation) {
          float rot = angleKF[iKF] - angleF[bestIdxF];
          if (rot < 0.0) rot += 360.0f;
          int bin = (int)std::round(rot * factor);
          if (bin == HISTO_LENGTH) bin = 0;
          recs.push_back(std::make_pair(bestIdxF, bin));
          hist[bin]++;
        }
        nmatches++;
      }
    }
    ++a; ++b;
  }
  if (check_orientation) {
    int ind1, ind2, ind3;
    three_maxima(hist, ind1, ind2, ind3);
    for (size_t r = 0; r < recs.size(); ++r)
      if (recs[r].second != ind1 && recs[r].second != ind2 && recs[r].second != ind3) { match[recs[r].first] = -1; nmatches--; }
  }
  for (int i = 0; i < nF; ++i) match_out[i] = match[i];
  return nmatches;
}

}  // extern "C"
