// liborb_b200.so - the LocalMapping-side consumers of the Hamming primitives (widening beyond SURVEY.md 8):
//   orb_load_frames                  keyframes of the host's map become the handle's resident batch (KeyFrame::KeyFrame copies the
//                                    Frame's keypoints, descriptors, mvuRight and grid: reference src/KeyFrame.cc:86-140)
//   ORBmatcher::SearchForTriangulation   reference src/ORBmatcher.cc:821-1042 (single-camera keyframes, Pinhole::epipolarConstrain
//                                        src/CameraModels/Pinhole.cpp:113-139 with the fundamental matrix supplied by the caller)
//   MapPoint::ComputeDistinctiveDescriptors   reference src/MapPoint.cc:367-431
//   ORBmatcher::SearchByBoW(KeyFrame*, KeyFrame*, ...)   reference src/ORBmatcher.cc:702-819 (loop closing / merging)
// (ORBmatcher::Fuse and the KeyFrame overload of SearchByProjection share the window scan of orb_match.cu and live there.)
#include <algorithm>
#include <cstring>

#include "orb_internal.h"
#include "orb_kb8_dev.cuh"

#define MP_TH_LOW 50      // ORBmatcher::TH_LOW (src/ORBmatcher.cc:36)
#define MP_HISTO 30       // ORBmatcher::HISTO_LENGTH (src/ORBmatcher.cc:37)

static __device__ __forceinline__ int mp_hamming256(const uint4 a0, const uint4 a1, const uint4* __restrict__ b) {
  const uint4 b0 = b[0], b1 = b[1];
  return __popc(a0.x ^ b0.x) + __popc(a0.y ^ b0.y) + __popc(a0.z ^ b0.z) + __popc(a0.w ^ b0.w) + __popc(a1.x ^ b1.x) +
         __popc(a1.y ^ b1.y) + __popc(a1.z ^ b1.z) + __popc(a1.w ^ b1.w);
}

// ---- ORBmatcher::SearchForTriangulation ---------------------------------------------------------------------------------------
// One CTA per keyframe pair. The reference walks the two FeatureVectors in step (:872-1013) and, inside a shared vocabulary node,
// lets every keypoint idx1 of pKF1 that has no map point scan the node's keypoints idx2 of pKF2 in order. vbMatched2 is never set
// in this code base (:868 initialises it, nothing writes it), so the scans are independent: one thread per FeatureVector entry of
// pKF1; the node of pKF2 is found by binary search over its sorted node ids. The rotation histogram is a set of counters
// (the bins' contents are only used to clear vMatches12[idx1], and idx1 is unique), ComputeThreeMaxima as in orb_match.cu.
struct SftSet {
  const orb_keypoint* kps;
  const uint8_t* desc;
  const float* uright;       // may be null: every entry -1
  const uint8_t* has_mp;
  const int* n;
  const unsigned int* fv_node;
  const int* fv_off;
  const unsigned int* fv_feat;
  const int* fv_n;
  int cap;
};
struct SftLevels { float sigma2[ORB_MAX_LEVELS], scale[ORB_MAX_LEVELS]; };

__global__ void __launch_bounds__(256) k_search_for_triangulation(SftSet S, const int* __restrict__ kf1, const int* __restrict__ kf2,
                                                                   const float* __restrict__ F12, const float* __restrict__ ep,
                                                                   SftLevels lv, int only_stereo, int coarse, int check_orientation,
                                                                   int* __restrict__ match12, int* __restrict__ nmatches) {
  __shared__ int s_hist[MP_HISTO];
  __shared__ int s_keep[3];
  __shared__ int s_nm;
  const int p = blockIdx.x, tid = threadIdx.x;
  const int a = kf1[p], b = kf2[p], cap = S.cap;
  const int n1 = min(S.n[a], cap);
  const orb_keypoint* kp1 = S.kps + (size_t)a * cap;
  const orb_keypoint* kp2 = S.kps + (size_t)b * cap;
  const uint8_t* d1 = S.desc + (size_t)a * cap * 32;
  const uint8_t* d2 = S.desc + (size_t)b * cap * 32;
  const float* ur1 = S.uright ? S.uright + (size_t)a * cap : nullptr;
  const float* ur2 = S.uright ? S.uright + (size_t)b * cap : nullptr;
  const uint8_t* mp1 = S.has_mp + (size_t)a * cap;
  const uint8_t* mp2 = S.has_mp + (size_t)b * cap;
  const unsigned int* node1 = S.fv_node + (size_t)a * cap;
  const unsigned int* node2 = S.fv_node + (size_t)b * cap;
  const int* off1 = S.fv_off + (size_t)a * (cap + 1);
  const int* off2 = S.fv_off + (size_t)b * (cap + 1);
  const unsigned int* feat1 = S.fv_feat + (size_t)a * cap;
  const unsigned int* feat2 = S.fv_feat + (size_t)b * cap;
  const int nn1 = min(S.fv_n[a], cap), nn2 = min(S.fv_n[b], cap);
  int* m12 = match12 + (size_t)p * cap;
  const float f00 = F12[p * 9 + 0], f01 = F12[p * 9 + 1], f02 = F12[p * 9 + 2], f10 = F12[p * 9 + 3], f11 = F12[p * 9 + 4],
              f12 = F12[p * 9 + 5], f20 = F12[p * 9 + 6], f21 = F12[p * 9 + 7], f22 = F12[p * 9 + 8];
  const float epx = ep[p * 2], epy = ep[p * 2 + 1];
  for (int i = tid; i < cap; i += 256) m12[i] = -1;
  if (tid < MP_HISTO) s_hist[tid] = 0;
  if (tid == 0) s_nm = 0;
  __syncthreads();
  const float factor = 1.0f / MP_HISTO;
  const int total1 = nn1 > 0 ? off1[nn1] : 0;
  int mine = 0;
  for (int t = tid; t < total1; t += 256) {
    // node of entry t: last j with off1[j] <= t
    int lo = 0, hi = nn1 - 1;
    while (lo < hi) {
      const int mid = (lo + hi + 1) >> 1;
      if (off1[mid] <= t) lo = mid; else hi = mid - 1;
    }
    const unsigned int node = node1[lo];
    // the same node in pKF2 (FeatureVector is a std::map: unique ascending ids)
    int l2 = 0, h2 = nn2 - 1, j2 = -1;
    while (l2 <= h2) {
      const int mid = (l2 + h2) >> 1;
      const unsigned int v = node2[mid];
      if (v == node) { j2 = mid; break; }
      if (v < node) l2 = mid + 1; else h2 = mid - 1;
    }
    if (j2 < 0) continue;
    const int idx1 = (int)feat1[t];
    if (idx1 >= n1 || mp1[idx1]) continue;                                   // pMP1 (:880-885)
    const bool stereo1 = ur1 && ur1[idx1] >= 0;                              // !mpCamera2 && mvuRight[idx1] >= 0 (:887)
    if (only_stereo && !stereo1) continue;
    const orb_keypoint k1 = kp1[idx1];
    const uint4* q = reinterpret_cast<const uint4*>(d1 + (size_t)idx1 * 32);
    const uint4 a0 = q[0], a1 = q[1];
    // epipolar line of kp1 in the second image (Pinhole.cpp:126-128): l = x1' F12
    const float la = __fadd_rn(__fadd_rn(__fmul_rn(k1.x, f00), __fmul_rn(k1.y, f10)), f20);
    const float lb = __fadd_rn(__fadd_rn(__fmul_rn(k1.x, f01), __fmul_rn(k1.y, f11)), f21);
    const float lc = __fadd_rn(__fadd_rn(__fmul_rn(k1.x, f02), __fmul_rn(k1.y, f12)), f22);
    const float den = __fadd_rn(__fmul_rn(la, la), __fmul_rn(lb, lb));
    int bestDist = MP_TH_LOW, bestIdx2 = -1;
    for (int u = off2[j2]; u < off2[j2 + 1]; ++u) {
      const int idx2 = (int)feat2[u];
      if (mp2[idx2]) continue;                                               // vbMatched2[idx2] (never set) || pMP2 (:913-916)
      const bool stereo2 = ur2 && ur2[idx2] >= 0;
      if (only_stereo && !stereo2) continue;
      const int dist = mp_hamming256(a0, a1, reinterpret_cast<const uint4*>(d2 + (size_t)idx2 * 32));
      if (dist > MP_TH_LOW || dist > bestDist) continue;                     // :926
      const orb_keypoint k2 = kp2[idx2];
      if (!stereo1 && !stereo2) {                                            // :943-950 (mpCamera2 == NULL here)
        const float distex = __fsub_rn(epx, k2.x), distey = __fsub_rn(epy, k2.y);
        if (__fadd_rn(__fmul_rn(distex, distex), __fmul_rn(distey, distey)) < __fmul_rn(100.f, lv.scale[k2.octave])) continue;
      }
      bool ok = coarse != 0;
      if (!ok) {                                                             // Pinhole::epipolarConstrain (Pinhole.cpp:130-138)
        const float num = __fadd_rn(__fadd_rn(__fmul_rn(la, k2.x), __fmul_rn(lb, k2.y)), lc);
        if (den != 0) {
          const float dsqr = __fdiv_rn(__fmul_rn(num, num), den);
          ok = (double)dsqr < __dmul_rn(3.84, (double)lv.sigma2[k2.octave]);
        }
      }
      if (ok) { bestIdx2 = idx2; bestDist = dist; }
    }
    if (bestIdx2 >= 0) {
      m12[idx1] = bestIdx2;
      ++mine;
      if (check_orientation) {
        float rot = __fsub_rn(k1.angle, kp2[bestIdx2].angle);
        if (rot < 0.0f) rot = __fadd_rn(rot, 360.0f);
        int bin = (int)roundf(__fmul_rn(rot, factor));
        if (bin == MP_HISTO) bin = 0;
        atomicAdd(&s_hist[bin], 1);
      }
    }
  }
  if (mine) atomicAdd(&s_nm, mine);
  __syncthreads();
  if (check_orientation) {
    if (tid == 0) {
      // ComputeThreeMaxima (:1844-1876)
      int ind1 = -1, ind2 = -1, ind3 = -1, max1 = 0, max2 = 0, max3 = 0;
      for (int i = 0; i < MP_HISTO; i++) {
        const int s = s_hist[i];
        if (s > max1) { max3 = max2; max2 = max1; max1 = s; ind3 = ind2; ind2 = ind1; ind1 = i; }
        else if (s > max2) { max3 = max2; max2 = s; ind3 = ind2; ind2 = i; }
        else if (s > max3) { max3 = s; ind3 = i; }
      }
      if ((float)max2 < __fmul_rn(0.1f, (float)max1)) { ind2 = -1; ind3 = -1; }
      else if ((float)max3 < __fmul_rn(0.1f, (float)max1)) { ind3 = -1; }
      s_keep[0] = ind1; s_keep[1] = ind2; s_keep[2] = ind3;
    }
    __syncthreads();
    // every match of a losing bin is taken back (:1025-1031); the bin of a match is recomputed from its two angles
    int drop = 0;
    for (int i = tid; i < n1; i += 256) {
      const int j = m12[i];
      if (j < 0) continue;
      float rot = __fsub_rn(kp1[i].angle, kp2[j].angle);
      if (rot < 0.0f) rot = __fadd_rn(rot, 360.0f);
      int bin = (int)roundf(__fmul_rn(rot, factor));
      if (bin == MP_HISTO) bin = 0;
      if (bin != s_keep[0] && bin != s_keep[1] && bin != s_keep[2]) { m12[i] = -1; ++drop; }
    }
    if (drop) atomicSub(&s_nm, drop);
    __syncthreads();
  }
  if (tid == 0) nmatches[p] = s_nm;
}

// ---- SearchForTriangulation between two-camera keyframes (mpCamera2 != NULL on both: the fisheye rig; :891-903, :935-981) --------
// The keyframes' keypoints are the left ones followed by the right ones (index idx < NLeft: mvKeys[idx], else mvKeysRight[idx - NLeft]),
// mDescriptors / mFeatVec / the map points cover that combined index space. bStereo1 / bStereo2 are false by definition
// (`!pKF->mpCamera2 && ...`), so bOnlyStereo skips everything and the epipole test never runs; the epipolar constraint is
// KannalaBrandt8::epipolarConstrain (src/CameraModels/KannalaBrandt8.cpp:229-236) = TriangulateMatches(...) > 0.0001f with the
// relative pose and the two cameras of the (left / right, left / right) combination: rigs[p][0..3] = ll, lr, rl, rr
// (cam1 = the camera of kp1, cam2 = the camera of kp2, R12 / t12 = Tll / Tlr / Trl / Trr of :846-855).
__global__ void __launch_bounds__(128) k_search_for_triangulation2(SftSet S, const int* __restrict__ nleft, const int* __restrict__ kf1,
                                                                   const int* __restrict__ kf2, const orb_kb8_rig* __restrict__ rigs,
                                                                   SftLevels lv, int only_stereo, int coarse, int check_orientation,
                                                                   int* __restrict__ match12, int* __restrict__ nmatches) {
  __shared__ int s_hist[MP_HISTO];
  __shared__ int s_keep[3];
  __shared__ int s_nm;
  __shared__ orb_kb8_rig s_rig[4];
  const int p = blockIdx.x, tid = threadIdx.x;
  const int a = kf1[p], b = kf2[p], cap = S.cap;
  const int n1 = min(S.n[a], cap), n2 = min(S.n[b], cap);
  const int nl1 = nleft[a], nl2 = nleft[b];
  const orb_keypoint* kp1 = S.kps + (size_t)a * cap;
  const orb_keypoint* kp2 = S.kps + (size_t)b * cap;
  const uint8_t* d1 = S.desc + (size_t)a * cap * 32;
  const uint8_t* d2 = S.desc + (size_t)b * cap * 32;
  const uint8_t* mp1 = S.has_mp + (size_t)a * cap;
  const uint8_t* mp2 = S.has_mp + (size_t)b * cap;
  const unsigned int* node1 = S.fv_node + (size_t)a * cap;
  const unsigned int* node2 = S.fv_node + (size_t)b * cap;
  const int* off1 = S.fv_off + (size_t)a * (cap + 1);
  const int* off2 = S.fv_off + (size_t)b * (cap + 1);
  const unsigned int* feat1 = S.fv_feat + (size_t)a * cap;
  const unsigned int* feat2 = S.fv_feat + (size_t)b * cap;
  const int nn1 = min(S.fv_n[a], cap), nn2 = min(S.fv_n[b], cap);
  int* m12 = match12 + (size_t)p * cap;
  for (int i = tid; i < cap; i += 128) m12[i] = -1;
  for (int i = tid; i < (int)(4 * sizeof(orb_kb8_rig) / 4); i += 128) reinterpret_cast<float*>(s_rig)[i] = reinterpret_cast<const float*>(rigs + (size_t)p * 4)[i];
  if (tid < MP_HISTO) s_hist[tid] = 0;
  if (tid == 0) s_nm = 0;
  __syncthreads();
  const float factor = 1.0f / MP_HISTO;
  const int total1 = (nn1 > 0 && !only_stereo) ? off1[nn1] : 0;               // bOnlyStereo: !bStereo1 always -> continue (:889-890)
  int mine = 0;
  for (int t = tid; t < total1; t += 128) {
    int lo = 0, hi = nn1 - 1;
    while (lo < hi) {
      const int mid = (lo + hi + 1) >> 1;
      if (off1[mid] <= t) lo = mid; else hi = mid - 1;
    }
    const unsigned int node = node1[lo];
    int l2 = 0, h2 = nn2 - 1, j2 = -1;
    while (l2 <= h2) {
      const int mid = (l2 + h2) >> 1;
      const unsigned int v = node2[mid];
      if (v == node) { j2 = mid; break; }
      if (v < node) l2 = mid + 1; else h2 = mid - 1;
    }
    if (j2 < 0) continue;
    const int idx1 = (int)feat1[t];
    if (idx1 >= n1 || mp1[idx1]) continue;
    const orb_keypoint k1 = kp1[idx1];
    const int right1 = idx1 >= nl1 ? 1 : 0;                                   // bRight1 (:898-899)
    const uint4* q = reinterpret_cast<const uint4*>(d1 + (size_t)idx1 * 32);
    const uint4 a0 = q[0], a1 = q[1];
    int bestDist = MP_TH_LOW, bestIdx2 = -1;
    for (int u = off2[j2]; u < off2[j2 + 1]; ++u) {
      const int idx2 = (int)feat2[u];
      if (idx2 >= n2 || mp2[idx2]) continue;
      const int dist = mp_hamming256(a0, a1, reinterpret_cast<const uint4*>(d2 + (size_t)idx2 * 32));
      if (dist > MP_TH_LOW || dist > bestDist) continue;                      // :926
      bool ok = coarse != 0;
      if (!ok) {
        const orb_keypoint k2 = kp2[idx2];
        const orb_kb8_rig& r = s_rig[2 * right1 + (idx2 >= nl2 ? 1 : 0)];     // :952-981
        float X[3];
        ok = kb8_triangulate_p(r.cam1, r.precision1, r.cam2, r.precision2, r.R12, r.t12, k1.x, k1.y, k2.x, k2.y, lv.sigma2[k1.octave],
                               lv.sigma2[k2.octave], X) > 0.0001f;
      }
      if (ok) { bestIdx2 = idx2; bestDist = dist; }
    }
    if (bestIdx2 >= 0) {
      m12[idx1] = bestIdx2;
      ++mine;
      if (check_orientation) {
        float rot = __fsub_rn(k1.angle, kp2[bestIdx2].angle);
        if (rot < 0.0f) rot = __fadd_rn(rot, 360.0f);
        int bin = (int)roundf(__fmul_rn(rot, factor));
        if (bin == MP_HISTO) bin = 0;
        atomicAdd(&s_hist[bin], 1);
      }
    }
  }
  if (mine) atomicAdd(&s_nm, mine);
  __syncthreads();
  if (check_orientation) {
    if (tid == 0) {
      int ind1 = -1, ind2 = -1, ind3 = -1, max1 = 0, max2 = 0, max3 = 0;      // ComputeThreeMaxima (:1844-1876)
      for (int i = 0; i < MP_HISTO; i++) {
        const int s = s_hist[i];
        if (s > max1) { max3 = max2; max2 = max1; max1 = s; ind3 = ind2; ind2 = ind1; ind1 = i; }
        else if (s > max2) { max3 = max2; max2 = s; ind3 = ind2; ind2 = i; }
        else if (s > max3) { max3 = s; ind3 = i; }
      }
      if ((float)max2 < __fmul_rn(0.1f, (float)max1)) { ind2 = -1; ind3 = -1; }
      else if ((float)max3 < __fmul_rn(0.1f, (float)max1)) { ind3 = -1; }
      s_keep[0] = ind1; s_keep[1] = ind2; s_keep[2] = ind3;
    }
    __syncthreads();
    int drop = 0;
    for (int i = tid; i < n1; i += 128) {
      const int j = m12[i];
      if (j < 0) continue;
      float rot = __fsub_rn(kp1[i].angle, kp2[j].angle);
      if (rot < 0.0f) rot = __fadd_rn(rot, 360.0f);
      int bin = (int)roundf(__fmul_rn(rot, factor));
      if (bin == MP_HISTO) bin = 0;
      if (bin != s_keep[0] && bin != s_keep[1] && bin != s_keep[2]) { m12[i] = -1; ++drop; }
    }
    if (drop) atomicSub(&s_nm, drop);
    __syncthreads();
  }
  if (tid == 0) nmatches[p] = s_nm;
}

// ---- ORBmatcher::SearchByBoW(KeyFrame *pKF1, KeyFrame *pKF2, vector<MapPoint*> &vpMatches12) (src/ORBmatcher.cc:702-819) ------------------
// One CTA per keyframe pair, one warp per shared vocabulary node (the features of different nodes are disjoint, so the nodes are
// independent); inside a node the keypoints idx1 of pKF1 are visited in order like the reference, the lanes scan the node's
// keypoints of pKF2 that hold a map point and are not matched yet (vbMatched2), best / second best as the two smallest keys
// (distance << 16 | position: the first minimum wins like the strict "<"), bestDist1 < TH_LOW, ratio test, lock, rotation histogram.
#define BK_NONE 0xffffffffu
__global__ void __launch_bounds__(256) k_search_by_bow_kf(SftSet S, const int* __restrict__ kf1, const int* __restrict__ kf2, float nnratio,
                                                           int check_orientation, int* __restrict__ match12, int* __restrict__ nmatches) {
  extern __shared__ __align__(16) unsigned char bk_raw[];
  __shared__ int s_hist[MP_HISTO];
  __shared__ int s_keep[3];
  __shared__ int s_nm;
  const int p = blockIdx.x, tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const int a = kf1[p], b = kf2[p], cap = S.cap;
  int* s_m12 = reinterpret_cast<int*>(bk_raw);
  unsigned char* s_matched2 = bk_raw + 4 * (size_t)cap;
  const int n1 = min(S.n[a], cap), n2 = min(S.n[b], cap);
  const orb_keypoint* kp1 = S.kps + (size_t)a * cap;
  const orb_keypoint* kp2 = S.kps + (size_t)b * cap;
  const uint8_t* d1 = S.desc + (size_t)a * cap * 32;
  const uint8_t* d2 = S.desc + (size_t)b * cap * 32;
  const uint8_t* mp1 = S.has_mp + (size_t)a * cap;
  const uint8_t* mp2 = S.has_mp + (size_t)b * cap;
  const unsigned int* node1 = S.fv_node + (size_t)a * cap;
  const unsigned int* node2 = S.fv_node + (size_t)b * cap;
  const int* off1 = S.fv_off + (size_t)a * (cap + 1);
  const int* off2 = S.fv_off + (size_t)b * (cap + 1);
  const unsigned int* feat1 = S.fv_feat + (size_t)a * cap;
  const unsigned int* feat2 = S.fv_feat + (size_t)b * cap;
  const int nn1 = min(S.fv_n[a], cap), nn2 = min(S.fv_n[b], cap);
  for (int i = tid; i < cap; i += 256) { s_m12[i] = -1; s_matched2[i] = 0; }
  if (tid < MP_HISTO) s_hist[tid] = 0;
  if (tid == 0) s_nm = 0;
  __syncthreads();
  const float factor = 1.0f / MP_HISTO;
  int nm = 0;   // per warp, uniform
  for (int j1 = wid; j1 < nn1; j1 += 8) {
    const unsigned int node = node1[j1];
    int lo = 0, hi = nn2;
    while (lo < hi) { const int mid = (lo + hi) >> 1; if (node2[mid] < node) lo = mid + 1; else hi = mid; }
    if (lo >= nn2 || node2[lo] != node) continue;
    const int f0 = off2[lo], f1 = off2[lo + 1];
    for (int t = off1[j1]; t < off1[j1 + 1]; ++t) {
      const int idx1 = (int)feat1[t];
      if (idx1 >= n1 || !mp1[idx1]) continue;                       // pMP1 missing or bad (:741-743)
      const uint4* q = reinterpret_cast<const uint4*>(d1 + (size_t)idx1 * 32);
      const uint4 a0 = q[0], a1 = q[1];
      unsigned int k0 = BK_NONE, k1 = BK_NONE;
      for (int u = f0 + lane; u < f1; u += 32) {
        const int idx2 = (int)feat2[u];
        if (idx2 >= n2 || !mp2[idx2] || s_matched2[idx2]) continue;  // vbMatched2 || !pMP2 || bad (:760-762)
        const unsigned int d = (unsigned int)mp_hamming256(a0, a1, reinterpret_cast<const uint4*>(d2 + (size_t)idx2 * 32));
        const unsigned int key = (d << 16) | (unsigned int)(u - f0);
        if (key < k0) { k1 = k0; k0 = key; } else if (key < k1) k1 = key;
      }
      const unsigned int m1 = __reduce_min_sync(0xffffffffu, k0);
      if (m1 == BK_NONE) continue;                                  // bestDist1 = 256
      if (k0 == m1) { k0 = k1; k1 = BK_NONE; }
      const unsigned int m2 = __reduce_min_sync(0xffffffffu, k0);
      const int bestDist1 = (int)(m1 >> 16), bestDist2 = m2 == BK_NONE ? 256 : (int)(m2 >> 16);
      __syncwarp();                                                 // the sweep's reads of s_matched2 come before lane 0's write
      if (bestDist1 < MP_TH_LOW && (float)bestDist1 < __fmul_rn(nnratio, (float)bestDist2)) {   // :777-779
        const int bestIdx2 = (int)feat2[f0 + (int)(m1 & 0xffffu)];
        if (lane == 0) {
          s_m12[idx1] = bestIdx2;
          s_matched2[bestIdx2] = 1;
          if (check_orientation) {
            float rot = __fsub_rn(kp1[idx1].angle, kp2[bestIdx2].angle);
            if (rot < 0.0f) rot = __fadd_rn(rot, 360.0f);
            int bin = (int)roundf(__fmul_rn(rot, factor));
            if (bin == MP_HISTO) bin = 0;
            atomicAdd(&s_hist[bin], 1);
          }
        }
        nm++;
      }
      __syncwarp();
    }
  }
  if (lane == 0 && nm) atomicAdd(&s_nm, nm);
  __syncthreads();
  if (check_orientation) {
    if (tid == 0) {
      int ind1 = -1, ind2 = -1, ind3 = -1, max1 = 0, max2 = 0, max3 = 0;
      for (int i = 0; i < MP_HISTO; i++) {
        const int s = s_hist[i];
        if (s > max1) { max3 = max2; max2 = max1; max1 = s; ind3 = ind2; ind2 = ind1; ind1 = i; }
        else if (s > max2) { max3 = max2; max2 = s; ind3 = ind2; ind2 = i; }
        else if (s > max3) { max3 = s; ind3 = i; }
      }
      if ((float)max2 < __fmul_rn(0.1f, (float)max1)) { ind2 = -1; ind3 = -1; }
      else if ((float)max3 < __fmul_rn(0.1f, (float)max1)) { ind3 = -1; }
      s_keep[0] = ind1; s_keep[1] = ind2; s_keep[2] = ind3;
    }
    __syncthreads();
    int drop = 0;
    for (int i = tid; i < n1; i += 256) {
      const int j = s_m12[i];
      if (j < 0) continue;
      float rot = __fsub_rn(kp1[i].angle, kp2[j].angle);
      if (rot < 0.0f) rot = __fadd_rn(rot, 360.0f);
      int bin = (int)roundf(__fmul_rn(rot, factor));
      if (bin == MP_HISTO) bin = 0;
      if (bin != s_keep[0] && bin != s_keep[1] && bin != s_keep[2]) { s_m12[i] = -1; ++drop; }
    }
    if (drop) atomicSub(&s_nm, drop);
    __syncthreads();
  }
  for (int i = tid; i < cap; i += 256) match12[(size_t)p * cap + i] = s_m12[i];
  if (tid == 0) nmatches[p] = s_nm;
}

// ---- MapPoint::ComputeDistinctiveDescriptors -------------------------------------------------------------------------------------
// One warp per map point. Row i of the distance matrix is built into a 257-bin histogram in the warp's shared memory (lanes stride
// over j), the median = the sorted row at index (size_t)(0.5 * (N - 1)) is read off the histogram's prefix sums, the smallest
// median with the lowest i wins (:417-429, strict "<"). The reference stores the distances as float; they are integers <= 256.
#define DD_WARPS 8
__global__ void __launch_bounds__(DD_WARPS * 32) k_distinctive(const uint8_t* __restrict__ desc, const int* __restrict__ off, int npoints,
                                                              int* __restrict__ best_out, int* __restrict__ median_out) {
  __shared__ int s_hist[DD_WARPS][288];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const int p = blockIdx.x * DD_WARPS + wid;
  if (p >= npoints) return;
  const int o = off[p], N = off[p + 1] - o;
  if (N <= 0) {
    if (lane == 0) { best_out[p] = -1; if (median_out) median_out[p] = -1; }
    return;
  }
  int* hist = s_hist[wid];
  const int kth = (int)(0.5 * (double)(N - 1));      // vDists[0.5 * (N - 1)] (:423)
  const uint8_t* D = desc + (size_t)o * 32;
  int bestMedian = 0x7fffffff, bestIdx = 0;
  for (int i = 0; i < N; ++i) {
    for (int k = lane; k < 288; k += 32) hist[k] = 0;
    __syncwarp();
    const uint4* q = reinterpret_cast<const uint4*>(D + (size_t)i * 32);
    const uint4 a0 = q[0], a1 = q[1];
    for (int j = lane; j < N; j += 32) {
      const int d = mp_hamming256(a0, a1, reinterpret_cast<const uint4*>(D + (size_t)j * 32));   // 0 on the diagonal (:408)
      atomicAdd(&hist[d], 1);
    }
    __syncwarp();
    // lane owns bins 9 * lane .. 9 * lane + 8 (257 bins, padded): first bin whose inclusive prefix exceeds kth
    int c[9], tot = 0;
#pragma unroll
    for (int k = 0; k < 9; ++k) { c[k] = hist[9 * lane + k]; tot += c[k]; }
    int incl = tot;
#pragma unroll
    for (int s = 1; s < 32; s <<= 1) {
      const int v = __shfl_up_sync(0xffffffffu, incl, s);
      if (lane >= s) incl += v;
    }
    int run = incl - tot, med = 0x7fffffff;
#pragma unroll
    for (int k = 0; k < 9; ++k) {
      run += c[k];
      if (med == 0x7fffffff && run > kth) med = 9 * lane + k;
    }
    med = __reduce_min_sync(0xffffffffu, med);
    if (med < bestMedian) { bestMedian = med; bestIdx = i; }
    __syncwarp();
  }
  if (lane == 0) { best_out[p] = bestIdx; if (median_out) median_out[p] = bestMedian; }
}

extern "C" {

int orb_load_frames(orb_handle* h, const orb_keypoint* kps, const uint8_t* desc, const float* uright, const int32_t* n, int batch, int cap,
                    int flags) {
  if (!h || !kps || !desc || !n || batch < 1 || cap < 1) return ORB_ERR_INVALID_ARG;
  int st;
  if ((st = orb_sync(h))) return st;                       // an asynchronous extraction still owns the resident buffers
  if ((st = orb_use_device(h))) return st;
  const int kcap = h->g.kcap;
  if (batch > h->g.batch_cap) return orb_set_error(h, ORB_ERR_CAPACITY, "more frames than the handle's max_batch");
  std::vector<int> hn(batch);
  if (flags & ORB_SRC_DEVICE) ORB_CUDA_CHECK(h, cudaMemcpy(hn.data(), n, (size_t)batch * 4, cudaMemcpyDeviceToHost));
  else std::memcpy(hn.data(), n, (size_t)batch * 4);
  for (int f = 0; f < batch; ++f)
    if (hn[f] < 0 || hn[f] > cap || hn[f] > kcap) return orb_set_error(h, ORB_ERR_CAPACITY, "a loaded frame holds more keypoints than cap / orb_keypoint_capacity()");
  if ((st = orb_ensure(h, h->d_kps, (size_t)batch * kcap * sizeof(orb_keypoint)))) return st;
  if ((st = orb_ensure(h, h->d_desc, (size_t)batch * kcap * 32))) return st;
  if ((st = orb_ensure(h, h->d_n, (size_t)batch * sizeof(int)))) return st;
  if (uright && (st = orb_ensure(h, h->d_uright, (size_t)batch * kcap * sizeof(float)))) return st;
  const int rows = std::min(cap, kcap);
  ORB_CUDA_CHECK(h, cudaMemcpy2DAsync(h->d_kps.p, (size_t)kcap * sizeof(orb_keypoint), kps, (size_t)cap * sizeof(orb_keypoint),
                                      (size_t)rows * sizeof(orb_keypoint), batch, cudaMemcpyDefault, h->stream));
  ORB_CUDA_CHECK(h, cudaMemcpy2DAsync(h->d_desc.p, (size_t)kcap * 32, desc, (size_t)cap * 32, (size_t)rows * 32, batch, cudaMemcpyDefault, h->stream));
  if (uright)
    ORB_CUDA_CHECK(h, cudaMemcpy2DAsync(h->d_uright.p, (size_t)kcap * sizeof(float), uright, (size_t)cap * sizeof(float), (size_t)rows * sizeof(float),
                                        batch, cudaMemcpyDefault, h->stream));
  ORB_CUDA_CHECK(h, cudaMemcpyAsync(h->d_n.p, hn.data(), (size_t)batch * 4, cudaMemcpyHostToDevice, h->stream));
  ORB_CUDA_CHECK(h, cudaStreamSynchronize(h->stream));     // hn is a stack-lifetime buffer
  h->cur_batch = batch;
  h->have_batch = true;
  h->have_stereo = uright != nullptr;
  h->have_grid = false; h->have_undist = false; h->have_bow = false; h->have_bow2 = false; h->have_fe = false; h->have_fe_tri = false;
  h->frames_loaded = true;
  return ORB_OK;
}

// device view of a keyframe set and its pair list (host inputs are staged into d_scratch)
static int stage_kf_pairs(orb_handle* h, const orb_kf_set* kfs, const int32_t* kf1, const int32_t* kf2, const float* F12, const float* ep,
                          int npairs, int flags, SftSet* S, const int** d_k1, const int** d_k2, const float** d_F, const float** d_ep) {
  const int cap = kfs->cap, cnt = kfs->count;
  const size_t nk = (size_t)cnt * cap;
  S->kps = kfs->kps; S->desc = kfs->desc; S->uright = kfs->uright; S->has_mp = kfs->has_mp; S->n = kfs->n; S->fv_node = kfs->fv_node;
  S->fv_off = kfs->fv_off; S->fv_feat = kfs->fv_feat; S->fv_n = kfs->fv_n; S->cap = cap;
  *d_k1 = kf1; *d_k2 = kf2; *d_F = F12; *d_ep = ep;
  if (flags & ORB_SRC_DEVICE) return ORB_OK;
  int st;
  for (int p = 0; p < npairs; ++p)
    if (kf1[p] < 0 || kf1[p] >= cnt || kf2[p] < 0 || kf2[p] >= cnt) return orb_set_error(h, ORB_ERR_INVALID_ARG, "pair index outside the keyframe set");
  // one staging buffer: kps | desc | uright | has_mp | n | node | off | feat | nn | kf1 | kf2 | F12 | ep
  const size_t bytes[13] = {nk * sizeof(orb_keypoint), nk * 32, kfs->uright ? nk * 4 : 0, nk, (size_t)cnt * 4, nk * 4,
                            (size_t)cnt * (cap + 1) * 4, nk * 4, (size_t)cnt * 4, (size_t)npairs * 4, (size_t)npairs * 4,
                            F12 ? (size_t)npairs * 36 : 0, ep ? (size_t)npairs * 8 : 0};
  const void* src[13] = {kfs->kps, kfs->desc, kfs->uright, kfs->has_mp, kfs->n, kfs->fv_node, kfs->fv_off, kfs->fv_feat, kfs->fv_n,
                         kf1, kf2, F12, ep};
  size_t o[14];
  o[0] = 0;
  for (int i = 0; i < 13; ++i) o[i + 1] = o[i] + ((bytes[i] + 255) & ~(size_t)255);
  if ((st = orb_ensure(h, h->d_scratch, o[13] + 256))) return st;
  uint8_t* base = h->d_scratch.as<uint8_t>();
  for (int i = 0; i < 13; ++i)
    if (bytes[i]) ORB_CUDA_CHECK(h, cudaMemcpyAsync(base + o[i], src[i], bytes[i], cudaMemcpyHostToDevice, h->stream));
  S->kps = (const orb_keypoint*)(base + o[0]); S->desc = base + o[1]; S->uright = kfs->uright ? (const float*)(base + o[2]) : nullptr;
  S->has_mp = base + o[3]; S->n = (const int*)(base + o[4]); S->fv_node = (const unsigned int*)(base + o[5]); S->fv_off = (const int*)(base + o[6]);
  S->fv_feat = (const unsigned int*)(base + o[7]); S->fv_n = (const int*)(base + o[8]);
  *d_k1 = (const int*)(base + o[9]); *d_k2 = (const int*)(base + o[10]);
  *d_F = F12 ? (const float*)(base + o[11]) : nullptr; *d_ep = ep ? (const float*)(base + o[12]) : nullptr;
  return ORB_OK;
}

static bool kf_set_ok(const orb_kf_set* kfs) {
  return kfs && kfs->kps && kfs->desc && kfs->has_mp && kfs->n && kfs->fv_node && kfs->fv_off && kfs->fv_feat && kfs->fv_n && kfs->count >= 1 &&
         kfs->cap >= 1;
}

int orb_search_for_triangulation(orb_handle* h, const orb_kf_set* kfs, const int32_t* kf1, const int32_t* kf2, const float* F12,
                                 const float* ep, int npairs, int only_stereo, int coarse, int check_orientation, int32_t* match12_out,
                                 int32_t* nmatches_out, int flags) {
  if (!h || !kf_set_ok(kfs) || !kf1 || !kf2 || !F12 || !ep || npairs < 1) return ORB_ERR_INVALID_ARG;
  int st;
  if ((st = orb_use_device(h))) return st;
  const int cap = kfs->cap;
  SftSet S;
  const int *d_k1, *d_k2;
  const float *d_F, *d_ep;
  if ((st = stage_kf_pairs(h, kfs, kf1, kf2, F12, ep, npairs, flags, &S, &d_k1, &d_k2, &d_F, &d_ep))) return st;
  if ((st = orb_ensure(h, h->d_scratch2, (size_t)npairs * cap * 4 + (size_t)npairs * 4))) return st;
  int* d_m = h->d_scratch2.as<int>();
  int* d_nm = d_m + (size_t)npairs * cap;
  SftLevels lv;
  for (int l = 0; l < ORB_MAX_LEVELS; ++l) {
    lv.sigma2[l] = l < (int)h->sigma2.size() ? h->sigma2[l] : 0.f;
    lv.scale[l] = l < (int)h->scale.size() ? h->scale[l] : 0.f;
  }
  k_search_for_triangulation<<<npairs, 256, 0, h->stream>>>(S, d_k1, d_k2, d_F, d_ep, lv, only_stereo, coarse, check_orientation, d_m, d_nm);
  h->launches++;
  ORB_CUDA_CHECK(h, cudaGetLastError());
  if (!(flags & ORB_NO_OUTPUT)) {
    if (match12_out) ORB_CUDA_CHECK(h, cudaMemcpyAsync(match12_out, d_m, (size_t)npairs * cap * 4, cudaMemcpyDefault, h->stream));
    if (nmatches_out) ORB_CUDA_CHECK(h, cudaMemcpyAsync(nmatches_out, d_nm, (size_t)npairs * 4, cudaMemcpyDefault, h->stream));
  }
  if (flags & ORB_ASYNC) return ORB_OK;
  ORB_CUDA_CHECK(h, cudaStreamSynchronize(h->stream));
  return ORB_OK;
}

int orb_search_for_triangulation_fisheye(orb_handle* h, const orb_kf_set* kfs, const int32_t* nleft, const int32_t* kf1, const int32_t* kf2,
                                         const orb_kb8_rig* rigs, int npairs, int only_stereo, int coarse, int check_orientation,
                                         int32_t* match12_out, int32_t* nmatches_out, int flags) {
  if (!h || !kf_set_ok(kfs) || !nleft || !kf1 || !kf2 || !rigs || npairs < 1) return ORB_ERR_INVALID_ARG;
  int st;
  if ((st = orb_use_device(h))) return st;
  const int cap = kfs->cap, cnt = kfs->count;
  SftSet S;
  const int *d_k1, *d_k2;
  const float *d_F, *d_ep;
  if ((st = stage_kf_pairs(h, kfs, kf1, kf2, nullptr, nullptr, npairs, flags, &S, &d_k1, &d_k2, &d_F, &d_ep))) return st;
  // results | nmatches | NLeft per keyframe | four rigs per pair
  const size_t b_m = (size_t)npairs * cap * 4, b_n = (size_t)npairs * 4, b_l = (size_t)cnt * 4, b_r = (size_t)npairs * 4 * sizeof(orb_kb8_rig);
  const size_t o_n = (b_m + 255) & ~(size_t)255, o_l = o_n + ((b_n + 255) & ~(size_t)255), o_r = o_l + ((b_l + 255) & ~(size_t)255);
  if ((st = orb_ensure(h, h->d_scratch2, o_r + b_r + 256))) return st;
  uint8_t* base = h->d_scratch2.as<uint8_t>();
  int* d_m = (int*)base;
  int* d_nm = (int*)(base + o_n);
  const int* d_nl = nleft;
  const orb_kb8_rig* d_rigs = rigs;
  if (!(flags & ORB_SRC_DEVICE)) {
    ORB_CUDA_CHECK(h, cudaMemcpyAsync(base + o_l, nleft, b_l, cudaMemcpyHostToDevice, h->stream));
    ORB_CUDA_CHECK(h, cudaMemcpyAsync(base + o_r, rigs, b_r, cudaMemcpyHostToDevice, h->stream));
    d_nl = (const int*)(base + o_l); d_rigs = (const orb_kb8_rig*)(base + o_r);
  }
  SftLevels lv;
  for (int l = 0; l < ORB_MAX_LEVELS; ++l) {
    lv.sigma2[l] = l < (int)h->sigma2.size() ? h->sigma2[l] : 0.f;
    lv.scale[l] = l < (int)h->scale.size() ? h->scale[l] : 0.f;
  }
  k_search_for_triangulation2<<<npairs, 128, 0, h->stream>>>(S, d_nl, d_k1, d_k2, d_rigs, lv, only_stereo, coarse, check_orientation, d_m, d_nm);
  h->launches++;
  ORB_CUDA_CHECK(h, cudaGetLastError());
  if (!(flags & ORB_NO_OUTPUT)) {
    if (match12_out) ORB_CUDA_CHECK(h, cudaMemcpyAsync(match12_out, d_m, b_m, cudaMemcpyDefault, h->stream));
    if (nmatches_out) ORB_CUDA_CHECK(h, cudaMemcpyAsync(nmatches_out, d_nm, b_n, cudaMemcpyDefault, h->stream));
  }
  if (flags & ORB_ASYNC) return ORB_OK;
  ORB_CUDA_CHECK(h, cudaStreamSynchronize(h->stream));
  return ORB_OK;
}

int orb_search_by_bow_kf(orb_handle* h, const orb_kf_set* kfs, const int32_t* kf1, const int32_t* kf2, int npairs, float nnratio,
                         int check_orientation, int32_t* match12_out, int32_t* nmatches_out, int flags) {
  if (!h || !kf_set_ok(kfs) || !kf1 || !kf2 || npairs < 1) return ORB_ERR_INVALID_ARG;
  int st;
  if ((st = orb_use_device(h))) return st;
  const int cap = kfs->cap;
  const size_t smem = (size_t)cap * 5 + 16;
  if (smem > 200 * 1024) return orb_set_error(h, ORB_ERR_CAPACITY, "too many keypoints per keyframe");
  SftSet S;
  const int *d_k1, *d_k2;
  const float *d_F, *d_ep;
  if ((st = stage_kf_pairs(h, kfs, kf1, kf2, nullptr, nullptr, npairs, flags, &S, &d_k1, &d_k2, &d_F, &d_ep))) return st;
  if ((st = orb_ensure(h, h->d_scratch2, (size_t)npairs * cap * 4 + (size_t)npairs * 4))) return st;
  int* d_m = h->d_scratch2.as<int>();
  int* d_nm = d_m + (size_t)npairs * cap;
  if ((st = orb_raise_dyn_smem(h, (const void*)k_search_by_bow_kf, smem))) return st;
  k_search_by_bow_kf<<<npairs, 256, smem, h->stream>>>(S, d_k1, d_k2, nnratio, check_orientation, d_m, d_nm);
  h->launches++;
  ORB_CUDA_CHECK(h, cudaGetLastError());
  if (!(flags & ORB_NO_OUTPUT)) {
    if (match12_out) ORB_CUDA_CHECK(h, cudaMemcpyAsync(match12_out, d_m, (size_t)npairs * cap * 4, cudaMemcpyDefault, h->stream));
    if (nmatches_out) ORB_CUDA_CHECK(h, cudaMemcpyAsync(nmatches_out, d_nm, (size_t)npairs * 4, cudaMemcpyDefault, h->stream));
  }
  if (flags & ORB_ASYNC) return ORB_OK;
  ORB_CUDA_CHECK(h, cudaStreamSynchronize(h->stream));
  return ORB_OK;
}

int orb_distinctive_descriptors(orb_handle* h, const uint8_t* desc, const int32_t* off, int npoints, int32_t* best_out, int32_t* median_out,
                                int flags) {
  if (!h || !desc || !off || npoints < 1 || !best_out) return ORB_ERR_INVALID_ARG;
  int st;
  if ((st = orb_use_device(h))) return st;
  const uint8_t* d_desc = desc;
  const int* d_off = off;
  if (!(flags & ORB_SRC_DEVICE)) {
    const int total = off[npoints];
    if (total < 0 || off[0] != 0) return orb_set_error(h, ORB_ERR_INVALID_ARG, "descriptor offsets must start at 0 and ascend");
    for (int p = 0; p < npoints; ++p)
      if (off[p + 1] < off[p]) return orb_set_error(h, ORB_ERR_INVALID_ARG, "descriptor offsets must start at 0 and ascend");
    const size_t b_d = (size_t)std::max(total, 1) * 32, b_o = (size_t)(npoints + 1) * 4;
    const size_t o_o = (b_d + 255) & ~(size_t)255;
    if ((st = orb_ensure(h, h->d_scratch, o_o + b_o))) return st;
    uint8_t* base = h->d_scratch.as<uint8_t>();
    if (total) ORB_CUDA_CHECK(h, cudaMemcpyAsync(base, desc, (size_t)total * 32, cudaMemcpyHostToDevice, h->stream));
    ORB_CUDA_CHECK(h, cudaMemcpyAsync(base + o_o, off, b_o, cudaMemcpyHostToDevice, h->stream));
    d_desc = base; d_off = (const int*)(base + o_o);
  }
  if ((st = orb_ensure(h, h->d_scratch2, (size_t)npoints * 8))) return st;
  int* d_best = h->d_scratch2.as<int>();
  int* d_med = d_best + npoints;
  k_distinctive<<<(npoints + DD_WARPS - 1) / DD_WARPS, DD_WARPS * 32, 0, h->stream>>>(d_desc, d_off, npoints, d_best, d_med);
  h->launches++;
  ORB_CUDA_CHECK(h, cudaGetLastError());
  if (!(flags & ORB_NO_OUTPUT)) {
    ORB_CUDA_CHECK(h, cudaMemcpyAsync(best_out, d_best, (size_t)npoints * 4, cudaMemcpyDefault, h->stream));
    if (median_out) ORB_CUDA_CHECK(h, cudaMemcpyAsync(median_out, d_med, (size_t)npoints * 4, cudaMemcpyDefault, h->stream));
  }
  if (flags & ORB_ASYNC) return ORB_OK;
  ORB_CUDA_CHECK(h, cudaStreamSynchronize(h->stream));
  return ORB_OK;
}

}  // extern "C"
