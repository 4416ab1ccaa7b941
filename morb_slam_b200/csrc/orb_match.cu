// liborb_b200.so - windowed matcher on the device-resident results of the extractor:
//   Frame::AssignFeaturesToGrid / PosInGrid          reference src/Frame.cc:501-528, 809-820
//   Frame::GetFeaturesInArea                          reference src/Frame.cc:742-807
//   ORBmatcher::SearchByProjection(Frame &CurrentFrame, const Frame &LastFrame, th, bMono)
//                                                     reference src/ORBmatcher.cc:1521-1733 (Nleft == -1), after the
//                                                     projection of the map point (host glue: pose and camera model)
//   ORBmatcher::ComputeThreeMaxima                    reference src/ORBmatcher.cc:1844-1876
//
// The reference loop is a greedy sequential assignment: last-frame keypoint i takes the best current keypoint of its
// window that is not LOCKED (already assigned to a map point with Observations() > 0 by an earlier i). Everything
// that does not depend on the lock is done in parallel, the lock itself is replayed in order:
//   k_grid_build   one CTA per frame: cell of every keypoint, block scan, stable fill (ascending keypoint index
//                  inside a cell, like the reference's push_back order) -> CSR in the reference's cell order ix * 48 + iy
//   k_sp_window    one warp per query: window cells in GetFeaturesInArea order, level / distance / uRight gates,
//                  256-bit Hamming; candidates with distance <= TH_HIGH (only those can ever be assigned) sorted by
//                  (distance, visiting order) - the reference's strict "<" scan picks the first minimum; the best
//                  SP_K are kept
//   k_sp_resolve   one CTA per frame: thread 0 replays the queries in order from shared memory (first unlocked
//                  candidate wins, overwrite of unlocked assignments like the reference, rotation histogram
//                  records incl. duplicates), then ComputeThreeMaxima and the removal of the losing bins.
//                  A query whose SP_K stored candidates are all locked while more exist re-scans its window
//                  (exact slow path, practically never taken).
#include <algorithm>
#include <cstring>

#include "orb_internal.h"

#define GRID_COLS 64      // FRAME_GRID_COLS (include/Frame.h:45)
#define GRID_ROWS 48      // FRAME_GRID_ROWS (include/Frame.h:44)
#define GRID_CELLS (GRID_COLS * GRID_ROWS)
#define SP_TH_HIGH 100    // ORBmatcher::TH_HIGH (src/ORBmatcher.cc:35)
#define SP_HISTO 30       // ORBmatcher::HISTO_LENGTH (src/ORBmatcher.cc:37)
#define SP_K 4            // sorted candidates kept per query
#define SP_WARPS 8
#define SP_LIST 64        // candidates (distance <= TH_HIGH) a warp can collect before it reports overflow

struct GridParams { float min_x, min_y, max_x, max_y, w_inv, h_inv; };

// ---- Frame::AssignFeaturesToGrid -------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_grid_build(const orb_keypoint* __restrict__ kps, const int* __restrict__ n_arr, int kcap,
                                                    GridParams gp, int* __restrict__ cell_off, unsigned short* __restrict__ cell_idx,
                                                    unsigned short* __restrict__ kp_cell) {
  __shared__ int s_cnt[GRID_CELLS];
  __shared__ int s_warp[8];
  __shared__ int s_carry;
  const int frame = blockIdx.x, tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const int n = min(n_arr[frame], kcap);
  const orb_keypoint* kp = kps + (size_t)frame * kcap;
  int* off = cell_off + (size_t)frame * (GRID_CELLS + 1);
  unsigned short* idx = cell_idx + (size_t)frame * kcap;
  unsigned short* kc = kp_cell + (size_t)frame * kcap;
  for (int i = tid; i < GRID_CELLS; i += 256) s_cnt[i] = 0;
  if (tid == 0) s_carry = 0;
  __syncthreads();
  for (int i = tid; i < n; i += 256) {
    // PosInGrid (:809-820): round((kp.pt.x - mnMinX) * mfGridElementWidthInv), half away from zero
    const int px = (int)roundf(__fmul_rn(__fsub_rn(kp[i].x, gp.min_x), gp.w_inv));
    const int py = (int)roundf(__fmul_rn(__fsub_rn(kp[i].y, gp.min_y), gp.h_inv));
    int c = 0xffff;
    if (px >= 0 && px < GRID_COLS && py >= 0 && py < GRID_ROWS) {
      c = px * GRID_ROWS + py;
      atomicAdd(&s_cnt[c], 1);
    }
    kc[i] = (unsigned short)c;
  }
  __syncthreads();
  // exclusive scan over the cells, 256 per sweep; s_cnt becomes the fill cursor
  for (int base = 0; base < GRID_CELLS; base += 256) {
    const int c = s_cnt[base + tid];
    int incl = c;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int v = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= o) incl += v;
    }
    if (lane == 31) s_warp[wid] = incl;
    __syncthreads();
    int before = s_carry, total = 0;
#pragma unroll
    for (int k = 0; k < 8; ++k) { if (k < wid) before += s_warp[k]; total += s_warp[k]; }
    const int start = before + incl - c;
    off[base + tid] = start;
    s_cnt[base + tid] = start;
    __syncthreads();
    if (tid == 0) s_carry += total;
    __syncthreads();
  }
  if (tid == 0) off[GRID_CELLS] = s_carry;
  // stable fill by ONE warp: keypoints in ascending order, lanes of the same cell ranked by lane
  if (wid == 0) {
    for (int base = 0; base < n; base += 32) {
      const int i = base + lane;
      const int c = i < n ? (int)kc[i] : 0xffff;
      const bool valid = c != 0xffff;
      const unsigned act = __ballot_sync(0xffffffffu, valid);
      if (valid) {
        const unsigned same = __match_any_sync(act, c);
        const int rank = __popc(same & ((1u << lane) - 1u));
        const int start = s_cnt[c];
        __syncwarp(act);
        if (rank == 0) s_cnt[c] = start + __popc(same);
        idx[start + rank] = (unsigned short)i;
      }
      __syncwarp();
    }
  }
}

// ---- window search of one query ------------------------------------------------------------------------------
static __device__ __forceinline__ int hamming256(const uint4 a0, const uint4 a1, const uint4* __restrict__ b) {
  const uint4 b0 = b[0], b1 = b[1];
  return __popc(a0.x ^ b0.x) + __popc(a0.y ^ b0.y) + __popc(a0.z ^ b0.z) + __popc(a0.w ^ b0.w) + __popc(a1.x ^ b1.x) +
         __popc(a1.y ^ b1.y) + __popc(a1.z ^ b1.z) + __popc(a1.w ^ b1.w);
}

struct SpWindow {
  int min_cx, min_cy, nx, ny;   // nx * ny cells, visited ix-major like GetFeaturesInArea
  int min_level, max_level;
  float u, v, radius, invz;
  bool ok;
};

// the part of the reference loop body before the candidate scan (:1541-1583) and GetFeaturesInArea's cell range
static __device__ __forceinline__ SpWindow sp_window(const orb_proj_query& q, const GridParams& gp, const float* __restrict__ scale,
                                                     float th, int mode) {
  SpWindow w;
  w.ok = false;
  if (!(q.flags & 1)) return w;                                   // no map point / outlier (:1541-1543)
  w.invz = (float)__ddiv_rn(1.0, (double)q.z);                    // const float invzc = 1.0 / x3Dc(2) (:1550)
  if (w.invz < 0) return w;
  w.u = q.u; w.v = q.v;
  if (w.u < gp.min_x || w.u > gp.max_x) return w;                 // :1556-1559
  if (w.v < gp.min_y || w.v > gp.max_y) return w;
  const int oct = q.octave;
  w.radius = __fmul_rn(th, scale[oct]);                           // :1567
  if (mode == 1) { w.min_level = oct; w.max_level = -1; }         // bForward  (:1571-1573)
  else if (mode == 2) { w.min_level = 0; w.max_level = oct; }     // bBackward (:1574-1576)
  else { w.min_level = oct - 1; w.max_level = oct + 1; }          // :1577-1579
  const float r = w.radius;
  const int minx = max(0, (int)floorf(__fmul_rn(__fsub_rn(__fsub_rn(w.u, gp.min_x), r), gp.w_inv)));
  if (minx >= GRID_COLS) return w;
  const int maxx = min(GRID_COLS - 1, (int)ceilf(__fmul_rn(__fadd_rn(__fsub_rn(w.u, gp.min_x), r), gp.w_inv)));
  if (maxx < 0) return w;
  const int miny = max(0, (int)floorf(__fmul_rn(__fsub_rn(__fsub_rn(w.v, gp.min_y), r), gp.h_inv)));
  if (miny >= GRID_ROWS) return w;
  const int maxy = min(GRID_ROWS - 1, (int)ceilf(__fmul_rn(__fadd_rn(__fsub_rn(w.v, gp.min_y), r), gp.h_inv)));
  if (maxy < 0) return w;
  w.min_cx = minx; w.min_cy = miny; w.nx = maxx - minx + 1; w.ny = maxy - miny + 1;
  w.ok = w.nx > 0 && w.ny > 0;
  return w;
}

// gates of GetFeaturesInArea (:787-799) and of the candidate loop (:1595-1599) that do not depend on the lock
static __device__ __forceinline__ bool sp_gate(const SpWindow& w, const orb_keypoint& k, float uright, float mbf) {
  const bool check_levels = (w.min_level > 0) || (w.max_level >= 0);
  if (check_levels) {
    if (k.octave < w.min_level) return false;
    if (w.max_level >= 0 && k.octave > w.max_level) return false;
  }
  const float distx = __fsub_rn(k.x, w.u), disty = __fsub_rn(k.y, w.v);
  if (!(fabsf(distx) < w.radius && fabsf(disty) < w.radius)) return false;
  if (uright > 0) {
    const float ur = __fsub_rn(w.u, __fmul_rn(mbf, w.invz));
    const float er = fabsf(__fsub_rn(ur, uright));
    if (er > w.radius) return false;
  }
  return true;
}

// candidate key: distance << 40 | cell visiting order << 28 | position inside the cell << 16 | keypoint index
static __device__ __forceinline__ unsigned long long sp_key(int dist, int c, int j, int i2) {
  return ((unsigned long long)dist << 40) | ((unsigned long long)c << 28) | ((unsigned long long)j << 16) | (unsigned long long)i2;
}

__global__ void __launch_bounds__(SP_WARPS * 32) k_sp_window(
    const orb_keypoint* __restrict__ kps, const uint8_t* __restrict__ desc, const float* __restrict__ uright, int kcap,
    const int* __restrict__ cell_off, const unsigned short* __restrict__ cell_idx, const orb_proj_query* __restrict__ queries,
    const uint8_t* __restrict__ qdesc, const int* __restrict__ nq_arr, int qcap, GridParams gp, OrbGeom g, float th,
    const float* __restrict__ tlc_z, float mb, int mono, float mbf, unsigned int* __restrict__ cand, unsigned char* __restrict__ cand_cnt) {
  __shared__ unsigned long long s_keys[SP_WARPS][SP_LIST];
  __shared__ int s_n[SP_WARPS];
  const int frame = blockIdx.y, lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const int qi = blockIdx.x * SP_WARPS + wid;
  if (qi >= min(nq_arr[frame], qcap)) return;
  const size_t qo = (size_t)frame * qcap + qi;
  const orb_proj_query q = queries[qo];
  const float tz = tlc_z[frame];
  const int mode = (tz > mb && !mono) ? 1 : ((-tz > mb && !mono) ? 2 : 0);   // bForward / bBackward (:1537-1538)
  const SpWindow w = sp_window(q, gp, g.scale, th, mode);
  if (lane == 0) s_n[wid] = 0;
  __syncwarp();
  if (w.ok) {
    const uint4* qd = reinterpret_cast<const uint4*>(qdesc + qo * 32);
    const uint4 a0 = qd[0], a1 = qd[1];
    const int* off = cell_off + (size_t)frame * (GRID_CELLS + 1);
    const unsigned short* idx = cell_idx + (size_t)frame * kcap;
    const orb_keypoint* kp = kps + (size_t)frame * kcap;
    const uint8_t* dc = desc + (size_t)frame * kcap * 32;
    const float* ur = uright ? uright + (size_t)frame * kcap : nullptr;
    const int nc = w.nx * w.ny;
    for (int c = lane; c < nc; c += 32) {
      const int cx = c / w.ny, cy = c - cx * w.ny;
      const int cell = (w.min_cx + cx) * GRID_ROWS + w.min_cy + cy;
      const int b = off[cell], e = off[cell + 1];
      for (int p = b; p < e; ++p) {
        const int i2 = idx[p];
        const orb_keypoint k = kp[i2];
        if (!sp_gate(w, k, ur ? ur[i2] : -1.f, mbf)) continue;
        const int d = hamming256(a0, a1, reinterpret_cast<const uint4*>(dc + (size_t)i2 * 32));
        if (d <= SP_TH_HIGH) {
          const int slot = atomicAdd(&s_n[wid], 1);
          if (slot < SP_LIST) s_keys[wid][slot] = sp_key(d, c, p - b, i2);
        }
      }
    }
  }
  __syncwarp();
  if (lane == 0) {
    const int n = s_n[wid];
    unsigned int* out = cand + qo * SP_K;
    if (n > SP_LIST) {
      cand_cnt[qo] = 255;   // overflow: the resolver re-scans this query's window
    } else {
      // selection of the SP_K smallest keys (n is 0..3 in practice)
      unsigned long long* a = s_keys[wid];
      const int keep = min(n, SP_K);
      for (int i = 0; i < keep; ++i) {
        int m = i;
        for (int j = i + 1; j < n; ++j)
          if (a[j] < a[m]) m = j;
        const unsigned long long t = a[i]; a[i] = a[m]; a[m] = t;
        out[i] = ((unsigned int)(a[i] >> 40) << 16) | (unsigned int)(a[i] & 0xffffu);
      }
      cand_cnt[qo] = (unsigned char)min(n, 254);
    }
  }
}

// exact re-scan of one query's window with the current locks (slow path of the resolver, one thread)
static __device__ int sp_rescan(const orb_proj_query& q, const uint8_t* __restrict__ qd8, const GridParams& gp, const OrbGeom& g, float th,
                                int mode, float mbf, const orb_keypoint* __restrict__ kp, const uint8_t* __restrict__ dc,
                                const float* __restrict__ ur, const int* __restrict__ off, const unsigned short* __restrict__ idx,
                                const short* assigned, const unsigned char* qlock, int* best_dist) {
  const SpWindow w = sp_window(q, gp, g.scale, th, mode);
  int bestDist = 256, bestIdx = -1;
  if (!w.ok) { *best_dist = bestDist; return -1; }
  const uint4* qd = reinterpret_cast<const uint4*>(qd8);
  const uint4 a0 = qd[0], a1 = qd[1];
  for (int cx = 0; cx < w.nx; ++cx)
    for (int cy = 0; cy < w.ny; ++cy) {
      const int cell = (w.min_cx + cx) * GRID_ROWS + w.min_cy + cy;
      for (int p = off[cell]; p < off[cell + 1]; ++p) {
        const int i2 = idx[p];
        const int a = assigned[i2];
        if (a >= 0 && qlock[a]) continue;
        const orb_keypoint k = kp[i2];
        if (!sp_gate(w, k, ur ? ur[i2] : -1.f, mbf)) continue;
        const int d = hamming256(a0, a1, reinterpret_cast<const uint4*>(dc + (size_t)i2 * 32));
        if (d < bestDist) { bestDist = d; bestIdx = i2; }
      }
    }
  *best_dist = bestDist;
  return bestIdx;
}

// dynamic shared memory: first[qcap] u32 | qangle[qcap] f32 | cangle[kcap] f32 | recs[qcap] u32 | assigned[kcap] i16 |
// cnt[qcap] u8 | qlock[qcap] u8
static size_t sp_resolve_smem(int qcap, int kcap) {
  return (size_t)qcap * 4 * 3 + (size_t)kcap * 4 + (((size_t)kcap * 2 + 3) & ~(size_t)3) + (size_t)qcap * 2 + 64;
}

__global__ void __launch_bounds__(128) k_sp_resolve(
    const orb_keypoint* __restrict__ kps, const uint8_t* __restrict__ desc, const float* __restrict__ uright, const int* __restrict__ n_arr,
    int kcap, const int* __restrict__ cell_off, const unsigned short* __restrict__ cell_idx, const orb_proj_query* __restrict__ queries,
    const uint8_t* __restrict__ qdesc, const int* __restrict__ nq_arr, int qcap, GridParams gp, OrbGeom g, float th,
    const float* __restrict__ tlc_z, float mb, int mono, float mbf, int check_orientation, const unsigned int* __restrict__ cand,
    const unsigned char* __restrict__ cand_cnt, int* __restrict__ match_out, int* __restrict__ nmatches_out) {
  extern __shared__ __align__(16) unsigned char s_raw[];
  unsigned int* s_first = reinterpret_cast<unsigned int*>(s_raw);
  float* s_qangle = reinterpret_cast<float*>(s_first + qcap);
  float* s_cangle = s_qangle + qcap;
  unsigned int* s_recs = reinterpret_cast<unsigned int*>(s_cangle + kcap);
  short* s_assigned = reinterpret_cast<short*>(s_recs + qcap);
  unsigned char* s_cnt = reinterpret_cast<unsigned char*>(s_assigned) + (((size_t)kcap * 2 + 3) & ~(size_t)3);
  unsigned char* s_qlock = s_cnt + qcap;
  __shared__ int s_hist[SP_HISTO];
  __shared__ int s_nm;
  const int frame = blockIdx.x, tid = threadIdx.x;
  const int nC = min(n_arr[frame], kcap), nq = min(nq_arr[frame], qcap);
  const orb_keypoint* kp = kps + (size_t)frame * kcap;
  const orb_proj_query* q = queries + (size_t)frame * qcap;
  const unsigned int* cd = cand + (size_t)frame * qcap * SP_K;
  const unsigned char* cc = cand_cnt + (size_t)frame * qcap;
  for (int i = tid; i < nq; i += 128) {
    s_first[i] = cd[(size_t)i * SP_K];
    s_cnt[i] = cc[i];
    s_qangle[i] = q[i].angle;
    s_qlock[i] = (q[i].flags & 2) ? 1 : 0;   // Observations() > 0
  }
  for (int i = tid; i < nC; i += 128) { s_cangle[i] = kp[i].angle; s_assigned[i] = -1; }
  if (tid < SP_HISTO) s_hist[tid] = 0;
  __syncthreads();
  if (tid == 0) {
    const float tz = tlc_z[frame];
    const int mode = (tz > mb && !mono) ? 1 : ((-tz > mb && !mono) ? 2 : 0);
    const float factor = 1.0f / SP_HISTO;
    int nm = 0, nrec = 0;
    for (int i = 0; i < nq; ++i) {
      const int cnt = s_cnt[i];
      if (cnt == 0) continue;
      int pick = -1;
      if (cnt != 255) {
        const int stored = min(cnt, SP_K);
        for (int k = 0; k < stored; ++k) {
          const unsigned int key = k == 0 ? s_first[i] : cd[(size_t)i * SP_K + k];
          const int i2 = (int)(key & 0xffffu);
          const int a = s_assigned[i2];
          if (a >= 0 && s_qlock[a]) continue;                     // :1592-1593
          pick = i2;
          break;
        }
      }
      if (pick < 0 && (cnt == 255 || cnt > SP_K)) {
        // every stored candidate is locked but the window holds more: exact re-scan (practically never)
        int bd;
        const int bi = sp_rescan(q[i], qdesc + ((size_t)frame * qcap + i) * 32, gp, g, th, mode, mbf, kp, desc + (size_t)frame * kcap * 32,
                                 uright ? uright + (size_t)frame * kcap : nullptr, cell_off + (size_t)frame * (GRID_CELLS + 1),
                                 cell_idx + (size_t)frame * kcap, s_assigned, s_qlock, &bd);
        if (bi >= 0 && bd <= SP_TH_HIGH) pick = bi;
      }
      if (pick >= 0) {                                            // bestDist <= TH_HIGH (:1610)
        s_assigned[pick] = (short)i;
        nm++;
        if (check_orientation) {
          float rot = __fsub_rn(s_qangle[i], s_cangle[pick]);
          if (rot < 0.0f) rot = __fadd_rn(rot, 360.0f);
          int bin = (int)roundf(__fmul_rn(rot, factor));
          if (bin == SP_HISTO) bin = 0;
          s_recs[nrec++] = (unsigned int)pick | ((unsigned int)bin << 16);
          s_hist[bin]++;
        }
      }
    }
    if (check_orientation) {
      // ComputeThreeMaxima (:1844-1876)
      int ind1 = -1, ind2 = -1, ind3 = -1, max1 = 0, max2 = 0, max3 = 0;
      for (int i = 0; i < SP_HISTO; i++) {
        const int s = s_hist[i];
        if (s > max1) { max3 = max2; max2 = max1; max1 = s; ind3 = ind2; ind2 = ind1; ind1 = i; }
        else if (s > max2) { max3 = max2; max2 = s; ind3 = ind2; ind2 = i; }
        else if (s > max3) { max3 = s; ind3 = i; }
      }
      if ((float)max2 < __fmul_rn(0.1f, (float)max1)) { ind2 = -1; ind3 = -1; }
      else if ((float)max3 < __fmul_rn(0.1f, (float)max1)) { ind3 = -1; }
      // every record of a losing bin clears its keypoint and takes one match back (:1718-1728), duplicates included
      for (int r = 0; r < nrec; ++r) {
        const int bin = (int)(s_recs[r] >> 16);
        if (bin != ind1 && bin != ind2 && bin != ind3) { s_assigned[s_recs[r] & 0xffffu] = -1; nm--; }
      }
    }
    s_nm = nm;
  }
  __syncthreads();
  for (int i = tid; i < kcap; i += 128) match_out[(size_t)frame * kcap + i] = i < nC ? (int)s_assigned[i] : -1;
  if (tid == 0) nmatches_out[frame] = s_nm;
}

// ---- ORBmatcher::SearchByProjection(Frame &F, const vector<MapPoint*> &vpMapPoints, th, bFarPoints, thFarPoints) ---------
// reference src/ORBmatcher.cc:42-209 (Nleft == -1) + RadiusByViewingCos (:211-216): the local map against one frame.
// Same split as above. What depends on the order of the map points is only the lock (a keypoint that holds a map point
// with observations is skipped BEFORE the distance is looked at, :86-87), and the locks that exist when the call starts
// are an input (locked0), so the window kernel already leaves those out. Of the reference's scan only two candidates
// matter: the first two unlocked ones in (distance, visiting order) - "best" is the first minimum of a strict "<" scan,
// "second" is the first arrival at the second-smallest value whichever way the scan reaches it.
//   k_sl_window    one warp per map point: window cells, level / distance / uRight gates, Hamming; every lane keeps its
//                  SL_K smallest keys (distance << 16 | CSR position: CSR positions ascend in visiting order) sorted in
//                  registers, SL_K rounds of redux.min merge them
//   k_sl_resolve   one CTA per frame: candidates staged in shared memory chunk by chunk, one thread replays the map points
//                  in order (first two unlocked candidates, ratio test on equal levels, overwrite rule), exact re-scan when
//                  the stored candidates run out while the window holds more
#define SL_K 4
#define SL_CHUNK 2048
#define SL_NONE 0xffffffffu

struct SlWindow {
  int min_cx, min_cy, nx, ny, min_level, max_level;
  float u, v, radius, xr;
  bool ok;
};

static __device__ __forceinline__ SlWindow sl_window(const orb_track_query& q, const GridParams& gp, const OrbGeom& g, float th) {
  SlWindow w;
  w.ok = false;
  if (!(q.flags & 1)) return w;                                   // :52-56
  const int lvl = q.level;
  if (lvl < 0 || lvl >= g.nlevels) return w;                      // the reference would index mvScaleFactors out of range
  float r = ((double)q.view_cos > 0.998) ? 2.5f : 4.0f;           // RadiusByViewingCos: float against a double literal
  if (th != 1.0f) r = __fmul_rn(r, th);                           // bFactor (:48,65)
  w.radius = __fmul_rn(r, g.scale[lvl]);                          // :69
  w.u = q.proj_x; w.v = q.proj_y; w.xr = q.proj_xr;
  w.min_level = lvl - 1; w.max_level = lvl;                       // :70
  const float rr = w.radius;
  const int minx = max(0, (int)floorf(__fmul_rn(__fsub_rn(__fsub_rn(w.u, gp.min_x), rr), gp.w_inv)));
  if (minx >= GRID_COLS) return w;
  const int maxx = min(GRID_COLS - 1, (int)ceilf(__fmul_rn(__fadd_rn(__fsub_rn(w.u, gp.min_x), rr), gp.w_inv)));
  if (maxx < 0) return w;
  const int miny = max(0, (int)floorf(__fmul_rn(__fsub_rn(__fsub_rn(w.v, gp.min_y), rr), gp.h_inv)));
  if (miny >= GRID_ROWS) return w;
  const int maxy = min(GRID_ROWS - 1, (int)ceilf(__fmul_rn(__fadd_rn(__fsub_rn(w.v, gp.min_y), rr), gp.h_inv)));
  if (maxy < 0) return w;
  w.min_cx = minx; w.min_cy = miny; w.nx = maxx - minx + 1; w.ny = maxy - miny + 1;
  w.ok = w.nx > 0 && w.ny > 0;
  return w;
}

// GetFeaturesInArea's gates (src/Frame.cc:787-799; bCheckLevels is always true here: maxLevel = level >= 0) and :89-92
static __device__ __forceinline__ bool sl_gate(const SlWindow& w, const orb_keypoint& k, float uright) {
  if (k.octave < w.min_level || k.octave > w.max_level) return false;
  const float distx = __fsub_rn(k.x, w.u), disty = __fsub_rn(k.y, w.v);
  if (!(fabsf(distx) < w.radius && fabsf(disty) < w.radius)) return false;
  if (uright > 0) {
    const float er = fabsf(__fsub_rn(w.xr, uright));
    if (er > w.radius) return false;
  }
  return true;
}

__global__ void __launch_bounds__(SP_WARPS * 32) k_sl_window(
    const orb_keypoint* __restrict__ kps, const uint8_t* __restrict__ desc, const float* __restrict__ uright, int kcap,
    const int* __restrict__ cell_off, const unsigned short* __restrict__ cell_idx, const orb_track_query* __restrict__ queries,
    const uint8_t* __restrict__ qdesc, const int* __restrict__ nq_arr, int qcap, const uint8_t* __restrict__ locked0, GridParams gp, OrbGeom g,
    float th, uint4* __restrict__ cand, unsigned char* __restrict__ cand_cnt) {
  const int frame = blockIdx.y, lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const int qi = blockIdx.x * SP_WARPS + wid;
  if (qi >= min(nq_arr[frame], qcap)) return;
  const size_t qo = (size_t)frame * qcap + qi;
  const orb_track_query q = queries[qo];
  const SlWindow w = sl_window(q, gp, g, th);
  unsigned int k0 = SL_NONE, k1 = SL_NONE, k2 = SL_NONE, k3 = SL_NONE;
  int cnt = 0;
  const unsigned short* idx = cell_idx + (size_t)frame * kcap;
  const orb_keypoint* kp = kps + (size_t)frame * kcap;
  if (w.ok) {
    const uint4* qd = reinterpret_cast<const uint4*>(qdesc + qo * 32);
    const uint4 a0 = qd[0], a1 = qd[1];
    const int* off = cell_off + (size_t)frame * (GRID_CELLS + 1);
    const uint8_t* dc = desc + (size_t)frame * kcap * 32;
    const float* ur = uright ? uright + (size_t)frame * kcap : nullptr;
    const uint8_t* lk = locked0 ? locked0 + (size_t)frame * kcap : nullptr;
    const int nc = w.nx * w.ny;
    for (int c = lane; c < nc; c += 32) {
      const int cx = c / w.ny, cy = c - cx * w.ny;
      const int cell = (w.min_cx + cx) * GRID_ROWS + w.min_cy + cy;
      const int b = off[cell], e = off[cell + 1];
      for (int p = b; p < e; ++p) {
        const int i2 = idx[p];
        if (lk && lk[i2]) continue;                               // :86-87 for the locks that exist before the call
        const orb_keypoint k = kp[i2];
        if (!sl_gate(w, k, ur ? ur[i2] : -1.f)) continue;
        const int d = hamming256(a0, a1, reinterpret_cast<const uint4*>(dc + (size_t)i2 * 32));
        unsigned int key = ((unsigned int)d << 16) | (unsigned int)p;
        cnt++;
        if (key < k3) {
          k3 = key;
          if (k3 < k2) { const unsigned int t = k2; k2 = k3; k3 = t; }
          if (k2 < k1) { const unsigned int t = k1; k1 = k2; k2 = t; }
          if (k1 < k0) { const unsigned int t = k0; k0 = k1; k1 = t; }
        }
      }
    }
  }
  cnt = __reduce_add_sync(0xffffffffu, cnt);
  unsigned int mine = SL_NONE;
#pragma unroll
  for (int r = 0; r < SL_K; ++r) {
    const unsigned int m = __reduce_min_sync(0xffffffffu, k0);
    if (m != SL_NONE && k0 == m) { k0 = k1; k1 = k2; k2 = k3; k3 = SL_NONE; }   // CSR positions are unique: one lane pops
    if (lane == r) mine = m;
  }
  unsigned int rec = SL_NONE;
  if (lane < SL_K && mine != SL_NONE) {
    const int i2 = idx[mine & 0xffffu];
    rec = ((mine >> 16) << 20) | ((unsigned int)(kp[i2].octave & 15) << 16) | (unsigned int)i2;
  }
  const unsigned int r1 = __shfl_sync(0xffffffffu, rec, 1), r2 = __shfl_sync(0xffffffffu, rec, 2), r3 = __shfl_sync(0xffffffffu, rec, 3);
  if (lane == 0) {
    cand[qo] = make_uint4(rec, r1, r2, r3);
    cand_cnt[qo] = (unsigned char)min(cnt, 255);
  }
}

// exact scan of one map point's window under the current locks (slow path of the resolver, one thread): :77-116
static __device__ void sl_rescan(const orb_track_query& q, const uint8_t* __restrict__ qd8, const GridParams& gp, const OrbGeom& g, float th,
                                 const orb_keypoint* __restrict__ kp, const uint8_t* __restrict__ dc, const float* __restrict__ ur,
                                 const int* __restrict__ off, const unsigned short* __restrict__ idx, const unsigned char* lock,
                                 int* bestDist_, int* bestLevel_, int* bestDist2_, int* bestLevel2_, int* bestIdx_) {
  int bestDist = 256, bestLevel = -1, bestDist2 = 256, bestLevel2 = -1, bestIdx = -1;
  const SlWindow w = sl_window(q, gp, g, th);
  if (w.ok) {
    const uint4* qd = reinterpret_cast<const uint4*>(qd8);
    const uint4 a0 = qd[0], a1 = qd[1];
    for (int cx = 0; cx < w.nx; ++cx)
      for (int cy = 0; cy < w.ny; ++cy) {
        const int cell = (w.min_cx + cx) * GRID_ROWS + w.min_cy + cy;
        for (int p = off[cell]; p < off[cell + 1]; ++p) {
          const int i2 = idx[p];
          if (lock[i2]) continue;
          const orb_keypoint k = kp[i2];
          if (!sl_gate(w, k, ur ? ur[i2] : -1.f)) continue;
          const int d = hamming256(a0, a1, reinterpret_cast<const uint4*>(dc + (size_t)i2 * 32));
          if (d < bestDist) { bestDist2 = bestDist; bestDist = d; bestLevel2 = bestLevel; bestLevel = k.octave; bestIdx = i2; }
          else if (d < bestDist2) { bestLevel2 = k.octave; bestDist2 = d; }
        }
      }
  }
  *bestDist_ = bestDist; *bestLevel_ = bestLevel; *bestDist2_ = bestDist2; *bestLevel2_ = bestLevel2; *bestIdx_ = bestIdx;
}

// dynamic shared memory: assigned[kcap] i32 | lock[kcap] u8
__global__ void __launch_bounds__(128) k_sl_resolve(
    const orb_keypoint* __restrict__ kps, const uint8_t* __restrict__ desc, const float* __restrict__ uright, const int* __restrict__ n_arr,
    int kcap, const int* __restrict__ cell_off, const unsigned short* __restrict__ cell_idx, const orb_track_query* __restrict__ queries,
    const uint8_t* __restrict__ qdesc, const int* __restrict__ nq_arr, int qcap, const uint8_t* __restrict__ locked0, GridParams gp, OrbGeom g,
    float th, float nnratio, const uint4* __restrict__ cand, const unsigned char* __restrict__ cand_cnt, int* __restrict__ match_out,
    int* __restrict__ nmatches_out) {
  extern __shared__ __align__(16) unsigned char s_raw[];
  int* s_assigned = reinterpret_cast<int*>(s_raw);
  unsigned char* s_lock = reinterpret_cast<unsigned char*>(s_assigned + kcap);
  __shared__ uint4 s_cand[SL_CHUNK];
  __shared__ unsigned char s_cnt[SL_CHUNK];
  __shared__ unsigned char s_obs[SL_CHUNK];
  __shared__ int s_nm;
  const int frame = blockIdx.x, tid = threadIdx.x;
  const int nC = min(n_arr[frame], kcap), nq = min(nq_arr[frame], qcap);
  const orb_track_query* q = queries + (size_t)frame * qcap;
  const uint4* cd = cand + (size_t)frame * qcap;
  const unsigned char* cc = cand_cnt + (size_t)frame * qcap;
  for (int i = tid; i < nC; i += 128) {
    s_assigned[i] = -1;
    s_lock[i] = locked0 ? locked0[(size_t)frame * kcap + i] : 0;
  }
  if (tid == 0) s_nm = 0;
  for (int base = 0; base < nq; base += SL_CHUNK) {
    const int m = min(SL_CHUNK, nq - base);
    __syncthreads();
    for (int i = tid; i < m; i += 128) {
      s_cand[i] = cd[base + i];
      s_cnt[i] = cc[base + i];
      s_obs[i] = (q[base + i].flags & 2) ? 1 : 0;                 // Observations() > 0
    }
    __syncthreads();
    if (tid == 0) {
      int nm = s_nm;
      for (int i = 0; i < m; ++i) {
        const int cnt = s_cnt[i];
        if (cnt == 0) continue;
        const uint4 c4 = s_cand[i];
        const unsigned int c[SL_K] = {c4.x, c4.y, c4.z, c4.w};
        unsigned int b1 = SL_NONE, b2 = SL_NONE;
#pragma unroll
        for (int k = 0; k < SL_K; ++k) {
          const unsigned int key = c[k];
          if (key == SL_NONE || b2 != SL_NONE) continue;
          if (s_lock[key & 0xffffu]) continue;                    // locked by an earlier map point of this call
          if (b1 == SL_NONE) b1 = key; else b2 = key;
        }
        int bestDist, bestLevel, bestDist2, bestLevel2, bestIdx;
        if (b2 == SL_NONE && cnt > SL_K) {
          // fewer than two unlocked candidates among the stored ones while the window holds more: exact re-scan
          sl_rescan(q[base + i], qdesc + ((size_t)frame * qcap + base + i) * 32, gp, g, th, kps + (size_t)frame * kcap,
                    desc + (size_t)frame * kcap * 32, uright ? uright + (size_t)frame * kcap : nullptr,
                    cell_off + (size_t)frame * (GRID_CELLS + 1), cell_idx + (size_t)frame * kcap, s_lock, &bestDist, &bestLevel, &bestDist2,
                    &bestLevel2, &bestIdx);
        } else {
          if (b1 == SL_NONE) continue;
          bestDist = (int)(b1 >> 20); bestLevel = (int)((b1 >> 16) & 15u); bestIdx = (int)(b1 & 0xffffu);
          bestDist2 = b2 == SL_NONE ? 256 : (int)(b2 >> 20);
          bestLevel2 = b2 == SL_NONE ? -1 : (int)((b2 >> 16) & 15u);
        }
        if (bestIdx < 0 || bestDist > SP_TH_HIGH) continue;       // :122
        // :123-126: the ratio only applies when best and second share the level; float * int -> float
        if (bestLevel == bestLevel2 && (float)bestDist > __fmul_rn(nnratio, (float)bestDist2)) continue;
        s_assigned[bestIdx] = base + i;                           // :127
        s_lock[bestIdx] = s_obs[i];
        nm++;
      }
      s_nm = nm;
    }
  }
  __syncthreads();
  for (int i = tid; i < kcap; i += 128) match_out[(size_t)frame * kcap + i] = i < nC ? s_assigned[i] : -1;
  if (tid == 0) nmatches_out[frame] = s_nm;
}

// ---- host side ---------------------------------------------------------------------------------------------
static GridParams to_gp(const orb_grid_params* p) {
  GridParams g;
  g.min_x = p->min_x; g.min_y = p->min_y; g.max_x = p->max_x; g.max_y = p->max_y; g.w_inv = p->grid_w_inv; g.h_inv = p->grid_h_inv;
  return g;
}

extern "C" {

int orb_assign_features_to_grid(orb_handle* h, const orb_grid_params* gp, int flags) {
  if (!h || !gp) return ORB_ERR_INVALID_ARG;
  if (!h->have_batch) return orb_set_error(h, ORB_ERR_STATE, "no extraction has run on this handle");
  int st;
  if ((st = orb_use_device(h))) return st;
  const int batch = h->cur_batch, kcap = h->g.kcap;
  if (kcap > 65535) return orb_set_error(h, ORB_ERR_CAPACITY, "more than 65535 keypoints per frame");
  if ((st = orb_ensure(h, h->d_grid_off, (size_t)batch * (GRID_CELLS + 1) * sizeof(int)))) return st;
  if ((st = orb_ensure(h, h->d_grid_idx, (size_t)batch * kcap * sizeof(unsigned short)))) return st;
  if ((st = orb_ensure(h, h->d_grid_cell, (size_t)batch * kcap * sizeof(unsigned short)))) return st;
  h->grid_params = *gp;
  k_grid_build<<<batch, 256, 0, h->stream>>>(h->d_kps.as<orb_keypoint>(), h->d_n.as<int>(), kcap, to_gp(gp), h->d_grid_off.as<int>(),
                                             h->d_grid_idx.as<unsigned short>(), h->d_grid_cell.as<unsigned short>());
  h->launches++;
  ORB_CUDA_CHECK(h, cudaGetLastError());
  h->have_grid = true;
  if (flags & ORB_ASYNC) return ORB_OK;
  ORB_CUDA_CHECK(h, cudaStreamSynchronize(h->stream));
  return ORB_OK;
}

int orb_debug_get_grid(orb_handle* h, int frame, int32_t* cell_off, int32_t* idx, int cap, int* n_out) {
  if (!h || !cell_off || !idx || !n_out) return ORB_ERR_INVALID_ARG;
  if (!h->have_grid) return orb_set_error(h, ORB_ERR_STATE, "orb_assign_features_to_grid has not run on this handle");
  if (frame < 0 || frame >= h->cur_batch) return ORB_ERR_INVALID_ARG;
  int st;
  if ((st = orb_use_device(h))) return st;
  ORB_CUDA_CHECK(h, cudaStreamSynchronize(h->stream));
  ORB_CUDA_CHECK(h, cudaMemcpy(cell_off, h->d_grid_off.as<int>() + (size_t)frame * (GRID_CELLS + 1), (GRID_CELLS + 1) * sizeof(int),
                               cudaMemcpyDeviceToHost));
  const int n = cell_off[GRID_CELLS];
  if (n > cap) return orb_set_error(h, ORB_ERR_CAPACITY, "grid index buffer too small");
  std::vector<unsigned short> tmp(std::max(n, 1));
  ORB_CUDA_CHECK(h, cudaMemcpy(tmp.data(), h->d_grid_idx.as<unsigned short>() + (size_t)frame * h->g.kcap, (size_t)n * sizeof(unsigned short),
                               cudaMemcpyDeviceToHost));
  for (int i = 0; i < n; ++i) idx[i] = tmp[i];
  *n_out = n;
  return ORB_OK;
}

int orb_search_by_projection(orb_handle* h, const orb_proj_query* queries, const uint8_t* qdesc, const int32_t* nq, int qcap, float th,
                             int mono, const float* tlc_z, float mb, float mbf, int check_orientation, int32_t* match_out,
                             int32_t* nmatches_out, int flags) {
  if (!h || !queries || !qdesc || !nq || !tlc_z || qcap < 1) return ORB_ERR_INVALID_ARG;
  if (!h->have_grid) return orb_set_error(h, ORB_ERR_STATE, "orb_assign_features_to_grid has not run on this handle");
  if (qcap > 65535) return orb_set_error(h, ORB_ERR_CAPACITY, "more than 65535 queries per frame");
  int st;
  if ((st = orb_use_device(h))) return st;
  const int batch = h->cur_batch, kcap = h->g.kcap;
  const size_t smem = sp_resolve_smem(qcap, kcap);
  if (smem > 200 * 1024) return orb_set_error(h, ORB_ERR_CAPACITY, "too many queries / keypoints per frame for the resolver");
  const size_t nqt = (size_t)batch * qcap;
  // device copies of the inputs when they live on the host
  const orb_proj_query* d_q = queries;
  const uint8_t* d_qd = qdesc;
  const int* d_nq = nq;
  const float* d_tz = tlc_z;
  if (!(flags & ORB_SRC_DEVICE)) {
    const size_t b_q = nqt * sizeof(orb_proj_query), b_d = nqt * 32, b_n = (size_t)batch * 4;
    const size_t o_d = (b_q + 255) & ~(size_t)255, o_n = o_d + ((b_d + 255) & ~(size_t)255), o_t = o_n + ((b_n + 255) & ~(size_t)255);
    if ((st = orb_ensure(h, h->d_scratch, o_t + b_n))) return st;
    uint8_t* base = h->d_scratch.as<uint8_t>();
    ORB_CUDA_CHECK(h, cudaMemcpyAsync(base, queries, b_q, cudaMemcpyHostToDevice, h->stream));
    ORB_CUDA_CHECK(h, cudaMemcpyAsync(base + o_d, qdesc, b_d, cudaMemcpyHostToDevice, h->stream));
    ORB_CUDA_CHECK(h, cudaMemcpyAsync(base + o_n, nq, b_n, cudaMemcpyHostToDevice, h->stream));
    ORB_CUDA_CHECK(h, cudaMemcpyAsync(base + o_t, tlc_z, b_n, cudaMemcpyHostToDevice, h->stream));
    d_q = (const orb_proj_query*)base; d_qd = base + o_d; d_nq = (const int*)(base + o_n); d_tz = (const float*)(base + o_t);
  }
  if ((st = orb_ensure(h, h->d_sp_cand, nqt * SP_K * sizeof(unsigned int)))) return st;
  if ((st = orb_ensure(h, h->d_sp_cnt, nqt))) return st;
  if ((st = orb_ensure(h, h->d_sp_match, (size_t)batch * kcap * sizeof(int)))) return st;
  if ((st = orb_ensure(h, h->d_sp_nm, (size_t)batch * sizeof(int)))) return st;
  const float* d_ur = h->have_stereo ? h->d_uright.as<float>() : nullptr;   // mvuRight = -1 without a stereo match
  const GridParams gp = to_gp(&h->grid_params);
  ORB_CUDA_CHECK(h, cudaMemsetAsync(h->d_sp_cnt.p, 0, nqt, h->stream));
  k_sp_window<<<dim3((qcap + SP_WARPS - 1) / SP_WARPS, batch), SP_WARPS * 32, 0, h->stream>>>(
      h->d_kps.as<orb_keypoint>(), h->d_desc.as<uint8_t>(), d_ur, kcap, h->d_grid_off.as<int>(), h->d_grid_idx.as<unsigned short>(), d_q, d_qd,
      d_nq, qcap, gp, h->g, th, d_tz, mb, mono, mbf, h->d_sp_cand.as<unsigned int>(), h->d_sp_cnt.as<unsigned char>());
  h->launches++;
  ORB_CUDA_CHECK(h, cudaFuncSetAttribute(k_sp_resolve, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)std::max(smem, (size_t)48 * 1024)));
  k_sp_resolve<<<batch, 128, smem, h->stream>>>(h->d_kps.as<orb_keypoint>(), h->d_desc.as<uint8_t>(), d_ur, h->d_n.as<int>(), kcap,
                                                h->d_grid_off.as<int>(), h->d_grid_idx.as<unsigned short>(), d_q, d_qd, d_nq, qcap, gp, h->g, th,
                                                d_tz, mb, mono, mbf, check_orientation, h->d_sp_cand.as<unsigned int>(),
                                                h->d_sp_cnt.as<unsigned char>(), h->d_sp_match.as<int>(), h->d_sp_nm.as<int>());
  h->launches++;
  ORB_CUDA_CHECK(h, cudaGetLastError());
  if (!(flags & ORB_NO_OUTPUT)) {
    if (match_out) ORB_CUDA_CHECK(h, cudaMemcpyAsync(match_out, h->d_sp_match.p, (size_t)batch * kcap * sizeof(int), cudaMemcpyDefault, h->stream));
    if (nmatches_out) ORB_CUDA_CHECK(h, cudaMemcpyAsync(nmatches_out, h->d_sp_nm.p, (size_t)batch * sizeof(int), cudaMemcpyDefault, h->stream));
  }
  if (flags & ORB_ASYNC) return ORB_OK;
  ORB_CUDA_CHECK(h, cudaStreamSynchronize(h->stream));
  return ORB_OK;
}

int orb_search_local_points(orb_handle* h, const orb_track_query* queries, const uint8_t* qdesc, const int32_t* nq, int qcap,
                            const uint8_t* locked0, float th, float nnratio, int32_t* match_out, int32_t* nmatches_out, int flags) {
  if (!h || !queries || !qdesc || !nq || qcap < 1) return ORB_ERR_INVALID_ARG;
  if (!h->have_grid) return orb_set_error(h, ORB_ERR_STATE, "orb_assign_features_to_grid has not run on this handle");
  int st;
  if ((st = orb_use_device(h))) return st;
  const int batch = h->cur_batch, kcap = h->g.kcap;
  if (h->g.nlevels > 16) return orb_set_error(h, ORB_ERR_CAPACITY, "more than 16 pyramid levels");
  const size_t smem = (size_t)kcap * 5 + 16;
  if (smem > 160 * 1024) return orb_set_error(h, ORB_ERR_CAPACITY, "too many keypoints per frame for the resolver");
  const size_t nqt = (size_t)batch * qcap;
  const orb_track_query* d_q = queries;
  const uint8_t* d_qd = qdesc;
  const int* d_nq = nq;
  const uint8_t* d_lk = locked0;
  if (!(flags & ORB_SRC_DEVICE)) {
    const size_t b_q = nqt * sizeof(orb_track_query), b_d = nqt * 32, b_n = (size_t)batch * 4, b_l = locked0 ? (size_t)batch * kcap : 0;
    const size_t o_d = (b_q + 255) & ~(size_t)255, o_n = o_d + ((b_d + 255) & ~(size_t)255), o_l = o_n + ((b_n + 255) & ~(size_t)255);
    if ((st = orb_ensure(h, h->d_scratch, o_l + b_l + 16))) return st;
    uint8_t* base = h->d_scratch.as<uint8_t>();
    ORB_CUDA_CHECK(h, cudaMemcpyAsync(base, queries, b_q, cudaMemcpyHostToDevice, h->stream));
    ORB_CUDA_CHECK(h, cudaMemcpyAsync(base + o_d, qdesc, b_d, cudaMemcpyHostToDevice, h->stream));
    ORB_CUDA_CHECK(h, cudaMemcpyAsync(base + o_n, nq, b_n, cudaMemcpyHostToDevice, h->stream));
    if (locked0) ORB_CUDA_CHECK(h, cudaMemcpyAsync(base + o_l, locked0, b_l, cudaMemcpyHostToDevice, h->stream));
    d_q = (const orb_track_query*)base; d_qd = base + o_d; d_nq = (const int*)(base + o_n); d_lk = locked0 ? base + o_l : nullptr;
  }
  if ((st = orb_ensure(h, h->d_sp_cand, nqt * SL_K * sizeof(unsigned int)))) return st;
  if ((st = orb_ensure(h, h->d_sp_cnt, nqt))) return st;
  if ((st = orb_ensure(h, h->d_sp_match, (size_t)batch * kcap * sizeof(int)))) return st;
  if ((st = orb_ensure(h, h->d_sp_nm, (size_t)batch * sizeof(int)))) return st;
  const float* d_ur = h->have_stereo ? h->d_uright.as<float>() : nullptr;   // mvuRight = -1 without a stereo match
  const GridParams gp = to_gp(&h->grid_params);
  k_sl_window<<<dim3((qcap + SP_WARPS - 1) / SP_WARPS, batch), SP_WARPS * 32, 0, h->stream>>>(
      h->d_kps.as<orb_keypoint>(), h->d_desc.as<uint8_t>(), d_ur, kcap, h->d_grid_off.as<int>(), h->d_grid_idx.as<unsigned short>(), d_q, d_qd,
      d_nq, qcap, d_lk, gp, h->g, th, h->d_sp_cand.as<uint4>(), h->d_sp_cnt.as<unsigned char>());
  h->launches++;
  ORB_CUDA_CHECK(h, cudaFuncSetAttribute(k_sl_resolve, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)std::max(smem, (size_t)8 * 1024)));
  k_sl_resolve<<<batch, 128, smem, h->stream>>>(h->d_kps.as<orb_keypoint>(), h->d_desc.as<uint8_t>(), d_ur, h->d_n.as<int>(), kcap,
                                                h->d_grid_off.as<int>(), h->d_grid_idx.as<unsigned short>(), d_q, d_qd, d_nq, qcap, d_lk, gp,
                                                h->g, th, nnratio, h->d_sp_cand.as<uint4>(), h->d_sp_cnt.as<unsigned char>(),
                                                h->d_sp_match.as<int>(), h->d_sp_nm.as<int>());
  h->launches++;
  ORB_CUDA_CHECK(h, cudaGetLastError());
  if (!(flags & ORB_NO_OUTPUT)) {
    if (match_out) ORB_CUDA_CHECK(h, cudaMemcpyAsync(match_out, h->d_sp_match.p, (size_t)batch * kcap * sizeof(int), cudaMemcpyDefault, h->stream));
    if (nmatches_out) ORB_CUDA_CHECK(h, cudaMemcpyAsync(nmatches_out, h->d_sp_nm.p, (size_t)batch * sizeof(int), cudaMemcpyDefault, h->stream));
  }
  if (flags & ORB_ASYNC) return ORB_OK;
  ORB_CUDA_CHECK(h, cudaStreamSynchronize(h->stream));
  return ORB_OK;
}

}  // extern "C"
