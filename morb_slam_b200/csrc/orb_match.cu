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
//   k_sp_window    one warp per query: the window is nx contiguous CSR ranges (one per grid column), flattened over the
//                  lanes; level / distance / uRight gates, 256-bit Hamming; only candidates with distance <= TH_HIGH
//                  can ever be assigned; keys (distance << 16 | CSR position) order them like the reference's strict "<"
//                  scan (CSR positions ascend in GetFeaturesInArea's visiting order); the best SL_K are kept
//   k_sp_resolve   one CTA per frame: warp 0 replays the queries in order, 32 per round, committing the lanes before the
//                  first lock conflict (first unlocked candidate wins, overwrite of unlocked assignments like the
//                  reference, rotation histogram records incl. duplicates), then ComputeThreeMaxima and the removal of
//                  the losing bins. A query whose stored candidates are all locked while more exist is scanned again by
//                  the whole warp under the current locks (exact).
#include <algorithm>
#include <cstring>
#include <cmath>

#include "orb_internal.h"

#define GRID_COLS 64      // FRAME_GRID_COLS (include/Frame.h:45)
#define GRID_ROWS 48      // FRAME_GRID_ROWS (include/Frame.h:44)
#define GRID_CELLS (GRID_COLS * GRID_ROWS)
#define SP_TH_HIGH 100    // ORBmatcher::TH_HIGH (src/ORBmatcher.cc:35)
#define SP_HISTO 30       // ORBmatcher::HISTO_LENGTH (src/ORBmatcher.cc:37)
#define SP_WARPS 8
#ifndef SP_MINB
#define SP_MINB 6   // warp-per-query gather chains: 40 registers for 48 resident warps per SM (0.242 -> 0.218 ms SearchByProjection, 0.459 -> 0.412 ms local map per 256 frames; 8: 0.235 / 0.425)
#endif

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
  float u, v, radius, invz, ur;
  bool ok;
};

// the part of the reference loop body before the candidate scan (:1541-1583) and GetFeaturesInArea's cell range
static __device__ __forceinline__ SpWindow sp_window(const orb_proj_query& q, const GridParams& gp, const float* __restrict__ scale,
                                                     float th, int mode, float mbf) {
  SpWindow w;
  w.ok = false;
  if (!(q.flags & 1)) return w;                                   // no map point / outlier (:1541-1543)
  w.invz = 0.f;
  if (mode != 3 && mode != 4) {                                   // modes 3 / 4 = the KeyFrame / Sim3 overloads: the caller tests the depth
    w.invz = (float)__ddiv_rn(1.0, (double)q.z);                  // const float invzc = 1.0 / x3Dc(2) (:1550)
    if (w.invz < 0) return w;
  }
  w.u = q.u; w.v = q.v;
  if (w.u < gp.min_x || w.u > gp.max_x) return w;                 // :1556-1559
  if (w.v < gp.min_y || w.v > gp.max_y) return w;
  const int oct = q.octave;
  w.radius = __fmul_rn(th, scale[oct]);                           // :1567
  if (mode == 1) { w.min_level = oct; w.max_level = -1; }         // bForward  (:1571-1573)
  else if (mode == 2) { w.min_level = 0; w.max_level = oct; }     // bBackward (:1574-1576)
  else if (mode == 4) { w.min_level = oct - 1; w.max_level = oct; }   // the Sim3 overloads (:469, :572): kpLevel in [l - 1, l]
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
  w.ur = __fsub_rn(w.u, __fmul_rn(mbf, w.invz));                 // const float ur = uv(0) - CurrentFrame.mbf * invzc (:1596)
  w.ok = w.nx > 0 && w.ny > 0;
  return w;
}

// gates of GetFeaturesInArea (:787-799) and of the candidate loop (:1595-1599) that do not depend on the lock
static __device__ __forceinline__ bool win_gate(const SpWindow& w, const orb_keypoint& k, float uright) {
  const bool check_levels = (w.min_level > 0) || (w.max_level >= 0);
  if (check_levels) {
    if (k.octave < w.min_level) return false;
    if (w.max_level >= 0 && k.octave > w.max_level) return false;
  }
  const float distx = __fsub_rn(k.x, w.u), disty = __fsub_rn(k.y, w.v);
  if (!(fabsf(distx) < w.radius && fabsf(disty) < w.radius)) return false;
  if (uright > 0) {
    const float er = fabsf(__fsub_rn(w.ur, uright));
    if (er > w.radius) return false;
  }
  return true;
}

#define SL_K 4            // sorted candidates kept per query
#define SL_CHUNK 2048
#define SL_NONE 0xffffffffu

// The window of a query is nx grid columns; the cells of one column are consecutive in the CSR (cell = ix * 48 + iy), so
// the window is nx contiguous CSR ranges visited in ascending position. The ranges are flattened over the warp's lanes.
struct SlRanges {
  int total;        // candidates in the window
  int start, len;   // this lane's column: CSR start and exclusive prefix / length
  int pre;
};

static __device__ __forceinline__ SlRanges sl_ranges(int min_cx, int min_cy, int nx, int ny, const int* __restrict__ off, int lane) {
  SlRanges r;
  r.start = 0; r.len = 0;
  if (lane < nx) {            // nx <= 64 columns; windows wider than 32 columns take a second sweep below
    const int c0 = (min_cx + lane) * GRID_ROWS + min_cy;
    r.start = off[c0];
    r.len = off[c0 + ny] - r.start;
  }
  int incl = r.len;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int v = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o) incl += v;
  }
  r.pre = incl - r.len;
  r.total = __shfl_sync(0xffffffffu, incl, 31);
  return r;
}

// CSR position of flat candidate t (t < total): the column whose prefix range holds t
static __device__ __forceinline__ int sl_position(const SlRanges& r, int t, int ncol) {
  int p = 0;
  for (int c = 0; c < ncol; ++c) {
    const int pre = __shfl_sync(0xffffffffu, r.pre, c), len = __shfl_sync(0xffffffffu, r.len, c), st = __shfl_sync(0xffffffffu, r.start, c);
    if (t >= pre && t < pre + len) p = st + (t - pre);
  }
  return p;
}

static __device__ __forceinline__ void sl_insert(unsigned int key, unsigned int& k0, unsigned int& k1, unsigned int& k2, unsigned int& k3) {
  if (key < k3) {
    k3 = key;
    if (k3 < k2) { const unsigned int t = k2; k2 = k3; k3 = t; }
    if (k2 < k1) { const unsigned int t = k1; k1 = k2; k2 = t; }
    if (k1 < k0) { const unsigned int t = k0; k0 = k1; k1 = t; }
  }
}

// one warp scans the window of one query: keys (distance << 16 | CSR position) of the candidates that pass the gates, are
// not locked and are no further than max_dist; every lane keeps its SL_K smallest sorted. Returns their number (warp total).
template <class W>
static __device__ __forceinline__ int win_scan(const W& w, int max_dist, const uint4 a0, const uint4 a1, const orb_keypoint* __restrict__ kp,
                                              const uint8_t* __restrict__ dc, const float* __restrict__ ur, const int* __restrict__ off,
                                              const unsigned short* __restrict__ idx, const unsigned char* lock, int lane, unsigned int& k0,
                                              unsigned int& k1, unsigned int& k2, unsigned int& k3, int* area = nullptr) {
  k0 = k1 = k2 = k3 = SL_NONE;
  int cnt = 0, narea = 0;
  if (w.ok) {                                                      // warp-uniform
    for (int cbase = 0; cbase < w.nx; cbase += 32) {
      const int ncol = min(32, w.nx - cbase);
      const SlRanges r = sl_ranges(w.min_cx + cbase, w.min_cy, ncol, w.ny, off, lane);
      for (int tb = 0; tb < r.total; tb += 32) {
        const int t = tb + lane;
        const int p = sl_position(r, t, ncol);
        if (t < r.total) {
          const int i2 = idx[p];
          if (!(lock && lock[i2])) {
            const orb_keypoint k = kp[i2];
            if (win_gate(w, k, ur ? ur[i2] : -1.f)) {
              narea++;
              const int d = hamming256(a0, a1, reinterpret_cast<const uint4*>(dc + (size_t)i2 * 32));
              if (d <= max_dist) {
                cnt++;
                sl_insert(((unsigned int)d << 16) | (unsigned int)p, k0, k1, k2, k3);
              }
            }
          }
        }
      }
    }
  }
  if (area) *area = __reduce_add_sync(0xffffffffu, narea);   // candidates that passed the window gates (no lock given: vIndices.size())
  return __reduce_add_sync(0xffffffffu, cnt);
}

// next smallest key of the warp (CSR positions are unique: exactly one lane pops)
static __device__ __forceinline__ unsigned int sl_pop(unsigned int& k0, unsigned int& k1, unsigned int& k2, unsigned int& k3) {
  const unsigned int m = __reduce_min_sync(0xffffffffu, k0);
  if (m != SL_NONE && k0 == m) { k0 = k1; k1 = k2; k2 = k3; k3 = SL_NONE; }
  return m;
}

__global__ void __launch_bounds__(SP_WARPS * 32, SP_MINB) k_sp_window(
    const orb_keypoint* __restrict__ kps, const uint8_t* __restrict__ desc, const float* __restrict__ uright, int kcap,
    const int* __restrict__ cell_off, const unsigned short* __restrict__ cell_idx, const orb_proj_query* __restrict__ queries,
    const uint8_t* __restrict__ qdesc, const int* __restrict__ nq_arr, int qcap, GridParams gp, OrbGeom g, float th,
    const float* __restrict__ tlc_z, float mb, int mono, float mbf, uint4* __restrict__ cand, unsigned char* __restrict__ cand_cnt,
    const unsigned char* __restrict__ locked0, int max_dist, int kf) {
  const int frame = blockIdx.y, lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const int qi = blockIdx.x * SP_WARPS + wid;
  if (qi >= min(nq_arr[frame], qcap)) return;
  const size_t qo = (size_t)frame * qcap + qi;
  const orb_proj_query q = queries[qo];
  const float tz = kf ? 0.f : tlc_z[frame];
  // bForward / bBackward (:1537-1538); kf: the KeyFrame overload (:1735-1842) always searches [level - 1, level + 1]
  const int mode = kf ? 2 + kf : ((tz > mb && !mono) ? 1 : ((-tz > mb && !mono) ? 2 : 0));   // kf 1 -> mode 3, kf 2 -> mode 4
  const SpWindow w = sp_window(q, gp, g.scale, th, mode, mbf);
  const unsigned short* idx = cell_idx + (size_t)frame * kcap;
  const uint4* qd = reinterpret_cast<const uint4*>(qdesc + qo * 32);
  unsigned int k0, k1, k2, k3;
  // only candidates with distance <= TH_HIGH (kf: <= ORBdist) can ever be assigned (:1610, :1805); keypoints that hold a map
  // point when the call starts (kf: locked0) never are
  const int cnt = win_scan(w, max_dist, qd[0], qd[1], kps + (size_t)frame * kcap, desc + (size_t)frame * kcap * 32,
                           uright ? uright + (size_t)frame * kcap : nullptr, cell_off + (size_t)frame * (GRID_CELLS + 1), idx,
                           locked0 ? locked0 + (size_t)frame * kcap : nullptr, lane, k0, k1, k2, k3);
  unsigned int mine = SL_NONE;
#pragma unroll
  for (int r = 0; r < SL_K; ++r) {
    const unsigned int m = sl_pop(k0, k1, k2, k3);
    if (lane == r) mine = m;
  }
  unsigned int rec = SL_NONE;
  if (lane < SL_K && mine != SL_NONE) rec = ((mine >> 16) << 20) | (unsigned int)idx[mine & 0xffffu];
  const unsigned int r1 = __shfl_sync(0xffffffffu, rec, 1), r2 = __shfl_sync(0xffffffffu, rec, 2), r3 = __shfl_sync(0xffffffffu, rec, 3);
  if (lane == 0) {
    cand[qo] = make_uint4(rec, r1, r2, r3);
    cand_cnt[qo] = (unsigned char)min(cnt, 255);
  }
}

// Resolver: the queries are replayed in order, 32 per round (see k_sl_resolve below for the argument): a lane's decision is
// the first stored candidate that is not locked, it stands unless an earlier lane of the round locks exactly that keypoint;
// the round commits the lanes before the first conflict. A query whose stored candidates are all locked while its window
// holds more is scanned again by the whole warp under the current locks when it is first in line. The rotation histogram
// and its records do not depend on the order (every record of a losing bin clears its keypoint and takes one match back).
// dynamic shared memory: assigned[kcap] i32 | cangle[kcap] f32 | recs[qcap] u32 | lock[kcap] u8
static size_t sp_resolve_smem(int qcap, int kcap) { return (size_t)kcap * 9 + (size_t)qcap * 4 + 32; }

__global__ void __launch_bounds__(128) k_sp_resolve(
    const orb_keypoint* __restrict__ kps, const uint8_t* __restrict__ desc, const float* __restrict__ uright, const int* __restrict__ n_arr,
    int kcap, const int* __restrict__ cell_off, const unsigned short* __restrict__ cell_idx, const orb_proj_query* __restrict__ queries,
    const uint8_t* __restrict__ qdesc, const int* __restrict__ nq_arr, int qcap, GridParams gp, OrbGeom g, float th,
    const float* __restrict__ tlc_z, float mb, int mono, float mbf, int check_orientation, const uint4* __restrict__ cand,
    const unsigned char* __restrict__ cand_cnt, int* __restrict__ match_out, int* __restrict__ nmatches_out,
    const unsigned char* __restrict__ locked0, int max_dist, int kf) {
  extern __shared__ __align__(16) unsigned char s_raw[];
  int* s_assigned = reinterpret_cast<int*>(s_raw);
  float* s_cangle = reinterpret_cast<float*>(s_assigned + kcap);
  unsigned int* s_recs = reinterpret_cast<unsigned int*>(s_cangle + kcap);
  unsigned char* s_lock = reinterpret_cast<unsigned char*>(s_recs + qcap);
  __shared__ uint4 s_cand[SL_CHUNK];
  __shared__ float s_qangle[SL_CHUNK];
  __shared__ unsigned char s_cnt[SL_CHUNK];
  __shared__ unsigned char s_obs[SL_CHUNK];
  __shared__ int s_hist[SP_HISTO];
  __shared__ int s_nm, s_nrec;
  const int frame = blockIdx.x, tid = threadIdx.x, lane = tid & 31;
  const int nC = min(n_arr[frame], kcap), nq = min(nq_arr[frame], qcap);
  const orb_keypoint* kp = kps + (size_t)frame * kcap;
  const orb_proj_query* q = queries + (size_t)frame * qcap;
  const uint4* cd = cand + (size_t)frame * qcap;
  const unsigned char* cc = cand_cnt + (size_t)frame * qcap;
  const unsigned short* idx = cell_idx + (size_t)frame * kcap;
  for (int i = tid; i < nC; i += 128) {
    s_cangle[i] = kp[i].angle; s_assigned[i] = -1;
    s_lock[i] = locked0 ? (locked0[(size_t)frame * kcap + i] ? 1 : 0) : 0;   // kf: CurrentFrame.mvpMapPoints[i2] != NULL at the start (:1793)
  }
  if (tid < SP_HISTO) s_hist[tid] = 0;
  const float tz = kf ? 0.f : tlc_z[frame];
  const int mode = kf ? 2 + kf : ((tz > mb && !mono) ? 1 : ((-tz > mb && !mono) ? 2 : 0));   // kf 1 -> mode 3, kf 2 -> mode 4
  const float factor = 1.0f / SP_HISTO;
  int nm = 0, nrec = 0;   // warp 0, uniform
  for (int base = 0; base < nq; base += SL_CHUNK) {
    const int m = min(SL_CHUNK, nq - base);
    __syncthreads();
    for (int i = tid; i < m; i += 128) {
      s_cand[i] = cd[base + i];
      s_cnt[i] = cc[base + i];
      s_qangle[i] = q[base + i].angle;
      s_obs[i] = (kf || (q[base + i].flags & 2)) ? 1 : 0;         // Observations() > 0; kf: every assignment locks (:1793)
    }
    __syncthreads();
    if (tid < 32) {
      int head = 0;
      while (head < m) {
        const int i = head + lane;
        int pick = -1, obs = 0;
        bool rescan = false;
        if (i < m) {
          const int cnt = s_cnt[i];
          obs = s_obs[i];
          if (cnt) {
            const uint4 c4 = s_cand[i];
            const unsigned int c[SL_K] = {c4.x, c4.y, c4.z, c4.w};
#pragma unroll
            for (int k = SL_K - 1; k >= 0; --k)
              if (c[k] != SL_NONE && !s_lock[c[k] & 0xffffu]) pick = (int)(c[k] & 0xffffu);   // :1592-1593, first unlocked
            if (pick < 0 && cnt > SL_K) rescan = true;
          }
        }
        const unsigned rs = __ballot_sync(0xffffffffu, rescan);
        int ncommit;
        if (rs & 1u) {
          // exact scan of the first query's window under the current locks by the whole warp
          const int qi = base + head;
          const SpWindow w = sp_window(q[qi], gp, g.scale, th, mode, mbf);
          const uint4* qd = reinterpret_cast<const uint4*>(qdesc + ((size_t)frame * qcap + qi) * 32);
          unsigned int k0, k1, k2, k3;
          win_scan(w, max_dist, qd[0], qd[1], kp, desc + (size_t)frame * kcap * 32, uright ? uright + (size_t)frame * kcap : nullptr,
                   cell_off + (size_t)frame * (GRID_CELLS + 1), idx, s_lock, lane, k0, k1, k2, k3);
          const unsigned int m1 = sl_pop(k0, k1, k2, k3);
          pick = (lane == 0 && m1 != SL_NONE) ? (int)idx[m1 & 0xffffu] : -1;
          ncommit = 1;
        } else {
          const int lockpick = (pick >= 0 && obs) ? pick : -1;
          bool bad = false;
#pragma unroll 8
          for (int k = 0; k < 31; ++k) {
            const int pk = __shfl_sync(0xffffffffu, lockpick, k);
            bad |= (k < lane) && (pk >= 0) && (pk == pick);
          }
          const unsigned stop = __ballot_sync(0xffffffffu, bad) | rs;
          ncommit = stop ? __ffs(stop) - 1 : 32;
        }
        const bool commit = lane < ncommit && pick >= 0;             // bestDist <= TH_HIGH (:1610)
        const unsigned cm = __ballot_sync(0xffffffffu, commit);
        __syncwarp();   // the lock reads of this round (all lanes) come before its lock writes
        if (commit) {
          const unsigned same = __match_any_sync(cm, pick);
          if (lane == 31 - __clz(same)) { s_assigned[pick] = base + i; s_lock[pick] = (unsigned char)obs; }
          if (check_orientation) {
            float rot = __fsub_rn(s_qangle[i], s_cangle[pick]);
            if (rot < 0.0f) rot = __fadd_rn(rot, 360.0f);
            int bin = (int)roundf(__fmul_rn(rot, factor));
            if (bin == SP_HISTO) bin = 0;
            s_recs[nrec + __popc(cm & ((1u << lane) - 1u))] = (unsigned int)pick | ((unsigned int)bin << 16);
            atomicAdd(&s_hist[bin], 1);
          }
        }
        nm += __popc(cm);
        nrec += __popc(cm);
        __syncwarp();
        head += ncommit;
      }
    }
  }
  if (tid == 0) { s_nm = nm; s_nrec = nrec; }
  __syncthreads();
  if (check_orientation && tid < 32) {
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
    int drop = 0;
    for (int r = lane; r < s_nrec; r += 32) {
      const int bin = (int)(s_recs[r] >> 16);
      if (bin != ind1 && bin != ind2 && bin != ind3) { s_assigned[s_recs[r] & 0xffffu] = -1; drop++; }
    }
    drop = __reduce_add_sync(0xffffffffu, drop);
    if (lane == 0) s_nm -= drop;
  }
  __syncthreads();
  for (int i = tid; i < kcap; i += 128) match_out[(size_t)frame * kcap + i] = i < nC ? s_assigned[i] : -1;
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
static __device__ __forceinline__ bool win_gate(const SlWindow& w, const orb_keypoint& k, float uright) {
  if (k.octave < w.min_level || k.octave > w.max_level) return false;
  const float distx = __fsub_rn(k.x, w.u), disty = __fsub_rn(k.y, w.v);
  if (!(fabsf(distx) < w.radius && fabsf(disty) < w.radius)) return false;
  if (uright > 0) {
    const float er = fabsf(__fsub_rn(w.xr, uright));
    if (er > w.radius) return false;
  }
  return true;
}

__global__ void __launch_bounds__(SP_WARPS * 32, SP_MINB) k_sl_window(
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
  const unsigned short* idx = cell_idx + (size_t)frame * kcap;
  const orb_keypoint* kp = kps + (size_t)frame * kcap;
  const uint4* qd = reinterpret_cast<const uint4*>(qdesc + qo * 32);
  unsigned int k0, k1, k2, k3;
  // :86-87 for the locks that exist before the call
  const int cnt = win_scan(w, 256, qd[0], qd[1], kp, desc + (size_t)frame * kcap * 32, uright ? uright + (size_t)frame * kcap : nullptr,
                          cell_off + (size_t)frame * (GRID_CELLS + 1), idx, locked0 ? locked0 + (size_t)frame * kcap : nullptr, lane, k0, k1, k2,
                          k3);
  unsigned int mine = SL_NONE;
#pragma unroll
  for (int r = 0; r < SL_K; ++r) {
    const unsigned int m = sl_pop(k0, k1, k2, k3);
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

// The resolver replays the map points in order, 32 at a time: every lane decides its map point against the locks as they
// stand at the start of the round; its decision only depends on the lock state of its best and second-best candidate
// (locks are only ever set, never cleared, and the stored candidates before them are locked already), so it stands
// unless an EARLIER lane of the round locks one of the two. The round commits the prefix of lanes before the first such
// conflict and the next round starts there (lane 0 is always right). A map point whose stored candidates run out while
// its window holds more is scanned again by the whole warp under the current locks when it is first in line.
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
  const int frame = blockIdx.x, tid = threadIdx.x, lane = tid & 31;
  const int nC = min(n_arr[frame], kcap), nq = min(nq_arr[frame], qcap);
  const orb_track_query* q = queries + (size_t)frame * qcap;
  const uint4* cd = cand + (size_t)frame * qcap;
  const unsigned char* cc = cand_cnt + (size_t)frame * qcap;
  const orb_keypoint* kp = kps + (size_t)frame * kcap;
  const unsigned short* idx = cell_idx + (size_t)frame * kcap;
  for (int i = tid; i < nC; i += 128) {
    s_assigned[i] = -1;
    s_lock[i] = locked0 ? locked0[(size_t)frame * kcap + i] : 0;
  }
  int nm = 0;   // warp 0, uniform
  for (int base = 0; base < nq; base += SL_CHUNK) {
    const int m = min(SL_CHUNK, nq - base);
    __syncthreads();
    for (int i = tid; i < m; i += 128) {
      s_cand[i] = cd[base + i];
      s_cnt[i] = cc[base + i];
      s_obs[i] = (q[base + i].flags & 2) ? 1 : 0;                 // Observations() > 0
    }
    __syncthreads();
    if (tid < 32) {
      int head = 0;
      while (head < m) {
        const int i = head + lane;
        // ---- every lane: decision of map point i under the current locks
        int pick = -1, dep1 = -1, dep2 = -1, obs = 0;
        bool rescan = false;
        if (i < m) {
          const int cnt = s_cnt[i];
          obs = s_obs[i];
          if (cnt) {
            const uint4 c4 = s_cand[i];
            const unsigned int c[SL_K] = {c4.x, c4.y, c4.z, c4.w};
            unsigned int b1 = SL_NONE, b2 = SL_NONE;
#pragma unroll
            for (int k = 0; k < SL_K; ++k) {
              const unsigned int key = c[k];
              if (key == SL_NONE || b2 != SL_NONE) continue;
              if (s_lock[key & 0xffffu]) continue;                // locked by an earlier map point of this call
              if (b1 == SL_NONE) b1 = key; else b2 = key;
            }
            // fewer than two unlocked stored candidates while the window holds more (all at least as far as the last
            // stored one): the exact scan is only needed when it can change the outcome - a best beyond TH_HIGH never
            // matches, and a best that passes the ratio against the last stored distance passes it against any later one
            bool more = b2 == SL_NONE && cnt > SL_K;
            if (more) {
              const int lastDist = (int)(c[SL_K - 1] >> 20);
              if (b1 == SL_NONE) { if (lastDist > SP_TH_HIGH) more = false; }
              else {
                const int bestDist = (int)(b1 >> 20);
                if (bestDist > SP_TH_HIGH || !((float)bestDist > __fmul_rn(nnratio, (float)lastDist))) more = false;
              }
            }
            if (more) rescan = true;
            else if (b1 != SL_NONE) {
              const int bestDist = (int)(b1 >> 20), bestLevel = (int)((b1 >> 16) & 15u);
              const int bestDist2 = b2 == SL_NONE ? 256 : (int)(b2 >> 20);
              const int bestLevel2 = b2 == SL_NONE ? -1 : (int)((b2 >> 16) & 15u);
              dep1 = (int)(b1 & 0xffffu);
              dep2 = b2 == SL_NONE ? -1 : (int)(b2 & 0xffffu);
              // :122-126: TH_HIGH, and the ratio only when best and second share the level; float * int -> float
              if (bestDist <= SP_TH_HIGH && !(bestLevel == bestLevel2 && (float)bestDist > __fmul_rn(nnratio, (float)bestDist2))) pick = dep1;
            }
          }
        }
        const unsigned rs = __ballot_sync(0xffffffffu, rescan);
        if (rs & 1u) {
          // the map point first in line needs the exact scan: the whole warp does it (:77-116 under the current locks)
          const int qi = base + head;
          const orb_track_query qq = q[qi];
          const SlWindow w = sl_window(qq, gp, g, th);
          const uint4* qd = reinterpret_cast<const uint4*>(qdesc + ((size_t)frame * qcap + qi) * 32);
          unsigned int k0, k1, k2, k3;
          win_scan(w, 256, qd[0], qd[1], kp, desc + (size_t)frame * kcap * 32, uright ? uright + (size_t)frame * kcap : nullptr,
                  cell_off + (size_t)frame * (GRID_CELLS + 1), idx, s_lock, lane, k0, k1, k2, k3);
          const unsigned int m1 = sl_pop(k0, k1, k2, k3), m2 = sl_pop(k0, k1, k2, k3);
          if (m1 != SL_NONE) {
            const int i1 = idx[m1 & 0xffffu];
            const int bestDist = (int)(m1 >> 16), bestLevel = kp[i1].octave;
            const int bestDist2 = m2 == SL_NONE ? 256 : (int)(m2 >> 16);
            const int bestLevel2 = m2 == SL_NONE ? -1 : kp[idx[m2 & 0xffffu]].octave;
            if (bestDist <= SP_TH_HIGH && !(bestLevel == bestLevel2 && (float)bestDist > __fmul_rn(nnratio, (float)bestDist2))) {
              if (lane == 0) { s_assigned[i1] = qi; s_lock[i1] = s_obs[head]; }
              nm++;
            }
          }
          __syncwarp();
          head += 1;
          continue;
        }
        // ---- conflicts: an earlier lane locks one of this lane's two candidates
        const int lockpick = (pick >= 0 && obs) ? pick : -1;
        bool bad = false;
#pragma unroll 8
        for (int k = 0; k < 31; ++k) {
          const int pk = __shfl_sync(0xffffffffu, lockpick, k);
          bad |= (k < lane) && (pk >= 0) && (pk == dep1 || pk == dep2);
        }
        const unsigned stop = __ballot_sync(0xffffffffu, bad) | rs;           // first conflict or first map point that needs the scan
        const int ncommit = stop ? __ffs(stop) - 1 : 32;                        // >= 1
        const bool commit = lane < ncommit && pick >= 0;
        const unsigned cm = __ballot_sync(0xffffffffu, commit);
        __syncwarp();   // the lock reads of this round (all lanes) come before its lock writes
        if (commit) {
          // several committed lanes on one keypoint: the earlier ones hold no lock, the last one stays (:127 overwrites)
          const unsigned same = __match_any_sync(cm, pick);
          if (lane == 31 - __clz(same)) { s_assigned[pick] = base + i; s_lock[pick] = (unsigned char)obs; }
        }
        nm += __popc(cm);
        __syncwarp();
        head += ncommit;
      }
    }
  }
  if (tid == 0) s_nm = nm;
  __syncthreads();
  for (int i = tid; i < kcap; i += 128) match_out[(size_t)frame * kcap + i] = i < nC ? s_assigned[i] : -1;
  if (tid == 0) nmatches_out[frame] = s_nm;
}

// ---- ORBmatcher::Fuse (both overloads), the search: reference src/ORBmatcher.cc:1131-1192 and :1277-1304 ----------------------
// One warp per map point: KeyFrame::GetFeaturesInArea (src/KeyFrame.cc:729-774: the window of Frame::GetFeaturesInArea without level
// arguments), the level gate [nPredictedLevel - 1, nPredictedLevel], the reprojection gates of the first overload, smallest
// (distance, visiting position) = the reference's strict "<" scan. No lock and no greedy assignment: the map surgery that follows
// in the reference (Replace / AddMapPoint) never changes a later search and stays with the caller.
struct FuLevels { float inv_sigma2[ORB_MAX_LEVELS]; };
struct FuWindow {
  int min_cx, min_cy, nx, ny;
  int level, mode;
  float u, v, ur, radius;
  const float* inv_sigma2;
  bool ok;
};
static __device__ __forceinline__ bool win_gate(const FuWindow& w, const orb_keypoint& k, float uright) {
  const float distx = __fsub_rn(k.x, w.u), disty = __fsub_rn(k.y, w.v);
  if (!(fabsf(distx) < w.radius && fabsf(disty) < w.radius)) return false;           // KeyFrame.cc:765-768
  if (k.octave < w.level - 1 || k.octave > w.level) return false;                    // :1147 / :1288
  if (w.mode == 0) {
    const float ex = __fsub_rn(w.u, k.x), ey = __fsub_rn(w.v, k.y);
    if (uright >= 0) {                                                               // :1149-1160
      const float er = __fsub_rn(w.ur, uright);
      const float e2 = __fadd_rn(__fadd_rn(__fmul_rn(ex, ex), __fmul_rn(ey, ey)), __fmul_rn(er, er));
      if ((double)__fmul_rn(e2, w.inv_sigma2[k.octave]) > 7.8) return false;
    } else {                                                                         // :1161-1169
      const float e2 = __fadd_rn(__fmul_rn(ex, ex), __fmul_rn(ey, ey));
      if ((double)__fmul_rn(e2, w.inv_sigma2[k.octave]) > 5.99) return false;
    }
  }
  return true;
}

__global__ void __launch_bounds__(SP_WARPS * 32, SP_MINB) k_fuse_search(
    const orb_keypoint* __restrict__ kps, const uint8_t* __restrict__ desc, const float* __restrict__ uright, int kcap,
    const int* __restrict__ cell_off, const unsigned short* __restrict__ cell_idx, const orb_fuse_query* __restrict__ queries,
    const uint8_t* __restrict__ qdesc, const int* __restrict__ nq_arr, int qcap, GridParams gp, OrbGeom g, FuLevels lv, float th, int mode,
    int* __restrict__ best_idx, int* __restrict__ best_dist) {
  const int frame = blockIdx.y, lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const int qi = blockIdx.x * SP_WARPS + wid;
  if (qi >= qcap) return;
  const size_t qo = (size_t)frame * qcap + qi;
  if (qi >= nq_arr[frame]) {
    if (lane == 0) { best_idx[qo] = -1; best_dist[qo] = 256; }
    return;
  }
  const orb_fuse_query q = queries[qo];
  FuWindow w;
  w.ok = false;
  w.level = q.level; w.mode = mode; w.u = q.u; w.v = q.v; w.ur = q.ur; w.inv_sigma2 = lv.inv_sigma2;
  w.min_cx = w.min_cy = w.nx = w.ny = 0; w.radius = 0.f;
  if ((q.flags & 1) && q.level >= 0 && q.level < g.nlevels) {
    const float r = __fmul_rn(th, g.scale[q.level]);                                   // :1131
    w.radius = r;
    const int minx = max(0, (int)floorf(__fmul_rn(__fsub_rn(__fsub_rn(q.u, gp.min_x), r), gp.w_inv)));
    const int maxx = min(GRID_COLS - 1, (int)ceilf(__fmul_rn(__fadd_rn(__fsub_rn(q.u, gp.min_x), r), gp.w_inv)));
    const int miny = max(0, (int)floorf(__fmul_rn(__fsub_rn(__fsub_rn(q.v, gp.min_y), r), gp.h_inv)));
    const int maxy = min(GRID_ROWS - 1, (int)ceilf(__fmul_rn(__fadd_rn(__fsub_rn(q.v, gp.min_y), r), gp.h_inv)));
    if (minx < GRID_COLS && maxx >= 0 && miny < GRID_ROWS && maxy >= 0) {
      w.min_cx = minx; w.min_cy = miny; w.nx = maxx - minx + 1; w.ny = maxy - miny + 1;
      w.ok = w.nx > 0 && w.ny > 0;
    }
  }
  const unsigned short* idx = cell_idx + (size_t)frame * kcap;
  const uint4* qd = reinterpret_cast<const uint4*>(qdesc + qo * 32);
  unsigned int k0, k1, k2, k3;
  // int bestDist = 256 with "dist < bestDist" (:1139, :1183): every distance up to 255 can win
  win_scan(w, 255, qd[0], qd[1], kps + (size_t)frame * kcap, desc + (size_t)frame * kcap * 32,
           uright ? uright + (size_t)frame * kcap : nullptr, cell_off + (size_t)frame * (GRID_CELLS + 1), idx, nullptr, lane, k0, k1, k2, k3);
  const unsigned int m = __reduce_min_sync(0xffffffffu, k0);
  if (lane == 0) {
    best_idx[qo] = m == SL_NONE ? -1 : (int)idx[m & 0xffffu];
    best_dist[qo] = m == SL_NONE ? 256 : (int)(m >> 16);
  }
}

// ---- ORBmatcher::SearchForInitialization(F1, F2, vbPrevMatched, vnMatches12, windowSize): reference src/ORBmatcher.cc:603-700 ----
// The monocular initialiser's matcher: every level-0 keypoint i1 of the initial frame F1 looks at the level-0 keypoints of the
// current frame F2 inside a square window around vbPrevMatched[i1]. Unlike the other window searches the sequential state is a
// DISTANCE per F2 keypoint: a candidate is skipped when an earlier i1 already holds it at a distance <= this one (:638), a better
// i1 steals it (:651-655). One CTA per frame:
//   phase 1 (all warps, i1 in parallel)  the window's candidates in GetFeaturesInArea order with their Hamming distances ->
//                                        list[i1][SFI_CAP] (distance << 16 | i2) + the true count
//   phase 2 (warp 0, i1 in order)        eligibility against vMatchedDistance, best / second best as the two smallest keys
//                                        (distance << 16 | visiting position: the strict "<" scan keeps the first minimum, the second
//                                        best is the second smallest value of the multiset), TH_LOW, ratio in float, steal, rotation bin;
//                                        a query whose window held more than SFI_CAP candidates is scanned again on the fly (exact)
//   then ComputeThreeMaxima, the removal of the losing bins (records of stolen matches included, like the reference's vectors)
//   and the update of vbPrevMatched (:694-697).
#define SFI_CAP 64
#define SFI_THREADS 256
#define SFI_TH_LOW 50     // ORBmatcher::TH_LOW (src/ORBmatcher.cc:36)
static size_t sfi_smem(int qcap, int kcap) { return (size_t)kcap * 6 + (size_t)qcap * 5 + 64; }

struct SfiWindow { int min_cx, min_cy, nx, ny; float x, y, r; bool ok; };

static __device__ __forceinline__ SfiWindow sfi_window(const orb_init_query& q, const GridParams& gp, float r) {
  SfiWindow w;
  w.ok = false; w.x = q.x; w.y = q.y; w.r = r;
  w.min_cx = w.min_cy = w.nx = w.ny = 0;
  if (q.octave > 0) return w;                                                                    // level1 > 0 (:621)
  const int minx = max(0, (int)floorf(__fmul_rn(__fsub_rn(__fsub_rn(q.x, gp.min_x), r), gp.w_inv)));
  if (minx >= GRID_COLS) return w;
  const int maxx = min(GRID_COLS - 1, (int)ceilf(__fmul_rn(__fadd_rn(__fsub_rn(q.x, gp.min_x), r), gp.w_inv)));
  if (maxx < 0) return w;
  const int miny = max(0, (int)floorf(__fmul_rn(__fsub_rn(__fsub_rn(q.y, gp.min_y), r), gp.h_inv)));
  if (miny >= GRID_ROWS) return w;
  const int maxy = min(GRID_ROWS - 1, (int)ceilf(__fmul_rn(__fadd_rn(__fsub_rn(q.y, gp.min_y), r), gp.h_inv)));
  if (maxy < 0) return w;
  w.min_cx = minx; w.min_cy = miny; w.nx = maxx - minx + 1; w.ny = maxy - miny + 1;
  w.ok = w.nx > 0 && w.ny > 0;
  return w;
}

// One warp walks the window of one query in GetFeaturesInArea order, 32 CSR entries per trip, and hands every trip to `sink`
// (pass: this lane's entry is a candidate; i2, d: its keypoint and Hamming distance). Candidates of a trip are in lane order.
template <class Sink>
static __device__ __forceinline__ void sfi_scan(const SfiWindow& w, int level, const uint4 a0, const uint4 a1, const orb_keypoint* __restrict__ kp,
                                                const uint8_t* __restrict__ dc, const int* __restrict__ off,
                                                const unsigned short* __restrict__ idx, int lane, Sink sink) {
  if (!w.ok) return;
  for (int cbase = 0; cbase < w.nx; cbase += 32) {
    const int ncol = min(32, w.nx - cbase);
    const SlRanges r = sl_ranges(w.min_cx + cbase, w.min_cy, ncol, w.ny, off, lane);
    for (int tb = 0; tb < r.total; tb += 32) {
      const int t = tb + lane;
      const int p = sl_position(r, t, ncol);
      bool pass = false;
      int i2 = 0, d = 0;
      if (t < r.total) {
        i2 = idx[p];
        const orb_keypoint k = kp[i2];
        // minLevel = maxLevel = level1 = 0: bCheckLevels holds (maxLevel >= 0), octave < 0 never, octave > 0 is skipped (Frame.cc:787-799)
        const float distx = __fsub_rn(k.x, w.x), disty = __fsub_rn(k.y, w.y);
        pass = !(k.octave < level) && !(k.octave > level) && fabsf(distx) < w.r && fabsf(disty) < w.r;
        if (pass) d = hamming256(a0, a1, reinterpret_cast<const uint4*>(dc + (size_t)i2 * 32));
      }
      sink(pass, i2, d);
    }
  }
}

__global__ void __launch_bounds__(SFI_THREADS) k_search_for_initialization(
    const orb_keypoint* __restrict__ kps, const uint8_t* __restrict__ desc, const int* __restrict__ n_arr, int kcap,
    const int* __restrict__ cell_off, const unsigned short* __restrict__ cell_idx, const orb_init_query* __restrict__ queries,
    const uint8_t* __restrict__ qdesc, const int* __restrict__ nq_arr, int qcap, GridParams gp, float radius, float nnratio,
    int check_orientation, unsigned int* __restrict__ list, int* __restrict__ list_cnt, int* __restrict__ match12_out,
    float2* __restrict__ prev_out, int* __restrict__ nmatches_out) {
  extern __shared__ __align__(16) unsigned char s_raw[];
  int* s_m21 = reinterpret_cast<int*>(s_raw);                                   // vnMatches21
  int* s_m12 = s_m21 + kcap;                                                    // vnMatches12
  unsigned short* s_md = reinterpret_cast<unsigned short*>(s_m12 + qcap);       // vMatchedDistance (0xffff = INT_MAX)
  unsigned char* s_bin = reinterpret_cast<unsigned char*>(s_md + kcap);         // rotation bin of i1's record (0xff: none)
  __shared__ int s_hist[SP_HISTO];
  __shared__ int s_nm;
  const int frame = blockIdx.x, tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const int n2 = min(n_arr[frame], kcap), nq = min(nq_arr[frame], qcap);
  const orb_keypoint* kp = kps + (size_t)frame * kcap;
  const uint8_t* dc = desc + (size_t)frame * kcap * 32;
  const int* off = cell_off + (size_t)frame * (GRID_CELLS + 1);
  const unsigned short* idx = cell_idx + (size_t)frame * kcap;
  const orb_init_query* q = queries + (size_t)frame * qcap;
  unsigned int* fl = list + (size_t)frame * qcap * SFI_CAP;
  int* fc = list_cnt + (size_t)frame * qcap;
  for (int i = tid; i < n2; i += SFI_THREADS) { s_m21[i] = -1; s_md[i] = 0xffffu; }
  for (int i = tid; i < qcap; i += SFI_THREADS) { s_m12[i] = -1; s_bin[i] = 0xffu; }
  if (tid < SP_HISTO) s_hist[tid] = 0;

  // ---- phase 1: candidate lists
  for (int i1 = wid; i1 < nq; i1 += SFI_THREADS / 32) {
    const SfiWindow w = sfi_window(q[i1], gp, radius);
    const uint4* qd = reinterpret_cast<const uint4*>(qdesc + ((size_t)frame * qcap + i1) * 32);
    unsigned int* my = fl + (size_t)i1 * SFI_CAP;
    int cnt = 0;
    sfi_scan(w, 0, qd[0], qd[1], kp, dc, off, idx, lane, [&](bool pass, int i2, int d) {
      const unsigned bal = __ballot_sync(0xffffffffu, pass);
      const int pos = cnt + __popc(bal & ((1u << lane) - 1u));
      if (pass && pos < SFI_CAP) my[pos] = ((unsigned int)d << 16) | (unsigned int)i2;
      cnt += __popc(bal);
    });
    if (lane == 0) fc[i1] = cnt;
  }
  __syncthreads();   // one CTA per frame: the lists are this block's own global writes

  // ---- phase 2: the reference's loop over i1, in order
  if (wid == 0) {
    int nm = 0;
    for (int i1 = 0; i1 < nq; ++i1) {
      const int cnt = fc[i1];
      if (cnt == 0) continue;                                                   // level1 > 0 or vIndices2.empty() (:621-626)
      unsigned int k1 = SL_NONE, k2 = SL_NONE;
      int j1 = -1, j2 = -1;
      auto take = [&](bool pass, int i2, int d, int seq) {
        if (pass && !((int)s_md[i2] <= d)) {                                    // if (vMatchedDistance[i2] <= dist) continue; (:638)
          const unsigned int key = ((unsigned int)d << 16) | (unsigned int)seq;
          if (key < k1) { k2 = k1; j2 = j1; k1 = key; j1 = i2; }
          else if (key < k2) { k2 = key; j2 = i2; }
        }
      };
      if (cnt <= SFI_CAP) {
        const unsigned int* my = fl + (size_t)i1 * SFI_CAP;
        for (int t = lane; t < cnt; t += 32) {
          const unsigned int e = my[t];
          take(true, (int)(e & 0xffffu), (int)(e >> 16), t);
        }
      } else {
        const SfiWindow w = sfi_window(q[i1], gp, radius);
        const uint4* qd = reinterpret_cast<const uint4*>(qdesc + ((size_t)frame * qcap + i1) * 32);
        int seq = 0;
        sfi_scan(w, 0, qd[0], qd[1], kp, dc, off, idx, lane, [&](bool pass, int i2, int d) {
          take(pass, i2, d, seq + lane);
          seq += 32;
        });
      }
      const unsigned int m1 = __reduce_min_sync(0xffffffffu, k1);
      if (m1 == SL_NONE) continue;                                              // bestDist = INT_MAX
      const bool own = k1 == m1;                                                // visiting positions are unique: one lane
      const int best = __shfl_sync(0xffffffffu, j1, __ffs(__ballot_sync(0xffffffffu, own)) - 1);
      if (own) { k1 = k2; j1 = j2; }
      const unsigned int m2 = __reduce_min_sync(0xffffffffu, k1);
      const int bestDist = (int)(m1 >> 16);
      const int bestDist2 = m2 == SL_NONE ? 2147483647 : (int)(m2 >> 16);
      if (bestDist <= SFI_TH_LOW && (float)bestDist < __fmul_rn((float)bestDist2, nnratio)) {   // :651-652
        if (lane == 0) {
          const int old = s_m21[best];
          if (old >= 0) { s_m12[old] = -1; nm--; }                              // :653-656
          s_m12[i1] = best; s_m21[best] = i1; s_md[best] = (unsigned short)bestDist;
          nm++;
          if (check_orientation) {
            float rot = __fsub_rn(q[i1].angle, kp[best].angle);                 // :663-668
            if (rot < 0.0f) rot = __fadd_rn(rot, 360.0f);
            int bin = (int)roundf(__fmul_rn(rot, 1.0f / SP_HISTO));
            if (bin == SP_HISTO) bin = 0;
            s_bin[i1] = (unsigned char)bin;
            s_hist[bin]++;
          }
        }
        __syncwarp();
      }
    }
    if (lane == 0) s_nm = nm;
    __syncwarp();
    if (check_orientation) {
      int ind1 = -1, ind2 = -1, ind3 = -1, max1 = 0, max2 = 0, max3 = 0;        // ComputeThreeMaxima (:1844-1876)
      for (int i = 0; i < SP_HISTO; i++) {
        const int s = s_hist[i];
        if (s > max1) { max3 = max2; max2 = max1; max1 = s; ind3 = ind2; ind2 = ind1; ind1 = i; }
        else if (s > max2) { max3 = max2; max2 = s; ind3 = ind2; ind2 = i; }
        else if (s > max3) { max3 = s; ind3 = i; }
      }
      if ((float)max2 < __fmul_rn(0.1f, (float)max1)) { ind2 = -1; ind3 = -1; }
      else if ((float)max3 < __fmul_rn(0.1f, (float)max1)) { ind3 = -1; }
      int drop = 0;
      for (int i1 = lane; i1 < nq; i1 += 32) {
        const int bin = s_bin[i1];
        // records of a losing bin: only those that still hold a match count (:684-687)
        if (bin != 0xff && bin != ind1 && bin != ind2 && bin != ind3 && s_m12[i1] >= 0) { s_m12[i1] = -1; drop++; }
      }
      drop = __reduce_add_sync(0xffffffffu, drop);
      if (lane == 0) s_nm -= drop;
    }
  }
  __syncthreads();
  for (int i = tid; i < qcap; i += SFI_THREADS) {
    const int m = i < nq ? s_m12[i] : -1;
    match12_out[(size_t)frame * qcap + i] = m;
    float2 p = make_float2(0.f, 0.f);
    if (i < nq) p = m >= 0 ? make_float2(kp[m].x, kp[m].y) : make_float2(q[i].x, q[i].y);   // :694-697
    prev_out[(size_t)frame * qcap + i] = p;
  }
  if (tid == 0) nmatches_out[frame] = s_nm;
}

// ---- ORBmatcher::SearchByBoW(KeyFrame *pKF, Frame &F, vector<MapPoint*> &vpMapPointMatches) ----------------------------------
// reference src/ORBmatcher.cc:218-395 (single camera): keyframe and frame keypoints that fall into the same vocabulary node
// are compared all against all; a keyframe keypoint with a map point takes the best frame keypoint of the node that no
// earlier one took (TH_LOW, ratio against the second best), then the rotation histogram. A frame keypoint belongs to
// exactly one node, so the nodes are independent: one warp per shared node, keyframe keypoints of the node in list order
// (that order decides who takes a contested keypoint), the node's frame keypoints spread over the lanes, best / second
// best by two redux.min rounds on (distance << 16 | position in the node's list) = the reference's strict "<" scan.
#define SBOW_TH_LOW 50    // ORBmatcher::TH_LOW (src/ORBmatcher.cc:36)

// dynamic shared memory: match[kcap] i32 | recs[kcap] u32
__global__ void __launch_bounds__(256) k_sbow(const orb_keypoint* __restrict__ kps, const uint8_t* __restrict__ desc, const int* __restrict__ n_arr,
                                              int kcap, const unsigned int* __restrict__ f_node, const int* __restrict__ f_off,
                                              const unsigned int* __restrict__ f_feat, const int* __restrict__ f_nn,
                                              const uint8_t* __restrict__ kf_desc, const float* __restrict__ kf_angle,
                                              const uint8_t* __restrict__ kf_flags, const int* __restrict__ kf_n, int qcap,
                                              const unsigned int* __restrict__ kf_node, const int* __restrict__ kf_off,
                                              const unsigned int* __restrict__ kf_feat, const int* __restrict__ kf_nn, float nnratio,
                                              int check_orientation, int* __restrict__ match_out, int* __restrict__ nmatches_out) {
  extern __shared__ __align__(16) unsigned char s_raw[];
  int* s_match = reinterpret_cast<int*>(s_raw);
  unsigned int* s_recs = reinterpret_cast<unsigned int*>(s_match + kcap);
  __shared__ int s_hist[SP_HISTO];
  __shared__ int s_nrec, s_nm, s_ind[3];
  const int frame = blockIdx.x, tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const int nF = min(n_arr[frame], kcap), nK = min(kf_n[frame], qcap);
  const int nnF = f_nn[frame], nnK = min(kf_nn[frame], qcap);
  const orb_keypoint* kp = kps + (size_t)frame * kcap;
  const uint8_t* dF = desc + (size_t)frame * kcap * 32;
  const unsigned int* fn = f_node + (size_t)frame * kcap;
  const int* fo = f_off + (size_t)frame * (kcap + 1);
  const unsigned int* ff = f_feat + (size_t)frame * kcap;
  const uint8_t* dK = kf_desc + (size_t)frame * qcap * 32;
  const float* aK = kf_angle + (size_t)frame * qcap;
  const uint8_t* flK = kf_flags + (size_t)frame * qcap;
  const unsigned int* kn = kf_node + (size_t)frame * qcap;
  const int* ko = kf_off + (size_t)frame * (qcap + 1);
  const unsigned int* kf = kf_feat + (size_t)frame * qcap;
  for (int i = tid; i < kcap; i += 256) s_match[i] = -1;
  if (tid < SP_HISTO) s_hist[tid] = 0;
  if (tid == 0) { s_nrec = 0; s_nm = 0; }
  __syncthreads();
  const float factor = 1.0f / SP_HISTO;
  int nm = 0;                                     // per warp, uniform
  for (int a = wid; a < nnK; a += 8) {
    // the frame's node with the same id (both lists ascend: the reference's merge with lower_bound visits exactly the equal ids)
    const unsigned int node = kn[a];
    int lo = 0, hi = nnF;
    while (lo < hi) { const int mid = (lo + hi) >> 1; if (fn[mid] < node) lo = mid + 1; else hi = mid; }
    if (lo >= nnF || fn[lo] != node) continue;
    const int f0 = fo[lo], f1 = fo[lo + 1];
    for (int t = ko[a]; t < ko[a + 1]; ++t) {
      const int iKF = (int)kf[t];
      if (iKF >= nK || !flK[iKF]) continue;       // no map point / bad map point (:246-250)
      const uint4* qd = reinterpret_cast<const uint4*>(dK + (size_t)iKF * 32);
      const uint4 a0 = qd[0], a1 = qd[1];
      unsigned int k0 = SL_NONE, k1 = SL_NONE;    // this lane's two smallest keys
      for (int u = f0 + lane; u < f1; u += 32) {
        const int iF = (int)ff[u];
        if (iF >= nF || s_match[iF] >= 0) continue;                 // already holds a map point (:266)
        const unsigned int d = (unsigned int)hamming256(a0, a1, reinterpret_cast<const uint4*>(dF + (size_t)iF * 32));
        const unsigned int key = (d << 16) | (unsigned int)(u - f0);
        if (key < k0) { k1 = k0; k0 = key; } else if (key < k1) k1 = key;
      }
      const unsigned int m1 = __reduce_min_sync(0xffffffffu, k0);
      if (m1 == SL_NONE) continue;                // warp-uniform
      if (k0 == m1) { k0 = k1; k1 = SL_NONE; }    // positions are unique: one lane pops
      const unsigned int m2 = __reduce_min_sync(0xffffffffu, k0);
      const int bestDist1 = (int)(m1 >> 16), bestDist2 = m2 == SL_NONE ? 256 : (int)(m2 >> 16);
      __syncwarp();                               // the sweep's reads of s_match come before lane 0's write
      if (bestDist1 <= SBOW_TH_LOW && (float)bestDist1 < __fmul_rn(nnratio, (float)bestDist2)) {   // :305-307
        const int bestIdxF = (int)ff[f0 + (int)(m1 & 0xffffu)];
        if (lane == 0) {
          s_match[bestIdxF] = iKF;
          if (check_orientation) {
            float rot = __fsub_rn(aK[iKF], kp[bestIdxF].angle);     // kp.angle - Fkp.angle (:320)
            if (rot < 0.0f) rot = __fadd_rn(rot, 360.0f);
            int bin = (int)roundf(__fmul_rn(rot, factor));
            if (bin == SP_HISTO) bin = 0;
            s_recs[atomicAdd(&s_nrec, 1)] = (unsigned int)bestIdxF | ((unsigned int)bin << 16);
            atomicAdd(&s_hist[bin], 1);
          }
        }
        nm++;
      }
      __syncwarp();                               // the lock is visible to the lanes before the next keyframe keypoint
    }
  }
  if (lane == 0 && nm) atomicAdd(&s_nm, nm);
  __syncthreads();
  if (check_orientation) {
    if (tid == 0) {
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
      s_ind[0] = ind1; s_ind[1] = ind2; s_ind[2] = ind3;
    }
    __syncthreads();
    int drop = 0;
    for (int r = tid; r < s_nrec; r += 256) {
      const int bin = (int)(s_recs[r] >> 16);
      if (bin != s_ind[0] && bin != s_ind[1] && bin != s_ind[2]) { s_match[s_recs[r] & 0xffffu] = -1; drop++; }   // :376-381
    }
    if (drop) atomicSub(&s_nm, drop);
    __syncthreads();
  }
  for (int i = tid; i < kcap; i += 256) match_out[(size_t)frame * kcap + i] = i < nF ? s_match[i] : -1;
  if (tid == 0) nmatches_out[frame] = s_nm;
}

// ---- SearchByBoW against a two-camera frame (src/ORBmatcher.cc:218-395 with F.Nleft != -1) ---------------------------------------
// The frame's FeatureVector is the one of orb_compute_bow_stereo: feature i < nL is left keypoint i, feature nL + j right keypoint j.
// Per keyframe keypoint the node's frame features are swept once, a best / second best is kept for the left camera and a best for
// the right one (:283-301); the right-camera match needs the LEFT best below TH_LOW and takes no ratio test of its own (`|| true`,
// :331-334). Nodes stay independent (a feature of the combined frame belongs to one node), so the warp-per-node scheme of k_sbow holds.
// dynamic shared memory: match[cap2] i32 | recs[2 cap2] u32
__global__ void __launch_bounds__(256) k_sbow2(const orb_keypoint* __restrict__ kpsL, const uint8_t* __restrict__ descL, const int* __restrict__ nL_arr,
                                               int kL, const orb_keypoint* __restrict__ kpsR, const uint8_t* __restrict__ descR,
                                               const int* __restrict__ nR_arr, int kR, const unsigned int* __restrict__ f_node,
                                               const int* __restrict__ f_off, const unsigned int* __restrict__ f_feat, const int* __restrict__ f_nn,
                                               const uint8_t* __restrict__ kf_desc, const float* __restrict__ kf_angle,
                                               const uint8_t* __restrict__ kf_flags, const int* __restrict__ kf_n, int qcap,
                                               const unsigned int* __restrict__ kf_node, const int* __restrict__ kf_off,
                                               const unsigned int* __restrict__ kf_feat, const int* __restrict__ kf_nn, float nnratio,
                                               int check_orientation, int* __restrict__ match_left, int* __restrict__ match_right,
                                               int* __restrict__ nmatches_out) {
  extern __shared__ __align__(16) unsigned char s_raw[];
  const int cap2 = kL + kR;
  int* s_match = reinterpret_cast<int*>(s_raw);
  unsigned int* s_recs = reinterpret_cast<unsigned int*>(s_match + cap2);
  __shared__ int s_hist[SP_HISTO];
  __shared__ int s_nrec, s_nm, s_ind[3];
  const int frame = blockIdx.x, tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const int nL = min(nL_arr[frame], kL), nR = min(nR_arr[frame], kR), nF = nL + nR, nK = min(kf_n[frame], qcap);
  const int nnF = f_nn[frame], nnK = min(kf_nn[frame], qcap);
  const orb_keypoint* kpL = kpsL + (size_t)frame * kL;
  const orb_keypoint* kpR = kpsR + (size_t)frame * kR;
  const uint8_t* dL = descL + (size_t)frame * kL * 32;
  const uint8_t* dR = descR + (size_t)frame * kR * 32;
  const unsigned int* fn = f_node + (size_t)frame * cap2;
  const int* fo = f_off + (size_t)frame * (cap2 + 1);
  const unsigned int* ff = f_feat + (size_t)frame * cap2;
  const uint8_t* dK = kf_desc + (size_t)frame * qcap * 32;
  const float* aK = kf_angle + (size_t)frame * qcap;
  const uint8_t* flK = kf_flags + (size_t)frame * qcap;
  const unsigned int* kn = kf_node + (size_t)frame * qcap;
  const int* ko = kf_off + (size_t)frame * (qcap + 1);
  const unsigned int* kf = kf_feat + (size_t)frame * qcap;
  for (int i = tid; i < cap2; i += 256) s_match[i] = -1;
  if (tid < SP_HISTO) s_hist[tid] = 0;
  if (tid == 0) { s_nrec = 0; s_nm = 0; }
  __syncthreads();
  const float factor = 1.0f / SP_HISTO;
  int nm = 0;                                     // per warp, uniform
  auto record = [&](int idxF, float angleK, float angleF) {
    float rot = __fsub_rn(angleK, angleF);
    if (rot < 0.0f) rot = __fadd_rn(rot, 360.0f);
    int bin = (int)roundf(__fmul_rn(rot, factor));
    if (bin == SP_HISTO) bin = 0;
    s_recs[atomicAdd(&s_nrec, 1)] = (unsigned int)idxF | ((unsigned int)bin << 16);
    atomicAdd(&s_hist[bin], 1);
  };
  for (int a = wid; a < nnK; a += 8) {
    const unsigned int node = kn[a];
    int lo = 0, hi = nnF;
    while (lo < hi) { const int mid = (lo + hi) >> 1; if (fn[mid] < node) lo = mid + 1; else hi = mid; }
    if (lo >= nnF || fn[lo] != node) continue;
    const int f0 = fo[lo], f1 = fo[lo + 1];
    for (int t = ko[a]; t < ko[a + 1]; ++t) {
      const int iKF = (int)kf[t];
      if (iKF >= nK || !flK[iKF]) continue;
      const uint4* qd = reinterpret_cast<const uint4*>(dK + (size_t)iKF * 32);
      const uint4 a0 = qd[0], a1 = qd[1];
      unsigned int k0 = SL_NONE, k1 = SL_NONE, r0 = SL_NONE;    // left: two smallest keys of this lane; right: the smallest
      for (int u = f0 + lane; u < f1; u += 32) {
        const int iF = (int)ff[u];
        if (iF >= nF || s_match[iF] >= 0) continue;                 // already holds a map point (:266 / :283)
        const uint8_t* dF = iF < nL ? dL + (size_t)iF * 32 : dR + (size_t)(iF - nL) * 32;
        const unsigned int d = (unsigned int)hamming256(a0, a1, reinterpret_cast<const uint4*>(dF));
        const unsigned int key = (d << 16) | (unsigned int)(u - f0);
        if (iF < nL) { if (key < k0) { k1 = k0; k0 = key; } else if (key < k1) k1 = key; }
        else if (key < r0) r0 = key;
      }
      const unsigned int m1 = __reduce_min_sync(0xffffffffu, k0);
      const unsigned int mr = __reduce_min_sync(0xffffffffu, r0);
      if (m1 == SL_NONE) continue;                // bestDist1 = 256 > TH_LOW: neither camera matches (:305)
      if (k0 == m1) { k0 = k1; k1 = SL_NONE; }
      const unsigned int m2 = __reduce_min_sync(0xffffffffu, k0);
      const int bestDist1 = (int)(m1 >> 16), bestDist2 = m2 == SL_NONE ? 256 : (int)(m2 >> 16);
      __syncwarp();                               // the sweep's reads of s_match come before lane 0's writes
      if (bestDist1 <= SBOW_TH_LOW) {
        if ((float)bestDist1 < __fmul_rn(nnratio, (float)bestDist2)) {
          const int bestIdxF = (int)ff[f0 + (int)(m1 & 0xffffu)];
          if (lane == 0) {
            s_match[bestIdxF] = iKF;
            if (check_orientation) record(bestIdxF, aK[iKF], kpL[bestIdxF].angle);
          }
          nm++;
        }
        if (mr != SL_NONE && (int)(mr >> 16) <= SBOW_TH_LOW) {
          const int bestIdxFR = (int)ff[f0 + (int)(mr & 0xffffu)];
          if (lane == 0) {
            s_match[bestIdxFR] = iKF;
            if (check_orientation) record(bestIdxFR, aK[iKF], kpR[bestIdxFR - nL].angle);
          }
          nm++;
        }
      }
      __syncwarp();
    }
  }
  if (lane == 0 && nm) atomicAdd(&s_nm, nm);
  __syncthreads();
  if (check_orientation) {
    if (tid == 0) {
      int ind1 = -1, ind2 = -1, ind3 = -1, max1 = 0, max2 = 0, max3 = 0;
      for (int i = 0; i < SP_HISTO; i++) {
        const int sz = s_hist[i];
        if (sz > max1) { max3 = max2; max2 = max1; max1 = sz; ind3 = ind2; ind2 = ind1; ind1 = i; }
        else if (sz > max2) { max3 = max2; max2 = sz; ind3 = ind2; ind2 = i; }
        else if (sz > max3) { max3 = sz; ind3 = i; }
      }
      if ((float)max2 < __fmul_rn(0.1f, (float)max1)) { ind2 = -1; ind3 = -1; }
      else if ((float)max3 < __fmul_rn(0.1f, (float)max1)) { ind3 = -1; }
      s_ind[0] = ind1; s_ind[1] = ind2; s_ind[2] = ind3;
    }
    __syncthreads();
    int drop = 0;
    for (int r = tid; r < s_nrec; r += 256) {
      const int bin = (int)(s_recs[r] >> 16);
      if (bin != s_ind[0] && bin != s_ind[1] && bin != s_ind[2]) { s_match[s_recs[r] & 0xffffu] = -1; drop++; }
    }
    if (drop) atomicSub(&s_nm, drop);
    __syncthreads();
  }
  for (int i = tid; i < kL; i += 256) match_left[(size_t)frame * kL + i] = i < nL ? s_match[i] : -1;
  for (int i = tid; i < kR; i += 256) match_right[(size_t)frame * kR + i] = i < nR ? s_match[nL + i] : -1;
  if (tid == 0) nmatches_out[frame] = s_nm;
}

// ---- Frame::UndistortKeyPoints (reference src/Frame.cc:829-857): cv::undistortPoints(mat, mat, K, mDistCoef, cv::Mat(), mK) =
// OpenCV's cvUndistortPointsInternal with 5 fixed iterations, all in double after widening the float inputs, narrowed to
// float at the end (restated and pinned against cv2 in oracle/shim: undistort_points_pinhole). One thread per keypoint;
// the file is compiled with -fmad=false, so every double product and sum rounds as in the scalar reference.
struct UndistortParams {
  double fx, fy, cx, cy, ifx, ify;
  double k[12];
  double rr[9];
};

__global__ void __launch_bounds__(256) k_undistort(const orb_keypoint* __restrict__ kps, const int* __restrict__ n_arr, int kcap,
                                                   UndistortParams p, orb_keypoint* __restrict__ out) {
  const int frame = blockIdx.y, i = blockIdx.x * 256 + threadIdx.x;
  if (i >= min(n_arr[frame], kcap)) return;
  orb_keypoint kp = kps[(size_t)frame * kcap + i];
  const double u = kp.x, v = kp.y;
  double x = (u - p.cx) * p.ifx, y = (v - p.cy) * p.ify;
  const double x0 = x, y0 = y;
  const double* k = p.k;
  for (int j = 0; j < 5; ++j) {                              // TermCriteria(MAX_ITER, 5, 0.01)
    const double r2 = x * x + y * y;
    const double icdist = (1 + ((k[7] * r2 + k[6]) * r2 + k[5]) * r2) / (1 + ((k[4] * r2 + k[1]) * r2 + k[0]) * r2);
    if (icdist < 0) { x = (u - p.cx) * p.ifx; y = (v - p.cy) * p.ify; break; }
    const double deltaX = 2 * k[2] * x * y + k[3] * (r2 + 2 * x * x) + k[8] * r2 + k[9] * r2 * r2;
    const double deltaY = k[2] * (r2 + 2 * y * y) + 2 * k[3] * x * y + k[10] * r2 + k[11] * r2 * r2;
    x = (x0 - deltaX) * icdist;
    y = (y0 - deltaY) * icdist;
  }
  const double xx = p.rr[0] * x + p.rr[1] * y + p.rr[2], yy = p.rr[3] * x + p.rr[4] * y + p.rr[5];
  const double ww = 1. / (p.rr[6] * x + p.rr[7] * y + p.rr[8]);
  kp.x = (float)(xx * ww);
  kp.y = (float)(yy * ww);
  out[(size_t)frame * kcap + i] = kp;                        // mvKeysUn[i] = mvKeys[i] with the new pt (:851-856)
}

// ---- host side ---------------------------------------------------------------------------------------------
// Frame::mvKeysUn: the undistorted keypoints when orb_undistort_keypoints ran on this batch, else mvKeys (:830-833)
static const orb_keypoint* orb_keys_un(const orb_handle* h) { return h->have_undist ? h->d_kps_un.as<orb_keypoint>() : h->d_kps.as<orb_keypoint>(); }

// ================================================================================================================================
// Two-camera frames (Frame::Nleft != -1, the fisheye rig): the right-camera halves of the two SearchByProjection overloads.
// The right camera is a second handle; its grid (orb_assign_features_to_grid on that handle) is mGridRight and a window scan on
// it is GetFeaturesInArea(..., bRight = true) (src/Frame.cc:783-790).
// ================================================================================================================================

// window of a frame-to-frame query in the RIGHT camera (:1639-1661): same radius and level gates as the left one, centred on
// project(Trl * x3Dc); no image-bounds test, no mvuRight gate
static __device__ __forceinline__ SpWindow sp_window_right(const orb_proj_query2& q, const GridParams& gp, const float* __restrict__ scale,
                                                           float th, int mode) {
  SpWindow w;
  w.ok = false;
  const int oct = q.octave;
  w.u = q.ur; w.v = q.vr;
  w.invz = 0.f; w.ur = 0.f;
  w.radius = __fmul_rn(th, scale[oct]);
  if (mode == 1) { w.min_level = oct; w.max_level = -1; }
  else if (mode == 2) { w.min_level = 0; w.max_level = oct; }
  else { w.min_level = oct - 1; w.max_level = oct + 1; }
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

static __device__ __forceinline__ orb_proj_query sp_left_part(const orb_proj_query2& q) {
  orb_proj_query a;
  a.u = q.u; a.v = q.v; a.z = q.z; a.angle = q.angle; a.octave = q.octave; a.flags = q.flags;
  return a;
}

// SIDE 0: left camera of a two-camera frame (no mvuRight gate; writes area[q] = the window held a keypoint, i.e. !vIndices2.empty());
// SIDE 1: right camera (only for queries whose left window held one: every earlier `continue` of the loop body skips the right half)
template <int SIDE>
__global__ void __launch_bounds__(SP_WARPS * 32, SP_MINB) k_sp2_window(
    const orb_keypoint* __restrict__ kps, const uint8_t* __restrict__ desc, int kcap, const int* __restrict__ cell_off,
    const unsigned short* __restrict__ cell_idx, const orb_proj_query2* __restrict__ queries, const uint8_t* __restrict__ qdesc,
    const int* __restrict__ nq_arr, int qcap, GridParams gp, OrbGeom g, float th, const float* __restrict__ tlc_z, float mb, int mono,
    unsigned char* __restrict__ area, uint4* __restrict__ cand, unsigned char* __restrict__ cand_cnt) {
  const int frame = blockIdx.y, lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const int qi = blockIdx.x * SP_WARPS + wid;
  if (qi >= min(nq_arr[frame], qcap)) return;
  const size_t qo = (size_t)frame * qcap + qi;
  const orb_proj_query2 q = queries[qo];
  const float tz = tlc_z[frame];
  const int mode = (tz > mb && !mono) ? 1 : ((-tz > mb && !mono) ? 2 : 0);
  SpWindow w;
  if (SIDE == 0) w = sp_window(sp_left_part(q), gp, g.scale, th, mode, 0.f);
  else {
    w = sp_window_right(q, gp, g.scale, th, mode);
    if (!area[qo]) w.ok = false;
  }
  const unsigned short* idx = cell_idx + (size_t)frame * kcap;
  const uint4* qd = reinterpret_cast<const uint4*>(qdesc + qo * 32);
  unsigned int k0, k1, k2, k3;
  int narea = 0;
  const int cnt = win_scan(w, SP_TH_HIGH, qd[0], qd[1], kps + (size_t)frame * kcap, desc + (size_t)frame * kcap * 32, nullptr,
                           cell_off + (size_t)frame * (GRID_CELLS + 1), idx, nullptr, lane, k0, k1, k2, k3, &narea);
  unsigned int mine = SL_NONE;
#pragma unroll
  for (int r = 0; r < SL_K; ++r) {
    const unsigned int m = sl_pop(k0, k1, k2, k3);
    if (lane == r) mine = m;
  }
  unsigned int rec = SL_NONE;
  if (lane < SL_K && mine != SL_NONE) rec = ((mine >> 16) << 20) | (unsigned int)idx[mine & 0xffffu];
  const unsigned int r1 = __shfl_sync(0xffffffffu, rec, 1), r2 = __shfl_sync(0xffffffffu, rec, 2), r3 = __shfl_sync(0xffffffffu, rec, 3);
  if (lane == 0) {
    cand[qo] = make_uint4(rec, r1, r2, r3);
    cand_cnt[qo] = (unsigned char)min(cnt, 255);
    if (SIDE == 0) area[qo] = narea > 0 ? 1 : 0;
  }
}

struct Sp2Side {
  const orb_keypoint* kps;
  const uint8_t* desc;
  const int* n_arr;
  int kcap;
  const int* cell_off;
  const unsigned short* cell_idx;
  const uint4* cand;
  const unsigned char* cand_cnt;
  int* match_out;
};

// Resolver of the two-camera frame-to-frame search: the left and the right camera are two independent greedy assignments (their
// keypoints and locks are disjoint index ranges; what couples them - the skipped right half - is already in the candidates), replayed
// one after the other with the round scheme of k_sp_resolve; the rotation histogram and its records are shared (:1711-1730).
// dynamic shared memory: assigned[kL + kR] i32 | cangle[kL + kR] f32 | recs[2 qcap] u32 | lock[kL + kR] u8
static size_t sp2_resolve_smem(int qcap, int kl, int kr) { return (size_t)(kl + kr) * 9 + (size_t)qcap * 8 + 32; }

__global__ void __launch_bounds__(128) k_sp2_resolve(Sp2Side SL, Sp2Side SR, const orb_proj_query2* __restrict__ queries,
                                                     const uint8_t* __restrict__ qdesc, const int* __restrict__ nq_arr, int qcap, GridParams gp,
                                                     OrbGeom g, float th, const float* __restrict__ tlc_z, float mb, int mono,
                                                     int check_orientation, int* __restrict__ nmatches_out) {
  extern __shared__ __align__(16) unsigned char s_raw[];
  const int ktot = SL.kcap + SR.kcap;
  int* s_assigned = reinterpret_cast<int*>(s_raw);
  float* s_cangle = reinterpret_cast<float*>(s_assigned + ktot);
  unsigned int* s_recs = reinterpret_cast<unsigned int*>(s_cangle + ktot);
  unsigned char* s_lock = reinterpret_cast<unsigned char*>(s_recs + 2 * qcap);
  __shared__ uint4 s_cand[SL_CHUNK];
  __shared__ float s_qangle[SL_CHUNK];
  __shared__ unsigned char s_cnt[SL_CHUNK];
  __shared__ unsigned char s_obs[SL_CHUNK];
  __shared__ int s_hist[SP_HISTO];
  __shared__ int s_nm, s_nrec;
  const int frame = blockIdx.x, tid = threadIdx.x, lane = tid & 31;
  const int nq = min(nq_arr[frame], qcap);
  const orb_proj_query2* q = queries + (size_t)frame * qcap;
  const float tz = tlc_z[frame];
  const int mode = (tz > mb && !mono) ? 1 : ((-tz > mb && !mono) ? 2 : 0);
  const float factor = 1.0f / SP_HISTO;
  if (tid < SP_HISTO) s_hist[tid] = 0;
  int nm = 0, nrec = 0;   // warp 0, uniform
  for (int side = 0; side < 2; ++side) {
    const Sp2Side& S = side ? SR : SL;
    const int koff = side ? SL.kcap : 0;                 // this camera's slice of the shared arrays
    const int nC = min(S.n_arr[frame], S.kcap);
    const orb_keypoint* kp = S.kps + (size_t)frame * S.kcap;
    const uint4* cd = S.cand + (size_t)frame * qcap;
    const unsigned char* cc = S.cand_cnt + (size_t)frame * qcap;
    const unsigned short* idx = S.cell_idx + (size_t)frame * S.kcap;
    int* assigned = s_assigned + koff;
    float* cangle = s_cangle + koff;
    unsigned char* lock = s_lock + koff;
    __syncthreads();
    for (int i = tid; i < nC; i += 128) { cangle[i] = kp[i].angle; assigned[i] = -1; lock[i] = 0; }
    for (int base = 0; base < nq; base += SL_CHUNK) {
      const int m = min(SL_CHUNK, nq - base);
      __syncthreads();
      for (int i = tid; i < m; i += 128) {
        s_cand[i] = cd[base + i];
        s_cnt[i] = cc[base + i];
        s_qangle[i] = q[base + i].angle;
        s_obs[i] = (q[base + i].flags & 2) ? 1 : 0;
      }
      __syncthreads();
      if (tid < 32) {
        int head = 0;
        while (head < m) {
          const int i = head + lane;
          int pick = -1, obs = 0;
          bool rescan = false;
          if (i < m) {
            const int cnt = s_cnt[i];
            obs = s_obs[i];
            if (cnt) {
              const uint4 c4 = s_cand[i];
              const unsigned int c[SL_K] = {c4.x, c4.y, c4.z, c4.w};
#pragma unroll
              for (int k = SL_K - 1; k >= 0; --k)
                if (c[k] != SL_NONE && !lock[c[k] & 0xffffu]) pick = (int)(c[k] & 0xffffu);
              if (pick < 0 && cnt > SL_K) rescan = true;
            }
          }
          const unsigned rs = __ballot_sync(0xffffffffu, rescan);
          int ncommit;
          if (rs & 1u) {
            const int qi = base + head;
            const orb_proj_query2 qq = q[qi];
            const SpWindow w = side ? sp_window_right(qq, gp, g.scale, th, mode) : sp_window(sp_left_part(qq), gp, g.scale, th, mode, 0.f);
            const uint4* qd = reinterpret_cast<const uint4*>(qdesc + ((size_t)frame * qcap + qi) * 32);
            unsigned int k0, k1, k2, k3;
            win_scan(w, SP_TH_HIGH, qd[0], qd[1], kp, S.desc + (size_t)frame * S.kcap * 32, nullptr, S.cell_off + (size_t)frame * (GRID_CELLS + 1),
                     idx, lock, lane, k0, k1, k2, k3);
            const unsigned int m1 = sl_pop(k0, k1, k2, k3);
            pick = (lane == 0 && m1 != SL_NONE) ? (int)idx[m1 & 0xffffu] : -1;
            ncommit = 1;
          } else {
            const int lockpick = (pick >= 0 && obs) ? pick : -1;
            bool bad = false;
#pragma unroll 8
            for (int k = 0; k < 31; ++k) {
              const int pk = __shfl_sync(0xffffffffu, lockpick, k);
              bad |= (k < lane) && (pk >= 0) && (pk == pick);
            }
            const unsigned stop = __ballot_sync(0xffffffffu, bad) | rs;
            ncommit = stop ? __ffs(stop) - 1 : 32;
          }
          const bool commit = lane < ncommit && pick >= 0;
          const unsigned cm = __ballot_sync(0xffffffffu, commit);
          __syncwarp();
          if (commit) {
            const unsigned same = __match_any_sync(cm, pick);
            if (lane == 31 - __clz(same)) { assigned[pick] = base + i; lock[pick] = (unsigned char)obs; }
            if (check_orientation) {
              float rot = __fsub_rn(s_qangle[i], cangle[pick]);
              if (rot < 0.0f) rot = __fadd_rn(rot, 360.0f);
              int bin = (int)roundf(__fmul_rn(rot, factor));
              if (bin == SP_HISTO) bin = 0;
              s_recs[nrec + __popc(cm & ((1u << lane) - 1u))] = (unsigned int)(pick + koff) | ((unsigned int)bin << 16);
              atomicAdd(&s_hist[bin], 1);
            }
          }
          nm += __popc(cm);
          nrec += __popc(cm);
          __syncwarp();
          head += ncommit;
        }
      }
    }
  }
  if (tid == 0) { s_nm = nm; s_nrec = nrec; }
  __syncthreads();
  if (check_orientation && tid < 32) {
    int ind1 = -1, ind2 = -1, ind3 = -1, max1 = 0, max2 = 0, max3 = 0;
    for (int i = 0; i < SP_HISTO; i++) {
      const int s = s_hist[i];
      if (s > max1) { max3 = max2; max2 = max1; max1 = s; ind3 = ind2; ind2 = ind1; ind1 = i; }
      else if (s > max2) { max3 = max2; max2 = s; ind3 = ind2; ind2 = i; }
      else if (s > max3) { max3 = s; ind3 = i; }
    }
    if ((float)max2 < __fmul_rn(0.1f, (float)max1)) { ind2 = -1; ind3 = -1; }
    else if ((float)max3 < __fmul_rn(0.1f, (float)max1)) { ind3 = -1; }
    int drop = 0;
    for (int r = lane; r < s_nrec; r += 32) {
      const int bin = (int)(s_recs[r] >> 16);
      if (bin != ind1 && bin != ind2 && bin != ind3) { s_assigned[s_recs[r] & 0xffffu] = -1; drop++; }
    }
    drop = __reduce_add_sync(0xffffffffu, drop);
    if (lane == 0) s_nm -= drop;
  }
  __syncthreads();
  {
    const int nL = min(SL.n_arr[frame], SL.kcap), nR = min(SR.n_arr[frame], SR.kcap);
    for (int i = tid; i < SL.kcap; i += 128) SL.match_out[(size_t)frame * SL.kcap + i] = i < nL ? s_assigned[i] : -1;
    for (int i = tid; i < SR.kcap; i += 128) SR.match_out[(size_t)frame * SR.kcap + i] = i < nR ? s_assigned[SL.kcap + i] : -1;
  }
  if (tid == 0) nmatches_out[frame] = s_nm;
}

// ---- local map against a two-camera frame (src/ORBmatcher.cc:42-209 with Nleft != -1) -------------------------------------------
// the query of orb_search_local_points_stereo split into the two cameras' orb_track_query records, so that k_sl_window serves both:
// right camera = (mTrackProjXR, mTrackProjYR, mTrackViewCosR, mnTrackScaleLevelR), visible when bit 2 is set
__global__ void k_split_track_queries(const orb_track_query2* __restrict__ q2, int n, orb_track_query* __restrict__ qL,
                                      orb_track_query* __restrict__ qR) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const orb_track_query2 q = q2[i];
  orb_track_query a, b;
  a.proj_x = q.proj_x; a.proj_y = q.proj_y; a.proj_xr = 0.f; a.view_cos = q.view_cos; a.level = q.level; a.flags = q.flags & 3;
  b.proj_x = q.proj_xr; b.proj_y = q.proj_yr; b.proj_xr = 0.f; b.view_cos = q.view_cos_r; b.level = q.level_r;
  b.flags = ((q.flags >> 2) & 1) | (q.flags & 2);
  qL[i] = a; qR[i] = b;
}

struct Sl2Side {
  const orb_keypoint* kps;
  const uint8_t* desc;
  const int* n_arr;
  int kcap;
  const int* cell_off;
  const unsigned short* cell_idx;
  const orb_track_query* queries;     // this camera's view of the map points (k_split_track_queries)
  const uint4* cand;
  const unsigned char* cand_cnt;
  const uint8_t* locked0;
  const int* partner;                 // mvLeftToRightMatch (left side) / mvRightToLeftMatch (right side)
  int* match_out;
};

// The two cameras' assignments are coupled (a match also gives the map point to the stereo partner in the other camera, an
// unconditional overwrite that may lock or RELEASE that keypoint, and a left ratio failure skips the right camera), so the map
// points are replayed strictly in order: the warp decides one camera of one map point at a time from the stored candidates (the
// first two unlocked ones under the current locks) and scans the window again when they run out while it holds more.
// dynamic shared memory: assigned[kL + kR] i32 | partner[kL + kR] i32 | lock[kL + kR] u8
#define SL2_CHUNK 1024
static size_t sl2_resolve_smem(int kl, int kr) { return (size_t)(kl + kr) * 9 + 32; }

__global__ void __launch_bounds__(128) k_sl2_resolve(Sl2Side SL, Sl2Side SR, const uint8_t* __restrict__ qdesc, const int* __restrict__ nq_arr,
                                                     int qcap, GridParams gp, OrbGeom g, float th, float nnratio,
                                                     int* __restrict__ nmatches_out) {
  extern __shared__ __align__(16) unsigned char s_raw[];
  const int ktot = SL.kcap + SR.kcap;
  int* s_assigned = reinterpret_cast<int*>(s_raw);
  int* s_partner = s_assigned + ktot;
  unsigned char* s_lock = reinterpret_cast<unsigned char*>(s_partner + ktot);
  __shared__ uint4 s_cand[2][SL2_CHUNK];
  __shared__ unsigned char s_cnt[2][SL2_CHUNK];
  __shared__ unsigned char s_obs[SL2_CHUNK];
  const int frame = blockIdx.x, tid = threadIdx.x, lane = tid & 31;
  const int nq = min(nq_arr[frame], qcap);
  for (int side = 0; side < 2; ++side) {
    const Sl2Side& S = side ? SR : SL;
    const int koff = side ? SL.kcap : 0, nC = min(S.n_arr[frame], S.kcap), ocap = side ? SL.kcap : SR.kcap;
    for (int i = tid; i < S.kcap; i += 128) {
      s_assigned[koff + i] = -1;
      s_lock[koff + i] = (i < nC && S.locked0) ? S.locked0[(size_t)frame * S.kcap + i] : 0;
      int pt = (i < nC && S.partner) ? S.partner[(size_t)frame * S.kcap + i] : -1;
      if (pt >= ocap) pt = -1;
      s_partner[koff + i] = pt;
    }
  }
  int nm = 0;   // warp 0, uniform
  for (int base = 0; base < nq; base += SL2_CHUNK) {
    const int m = min(SL2_CHUNK, nq - base);
    __syncthreads();
    for (int i = tid; i < m; i += 128) {
      const size_t qo = (size_t)frame * qcap + base + i;
      s_cand[0][i] = SL.cand[qo]; s_cand[1][i] = SR.cand[qo];
      s_cnt[0][i] = SL.cand_cnt[qo]; s_cnt[1][i] = SR.cand_cnt[qo];
      s_obs[i] = (unsigned char)((SL.queries[qo].flags >> 1) & 1);
    }
    __syncthreads();
    if (tid < 32) {
      for (int i = 0; i < m; ++i) {
        const size_t qo = (size_t)frame * qcap + base + i;
        const int obs = s_obs[i];
        bool skip_right = false;
        for (int side = 0; side < 2; ++side) {
          if (side == 1 && skip_right) break;
          const int cnt = s_cnt[side][i];                 // 0 also for a map point that is not in view of this camera
          if (!cnt) continue;
          const Sl2Side& S = side ? SR : SL;
          const int koff = side ? SL.kcap : 0, ooff = side ? 0 : SL.kcap;
          const uint4 c4 = s_cand[side][i];
          const unsigned int c[SL_K] = {c4.x, c4.y, c4.z, c4.w};
          const unsigned char* lock = s_lock + koff;
          unsigned int b1 = SL_NONE, b2 = SL_NONE;
#pragma unroll
          for (int k = 0; k < SL_K; ++k) {
            const unsigned int key = c[k];
            if (key == SL_NONE || b2 != SL_NONE) continue;
            if (lock[key & 0xffffu]) continue;
            if (b1 == SL_NONE) b1 = key; else b2 = key;
          }
          int bestDist = 256, bestLevel = -1, bestDist2 = 256, bestLevel2 = -1, bestIdx = -1;
          bool more = b2 == SL_NONE && cnt > SL_K;
          if (more) {   // the exact scan is only needed when it can change the outcome (see k_sl_resolve)
            const int lastDist = (int)(c[SL_K - 1] >> 20);
            if (b1 == SL_NONE) { if (lastDist > SP_TH_HIGH) more = false; }
            else {
              const int bd = (int)(b1 >> 20);
              if (bd > SP_TH_HIGH || !((float)bd > __fmul_rn(nnratio, (float)lastDist))) more = false;
            }
          }
          if (more) {
            const orb_track_query qq = S.queries[qo];
            const SlWindow w = sl_window(qq, gp, g, side ? 1.0f : th);          // no th factor in the right camera (:131)
            const uint4* qd = reinterpret_cast<const uint4*>(qdesc + qo * 32);
            const orb_keypoint* kp = S.kps + (size_t)frame * S.kcap;
            const unsigned short* idx = S.cell_idx + (size_t)frame * S.kcap;
            unsigned int k0, k1, k2, k3;
            win_scan(w, 256, qd[0], qd[1], kp, S.desc + (size_t)frame * S.kcap * 32, nullptr, S.cell_off + (size_t)frame * (GRID_CELLS + 1), idx,
                     lock, lane, k0, k1, k2, k3);
            const unsigned int m1 = sl_pop(k0, k1, k2, k3), m2 = sl_pop(k0, k1, k2, k3);
            if (m1 != SL_NONE) {
              bestIdx = idx[m1 & 0xffffu];
              bestDist = (int)(m1 >> 16); bestLevel = kp[bestIdx].octave;
              if (m2 != SL_NONE) { bestDist2 = (int)(m2 >> 16); bestLevel2 = kp[idx[m2 & 0xffffu]].octave; }
            }
          } else if (b1 != SL_NONE) {
            bestIdx = (int)(b1 & 0xffffu);
            bestDist = (int)(b1 >> 20); bestLevel = (int)((b1 >> 16) & 15u);
            if (b2 != SL_NONE) { bestDist2 = (int)(b2 >> 20); bestLevel2 = (int)((b2 >> 16) & 15u); }
          }
          if (bestIdx >= 0 && bestDist <= SP_TH_HIGH) {
            if (bestLevel == bestLevel2 && (float)bestDist > __fmul_rn(nnratio, (float)bestDist2)) {
              if (side == 0) skip_right = true;            // :123 `continue` leaves the loop body: no right camera for this map point
              continue;
            }
            const int partner = s_partner[koff + bestIdx];
            __syncwarp();
            if (lane == 0) {
              s_assigned[koff + bestIdx] = base + i; s_lock[koff + bestIdx] = (unsigned char)obs;
              if (partner >= 0) { s_assigned[ooff + partner] = base + i; s_lock[ooff + partner] = (unsigned char)obs; }
            }
            nm += 1 + (partner >= 0 ? 1 : 0);
            __syncwarp();
          }
        }
      }
    }
  }
  __syncthreads();
  {
    const int nL = min(SL.n_arr[frame], SL.kcap), nR = min(SR.n_arr[frame], SR.kcap);
    for (int i = tid; i < SL.kcap; i += 128) SL.match_out[(size_t)frame * SL.kcap + i] = i < nL ? s_assigned[i] : -1;
    for (int i = tid; i < SR.kcap; i += 128) SR.match_out[(size_t)frame * SR.kcap + i] = i < nR ? s_assigned[SL.kcap + i] : -1;
  }
  if (tid == 0) nmatches_out[frame] = nm;
}


static GridParams to_gp(const orb_grid_params* p) {
  GridParams g;
  g.min_x = p->min_x; g.min_y = p->min_y; g.max_x = p->max_x; g.max_y = p->max_y; g.w_inv = p->grid_w_inv; g.h_inv = p->grid_h_inv;
  return g;
}

extern "C" {

int orb_undistort_keypoints(orb_handle* h, const float* K, const float* dist, int ndist, const float* P, orb_keypoint* kps_un_out, int cap,
                            int flags) {
  if (!h || !K || !P || (ndist > 0 && !dist) || ndist < 0) return ORB_ERR_INVALID_ARG;
  if (ndist > 12) return orb_set_error(h, ORB_ERR_INVALID_ARG, "tilted sensor models (more than 12 distortion coefficients) are not supported");
  if (!h->have_batch) return orb_set_error(h, ORB_ERR_STATE, "no extraction has run on this handle");
  int st;
  if ((st = orb_use_device(h))) return st;
  const int batch = h->cur_batch, kcap = h->g.kcap;
  if (ndist == 0 || dist[0] == 0.0f) {                       // mvKeysUn = mvKeys (:830-833)
    h->have_undist = false;
    h->have_grid = false;
  } else {
    if ((st = orb_ensure(h, h->d_kps_un, (size_t)batch * kcap * sizeof(orb_keypoint)))) return st;
    UndistortParams p;
    p.fx = K[0]; p.fy = K[4]; p.cx = K[2]; p.cy = K[5];
    p.ifx = 1. / p.fx; p.ify = 1. / p.fy;
    for (int i = 0; i < 12; ++i) p.k[i] = i < ndist ? (double)dist[i] : 0.0;
    for (int i = 0; i < 9; ++i) p.rr[i] = P[i];
    k_undistort<<<dim3((kcap + 255) / 256, batch), 256, 0, h->stream>>>(h->d_kps.as<orb_keypoint>(), h->d_n.as<int>(), kcap, p,
                                                                       h->d_kps_un.as<orb_keypoint>());
    h->launches++;
    ORB_CUDA_CHECK(h, cudaGetLastError());
    h->have_undist = true;
    h->have_grid = false;
  }
  if (kps_un_out && !(flags & ORB_NO_OUTPUT)) {
    const int rows = std::min(cap, kcap);
    ORB_CUDA_CHECK(h, cudaMemcpy2DAsync(kps_un_out, (size_t)cap * sizeof(orb_keypoint), orb_keys_un(h), (size_t)kcap * sizeof(orb_keypoint),
                                        (size_t)rows * sizeof(orb_keypoint), batch, cudaMemcpyDefault, h->stream));
  }
  if (flags & ORB_ASYNC) return ORB_OK;
  ORB_CUDA_CHECK(h, cudaStreamSynchronize(h->stream));
  return ORB_OK;
}

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
  k_grid_build<<<batch, 256, 0, h->stream>>>(orb_keys_un(h), h->d_n.as<int>(), kcap, to_gp(gp), h->d_grid_off.as<int>(),
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
  if (smem > 160 * 1024) return orb_set_error(h, ORB_ERR_CAPACITY, "too many queries / keypoints per frame for the resolver");
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
  if ((st = orb_ensure(h, h->d_sp_cand, nqt * SL_K * sizeof(unsigned int)))) return st;
  if ((st = orb_ensure(h, h->d_sp_cnt, nqt))) return st;
  if ((st = orb_ensure(h, h->d_sp_match, (size_t)batch * kcap * sizeof(int)))) return st;
  if ((st = orb_ensure(h, h->d_sp_nm, (size_t)batch * sizeof(int)))) return st;
  const float* d_ur = h->have_stereo ? h->d_uright.as<float>() : nullptr;   // mvuRight = -1 without a stereo match
  const GridParams gp = to_gp(&h->grid_params);
  k_sp_window<<<dim3((qcap + SP_WARPS - 1) / SP_WARPS, batch), SP_WARPS * 32, 0, h->stream>>>(
      orb_keys_un(h), h->d_desc.as<uint8_t>(), d_ur, kcap, h->d_grid_off.as<int>(), h->d_grid_idx.as<unsigned short>(), d_q, d_qd,
      d_nq, qcap, gp, h->g, th, d_tz, mb, mono, mbf, h->d_sp_cand.as<uint4>(), h->d_sp_cnt.as<unsigned char>(), nullptr, SP_TH_HIGH, 0);
  h->launches++;
  { const int st_a = orb_raise_dyn_smem(h, (const void*)k_sp_resolve, smem); if (st_a) return st_a; }
  k_sp_resolve<<<batch, 128, smem, h->stream>>>(orb_keys_un(h), h->d_desc.as<uint8_t>(), d_ur, h->d_n.as<int>(), kcap,
                                                h->d_grid_off.as<int>(), h->d_grid_idx.as<unsigned short>(), d_q, d_qd, d_nq, qcap, gp, h->g, th,
                                                d_tz, mb, mono, mbf, check_orientation, h->d_sp_cand.as<uint4>(),
                                                h->d_sp_cnt.as<unsigned char>(), h->d_sp_match.as<int>(), h->d_sp_nm.as<int>(), nullptr,
                                                SP_TH_HIGH, 0);
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
      orb_keys_un(h), h->d_desc.as<uint8_t>(), d_ur, kcap, h->d_grid_off.as<int>(), h->d_grid_idx.as<unsigned short>(), d_q, d_qd,
      d_nq, qcap, d_lk, gp, h->g, th, h->d_sp_cand.as<uint4>(), h->d_sp_cnt.as<unsigned char>());
  h->launches++;
  { const int st_a = orb_raise_dyn_smem(h, (const void*)k_sl_resolve, smem); if (st_a) return st_a; }
  k_sl_resolve<<<batch, 128, smem, h->stream>>>(orb_keys_un(h), h->d_desc.as<uint8_t>(), d_ur, h->d_n.as<int>(), kcap,
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

int orb_search_by_bow(orb_handle* h, const orb_bow_keyframes* kf, float nnratio, int check_orientation, int32_t* match_out,
                      int32_t* nmatches_out, int flags) {
  if (!h || !kf || !kf->desc || !kf->angle || !kf->flags || !kf->n || !kf->fv_node || !kf->fv_off || !kf->fv_feat || !kf->fv_n || kf->cap < 1)
    return ORB_ERR_INVALID_ARG;
  if (!h->have_batch || !h->have_bow) return orb_set_error(h, ORB_ERR_STATE, "orb_compute_bow has not run on this handle's batch");
  int st;
  if ((st = orb_use_device(h))) return st;
  const int batch = h->cur_batch, kcap = h->g.kcap, qcap = kf->cap;
  if (kcap > 65535) return orb_set_error(h, ORB_ERR_CAPACITY, "more than 65535 keypoints per frame");
  const size_t smem = (size_t)kcap * 8;
  if (smem > 200 * 1024) return orb_set_error(h, ORB_ERR_CAPACITY, "too many keypoints per frame");
  const size_t nq = (size_t)batch * qcap;
  const uint8_t *d_desc = kf->desc, *d_flags = kf->flags;
  const float* d_angle = kf->angle;
  const int *d_n = kf->n, *d_off = kf->fv_off, *d_nn = kf->fv_n;
  const unsigned int *d_node = kf->fv_node, *d_feat = kf->fv_feat;
  if (!(flags & ORB_SRC_DEVICE)) {
    // one staging buffer: desc | angle | node | feat | off | flags | n | nn
    size_t o[9];
    const size_t bytes[8] = {nq * 32, nq * 4, nq * 4, nq * 4, (size_t)batch * (qcap + 1) * 4, nq, (size_t)batch * 4, (size_t)batch * 4};
    o[0] = 0;
    for (int i = 0; i < 8; ++i) o[i + 1] = o[i] + ((bytes[i] + 255) & ~(size_t)255);
    if ((st = orb_ensure(h, h->d_scratch, o[8]))) return st;
    uint8_t* base = h->d_scratch.as<uint8_t>();
    const void* src[8] = {kf->desc, kf->angle, kf->fv_node, kf->fv_feat, kf->fv_off, kf->flags, kf->n, kf->fv_n};
    for (int i = 0; i < 8; ++i) ORB_CUDA_CHECK(h, cudaMemcpyAsync(base + o[i], src[i], bytes[i], cudaMemcpyHostToDevice, h->stream));
    d_desc = base + o[0]; d_angle = (const float*)(base + o[1]); d_node = (const unsigned int*)(base + o[2]);
    d_feat = (const unsigned int*)(base + o[3]); d_off = (const int*)(base + o[4]); d_flags = base + o[5];
    d_n = (const int*)(base + o[6]); d_nn = (const int*)(base + o[7]);
  }
  if ((st = orb_ensure(h, h->d_sp_match, (size_t)batch * kcap * sizeof(int)))) return st;
  if ((st = orb_ensure(h, h->d_sp_nm, (size_t)batch * sizeof(int)))) return st;
  { const int st_a = orb_raise_dyn_smem(h, (const void*)k_sbow, smem); if (st_a) return st_a; }
  // F.mvKeys (not mvKeysUn) supplies the frame keypoint's angle (:314-318); the two hold the same angle anyway
  k_sbow<<<batch, 256, smem, h->stream>>>(h->d_kps.as<orb_keypoint>(), h->d_desc.as<uint8_t>(), h->d_n.as<int>(), kcap,
                                          h->d_fv_node.as<unsigned int>(), h->d_fv_off.as<int>(), h->d_fv_feat.as<unsigned int>(),
                                          h->d_bow_n.as<int>() + batch, d_desc, d_angle, d_flags, d_n, qcap, d_node, d_off, d_feat, d_nn, nnratio,
                                          check_orientation, h->d_sp_match.as<int>(), h->d_sp_nm.as<int>());
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

static int search_by_projection_kf_impl(orb_handle* h, const orb_proj_query* queries, const uint8_t* qdesc, const int32_t* nq, int qcap,
                                        const uint8_t* locked0, float th, int orb_dist, int check_orientation, int32_t* match_out,
                                        int32_t* nmatches_out, int flags, int kf) {
  if (!h || !queries || !qdesc || !nq || qcap < 1 || orb_dist < 0) return ORB_ERR_INVALID_ARG;
  if (!h->have_grid) return orb_set_error(h, ORB_ERR_STATE, "orb_assign_features_to_grid has not run on this handle");
  if (qcap > 65535) return orb_set_error(h, ORB_ERR_CAPACITY, "more than 65535 queries per frame");
  int st;
  if ((st = orb_use_device(h))) return st;
  const int batch = h->cur_batch, kcap = h->g.kcap;
  const size_t smem = sp_resolve_smem(qcap, kcap);
  if (smem > 160 * 1024) return orb_set_error(h, ORB_ERR_CAPACITY, "too many queries / keypoints per frame for the resolver");
  const size_t nqt = (size_t)batch * qcap;
  const orb_proj_query* d_q = queries;
  const uint8_t* d_qd = qdesc;
  const int* d_nq = nq;
  const uint8_t* d_lk = locked0;
  if (!(flags & ORB_SRC_DEVICE)) {
    const size_t b_q = nqt * sizeof(orb_proj_query), b_d = nqt * 32, b_n = (size_t)batch * 4, b_l = locked0 ? (size_t)batch * kcap : 0;
    const size_t o_d = (b_q + 255) & ~(size_t)255, o_n = o_d + ((b_d + 255) & ~(size_t)255), o_l = o_n + ((b_n + 255) & ~(size_t)255);
    if ((st = orb_ensure(h, h->d_scratch, o_l + b_l + 16))) return st;
    uint8_t* base = h->d_scratch.as<uint8_t>();
    ORB_CUDA_CHECK(h, cudaMemcpyAsync(base, queries, b_q, cudaMemcpyHostToDevice, h->stream));
    ORB_CUDA_CHECK(h, cudaMemcpyAsync(base + o_d, qdesc, b_d, cudaMemcpyHostToDevice, h->stream));
    ORB_CUDA_CHECK(h, cudaMemcpyAsync(base + o_n, nq, b_n, cudaMemcpyHostToDevice, h->stream));
    if (locked0) ORB_CUDA_CHECK(h, cudaMemcpyAsync(base + o_l, locked0, b_l, cudaMemcpyHostToDevice, h->stream));
    d_q = (const orb_proj_query*)base; d_qd = base + o_d; d_nq = (const int*)(base + o_n); d_lk = locked0 ? base + o_l : nullptr;
  }
  if ((st = orb_ensure(h, h->d_sp_cand, nqt * SL_K * sizeof(unsigned int)))) return st;
  if ((st = orb_ensure(h, h->d_sp_cnt, nqt))) return st;
  if ((st = orb_ensure(h, h->d_sp_match, (size_t)batch * kcap * sizeof(int)))) return st;
  if ((st = orb_ensure(h, h->d_sp_nm, (size_t)batch * sizeof(int)))) return st;
  const GridParams gp = to_gp(&h->grid_params);
  const int max_dist = std::min(orb_dist, 255);   // bestDist starts at 256 with "dist < bestDist" (:1786-1800)
  // the KeyFrame overload has no mvuRight gate and no forward / backward modes: uright = NULL, kf = 1
  k_sp_window<<<dim3((qcap + SP_WARPS - 1) / SP_WARPS, batch), SP_WARPS * 32, 0, h->stream>>>(
      orb_keys_un(h), h->d_desc.as<uint8_t>(), nullptr, kcap, h->d_grid_off.as<int>(), h->d_grid_idx.as<unsigned short>(), d_q, d_qd, d_nq,
      qcap, gp, h->g, th, nullptr, 0.f, 1, 0.f, h->d_sp_cand.as<uint4>(), h->d_sp_cnt.as<unsigned char>(), d_lk, max_dist, kf);
  h->launches++;
  { const int st_a = orb_raise_dyn_smem(h, (const void*)k_sp_resolve, smem); if (st_a) return st_a; }
  k_sp_resolve<<<batch, 128, smem, h->stream>>>(orb_keys_un(h), h->d_desc.as<uint8_t>(), nullptr, h->d_n.as<int>(), kcap,
                                                h->d_grid_off.as<int>(), h->d_grid_idx.as<unsigned short>(), d_q, d_qd, d_nq, qcap, gp, h->g, th,
                                                nullptr, 0.f, 1, 0.f, check_orientation, h->d_sp_cand.as<uint4>(),
                                                h->d_sp_cnt.as<unsigned char>(), h->d_sp_match.as<int>(), h->d_sp_nm.as<int>(), d_lk, max_dist, kf);
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

int orb_search_by_projection_kf(orb_handle* h, const orb_proj_query* queries, const uint8_t* qdesc, const int32_t* nq, int qcap,
                                const uint8_t* locked0, float th, int orb_dist, int check_orientation, int32_t* match_out,
                                int32_t* nmatches_out, int flags) {
  return search_by_projection_kf_impl(h, queries, qdesc, nq, qcap, locked0, th, orb_dist, check_orientation, match_out, nmatches_out, flags, 1);
}

int orb_search_by_projection_sim3(orb_handle* h, const orb_proj_query* queries, const uint8_t* qdesc, const int32_t* nq, int qcap,
                                  const uint8_t* matched0, float th, float ratio_hamming, int32_t* match_out, int32_t* nmatches_out,
                                  int flags) {
  // bestDist <= TH_LOW * ratioHamming (:487, :594): an int against a float product
  const float lim = 50.0f * ratio_hamming;
  if (!(lim >= 0.f)) return ORB_ERR_INVALID_ARG;
  return search_by_projection_kf_impl(h, queries, qdesc, nq, qcap, matched0, th, (int)std::floor(std::min(lim, 255.0f)), 0, match_out, nmatches_out,
                                      flags, 2);
}

int orb_fuse_search(orb_handle* h, const orb_fuse_query* queries, const uint8_t* qdesc, const int32_t* nq, int qcap, float th, int mode,
                    int32_t* best_idx_out, int32_t* best_dist_out, int flags) {
  if (!h || !queries || !qdesc || !nq || qcap < 1 || (mode != 0 && mode != 1)) return ORB_ERR_INVALID_ARG;
  if (!h->have_grid) return orb_set_error(h, ORB_ERR_STATE, "orb_assign_features_to_grid has not run on this handle");
  int st;
  if ((st = orb_use_device(h))) return st;
  const int batch = h->cur_batch, kcap = h->g.kcap;
  const size_t nqt = (size_t)batch * qcap;
  const orb_fuse_query* d_q = queries;
  const uint8_t* d_qd = qdesc;
  const int* d_nq = nq;
  if (!(flags & ORB_SRC_DEVICE)) {
    const size_t b_q = nqt * sizeof(orb_fuse_query), b_d = nqt * 32, b_n = (size_t)batch * 4;
    const size_t o_d = (b_q + 255) & ~(size_t)255, o_n = o_d + ((b_d + 255) & ~(size_t)255);
    if ((st = orb_ensure(h, h->d_scratch, o_n + b_n + 16))) return st;
    uint8_t* base = h->d_scratch.as<uint8_t>();
    ORB_CUDA_CHECK(h, cudaMemcpyAsync(base, queries, b_q, cudaMemcpyHostToDevice, h->stream));
    ORB_CUDA_CHECK(h, cudaMemcpyAsync(base + o_d, qdesc, b_d, cudaMemcpyHostToDevice, h->stream));
    ORB_CUDA_CHECK(h, cudaMemcpyAsync(base + o_n, nq, b_n, cudaMemcpyHostToDevice, h->stream));
    d_q = (const orb_fuse_query*)base; d_qd = base + o_d; d_nq = (const int*)(base + o_n);
  }
  // results: best index | best distance, qcap each per frame (d_sp_cand is free here: no candidate lists)
  if ((st = orb_ensure(h, h->d_sp_cand, nqt * 2 * sizeof(int)))) return st;
  int* d_bi = h->d_sp_cand.as<int>();
  int* d_bd = d_bi + nqt;
  FuLevels lv;
  for (int l = 0; l < ORB_MAX_LEVELS; ++l) lv.inv_sigma2[l] = l < (int)h->inv_sigma2.size() ? h->inv_sigma2[l] : 0.f;
  const float* d_ur = h->have_stereo ? h->d_uright.as<float>() : nullptr;   // mvuRight = -1 everywhere for a monocular keyframe
  k_fuse_search<<<dim3((qcap + SP_WARPS - 1) / SP_WARPS, batch), SP_WARPS * 32, 0, h->stream>>>(
      orb_keys_un(h), h->d_desc.as<uint8_t>(), d_ur, kcap, h->d_grid_off.as<int>(), h->d_grid_idx.as<unsigned short>(), d_q, d_qd, d_nq, qcap,
      to_gp(&h->grid_params), h->g, lv, th, mode, d_bi, d_bd);
  h->launches++;
  ORB_CUDA_CHECK(h, cudaGetLastError());
  if (!(flags & ORB_NO_OUTPUT)) {
    if (best_idx_out) ORB_CUDA_CHECK(h, cudaMemcpyAsync(best_idx_out, d_bi, nqt * sizeof(int), cudaMemcpyDefault, h->stream));
    if (best_dist_out) ORB_CUDA_CHECK(h, cudaMemcpyAsync(best_dist_out, d_bd, nqt * sizeof(int), cudaMemcpyDefault, h->stream));
  }
  if (flags & ORB_ASYNC) return ORB_OK;
  ORB_CUDA_CHECK(h, cudaStreamSynchronize(h->stream));
  return ORB_OK;
}

// ---- two-camera searches (see the kernel section above) ----
static int check_pair(orb_handle* hL, orb_handle* hR) {
  if (!hL->have_batch || !hR->have_batch) return orb_set_error(hL, ORB_ERR_STATE, "the two-camera search needs an extraction on both handles");
  if (!hL->have_grid || !hR->have_grid) return orb_set_error(hL, ORB_ERR_STATE, "orb_assign_features_to_grid has not run on both handles");
  if (hL->device != hR->device || hL->cur_batch != hR->cur_batch || hL == hR)
    return orb_set_error(hL, ORB_ERR_INVALID_ARG, "left / right handles must be two handles of one device with equal batches");
  if (std::memcmp(&hL->grid_params, &hR->grid_params, sizeof(orb_grid_params)) != 0)
    return orb_set_error(hL, ORB_ERR_INVALID_ARG, "both cameras share the grid bounds (Frame::mnMinX ... are static members)");
  if (hL->g.kcap > 65535 || hR->g.kcap > 65535) return orb_set_error(hL, ORB_ERR_CAPACITY, "more than 65535 keypoints per frame");
  return ORB_OK;
}

// carve `n` regions out of one device buffer (256-byte aligned)
static int carve(orb_handle* h, DevBuf& buf, const size_t* bytes, int n, uint8_t** out) {
  size_t off = 0;
  std::vector<size_t> o(n);
  for (int i = 0; i < n; ++i) { o[i] = off; off += (bytes[i] + 255) & ~(size_t)255; }
  const int st = orb_ensure(h, buf, off + 256);
  if (st) return st;
  for (int i = 0; i < n; ++i) out[i] = buf.as<uint8_t>() + o[i];
  return ORB_OK;
}

int orb_search_by_projection_stereo(orb_handle* hL, orb_handle* hR, const orb_proj_query2* queries, const uint8_t* qdesc, const int32_t* nq,
                                    int qcap, float th, int mono, const float* tlc_z, float mb, int check_orientation,
                                    int32_t* match_left_out, int32_t* match_right_out, int32_t* nmatches_out, int flags) {
  if (!hL || !hR || !queries || !qdesc || !nq || !tlc_z || qcap < 1) return ORB_ERR_INVALID_ARG;
  if (qcap > 32767) return orb_set_error(hL, ORB_ERR_CAPACITY, "more than 32767 queries per frame");
  int st;
  if ((st = check_pair(hL, hR))) return st;
  if ((st = orb_use_device(hL))) return st;
  const int batch = hL->cur_batch, kL = hL->g.kcap, kR = hR->g.kcap;
  if (kL + kR > 65535) return orb_set_error(hL, ORB_ERR_CAPACITY, "more than 65535 keypoints in both cameras");
  const size_t smem = sp2_resolve_smem(qcap, kL, kR);
  if (smem > 160 * 1024) return orb_set_error(hL, ORB_ERR_CAPACITY, "too many queries / keypoints per frame for the resolver");
  const size_t nqt = (size_t)batch * qcap;
  const bool host_in = !(flags & ORB_SRC_DEVICE);
  // regions: queries, descriptors, nq, tlc_z (host inputs) | cand L, cnt L, cand R, cnt R, area, match L, match R, nmatches
  const size_t bytes[12] = {host_in ? nqt * sizeof(orb_proj_query2) : 0, host_in ? nqt * 32 : 0, host_in ? (size_t)batch * 4 : 0,
                            host_in ? (size_t)batch * 4 : 0, nqt * 16, nqt, nqt * 16, nqt, nqt, (size_t)batch * kL * 4, (size_t)batch * kR * 4,
                            (size_t)batch * 4};
  uint8_t* r[12];
  if ((st = carve(hL, hL->d_sp2, bytes, 12, r))) return st;
  const orb_proj_query2* d_q = queries; const uint8_t* d_qd = qdesc; const int* d_nq = nq; const float* d_tz = tlc_z;
  if (host_in) {
    ORB_CUDA_CHECK(hL, cudaMemcpyAsync(r[0], queries, bytes[0], cudaMemcpyHostToDevice, hL->stream));
    ORB_CUDA_CHECK(hL, cudaMemcpyAsync(r[1], qdesc, bytes[1], cudaMemcpyHostToDevice, hL->stream));
    ORB_CUDA_CHECK(hL, cudaMemcpyAsync(r[2], nq, bytes[2], cudaMemcpyHostToDevice, hL->stream));
    ORB_CUDA_CHECK(hL, cudaMemcpyAsync(r[3], tlc_z, bytes[3], cudaMemcpyHostToDevice, hL->stream));
    d_q = (const orb_proj_query2*)r[0]; d_qd = r[1]; d_nq = (const int*)r[2]; d_tz = (const float*)r[3];
  }
  if ((st = orb_peer_read_begin(hL, hR))) return st;
  const GridParams gp = to_gp(&hL->grid_params);
  const dim3 wgrid((qcap + SP_WARPS - 1) / SP_WARPS, batch);
  k_sp2_window<0><<<wgrid, SP_WARPS * 32, 0, hL->stream>>>(orb_keys_un(hL), hL->d_desc.as<uint8_t>(), kL, hL->d_grid_off.as<int>(),
                                                          hL->d_grid_idx.as<unsigned short>(), d_q, d_qd, d_nq, qcap, gp, hL->g, th, d_tz, mb, mono,
                                                          r[8], (uint4*)r[4], r[5]);
  k_sp2_window<1><<<wgrid, SP_WARPS * 32, 0, hL->stream>>>(orb_keys_un(hR), hR->d_desc.as<uint8_t>(), kR, hR->d_grid_off.as<int>(),
                                                          hR->d_grid_idx.as<unsigned short>(), d_q, d_qd, d_nq, qcap, gp, hL->g, th, d_tz, mb, mono,
                                                          r[8], (uint4*)r[6], r[7]);
  Sp2Side SL = {orb_keys_un(hL), hL->d_desc.as<uint8_t>(), hL->d_n.as<int>(), kL, hL->d_grid_off.as<int>(), hL->d_grid_idx.as<unsigned short>(),
                (const uint4*)r[4], r[5], (int*)r[9]};
  Sp2Side SR = {orb_keys_un(hR), hR->d_desc.as<uint8_t>(), hR->d_n.as<int>(), kR, hR->d_grid_off.as<int>(), hR->d_grid_idx.as<unsigned short>(),
                (const uint4*)r[6], r[7], (int*)r[10]};
  if ((st = orb_raise_dyn_smem(hL, (const void*)k_sp2_resolve, smem))) return st;
  k_sp2_resolve<<<batch, 128, smem, hL->stream>>>(SL, SR, d_q, d_qd, d_nq, qcap, gp, hL->g, th, d_tz, mb, mono, check_orientation, (int*)r[11]);
  hL->launches += 3;
  ORB_CUDA_CHECK(hL, cudaGetLastError());
  if ((st = orb_peer_read_end(hL, hR))) return st;
  if (!(flags & ORB_NO_OUTPUT)) {
    if (match_left_out) ORB_CUDA_CHECK(hL, cudaMemcpyAsync(match_left_out, r[9], bytes[9], cudaMemcpyDefault, hL->stream));
    if (match_right_out) ORB_CUDA_CHECK(hL, cudaMemcpyAsync(match_right_out, r[10], bytes[10], cudaMemcpyDefault, hL->stream));
    if (nmatches_out) ORB_CUDA_CHECK(hL, cudaMemcpyAsync(nmatches_out, r[11], bytes[11], cudaMemcpyDefault, hL->stream));
  }
  if (flags & ORB_ASYNC) return ORB_OK;
  ORB_CUDA_CHECK(hL, cudaStreamSynchronize(hL->stream));
  return ORB_OK;
}

int orb_search_local_points_stereo(orb_handle* hL, orb_handle* hR, const orb_track_query2* queries, const uint8_t* qdesc, const int32_t* nq,
                                   int qcap, const uint8_t* locked0_left, const uint8_t* locked0_right, const int32_t* left_to_right,
                                   const int32_t* right_to_left, float th, float nnratio, int32_t* match_left_out,
                                   int32_t* match_right_out, int32_t* nmatches_out, int flags) {
  if (!hL || !hR || !queries || !qdesc || !nq || qcap < 1) return ORB_ERR_INVALID_ARG;
  if ((left_to_right == nullptr) != (right_to_left == nullptr)) return ORB_ERR_INVALID_ARG;
  int st;
  if ((st = check_pair(hL, hR))) return st;
  if (!left_to_right && !hL->have_fe_tri)
    return orb_set_error(hL, ORB_ERR_STATE, "no stereo pairing: pass left_to_right / right_to_left or run orb_stereo_fisheye_triangulate_batch");
  if ((st = orb_use_device(hL))) return st;
  if (hL->g.nlevels > 16) return orb_set_error(hL, ORB_ERR_CAPACITY, "more than 16 pyramid levels");
  const int batch = hL->cur_batch, kL = hL->g.kcap, kR = hR->g.kcap;
  if (kL + kR > 65535) return orb_set_error(hL, ORB_ERR_CAPACITY, "more than 65535 keypoints in both cameras");
  const size_t smem = sl2_resolve_smem(kL, kR);
  if (smem > 120 * 1024) return orb_set_error(hL, ORB_ERR_CAPACITY, "too many keypoints per frame for the resolver");
  const size_t nqt = (size_t)batch * qcap;
  const bool host_in = !(flags & ORB_SRC_DEVICE);
  // regions: queries, descriptors, nq, locked L, locked R, l2r, r2l (host inputs) | split queries L / R, cand L, cnt L, cand R, cnt R,
  //          match L, match R, nmatches
  const size_t bytes[16] = {host_in ? nqt * sizeof(orb_track_query2) : 0, host_in ? nqt * 32 : 0, host_in ? (size_t)batch * 4 : 0,
                            (host_in && locked0_left) ? (size_t)batch * kL : 0, (host_in && locked0_right) ? (size_t)batch * kR : 0,
                            (host_in && left_to_right) ? (size_t)batch * kL * 4 : 0, (host_in && right_to_left) ? (size_t)batch * kR * 4 : 0,
                            nqt * sizeof(orb_track_query), nqt * sizeof(orb_track_query), nqt * 16, nqt, nqt * 16, nqt,
                            (size_t)batch * kL * 4, (size_t)batch * kR * 4, (size_t)batch * 4};
  uint8_t* r[16];
  if ((st = carve(hL, hL->d_sp2, bytes, 16, r))) return st;
  const orb_track_query2* d_q = queries; const uint8_t* d_qd = qdesc; const int* d_nq = nq;
  const uint8_t *d_lkL = locked0_left, *d_lkR = locked0_right;
  const int *d_l2r = left_to_right, *d_r2l = right_to_left;
  if (host_in) {
    const void* src[7] = {queries, qdesc, nq, locked0_left, locked0_right, left_to_right, right_to_left};
    for (int i = 0; i < 7; ++i)
      if (bytes[i]) ORB_CUDA_CHECK(hL, cudaMemcpyAsync(r[i], src[i], bytes[i], cudaMemcpyHostToDevice, hL->stream));
    d_q = (const orb_track_query2*)r[0]; d_qd = r[1]; d_nq = (const int*)r[2];
    d_lkL = locked0_left ? r[3] : nullptr; d_lkR = locked0_right ? r[4] : nullptr;
    if (left_to_right) { d_l2r = (const int*)r[5]; d_r2l = (const int*)r[6]; }
  }
  if (!left_to_right) { d_l2r = hL->d_fe_l2r.as<int>(); d_r2l = hL->d_fe_r2l.as<int>(); }   // [batch][kL] / [batch][kR] of the triangulation
  if ((st = orb_peer_read_begin(hL, hR))) return st;
  const GridParams gp = to_gp(&hL->grid_params);
  orb_track_query* qL = (orb_track_query*)r[7];
  orb_track_query* qR = (orb_track_query*)r[8];
  k_split_track_queries<<<(unsigned)((nqt + 255) / 256), 256, 0, hL->stream>>>(d_q, (int)nqt, qL, qR);
  const dim3 wgrid((qcap + SP_WARPS - 1) / SP_WARPS, batch);
  // no lock is applied in the window kernels: in a two-camera frame a lock can be released again (partner overwrite)
  k_sl_window<<<wgrid, SP_WARPS * 32, 0, hL->stream>>>(orb_keys_un(hL), hL->d_desc.as<uint8_t>(), nullptr, kL, hL->d_grid_off.as<int>(),
                                                      hL->d_grid_idx.as<unsigned short>(), qL, d_qd, d_nq, qcap, nullptr, gp, hL->g, th,
                                                      (uint4*)r[9], r[10]);
  k_sl_window<<<wgrid, SP_WARPS * 32, 0, hL->stream>>>(orb_keys_un(hR), hR->d_desc.as<uint8_t>(), nullptr, kR, hR->d_grid_off.as<int>(),
                                                      hR->d_grid_idx.as<unsigned short>(), qR, d_qd, d_nq, qcap, nullptr, gp, hL->g, 1.0f,
                                                      (uint4*)r[11], r[12]);
  Sl2Side SL = {orb_keys_un(hL), hL->d_desc.as<uint8_t>(), hL->d_n.as<int>(), kL, hL->d_grid_off.as<int>(), hL->d_grid_idx.as<unsigned short>(), qL,
                (const uint4*)r[9], r[10], d_lkL, d_l2r, (int*)r[13]};
  Sl2Side SR = {orb_keys_un(hR), hR->d_desc.as<uint8_t>(), hR->d_n.as<int>(), kR, hR->d_grid_off.as<int>(), hR->d_grid_idx.as<unsigned short>(), qR,
                (const uint4*)r[11], r[12], d_lkR, d_r2l, (int*)r[14]};
  if ((st = orb_raise_dyn_smem(hL, (const void*)k_sl2_resolve, smem))) return st;
  k_sl2_resolve<<<batch, 128, smem, hL->stream>>>(SL, SR, d_qd, d_nq, qcap, gp, hL->g, th, nnratio, (int*)r[15]);
  hL->launches += 4;
  ORB_CUDA_CHECK(hL, cudaGetLastError());
  if ((st = orb_peer_read_end(hL, hR))) return st;
  if (!(flags & ORB_NO_OUTPUT)) {
    if (match_left_out) ORB_CUDA_CHECK(hL, cudaMemcpyAsync(match_left_out, r[13], bytes[13], cudaMemcpyDefault, hL->stream));
    if (match_right_out) ORB_CUDA_CHECK(hL, cudaMemcpyAsync(match_right_out, r[14], bytes[14], cudaMemcpyDefault, hL->stream));
    if (nmatches_out) ORB_CUDA_CHECK(hL, cudaMemcpyAsync(nmatches_out, r[15], bytes[15], cudaMemcpyDefault, hL->stream));
  }
  if (flags & ORB_ASYNC) return ORB_OK;
  ORB_CUDA_CHECK(hL, cudaStreamSynchronize(hL->stream));
  return ORB_OK;
}

int orb_search_by_bow_stereo(orb_handle* hL, orb_handle* hR, const orb_bow_keyframes* kf, float nnratio, int check_orientation,
                             int32_t* match_left_out, int32_t* match_right_out, int32_t* nmatches_out, int flags) {
  if (!hL || !hR || hL == hR || !kf || !kf->desc || !kf->angle || !kf->flags || !kf->n || !kf->fv_node || !kf->fv_off || !kf->fv_feat ||
      !kf->fv_n || kf->cap < 1)
    return ORB_ERR_INVALID_ARG;
  if (!hL->have_batch || !hR->have_batch || !hL->have_bow2)
    return orb_set_error(hL, ORB_ERR_STATE, "orb_compute_bow_stereo has not run on this pair's batch");
  if (hL->device != hR->device || hL->cur_batch != hR->cur_batch) return orb_set_error(hL, ORB_ERR_INVALID_ARG, "left / right handles differ in device or batch");
  int st;
  if ((st = orb_use_device(hL))) return st;
  const int batch = hL->cur_batch, kL = hL->g.kcap, kR = hR->g.kcap, cap2 = kL + kR, qcap = kf->cap;
  if (cap2 != hL->bow2_cap) return orb_set_error(hL, ORB_ERR_STATE, "the pair's capacity changed since orb_compute_bow_stereo");
  if (cap2 > 65535) return orb_set_error(hL, ORB_ERR_CAPACITY, "more than 65535 keypoints in both cameras");
  const size_t smem = (size_t)cap2 * 12;
  if (smem > 200 * 1024) return orb_set_error(hL, ORB_ERR_CAPACITY, "too many keypoints per frame");
  const size_t nq = (size_t)batch * qcap;
  const uint8_t *d_desc = kf->desc, *d_flags = kf->flags;
  const float* d_angle = kf->angle;
  const int *d_n = kf->n, *d_off = kf->fv_off, *d_nn = kf->fv_n;
  const unsigned int *d_node = kf->fv_node, *d_feat = kf->fv_feat;
  // regions: keyframe inputs (desc | angle | node | feat | off | flags | n | nn) | match L, match R, nmatches
  const bool host_in = !(flags & ORB_SRC_DEVICE);
  const size_t bytes[11] = {host_in ? nq * 32 : 0, host_in ? nq * 4 : 0, host_in ? nq * 4 : 0, host_in ? nq * 4 : 0,
                            host_in ? (size_t)batch * (qcap + 1) * 4 : 0, host_in ? nq : 0, host_in ? (size_t)batch * 4 : 0,
                            host_in ? (size_t)batch * 4 : 0, (size_t)batch * kL * 4, (size_t)batch * kR * 4, (size_t)batch * 4};
  uint8_t* r[11];
  if ((st = carve(hL, hL->d_sp2, bytes, 11, r))) return st;
  if (host_in) {
    const void* src[8] = {kf->desc, kf->angle, kf->fv_node, kf->fv_feat, kf->fv_off, kf->flags, kf->n, kf->fv_n};
    for (int i = 0; i < 8; ++i) ORB_CUDA_CHECK(hL, cudaMemcpyAsync(r[i], src[i], bytes[i], cudaMemcpyHostToDevice, hL->stream));
    d_desc = r[0]; d_angle = (const float*)r[1]; d_node = (const unsigned int*)r[2]; d_feat = (const unsigned int*)r[3];
    d_off = (const int*)r[4]; d_flags = r[5]; d_n = (const int*)r[6]; d_nn = (const int*)r[7];
  }
  if ((st = orb_peer_read_begin(hL, hR))) return st;
  if ((st = orb_raise_dyn_smem(hL, (const void*)k_sbow2, smem))) return st;
  uint8_t** b2 = hL->bow2_r;
  k_sbow2<<<batch, 256, smem, hL->stream>>>(hL->d_kps.as<orb_keypoint>(), hL->d_desc.as<uint8_t>(), hL->d_n.as<int>(), kL,
                                            hR->d_kps.as<orb_keypoint>(), hR->d_desc.as<uint8_t>(), hR->d_n.as<int>(), kR,
                                            (const unsigned int*)b2[13], (const int*)b2[14], (const unsigned int*)b2[15], (const int*)b2[10] + batch,
                                            d_desc, d_angle, d_flags, d_n, qcap, d_node, d_off, d_feat, d_nn, nnratio, check_orientation,
                                            (int*)r[8], (int*)r[9], (int*)r[10]);
  hL->launches++;
  ORB_CUDA_CHECK(hL, cudaGetLastError());
  if ((st = orb_peer_read_end(hL, hR))) return st;
  if (!(flags & ORB_NO_OUTPUT)) {
    if (match_left_out) ORB_CUDA_CHECK(hL, cudaMemcpyAsync(match_left_out, r[8], bytes[8], cudaMemcpyDefault, hL->stream));
    if (match_right_out) ORB_CUDA_CHECK(hL, cudaMemcpyAsync(match_right_out, r[9], bytes[9], cudaMemcpyDefault, hL->stream));
    if (nmatches_out) ORB_CUDA_CHECK(hL, cudaMemcpyAsync(nmatches_out, r[10], bytes[10], cudaMemcpyDefault, hL->stream));
  }
  if (flags & ORB_ASYNC) return ORB_OK;
  ORB_CUDA_CHECK(hL, cudaStreamSynchronize(hL->stream));
  return ORB_OK;
}

int orb_search_for_initialization(orb_handle* h, const orb_init_query* queries, const uint8_t* qdesc, const int32_t* nq, int qcap,
                                  int window_size, float nnratio, int check_orientation, int32_t* matches12_out, float* prev_matched_out,
                                  int32_t* nmatches_out, int flags) {
  if (!h || !queries || !qdesc || !nq || qcap < 1 || window_size < 0) return ORB_ERR_INVALID_ARG;
  if (!h->have_grid) return orb_set_error(h, ORB_ERR_STATE, "orb_assign_features_to_grid has not run on this handle");
  int st;
  if ((st = orb_use_device(h))) return st;
  const int batch = h->cur_batch, kcap = h->g.kcap;
  if (qcap > 65535 || kcap > 65535) return orb_set_error(h, ORB_ERR_CAPACITY, "more than 65535 keypoints per frame");
  const size_t smem = sfi_smem(qcap, kcap);
  if (smem > 160 * 1024) return orb_set_error(h, ORB_ERR_CAPACITY, "too many keypoints per frame for the initialisation matcher");
  const size_t nqt = (size_t)batch * qcap;
  const orb_init_query* d_q = queries;
  const uint8_t* d_qd = qdesc;
  const int* d_nq = nq;
  if (!(flags & ORB_SRC_DEVICE)) {
    const size_t b_q = nqt * sizeof(orb_init_query), b_d = nqt * 32, b_n = (size_t)batch * 4;
    const size_t o_d = (b_q + 255) & ~(size_t)255, o_n = o_d + ((b_d + 255) & ~(size_t)255);
    if ((st = orb_ensure(h, h->d_scratch, o_n + b_n + 16))) return st;
    uint8_t* base = h->d_scratch.as<uint8_t>();
    ORB_CUDA_CHECK(h, cudaMemcpyAsync(base, queries, b_q, cudaMemcpyHostToDevice, h->stream));
    ORB_CUDA_CHECK(h, cudaMemcpyAsync(base + o_d, qdesc, b_d, cudaMemcpyHostToDevice, h->stream));
    ORB_CUDA_CHECK(h, cudaMemcpyAsync(base + o_n, nq, b_n, cudaMemcpyHostToDevice, h->stream));
    d_q = (const orb_init_query*)base; d_qd = base + o_d; d_nq = (const int*)(base + o_n);
  }
  // candidate lists | counts | vnMatches12 | vbPrevMatched | nmatches
  const size_t bytes[5] = {nqt * SFI_CAP * sizeof(unsigned int), nqt * sizeof(int), nqt * sizeof(int), nqt * sizeof(float2), (size_t)batch * sizeof(int)};
  uint8_t* r[5];
  if ((st = carve(h, h->d_sp_cand, bytes, 5, r))) return st;
  { const int st_a = orb_raise_dyn_smem(h, (const void*)k_search_for_initialization, smem); if (st_a) return st_a; }
  k_search_for_initialization<<<batch, SFI_THREADS, smem, h->stream>>>(
      orb_keys_un(h), h->d_desc.as<uint8_t>(), h->d_n.as<int>(), kcap, h->d_grid_off.as<int>(), h->d_grid_idx.as<unsigned short>(), d_q, d_qd,
      d_nq, qcap, to_gp(&h->grid_params), (float)window_size, nnratio, check_orientation, (unsigned int*)r[0], (int*)r[1], (int*)r[2],
      (float2*)r[3], (int*)r[4]);
  h->launches++;
  ORB_CUDA_CHECK(h, cudaGetLastError());
  if (!(flags & ORB_NO_OUTPUT)) {
    if (matches12_out) ORB_CUDA_CHECK(h, cudaMemcpyAsync(matches12_out, r[2], bytes[2], cudaMemcpyDefault, h->stream));
    if (prev_matched_out) ORB_CUDA_CHECK(h, cudaMemcpyAsync(prev_matched_out, r[3], bytes[3], cudaMemcpyDefault, h->stream));
    if (nmatches_out) ORB_CUDA_CHECK(h, cudaMemcpyAsync(nmatches_out, r[4], bytes[4], cudaMemcpyDefault, h->stream));
  }
  if (flags & ORB_ASYNC) return ORB_OK;
  ORB_CUDA_CHECK(h, cudaStreamSynchronize(h->stream));
  return ORB_OK;
}

}  // extern "C"
