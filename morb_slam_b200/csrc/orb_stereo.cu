// liborb_b200.so - rectified stereo matching: Frame::ComputeStereoMatches (reference src/Frame.cc:889-1047)
// restated per left keypoint (order-free form, SURVEY.md a12'):
//   k_stereo_rows     row table of the right keypoints (:898-912): per image row the entries {index | octave << 16, x} of the
//                     keypoints whose band covers it - what the matcher's gates need, so it never loads a right keypoint record;
//   k_stereo_match_h  one HALF-WARP per left keypoint (default): octave / disparity-range gates over the entries of its row,
//                     256-bit Hamming (uint4 loads + __popc), arg-min of (dist, iR) over the half, then the 11x11 SAD over 11
//                     shifts on the un-blurred pyramids and the parabola fit; k_stereo_match = the same with one warp per
//                     keypoint (ORB_B200_STEREO=warp, kept for A/B measurements);
//   k_stereo_gate   one CTA per frame: median of the accepted SADs by a two-level radix select,
//                   rejection of matches with SAD >= 1.5 * 1.4 * median (:1035-1046).
#include <algorithm>
#include <cstdlib>
#include <cstring>

#include "orb_internal.h"

#define ST_WARPS 8
#define TH_HIGH 100      // src/ORBmatcher.cc:34
#define TH_ORB_DIST 75   // (TH_HIGH + TH_LOW) / 2, src/Frame.cc:893

static __device__ __forceinline__ const uint8_t* st_lvl_ptr(const OrbGeom& g, const uint8_t* base, int frame, int l) {
  return base + g.level_base[l] + (size_t)frame * g.level_fstride[l];
}

// Row table of the right keypoints (:898-912): keypoint iR is registered in every image row of
// [floor(y - r), ceil(y + r)], r = 2 * scaleFactor[octave]. One CTA per frame: histogram of rows in shared
// memory, block scan, fill. The order inside a row is irrelevant because the matcher takes the
// lexicographic minimum of (distance, iR), which equals the reference's ascending scan with strict <.
#define SR_THREADS 1024   // one CTA per frame: the two atomic passes over (keypoint, row) pairs are the kernel (a single pair's latency path)
__global__ void __launch_bounds__(SR_THREADS) k_stereo_rows(OrbGeom gL, int kcapR, int items_cap, const orb_keypoint* __restrict__ kpsR,
                                                            const int* __restrict__ nR_arr, int* __restrict__ row_off,
                                                            uint2* __restrict__ row_items) {
  extern __shared__ int s_hist[];  // [H + 1] counts, then cursors
  __shared__ int s_warp[SR_THREADS / 32];
  __shared__ int s_carry;
  const int frame = blockIdx.x, tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const int H = gL.h[0];
  const int nR = nR_arr[frame];
  const orb_keypoint* kR = kpsR + (size_t)frame * kcapR;
  int* off = row_off + (size_t)frame * (H + 1);
  uint2* items = row_items + (size_t)frame * items_cap;
  for (int i = tid; i <= H; i += SR_THREADS) s_hist[i] = 0;
  if (tid == 0) s_carry = 0;
  __syncthreads();
  int minr = 0, maxr = -1, minr2 = 0, maxr2 = -1;   // a thread's (at most two) keypoints: nR <= 2 * SR_THREADS in practice, more loop below
  for (int i = tid, k = 0; i < nR; i += SR_THREADS, ++k) {
    const float yR = kR[i].y;
    const float r = __fmul_rn(2.0f, gL.scale[kR[i].octave]);
    const int hi = min((int)ceilf(__fadd_rn(yR, r)), H - 1);
    const int lo = max((int)floorf(__fsub_rn(yR, r)), 0);
    if (k == 0) { minr = lo; maxr = hi; } else if (k == 1) { minr2 = lo; maxr2 = hi; }
    for (int y = lo; y <= hi; ++y) atomicAdd(&s_hist[y], 1);
  }
  __syncthreads();
  // exclusive scan over rows, SR_THREADS rows per sweep
  for (int base = 0; base < H; base += SR_THREADS) {
    const int y = base + tid;
    const int c = (y < H) ? s_hist[y] : 0;
    int incl = c;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int v = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= o) incl += v;
    }
    if (lane == 31) s_warp[wid] = incl;
    __syncthreads();
    int before = s_carry, total = 0;
    for (int k = 0; k < SR_THREADS / 32; ++k) { const int v = s_warp[k]; if (k < wid) before += v; total += v; }
    const int start = before + incl - c;
    if (y < H) { off[y] = start; s_hist[y] = start; }
    __syncthreads();
    if (tid == 0) s_carry += total;
    __syncthreads();
  }
  if (tid == 0) off[H] = s_carry;
  for (int i = tid, k = 0; i < nR; i += SR_THREADS, ++k) {
    int lo, hi;
    if (k == 0) { lo = minr; hi = maxr; }
    else if (k == 1) { lo = minr2; hi = maxr2; }
    else {
      const float yR = kR[i].y;
      const float r = __fmul_rn(2.0f, gL.scale[kR[i].octave]);
      hi = min((int)ceilf(__fadd_rn(yR, r)), H - 1);
      lo = max((int)floorf(__fsub_rn(yR, r)), 0);
    }
    // an entry carries what the matcher's gates need (index | octave << 16, x): the matcher never touches the keypoint record
    const uint2 ent = make_uint2((uint32_t)i | ((uint32_t)kR[i].octave << 16), __float_as_uint(kR[i].x));
    for (int y = lo; y <= hi; ++y) {
      const int pos = atomicAdd(&s_hist[y], 1);
      if (pos < items_cap) items[pos] = ent;
    }
  }
}

#ifndef ST_MINB
#define ST_MINB 8   // latency-bound kernel: 32 registers (40 bytes of spill) for 64 resident warps per SM: 0.346 -> 0.271 ms per 256 pairs (6: 0.319)
#endif
__global__ void __launch_bounds__(ST_WARPS * 32, ST_MINB) k_stereo_match(
    OrbGeom gL, OrbGeom gR, const uint8_t* __restrict__ pyrL, const uint8_t* __restrict__ pyrR,
    const orb_keypoint* __restrict__ kpsL, const uint8_t* __restrict__ descL, const int* __restrict__ nL_arr,
    const orb_keypoint* __restrict__ kpsR, const uint8_t* __restrict__ descR, const int* __restrict__ nR_arr,
    float mbf, float maxD, const int* __restrict__ row_off, const uint2* __restrict__ row_items, int items_cap,
    float* __restrict__ uright, float* __restrict__ depth, int* __restrict__ sad_out, int* __restrict__ best_idx,
    int* __restrict__ best_dist) {
  __shared__ uint32_t s_il[ST_WARPS][11 * 4];      // left patch rows as 4 aligned words
  __shared__ uint32_t s_ir[ST_WARPS][11 * 7 + 3];  // right strip rows as 7 aligned words
  const int frame = blockIdx.y;
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const int iL = blockIdx.x * ST_WARPS + wid;
  const int nL = nL_arr[frame], nR = nR_arr[frame];
  if (iL >= nL) return;
  const size_t oL = (size_t)frame * gL.kcap + iL;
  if (lane == 0) { uright[oL] = -1.f; depth[oL] = -1.f; sad_out[oL] = -1; best_idx[oL] = -1; best_dist[oL] = -1; }
  const orb_keypoint kL = kpsL[oL];
  const int levelL = kL.octave;
  const float vL = kL.y, uL = kL.x;
  const int row = (int)vL;                 // vRowIndices[vL] (:929): truncation
  const float minU = __fsub_rn(uL, maxD);  // :933
  const float maxU = uL;                   // uL - minD, minD = 0
  if (maxU < 0) return;                    // :936
  const uint4* dl = reinterpret_cast<const uint4*>(descL + oL * 32);
  const uint4 a0 = dl[0], a1 = dl[1];
  const orb_keypoint* kR = kpsR + (size_t)frame * gR.kcap;
  const uint8_t* dR = descR + (size_t)frame * gR.kcap * 32;
  uint32_t best = 0xffffffffu;
  (void)nR;
  const int H0 = gL.h[0];
  if (row < 0 || row >= H0) return;  // vRowIndices[vL] is only defined for rows of the image
  const int* off = row_off + (size_t)frame * (H0 + 1);
  const uint2* items = row_items + (size_t)frame * items_cap;
  const int c0 = off[row], c1 = min(off[row + 1], items_cap);
  for (int ic = c0 + lane; ic < c1; ic += 32) {
    const int iR = (int)(items[ic].x & 0xffffu);
    const float uR = kR[iR].x;
    const int octR = kR[iR].octave;
    if (octR < levelL - 1 || octR > levelL + 1) continue;     // :948
    if (!(uR >= minU && uR <= maxU)) continue;                // :952
    const uint4* dr = reinterpret_cast<const uint4*>(dR + (size_t)iR * 32);
    const uint4 b0 = dr[0], b1 = dr[1];
    const int d = __popc(a0.x ^ b0.x) + __popc(a0.y ^ b0.y) + __popc(a0.z ^ b0.z) + __popc(a0.w ^ b0.w) +
                  __popc(a1.x ^ b1.x) + __popc(a1.y ^ b1.y) + __popc(a1.z ^ b1.z) + __popc(a1.w ^ b1.w);
    if (d < TH_HIGH) best = min(best, ((uint32_t)d << 16) | (uint32_t)iR);  // min (d, iR) == ascending scan with strict <
  }
  best = __reduce_min_sync(0xffffffffu, best);
  if (best == 0xffffffffu) return;
  const int bestDist = (int)(best >> 16), bestR = (int)(best & 0xffffu);
  if (lane == 0) { best_idx[oL] = bestR; best_dist[oL] = bestDist; }
  if (!(bestDist < TH_ORB_DIST)) return;   // :964

  // ---- sub-pixel refinement by SAD at the left keypoint's pyramid level (:966-1003)
  const float uR0 = kR[bestR].x;
  const float sf = gL.inv_scale[levelL];
  const float scaleduL = roundf(__fmul_rn(kL.x, sf));   // std::round: half away from zero
  const float scaledvL = roundf(__fmul_rn(kL.y, sf));
  const float scaleduR0 = roundf(__fmul_rn(uR0, sf));
  const int w = 5, L = 5;
  const float iniu = __fsub_rn(__fadd_rn(scaleduR0, (float)L), (float)w);
  const float endu = __fadd_rn(__fadd_rn(__fadd_rn(scaleduR0, (float)L), (float)w), 1.f);
  const int WL = gL.w[levelL], HL = gL.h[levelL], WR = gR.w[levelL], HR = gR.h[levelL];
  if (iniu < 0 || endu >= (float)WR) return;            // :984-988 verbatim
  const int cy = (int)scaledvL, cxl = (int)scaleduL, cxr = (int)scaleduR0;
  // the reference would throw (cv::Mat range assert) outside these bounds; cannot happen for extractor output
  if (cy - w < 0 || cy + w >= HL || cy + w >= HR || cxl - w < 0 || cxl + w >= WL || cxr - L - w < 0 || cxr + L + w >= WR) return;
  const uint8_t* IL = st_lvl_ptr(gL, pyrL, frame, levelL);
  const uint8_t* IR = st_lvl_ptr(gR, pyrR, frame, levelL);
  const int PL = gL.pitch[levelL], PR = gR.pitch[levelL];
  // Stage the 11 x 11 left patch and the 11 x 21 right strip as ALIGNED words (rows are 16-byte aligned): 4 words
  // per left row from column xl0, 7 words per right row from column xr0 (reads may run a few bytes past the
  // patch, inside the padded pyramid buffer; those bytes are masked out below).
  const int xl0 = (cxl - w) & ~3, xr0 = (cxr - L - w) & ~3;
  uint32_t* sl = s_il[wid];
  uint32_t* sr = s_ir[wid];
  for (int i = lane; i < 11 * 11; i += 32) {
    if (i < 44) {
      const int dy = i >> 2, j = i & 3;
      sl[i] = *reinterpret_cast<const uint32_t*>(IL + (size_t)(cy - w + dy) * PL + xl0 + 4 * j);
    } else {
      const int t = i - 44, dy = t / 7, j = t - dy * 7;
      sr[t] = *reinterpret_cast<const uint32_t*>(IR + (size_t)(cy - w + dy) * PR + xr0 + 4 * j);
    }
  }
  __syncwarp();
  // lane = (shift k = lane & 15, row parity g = lane >> 4): SAD of rows g, g+2, .. for shift k with byte-SIMD
  // |a - b| accumulation (VABSDIFF4.ACC), 4 pixels per instruction; the 11th..12th bytes of a row are masked
  const int k = lane & 15, gpar = lane >> 4;
  const int shl = 8 * ((cxl - w) - xl0);                 // bit offset of the patch inside the left words
  const int sR = (cxr - L - w) - xr0 + min(k, 10);       // byte offset of shift k inside the right words (0..13)
  const int wR = sR >> 2, shr = 8 * (sR & 3);
  uint32_t sad = 0;
#pragma unroll
  for (int dy2 = 0; dy2 < 6; ++dy2) {
    const int dy = 2 * dy2 + gpar;
    if (dy < 11) {
      const uint32_t* a = sl + 4 * dy;
      const uint32_t* b = sr + 7 * dy + wR;
      const uint32_t a0w = a[0], a1w = a[1], a2w = a[2], a3w = a[3];
      const uint32_t b0w = b[0], b1w = b[1], b2w = b[2], b3w = b[3];
      sad = __vsadu4(__funnelshift_r(a0w, a1w, shl), __funnelshift_r(b0w, b1w, shr)) + sad;
      sad = __vsadu4(__funnelshift_r(a1w, a2w, shl), __funnelshift_r(b1w, b2w, shr)) + sad;
      sad = __vsadu4(__funnelshift_r(a2w, a3w, shl) & 0x00ffffffu, __funnelshift_r(b2w, b3w, shr) & 0x00ffffffu) + sad;
    }
  }
  sad += __shfl_xor_sync(0xffffffffu, sad, 16);
  // best shift: ascending scan with strict < (:1001-1004) == lexicographic minimum of (sad, k); sad <= 121 * 255
  const uint32_t key = __reduce_min_sync(0xffffffffu, k < 11 ? ((sad << 4) | (uint32_t)k) : 0xffffffffu);
  const int bestSad = (int)(key >> 4), bestK = (int)(key & 15u), bestInc = bestK - L;
  // lanes 0..10 hold the SAD of shift k; the parabola needs the neighbours of the best one
  const float d1 = (float)__shfl_sync(0xffffffffu, sad, max(bestK - 1, 0));
  const float d2 = (float)bestSad;
  const float d3 = (float)__shfl_sync(0xffffffffu, sad, min(bestK + 1, 10));
  if (bestInc == -L || bestInc == L) return;            // :1005
  // deltaR = (dist1 - dist3) / (2.0f * (dist1 + dist3 - 2.0f * dist2))  (:1012-1013)
  const float deltaR = __fdiv_rn(__fsub_rn(d1, d3), __fmul_rn(2.0f, __fsub_rn(__fadd_rn(d1, d3), __fmul_rn(2.0f, d2))));
  if (deltaR < -1 || deltaR > 1) return;                // NaN falls through, rejected by the range test below
  float bestuR = __fmul_rn(gL.scale[levelL], __fadd_rn(__fadd_rn(scaleduR0, (float)bestInc), deltaR));
  float disparity = __fsub_rn(uL, bestuR);
  if (disparity >= 0.f && disparity < maxD) {           // minD = 0
    if (disparity <= 0) {
      disparity = 0.01f;                                // (float)0.01
      bestuR = (float)((double)uL - 0.01);
    }
    if (lane == 0) {
      depth[oL] = __fdiv_rn(mbf, disparity);
      uright[oL] = bestuR;
      sad_out[oL] = bestSad;
    }
  }
}

// ---- the same search with ONE HALF-WARP per left keypoint (the default): a row of the table holds about 20 right keypoints, so a
// full warp ran one mostly empty trip of the candidate loop and the scalar part of the kernel (geometry, gates, parabola) once per
// keypoint; two keypoints per warp halve that part. The SAD stage keeps one lane per shift (11 of 16 lanes) and walks the 11 rows:
// the left patch is staged already ALIGNED (3 words per row, funnel shift done once at staging instead of once per shift), one
// LDS.128 per row for it. All warp-level primitives carry the half's mask; the two halves leave the search at different gates.
#define SH_SUBS (ST_WARPS * 2)
#define SH_PITCH 12    // words per staged row: [3 aligned left words, 1 pad | 7 right words, 1 pad]
#define SH_WORDS 144   // 11 * 12 + 12: consecutive halves sit 16 banks apart (their right-strip loads never share a bank)
__global__ void __launch_bounds__(ST_WARPS * 32, ST_MINB) k_stereo_match_h(
    OrbGeom gL, OrbGeom gR, const uint8_t* __restrict__ pyrL, const uint8_t* __restrict__ pyrR,
    const orb_keypoint* __restrict__ kpsL, const uint8_t* __restrict__ descL, const int* __restrict__ nL_arr,
    const orb_keypoint* __restrict__ kpsR, const uint8_t* __restrict__ descR,
    float mbf, float maxD, const int* __restrict__ row_off, const uint2* __restrict__ row_items, int items_cap,
    float* __restrict__ uright, float* __restrict__ depth, int* __restrict__ sad_out, int* __restrict__ best_idx,
    int* __restrict__ best_dist) {
  __shared__ __align__(16) uint32_t s_p[SH_SUBS][SH_WORDS];
  const int frame = blockIdx.y;
  const int lane = threadIdx.x & 31, hl = lane & 15, hbase = lane & 16, wid = threadIdx.x >> 5;
  const unsigned hmask = 0xffffu << hbase;
  const int sub = 2 * wid + (hbase >> 4);
  const int iL = blockIdx.x * SH_SUBS + sub;
  if (iL >= nL_arr[frame]) return;
  const size_t oL = (size_t)frame * gL.kcap + iL;
  float r_u = -1.f, r_depth = -1.f;
  int r_sad = -1, r_idx = -1, r_dist = -1;
  do {
    const orb_keypoint kL = kpsL[oL];
    const int levelL = kL.octave;
    const float vL = kL.y, uL = kL.x;
    const int row = (int)vL;                 // vRowIndices[vL] (:929): truncation
    const float minU = __fsub_rn(uL, maxD);  // :933
    const float maxU = uL;                   // uL - minD, minD = 0
    if (maxU < 0) break;                     // :936
    const int H0 = gL.h[0];
    if (row < 0 || row >= H0) break;         // vRowIndices[vL] is only defined for rows of the image
    const uint4* dl = reinterpret_cast<const uint4*>(descL + oL * 32);
    const uint4 a0 = dl[0], a1 = dl[1];
    const uint8_t* dR = descR + (size_t)frame * gR.kcap * 32;
    const int* off = row_off + (size_t)frame * (H0 + 1);
    const uint2* items = row_items + (size_t)frame * items_cap;
    const int c0 = off[row], c1 = min(off[row + 1], items_cap);
    uint32_t best = 0xffffffffu;
    float bu = 0.f;                          // x of this lane's best candidate
    for (int ic = c0 + hl; ic < c1; ic += 16) {
      const uint2 ent = items[ic];           // index | octave << 16, x: the gates need no load of the keypoint record
      const int iR = (int)(ent.x & 0xffffu), octR = (int)(ent.x >> 16);
      const float uR = __uint_as_float(ent.y);
      if (octR < levelL - 1 || octR > levelL + 1) continue;     // :948
      if (!(uR >= minU && uR <= maxU)) continue;                // :952
      const uint4* dr = reinterpret_cast<const uint4*>(dR + (size_t)iR * 32);
      const uint4 b0 = dr[0], b1 = dr[1];
      const int d = __popc(a0.x ^ b0.x) + __popc(a0.y ^ b0.y) + __popc(a0.z ^ b0.z) + __popc(a0.w ^ b0.w) +
                    __popc(a1.x ^ b1.x) + __popc(a1.y ^ b1.y) + __popc(a1.z ^ b1.z) + __popc(a1.w ^ b1.w);
      const uint32_t key = ((uint32_t)d << 16) | (uint32_t)iR;   // min (d, iR) == ascending scan with strict <
      if (d < TH_HIGH && key < best) { best = key; bu = uR; }
    }
    const uint32_t mine = best;
    best = __reduce_min_sync(hmask, best);
    if (best == 0xffffffffu) break;
    // keys are unique (they hold iR): exactly one lane of the half owns the winner and its x
    const float uR0 = __shfl_sync(hmask, bu, __ffs(__ballot_sync(hmask, mine == best)) - 1);
    const int bestDist = (int)(best >> 16), bestR = (int)(best & 0xffffu);
    r_idx = bestR; r_dist = bestDist;
    if (!(bestDist < TH_ORB_DIST)) break;    // :964

    // ---- sub-pixel refinement by SAD at the left keypoint's pyramid level (:966-1003)
    const float sf = gL.inv_scale[levelL];
    const float scaleduL = roundf(__fmul_rn(kL.x, sf));   // std::round: half away from zero
    const float scaledvL = roundf(__fmul_rn(kL.y, sf));
    const float scaleduR0 = roundf(__fmul_rn(uR0, sf));
    const int w = 5, L = 5;
    const float iniu = __fsub_rn(__fadd_rn(scaleduR0, (float)L), (float)w);
    const float endu = __fadd_rn(__fadd_rn(__fadd_rn(scaleduR0, (float)L), (float)w), 1.f);
    const int WL = gL.w[levelL], HL = gL.h[levelL], WR = gR.w[levelL], HR = gR.h[levelL];
    if (iniu < 0 || endu >= (float)WR) break;             // :984-988 verbatim
    const int cy = (int)scaledvL, cxl = (int)scaleduL, cxr = (int)scaleduR0;
    // the reference would throw (cv::Mat range assert) outside these bounds; cannot happen for extractor output
    if (cy - w < 0 || cy + w >= HL || cy + w >= HR || cxl - w < 0 || cxl + w >= WL || cxr - L - w < 0 || cxr + L + w >= WR) break;
    const int PL = gL.pitch[levelL], PR = gR.pitch[levelL];
    // Staging as ALIGNED word loads (rows are 16-byte aligned; reads may run a few bytes past the patch, inside the padded pyramid
    // buffer, exactly the words the one-warp kernel reads): item (dy, j) of 11 x 10, j < 3 = left word j of the row shifted to the
    // patch's first byte (two loads + funnel shift, the 12th byte masked), j >= 3 = right word j - 3 from column xr0.
    const int xl0 = (cxl - w) & ~3, xr0 = (cxr - L - w) & ~3;
    const uint32_t shl = 8u * (uint32_t)((cxl - w) - xl0);    // bit offset of the patch inside the left words
    const uint8_t* IL = st_lvl_ptr(gL, pyrL, frame, levelL) + (size_t)(cy - w) * PL + xl0;
    const uint8_t* IR = st_lvl_ptr(gR, pyrR, frame, levelL) + (size_t)(cy - w) * PR + xr0 - 12;   // - 12: word j - 3
    uint32_t* sp = s_p[sub];
#pragma unroll
    for (int t = 0; t < 7; ++t) {
      const int i = hl + 16 * t;
      const int dy = (i * 205) >> 11, j = i - 10 * dy;      // i / 10, i % 10 for i < 128
      if (i < 110) {
        const bool left = j < 3;
        const uint8_t* p = (left ? IL : IR) + (size_t)dy * (left ? PL : PR) + 4 * j;
        const uint32_t w0 = *reinterpret_cast<const uint32_t*>(p);
        uint32_t v = w0;
        if (left) {
          const uint32_t w1 = *reinterpret_cast<const uint32_t*>(p + 4);
          v = __funnelshift_r(w0, w1, shl);
          if (j == 2) v &= 0x00ffffffu;
        }
        sp[SH_PITCH * dy + j + (left ? 0 : 1)] = v;
      }
    }
    __syncwarp(hmask);
    // lane = shift k (11 of the 16 lanes): SAD over the 11 rows with byte-SIMD |a - b| accumulation (VABSDIFF4.ACC)
    const int k = min(hl, 10);
    const int sR = (cxr - L - w) - xr0 + k;                 // byte offset of shift k inside the right words (0..13)
    const uint32_t shr = 8u * (uint32_t)(sR & 3);
    const uint32_t* bp = sp + 4 + (sR >> 2);
    uint32_t sad = 0;
#pragma unroll
    for (int dy = 0; dy < 11; ++dy) {
      const uint4 a = *reinterpret_cast<const uint4*>(sp + SH_PITCH * dy);
      const uint32_t* b = bp + SH_PITCH * dy;
      const uint32_t b0w = b[0], b1w = b[1], b2w = b[2], b3w = b[3];
      sad = __vsadu4(a.x, __funnelshift_r(b0w, b1w, shr)) + sad;
      sad = __vsadu4(a.y, __funnelshift_r(b1w, b2w, shr)) + sad;
      sad = __vsadu4(a.z, __funnelshift_r(b2w, b3w, shr) & 0x00ffffffu) + sad;
    }
    // best shift: ascending scan with strict < (:1001-1004) == lexicographic minimum of (sad, k); sad <= 121 * 255
    const uint32_t key = __reduce_min_sync(hmask, hl < 11 ? ((sad << 4) | (uint32_t)hl) : 0xffffffffu);
    const int bestSad = (int)(key >> 4), bestK = (int)(key & 15u), bestInc = bestK - L;
    // lanes 0..10 of the half hold the SAD of shift k; the parabola needs the neighbours of the best one
    const float d1 = (float)__shfl_sync(hmask, sad, hbase + max(bestK - 1, 0));
    const float d2 = (float)bestSad;
    const float d3 = (float)__shfl_sync(hmask, sad, hbase + min(bestK + 1, 10));
    if (bestInc == -L || bestInc == L) break;             // :1005
    // deltaR = (dist1 - dist3) / (2.0f * (dist1 + dist3 - 2.0f * dist2))  (:1012-1013)
    const float deltaR = __fdiv_rn(__fsub_rn(d1, d3), __fmul_rn(2.0f, __fsub_rn(__fadd_rn(d1, d3), __fmul_rn(2.0f, d2))));
    if (deltaR < -1 || deltaR > 1) break;                 // NaN falls through, rejected by the range test below
    float bestuR = __fmul_rn(gL.scale[levelL], __fadd_rn(__fadd_rn(scaleduR0, (float)bestInc), deltaR));
    float disparity = __fsub_rn(uL, bestuR);
    if (disparity >= 0.f && disparity < maxD) {           // minD = 0
      if (disparity <= 0) {
        disparity = 0.01f;                                // (float)0.01
        bestuR = (float)((double)uL - 0.01);
      }
      r_depth = __fdiv_rn(mbf, disparity);
      r_u = bestuR;
      r_sad = bestSad;
    }
  } while (0);
  if (hl == 0) { uright[oL] = r_u; depth[oL] = r_depth; sad_out[oL] = r_sad; best_idx[oL] = r_idx; best_dist[oL] = r_dist; }
}

// host_uright / host_depth (small batches, page-locked result buffers of the caller, host_cap entries per frame): the final values
// also go straight to the host from here instead of two device-to-host copies
__global__ void __launch_bounds__(256) k_stereo_gate(int kcap, const int* __restrict__ nL_arr, const int* __restrict__ sad,
                                                     float* __restrict__ uright, float* __restrict__ depth, float* __restrict__ host_uright,
                                                     float* __restrict__ host_depth, int host_cap) {
  __shared__ int hist[256];
  __shared__ int s_total, s_bin, s_rank, s_median;
  const int frame = blockIdx.x, tid = threadIdx.x;
  const int nL = nL_arr[frame];
  const int* sd = sad + (size_t)frame * kcap;
  hist[tid] = 0;
  if (tid == 0) s_total = 0;
  __syncthreads();
  int local = 0;
  for (int i = tid; i < nL; i += 256) {
    const int v = sd[i];
    if (v >= 0) { atomicAdd(&hist[min(v >> 7, 255)], 1); ++local; }
  }
  if (local) atomicAdd(&s_total, local);
  __syncthreads();
  const int n = s_total;
  if (n != 0) {   // n == 0: the reference reads vDistIdx[0] of an empty vector here; nothing to gate
  if (tid == 0) {
    int k = n / 2, cum = 0, b = 0;  // vDistIdx[size / 2] after the ascending sort (:1036)
    for (; b < 256; ++b) { if (cum + hist[b] > k) break; cum += hist[b]; }
    s_bin = b; s_rank = k - cum;
  }
  __syncthreads();
  const int bin = s_bin;
  hist[tid] = 0;
  __syncthreads();
  for (int i = tid; i < nL; i += 256) {
    const int v = sd[i];
    if (v >= 0 && min(v >> 7, 255) == bin) atomicAdd(&hist[v & 127], 1);
  }
  __syncthreads();
  if (tid == 0) {
    int k = s_rank, cum = 0, b = 0;
    for (; b < 128; ++b) { if (cum + hist[b] > k) break; cum += hist[b]; }
    s_median = (bin << 7) | b;
  }
  __syncthreads();
  const float median = (float)s_median;
  const float thDist = __fmul_rn(__fmul_rn(1.5f, 1.4f), median);  // 1.5f * 1.4f * median (:1037)
  for (int i = tid; i < nL; i += 256) {
    const int v = sd[i];
    if (v >= 0 && !((float)v < thDist)) {
      uright[(size_t)frame * kcap + i] = -1.f;
      depth[(size_t)frame * kcap + i] = -1.f;
    }
  }
  }
  if (host_uright) {
    __syncthreads();   // (the gate above wrote with the same thread -> index mapping; the barrier keeps the code obviously ordered)
    for (int i = tid; i < min(kcap, host_cap); i += 256) {
      host_uright[(size_t)frame * host_cap + i] = uright[(size_t)frame * kcap + i];
      host_depth[(size_t)frame * host_cap + i] = depth[(size_t)frame * kcap + i];
    }
  }
}

static int stereo_buffers(orb_handle* h, int batch) {
  int st;
  const size_t n = (size_t)std::max(batch, h->max_batch) * h->g.kcap;
  if ((st = orb_ensure(h, h->d_uright, n * sizeof(float)))) return st;
  if ((st = orb_ensure(h, h->d_depth, n * sizeof(float)))) return st;
  if ((st = orb_ensure(h, h->d_sad, n * sizeof(int)))) return st;
  if ((st = orb_ensure(h, h->d_best_idx, n * sizeof(int)))) return st;
  if ((st = orb_ensure(h, h->d_best_dist, n * sizeof(int)))) return st;
  return ORB_OK;
}

// capacity of the row table of one frame: every right keypoint appears in at most 2 * ceil(r) + 3 rows
static int row_items_cap(const orb_handle* hL, const orb_handle* hR) {
  float smax = 1.f;
  for (int l = 0; l < hL->g.nlevels; ++l) smax = std::max(smax, hL->g.scale[l]);
  const int band = 2 * (int)std::ceil(2.0f * smax) + 3;
  return hR->g.kcap * band;
}

static int stereo_launch(orb_handle* hL, orb_handle* hR, int batch, float mbf, float max_d, float* host_uright = nullptr,
                         float* host_depth = nullptr, int host_cap = 0) {
  int st;
  if (hR->g.kcap > 65535) return orb_set_error(hL, ORB_ERR_CAPACITY, "more than 65535 keypoints per frame");
  if ((st = stereo_buffers(hL, batch))) return st;
  // order hL's stream after everything queued on hR's stream
  if ((st = orb_peer_read_begin(hL, hR))) return st;
  if (hL->stage_timing) cudaEventRecord(hL->ev_stage[7], hL->stream);
  const OrbGeom& gL = hL->g;
  const int H0 = gL.h[0];
  const int items_cap = row_items_cap(hL, hR);
  const int bcap = std::max(batch, hL->max_batch);
  if ((st = orb_ensure(hL, hL->d_rband, (size_t)bcap * (H0 + 1) * sizeof(int)))) return st;
  if ((st = orb_ensure(hL, hL->d_row_items, (size_t)bcap * items_cap * sizeof(uint2)))) return st;
  k_stereo_rows<<<batch, SR_THREADS, (size_t)(H0 + 1) * sizeof(int), hL->stream>>>(gL, hR->g.kcap, items_cap, hR->d_kps.as<orb_keypoint>(),
                                                                            hR->d_n.as<int>(), hL->d_rband.as<int>(),
                                                                            hL->d_row_items.as<uint2>());
  hL->launches++;
  static const bool warp_per_keypoint = [] { const char* e = getenv("ORB_B200_STEREO"); return e && !strcmp(e, "warp"); }();   // measurement switch
  if (warp_per_keypoint)
    k_stereo_match<<<dim3((gL.kcap + ST_WARPS - 1) / ST_WARPS, batch), ST_WARPS * 32, 0, hL->stream>>>(
        gL, hR->g, hL->d_pyr.as<uint8_t>(), hR->d_pyr.as<uint8_t>(), hL->d_kps.as<orb_keypoint>(), hL->d_desc.as<uint8_t>(),
        hL->d_n.as<int>(), hR->d_kps.as<orb_keypoint>(), hR->d_desc.as<uint8_t>(), hR->d_n.as<int>(), mbf, max_d,
        hL->d_rband.as<int>(), hL->d_row_items.as<uint2>(), items_cap, hL->d_uright.as<float>(), hL->d_depth.as<float>(), hL->d_sad.as<int>(), hL->d_best_idx.as<int>(),
        hL->d_best_dist.as<int>());
  else
    k_stereo_match_h<<<dim3((gL.kcap + SH_SUBS - 1) / SH_SUBS, batch), ST_WARPS * 32, 0, hL->stream>>>(
        gL, hR->g, hL->d_pyr.as<uint8_t>(), hR->d_pyr.as<uint8_t>(), hL->d_kps.as<orb_keypoint>(), hL->d_desc.as<uint8_t>(),
        hL->d_n.as<int>(), hR->d_kps.as<orb_keypoint>(), hR->d_desc.as<uint8_t>(), mbf, max_d,
        hL->d_rband.as<int>(), hL->d_row_items.as<uint2>(), items_cap, hL->d_uright.as<float>(), hL->d_depth.as<float>(), hL->d_sad.as<int>(), hL->d_best_idx.as<int>(),
        hL->d_best_dist.as<int>());
  if (hL->stage_timing) cudaEventRecord(hL->ev_stage[8], hL->stream);
  // last kernel that reads hR's pyramid / keypoints / descriptors: hR's next extraction waits for it
  if ((st = orb_peer_read_end(hL, hR))) return st;
  k_stereo_gate<<<batch, 256, 0, hL->stream>>>(gL.kcap, hL->d_n.as<int>(), hL->d_sad.as<int>(), hL->d_uright.as<float>(),
                                               hL->d_depth.as<float>(), host_uright, host_depth, host_cap);
  if (hL->stage_timing) cudaEventRecord(hL->ev_stage[9], hL->stream);
  hL->launches += 2;
  ORB_CUDA_CHECK(hL, cudaGetLastError());
  hL->have_stereo = true;
  return ORB_OK;
}

extern "C" {

int orb_stereo_match_batch(orb_handle* hL, orb_handle* hR, float mbf, float max_d, float* uright_out, float* depth_out,
                           int cap, int flags) {
  if (!hL || !hR) return ORB_ERR_INVALID_ARG;
  if (!hL->have_batch || !hR->have_batch) return orb_set_error(hL, ORB_ERR_STATE, "stereo match needs an extraction on both handles");
  if (hL->frames_loaded || hR->frames_loaded) return orb_set_error(hL, ORB_ERR_STATE, "stereo match needs the pyramids of an extraction (orb_load_frames keeps none)");
  if (hL->device != hR->device) return orb_set_error(hL, ORB_ERR_INVALID_ARG, "both handles must live on the same device");
  if (hL->cur_batch != hR->cur_batch || hL->g.nlevels != hR->g.nlevels)
    return orb_set_error(hL, ORB_ERR_INVALID_ARG, "left/right batches differ");
  int st;
  if ((st = orb_use_device(hL))) return st;
  const int batch = hL->cur_batch;
  // a few frames with page-locked result buffers: the gate kernel writes mvuRight / mvDepth into them itself
  const bool zero_copy = batch <= ORB_SMALL_BATCH && !(flags & (ORB_DST_DEVICE | ORB_NO_OUTPUT)) && uright_out && depth_out && cap > 0 &&
                         orb_host_buffer_is_device_writable(uright_out) && orb_host_buffer_is_device_writable(depth_out);
  if ((st = stereo_launch(hL, hR, batch, mbf, max_d, zero_copy ? uright_out : nullptr, zero_copy ? depth_out : nullptr, cap))) return st;
  if (!(flags & ORB_NO_OUTPUT) && !zero_copy) {
    const int kcap = hL->g.kcap;
    const int rows = std::min(cap, kcap);
    if (cap == kcap) {
      if (uright_out) ORB_CUDA_CHECK(hL, cudaMemcpyAsync(uright_out, hL->d_uright.p, (size_t)batch * kcap * 4, cudaMemcpyDefault, hL->stream));
      if (depth_out) ORB_CUDA_CHECK(hL, cudaMemcpyAsync(depth_out, hL->d_depth.p, (size_t)batch * kcap * 4, cudaMemcpyDefault, hL->stream));
    } else {
      if (uright_out)
        ORB_CUDA_CHECK(hL, cudaMemcpy2DAsync(uright_out, (size_t)cap * 4, hL->d_uright.p, (size_t)kcap * 4, (size_t)rows * 4, batch,
                                             cudaMemcpyDefault, hL->stream));
      if (depth_out)
        ORB_CUDA_CHECK(hL, cudaMemcpy2DAsync(depth_out, (size_t)cap * 4, hL->d_depth.p, (size_t)kcap * 4, (size_t)rows * 4, batch,
                                             cudaMemcpyDefault, hL->stream));
    }
  }
  if (flags & ORB_ASYNC) return ORB_OK;
  ORB_CUDA_CHECK(hL, cudaStreamSynchronize(hL->stream));
  if (hL->stage_timing) {
    cudaEventElapsedTime(&hL->stage_ms[6], hL->ev_stage[7], hL->ev_stage[8]);
    cudaEventElapsedTime(&hL->stage_ms[7], hL->ev_stage[8], hL->ev_stage[9]);
  }
  return ORB_OK;
}

int orb_stereo_match(orb_handle* hL, orb_handle* hR, const orb_keypoint* kpsL, const uint8_t* descL, int nL,
                     const orb_keypoint* kpsR, const uint8_t* descR, int nR, float mbf, float max_d, float* uright_out,
                     float* depth_out) {
  if (!hL || !hR || nL < 0 || nR < 0 || (nL && (!kpsL || !descL)) || (nR && (!kpsR || !descR))) return ORB_ERR_INVALID_ARG;
  if (!hL->have_batch || !hR->have_batch) return orb_set_error(hL, ORB_ERR_STATE, "stereo match needs an extraction on both handles");
  if (hL->frames_loaded || hR->frames_loaded) return orb_set_error(hL, ORB_ERR_STATE, "stereo match needs the pyramids of an extraction (orb_load_frames keeps none)");
  if (hL->device != hR->device) return orb_set_error(hL, ORB_ERR_INVALID_ARG, "both handles must live on the same device");
  if (nL > hL->g.kcap || nR > hR->g.kcap || nR > 65535) return orb_set_error(hL, ORB_ERR_CAPACITY, "more keypoints than the handle holds");
  int st;
  if ((st = orb_use_device(hL))) return st;
  if ((st = orb_sync(hL)) || (st = orb_sync(hR))) return st;
  // the caller's keypoints/descriptors replace frame 0 of the device-resident results
  ORB_CUDA_CHECK(hL, cudaMemcpyAsync(hL->d_kps.p, kpsL, (size_t)nL * sizeof(orb_keypoint), cudaMemcpyHostToDevice, hL->stream));
  ORB_CUDA_CHECK(hL, cudaMemcpyAsync(hL->d_desc.p, descL, (size_t)nL * 32, cudaMemcpyHostToDevice, hL->stream));
  ORB_CUDA_CHECK(hL, cudaMemcpyAsync(hL->d_n.p, &nL, sizeof(int), cudaMemcpyHostToDevice, hL->stream));
  ORB_CUDA_CHECK(hL, cudaMemcpyAsync(hR->d_kps.p, kpsR, (size_t)nR * sizeof(orb_keypoint), cudaMemcpyHostToDevice, hL->stream));
  ORB_CUDA_CHECK(hL, cudaMemcpyAsync(hR->d_desc.p, descR, (size_t)nR * 32, cudaMemcpyHostToDevice, hL->stream));
  ORB_CUDA_CHECK(hL, cudaMemcpyAsync(hR->d_n.p, &nR, sizeof(int), cudaMemcpyHostToDevice, hL->stream));
  ORB_CUDA_CHECK(hL, cudaStreamSynchronize(hL->stream));
  if ((st = stereo_launch(hL, hR, 1, mbf, max_d))) return st;
  if (uright_out && nL) ORB_CUDA_CHECK(hL, cudaMemcpyAsync(uright_out, hL->d_uright.p, (size_t)nL * 4, cudaMemcpyDeviceToHost, hL->stream));
  if (depth_out && nL) ORB_CUDA_CHECK(hL, cudaMemcpyAsync(depth_out, hL->d_depth.p, (size_t)nL * 4, cudaMemcpyDeviceToHost, hL->stream));
  ORB_CUDA_CHECK(hL, cudaStreamSynchronize(hL->stream));
  return ORB_OK;
}

int orb_debug_get_stereo_best(orb_handle* hL, int frame, int32_t* best_idx, int32_t* best_dist, int cap) {
  if (!hL || frame < 0) return ORB_ERR_INVALID_ARG;
  if (!hL->have_stereo) return orb_set_error(hL, ORB_ERR_STATE, "no stereo match has run");
  int st;
  if ((st = orb_use_device(hL))) return st;
  const int kcap = hL->g.kcap;
  const int rows = std::min(cap, kcap);
  ORB_CUDA_CHECK(hL, cudaStreamSynchronize(hL->stream));
  if (best_idx) ORB_CUDA_CHECK(hL, cudaMemcpy(best_idx, hL->d_best_idx.as<int>() + (size_t)frame * kcap, (size_t)rows * 4, cudaMemcpyDeviceToHost));
  if (best_dist) ORB_CUDA_CHECK(hL, cudaMemcpy(best_dist, hL->d_best_dist.as<int>() + (size_t)frame * kcap, (size_t)rows * 4, cudaMemcpyDeviceToHost));
  return ORB_OK;
}

}  // extern "C"
