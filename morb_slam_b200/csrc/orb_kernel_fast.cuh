// FAST-9/16 per 35-px cell (reference src/ORBextractor.cc:744-820: cv::FAST at iniThFAST, fallback to
// minThFAST when the cell is empty, on the cell ROI [iniX, maxX) x [iniY, maxY)).
//
// One CTA per cell. Formulation (equivalence with the two cv::FAST calls: SURVEY.md Appendix A.3):
//   score S(p) = OpenCV cornerScore<16> for pixels that are corners at minThFAST, else 0;
//   local maximum  <=> S(p) > S(q) for the 8 neighbours q, pixels outside the cell interior count as 0;
//   cell threshold = iniThFAST if any local maximum reaches it, else minThFAST;
//   output = local maxima with S >= threshold, row-major (the order is part of the contract).
//
// The kernel is instruction-bound, not HBM-bound (profiles/README_r1.md), so it is organised to keep
// lanes busy on the expensive steps:
//   load   the ROI is staged with aligned 32-bit loads and re-aligned (funnel shift with the neighbour lane's
//          word) so that interior column 0 sits on a word boundary of the shared-memory tile;
//   pass A every pixel, 4 per thread in byte-SIMD: compass points 0/4/8/12 against v +- t; a 9-arc of the
//          16-ring contains at least two of them, so "fewer than two brighter and fewer than two darker"
//          rules a pixel out; survivors (with the polarities still possible) go to a shared-memory list;
//   pass B 16-bit arc mask of the possible polarity on the dense list -> corners at minThFAST;
//   pass C exact score on the dense corner list;
//   pass D NMS, corners only, sets bits in per-row masks;
//   pass E ordered output from the mask words with one block scan.
#pragma once

#define FAST_TW 23                  // tile pitch in words: 1 + interior/ring columns (<= 80) / 4 + spare
#define FAST_TPB (FAST_TW * 4)      // tile pitch in bytes
#define FAST_SP ORB_ROI_MAX         // score map pitch
#define FAST_WPR 3                  // mask words per interior row (interior width <= 74)
#define FAST_THREADS 128

static __device__ __forceinline__ bool has_arc9(uint32_t m16) {
  const uint32_t d = m16 | (m16 << 16);
  uint32_t m = d & (d >> 1);  // 2 contiguous
  m &= m >> 2;                // 4
  m &= m >> 4;                // 8
  m &= d >> 8;                // 9
  return (m & 0xffffu) != 0;
}

// per-byte unsigned a > b, result in bit 7 of every byte: carry out of a + ~b
static __device__ __forceinline__ uint32_t swar_gt(uint32_t a, uint32_t b) {
  const uint32_t nb = ~b;
  const uint32_t t = (a & 0x7f7f7f7fu) + (nb & 0x7f7f7f7fu);
  return (a & nb) | ((a | nb) & t);  // majority(a7, ~b7, carry into bit 7)
}
// Four compass flags in ring order (0, 4, 8, 12): a 9-arc holds 2 or 3 compass points, and two of them are
// always ring-adjacent, so "some adjacent pair set" is necessary for an arc (stricter than "any two").
static __device__ __forceinline__ uint32_t swar_adjacent_pair(uint32_t p0, uint32_t p4, uint32_t p8, uint32_t p12) {
  return ((p0 | p8) & (p4 | p12));  // (p0&p4)|(p4&p8)|(p8&p12)|(p12&p0)
}

// One launch per pyramid level: grid = (cell columns, cell rows, frames), so the cell geometry comes straight
// from blockIdx without divisions, table look-ups or dependent loads in the prologue.
// Dynamic shared memory (sized by the host for the level's largest cell, see launch_pipeline in orb_extract.cu):
//   tile words [rh_max][FAST_TW] | score bytes [(ih_max + 2)][FAST_SP] | m_ini, m_min words [ih_max][FAST_WPR] |
//   list1, list2 u16 [list_cap]
__global__ void __launch_bounds__(FAST_THREADS, 12) k_fast_cells(OrbGeom g, int l, const uint8_t* __restrict__ pyr,
                                                                 int* __restrict__ cell_count, uint32_t* __restrict__ cell_keys,
                                                                 int cells_per_frame, int rh_max, int list_cap,
                                                                 int* __restrict__ status) {
  extern __shared__ __align__(16) uint32_t s_dyn[];
  __shared__ int s_cnt1, s_cnt2, s_any_ini;
  __shared__ int s_wsum[FAST_THREADS / 32];
  const int ih_max = rh_max - 6;
  uint32_t* tile_w = s_dyn;
  uint8_t* sc = reinterpret_cast<uint8_t*>(tile_w + rh_max * FAST_TW);   // interior scores with a 1-px zero ring
  uint32_t* m_ini = reinterpret_cast<uint32_t*>(sc + (ih_max + 2) * FAST_SP);
  uint32_t* m_min = m_ini + ih_max * FAST_WPR;
  uint16_t* list1 = reinterpret_cast<uint16_t*>(m_min + ih_max * FAST_WPR);
  uint16_t* list2 = list1 + list_cap;

  const int ci_j = blockIdx.x, ci_i = blockIdx.y, frame = blockIdx.z;
  const int cell = g.cell_start[l] + ci_i * g.ncols[l] + ci_j;
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const uint32_t lt = (1u << lane) - 1u;
  const int W = g.w[l], H = g.h[l], P = g.pitch[l];
  const int maxBX = W - ORB_EDGE + 3, maxBY = H - ORB_EDGE + 3;
  const int iniY = ORB_BORDER + ci_i * g.hcell[l];
  const int iniX = ORB_BORDER + ci_j * g.wcell[l];
  int* out_count = cell_count + (size_t)frame * cells_per_frame + cell;
  uint32_t* out_keys = cell_keys + ((size_t)frame * cells_per_frame + cell) * ORB_CELL_CAP;
  const int maxY = min(iniY + g.hcell[l] + 6, maxBY), maxX = min(iniX + g.wcell[l] + 6, maxBX);
  const int rw = maxX - iniX, rh = maxY - iniY;
  const int iw = rw - 6, ih = rh - 6;  // interior: the pixels FAST actually tests
  if (iniY >= maxBY - 3 || iniX >= maxBX - 6 || iw <= 0 || ih <= 0) {  // :767, :773
    if (tid == 0) *out_count = 0;
    return;
  }
  // ---- stage the ROI. Tile column d holds image column iniX - 1 + d, so interior column 0 (ROI column 3) is
  //      tile column 4. Tile word j = image bytes [xs + 4j, xs + 4j + 3], xs = iniX - 1.
  //      Step 1: all aligned source words of the ROI go to a raw staging area with 4-byte cp.async copies
  //      (every load of the CTA in flight at once: one global latency instead of one per row); the staging area
  //      aliases the candidate lists, which are not live yet. Step 2: re-align by funnel-shifting neighbours.
  {
    const int xs = iniX - 1;
    const int xa = xs & ~3, sh = (xs - xa) * 8;
    const int nw = (rw + 1 + 3) >> 2;  // tile words per row that carry ROI data (<= 21)
    const int rpw = nw + 1;            // raw words per row
    uint32_t* raw = reinterpret_cast<uint32_t*>(list1);
    const uint8_t* __restrict__ src = lvl_ptr(g, pyr, frame, l) + (size_t)iniY * P + xa;
    for (int y = wid; y < rh; y += FAST_THREADS / 32)
      if (lane < rpw) {
        const unsigned sa = (unsigned)__cvta_generic_to_shared(&raw[y * rpw + lane]);
        asm volatile("cp.async.ca.shared.global [%0], [%1], 4;\n" ::"r"(sa), "l"(src + (size_t)y * P + 4 * lane));  // inside the padded row
      }
    asm volatile("cp.async.commit_group;\n" ::);
    uint32_t* scw = reinterpret_cast<uint32_t*>(sc);
    const int zw = ((iw + 2 + 3) >> 2);  // words per score row that can ever be read
    for (int y = wid; y < ih + 2; y += FAST_THREADS / 32)
      if (lane < zw) scw[y * (FAST_SP / 4) + lane] = 0u;
    for (int i = tid; i < ih * FAST_WPR; i += FAST_THREADS) { m_ini[i] = 0u; m_min[i] = 0u; }
    if (tid == 0) { s_cnt1 = 0; s_cnt2 = 0; s_any_ini = 0; }
    asm volatile("cp.async.wait_group 0;\n" ::);
    __syncthreads();
    for (int y = wid; y < rh; y += FAST_THREADS / 32)
      if (lane < nw) tile_w[y * FAST_TW + lane] = __funnelshift_r(raw[y * rpw + lane], raw[y * rpw + lane + 1], sh);
  }
  __syncthreads();
  const int th = g.min_th;

  // ---- pass A (byte-SIMD): item = (interior row y, word wx) = interior columns 4wx .. 4wx+3
  {
    const int wpi = (iw + 3) >> 2;            // words per interior row
    const int nitems = ih * wpi;
    const uint32_t th4 = (uint32_t)th * 0x01010101u;
    const int sy = FAST_THREADS / wpi, sx = FAST_THREADS - sy * wpi;
    int y = tid / wpi, wx = tid - y * wpi;
    for (int it = tid; it < ((nitems + 31) & ~31); it += FAST_THREADS) {
      uint32_t pb = 0, pd = 0;  // per-byte flags (bit 7): brighter / darker arc still possible
      if (it < nitems) {
        const uint32_t* c = &tile_w[(y + 3) * FAST_TW + 1 + wx];
        const uint32_t C = c[0];
        const uint32_t T = c[3 * FAST_TW], B = c[-3 * FAST_TW];     // ring points 0 (0,+3) and 8 (0,-3)
        const uint32_t R = __funnelshift_r(C, c[1], 24);            // ring point 4 (+3,0)
        const uint32_t L = __funnelshift_r(c[-1], C, 8);            // ring point 12 (-3,0)
        // hi = min(C + t, 255), lo = max(C - t, 0) per byte (saturation keeps "r > hi" / "r < lo" exact)
        const uint32_t hi = __vaddus4(C, th4), lo = __vsubus4(C, th4);
        pb = swar_adjacent_pair(swar_gt(T, hi), swar_gt(R, hi), swar_gt(B, hi), swar_gt(L, hi));
        pd = swar_adjacent_pair(swar_gt(lo, T), swar_gt(lo, R), swar_gt(lo, B), swar_gt(lo, L));
        // drop the columns past the interior in the last word of a row
        const int valid = min(iw - 4 * wx, 4);
        const uint32_t vm = valid >= 4 ? 0x80808080u : ((1u << (8 * valid)) - 1u) & 0x80808080u;
        pb &= vm; pd &= vm;
      }
      const uint32_t any = (pb | pd) & 0x80808080u;
      // warp-aggregated append: one ballot per byte position, one shared atomic per warp
      const uint32_t b0 = __ballot_sync(0xffffffffu, any & 0x00000080u), b1 = __ballot_sync(0xffffffffu, any & 0x00008000u);
      const uint32_t b2 = __ballot_sync(0xffffffffu, any & 0x00800000u), b3 = __ballot_sync(0xffffffffu, any & 0x80000000u);
      const int n0 = __popc(b0), n1 = __popc(b1), n2 = __popc(b2), n3 = __popc(b3);
      if (b0 | b1 | b2 | b3) {
        int base = 0;
        if (lane == 0) base = atomicAdd(&s_cnt1, n0 + n1 + n2 + n3);
        base = __shfl_sync(0xffffffffu, base, 0);
        const uint32_t code0 = (uint32_t)((y << 7) | (4 * wx));
        if (any & 0x00000080u) list1[base + __popc(b0 & lt)] = (uint16_t)(code0 | ((pb >> 7) & 1u) << 14 | ((pd >> 7) & 1u) << 15);
        base += n0;
        if (any & 0x00008000u) list1[base + __popc(b1 & lt)] = (uint16_t)((code0 + 1) | ((pb >> 15) & 1u) << 14 | ((pd >> 15) & 1u) << 15);
        base += n1;
        if (any & 0x00800000u) list1[base + __popc(b2 & lt)] = (uint16_t)((code0 + 2) | ((pb >> 23) & 1u) << 14 | ((pd >> 23) & 1u) << 15);
        base += n2;
        if (any & 0x80000000u) list1[base + __popc(b3 & lt)] = (uint16_t)((code0 + 3) | ((pb >> 31) & 1u) << 14 | ((pd >> 31) & 1u) << 15);
      }
      wx += sx; y += sy;
      if (wx >= wpi) { wx -= wpi; ++y; }
    }
  }
  __syncthreads();
  const uint8_t* tile = reinterpret_cast<const uint8_t*>(tile_w) + 3 * FAST_TPB + 4;  // interior origin

  // ---- pass B: 16-ring arc masks of the polarities still possible; corners at minThFAST go to list2
  //      (bit 15 = the arc is brighter than the centre)
  {
    const int n1 = s_cnt1;
    for (int i = tid; i < ((n1 + 31) & ~31); i += FAST_THREADS) {
      bool pass = false, bright = false;
      int code = 0;
      if (i < n1) {
        code = list1[i];
        const uint8_t* c = tile + ((code >> 7) & 127) * FAST_TPB + (code & 127);
        const int v = c[0];
        int r[16];
        r[0] = c[3 * FAST_TPB];       r[1] = c[3 * FAST_TPB + 1];   r[2] = c[2 * FAST_TPB + 2];    r[3] = c[FAST_TPB + 3];
        r[4] = c[3];                  r[5] = c[-FAST_TPB + 3];      r[6] = c[-2 * FAST_TPB + 2];   r[7] = c[-3 * FAST_TPB + 1];
        r[8] = c[-3 * FAST_TPB];      r[9] = c[-3 * FAST_TPB - 1];  r[10] = c[-2 * FAST_TPB - 2];  r[11] = c[-FAST_TPB - 3];
        r[12] = c[-3];                r[13] = c[FAST_TPB - 3];      r[14] = c[2 * FAST_TPB - 2];   r[15] = c[3 * FAST_TPB - 1];
        if (code & 0x4000) {
          const int hi = v + th;
          uint32_t m = 0;
#pragma unroll
          for (int k = 0; k < 16; ++k) m |= (uint32_t)(r[k] > hi) << k;
          bright = has_arc9(m);
        }
        pass = bright;
        if (!bright && (code & 0x8000)) {
          const int lo = v - th;
          uint32_t m = 0;
#pragma unroll
          for (int k = 0; k < 16; ++k) m |= (uint32_t)(r[k] < lo) << k;
          pass = has_arc9(m);
        }
      }
      const uint32_t b = __ballot_sync(0xffffffffu, pass);
      if (b) {
        int base = 0;
        if (lane == 0) base = atomicAdd(&s_cnt2, __popc(b));
        base = __shfl_sync(0xffffffffu, base, 0);
        if (pass) list2[base + __popc(b & lt)] = (uint16_t)((code & 0x3fff) | (bright ? 0x8000 : 0));
      }
    }
  }
  __syncthreads();

  // ---- pass C: exact score of every corner: max over the 16 arcs of 9 of the minimum |difference|, minus 1
  //      (only one polarity can hold a 9-arc, the other cannot exceed the threshold)
  const int n2 = s_cnt2;
  for (int i = tid; i < n2; i += FAST_THREADS) {
    const int code = list2[i];
    const int y = (code >> 7) & 127, x = code & 127;
    const uint8_t* c = tile + y * FAST_TPB + x;
    const int v = c[0];
    const int sgn = (code & 0x8000) ? 1 : -1;
    int e[16];
#define FAST_E(k, off) e[k] = sgn * ((int)c[off] - v);
    FAST_E(0, 3 * FAST_TPB)      FAST_E(1, 3 * FAST_TPB + 1)   FAST_E(2, 2 * FAST_TPB + 2)   FAST_E(3, FAST_TPB + 3)
    FAST_E(4, 3)                 FAST_E(5, -FAST_TPB + 3)      FAST_E(6, -2 * FAST_TPB + 2)  FAST_E(7, -3 * FAST_TPB + 1)
    FAST_E(8, -3 * FAST_TPB)     FAST_E(9, -3 * FAST_TPB - 1)  FAST_E(10, -2 * FAST_TPB - 2) FAST_E(11, -FAST_TPB - 3)
    FAST_E(12, -3)               FAST_E(13, FAST_TPB - 3)      FAST_E(14, 2 * FAST_TPB - 2)  FAST_E(15, 3 * FAST_TPB - 1)
#undef FAST_E
    int m2[16], m4[16], m8[16];
#pragma unroll
    for (int k = 0; k < 16; ++k) m2[k] = min(e[k], e[(k + 1) & 15]);
#pragma unroll
    for (int k = 0; k < 16; ++k) m4[k] = min(m2[k], m2[(k + 2) & 15]);
#pragma unroll
    for (int k = 0; k < 16; ++k) m8[k] = min(m4[k], m4[(k + 4) & 15]);
    int best = 0;
#pragma unroll
    for (int k = 0; k < 16; ++k) best = max(best, min(m8[k], e[(k + 8) & 15]));
    sc[(y + 1) * FAST_SP + (x + 1)] = (uint8_t)(best - 1);
  }
  __syncthreads();

  // ---- pass D: 3x3 strict non-max suppression, corners only; survivors set a bit in their row mask
  for (int i = tid; i < n2; i += FAST_THREADS) {
    const int code = list2[i];
    const int y = (code >> 7) & 127, x = code & 127;
    const uint8_t* s = &sc[(y + 1) * FAST_SP + (x + 1)];
    const int v = s[0];
    const bool lm = v > s[-1] && v > s[1] && v > s[-FAST_SP - 1] && v > s[-FAST_SP] && v > s[-FAST_SP + 1] &&
                    v > s[FAST_SP - 1] && v > s[FAST_SP] && v > s[FAST_SP + 1];
    if (lm) {
      atomicOr(&m_min[y * FAST_WPR + (x >> 5)], 1u << (x & 31));
      if (v >= g.ini_th) { atomicOr(&m_ini[y * FAST_WPR + (x >> 5)], 1u << (x & 31)); s_any_ini = 1; }
    }
  }
  __syncthreads();

  // ---- pass E: ordered output. Mask words are in row-major order; one block scan of their popcounts.
  const uint32_t* mask = s_any_ini ? m_ini : m_min;
  const int nwords = ih * FAST_WPR;
  int carry = 0;
  for (int base = 0; base < nwords; base += FAST_THREADS) {
    const int t = base + tid;
    uint32_t w = (t < nwords) ? mask[t] : 0u;
    const int c = __popc(w);
    int incl = c;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int v = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= o) incl += v;
    }
    if (lane == 31) s_wsum[wid] = incl;
    __syncthreads();
    int off = carry, total = 0;
#pragma unroll
    for (int k = 0; k < FAST_THREADS / 32; ++k) {
      if (k < wid) off += s_wsum[k];
      total += s_wsum[k];
    }
    int pos = off + incl - c;
    if (w) {
      const int y = t / FAST_WPR, xw = (t - y * FAST_WPR) * 32;
      while (w) {
        const int bit = __ffs(w) - 1;
        w &= w - 1;
        const int x = xw + bit;
        if (pos < ORB_CELL_CAP)
          out_keys[pos] = orb_pack(iniX + 3 + x - ORB_BORDER, iniY + 3 + y - ORB_BORDER, sc[(y + 1) * FAST_SP + (x + 1)]);
        ++pos;
      }
    }
    carry += total;
    __syncthreads();
  }
  if (tid == 0) {
    *out_count = min(carry, ORB_CELL_CAP);
    if (carry > ORB_CELL_CAP) atomicOr(status + frame, ORB_ST_CELL_OVERFLOW);
  }
}
