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
// lanes busy on the expensive steps: the ROI is staged with aligned 32-bit loads; a cheap 4-compass-point
// test runs on every pixel (division-free indexing) and appends survivors to a shared-memory list; the
// 16-bit arc masks and the exact score then run on dense lists; NMS visits corners only and sets bits in
// per-row masks; the ordered output is produced from mask words with one block scan.
#pragma once

#define FAST_TPB 88                 // tile pitch in bytes (22 words): ROI <= 80 plus up to 3 bytes of alignment offset
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

// dynamic shared memory: [list1 u16 x cap][list2 u16 x cap], cap = largest cell interior of the geometry
__global__ void __launch_bounds__(FAST_THREADS) k_fast_cells(OrbGeom g, const uint8_t* __restrict__ pyr,
                                                             int* __restrict__ cell_count, uint32_t* __restrict__ cell_keys,
                                                             int cells_per_frame, int list_cap, int* __restrict__ status) {
  __shared__ __align__(16) uint32_t tile_w[ORB_ROI_MAX * (FAST_TPB / 4)];
  __shared__ __align__(16) uint8_t sc[(ORB_ROI_MAX + 2) * FAST_SP];  // interior scores with a 1-px zero ring
  __shared__ uint32_t m_ini[ORB_ROI_MAX * FAST_WPR], m_min[ORB_ROI_MAX * FAST_WPR];
  __shared__ int s_cnt1, s_cnt2, s_any_ini;
  __shared__ int s_wsum[FAST_THREADS / 32];
  extern __shared__ __align__(16) uint16_t s_lists[];
  uint16_t* list1 = s_lists;
  uint16_t* list2 = s_lists + list_cap;

  const int cell = blockIdx.x, frame = blockIdx.y;
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const uint32_t lt = (1u << lane) - 1u;
  int l = 0;
  while (cell >= g.cell_start[l + 1]) ++l;
  const int ci = cell - g.cell_start[l];
  const int ci_i = ci / g.ncols[l], ci_j = ci - ci_i * g.ncols[l];
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
  // ---- stage the ROI: aligned 32-bit loads (level rows are 16-byte aligned), lanes = words of a row
  const int x0a = iniX & ~3, xoff = iniX - x0a;
  const int wpr = (xoff + rw + 3) >> 2;  // <= 21
  {
    const uint8_t* __restrict__ src = lvl_ptr(g, pyr, frame, l) + (size_t)iniY * P + x0a;
    for (int y = wid; y < rh; y += FAST_THREADS / 32)
      if (lane < wpr) tile_w[y * (FAST_TPB / 4) + lane] = *reinterpret_cast<const uint32_t*>(src + (size_t)y * P + 4 * lane);
    uint32_t* scw = reinterpret_cast<uint32_t*>(sc);
    for (int i = tid; i < (ih + 2) * (FAST_SP / 4); i += FAST_THREADS) scw[i] = 0u;
    for (int i = tid; i < ih * FAST_WPR; i += FAST_THREADS) { m_ini[i] = 0u; m_min[i] = 0u; }
    if (tid == 0) { s_cnt1 = 0; s_cnt2 = 0; s_any_ini = 0; }
  }
  __syncthreads();
  const uint8_t* tile = reinterpret_cast<const uint8_t*>(tile_w) + xoff + 3 * FAST_TPB + 3;  // interior origin
  const int th = g.min_th;
  const int npix = iw * ih;

  // ---- pass A: every interior pixel, compass points 0 (0,+3), 4 (+3,0), 8 (0,-3), 12 (-3,0).
  //      A 9-arc of the 16-ring contains at least two of them, so fewer than two brighter AND fewer than
  //      two darker compass points rules the pixel out.
  {
    const int sy = FAST_THREADS / iw, sx = FAST_THREADS - sy * iw;
    int y = tid / iw, x = tid - y * iw;
    for (int p = tid; p < ((npix + 31) & ~31); p += FAST_THREADS) {
      bool pass = false;
      if (p < npix) {
        const uint8_t* c = tile + y * FAST_TPB + x;
        const int v = c[0];
        const int hi = v + th, lo = v - th;
        const int r0 = c[3 * FAST_TPB], r4 = c[3], r8 = c[-3 * FAST_TPB], r12 = c[-3];
        const int nb = (r0 > hi) + (r4 > hi) + (r8 > hi) + (r12 > hi);
        const int nd = (r0 < lo) + (r4 < lo) + (r8 < lo) + (r12 < lo);
        pass = (nb >= 2) | (nd >= 2);
      }
      const uint32_t b = __ballot_sync(0xffffffffu, pass);
      if (b) {
        int base = 0;
        if (lane == 0) base = atomicAdd(&s_cnt1, __popc(b));
        base = __shfl_sync(0xffffffffu, base, 0);
        if (pass) list1[base + __popc(b & lt)] = (uint16_t)((y << 7) | x);
      }
      x += sx; y += sy;
      if (x >= iw) { x -= iw; ++y; }
    }
  }
  __syncthreads();

  // ---- pass B: 16-ring masks on the survivors; corners at minThFAST go to list2 (bit 15 = brighter arc)
  {
    const int n1 = s_cnt1;
    for (int i = tid; i < ((n1 + 31) & ~31); i += FAST_THREADS) {
      bool pass = false, bright = false;
      int code = 0;
      if (i < n1) {
        code = list1[i];
        const uint8_t* c = tile + (code >> 7) * FAST_TPB + (code & 127);
        const int v = c[0];
        const int hi = v + th, lo = v - th;
        uint32_t mb = 0, md = 0;
#define FAST_RING(k, off) { const int r = c[off]; mb |= (uint32_t)(r > hi) << k; md |= (uint32_t)(r < lo) << k; }
        FAST_RING(0, 3 * FAST_TPB)      FAST_RING(1, 3 * FAST_TPB + 1)   FAST_RING(2, 2 * FAST_TPB + 2)   FAST_RING(3, FAST_TPB + 3)
        FAST_RING(4, 3)                 FAST_RING(5, -FAST_TPB + 3)      FAST_RING(6, -2 * FAST_TPB + 2)  FAST_RING(7, -3 * FAST_TPB + 1)
        FAST_RING(8, -3 * FAST_TPB)     FAST_RING(9, -3 * FAST_TPB - 1)  FAST_RING(10, -2 * FAST_TPB - 2) FAST_RING(11, -FAST_TPB - 3)
        FAST_RING(12, -3)               FAST_RING(13, FAST_TPB - 3)      FAST_RING(14, 2 * FAST_TPB - 2)  FAST_RING(15, 3 * FAST_TPB - 1)
#undef FAST_RING
        bright = has_arc9(mb);
        pass = bright || has_arc9(md);
      }
      const uint32_t b = __ballot_sync(0xffffffffu, pass);
      if (b) {
        int base = 0;
        if (lane == 0) base = atomicAdd(&s_cnt2, __popc(b));
        base = __shfl_sync(0xffffffffu, base, 0);
        if (pass) list2[base + __popc(b & lt)] = (uint16_t)(code | (bright ? 0x8000 : 0));
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
