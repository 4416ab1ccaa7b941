// FAST-9/16 per 35-px cell (reference src/ORBextractor.cc:744-820: cv::FAST at iniThFAST, fallback to
// minThFAST when the cell is empty, on the cell ROI [iniX, maxX) x [iniY, maxY)).
//
// One CTA per TILE of nbx x nby cells (one launch per pyramid level, frames in grid.z). The corner score does
// not depend on the cell, only non-maximum suppression and the threshold choice do, so the expensive part runs
// on a large tile and amortises the per-CTA costs; the per-cell part runs on the sparse corner list.
// Formulation (equivalence with the two cv::FAST calls: SURVEY.md Appendix A.3):
//   score S(p) = OpenCV cornerScore<16> for pixels that are corners at minThFAST, else 0;
//   local maximum  <=> S(p) > S(q) for the 8 neighbours q, pixels outside the CELL interior count as 0;
//   cell threshold = iniThFAST if any local maximum of the cell reaches it, else minThFAST;
//   output = local maxima with S >= threshold, row-major inside the cell (the order is part of the contract).
// The interiors of the cells ([iniX + 3, maxX - 3)) tile [19, W - 19) x [19, H - 19) without overlap.
//
// The kernel is instruction-bound, not HBM-bound (profiles/README_r1.md), so everything that touches every
// pixel is byte-SIMD on aligned 32-bit words (4 pixels per instruction):
//   load   ONE TMA tensor copy (cp.async.bulk.tensor.2d) per CTA brings the tile + 3-px ring into shared memory -
//          no per-thread staging instructions at all. The box must start on a 16-byte boundary of the row
//          (measured: other start columns raise "illegal instruction"), so the first interior column sits at
//          byte o = 4..19 of a tile row; the passes work in "xt" columns counted from the word that holds it
//          (xt = x + (o & 3)), which keeps every SIMD word aligned in shared memory;
//   pass A every word: compass points 0/4/8/12 against v +- t; a 9-arc of the 16-ring contains one end of every
//          diameter, so "(p0 | p8) & (p4 | p12)" of one polarity is necessary; words with a surviving byte
//          (35 % on the synthetic frames) go to a shared-memory list (one ballot per warp);
//   pass B listed words only: the full 16-ring test in byte-SIMD for both polarities (16 funnel-shifted ring
//          words, 32 per-byte compares, "9 contiguous" as AND/OR trees on the flag words) -> corner pixels
//          with their polarity go to the corner list;
//   pass C exact score on the dense corner list (3-input min / max);
//   pass D NMS inside the corner's cell, corners only, sets bits in per-cell row masks;
//   pass E one warp per cell: ordered output from the mask words with a warp scan.
#pragma once
#include "orb_tma.cuh"

#define FT_MAXW 124                 // tile interior width limit (7-bit xt in the list codes: xt <= FT_MAXW + 2)
#define FT_MAXH 127                 // hard limit of the list codes (7-bit y); the host picks nby below FT_TILE_H
#define FT_TILE_H 80                // preferred tile interior height
#define FT_TP 160                   // tile pitch in bytes = TMA box width: 19 + FT_MAXW + 3, word reads up to +10
#define FT_TW (FT_TP / 4)
#ifndef FT_THREADS
#define FT_THREADS 256
#endif
#define FT_MINB (1024 / FT_THREADS)
#define FT_MAXCELLS 16              // cells per tile (nbx <= 3 since cells are >= 35 px wide)

// Dynamic shared memory (all carved from one 128-byte aligned block, sized by fast_tile_smem()):
//   tile bytes [bh][FT_TP] (TMA destination) | control words [FT_CTL_WORDS] | score bytes [(ih_max + 2)][sp] |
//   m_ini, m_min words [cells][hcell][wpr] | list2 u16 [list_cap] (corner pixels) | list1 u16 [list1_cap] (words)
#define FT_CTL_WORDS 64   // 0..1 mbarrier, 2 list1 count, 3 list2 count, 4..19 any-ini flag per cell, 32..63 valid-byte masks
static size_t fast_tile_smem(const FastTileGeom& t, int hcell) {
  size_t b = (size_t)t.bh * FT_TP + FT_CTL_WORDS * 4;
  b += (size_t)(t.nby * hcell + 2) * t.sp;
  b = (b + 15) & ~(size_t)15;
  b += 2 * (size_t)t.nbx * t.nby * hcell * t.wpr * 4;
  b += (size_t)t.list_cap * 2 + (size_t)t.list1_cap * 2;
  return (b + 15) & ~(size_t)15;
}

// per-byte r > hi and lo > r (bit 7 of every byte; the other bits are garbage), with the shared sub-expressions
// hoisted: r7 = r & 0x7f.., nh7 = ~hi & 0x7f.., l7 = (lo & 0x7f..) + 0x7f..
static __device__ __forceinline__ void swar_cmp2(uint32_t r, uint32_t hi, uint32_t nh7, uint32_t lo, uint32_t l7,
                                                 uint32_t& brighter, uint32_t& darker) {
  const uint32_t r7 = r & 0x7f7f7f7fu;
  const uint32_t tb = r7 + nh7;   // bit 7: low 7 bits of r exceed those of hi
  const uint32_t td = l7 - r7;    // bit 7: low 7 bits of lo exceed those of r (no borrow between bytes)
  brighter = (r & ~hi) | ((r | ~hi) & tb);
  darker = (lo & ~r) | ((lo | ~r) & td);
}

// shared-memory atomic add without the compiler's generic warp-aggregation wrapper (callers aggregate per warp)
static __device__ __forceinline__ int smem_add(int* p, int v) {
  int old;
  asm volatile("atom.shared.add.u32 %0, [%1], %2;\n" : "=r"(old) : "r"((uint32_t)__cvta_generic_to_shared(p)), "r"(v) : "memory");
  return old;
}

__global__ void __launch_bounds__(FT_THREADS, FT_MINB) k_fast_tiles(const __grid_constant__ CUtensorMap tmap, OrbGeom g, int l,
                                                           FastTileGeom tg, int* __restrict__ cell_count,
                                                           uint32_t* __restrict__ cell_keys, int cells_per_frame,
                                                           int* __restrict__ status) {
  extern __shared__ __align__(128) uint8_t s_dyn[];
  const int wc = g.wcell[l], hc = g.hcell[l];
  const int ihm = tg.nby * hc;
  const int SP = tg.sp, WPR = tg.wpr;
  uint32_t* tile_w = reinterpret_cast<uint32_t*>(s_dyn);
  uint32_t* ctl = tile_w + tg.bh * FT_TW;
  int* s_cnt1 = reinterpret_cast<int*>(ctl + 2);
  int* s_cnt2 = reinterpret_cast<int*>(ctl + 3);
  int* s_any_ini = reinterpret_cast<int*>(ctl + 4);
  uint32_t* vm_tab = ctl + 32;
  uint8_t* sc = reinterpret_cast<uint8_t*>(ctl + FT_CTL_WORDS);  // interior scores with a 1-px zero ring
  uint32_t* m_ini = reinterpret_cast<uint32_t*>(s_dyn + ((tg.bh * FT_TP + FT_CTL_WORDS * 4 + (ihm + 2) * SP + 15) & ~15));
  const int mask_words = tg.nbx * tg.nby * hc * WPR;
  uint32_t* m_min = m_ini + mask_words;
  uint16_t* list2 = reinterpret_cast<uint16_t*>(m_min + mask_words);
  uint16_t* list1 = list2 + tg.list_cap;

  const int frame = blockIdx.z;
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const uint32_t lt = (1u << lane) - 1u;
  const int W = g.w[l], H = g.h[l];
  const int j0 = blockIdx.x * tg.nbx, i0 = blockIdx.y * tg.nby;
  const int ncx = min(tg.nbx, g.ncols[l] - j0), ncy = min(tg.nby, g.nrows[l] - i0);
  const int X0 = ORB_EDGE + j0 * wc, Y0 = ORB_EDGE + i0 * hc;          // first interior pixel of the tile
  const int iw = min(X0 + ncx * wc, W - ORB_EDGE) - X0;                // interior = pixels FAST actually tests
  const int ih = min(Y0 + ncy * hc, H - ORB_EDGE) - Y0;
  const size_t cell_base = (size_t)frame * cells_per_frame + g.cell_start[l];
  if (iw <= 0 || ih <= 0) {  // :767, :773 - cells without a testable pixel
    if (tid < ncx * ncy) {
      const int cy = tid / ncx, cx = tid - cy * ncx;
      cell_count[cell_base + (size_t)(i0 + cy) * g.ncols[l] + j0 + cx] = 0;
    }
    return;
  }

  // ---- stage the tile: box column 0 = image column xa <= X0 - 4, box row 0 = image row Y0 - 3 of this frame (the
  //      level's frames are stacked in the tensor's second dimension). One thread arms the mbarrier and issues
  //      the copy; meanwhile everybody clears the score map and the masks.
  const int xa = (X0 - 4) & ~15;       // box column 0 (16-byte aligned)
  const int ow = (X0 - xa) & ~3;       // tile byte of xt = 0
  const int sh = (X0 - xa) & 3;        // xt of interior column 0
  const int wpi = (sh + iw + 3) >> 2;  // words per interior row (<= 32)
  if (tid == 0) {
    tma_load_tile(s_dyn, &tmap, xa, frame * H + Y0 - 3, ctl, (uint32_t)(tg.bh * FT_TP));
    *s_cnt1 = 0;
    *s_cnt2 = 0;
  }
  if (tid < FT_MAXCELLS) s_any_ini[tid] = 0;
  if (tid >= 32 && tid < 64) {
    // bit 7 of byte b of word wx is set when column xt = 4wx + b belongs to the interior
    const int wx = tid - 32;
    const int v0 = min(max(sh - 4 * wx, 0), 4), v1 = min(max(sh + iw - 4 * wx, 0), 4);
    const uint32_t m1 = v1 >= 4 ? 0x80808080u : ((1u << (8 * v1)) - 1u) & 0x80808080u;
    const uint32_t m0 = v0 >= 4 ? 0xffffffffu : ((1u << (8 * v0)) - 1u);
    vm_tab[wx] = m1 & ~m0;
  }
  {
    // score map and masks are contiguous and 16-byte aligned: clear them with 128-bit stores
    uint4* z = reinterpret_cast<uint4*>(sc);
    const int nz = (int)(reinterpret_cast<uint8_t*>(m_min + mask_words) - sc + 15) >> 4;
    for (int i = tid; i < nz; i += FT_THREADS) z[i] = make_uint4(0u, 0u, 0u, 0u);
  }
  __syncthreads();  // mbarrier initialised for everybody, clears done
  tma_wait(ctl);
  const uint32_t th4 = (uint32_t)g.min_th * 0x01010101u;   // 1 <= minThFAST <= 127 (checked by orb_create)
  const uint32_t* tw0 = tile_w + (ow >> 2) + 3 * FT_TW;    // interior row 0, xt = 0

  // ---- pass A (byte-SIMD filter): item = (interior row y, word wx) = columns xt = 4wx .. 4wx+3
  {
    const int nitems = ih * wpi;
    const int sy = FT_THREADS / wpi, sx = FT_THREADS - sy * wpi;
    // two words per thread and trip (rows FT_THREADS / wpi apart): the two compare chains are independent, which gives the
    // scheduler something to issue while the other chain waits for its shared-memory loads
    auto filter = [&](int y, int wx) -> uint32_t {
      const uint32_t* c = &tw0[y * FT_TW + wx];
      const uint32_t C = c[0];
      const uint32_t T = c[3 * FT_TW], B = c[-3 * FT_TW];         // ring points 0 (0,+3) and 8 (0,-3)
      const uint32_t R = __funnelshift_r(C, c[1], 24);            // ring point 4 (+3,0)
      const uint32_t L = __funnelshift_r(c[-1], C, 8);            // ring point 12 (-3,0)
      // hi = C + t and lo = C - t per byte, modulo 256, with overflow / underflow flags (bit 7): a pixel whose
      // hi overflows has nothing brighter, one whose lo underflows nothing darker
      const uint32_t s7 = (C & 0x7f7f7f7fu) + th4;
      const uint32_t ov = C & s7, hi = s7 ^ (C & 0x80808080u);
      const uint32_t u = (C | 0x80808080u) - th4;
      const uint32_t lo = u & (C | 0x7f7f7f7fu), un = ~(C | u);
      const uint32_t nh7 = ~hi & 0x7f7f7f7fu, l7 = (lo & 0x7f7f7f7fu) + 0x7f7f7f7fu;
      uint32_t b0, b4, b8, b12, d0, d4, d8, d12;
      swar_cmp2(T, hi, nh7, lo, l7, b0, d0);
      swar_cmp2(R, hi, nh7, lo, l7, b4, d4);
      swar_cmp2(B, hi, nh7, lo, l7, b8, d8);
      swar_cmp2(L, hi, nh7, lo, l7, b12, d12);
      const uint32_t pb = ((b0 | b8) & ~ov) & (b4 | b12);
      const uint32_t pd = ((d0 | d8) & ~un) & (d4 | d12);
      return (pb | pd) & vm_tab[wx];
    };
    int y = tid / wpi, wx = tid - y * wpi;
    for (int it = tid; it < ((nitems + 31) & ~31); it += 2 * FT_THREADS) {
      int y1 = y + sy, wx1 = wx + sx;
      if (wx1 >= wpi) { wx1 -= wpi; ++y1; }
      const uint32_t any0 = it < nitems ? filter(y, wx) : 0u;
      const uint32_t any1 = it + FT_THREADS < nitems ? filter(y1, wx1) : 0u;
      // warp-aggregated append of the surviving words
      const uint32_t bal0 = __ballot_sync(0xffffffffu, any0 != 0), bal1 = __ballot_sync(0xffffffffu, any1 != 0);
      if (bal0 | bal1) {
        int base = 0;
        if (lane == 0) base = smem_add(s_cnt1, __popc(bal0) + __popc(bal1));
        base = __shfl_sync(0xffffffffu, base, 0);
        if (any0) list1[base + __popc(bal0 & lt)] = (uint16_t)((y << 5) | wx);
        if (any1) list1[base + __popc(bal0) + __popc(bal1 & lt)] = (uint16_t)((y1 << 5) | wx1);
      }
      wx = wx1 + sx; y = y1 + sy;
      if (wx >= wpi) { wx -= wpi; ++y; }
    }
  }
  __syncthreads();

  // ---- pass B (byte-SIMD, listed words): full 16-ring test, both polarities; corner pixels go to list2 as
  //      y << 7 | xt, bit 15 = the arc is brighter than the centre
  {
    const int n1 = *s_cnt1;
    for (int i = tid; i < ((n1 + 31) & ~31); i += FT_THREADS) {
      uint32_t cb = 0, cd = 0;
      uint32_t code0 = 0;
      if (i < n1) {
        const int e = list1[i];
        const int y = e >> 5, wx = e & 31;
        code0 = (uint32_t)((y << 7) | (4 * wx));
        const uint32_t* c = &tw0[y * FT_TW + wx];
        const uint32_t C = c[0];
        const uint32_t s7 = (C & 0x7f7f7f7fu) + th4;
        const uint32_t ov = C & s7, hi = s7 ^ (C & 0x80808080u);
        const uint32_t u = (C | 0x80808080u) - th4;
        const uint32_t lo = u & (C | 0x7f7f7f7fu), un = ~(C | u);
        const uint32_t nh7 = ~hi & 0x7f7f7f7fu, l7 = (lo & 0x7f7f7f7fu) + 0x7f7f7f7fu;
        uint32_t fb[16], fd[16];
#define FT_RING(k, dy, expr)                                                                  \
  {                                                                                           \
    const uint32_t wm = c[(dy) * FT_TW - 1], w0 = c[(dy) * FT_TW], wp = c[(dy) * FT_TW + 1];  \
    (void)wm; (void)wp;                                                                       \
    swar_cmp2((expr), hi, nh7, lo, l7, fb[k], fd[k]);                                         \
  }
        FT_RING(0, 3, w0)
        FT_RING(1, 3, __funnelshift_r(w0, wp, 8))
        FT_RING(2, 2, __funnelshift_r(w0, wp, 16))
        FT_RING(3, 1, __funnelshift_r(w0, wp, 24))
        FT_RING(4, 0, __funnelshift_r(w0, wp, 24))
        FT_RING(5, -1, __funnelshift_r(w0, wp, 24))
        FT_RING(6, -2, __funnelshift_r(w0, wp, 16))
        FT_RING(7, -3, __funnelshift_r(w0, wp, 8))
        FT_RING(8, -3, w0)
        FT_RING(9, -3, __funnelshift_r(wm, w0, 24))
        FT_RING(10, -2, __funnelshift_r(wm, w0, 16))
        FT_RING(11, -1, __funnelshift_r(wm, w0, 8))
        FT_RING(12, 0, __funnelshift_r(wm, w0, 8))
        FT_RING(13, 1, __funnelshift_r(wm, w0, 8))
        FT_RING(14, 2, __funnelshift_r(wm, w0, 16))
        FT_RING(15, 3, __funnelshift_r(wm, w0, 24))
#undef FT_RING
        // 9 contiguous ring points: a3[k] = f[k] & f[k+1] & f[k+2], arc at k = a3[k] & a3[k+3] & a3[k+6]
        uint32_t a3[16];
#pragma unroll
        for (int k = 0; k < 16; ++k) a3[k] = fb[k] & fb[(k + 1) & 15] & fb[(k + 2) & 15];
#pragma unroll
        for (int k = 0; k < 16; ++k) cb |= a3[k] & a3[(k + 3) & 15] & a3[(k + 6) & 15];
#pragma unroll
        for (int k = 0; k < 16; ++k) a3[k] = fd[k] & fd[(k + 1) & 15] & fd[(k + 2) & 15];
#pragma unroll
        for (int k = 0; k < 16; ++k) cd |= a3[k] & a3[(k + 3) & 15] & a3[(k + 6) & 15];
        const uint32_t vm = vm_tab[wx];
        cb &= ~ov & vm;
        cd &= ~un & vm;
      }
      const uint32_t any = cb | cd;
      // warp-aggregated append: one ballot per byte position, one shared atomic per warp
      const uint32_t b0 = __ballot_sync(0xffffffffu, any & 0x00000080u), b1 = __ballot_sync(0xffffffffu, any & 0x00008000u);
      const uint32_t b2 = __ballot_sync(0xffffffffu, any & 0x00800000u), b3 = __ballot_sync(0xffffffffu, any & 0x80000000u);
      const int n0 = __popc(b0), n1b = __popc(b1), n2b = __popc(b2), n3 = __popc(b3);
      if (b0 | b1 | b2 | b3) {
        int base = 0;
        if (lane == 0) base = smem_add(s_cnt2, n0 + n1b + n2b + n3);
        base = __shfl_sync(0xffffffffu, base, 0);
        if (any & 0x00000080u) list2[base + __popc(b0 & lt)] = (uint16_t)(code0 | ((cb << 8) & 0x8000u));
        base += n0;
        if (any & 0x00008000u) list2[base + __popc(b1 & lt)] = (uint16_t)((code0 + 1) | (cb & 0x8000u));
        base += n1b;
        if (any & 0x00800000u) list2[base + __popc(b2 & lt)] = (uint16_t)((code0 + 2) | ((cb >> 8) & 0x8000u));
        base += n2b;
        if (any & 0x80000000u) list2[base + __popc(b3 & lt)] = (uint16_t)((code0 + 3) | ((cb >> 16) & 0x8000u));
      }
    }
  }
  __syncthreads();
  const uint8_t* tile = reinterpret_cast<const uint8_t*>(tw0);  // interior row 0, xt = 0

  // ---- pass C: exact score of every corner: max over the 16 arcs of 9 of the minimum |difference|, minus 1
  //      (only one polarity can hold a 9-arc, the other cannot exceed the threshold)
  const int n2 = *s_cnt2;
  for (int i = tid; i < n2; i += FT_THREADS) {
    const int code = list2[i];
    const int y = (code >> 7) & 127, x = code & 127;
    const uint8_t* c = tile + y * FT_TP + x;
    const int v = c[0];
    const int sgn = (code & 0x8000) ? 1 : -1;
    int e[16];
#define FAST_E(k, off) e[k] = sgn * ((int)c[off] - v);
    FAST_E(0, 3 * FT_TP)      FAST_E(1, 3 * FT_TP + 1)   FAST_E(2, 2 * FT_TP + 2)   FAST_E(3, FT_TP + 3)
    FAST_E(4, 3)              FAST_E(5, -FT_TP + 3)      FAST_E(6, -2 * FT_TP + 2)  FAST_E(7, -3 * FT_TP + 1)
    FAST_E(8, -3 * FT_TP)     FAST_E(9, -3 * FT_TP - 1)  FAST_E(10, -2 * FT_TP - 2) FAST_E(11, -FT_TP - 3)
    FAST_E(12, -3)            FAST_E(13, FT_TP - 3)      FAST_E(14, 2 * FT_TP - 2)  FAST_E(15, 3 * FT_TP - 1)
#undef FAST_E
    int m3[16];
#pragma unroll
    for (int k = 0; k < 16; ++k) m3[k] = min(min(e[k], e[(k + 1) & 15]), e[(k + 2) & 15]);
    int best = 0;
#pragma unroll
    for (int k = 0; k < 16; k += 2)
      best = max(max(best, min(min(m3[k], m3[(k + 3) & 15]), m3[(k + 6) & 15])),
                 min(min(m3[k + 1], m3[(k + 4) & 15]), m3[(k + 7) & 15]));
    sc[(y + 1) * SP + (x + 1)] = (uint8_t)(best - 1);
  }
  __syncthreads();

  // ---- pass D: 3x3 strict non-max suppression inside the corner's cell (neighbours that belong to another
  //      cell count as 0, like the untested border of the cell's cv::FAST call); survivors set a bit in the
  //      cell's row mask
  for (int i = tid; i < n2; i += FT_THREADS) {
    const int code = list2[i];
    const int y = (code >> 7) & 127, xt = code & 127, x = xt - sh;
    const int cj = (int)(((unsigned)x * tg.mul_w) >> 16), ci = (int)(((unsigned)y * tg.mul_h) >> 16);
    const int xr = x - cj * wc, yr = y - ci * hc;
    const uint8_t* s = &sc[(y + 1) * SP + (xt + 1)];
    const int v = s[0];
    const bool okl = xr > 0, okr = xr < wc - 1, oku = yr > 0, okd = yr < hc - 1;
    const int a0 = (okl && oku) ? s[-SP - 1] : 0, a1 = oku ? s[-SP] : 0, a2 = (okr && oku) ? s[-SP + 1] : 0;
    const int a3 = okl ? s[-1] : 0, a4 = okr ? s[1] : 0;
    const int a5 = (okl && okd) ? s[SP - 1] : 0, a6 = okd ? s[SP] : 0, a7 = (okr && okd) ? s[SP + 1] : 0;
    const int mx = max(max(max(a0, a1), max(a2, a3)), max(max(a4, a5), max(a6, a7)));
    if (v > mx) {
      const int cell = ci * tg.nbx + cj;
      const int widx = (cell * hc + yr) * WPR + (xr >> 5);
      atomicOr(&m_min[widx], 1u << (xr & 31));
      if (v >= g.ini_th) { atomicOr(&m_ini[widx], 1u << (xr & 31)); s_any_ini[cell] = 1; }
    }
  }
  __syncthreads();

  // ---- pass E: ordered output, one warp per cell; a lane owns one cell row (mask words are in row-major order)
  for (int cl = wid; cl < ncx * ncy; cl += FT_THREADS / 32) {
    const int cy = cl / ncx, cx = cl - cy * ncx;
    const int cell = cy * tg.nbx + cx;
    const uint32_t* mask = (s_any_ini[cell] ? m_ini : m_min) + cell * hc * WPR;
    const size_t gc = cell_base + (size_t)(i0 + cy) * g.ncols[l] + j0 + cx;
    uint32_t* out_keys = cell_keys + gc * ORB_CELL_CAP;
    const int kx = X0 + cx * wc - ORB_BORDER, ky = Y0 + cy * hc - ORB_BORDER;
    const uint8_t* scell = sc + (cy * hc + 1) * SP + cx * wc + sh + 1;
    int carry = 0;
    for (int base = 0; base < hc; base += 32) {
      const int row = base + lane;
      uint32_t w0 = 0, w1 = 0, w2 = 0;
      if (row < hc) {
        const uint32_t* mr = mask + row * WPR;
        w0 = mr[0];
        if (WPR > 1) w1 = mr[1];
        if (WPR > 2) w2 = mr[2];
      }
      const int c = __popc(w0) + __popc(w1) + __popc(w2);
      int incl = c;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const int v = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += v;
      }
      int pos = carry + incl - c;
      carry += __shfl_sync(0xffffffffu, incl, 31);
#pragma unroll
      for (int k = 0; k < 3; ++k) {
        uint32_t w = k == 0 ? w0 : (k == 1 ? w1 : w2);
        while (w) {
          const int bit = __ffs(w) - 1;
          w &= w - 1;
          const int x = 32 * k + bit;
          if (pos < ORB_CELL_CAP) out_keys[pos] = orb_pack(kx + x, ky + row, scell[row * SP + x]);
          ++pos;
        }
      }
    }
    if (lane == 0) {
      cell_count[gc] = min(carry, ORB_CELL_CAP);
      if (carry > ORB_CELL_CAP) atomicOr(status + frame, ORB_ST_CELL_OVERFLOW);
    }
  }
}
