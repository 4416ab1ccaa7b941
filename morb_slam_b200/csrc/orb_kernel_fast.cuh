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
// The kernel is instruction-bound, not HBM-bound (profiles/README_r1.md):
//   load   ONE TMA tensor copy (cp.async.bulk.tensor.2d) per CTA brings the tile + 3-px ring into shared memory -
//          no per-thread staging instructions at all. The box must start on a 16-byte boundary of the row
//          (measured: other start columns raise "illegal instruction"), so the first interior column sits at
//          byte o = 4..19 of a tile row; the passes work in "xt" columns counted from the word that holds it
//          (xt = x + (o & 3)), which keeps every SIMD word aligned in shared memory;
//   pass A every pixel, 4 per thread in byte-SIMD: compass points 0/4/8/12 against v +- t; a 9-arc of the
//          16-ring contains two ring-adjacent compass points, so "no adjacent pair brighter and none darker"
//          rules a pixel out; survivors (with the polarities still possible) go to a shared-memory list;
//   pass B 16-bit arc mask of the possible polarity on the dense list -> corners at minThFAST;
//   pass C exact score on the dense corner list;
//   pass D NMS inside the corner's cell, corners only, sets bits in per-cell row masks;
//   pass E one warp per cell: ordered output from the mask words with a warp scan.
#pragma once
#include <cuda.h>

#define FT_MAXW 124                 // tile interior width limit (7-bit xt in the list codes: xt <= FT_MAXW + 2)
#define FT_MAXH 127                 // hard limit of the list codes (7-bit y); the host picks nby below FT_TILE_H
#define FT_TILE_H 80                // preferred tile interior height
#define FT_TP 160                   // tile pitch in bytes = TMA box width: 19 + FT_MAXW + 3, word reads up to +10
#define FT_TW (FT_TP / 4)
#define FT_THREADS 256
#define FT_MAXCELLS 16              // cells per tile (nbx <= 3 since cells are >= 35 px wide)

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

// Dynamic shared memory (all carved from one 128-byte aligned block, sized by fast_tile_smem()):
//   tile bytes [bh][FT_TP] (TMA destination) | score bytes [(ih_max + 2)][sp] | m_ini, m_min words
//   [cells][hcell][wpr] | list1, list2 u16 [list_cap] | control words
static size_t fast_tile_smem(const FastTileGeom& t, int hcell) {
  size_t b = (size_t)t.bh * FT_TP;
  b += (size_t)(t.nby * hcell + 2) * t.sp;
  b = (b + 15) & ~(size_t)15;
  b += 2 * (size_t)t.nbx * t.nby * hcell * t.wpr * 4;
  b += 2 * (size_t)t.list_cap * 2;
  b = (b + 15) & ~(size_t)15;
  b += 128;  // mbarrier, counters, per-cell flags
  return b;
}

__global__ void __launch_bounds__(FT_THREADS) k_fast_tiles(const __grid_constant__ CUtensorMap tmap, OrbGeom g, int l,
                                                           FastTileGeom tg, int* __restrict__ cell_count,
                                                           uint32_t* __restrict__ cell_keys, int cells_per_frame,
                                                           int* __restrict__ status) {
  extern __shared__ __align__(128) uint8_t s_dyn[];
  const int wc = g.wcell[l], hc = g.hcell[l];
  const int ihm = tg.nby * hc;
  const int SP = tg.sp, WPR = tg.wpr;
  uint32_t* tile_w = reinterpret_cast<uint32_t*>(s_dyn);
  uint8_t* sc = s_dyn + tg.bh * FT_TP;  // interior scores with a 1-px zero ring
  uint32_t* m_ini = reinterpret_cast<uint32_t*>(s_dyn + ((tg.bh * FT_TP + (ihm + 2) * SP + 15) & ~15));
  const int mask_words = tg.nbx * tg.nby * hc * WPR;
  uint32_t* m_min = m_ini + mask_words;
  uint16_t* list1 = reinterpret_cast<uint16_t*>(m_min + mask_words);
  uint16_t* list2 = list1 + tg.list_cap;
  uint32_t* ctl = reinterpret_cast<uint32_t*>((reinterpret_cast<uintptr_t>(list2 + tg.list_cap) + 15) & ~(uintptr_t)15);
  // ctl[0..1] mbarrier, ctl[2] list1 count, ctl[3] list2 count, ctl[4 .. 4 + FT_MAXCELLS) any-ini flag per cell
  int* s_cnt1 = reinterpret_cast<int*>(ctl + 2);
  int* s_cnt2 = reinterpret_cast<int*>(ctl + 3);
  int* s_any_ini = reinterpret_cast<int*>(ctl + 4);

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
  const uint32_t bar = (uint32_t)__cvta_generic_to_shared(ctl);
  if (tid == 0) {
    const uint32_t dst = (uint32_t)__cvta_generic_to_shared(s_dyn);
    const uint32_t bytes = (uint32_t)(tg.bh * FT_TP);
    const int cx = xa, cy = frame * H + Y0 - 3;
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;\n" ::"r"(bar));
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(bar), "r"(bytes) : "memory");
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];\n" ::"r"(dst),
        "l"(reinterpret_cast<uint64_t>(&tmap)), "r"(cx), "r"(cy), "r"(bar)
        : "memory");
    *s_cnt1 = 0;
    *s_cnt2 = 0;
  }
  if (tid < FT_MAXCELLS) s_any_ini[tid] = 0;
  {
    uint32_t* scw = reinterpret_cast<uint32_t*>(sc);
    const int nz = ((ih + 2) * SP) >> 2;
    for (int i = tid; i < nz; i += FT_THREADS) scw[i] = 0u;
    for (int i = tid; i < 2 * mask_words; i += FT_THREADS) m_ini[i] = 0u;
  }
  __syncthreads();  // mbarrier initialised for everybody, clears done
  {
    uint32_t done;
    do {
      asm volatile(
          "{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n selp.u32 %0, 1, 0, p;\n}\n"
          : "=r"(done)
          : "r"(bar)
          : "memory");
    } while (!done);
  }
  const int th = g.min_th;

  // ---- pass A (byte-SIMD): item = (interior row y, word wx) = columns xt = 4wx .. 4wx+3 = tile bytes ow + 4wx ..
  {
    const uint32_t* tw0 = tile_w + (ow >> 2);
    const int wpi = (sh + iw + 3) >> 2;       // words per interior row
    const int nitems = ih * wpi;
    const uint32_t th4 = (uint32_t)th * 0x01010101u;
    const int sy = FT_THREADS / wpi, sx = FT_THREADS - sy * wpi;
    int y = tid / wpi, wx = tid - y * wpi;
    for (int it = tid; it < ((nitems + 31) & ~31); it += FT_THREADS) {
      uint32_t pb = 0, pd = 0;  // per-byte flags (bit 7): brighter / darker arc still possible
      if (it < nitems) {
        const uint32_t* c = &tw0[(y + 3) * FT_TW + wx];
        const uint32_t C = c[0];
        const uint32_t T = c[3 * FT_TW], B = c[-3 * FT_TW];         // ring points 0 (0,+3) and 8 (0,-3)
        const uint32_t R = __funnelshift_r(C, c[1], 24);            // ring point 4 (+3,0)
        const uint32_t L = __funnelshift_r(c[-1], C, 8);            // ring point 12 (-3,0)
        // hi = min(C + t, 255), lo = max(C - t, 0) per byte (saturation keeps "r > hi" / "r < lo" exact)
        const uint32_t hi = __vaddus4(C, th4), lo = __vsubus4(C, th4);
        pb = swar_adjacent_pair(swar_gt(T, hi), swar_gt(R, hi), swar_gt(B, hi), swar_gt(L, hi));
        pd = swar_adjacent_pair(swar_gt(lo, T), swar_gt(lo, R), swar_gt(lo, B), swar_gt(lo, L));
        // drop the columns before / past the interior in the first / last word of a row
        const int v0 = max(sh - 4 * wx, 0), v1 = min(sh + iw - 4 * wx, 4);
        uint32_t vm = v1 >= 4 ? 0x80808080u : ((1u << (8 * v1)) - 1u) & 0x80808080u;
        vm &= ~((1u << (8 * v0)) - 1u);
        pb &= vm; pd &= vm;
      }
      const uint32_t any = (pb | pd) & 0x80808080u;
      // warp-aggregated append: one ballot per byte position, one shared atomic per warp
      const uint32_t b0 = __ballot_sync(0xffffffffu, any & 0x00000080u), b1 = __ballot_sync(0xffffffffu, any & 0x00008000u);
      const uint32_t b2 = __ballot_sync(0xffffffffu, any & 0x00800000u), b3 = __ballot_sync(0xffffffffu, any & 0x80000000u);
      const int n0 = __popc(b0), n1 = __popc(b1), n2 = __popc(b2), n3 = __popc(b3);
      if (b0 | b1 | b2 | b3) {
        int base = 0;
        if (lane == 0) base = atomicAdd(s_cnt1, n0 + n1 + n2 + n3);
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
  const uint8_t* tile = s_dyn + 3 * FT_TP + ow;  // interior row 0, xt = 0

  // ---- pass B: 16-ring arc masks of the polarities still possible; corners at minThFAST go to list2
  //      (bit 15 = the arc is brighter than the centre)
  {
    const int n1 = *s_cnt1;
    for (int i = tid; i < ((n1 + 31) & ~31); i += FT_THREADS) {
      bool pass = false, bright = false;
      int code = 0;
      if (i < n1) {
        code = list1[i];
        const uint8_t* c = tile + ((code >> 7) & 127) * FT_TP + (code & 127);
        const int v = c[0];
        int r[16];
        r[0] = c[3 * FT_TP];       r[1] = c[3 * FT_TP + 1];   r[2] = c[2 * FT_TP + 2];    r[3] = c[FT_TP + 3];
        r[4] = c[3];               r[5] = c[-FT_TP + 3];      r[6] = c[-2 * FT_TP + 2];   r[7] = c[-3 * FT_TP + 1];
        r[8] = c[-3 * FT_TP];      r[9] = c[-3 * FT_TP - 1];  r[10] = c[-2 * FT_TP - 2];  r[11] = c[-FT_TP - 3];
        r[12] = c[-3];             r[13] = c[FT_TP - 3];      r[14] = c[2 * FT_TP - 2];   r[15] = c[3 * FT_TP - 1];
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
        if (lane == 0) base = atomicAdd(s_cnt2, __popc(b));
        base = __shfl_sync(0xffffffffu, base, 0);
        if (pass) list2[base + __popc(b & lt)] = (uint16_t)((code & 0x3fff) | (bright ? 0x8000 : 0));
      }
    }
  }
  __syncthreads();

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
