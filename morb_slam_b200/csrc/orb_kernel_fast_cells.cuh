// FAST-9/16 per 35-px cell, warp-autonomous form (reference src/ORBextractor.cc:744-820: cv::FAST at iniThFAST on the
// cell ROI [iniX, maxX) x [iniY, maxY), cv::FAST at minThFAST when that returns nothing).
//
// ONE launch covers every level and frame of the batch. The unit of work is an ITEM = up to two horizontally adjacent
// cells of one cell row; one WARP owns an item from the tile load to the ordered output, so there is no block-wide
// barrier anywhere: warps of a CTA only share the launch. While one warp runs the ALU-heavy ring tests another one
// waits for its tile or walks its corner list, which is what keeps the issue slots busy (profiles/README_r2.md).
// Items are handed out statically (item = global warp + k * warps of the grid).
//
// Formulation (equivalence with the two cv::FAST calls: SURVEY.md Appendix A.3):
//   score S(p)     = OpenCV cornerScore<16> of a pixel that is a corner at threshold t (S >= t), else 0;
//   local maximum <=> S(p) > S(q) for the 8 neighbours q; pixels outside the cell interior count as 0
//                     (every cell has its own score map with a zero ring);
//   run 1          t = iniThFAST; a cell with at least one local maximum is finished (cv::FAST(ini) non-empty);
//   run 2          only for cells without one: t = minThFAST on that cell's columns, scores recomputed from scratch;
//   output         local maxima row-major inside the cell (the order is part of the contract).
// Running the high threshold first is what the reference does, and it is cheaper than one pass at minThFAST: 23 % of the
// 4-pixel words survive the filter at 20 against 36 % at 7 (synthetic EuRoC frames), and about 6 % of the cells need
// run 2.
//
// Passes of a run (all warp-local; lists live in the warp's shared-memory slice):
//   load   one TMA tensor copy per item (cp.async.bulk.tensor.2d -> the warp's own mbarrier): box of PW words x
//          (hcell + 6) rows whose column 0 is the 16-byte aligned image column xa <= X0 - 3. The NEXT item's copy is
//          issued as soon as the tile is no longer needed, underneath the ordered output of the current one.
//   A      every word of the interior rows, flattened over the box (lane i of a trip = words 64 k + 2 i and 64 k + 2 i + 1):
//          consecutive lanes read consecutive words for all five loads, so pass A has no shared-memory bank conflict by
//          construction. Filter = necessary condition "one end of the vertical and one end of the horizontal diameter
//          differs from the centre by more than t" on |r - v| (VABSDIFF4, polarity-free; measured 36.4 % of the words
//          against 35.6 % for the polarity-exact compass test, for half the ALU instructions). Surviving words go to
//          the word list (one ballot per 32 words), which borrows the score maps' memory until pass C.
//   B      the word list, 32 words per trip: the full 16-ring test in byte-SIMD for both polarities (exact: per-byte
//          r > v + t / r < v - t with explicit overflow / underflow flags), "9 contiguous" as AND / OR trees on the flag
//          words -> corner pixels with their polarity are appended to the corner list.
//   C      the corner list, 32 corners per trip: exact score (3-input min / max) into the cell's score map.
//   D      3x3 strict non-maximum suppression over the corner list -> bit in the cell's row mask.
//   E      ordered output from the mask words with a warp scan (lane = cell row).
// The passes are plain loops with warp-uniform trip counts (a first version interleaved them through ring queues: its
// control flow cost more instructions than the filter, profiles/README_r2.md).
// The corner list keeps its first FC_L2S entries in shared memory and spills the rest to a per-warp global buffer
// sized for "every pixel is a corner", so no input can overflow it.
#pragma once
#include "orb_tma.cuh"

#define FC_MAXG 2          // cells per item
#define FC_L2S 256         // corner-list entries in shared memory per warp
#ifndef FC_MAX_WARPS
#define FC_MAX_WARPS 12    // most warps per CTA (the host picks the number from the slice size and the batch)
#endif
#ifndef FC_MINB
#define FC_MINB 2          // CTAs per SM the register budget is compiled for (85 registers per thread)
#endif
#define FC_SLICE_BUDGET 8500   // tile + score maps + masks of a level above which its items hold one cell instead of two

static __device__ __forceinline__ void fc_mbar_init(void* bar_ptr) {
  const uint32_t bar = (uint32_t)__cvta_generic_to_shared(bar_ptr);
  asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;\n" ::"r"(bar));
  asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
}
// arm the barrier for `bytes` and issue the box copy (call from one lane)
static __device__ __forceinline__ void fc_tma_issue(void* dst, const CUtensorMap* tmap, int x, int y, void* bar_ptr, uint32_t bytes) {
  const uint32_t bar = (uint32_t)__cvta_generic_to_shared(bar_ptr);
  const uint32_t d = (uint32_t)__cvta_generic_to_shared(dst);
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(bar), "r"(bytes) : "memory");
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];\n" ::"r"(d),
               "l"(reinterpret_cast<uint64_t>(tmap)), "r"(x), "r"(y), "r"(bar)
               : "memory");
}
static __device__ __forceinline__ void fc_mbar_wait(void* bar_ptr, uint32_t phase) {
  const uint32_t bar = (uint32_t)__cvta_generic_to_shared(bar_ptr);
  uint32_t done;
  do {
    asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}\n"
                 : "=r"(done)
                 : "r"(bar), "r"(phase)
                 : "memory");
  } while (!done);
}

// per-byte |a - b| (native VABSDIFF4.U8 on sm_100a)
static __device__ __forceinline__ uint32_t fc_absdiff4(uint32_t a, uint32_t b) { return __vabsdiffu4(a, b); }

// item code: level | cell row << 4 | first cell column << 12 | cells << 20
static __host__ __device__ inline uint32_t fc_item_code(int l, int row, int col, int n) {
  return (uint32_t)l | ((uint32_t)row << 4) | ((uint32_t)col << 12) | ((uint32_t)n << 20);
}

// BIGT: iniThFAST >= 128 (the per-byte threshold arithmetic then needs the carry of bit 7; never the case for the
// reference's settings, so the common instantiation folds it away)
template <bool BIGT>
__global__ void __launch_bounds__(FC_MAX_WARPS * 32, FC_MINB)
k_fast_cells(const __grid_constant__ FastMaps maps, OrbGeom g, FastCellGeom fg, const uint32_t* __restrict__ items, int total_items,
             int* __restrict__ cell_count, uint32_t* __restrict__ cell_keys, int cells_per_frame, uint16_t* __restrict__ spill,
             int* __restrict__ status, int* __restrict__ work_counter) {
  // work_counter (large batches; zeroed before the launch): after its first item (= its global warp index) a warp takes the next
  // unclaimed item instead of a fixed stride - items differ in cost by what survives the filter, and with 25 items per warp the
  // slowest warp of a static assignment ends well after the average one. nullptr: static stride (small batches: an item per warp).
  extern __shared__ __align__(128) uint8_t s_dyn[];
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5, nwb = blockDim.x >> 5;
  const uint32_t lt = (1u << lane) - 1u;
  uint8_t* wbase = s_dyn + (size_t)wib * fg.warp_stride;
  uint32_t* vm_tab = reinterpret_cast<uint32_t*>(wbase);                // 128 bytes in front of the tile: tile word -1 is readable
  uint32_t* tile_w = reinterpret_cast<uint32_t*>(wbase + 128);
  uint8_t* sc = wbase + fg.score_off;
  uint32_t* mask = reinterpret_cast<uint32_t*>(wbase + fg.mask_off);
  uint16_t* sl2 = reinterpret_cast<uint16_t*>(wbase + fg.list_off);
  uint32_t* surv = reinterpret_cast<uint32_t*>(wbase + fg.list_off + FC_L2S * 2);   // [2 cells][32] local maxima: y << 16 | x << 8 | score
  uint64_t* bar = reinterpret_cast<uint64_t*>(wbase + fg.bar_off);
  int* ctr = reinterpret_cast<int*>(bar + 1);                                       // [0] corner-list length, [1..2] local maxima per cell
  const int gwarp = blockIdx.x * nwb + wib, gstride = gridDim.x * nwb;
  uint16_t* gl2 = spill + (size_t)gwarp * fg.spill_cap;
  auto l2_put = [&](int i, uint32_t v) {
    if (i < FC_L2S) sl2[i] = (uint16_t)v;
    else gl2[i - FC_L2S] = (uint16_t)v;
  };
  auto l2_get = [&](int i) -> uint32_t { return i < FC_L2S ? (uint32_t)sl2[i] : (uint32_t)__ldcg(gl2 + (i - FC_L2S)); };

  // decode + issue the tile copy of an item (lane 0)
  auto issue = [&](int frame, int li) {
    const uint32_t e = items[li];
    const int l = e & 15, ci = (e >> 4) & 0xff, j0 = (e >> 12) & 0xff;
    const int X0 = ORB_EDGE + j0 * g.wcell[l], Y0 = ORB_EDGE + ci * g.hcell[l];
    fc_tma_issue(tile_w, &maps.m[l], (X0 - 3) & ~15, frame * g.h[l] + Y0 - 3, bar, (uint32_t)(fg.PW[l] * 4 * (g.hcell[l] + 6)));
  };

  if (lane == 0) fc_mbar_init(bar);
  __syncwarp();
  uint32_t phase = 0;
  int item = gwarp;
  // (frame, item inside the frame) advance by a constant stride: one division per kernel instead of two per item
  const int ipf = fg.items_per_frame, sq = gstride / ipf, sr = gstride - sq * ipf;
  int frame = item / ipf, li = item - frame * ipf;
  if (item < total_items && lane == 0) {
    asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
    issue(frame, li);
  }

  while (item < total_items) {
    const uint32_t ecode = items[li];
    int next_item, nframe, nli;
    if (work_counter) {   // claimed now, used when this item's tile is free: the atomic's latency hides behind the passes
      int v = 0;
      if (lane == 0) v = gstride + atomicAdd(work_counter, 1);
      next_item = __shfl_sync(0xffffffffu, v, 0);
      nframe = next_item / ipf; nli = next_item - nframe * ipf;
    } else {
      next_item = item + gstride;
      nframe = frame + sq; nli = li + sr;
      if (nli >= ipf) { nli -= ipf; ++nframe; }
    }
    const int l = ecode & 15, ci = (ecode >> 4) & 0xff, j0 = (ecode >> 12) & 0xff, ncx = (ecode >> 20) & 3;
    const int W = g.w[l], H = g.h[l], wc = g.wcell[l], hc = g.hcell[l];
    const int PW = fg.PW[l], BW = PW * 4, SP = fg.SP[l], SCELL = fg.SCELL[l], WPR = fg.WPR[l];
    const uint32_t MPW = fg.MPW[l];                                         // ceil(2^20 / PW): idx / PW == (idx * MPW) >> 20
    const int X0 = ORB_EDGE + j0 * wc, Y0 = ORB_EDGE + ci * hc;             // first interior pixel of the item
    const int iw = min(X0 + ncx * wc, W - ORB_EDGE) - X0;                   // interior = pixels FAST actually tests
    const int ih = min(Y0 + hc, H - ORB_EDGE) - Y0;
    const int o = X0 - ((X0 - 3) & ~15);                                    // box byte of interior column 0 (3..18)
    const size_t cell_base = (size_t)frame * cells_per_frame + g.cell_start[l] + (size_t)ci * g.ncols[l] + j0;

    // clear the masks of the item while the tile is in flight
    for (int i = lane; i < ncx * hc * WPR; i += 32) mask[i] = 0u;
    if (lane < 3) ctr[lane] = 0;
    fc_mbar_wait(bar, phase);
    phase ^= 1u;

    int cnt0 = 0, cnt1 = 0;          // local maxima per cell (warp-uniform)
    uint32_t todo = 0;               // cells of the current run (bit per cell)
    if (ih > 0) todo = (iw > 0 ? 1u : 0u) | ((ncx > 1 && iw > wc) ? 2u : 0u);
    const uint32_t* tw3 = tile_w + 3 * PW;                                  // interior row 0, box word 0
    const uint8_t* tb3 = reinterpret_cast<const uint8_t*>(tw3);

    for (int run = 0; run < 2 && todo; ++run) {
      const int th = run == 0 ? g.ini_th : g.min_th;
      // ---- byte range of the run inside a box row, and the valid-byte masks of its words
      const int c_lo = (todo & 1u) ? 0 : 1, c_hi = (todo & 2u) ? 2 : 1;
      {
        const int bx0 = o + c_lo * wc, bx1 = o + min(c_hi * wc, iw);
        const int v0 = min(max(bx0 - 4 * lane, 0), 4), v1 = min(max(bx1 - 4 * lane, 0), 4);
        const uint32_t m1 = v1 >= 4 ? 0x80808080u : ((1u << (8 * v1)) - 1u) & 0x80808080u;
        const uint32_t m0 = v0 >= 4 ? 0xffffffffu : ((1u << (8 * v0)) - 1u);
        vm_tab[lane] = m1 & ~m0;
      }
      __syncwarp();
      const uint32_t t7 = (uint32_t)(th & 0x7f) * 0x01010101u;
      const uint32_t T80 = (BIGT && th >= 128) ? 0x80808080u : 0u;
      const uint32_t KA = (uint32_t)(0x7f - min(th, 127)) * 0x01010101u;   // pass A threshold (conservative above 127)

      // ---- pass A: lane = box words base + lane and base + 32 + lane of the interior rows. The word list borrows the
      //      score maps of the run's cells (dead until pass C; it only ever holds words with a valid byte, which fit)
      uint16_t* list1 = reinterpret_cast<uint16_t*>(sc + c_lo * SCELL);
      int n1 = 0;
      // |r - v| > t per byte: bit 7 of ((d & 0x7f) + (0x7f - t)) | d, for the four compass points of a word
      auto filter = [&](uint32_t C, uint32_t T, uint32_t B, uint32_t R, uint32_t L) -> uint32_t {
        const uint32_t d0 = fc_absdiff4(T, C), d8 = fc_absdiff4(B, C), d4 = fc_absdiff4(R, C), d12 = fc_absdiff4(L, C);
        const uint32_t x0 = (d0 & 0x7f7f7f7fu) + KA, x8 = (d8 & 0x7f7f7f7fu) + KA;
        const uint32_t x4 = (d4 & 0x7f7f7f7fu) + KA, x12 = (d12 & 0x7f7f7f7fu) + KA;
        return (x0 | d0 | x8 | d8) & (x4 | d4 | x12 | d12);
      };
      // a lane takes two adjacent words (64-bit loads of the centre, upper and lower row; the pitch is even, so both are in one row).
      // Only the words that hold a byte of the run are visited: the box starts up to 15 bytes before the interior and ends on a
      // 16-byte boundary after it, a quarter of its words on average hold nothing to test.
      const int w_lo = ((o + c_lo * wc) >> 2) & ~1, w_hi = (((o + min(c_hi * wc, iw) + 3) >> 2) + 1) & ~1;
      const int NW = max(w_hi - w_lo, 2), nrun = ih * NW;
      const uint32_t MNW = ((1u << 20) + NW - 1) / NW;                       // t / NW == (t * MNW) >> 20 for t < 2^15
      for (int base = 0; base < nrun; base += 64) {
        const int t = base + 2 * lane;
        int idx = 0;
        uint32_t any0 = 0, any1 = 0;
        if (t < nrun) {
          const int row = (int)(((uint32_t)t * MNW) >> 20);
          const int w = w_lo + t - row * NW;
          idx = row * PW + w;
          const uint32_t* c = tw3 + idx;
          const uint2 C = *reinterpret_cast<const uint2*>(c);
          const uint2 T = *reinterpret_cast<const uint2*>(c + 3 * PW), B = *reinterpret_cast<const uint2*>(c - 3 * PW);   // ring points 0 (0,+3), 8 (0,-3)
          const uint32_t Lw = c[-1], Rw = c[2];
          const uint2 vm = *reinterpret_cast<const uint2*>(vm_tab + w);
          // ring points 4 (+3,0) and 12 (-3,0) by funnel shifts over the neighbouring words
          any0 = filter(C.x, T.x, B.x, __funnelshift_r(C.x, C.y, 24), __funnelshift_r(Lw, C.x, 8)) & vm.x;
          any1 = filter(C.y, T.y, B.y, __funnelshift_r(C.y, Rw, 24), __funnelshift_r(C.x, C.y, 8)) & vm.y;
        }
        const uint32_t bal0 = __ballot_sync(0xffffffffu, any0 != 0), bal1 = __ballot_sync(0xffffffffu, any1 != 0);
        const int p0 = __popc(bal0);
        if (any0) list1[n1 + __popc(bal0 & lt)] = (uint16_t)idx;
        if (any1) list1[n1 + p0 + __popc(bal1 & lt)] = (uint16_t)(idx + 1);
        n1 += p0 + __popc(bal1);
      }
      __syncwarp();

      // ---- pass A2: the same necessary condition on the two DIAGONAL diameters (ring points 2 / 10 and 6 / 14) of the listed words,
      //      compacting the list in place: 24.5 % of the words pass A at t = 20, 16.4 % pass both (synthetic EuRoC frames; 4.5 % hold
      //      a corner), for 38 instructions per listed word against 205 of the ring test it spares (profiles/README_r2.md)
      {
        int n1b = 0;
        for (int i0 = 0; i0 < n1; i0 += 32) {
          uint32_t keep = 0;
          int idx = 0;
          if (i0 + lane < n1) {
            idx = list1[i0 + lane];
            const int ey = (int)(((uint32_t)idx * MPW) >> 20), ew = idx - ey * PW;
            const uint32_t* c = tw3 + idx;
            const uint32_t* up = c + 2 * PW;
            const uint32_t* dn = c - 2 * PW;
            const uint32_t C = c[0];
            const uint32_t u0 = up[0], d0w = dn[0];
            const uint32_t p2 = __funnelshift_r(u0, up[1], 16), p14 = __funnelshift_r(up[-1], u0, 16);     // (+2, +2), (-2, +2)
            const uint32_t p6 = __funnelshift_r(d0w, dn[1], 16), p10 = __funnelshift_r(dn[-1], d0w, 16);   // (+2, -2), (-2, -2)
            const uint32_t e2 = fc_absdiff4(p2, C), e10 = fc_absdiff4(p10, C), e6 = fc_absdiff4(p6, C), e14 = fc_absdiff4(p14, C);
            const uint32_t y2 = (e2 & 0x7f7f7f7fu) + KA, y10 = (e10 & 0x7f7f7f7fu) + KA;
            const uint32_t y6 = (e6 & 0x7f7f7f7fu) + KA, y14 = (e14 & 0x7f7f7f7fu) + KA;
            keep = (y2 | e2 | y10 | e10) & (y6 | e6 | y14 | e14) & vm_tab[ew];
          }
          const uint32_t bal = __ballot_sync(0xffffffffu, keep != 0);
          __syncwarp();   // the trip's reads of the list (all lanes) come before its writes
          if (keep) list1[n1b + __popc(bal & lt)] = (uint16_t)idx;   // behind the read position: entries of later trips stay intact
          n1b += __popc(bal);
        }
        n1 = n1b;
      }
      __syncwarp();

      // ---- pass B: full 16-ring test of the listed words; corner pixels go to the corner list as
      //      y << 7 | box byte column, bit 15 = the arc is brighter than the centre
      for (int i0 = 0; i0 < n1; i0 += 32) {
        uint32_t cb = 0, cd = 0, code0 = 0;
        if (i0 + lane < n1) {
          const int idx = list1[i0 + lane];
          const int ey = (int)(((uint32_t)idx * MPW) >> 20), ew = idx - ey * PW;
          code0 = (uint32_t)((ey << 7) | (4 * ew));
          const uint32_t* c = tw3 + idx;
          const uint32_t* cp1 = c + PW; const uint32_t* cp2 = cp1 + PW; const uint32_t* cp3 = cp2 + PW;
          const uint32_t* cm1 = c - PW; const uint32_t* cm2 = cm1 - PW; const uint32_t* cm3 = cm2 - PW;
          const uint32_t C = c[0];
          // hi = C + t, lo = C - t per byte modulo 256 with overflow / underflow flags (bit 7): a pixel whose hi
          // overflows has nothing brighter, one whose lo underflows nothing darker (any t in 1..254 with BIGT)
          const uint32_t s7 = (C & 0x7f7f7f7fu) + t7;
          const uint32_t u = (C | 0x80808080u) - t7;
          uint32_t hi, ov, lo, un;
          if (BIGT) {
            hi = s7 ^ (C & 0x80808080u) ^ T80;
            ov = (C & s7) | (T80 & (C | s7));
            lo = (u & 0x7f7f7f7fu) | ((C ^ T80 ^ ~u) & 0x80808080u);
            un = (~C & T80) | (~(C ^ T80) & ~u);
          } else {
            hi = s7 ^ (C & 0x80808080u);
            ov = C & s7;
            lo = u & (C | 0x7f7f7f7fu);
            un = ~(C | u);
          }
          const uint32_t nh7 = ~hi & 0x7f7f7f7fu, l7 = (lo & 0x7f7f7f7fu) + 0x7f7f7f7fu;
          uint32_t fb[16], fd[16];
#define FC_RING(k, rp, expr)                                  \
  {                                                           \
    const uint32_t wm = (rp)[-1], w0 = (rp)[0], wp = (rp)[1]; \
    (void)wm; (void)wp;                                       \
    swar_cmp2((expr), hi, nh7, lo, l7, fb[k], fd[k]);         \
  }
          FC_RING(0, cp3, w0)
          FC_RING(1, cp3, __funnelshift_r(w0, wp, 8))
          FC_RING(2, cp2, __funnelshift_r(w0, wp, 16))
          FC_RING(3, cp1, __funnelshift_r(w0, wp, 24))
          FC_RING(4, c, __funnelshift_r(w0, wp, 24))
          FC_RING(5, cm1, __funnelshift_r(w0, wp, 24))
          FC_RING(6, cm2, __funnelshift_r(w0, wp, 16))
          FC_RING(7, cm3, __funnelshift_r(w0, wp, 8))
          FC_RING(8, cm3, w0)
          FC_RING(9, cm3, __funnelshift_r(wm, w0, 24))
          FC_RING(10, cm2, __funnelshift_r(wm, w0, 16))
          FC_RING(11, cm1, __funnelshift_r(wm, w0, 8))
          FC_RING(12, c, __funnelshift_r(wm, w0, 8))
          FC_RING(13, cp1, __funnelshift_r(wm, w0, 8))
          FC_RING(14, cp2, __funnelshift_r(wm, w0, 16))
          FC_RING(15, cp3, __funnelshift_r(wm, w0, 24))
#undef FC_RING
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
          const uint32_t vm = vm_tab[ew];
          cb &= ~ov & vm;
          cd &= ~un & vm;
        }
        // append: a lane's corners (up to 4, 9 % of the lanes have any) go behind a shared-memory counter; the order of the
        // corner list does not matter (the output order comes from the sort / masks of pass E)
        const uint32_t any = cb | cd;
        if (any) {
          int pos = smem_add(ctr, __popc(any));
          if (pos + 4 <= FC_L2S) {   // common case: everything stays in shared memory
            if (any & 0x00000080u) sl2[pos++] = (uint16_t)(code0 | ((cb << 8) & 0x8000u));
            if (any & 0x00008000u) sl2[pos++] = (uint16_t)((code0 + 1) | (cb & 0x8000u));
            if (any & 0x00800000u) sl2[pos++] = (uint16_t)((code0 + 2) | ((cb >> 8) & 0x8000u));
            if (any & 0x80000000u) sl2[pos++] = (uint16_t)((code0 + 3) | ((cb >> 16) & 0x8000u));
          } else {
            if (any & 0x00000080u) l2_put(pos++, code0 | ((cb << 8) & 0x8000u));
            if (any & 0x00008000u) l2_put(pos++, (code0 + 1) | (cb & 0x8000u));
            if (any & 0x00800000u) l2_put(pos++, (code0 + 2) | ((cb >> 8) & 0x8000u));
            if (any & 0x80000000u) l2_put(pos++, (code0 + 3) | ((cb >> 16) & 0x8000u));
          }
        }
      }
      __syncwarp();
      const int n2 = ctr[0];
      const bool spilled = n2 > FC_L2S;

      // ---- score maps of the run's cells start from zero (they held the word list until here)
      {
        uint4* z = reinterpret_cast<uint4*>(sc + c_lo * SCELL);
        const int nz = ((c_hi - c_lo) * SCELL) >> 4;
        for (int i = lane; i < nz; i += 32) z[i] = make_uint4(0u, 0u, 0u, 0u);
      }
      if (run == 1) {   // run 2 starts from scratch in its cells
        for (int cj = c_lo; cj < c_hi; ++cj) {
          uint32_t* zm = mask + cj * hc * WPR;
          for (int i = lane; i < hc * WPR; i += 32) zm[i] = 0u;
        }
      }
      __syncwarp();
      if (lane == 0) ctr[0] = 0;   // every lane has read it: free for the next run

      // ---- pass C: exact score of every corner: max over the 16 arcs of 9 of the minimum |difference|, minus 1
      //      (only one polarity can hold a 9-arc, the other cannot exceed the threshold)
      for (int i0 = 0; i0 < n2; i0 += 32) {
        if (i0 + lane < n2) {
          const uint32_t code = spilled ? l2_get(i0 + lane) : (uint32_t)sl2[i0 + lane];
          const int cy = (code >> 7) & 127, xb = code & 127;
          const uint8_t* c = tb3 + cy * BW + xb;
          const int v = c[0];
          const int sgn = (code & 0x8000u) ? 1 : -1;
          const int B2 = 2 * BW, B3 = 3 * BW;
          int e[16];
#define FC_E(k, off) e[k] = sgn * ((int)c[off] - v);
          FC_E(0, B3)        FC_E(1, B3 + 1)     FC_E(2, B2 + 2)     FC_E(3, BW + 3)
          FC_E(4, 3)         FC_E(5, -BW + 3)    FC_E(6, -B2 + 2)    FC_E(7, -B3 + 1)
          FC_E(8, -B3)       FC_E(9, -B3 - 1)    FC_E(10, -B2 - 2)   FC_E(11, -BW - 3)
          FC_E(12, -3)       FC_E(13, BW - 3)    FC_E(14, B2 - 2)    FC_E(15, B3 - 1)
#undef FC_E
          int m3[16];
#pragma unroll
          for (int k = 0; k < 16; ++k) m3[k] = min(min(e[k], e[(k + 1) & 15]), e[(k + 2) & 15]);
          int best = 0;
#pragma unroll
          for (int k = 0; k < 16; k += 2)
            best = max(max(best, min(min(m3[k], m3[(k + 3) & 15]), m3[(k + 6) & 15])),
                       min(min(m3[k + 1], m3[(k + 4) & 15]), m3[(k + 7) & 15]));
          const int xi = xb - o;
          const int cj = xi >= wc ? 1 : 0;
          sc[cj * SCELL + (cy + 1) * SP + (xi - cj * wc) + 1] = (uint8_t)(best - 1);
        }
      }
      __syncwarp();

      // ---- pass D: 3x3 strict non-maximum suppression over the corner list; every cell's score map has a zero ring,
      //      so neighbours that belong to another cell count as 0 like the untested border of the cell's cv::FAST call
      for (int i0 = 0; i0 < n2; i0 += 32) {
        const int i = i0 + lane;
        if (i < n2) {
          const uint32_t code = spilled ? l2_get(i) : (uint32_t)sl2[i];
          const int cy = (code >> 7) & 127, xi = (int)(code & 127) - o;
          const int cj = xi >= wc ? 1 : 0;
          const int xr = xi - cj * wc;
          const uint8_t* s = sc + cj * SCELL + (cy + 1) * SP + xr + 1;
          const int v = s[0];
          const int mx = max(max(max((int)s[-SP - 1], (int)s[-SP]), max((int)s[-SP + 1], (int)s[-1])),
                             max(max((int)s[1], (int)s[SP - 1]), max((int)s[SP], (int)s[SP + 1])));
          if (v > mx) {
            // a cell's first 32 local maxima also go to a small list (sorted by rank in pass E); the row masks serve cells with more
            atomicOr(&mask[(cj * hc + cy) * WPR + (xr >> 5)], 1u << (xr & 31));
            const int p = smem_add(ctr + 1 + cj, 1);
            if (p < 32) surv[cj * 32 + p] = ((uint32_t)cy << 16) | ((uint32_t)xr << 8) | (uint32_t)v;
          }
        }
      }
      __syncwarp();
      cnt0 = ctr[1]; cnt1 = ctr[2];
      // cells without a local maximum at iniThFAST go through run 2 (cv::FAST(ini) returned nothing, :778-796)
      uint32_t next = 0;
      if (run == 0 && g.min_th != g.ini_th) {
        if ((todo & 1u) && cnt0 == 0) next |= 1u;
        if ((todo & 2u) && cnt1 == 0) next |= 2u;
      }
      todo = next;   // (the counters of cells that go on are still 0)
    }

    // ---- the tile is free: start the next item's copy underneath the ordered output
    __syncwarp();
    if (next_item < total_items && lane == 0) issue(nframe, nli);

    // ---- pass E: ordered output; a lane owns one cell row (mask words are in row-major order)
    for (int cj = 0; cj < ncx; ++cj) {
      const uint32_t* cmask = mask + cj * hc * WPR;
      const size_t gc = cell_base + cj;
      uint32_t* out_keys = cell_keys + gc * ORB_CELL_CAP;
      if ((cj ? cnt1 : cnt0) == 0) {
        if (lane == 0) cell_count[gc] = 0;
        continue;
      }
      const int kx = X0 + cj * wc - ORB_BORDER, ky = Y0 - ORB_BORDER;
      const int ncell = cj ? cnt1 : cnt0;
      if (ncell <= 32) {
        // few local maxima (the usual case): rank of a lane's entry = entries with a smaller (row, column)
        const uint32_t* sv = surv + cj * 32;
        const uint32_t my = lane < ncell ? sv[lane] : 0xffffffffu;
        int rank = 0;
        for (int j = 0; j < ncell; ++j) rank += sv[j] < my ? 1 : 0;   // entries differ in (row, column): bits 8..31
        if (lane < ncell) out_keys[rank] = orb_pack(kx + (int)((my >> 8) & 0xff), ky + (int)(my >> 16), (int)(my & 0xff));
        if (lane == 0) cell_count[gc] = ncell;
        continue;
      }
      const uint8_t* scell = sc + cj * SCELL + SP + 1;
      int carry = 0;
      for (int base = 0; base < hc; base += 32) {
        const int row = base + lane;
        uint32_t w0 = 0, w1 = 0, w2 = 0;
        if (row < hc) {
          const uint32_t* mr = cmask + row * WPR;
          w0 = mr[0];
          if (WPR > 1) w1 = mr[1];
          if (WPR > 2) w2 = mr[2];
        }
        const int c = __popc(w0) + __popc(w1) + __popc(w2);
        int incl = c;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
          const int v = __shfl_up_sync(0xffffffffu, incl, d);
          if (lane >= d) incl += v;
        }
        int pos = carry + incl - c;
        carry += __shfl_sync(0xffffffffu, incl, 31);
#pragma unroll
        for (int k = 0; k < 3; ++k) {
          uint32_t wd = k == 0 ? w0 : (k == 1 ? w1 : w2);
          while (wd) {
            const int bit = __ffs(wd) - 1;
            wd &= wd - 1;
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
    __syncwarp();
    item = next_item; frame = nframe; li = nli;
  }
}
