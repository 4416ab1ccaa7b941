// CUDA kernels of the extraction path (sm_100a). Included by orb_extract.cu only.
//
//   k_resize_level   ComputePyramid              src/ORBextractor.cc:1088-1112 (cv::resize INTER_LINEAR 8U)
//   k_blur7          per-level GaussianBlur      src/ORBextractor.cc:1049-1050 (7x7, sigma 2, REFLECT_101)
//   k_fast_cells     per-cell FAST-9 + NMS       src/ORBextractor.cc:744-820   (cv::FAST ini/min threshold)
//   k_octree         DistributeOctTree           src/ORBextractor.cc:540-738, DivideNode :475-523
//   k_assemble       output slot assignment      src/ORBextractor.cc:1041,1066-1079 (mono / lapping order)
//   k_orient_describe IC_Angle + rBRIEF          src/ORBextractor.cc:75-145,466-473
//
// All pixel / index arithmetic is integer; the few float expressions use explicit round-to-nearest
// intrinsics so that no FMA contraction can change a bit with respect to the reference build.
#pragma once
#include "orb_internal.h"
#include "orb_tma.cuh"

__constant__ int8_t c_pattern[1024];  // rBRIEF sampling pattern (orb_pattern_31.inc)
__constant__ int c_umax[16];          // umax of the reference ctor (src/ORBextractor.cc:451-463)

static __device__ __forceinline__ const uint8_t* lvl_ptr(const OrbGeom& g, const uint8_t* base, int frame, int l) {
  return base + g.level_base[l] + (size_t)frame * g.level_fstride[l];
}
static __device__ __forceinline__ uint8_t* lvl_ptr(const OrbGeom& g, uint8_t* base, int frame, int l) {
  return base + g.level_base[l] + (size_t)frame * g.level_fstride[l];
}

// -------------------------------------------------------------------------------------------------
// Pyramid: one launch per level (level l is resized from the RESULT of level l-1, :1101), all frames
// of the batch in grid.z. Each thread produces 4 horizontally adjacent pixels and stores them as one
// 32-bit word. Tables hold, per destination column/row, the source offset and the two 11-bit
// fixed-point weights exactly as OpenCV computes them (host side, orb_extract.cu).
// -------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_resize_level(OrbGeom g, uint8_t* __restrict__ pyr, int l,
                                                      const int2* __restrict__ xtab, const int2* __restrict__ ytab,
                                                      int area2x) {
  const int frame = blockIdx.z;
  const int dw = g.w[l], dh = g.h[l], dp = g.pitch[l];
  const int sw = g.w[l - 1], sh = g.h[l - 1], sp = g.pitch[l - 1];
  const int y = blockIdx.y * blockDim.y + threadIdx.y;
  const int x0 = (blockIdx.x * blockDim.x + threadIdx.x) * 4;
  if (y >= dh || x0 >= dw) return;
  const uint8_t* __restrict__ src = lvl_ptr(g, (const uint8_t*)pyr, frame, l - 1);
  uint8_t* dst = lvl_ptr(g, pyr, frame, l);
  uint32_t packed = 0;
  if (area2x) {
    const uint8_t* s0 = src + (size_t)(2 * y) * sp;
    const uint8_t* s1 = s0 + sp;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      int x = x0 + i;
      if (x < dw) {
        int v = (s0[2 * x] + s0[2 * x + 1] + s1[2 * x] + s1[2 * x + 1] + 2) >> 2;
        packed |= (uint32_t)v << (8 * i);
      }
    }
  } else {
    const int2 ty = ytab[y];
    const int sy0 = min(max(ty.x, 0), sh - 1), sy1 = min(max(ty.x + 1, 0), sh - 1);
    const int b0 = ty.y & 0xffff, b1 = ty.y >> 16;
    const uint8_t* r0 = src + (size_t)sy0 * sp;
    const uint8_t* r1 = src + (size_t)sy1 * sp;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      int x = x0 + i;
      if (x < dw) {
        const int2 tx = xtab[x];
        const int sx = tx.x, sx1 = min(sx + 1, sw - 1);
        const int a0 = tx.y & 0xffff, a1 = tx.y >> 16;
        const int h0 = r0[sx] * a0 + r0[sx1] * a1;
        const int h1 = r1[sx] * a0 + r1[sx1] * a1;
        int v = (((b0 * (h0 >> 4)) >> 16) + ((b1 * (h1 >> 4)) >> 16) + 2) >> 2;
        v = min(max(v, 0), 255);
        packed |= (uint32_t)v << (8 * i);
      }
    }
  }
  if (x0 + 3 < dw) {
    *reinterpret_cast<uint32_t*>(dst + (size_t)y * dp + x0) = packed;  // pitch and x0 are multiples of 4
  } else {
    for (int i = 0; i < 4 && x0 + i < dw; ++i) dst[(size_t)y * dp + x0 + i] = (uint8_t)(packed >> (8 * i));
  }
}

// -------------------------------------------------------------------------------------------------
// Pyramid, tile version (all levels whose ratio is not exactly 2): one CTA produces RS_OW x RS_OH pixels
// of level l. The source window of level l-1 arrives with ONE TMA tensor copy; a thread owns 4 adjacent
// destination columns (their source offsets and weights stay in registers) and walks down RS_ROWS
// destination rows; the horizontal interpolation of a source row, (r[sx] * a0 + r[sx + 1] * a1) >> 4, is
// computed once per source row and reused by the (on average 1.67) destination rows that blend it.
// Arithmetic exactly as cv::resize INTER_LINEAR 8U, same tables as k_resize_level.
// -------------------------------------------------------------------------------------------------
#define RS_OW 128
#ifndef RS_ROWS
#define RS_ROWS 8
#endif
#ifndef RS_WARPS
#define RS_WARPS 4   // 128 x 32 destination pixels per CTA: measured 0.208 -> 0.202 ms per 256 images against 8 warps (12 / 16 rows per thread: 0.197 / 0.199, longer CTAs on the single-pair path)
#endif
#define RS_OH (RS_ROWS * RS_WARPS)

// WORDS (every level whose 4-column groups span at most 8 source bytes, i.e. ratios below about 2): the horizontal pass of a source
// row reads THREE aligned words (12 bytes from the word that holds the thread's first source byte) instead of 8 single bytes, brings
// the 8-byte window to its first byte with two funnel shifts, gathers the (left, right) byte pairs of two columns with one PRMT and
// forms byte * a0 + byte * a1 as a 2-way dot product of the table's packed 16-bit weight pair (IDP.2A): 16 instructions per source
// row and thread against 28, and a third of the shared-memory wavefronts.
static __device__ __forceinline__ uint32_t dp2a_lo_uu(uint32_t w16, uint32_t px) {
  uint32_t d;
  asm("dp2a.lo.u32.u32 %0, %1, %2, %3;" : "=r"(d) : "r"(w16), "r"(px), "r"(0u));
  return d;
}
static __device__ __forceinline__ uint32_t dp2a_hi_uu(uint32_t w16, uint32_t px) {
  uint32_t d;
  asm("dp2a.hi.u32.u32 %0, %1, %2, %3;" : "=r"(d) : "r"(w16), "r"(px), "r"(0u));
  return d;
}

template <bool WORDS>
__global__ void __launch_bounds__(RS_WARPS * 32) k_resize_tiles(const __grid_constant__ CUtensorMap tmap, OrbGeom g,
                                                               uint8_t* __restrict__ pyr, int l,
                                                               const int2* __restrict__ xtab, const int2* __restrict__ ytab,
                                                               int bw, int bh) {
  extern __shared__ __align__(128) uint8_t s_rs[];
  uint64_t* bar = reinterpret_cast<uint64_t*>(s_rs + bw * bh);
  const int frame = blockIdx.z;
  const int dw = g.w[l], dh = g.h[l], dp = g.pitch[l];
  const int sw = g.w[l - 1], sh = g.h[l - 1];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const int ox0 = blockIdx.x * RS_OW, oy0 = blockIdx.y * RS_OH;
  // source window: columns from the 16-byte boundary at or before the first source column, rows from the first source row
  const int xa = xtab[ox0].x & ~15;
  const int ya = min(max(ytab[oy0].x, 0), sh - 1);
  if (threadIdx.x == 0) tma_load_tile(s_rs, &tmap, xa, frame * sh + ya, bar, (uint32_t)(bw * bh));
  // this thread's 4 destination columns (clamped for the table look-up; invalid ones are not stored)
  const int x0 = ox0 + 4 * lane;
  int lx0[4], lx1[4], a0[4], a1[4];
  uint32_t wgt[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int2 tx = xtab[min(x0 + i, dw - 1)];
    lx0[i] = tx.x - xa;
    lx1[i] = min(tx.x + 1, sw - 1) - xa;
    a0[i] = tx.y & 0xffff;
    a1[i] = tx.y >> 16;
    wgt[i] = (uint32_t)tx.y;   // a0 | a1 << 16
  }
  // WORDS: window of 8 bytes from the thread's first source byte; PRMT selectors of the byte pairs of columns (0, 1) and (2, 3)
  const int wofs = lx0[0] & ~3;
  const uint32_t wsh = 8u * (uint32_t)(lx0[0] & 3);
  const uint32_t sel01 = (uint32_t)((lx0[0] - lx0[0]) | ((lx1[0] - lx0[0]) << 4) | ((lx0[1] - lx0[0]) << 8) | ((lx1[1] - lx0[0]) << 12));
  const uint32_t sel23 = (uint32_t)((lx0[2] - lx0[0]) | ((lx1[2] - lx0[0]) << 4) | ((lx0[3] - lx0[0]) << 8) | ((lx1[3] - lx0[0]) << 12));
  __syncthreads();
  tma_wait(bar);
  if (x0 >= dw) return;
  uint8_t* dst = lvl_ptr(g, pyr, frame, l) + x0;
  int rb = -1;               // source row (window-relative) whose horizontal pass is held in hb
  int ha[4], hb[4];
  const int y_begin = oy0 + wid * RS_ROWS, y_end = min(y_begin + RS_ROWS, dh);
  for (int y = y_begin; y < y_end; ++y) {
    const int2 ty = ytab[y];
    const int r0 = min(max(ty.x, 0), sh - 1) - ya, r1 = min(max(ty.x + 1, 0), sh - 1) - ya;
    const int b0 = ty.y & 0xffff, b1 = ty.y >> 16;
    // the upper row of this destination row is usually the lower row of the previous one (ratio < 2)
    if (r0 == rb) {
#pragma unroll
      for (int i = 0; i < 4; ++i) ha[i] = hb[i];
    } else {
      const uint8_t* p = s_rs + r0 * bw;
      if (WORDS) {
        const uint32_t* q = reinterpret_cast<const uint32_t*>(p + wofs);
        const uint32_t w0 = q[0], w1 = q[1], w2 = q[2];
        const uint32_t lo = __funnelshift_r(w0, w1, wsh), hi = __funnelshift_r(w1, w2, wsh);
        const uint32_t p01 = __byte_perm(lo, hi, sel01), p23 = __byte_perm(lo, hi, sel23);
        ha[0] = (int)(dp2a_lo_uu(wgt[0], p01) >> 4); ha[1] = (int)(dp2a_hi_uu(wgt[1], p01) >> 4);
        ha[2] = (int)(dp2a_lo_uu(wgt[2], p23) >> 4); ha[3] = (int)(dp2a_hi_uu(wgt[3], p23) >> 4);
      } else {
#pragma unroll
        for (int i = 0; i < 4; ++i) ha[i] = (p[lx0[i]] * a0[i] + p[lx1[i]] * a1[i]) >> 4;
      }
    }
    {
      const uint8_t* p = s_rs + r1 * bw;
      if (WORDS) {
        const uint32_t* q = reinterpret_cast<const uint32_t*>(p + wofs);
        const uint32_t w0 = q[0], w1 = q[1], w2 = q[2];
        const uint32_t lo = __funnelshift_r(w0, w1, wsh), hi = __funnelshift_r(w1, w2, wsh);
        const uint32_t p01 = __byte_perm(lo, hi, sel01), p23 = __byte_perm(lo, hi, sel23);
        hb[0] = (int)(dp2a_lo_uu(wgt[0], p01) >> 4); hb[1] = (int)(dp2a_hi_uu(wgt[1], p01) >> 4);
        hb[2] = (int)(dp2a_lo_uu(wgt[2], p23) >> 4); hb[3] = (int)(dp2a_hi_uu(wgt[3], p23) >> 4);
      } else {
#pragma unroll
        for (int i = 0; i < 4; ++i) hb[i] = (p[lx0[i]] * a0[i] + p[lx1[i]] * a1[i]) >> 4;
      }
      rb = r1;
    }
    // no clamp needed: the weights of an axis sum to 2048 (+-1), so each term is <= 1020 and the sum < 1024
    uint32_t packed = 0;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int v = (((b0 * ha[i]) >> 16) + ((b1 * hb[i]) >> 16) + 2) >> 2;
      packed |= (uint32_t)v << (8 * i);
    }
    // the pitch is a multiple of 16 and x0 of 4: the padded tail of a row may be overwritten freely
    *reinterpret_cast<uint32_t*>(dst + (size_t)y * dp) = packed;
  }
}

#include "orb_libm_glibc.cuh"
#include "orb_kernel_remap.cuh"
#include "orb_kernel_blur.cuh"
#include "orb_kernel_fast.cuh"
#include "orb_kernel_fast_cells.cuh"

// -------------------------------------------------------------------------------------------------
// Quad-tree distribution. One warp per (frame, level) runs the reference's list algorithm exactly:
// the std::list is a doubly linked list of node slots in shared memory, every node owns a contiguous
// segment of the key array and DivideNode is a warp-cooperative stable 4-way partition (ballot +
// prefix popcount) between two ping-pong key buffers. The std::sort of Phase B (:667) is emulated
// step by step (libstdc++ introsort: median-of-3 to first, unguarded partition, threshold 16,
// heap-sort fallback, final insertion sort) on (count << 16 | UL.x, node) records because the
// reference's result depends on how std::sort permutes equivalent elements.
// Control flow is warp-uniform: every lane carries the same scalars; shared-memory writes of the
// list / node state are done by lane 0 and published with __syncwarp().
// -------------------------------------------------------------------------------------------------
#define TREE_NULL 0xffffu

struct TreeSmem {
  uint32_t* keys[2];
  uint32_t* n_bc;    // begin | count << 16
  uint32_t* n_x;     // UL.x | UR.x << 16
  uint32_t* n_y;     // UL.y | BR.y << 16
  uint32_t* n_ln;    // prev | next << 16
  uint8_t* n_fl;     // bit 0: bNoMore, bit 1: key buffer id
  uint16_t* free_list;
  unsigned long long* rec;   // vSizeAndPointerToNode: key << 32 | node
  unsigned long long* prev;  // vPrevSizeAndPointerToNode
};

static __device__ __forceinline__ bool rec_less(unsigned long long a, unsigned long long b) {
  return (uint32_t)(a >> 32) < (uint32_t)(b >> 32);  // compareNodes: (nKeys, UL.x) lexicographic, payload ignored
}

static __device__ void dev_adjust_heap(unsigned long long* a, int hole, int len, unsigned long long value) {
  const int top = hole;
  int child = hole;
  while (child < (len - 1) / 2) {
    child = 2 * (child + 1);
    if (rec_less(a[child], a[child - 1])) child--;
    a[hole] = a[child];
    hole = child;
  }
  if ((len & 1) == 0 && child == (len - 2) / 2) {
    child = 2 * (child + 1);
    a[hole] = a[child - 1];
    hole = child - 1;
  }
  int parent = (hole - 1) / 2;
  while (hole > top && rec_less(a[parent], value)) {
    a[hole] = a[parent];
    hole = parent;
    parent = (hole - 1) / 2;
  }
  a[hole] = value;
}

static __device__ void dev_heap_sort(unsigned long long* a, int n) {
  if (n >= 2) {
    int parent = (n - 2) / 2;
    while (true) {
      unsigned long long v = a[parent];
      dev_adjust_heap(a, parent, n, v);
      if (parent == 0) break;
      parent--;
    }
  }
  int last = n;
  while (last > 1) {
    --last;
    unsigned long long v = a[last];
    a[last] = a[0];
    dev_adjust_heap(a, 0, last, v);
  }
}

// single-thread emulation of libstdc++ std::sort(first, last, compareNodes) in its two halves: __introsort_loop (quicksort down to
// ranges of at most 16 elements, heap sort below the depth limit) and __final_insertion_sort. The insertion sorts only ever move
// an element in front of strictly greater ones, so the second half is a STABLE sort of whatever the first half left behind - the
// block-parallel quad-tree kernel replaces it by a rank computation over all threads and runs the first half on a whole warp
// (orb_kernel_octree_passes.cuh: op_introsort_loop_warp); this single-thread form serves the one-warp kernel k_octree.
static __device__ void dev_introsort_loop(unsigned long long* a, int n) {
  if (n <= 16) return;
  int stack_first[40], stack_last[40], stack_depth[40];
  int sp = 0;
  int first = 0, last = n, depth = 2 * (31 - __clz(n));
  while (true) {
    while (last - first > 16) {
      if (depth == 0) { dev_heap_sort(a + first, last - first); break; }
      --depth;
      const int mid = first + (last - first) / 2;
      int ia = first + 1, ib = mid, ic = last - 1, pick;
      const unsigned long long va = a[ia], vb = a[ib], vc = a[ic];
      if (rec_less(va, vb)) {
        if (rec_less(vb, vc)) pick = ib;
        else if (rec_less(va, vc)) pick = ic;
        else pick = ia;
      } else if (rec_less(va, vc)) pick = ia;
      else if (rec_less(vb, vc)) pick = ic;
      else pick = ib;
      unsigned long long t = a[first]; a[first] = a[pick]; a[pick] = t;
      const unsigned long long pivot = a[first];
      int lo = first + 1, hi = last;
      while (true) {
        while (rec_less(a[lo], pivot)) ++lo;
        --hi;
        while (rec_less(pivot, a[hi])) --hi;
        if (!(lo < hi)) break;
        t = a[lo]; a[lo] = a[hi]; a[hi] = t;
        ++lo;
      }
      // recurse on [lo, last) (deferred on the stack), continue with [first, lo)
      stack_first[sp] = lo; stack_last[sp] = last; stack_depth[sp] = depth; ++sp;
      last = lo;
    }
    if (sp == 0) break;
    --sp;
    first = stack_first[sp]; last = stack_last[sp]; depth = stack_depth[sp];
  }
}

static __device__ void dev_std_sort(unsigned long long* a, int n) {
  if (n <= 1) return;
  dev_introsort_loop(a, n);
  // __final_insertion_sort
  const int guarded = n > 16 ? 16 : n;
  for (int i = 1; i < guarded; ++i) {
    const unsigned long long v = a[i];
    if (rec_less(v, a[0])) {
      for (int j = i; j > 0; --j) a[j] = a[j - 1];
      a[0] = v;
    } else {
      int j = i;
      while (rec_less(v, a[j - 1])) { a[j] = a[j - 1]; --j; }
      a[j] = v;
    }
  }
  for (int i = guarded; i < n; ++i) {
    const unsigned long long v = a[i];
    int j = i;
    while (rec_less(v, a[j - 1])) { a[j] = a[j - 1]; --j; }
    a[j] = v;
  }
}

struct TreeState {
  int head, size, free_top, nrec;
  bool overflow;
};

static __device__ __forceinline__ void tree_push_front(const TreeSmem& S, TreeState& T, int c, int lane) {
  if (lane == 0) {
    S.n_ln[c] = TREE_NULL | ((uint32_t)(T.head < 0 ? TREE_NULL : T.head) << 16);
    if (T.head >= 0) S.n_ln[T.head] = (S.n_ln[T.head] & 0xffff0000u) | (uint32_t)c;
  }
  T.head = c;
  T.size++;
  __syncwarp();
}

// unlink node n, recycle its slot, return the following node (or -1)
static __device__ __forceinline__ int tree_erase(const TreeSmem& S, TreeState& T, int n, int lane) {
  const uint32_t ln = S.n_ln[n];
  const int p = (ln & 0xffffu) == TREE_NULL ? -1 : (int)(ln & 0xffffu);
  const int q = (ln >> 16) == TREE_NULL ? -1 : (int)(ln >> 16);
  __syncwarp();
  if (lane == 0) {
    if (p >= 0) S.n_ln[p] = (S.n_ln[p] & 0x0000ffffu) | ((uint32_t)(q < 0 ? TREE_NULL : q) << 16);
    if (q >= 0) S.n_ln[q] = (S.n_ln[q] & 0xffff0000u) | (uint32_t)(p < 0 ? TREE_NULL : p);
    S.free_list[T.free_top] = (uint16_t)n;
  }
  if (p < 0) T.head = q;
  T.free_top++;
  T.size--;
  __syncwarp();
  return q;
}

// DivideNode (:475-523): stable partition of the node's keys into n1 (UL), n2 (UR), n3 (BL), n4 (BR);
// creates the non-empty children (slots in child[]), records multi-key children, does not touch the list.
static __device__ __forceinline__ void tree_divide(const TreeSmem& S, TreeState& T, int n, int child[4], int lane) {
  const uint32_t bc = S.n_bc[n], nx = S.n_x[n], ny = S.n_y[n];
  const int begin = bc & 0xffff, count = bc >> 16;
  const int ulx = nx & 0xffff, urx = nx >> 16, uly = ny & 0xffff, bry = ny >> 16;
  const int buf = (S.n_fl[n] >> 1) & 1;
  const int midX = ulx + ((urx - ulx + 1) >> 1);  // UL.x + ceil((UR.x - UL.x) / 2)
  const int midY = uly + ((bry - uly + 1) >> 1);
  const uint32_t* __restrict__ src = S.keys[buf] + begin;
  uint32_t* dst = S.keys[buf ^ 1] + begin;
  const uint32_t lt = (1u << lane) - 1u;
  int c[4] = {0, 0, 0, 0};
  if (count <= 32) {
    const bool valid = lane < count;
    const uint32_t k = valid ? src[lane] : 0u;
    const int q = (orb_px(k) < midX ? 0 : 1) + (orb_py(k) < midY ? 0 : 2);
    uint32_t b[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) { b[i] = __ballot_sync(0xffffffffu, valid && q == i); c[i] = __popc(b[i]); }
    if (valid) {
      const int off = (q > 0 ? c[0] : 0) + (q > 1 ? c[1] : 0) + (q > 2 ? c[2] : 0);
      dst[off + __popc(b[q] & lt)] = k;
    }
  } else {
    for (int base = 0; base < count; base += 32) {
      const int i = base + lane;
      const bool valid = i < count;
      const uint32_t k = valid ? src[i] : 0u;
      const int q = (orb_px(k) < midX ? 0 : 1) + (orb_py(k) < midY ? 0 : 2);
#pragma unroll
      for (int j = 0; j < 4; ++j) c[j] += __popc(__ballot_sync(0xffffffffu, valid && q == j));
    }
    int run[4] = {0, c[0], c[0] + c[1], c[0] + c[1] + c[2]};
    for (int base = 0; base < count; base += 32) {
      const int i = base + lane;
      const bool valid = i < count;
      const uint32_t k = valid ? src[i] : 0u;
      const int q = (orb_px(k) < midX ? 0 : 1) + (orb_py(k) < midY ? 0 : 2);
      uint32_t b[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) b[j] = __ballot_sync(0xffffffffu, valid && q == j);
      if (valid) dst[run[q] + __popc(b[q] & lt)] = k;
#pragma unroll
      for (int j = 0; j < 4; ++j) run[j] += __popc(b[j]);
    }
  }
  int off = 0;
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    child[q] = -1;
    if (c[q] > 0) {
      if (T.free_top <= 0) { T.overflow = true; }
      else {
        const int slot = S.free_list[T.free_top - 1];
        T.free_top--;
        child[q] = slot;
        if (lane == 0) {
          const int cx0 = (q & 1) ? midX : ulx, cx1 = (q & 1) ? urx : midX;
          const int cy0 = (q & 2) ? midY : uly, cy1 = (q & 2) ? bry : midY;
          S.n_bc[slot] = (uint32_t)(begin + off) | ((uint32_t)c[q] << 16);
          S.n_x[slot] = (uint32_t)cx0 | ((uint32_t)cx1 << 16);
          S.n_y[slot] = (uint32_t)cy0 | ((uint32_t)cy1 << 16);
          S.n_fl[slot] = (uint8_t)((c[q] == 1 ? 1 : 0) | ((buf ^ 1) << 1));
          if (c[q] > 1) S.rec[T.nrec] = ((unsigned long long)(((uint32_t)c[q] << 16) | (uint32_t)cx0) << 32) | (uint32_t)slot;
        }
        if (c[q] > 1) T.nrec++;
      }
    }
    off += c[q];
  }
  __syncwarp();
}

// Candidate gather (fused compaction): the per-cell lists of one (frame, level) are concatenated in reference order (cells
// row-major, src/ORBextractor.cc:811-818) straight into the quad-tree's key buffer - shared memory when they fit, the level's
// global ping-pong buffer otherwise. Warp 0 scans the cell counts into an offset table (`offs`, cells + 1 ints of scratch: the
// second key buffer, which is free until the roots are split); then every thread takes candidates i, i + threads, ... and finds
// the cell of each with a binary search over the offsets - the loads of different candidates are independent, so a single warp
// keeps many of them in flight (a per-cell copy loop is a chain of dependent global loads: 0.1 ms per 256 frames).
// Returns the number of candidates (the same value in every thread); -1 if they exceed the level's capacity.
template <bool BLOCK>
static __device__ int tree_gather_cells(const OrbGeom& g, int l, int frame, const int* __restrict__ cell_count,
                                        const uint32_t* __restrict__ cell_keys, int cells_per_frame, uint32_t* dst_smem, int smem_keys,
                                        uint32_t* dst_glob, int* offs, int tid, int nthreads) {
  const int c0 = g.cell_start[l], nc = g.cell_start[l + 1] - c0, lane = tid & 31;
  const int* cc = cell_count + (size_t)frame * cells_per_frame + c0;
  if (tid < 32) {
    int carry = 0;
    for (int base = 0; base < nc; base += 32) {
      const int cnt = (base + lane < nc) ? cc[base + lane] : 0;
      int incl = cnt;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const int v = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += v;
      }
      if (base + lane < nc) offs[base + lane] = carry + incl - cnt;
      carry += __shfl_sync(0xffffffffu, incl, 31);
    }
    if (lane == 0) offs[nc] = carry;
  }
  if (BLOCK) __syncthreads(); else __syncwarp();
  const int total = offs[nc];
  if (total > g.level_cap[l]) return -1;
  uint32_t* dst = total > smem_keys ? dst_glob : dst_smem;
  const uint32_t* ck = cell_keys + ((size_t)frame * cells_per_frame + c0) * ORB_CELL_CAP;
#pragma unroll 4
  for (int i = tid; i < total; i += nthreads) {
    int lo = 0, hi = nc - 1;   // largest cell with offs[cell] <= i (empty cells share an offset with their successor)
    while (lo < hi) {
      const int mid = (lo + hi + 1) >> 1;
      if (offs[mid] <= i) lo = mid; else hi = mid - 1;
    }
    dst[i] = ck[(size_t)lo * ORB_CELL_CAP + (i - offs[lo])];
  }
  if (BLOCK) __syncthreads(); else __syncwarp();   // the offset table's memory is handed back to the caller
  return total;
}

__global__ void __launch_bounds__(32) k_octree(OrbGeom g, const int* __restrict__ cell_count,
                                               const uint32_t* __restrict__ cell_keys, int cells_per_frame,
                                               uint32_t* __restrict__ tree_scratch, int* __restrict__ lvl_count,
                                               int* __restrict__ sel_count, uint32_t* __restrict__ sel_keys,
                                               int* __restrict__ status,
                                               // sizing: level (-1: blockIdx.y), node slots, keys that fit shared memory
                                               int l_arg, int NC, int smem_keys,
                                               // debug entry: explicit candidate list instead of the cell slots
                                               const uint32_t* __restrict__ dbg_keys, int dbg_n) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int frame = blockIdx.x, lane = threadIdx.x;
  const int l = l_arg >= 0 ? l_arg : (int)blockIdx.y;
  TreeSmem S;
  {
    unsigned char* p = smem_raw;
    S.rec = (unsigned long long*)p; p += sizeof(unsigned long long) * NC;
    S.prev = (unsigned long long*)p; p += sizeof(unsigned long long) * NC;
    S.keys[0] = (uint32_t*)p; p += sizeof(uint32_t) * smem_keys;
    S.keys[1] = (uint32_t*)p; p += sizeof(uint32_t) * smem_keys;
    S.n_bc = (uint32_t*)p; p += sizeof(uint32_t) * NC;
    S.n_x = (uint32_t*)p; p += sizeof(uint32_t) * NC;
    S.n_y = (uint32_t*)p; p += sizeof(uint32_t) * NC;
    S.n_ln = (uint32_t*)p; p += sizeof(uint32_t) * NC;
    S.free_list = (uint16_t*)p; p += sizeof(uint16_t) * NC;
    S.n_fl = (uint8_t*)p;
  }
  const int N = g.nfeat[l];
  int* out_count = sel_count + (size_t)frame * g.nlevels + l;
  uint32_t* out_keys = sel_keys + ((size_t)frame * g.nlevels + l) * g.lvl_kcap;

  // ---- the level's candidates in reference order (cells row-major, row-major inside a cell): gathered from the FAST
  //      kernel's per-cell lists into the key buffer (shared memory, or this (frame, level)'s global buffer A when they
  //      do not fit: rare, same code through generic pointers)
  uint32_t* gA = tree_scratch + (size_t)frame * g.scratch_frame + g.scratch_off[l];
  int n;
  if (dbg_keys) {
    n = dbg_n;
    if (n > g.level_cap[l]) n = -1;
  } else {
    const int ncell = g.cell_start[l + 1] - g.cell_start[l];
    int* offs = ncell + 1 <= smem_keys ? (int*)S.keys[1] : (int*)(gA + g.level_cap[l]);
    n = tree_gather_cells<false>(g, l, frame, cell_count, cell_keys, cells_per_frame, S.keys[0], smem_keys, gA, offs, lane, 32);
    if (lane == 0) lvl_count[(size_t)frame * g.nlevels + l] = max(n, 0);
  }
  if (n < 0) {
    if (lane == 0) { atomicOr(status + frame, ORB_ST_LEVEL_OVERFLOW); *out_count = 0; }
    return;
  }
  if (n == 0) {
    if (lane == 0) *out_count = 0;
    return;
  }
  if (n > smem_keys) {
    S.keys[0] = gA;
    S.keys[1] = gA + g.level_cap[l];
  }
  if (dbg_keys)
    for (int i = lane; i < n; i += 32) S.keys[0][i] = dbg_keys[i];
  __syncwarp();

  // ---- roots (:545-582): key -> root (int)(pt.x / hX), stable; empty roots dropped
  const int nIni = g.nini[l];
  const float hX = g.hx[l];
  const int regW = g.w[l] - 2 * ORB_BORDER, regH = g.h[l] - 2 * ORB_BORDER;
  TreeState T;
  T.head = -1; T.size = 0; T.free_top = NC; T.nrec = 0; T.overflow = false;
  for (int i = lane; i < NC; i += 32) S.free_list[i] = (uint16_t)(NC - 1 - i);  // pop order 0,1,2,...
  __syncwarp();
  {
    const uint32_t lt = (1u << lane) - 1u;
    int wpos = 0, tail = -1;
    for (int r = 0; r < nIni; ++r) {
      const int rbegin = wpos;
      for (int base = 0; base < n; base += 32) {
        const int i = base + lane;
        bool f = false;
        uint32_t k = 0;
        if (i < n) {
          k = S.keys[0][i];
          int rr = (int)__fdiv_rn((float)orb_px(k), hX);
          rr = min(rr, nIni - 1);
          f = (rr == r);
        }
        const uint32_t b = __ballot_sync(0xffffffffu, f);
        if (f) S.keys[1][wpos + __popc(b & lt)] = k;
        wpos += __popc(b);
      }
      const int cnt = wpos - rbegin;
      if (cnt == 0) continue;
      const int slot = S.free_list[T.free_top - 1];
      T.free_top--;
      if (lane == 0) {
        const int ulx = (int)__fmul_rn(hX, (float)r), urx = (int)__fmul_rn(hX, (float)(r + 1));
        S.n_bc[slot] = (uint32_t)rbegin | ((uint32_t)cnt << 16);
        S.n_x[slot] = (uint32_t)ulx | ((uint32_t)urx << 16);
        S.n_y[slot] = 0u | ((uint32_t)regH << 16);
        S.n_fl[slot] = (uint8_t)((cnt == 1 ? 1 : 0) | (1 << 1));  // keys live in buffer 1
        S.n_ln[slot] = (uint32_t)(tail < 0 ? TREE_NULL : tail) | ((uint32_t)TREE_NULL << 16);
        if (tail >= 0) S.n_ln[tail] = (S.n_ln[tail] & 0x0000ffffu) | ((uint32_t)slot << 16);
      }
      if (tail < 0) T.head = slot;
      tail = slot;
      T.size++;
      __syncwarp();
    }
    (void)regW;
  }

  // ---- main loop (:591-716)
  bool finish = false;
  while (!finish && !T.overflow) {
    const int prev_size = T.size;
    T.nrec = 0;
    int n_to_expand = 0;
    int it = T.head;
    while (it >= 0 && !T.overflow) {
      if (S.n_fl[it] & 1) {  // bNoMore
        const uint32_t ln = S.n_ln[it];
        it = (ln >> 16) == TREE_NULL ? -1 : (int)(ln >> 16);
        continue;
      }
      int ch[4];
      const int before = T.nrec;
      tree_divide(S, T, it, ch, lane);
      n_to_expand += T.nrec - before;
#pragma unroll
      for (int q = 0; q < 4; ++q)
        if (ch[q] >= 0) tree_push_front(S, T, ch[q], lane);
      it = tree_erase(S, T, it, lane);
    }
    if (T.size >= N || T.size == prev_size) {
      finish = true;
    } else if (T.size + n_to_expand * 3 > N) {
      while (!finish && !T.overflow) {
        const int psize = T.size;
        const int np = T.nrec;
        for (int i = lane; i < np; i += 32) S.prev[i] = S.rec[i];
        T.nrec = 0;
        __syncwarp();
        if (lane == 0) dev_std_sort(S.prev, np);
        __syncwarp();
        for (int j = np - 1; j >= 0 && !T.overflow; --j) {
          const int node = (int)(uint32_t)(S.prev[j] & 0xffffffffull);
          int ch[4];
          tree_divide(S, T, node, ch, lane);
#pragma unroll
          for (int q = 0; q < 4; ++q)
            if (ch[q] >= 0) tree_push_front(S, T, ch[q], lane);
          tree_erase(S, T, node, lane);
          if (T.size >= N) break;
        }
        if (T.size >= N || T.size == psize) finish = true;
      }
    }
  }
  if (T.overflow) {
    if (lane == 0) { atomicOr(status + frame, ORB_ST_NODE_OVERFLOW); *out_count = 0; }
    return;
  }

  // ---- best response per leaf, first maximum wins, list order (:718-735)
  uint16_t* order = (uint16_t*)S.prev;
  {
    int it = T.head, i = 0;
    while (it >= 0) {
      if (lane == 0) order[i] = (uint16_t)it;
      const uint32_t ln = S.n_ln[it];
      it = (ln >> 16) == TREE_NULL ? -1 : (int)(ln >> 16);
      ++i;
    }
  }
  __syncwarp();
  const int nout = T.size;
  if (nout > g.lvl_kcap) {
    if (lane == 0) { atomicOr(status + frame, ORB_ST_OUT_OVERFLOW); *out_count = 0; }
    return;
  }
  for (int i = lane; i < nout; i += 32) {
    const int nd = order[i];
    const uint32_t bc = S.n_bc[nd];
    const int begin = bc & 0xffff, count = bc >> 16;
    const uint32_t* ks = S.keys[(S.n_fl[nd] >> 1) & 1] + begin;
    uint32_t best = ks[0];
    for (int k = 1; k < count; ++k) {
      const uint32_t kk = ks[k];
      if (orb_ps(kk) > orb_ps(best)) best = kk;
    }
    out_keys[i] = best;
  }
  if (lane == 0) *out_count = nout;
}

// -------------------------------------------------------------------------------------------------
// Output slot assignment (:1041, :1066-1079): keypoints are visited level by level in list order;
// one whose scaled x lies in [lap0, lap1] goes to the back (stereoIndex--), the others to the front
// (monoIndex++). One CTA per frame performs the ordered scan.
// -------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_assemble(OrbGeom g, const int* __restrict__ sel_count,
                                                  const uint32_t* __restrict__ sel_keys, int lap0, int lap1,
                                                  int* __restrict__ ord_src, int* __restrict__ ord_dst,
                                                  int* __restrict__ n_out, int* __restrict__ mono_out,
                                                  int* __restrict__ status, int* __restrict__ host_n, int* __restrict__ host_mono,
                                                  int* __restrict__ host_status) {
  // host_*: the handle's pinned result words, written straight from here (device-accessible under unified addressing): the frame's
  // keypoint count, monoIndex and capacity status are final once this kernel ends - no device-to-host copies for them
  __shared__ int lvl_off[ORB_MAX_LEVELS + 1];
  __shared__ int warp_sum[8];
  __shared__ int carry;
  const int frame = blockIdx.x, tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  if (tid == 0) {
    int o = 0;
    for (int l = 0; l < g.nlevels; ++l) { lvl_off[l] = o; o += sel_count[(size_t)frame * g.nlevels + l]; }
    lvl_off[g.nlevels] = o;
    carry = 0;
  }
  __syncthreads();
  const int n = lvl_off[g.nlevels];
  if (n > g.kcap) {
    if (tid == 0) {
      const int st0 = atomicOr(status + frame, ORB_ST_OUT_OVERFLOW) | ORB_ST_OUT_OVERFLOW;
      n_out[frame] = 0; mono_out[frame] = 0;
      host_n[frame] = 0; host_mono[frame] = 0; host_status[frame] = st0;
      status[frame] = 0;   // for the next extraction
    }
    return;
  }
  const float flap0 = (float)lap0, flap1 = (float)lap1;
  for (int base = 0; base < n; base += 256) {
    const int ord = base + tid;
    int lap = 0, l = 0, idx = 0;
    uint32_t key = 0;
    if (ord < n) {
      while (ord >= lvl_off[l + 1]) ++l;
      idx = ord - lvl_off[l];
      const uint32_t k = sel_keys[((size_t)frame * g.nlevels + l) * g.lvl_kcap + idx];
      key = k;
      float x = (float)(orb_px(k) + ORB_BORDER);
      if (l != 0) x = __fmul_rn(x, g.scale[l]);
      lap = (x >= flap0 && x <= flap1) ? 1 : 0;
    }
    // block-wide exclusive scan of lap
    const uint32_t b = __ballot_sync(0xffffffffu, lap);
    if (lane == 0) warp_sum[wid] = __popc(b);
    __syncthreads();
    int before = carry;
    for (int w = 0; w < wid; ++w) before += warp_sum[w];
    before += __popc(b & ((1u << lane) - 1u));
    if (ord < n) {
      ord_src[(size_t)frame * g.kcap + ord] = (int)key;
      // lapping keypoint number `before` goes to n-1-before; a mono keypoint to ord - before
      ord_dst[(size_t)frame * g.kcap + ord] = ((lap ? (n - 1 - before) : (ord - before)) << 4) | l;
    }
    __syncthreads();
    if (tid == 0) {
      int s = 0;
      for (int w = 0; w < 8; ++w) s += warp_sum[w];
      carry += s;
    }
    __syncthreads();
  }
  if (tid == 0) {
    n_out[frame] = n; mono_out[frame] = n - carry;
    host_n[frame] = n; host_mono[frame] = n - carry; host_status[frame] = status[frame];
    status[frame] = 0;     // for the next extraction (no memset at the head of the pipeline)
  }
}

// -------------------------------------------------------------------------------------------------
// Orientation + descriptor. One warp per keypoint.
//   IC_Angle (:75-99): first-order moments over the radius-15 disc of the UN-blurred level, then
//   cv::fastAtan2 (OpenCV's degree polynomial, replicated operation by operation).
//   computeOrbDescriptor (:102-145): 256 comparisons of the BLURRED level sampled at the pattern
//   rotated by the angle: a = cosf, b = sinf (glibc's float routines, restated in double exactly),
//   row = cvRound(x*b + y*a), col = cvRound(x*a - y*b) with separate roundings (no FMA).
// Lane i builds descriptor byte i; the 32 bytes leave the warp as two 16-byte stores.
// -------------------------------------------------------------------------------------------------
static __device__ __forceinline__ float dev_fast_atan2(float y, float x) {
  const float rad2deg = 57.29577951308232f;  // (float)(180 / CV_PI)
  const float p1 = __fmul_rn(0.9997878412794807f, rad2deg);
  const float p3 = __fmul_rn(-0.3258083974640975f, rad2deg);
  const float p5 = __fmul_rn(0.1555786518463281f, rad2deg);
  const float p7 = __fmul_rn(-0.04432655554792128f, rad2deg);
  const float eps = 2.220446049250313e-16f;  // (float)DBL_EPSILON
  const float ax = fabsf(x), ay = fabsf(y);
  float a, c, c2;
  if (ax >= ay) {
    c = __fdiv_rn(ay, __fadd_rn(ax, eps));
    c2 = __fmul_rn(c, c);
    a = __fmul_rn(__fadd_rn(__fmul_rn(__fadd_rn(__fmul_rn(__fadd_rn(__fmul_rn(p7, c2), p5), c2), p3), c2), p1), c);
  } else {
    c = __fdiv_rn(ax, __fadd_rn(ay, eps));
    c2 = __fmul_rn(c, c);
    a = __fsub_rn(90.f, __fmul_rn(__fadd_rn(__fmul_rn(__fadd_rn(__fmul_rn(__fadd_rn(__fmul_rn(p7, c2), p5), c2), p3), c2), p1), c));
  }
  if (x < 0) a = __fsub_rn(180.f, a);
  if (y < 0) a = __fsub_rn(360.f, a);
  return a;
}


#include "orb_kernel_describe.cuh"
