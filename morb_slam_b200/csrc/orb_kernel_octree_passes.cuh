// DistributeOctTree (reference src/ORBextractor.cc:540-738) in PASS form, one CTA of 8 warps per (frame, level); the default quad-tree
// kernel since round 2 (ORB_B200_OCTREE=warp selects the one-warp list kernel k_octree; both are parity-tested). The formulation is
// restated and checked on the CPU as oracle/orb_oracle.cc::distribute_octree_passes.
//   * the list is an ARRAY of node records in list order (double-buffered: every pass writes the next array), no links, no free list;
//   * a pass = (1) one warp per divisible node counts its quadrants, (2) warp 0 scans: children of the LAST divided node come
//     first (each group as n4 n3 n2 n1), then the undivided nodes in their old order; record entries in visiting order,
//     (3) one warp per node partitions the keys (stable, into the other key buffer, same segment) and writes the child records at
//     their final positions, (4) survivors are copied behind them;
//   * the final phase sorts the record list like std::sort (the quicksort half by one warp with ballot-built partitions, the insertion-sort half as a
//     block-wide stable rank computation: op_block_std_sort), counts the children of all of its nodes,
//     finds by a prefix sum how many divisions bring the list to N nodes, and performs exactly those.
#pragma once

#ifndef OP_WARPS
#define OP_WARPS 8
#endif
#define OP_THREADS (OP_WARPS * 32)

struct OpSmem {
  uint32_t* keys[2];
  uint32_t* bc[2];   // begin | count << 16
  uint32_t* nx[2];   // UL.x | UR.x << 16
  uint32_t* ny[2];   // UL.y | BR.y << 16
  uint8_t* fl[2];    // bit 0 bNoMore, bit 1 key buffer
  unsigned long long* rec;
  unsigned long long* prev;
  uint32_t* cA;      // children 0 | 1 << 16 of the node at a position (main pass) / of division t (final phase)
  uint32_t* cB;      // children 2 | 3 << 16
  int* pos_child;    // first position of the node's children in the next array
  int* pos_rec;      // first record index of the node's children
  int* pos_surv;     // position of an undivided node in the next array (-1: divided)
  int* tnode;        // final phase: node position of division t
};

static size_t octree_passes_smem_bytes(int node_cap, int smem_keys) {
  return (size_t)node_cap * (2 * (4 + 4 + 4 + 1) + 8 + 8 + 4 + 4 + 4 + 4 + 4 + 4) + 2 * sizeof(uint32_t) * (size_t)smem_keys + 256;
}

static __device__ __forceinline__ int op_quadrant(uint32_t k, int midX, int midY) {
  return (orb_px(k) < midX ? 0 : 1) + (orb_py(k) < midY ? 0 : 2);
}

// libstdc++'s __introsort_loop by ONE WARP (the block-parallel kernel's replacement of dev_introsort_loop, same result).
// __unguarded_partition(first + 1, last, pivot) swaps the k-th record from the left that is not less than the pivot (L_k) with the
// k-th record from the right that is not greater (R_k) for as long as L_k < R_k: the scans never re-read a swapped record (the
// pointers have passed it), so L_k / R_k are properties of the range as it is when the call starts. With nsw such swaps the cut
// (the returned `first`) is min(L_{nsw + 1}, R_{nsw}) - the left scan stops at the next such record or at the record the last swap
// put there (R_0 = last). The warp builds both lists with ballots, counts nsw, swaps in parallel. Median-of-three and the range
// stack are warp-uniform; the heap-sort fallback below the depth limit stays on one lane. Lp / Rp: n ints each.
static __device__ void op_introsort_loop_warp(unsigned long long* a, int n, int* Lp, int* Rp, int lane) {
  if (n <= 16) return;
  int stack_first[40], stack_last[40], stack_depth[40];
  int sp = 0;
  int first = 0, last = n, depth = 2 * (31 - __clz(n));
  const uint32_t lt = (1u << lane) - 1u;
  while (true) {
    while (last - first > 16) {
      if (depth == 0) {
        if (lane == 0) dev_heap_sort(a + first, last - first);
        __syncwarp();
        break;
      }
      --depth;
      const int mid = first + (last - first) / 2;
      const int ia = first + 1, ib = mid, ic = last - 1;
      int pick;
      const unsigned long long va = a[ia], vb = a[ib], vc = a[ic], vf = a[first];
      if (rec_less(va, vb)) {
        if (rec_less(vb, vc)) pick = ib;
        else if (rec_less(va, vc)) pick = ic;
        else pick = ia;
      } else if (rec_less(va, vc)) pick = ia;
      else if (rec_less(vb, vc)) pick = ic;
      else pick = ib;
      const unsigned long long pivot = pick == ia ? va : (pick == ib ? vb : vc);
      __syncwarp();
      if (lane == 0) { a[first] = pivot; a[pick] = vf; }
      __syncwarp();
      int nL = 0, nR = 0;
      for (int base = first + 1; base < last; base += 32) {
        const int p = base + lane;
        const bool f = p < last && !rec_less(a[p], pivot);
        const uint32_t b = __ballot_sync(0xffffffffu, f);
        if (f) Lp[nL + __popc(b & lt)] = p;
        nL += __popc(b);
      }
      for (int base = last - 1; base > first; base -= 32) {
        const int p = base - lane;
        const bool f = p > first && !rec_less(pivot, a[p]);
        const uint32_t b = __ballot_sync(0xffffffffu, f);
        if (f) Rp[nR + __popc(b & lt)] = p;
        nR += __popc(b);
      }
      __syncwarp();
      const int K = min(nL, nR);
      int nsw = 0;
      for (int base = 0; base < K; base += 32) {
        const int k = base + lane;
        const bool f = k < K && Lp[k] < Rp[k];
        const uint32_t b = __ballot_sync(0xffffffffu, f);
        nsw += __popc(b);
        if (b != 0xffffffffu) break;     // L ascends, R descends: the condition holds for a prefix
      }
      for (int k = lane; k < nsw; k += 32) {
        const int pl = Lp[k], pr = Rp[k];
        const unsigned long long t = a[pl]; a[pl] = a[pr]; a[pr] = t;
      }
      int cut = nsw > 0 ? Rp[nsw - 1] : last;
      if (nsw < nL) cut = min(cut, Lp[nsw]);
      __syncwarp();
      stack_first[sp] = cut; stack_last[sp] = last; stack_depth[sp] = depth; ++sp;   // [cut, last) later, [first, cut) now
      last = cut;
    }
    if (sp == 0) break;
    --sp;
    first = stack_first[sp]; last = stack_last[sp]; depth = stack_depth[sp];
  }
}

// std::sort(rec, rec + np, compareNodes) by the whole block: rec -> prev, __introsort_loop on prev by warp 0, then
// __final_insertion_sort, which is a STABLE sort of the loop's result (orb_kernels_extract.cuh), as a rank computation of every
// thread's records among all of them, scattered back into rec. Lp / Rp: scratch of np ints each. Ends with a block barrier.
static __device__ __forceinline__ void op_block_std_sort(unsigned long long* rec, unsigned long long* prev, int np, int* Lp, int* Rp, int tid) {
  for (int i = tid; i < np; i += OP_THREADS) prev[i] = rec[i];
  __syncthreads();
  if (tid < 32) op_introsort_loop_warp(prev, np, Lp, Rp, tid);
  __syncthreads();
  for (int i = tid; i < np; i += OP_THREADS) {
    const unsigned long long v = prev[i];
    const uint32_t key = (uint32_t)(v >> 32);
    int rank = 0;
    for (int j = 0; j < np; ++j) {
      const uint32_t kj = (uint32_t)(prev[j] >> 32);
      rank += (kj < key || (kj == key && j < i)) ? 1 : 0;
    }
    rec[rank] = v;
  }
  __syncthreads();
}

// test entry (orb_debug_std_sort): the block's std::sort on arbitrary (key, payload) records
__global__ void __launch_bounds__(OP_THREADS) k_debug_std_sort(const uint32_t* __restrict__ keys, int n, uint32_t* __restrict__ keys_out,
                                                               uint32_t* __restrict__ payload_out) {
  extern __shared__ __align__(16) unsigned char op_raw[];
  unsigned long long* rec = (unsigned long long*)op_raw;
  unsigned long long* prev = rec + n;
  int* Lp = (int*)(prev + n);
  int* Rp = Lp + n;
  for (int i = threadIdx.x; i < n; i += OP_THREADS) rec[i] = ((unsigned long long)keys[i] << 32) | (uint32_t)i;
  __syncthreads();
  op_block_std_sort(rec, prev, n, Lp, Rp, threadIdx.x);
  for (int i = threadIdx.x; i < n; i += OP_THREADS) { keys_out[i] = (uint32_t)(rec[i] >> 32); payload_out[i] = (uint32_t)rec[i]; }
}

// quadrant counts of one node (warp-cooperative, no key moves)
static __device__ __forceinline__ void op_count(const OpSmem& S, int cur, int node, int lane, int c[4]) {
  const uint32_t bc = S.bc[cur][node], nx = S.nx[cur][node], ny = S.ny[cur][node];
  const int begin = bc & 0xffff, count = bc >> 16;
  const int ulx = nx & 0xffff, urx = nx >> 16, uly = ny & 0xffff, bry = ny >> 16;
  const int midX = ulx + ((urx - ulx + 1) >> 1), midY = uly + ((bry - uly + 1) >> 1);
  const uint32_t* src = S.keys[(S.fl[cur][node] >> 1) & 1] + begin;
  c[0] = c[1] = c[2] = c[3] = 0;
  for (int base = 0; base < count; base += 32) {
    const int i = base + lane;
    const bool valid = i < count;
    const int q = valid ? op_quadrant(src[i], midX, midY) : -1;
#pragma unroll
    for (int j = 0; j < 4; ++j) c[j] += __popc(__ballot_sync(0xffffffffu, q == j));
  }
}

// DivideNode of one node: stable partition into the other key buffer, child records at nxt positions pos .. (n4 first),
// record entries at rec[rbase ..] in n1 .. n4 order
static __device__ __forceinline__ void op_divide(const OpSmem& S, int cur, int node, int lane, const int c[4], int pos, int rbase) {
  const uint32_t bc = S.bc[cur][node], nx = S.nx[cur][node], ny = S.ny[cur][node];
  const int begin = bc & 0xffff, count = bc >> 16;
  const int ulx = nx & 0xffff, urx = nx >> 16, uly = ny & 0xffff, bry = ny >> 16;
  const int midX = ulx + ((urx - ulx + 1) >> 1), midY = uly + ((bry - uly + 1) >> 1);
  const int buf = (S.fl[cur][node] >> 1) & 1;
  const uint32_t* src = S.keys[buf] + begin;
  uint32_t* dst = S.keys[buf ^ 1] + begin;
  const uint32_t lt = (1u << lane) - 1u;
  int run[4] = {0, c[0], c[0] + c[1], c[0] + c[1] + c[2]};
  for (int base = 0; base < count; base += 32) {
    const int i = base + lane;
    const bool valid = i < count;
    const uint32_t k = valid ? src[i] : 0u;
    const int q = valid ? op_quadrant(k, midX, midY) : -1;
    uint32_t b[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) b[j] = __ballot_sync(0xffffffffu, q == j);
    if (valid) dst[run[q] + __popc(b[q] & lt)] = k;
#pragma unroll
    for (int j = 0; j < 4; ++j) run[j] += __popc(b[j]);
  }
  if (lane < 4) {   // lane q writes child q: list order of the group n4 n3 n2 n1, records in n1 .. n4 order
    const int q = lane, nxt = cur ^ 1;
    const int cq = q == 0 ? c[0] : (q == 1 ? c[1] : (q == 2 ? c[2] : c[3]));
    if (cq != 0) {
      const int offq = (q > 0 ? c[0] : 0) + (q > 1 ? c[1] : 0) + (q > 2 ? c[2] : 0);
      const int p = pos + (q < 3 && c[3] != 0) + (q < 2 && c[2] != 0) + (q < 1 && c[1] != 0);
      const int cx0 = (q & 1) ? midX : ulx, cx1 = (q & 1) ? urx : midX;
      const int cy0 = (q & 2) ? midY : uly, cy1 = (q & 2) ? bry : midY;
      S.bc[nxt][p] = (uint32_t)(begin + offq) | ((uint32_t)cq << 16);
      S.nx[nxt][p] = (uint32_t)cx0 | ((uint32_t)cx1 << 16);
      S.ny[nxt][p] = (uint32_t)cy0 | ((uint32_t)cy1 << 16);
      S.fl[nxt][p] = (uint8_t)((cq == 1 ? 1 : 0) | ((buf ^ 1) << 1));
      if (cq > 1) {   // the node id of a record is its position in the next array
        const int r = rbase + (q > 0 && c[0] > 1) + (q > 1 && c[1] > 1) + (q > 2 && c[2] > 1);
        S.rec[r] = ((unsigned long long)(((uint32_t)cq << 16) | (uint32_t)cx0) << 32) | (uint32_t)p;
      }
    }
  }
}

#ifndef OP_MINB
#define OP_MINB 1
#endif
__global__ void __launch_bounds__(OP_THREADS, OP_MINB) k_octree_passes(OrbGeom g, const int* __restrict__ cell_count,
                                                              const uint32_t* __restrict__ cell_keys, int cells_per_frame,
                                                              uint32_t* __restrict__ tree_scratch, int* __restrict__ lvl_count,
                                                              int* __restrict__ sel_count, uint32_t* __restrict__ sel_keys,
                                                              int* __restrict__ status, int l_arg, int NC, int smem_keys,
                                                              const uint32_t* __restrict__ dbg_keys, int dbg_n) {
  extern __shared__ __align__(16) unsigned char op_raw[];
  __shared__ int s_size, s_nrec, s_used, s_state;   // s_state: 0 run main pass, 1 final phase, 2 finished, 3 overflow
  const int frame = blockIdx.x, tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const int l = l_arg >= 0 ? l_arg : (int)blockIdx.y;
  OpSmem S;
  {
    unsigned char* p = op_raw;
    S.rec = (unsigned long long*)p; p += 8 * (size_t)NC;
    S.prev = (unsigned long long*)p; p += 8 * (size_t)NC;
    S.keys[0] = (uint32_t*)p; p += 4 * (size_t)smem_keys;
    S.keys[1] = (uint32_t*)p; p += 4 * (size_t)smem_keys;
    for (int b = 0; b < 2; ++b) {
      S.bc[b] = (uint32_t*)p; p += 4 * (size_t)NC;
      S.nx[b] = (uint32_t*)p; p += 4 * (size_t)NC;
      S.ny[b] = (uint32_t*)p; p += 4 * (size_t)NC;
    }
    S.cA = (uint32_t*)p; p += 4 * (size_t)NC;
    S.cB = (uint32_t*)p; p += 4 * (size_t)NC;
    S.pos_child = (int*)p; p += 4 * (size_t)NC;
    S.pos_rec = (int*)p; p += 4 * (size_t)NC;
    S.pos_surv = (int*)p; p += 4 * (size_t)NC;
    S.tnode = (int*)p; p += 4 * (size_t)NC;
    S.fl[0] = (uint8_t*)p; p += (size_t)NC;
    S.fl[1] = (uint8_t*)p;
  }
  const int N = g.nfeat[l];
  int* out_count = sel_count + (size_t)frame * g.nlevels + l;
  uint32_t* out_keys = sel_keys + ((size_t)frame * g.nlevels + l) * g.lvl_kcap;
  uint32_t* gA = tree_scratch + (size_t)frame * g.scratch_frame + g.scratch_off[l];
  // candidates in reference order, gathered from the FAST kernel's per-cell lists (tree_gather_cells)
  int n;
  if (dbg_keys) {
    n = dbg_n;
    if (n > g.level_cap[l]) n = -1;
  } else {
    const int ncell = g.cell_start[l + 1] - g.cell_start[l];
    int* offs = ncell + 1 <= smem_keys ? (int*)S.keys[1] : (int*)(gA + g.level_cap[l]);
    n = tree_gather_cells<true>(g, l, frame, cell_count, cell_keys, cells_per_frame, S.keys[0], smem_keys, gA, offs, tid, OP_THREADS);
    if (tid == 0) lvl_count[(size_t)frame * g.nlevels + l] = max(n, 0);
  }
  if (n < 0) {
    if (tid == 0) { atomicOr(status + frame, ORB_ST_LEVEL_OVERFLOW); *out_count = 0; }
    return;
  }
  if (n == 0) {
    if (tid == 0) *out_count = 0;
    return;
  }
  if (n > smem_keys) {
    S.keys[0] = gA;
    S.keys[1] = gA + g.level_cap[l];
  }
  if (dbg_keys)
    for (int i = tid; i < n; i += OP_THREADS) S.keys[0][i] = dbg_keys[i];
  __syncthreads();

  // ---- roots (:545-582) by warp 0: key -> root (int)(pt.x / hX), stable; records at positions 0 .. in root order
  int cur = 0;
  if (wid == 0) {
    const int nIni = g.nini[l];
    const float hX = g.hx[l];
    const int regH = g.h[l] - 2 * ORB_BORDER;
    const uint32_t lt = (1u << lane) - 1u;
    int wpos = 0, nroot = 0;
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
      if (lane == 0) {
        const int ulx = (int)__fmul_rn(hX, (float)r), urx = (int)__fmul_rn(hX, (float)(r + 1));
        S.bc[0][nroot] = (uint32_t)rbegin | ((uint32_t)cnt << 16);
        S.nx[0][nroot] = (uint32_t)ulx | ((uint32_t)urx << 16);
        S.ny[0][nroot] = 0u | ((uint32_t)regH << 16);
        S.fl[0][nroot] = (uint8_t)((cnt == 1 ? 1 : 0) | (1 << 1));
      }
      ++nroot;
    }
    if (lane == 0) { s_size = nroot; s_nrec = 0; s_state = 0; }
  }
  __syncthreads();

  while (true) {
    const int state = s_state;
    if (state >= 2) break;
    const int size = s_size;
    int ndiv;                       // nodes this round may divide: all positions (main pass) or the sorted records (final phase)
    if (state == 0) {
      ndiv = size;
      for (int i = tid; i < size; i += OP_THREADS) S.tnode[i] = i;
    } else {
      // std::sort of the record list (:665-667): __introsort_loop by warp 0 (parallel partitions), then __final_insertion_sort as a
      // block-wide stable rank computation into S.rec, whose old content is dead until the divisions of this round write the next records
      const int np = s_nrec;
      op_block_std_sort(S.rec, S.prev, np, S.pos_child, S.pos_rec, tid);   // the position arrays are dead until the placement below
      ndiv = np;
      for (int t = tid; t < np; t += OP_THREADS) S.tnode[t] = (int)(uint32_t)(S.rec[np - 1 - t] & 0xffffffffull);   // t = 0 is divided first
    }
    __syncthreads();
    // (1) quadrant counts, one warp per candidate division
    for (int t = wid; t < ndiv; t += OP_WARPS) {
      const int node = S.tnode[t];
      int c[4] = {0, 0, 0, 0};
      if (!(S.fl[cur][node] & 1)) op_count(S, cur, node, lane, c);
      if (lane == 0) { S.cA[t] = (uint32_t)c[0] | ((uint32_t)c[1] << 16); S.cB[t] = (uint32_t)c[2] | ((uint32_t)c[3] << 16); }
    }
    __syncthreads();
    // (2) placement by warp 0 (sequential over chunks of 32 with warp scans)
    if (wid == 0) {
      // how many divisions happen: all divisible ones (main pass) / the first `used` that bring the list to N (final phase)
      int used = ndiv;
      if (state == 1) {
        int sz = size;
        used = ndiv;
        for (int base = 0; base < ndiv && used == ndiv; base += 32) {
          const int t = base + lane;
          int d = 0;
          if (t < ndiv) {
            const uint32_t a = S.cA[t], b = S.cB[t];
            d = ((a & 0xffff) != 0) + ((a >> 16) != 0) + ((b & 0xffff) != 0) + ((b >> 16) != 0) - 1;
          }
          int incl = d;
#pragma unroll
          for (int o = 1; o < 32; o <<= 1) { const int v = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += v; }
          const uint32_t hit = __ballot_sync(0xffffffffu, t < ndiv && sz + incl >= N);
          if (hit) used = base + __ffs(hit);           // the first division after which the list holds N nodes, inclusive
          sz += __shfl_sync(0xffffffffu, incl, 31);
        }
      }
      // total number of children of the divisions that happen, and per division the inclusive prefix
      int run_child = 0, run_rec = 0;
      for (int base = 0; base < used; base += 32) {
        const int t = base + lane;
        int nch = 0, nrc = 0;
        if (t < used) {
          const uint32_t a = S.cA[t], b = S.cB[t];
          const int c0 = a & 0xffff, c1 = a >> 16, c2 = b & 0xffff, c3 = b >> 16;
          nch = (c0 != 0) + (c1 != 0) + (c2 != 0) + (c3 != 0);
          nrc = (c0 > 1) + (c1 > 1) + (c2 > 1) + (c3 > 1);
        }
        int ic = nch, ir = nrc;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
          const int v = __shfl_up_sync(0xffffffffu, ic, o), w = __shfl_up_sync(0xffffffffu, ir, o);
          if (lane >= o) { ic += v; ir += w; }
        }
        if (t < used) { S.pos_child[t] = run_child + ic; S.pos_rec[t] = run_rec + ir - nrc; }   // pos_child: inclusive prefix for now
        run_child += __shfl_sync(0xffffffffu, ic, 31);
        run_rec += __shfl_sync(0xffffffffu, ir, 31);
      }
      // children of division t start at (total - inclusive prefix): the last division comes first
      for (int t = lane; t < used; t += 32) S.pos_child[t] = run_child - S.pos_child[t];
      // survivors: every position that is not divided, in the old order, behind the children
      for (int i = lane; i < size; i += 32) S.pos_surv[i] = 0;
      __syncwarp();
      for (int t = lane; t < used; t += 32) {
        const uint32_t a = S.cA[t], b = S.cB[t];
        if ((a | b) != 0) S.pos_surv[S.tnode[t]] = -1;       // a node with keys always has a child: divided
      }
      __syncwarp();
      int run_s = run_child;
      for (int base = 0; base < size; base += 32) {
        const int i = base + lane;
        const bool surv = i < size && S.pos_surv[i] == 0;
        const uint32_t b = __ballot_sync(0xffffffffu, surv);
        if (surv) S.pos_surv[i] = run_s + __popc(b & ((1u << lane) - 1u));
        run_s += __popc(b);
      }
      if (lane == 0) { s_used = used; s_nrec = run_rec; if (run_s > NC) s_state = 3; else s_size = run_s; }
    }
    __syncthreads();
    if (s_state == 3) break;
    const int used = s_used;
    // (3) the divisions, one warp per node, and (4) the survivors
    for (int t = wid; t < used; t += OP_WARPS) {
      const uint32_t a = S.cA[t], b = S.cB[t];
      if ((a | b) == 0) continue;                       // bNoMore node of a main pass: nothing to divide
      const int c[4] = {(int)(a & 0xffff), (int)(a >> 16), (int)(b & 0xffff), (int)(b >> 16)};
      op_divide(S, cur, S.tnode[t], lane, c, S.pos_child[t], S.pos_rec[t]);
    }
    for (int i = tid; i < size; i += OP_THREADS) {
      const int p = S.pos_surv[i];
      if (p >= 0) {
        S.bc[cur ^ 1][p] = S.bc[cur][i]; S.nx[cur ^ 1][p] = S.nx[cur][i]; S.ny[cur ^ 1][p] = S.ny[cur][i]; S.fl[cur ^ 1][p] = S.fl[cur][i];
      }
    }
    __syncthreads();
    cur ^= 1;
    if (tid == 0) {
      const int nsize = s_size, nrec = s_nrec;
      if (nsize >= N || nsize == size) s_state = 2;                       // :654 / :711
      else if (state == 0 && nsize + nrec * 3 > N) s_state = 1;           // :659 (nToExpand = records of this pass)
    }
    __syncthreads();
  }
  if (s_state == 3) {
    if (tid == 0) { atomicOr(status + frame, ORB_ST_NODE_OVERFLOW); *out_count = 0; }
    return;
  }
  // ---- best response per leaf, first maximum wins, list order (:718-735)
  const int nout = s_size;
  if (nout > g.lvl_kcap) {
    if (tid == 0) { atomicOr(status + frame, ORB_ST_OUT_OVERFLOW); *out_count = 0; }
    return;
  }
  for (int i = tid; i < nout; i += OP_THREADS) {
    const uint32_t bc = S.bc[cur][i];
    const int begin = bc & 0xffff, count = bc >> 16;
    const uint32_t* ks = S.keys[(S.fl[cur][i] >> 1) & 1] + begin;
    uint32_t best = ks[0];
    for (int k = 1; k < count; ++k) {
      const uint32_t kk = ks[k];
      if (orb_ps(kk) > orb_ps(best)) best = kk;
    }
    out_keys[i] = best;
  }
  if (tid == 0) *out_count = nout;
}
