// liborb_b200.so - batched brute-force top-2 Hamming kNN.
// Replaces cv::BFMatcher(NORM_HAMMING).knnMatch(q, db, k = 2) as used by
// Frame::ComputeStereoFishEyeMatches (reference src/Frame.cc:46, :1242) and the scalar best/second-best
// idiom of ORBmatcher (src/ORBmatcher.cc:98-115); tie rule: the lower database index wins
// (ascending scan with strict <).
//
//   k_knn2_scan   grid (query tiles, db chunks). A thread keeps QPT query descriptors (8 x u32 each) in
//                 registers; database rows stream through a double-buffered shared-memory tile filled with
//                 16-byte cp.async copies and are read back as warp-wide broadcast uint4 loads; the
//                 distance is 8 XOR + one carry-save level + 6 POPC per pair; each thread keeps a running (best,
//                 second) per query. Bound by the INT/POPC issue rate, not by HBM: the database is read once
//                 per query tile and stays L2-resident.
//   k_knn2_merge  per query, lexicographic (distance, index) top-2 over the partial lists of all chunks /
//                 all ranks.
#include <algorithm>
#include <cstdlib>
#include <cstring>

#include "orb_internal.h"

#define KNN_QPT_MAX 6
#define KNN_TILE_ROWS 256
#define KNN_NONE 0xffffffffffffffffull

static __device__ __forceinline__ void cp_async16(void* smem, const void* gmem) {
  unsigned s = (unsigned)__cvta_generic_to_shared(smem);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(s), "l"(gmem));
}
static __device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N>
static __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N)); }

// 256-bit Hamming distance with one level of carry-save adders: six of the eight XOR words are compressed bit-wise
// (x0+x1+x2 = s + 2c per bit position) so that six population counts remain:
//   sum_i popc(x_i) = popc(s0) + popc(s1) + popc(x6) + popc(x7) + 2 * (popc(c0) + popc(c1)).
// POPC issues at 16 lanes/clk/SM (XU pipe) while LOP3 issues at 64, so trading 2 POPC for 4 LOP3 lifts the
// kernel above the plain "8 POPC per pair" bound without overloading the ALU pipe. Exact integer arithmetic, same result as the reference's
// per-word popcount (src/ORBmatcher.cc:1880-1894).
static __device__ __forceinline__ void csa(uint32_t a, uint32_t b, uint32_t c, uint32_t& s, uint32_t& cy) {
  s = a ^ b ^ c;
  cy = (a & b) | (a & c) | (b & c);
}
static __device__ __forceinline__ uint32_t hamming256_csa(const uint4 a, const uint4 b, const uint32_t* q) {
  const uint32_t x0 = a.x ^ q[0], x1 = a.y ^ q[1], x2 = a.z ^ q[2], x3 = a.w ^ q[3];
  const uint32_t x4 = b.x ^ q[4], x5 = b.y ^ q[5], x6 = b.z ^ q[6], x7 = b.w ^ q[7];
  // one tree level: 6 POPC + 4 LOP3. POPC issues at 16 lanes/clk/SM (XU pipe), LOP3 at 64 (ALU pipe), and the ALU pipe also carries
  // the XORs and the top-2 update, so the split between the two pipes is measured, not derived (1200 x 1.25 M pairs, B200):
  // 8 POPC / no tree 2.96 ms (XU-bound), 7 / 2 LOP3 2.64 ms, 6 / 4 2.52 ms, 5 / 6 2.82 ms, 4 / 8 2.91 ms (ALU pipe 95 %).
  uint32_t s0, c0, s1, c1;
  csa(x0, x1, x2, s0, c0);
  csa(x3, x4, x5, s1, c1);
  return __popc(s0) + __popc(s1) + __popc(x6) + __popc(x7) + 2u * (__popc(c0) + __popc(c1));
}

// partial[chunk][query][2] packed keys: dist << 32 | global index
// Block sizes are multiples of 4 warps (128 or 256 threads) so that every SM sub-partition carries the same
// number of warps of a block: with 5 warps per block one scheduler held two of them and the other warps idled
// at the per-tile barrier for 40 % of the time (profiles/README_r1.md).
template <int KNN_QPT>
__global__ void __launch_bounds__(256) k_knn2_scan(const uint8_t* __restrict__ q, int nq, int q_per_tile,
                                                   const uint8_t* __restrict__ db, long long ndb, long long rows_per_chunk,
                                                   int index_base, unsigned long long* __restrict__ partial) {
  __shared__ __align__(16) uint4 tile[2][KNN_TILE_ROWS * 2];
  const int tid = threadIdx.x, nthr = blockDim.x;
  const int qbase = blockIdx.x * q_per_tile;
  const long long r0 = (long long)blockIdx.y * rows_per_chunk;
  const long long r1 = min(r0 + rows_per_chunk, ndb);
  // queries of this thread: qbase + tid + j * nthr (coalesced-ish loads, any nq)
  uint32_t Q[KNN_QPT][8];
  bool qv[KNN_QPT];
#pragma unroll
  for (int j = 0; j < KNN_QPT; ++j) {
    const int qi = qbase + tid + j * nthr;
    qv[j] = qi < nq && (tid + j * nthr) < q_per_tile;
    if (qv[j]) {
      const uint4* p = reinterpret_cast<const uint4*>(q + (size_t)qi * 32);
      const uint4 a = p[0], b = p[1];
      Q[j][0] = a.x; Q[j][1] = a.y; Q[j][2] = a.z; Q[j][3] = a.w;
      Q[j][4] = b.x; Q[j][5] = b.y; Q[j][6] = b.z; Q[j][7] = b.w;
    } else {
#pragma unroll
      for (int k = 0; k < 8; ++k) Q[j][k] = 0;
    }
  }
  uint32_t d0[KNN_QPT], d1[KNN_QPT], i0[KNN_QPT], i1[KNN_QPT];
#pragma unroll
  for (int j = 0; j < KNN_QPT; ++j) { d0[j] = d1[j] = 0xffffffffu; i0[j] = i1[j] = 0xffffffffu; }

  const long long nrows = r1 - r0;
  const int ntiles = (int)((nrows + KNN_TILE_ROWS - 1) / KNN_TILE_ROWS);
  auto issue = [&](int t, int buf) {
    const long long base = r0 + (long long)t * KNN_TILE_ROWS;
    const int rows = (int)min((long long)KNN_TILE_ROWS, r1 - base);
    const uint4* src = reinterpret_cast<const uint4*>(db + (size_t)base * 32);
    for (int i = tid; i < rows * 2; i += nthr) cp_async16(&tile[buf][i], src + i);
    cp_async_commit();
  };
  if (ntiles > 0) issue(0, 0);
  for (int t = 0; t < ntiles; ++t) {
    const int buf = t & 1;
    if (t + 1 < ntiles) { issue(t + 1, buf ^ 1); cp_async_wait<1>(); }
    else cp_async_wait<0>();
    __syncthreads();
    const long long base = r0 + (long long)t * KNN_TILE_ROWS;
    const int rows = (int)min((long long)KNN_TILE_ROWS, r1 - base);
    const uint32_t gidx0 = (uint32_t)(index_base + base);
#pragma unroll 2
    for (int r = 0; r < rows; ++r) {
      const uint4 a = tile[buf][2 * r], b = tile[buf][2 * r + 1];
#pragma unroll
      for (int j = 0; j < KNN_QPT; ++j) {
        const uint32_t d = hamming256_csa(a, b, Q[j]);
        if (d < d1[j]) {  // rare after warm-up
          if (d < d0[j]) { d1[j] = d0[j]; i1[j] = i0[j]; d0[j] = d; i0[j] = gidx0 + r; }
          else { d1[j] = d; i1[j] = gidx0 + r; }
        }
      }
    }
    __syncthreads();
  }
#pragma unroll
  for (int j = 0; j < KNN_QPT; ++j) {
    if (!qv[j]) continue;
    const int qi = qbase + tid + j * nthr;
    unsigned long long* o = partial + ((size_t)blockIdx.y * nq + qi) * 2;
    o[0] = d0[j] == 0xffffffffu ? KNN_NONE : (((unsigned long long)d0[j] << 32) | i0[j]);
    o[1] = d1[j] == 0xffffffffu ? KNN_NONE : (((unsigned long long)d1[j] << 32) | i1[j]);
  }
}

// ---- Frame::ComputeStereoFishEyeMatches (reference src/Frame.cc:1222-1250), the part before the triangulation, for a whole
// batch: per frame, BFmatcher.knnMatch(mDescriptors.rowRange(monoLeft, N), mDescriptorsRight.rowRange(monoRight, Nright), 2) and
// Lowe's ratio `m[0].distance < m[1].distance * 0.7` (float * double: evaluated in double) on the device-resident descriptors
// of the two handles. Grid (query tiles, frames): a thread keeps FE_QPT queries in registers, the frame's right descriptors
// stream through the same double-buffered cp.async tile as k_knn2_scan; ascending scan with strict "<" keeps the lower train index
// on ties like cv::BFMatcher.
#define FE_QPT 3
#define FE_SMALL_BATCH 8    // batches up to this size split the right keypoints over blockIdx.z
#define FE_CHUNK_ROWS 96
#ifndef FE_THREADS
#define FE_THREADS 128   // 4 blocks of 384 queries per TUM-VI frame: 1024 blocks per 256 frames spread evenly over the 148 SMs
#endif
__global__ void __launch_bounds__(FE_THREADS) k_fisheye_knn2(const uint8_t* __restrict__ descL, const int* __restrict__ nL, const int* __restrict__ monoL,
                                                      int kcapL, const uint8_t* __restrict__ descR, const int* __restrict__ nR,
                                                      const int* __restrict__ monoR, int kcapR, int out_cap, int32_t* __restrict__ idx_out,
                                                      int32_t* __restrict__ dist_out, uint8_t* __restrict__ pass_out, int chunk_rows,
                                                      unsigned long long* __restrict__ partial) {
  // chunk_rows > 0 (small batches, a single pair's latency path): blockIdx.z scans right rows [z * chunk_rows, (z + 1) * chunk_rows) and
  // leaves its top-2 as keys (distance << 32 | index) in partial[z][frame][query][2]; k_fisheye_merge combines the chunks
  __shared__ __align__(16) uint4 tile[2][KNN_TILE_ROWS * 2];
  const int frame = blockIdx.y, tid = threadIdx.x;
  const int q0 = max(monoL[frame], 0), nq = max(min(nL[frame], kcapL) - q0, 0);
  const int t0 = max(monoR[frame], 0), nt = max(min(nR[frame], kcapR) - t0, 0);
  const int qbase = blockIdx.x * FE_THREADS * FE_QPT;
  if (qbase >= nq) return;                                   // whole block
  const int r_begin = chunk_rows > 0 ? (int)blockIdx.z * chunk_rows : 0;
  const int r_end = chunk_rows > 0 ? min(nt, r_begin + chunk_rows) : nt;
  if (chunk_rows > 0 && r_begin >= nt) return;               // whole block: the merge only reads the chunks that hold rows
  const uint8_t* qd = descL + ((size_t)frame * kcapL + q0) * 32;
  const uint8_t* db = descR + ((size_t)frame * kcapR + t0) * 32;
  uint32_t Q[FE_QPT][8];
  bool qv[FE_QPT];
#pragma unroll
  for (int j = 0; j < FE_QPT; ++j) {
    const int qi = qbase + tid + j * FE_THREADS;
    qv[j] = qi < nq;
    const uint4* p = reinterpret_cast<const uint4*>(qd + (size_t)(qv[j] ? qi : 0) * 32);
    const uint4 a = p[0], b = p[1];
    Q[j][0] = a.x; Q[j][1] = a.y; Q[j][2] = a.z; Q[j][3] = a.w; Q[j][4] = b.x; Q[j][5] = b.y; Q[j][6] = b.z; Q[j][7] = b.w;
  }
  uint32_t d0[FE_QPT], d1[FE_QPT], i0[FE_QPT], i1[FE_QPT];
#pragma unroll
  for (int j = 0; j < FE_QPT; ++j) { d0[j] = d1[j] = 0xffffffffu; i0[j] = i1[j] = 0xffffffffu; }
  const int ntiles = (r_end - r_begin + KNN_TILE_ROWS - 1) / KNN_TILE_ROWS;
  auto issue = [&](int t, int buf) {
    const int base = r_begin + t * KNN_TILE_ROWS, rows = min(KNN_TILE_ROWS, r_end - base);
    const uint4* src = reinterpret_cast<const uint4*>(db + (size_t)base * 32);
    for (int i = tid; i < rows * 2; i += FE_THREADS) cp_async16(&tile[buf][i], src + i);
    cp_async_commit();
  };
  if (ntiles > 0) issue(0, 0);
  for (int t = 0; t < ntiles; ++t) {
    const int buf = t & 1;
    if (t + 1 < ntiles) { issue(t + 1, buf ^ 1); cp_async_wait<1>(); }
    else cp_async_wait<0>();
    __syncthreads();
    const int base = r_begin + t * KNN_TILE_ROWS, rows = min(KNN_TILE_ROWS, r_end - base);
#pragma unroll 2
    for (int r = 0; r < rows; ++r) {
      const uint4 a = tile[buf][2 * r], b = tile[buf][2 * r + 1];
#pragma unroll
      for (int j = 0; j < FE_QPT; ++j) {
        const uint32_t d = hamming256_csa(a, b, Q[j]);
        if (d < d1[j]) {
          if (d < d0[j]) { d1[j] = d0[j]; i1[j] = i0[j]; d0[j] = d; i0[j] = (uint32_t)(base + r); }
          else { d1[j] = d; i1[j] = (uint32_t)(base + r); }
        }
      }
    }
    __syncthreads();
  }
#pragma unroll
  for (int j = 0; j < FE_QPT; ++j) {
    const int qi = qbase + tid + j * FE_THREADS;
    if (!qv[j] || qi >= out_cap) continue;
    const size_t o = (size_t)frame * out_cap + qi;
    if (partial) {
      unsigned long long* pk = partial + (((size_t)blockIdx.z * gridDim.y + frame) * out_cap + qi) * 2;
      pk[0] = d0[j] == 0xffffffffu ? ~0ull : (((unsigned long long)d0[j] << 32) | i0[j]);
      pk[1] = d1[j] == 0xffffffffu ? ~0ull : (((unsigned long long)d1[j] << 32) | i1[j]);
      continue;
    }
    idx_out[2 * o] = d0[j] == 0xffffffffu ? -1 : (int)i0[j];
    idx_out[2 * o + 1] = d1[j] == 0xffffffffu ? -1 : (int)i1[j];
    dist_out[2 * o] = d0[j] == 0xffffffffu ? -1 : (int)d0[j];
    dist_out[2 * o + 1] = d1[j] == 0xffffffffu ? -1 : (int)d1[j];
    // (*it).size() >= 2 && (*it)[0].distance < (*it)[1].distance * 0.7 (:1249-1250)
    pass_out[o] = (d1[j] != 0xffffffffu && (double)(float)d0[j] < (double)(float)d1[j] * 0.7) ? 1 : 0;
  }
}

// the chunks' top-2 lists of a query merged by (distance, index): the same two rows an ascending scan with strict "<" keeps
__global__ void __launch_bounds__(128) k_fisheye_merge(const unsigned long long* __restrict__ partial, int chunk_rows, const int* __restrict__ nL,
                                                       const int* __restrict__ monoL, int kcapL, const int* __restrict__ nR,
                                                       const int* __restrict__ monoR, int kcapR, int out_cap, int32_t* __restrict__ idx_out,
                                                       int32_t* __restrict__ dist_out, uint8_t* __restrict__ pass_out) {
  const int frame = blockIdx.y, qi = blockIdx.x * 128 + threadIdx.x;
  const int q0 = max(monoL[frame], 0), nq = max(min(nL[frame], kcapL) - q0, 0);
  const int t0 = max(monoR[frame], 0), nt = max(min(nR[frame], kcapR) - t0, 0);
  if (qi >= out_cap) return;
  if (qi >= nq) {   // entries behind the frame's queries: the defaults (this launch replaces the memsets of the large-batch path)
    const size_t o = (size_t)frame * out_cap + qi;
    idx_out[2 * o] = -1; idx_out[2 * o + 1] = -1; dist_out[2 * o] = -1; dist_out[2 * o + 1] = -1; pass_out[o] = 0;
    return;
  }
  const int nch = (nt + chunk_rows - 1) / chunk_rows;
  unsigned long long k0 = ~0ull, k1 = ~0ull;
  for (int c = 0; c < nch; ++c) {
    const unsigned long long* pk = partial + (((size_t)c * gridDim.y + frame) * out_cap + qi) * 2;
    const unsigned long long a = pk[0], b = pk[1];
    if (a < k0) { k1 = min(k0, b); k0 = a; }
    else if (a < k1) k1 = a;
  }
  const size_t o = (size_t)frame * out_cap + qi;
  const uint32_t d0 = (uint32_t)(k0 >> 32), d1 = (uint32_t)(k1 >> 32);
  idx_out[2 * o] = k0 == ~0ull ? -1 : (int)(uint32_t)k0;
  idx_out[2 * o + 1] = k1 == ~0ull ? -1 : (int)(uint32_t)k1;
  dist_out[2 * o] = k0 == ~0ull ? -1 : (int)d0;
  dist_out[2 * o + 1] = k1 == ~0ull ? -1 : (int)d1;
  pass_out[o] = (k1 != ~0ull && (double)(float)d0 < (double)(float)d1 * 0.7) ? 1 : 0;
}

// partial lists as packed keys: [part][query][2]
__global__ void k_knn2_merge_keys(const unsigned long long* __restrict__ partial, int nparts, int nq,
                                  int32_t* __restrict__ idx_out, int32_t* __restrict__ dist_out) {
  const int qi = blockIdx.x * blockDim.x + threadIdx.x;
  if (qi >= nq) return;
  unsigned long long b0 = KNN_NONE, b1 = KNN_NONE;
  for (int p = 0; p < nparts; ++p) {
    const unsigned long long* o = partial + ((size_t)p * nq + qi) * 2;
#pragma unroll
    for (int k = 0; k < 2; ++k) {
      const unsigned long long v = o[k];
      if (v < b0) { b1 = b0; b0 = v; }
      else if (v < b1) b1 = v;
    }
  }
  idx_out[2 * qi] = b0 == KNN_NONE ? -1 : (int32_t)(uint32_t)(b0 & 0xffffffffu);
  dist_out[2 * qi] = b0 == KNN_NONE ? -1 : (int32_t)(b0 >> 32);
  idx_out[2 * qi + 1] = b1 == KNN_NONE ? -1 : (int32_t)(uint32_t)(b1 & 0xffffffffu);
  dist_out[2 * qi + 1] = b1 == KNN_NONE ? -1 : (int32_t)(b1 >> 32);
}

// partial lists as (idx, dist) int32 arrays gathered from several ranks: [part][query][2]
__global__ void k_knn2_merge_parts(const int32_t* __restrict__ idx_parts, const int32_t* __restrict__ dist_parts, int nparts,
                                   int nq, int32_t* __restrict__ idx_out, int32_t* __restrict__ dist_out) {
  const int qi = blockIdx.x * blockDim.x + threadIdx.x;
  if (qi >= nq) return;
  unsigned long long b0 = KNN_NONE, b1 = KNN_NONE;
  for (int p = 0; p < nparts; ++p) {
#pragma unroll
    for (int k = 0; k < 2; ++k) {
      const size_t o = ((size_t)p * nq + qi) * 2 + k;
      const int32_t id = idx_parts[o], di = dist_parts[o];
      if (id < 0 || di < 0) continue;
      const unsigned long long v = ((unsigned long long)(uint32_t)di << 32) | (uint32_t)id;
      if (v < b0) { b1 = b0; b0 = v; }
      else if (v < b1) b1 = v;
    }
  }
  idx_out[2 * qi] = b0 == KNN_NONE ? -1 : (int32_t)(uint32_t)(b0 & 0xffffffffu);
  dist_out[2 * qi] = b0 == KNN_NONE ? -1 : (int32_t)(b0 >> 32);
  idx_out[2 * qi + 1] = b1 == KNN_NONE ? -1 : (int32_t)(uint32_t)(b1 & 0xffffffffu);
  dist_out[2 * qi + 1] = b1 == KNN_NONE ? -1 : (int32_t)(b1 >> 32);
}

// Lowe ratio gate (src/Frame.cc:1250): float distance compared against float * double(0.7), in double
__global__ void k_ratio_test(const int32_t* __restrict__ dist, int nq, uint8_t* __restrict__ pass) {
  const int qi = blockIdx.x * blockDim.x + threadIdx.x;
  if (qi >= nq) return;
  const int32_t a = dist[2 * qi], b = dist[2 * qi + 1];
  bool ok = false;
  if (a >= 0 && b >= 0) ok = (double)(float)a < __dmul_rn((double)(float)b, 0.7);
  pass[qi] = ok ? 1 : 0;
}

// Launches the scan of one (device-resident) database shard: picks the tiling, makes room for the partial lists at the start of
// d_scratch (+ `extra_bytes` behind them, 256-byte aligned: *extra_out) and runs k_knn2_scan. partial[chunk][query][2] keys.
static int knn2_scan_to_parts(orb_handle* h, const uint8_t* d_q, int nq, const uint8_t* d_db, int64_t ndb, int32_t index_base, size_t extra_bytes,
                              unsigned long long** part_out, int* nchunks_out, uint8_t** extra_out) {
  int st;
  // tiling: pick (query tiles, threads in {128, 256}, queries per thread in 2..6) with the least idle query slots
  int qtiles = 1, threads = 128, qpt = 2, q_per_tile = nq;
  {
    double best = 1e30;
    for (int tiles = 1; tiles <= std::max(1, (nq + 255) / 256); ++tiles) {
      const int per = (nq + tiles - 1) / tiles;
      for (int thr = 128; thr <= 256; thr += 128)
        for (int k = 2; k <= KNN_QPT_MAX; ++k) {
          if (thr * k < per) continue;
          // cost ~ slots processed per database row, with a mild preference for more queries per thread
          const double cost = (double)tiles * thr * k * (1.0 + 0.25 / k);
          if (cost < best) { best = cost; qtiles = tiles; threads = thr; qpt = k; q_per_tile = per; }
        }
    }
  }
  int sms = 148;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, h->device);
  // one full wave of resident blocks: database chunks = resident blocks per SM x SMs / query tiles
  int occ = 4;
  switch (qpt) {
    case 2: cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_knn2_scan<2>, threads, 0); break;
    case 3: cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_knn2_scan<3>, threads, 0); break;
    case 4: cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_knn2_scan<4>, threads, 0); break;
    case 5: cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_knn2_scan<5>, threads, 0); break;
    default: cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_knn2_scan<6>, threads, 0); break;
  }
  occ = std::max(occ, 1);
  long long want_chunks = std::max(1LL, (long long)occ * sms / qtiles);
  long long rows_per_chunk = std::max((long long)KNN_TILE_ROWS * 4, (long long)((ndb + want_chunks - 1) / want_chunks));
  rows_per_chunk = (rows_per_chunk + 31) / 32 * 32;
  const int nchunks = (int)std::max(1LL, (long long)((ndb + rows_per_chunk - 1) / rows_per_chunk));
  const size_t part_bytes = (size_t)nchunks * nq * 2 * sizeof(unsigned long long);
  if ((st = orb_ensure(h, h->d_scratch, part_bytes + extra_bytes + 512))) return st;
  unsigned long long* d_part = h->d_scratch.as<unsigned long long>();
  if (ndb == 0) {
    ORB_CUDA_CHECK(h, cudaMemsetAsync(d_part, 0xff, part_bytes, h->stream));
  } else {
    const dim3 grd(qtiles, nchunks);
#define KNN_LAUNCH(K) k_knn2_scan<K><<<grd, threads, 0, h->stream>>>(d_q, nq, q_per_tile, d_db, (long long)ndb, rows_per_chunk, index_base, d_part)
    switch (qpt) {
      case 2: KNN_LAUNCH(2); break;
      case 3: KNN_LAUNCH(3); break;
      case 4: KNN_LAUNCH(4); break;
      case 5: KNN_LAUNCH(5); break;
      default: KNN_LAUNCH(6); break;
    }
#undef KNN_LAUNCH
    h->launches++;
    ORB_CUDA_CHECK(h, cudaGetLastError());
  }
  *part_out = d_part;
  *nchunks_out = nchunks;
  if (extra_out) *extra_out = h->d_scratch.as<uint8_t>() + (part_bytes + 255) / 256 * 256;
  return ORB_OK;
}

// ---- sharded search with the exchange fused over peer memory (include/orb_b200.h: orb_knn_exchange_*) ----
// Every rank owns one exchange buffer that all ranks can write (same-process pointers or CUDA IPC mappings over NVLink):
//   keys[2 slots][world][max_nq][2] (dist << 32 | global index), flags[world], block counter, status.
// k_knn2_merge_push: thread per query, merges the chunk lists of the local scan and STORES the rank's top-2 straight into every
//   rank's buffer (16-byte peer stores); the last block to finish publishes the epoch in flags[rank] of every rank (fence + store).
// k_knn2_merge_wait: waits until all ranks' flags reached the epoch (bounded spin), then merges the `world` lists of its own buffer
//   by (distance, index). No collective library call, no host synchronisation between the scan and the result.
struct KnnPeers {
  unsigned long long* keys[ORB_KNN_MAX_RANKS];
  unsigned int* flags[ORB_KNN_MAX_RANKS];
};

__global__ void __launch_bounds__(128) k_knn2_merge_push(const unsigned long long* __restrict__ partial, int nparts, int nq, KnnPeers peers, int rank,
                                                        int world, int max_nq, int slot, unsigned int epoch, unsigned int* counter) {
  const int qi = blockIdx.x * blockDim.x + threadIdx.x;
  if (qi < nq) {
    unsigned long long b0 = KNN_NONE, b1 = KNN_NONE;
    for (int p = 0; p < nparts; ++p) {
      const unsigned long long* o = partial + ((size_t)p * nq + qi) * 2;
#pragma unroll
      for (int k = 0; k < 2; ++k) {
        const unsigned long long v = o[k];
        if (v < b0) { b1 = b0; b0 = v; }
        else if (v < b1) b1 = v;
      }
    }
    const size_t off = (((size_t)slot * world + rank) * max_nq + qi) * 2;
    for (int p = 0; p < world; ++p) {
      const int peer = (rank + p) % world;   // own buffer first, then staggered so that the ranks do not all hit one GPU at once
      *reinterpret_cast<ulonglong2*>(peers.keys[peer] + off) = make_ulonglong2(b0, b1);
    }
  }
  __threadfence_system();
  __syncthreads();
  if (threadIdx.x == 0) {
    const unsigned int t = atomicAdd(counter, 1u);
    if (t == gridDim.x - 1) {
      *counter = 0;
      __threadfence_system();
      for (int p = 0; p < world; ++p) {
        volatile unsigned int* f = peers.flags[(rank + p) % world] + rank;
        *f = epoch;
      }
      __threadfence_system();
    }
  }
}

__global__ void __launch_bounds__(128) k_knn2_merge_wait(const unsigned long long* __restrict__ keys, const unsigned int* flags, int world, int nq,
                                                        int max_nq, int slot, unsigned int epoch, long long timeout_cycles, int* status,
                                                        int32_t* __restrict__ idx_out, int32_t* __restrict__ dist_out) {
  __shared__ int s_ok;
  if (threadIdx.x == 0) s_ok = 1;
  __syncthreads();
  if (threadIdx.x < world) {
    const volatile unsigned int* f = flags + threadIdx.x;
    const long long t0 = clock64();
    while ((int)(*f - epoch) < 0) {
      if (clock64() - t0 > timeout_cycles) { s_ok = 0; atomicExch(status, 1); break; }
      __nanosleep(100);
    }
    __threadfence_system();
  }
  __syncthreads();
  if (!s_ok) return;
  const int qi = blockIdx.x * blockDim.x + threadIdx.x;
  if (qi >= nq) return;
  unsigned long long b0 = KNN_NONE, b1 = KNN_NONE;
  for (int r = 0; r < world; ++r) {
    const volatile unsigned long long* o = keys + (((size_t)slot * world + r) * max_nq + qi) * 2;
#pragma unroll
    for (int k = 0; k < 2; ++k) {
      const unsigned long long v = o[k];
      if (v < b0) { b1 = b0; b0 = v; }
      else if (v < b1) b1 = v;
    }
  }
  idx_out[2 * qi] = b0 == KNN_NONE ? -1 : (int32_t)(uint32_t)(b0 & 0xffffffffu);
  dist_out[2 * qi] = b0 == KNN_NONE ? -1 : (int32_t)(b0 >> 32);
  idx_out[2 * qi + 1] = b1 == KNN_NONE ? -1 : (int32_t)(uint32_t)(b1 & 0xffffffffu);
  dist_out[2 * qi + 1] = b1 == KNN_NONE ? -1 : (int32_t)(b1 >> 32);
}

struct orb_knn_exchange {
  orb_handle* h = nullptr;
  int device = 0;                     // kept separately: the exchange may be destroyed after its handle
  int rank = 0, world = 1, max_nq = 0;
  uint8_t* base = nullptr;            // this rank's buffer (cudaMalloc)
  size_t keys_bytes = 0, bytes = 0;
  unsigned int epoch = 0;
  long long timeout_cycles = 0;       // ORB_KNN_PEER_TIMEOUT_S in SM clocks (queried once: the clock-rate attribute is slow to read)
  bool connected = false;
  void* mapped[ORB_KNN_MAX_RANKS] = {nullptr};   // cudaIpcOpenMemHandle mappings to close
  KnnPeers peers{};
  unsigned long long* keys() const { return (unsigned long long*)base; }
  unsigned int* flags() const { return (unsigned int*)(base + keys_bytes); }
  unsigned int* counter() const { return flags() + ORB_KNN_MAX_RANKS; }
  int* status() const { return (int*)(flags() + ORB_KNN_MAX_RANKS + 1); }
  void set_peer(int r, uint8_t* b) { peers.keys[r] = (unsigned long long*)b; peers.flags[r] = (unsigned int*)(b + keys_bytes); }
};

extern "C" {

int orb_hamming_knn2(orb_handle* h, const uint8_t* q, int nq, const uint8_t* db, int64_t ndb, int32_t index_base,
                     int32_t* idx_out, int32_t* dist_out, int flags) {
  if (!h || !q || nq < 1 || ndb < 0 || (ndb && !db) || !idx_out || !dist_out) return ORB_ERR_INVALID_ARG;
  if ((int64_t)index_base + ndb > 0x7fffffffLL) return orb_set_error(h, ORB_ERR_CAPACITY, "database index exceeds int32");
  int st;
  if ((st = orb_use_device(h))) return st;
  // device staging
  const uint8_t* d_q = q;
  const uint8_t* d_db = db;
  size_t need2 = 0;
  if (!(flags & ORB_SRC_DEVICE)) need2 = (size_t)nq * 32 + (size_t)ndb * 32 + 512;
  if (need2) {
    if ((st = orb_ensure(h, h->d_scratch2, need2))) return st;
    uint8_t* base = h->d_scratch2.as<uint8_t>();
    ORB_CUDA_CHECK(h, cudaMemcpyAsync(base, q, (size_t)nq * 32, cudaMemcpyHostToDevice, h->stream));
    uint8_t* dbp = base + ((size_t)nq * 32 + 255) / 256 * 256;
    if (ndb) ORB_CUDA_CHECK(h, cudaMemcpyAsync(dbp, db, (size_t)ndb * 32, cudaMemcpyHostToDevice, h->stream));
    d_q = base; d_db = dbp;
  }
  const size_t out_bytes = (size_t)nq * 2 * sizeof(int32_t);
  unsigned long long* d_part = nullptr;
  int nchunks = 0;
  uint8_t* extra = nullptr;
  if ((st = knn2_scan_to_parts(h, d_q, nq, d_db, ndb, index_base, 2 * out_bytes, &d_part, &nchunks, &extra))) return st;
  int32_t* d_idx = (flags & ORB_DST_DEVICE) ? idx_out : (int32_t*)extra;
  int32_t* d_dist = (flags & ORB_DST_DEVICE) ? dist_out : d_idx + (size_t)nq * 2;
  k_knn2_merge_keys<<<(nq + 127) / 128, 128, 0, h->stream>>>(d_part, nchunks, nq, d_idx, d_dist);
  h->launches++;
  ORB_CUDA_CHECK(h, cudaGetLastError());
  if (!(flags & ORB_DST_DEVICE)) {
    ORB_CUDA_CHECK(h, cudaMemcpyAsync(idx_out, d_idx, out_bytes, cudaMemcpyDeviceToHost, h->stream));
    ORB_CUDA_CHECK(h, cudaMemcpyAsync(dist_out, d_dist, out_bytes, cudaMemcpyDeviceToHost, h->stream));
  }
  if (!(flags & ORB_ASYNC)) ORB_CUDA_CHECK(h, cudaStreamSynchronize(h->stream));
  return ORB_OK;
}

int orb_knn_exchange_create(orb_handle* h, int rank, int world, int max_nq, orb_knn_exchange** out, uint8_t* ipc_handle_out) {
  if (!h || !out || world < 1 || world > ORB_KNN_MAX_RANKS || rank < 0 || rank >= world || max_nq < 1) return ORB_ERR_INVALID_ARG;
  int st;
  if ((st = orb_use_device(h))) return st;
  orb_knn_exchange* x = new orb_knn_exchange();
  x->h = h; x->device = h->device; x->rank = rank; x->world = world; x->max_nq = max_nq;
  x->keys_bytes = ((size_t)2 * world * max_nq * 2 * sizeof(unsigned long long) + 255) / 256 * 256;
  x->bytes = x->keys_bytes + (ORB_KNN_MAX_RANKS + 2) * sizeof(unsigned int);
  cudaError_t e = cudaMalloc((void**)&x->base, x->bytes);
  if (e == cudaSuccess) e = cudaMemset(x->base, 0, x->bytes);
  if (e == cudaSuccess && ipc_handle_out) {
    cudaIpcMemHandle_t hnd;
    static_assert(sizeof(cudaIpcMemHandle_t) == ORB_IPC_HANDLE_BYTES, "IPC handle size");
    e = cudaIpcGetMemHandle(&hnd, x->base);
    if (e == cudaSuccess) memcpy(ipc_handle_out, &hnd, sizeof(hnd));
  }
  if (e != cudaSuccess) {
    if (x->base) cudaFree(x->base);
    delete x;
    return orb_set_error(h, ORB_ERR_CUDA, std::string("knn exchange buffer: ") + cudaGetErrorString(e));
  }
  int khz = 1965000;
  cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, h->device);
  int timeout_s = ORB_KNN_PEER_TIMEOUT_S;
  if (const char* e = getenv("ORB_B200_KNN_TIMEOUT_S")) { const int v = atoi(e); if (v >= 1 && v <= 3600) timeout_s = v; }
  x->timeout_cycles = (long long)khz * 1000LL * timeout_s;
  x->set_peer(rank, x->base);
  x->connected = world == 1;
  *out = x;
  return ORB_OK;
}

int orb_knn_exchange_connect(orb_knn_exchange* x, const uint8_t* all_handles) {
  if (!x || !all_handles) return ORB_ERR_INVALID_ARG;
  int st;
  if ((st = orb_use_device(x->h))) return st;
  for (int r = 0; r < x->world; ++r) {
    if (r == x->rank) continue;
    cudaIpcMemHandle_t hnd;
    memcpy(&hnd, all_handles + (size_t)r * ORB_IPC_HANDLE_BYTES, sizeof(hnd));
    void* p = nullptr;
    ORB_CUDA_CHECK(x->h, cudaIpcOpenMemHandle(&p, hnd, cudaIpcMemLazyEnablePeerAccess));
    x->mapped[r] = p;
    x->set_peer(r, (uint8_t*)p);
  }
  x->connected = true;
  return ORB_OK;
}

int orb_knn_exchange_connect_local(orb_knn_exchange* x, orb_knn_exchange* const* all) {
  if (!x || !all) return ORB_ERR_INVALID_ARG;
  for (int r = 0; r < x->world; ++r) {
    if (!all[r] || all[r]->world != x->world || all[r]->max_nq != x->max_nq || all[r]->rank != r)
      return orb_set_error(x->h, ORB_ERR_INVALID_ARG, "knn exchange: the local peers do not form one group");
    if (all[r]->h->device != x->h->device) {
      int can = 0;
      cudaDeviceCanAccessPeer(&can, x->h->device, all[r]->h->device);
      if (!can) return orb_set_error(x->h, ORB_ERR_CUDA, "knn exchange: no peer access between the two devices");
      int st;
      if ((st = orb_use_device(x->h))) return st;
      cudaError_t e = cudaDeviceEnablePeerAccess(all[r]->h->device, 0);
      if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) return orb_set_error(x->h, ORB_ERR_CUDA, cudaGetErrorString(e));
      cudaGetLastError();
    }
    x->set_peer(r, all[r]->base);
  }
  x->connected = true;
  return ORB_OK;
}

// completes the searches enqueued with ORB_ASYNC and reports a peer time-out that happened in any of them
int orb_knn_exchange_check(orb_knn_exchange* x) {
  if (!x) return ORB_ERR_INVALID_ARG;
  orb_handle* h = x->h;
  int st;
  if ((st = orb_use_device(h))) return st;
  ORB_CUDA_CHECK(h, cudaStreamSynchronize(h->stream));
  int status = 0;
  ORB_CUDA_CHECK(h, cudaMemcpy(&status, x->status(), sizeof(int), cudaMemcpyDeviceToHost));
  if (status) return orb_set_error(h, ORB_ERR_STATE, "knn exchange: a peer did not deliver its top-2 lists in time");
  return ORB_OK;
}

int orb_knn_exchange_destroy(orb_knn_exchange* x) {
  if (!x) return ORB_ERR_INVALID_ARG;
  cudaSetDevice(x->device);
  cudaDeviceSynchronize();            // not the handle's stream: the handle may be gone already
  for (int r = 0; r < ORB_KNN_MAX_RANKS; ++r)
    if (x->mapped[r]) cudaIpcCloseMemHandle(x->mapped[r]);
  if (x->base) cudaFree(x->base);
  delete x;
  return ORB_OK;
}

int orb_hamming_knn2_sharded(orb_handle* h, orb_knn_exchange* x, const uint8_t* q, int nq, const uint8_t* db_local, int64_t ndb_local,
                             int32_t index_base, int32_t* idx_out, int32_t* dist_out, int flags) {
  if (!h || !x || x->h != h || !q || nq < 1 || ndb_local < 0 || (ndb_local && !db_local) || !idx_out || !dist_out) return ORB_ERR_INVALID_ARG;
  if (!(flags & ORB_SRC_DEVICE) || !(flags & ORB_DST_DEVICE))
    return orb_set_error(h, ORB_ERR_INVALID_ARG, "the sharded search takes device-resident queries / shard / outputs (ORB_SRC_DEVICE | ORB_DST_DEVICE)");
  if (!x->connected) return orb_set_error(h, ORB_ERR_STATE, "knn exchange: connect the peers first");
  if (nq > x->max_nq) return orb_set_error(h, ORB_ERR_CAPACITY, "knn exchange: more queries than the exchange was created for");
  if ((int64_t)index_base + ndb_local > 0x7fffffffLL) return orb_set_error(h, ORB_ERR_CAPACITY, "database index exceeds int32");
  int st;
  if ((st = orb_use_device(h))) return st;
  unsigned long long* d_part = nullptr;
  int nchunks = 0;
  if ((st = knn2_scan_to_parts(h, q, nq, db_local, ndb_local, index_base, 0, &d_part, &nchunks, nullptr))) return st;
  const unsigned int epoch = ++x->epoch;
  const int slot = (int)(epoch & 1u);
  const int blocks = (nq + 127) / 128;
  k_knn2_merge_push<<<blocks, 128, 0, h->stream>>>(d_part, nchunks, nq, x->peers, x->rank, x->world, x->max_nq, slot, epoch, x->counter());
  k_knn2_merge_wait<<<blocks, 128, 0, h->stream>>>(x->keys(), x->flags(), x->world, nq, x->max_nq, slot, epoch, x->timeout_cycles, x->status(),
                                                   idx_out, dist_out);
  h->launches += 2;
  ORB_CUDA_CHECK(h, cudaGetLastError());
  if (flags & ORB_ASYNC) return ORB_OK;
  return orb_knn_exchange_check(x);
}

int orb_knn2_merge(orb_handle* h, const int32_t* idx_parts, const int32_t* dist_parts, int nparts, int nq, int32_t* idx_out,
                   int32_t* dist_out, int flags) {
  if (!h || !idx_parts || !dist_parts || nparts < 1 || nq < 1 || !idx_out || !dist_out) return ORB_ERR_INVALID_ARG;
  int st;
  if ((st = orb_use_device(h))) return st;
  const size_t in_bytes = (size_t)nparts * nq * 2 * sizeof(int32_t), out_bytes = (size_t)nq * 2 * sizeof(int32_t);
  const int32_t *di = idx_parts, *dd = dist_parts;
  int32_t *oi = idx_out, *od = dist_out;
  const bool src_dev = flags & ORB_SRC_DEVICE, dst_dev = flags & ORB_DST_DEVICE;
  if (!src_dev || !dst_dev) {
    if ((st = orb_ensure(h, h->d_scratch, 2 * in_bytes + 2 * out_bytes + 1024))) return st;
    uint8_t* b = h->d_scratch.as<uint8_t>();
    if (!src_dev) {
      ORB_CUDA_CHECK(h, cudaMemcpyAsync(b, idx_parts, in_bytes, cudaMemcpyHostToDevice, h->stream));
      ORB_CUDA_CHECK(h, cudaMemcpyAsync(b + in_bytes, dist_parts, in_bytes, cudaMemcpyHostToDevice, h->stream));
      di = (const int32_t*)b; dd = (const int32_t*)(b + in_bytes);
    }
    if (!dst_dev) { oi = (int32_t*)(b + 2 * in_bytes); od = (int32_t*)(b + 2 * in_bytes + out_bytes); }
  }
  k_knn2_merge_parts<<<(nq + 127) / 128, 128, 0, h->stream>>>(di, dd, nparts, nq, oi, od);
  h->launches++;
  ORB_CUDA_CHECK(h, cudaGetLastError());
  if (!dst_dev) {
    ORB_CUDA_CHECK(h, cudaMemcpyAsync(idx_out, oi, out_bytes, cudaMemcpyDeviceToHost, h->stream));
    ORB_CUDA_CHECK(h, cudaMemcpyAsync(dist_out, od, out_bytes, cudaMemcpyDeviceToHost, h->stream));
  }
  if (!(flags & ORB_ASYNC)) ORB_CUDA_CHECK(h, cudaStreamSynchronize(h->stream));
  return ORB_OK;
}

int orb_ratio_test(orb_handle* h, const int32_t* dist, int nq, uint8_t* pass_out, int flags) {
  if (!h || !dist || nq < 1 || !pass_out) return ORB_ERR_INVALID_ARG;
  int st;
  if ((st = orb_use_device(h))) return st;
  const int32_t* dd = dist;
  uint8_t* dp = pass_out;
  const bool src_dev = flags & ORB_SRC_DEVICE, dst_dev = flags & ORB_DST_DEVICE;
  if (!src_dev || !dst_dev) {
    if ((st = orb_ensure(h, h->d_scratch, (size_t)nq * 8 + nq + 512))) return st;
    uint8_t* b = h->d_scratch.as<uint8_t>();
    if (!src_dev) { ORB_CUDA_CHECK(h, cudaMemcpyAsync(b, dist, (size_t)nq * 8, cudaMemcpyHostToDevice, h->stream)); dd = (const int32_t*)b; }
    if (!dst_dev) dp = b + (size_t)nq * 8;
  }
  k_ratio_test<<<(nq + 127) / 128, 128, 0, h->stream>>>(dd, nq, dp);
  h->launches++;
  ORB_CUDA_CHECK(h, cudaGetLastError());
  if (!dst_dev) ORB_CUDA_CHECK(h, cudaMemcpyAsync(pass_out, dp, nq, cudaMemcpyDeviceToHost, h->stream));
  if (!(flags & ORB_ASYNC)) ORB_CUDA_CHECK(h, cudaStreamSynchronize(h->stream));
  return ORB_OK;
}

int orb_stereo_fisheye_match_batch(orb_handle* hL, orb_handle* hR, int32_t* idx_out, int32_t* dist_out, uint8_t* pass_out, int cap, int flags) {
  if (!hL || !hR) return ORB_ERR_INVALID_ARG;
  if (!hL->have_batch || !hR->have_batch) return orb_set_error(hL, ORB_ERR_STATE, "fisheye stereo match needs an extraction on both handles");
  if (hL->device != hR->device) return orb_set_error(hL, ORB_ERR_INVALID_ARG, "both handles must live on the same device");
  if (hL->cur_batch != hR->cur_batch) return orb_set_error(hL, ORB_ERR_INVALID_ARG, "left/right batches differ");
  int st;
  if ((st = orb_use_device(hL))) return st;
  const int batch = hL->cur_batch, kcap = hL->g.kcap;
  const size_t n = (size_t)batch * kcap;
  if ((st = orb_ensure(hL, hL->d_fe_idx, n * 2 * sizeof(int))) || (st = orb_ensure(hL, hL->d_fe_dist, n * 2 * sizeof(int))) ||
      (st = orb_ensure(hL, hL->d_fe_pass, n)))
    return st;
  if ((st = orb_peer_read_begin(hL, hR))) return st;   // order hL's stream after everything queued on hR's stream
  if (batch > FE_SMALL_BATCH) {   // (the small-batch path's merge kernel writes every entry itself)
    ORB_CUDA_CHECK(hL, cudaMemsetAsync(hL->d_fe_pass.p, 0, n, hL->stream));
    ORB_CUDA_CHECK(hL, cudaMemsetAsync(hL->d_fe_idx.p, 0xff, n * 2 * sizeof(int), hL->stream));
    ORB_CUDA_CHECK(hL, cudaMemsetAsync(hL->d_fe_dist.p, 0xff, n * 2 * sizeof(int), hL->stream));
  }
  const int qblocks = (kcap + FE_THREADS * FE_QPT - 1) / (FE_THREADS * FE_QPT);
  if (batch <= FE_SMALL_BATCH) {
    // a few frames leave the GPU empty with one block per 384 queries (one TUM-VI pair: 4 blocks scanning 1500 rows each, 0.2 ms):
    // split the right keypoints into chunks of FE_CHUNK_ROWS over blockIdx.z and merge the chunks' top-2 lists
    const int nch = (hR->g.kcap + FE_CHUNK_ROWS - 1) / FE_CHUNK_ROWS;
    if ((st = orb_ensure(hL, hL->d_fe_part, (size_t)nch * n * 2 * sizeof(unsigned long long)))) return st;
    k_fisheye_knn2<<<dim3(qblocks, batch, nch), FE_THREADS, 0, hL->stream>>>(
        hL->d_desc.as<uint8_t>(), hL->d_n.as<int>(), hL->d_mono.as<int>(), kcap, hR->d_desc.as<uint8_t>(), hR->d_n.as<int>(), hR->d_mono.as<int>(),
        hR->g.kcap, kcap, nullptr, nullptr, nullptr, FE_CHUNK_ROWS, hL->d_fe_part.as<unsigned long long>());
    k_fisheye_merge<<<dim3((kcap + 127) / 128, batch), 128, 0, hL->stream>>>(
        hL->d_fe_part.as<unsigned long long>(), FE_CHUNK_ROWS, hL->d_n.as<int>(), hL->d_mono.as<int>(), kcap, hR->d_n.as<int>(), hR->d_mono.as<int>(),
        hR->g.kcap, kcap, hL->d_fe_idx.as<int32_t>(), hL->d_fe_dist.as<int32_t>(), hL->d_fe_pass.as<uint8_t>());
    hL->launches += 2;
  } else {
    k_fisheye_knn2<<<dim3(qblocks, batch), FE_THREADS, 0, hL->stream>>>(
        hL->d_desc.as<uint8_t>(), hL->d_n.as<int>(), hL->d_mono.as<int>(), kcap, hR->d_desc.as<uint8_t>(), hR->d_n.as<int>(), hR->d_mono.as<int>(),
        hR->g.kcap, kcap, hL->d_fe_idx.as<int32_t>(), hL->d_fe_dist.as<int32_t>(), hL->d_fe_pass.as<uint8_t>(), 0, nullptr);
    hL->launches++;
  }
  if ((st = orb_peer_read_end(hL, hR))) return st;     // hR's next extraction waits for this kernel
  hL->have_fe = true;
  ORB_CUDA_CHECK(hL, cudaGetLastError());
  if (!(flags & ORB_NO_OUTPUT)) {
    const int rows = std::min(cap, kcap);
    if (idx_out)
      ORB_CUDA_CHECK(hL, cudaMemcpy2DAsync(idx_out, (size_t)cap * 8, hL->d_fe_idx.p, (size_t)kcap * 8, (size_t)rows * 8, batch, cudaMemcpyDefault, hL->stream));
    if (dist_out)
      ORB_CUDA_CHECK(hL, cudaMemcpy2DAsync(dist_out, (size_t)cap * 8, hL->d_fe_dist.p, (size_t)kcap * 8, (size_t)rows * 8, batch, cudaMemcpyDefault, hL->stream));
    if (pass_out)
      ORB_CUDA_CHECK(hL, cudaMemcpy2DAsync(pass_out, (size_t)cap, hL->d_fe_pass.p, (size_t)kcap, (size_t)rows, batch, cudaMemcpyDefault, hL->stream));
  }
  if (flags & ORB_ASYNC) return ORB_OK;
  ORB_CUDA_CHECK(hL, cudaStreamSynchronize(hL->stream));
  return ORB_OK;
}

}  // extern "C"
