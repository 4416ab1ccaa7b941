// Orientation + descriptor, one warp per keypoint (reference src/ORBextractor.cc:75-145, 466-473).
//   IC_Angle (:75-99): first-order moments over the radius-15 disc of the UN-blurred level, then cv::fastAtan2
//   (dev_fast_atan2: OpenCV's degree polynomial, operation by operation).
//   computeOrbDescriptor (:102-145): 256 comparisons of the BLURRED level sampled at the pattern rotated by the
//   angle: a = cosf, b = sinf (glibc's float routines restated in double), row = cvRound(x*b + y*a),
//   col = cvRound(x*a - y*b) with separate roundings (no FMA). Lane i builds descriptor byte i.
//
// Round 2 (profiles/README_r2.md): the kernel issued 905 warp instructions per keypoint, a third of them the byte loads
// and address arithmetic of the moments, and a first rewrite that only cut instructions ran no faster: it had moved the
// bound to the load / store unit (about 235 L1 wavefronts per keypoint). Now
//   patch     the blurred patch arrives with ONE TMA box copy per warp (64 x 37 bytes from the 16-byte boundary at or
//             before cx - 18; the warp's own mbarrier), issued first and awaited after the orientation is known: no
//             load / store instructions, no wavefronts, and its latency hides behind the moments;
//   moments   the disc of the un-blurred level arrives the same way (box of 48 x 31 bytes from the 16-byte boundary at or before
//             cx - 15, a second mbarrier): 31 rows x 12 words = 372 (row, word) items in the order they lie in shared memory, 12 per
//             lane, so a warp instruction reads 32 consecutive words (no bank conflict, no address arithmetic: one base register
//             and immediate offsets); an item is two byte dot products (IDP.4A) of the pixel word with the u weights and with the
//             v weights (host table per byte offset of cx - 15 inside the box: 16 x 384 entries of 8 bytes, 0 outside the disc);
//   pattern   the lane's 16 sampling points come as floats from a transposed table (no int8 -> float conversions).
#pragma once

#define DESC_WARPS 8
#define DESC_BOXW 64    // TMA box: 64 bytes x 37 rows
#define DESC_SLOT 2432  // 37 * 64 = 2368 rounded up to a multiple of 128 (TMA destination alignment)
#define ORB_IC_ITEMS 384   // 372 (row, word) items of the moment disc's box, padded to 12 per lane
#define ORB_IC_BOXW 48     // TMA box of the moment disc: 48 bytes x 31 rows
#define ORB_IC_SLOT 1536   // 384 words

#ifndef DESC_MINB
#define DESC_MINB 5
#endif

// dot product of 4 unsigned pixel bytes with 4 signed weight bytes
static __device__ __forceinline__ int dp4a_us(uint32_t px, uint32_t w, int acc) {
  int d;
  asm("dp4a.u32.s32 %0, %1, %2, %3;" : "=r"(d) : "r"(px), "r"(w), "r"(acc));
  return d;
}

__global__ void __launch_bounds__(DESC_WARPS * 32, DESC_MINB) k_orient_describe(
    const __grid_constant__ BlurMaps dmaps, const __grid_constant__ BlurMaps imaps, OrbGeom g, const int* __restrict__ n_arr,
    const uint32_t* __restrict__ ord_key, const int* __restrict__ ord_slot, const float4* __restrict__ patf,
    const uint2* __restrict__ ictab, orb_keypoint* __restrict__ kps, uint8_t* __restrict__ desc, orb_keypoint* __restrict__ host_kps,
    uint8_t* __restrict__ host_desc, int host_cap) {
  // host_kps / host_desc (small batches, page-locked result buffers of the caller, host_cap records per frame): the results also go
  // straight to the host from here, so the extraction ends without device-to-host copies
  __shared__ __align__(128) uint8_t s_patch[DESC_WARPS][DESC_SLOT];
  __shared__ __align__(128) uint8_t s_disc[DESC_WARPS][ORB_IC_SLOT];
  __shared__ __align__(8) uint64_t s_bar[DESC_WARPS][2];
  const int frame = blockIdx.y;
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const int ord = blockIdx.x * DESC_WARPS + wid;
  if (ord >= n_arr[frame]) return;
  const uint32_t k = ord_key[(size_t)frame * g.kcap + ord];
  const int sl = ord_slot[(size_t)frame * g.kcap + ord];
  const int l = sl & 15, slot = sl >> 4;
  const int cx = orb_px(k) + ORB_BORDER, cy = orb_py(k) + ORB_BORDER;
  uint8_t* patch = s_patch[wid];
  const uint32_t* disc = reinterpret_cast<const uint32_t*>(s_disc[wid]);
  // ---- the moment disc of the un-blurred level and the 37x37 blurred patch (pattern radius <= 18.4 -> rounded offsets within
  //      +-18): one box copy each, columns from the 16-byte boundary at or before cx - 15 / cx - 18 (off <= 15: 15 + 31 <= 48,
  //      15 + 37 <= 64), rows cy - 15 .. cy + 15 / cy - 18 .. cy + 18 of this frame
  const int x15 = cx - ORB_HALF_PATCH, xd = x15 & ~15;
  const int xs = cx - 18, xa = xs & ~15, off = xs - xa;
  if (lane == 0) {
    tma_load_tile(s_disc[wid], &imaps.m[l], xd, frame * g.h[l] + cy - ORB_HALF_PATCH, &s_bar[wid][0], ORB_IC_BOXW * 31);
    tma_load_tile(patch, &dmaps.m[l], xa, frame * g.h[l] + cy - 18, &s_bar[wid][1], DESC_BOXW * 37);
  }
  __syncwarp();          // lane 0 has armed the barriers and issued the copies
  // ---- IC_Angle: 12 (row, word) items per lane in shared-memory order; weights of this byte offset from the table
  int m10 = 0, m01 = 0;
  {
    const uint2* __restrict__ tab = ictab + (x15 - xd) * ORB_IC_ITEMS + lane;
    tma_wait(&s_bar[wid][0]);
#pragma unroll
    for (int i = 0; i < ORB_IC_ITEMS / 32; ++i) {
      const uint2 t = tab[32 * i];                       // x: u weights, y: v weights (both 0 outside the disc)
      const uint32_t px = disc[lane + 32 * i];
      m10 = dp4a_us(px, t.x, m10);
      m01 = dp4a_us(px, t.y, m01);
    }
    m10 = __reduce_add_sync(0xffffffffu, m10);
    m01 = __reduce_add_sync(0xffffffffu, m01);
  }
  const float angle = dev_fast_atan2((float)m01, (float)m10);
  const float factorPI = 0.017453292519943295f;  // (float)(CV_PI / 180.f)
  float a, b;
  dev_glibc_sincosf(__fmul_rn(angle, factorPI), &b, &a);
  tma_wait(&s_bar[wid][1]);
  // ---- 8 comparisons of this lane: comparison 8 * lane + j = entry j * 32 + lane of the transposed float pattern
  const uint8_t* pc = patch + 18 * DESC_BOXW + off + 18;
  uint32_t val = 0;
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const float4 pt = patf[j * 32 + lane];   // x0 y0 x1 y1
    const int r0 = __float2int_rn(__fadd_rn(__fmul_rn(pt.x, b), __fmul_rn(pt.y, a)));
    const int q0 = __float2int_rn(__fsub_rn(__fmul_rn(pt.x, a), __fmul_rn(pt.y, b)));
    const int r1 = __float2int_rn(__fadd_rn(__fmul_rn(pt.z, b), __fmul_rn(pt.w, a)));
    const int q1 = __float2int_rn(__fsub_rn(__fmul_rn(pt.z, a), __fmul_rn(pt.w, b)));
    const int t0 = pc[r0 * DESC_BOXW + q0], t1 = pc[r1 * DESC_BOXW + q1];
    val |= (uint32_t)(t0 < t1) << j;
  }
  // gather 32 bytes -> 8 words -> two uint4 stores
  uint32_t word = 0;
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const uint32_t bj = __shfl_sync(0xffffffffu, val, (lane & 7) * 4 + j);
    word |= bj << (8 * j);
  }
  uint4 q;
  const int base = (lane & 1) * 4;
  q.x = __shfl_sync(0xffffffffu, word, base + 0);
  q.y = __shfl_sync(0xffffffffu, word, base + 1);
  q.z = __shfl_sync(0xffffffffu, word, base + 2);
  q.w = __shfl_sync(0xffffffffu, word, base + 3);
  uint8_t* d = desc + ((size_t)frame * g.kcap + slot) * 32;
  if (lane < 2) reinterpret_cast<uint4*>(d)[lane] = q;
  const bool to_host = host_desc != nullptr && slot < host_cap;
  if (to_host && lane < 2) reinterpret_cast<uint4*>(host_desc + ((size_t)frame * host_cap + slot) * 32)[lane] = q;
  // ---- keypoint record (:829-838, :1066-1068)
  if (lane == 0) {
    float fx = (float)cx, fy = (float)cy;
    if (l != 0) { fx = __fmul_rn(fx, g.scale[l]); fy = __fmul_rn(fy, g.scale[l]); }
    orb_keypoint kp;
    kp.x = fx; kp.y = fy;
    kp.size = (float)g.patch_size[l];
    kp.angle = angle;
    kp.response = (float)orb_ps(k);
    kp.octave = l;
    kp.class_id = -1;
    kps[(size_t)frame * g.kcap + slot] = kp;
    if (to_host) host_kps[(size_t)frame * host_cap + slot] = kp;
  }
}
