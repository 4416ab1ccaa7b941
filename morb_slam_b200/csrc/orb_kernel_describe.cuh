// Orientation + descriptor, one warp per keypoint (reference src/ORBextractor.cc:75-145, 466-473).
//   IC_Angle (:75-99): first-order moments over the radius-15 disc of the UN-blurred level (lane = column),
//   then cv::fastAtan2 (dev_fast_atan2: OpenCV's degree polynomial, operation by operation).
//   computeOrbDescriptor (:102-145): 256 comparisons of the BLURRED level sampled at the pattern rotated by the
//   angle: a = cosf, b = sinf (glibc's float routines restated in double), row = cvRound(x*b + y*a),
//   col = cvRound(x*a - y*b) with separate roundings (no FMA). Lane i builds descriptor byte i.
//
// The kernel is latency-bound (short dependent chain per keypoint), so it is written for occupancy and few
// memory round trips: one compact (key, slot|level) record per keypoint from k_assemble, the 37x37 blurred
// patch staged in shared memory with aligned 32-bit loads (2 rows per warp instruction), the lane's 16
// pattern points held as packed int8 in 8 registers (two coalesced 16-byte loads), 32 bytes out as two
// 16-byte stores.
#pragma once

#define DESC_WARPS 8
#define DESC_PW 12  // words per staged patch row (37 bytes + alignment offset <= 40 bytes, padded)

static __device__ __forceinline__ float s8_to_float(uint32_t w, int k) {
  return (float)(int)(int8_t)(w >> (8 * k));
}

#ifndef DESC_MINB
#define DESC_MINB 5
#endif
__global__ void __launch_bounds__(DESC_WARPS * 32, DESC_MINB) k_orient_describe(
    OrbGeom g, const uint8_t* __restrict__ pyr, const uint8_t* __restrict__ blur, const int* __restrict__ n_arr,
    const uint32_t* __restrict__ ord_key, const int* __restrict__ ord_slot, const uint4* __restrict__ pattern,
    orb_keypoint* __restrict__ kps, uint8_t* __restrict__ desc) {
  __shared__ uint32_t s_patch[DESC_WARPS][38 * DESC_PW];
  const int frame = blockIdx.y;
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const int ord = blockIdx.x * DESC_WARPS + wid;
  if (ord >= n_arr[frame]) return;
  const uint32_t k = ord_key[(size_t)frame * g.kcap + ord];
  const int sl = ord_slot[(size_t)frame * g.kcap + ord];
  // this lane's 16 pattern points: 32 int8 (x0,y0,x1,y1,...) = two 16-byte loads, coalesced over the warp
  const uint4 pa = pattern[2 * lane], pb = pattern[2 * lane + 1];
  const int l = sl & 15, slot = sl >> 4;
  const int cx = orb_px(k) + ORB_BORDER, cy = orb_py(k) + ORB_BORDER;
  const int P = g.pitch[l];
  uint32_t* patch_w = s_patch[wid];
  // ---- stage the 37x37 blurred patch (pattern radius <= 18.4 -> rounded offsets within +-18): 16 lanes per
  //      row, two rows per instruction; rows of a level are 16-byte aligned, so the word loads are aligned
  const int xs = cx - 18, xa = xs & ~3, off = xs - xa;
  {
    const int ncols = (off + 37 + 3) >> 2;  // <= 11
    const int c = lane & 15, rsub = lane >> 4;
    const uint8_t* __restrict__ b0 = lvl_ptr(g, blur, frame, l) + (size_t)(cy - 18 + rsub) * P + xa + 4 * c;
    if (c < ncols) {
#pragma unroll
      for (int i = 0; i < 19; ++i) {
        const int r = 2 * i + rsub;
        if (r < 37) patch_w[r * DESC_PW + c] = *reinterpret_cast<const uint32_t*>(b0 + (size_t)(2 * i) * P);
      }
    }
  }
  // ---- IC_Angle on the un-blurred level: lane u handles column offset u - 15 (lane 31 idles)
  int m10 = 0, m01 = 0;
  {
    const uint8_t* __restrict__ c = lvl_ptr(g, pyr, frame, l) + (size_t)cy * P + cx;
    const int u = lane - ORB_HALF_PATCH;
    if (lane < 31) {
      const int au = u < 0 ? -u : u;
      m10 = u * c[u];
#pragma unroll
      for (int v = 1; v <= ORB_HALF_PATCH; ++v) {
        if (au <= c_umax[v]) {
          const int vp = c[u + v * P], vm = c[u - v * P];
          m10 += u * (vp + vm);
          m01 += v * (vp - vm);
        }
      }
    }
    m10 = __reduce_add_sync(0xffffffffu, m10);
    m01 = __reduce_add_sync(0xffffffffu, m01);
  }
  const float angle = dev_fast_atan2((float)m01, (float)m10);
  const float factorPI = 0.017453292519943295f;  // (float)(CV_PI / 180.f)
  float a, b;
  dev_glibc_sincosf(__fmul_rn(angle, factorPI), &b, &a);
  __syncwarp();
  // ---- 8 comparisons of this lane
  const uint8_t* pc = reinterpret_cast<const uint8_t*>(patch_w) + 18 * (DESC_PW * 4) + off + 18;
  const uint32_t pw[8] = {pa.x, pa.y, pa.z, pa.w, pb.x, pb.y, pb.z, pb.w};
  uint32_t val = 0;
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const float x0 = s8_to_float(pw[j], 0), y0 = s8_to_float(pw[j], 1), x1 = s8_to_float(pw[j], 2), y1 = s8_to_float(pw[j], 3);
    const int r0 = __float2int_rn(__fadd_rn(__fmul_rn(x0, b), __fmul_rn(y0, a)));
    const int q0 = __float2int_rn(__fsub_rn(__fmul_rn(x0, a), __fmul_rn(y0, b)));
    const int r1 = __float2int_rn(__fadd_rn(__fmul_rn(x1, b), __fmul_rn(y1, a)));
    const int q1 = __float2int_rn(__fsub_rn(__fmul_rn(x1, a), __fmul_rn(y1, b)));
    const int t0 = pc[r0 * (DESC_PW * 4) + q0], t1 = pc[r1 * (DESC_PW * 4) + q1];
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
  }
}
