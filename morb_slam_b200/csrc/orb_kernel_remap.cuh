// Input rectification on the device: System::TrackStereo's cv::remap(imLeft, imLeftToFeed, M1l, M2l, cv::INTER_LINEAR)
// (reference src/System.cc:254-261; maps from cv::initUndistortRectifyMap(..., CV_32F, ...), src/Settings.cc:540-545),
// 8UC1, BORDER_CONSTANT 0. OpenCV's fixed-point path, restated and pinned against cv2 in oracle/shim (remap_linear_8u):
//   sx = cvRound(mapx * 32) (cvtss2si: 0x80000000 outside the int range / NaN), integer part saturated to short, the
//   5-bit fractions select four 15-bit weights ((32 - fy)(32 - fx) * 32, ...; fractions (0, 0): {32767, 0, 0, 1}),
//   dst = (sum w_i * p_i + 2^14) >> 15, taps outside the source count as 0.
// One thread = 4 adjacent destination pixels (one 32-bit store) of RM_FRAMES frames: the maps are the same for every
// frame of the batch, so the per-pixel tap offsets and weights are computed once and reused; a tap outside the source
// gets weight 0 and a clamped (valid) address.
#pragma once

#define RM_FRAMES 8

static __device__ __forceinline__ int rm_round(float v) {
  return (v >= -2147483648.f && v < 2147483648.f) ? __float2int_rn(v) : (int)0x80000000;
}

__global__ void __launch_bounds__(256) k_remap(const uint8_t* __restrict__ raw, int sw, int sh, size_t sstride, size_t sframe,
                                              const float* __restrict__ mapx, const float* __restrict__ mapy, int dw, int dh,
                                              uint8_t* __restrict__ dst, int dpitch, size_t dframe, int batch) {
  const int wpr = (dw + 3) >> 2;
  const int wi = blockIdx.x * 256 + threadIdx.x;
  if (wi >= wpr * dh) return;
  const int y = wi / wpr, x0 = (wi - y * wpr) * 4;
  int row0[4], row1[4], xs[4];
  unsigned w01[4], w23[4];
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const int x = min(x0 + k, dw - 1);   // the padded tail of a row repeats the last pixel (never read downstream)
    const int fsx = rm_round(__fmul_rn(mapx[(size_t)y * dw + x], 32.f)), fsy = rm_round(__fmul_rn(mapy[(size_t)y * dw + x], 32.f));
    const int sx = min(max(fsx >> 5, -32768), 32767), sy = min(max(fsy >> 5, -32768), 32767);
    const int fx = fsx & 31, fy = fsy & 31;
    int w0 = (32 - fy) * (32 - fx) * 32, w1 = (32 - fy) * fx * 32, w2 = fy * (32 - fx) * 32, w3 = fy * fx * 32;
    if ((fx | fy) == 0) { w0 = 32767; w3 = 1; }
    const bool vx0 = sx >= 0 && sx < sw, vx1 = sx + 1 >= 0 && sx + 1 < sw, vy0 = sy >= 0 && sy < sh, vy1 = sy + 1 >= 0 && sy + 1 < sh;
    if (!(vx0 && vy0)) w0 = 0;
    if (!(vx1 && vy0)) w1 = 0;
    if (!(vx0 && vy1)) w2 = 0;
    if (!(vx1 && vy1)) w3 = 0;
    const int cx0 = min(max(sx, 0), sw - 1), cx1 = min(max(sx + 1, 0), sw - 1);
    row0[k] = min(max(sy, 0), sh - 1);
    row1[k] = min(max(sy + 1, 0), sh - 1);
    xs[k] = cx0 | (cx1 << 16);
    w01[k] = (unsigned)w0 | ((unsigned)w1 << 16);
    w23[k] = (unsigned)w2 | ((unsigned)w3 << 16);
  }
  const int f0 = blockIdx.y * RM_FRAMES, f1 = min(f0 + RM_FRAMES, batch);
  for (int f = f0; f < f1; ++f) {
    const uint8_t* s = raw + (size_t)f * sframe;
    unsigned out = 0;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const uint8_t* r0 = s + (size_t)row0[k] * sstride;
      const uint8_t* r1 = s + (size_t)row1[k] * sstride;
      const int cx0 = xs[k] & 0xffff, cx1 = xs[k] >> 16;
      const unsigned acc = r0[cx0] * (w01[k] & 0xffffu) + r0[cx1] * (w01[k] >> 16) + r1[cx0] * (w23[k] & 0xffffu) + r1[cx1] * (w23[k] >> 16);
      out |= min((acc + (1u << 14)) >> 15, 255u) << (8 * k);
    }
    *reinterpret_cast<unsigned*>(dst + (size_t)f * dframe + (size_t)y * dpitch + x0) = out;
  }
}
