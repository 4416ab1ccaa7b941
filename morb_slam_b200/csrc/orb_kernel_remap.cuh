// Input rectification on the device: System::TrackStereo's cv::remap(imLeft, imLeftToFeed, M1l, M2l, cv::INTER_LINEAR)
// (reference src/System.cc:254-261; maps from cv::initUndistortRectifyMap(..., CV_32F, ...), src/Settings.cc:540-545),
// 8UC1, BORDER_CONSTANT 0. OpenCV's fixed-point path, restated and pinned against cv2 in oracle/shim (remap_linear_8u):
//   sx = cvRound(mapx * 32) (cvtss2si: 0x80000000 outside the int range / NaN), integer part saturated to short, the
//   5-bit fractions select four 15-bit weights ((32 - fy)(32 - fx) * 32, ...; fractions (0, 0): {32767, 0, 0, 1}),
//   dst = (sum w_i * p_i + 2^14) >> 15, taps outside the source count as 0.
// CTA = a tile of 64 x 16 destination pixels, thread = 4 adjacent pixels (one 32-bit store), RM_FRAMES frames per CTA:
// the maps are the same for every frame of the batch, so tap positions and weights are computed once and reused.
// Rectification maps are smooth: the taps of a tile fall into a small source box (its integer bounds are computed on the
// host when the maps are set). The CTA stages that box in shared memory with aligned 32-bit loads and gathers from
// there - 4 byte loads, 2 PRMT and 2 IDP.2A (dp2a: 16-bit weights x 8-bit pixels) per pixel, no divergence; a tap
// outside the source has weight 0 and a clamped address. Tiles whose box does not fit (wild maps) or unaligned raw
// buffers gather from global memory instead.
#pragma once

#define RM_FRAMES 32
#define RM_TW 64
#define RM_TH 16
#define RM_BOXW 144     // bytes per staged row (multiple of 4)
#define RM_BOXH 40

static __host__ __device__ __forceinline__ int rm_sat_short(int v) { return v < -32768 ? -32768 : (v > 32767 ? 32767 : v); }

static __device__ __forceinline__ void rm_cp16(void* smem, const void* gmem) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"((uint32_t)__cvta_generic_to_shared(smem)), "l"(gmem));
}
static __device__ __forceinline__ void rm_cp4(void* smem, const void* gmem) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4;\n" ::"r"((uint32_t)__cvta_generic_to_shared(smem)), "l"(gmem));
}

static __device__ __forceinline__ int rm_round(float v) {
  return (v >= -2147483648.f && v < 2147483648.f) ? __float2int_rn(v) : (int)0x80000000;
}

__global__ void __launch_bounds__(256, 4) k_remap(const uint8_t* __restrict__ raw, int sw, int sh, size_t sstride, size_t sframe,
                                                 const float* __restrict__ mapx, const float* __restrict__ mapy, int dw, int dh,
                                                 const int4* __restrict__ tiles, int tiles_x, uint8_t* __restrict__ dst, int dpitch,
                                                 size_t dframe, int batch) {
  __shared__ __align__(16) uint8_t s_boxes[2][RM_BOXH * RM_BOXW];   // double buffer: frame f + 1 lands while frame f is gathered
  const int tid = threadIdx.x;
  const int ty = blockIdx.x / tiles_x, tx = blockIdx.x - ty * tiles_x;
  const int x0 = tx * RM_TW + 4 * (tid & 15), y = ty * RM_TH + (tid >> 4);
  const bool active = x0 < dw && y < dh;
  // source box of the tile (taps clamped into the source like the per-pixel addresses below)
  const int4 t = tiles[blockIdx.x];
  const int bx0 = min(max(t.x, 0), sw - 1) & ~3, bx1 = min(max(t.y + 1, 0), sw - 1);
  const int by0 = min(max(t.z, 0), sh - 1), by1 = min(max(t.w + 1, 0), sh - 1);
  const int bw = bx1 - bx0 + 1, bh = by1 - by0 + 1;
  const bool staged = bw <= RM_BOXW && bh <= RM_BOXH && ((sstride | sframe | (size_t)raw) & 3) == 0;
  int row0[4], row1[4], xs[4];
  unsigned w01[4], w23[4];
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const int x = min(x0 + k, dw - 1), yy = min(y, dh - 1);   // the padded tail of a row repeats the last pixel (never read downstream)
    const int fsx = rm_round(__fmul_rn(mapx[(size_t)yy * dw + x], 32.f)), fsy = rm_round(__fmul_rn(mapy[(size_t)yy * dw + x], 32.f));
    const int sx = rm_sat_short(fsx >> 5), sy = rm_sat_short(fsy >> 5);
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
  if (staged) {                                   // CTA-uniform
    // shared-memory offsets of the four taps of every pixel: (row0, cx0) | (row0, cx1) << 16 and the same for row1
    unsigned o0[4], o1[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int c0 = (xs[k] & 0xffff) - bx0, c1 = (xs[k] >> 16) - bx0;
      const int r0 = (row0[k] - by0) * RM_BOXW, r1 = (row1[k] - by0) * RM_BOXW;
      o0[k] = (unsigned)(r0 + c0) | ((unsigned)(r0 + c1) << 16);
      o1[k] = (unsigned)(r1 + c0) | ((unsigned)(r1 + c1) << 16);
    }
    // staging: 16-byte loads when the raw rows allow it (the box then starts at a 16-byte boundary, bw16 <= RM_BOXW is
    // checked here), else 4-byte loads; thread = (row tid >> 4 (+ 16 per pass), column tid & 15 (+ 16 per pass))
    const int bx16 = bx0 & ~15, shift16 = bx0 - bx16;
    const bool vec16 = ((sstride | sframe | (size_t)raw) & 15) == 0 && (bx1 - bx16 + 1) <= RM_BOXW;
    const int ncol = vec16 ? (bx1 - bx16 + 16) >> 4 : (bw + 3) >> 2;
    const int srow = tid >> 4, scol = tid & 15;
    if (vec16) {
#pragma unroll
      for (int k = 0; k < 4; ++k) { o0[k] += (unsigned)shift16 * 0x10001u; o1[k] += (unsigned)shift16 * 0x10001u; }
    }
    const size_t box_org = (size_t)by0 * sstride + (vec16 ? bx16 : bx0);
    // thread = (row tid >> 4 (+ 16 per pass), column tid & 15); ncol <= 16 with 16-byte copies, <= 36 with 4-byte ones
    const int unit = vec16 ? 16 : 4;
    const size_t g_thread = (size_t)srow * sstride + (size_t)unit * scol;
    const int s_thread = srow * RM_BOXW + unit * scol;
    auto stage = [&](int f, uint8_t* buf) {
      const uint8_t* s = raw + (size_t)f * sframe + box_org;
      if (vec16) {
        if (scol < ncol) {
          const uint8_t* gp = s + g_thread;
          uint8_t* sp = buf + s_thread;
          for (int r = srow; r < bh; r += 16, gp += 16 * sstride, sp += 16 * RM_BOXW) rm_cp16(sp, gp);
        }
      } else {
        for (int c = scol; c < ncol; c += 16) {
          const uint8_t* gp = s + g_thread + 4 * (c - scol);
          uint8_t* sp = buf + s_thread + 4 * (c - scol);
          for (int r = srow; r < bh; r += 16, gp += 16 * sstride, sp += 16 * RM_BOXW) rm_cp4(sp, gp);
        }
      }
      asm volatile("cp.async.commit_group;\n" ::);
    };
    stage(f0, s_boxes[0]);
    for (int f = f0; f < f1; ++f) {
      const uint8_t* s_box = s_boxes[(f - f0) & 1];
      asm volatile("cp.async.wait_group 0;\n" ::);
      __syncthreads();                            // frame f has landed for everybody; the gathers of frame f - 1 are done
      if (f + 1 < f1) stage(f + 1, s_boxes[(f + 1 - f0) & 1]);
      if (active) {
        unsigned out = 0;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const unsigned p0 = __byte_perm(s_box[o0[k] & 0xffffu], s_box[o0[k] >> 16], 0x0040);
          const unsigned p1 = __byte_perm(s_box[o1[k] & 0xffffu], s_box[o1[k] >> 16], 0x0040);
          const unsigned acc = __dp2a_lo(w23[k], p1, __dp2a_lo(w01[k], p0, 1u << 14));
          out |= (acc >> 15) << (8 * k);          // acc < 255 * 32768 + 2^15: no clamp needed
        }
        *reinterpret_cast<unsigned*>(dst + (size_t)f * dframe + (size_t)y * dpitch + x0) = out;
      }
    }
    return;
  }
  if (!active) return;
  for (int f = f0; f < f1; ++f) {
    const uint8_t* s = raw + (size_t)f * sframe;
    unsigned out = 0;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const uint8_t* r0 = s + (size_t)row0[k] * sstride;
      const uint8_t* r1 = s + (size_t)row1[k] * sstride;
      const int cx0 = xs[k] & 0xffff, cx1 = xs[k] >> 16;
      const unsigned acc = r0[cx0] * (w01[k] & 0xffffu) + r0[cx1] * (w01[k] >> 16) + r1[cx0] * (w23[k] & 0xffffu) + r1[cx1] * (w23[k] >> 16);
      out |= ((acc + (1u << 14)) >> 15) << (8 * k);
    }
    *reinterpret_cast<unsigned*>(dst + (size_t)f * dframe + (size_t)y * dpitch + x0) = out;
  }
}

// Input resize on the device: System::TrackStereo's cv::resize(imLeft, imLeftToFeed, settings_->newImSize()) for settings
// with needToResize() (reference src/System.cc:262-264; also TrackMonocular / TrackRGBD). cv::resize INTER_LINEAR 8U with
// the tables of SURVEY.md A.1 (11-bit weights, host-built like the pyramid's), an exact 2x shrink is OpenCV's 2x2 box
// filter. Same arithmetic as k_resize_level, explicit source (tight raw frames) and destination (level 0).
__global__ void __launch_bounds__(256) k_resize_input(const uint8_t* __restrict__ raw, int sw, int sh, size_t sframe, uint8_t* __restrict__ dst,
                                                      int dw, int dh, int dp, size_t dframe, const int2* __restrict__ xtab,
                                                      const int2* __restrict__ ytab, int area2x) {
  const int frame = blockIdx.z;
  const int y = blockIdx.y * blockDim.y + threadIdx.y;
  const int x0 = (blockIdx.x * blockDim.x + threadIdx.x) * 4;
  if (y >= dh || x0 >= dw) return;
  const uint8_t* __restrict__ src = raw + (size_t)frame * sframe;
  uint32_t packed = 0;
  if (area2x) {
    const uint8_t* s0 = src + (size_t)(2 * y) * sw;
    const uint8_t* s1 = s0 + sw;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int x = min(x0 + i, dw - 1);
      packed |= (uint32_t)((s0[2 * x] + s0[2 * x + 1] + s1[2 * x] + s1[2 * x + 1] + 2) >> 2) << (8 * i);
    }
  } else {
    const int2 ty = ytab[y];
    const int sy0 = min(max(ty.x, 0), sh - 1), sy1 = min(max(ty.x + 1, 0), sh - 1);
    const int b0 = ty.y & 0xffff, b1 = ty.y >> 16;
    const uint8_t* r0 = src + (size_t)sy0 * sw;
    const uint8_t* r1 = src + (size_t)sy1 * sw;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int2 tx = xtab[min(x0 + i, dw - 1)];
      const int sx = tx.x, sx1 = min(sx + 1, sw - 1);
      const int a0 = tx.y & 0xffff, a1 = tx.y >> 16;
      const int h0 = r0[sx] * a0 + r0[sx1] * a1;
      const int h1 = r1[sx] * a0 + r1[sx1] * a1;
      const int v = (((b0 * (h0 >> 4)) >> 16) + ((b1 * (h1 >> 4)) >> 16) + 2) >> 2;
      packed |= (uint32_t)min(max(v, 0), 255) << (8 * i);
    }
  }
  // the destination pitch is a multiple of 16: the padded tail of a row may be written (it repeats the last pixel)
  *reinterpret_cast<uint32_t*>(dst + (size_t)frame * dframe + (size_t)y * dp + x0) = packed;
}
