// 7x7 sigma-2 Gaussian blur of every pyramid level (reference src/ORBextractor.cc:1049-1050:
// GaussianBlur(level.clone(), Size(7,7), 2, 2, BORDER_REFLECT_101)), OpenCV's 8-bit fixed-point path:
// separable kernel [18 34 48 56 48 34 18] / 256, 16-bit horizontal intermediate (max 255*256), one rounding
// (+32768 >> 16) after the vertical pass. SURVEY.md Appendix A.2.
//
// Tiles of 128 x 32 outputs (tiles of all levels flattened into blockIdx.x, frames in blockIdx.y).
// Everything moves as 32-bit words: aligned word loads of the source (level rows are 16-byte aligned; only
// words that straddle the image border are assembled byte-wise with reflect-101), each thread produces 4
// adjacent outputs per step in both passes, stores are aligned words.
#pragma once

#define BLUR_TW 128
#define BLUR_TH 32
#define BLUR_RAW_WORDS 36   // 34 used: 4-byte left margin + 128 + 4, padded
#define BLUR_HS_WORDS 66    // 128 u16 = 64 words, padded

static __device__ __forceinline__ int reflect101(int p, int len) {
  if (p < 0) p = -p;
  if (p >= len) p = 2 * len - 2 - p;
  return p;
}

__global__ void __launch_bounds__(256) k_blur7(OrbGeom g, const uint8_t* __restrict__ pyr, uint8_t* __restrict__ blur) {
  __shared__ uint32_t raw[BLUR_TH + 6][BLUR_RAW_WORDS];
  __shared__ uint32_t hs[BLUR_TH + 6][BLUR_HS_WORDS];
  const int frame = blockIdx.y;
  int l = 0;
  while ((int)blockIdx.x >= g.blur_tile_start[l + 1]) ++l;
  const int t = blockIdx.x - g.blur_tile_start[l];
  const int tiles_x = g.blur_tiles_x[l];
  const int ty = t / tiles_x, tx = t - ty * tiles_x;
  const int W = g.w[l], H = g.h[l], P = g.pitch[l];
  const uint8_t* __restrict__ src = lvl_ptr(g, pyr, frame, l);
  uint8_t* dst = lvl_ptr(g, blur, frame, l);
  const int ox = tx * BLUR_TW, oy = ty * BLUR_TH;
  const int tid = threadIdx.x;

  // ---- stage rows oy-3 .. oy+34, columns ox-4 .. ox+131 (word j of a row starts at column ox - 4 + 4j)
  for (int i = tid; i < (BLUR_TH + 6) * 34; i += 256) {
    const int r = i / 34, j = i - r * 34;
    const int sy = reflect101(oy + r - 3, H);
    const int x = ox - 4 + 4 * j;
    const uint8_t* row = src + (size_t)sy * P;
    uint32_t w;
    if (x >= 0 && x + 3 < W) {
      w = *reinterpret_cast<const uint32_t*>(row + x);
    } else {
      // border word: reflect each column; columns further than the 3-px halo from the tile are never used
      w = 0;
#pragma unroll
      for (int b = 0; b < 4; ++b) {
        int xx = x + b;
        xx = min(max(xx, -(W - 1)), 2 * W - 2);
        w |= (uint32_t)row[reflect101(xx, W)] << (8 * b);
      }
    }
    raw[r][j] = w;
  }
  __syncthreads();

  // ---- horizontal pass: item = (row, quad of 4 outputs); outputs 4q..4q+3 need staged bytes 4q+1 .. 4q+10
  for (int i = tid; i < (BLUR_TH + 6) * 32; i += 256) {
    const int r = i >> 5, q = i & 31;
    const uint32_t w0 = raw[r][q], w1 = raw[r][q + 1], w2 = raw[r][q + 2];
    int b[12];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      b[k] = (w0 >> (8 * k)) & 0xff;
      b[4 + k] = (w1 >> (8 * k)) & 0xff;
      b[8 + k] = (w2 >> (8 * k)) & 0xff;
    }
    uint32_t o[4];
#pragma unroll
    for (int j = 0; j < 4; ++j)
      o[j] = 18 * (b[j + 1] + b[j + 7]) + 34 * (b[j + 2] + b[j + 6]) + 48 * (b[j + 3] + b[j + 5]) + 56 * b[j + 4];
    uint2 v;
    v.x = o[0] | (o[1] << 16);
    v.y = o[2] | (o[3] << 16);
    *reinterpret_cast<uint2*>(&hs[r][2 * q]) = v;
  }
  __syncthreads();

  // ---- vertical pass: thread = (group of 4 rows, quad of 4 columns)
  {
    const int rg = tid >> 5, q = tid & 31;
    const int r0 = rg * 4;
    const int x = ox + 4 * q;
    if (x < W) {
      int h[10][4];
#pragma unroll
      for (int k = 0; k < 10; ++k) {
        const uint2 v = *reinterpret_cast<const uint2*>(&hs[r0 + k][2 * q]);
        h[k][0] = v.x & 0xffff; h[k][1] = v.x >> 16; h[k][2] = v.y & 0xffff; h[k][3] = v.y >> 16;
      }
#pragma unroll
      for (int r = 0; r < 4; ++r) {
        const int y = oy + r0 + r;
        if (y < H) {
          uint32_t packed = 0;
#pragma unroll
          for (int c = 0; c < 4; ++c) {
            const int acc = 18 * (h[r][c] + h[r + 6][c]) + 34 * (h[r + 1][c] + h[r + 5][c]) + 48 * (h[r + 2][c] + h[r + 4][c]) +
                            56 * h[r + 3][c];
            packed |= (uint32_t)((acc + 32768) >> 16) << (8 * c);
          }
          // pitch is a multiple of 16 and x of 4: the padded tail of a row may be overwritten freely
          *reinterpret_cast<uint32_t*>(dst + (size_t)y * P + x) = packed;
        }
      }
    }
  }
}
