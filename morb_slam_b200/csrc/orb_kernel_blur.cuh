// 7x7 sigma-2 Gaussian blur of every pyramid level (reference src/ORBextractor.cc:1049-1050:
// GaussianBlur(level.clone(), Size(7,7), 2, 2, BORDER_REFLECT_101)), OpenCV's 8-bit fixed-point path:
// separable kernel [18 34 48 56 48 34 18] / 256, 16-bit horizontal intermediate (max 255*256), one rounding
// (+32768 >> 16) after the vertical pass. SURVEY.md Appendix A.2. All sums are exact integers.
//
// Tiles of 128 x 64 outputs (tiles of all levels flattened into blockIdx.x, frames in blockIdx.y).
//   load        ONE TMA tensor copy per CTA: box of 160 x 70 bytes at (ox - 16, oy - 3) (16-byte aligned start);
//               pixels outside the image are then patched in shared memory with REFLECT_101 (edge tiles only);
//   horizontal  item = (row, 4 outputs): the 7 taps of the 4 outputs are byte dot products of the three source words
//               with constant coefficient words (IDP.4A: 10 per 4 outputs, no unpacking, no shifts - the work runs on
//               the FMA pipe, which the byte-SIMD kernels of this path leave idle); results stored as 4 x 32 bit;
//   vertical    thread = 4 columns x 8 rows: 14 intermediate rows in registers, symmetric taps
//               (3 adds + 4 multiply-adds per output; the rounding constant rides in the horizontal sums), one 32-bit store per row.
#pragma once

#define BLUR_TW 128
#define BLUR_TH 64
#define BLUR_TP 160                  // raw tile pitch = TMA box width: 16 + 128 + 16
#define BLUR_TR (BLUR_TH + 6)        // raw tile rows
#define BLUR_ROWS 8                  // output rows per thread (8 warps x 8 rows)
#define BLUR_SMEM (BLUR_TR * BLUR_TP + BLUR_TR * 32 * 16 + 16)

static __device__ __forceinline__ int reflect101(int p, int len) {
  if (p < 0) p = -p;
  if (p >= len) p = 2 * len - 2 - p;
  return p;
}

__global__ void __launch_bounds__(256) k_blur7(const __grid_constant__ BlurMaps maps, OrbGeom g, const uint32_t* __restrict__ tile_tab,
                                              uint8_t* __restrict__ blur) {
  extern __shared__ __align__(128) uint8_t s_bl[];
  uint8_t* raw = s_bl;
  uint4* hs = reinterpret_cast<uint4*>(s_bl + BLUR_TR * BLUR_TP);          // [BLUR_TR][32] horizontal sums of 4 columns
  uint64_t* bar = reinterpret_cast<uint64_t*>(s_bl + BLUR_TR * BLUR_TP + BLUR_TR * 32 * 16);
  const int frame = blockIdx.y;
  const uint32_t tcode = tile_tab[blockIdx.x];   // level | tile column << 4 | tile row << 16
  const int l = tcode & 15, tx = (tcode >> 4) & 0xfff, ty = tcode >> 16;
  const int W = g.w[l], H = g.h[l], P = g.pitch[l];
  const int ox = tx * BLUR_TW, oy = ty * BLUR_TH;
  const int tid = threadIdx.x;

  // ---- stage rows oy-3 .. oy+66, columns ox-16 .. ox+143 (tile byte c of a row = image column ox - 16 + c)
  if (tid == 0) tma_load_tile(raw, &maps.m[l], ox - 16, frame * H + oy - 3, bar, BLUR_TR * BLUR_TP);
  __syncthreads();
  tma_wait(bar);
  // rows the tile really needs (the last tile of a level may be short)
  const int nrows = min(BLUR_TH, H - oy) + 6;
  {
    // REFLECT_101 patch of the pixels outside the image (edge tiles only): up to 3 rows above / below and
    // 3 columns left / right of it; only the candidates that exist for this tile are enumerated
    const int rb = H - oy + 3, cb = W - ox + 16;      // first tile row / column past the image
    const int top = oy == 0 ? 3 : 0, nfr = top + (rb < nrows ? 3 : 0);
    const int left = ox == 0 ? 3 : 0, nfc = left + (cb < BLUR_TP - 13 ? 3 : 0);
    const int nfix = nfr * BLUR_TP + nrows * nfc;
    if (nfix) {
      for (int i = tid; i < nfix; i += 256) {
        int r, c;
        if (i < nfr * BLUR_TP) {
          const int rr = i / BLUR_TP;
          c = i - rr * BLUR_TP;
          r = rr < top ? rr : rb + rr - top;
        } else {
          const int j = i - nfr * BLUR_TP;
          r = j / nfc;
          const int cc = j - r * nfc;
          c = cc < left ? 13 + cc : cb + cc - left;
        }
        const int y = oy - 3 + r, x = ox - 16 + c;
        if (r < nrows && (y < 0 || y >= H || x < 0 || x >= W)) {
          const int sy = reflect101(y, H) - oy + 3, sx = reflect101(min(max(x, -(W - 1)), 2 * W - 2), W) - ox + 16;
          if (sy >= 0 && sy < BLUR_TR && sx >= 0 && sx < BLUR_TP) raw[r * BLUR_TP + c] = raw[sy * BLUR_TP + sx];
        }
      }
      __syncthreads();
    }
  }

  // ---- horizontal pass: item = (row r, quad q); outputs 4q..4q+3 sit at tile bytes 16+4q .. 19+4q = word 4+q, their taps
  //      are bytes 1..10 of the words 3+q, 4+q, 5+q. Coefficient words: byte i multiplies byte i of the source word.
  {
    const uint32_t* raw_w = reinterpret_cast<const uint32_t*>(raw);
    for (int i = tid; i < nrows * 32; i += 256) {
      const int r = i >> 5, q = i & 31;
      const uint32_t* w = raw_w + r * (BLUR_TP / 4) + 3 + q;
      const uint32_t w0 = w[0], w1 = w[1], w2 = w[2];
      uint4 v;
      // every sum starts at 128: the vertical taps add up to 256, so the final rounding constant 32768 is already inside
      v.x = __dp4a(w1, 0x12223038u, __dp4a(w0, 0x30221200u, 128u));                              // taps b1..b7
      v.y = __dp4a(w2, 0x00000012u, __dp4a(w1, 0x22303830u, __dp4a(w0, 0x22120000u, 128u)));     // b2..b8
      v.z = __dp4a(w2, 0x00001222u, __dp4a(w1, 0x30383022u, __dp4a(w0, 0x12000000u, 128u)));     // b3..b9
      v.w = __dp4a(w2, 0x00122230u, __dp4a(w1, 0x38302212u, 128u));                              // b4..b10
      hs[r * 32 + q] = v;
    }
  }
  __syncthreads();

  // ---- vertical pass: thread = (strip of 8 rows, quad of 4 columns)
  {
    const int strip = tid >> 5, q = tid & 31;
    const int r0 = strip * BLUR_ROWS;
    const int x = ox + 4 * q;
    if (x < W && oy + r0 < H) {
      int h[BLUR_ROWS + 6][4];
#pragma unroll
      for (int k = 0; k < BLUR_ROWS + 6; ++k) {
        // rows past the tile's last needed row are never used by a stored output; clamp the index to stay in the tile
        const uint4 v = hs[min(r0 + k, BLUR_TR - 1) * 32 + q];
        h[k][0] = v.x; h[k][1] = v.y; h[k][2] = v.z; h[k][3] = v.w;
      }
      uint8_t* dst = blur + g.level_base[l] + (size_t)frame * g.level_fstride[l] + x;
#pragma unroll
      for (int r = 0; r < BLUR_ROWS; ++r) {
        const int y = oy + r0 + r;
        uint32_t acc[4];
#pragma unroll
        for (int c = 0; c < 4; ++c)
          acc[c] = 18 * (h[r][c] + h[r + 6][c]) + 34 * (h[r + 1][c] + h[r + 5][c]) + 48 * (h[r + 2][c] + h[r + 4][c]) + 56 * h[r + 3][c];
        // byte 2 of every accumulator is the rounded result (acc < 2^24, the rounding constant came with the horizontal sums)
        const uint32_t p01 = __byte_perm(acc[0], acc[1], 0x0062), p23 = __byte_perm(acc[2], acc[3], 0x0062);
        // pitch is a multiple of 16 and x of 4: the padded tail of a row may be overwritten freely
        if (y < H) *reinterpret_cast<uint32_t*>(dst + (size_t)y * P) = __byte_perm(p01, p23, 0x5410);
      }
    }
  }
}
