// 7x7 sigma-2 Gaussian blur of every pyramid level (reference src/ORBextractor.cc:1049-1050:
// GaussianBlur(level.clone(), Size(7,7), 2, 2, BORDER_REFLECT_101)), OpenCV's 8-bit fixed-point path:
// separable kernel [18 34 48 56 48 34 18] / 256, 16-bit horizontal intermediate (max 255*256), one rounding
// (+32768 >> 16) after the vertical pass. SURVEY.md Appendix A.2. All sums are exact integers.
//
// Tiles of 128 x 64 outputs (tiles of all levels flattened into blockIdx.x, frames in blockIdx.y).
//   load        ONE TMA tensor copy per CTA: box of 160 x 70 bytes at (ox - 16, oy - 3) (16-byte aligned start);
//               pixels outside the image are then patched in shared memory with REFLECT_101 (edge tiles only);
//   horizontal  item = (row pair, 4 outputs): the 7 taps of the 4 outputs are byte dot products of the three source words
//               with constant coefficient words (IDP.4A: 10 per 4 outputs, no unpacking, no shifts - the work runs on
//               the FMA pipe, which the byte-SIMD kernels of this path leave idle); the sums fit 16 bits (<= 255 * 256 + 128)
//               and are stored as VERTICAL pairs: word = row 2p | row 2p + 1 << 16 of one column;
//   vertical    thread = 4 columns x 8 rows: 7 row pairs in registers; an output is 4 two-way dot products (IDP.2A) of
//               row pairs with coefficient byte pairs - even rows take the pairs from their own row with (18,34) (48,56)
//               (48,34) (18,0), odd rows from the row above with (0,18) (34,48) (56,48) (34,18): the .lo / .hi halves of
//               four constant words (4 instructions per output instead of 3 adds + 4 multiply-adds; the rounding
//               constant rides in the horizontal sums), one 32-bit store per row.
#pragma once

#define BLUR_TW 128
#define BLUR_TH 64
#define BLUR_TP 160                  // raw tile pitch = TMA box width: 16 + 128 + 16
#define BLUR_TR (BLUR_TH + 6)        // raw tile rows
#define BLUR_ROWS 8                  // output rows per thread (8 warps x 8 rows)
#define BLUR_PAIRS (BLUR_TR / 2)      // vertical pairs of horizontal sums
#define BLUR_SMEM (BLUR_TR * BLUR_TP + BLUR_PAIRS * 32 * 16 + 16)

// 16-bit pair (two rows of one column) times a byte pair of coefficients, accumulated
static __device__ __forceinline__ uint32_t dp2a_lo_acc(uint32_t pair16, uint32_t coef8, uint32_t acc) {
  uint32_t d;
  asm("dp2a.lo.u32.u32 %0, %1, %2, %3;" : "=r"(d) : "r"(pair16), "r"(coef8), "r"(acc));
  return d;
}
static __device__ __forceinline__ uint32_t dp2a_hi_acc(uint32_t pair16, uint32_t coef8, uint32_t acc) {
  uint32_t d;
  asm("dp2a.hi.u32.u32 %0, %1, %2, %3;" : "=r"(d) : "r"(pair16), "r"(coef8), "r"(acc));
  return d;
}

static __device__ __forceinline__ int reflect101(int p, int len) {
  if (p < 0) p = -p;
  if (p >= len) p = 2 * len - 2 - p;
  return p;
}

__global__ void __launch_bounds__(256) k_blur7(const __grid_constant__ BlurMaps maps, OrbGeom g, const uint32_t* __restrict__ tile_tab,
                                              uint8_t* __restrict__ blur) {
  extern __shared__ __align__(128) uint8_t s_bl[];
  uint8_t* raw = s_bl;
  uint4* hs = reinterpret_cast<uint4*>(s_bl + BLUR_TR * BLUR_TP);          // [BLUR_PAIRS][32] horizontal sums of 4 columns x 2 rows
  uint64_t* bar = reinterpret_cast<uint64_t*>(s_bl + BLUR_TR * BLUR_TP + BLUR_PAIRS * 32 * 16);
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

  // ---- horizontal pass: item = (row pair p, quad q); outputs 4q..4q+3 sit at tile bytes 16+4q .. 19+4q = word 4+q, their taps
  //      are bytes 1..10 of the words 3+q, 4+q, 5+q. Coefficient words: byte i multiplies byte i of the source word.
  {
    const uint32_t* raw_w = reinterpret_cast<const uint32_t*>(raw);
    const int npairs = (nrows + 1) >> 1;   // an odd last row pairs with a tile row no stored output uses
    for (int i = tid; i < npairs * 32; i += 256) {
      const int p = i >> 5, q = i & 31;
      const uint32_t* w = raw_w + 2 * p * (BLUR_TP / 4) + 3 + q;
      uint4 v[2];
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const uint32_t w0 = w[e * (BLUR_TP / 4)], w1 = w[e * (BLUR_TP / 4) + 1], w2 = w[e * (BLUR_TP / 4) + 2];
        // every sum starts at 128: the vertical taps add up to 256, so the final rounding constant 32768 is already inside
        v[e].x = __dp4a(w1, 0x12223038u, __dp4a(w0, 0x30221200u, 128u));                              // taps b1..b7
        v[e].y = __dp4a(w2, 0x00000012u, __dp4a(w1, 0x22303830u, __dp4a(w0, 0x22120000u, 128u)));     // b2..b8
        v[e].z = __dp4a(w2, 0x00001222u, __dp4a(w1, 0x30383022u, __dp4a(w0, 0x12000000u, 128u)));     // b3..b9
        v[e].w = __dp4a(w2, 0x00122230u, __dp4a(w1, 0x38302212u, 128u));                              // b4..b10
      }
      uint4 o;
      o.x = __byte_perm(v[0].x, v[1].x, 0x5410); o.y = __byte_perm(v[0].y, v[1].y, 0x5410);
      o.z = __byte_perm(v[0].z, v[1].z, 0x5410); o.w = __byte_perm(v[0].w, v[1].w, 0x5410);
      hs[p * 32 + q] = o;
    }
  }
  __syncthreads();

  // ---- vertical pass: thread = (strip of 8 rows, quad of 4 columns)
  {
    const int strip = tid >> 5, q = tid & 31;
    const int r0 = strip * BLUR_ROWS;
    const int x = ox + 4 * q;
    if (x < W && oy + r0 < H) {
      uint32_t h[BLUR_ROWS / 2 + 3][4];     // pairs of tile rows (r0 + 2k, r0 + 2k + 1), k = 0..6
#pragma unroll
      for (int k = 0; k < BLUR_ROWS / 2 + 3; ++k) {
        const uint4 v = hs[(r0 / 2 + k) * 32 + q];
        h[k][0] = v.x; h[k][1] = v.y; h[k][2] = v.z; h[k][3] = v.w;
      }
      uint8_t* dst = blur + g.level_base[l] + (size_t)frame * g.level_fstride[l] + x;
      // coefficient byte pairs: .lo = even output row (taps start at the pair's first row), .hi = odd (at its second)
      const uint32_t C0 = 0x12002212u, C1 = 0x30223830u, C2 = 0x30382230u, C3 = 0x12220012u;
#pragma unroll
      for (int r = 0; r < BLUR_ROWS; ++r) {
        const int y = oy + r0 + r;
        const int m = r >> 1;
        uint32_t acc[4];
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          if (r & 1) acc[c] = dp2a_hi_acc(h[m + 3][c], C3, dp2a_hi_acc(h[m + 2][c], C2, dp2a_hi_acc(h[m + 1][c], C1, dp2a_hi_acc(h[m][c], C0, 0u))));
          else       acc[c] = dp2a_lo_acc(h[m + 3][c], C3, dp2a_lo_acc(h[m + 2][c], C2, dp2a_lo_acc(h[m + 1][c], C1, dp2a_lo_acc(h[m][c], C0, 0u))));
        }
        // byte 2 of every accumulator is the rounded result (acc < 2^24, the rounding constant came with the horizontal sums)
        const uint32_t p01 = __byte_perm(acc[0], acc[1], 0x0062), p23 = __byte_perm(acc[2], acc[3], 0x0062);
        // pitch is a multiple of 16 and x of 4: the padded tail of a row may be overwritten freely
        if (y < H) *reinterpret_cast<uint32_t*>(dst + (size_t)y * P) = __byte_perm(p01, p23, 0x5410);
      }
    }
  }
}
