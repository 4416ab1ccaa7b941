// TEST INFRASTRUCTURE ONLY (oracle). Restatement of glibc 2.39's float sinf/cosf for |x| < 120
// (sysdeps/ieee754/flt-32/s_sinf.c, s_cosf.c, sincosf.h; the ARM "optimized routines" algorithm):
// double-precision argument reduction by pi/2 followed by degree-7 (sine) / degree-8 (cosine)
// double polynomials, one rounding to float at the end.
//
// The reference calls cosf/sinf at src/ORBextractor.cc:104-105; the CUDA kernel cannot call glibc, so
// it carries a device copy of this restatement. tests/test_oracle_primitives.py pins this file
// against the libm of this image over every float in [0, 2*pi] (and a sample outside).
#pragma once
#include <cstdint>
#include <cstring>

namespace sincosf_restate {

struct Tab {
  double sign[4];
  double hpi_inv;  // 2/pi * 2^24
  double hpi;      // pi/2
  double c0, c1, c2, c3, c4;
  double s1, s2, s3;
};

static const Tab kTab[2] = {
    {{1.0, -1.0, -1.0, 1.0},
     0x1.45F306DC9C883p+23,
     0x1.921FB54442D18p0,
     0x1p0, -0x1.ffffffd0c621cp-2, 0x1.55553e1068f19p-5, -0x1.6c087e89a359dp-10, 0x1.99343027bf8c3p-16,
     -0x1.555545995a603p-3, 0x1.1107605230bc4p-7, -0x1.994eb3774cf24p-13},
    {{1.0, -1.0, -1.0, 1.0},
     0x1.45F306DC9C883p+23,
     0x1.921FB54442D18p0,
     -0x1p0, 0x1.ffffffd0c621cp-2, -0x1.55553e1068f19p-5, 0x1.6c087e89a359dp-10, -0x1.99343027bf8c3p-16,
     -0x1.555545995a603p-3, 0x1.1107605230bc4p-7, -0x1.994eb3774cf24p-13}};

static inline uint32_t abstop12(float x) {
  uint32_t u;
  std::memcpy(&u, &x, 4);
  return (u >> 20) & 0x7ff;
}

static inline float poly(double x, double x2, const Tab* p, int n) {
  if ((n & 1) == 0) {
    double x3 = x * x2;
    double s1 = p->s2 + x2 * p->s3;
    double x7 = x3 * x2;
    double s = x + x3 * p->s1;
    return (float)(s + x7 * s1);
  } else {
    double x4 = x2 * x2;
    double c2 = p->c3 + x2 * p->c4;
    double c1 = p->c0 + x2 * p->c1;
    double x6 = x4 * x2;
    double c = c1 + x4 * p->c2;
    return (float)(c + x6 * c2);
  }
}

static inline double reduce_fast(double x, const Tab* p, int* np) {
  double r = x * p->hpi_inv;
  int n = ((int32_t)r + 0x800000) >> 24;
  *np = n;
  return x - n * p->hpi;
}

// valid for |y| < 120 (the reference only passes angles in [0, 2*pi])
static inline float sinf_r(float y) {
  double x = y;
  const Tab* p = &kTab[0];
  if (abstop12(y) < abstop12(0x1.921FB6p-1f)) {  // |y| < pi/4
    double s = x * x;
    if (abstop12(y) < abstop12(0x1p-12f)) return y;
    return poly(x, s, p, 0);
  }
  int n;
  x = reduce_fast(x, p, &n);
  double s = p->sign[n & 3];
  if (n & 2) p = &kTab[1];
  return poly(x * s, x * x, p, n);
}

static inline float cosf_r(float y) {
  double x = y;
  const Tab* p = &kTab[0];
  if (abstop12(y) < abstop12(0x1.921FB6p-1f)) {
    double x2 = x * x;
    if (abstop12(y) < abstop12(0x1p-12f)) return 1.0f;
    return poly(x, x2, p, 1);
  }
  int n;
  x = reduce_fast(x, p, &n);
  double s = p->sign[n & 3];
  if (n & 2) p = &kTab[1];
  return poly(x * s, x * x, p, n ^ 1);
}

}  // namespace sincosf_restate
