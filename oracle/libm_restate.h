// TEST INFRASTRUCTURE ONLY (oracle). Restatement of glibc 2.39's float tanf / atanf / atan2f (the fdlibm-derived routines of
// sysdeps/ieee754/flt-32: k_tanf.c, s_tanf.c, e_rem_pio2f.c, s_atanf.c, e_atan2f.c) - float arithmetic, no contraction.
// KannalaBrandt8::unproject calls tanf (reference src/CameraModels/KannalaBrandt8.cpp:141), project calls atan2f (:70-71); the CUDA
// kernel cannot call glibc, so it carries a device copy of this file (morb_slam_b200/csrc/orb_libm_glibc.cuh).
// Pinned against the libm of this image (tests/test_oracle_primitives.py runs the sampled form; the exhaustive runs - tanf on
// all 1 075 235 812 floats of [0, 3 pi / 4), atanf on all 2 139 095 040 positive finite floats, atan2f on 2e8 random pairs: 0
// differences - are tools/probe/libm_check.cc).
#pragma once
#include <cstdint>
#include <cstring>
#include <cmath>
#ifndef LR_FN
#define LR_FN static inline
#endif
namespace libm_restate {
LR_FN int32_t fw(float x) { int32_t i; std::memcpy(&i, &x, 4); return i; }
LR_FN float wf(int32_t i) { float x; std::memcpy(&x, &i, 4); return x; }

LR_FN float kernel_tanf(float x, float y, int iy) {
  const float one = 1.0f, pio4 = 7.8539812565e-01f, pio4lo = 3.7748947079e-08f;
  const float T[13] = {3.3333334327e-01f, 1.3333334029e-01f, 5.3968254477e-02f, 2.1869488060e-02f, 8.8632395491e-03f, 3.5920790397e-03f,
                       1.4562094584e-03f, 5.8804126456e-04f, 2.4646313977e-04f, 7.8179444245e-05f, 7.1407252108e-05f, -1.8558637748e-05f,
                       2.5907305826e-05f};
  float z, r, v, w, s;
  const int32_t hx = fw(x), ix = hx & 0x7fffffff;
  if (ix < 0x39000000) {  // |x| < 2**-13
    if ((int)x == 0) {
      if ((ix | (iy + 1)) == 0) return one / fabsf(x);
      else if (iy == 1) return x;
      else return -one / x;
    }
  }
  if (ix >= 0x3f2ca140) {  // |x| >= 0.6744
    if (hx < 0) { x = -x; y = -y; }
    z = pio4 - x;
    w = pio4lo - y;
    x = z + w; y = 0.0f;
    if (fabsf(x) < 0x1p-13f) return (1 - ((hx >> 30) & 2)) * iy * (1.0f - 2 * iy * x);
  }
  z = x * x;
  w = z * z;
  r = T[1] + w * (T[3] + w * (T[5] + w * (T[7] + w * (T[9] + w * T[11]))));
  v = z * (T[2] + w * (T[4] + w * (T[6] + w * (T[8] + w * (T[10] + w * T[12])))));
  s = z * x;
  r = y + z * (s * (r + v) + y);
  r += T[0] * s;
  w = x + r;
  if (ix >= 0x3f2ca140) {
    v = (float)iy;
    return (float)(1 - ((hx >> 30) & 2)) * (v - 2.0f * (x - (w * w / (w + v) - r)));
  }
  if (iy == 1) return w;
  else {
    float a, t;
    z = wf(fw(w) & 0xfffff000);
    v = r - (z - x);
    t = a = -1.0f / w;
    t = wf(fw(t) & 0xfffff000);
    s = 1.0f + t * z;
    return t + a * (s + t * v);
  }
}

// valid for |x| < 3 pi / 4 (the reference calls tanf on an angle in [0, pi/2]); __ieee754_rem_pio2f reduces in double
// (x -+ pi/2 rounded once) and hands head + tail to the float kernel - verified against this image's libm on every float of the range
LR_FN float tanf_r(float x) {
  const int32_t hx = fw(x), ix = hx & 0x7fffffff;
  if (ix <= 0x3f490fda) return kernel_tanf(x, 0.0f, 1);
  const double yd = hx > 0 ? (double)x - 1.57079632679489655800e+00 : (double)x + 1.57079632679489655800e+00;
  const float y0 = (float)yd, y1 = (float)(yd - (double)y0);
  return kernel_tanf(y0, y1, -1);
}

LR_FN float atanf_r(float x) {
  const float atanhi[4] = {4.6364760399e-01f, 7.8539812565e-01f, 9.8279368877e-01f, 1.5707962513e+00f};
  const float atanlo[4] = {5.0121582440e-09f, 3.7748947079e-08f, 3.4473217170e-08f, 7.5497894159e-08f};
  const float aT[11] = {3.3333334327e-01f, -2.0000000298e-01f, 1.4285714924e-01f, -1.1111110449e-01f, 9.0908870101e-02f, -7.6918758452e-02f,
                        6.6610731184e-02f, -5.8335702866e-02f, 4.9768779427e-02f, -3.6531571299e-02f, 1.6285819933e-02f};
  const float one = 1.0f;
  float w, s1, s2, z;
  int id;
  const int32_t hx = fw(x), ix = hx & 0x7fffffff;
  if (ix >= 0x4c000000) {  // |x| >= 2^25
    if (ix > 0x7f800000) return x + x;
    if (hx > 0) return atanhi[3] + atanlo[3];
    else return -atanhi[3] - atanlo[3];
  }
  if (ix < 0x3ee00000) {  // |x| < 0.4375
    if (ix < 0x31000000) return x;  // |x| < 2^-29
    id = -1;
  } else {
    x = fabsf(x);
    if (ix < 0x3f980000) {
      if (ix < 0x3f300000) { id = 0; x = (2.0f * x - one) / (2.0f + x); }
      else { id = 1; x = (x - one) / (x + one); }
    } else {
      if (ix < 0x401c0000) { id = 2; x = (x - 1.5f) / (one + 1.5f * x); }
      else { id = 3; x = -1.0f / x; }
    }
  }
  z = x * x;
  w = z * z;
  s1 = z * (aT[0] + w * (aT[2] + w * (aT[4] + w * (aT[6] + w * (aT[8] + w * aT[10])))));
  s2 = w * (aT[1] + w * (aT[3] + w * (aT[5] + w * (aT[7] + w * aT[9]))));
  if (id < 0) return x - x * (s1 + s2);
  z = atanhi[id] - ((x * (s1 + s2) - atanlo[id]) - x);
  return (hx < 0) ? -z : z;
}

// finite arguments only
LR_FN float atan2f_r(float y, float x) {
  const float tiny = 1.0e-30f, pi_o_2 = 1.5707963705e+00f, pi = 3.1415927410e+00f, pi_lo = -8.7422776573e-08f;
  float z;
  const int32_t hx = fw(x), ix = hx & 0x7fffffff, hy = fw(y), iy = hy & 0x7fffffff;
  if (hx == 0x3f800000) return atanf_r(y);
  const int m = ((hy >> 31) & 1) | ((hx >> 30) & 2);
  if (iy == 0) {
    switch (m) {
      case 0: case 1: return y;
      case 2: return pi + tiny;
      default: return -pi - tiny;
    }
  }
  if (ix == 0) return (hy < 0) ? -pi_o_2 - tiny : pi_o_2 + tiny;
  const int k = (iy - ix) >> 23;
  if (k > 60) z = pi_o_2 + 0.5f * pi_lo;
  else if (hx < 0 && k < -60) z = 0.0f;
  else z = atanf_r(fabsf(y / x));
  switch (m) {
    case 0: return z;
    case 1: return wf(fw(z) ^ (int32_t)0x80000000);
    case 2: return pi - (z - pi_lo);
    default: return (z - pi_lo) - pi;
  }
}
}  // namespace libm_restate
