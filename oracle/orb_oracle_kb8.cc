// TEST INFRASTRUCTURE ONLY (the checker; never linked into or called by the product).
// CPU restatement of the fisheye stereo triangulation in the CUDA kernel's formulation (flat float expressions, no matrix
// types): KannalaBrandt8::unproject / project / TriangulateMatches / Triangulate (reference
// src/CameraModels/KannalaBrandt8.cpp:116-147, :68-94, :323-395, :415-428) and the acceptance loop of
// Frame::ComputeStereoFishEyeMatches (src/Frame.cc:1244-1273). tests/test_oracle_kb8.py checks it bit for bit against
// oracle/_ref/libmorb_ref_kb8.so (the reference's own lines on the mini Eigen stand-in). Both share orb_eigen_jacobi_svd4f, Eigen's
// two-sided float Jacobi SVD restated from its published algorithm; Eigen itself is absent from this image, so that restatement is
// not pinned against the library. tanf / atan2f / cosf / sinf are this image's glibc (restated for the kernel in libm_restate.h).
#include <math.h>
#include <cmath>
#include <cstdint>
#include <cstring>

#include "orb_oracle_kb8.h"
#include "libm_restate.h"

namespace {
struct Rig {
  const float *cam1, *cam2, *R, *t;
  float prec1, prec2;
};

inline float sum3(float a, float b, float c) { return a + (b + c); }   // the stand-in's (= Eigen's unrolled) reduction order

void unproject(const float* P, float prec, float x, float y, float& rx, float& ry) {
  const float pwx = (x - P[2]) / P[0], pwy = (y - P[3]) / P[1];
  float scale = 1.f;
  float theta_d = sqrtf(pwx * pwx + pwy * pwy);
  theta_d = fminf(fmaxf(-3.1415926535897932384626433832795 / 2.f, theta_d), 3.1415926535897932384626433832795 / 2.f);
  if (theta_d > 1e-8) {
    float theta = theta_d;
    for (int j = 0; j < 10; j++) {
      const float theta2 = theta * theta, theta4 = theta2 * theta2, theta6 = theta4 * theta2, theta8 = theta4 * theta4;
      const float k0 = P[4] * theta2, k1 = P[5] * theta4, k2 = P[6] * theta6, k3 = P[7] * theta8;
      const float fix = (theta * (1 + k0 + k1 + k2 + k3) - theta_d) / (1 + 3 * k0 + 5 * k1 + 7 * k2 + 9 * k3);
      theta = theta - fix;
      if (fabsf(fix) < prec) break;
    }
    scale = tanf(theta) / theta_d;
  }
  rx = pwx * scale;
  ry = pwy * scale;
}

void project(const float* P, float X, float Y, float Z, float& u, float& v) {
  const float x2y2 = X * X + Y * Y;
  const float theta = atan2f(sqrtf(x2y2), Z);
  const float psi = atan2f(Y, X);
  const float theta2 = theta * theta, theta3 = theta * theta2, theta5 = theta3 * theta2, theta7 = theta5 * theta2, theta9 = theta7 * theta2;
  const float r = theta + P[4] * theta3 + P[5] * theta5 + P[6] * theta7 + P[7] * theta9;
  u = P[0] * r * cosf(psi) + P[2];
  v = P[1] * r * sinf(psi) + P[3];
}

// q (optional, 7 floats): the compared quantities cos, z1, z2, err1, 5.991 sigma1, err2, 5.991 sigma2 (NaN = not reached)
float triangulate(const Rig& g, float x1, float y1, float x2, float y2, float sigma1, float sigma2, float* p3d, float* q) {
  if (q) for (int i = 0; i < 7; ++i) q[i] = NAN;
  float r1x, r1y, r2x, r2y;
  unproject(g.cam1, g.prec1, x1, y1, r1x, r1y);
  unproject(g.cam2, g.prec2, x2, y2, r2x, r2y);
  const float* R = g.R;
  const float qx = sum3(R[0] * r2x, R[1] * r2y, R[2] * 1.f);
  const float qy = sum3(R[3] * r2x, R[4] * r2y, R[5] * 1.f);
  const float qz = sum3(R[6] * r2x, R[7] * r2y, R[8] * 1.f);
  const float n1 = sqrtf(sum3(r1x * r1x, r1y * r1y, 1.f * 1.f)), n2 = sqrtf(sum3(qx * qx, qy * qy, qz * qz));
  const float cosp = sum3(r1x * qx, r1y * qy, 1.f * qz) / (n1 * n2);
  if (q) q[0] = cosp;
  if (cosp > 0.9998) return -1.f;
  float T2[12];
  for (int i = 0; i < 3; ++i) {
    T2[4 * i + 0] = R[0 * 3 + i];
    T2[4 * i + 1] = R[1 * 3 + i];
    T2[4 * i + 2] = R[2 * 3 + i];
    T2[4 * i + 3] = sum3((-R[0 * 3 + i]) * g.t[0], (-R[1 * 3 + i]) * g.t[1], (-R[2 * 3 + i]) * g.t[2]);
  }
  float A[16];
  A[0] = r1x * 0.f - 1.f; A[1] = r1x * 0.f - 0.f; A[2] = r1x * 1.f - 0.f; A[3] = r1x * 0.f - 0.f;
  A[4] = r1y * 0.f - 0.f; A[5] = r1y * 0.f - 1.f; A[6] = r1y * 1.f - 0.f; A[7] = r1y * 0.f - 0.f;
  for (int j = 0; j < 4; ++j) {
    A[8 + j] = r2x * T2[8 + j] - T2[j];
    A[12 + j] = r2y * T2[8 + j] - T2[4 + j];
  }
  float V[16];
  orb_eigen_jacobi_svd4f(A, V);   // Eigen::JacobiSVD<Matrix4f>(A, ComputeFullV).matrixV(), float (orb_oracle_kb8.h)
  const float h0 = V[3], h1 = V[7], h2 = V[11], h3 = V[15];
  const float X = h0 / h3, Y = h1 / h3, Z = h2 / h3;
  const float z1 = Z;
  if (q) q[1] = z1;
  if (z1 <= 0) return -2.f;
  const float z2 = sum3(T2[8] * X, T2[9] * Y, T2[10] * Z) + T2[11];
  if (q) q[2] = z2;
  if (z2 <= 0) return -3.f;
  float u, v;
  project(g.cam1, X, Y, Z, u, v);
  const float e1x = u - x1, e1y = v - y1;
  if (q) { q[3] = e1x * e1x + e1y * e1y; q[4] = (float)(5.991 * sigma1); }
  if ((e1x * e1x + e1y * e1y) > 5.991 * sigma1) return -4.f;
  const float X2 = sum3(T2[0] * X, T2[1] * Y, T2[2] * Z) + T2[3];
  const float Y2 = sum3(T2[4] * X, T2[5] * Y, T2[6] * Z) + T2[7];
  const float Z2 = sum3(T2[8] * X, T2[9] * Y, T2[10] * Z) + T2[11];
  project(g.cam2, X2, Y2, Z2, u, v);
  const float e2x = u - x2, e2y = v - y2;
  if (q) { q[5] = e2x * e2x + e2y * e2y; q[6] = (float)(5.991 * sigma2); }
  if ((e2x * e2x + e2y * e2y) > 5.991 * sigma2) return -5.f;
  p3d[0] = X; p3d[1] = Y; p3d[2] = Z;
  return z1;
}
}  // namespace

struct OracleKp { float x, y, size, angle, response; int32_t octave, class_id; };

extern "C" {
void oracle_kb8_triangulate(const float* cam1, float prec1, const float* cam2, float prec2, const float* R12, const float* t12, const float* xy1,
                            const float* xy2, const float* s1, const float* s2, int n, float* ret, float* p3d, float* quantities) {
  const Rig g{cam1, cam2, R12, t12, prec1, prec2};
  for (int i = 0; i < n; ++i) {
    float X[3] = {0.f, 0.f, 0.f};
    ret[i] = triangulate(g, xy1[2 * i], xy1[2 * i + 1], xy2[2 * i], xy2[2 * i + 1], s1[i], s2[i], X, quantities ? quantities + 7 * i : nullptr);
    p3d[3 * i] = X[0]; p3d[3 * i + 1] = X[1]; p3d[3 * i + 2] = X[2];
  }
}
void oracle_kb8_unproject(const float* cam, float prec, const float* xy, int n, float* rays) {
  for (int i = 0; i < n; ++i) { unproject(cam, prec, xy[2 * i], xy[2 * i + 1], rays[3 * i], rays[3 * i + 1]); rays[3 * i + 2] = 1.f; }
}
void oracle_kb8_project(const float* cam, const float* xyz, int n, float* uv) {
  for (int i = 0; i < n; ++i) project(cam, xyz[3 * i], xyz[3 * i + 1], xyz[3 * i + 2], uv[2 * i], uv[2 * i + 1]);
}
void oracle_svd4_v(const float* A, double* V, double* sv) { orb_oracle_svd4_v(A, V, sv); }
void oracle_jacobi_svd4f(const float* A, float* V, float* sv) { orb_eigen_jacobi_svd4f(A, V, sv); }
// glibc's float routines against their restatements (the kernel's copies): tests/test_oracle_primitives.py
float restated_tanf(float x) { return libm_restate::tanf_r(x); }
float restated_atanf(float x) { return libm_restate::atanf_r(x); }
float restated_atan2f(float y, float x) { return libm_restate::atan2f_r(y, x); }
float libm_tanf(float x) { return tanf(x); }
float libm_atanf(float x) { return atanf(x); }
float libm_atan2f(float y, float x) { return atan2f(y, x); }
// counts the floats in [lo, hi) (bit patterns) whose restated tanf / atanf differs from libm's
long restated_tanf_mismatches(unsigned lo, unsigned hi, unsigned step) {
  long bad = 0;
  for (unsigned long u = lo; u < hi; u += step) { const float x = libm_restate::wf((int32_t)u); bad += libm_restate::fw(tanf(x)) != libm_restate::fw(libm_restate::tanf_r(x)); }
  return bad;
}
long restated_atanf_mismatches(unsigned lo, unsigned hi, unsigned step) {
  long bad = 0;
  for (unsigned long u = lo; u < hi; u += step) { const float x = libm_restate::wf((int32_t)u); bad += libm_restate::fw(atanf(x)) != libm_restate::fw(libm_restate::atanf_r(x)); }
  return bad;
}
// Frame::ComputeStereoFishEyeMatches from :1244: ratio gate, triangulation, depth > 0.0001f; code as in include/orb_b200.h,
// quantities (optional): 7 floats per LEFT keypoint
void oracle_fisheye_accept(const float* cam1, float prec1, const float* cam2, float prec2, const float* R12, const float* t12, const OracleKp* kL,
                           int nL, int monoL, const OracleKp* kR, int nR, int monoR, const float* sigma2, int nlev, const int* knn_idx,
                           const int* knn_dist, int nq, int* l2r, int* r2l, float* depth, float* p3d, int8_t* code, float* quantities) {
  const Rig g{cam1, cam2, R12, t12, prec1, prec2};
  for (int i = 0; i < nL; ++i) { l2r[i] = -1; depth[i] = -1.f; p3d[3 * i] = p3d[3 * i + 1] = p3d[3 * i + 2] = 0.f; code[i] = 0; }
  for (int i = 0; i < nR; ++i) r2l[i] = -1;
  if (quantities) for (int i = 0; i < 7 * nL; ++i) quantities[i] = NAN;
  for (int i = 0; i < nq; ++i) {
    if (knn_idx[2 * i] < 0 || knn_idx[2 * i + 1] < 0) continue;                              // (*it).size() >= 2
    if (!((float)knn_dist[2 * i] < (float)knn_dist[2 * i + 1] * 0.7)) continue;            // float * double -> double
    const int l = i + monoL, r = knn_idx[2 * i] + monoR;
    float X[3] = {0.f, 0.f, 0.f};
    const float d = triangulate(g, kL[l].x, kL[l].y, kR[r].x, kR[r].y, sigma2[kL[l].octave], sigma2[kR[r].octave], X,
                                quantities ? quantities + 7 * l : nullptr);
    if (d > 0.0001f) {
      l2r[l] = r; r2l[r] = l; depth[l] = d; code[l] = 1;
      p3d[3 * l] = X[0]; p3d[3 * l + 1] = X[1]; p3d[3 * l + 2] = X[2];
    } else {
      code[l] = d < 0.f ? (int8_t)d : (int8_t)-6;
    }
  }
}
}
