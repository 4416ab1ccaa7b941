// TEST INFRASTRUCTURE ONLY (the checker; never linked into or called by the product).
// Shared by the reference driver (oracle/ref_driver_kb8.cc through oracle/shim_eigen/mini_eigen.h) and the restatement
// (oracle/orb_oracle_kb8.cc): V of the singular value decomposition of a 4 x 4 float matrix, one-sided Jacobi (Hestenes) in
// double, columns ordered by descending singular value like Eigen::JacobiSVD::matrixV() (only the order matters to the caller:
// reference src/CameraModels/KannalaBrandt8.cpp:425-426 takes col(3)). Stand-in for Eigen, which is absent here: UNPINNED,
// checked against numpy.linalg.svd (LAPACK) in tests/test_oracle_kb8.py.
#pragma once
#include <cfloat>
#include <cmath>

static inline void orb_oracle_svd4_v(const float* A /* row-major */, double* Vout /* row-major */, double* sv = nullptr) {
  double U[4][4], V[4][4];
  for (int i = 0; i < 4; ++i)
    for (int j = 0; j < 4; ++j) { U[i][j] = (double)A[4 * i + j]; V[i][j] = i == j ? 1.0 : 0.0; }
  for (int sweep = 0; sweep < 60; ++sweep) {
    bool rotated = false;
    for (int p = 0; p < 3; ++p)
      for (int q = p + 1; q < 4; ++q) {
        double a = 0.0, b = 0.0, c = 0.0;
        for (int i = 0; i < 4; ++i) { a += U[i][p] * U[i][p]; b += U[i][q] * U[i][q]; c += U[i][p] * U[i][q]; }
        if (c != 0.0 && std::fabs(c) > 1e-15 * std::sqrt(a * b)) {
          rotated = true;
          const double zeta = (b - a) / (2.0 * c);
          const double t = (zeta >= 0.0 ? 1.0 : -1.0) / (std::fabs(zeta) + std::sqrt(1.0 + zeta * zeta));
          const double cs = 1.0 / std::sqrt(1.0 + t * t), sn = cs * t;
          for (int i = 0; i < 4; ++i) {
            const double up = U[i][p], uq = U[i][q];
            U[i][p] = cs * up - sn * uq;
            U[i][q] = sn * up + cs * uq;
            const double vp = V[i][p], vq = V[i][q];
            V[i][p] = cs * vp - sn * vq;
            V[i][q] = sn * vp + cs * vq;
          }
        }
      }
    if (!rotated) break;
  }
  double n[4];
  int order[4] = {0, 1, 2, 3};
  for (int j = 0; j < 4; ++j) { n[j] = 0.0; for (int i = 0; i < 4; ++i) n[j] += U[i][j] * U[i][j]; }
  for (int i = 0; i < 4; ++i)   // stable selection by descending norm
    for (int j = i + 1; j < 4; ++j)
      if (n[order[j]] > n[order[i]]) { const int t = order[i]; order[i] = order[j]; order[j] = t; }
  for (int j = 0; j < 4; ++j) {
    for (int i = 0; i < 4; ++i) Vout[4 * i + j] = V[i][order[j]];
    if (sv) sv[j] = std::sqrt(n[order[j]]);
  }
}


// Eigen::JacobiSVD<Eigen::Matrix4f>(A, ComputeFullV).matrixV(), restated from Eigen's published algorithm (Eigen/src/SVD/JacobiSVD.h
// compute(), Eigen/src/misc/RealSvd2x2.h real_2x2_jacobi_svd, Eigen/src/Jacobi/Jacobi.h makeJacobi / apply_rotation_in_the_plane;
// the same text in 3.3.x and 3.4.0) in FLOAT, operation by operation: scaling by the largest |coefficient|, two-sided Jacobi sweeps
// over (p, q) = (1,0), (2,0), (2,1), (3,0), (3,1), (3,2) until every off-diagonal pair is below max(FLT_MIN, 2 eps maxDiag), singular
// values = |diagonal| * scale sorted in descending order with the columns of V swapped along. A square matrix takes no QR
// preconditioner. Rotations are applied as x' = c x + s y, y' = -s x + c y (two products, one sum, no contraction).
// Eigen itself is absent from this image, so this restatement is NOT pinned against the library; it is what the
// reference driver's stand-in, the CPU restatement and the CUDA kernel all run, bit for bit.
struct OrbJacobiRot { float c, s; };
static inline OrbJacobiRot orb_make_jacobi(float x, float y, float z) {
  OrbJacobiRot r;
  const float deno = 2.f * std::fabs(y);
  if (deno < FLT_MIN) { r.c = 1.f; r.s = 0.f; return r; }
  const float tau = (x - z) / deno;
  const float w = std::sqrt(tau * tau + 1.f);
  float t;
  if (tau > 0.f) t = 1.f / (tau + w);
  else t = 1.f / (tau - w);
  const float sign_t = t > 0.f ? 1.f : -1.f;
  const float n = 1.f / std::sqrt(t * t + 1.f);
  r.s = -sign_t * (y / std::fabs(y)) * std::fabs(t) * n;
  r.c = n;
  return r;
}
static inline void orb_rot_apply(float& x, float& y, float c, float s) {   // apply_rotation_in_the_plane
  if (c == 1.f && s == 0.f) return;
  const float xi = x, yi = y;
  x = c * xi + s * yi;
  y = -s * xi + c * yi;
}
static inline void orb_eigen_jacobi_svd4f(const float* A /* row-major */, float* Vout /* row-major */, float* sv = nullptr) {
  float W[4][4], V[4][4], S[4];
  float scale = 0.f;
  for (int i = 0; i < 16; ++i) scale = std::fmax(scale, std::fabs(A[i]));
  if (scale == 0.f) scale = 1.f;
  for (int i = 0; i < 4; ++i)
    for (int j = 0; j < 4; ++j) { W[i][j] = A[4 * i + j] / scale; V[i][j] = i == j ? 1.f : 0.f; }
  const float precision = 2.f * FLT_EPSILON, considerAsZero = FLT_MIN;
  float maxDiag = 0.f;
  for (int i = 0; i < 4; ++i) maxDiag = std::fmax(maxDiag, std::fabs(W[i][i]));
  bool finished = false;
  for (int sweep = 0; !finished && sweep < 1000; ++sweep) {
    finished = true;
    for (int p = 1; p < 4; ++p)
      for (int q = 0; q < p; ++q) {
        const float threshold = std::fmax(considerAsZero, precision * maxDiag);
        if (std::fabs(W[p][q]) > threshold || std::fabs(W[q][p]) > threshold) {
          finished = false;
          // real_2x2_jacobi_svd on m = [W(p,p) W(p,q); W(q,p) W(q,q)]
          float m00 = W[p][p], m01 = W[p][q], m10 = W[q][p], m11 = W[q][q];
          OrbJacobiRot rot1;
          const float t = m00 + m11, d = m10 - m01;
          if (std::fabs(d) < FLT_MIN) { rot1.s = 0.f; rot1.c = 1.f; }
          else {
            const float u = t / d;
            const float tmp = std::sqrt(1.f + u * u);
            rot1.s = 1.f / tmp;
            rot1.c = u / tmp;
          }
          orb_rot_apply(m00, m10, rot1.c, rot1.s);      // m.applyOnTheLeft(0, 1, rot1): rows 0 and 1, column by column
          orb_rot_apply(m01, m11, rot1.c, rot1.s);
          const OrbJacobiRot jr = orb_make_jacobi(m00, m01, m11);
          // j_left = rot1 * j_right.transpose(); transpose() = (c, -s)
          OrbJacobiRot jl;
          jl.c = rot1.c * jr.c - rot1.s * (-jr.s);
          jl.s = rot1.c * (-jr.s) + rot1.s * jr.c;
          for (int k = 0; k < 4; ++k) orb_rot_apply(W[p][k], W[q][k], jl.c, jl.s);      // W.applyOnTheLeft(p, q, j_left)
          for (int k = 0; k < 4; ++k) orb_rot_apply(W[k][p], W[k][q], jr.c, -jr.s);     // W.applyOnTheRight(p, q, j_right): with j.transpose()
          for (int k = 0; k < 4; ++k) orb_rot_apply(V[k][p], V[k][q], jr.c, -jr.s);     // V.applyOnTheRight(p, q, j_right)
          maxDiag = std::fmax(maxDiag, std::fmax(std::fabs(W[p][p]), std::fabs(W[q][q])));
        }
      }
  }
  for (int i = 0; i < 4; ++i) S[i] = std::fabs(W[i][i]) * scale;    // (m_singularValues *= scale)
  for (int i = 0; i < 4; ++i) {   // descending order, columns of V swapped along
    int pos = i;
    for (int j = i + 1; j < 4; ++j)
      if (S[j] > S[pos]) pos = j;
    if (S[pos] == 0.f) break;
    if (pos != i) {
      const float ts = S[i]; S[i] = S[pos]; S[pos] = ts;
      for (int k = 0; k < 4; ++k) { const float tv = V[k][i]; V[k][i] = V[k][pos]; V[k][pos] = tv; }
    }
  }
  for (int i = 0; i < 4; ++i)
    for (int j = 0; j < 4; ++j) Vout[4 * i + j] = V[i][j];
  if (sv) for (int i = 0; i < 4; ++i) sv[i] = S[i];
}
