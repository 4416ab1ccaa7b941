// TEST INFRASTRUCTURE ONLY (the checker; never linked into or called by the product).
// Shared by the reference driver (oracle/ref_driver_kb8.cc through oracle/shim_eigen/mini_eigen.h) and the restatement
// (oracle/orb_oracle_kb8.cc): V of the singular value decomposition of a 4 x 4 float matrix, one-sided Jacobi (Hestenes) in
// double, columns ordered by descending singular value like Eigen::JacobiSVD::matrixV() (only the order matters to the caller:
// reference src/CameraModels/KannalaBrandt8.cpp:425-426 takes col(3)). Stand-in for Eigen, which is absent here: UNPINNED,
// checked against numpy.linalg.svd (LAPACK) in tests/test_oracle_kb8.py.
#pragma once
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
