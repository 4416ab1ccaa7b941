// KannalaBrandt8 on the device (reference src/CameraModels/KannalaBrandt8.cpp): unproject (:116-147), project (:68-94), Triangulate
// (:415-428, Eigen's two-sided float Jacobi SVD restated) and TriangulateMatches (:323-395). Shared by the fisheye stereo matcher
// (orb_fisheye.cu) and the two-camera SearchForTriangulation (orb_mapping.cu: KannalaBrandt8::epipolarConstrain :229-236 is
// TriangulateMatches(...) > 0.0001f). Float expressions in the reference's order with plain operators (-fmad=false: nothing
// contracts); libm calls are glibc's routines restated (orb_libm_glibc.cuh). See orb_fisheye.cu for what is and is not pinned.
#pragma once
#include <cfloat>

#include "orb_internal.h"
#include "orb_libm_glibc.cuh"

struct Kb8RigDev {
  float cam1[8], cam2[8];
  float prec1, prec2;
  float R12[9], t12[3];
  float sig1[ORB_MAX_LEVELS], sig2[ORB_MAX_LEVELS];   // mvLevelSigma2 of the left / right extractor
};

// a fixed-size reduction of three terms in the order the oracle's Eigen stand-in uses (oracle/shim_eigen/mini_eigen.h)
static __device__ __forceinline__ float sum3(float a, float b, float c) { return a + (b + c); }

// KannalaBrandt8::unproject (:116-147)
static __device__ __forceinline__ void kb8_unproject(const float* P, float prec, float x, float y, float& rx, float& ry) {
  const float pwx = (x - P[2]) / P[0], pwy = (y - P[3]) / P[1];
  float scale = 1.f;
  float theta_d = sqrtf(pwx * pwx + pwy * pwy);
  const float hp = (float)(3.1415926535897932384626433832795 / 2.0);
  theta_d = fminf(fmaxf(-hp, theta_d), hp);
  if ((double)theta_d > 1e-8) {
    float theta = theta_d;
    for (int j = 0; j < 10; ++j) {
      const float theta2 = theta * theta, theta4 = theta2 * theta2, theta6 = theta4 * theta2, theta8 = theta4 * theta4;
      const float k0 = P[4] * theta2, k1 = P[5] * theta4, k2 = P[6] * theta6, k3 = P[7] * theta8;
      const float fix = (theta * (1 + k0 + k1 + k2 + k3) - theta_d) / (1 + 3 * k0 + 5 * k1 + 7 * k2 + 9 * k3);
      theta = theta - fix;
      if (fabsf(fix) < prec) break;
    }
    // glibc's tanf (restated); the restatement covers |theta| < 3 pi / 4, far beyond what a converged theta <= pi / 2 can reach
    scale = (fabsf(theta) < 2.35f ? dev_libm::tanf_r(theta) : tanf(theta)) / theta_d;
  }
  rx = pwx * scale;
  ry = pwy * scale;
}

// KannalaBrandt8::project(const Eigen::Vector3f&) (:68-94)
static __device__ __forceinline__ void kb8_project(const float* P, float X, float Y, float Z, float& u, float& v) {
  const float x2y2 = X * X + Y * Y;
  const float theta = dev_libm::atan2f_r(sqrtf(x2y2), Z);
  const float psi = dev_libm::atan2f_r(Y, X);
  float sin_psi, cos_psi;
  dev_glibc_sincosf(psi, &sin_psi, &cos_psi);
  const float theta2 = theta * theta, theta3 = theta * theta2, theta5 = theta3 * theta2, theta7 = theta5 * theta2, theta9 = theta7 * theta2;
  const float r = theta + P[4] * theta3 + P[5] * theta5 + P[6] * theta7 + P[7] * theta9;
  u = P[0] * r * cos_psi + P[2];
  v = P[1] * r * sin_psi + P[3];
}

// Eigen::JacobiSVD<Matrix4f>(A, ComputeFullV).matrixV().col(3), Eigen's algorithm restated in float operation by operation
// (Eigen/src/SVD/JacobiSVD.h compute(), src/misc/RealSvd2x2.h, src/Jacobi/Jacobi.h; host twin: oracle/orb_oracle_kb8.h
// orb_eigen_jacobi_svd4f): scale by the largest |coefficient|, two-sided Jacobi sweeps over (p, q) = (1,0) (2,0) (2,1) (3,0) (3,1)
// (3,2) until every off-diagonal pair is below max(FLT_MIN, 2 eps maxDiag), singular values sorted in descending order with the
// columns of V swapped along. Rotation: x' = c x + s y, y' = -s x + c y.
static __device__ __forceinline__ void kb8_rot(float& x, float& y, float c, float s) {
  if (c == 1.f && s == 0.f) return;
  const float xi = x, yi = y;
  x = c * xi + s * yi;
  y = -s * xi + c * yi;
}
static __device__ void kb8_jacobi_v3(const float* A, float* x) {
  float W[4][4], V[4][4], S[4];
  float scale = 0.f;
#pragma unroll
  for (int i = 0; i < 16; ++i) scale = fmaxf(scale, fabsf(A[i]));
  if (scale == 0.f) scale = 1.f;
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) { W[i][j] = A[4 * i + j] / scale; V[i][j] = i == j ? 1.f : 0.f; }
  const float precision = 2.f * FLT_EPSILON, considerAsZero = FLT_MIN;
  float maxDiag = 0.f;
#pragma unroll
  for (int i = 0; i < 4; ++i) maxDiag = fmaxf(maxDiag, fabsf(W[i][i]));
  bool finished = false;
  for (int sweep = 0; !finished && sweep < 1000; ++sweep) {
    finished = true;
#pragma unroll
    for (int p = 1; p < 4; ++p)
#pragma unroll
      for (int q = 0; q < p; ++q) {
        const float threshold = fmaxf(considerAsZero, precision * maxDiag);
        if (fabsf(W[p][q]) > threshold || fabsf(W[q][p]) > threshold) {
          finished = false;
          float m00 = W[p][p], m01 = W[p][q], m10 = W[q][p], m11 = W[q][q];
          float r1c, r1s;
          const float t = m00 + m11, d = m10 - m01;
          if (fabsf(d) < FLT_MIN) { r1s = 0.f; r1c = 1.f; }
          else {
            const float u = t / d;
            const float tmp = sqrtf(1.f + u * u);
            r1s = 1.f / tmp;
            r1c = u / tmp;
          }
          kb8_rot(m00, m10, r1c, r1s);
          kb8_rot(m01, m11, r1c, r1s);
          float jc, js;   // makeJacobi(m00, m01, m11)
          const float deno = 2.f * fabsf(m01);
          if (deno < FLT_MIN) { jc = 1.f; js = 0.f; }
          else {
            const float tau = (m00 - m11) / deno;
            const float w = sqrtf(tau * tau + 1.f);
            float tt;
            if (tau > 0.f) tt = 1.f / (tau + w);
            else tt = 1.f / (tau - w);
            const float sign_t = tt > 0.f ? 1.f : -1.f;
            const float n = 1.f / sqrtf(tt * tt + 1.f);
            js = -sign_t * (m01 / fabsf(m01)) * fabsf(tt) * n;
            jc = n;
          }
          const float lc = r1c * jc - r1s * (-js), ls = r1c * (-js) + r1s * jc;   // j_left = rot1 * j_right.transpose()
#pragma unroll
          for (int k = 0; k < 4; ++k) kb8_rot(W[p][k], W[q][k], lc, ls);
#pragma unroll
          for (int k = 0; k < 4; ++k) kb8_rot(W[k][p], W[k][q], jc, -js);
#pragma unroll
          for (int k = 0; k < 4; ++k) kb8_rot(V[k][p], V[k][q], jc, -js);
          maxDiag = fmaxf(maxDiag, fmaxf(fabsf(W[p][p]), fabsf(W[q][q])));
        }
      }
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) S[i] = fabsf(W[i][i]) * scale;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    int pos = i;
#pragma unroll
    for (int j = i + 1; j < 4; ++j)
      if (j > i && S[j] > S[pos]) pos = j;
    if (S[pos] == 0.f) break;
    if (pos != i) {
      const float ts = S[i]; S[i] = S[pos]; S[pos] = ts;
#pragma unroll
      for (int k = 0; k < 4; ++k) { const float tv = V[k][i]; V[k][i] = V[k][pos]; V[k][pos] = tv; }
    }
  }
#pragma unroll
  for (int k = 0; k < 4; ++k) x[k] = V[k][3];
}

// KannalaBrandt8::TriangulateMatches (:323-395) with Triangulate (:415-428); returns the depth or -1 .. -5
// cam1 / cam2: the 8 parameters of this camera / pCamera2, R = R12 row-major, t = t12
static __device__ float kb8_triangulate_p(const float* cam1, float prec1, const float* cam2, float prec2, const float* R, const float* t12,
                                          float x1, float y1, float x2, float y2, float sigma1, float sigma2, float* p3d) {
  float r1x, r1y, r2x, r2y;
  kb8_unproject(cam1, prec1, x1, y1, r1x, r1y);
  kb8_unproject(cam2, prec2, x2, y2, r2x, r2y);
  // r21 = R12 * r2, rays have z = 1
  const float qx = sum3(R[0] * r2x, R[1] * r2y, R[2] * 1.f);
  const float qy = sum3(R[3] * r2x, R[4] * r2y, R[5] * 1.f);
  const float qz = sum3(R[6] * r2x, R[7] * r2y, R[8] * 1.f);
  const float n1 = sqrtf(sum3(r1x * r1x, r1y * r1y, 1.f * 1.f)), n2 = sqrtf(sum3(qx * qx, qy * qy, qz * qz));
  const float cosp = sum3(r1x * qx, r1y * qy, 1.f * qz) / (n1 * n2);
  if ((double)cosp > 0.9998) return -1.f;
  // Tcw1 = [I | 0], Tcw2 = [R21 | -R21 * t12], R21 = R12^T
  float T2[12];
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    T2[4 * i + 0] = R[0 * 3 + i];
    T2[4 * i + 1] = R[1 * 3 + i];
    T2[4 * i + 2] = R[2 * 3 + i];
    T2[4 * i + 3] = sum3((-R[0 * 3 + i]) * t12[0], (-R[1 * 3 + i]) * t12[1], (-R[2 * 3 + i]) * t12[2]);
  }
  float A[16];
  A[0] = r1x * 0.f - 1.f; A[1] = r1x * 0.f - 0.f; A[2] = r1x * 1.f - 0.f; A[3] = r1x * 0.f - 0.f;
  A[4] = r1y * 0.f - 0.f; A[5] = r1y * 0.f - 1.f; A[6] = r1y * 1.f - 0.f; A[7] = r1y * 0.f - 0.f;
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    A[8 + j] = r2x * T2[8 + j] - T2[j];
    A[12 + j] = r2y * T2[8 + j] - T2[4 + j];
  }
  float xh[4];
  kb8_jacobi_v3(A, xh);
  const float h3 = xh[3];   // x3D = x3D_h.head(3) / x3D_h(3) (:426-427)
  const float X = xh[0] / h3, Y = xh[1] / h3, Z = xh[2] / h3;
  const float z1 = Z;
  if (z1 <= 0) return -2.f;
  const float z2 = sum3(T2[8] * X, T2[9] * Y, T2[10] * Z) + T2[11];
  if (z2 <= 0) return -3.f;
  float u, v;
  kb8_project(cam1, X, Y, Z, u, v);
  const float e1x = u - x1, e1y = v - y1;
  if ((double)(e1x * e1x + e1y * e1y) > 5.991 * (double)sigma1) return -4.f;
  const float X2 = sum3(T2[0] * X, T2[1] * Y, T2[2] * Z) + T2[3];
  const float Y2 = sum3(T2[4] * X, T2[5] * Y, T2[6] * Z) + T2[7];
  const float Z2 = sum3(T2[8] * X, T2[9] * Y, T2[10] * Z) + T2[11];
  kb8_project(cam2, X2, Y2, Z2, u, v);
  const float e2x = u - x2, e2y = v - y2;
  if ((double)(e2x * e2x + e2y * e2y) > 5.991 * (double)sigma2) return -5.f;
  p3d[0] = X; p3d[1] = Y; p3d[2] = Z;
  return z1;
}

static __device__ __forceinline__ float kb8_triangulate(const Kb8RigDev& rig, float x1, float y1, float x2, float y2, float sigma1, float sigma2,
                                                        float* p3d) {
  return kb8_triangulate_p(rig.cam1, rig.prec1, rig.cam2, rig.prec2, rig.R12, rig.t12, x1, y1, x2, y2, sigma1, sigma2, p3d);
}
