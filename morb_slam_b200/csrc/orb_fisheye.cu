// The second half of Frame::ComputeStereoFishEyeMatches (reference src/Frame.cc:1252-1277): every left keypoint whose best
// match passed the ratio test (k_fisheye_knn2, orb_knn.cu) is triangulated with KannalaBrandt8::TriangulateMatches
// (src/CameraModels/KannalaBrandt8.cpp:323-395) and accepted when the returned depth exceeds 0.0001f.
//
// One thread per left keypoint; the work per match is ~1.5 k flops on 2 x 28 bytes of input, so the kernel is a single wave of
// latency (a TUM-VI frame has <= 1500 candidates, a batch of 256 frames 384 k threads). Float expressions are written in the
// reference's order with plain operators (the library is built with -fmad=false, nothing contracts); tanf / atan2f / cosf / sinf
// are glibc 2.39's float routines restated for the device (orb_libm_glibc.cuh, pinned exhaustively against the image's libm on the
// host); Eigen::JacobiSVD<Matrix4f> (:425) is Eigen's published two-sided Jacobi algorithm restated in float (kb8_jacobi_v3 below).
// Accept / reject codes, depths and 3-D points equal the oracle's bit for bit (tests/test_gpu_fisheye.py). What cannot be checked in
// this image is the oracle's Eigen stand-in against the real library (Eigen is an un-vendored dependency of the reference, absent
// here): fixed-size reduction order and JacobiSVD are restated from Eigen's source as published, not pinned against a build of it.
#include <algorithm>

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
static __device__ float kb8_triangulate(const Kb8RigDev& rig, float x1, float y1, float x2, float y2, float sigma1, float sigma2, float* p3d) {
  float r1x, r1y, r2x, r2y;
  kb8_unproject(rig.cam1, rig.prec1, x1, y1, r1x, r1y);
  kb8_unproject(rig.cam2, rig.prec2, x2, y2, r2x, r2y);
  const float* R = rig.R12;
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
    T2[4 * i + 3] = sum3((-R[0 * 3 + i]) * rig.t12[0], (-R[1 * 3 + i]) * rig.t12[1], (-R[2 * 3 + i]) * rig.t12[2]);
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
  kb8_project(rig.cam1, X, Y, Z, u, v);
  const float e1x = u - x1, e1y = v - y1;
  if ((double)(e1x * e1x + e1y * e1y) > 5.991 * (double)sigma1) return -4.f;
  const float X2 = sum3(T2[0] * X, T2[1] * Y, T2[2] * Z) + T2[3];
  const float Y2 = sum3(T2[4] * X, T2[5] * Y, T2[6] * Z) + T2[7];
  const float Z2 = sum3(T2[8] * X, T2[9] * Y, T2[10] * Z) + T2[11];
  kb8_project(rig.cam2, X2, Y2, Z2, u, v);
  const float e2x = u - x2, e2y = v - y2;
  if ((double)(e2x * e2x + e2y * e2y) > 5.991 * (double)sigma2) return -5.f;
  p3d[0] = X; p3d[1] = Y; p3d[2] = Z;
  return z1;
}

// One thread per left keypoint j of the frame writes the defaults; the keypoints whose match passed the ratio test (about half of the
// stereo keypoints) are compacted into a per-block list first, so that the expensive triangulation runs in fully populated warps
// instead of half-empty ones (0.19 -> see profiles/README_r1.md per 256 TUM-VI frames). mvRightToLeftMatch keeps the LAST accepted
// query of the reference's loop = the largest left index, hence atomicMax on a -1-initialised array.
#define FT_TRI_THREADS 128
__global__ void __launch_bounds__(FT_TRI_THREADS) k_fisheye_triangulate(Kb8RigDev rig, const orb_keypoint* __restrict__ kpsL, const int* __restrict__ nL,
                                                             const int* __restrict__ monoL, int kcapL, const orb_keypoint* __restrict__ kpsR,
                                                             const int* __restrict__ nR, const int* __restrict__ monoR, int kcapR,
                                                             const int32_t* __restrict__ fe_idx, const uint8_t* __restrict__ fe_pass,
                                                             int32_t* __restrict__ l2r, int32_t* __restrict__ r2l, float* __restrict__ depth,
                                                             float* __restrict__ p3d, int8_t* __restrict__ code) {
  __shared__ int s_item[FT_TRI_THREADS];     // (left keypoint offset inside the block) << 20 | right keypoint index
  __shared__ int s_warp[FT_TRI_THREADS / 32];
  const int frame = blockIdx.y, tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const int j = blockIdx.x * FT_TRI_THREADS + tid;
  const size_t o = (size_t)frame * kcapL + j;
  const int n = min(nL[frame], kcapL), m0 = max(monoL[frame], 0);
  const int i = j - m0;   // query index of the kNN
  int t = -1;
  if (j < kcapL && j < n && i >= 0 && fe_pass[(size_t)frame * kcapL + i]) {
    t = fe_idx[((size_t)frame * kcapL + i) * 2] + max(monoR[frame], 0);
    if (t < 0 || t >= min(nR[frame], kcapR)) t = -1;
  }
  if (j < kcapL) {   // defaults; the worker threads below overwrite the entries of the triangulated keypoints after the barrier
    l2r[o] = -1;
    depth[o] = -1.f;
    p3d[3 * o] = 0.f; p3d[3 * o + 1] = 0.f; p3d[3 * o + 2] = 0.f;
    code[o] = 0;
  }
  // order-preserving compaction of the work items (ballot + prefix over the warps)
  const uint32_t bal = __ballot_sync(0xffffffffu, t >= 0);
  if (lane == 0) s_warp[wid] = __popc(bal);
  __syncthreads();
  int before = 0, total = 0;
#pragma unroll
  for (int k = 0; k < FT_TRI_THREADS / 32; ++k) { if (k < wid) before += s_warp[k]; total += s_warp[k]; }
  if (t >= 0) s_item[before + __popc(bal & ((1u << lane) - 1u))] = (tid << 20) | t;
  __syncthreads();
  if (tid >= total) return;
  const int it = s_item[tid];
  const int jj = blockIdx.x * FT_TRI_THREADS + (it >> 20), tt = it & 0xfffff;
  const size_t oo = (size_t)frame * kcapL + jj;
  const orb_keypoint a = kpsL[oo], b = kpsR[(size_t)frame * kcapR + tt];
  float X[3];
  const float d = kb8_triangulate(rig, a.x, a.y, b.x, b.y, rig.sig1[a.octave], rig.sig2[b.octave], X);
  if (d > 0.0001f) {
    l2r[oo] = tt;
    depth[oo] = d;
    p3d[3 * oo] = X[0]; p3d[3 * oo + 1] = X[1]; p3d[3 * oo + 2] = X[2];
    code[oo] = 1;
    atomicMax(&r2l[(size_t)frame * kcapR + tt], jj);
  } else {
    code[oo] = (int8_t)(d < 0.f ? (int)d : -6);
  }
}

__global__ void k_kb8_triangulate_pairs(Kb8RigDev rig, const float* __restrict__ xy1, const float* __restrict__ xy2, const float* __restrict__ s1,
                                        const float* __restrict__ s2, int n, float* __restrict__ ret, float* __restrict__ p3d) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float X[3] = {0.f, 0.f, 0.f};
  ret[i] = kb8_triangulate(rig, xy1[2 * i], xy1[2 * i + 1], xy2[2 * i], xy2[2 * i + 1], s1[i], s2[i], X);
  p3d[3 * i] = X[0]; p3d[3 * i + 1] = X[1]; p3d[3 * i + 2] = X[2];
}

static void fill_rig(Kb8RigDev& d, const orb_kb8_rig* rig, const orb_handle* hL, const orb_handle* hR) {
  for (int i = 0; i < 8; ++i) { d.cam1[i] = rig->cam1[i]; d.cam2[i] = rig->cam2[i]; }
  d.prec1 = rig->precision1; d.prec2 = rig->precision2;
  for (int i = 0; i < 9; ++i) d.R12[i] = rig->R12[i];
  for (int i = 0; i < 3; ++i) d.t12[i] = rig->t12[i];
  for (int i = 0; i < ORB_MAX_LEVELS; ++i) {
    d.sig1[i] = hL && i < (int)hL->sigma2.size() ? hL->sigma2[i] : 1.f;
    d.sig2[i] = hR && i < (int)hR->sigma2.size() ? hR->sigma2[i] : 1.f;
  }
}

extern "C" {

int orb_stereo_fisheye_triangulate_batch(orb_handle* hL, orb_handle* hR, const orb_kb8_rig* rig, int32_t* left_to_right, int32_t* right_to_left,
                                         float* depth, float* p3d, int8_t* code, int cap, int flags) {
  if (!hL || !hR || !rig) return ORB_ERR_INVALID_ARG;
  if (!hL->have_batch || !hR->have_batch || !hL->have_fe)
    return orb_set_error(hL, ORB_ERR_STATE, "fisheye triangulation needs orb_stereo_fisheye_match_batch on the current batches");
  if (hL->device != hR->device || hL->cur_batch != hR->cur_batch) return orb_set_error(hL, ORB_ERR_INVALID_ARG, "left/right handles differ in device or batch");
  int st;
  if ((st = orb_use_device(hL))) return st;
  const int batch = hL->cur_batch, kL = hL->g.kcap, kR = hR->g.kcap;
  const size_t nl = (size_t)batch * kL, nr = (size_t)batch * kR;
  if ((st = orb_ensure(hL, hL->d_fe_l2r, nl * 4)) || (st = orb_ensure(hL, hL->d_fe_r2l, nr * 4)) || (st = orb_ensure(hL, hL->d_fe_depth, nl * 4)) ||
      (st = orb_ensure(hL, hL->d_fe_p3d, nl * 12)) || (st = orb_ensure(hL, hL->d_fe_code, nl)))
    return st;
  Kb8RigDev d;
  fill_rig(d, rig, hL, hR);
  if ((st = orb_peer_read_begin(hL, hR))) return st;   // the right keypoints are produced on hR's stream
  ORB_CUDA_CHECK(hL, cudaMemsetAsync(hL->d_fe_r2l.p, 0xff, nr * 4, hL->stream));
  k_fisheye_triangulate<<<dim3((kL + FT_TRI_THREADS - 1) / FT_TRI_THREADS, batch), FT_TRI_THREADS, 0, hL->stream>>>(
      d, hL->d_kps.as<orb_keypoint>(), hL->d_n.as<int>(), hL->d_mono.as<int>(), kL, hR->d_kps.as<orb_keypoint>(), hR->d_n.as<int>(),
      hR->d_mono.as<int>(), kR, hL->d_fe_idx.as<int32_t>(), hL->d_fe_pass.as<uint8_t>(), hL->d_fe_l2r.as<int32_t>(), hL->d_fe_r2l.as<int32_t>(),
      hL->d_fe_depth.as<float>(), hL->d_fe_p3d.as<float>(), hL->d_fe_code.as<int8_t>());
  hL->launches++;
  if ((st = orb_peer_read_end(hL, hR))) return st;     // hR's next extraction waits for this kernel
  hL->have_fe_tri = true;
  ORB_CUDA_CHECK(hL, cudaGetLastError());
  if (!(flags & ORB_NO_OUTPUT)) {
    const int rl = std::min(cap, kL), rr = std::min(cap, kR);
    if (left_to_right)
      ORB_CUDA_CHECK(hL, cudaMemcpy2DAsync(left_to_right, (size_t)cap * 4, hL->d_fe_l2r.p, (size_t)kL * 4, (size_t)rl * 4, batch, cudaMemcpyDefault, hL->stream));
    if (right_to_left)
      ORB_CUDA_CHECK(hL, cudaMemcpy2DAsync(right_to_left, (size_t)cap * 4, hL->d_fe_r2l.p, (size_t)kR * 4, (size_t)rr * 4, batch, cudaMemcpyDefault, hL->stream));
    if (depth)
      ORB_CUDA_CHECK(hL, cudaMemcpy2DAsync(depth, (size_t)cap * 4, hL->d_fe_depth.p, (size_t)kL * 4, (size_t)rl * 4, batch, cudaMemcpyDefault, hL->stream));
    if (p3d)
      ORB_CUDA_CHECK(hL, cudaMemcpy2DAsync(p3d, (size_t)cap * 12, hL->d_fe_p3d.p, (size_t)kL * 12, (size_t)rl * 12, batch, cudaMemcpyDefault, hL->stream));
    if (code)
      ORB_CUDA_CHECK(hL, cudaMemcpy2DAsync(code, (size_t)cap, hL->d_fe_code.p, (size_t)kL, (size_t)rl, batch, cudaMemcpyDefault, hL->stream));
  }
  if (flags & ORB_ASYNC) return ORB_OK;
  ORB_CUDA_CHECK(hL, cudaStreamSynchronize(hL->stream));
  return ORB_OK;
}

int orb_kb8_triangulate_matches(orb_handle* h, const orb_kb8_rig* rig, const float* xy1, const float* xy2, const float* sigma1, const float* sigma2,
                                int n, float* ret, float* p3d) {
  if (!h || !rig || n < 0 || (n > 0 && (!xy1 || !xy2 || !sigma1 || !sigma2 || !ret || !p3d))) return ORB_ERR_INVALID_ARG;
  if (n == 0) return ORB_OK;
  int st;
  if ((st = orb_use_device(h))) return st;
  // scratch: xy1 | xy2 | s1 | s2 (6 n floats), scratch2: ret | p3d (4 n floats)
  if ((st = orb_ensure(h, h->d_scratch, (size_t)n * 24)) || (st = orb_ensure(h, h->d_scratch2, (size_t)n * 16))) return st;
  float* in = h->d_scratch.as<float>();
  float* out = h->d_scratch2.as<float>();
  ORB_CUDA_CHECK(h, cudaMemcpyAsync(in, xy1, (size_t)n * 8, cudaMemcpyHostToDevice, h->stream));
  ORB_CUDA_CHECK(h, cudaMemcpyAsync(in + 2 * (size_t)n, xy2, (size_t)n * 8, cudaMemcpyHostToDevice, h->stream));
  ORB_CUDA_CHECK(h, cudaMemcpyAsync(in + 4 * (size_t)n, sigma1, (size_t)n * 4, cudaMemcpyHostToDevice, h->stream));
  ORB_CUDA_CHECK(h, cudaMemcpyAsync(in + 5 * (size_t)n, sigma2, (size_t)n * 4, cudaMemcpyHostToDevice, h->stream));
  Kb8RigDev d;
  fill_rig(d, rig, nullptr, nullptr);
  k_kb8_triangulate_pairs<<<(n + 127) / 128, 128, 0, h->stream>>>(d, in, in + 2 * (size_t)n, in + 4 * (size_t)n, in + 5 * (size_t)n, n, out,
                                                                  out + (size_t)n);
  h->launches++;
  ORB_CUDA_CHECK(h, cudaGetLastError());
  ORB_CUDA_CHECK(h, cudaMemcpyAsync(ret, out, (size_t)n * 4, cudaMemcpyDeviceToHost, h->stream));
  ORB_CUDA_CHECK(h, cudaMemcpyAsync(p3d, out + (size_t)n, (size_t)n * 12, cudaMemcpyDeviceToHost, h->stream));
  ORB_CUDA_CHECK(h, cudaStreamSynchronize(h->stream));
  return ORB_OK;
}

}  // extern "C"
