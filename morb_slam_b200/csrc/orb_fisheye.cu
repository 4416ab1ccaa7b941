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
#include "orb_kb8_dev.cuh"

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
                                                             float* __restrict__ p3d, int8_t* __restrict__ code, int host_cap,
                                                             int32_t* __restrict__ host_l2r, float* __restrict__ host_depth,
                                                             float* __restrict__ host_p3d, int8_t* __restrict__ host_code) {
  // host_* (a few frames, page-locked result buffers of the caller, host_cap entries per frame; host_cap == 0: none): the block copies
  // its own entries there at the end, so those four results need no device-to-host copies (mvRightToLeftMatch is an atomicMax across
  // blocks and keeps its copy)
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
  if (tid < total) {
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
  if (host_cap > 0) {
    __syncthreads();   // the entries of this block's keypoints were written by this block's threads
    if (j < kcapL && j < host_cap) {
      const size_t ho = (size_t)frame * host_cap + j;
      if (host_l2r) host_l2r[ho] = l2r[o];
      if (host_depth) host_depth[ho] = depth[o];
      if (host_p3d) { host_p3d[3 * ho] = p3d[3 * o]; host_p3d[3 * ho + 1] = p3d[3 * o + 1]; host_p3d[3 * ho + 2] = p3d[3 * o + 2]; }
      if (host_code) host_code[ho] = code[o];
    }
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
  // a few frames with page-locked result buffers: the kernel writes the per-left-keypoint results into them itself
  auto writable_or_null = [](const void* p) { return !p || orb_host_buffer_is_device_writable(p); };
  const bool zero_copy = batch <= ORB_SMALL_BATCH && !(flags & (ORB_DST_DEVICE | ORB_NO_OUTPUT)) && cap > 0 && (left_to_right || depth || p3d || code) &&
                         writable_or_null(left_to_right) && writable_or_null(depth) && writable_or_null(p3d) && writable_or_null(code);
  if ((st = orb_peer_read_begin(hL, hR))) return st;   // the right keypoints are produced on hR's stream
  ORB_CUDA_CHECK(hL, cudaMemsetAsync(hL->d_fe_r2l.p, 0xff, nr * 4, hL->stream));
  k_fisheye_triangulate<<<dim3((kL + FT_TRI_THREADS - 1) / FT_TRI_THREADS, batch), FT_TRI_THREADS, 0, hL->stream>>>(
      d, hL->d_kps.as<orb_keypoint>(), hL->d_n.as<int>(), hL->d_mono.as<int>(), kL, hR->d_kps.as<orb_keypoint>(), hR->d_n.as<int>(),
      hR->d_mono.as<int>(), kR, hL->d_fe_idx.as<int32_t>(), hL->d_fe_pass.as<uint8_t>(), hL->d_fe_l2r.as<int32_t>(), hL->d_fe_r2l.as<int32_t>(),
      hL->d_fe_depth.as<float>(), hL->d_fe_p3d.as<float>(), hL->d_fe_code.as<int8_t>(), zero_copy ? cap : 0, zero_copy ? left_to_right : nullptr,
      zero_copy ? depth : nullptr, zero_copy ? p3d : nullptr, zero_copy ? code : nullptr);
  hL->launches++;
  if ((st = orb_peer_read_end(hL, hR))) return st;     // hR's next extraction waits for this kernel
  hL->have_fe_tri = true;
  ORB_CUDA_CHECK(hL, cudaGetLastError());
  if (!(flags & ORB_NO_OUTPUT)) {
    const int rl = std::min(cap, kL), rr = std::min(cap, kR);
    if (left_to_right && !zero_copy)
      ORB_CUDA_CHECK(hL, cudaMemcpy2DAsync(left_to_right, (size_t)cap * 4, hL->d_fe_l2r.p, (size_t)kL * 4, (size_t)rl * 4, batch, cudaMemcpyDefault, hL->stream));
    if (right_to_left)
      ORB_CUDA_CHECK(hL, cudaMemcpy2DAsync(right_to_left, (size_t)cap * 4, hL->d_fe_r2l.p, (size_t)kR * 4, (size_t)rr * 4, batch, cudaMemcpyDefault, hL->stream));
    if (depth && !zero_copy)
      ORB_CUDA_CHECK(hL, cudaMemcpy2DAsync(depth, (size_t)cap * 4, hL->d_fe_depth.p, (size_t)kL * 4, (size_t)rl * 4, batch, cudaMemcpyDefault, hL->stream));
    if (p3d && !zero_copy)
      ORB_CUDA_CHECK(hL, cudaMemcpy2DAsync(p3d, (size_t)cap * 12, hL->d_fe_p3d.p, (size_t)kL * 12, (size_t)rl * 12, batch, cudaMemcpyDefault, hL->stream));
    if (code && !zero_copy)
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
