// the byte fragments KeyFrame::serialize (reference include/KeyFrame.h:116-124) writes for the front-end's
// results through serializeVectorKeyPoints and serializeMatrix (include/SerializationUtils.h:74-152) into the binary Atlas file
// (boost::archive::binary_oarchive, src/System.cc:1434: primitives and make_array() blocks are stored as their native bytes):
//   keypoints:   int32 NumEl, then per keypoint float angle, response, size, pt.x, pt.y, int32 class_id, octave   (4 + 28 N bytes)
//   descriptors: int32 cols, rows, type, 1-byte bool continuous, then rows * cols * elemSize bytes                  (13 + 32 N bytes)
// The keypoint records are permuted on the device (one thread per keypoint, seven 4-byte words in, seven out) so the fragment leaves
// the GPU as one contiguous copy; the descriptor payload is the resident N x 32 matrix as it is.
#include <algorithm>
#include <cstring>

#include "orb_internal.h"

__global__ void k_serialize_keypoints(const orb_keypoint* __restrict__ kps, int n, uint32_t* __restrict__ out /* 1 + 7 n words */) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i == 0) out[0] = (uint32_t)n;
  if (i >= n) return;
  const uint32_t* s = reinterpret_cast<const uint32_t*>(kps + i);   // x y size angle response octave class_id
  uint32_t* d = out + 1 + 7 * (size_t)i;
  d[0] = s[3]; d[1] = s[4]; d[2] = s[2]; d[3] = s[0]; d[4] = s[1]; d[5] = s[6]; d[6] = s[5];
}

extern "C" {

size_t orb_serialized_keypoints_size(int n) { return 4 + 28 * (size_t)(n > 0 ? n : 0); }
size_t orb_serialized_descriptors_size(int n) { return 13 + 32 * (size_t)(n > 0 ? n : 0); }

int orb_serialize_frame(orb_handle* h, int frame, int what, uint8_t* out, size_t cap, size_t* written) {
  if (!h || !out || !written) return ORB_ERR_INVALID_ARG;
  if (!h->have_batch) return orb_set_error(h, ORB_ERR_STATE, "no extraction to serialise");
  if (frame < 0 || frame >= h->cur_batch) return orb_set_error(h, ORB_ERR_INVALID_ARG, "frame index outside the last batch");
  if (what == ORB_SER_KEYS_UN && !h->have_undist) return orb_set_error(h, ORB_ERR_STATE, "mvKeysUn: orb_undistort_keypoints has not run on this batch");
  int st;
  if ((st = orb_use_device(h))) return st;
  if ((st = orb_sync(h))) return st;     // completes a pending asynchronous extraction; h_n holds the keypoint counts
  const int n = std::min(h->h_n[frame], h->g.kcap), kcap = h->g.kcap;
  if (what == ORB_SER_KEYS || what == ORB_SER_KEYS_UN) {
    const size_t bytes = orb_serialized_keypoints_size(n);
    if (cap < bytes) return orb_set_error(h, ORB_ERR_CAPACITY, "serialisation buffer too small");
    if ((st = orb_ensure(h, h->d_scratch, bytes))) return st;
    const orb_keypoint* src = (what == ORB_SER_KEYS ? h->d_kps : h->d_kps_un).as<orb_keypoint>() + (size_t)frame * kcap;
    k_serialize_keypoints<<<(std::max(n, 1) + 127) / 128, 128, 0, h->stream>>>(src, n, h->d_scratch.as<uint32_t>());
    h->launches++;
    ORB_CUDA_CHECK(h, cudaGetLastError());
    ORB_CUDA_CHECK(h, cudaMemcpyAsync(out, h->d_scratch.p, bytes, cudaMemcpyDeviceToHost, h->stream));
    ORB_CUDA_CHECK(h, cudaStreamSynchronize(h->stream));
    *written = bytes;
    return ORB_OK;
  }
  if (what == ORB_SER_DESCRIPTORS) {
    const size_t bytes = orb_serialized_descriptors_size(n);
    if (cap < bytes) return orb_set_error(h, ORB_ERR_CAPACITY, "serialisation buffer too small");
    const int32_t hdr[3] = {32, n, 0 /* CV_8UC1 */};   // the empty matrix of a frame without keypoints keeps cols = 32 here (Mat(0, 32, CV_8U))
    memcpy(out, hdr, 12);
    out[12] = 1;                                        // isContinuous()
    if (n > 0) {
      ORB_CUDA_CHECK(h, cudaMemcpyAsync(out + 13, h->d_desc.as<uint8_t>() + (size_t)frame * kcap * 32, (size_t)n * 32, cudaMemcpyDeviceToHost, h->stream));
      ORB_CUDA_CHECK(h, cudaStreamSynchronize(h->stream));
    }
    *written = bytes;
    return ORB_OK;
  }
  return orb_set_error(h, ORB_ERR_INVALID_ARG, "unknown fragment kind");
}

// host-side loaders of the same fragments (no device work): the loading branch of the two reference templates
int orb_deserialize_keypoints(const uint8_t* in, size_t len, orb_keypoint* kps, int cap, int* n_out) {
  if (!in || !n_out || len < 4) return ORB_ERR_INVALID_ARG;
  int32_t n;
  memcpy(&n, in, 4);
  if (n < 0 || len < 4 + 28 * (size_t)n) return ORB_ERR_INVALID_ARG;
  *n_out = n;
  if (n > cap || (n > 0 && !kps)) return ORB_ERR_CAPACITY;
  for (int i = 0; i < n; ++i) {
    uint32_t w[7];
    memcpy(w, in + 4 + 28 * (size_t)i, 28);
    const uint32_t r[7] = {w[3], w[4], w[2], w[0], w[1], w[6], w[5]};
    memcpy(kps + i, r, 28);
  }
  return ORB_OK;
}

int orb_deserialize_descriptors(const uint8_t* in, size_t len, uint8_t* desc, int cap_rows, int* rows_out) {
  if (!in || !rows_out || len < 13) return ORB_ERR_INVALID_ARG;
  int32_t hdr[3];
  memcpy(hdr, in, 12);
  const int cols = hdr[0], rows = hdr[1], type = hdr[2];
  if (rows < 0 || (rows > 0 && (cols != 32 || type != 0)) || in[12] != 1 || len < 13 + (size_t)rows * 32) return ORB_ERR_INVALID_ARG;
  *rows_out = rows;
  if (rows > cap_rows || (rows > 0 && !desc)) return ORB_ERR_CAPACITY;
  memcpy(desc, in + 13, (size_t)rows * 32);
  return ORB_OK;
}

}  // extern "C"
