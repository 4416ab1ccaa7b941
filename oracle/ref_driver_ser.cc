// TEST INFRASTRUCTURE ONLY (oracle/_ref). Not part of the product path.
//
// The reference's serializeMatrix / serializeVectorKeyPoints templates (include/SerializationUtils.h:74-152), cut out by line
// range at build time (oracle/Makefile) and instantiated with the raw-bytes archives below. Boost is not in this image; what the
// stand-in reproduces is the documented behaviour of boost::archive::binary_oarchive / binary_iarchive for the only things these
// templates hand to it: arithmetic primitives (int, bool, float) and boost::serialization::make_array blocks are stored as their
// sizeof(T) native bytes, in call order, with no per-item framing. Field order, field types and the continuous / per-row branch
// are the reference's own code. Nothing of the reference is copied into the repository.
#include <cstdint>
#include <cstring>
#include <type_traits>
#include <vector>

#include <opencv2/core/core.hpp>   // the oracle's shim

namespace boost { namespace serialization {
struct RawArray { void* p; size_t n; };
template <class T> RawArray make_array(T* p, size_t n) { return RawArray{(void*)p, n * sizeof(T)}; }
} }

struct RawOut {
  typedef std::true_type is_saving;
  typedef std::false_type is_loading;
  std::vector<uint8_t> bytes;
  template <class T> typename std::enable_if<std::is_arithmetic<T>::value, RawOut&>::type operator&(T& v) {
    const uint8_t* p = (const uint8_t*)&v; bytes.insert(bytes.end(), p, p + sizeof(T)); return *this;
  }
  RawOut& operator&(const boost::serialization::RawArray& a) { const uint8_t* p = (const uint8_t*)a.p; bytes.insert(bytes.end(), p, p + a.n); return *this; }
};
struct RawIn {
  typedef std::false_type is_saving;
  typedef std::true_type is_loading;
  const uint8_t* cur;
  template <class T> typename std::enable_if<std::is_arithmetic<T>::value, RawIn&>::type operator&(T& v) { memcpy(&v, cur, sizeof(T)); cur += sizeof(T); return *this; }
  RawIn& operator&(const boost::serialization::RawArray& a) { memcpy(a.p, cur, a.n); cur += a.n; return *this; }
};

namespace ORB_SLAM3 {
#include "serialization_utils.inc"
}

extern "C" {
// returns the fragment size; out may be NULL to query it
size_t ref_serialize_keypoints(const cv::KeyPoint* kps, int n, uint8_t* out) {
  std::vector<cv::KeyPoint> v(kps, kps + n);
  RawOut ar;
  ORB_SLAM3::serializeVectorKeyPoints(ar, v, 0);
  if (out) memcpy(out, ar.bytes.data(), ar.bytes.size());
  return ar.bytes.size();
}
size_t ref_serialize_matrix(const uint8_t* data, int rows, int cols, size_t stride, uint8_t* out) {
  cv::Mat m(rows, cols, CV_8UC1, (void*)data, stride);
  RawOut ar;
  ORB_SLAM3::serializeMatrix(ar, m, 0);
  if (out) memcpy(out, ar.bytes.data(), ar.bytes.size());
  return ar.bytes.size();
}
int ref_deserialize_keypoints(const uint8_t* in, cv::KeyPoint* kps, int cap) {
  std::vector<cv::KeyPoint> v;
  RawIn ar{in};
  ORB_SLAM3::serializeVectorKeyPoints(ar, v, 0);
  for (int i = 0; i < (int)v.size() && i < cap; ++i) kps[i] = v[i];
  return (int)v.size();
}
int ref_deserialize_matrix(const uint8_t* in, uint8_t* data, int cap_bytes, int* cols) {
  cv::Mat m;
  RawIn ar{in};
  ORB_SLAM3::serializeMatrix(ar, m, 0);
  *cols = m.cols;
  if (m.rows * m.cols <= cap_bytes && m.rows * m.cols > 0) memcpy(data, m.data, (size_t)m.rows * m.cols);
  return m.rows;
}
}
