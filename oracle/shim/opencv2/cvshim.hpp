// TEST INFRASTRUCTURE ONLY (oracle). Not part of the product path.
//
// Minimal stand-in for the slice of the OpenCV C++ API that the reference's hot
// path touches (reference: src/ORBextractor.cc, src/Frame.cc:889-1047,1222-1274).
// OpenCV itself is an un-vendored third-party dependency of the reference
// (CMakeLists.txt:27 wants >= 4.4, CI pins 4.5.2); it is absent from this image as
// a C++ library, so its published algorithms are restated here and every
// primitive is pinned bit-exact against the Python cv2 4.13.0 wheel by
// tests/test_oracle_primitives.py.
//
// The container types (Mat, KeyPoint, ...) carry only what the reference uses.
// The numeric primitives (resize, GaussianBlur, FAST, fastAtan2, copyMakeBorder,
// norm, BFMatcher) are written from the OpenCV algorithm descriptions in
// SURVEY.md Appendix A, not copied from OpenCV source.
#pragma once
#include <algorithm>
#include <cassert>
#include <cfloat>
#include <climits>
#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <vector>

#define CV_PI 3.1415926535897932384626433832795
#define CV_8U 0
#define CV_8UC1 0

typedef unsigned char uchar;

static inline int cvRound(float v) { return (int)lrintf(v); }   // round-half-even (SSE cvtss2si)
static inline int cvRound(double v) { return (int)lrint(v); }  // round-half-even (SSE cvtsd2si)
static inline int cvRound(int v) { return v; }
static inline int cvFloor(float v) { int i = (int)v; return i - (i > v); }
static inline int cvFloor(double v) { int i = (int)v; return i - (i > v); }
static inline int cvCeil(float v) { int i = (int)v; return i + (i < v); }
static inline int cvCeil(double v) { int i = (int)v; return i + (i < v); }

namespace cv {

typedef ::uchar uchar;

enum { BORDER_CONSTANT = 0, BORDER_REPLICATE = 1, BORDER_REFLECT = 2, BORDER_WRAP = 3,
       BORDER_REFLECT_101 = 4, BORDER_REFLECT101 = 4, BORDER_DEFAULT = 4, BORDER_ISOLATED = 16 };
enum { INTER_NEAREST = 0, INTER_LINEAR = 1, INTER_CUBIC = 2, INTER_AREA = 3 };
enum { NORM_INF = 1, NORM_L1 = 2, NORM_L2 = 4, NORM_HAMMING = 6 };

template <typename T>
struct Point_ {
  T x, y;
  Point_() : x(0), y(0) {}
  Point_(T _x, T _y) : x(_x), y(_y) {}
  template <typename U>
  Point_(const Point_<U>& o) : x((T)o.x), y((T)o.y) {}
};
typedef Point_<int> Point2i;
typedef Point_<int> Point;
typedef Point_<float> Point2f;

// cv::Point_<float> *= float multiplies in float (saturate_cast<float> is the identity)
static inline Point2f& operator*=(Point2f& a, float b) { a.x = a.x * b; a.y = a.y * b; return a; }

struct Size {
  int width, height;
  Size() : width(0), height(0) {}
  Size(int w, int h) : width(w), height(h) {}
};

struct Rect {
  int x, y, width, height;
  Rect() : x(0), y(0), width(0), height(0) {}
  Rect(int _x, int _y, int w, int h) : x(_x), y(_y), width(w), height(h) {}
};

// 28-byte layout identical to cv::KeyPoint (pt.x, pt.y, size, angle, response, octave, class_id)
struct KeyPoint {
  Point2f pt;
  float size;
  float angle;
  float response;
  int octave;
  int class_id;
  KeyPoint() : pt(0, 0), size(0), angle(-1), response(0), octave(0), class_id(-1) {}
  KeyPoint(float x, float y, float _size, float _angle = -1, float _response = 0, int _octave = 0,
           int _class_id = -1)
      : pt(x, y), size(_size), angle(_angle), response(_response), octave(_octave), class_id(_class_id) {}
};
static_assert(sizeof(KeyPoint) == 28, "cv::KeyPoint layout");

struct MatStep {
  size_t v;
  MatStep() : v(0) {}
  MatStep(size_t s) : v(s) {}
  operator size_t() const { return v; }
};

class Mat {
 public:
  int rows, cols;
  uchar* data;
  MatStep step;
  std::shared_ptr<std::vector<uchar>> buf;  // keeps the allocation alive for ROI views

  Mat() : rows(0), cols(0), data(nullptr) {}
  Mat(int r, int c, int type) { alloc(r, c, type); }
  Mat(Size s, int type) { alloc(s.height, s.width, type); }
  // wrap user memory (no ownership), like cv::Mat(rows, cols, type, void*, step)
  Mat(int r, int c, int type, void* p, size_t st = 0) : rows(r), cols(c), data((uchar*)p), step(st ? st : (size_t)c) {
    (void)type;
  }
  void alloc(int r, int c, int type) {
    assert(type == CV_8UC1);
    (void)type;
    rows = r; cols = c; step = MatStep((size_t)c);
    buf = std::make_shared<std::vector<uchar>>((size_t)r * c);
    data = buf->data();
  }
  void create(int r, int c, int type) {
    if (data && rows == r && cols == c) return;  // OpenCV: no-op when size/type already match
    alloc(r, c, type);
  }
  void release() { rows = cols = 0; data = nullptr; buf.reset(); }
  bool empty() const { return data == nullptr || rows == 0 || cols == 0; }
  int type() const { return CV_8UC1; }
  size_t elemSize() const { return 1; }   // 8UC1 only
  size_t step1() const { return step.v; }
  bool isContinuous() const { return step.v == (size_t)cols || rows == 1; }

  Mat operator()(const Rect& r) const {
    Mat m;
    m.rows = r.height; m.cols = r.width; m.step = step; m.buf = buf;
    m.data = data + (size_t)r.y * step.v + r.x;
    return m;
  }
  Mat rowRange(int a, int b) const { return (*this)(Rect(0, a, cols, b - a)); }
  Mat colRange(int a, int b) const { return (*this)(Rect(a, 0, b - a, rows)); }
  Mat row(int y) const { return rowRange(y, y + 1); }
  Mat clone() const {
    Mat m(rows, cols, CV_8UC1);
    for (int y = 0; y < rows; ++y) std::memcpy(m.data + (size_t)y * m.step.v, data + (size_t)y * step.v, cols);
    return m;
  }
  void copyTo(const Mat& dst) const {  // destination must already have the right size (all call sites do)
    assert(dst.rows == rows && dst.cols == cols);
    for (int y = 0; y < rows; ++y) std::memcpy(dst.data + (size_t)y * dst.step.v, data + (size_t)y * step.v, cols);
  }
  template <typename T> T& at(int y, int x) { return *(T*)(data + (size_t)y * step.v + x * sizeof(T)); }
  template <typename T> const T& at(int y, int x) const { return *(const T*)(data + (size_t)y * step.v + x * sizeof(T)); }
  uchar* ptr(int y = 0) { return data + (size_t)y * step.v; }
  const uchar* ptr(int y = 0) const { return data + (size_t)y * step.v; }
  template <typename T> T* ptr(int y = 0) { return (T*)(data + (size_t)y * step.v); }
  template <typename T> const T* ptr(int y = 0) const { return (const T*)(data + (size_t)y * step.v); }
  static Mat zeros(int r, int c, int type) {
    Mat m(r, c, type);
    std::memset(m.data, 0, (size_t)r * c);
    return m;
  }
};

class _InputArray {
 public:
  const Mat* m;
  _InputArray(const Mat& mm) : m(&mm) {}
  Mat getMat() const { return *m; }
  bool empty() const { return m->empty(); }
};
class _OutputArray {
 public:
  Mat* m;
  _OutputArray(Mat& mm) : m(&mm) {}
  _OutputArray(const Mat& mm) : m(const_cast<Mat*>(&mm)) {}  // fixed-size outputs (ROI temporaries)
  Mat getMat() const { return *m; }
  void create(int r, int c, int type) const { m->create(r, c, type); }
  void create(Size s, int type) const { m->create(s.height, s.width, type); }
  void release() const { m->release(); }
  bool empty() const { return m->empty(); }
};
typedef const _InputArray& InputArray;
typedef const _OutputArray& OutputArray;

// ---- border index, BORDER_REFLECT_101: gfedcb|abcdefgh|gfedcba --------------------------------
static inline int reflect101(int p, int len) {
  if (len == 1) return 0;
  while (p < 0 || p >= len) {
    if (p < 0) p = -p;
    else p = 2 * len - 2 - p;
  }
  return p;
}

// cv::copyMakeBorder for the two call shapes of ORBextractor.cc:1104-1109. When dst already is the
// bordered buffer around src (in-place ROI call with BORDER_ISOLATED) only the margins change.
static inline void copyMakeBorder(InputArray _src, OutputArray _dst, int top, int bottom, int left, int right,
                                  int borderType) {
  Mat src = _src.getMat();
  assert((borderType & ~BORDER_ISOLATED) == BORDER_REFLECT_101);
  (void)borderType;
  _dst.create(src.rows + top + bottom, src.cols + left + right, CV_8UC1);
  Mat dst = _dst.getMat();
  // interior first (a no-op when src aliases the interior of dst)
  for (int y = 0; y < src.rows; ++y) {
    uchar* d = dst.data + (size_t)(y + top) * dst.step.v + left;
    const uchar* s = src.data + (size_t)y * src.step.v;
    if (d != s) std::memmove(d, s, src.cols);
  }
  for (int y = 0; y < dst.rows; ++y) {
    int sy = reflect101(y - top, src.rows);
    const uchar* s = src.data + (size_t)sy * src.step.v;
    uchar* d = dst.data + (size_t)y * dst.step.v;
    const bool interior_row = (y >= top && y < top + src.rows);
    for (int x = 0; x < left; ++x) d[x] = s[reflect101(x - left, src.cols)];
    if (!interior_row)
      for (int x = left; x < left + src.cols; ++x) d[x] = s[x - left];
    for (int x = left + src.cols; x < dst.cols; ++x) d[x] = s[reflect101(x - left, src.cols)];
  }
}

// ---- cv::resize, INTER_LINEAR, 8UC1 -----------------------------------
// 11-bit fixed-point coefficients, int32 horizontal pass, vertical pass with the >>4, >>16, +2, >>2
// rounding chain of OpenCV's 8U linear resizer.
namespace shim_detail {
struct AxisTab {
  std::vector<int> ofs;
  std::vector<short> c0, c1;
};
static inline short sat_short(float v) {
  int i = cvRound(v);
  return (short)std::min(std::max(i, (int)SHRT_MIN), (int)SHRT_MAX);
}
// horizontal rule: offsets clamped and the fraction zeroed at both ends
static inline AxisTab axis_tab_h(int S, int D) {
  AxisTab t; t.ofs.resize(D); t.c0.resize(D); t.c1.resize(D);
  double inv_scale = (double)D / S;
  double scale = 1.0 / inv_scale;
  for (int d = 0; d < D; ++d) {
    float f = (float)((d + 0.5) * scale - 0.5);
    int s = cvFloor(f);
    f -= s;
    if (s < 0) { s = 0; f = 0.f; }
    if (s >= S - 1) { s = S - 1; f = 0.f; }
    t.ofs[d] = s;
    t.c0[d] = sat_short((1.f - f) * 2048.f);
    t.c1[d] = sat_short(f * 2048.f);
  }
  return t;
}
// vertical rule: the fraction is kept, the two source rows are clipped into the image
static inline AxisTab axis_tab_v(int S, int D) {
  AxisTab t; t.ofs.resize(D); t.c0.resize(D); t.c1.resize(D);
  double inv_scale = (double)D / S;
  double scale = 1.0 / inv_scale;
  for (int d = 0; d < D; ++d) {
    float f = (float)((d + 0.5) * scale - 0.5);
    int s = cvFloor(f);
    f -= s;
    t.ofs[d] = s;
    t.c0[d] = sat_short((1.f - f) * 2048.f);
    t.c1[d] = sat_short(f * 2048.f);
  }
  return t;
}
}  // namespace shim_detail

// ---- cv::undistortPoints(src, dst, cameraMatrix, distCoeffs, noArray(), P) as Frame::UndistortKeyPoints calls it
// (reference src/Frame.cc:845-846): OpenCV's cvUndistortPointsInternal with TermCriteria(MAX_ITER, 5, 0.01), R = identity,
// no tilt terms; everything in double after the float inputs are widened, results narrowed to float. K / P: 3 x 3 row-major
// float (only fx, fy, cx, cy of K are used, like OpenCV); dist: k1 k2 p1 p2 [k3 [k4 k5 k6 [s1 s2 s3 s4]]].
static inline void undistort_points_pinhole(const float* pts, int n, const float* K, const float* dist, int ndist, const float* P, float* out) {
  double k[14] = {0};
  for (int i = 0; i < ndist && i < 14; ++i) k[i] = dist[i];
  const double fx = K[0], fy = K[4], cx = K[2], cy = K[5], ifx = 1. / fx, ify = 1. / fy;
  double RR[9];
  for (int i = 0; i < 9; ++i) RR[i] = P[i];                   // P * identity
  for (int i = 0; i < n; ++i) {
    const double u = pts[2 * i], v = pts[2 * i + 1];
    double x = (u - cx) * ifx, y = (v - cy) * ify;
    const double x0 = x, y0 = y;
    for (int j = 0; j < 5; ++j) {
      const double r2 = x * x + y * y;
      const double icdist = (1 + ((k[7] * r2 + k[6]) * r2 + k[5]) * r2) / (1 + ((k[4] * r2 + k[1]) * r2 + k[0]) * r2);
      if (icdist < 0) { x = (u - cx) * ifx; y = (v - cy) * ify; break; }
      const double deltaX = 2 * k[2] * x * y + k[3] * (r2 + 2 * x * x) + k[8] * r2 + k[9] * r2 * r2;
      const double deltaY = k[2] * (r2 + 2 * y * y) + 2 * k[3] * x * y + k[10] * r2 + k[11] * r2 * r2;
      x = (x0 - deltaX) * icdist;
      y = (y0 - deltaY) * icdist;
    }
    const double xx = RR[0] * x + RR[1] * y + RR[2], yy = RR[3] * x + RR[4] * y + RR[5], ww = 1. / (RR[6] * x + RR[7] * y + RR[8]);
    out[2 * i] = (float)(xx * ww);
    out[2 * i + 1] = (float)(yy * ww);
  }
}

// ---- cv::remap(src, dst, map1 CV_32FC1, map2 CV_32FC1, INTER_LINEAR, BORDER_CONSTANT, 0) on 8UC1 (OpenCV imgwarp.cpp):
// the float maps become fixed point with INTER_BITS = 5 (sx = cvRound(x * 32), integer part saturated to short, 5-bit
// fractions index a 32 x 32 table of four 15-bit weights), result = (sum w_i * p_i + 2^14) >> 15; pixels of the 2 x 2
// footprint outside the source count as 0. The weights are exact multiples of 32 except the entry for fractions (0, 0):
// 1.0 * 32768 saturates to 32767 and OpenCV's sum correction puts the missing 1 on the LAST weight (its search loop
// starts at the last element for a 2 x 2 kernel), i.e. {32767, 0, 0, 1}.
namespace shim_detail {
static inline void remap_weights(int fx, int fy, int w[4]) {
  w[0] = (32 - fy) * (32 - fx) * 32; w[1] = (32 - fy) * fx * 32; w[2] = fy * (32 - fx) * 32; w[3] = fy * fx * 32;
  if (fx == 0 && fy == 0) { w[0] = 32767; w[3] = 1; }
}
static inline int sat_short_i(int v) { return v < -32768 ? -32768 : (v > 32767 ? 32767 : v); }
// cvtss2si / cvtps2dq: values outside the int range and NaN give the "integer indefinite" 0x80000000
static inline int round_i32(float v) { return (v >= -2147483648.f && v < 2147483648.f) ? (int)lrintf(v) : INT_MIN; }
}  // namespace shim_detail

static inline void remap_linear_8u(const uint8_t* src, int sw, int sh, int sstride, const float* mapx, const float* mapy, int dw, int dh,
                                   uint8_t* dst, int dstride) {
  for (int y = 0; y < dh; ++y)
    for (int x = 0; x < dw; ++x) {
      const int fsx = shim_detail::round_i32(mapx[(size_t)y * dw + x] * 32.f), fsy = shim_detail::round_i32(mapy[(size_t)y * dw + x] * 32.f);
      const int sx = shim_detail::sat_short_i(fsx >> 5), sy = shim_detail::sat_short_i(fsy >> 5);
      int w[4];
      shim_detail::remap_weights(fsx & 31, fsy & 31, w);
      int v[4];
      for (int k = 0; k < 4; ++k) {
        const int px = sx + (k & 1), py = sy + (k >> 1);
        v[k] = (px >= 0 && px < sw && py >= 0 && py < sh) ? src[(size_t)py * sstride + px] : 0;
      }
      const int acc = v[0] * w[0] + v[1] * w[1] + v[2] * w[2] + v[3] * w[3];
      int r = (acc + (1 << 14)) >> 15;
      dst[(size_t)y * dstride + x] = (uint8_t)(r < 0 ? 0 : (r > 255 ? 255 : r));
    }
}

static inline void resize(InputArray _src, OutputArray _dst, Size dsize, double fx = 0, double fy = 0,
                          int interpolation = INTER_LINEAR) {
  (void)fx; (void)fy;
  assert(interpolation == INTER_LINEAR);
  (void)interpolation;
  Mat src = _src.getMat();
  _dst.create(dsize, CV_8UC1);
  Mat dst = _dst.getMat();
  const int SW = src.cols, SH = src.rows, DW = dsize.width, DH = dsize.height;
  if (SW == 2 * DW && SH == 2 * DH) {
    // OpenCV switches an exact 2x linear shrink to the INTER_AREA 2x2 box filter
    for (int y = 0; y < DH; ++y) {
      const uchar* s0 = src.ptr(2 * y); const uchar* s1 = src.ptr(2 * y + 1);
      uchar* d = dst.ptr(y);
      for (int x = 0; x < DW; ++x) d[x] = (uchar)((s0[2 * x] + s0[2 * x + 1] + s1[2 * x] + s1[2 * x + 1] + 2) >> 2);
    }
    return;
  }
  shim_detail::AxisTab tx = shim_detail::axis_tab_h(SW, DW), ty = shim_detail::axis_tab_v(SH, DH);
  std::vector<int> h0(DW), h1(DW);
  auto hrow = [&](int sy, std::vector<int>& out) {
    const uchar* s = src.ptr(sy);
    for (int d = 0; d < DW; ++d) {
      int sx = tx.ofs[d];
      int sx1 = std::min(sx + 1, SW - 1);
      out[d] = s[sx] * tx.c0[d] + s[sx1] * tx.c1[d];
    }
  };
  for (int y = 0; y < DH; ++y) {
    int sy0 = std::min(std::max(ty.ofs[y], 0), SH - 1);
    int sy1 = std::min(std::max(ty.ofs[y] + 1, 0), SH - 1);
    hrow(sy0, h0);
    hrow(sy1, h1);
    int b0 = ty.c0[y], b1 = ty.c1[y];
    uchar* d = dst.ptr(y);
    for (int x = 0; x < DW; ++x) {
      int v = (((b0 * (h0[x] >> 4)) >> 16) + ((b1 * (h1[x] >> 4)) >> 16) + 2) >> 2;
      d[x] = (uchar)std::min(std::max(v, 0), 255);
    }
  }
}

// ---- cv::GaussianBlur(7x7, sigma 2, BORDER_REFLECT_101), 8UC1 ----------
// 8.8 fixed-point separable kernel [18 34 48 56 48 34 18]/256, one rounding after both passes.
static inline void GaussianBlur(InputArray _src, OutputArray _dst, Size ksize, double sx, double sy = 0,
                                int borderType = BORDER_DEFAULT) {
  assert(ksize.width == 7 && ksize.height == 7 && sx == 2.0 && (sy == 2.0 || sy == 0.0));
  assert((borderType & ~BORDER_ISOLATED) == BORDER_REFLECT_101);
  (void)ksize; (void)sx; (void)sy; (void)borderType;
  Mat src = _src.getMat().clone();  // in-place calls are legal
  _dst.create(src.rows, src.cols, CV_8UC1);
  Mat dst = _dst.getMat();
  const int W = src.cols, H = src.rows;
  // horizontal pass into a 16-bit buffer (max 255 * 256), interior without border look-ups
  std::vector<unsigned short> hbuf((size_t)W * H);
  for (int y = 0; y < H; ++y) {
    const uchar* s = src.ptr(y);
    unsigned short* hrow = &hbuf[(size_t)y * W];
    for (int x = 0; x < W; ++x) {
      if (x >= 3 && x + 3 < W) {
        hrow[x] = (unsigned short)(18 * (s[x - 3] + s[x + 3]) + 34 * (s[x - 2] + s[x + 2]) + 48 * (s[x - 1] + s[x + 1]) + 56 * s[x]);
      } else {
        static const int k[7] = {18, 34, 48, 56, 48, 34, 18};
        int acc = 0;
        for (int i = 0; i < 7; ++i) acc += k[i] * s[reflect101(x + i - 3, W)];
        hrow[x] = (unsigned short)acc;
      }
    }
  }
  for (int y = 0; y < H; ++y) {
    uchar* d = dst.ptr(y);
    const unsigned short* r[7];
    for (int j = 0; j < 7; ++j) r[j] = &hbuf[(size_t)reflect101(y + j - 3, H) * W];
    for (int x = 0; x < W; ++x) {
      const int acc = 18 * (r[0][x] + r[6][x]) + 34 * (r[1][x] + r[5][x]) + 48 * (r[2][x] + r[4][x]) + 56 * r[3][x];
      d[x] = (uchar)((acc + 32768) >> 16);
    }
  }
}

// ---- cv::FAST(img, kps, threshold, true): FAST-9/16 + 3x3 strict NMS ---
namespace shim_detail {
static const int ring_dx[16] = {0, 1, 2, 3, 3, 3, 2, 1, 0, -1, -2, -3, -3, -3, -2, -1};
static const int ring_dy[16] = {3, 3, 2, 1, 0, -1, -2, -3, -3, -3, -2, -1, 0, 1, 2, 3};
// max threshold for which p stays a corner, minus nothing: returns max over 9-arcs of min(d) for both
// polarities, minus 1 (0 if p is not a corner for any threshold >= 0)
static inline int fast_score(const uchar* p, size_t step) {
  int d[16];
  const int v = p[0];
  for (int k = 0; k < 16; ++k) d[k] = v - p[(ptrdiff_t)ring_dy[k] * (ptrdiff_t)step + ring_dx[k]];
  // max over the 16 arcs of 9 of min(d) and of min(-d): window minima by doubling (2, 4, 8, then +1)
  int best = INT_MIN;
  for (int pol = 0; pol < 2; ++pol) {
    int e[16], m2[16], m4[16], m8[16];
    for (int k = 0; k < 16; ++k) e[k] = pol ? -d[k] : d[k];
    for (int k = 0; k < 16; ++k) m2[k] = std::min(e[k], e[(k + 1) & 15]);
    for (int k = 0; k < 16; ++k) m4[k] = std::min(m2[k], m2[(k + 2) & 15]);
    for (int k = 0; k < 16; ++k) m8[k] = std::min(m4[k], m4[(k + 4) & 15]);
    for (int k = 0; k < 16; ++k) best = std::max(best, std::min(m8[k], e[(k + 8) & 15]));
  }
  return best - 1;
}
}  // namespace shim_detail

static inline void FAST(InputArray _img, std::vector<KeyPoint>& kps, int threshold, bool nms = true) {
  Mat img = _img.getMat();
  kps.clear();
  const int W = img.cols, H = img.rows;
  if (W < 7 || H < 7) return;
  threshold = std::min(std::max(threshold, 0), 255);
  std::vector<int> score((size_t)W * H, 0);
  const ptrdiff_t st = (ptrdiff_t)img.step.v;
  ptrdiff_t ring_off[16];
  for (int k = 0; k < 16; ++k) ring_off[k] = (ptrdiff_t)shim_detail::ring_dy[k] * st + shim_detail::ring_dx[k];
  for (int y = 3; y < H - 3; ++y) {
    const uchar* row = img.ptr(y);
    for (int x = 3; x < W - 3; ++x) {
      // cheap exact pre-test (like OpenCV's own): a 9-arc of the 16-ring holds at least two of the four compass
      // points, so fewer than two brighter and fewer than two darker ones rule the pixel out
      const uchar* p = row + x;
      const int v = p[0], hi = v + threshold, lo = v - threshold;
      const int r0 = p[3 * st], r4 = p[3], r8 = p[-3 * st], r12 = p[-3];
      const int nb = (r0 > hi) + (r4 > hi) + (r8 > hi) + (r12 > hi);
      const int nd = (r0 < lo) + (r4 < lo) + (r8 < lo) + (r12 < lo);
      if (nb < 2 && nd < 2) continue;
      {  // is there an arc of 9 contiguous ring pixels all brighter than hi or all darker than lo?
        unsigned mb = 0, md = 0;
        for (int k = 0; k < 16; ++k) {
          const int r = p[ring_off[k]];
          mb |= (unsigned)(r > hi) << k;
          md |= (unsigned)(r < lo) << k;
        }
        auto arc9 = [](unsigned m16) {
          const unsigned d = m16 | (m16 << 16);
          unsigned m = d & (d >> 1);
          m &= m >> 2; m &= m >> 4; m &= d >> 8;
          return (m & 0xffffu) != 0;
        };
        if (!arc9(mb) && !arc9(md)) continue;
      }
      int s = shim_detail::fast_score(p, img.step.v);
      // corner at threshold t  <=>  9 contiguous ring pixels all differ from the centre by more than t
      if (s >= threshold) score[(size_t)y * W + x] = s;
    }
  }
  for (int y = 3; y < H - 3; ++y)
    for (int x = 3; x < W - 3; ++x) {
      int s = score[(size_t)y * W + x];
      if (s < threshold || s == 0) continue;  // not a corner: the score buffer holds 0 there
      bool keep = true;
      if (nms) {
        for (int dy = -1; dy <= 1 && keep; ++dy)
          for (int dx = -1; dx <= 1; ++dx) {
            if (!dx && !dy) continue;
            if (!(s > score[(size_t)(y + dy) * W + (x + dx)])) { keep = false; break; }
          }
      }
      if (keep) kps.push_back(KeyPoint((float)x, (float)y, 7.f, -1.f, (float)s));
    }
}

// ---- cv::fastAtan2: degree-valued 7th-order odd polynomial ---------------
static inline float fastAtan2(float y, float x) {
  static const float p1 = 0.9997878412794807f * (float)(180 / CV_PI);
  static const float p3 = -0.3258083974640975f * (float)(180 / CV_PI);
  static const float p5 = 0.1555786518463281f * (float)(180 / CV_PI);
  static const float p7 = -0.04432655554792128f * (float)(180 / CV_PI);
  float ax = std::abs(x), ay = std::abs(y);
  float a, c, c2;
  if (ax >= ay) {
    c = ay / (ax + (float)DBL_EPSILON);
    c2 = c * c;
    a = (((p7 * c2 + p5) * c2 + p3) * c2 + p1) * c;
  } else {
    c = ax / (ay + (float)DBL_EPSILON);
    c2 = c * c;
    a = 90.f - (((p7 * c2 + p5) * c2 + p3) * c2 + p1) * c;
  }
  if (x < 0) a = 180.f - a;
  if (y < 0) a = 360.f - a;
  return a;
}

// ---- cv::norm(a, b, NORM_L1) on 8U: exact integer sum of absolute differences -------------------
static inline double norm(InputArray _a, InputArray _b, int normType) {
  assert(normType == NORM_L1);
  (void)normType;
  Mat a = _a.getMat(), b = _b.getMat();
  long long s = 0;
  for (int y = 0; y < a.rows; ++y) {
    const uchar* pa = a.ptr(y); const uchar* pb = b.ptr(y);
    for (int x = 0; x < a.cols; ++x) s += std::abs((int)pa[x] - (int)pb[x]);
  }
  return (double)s;
}

// ---- cv::BFMatcher(NORM_HAMMING).knnMatch ------------------------------
struct DMatch {
  int queryIdx, trainIdx, imgIdx;
  float distance;
  DMatch() : queryIdx(-1), trainIdx(-1), imgIdx(-1), distance(FLT_MAX) {}
  DMatch(int q, int t, float d) : queryIdx(q), trainIdx(t), imgIdx(0), distance(d) {}
};

class BFMatcher {
 public:
  int normType;
  BFMatcher(int nt = NORM_HAMMING) : normType(nt) {}
  // ascending train index scan, insert when d < current k-th best: ties keep the lower train index
  void knnMatch(InputArray _q, InputArray _t, std::vector<std::vector<DMatch>>& matches, int k) const {
    Mat q = _q.getMat(), t = _t.getMat();
    matches.assign(q.rows, std::vector<DMatch>());
    for (int i = 0; i < q.rows; ++i) {
      std::vector<DMatch>& m = matches[i];
      for (int j = 0; j < t.rows; ++j) {
        int d = 0;
        const uchar* a = q.ptr(i); const uchar* b = t.ptr(j);
        for (int c = 0; c < q.cols; ++c) d += __builtin_popcount((unsigned)(a[c] ^ b[c]));
        if ((int)m.size() < k || (float)d < m.back().distance) {
          if ((int)m.size() < k) m.push_back(DMatch());
          int pos = (int)m.size() - 1;
          while (pos > 0 && m[pos - 1].distance > (float)d) { m[pos] = m[pos - 1]; --pos; }
          m[pos] = DMatch(i, j, (float)d);
        }
      }
    }
  }
};

struct KeyPointsFilter {
  // only has to compile (dead code in the reference, src/ORBextractor.cc:972,987)
  static void retainBest(std::vector<KeyPoint>& kps, int n) {
    if (n >= 0 && (int)kps.size() > n) {
      std::stable_sort(kps.begin(), kps.end(), [](const KeyPoint& a, const KeyPoint& b) { return a.response > b.response; });
      kps.resize(n);
    }
  }
};

}  // namespace cv
