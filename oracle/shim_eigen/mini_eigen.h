// TEST INFRASTRUCTURE ONLY. A value-semantics stand-in for the handful of Eigen expressions that
// KannalaBrandt8::TriangulateMatches / Triangulate / project / unprojectEig use (reference
// src/CameraModels/KannalaBrandt8.cpp:68-94,111-114,323-395,415-428). Eigen (find_package(Eigen3 3.1.0), reference
// CMakeLists.txt) is an un-vendored dependency and absent from this image, so two things here are RESTATED, NOT PINNED:
//   * the float evaluation order of fixed-size products / reductions: coefficient-wise, a size-3 sum as a0 + (a1 + a2)
//     (Eigen 3's unrolled non-vectorised reduction splits the range in halves; written from memory, unverifiable here);
//   * Eigen::JacobiSVD<Matrix4f>: Eigen's published algorithm (two-sided Jacobi, float) restated in orb_oracle_kb8.h
//     (orb_eigen_jacobi_svd4f), checked against numpy.linalg.svd and, as the whole DLT step, against cv2.triangulatePoints in
//     tests/test_oracle_kb8.py. Only matrixV().col(3) (the smallest singular value's vector) is used.
// The CUDA kernel runs the same restatements, so CUDA == this stand-in bit for bit; whether the stand-in == real Eigen cannot be
// checked in this image.
#pragma once
#include <cmath>
#include <type_traits>
#include "orb_oracle_kb8.h"

namespace Eigen {
enum { ComputeFullV = 16 };

template <int N>
static inline float redux_sum(const float* v) {   // halves, like Eigen's unrolled reduction
  if constexpr (N == 1) return v[0];
  else return redux_sum<N / 2>(v) + redux_sum<N - N / 2>(v + N / 2);
}

template <int R, int C>
struct Mat {
  float d[R * C];   // row-major
  Mat() { for (float& x : d) x = 0.f; }
  template <int Z = C, typename = typename std::enable_if<Z == 1 && R == 3>::type>
  Mat(float x, float y, float z) { d[0] = x; d[1] = y; d[2] = z; }
  template <int Z = C, typename = typename std::enable_if<Z == 1 && R == 2>::type>
  Mat(float x, float y) { d[0] = x; d[1] = y; }
  float& operator()(int i, int j) { return d[i * C + j]; }
  float operator()(int i, int j) const { return d[i * C + j]; }
  float& operator()(int i) { return d[i]; }
  float operator()(int i) const { return d[i]; }
  float& operator[](int i) { return d[i]; }
  float operator[](int i) const { return d[i]; }
  static Mat Identity() { Mat m; for (int i = 0; i < (R < C ? R : C); ++i) m(i, i) = 1.f; return m; }
  static Mat Zero() { return Mat(); }
  Mat<C, R> transpose() const { Mat<C, R> t; for (int i = 0; i < R; ++i) for (int j = 0; j < C; ++j) t(j, i) = (*this)(i, j); return t; }
  struct RowRef {
    Mat* m; int i;
    RowRef& operator=(const Mat<1, C>& r) { for (int j = 0; j < C; ++j) (*m)(i, j) = r.d[j]; return *this; }
    operator Mat<1, C>() const { Mat<1, C> r; for (int j = 0; j < C; ++j) r.d[j] = (*m)(i, j); return r; }
    template <int R2, int C2> float dot(const Mat<R2, C2>& o) const { return ((Mat<1, C>)*this).dot(o); }
  };
  RowRef row(int i) { return RowRef{this, i}; }
  Mat<1, C> row(int i) const { Mat<1, C> r; for (int j = 0; j < C; ++j) r.d[j] = (*this)(i, j); return r; }
  Mat<R, 1> col(int j) const { Mat<R, 1> c; for (int i = 0; i < R; ++i) c.d[i] = (*this)(i, j); return c; }
  template <int R2, int C2>
  float dot(const Mat<R2, C2>& o) const {
    static_assert(R2 * C2 == R * C, "dot: sizes");
    float p[R * C];
    for (int i = 0; i < R * C; ++i) p[i] = d[i] * o.d[i];
    return redux_sum<R * C>(p);
  }
  float norm() const { return std::sqrt(this->dot(*this)); }
  Mat<3, 1> head(int n) const { Mat<3, 1> h; for (int i = 0; i < 3 && i < n; ++i) h.d[i] = d[i]; return h; }
  // comma initialiser: blocks fill the matrix left to right, then top to bottom (block rows)
  struct Comma {
    Mat* m; int r0, c0, rh;
    template <int R2, int C2>
    Comma& operator,(const Mat<R2, C2>& b) {
      if (c0 == C) { r0 += rh; c0 = 0; }
      for (int i = 0; i < R2; ++i) for (int j = 0; j < C2; ++j) (*m)(r0 + i, c0 + j) = b(i, j);
      c0 += C2; rh = R2;
      return *this;
    }
  };
  template <int R2, int C2>
  Comma operator<<(const Mat<R2, C2>& b) { Comma c{this, 0, 0, R2}; c, b; return c; }
};

template <int R, int K, int C>
static inline Mat<R, C> operator*(const Mat<R, K>& a, const Mat<K, C>& b) {
  Mat<R, C> r;
  for (int i = 0; i < R; ++i)
    for (int j = 0; j < C; ++j) {
      float p[K];
      for (int k = 0; k < K; ++k) p[k] = a(i, k) * b(k, j);
      r(i, j) = redux_sum<K>(p);
    }
  return r;
}
template <int R, int C> static inline Mat<R, C> operator+(const Mat<R, C>& a, const Mat<R, C>& b) { Mat<R, C> r; for (int i = 0; i < R * C; ++i) r.d[i] = a.d[i] + b.d[i]; return r; }
template <int R, int C> static inline Mat<R, C> operator-(const Mat<R, C>& a, const Mat<R, C>& b) { Mat<R, C> r; for (int i = 0; i < R * C; ++i) r.d[i] = a.d[i] - b.d[i]; return r; }
template <int R, int C> static inline Mat<R, C> operator-(const Mat<R, C>& a) { Mat<R, C> r; for (int i = 0; i < R * C; ++i) r.d[i] = -a.d[i]; return r; }
template <int R, int C> static inline Mat<R, C> operator*(float s, const Mat<R, C>& a) { Mat<R, C> r; for (int i = 0; i < R * C; ++i) r.d[i] = s * a.d[i]; return r; }
template <int R, int C> static inline Mat<R, C> operator/(const Mat<R, C>& a, float s) { Mat<R, C> r; for (int i = 0; i < R * C; ++i) r.d[i] = a.d[i] / s; return r; }

template <typename T, int R, int C> struct MatrixSel;
template <int R, int C> struct MatrixSel<float, R, C> { typedef Mat<R, C> type; };
template <typename T, int R, int C> using Matrix = typename MatrixSel<T, R, C>::type;
typedef Mat<3, 3> Matrix3f;
typedef Mat<4, 4> Matrix4f;
typedef Mat<3, 1> Vector3f;
typedef Mat<2, 1> Vector2f;
typedef Mat<4, 1> Vector4f;

template <typename M>
struct JacobiSVD {
  Matrix4f V;
  JacobiSVD(const Matrix4f& A, int) {
    orb_eigen_jacobi_svd4f(A.d, V.d);   // Eigen's two-sided Jacobi SVD in float, restated (orb_oracle_kb8.h)
  }
  const Matrix4f& matrixV() const { return V; }
};
}  // namespace Eigen
