// Minimal fixed-size Eigen stand-in for environments without Eigen (like this build image).
// When real Eigen is on the include path, define SFX_USE_EIGEN and <Eigen/Core> is used instead;
// the sym:: layer only needs: fixed-size column-major Matrix<Scalar,R,C> with data(), operator(),
// operator[], Zero(), Identity(), Constant(), Ones(), norm(), size(), RowsAtCompileTime / ColsAtCompileTime.
#pragma once
#if defined(SFX_USE_EIGEN)
#include <Eigen/Core>
#else
#include <cmath>
#include <cstring>
#include <initializer_list>
namespace Eigen {
template <typename Scalar_, int R, int C>
class Matrix {
 public:
  using Scalar = Scalar_;
  enum { RowsAtCompileTime = R, ColsAtCompileTime = C, SizeAtCompileTime = R * C };
  Matrix() { std::memset(d_, 0, sizeof(d_)); }
  template <typename... T, typename = typename std::enable_if<sizeof...(T) == R * C && (R * C > 1)>::type>
  Matrix(T... v) : d_{static_cast<Scalar>(v)...} {}
  static Matrix Zero() { return Matrix(); }
  static Matrix Constant(Scalar v) {
    Matrix m;
    for (int i = 0; i < R * C; ++i) m.d_[i] = v;
    return m;
  }
  static Matrix Ones() { return Constant(Scalar(1)); }
  Scalar squaredNorm() const {
    Scalar s = 0;
    for (int i = 0; i < R * C; ++i) s += d_[i] * d_[i];
    return s;
  }
  Scalar norm() const { return std::sqrt(squaredNorm()); }
  static Matrix Identity() {
    Matrix m;
    for (int i = 0; i < (R < C ? R : C); ++i) m(i, i) = 1;
    return m;
  }
  static Matrix FromData(const Scalar* p) {
    Matrix m;
    std::memcpy(m.d_, p, sizeof(m.d_));
    return m;
  }
  Scalar& operator()(int r, int c) { return d_[r + c * R]; }
  const Scalar& operator()(int r, int c) const { return d_[r + c * R]; }
  Scalar& operator[](int i) { return d_[i]; }
  const Scalar& operator[](int i) const { return d_[i]; }
  Scalar* data() { return d_; }
  const Scalar* data() const { return d_; }
  static constexpr int size() { return R * C; }
  Matrix operator*(Scalar s) const {
    Matrix m;
    for (int i = 0; i < R * C; ++i) m.d_[i] = d_[i] * s;
    return m;
  }
  Matrix operator/(Scalar s) const {
    Matrix m;
    for (int i = 0; i < R * C; ++i) m.d_[i] = d_[i] / s;
    return m;
  }
  Matrix operator+(const Matrix& o) const {
    Matrix m;
    for (int i = 0; i < R * C; ++i) m.d_[i] = d_[i] + o.d_[i];
    return m;
  }

 private:
  Scalar d_[R * C];
};
using Vector2d = Matrix<double, 2, 1>;
using Vector3d = Matrix<double, 3, 1>;
using Vector4d = Matrix<double, 4, 1>;
}  // namespace Eigen
#endif
