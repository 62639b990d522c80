// Stand-in for gen/cpp/sym/linear_camera_cal.h: the generated factor headers only call .Data().
#pragma once
#include <Eigen/Core>
namespace sym {
template <typename Scalar>
class LinearCameraCal {
 public:
  using DataVec = Eigen::Matrix<Scalar, 4, 1>;
  explicit LinearCameraCal(const Scalar* p) : data_(p) {}
  const DataVec& Data() const { return data_; }
 private:
  DataVec data_;
};
}  // namespace sym
