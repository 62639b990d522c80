// Stand-in for gen/cpp/sym/pose3.h: the generated factor headers only call .Data().
#pragma once
#include <Eigen/Core>
namespace sym {
template <typename Scalar>
class Pose3 {
 public:
  using DataVec = Eigen::Matrix<Scalar, 7, 1>;
  explicit Pose3(const Scalar* p) : data_(p) {}
  const DataVec& Data() const { return data_; }
 private:
  DataVec data_;
};
}  // namespace sym
