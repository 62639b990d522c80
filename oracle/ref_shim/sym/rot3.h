// Stand-in for gen/cpp/sym/rot3.h: the generated factor headers only call .Data().
#pragma once
#include <Eigen/Core>
namespace sym {
template <typename Scalar>
class Rot3 {
 public:
  using DataVec = Eigen::Matrix<Scalar, 4, 1>;
  explicit Rot3(const Scalar* p) : data_(p) {}
  const DataVec& Data() const { return data_; }
 private:
  DataVec data_;
};
}  // namespace sym
