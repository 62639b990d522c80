// sym:: header layer -- the reference's C++ surface for the sparse LM path, lowered onto the C ABI
// of include/sfx.h (libsfx.so: CUDA sm_100a).  Same class / function names, argument meaning and
// error behaviour as the reference for THIS path:
//   sym::Key                         symforce/opt/key.h:26-103
//   sym::Values<Scalar>              symforce/opt/values.h:31-324   (Set / At / Has / CreateIndex / Data ...)
//   sym::Factor<Scalar>::Hessian     symforce/opt/factor.h:231-234  (function-pointer flavour)
//   sym::optimizer_params_t, DefaultOptimizerParams()   lcmtypes/symforce.lcm:134-200, opt/optimizer.cc:8-56
//   sym::OptimizationStats           symforce/opt/optimization_stats.h:22-94
//   sym::Optimizer<Scalar>           symforce/opt/optimizer.h:72-325
//   sym::Optimize(params, factors, values, eps)          symforce/opt/optimizer.h:334-340
// What is different, by design: factors must be generated functions with a device implementation
// (all of their inputs are Values keys); anything else is a std::runtime_error at Factor creation --
// there is no CPU fallback.  `Optimize()` moves Values::Data() to the GPU, runs the whole LM loop
// there and copies the best values + per-iteration stats back.
#pragma once
#include <algorithm>
#include <array>
#include <chrono>
#include <cstdio>
#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <exception>
#include <limits>
#include <map>
#include <memory>
#include <optional>
#include <random>
#include <sstream>
#include <stdexcept>
#include <string>
#include <thread>
#include <unordered_map>
#include <unordered_set>
#include <utility>
#include <vector>

#include "../sfx.h"
#include "mini_eigen.h"

// SYM_ASSERT -> std::runtime_error (symforce/opt/assert.h:40-92)
#define SYM_ASSERT(expr, ...)                                                                   \
  do {                                                                                          \
    if (!(expr)) {                                                                              \
      std::ostringstream o__;                                                                   \
      o__ << "SYM_ASSERT: " #expr "\n    --> " << __func__ << "\n    --> " << __FILE__ << ":" << __LINE__; \
      throw std::runtime_error(o__.str());                                                      \
    }                                                                                           \
  } while (0)

namespace sym {

template <typename Scalar>
constexpr Scalar kDefaultEpsilon = Scalar(10) * std::numeric_limits<Scalar>::epsilon();  // sym/util/epsilon.h:31
constexpr double kDefaultEpsilond = kDefaultEpsilon<double>;

template <typename S>
using Vector1 = Eigen::Matrix<S, 1, 1>;
template <typename S>
using Vector2 = Eigen::Matrix<S, 2, 1>;
template <typename S>
using Vector3 = Eigen::Matrix<S, 3, 1>;
template <typename S>
using Vector4 = Eigen::Matrix<S, 4, 1>;
template <typename S>
using Vector5 = Eigen::Matrix<S, 5, 1>;
template <typename S>
using Vector6 = Eigen::Matrix<S, 6, 1>;
template <typename S>
using Vector7 = Eigen::Matrix<S, 7, 1>;
template <typename S>
using Matrix33 = Eigen::Matrix<S, 3, 3>;
template <typename S>
using Matrix66 = Eigen::Matrix<S, 6, 6>;
using Vector3d = Vector3<double>;
using Vector5d = Vector5<double>;
using Vector6d = Vector6<double>;
using Vector7d = Vector7<double>;
using Matrix66d = Matrix66<double>;

// sym::Random<Matrix>(gen) (gen/cpp/sym/ops/matrix/storage_ops.h:211-218): a fresh standard normal distribution per
// call, entries drawn in storage order -- what the reference's tests build their sample data with
template <typename T, typename Generator>
T Random(Generator& gen) {
  std::normal_distribution<typename T::Scalar> distribution{};
  T m;
  for (int i = 0; i < static_cast<int>(T::SizeAtCompileTime); ++i) m.data()[i] = distribution(gen);
  return m;
}

// ------------------------------------------------------------------------------------------------
// Key (symforce/opt/key.h)
// ------------------------------------------------------------------------------------------------
class Key {
 public:
  using subscript_t = std::int64_t;
  using superscript_t = std::int64_t;
  static constexpr char kInvalidLetter = static_cast<char>(0);
  static constexpr subscript_t kInvalidSub = std::numeric_limits<subscript_t>::min();
  static constexpr superscript_t kInvalidSuper = std::numeric_limits<superscript_t>::min();

  constexpr Key(const char letter, const subscript_t sub = kInvalidSub, const superscript_t super = kInvalidSuper)
      : letter_(letter), sub_(sub), super_(super) {}
  constexpr Key() = default;
  constexpr char Letter() const { return letter_; }
  constexpr subscript_t Sub() const { return sub_; }
  constexpr superscript_t Super() const { return super_; }
  constexpr Key WithLetter(const char letter) const { return Key(letter, sub_, super_); }
  constexpr Key WithSub(const subscript_t sub) const { return Key(letter_, sub, super_); }
  constexpr Key WithSuper(const superscript_t super) const { return Key(letter_, sub_, super); }
  constexpr bool operator==(const Key& o) const { return letter_ == o.letter_ && sub_ == o.sub_ && super_ == o.super_; }
  constexpr bool operator!=(const Key& o) const { return !(*this == o); }
  // symforce/opt/key.cc:18-21
  static bool LexicalLessThan(const Key& a, const Key& b) {
    return std::make_tuple(a.Letter(), a.Sub(), a.Super()) < std::make_tuple(b.Letter(), b.Sub(), b.Super());
  }
  std::string str() const {
    if (letter_ == kInvalidLetter) return "NULLKEY";
    std::ostringstream os;
    os << letter_;
    if (sub_ != kInvalidSub) os << '_' << (sub_ < 0 ? "n" : "") << std::llabs(sub_);
    if (super_ != kInvalidSuper) os << '_' << (super_ < 0 ? "n" : "") << std::llabs(super_);
    return os.str();
  }

 private:
  char letter_{kInvalidLetter};
  subscript_t sub_{kInvalidSub};
  superscript_t super_{kInvalidSuper};
};
struct KeyHash {
  std::size_t operator()(const Key& k) const {
    std::size_t h = std::hash<char>()(k.Letter());
    h ^= std::hash<std::int64_t>()(k.Sub()) + 0x9e3779b97f4a7c15ull + (h << 6) + (h >> 2);
    h ^= std::hash<std::int64_t>()(k.Super()) + 0x9e3779b97f4a7c15ull + (h << 6) + (h >> 2);
    return h;
  }
};

// ------------------------------------------------------------------------------------------------
// Geo types needed by the path (storage + retract semantics live on the device; these are the
// host-side value types of gen/cpp/sym/{rot3,pose3,linear_camera_cal}.h)
// ------------------------------------------------------------------------------------------------
template <typename Scalar>
class Rot3 {
 public:
  using DataVec = Eigen::Matrix<Scalar, 4, 1>;
  Rot3() { data_[3] = 1; }
  explicit Rot3(const DataVec& d, bool normalize = true) : data_(d) {
    if (normalize) Normalize();
  }
  static Rot3 Identity() { return Rot3(); }
  // gen/cpp/sym/ops/rot3/lie_group_ops.cc:15-37
  static Rot3 FromTangent(const Vector3<Scalar>& v, Scalar epsilon = kDefaultEpsilon<Scalar>) {
    const Scalar t0 = std::sqrt(epsilon * epsilon + v[0] * v[0] + v[1] * v[1] + v[2] * v[2]);
    const Scalar t1 = Scalar(0.5) * t0, s = std::sin(t1) / t0;
    DataVec d;
    d[0] = s * v[0];
    d[1] = s * v[1];
    d[2] = s * v[2];
    d[3] = std::cos(t1);
    return Rot3(d);
  }
  Rot3 Compose(const Rot3& b) const {
    const DataVec& a = data_;
    DataVec r;
    r[0] = a[3] * b.data_[0] + a[0] * b.data_[3] + a[1] * b.data_[2] - a[2] * b.data_[1];
    r[1] = a[3] * b.data_[1] - a[0] * b.data_[2] + a[1] * b.data_[3] + a[2] * b.data_[0];
    r[2] = a[3] * b.data_[2] + a[0] * b.data_[1] - a[1] * b.data_[0] + a[2] * b.data_[3];
    r[3] = a[3] * b.data_[3] - a[0] * b.data_[0] - a[1] * b.data_[1] - a[2] * b.data_[2];
    return Rot3(r);
  }
  Rot3 Retract(const Vector3<Scalar>& v, Scalar epsilon = kDefaultEpsilon<Scalar>) const {
    return Compose(Rot3(FromTangent(v, epsilon).Data(), false));
  }
  Rot3 Inverse() const {
    DataVec d = data_;
    d[0] = -d[0];
    d[1] = -d[1];
    d[2] = -d[2];
    return Rot3(d, false);
  }
  Rot3 Between(const Rot3& b) const { return Inverse().Compose(b); }
  // gen/cpp/sym/ops/rot3/lie_group_ops.cc:39-60: log map with |w| clamped to 1 - epsilon, sign taken from w
  Vector3<Scalar> ToTangent(Scalar epsilon = kDefaultEpsilon<Scalar>) const {
    const Scalar w = data_[3];
    const Scalar wc = std::min<Scalar>(std::fabs(w), Scalar(1) - epsilon);
    const Scalar sign = w < 0 ? Scalar(-1) : Scalar(1);
    const Scalar s = 2 * sign * std::acos(wc) / std::sqrt(Scalar(1) - wc * wc);
    Vector3<Scalar> v;
    for (int i = 0; i < 3; ++i) v[i] = s * data_[i];
    return v;
  }
  // tangent v with this->Retract(v) == b
  Vector3<Scalar> LocalCoordinates(const Rot3& b, Scalar epsilon = kDefaultEpsilon<Scalar>) const {
    return Between(b).ToTangent(epsilon);
  }
  Vector3<Scalar> Rotate(const Vector3<Scalar>& p) const {
    const Scalar x = data_[0], y = data_[1], z = data_[2], w = data_[3];
    const Scalar tx = 2 * (y * p[2] - z * p[1]), ty = 2 * (z * p[0] - x * p[2]), tz = 2 * (x * p[1] - y * p[0]);
    Vector3<Scalar> r;
    r[0] = p[0] + w * tx + (y * tz - z * ty);
    r[1] = p[1] + w * ty + (z * tx - x * tz);
    r[2] = p[2] + w * tz + (x * ty - y * tx);
    return r;
  }
  const DataVec& Data() const { return data_; }

 private:
  void Normalize() {
    const Scalar n2 = data_[0] * data_[0] + data_[1] * data_[1] + data_[2] * data_[2] + data_[3] * data_[3];
    if (n2 > 0) {
      const Scalar n = std::sqrt(n2);
      for (int i = 0; i < 4; ++i) data_[i] /= n;
    }
  }
  DataVec data_;
};

template <typename Scalar>
class Pose3 {
 public:
  using DataVec = Eigen::Matrix<Scalar, 7, 1>;
  Pose3() { data_[3] = 1; }
  explicit Pose3(const DataVec& d, bool normalize = true) : data_(d) {
    if (normalize) {
      typename Rot3<Scalar>::DataVec q;
      for (int i = 0; i < 4; ++i) q[i] = d[i];
      Rot3<Scalar> r(q, true);
      for (int i = 0; i < 4; ++i) data_[i] = r.Data()[i];
    }
  }
  Pose3(const Rot3<Scalar>& R, const Vector3<Scalar>& t) {
    for (int i = 0; i < 4; ++i) data_[i] = R.Data()[i];
    for (int i = 0; i < 3; ++i) data_[4 + i] = t[i];
  }
  static Pose3 Identity() { return Pose3(); }
  Rot3<Scalar> Rotation() const {
    typename Rot3<Scalar>::DataVec q;
    for (int i = 0; i < 4; ++i) q[i] = data_[i];
    return Rot3<Scalar>(q, false);
  }
  Vector3<Scalar> Position() const {
    Vector3<Scalar> t;
    for (int i = 0; i < 3; ++i) t[i] = data_[4 + i];
    return t;
  }
  // gen/cpp/sym/ops/pose3/lie_group_ops.cc:70-103: rotation retracts on the right, translation adds
  Pose3 Retract(const Vector6<Scalar>& v, Scalar epsilon = kDefaultEpsilon<Scalar>) const {
    Vector3<Scalar> w, t = Position();
    for (int i = 0; i < 3; ++i) {
      w[i] = v[i];
      t[i] += v[3 + i];
    }
    return Pose3(Rotation().Retract(w, epsilon), t);
  }
  Pose3 Inverse() const {
    const Rot3<Scalar> Ri = Rotation().Inverse();
    const Vector3<Scalar> t = Ri.Rotate(Position());
    return Pose3(Ri, Vector3<Scalar>(-t[0], -t[1], -t[2]));
  }
  Pose3 Compose(const Pose3& b) const {
    const Vector3<Scalar> rt = Rotation().Rotate(b.Position()), t = Position();
    return Pose3(Rotation().Compose(b.Rotation()), Vector3<Scalar>(rt[0] + t[0], rt[1] + t[1], rt[2] + t[2]));
  }
  // gen/cpp/sym/ops/pose3/lie_group_ops.cc:105-140: [rotation local coordinates, translation difference]
  Vector6<Scalar> LocalCoordinates(const Pose3& b, Scalar epsilon = kDefaultEpsilon<Scalar>) const {
    const Vector3<Scalar> w = Rotation().LocalCoordinates(b.Rotation(), epsilon);
    Vector6<Scalar> v;
    for (int i = 0; i < 3; ++i) {
      v[i] = w[i];
      v[3 + i] = b.data_[4 + i] - data_[4 + i];
    }
    return v;
  }
  const DataVec& Data() const { return data_; }

 private:
  DataVec data_;
};

template <typename Scalar>
class LinearCameraCal {
 public:
  using DataVec = Eigen::Matrix<Scalar, 4, 1>;
  LinearCameraCal(const Vector2<Scalar>& focal, const Vector2<Scalar>& principal) {
    data_[0] = focal[0];
    data_[1] = focal[1];
    data_[2] = principal[0];
    data_[3] = principal[1];
  }
  explicit LinearCameraCal(const DataVec& d) : data_(d) {}
  const DataVec& Data() const { return data_; }

 private:
  DataVec data_;
};
using Rot3d = Rot3<double>;
using Pose3d = Pose3<double>;
using LinearCameraCald = LinearCameraCal<double>;

// ------------------------------------------------------------------------------------------------
// Values (symforce/opt/values.h): flat insertion-ordered data + key -> index entry
// ------------------------------------------------------------------------------------------------
enum class type_t : int32_t { INVALID = 0, SCALAR, ROT3, POSE3, VECTOR, MATRIX, CAMERA_CAL };

struct index_entry_t {
  Key key;
  type_t type{type_t::INVALID};
  int32_t offset{0};
  int32_t storage_dim{0};
  int32_t tangent_dim{0};
};
struct index_t {
  int32_t storage_dim{0};
  int32_t tangent_dim{0};
  std::vector<index_entry_t> entries;
};

namespace internal {
// fn(begin, end) over contiguous chunks of [0, n) on up to 8 host threads (SFX_HOST_THREADS overrides); read-only
// lookups in the key maps of a multi-million-factor problem are bound by cache misses and scale with threads
template <typename Fn>
inline void ParallelFor(size_t n, const Fn& fn) {
  size_t nt = std::min<size_t>(std::min<unsigned>(8u, std::max(1u, std::thread::hardware_concurrency())), (n + 65535) / 65536);
  if (const char* e = std::getenv("SFX_HOST_THREADS")) nt = static_cast<size_t>(std::max(1, std::atoi(e)));
  if (nt <= 1) {
    fn(size_t{0}, n);
    return;
  }
  std::vector<std::thread> th;
  std::vector<std::exception_ptr> err(nt);
  for (size_t t = 0; t < nt; ++t)
    th.emplace_back([&, t] {
      try {
        fn(n * t / nt, n * (t + 1) / nt);
      } catch (...) {
        err[t] = std::current_exception();
      }
    });
  for (auto& x : th) x.join();
  for (auto& e : err)
    if (e) std::rethrow_exception(e);
}
}  // namespace internal

namespace internal {
template <typename T, typename = void>
struct StorageTraits;  // storage pointer / dims / type of a value type
template <>
struct StorageTraits<double> {
  static constexpr int kStorage = 1, kTangent = 1;
  static constexpr type_t kType = type_t::SCALAR;
  static const double* Ptr(const double& v) { return &v; }
  static double From(const double* p) { return *p; }
};
template <typename S>
struct StorageTraits<Rot3<S>> {
  static constexpr int kStorage = 4, kTangent = 3;
  static constexpr type_t kType = type_t::ROT3;
  static const S* Ptr(const Rot3<S>& v) { return v.Data().data(); }
  static Rot3<S> From(const S* p) { return Rot3<S>(Rot3<S>::DataVec::FromData(p), false); }
};
template <typename S>
struct StorageTraits<Pose3<S>> {
  static constexpr int kStorage = 7, kTangent = 6;
  static constexpr type_t kType = type_t::POSE3;
  static const S* Ptr(const Pose3<S>& v) { return v.Data().data(); }
  static Pose3<S> From(const S* p) { return Pose3<S>(Pose3<S>::DataVec::FromData(p), false); }
};
template <typename S>
struct StorageTraits<LinearCameraCal<S>> {
  static constexpr int kStorage = 4, kTangent = 4;
  static constexpr type_t kType = type_t::CAMERA_CAL;
  static const S* Ptr(const LinearCameraCal<S>& v) { return v.Data().data(); }
  static LinearCameraCal<S> From(const S* p) { return LinearCameraCal<S>(LinearCameraCal<S>::DataVec::FromData(p)); }
};
template <typename S, int R, int C>
struct StorageTraits<Eigen::Matrix<S, R, C>> {
  static constexpr int kStorage = R * C, kTangent = R * C;
  static constexpr type_t kType = (C == 1) ? type_t::VECTOR : type_t::MATRIX;
  static const S* Ptr(const Eigen::Matrix<S, R, C>& v) { return v.data(); }
  static Eigen::Matrix<S, R, C> From(const S* p) { return Eigen::Matrix<S, R, C>::FromData(p); }
};
}  // namespace internal

template <typename Scalar>
class Values {
 public:
  using MapType = std::unordered_map<Key, index_entry_t, KeyHash>;

  bool Has(const Key& key) const { return map_.find(key) != map_.end(); }

  // Values::Set (values.tcc:60-100): new key appends, existing key overwrites (type must match)
  template <typename T>
  bool Set(const Key& key, const T& value) {
    using Tr = internal::StorageTraits<T>;
    auto it = map_.find(key);
    if (it == map_.end()) {
      index_entry_t e;
      e.key = key;
      e.type = Tr::kType;
      e.offset = static_cast<int32_t>(data_.size());
      e.storage_dim = Tr::kStorage;
      e.tangent_dim = Tr::kTangent;
      map_[key] = e;
      data_.insert(data_.end(), Tr::Ptr(value), Tr::Ptr(value) + Tr::kStorage);
      return true;
    }
    if (it->second.type != Tr::kType || it->second.storage_dim != Tr::kStorage)
      throw std::runtime_error("Trying to set key " + key.str() + " with a different type than it was created with");
    std::copy(Tr::Ptr(value), Tr::Ptr(value) + Tr::kStorage, data_.begin() + it->second.offset);
    return false;
  }
  bool Set(const Key& key, const int value) { return Set<double>(key, static_cast<double>(value)); }

  // Values::At (values.tcc:27-58): "Key not found" / type mismatch are std::runtime_error
  template <typename T>
  T At(const Key& key) const {
    using Tr = internal::StorageTraits<T>;
    const index_entry_t& e = IndexEntryAt(key);
    if (e.type != Tr::kType || e.storage_dim != Tr::kStorage)
      throw std::runtime_error("Mismatched types; index entry for key " + key.str() + " has a different type");
    return Tr::From(data_.data() + e.offset);
  }
  const index_entry_t& IndexEntryAt(const Key& key) const {
    auto it = map_.find(key);
    if (it == map_.end()) throw std::runtime_error("Key not found: " + key.str());
    return it->second;
  }
  // values.cc:210-234
  index_t CreateIndex(const std::vector<Key>& keys) const {
    index_t index;
    for (const Key& key : keys) {
      auto it = map_.find(key);
      if (it == map_.end()) throw std::runtime_error("Tried to create index for key " + key.str() + " not in values");
      index.entries.push_back(it->second);
      index.storage_dim += it->second.storage_dim;
      index.tangent_dim += it->second.tangent_dim;
    }
    return index;
  }
  size_t NumEntries() const { return map_.size(); }
  bool Empty() const { return map_.empty(); }
  const std::vector<Scalar>& Data() const { return data_; }
  Scalar* DataPointer() { return data_.data(); }
  const MapType& Items() const { return map_; }

  // values.cc:69-119: all keys, by default in storage order
  std::vector<Key> Keys(const bool sort_by_offset = true) const {
    std::vector<index_entry_t> entries = CreateIndex(sort_by_offset).entries;
    std::vector<Key> keys;
    keys.reserve(entries.size());
    for (const index_entry_t& e : entries) keys.push_back(e.key);
    return keys;
  }
  // values.cc:180-206: index of every key (tangent_dim is -1 when a key has no tangent space)
  index_t CreateIndex(const bool sort_by_offset) const {
    index_t index;
    index.entries.reserve(map_.size());
    for (const auto& kv : map_) {
      index.entries.push_back(kv.second);
      index.storage_dim += kv.second.storage_dim;
      if (index.tangent_dim >= 0) index.tangent_dim = kv.second.tangent_dim >= 0 ? index.tangent_dim + kv.second.tangent_dim : -1;
    }
    if (sort_by_offset)
      std::sort(index.entries.begin(), index.entries.end(),
                [](const index_entry_t& a, const index_entry_t& b) { return a.offset < b.offset; });
    return index;
  }
  std::optional<index_entry_t> MaybeIndexEntryAt(const Key& key) const {
    auto it = map_.find(key);
    if (it == map_.end()) return {};
    return it->second;
  }
  // values.tcc: SetNew asserts that the key is new
  template <typename T>
  void SetNew(const Key& key, const T& value) {
    const bool added = Set<T>(key, value);
    if (!added) throw std::runtime_error("SYM_ASSERT: added -- key " + key.str() + " already exists");
  }
  // At / Set through an index entry (values.tcc:102-140)
  template <typename T>
  T At(const index_entry_t& entry) const {
    using Tr = internal::StorageTraits<T>;
    if (entry.type != Tr::kType || entry.storage_dim != Tr::kStorage)
      throw std::runtime_error("Mismatched types; index entry is type " + std::to_string(static_cast<int>(entry.type)));
    return Tr::From(data_.data() + entry.offset);
  }
  template <typename T>
  void Set(const index_entry_t& entry, const T& value) {
    using Tr = internal::StorageTraits<T>;
    if (entry.type != Tr::kType || entry.storage_dim != Tr::kStorage)
      throw std::runtime_error("Trying to set index entry of type " + std::to_string(static_cast<int>(entry.type)) +
                               " with a value of another type");
    SYM_ASSERT(entry.offset >= 0 && static_cast<size_t>(entry.offset + entry.storage_dim) <= data_.size());
    std::copy(Tr::Ptr(value), Tr::Ptr(value) + Tr::kStorage, data_.begin() + entry.offset);
  }
  // values.cc:160-178: Remove drops the key only; Cleanup compacts the data array and returns the scalars freed
  bool Remove(const Key& key) { return map_.erase(key) > 0; }
  void RemoveAll() {
    map_.clear();
    data_.clear();
  }
  size_t Cleanup() {
    const std::vector<Scalar> data_copy = data_;
    const index_t full_index = CreateIndex(/* sort_by_offset = */ true);
    data_.resize(full_index.storage_dim);
    SYM_ASSERT(data_copy.size() >= data_.size());
    size_t new_offset = 0;
    for (const index_entry_t& entry : full_index.entries) {
      std::copy_n(data_copy.begin() + entry.offset, entry.storage_dim, data_.begin() + new_offset);
      map_[entry.key].offset = static_cast<int32_t>(new_offset);
      new_offset += entry.storage_dim;
    }
    return data_copy.size() - data_.size();
  }
  // values.cc:262-287: copy the indexed entries from another Values (same layout / separately indexed layouts)
  void Update(const index_t& index, const Values& other) {
    SYM_ASSERT(data_.size() == other.data_.size());
    for (const index_entry_t& entry : index.entries)
      std::copy_n(other.data_.begin() + entry.offset, entry.storage_dim, data_.begin() + entry.offset);
  }
  void Update(const index_t& index_this, const index_t& index_other, const Values& other) {
    SYM_ASSERT(index_this.entries.size() == index_other.entries.size());
    for (size_t i = 0; i < index_this.entries.size(); ++i) {
      const index_entry_t& entry_this = index_this.entries[i];
      const index_entry_t& entry_other = index_other.entries[i];
      SYM_ASSERT(entry_this.storage_dim == entry_other.storage_dim);
      SYM_ASSERT(entry_this.key == entry_other.key);
      std::copy_n(other.data_.begin() + entry_other.offset, entry_this.storage_dim, data_.begin() + entry_this.offset);
    }
  }
  // values.cc:45-62: like Update for keys that exist here, appends the others
  void UpdateOrSet(const index_t& index, const Values& other) {
    for (const index_entry_t& entry_other : index.entries) {
      auto it = map_.find(entry_other.key);
      if (it == map_.end()) {
        index_entry_t e = entry_other;
        e.offset = static_cast<int32_t>(data_.size());
        map_[e.key] = e;
        data_.insert(data_.end(), other.data_.begin() + entry_other.offset,
                     other.data_.begin() + entry_other.offset + entry_other.storage_dim);
      } else {
        SYM_ASSERT(it->second.storage_dim == entry_other.storage_dim);
        std::copy_n(other.data_.begin() + entry_other.offset, entry_other.storage_dim, data_.begin() + it->second.offset);
      }
    }
  }
  // values.cc:315-327: x <- x (+) delta for the indexed keys, delta laid out in index order
  void Retract(const index_t& index, const Scalar* delta, const Scalar epsilon) {
    SYM_ASSERT(index.tangent_dim >= 0);
    size_t tangent_inx = 0;
    for (const index_entry_t& entry : index.entries) {
      Scalar* t_ptr = data_.data() + entry.offset;
      const Scalar* d = delta + tangent_inx;
      if (entry.type == type_t::ROT3) {
        const Rot3<Scalar> r = internal::StorageTraits<Rot3<Scalar>>::From(t_ptr).Retract(Vector3<Scalar>(d[0], d[1], d[2]), epsilon);
        std::copy(r.Data().data(), r.Data().data() + 4, t_ptr);
      } else if (entry.type == type_t::POSE3) {
        const Pose3<Scalar> p = internal::StorageTraits<Pose3<Scalar>>::From(t_ptr).Retract(
            Vector6<Scalar>(d[0], d[1], d[2], d[3], d[4], d[5]), epsilon);
        std::copy(p.Data().data(), p.Data().data() + 7, t_ptr);
      } else {
        for (int32_t i = 0; i < entry.tangent_dim; ++i) t_ptr[i] += d[i];
      }
      tangent_inx += entry.tangent_dim;
    }
  }
  // values.cc:329-350: tangent vector from this to `others` for the indexed keys
  std::vector<Scalar> LocalCoordinates(const Values& others, const index_t& index, const Scalar epsilon) const {
    SYM_ASSERT(index.tangent_dim >= 0);
    std::vector<Scalar> tangent_vec(index.tangent_dim);
    size_t tangent_inx = 0;
    for (const index_entry_t& entry : index.entries) {
      const Scalar* a = data_.data() + entry.offset;
      const Scalar* b = others.data_.data() + entry.offset;
      Scalar* out = tangent_vec.data() + tangent_inx;
      if (entry.type == type_t::ROT3) {
        const Vector3<Scalar> v = internal::StorageTraits<Rot3<Scalar>>::From(a).LocalCoordinates(
            internal::StorageTraits<Rot3<Scalar>>::From(b), epsilon);
        for (int i = 0; i < 3; ++i) out[i] = v[i];
      } else if (entry.type == type_t::POSE3) {
        const Vector6<Scalar> v = internal::StorageTraits<Pose3<Scalar>>::From(a).LocalCoordinates(
            internal::StorageTraits<Pose3<Scalar>>::From(b), epsilon);
        for (int i = 0; i < 6; ++i) out[i] = v[i];
      } else {
        for (int32_t i = 0; i < entry.tangent_dim; ++i) out[i] = b[i] - a[i];
      }
      tangent_inx += entry.tangent_dim;
    }
    return tangent_vec;
  }

 private:
  MapType map_;
  std::vector<Scalar> data_;
};
using Valuesd = Values<double>;

// ------------------------------------------------------------------------------------------------
// Params / stats (lcmtypes/symforce.lcm)
// ------------------------------------------------------------------------------------------------
enum class lambda_update_type_t : int32_t { INVALID = 0, STATIC = 1, DYNAMIC = 2 };
enum class optimization_status_t : int32_t { INVALID = 0, SUCCESS = 1, HIT_ITERATION_LIMIT = 2, FAILED = 3 };
enum class levenberg_marquardt_solver_failure_reason_t : int32_t {
  INVALID = 0,
  LAMBDA_OUT_OF_BOUNDS = 1,
  INITIAL_ERROR_NOT_FINITE = 2
};

struct optimizer_params_t {
  bool verbose{false};
  bool debug_stats{false};
  bool check_derivatives{false};
  bool include_jacobians{false};
  bool debug_checks{false};
  double initial_lambda{1.0};
  double lambda_lower_bound{0.0};
  double lambda_upper_bound{1000000.0};
  lambda_update_type_t lambda_update_type{lambda_update_type_t::STATIC};
  double lambda_up_factor{4.0};
  double lambda_down_factor{0.25};
  double dynamic_lambda_update_beta{2.0};
  double dynamic_lambda_update_gamma{3.0};
  int32_t dynamic_lambda_update_p{3};
  bool use_diagonal_damping{false};
  bool use_unit_damping{true};
  bool keep_max_diagonal_damping{false};
  double diagonal_damping_min{1e-6};
  int32_t iterations{50};
  double early_exit_min_reduction{1e-6};
  double early_exit_min_absolute_error{0.0};
  bool enable_bold_updates{false};
};
inline optimizer_params_t DefaultOptimizerParams() { return optimizer_params_t{}; }  // optimizer.cc:8-56

struct optimization_iteration_t {
  int16_t iteration{0};
  double current_lambda{0};
  double new_error_linear{0};
  double new_error{0};
  double relative_reduction{0};
  bool update_accepted{false};
  double update_angle_change{0};
  // debug_stats only (symforce.lcm:247-257; levenberg_marquardt_solver.tcc:115-122, 165-176): the update this iteration
  // tried (empty in the record of iteration -1), the data of the Values buffer, the residual, and -- with
  // include_jacobians -- the values of the Jacobian in the order of OptimizationStats::jacobian_sparsity
  std::vector<double> update, values, residual, jacobian_values;
};

// lcmtypes/symforce.lcm:268-277
struct sparse_matrix_structure_t {
  std::vector<int32_t> row_indices;      // size nnz
  std::vector<int32_t> column_pointers;  // size cols + 1 (linearization.h:94-100)
  std::vector<int64_t> shape;            // (rows, cols); empty when not filled
};

// Linearization container (symforce/opt/linearization.h:27-77) with the CSC pieces the reference's
// Eigen::SparseMatrix exposes (valuePtr / innerIndexPtr / outerIndexPtr / nonZeros / rows / cols)
struct SparseMatrixCsc {
  int rows_{0}, cols_{0};
  std::vector<int32_t> outer, inner;
  std::vector<double> values;
  int rows() const { return rows_; }
  int cols() const { return cols_; }
  int64_t nonZeros() const { return static_cast<int64_t>(values.size()); }
  const double* valuePtr() const { return values.data(); }
  const int32_t* innerIndexPtr() const { return inner.data(); }
  const int32_t* outerIndexPtr() const { return outer.data(); }
};
struct SparseLinearization {
  std::vector<double> residual;
  SparseMatrixCsc hessian_lower;
  SparseMatrixCsc jacobian;  // M x N, filled by Optimizer::Linearize when optimizer_params_t::include_jacobians is set
  std::vector<double> rhs;
  bool IsInitialized() const { return !rhs.empty(); }
  double Error() const {
    double s = 0;
    for (double r : residual) s += r * r;
    return 0.5 * s;
  }
  // linearization.h:62-67: change in error the linear model predicts for an update solved with the given damping
  double LinearDeltaError(const std::vector<double>& x_update, const std::vector<double>& damping_vector) const {
    SYM_ASSERT(x_update.size() == rhs.size() && damping_vector.size() == rhs.size());
    double s = 0;
    for (size_t i = 0; i < rhs.size(); ++i) s += x_update[i] * (rhs[i] - damping_vector[i] * x_update[i]);
    return 0.5 * s;
  }
};

// Dynamic dense matrix (column-major) for covariance blocks: Eigen::MatrixXd when real Eigen is used
#if defined(SFX_USE_EIGEN)
template <typename Scalar>
using MatrixX = Eigen::Matrix<Scalar, Eigen::Dynamic, Eigen::Dynamic>;
using ComputationInfo = Eigen::ComputationInfo;
constexpr ComputationInfo kSuccess = Eigen::Success;
constexpr ComputationInfo kNumericalIssue = Eigen::NumericalIssue;
#else
template <typename Scalar>
class MatrixX {
 public:
  MatrixX() = default;
  MatrixX(int r, int c) : r_(r), c_(c), d_(static_cast<size_t>(r) * c, Scalar(0)) {}
  void resize(int r, int c) {
    r_ = r;
    c_ = c;
    d_.assign(static_cast<size_t>(r) * c, Scalar(0));
  }
  int rows() const { return r_; }
  int cols() const { return c_; }
  Scalar& operator()(int r, int c) { return d_[r + static_cast<size_t>(c) * r_]; }
  const Scalar& operator()(int r, int c) const { return d_[r + static_cast<size_t>(c) * r_]; }
  Scalar* data() { return d_.data(); }
  const Scalar* data() const { return d_.data(); }
  // covariance_block.block(offset, offset, dim, dim) of SplitCovariancesByKey (covariance_utils.h:180-181)
  MatrixX block(int r0, int c0, int nr, int nc) const {
    MatrixX m(nr, nc);
    for (int c = 0; c < nc; ++c)
      for (int r = 0; r < nr; ++r) m(r, c) = (*this)(r0 + r, c0 + c);
    return m;
  }

 private:
  int r_{0}, c_{0};
  std::vector<Scalar> d_;
};
enum ComputationInfo { kSuccess = 0, kNumericalIssue = 1, kNoConvergence = 2, kInvalidInput = 3 };  // Eigen::ComputationInfo
#endif

struct OptimizationStats {
  std::vector<optimization_iteration_t> iterations;
  int32_t best_index{0};
  optimization_status_t status{optimization_status_t::INVALID};
  int32_t failure_reason{0};
  std::optional<SparseLinearization> best_linearization{};
  // debug_stats only (optimization_stats.h:40-60).  jacobian_sparsity additionally needs include_jacobians.
  // linear_solver_ordering: elimination position -> scalar index of the system the linear solver factors (the reduced
  // camera system under Schur elimination).  cholesky_factor_sparsity stays default constructed, as the reference
  // leaves it for a solver that does not expose L(): the factor here is a supernodal LL^T held as dense fronts.
  sparse_matrix_structure_t jacobian_sparsity{};
  std::vector<int32_t> linear_solver_ordering{};
  sparse_matrix_structure_t cholesky_factor_sparsity{};

  // optimization_stats.h:67-75: the Jacobian of an iteration record (a copy; the reference returns a Map)
  SparseMatrixCsc JacobianView(const optimization_iteration_t& iteration) const {
    SYM_ASSERT(jacobian_sparsity.shape.size() == 2 &&
               "Jacobian sparsity is empty, did you set debug_stats = true and include_jacobians = true?");
    SYM_ASSERT(jacobian_sparsity.row_indices.size() == iteration.jacobian_values.size());
    SparseMatrixCsc J;
    J.rows_ = static_cast<int>(jacobian_sparsity.shape[0]);
    J.cols_ = static_cast<int>(jacobian_sparsity.shape[1]);
    J.outer = jacobian_sparsity.column_pointers;
    J.inner = jacobian_sparsity.row_indices;
    J.values = iteration.jacobian_values;
    return J;
  }
};

// ------------------------------------------------------------------------------------------------
// Generated factor functions with a device implementation.  Host-side they are only *identities*:
// Factor::Hessian recognises them by address.  When the reference's generated headers are also
// included (real Eigen build) these are the same entities (matching redeclarations); otherwise the
// definitions below stand in and refuse to run on the host.
// ------------------------------------------------------------------------------------------------
namespace internal {
[[noreturn]] inline void DeviceOnly(const char* name) {
  throw std::runtime_error(std::string(name) + " is evaluated on the GPU by sym::Optimizer; there is no host implementation");
}
}  // namespace internal

#ifndef SFX_HAVE_REFERENCE_FACTOR_HEADERS
template <typename Scalar>
void SnavelyReprojectionFactor(const Pose3<Scalar>&, const Eigen::Matrix<Scalar, 3, 1>&, const Eigen::Matrix<Scalar, 3, 1>&,
                               const Eigen::Matrix<Scalar, 2, 1>&, const Scalar, Eigen::Matrix<Scalar, 2, 1>* const = nullptr,
                               Eigen::Matrix<Scalar, 2, 12>* const = nullptr, Eigen::Matrix<Scalar, 12, 12>* const = nullptr,
                               Eigen::Matrix<Scalar, 12, 1>* const = nullptr) {
  internal::DeviceOnly("SnavelyReprojectionFactor");
}
template <typename Scalar>
void BetweenFactorPose3(const Pose3<Scalar>&, const Pose3<Scalar>&, const Pose3<Scalar>&, const Eigen::Matrix<Scalar, 6, 6>&,
                        const Scalar, Eigen::Matrix<Scalar, 6, 1>* const = nullptr,
                        Eigen::Matrix<Scalar, 6, 12>* const = nullptr, Eigen::Matrix<Scalar, 12, 12>* const = nullptr,
                        Eigen::Matrix<Scalar, 12, 1>* const = nullptr) {
  internal::DeviceOnly("BetweenFactorPose3");
}
template <typename Scalar>
void PriorFactorPose3(const Pose3<Scalar>&, const Pose3<Scalar>&, const Eigen::Matrix<Scalar, 6, 6>&, const Scalar,
                      Eigen::Matrix<Scalar, 6, 1>* const = nullptr, Eigen::Matrix<Scalar, 6, 6>* const = nullptr,
                      Eigen::Matrix<Scalar, 6, 6>* const = nullptr, Eigen::Matrix<Scalar, 6, 1>* const = nullptr) {
  internal::DeviceOnly("PriorFactorPose3");
}
template <typename Scalar>
void MatchingFactor(const Pose3<Scalar>&, const Eigen::Matrix<Scalar, 3, 1>&, const Eigen::Matrix<Scalar, 3, 1>&,
                    const Scalar, Eigen::Matrix<Scalar, 3, 1>* const = nullptr, Eigen::Matrix<Scalar, 3, 6>* const = nullptr,
                    Eigen::Matrix<Scalar, 6, 6>* const = nullptr, Eigen::Matrix<Scalar, 6, 1>* const = nullptr) {
  internal::DeviceOnly("MatchingFactor");
}
template <typename Scalar>
void OdometryFactor(const Pose3<Scalar>&, const Pose3<Scalar>&, const Pose3<Scalar>&, const Eigen::Matrix<Scalar, 6, 1>&,
                    const Scalar, Eigen::Matrix<Scalar, 6, 1>* const = nullptr, Eigen::Matrix<Scalar, 6, 12>* const = nullptr,
                    Eigen::Matrix<Scalar, 12, 12>* const = nullptr, Eigen::Matrix<Scalar, 12, 1>* const = nullptr) {
  internal::DeviceOnly("OdometryFactor");
}
template <typename Scalar>
void InverseRangeLandmarkLinearGncFactor(const Pose3<Scalar>&, const LinearCameraCal<Scalar>&, const Pose3<Scalar>&,
                                         const LinearCameraCal<Scalar>&, const Scalar, const Eigen::Matrix<Scalar, 2, 1>&,
                                         const Eigen::Matrix<Scalar, 2, 1>&, const Scalar, const Scalar, const Scalar,
                                         const Scalar, Eigen::Matrix<Scalar, 2, 1>* const = nullptr,
                                         Eigen::Matrix<Scalar, 2, 13>* const = nullptr,
                                         Eigen::Matrix<Scalar, 13, 13>* const = nullptr,
                                         Eigen::Matrix<Scalar, 13, 1>* const = nullptr) {
  internal::DeviceOnly("InverseRangeLandmarkLinearGncFactor");
}
template <typename Scalar>
void InverseRangeLandmarkPriorFactor(const Scalar, const Scalar, const Scalar, const Scalar, const Scalar,
                                     Eigen::Matrix<Scalar, 1, 1>* const = nullptr, Eigen::Matrix<Scalar, 1, 1>* const = nullptr,
                                     Eigen::Matrix<Scalar, 1, 1>* const = nullptr, Eigen::Matrix<Scalar, 1, 1>* const = nullptr) {
  internal::DeviceOnly("InverseRangeLandmarkPriorFactor");
}
template <typename Scalar>
void BetweenFactorRot3(const Rot3<Scalar>&, const Rot3<Scalar>&, const Rot3<Scalar>&, const Eigen::Matrix<Scalar, 3, 3>&,
                       const Scalar, Eigen::Matrix<Scalar, 3, 1>* const = nullptr, Eigen::Matrix<Scalar, 3, 6>* const = nullptr,
                       Eigen::Matrix<Scalar, 6, 6>* const = nullptr, Eigen::Matrix<Scalar, 6, 1>* const = nullptr) {
  internal::DeviceOnly("BetweenFactorRot3");
}
template <typename Scalar>
void PriorFactorRot3(const Rot3<Scalar>&, const Rot3<Scalar>&, const Eigen::Matrix<Scalar, 3, 3>&, const Scalar,
                     Eigen::Matrix<Scalar, 3, 1>* const = nullptr, Eigen::Matrix<Scalar, 3, 3>* const = nullptr,
                     Eigen::Matrix<Scalar, 3, 3>* const = nullptr, Eigen::Matrix<Scalar, 3, 1>* const = nullptr) {
  internal::DeviceOnly("PriorFactorRot3");
}
#endif
}  // namespace sym

// gnc_factors::BarronFactor (test/symforce_function_codegen_test_data/symengine/gnc_test_data/cpp/symforce/gnc_factors/
// barron_factor.h:33-40): the factor of the reference's GNC test, a device kind like the functions above
namespace gnc_factors {
template <typename Scalar>
void BarronFactor(const Eigen::Matrix<Scalar, 5, 1>&, const Eigen::Matrix<Scalar, 5, 1>&, const Scalar, const Scalar,
                  Eigen::Matrix<Scalar, 5, 1>* const = nullptr, Eigen::Matrix<Scalar, 5, 5>* const = nullptr,
                  Eigen::Matrix<Scalar, 5, 5>* const = nullptr, Eigen::Matrix<Scalar, 5, 1>* const = nullptr) {
  sym::internal::DeviceOnly("BarronFactor");
}
}  // namespace gnc_factors

namespace sym {

// ------------------------------------------------------------------------------------------------
// Factor (symforce/opt/factor.h)
// ------------------------------------------------------------------------------------------------
namespace internal {
struct KindInfo {
  int kind, n_args, n_opt;
  int opt_args[3];
};
inline const std::unordered_map<const void*, KindInfo>& KindRegistry() {
  static const std::unordered_map<const void*, KindInfo> reg = [] {
    std::unordered_map<const void*, KindInfo> r;
    auto add = [&](const void* f, KindInfo k) { r[f] = k; };
    add(reinterpret_cast<const void*>(&SnavelyReprojectionFactor<double>), {SFX_KIND_SNAVELY, 5, 3, {0, 1, 2}});
    add(reinterpret_cast<const void*>(&BetweenFactorPose3<double>), {SFX_KIND_BETWEEN_POSE3, 5, 2, {0, 1, -1}});
    add(reinterpret_cast<const void*>(&PriorFactorPose3<double>), {SFX_KIND_PRIOR_POSE3, 4, 1, {0, -1, -1}});
    add(reinterpret_cast<const void*>(&MatchingFactor<double>), {SFX_KIND_MATCHING, 4, 1, {0, -1, -1}});
    add(reinterpret_cast<const void*>(&OdometryFactor<double>), {SFX_KIND_ODOMETRY, 5, 2, {0, 1, -1}});
    add(reinterpret_cast<const void*>(&InverseRangeLandmarkLinearGncFactor<double>),
        {SFX_KIND_IRL_LINEAR_GNC, 11, 3, {0, 2, 4}});
    add(reinterpret_cast<const void*>(&InverseRangeLandmarkPriorFactor<double>), {SFX_KIND_IRL_PRIOR, 5, 1, {0, -1, -1}});
    add(reinterpret_cast<const void*>(&BetweenFactorRot3<double>), {SFX_KIND_BETWEEN_ROT3, 5, 2, {0, 1, -1}});
    add(reinterpret_cast<const void*>(&PriorFactorRot3<double>), {SFX_KIND_PRIOR_ROT3, 4, 1, {0, -1, -1}});
    add(reinterpret_cast<const void*>(&gnc_factors::BarronFactor<double>), {SFX_KIND_BARRON, 4, 1, {0, -1, -1}});
    return r;
  }();
  return reg;
}
}  // namespace internal

template <typename Scalar>
class Factor {
 public:
  // Factor::Hessian(func, keys_to_func, keys_to_optimize) (factor.h:231-234).  `func` must be one
  // of the generated functions with a device implementation (recognised by address).
  template <typename Functor>
  static Factor Hessian(Functor&& func, const std::vector<Key>& keys_to_func,
                        const std::vector<Key>& keys_to_optimize = {}, bool /*requires_jacobian*/ = false) {
    const void* addr = AddressOf(std::forward<Functor>(func));
    const auto& reg = internal::KindRegistry();
    auto it = addr ? reg.find(addr) : reg.end();
    if (it == reg.end())
      throw std::runtime_error(
          "sym::Factor: this functor has no GPU implementation (only generated factor functions whose inputs are all "
          "Values keys run on the device; there is no CPU fallback)");
    Factor f;
    f.kind_ = it->second.kind;
    if (static_cast<int>(keys_to_func.size()) != it->second.n_args)
      throw std::runtime_error("SYM_ASSERT: keys_to_func.size() == number of function arguments");
    f.keys_ = keys_to_func;
    // keys_to_optimize empty == all keys are optimized (factor.h:190-192); for generated
    // linearization functions that means the arguments the derivative is taken with respect to
    if (keys_to_optimize.empty()) {
      for (int o = 0; o < it->second.n_opt; ++o) f.keys_to_optimize_.push_back(keys_to_func[it->second.opt_args[o]]);
    } else {
      f.keys_to_optimize_ = keys_to_optimize;
    }
    if (static_cast<int>(f.keys_to_optimize_.size()) != it->second.n_opt)
      throw std::runtime_error("SYM_ASSERT: keys_to_optimize must list the linearized arguments of the function");
    for (int o = 0; o < it->second.n_opt; ++o)
      if (!(f.keys_to_optimize_[o] == keys_to_func[it->second.opt_args[o]]))
        throw std::runtime_error("SYM_ASSERT: keys_to_optimize must be the linearized arguments, in argument order");
    return f;
  }
  // Factor::Jacobian(func, keys_to_func, keys_to_optimize) (factor.h:195-196): the reference evaluates residual and
  // Jacobian with `func` and forms H = J^T J, rhs = J^T r itself (factor.tcc:24-81).  The device kinds do exactly
  // that -- their kernels compute (residual, J) and accumulate J^T J / J^T r -- so a generated function with a device
  // implementation lowers to the same kind whether it is handed to Hessian() or to Jacobian().  Host functors
  // (lambdas, std::function) have no device implementation and are rejected like in Hessian(): no CPU fallback.
  template <typename Functor>
  static Factor Jacobian(Functor&& func, const std::vector<Key>& keys_to_func,
                         const std::vector<Key>& keys_to_optimize = {}) {
    return Hessian(std::forward<Functor>(func), keys_to_func, keys_to_optimize, /*requires_jacobian=*/true);
  }
  const std::vector<Key>& AllKeys() const { return keys_; }
  const std::vector<Key>& OptimizedKeys() const { return keys_to_optimize_; }
  int Kind() const { return kind_; }

 private:
  // functions (and pointers to functions) decay to a function pointer; anything else has no address
  template <typename F>
  static const void* AddressOf(F&& f) {
    using D = typename std::decay<F>::type;
    if constexpr (std::is_pointer<D>::value && std::is_function<typename std::remove_pointer<D>::type>::value)
      return reinterpret_cast<const void*>(static_cast<D>(f));
    else
      return nullptr;
  }
  int kind_{-1};
  std::vector<Key> keys_, keys_to_optimize_;
};
using Factord = Factor<double>;

// ComputeKeysToOptimize (factor.h:424-449): union of optimized keys, LexicalLessThan order
template <typename Scalar>
std::vector<Key> ComputeKeysToOptimize(const std::vector<Factor<Scalar>>& factors) {
  std::vector<Key> keys;
  std::unordered_set<Key, KeyHash> seen;
  seen.reserve(factors.size());
  for (const auto& f : factors)
    for (const Key& k : f.OptimizedKeys())
      if (seen.insert(k).second) keys.push_back(k);
  std::sort(keys.begin(), keys.end(), Key::LexicalLessThan);
  return keys;
}

// Extension point for the GPU path: which linear solver the LM loop uses.
struct GpuSolverOptions {
  enum Solver { AUTO, CHOLESKY, SCHUR };
  Solver solver = AUTO;        // AUTO: Schur when a trailing run of <=3-dim vector keys is pairwise
                               // unconnected (the block-diagonal C of sparse_schur_solver.h:20-31)
  int schur_num_keys = 0;      // SCHUR: number of trailing keys to eliminate
  int ordering = SFX_ORDERING_METIS_SCALAR;
  int device = 0;
};

// ------------------------------------------------------------------------------------------------
// Optimizer (symforce/opt/optimizer.h:72-325)
// ------------------------------------------------------------------------------------------------
template <typename ScalarType>
class Optimizer {
 public:
  using Scalar = ScalarType;
  using Stats = OptimizationStats;
  static_assert(std::is_same<Scalar, double>::value, "the GPU path computes in fp64");

  Optimizer(const optimizer_params_t& params, std::vector<Factor<Scalar>> factors,
            const std::string& name = "sym::Optimize", std::vector<Key> keys = {},
            const Scalar epsilon = kDefaultEpsilon<Scalar>, const GpuSolverOptions& gpu = GpuSolverOptions())
      : params_(params), factors_(std::move(factors)), name_(name), keys_(std::move(keys)), epsilon_(epsilon), gpu_(gpu) {
    if (keys_.empty()) keys_ = ComputeKeysToOptimize(factors_);
    SYM_ASSERT(!factors_.empty());
    SYM_ASSERT(!keys_.empty());
    SYM_ASSERT(!params.check_derivatives || params.include_jacobians);  // optimizer.tcc:39, 62
  }
  Optimizer(const Optimizer&) = delete;
  Optimizer& operator=(const Optimizer&) = delete;
  virtual ~Optimizer() {
    if (handle_) sfx_problem_destroy(handle_);
    if (cov_handle_) sfx_problem_destroy(cov_handle_);
    if (check_handle_) sfx_problem_destroy(check_handle_);
  }

  // Optimize(values, num_iterations, populate_best_linearization) (optimizer.tcc:79-89)
  Stats Optimize(Values<Scalar>& values, int num_iterations = -1, bool populate_best_linearization = false) {
    Stats stats;
    Optimize(values, num_iterations, populate_best_linearization, stats);
    return stats;
  }
  virtual void Optimize(Values<Scalar>& values, int num_iterations, bool populate_best_linearization, Stats& stats) {
    OptimizeImpl(values, num_iterations, populate_best_linearization, stats, /*continue_previous=*/false);
  }
  // the two shorter forms of optimizer.h:159, 172
  void Optimize(Values<Scalar>& values, int num_iterations, Stats& stats) { Optimize(values, num_iterations, false, stats); }
  void Optimize(Values<Scalar>& values, Stats& stats) { Optimize(values, -1, false, stats); }
  const std::string& GetName() const { return name_; }

 protected:
  // OptimizeImpl (internal/optimizer_utils.h:78-101), or -- continue_previous -- GncOptimizer::OptimizeContinue
  // (gnc_optimizer.h:133-142): ResetState(values) + IterateToConvergence, stats keep accumulating
  void OptimizeImpl(Values<Scalar>& values, int num_iterations, bool populate_best_linearization, Stats& stats,
                    bool continue_previous) {
    Initialize(values);
    // check_derivatives (optimizer.tcc:261-272): the reference asserts inside its linearize function, i.e. at the values
    // of every linearization; here the same assertion runs on a sibling problem (the LM state is not disturbed) at the
    // initial values before the run and, after it, at the values of every record (debug_stats) or at the best values
    if (params_.check_derivatives) CheckDerivativesAt(values.Data().data(), values.Data().size());
    Check(sfx_set_values(handle_, values.Data().data(), static_cast<int64_t>(values.Data().size())));
    sfx_stats st{};
    Check(continue_previous ? sfx_optimize_continue(handle_, num_iterations, &st) : sfx_optimize(handle_, num_iterations, &st));
    std::vector<sfx_iteration> its(st.n_iterations);
    int32_t n = 0;
    Check(sfx_get_iterations(handle_, its.data(), st.n_iterations, &n));
    stats.iterations.clear();
    for (const auto& it : its) {
      optimization_iteration_t o;
      o.iteration = static_cast<int16_t>(it.iteration);
      o.current_lambda = it.current_lambda;
      o.new_error_linear = it.new_error_linear;
      o.new_error = it.new_error;
      o.relative_reduction = it.relative_reduction;
      o.update_accepted = it.update_accepted != 0;
      o.update_angle_change = it.update_angle_change;
      stats.iterations.push_back(o);
    }
    stats.jacobian_sparsity = {};
    stats.linear_solver_ordering.clear();
    stats.cholesky_factor_sparsity = {};
    if (params_.debug_stats) {
      int32_t N = 0, M = 0;
      int64_t nnz = 0, jnnz = 0;
      Check(sfx_get_dims(handle_, &N, &M, &nnz));
      if (params_.include_jacobians) {  // levenberg_marquardt_solver.tcc:172-175
        Check(sfx_get_jacobian_pattern(handle_, &jnnz, nullptr, nullptr));
        stats.jacobian_sparsity.row_indices.resize(jnnz);
        stats.jacobian_sparsity.column_pointers.resize(N + 1);
        stats.jacobian_sparsity.shape = {M, N};
        Check(sfx_get_jacobian_pattern(handle_, nullptr, stats.jacobian_sparsity.column_pointers.data(),
                                       stats.jacobian_sparsity.row_indices.data()));
      }
      for (size_t i = 0; i < stats.iterations.size(); ++i) {
        auto& rec = stats.iterations[i];
        rec.values.resize(values.Data().size());
        rec.residual.resize(M);
        Check(sfx_get_iteration_debug(handle_, static_cast<int32_t>(i), rec.values.data(), rec.residual.data()));
        if (rec.iteration >= 0) {
          rec.update.resize(N);
          Check(sfx_get_iteration_update(handle_, static_cast<int32_t>(i), rec.update.data()));
        }
        // (the reference fills the jacobian of records >= 0 whenever the linearization carries one, tcc:120-121)
        if (params_.include_jacobians) {
          rec.jacobian_values.resize(jnnz);
          Check(sfx_get_iteration_jacobian(handle_, static_cast<int32_t>(i), rec.jacobian_values.data()));
        }
      }
      int32_t n_ord = 0;  // levenberg_marquardt_solver.tcc:209-212
      Check(sfx_get_ordering(handle_, nullptr, 0, &n_ord));
      stats.linear_solver_ordering.resize(n_ord);
      Check(sfx_get_ordering(handle_, stats.linear_solver_ordering.data(), n_ord, &n_ord));
    }
    stats.best_index = st.best_index;
    stats.status = static_cast<optimization_status_t>(st.status);
    stats.failure_reason = st.failure_reason;
    // values = nonlinear_solver.GetBestValues() (internal/optimizer_utils.h:69)
    // (only the optimized keys changed: Values::Update semantics move a fraction of the buffer over PCIe)
    Check(sfx_update_best_values(handle_, values.DataPointer(), static_cast<int64_t>(values.Data().size()), nullptr));
    if (params_.check_derivatives) {
      if (params_.debug_stats)
        for (size_t i = 1; i < stats.iterations.size(); ++i)
          CheckDerivativesAt(stats.iterations[i].values.data(), stats.iterations[i].values.size());
      else
        CheckDerivativesAt(values.Data().data(), values.Data().size());
    }
    if (populate_best_linearization) {
      SparseLinearization lin;
      FillPattern(lin);
      Check(sfx_get_best_linearization(handle_, lin.residual.data(), lin.rhs.data(), lin.hessian_lower.values.data()));
      stats.best_linearization = std::move(lin);
    } else {
      stats.best_linearization.reset();
    }
  }
  void RelaxDampingToInitial() { Check(sfx_relax_damping_to_initial(handle_)); }

 public:

  // Linearize(values) (optimizer.h:177)
  SparseLinearization Linearize(const Values<Scalar>& values) {
    Initialize(values);
    Check(sfx_set_values(handle_, values.Data().data(), static_cast<int64_t>(values.Data().size())));
    SparseLinearization lin;
    FillPattern(lin);
    Check(sfx_linearize(handle_, lin.residual.data(), lin.rhs.data(), lin.hessian_lower.values.data()));
    if (params_.include_jacobians) {  // linearizer.cc:252-259, 297-313
      int64_t nnz = 0;
      Check(sfx_get_jacobian_pattern(handle_, &nnz, nullptr, nullptr));
      lin.jacobian.rows_ = static_cast<int>(lin.residual.size());
      lin.jacobian.cols_ = static_cast<int>(lin.rhs.size());
      lin.jacobian.outer.resize(lin.rhs.size() + 1);
      lin.jacobian.inner.resize(nnz);
      lin.jacobian.values.resize(nnz);
      Check(sfx_get_jacobian_pattern(handle_, nullptr, lin.jacobian.outer.data(), lin.jacobian.inner.data()));
      Check(sfx_linearize_jacobian(handle_, lin.jacobian.values.data()));
    }
    if (params_.check_derivatives) CheckDerivativesAt(values.Data().data(), values.Data().size());
    return lin;
  }

  // internal::CheckDerivatives (internal/derivative_checker.h:32-123) at the given Values data, SYM_ASSERTed as the
  // reference's linearize function does (optimizer.tcc:266-268); rel_errors (jacobian, hessian, rhs) for callers that
  // want the numbers.  Runs on a sibling problem with the same structure.
  bool CheckDerivatives(const Values<Scalar>& values, std::array<double, 3>* rel_errors = nullptr) {
    Initialize(values);
    return CheckDerivativesImpl(values.Data().data(), values.Data().size(), rel_errors);
  }

  // ComputeAllCovariances (optimizer.tcc:113-121): (H + eps I)^-1 split by key
  void ComputeAllCovariances(const SparseLinearization& linearization,
                             std::unordered_map<Key, MatrixX<Scalar>, KeyHash>& covariances_by_key) {
    MatrixX<Scalar> cov;
    ComputeFullCovariance(linearization, cov);
    SplitCovariancesByKey(cov, keys_, covariances_by_key);
  }

  // ComputeCovariances (optimizer.tcc:177-199): marginal covariance of `keys`, which must be the first
  // keys of Keys() in order; every later key is eliminated with the Schur complement.
  // c_is_block_diagonal = true: each eliminated key is a vector of dim <= 3 that shares no factor with another
  // eliminated key (SparseSchurSolver on the device).  false: C may have any structure
  // (covariance_utils.h:41-103); the block comes from the sparse Cholesky factor of the whole matrix, damped on C.
  ComputationInfo ComputeCovariances(const SparseLinearization& linearization, const std::vector<Key>& keys,
                                     std::unordered_map<Key, MatrixX<Scalar>, KeyHash>& covariances_by_key,
                                     const bool c_is_block_diagonal = true) {
    SYM_ASSERT(IsInitialized());
    SYM_ASSERT(!keys.empty() && keys.size() <= keys_.size());
    for (size_t i = 0; i < keys.size(); ++i) SYM_ASSERT(keys[i] == keys_[i]);  // CheckKeyOrderMatchesLinearizerKeysStart
    if (keys.size() == keys_.size()) {
      ComputeAllCovariances(linearization, covariances_by_key);
      return kSuccess;
    }
    int block_dim = 0;
    for (size_t i = 0; i < keys.size(); ++i) block_dim += kentries_[i].tangent_dim;
    const int n_elim = c_is_block_diagonal ? static_cast<int>(keys_.size() - keys.size()) : 0;
    sfx_problem* h = handle_;
    if (!c_is_block_diagonal) {
      if (solver_ != SFX_SOLVER_CHOLESKY) {  // sibling problem that factors the whole H, as for ComputeFullCovariance
        if (cov_handle_ && cov_schur_keys_ != 0) {
          sfx_problem_destroy(cov_handle_);
          cov_handle_ = nullptr;
        }
        if (!cov_handle_) {
          cov_handle_ = Create(SFX_SOLVER_CHOLESKY, 0);
          cov_schur_keys_ = 0;
        }
        h = cov_handle_;
      }
    } else if (!(solver_ == SFX_SOLVER_SCHUR && schur_keys_ == n_elim)) {
      // the LM solve uses another split (or none): a sibling problem with the same structure and the
      // requested Schur split, built once (the reference builds a SparseSchurSolver per call, covariance_utils.h:141-145)
      if (cov_handle_ && cov_schur_keys_ != n_elim) {
        sfx_problem_destroy(cov_handle_);
        cov_handle_ = nullptr;
      }
      if (!cov_handle_) {
        cov_handle_ = Create(SFX_SOLVER_SCHUR, n_elim);
        cov_schur_keys_ = n_elim;
      }
      h = cov_handle_;
    }
    MatrixX<Scalar> cov(block_dim, block_dim);
    const sfx_status st = sfx_compute_covariance(h, linearization.hessian_lower.values.data(), block_dim, cov.data());
    if (st == SFX_ERR_NUMERICAL) return kNumericalIssue;
    if (st != SFX_OK) throw std::runtime_error(std::string("sym::Optimizer<") + name_ + ">: " + sfx_last_error(h));
    SplitCovariancesByKey(cov, keys, covariances_by_key);
    return kSuccess;
  }

  // ComputeFullCovariance (optimizer.tcc:201-206 -> LevenbergMarquardtSolver::ComputeCovariance)
  void ComputeFullCovariance(const SparseLinearization& linearization, MatrixX<Scalar>& covariance) {
    SYM_ASSERT(IsInitialized());
    int N = 0;
    for (const auto& e : kentries_) N += e.tangent_dim;
    sfx_problem* h = handle_;
    if (solver_ != SFX_SOLVER_CHOLESKY) {
      // the full inverse needs the factorization of the whole H: sibling problem without Schur elimination
      if (cov_handle_ && cov_schur_keys_ != 0) {
        sfx_problem_destroy(cov_handle_);
        cov_handle_ = nullptr;
      }
      if (!cov_handle_) {
        cov_handle_ = Create(SFX_SOLVER_CHOLESKY, 0);
        cov_schur_keys_ = 0;
      }
      h = cov_handle_;
    }
    covariance.resize(N, N);
    const sfx_status st = sfx_compute_covariance(h, linearization.hessian_lower.values.data(), N, covariance.data());
    if (st != SFX_OK) throw std::runtime_error(std::string("sym::Optimizer<") + name_ + ">: " + sfx_last_error(h));
  }

  bool IsInitialized() const { return handle_ != nullptr; }
  const std::vector<Key>& Keys() const { return keys_; }
  const std::vector<Factor<Scalar>>& Factors() const { return factors_; }
  const optimizer_params_t& Params() const { return params_; }
  void UpdateParams(const optimizer_params_t& params) {
    params_ = params;
    if (handle_) {
      sfx_params p = ToC(params_);
      Check(sfx_update_params(handle_, &p));
    }
  }
  sfx_problem* Handle() { return handle_; }

 private:
  static sfx_params ToC(const optimizer_params_t& q) {
    sfx_params p{};
    p.verbose = q.verbose;
    p.debug_stats = q.debug_stats;
    p.check_derivatives = q.check_derivatives;
    p.include_jacobians = q.include_jacobians;
    p.debug_checks = q.debug_checks;
    p.initial_lambda = q.initial_lambda;
    p.lambda_lower_bound = q.lambda_lower_bound;
    p.lambda_upper_bound = q.lambda_upper_bound;
    p.lambda_update_type = static_cast<int32_t>(q.lambda_update_type);
    p.lambda_up_factor = q.lambda_up_factor;
    p.lambda_down_factor = q.lambda_down_factor;
    p.dynamic_lambda_update_beta = q.dynamic_lambda_update_beta;
    p.dynamic_lambda_update_gamma = q.dynamic_lambda_update_gamma;
    p.dynamic_lambda_update_p = q.dynamic_lambda_update_p;
    p.use_diagonal_damping = q.use_diagonal_damping;
    p.use_unit_damping = q.use_unit_damping;
    p.keep_max_diagonal_damping = q.keep_max_diagonal_damping;
    p.diagonal_damping_min = q.diagonal_damping_min;
    p.iterations = q.iterations;
    p.early_exit_min_reduction = q.early_exit_min_reduction;
    p.early_exit_min_absolute_error = q.early_exit_min_absolute_error;
    p.enable_bold_updates = q.enable_bold_updates;
    return p;
  }
  void Check(sfx_status s) const { Check(s, handle_); }
  void Check(sfx_status s, sfx_problem* h) const {
    if (s != SFX_OK) throw std::runtime_error(std::string("sym::Optimizer<") + name_ + ">: " + sfx_last_error(h));
  }
  static int DeviceType(type_t t) {
    switch (t) {
      case type_t::ROT3: return SFX_TYPE_ROT3;
      case type_t::POSE3: return SFX_TYPE_POSE3;
      default: return SFX_TYPE_VECTOR;
    }
  }
  // Lazy first-call initialisation like Optimizer::Initialize (optimizer.tcc:273-278): index the
  // keys, lower every factor to argument offsets, create the device problem.
  void Initialize(const Values<Scalar>& values) {
    if (handle_) {
      SYM_ASSERT(static_cast<int64_t>(values.Data().size()) == n_values_);
      return;
    }
    const auto t_index = std::chrono::steady_clock::now();
    std::unordered_map<Key, int, KeyHash> key_index;
    kentries_.clear();
    for (size_t i = 0; i < keys_.size(); ++i) {
      const index_entry_t& e = values.IndexEntryAt(keys_[i]);
      kentries_.push_back(sfx_key_entry{DeviceType(e.type), e.offset, e.storage_dim, e.tangent_dim});
      key_index[keys_[i]] = static_cast<int>(i);
    }
    // Factors grouped by device kind (batches in ascending kind, slots in factor order).  Pass 1 (serial) numbers the
    // slots; pass 2 looks every key up -- ~8 hash lookups per factor in maps far larger than the caches -- on up to 8
    // host threads, writing straight into the flat [argument][slot] arrays the C ABI takes.
    const size_t nf = factors_.size();
    std::map<int, int> kind2batch;
    for (const auto& f : factors_) kind2batch.emplace(f.Kind(), 0);
    batch_kind_.clear();
    for (auto& kv : kind2batch) {
      kv.second = static_cast<int>(batch_kind_.size());
      batch_kind_.push_back(kv.first);
    }
    const size_t nb = batch_kind_.size();
    std::vector<int> b_args(nb, -1), b_opt(nb, -1);
    std::vector<int32_t> batch_of(nf), slot_of(nf), b_n(nb, 0);
    {
      int last_kind = -1, last_batch = -1;
      for (size_t fi = 0; fi < nf; ++fi) {
        const Factor<Scalar>& f = factors_[fi];
        if (f.Kind() != last_kind) {
          last_kind = f.Kind();
          last_batch = kind2batch[last_kind];
        }
        const int bi = last_batch;
        if (b_args[bi] < 0) {
          b_args[bi] = static_cast<int>(f.AllKeys().size());
          b_opt[bi] = static_cast<int>(f.OptimizedKeys().size());
        }
        SYM_ASSERT(b_args[bi] == static_cast<int>(f.AllKeys().size()) && b_opt[bi] == static_cast<int>(f.OptimizedKeys().size()));
        batch_of[fi] = bi;
        slot_of[fi] = b_n[bi]++;
      }
    }
    flat_args_.assign(nb, {});
    flat_opt_.assign(nb, {});
    batch_fidx_.assign(nb, {});
    for (size_t bi = 0; bi < nb; ++bi) {
      flat_args_[bi].resize(static_cast<size_t>(b_args[bi]) * b_n[bi]);
      flat_opt_[bi].resize(static_cast<size_t>(b_opt[bi]) * b_n[bi]);
      batch_fidx_[bi].resize(b_n[bi]);
    }
    internal::ParallelFor(nf, [&](size_t f0, size_t f1) {
      for (size_t fi = f0; fi < f1; ++fi) {
        const Factor<Scalar>& f = factors_[fi];
        const int bi = batch_of[fi];
        const size_t n = b_n[bi], s = slot_of[fi];
        for (int a = 0; a < b_args[bi]; ++a) flat_args_[bi][a * n + s] = values.IndexEntryAt(f.AllKeys()[a]).offset;
        for (int o = 0; o < b_opt[bi]; ++o) {
          auto it = key_index.find(f.OptimizedKeys()[o]);
          flat_opt_[bi][o * n + s] = it == key_index.end() ? -1 : it->second;
        }
        batch_fidx_[bi][s] = static_cast<int32_t>(fi);
      }
    });
    n_values_ = static_cast<int64_t>(values.Data().size());
    int schur_keys = 0;
    if (gpu_.solver == GpuSolverOptions::SCHUR) schur_keys = gpu_.schur_num_keys;
    if (gpu_.solver == GpuSolverOptions::AUTO) schur_keys = AutoSchurKeys(values);
    solver_ = schur_keys > 0 ? SFX_SOLVER_SCHUR : SFX_SOLVER_CHOLESKY;
    schur_keys_ = schur_keys;
    if (std::getenv("SFX_TIMING"))
      std::fprintf(stderr, "[sym] indexed %zu factors over %zu optimized keys in %.2f s\n", nf, keys_.size(),
                   std::chrono::duration<double>(std::chrono::steady_clock::now() - t_index).count());
    handle_ = Create(solver_, schur_keys_);
  }
  // Device problem for the indexed factor graph with the given linear solver (the LM problem, or the
  // sibling ComputeCovariances needs when it eliminates a different set of keys).
  bool CheckDerivativesImpl(const Scalar* data, size_t n, std::array<double, 3>* rel_errors) {
    if (!check_handle_) check_handle_ = Create(solver_, schur_keys_);
    Check(sfx_set_values(check_handle_, data, static_cast<int64_t>(n)), check_handle_);
    int32_t ok = 0;
    double err[3] = {0, 0, 0};
    Check(sfx_check_derivatives(check_handle_, err, &ok, nullptr), check_handle_);
    if (rel_errors) *rel_errors = {err[0], err[1], err[2]};
    return ok != 0;
  }
  void CheckDerivativesAt(const Scalar* data, size_t n) {
    SYM_ASSERT(CheckDerivativesImpl(data, n, nullptr) && "internal::CheckDerivatives(linearizer_, values, index_, linearization, epsilon_)");
  }
  sfx_problem* Create(int solver, int schur_keys) {
    std::vector<sfx_factor_batch> fb;
    for (size_t bi = 0; bi < batch_kind_.size(); ++bi) {
      sfx_factor_batch x{};
      x.kind = batch_kind_[bi];
      x.n = static_cast<int32_t>(batch_fidx_[bi].size());
      x.arg_offsets = flat_args_[bi].data();
      x.opt_keys = flat_opt_[bi].data();
      x.factor_index = batch_fidx_[bi].data();
      fb.push_back(x);
    }
    sfx_problem_desc d{};
    d.abi_version = SFX_ABI_VERSION;
    d.params = ToC(params_);
    d.epsilon = epsilon_;
    d.n_values = n_values_;
    d.n_keys = static_cast<int32_t>(kentries_.size());
    d.keys = kentries_.data();
    d.n_batches = static_cast<int32_t>(fb.size());
    d.batches = fb.data();
    d.n_factors = static_cast<int32_t>(factors_.size());
    d.ordering = gpu_.ordering;
    d.device = gpu_.device;
    d.rank = 0;
    d.world = 1;
    d.comm = nullptr;
    d.solver = solver;
    d.schur_num_keys = schur_keys;
    sfx_problem* h = nullptr;
    sfx_status s = sfx_problem_create(&d, &h);
    if (s != SFX_OK) throw std::runtime_error(std::string("sym::Optimizer<") + name_ + ">: " + sfx_last_error(nullptr));
    return h;
  }
  // internal::SplitCovariancesByKey (covariance_utils.h:172-191)
  void SplitCovariancesByKey(const MatrixX<Scalar>& covariance_block, const std::vector<Key>& keys,
                             std::unordered_map<Key, MatrixX<Scalar>, KeyHash>& covariances_by_key) const {
    int offset = 0;
    for (size_t i = 0; i < keys.size(); ++i) {
      const int dim = kentries_[i].tangent_dim;
      covariances_by_key[keys[i]] = covariance_block.block(offset, offset, dim, dim);
      offset += dim;
    }
    SYM_ASSERT(covariances_by_key.size() == keys.size());
  }
  // Longest trailing run of keys (in keys_ order) that are vectors of dim <= 3 and never share a
  // factor with each other: eliminating them per block is exactly SparseSchurSolver's C.
  int AutoSchurKeys(const Values<Scalar>& values) const {
    const int nk = static_cast<int>(keys_.size());
    int first = nk;
    while (first > 0) {
      const index_entry_t& e = values.IndexEntryAt(keys_[first - 1]);
      if ((e.type == type_t::VECTOR || e.type == type_t::SCALAR) && e.tangent_dim <= 3)
        --first;
      else
        break;
    }
    if (first == nk) return 0;
    // shrink the run until no factor touches two of its keys (BAL: intrinsics and points are both small vectors
    // and meet in every factor; the run must start behind the last intrinsics key)
    for (size_t bi = 0; bi < flat_opt_.size(); ++bi) {
      const size_t n = batch_fidx_[bi].size();
      const size_t n_opt = n ? flat_opt_[bi].size() / n : 0;
      for (size_t s = 0; s < n; ++s) {
        int largest = -1, second = -1;
        for (size_t o = 0; o < n_opt; ++o) {
          const int k = flat_opt_[bi][o * n + s];
          if (k < first) continue;
          if (k > largest) {
            second = largest;
            largest = k;
          } else if (k > second) {
            second = k;
          }
        }
        if (second >= 0) first = std::max(first, second + 1);
      }
    }
    if (first == 0 || first >= nk) return 0;
    return (nk - first) >= nk / 2 ? nk - first : 0;
  }
  void FillPattern(SparseLinearization& lin) {
    int32_t N = 0, M = 0;
    int64_t nnz = 0;
    Check(sfx_get_dims(handle_, &N, &M, &nnz));
    lin.residual.resize(M);
    lin.rhs.resize(N);
    lin.hessian_lower.rows_ = lin.hessian_lower.cols_ = N;
    lin.hessian_lower.outer.resize(N + 1);
    lin.hessian_lower.inner.resize(nnz);
    lin.hessian_lower.values.resize(nnz);
    Check(sfx_get_hessian_pattern(handle_, lin.hessian_lower.outer.data(), lin.hessian_lower.inner.data()));
  }

  optimizer_params_t params_;
  std::vector<Factor<Scalar>> factors_;
  std::string name_;
  std::vector<Key> keys_;
  Scalar epsilon_;
  GpuSolverOptions gpu_;
  sfx_problem* handle_{nullptr};
  int64_t n_values_{0};
  // the indexed problem (kept so that ComputeCovariances can build a sibling with another Schur split)
  std::vector<sfx_key_entry> kentries_;
  std::vector<int> batch_kind_;
  std::vector<std::vector<int32_t>> flat_args_, flat_opt_, batch_fidx_;
  int solver_{SFX_SOLVER_CHOLESKY}, schur_keys_{0};
  sfx_problem* cov_handle_{nullptr};
  sfx_problem* check_handle_{nullptr};  // check_derivatives: same structure and solver, its own LM state
  int cov_schur_keys_{-1};
};
using Optimizerd = Optimizer<double>;

// optimizer_gnc_params_t (lcmtypes/symforce.lcm:203-226)
struct optimizer_gnc_params_t {
  double gnc_update_min_reduction{0};
  double mu_initial{0};
  double mu_step{0};
  double mu_max{0};
};

// GncOptimizer (symforce/opt/gnc_optimizer.h:15-148): optimizes to convergence with the convex cost, then steps
// the convexity parameter mu (a scalar in the Values) towards mu_max, relaxing the damping and continuing the
// optimization after every step.  The iteration budget is the total over all stages.
template <typename BaseOptimizerType>
class GncOptimizer : public BaseOptimizerType {
 public:
  using BaseOptimizer = BaseOptimizerType;
  using Scalar = typename BaseOptimizer::Scalar;

  template <typename... OptimizerArgs>
  GncOptimizer(const optimizer_params_t& optimizer_params, const optimizer_gnc_params_t& gnc_params,
               const Key& gnc_mu_key, OptimizerArgs&&... args)
      : BaseOptimizer(optimizer_params, std::forward<OptimizerArgs>(args)...),
        gnc_params_(gnc_params),
        gnc_mu_key_(gnc_mu_key) {}

  using BaseOptimizerType::Optimize;
  void Optimize(Values<Scalar>& values, int num_iterations, bool populate_best_linearization,
                typename BaseOptimizer::Stats& stats) override {
    if (num_iterations < 0) num_iterations = this->Params().iterations;
    bool updating_gnc = gnc_params_.mu_initial < gnc_params_.mu_max && gnc_params_.mu_step > 0.0;
    values.template Set<Scalar>(gnc_mu_key_, static_cast<Scalar>(gnc_params_.mu_initial));
    optimizer_params_t optimizer_params = this->Params();
    const double early_exit_min_reduction = optimizer_params.early_exit_min_reduction;
    if (updating_gnc) optimizer_params.early_exit_min_reduction = gnc_params_.gnc_update_min_reduction;
    this->UpdateParams(optimizer_params);
    this->OptimizeImpl(values, num_iterations, populate_best_linearization, stats, false);
    while (static_cast<int>(stats.iterations.size()) < num_iterations) {
      if (stats.status != optimization_status_t::SUCCESS) return;  // the previous stage did not converge
      if (!updating_gnc) return;
      values.template Set<Scalar>(gnc_mu_key_,
                                  values.template At<Scalar>(gnc_mu_key_) + static_cast<Scalar>(gnc_params_.mu_step));
      this->RelaxDampingToInitial();
      if (values.template At<Scalar>(gnc_mu_key_) >= gnc_params_.mu_max) {
        values.template Set<Scalar>(gnc_mu_key_, static_cast<Scalar>(gnc_params_.mu_max));
        optimizer_params.early_exit_min_reduction = early_exit_min_reduction;
        this->UpdateParams(optimizer_params);
        updating_gnc = false;
      }
      this->OptimizeImpl(values, num_iterations - static_cast<int>(stats.iterations.size()),
                         populate_best_linearization, stats, true);
    }
  }

 private:
  optimizer_gnc_params_t gnc_params_;
  Key gnc_mu_key_;
};

// sym::Optimize (optimizer.h:334-340)
template <typename Scalar>
OptimizationStats Optimize(const optimizer_params_t& params, std::vector<Factor<Scalar>> factors, Values<Scalar>& values,
                           const Scalar epsilon = kDefaultEpsilon<Scalar>) {
  Optimizer<Scalar> optimizer(params, std::move(factors), "sym::Optimize", {}, epsilon);
  return optimizer.Optimize(values);
}

}  // namespace sym
