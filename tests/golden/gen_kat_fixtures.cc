// Generates tests/golden/kat_initial_values.json: the random initial values of the reference's
// optimizer known-answer tests, reproduced with the same libstdc++ facilities the reference uses
// (std::mt19937(42) + a fresh std::normal_distribution<double> per sym::Random<VectorN> call,
// gen/cpp/sym/ops/matrix/storage_ops.h:211-218), then pushed through retract / FromTangent with
// sym::kDefaultEpsilon (gen/cpp/sym/ops/{pose3,rot3}/lie_group_ops.cc).
//   pose_smoothing      test/symforce_optimizer_test.cc:79-134   (10 Pose3, prior_start.Retract(0.4*N6))
//   rotation_smoothing  test/symforce_optimizer_test.cc:183-236  (10 Rot3, identity.Retract(0.4*N3))
//   frozen_keys         test/symforce_optimizer_test.cc:276-312  (3 Rot3, FromTangent(0.4*N3))
//   gnc_test            test/symforce_gnc_test.cc:33-42          (20 Vector5, 3 outliers)
// Build & run:  g++ -O2 -std=c++17 gen_kat_fixtures.cc -o /tmp/gen_kat && /tmp/gen_kat > kat_initial_values.json
#include <cmath>
#include <cstdio>
#include <limits>
#include <random>
#include <vector>

static const double kEps = 10 * std::numeric_limits<double>::epsilon();

template <int N>
static void random_vec(std::mt19937& gen, double* v) {
  std::normal_distribution<double> d{};
  for (int i = 0; i < N; ++i) v[i] = d(gen);
}

static void normalize4(double* q) {
  double n2 = q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3];
  if (n2 > 0) {
    double n = std::sqrt(n2);
    for (int i = 0; i < 4; ++i) q[i] /= n;
  }
}

static void rot3_retract(const double* a, const double* v, double* out) {
  const double t0 = std::sqrt(kEps * kEps + v[0] * v[0] + v[1] * v[1] + v[2] * v[2]);
  const double t1 = 0.5 * t0;
  const double s = std::sin(t1) / t0, c = std::cos(t1);
  out[0] = a[0] * c + a[1] * (s * v[2]) + (a[3] * s) * v[0] - (a[2] * s) * v[1];
  out[1] = a[1] * c + (a[3] * s) * v[1] + (a[2] * s) * v[0] - (a[0] * s) * v[2];
  out[2] = a[2] * c + a[3] * (s * v[2]) + (a[0] * s) * v[1] - (a[1] * s) * v[0];
  out[3] = -a[2] * (s * v[2]) + a[3] * c - (a[0] * s) * v[0] - (a[1] * s) * v[1];
  normalize4(out);
}

static void rot3_from_tangent(const double* v, double* out) {
  const double t0 = std::sqrt(kEps * kEps + v[0] * v[0] + v[1] * v[1] + v[2] * v[2]);
  const double t1 = 0.5 * t0;
  const double s = std::sin(t1) / t0;
  out[0] = s * v[0];
  out[1] = s * v[1];
  out[2] = s * v[2];
  out[3] = std::cos(t1);
  normalize4(out);
}

static void print_arr(const char* name, const std::vector<double>& v, bool last) {
  std::printf("  \"%s\": [", name);
  for (size_t i = 0; i < v.size(); ++i) std::printf("%s%.17g", i ? ", " : "", v[i]);
  std::printf("]%s\n", last ? "" : ",");
}

int main() {
  std::printf("{\n");
  {
    std::mt19937 gen(42);
    std::vector<double> out;
    const double id[7] = {0, 0, 0, 1, 0, 0, 0};
    for (int i = 0; i < 10; ++i) {
      double v[6], p[7];
      random_vec<6>(gen, v);
      for (double& x : v) x *= 0.4;
      rot3_retract(id, v, p);
      p[4] = id[4] + v[3];
      p[5] = id[5] + v[4];
      p[6] = id[6] + v[5];
      out.insert(out.end(), p, p + 7);
    }
    print_arr("pose_smoothing", out, false);
  }
  {
    std::mt19937 gen(42);
    std::vector<double> out;
    const double id[4] = {0, 0, 0, 1};
    for (int i = 0; i < 10; ++i) {
      double v[3], q[4];
      random_vec<3>(gen, v);
      for (double& x : v) x *= 0.4;
      rot3_retract(id, v, q);
      out.insert(out.end(), q, q + 4);
    }
    print_arr("rotation_smoothing", out, false);
  }
  {
    std::mt19937 gen(42);
    std::vector<double> out;
    for (int i = 0; i < 3; ++i) {
      double v[3], q[4];
      random_vec<3>(gen, v);
      for (double& x : v) x *= 0.4;
      rot3_from_tangent(v, q);
      out.insert(out.end(), q, q + 4);
    }
    print_arr("frozen_keys", out, false);
  }
  {
    // test/symforce_gnc_test.cc:33-42: 20 Vector5 samples y_i, the first 3 are outliers around 10
    std::mt19937 gen(42);
    std::vector<double> out;
    for (int i = 0; i < 20; ++i) {
      double v[5];
      random_vec<5>(gen, v);
      for (double& x : v) x = (i < 3 ? 10.0 : 0.0) + 0.1 * x;
      out.insert(out.end(), v, v + 5);
    }
    print_arr("gnc_test", out, true);
  }
  std::printf("}\n");
  return 0;
}
