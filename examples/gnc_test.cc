// Known-answer check of sym::GncOptimizer on the GPU path, on the problem of the reference's GNC test
// (test/symforce_gnc_test.cc:23-86): a 5-vector x pulled towards 20 samples through gnc_factors::BarronFactor (a device
// kind), three of the samples being outliers near 10.  What the reference test asserts is evaluated at the end:
//   9 iteration records in total, |x| < 0.1, |x| at least 5 times smaller than without GNC, status SUCCESS.
// Prints GNC_TEST_OK when all four hold.
#include <cstdio>
#include <cstdlib>
#include <random>
#include <string>
#include <vector>

#include <sym/sym.h>

namespace {

constexpr int kSamples = 20;
constexpr int kOutliers = 3;
constexpr double kOptimizerEpsilon = 1e-12;
const sym::Key kX('x'), kMu('u'), kEps('e');

sym::Key SampleKey(int i) { return sym::Key('y', i); }

// the sample stream of the reference test: std::mt19937(42), one sym::Random<Vector5d> per sample, scaled by 0.1,
// outliers shifted to 10
sym::Valuesd MakeValues() {
  sym::Valuesd values;
  values.Set<sym::Vector5d>(kX, sym::Vector5d::Ones());
  values.Set(kEps, sym::kDefaultEpsilond);
  std::mt19937 rng(42);
  for (int i = 0; i < kSamples; ++i) {
    const sym::Vector5d noise = sym::Random<sym::Vector5d>(rng) * 0.1;
    values.Set<sym::Vector5d>(SampleKey(i), i < kOutliers ? sym::Vector5d::Constant(10) + noise : noise);
  }
  return values;
}

std::vector<sym::Factord> MakeFactors() {
  std::vector<sym::Factord> factors;
  factors.reserve(kSamples);
  for (int i = 0; i < kSamples; ++i)
    factors.push_back(sym::Factord::Hessian(gnc_factors::BarronFactor<double>, {kX, SampleKey(i), kMu, kEps}, {kX}));
  return factors;
}

struct Check {
  const char* what;
  bool ok;
};

}  // namespace

int main() {
  const sym::Valuesd initial = MakeValues();
  if (std::getenv("GNC_TEST_PRINT_VALUES")) {  // host-only: lets the CPU tests compare the sample data with their fixture
    for (double v : initial.Data()) std::printf("%.17g\n", v);
    return 0;
  }
  const std::vector<sym::Factord> factors = MakeFactors();
  const sym::optimizer_params_t params = sym::DefaultOptimizerParams();

  sym::optimizer_gnc_params_t gnc{};
  gnc.mu_initial = 0.0;
  gnc.mu_step = 0.33;
  gnc.mu_max = 0.99;
  gnc.gnc_update_min_reduction = 1e-3;

  // with graduated non-convexity: mu walks 0 -> 0.33 -> 0.66 -> 0.99
  sym::Valuesd with_gnc = initial;
  sym::GncOptimizer<sym::Optimizerd> gnc_optimizer(params, gnc, kMu, factors, "sym::Optimize", std::vector<sym::Key>{},
                                                   kOptimizerEpsilon);
  const auto stats = gnc_optimizer.Optimize(with_gnc);

  // without: the convex (mu = 0) cost only
  sym::Valuesd without_gnc = initial;
  without_gnc.Set(kMu, 0.0);
  sym::Optimize(params, factors, without_gnc);

  const double x_gnc = with_gnc.At<sym::Vector5d>(kX).norm();
  const double x_plain = without_gnc.At<sym::Vector5d>(kX).norm();
  std::printf("records %zu  |x| with GNC %.6f  without %.6f  status %d\n", stats.iterations.size(), x_gnc, x_plain,
              static_cast<int>(stats.status));

  const Check checks[] = {
      {"9 iteration records", stats.iterations.size() == 9},
      {"|x| < 0.1", x_gnc < 0.1},
      {"5 x closer to zero than the plain optimization", x_gnc * 5 < x_plain},
      {"status SUCCESS", stats.status == sym::optimization_status_t::SUCCESS},
  };
  bool all = true;
  for (const Check& c : checks) {
    if (!c.ok) std::printf("FAILED: %s\n", c.what);
    all = all && c.ok;
  }
  std::printf(all ? "GNC_TEST_OK\n" : "GNC_TEST_FAILED\n");
  return all ? 0 : 1;
}
