// The reference's GNC test (test/symforce_gnc_test.cc:14-86) against the GPU path: the same Values, factors
// (gnc_factors::BarronFactor, a device kind), parameters and checks, written with the sym:: layer of include/sym/sym.h.
// Prints GNC_TEST_OK when every CHECK of the reference test holds.
#include <cstdio>
#include <random>

#include <sym/sym.h>

sym::optimizer_gnc_params_t DefaultGncParams() {
  sym::optimizer_gnc_params_t params{};
  params.mu_initial = 0;
  params.mu_max = 0.99;
  params.mu_step = 0.33;
  params.gnc_update_min_reduction = 1e-3;
  return params;
}

int main() {
  static constexpr const double kEpsilon = 1e-12;
  const int n_residuals = 20;
  const int n_outliers = 3;

  // Create values
  sym::Valuesd initial_values;
  initial_values.Set<sym::Vector5d>('x', sym::Vector5d::Ones());
  initial_values.Set('e', sym::kDefaultEpsilond);

  // Pick random normal samples, with some outliers
  std::mt19937 gen(42);
  for (int i = 0; i < n_residuals; i++) {
    if (i < n_outliers) {
      initial_values.Set<sym::Vector5d>({'y', i}, sym::Vector5d::Constant(10) + sym::Random<sym::Vector5d>(gen) * 0.1);
    } else {
      initial_values.Set<sym::Vector5d>({'y', i}, sym::Random<sym::Vector5d>(gen) * 0.1);
    }
  }

  std::vector<sym::Factord> factors;
  for (int i = 0; i < n_residuals; i++) {
    factors.push_back(sym::Factord::Hessian(gnc_factors::BarronFactor<double>, {'x', {'y', i}, 'u', 'e'}, {'x'}));
  }

  auto params = sym::DefaultOptimizerParams();

  sym::GncOptimizer<sym::Optimizerd> gnc_optimizer(params, DefaultGncParams(), 'u', factors, "sym::Optimize",
                                                   /* keys */ std::vector<sym::Key>{}, kEpsilon);

  sym::Valuesd gnc_optimized_values = initial_values;
  const auto gnc_stats = gnc_optimizer.Optimize(gnc_optimized_values);

  sym::Valuesd regular_optimized_values = initial_values;
  regular_optimized_values.Set('u', 0.0);
  sym::Optimize(params, factors, regular_optimized_values);

  const sym::Vector5d gnc_optimized_x = gnc_optimized_values.At<sym::Vector5d>('x');
  const sym::Vector5d regular_optimized_x = regular_optimized_values.At<sym::Vector5d>('x');
  std::printf("iterations %zu, |x_gnc| %.6f, |x_regular| %.6f, status %d\n", gnc_stats.iterations.size(),
              gnc_optimized_x.norm(), regular_optimized_x.norm(), static_cast<int>(gnc_stats.status));
  bool ok = true;
  ok = ok && gnc_stats.iterations.size() == 9;
  ok = ok && gnc_optimized_x.norm() < 0.1;
  ok = ok && gnc_optimized_x.norm() * 5 < regular_optimized_x.norm();
  ok = ok && gnc_stats.status == sym::optimization_status_t::SUCCESS;
  std::printf(ok ? "GNC_TEST_OK\n" : "GNC_TEST_FAILED\n");
  return ok ? 0 : 1;
}
