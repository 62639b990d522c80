// robot_3d_localization on the GPU path, written against the sym:: API exactly like the reference
// example (symforce/examples/robot_3d_localization/run_dynamic_size.cc:25-105, common.h:24-62).
#include <algorithm>
#include <cmath>
#include <cstdio>

#include <sym/sym.h>

#include "robot3d_data.inc"

namespace Keys {
static const sym::Key WORLD_T_BODY = 'w';
static const sym::Key WORLD_T_LANDMARK = 'W';
static const sym::Key ODOMETRY_DIAGONAL_SIGMAS = 'o';
static const sym::Key ODOMETRY_RELATIVE_POSE_MEASUREMENTS = 'O';
static const sym::Key MATCHING_SIGMA = 'm';
static const sym::Key BODY_T_LANDMARK_MEASUREMENTS = 'b';
static const sym::Key EPSILON = 'e';
}  // namespace Keys

static sym::Factord CreateMatchingFactor(int i, int j) {
  return sym::Factord::Hessian(sym::MatchingFactor<double>,
                               {Keys::WORLD_T_BODY.WithSuper(i), Keys::WORLD_T_LANDMARK.WithSuper(j),
                                {Keys::BODY_T_LANDMARK_MEASUREMENTS.Letter(), i, j}, Keys::MATCHING_SIGMA},
                               {Keys::WORLD_T_BODY.WithSuper(i)});
}
static sym::Factord CreateOdometryFactor(int i) {
  return sym::Factord::Hessian(
      sym::OdometryFactor<double>,
      {Keys::WORLD_T_BODY.WithSuper(i), Keys::WORLD_T_BODY.WithSuper(i + 1),
       Keys::ODOMETRY_RELATIVE_POSE_MEASUREMENTS.WithSuper(i), Keys::ODOMETRY_DIAGONAL_SIGMAS, Keys::EPSILON},
      {Keys::WORLD_T_BODY.WithSuper(i), Keys::WORLD_T_BODY.WithSuper(i + 1)});
}

int main() {
  sym::Valuesd values;
  for (int i = 0; i < kNumPoses; i++) values.Set(Keys::WORLD_T_BODY.WithSuper(i), sym::Pose3d());
  for (int i = 0; i < kNumLandmarks; i++)
    values.Set(Keys::WORLD_T_LANDMARK.WithSuper(i), sym::Vector3d::FromData(kLandmarks + 3 * i));
  values.Set(Keys::ODOMETRY_DIAGONAL_SIGMAS, sym::Vector6d(0.05, 0.05, 0.05, 0.2, 0.2, 0.2));
  for (int i = 0; i < kNumPoses - 1; i++)
    values.Set(Keys::ODOMETRY_RELATIVE_POSE_MEASUREMENTS.WithSuper(i), sym::Pose3d(sym::Vector7d::FromData(kOdometry + 7 * i)));
  values.Set(Keys::MATCHING_SIGMA, 0.1);
  for (int i = 0; i < kNumPoses; i++)
    for (int j = 0; j < kNumLandmarks; j++)
      values.Set({Keys::BODY_T_LANDMARK_MEASUREMENTS.Letter(), i, j},
                 sym::Vector3d::FromData(kBodyTLandmark + 3 * (i * kNumLandmarks + j)));
  values.Set(Keys::EPSILON, sym::kDefaultEpsilond);

  const sym::Valuesd initial_values = values;

  std::vector<sym::Factord> factors;
  for (int i = 0; i < kNumPoses; i++)
    for (int j = 0; j < kNumLandmarks; j++) factors.push_back(CreateMatchingFactor(i, j));
  for (int i = 0; i < kNumPoses - 1; i++) factors.push_back(CreateOdometryFactor(i));

  sym::optimizer_params_t params = sym::DefaultOptimizerParams();
  params.initial_lambda = 1e4;
  params.lambda_down_factor = 1 / 2.;
  sym::Optimizer<double> optimizer(params, factors, "Robot3DScanMatchingOptimizerDynamic");
  const auto stats = optimizer.Optimize(values);

  const auto& first_iter = stats.iterations.front();
  const auto& last_iter = stats.iterations.back();
  const auto& best_iter = stats.iterations[stats.best_index];
  std::printf("Iterations: %d\nLambda: %.6g\nInitial error: %.10f\nFinal error: %.10f\nStatus: %d\n", last_iter.iteration,
              last_iter.current_lambda, first_iter.new_error, best_iter.new_error, static_cast<int>(stats.status));
  for (int i = 0; i < kNumPoses; i++) {
    const auto p = values.At<sym::Pose3d>(Keys::WORLD_T_BODY.WithSuper(i));
    std::printf("Pose %d: t = [%.6f %.6f %.6f]\n", i, p.Data()[4], p.Data()[5], p.Data()[6]);
  }
  // marginal covariances of every optimized key at the optimum (Optimizer::ComputeAllCovariances), and of the
  // first pose alone with everything else eliminated... the other poses are 6-dim, so that split is not
  // block diagonal: the full inverse is the path for this problem
  {
    const auto lin = optimizer.Linearize(values);
    std::unordered_map<sym::Key, sym::MatrixX<double>, sym::KeyHash> covs;
    optimizer.ComputeAllCovariances(lin, covs);
    for (int i = 0; i < kNumPoses; i++) {
      const auto& c = covs.at(Keys::WORLD_T_BODY.WithSuper(i));
      double tr = 0;
      for (int d = 0; d < c.rows(); d++) tr += c(d, d);
      std::printf("Covariance trace %d: %.12e\n", i, tr);
    }
  }
  // debug_stats + include_jacobians (optimization_stats.h:40-75, levenberg_marquardt_solver.tcc:115-122, 165-176):
  // every record carries update / values / residual / jacobian_values; J^T r of the best record, rebuilt through
  // JacobianView, is the rhs of the linearization at the optimized values
  bool debug_ok = false;
  {
    params.debug_stats = true;
    params.include_jacobians = true;
    params.check_derivatives = true;  // SYM_ASSERTs internal::CheckDerivatives at the values of every record
    sym::Optimizer<double> debug_optimizer(params, factors, "Robot3DDebugStats");
    sym::Valuesd debug_values = initial_values;
    const auto debug_stats = debug_optimizer.Optimize(debug_values);
    const auto& rec = debug_stats.iterations[debug_stats.best_index];
    const auto J = debug_stats.JacobianView(rec);
    const auto lin = debug_optimizer.Linearize(debug_values);
    double worst = 0, scale = 0;
    for (int c = 0; c < J.cols(); ++c) {
      double jtr = 0;
      for (int k = J.outerIndexPtr()[c]; k < J.outerIndexPtr()[c + 1]; ++k)
        jtr += J.valuePtr()[k] * rec.residual[J.innerIndexPtr()[k]];
      worst = std::max(worst, std::abs(jtr - lin.rhs[c]));
      scale = std::max(scale, std::abs(lin.rhs[c]));
    }
    const size_t N = lin.rhs.size();
    std::array<double, 3> derivative_errors{};
    const bool derivatives_ok = debug_optimizer.CheckDerivatives(debug_values, &derivative_errors);
    std::printf("Derivative check: numerical J %.3e, J^T J %.3e, J^T r %.3e\n", derivative_errors[0], derivative_errors[1],
                derivative_errors[2]);
    debug_ok = derivatives_ok && debug_stats.iterations.size() == stats.iterations.size() && worst <= 1e-9 * scale &&
               J.rows() == static_cast<int>(rec.residual.size()) && J.cols() == static_cast<int>(N) &&
               J.nonZeros() == lin.jacobian.nonZeros() && rec.update.size() == N &&
               debug_stats.iterations.front().update.empty() &&
               debug_stats.iterations.front().jacobian_values.size() == rec.jacobian_values.size() &&
               debug_stats.linear_solver_ordering.size() == N && debug_stats.cholesky_factor_sparsity.shape.empty();
    std::printf("Debug stats: %zu records, J %d x %d nnz %lld, |J^T r - rhs| %.3e of %.3e: %s\n",
                debug_stats.iterations.size(), J.rows(), J.cols(), static_cast<long long>(J.nonZeros()), worst, scale,
                debug_ok ? "DEBUG_STATS_OK" : "DEBUG_STATS_MISMATCH");
  }
  // same acceptance check as test/symforce_examples_robot_3d_localization_test.py:49-51
  return (debug_ok && stats.status == sym::optimization_status_t::SUCCESS && best_iter.new_error < 140) ? 0 : 1;
}
