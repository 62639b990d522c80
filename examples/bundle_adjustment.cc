// The reference's bundle_adjustment example (symforce/examples/bundle_adjustment/run_bundle_adjustment.cc:20-190,
// build_example_state.cc) on the GPU path: two views, inverse-range landmarks, relative-pose priors, Barron-robust
// reprojection factors whose convexity is the Values scalar GNC_MU.  Runs it once with sym::Optimizer like the
// reference example, and once with sym::GncOptimizer (symforce/opt/gnc_optimizer.h), which steps GNC_MU from 0 to
// 0.99 and continues the optimization after every step.
#include <cmath>
#include <cstdio>
#include <random>

#include <sym/sym.h>

namespace Var {
constexpr char VIEW = 'v', CALIBRATION = 'c', POSE_PRIOR_T = 'T', POSE_PRIOR_SQRT_INFO = 's', LANDMARK = 'l',
               LANDMARK_PRIOR = 'P', LANDMARK_PRIOR_SIGMA = 'S', MATCH_SOURCE_COORDS = 'm', MATCH_TARGET_COORDS = 'M',
               MATCH_WEIGHT = 'W', GNC_MU = 'u', GNC_SCALE = 'C', EPSILON = 'e';
}

static constexpr int kNumViews = 2, kNumLandmarks = 20;
static constexpr double kEpsilon = 1e-10;

static std::vector<sym::Factord> BuildFactors() {
  std::vector<sym::Factord> factors;
  for (int i = 0; i < kNumViews; i++)
    for (int j = 0; j < kNumViews; j++) {
      if (i == j) continue;
      factors.push_back(sym::Factord::Hessian(
          sym::BetweenFactorPose3<double>,
          {{Var::VIEW, i}, {Var::VIEW, j}, {Var::POSE_PRIOR_T, i, j}, {Var::POSE_PRIOR_SQRT_INFO, i, j}, Var::EPSILON},
          {{Var::VIEW, i}, {Var::VIEW, j}}));
    }
  for (int i = 1; i < kNumViews; i++)
    for (int l = 0; l < kNumLandmarks; l++)
      factors.push_back(sym::Factord::Hessian(sym::InverseRangeLandmarkPriorFactor<double>,
                                              {{Var::LANDMARK, l},
                                               {Var::LANDMARK_PRIOR, i, l},
                                               {Var::MATCH_WEIGHT, i, l},
                                               {Var::LANDMARK_PRIOR_SIGMA, i, l},
                                               Var::EPSILON},
                                              {{Var::LANDMARK, l}}));
  for (int i = 1; i < kNumViews; i++)
    for (int l = 0; l < kNumLandmarks; l++)
      factors.push_back(sym::Factord::Hessian(sym::InverseRangeLandmarkLinearGncFactor<double>,
                                              {{Var::VIEW, 0},
                                               {Var::CALIBRATION, 0},
                                               {Var::VIEW, i},
                                               {Var::CALIBRATION, i},
                                               {Var::LANDMARK, l},
                                               {Var::MATCH_SOURCE_COORDS, i, l},
                                               {Var::MATCH_TARGET_COORDS, i, l},
                                               {Var::MATCH_WEIGHT, i, l},
                                               Var::GNC_MU,
                                               Var::GNC_SCALE,
                                               Var::EPSILON},
                                              {{Var::VIEW, 0}, {Var::VIEW, i}, {Var::LANDMARK, l}}));
  return factors;
}

static sym::Valuesd BuildValues(std::mt19937& gen) {
  std::normal_distribution<double> n(0, 1);
  std::uniform_real_distribution<double> u(0, 1);
  sym::Valuesd values;
  values.Set(Var::EPSILON, kEpsilon);
  values.Set(Var::GNC_SCALE, 10.0);
  values.Set(Var::GNC_MU, 0.0);
  const double fx = 740, fy = 740, cx = 639.5, cy = 359.5;
  for (int i = 0; i < kNumViews; i++)
    values.Set({Var::CALIBRATION, i}, sym::LinearCameraCald(Eigen::Vector2d(fx, fy), Eigen::Vector2d(cx, cy)));
  // view 0 is the world frame; view 1 is displaced
  const sym::Pose3d view0;
  const sym::Pose3d view1 = view0.Retract(sym::Vector6d(0.03, -0.06, 0.03, 0.63, 0.12, -0.06));
  values.Set({Var::VIEW, 0}, view0);
  values.Set({Var::VIEW, 1}, view1.Retract(sym::Vector6d(0.03 * n(gen), 0.03 * n(gen), 0.03 * n(gen), 0.03 * n(gen),
                                                         0.03 * n(gen), 0.03 * n(gen))));
  // relative pose priors in both directions: between(view_i, view_j) with a weak information matrix
  const Eigen::Vector3d t1 = view1.Position();
  const auto q1 = view1.Rotation().Data();
  const sym::Rot3d r1_inv(Eigen::Vector4d(-q1[0], -q1[1], -q1[2], q1[3]), false);
  const Eigen::Vector3d mt = r1_inv.Rotate(Eigen::Vector3d(-t1[0], -t1[1], -t1[2]));
  values.Set({Var::POSE_PRIOR_T, 0, 1}, view1);                                // 0_T_1 = view1 (view0 = identity)
  values.Set({Var::POSE_PRIOR_T, 1, 0}, sym::Pose3d(r1_inv, mt));              // 1_T_0
  Eigen::Matrix<double, 6, 6> sqrt_info = Eigen::Matrix<double, 6, 6>::Identity() * (1.0 / 0.3);
  values.Set({Var::POSE_PRIOR_SQRT_INFO, 0, 1}, sqrt_info);
  values.Set({Var::POSE_PRIOR_SQRT_INFO, 1, 0}, sqrt_info);
  // correspondences: pixels in view 0, inverse ranges 1 / U(2.5, 30), projected into view 1 with 1 px noise
  for (int l = 0; l < kNumLandmarks; l++) {
    const double px = 100 + 1000 * u(gen), py = 100 + 500 * u(gen);
    const double inv_range = 1.0 / (2.5 + 27.5 * u(gen));
    Eigen::Vector3d ray((px - cx) / fx, (py - cy) / fy, 1.0);
    const double nrm = std::sqrt(ray[0] * ray[0] + ray[1] * ray[1] + ray[2] * ray[2]);
    const Eigen::Vector3d pw(ray[0] / nrm / inv_range, ray[1] / nrm / inv_range, ray[2] / nrm / inv_range);
    const Eigen::Vector3d pc = r1_inv.Rotate(Eigen::Vector3d(pw[0] - t1[0], pw[1] - t1[1], pw[2] - t1[2]));
    double tx = fx * pc[0] / pc[2] + cx + n(gen), ty = fy * pc[1] / pc[2] + cy + n(gen);
    if (l < 2) {  // two gross outliers: what the robust cost is for
      tx += 80;
      ty -= 60;
    }
    values.Set({Var::LANDMARK, l}, inv_range * std::min(2.0, std::max(0.5, 1 + 0.5 * n(gen))));
    values.Set({Var::MATCH_SOURCE_COORDS, 1, l}, Eigen::Vector2d(px, py));
    values.Set({Var::MATCH_TARGET_COORDS, 1, l}, Eigen::Vector2d(tx, ty));
    values.Set({Var::MATCH_WEIGHT, 1, l}, 1.0);
    values.Set({Var::LANDMARK_PRIOR, 1, l}, inv_range);
    values.Set({Var::LANDMARK_PRIOR_SIGMA, 1, l}, 100.0);
  }
  return values;
}

int main() {
  std::mt19937 gen(42);
  const sym::Valuesd initial = BuildValues(gen);
  const std::vector<sym::Factord> factors = BuildFactors();
  std::vector<sym::Key> keys;  // ComputeKeysToOptimizeWithoutView0 (run_bundle_adjustment.cc:107-125)
  for (const auto& k : sym::ComputeKeysToOptimize(factors))
    if (!(k == sym::Key(Var::VIEW, 0))) keys.push_back(k);

  auto params = sym::DefaultOptimizerParams();  // example_utils::OptimizerParams()
  params.iterations = 50;
  params.lambda_up_factor = 10.0;
  params.lambda_down_factor = 0.1;
  params.lambda_lower_bound = 1e-8;

  sym::Valuesd v_plain = initial;
  sym::Optimizerd optimizer(params, factors, "BundleAdjustmentOptimizer", keys, kEpsilon);
  const auto stats = optimizer.Optimize(v_plain);
  const auto& best = stats.iterations[stats.best_index];
  std::printf("Optimizer: %zu records, status %d, initial error %.9e, best error %.9e\n", stats.iterations.size(),
              (int)stats.status, stats.iterations.front().new_error, best.new_error);

  sym::optimizer_gnc_params_t gnc{};  // test/symforce_gnc_test.cc:14-21
  gnc.mu_initial = 0;
  gnc.mu_max = 0.99;
  gnc.mu_step = 0.33;
  gnc.gnc_update_min_reduction = 1e-3;
  sym::Valuesd v_gnc = initial;
  sym::GncOptimizer<sym::Optimizerd> gnc_optimizer(params, gnc, Var::GNC_MU, factors, "GncBundleAdjustment", keys, kEpsilon);
  const auto gstats = gnc_optimizer.Optimize(v_gnc);
  bool numbering = true;
  for (size_t i = 0; i < gstats.iterations.size(); i++) numbering = numbering && gstats.iterations[i].iteration == (int)i - 1;
  std::printf("GncOptimizer: %zu records, status %d, final mu %.2f, best error %.9e, numbering continuous %d\n",
              gstats.iterations.size(), (int)gstats.status, v_gnc.At<double>(Var::GNC_MU),
              gstats.iterations[gstats.best_index].new_error, (int)numbering);
  // the robust cost (mu = 0.99) discounts the two outliers: their landmarks' reprojection no longer drags view 1
  const auto p_plain = v_plain.At<sym::Pose3d>({Var::VIEW, 1}).Position();
  const auto p_gnc = v_gnc.At<sym::Pose3d>({Var::VIEW, 1}).Position();
  std::printf("view 1 position: plain [%.4f %.4f %.4f]  gnc [%.4f %.4f %.4f]  truth [0.63 0.12 -0.06]\n", p_plain[0], p_plain[1],
              p_plain[2], p_gnc[0], p_gnc[1], p_gnc[2]);
  const bool ok = stats.status == sym::optimization_status_t::SUCCESS && gstats.status == sym::optimization_status_t::SUCCESS &&
                  std::fabs(v_gnc.At<double>(Var::GNC_MU) - 0.99) < 1e-12 && numbering &&
                  gstats.iterations.size() > stats.iterations.size() && std::isfinite(best.new_error);
  std::printf(ok ? "GNC_OK\n" : "GNC_FAIL\n");
  return ok ? 0 : 1;
}
