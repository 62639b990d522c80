// Marginal covariances through the sym:: API on the GPU path (Optimizer::ComputeCovariances /
// ComputeAllCovariances, symforce/opt/optimizer.h:190-233): a small BAL-shaped problem with a pose prior per
// camera (which removes the gauge freedom), solved once with the Schur solver and once with the Cholesky
// solver.  Checks the property the reference's test checks (test/symforce_covariance_utils_test.cc:100-140):
// the Schur-complement marginal equals the corresponding block of the full inverse, on every code path:
//   A: Schur problem, ComputeCovariances(cameras)        -> the LM problem's own Schur solver
//   B: Cholesky problem, ComputeCovariances(cameras)     -> sibling problem with the requested split
//   C: Cholesky problem, ComputeAllCovariances           -> (H + eps I)^-1
//   D: Schur problem, ComputeCovariances(cameras, c_is_block_diagonal = false) -> general-C path on a sibling problem
//   E: Cholesky problem, ComputeCovariances(first three poses, false): C holds poses, intrinsics and points
#include <cmath>
#include <cstdio>
#include <random>

#include <sym/sym.h>

static const sym::Key CAM_T_WORLD = 'c';
static const sym::Key INTRINSICS = 'i';
static const sym::Key POINT = 'p';
static const sym::Key PIXEL = 'P';
static const sym::Key PRIOR = 'q';
static const sym::Key SQRT_INFO = 'S';
static const sym::Key EPSILON = 'e';

using CovMap = std::unordered_map<sym::Key, sym::MatrixX<double>, sym::KeyHash>;

static double MaxRelDiff(const CovMap& a, const CovMap& b, const std::vector<sym::Key>& keys) {
  double worst = 0;
  for (const auto& k : keys) {
    const auto &x = a.at(k), &y = b.at(k);
    double scale = 0, diff = 0;
    for (int c = 0; c < x.cols(); c++)
      for (int r = 0; r < x.rows(); r++) {
        scale = std::max(scale, std::fabs(y(r, c)));
        diff = std::max(diff, std::fabs(x(r, c) - y(r, c)));
      }
    worst = std::max(worst, diff / scale);
  }
  return worst;
}

int main() {
  const int cams = 8, pts = 200, per_pt = 4;
  std::mt19937_64 gen(7);
  std::uniform_real_distribution<double> u(-1, 1);
  std::normal_distribution<double> n(0, 1);
  const double kPi = 3.14159265358979323846;
  sym::Valuesd values;
  std::vector<sym::Factord> factors;
  std::vector<sym::Pose3d> pose;
  std::vector<Eigen::Vector3d> intr, pt;
  for (int i = 0; i < cams; i++) {
    const double th = 2 * kPi * i / cams;
    const sym::Rot3d R = sym::Rot3d::FromTangent(Eigen::Vector3d(kPi, 0, 0)).Compose(
        sym::Rot3d::FromTangent(Eigen::Vector3d(0, 0, 0.3 * std::sin(th))));
    const Eigen::Vector3d RC = R.Rotate(Eigen::Vector3d(5 * std::cos(th), 5 * std::sin(th), 0));
    pose.emplace_back(R, Eigen::Vector3d(-RC[0], -RC[1], -RC[2]));
    intr.emplace_back(1000.0 + 200 * u(gen), 1e-3 * n(gen), 1e-5 * n(gen));
  }
  for (int j = 0; j < pts; j++) pt.emplace_back(10 * u(gen), 10 * u(gen), 20 + 10 * u(gen));
  int obs = 0;
  for (int c = 0; c < cams; c++)
    for (int j = 0; j < pts; j++) {
      int d = std::abs(c - (int)((long long)j * cams / pts));
      d = std::min(d, cams - d);
      if (d > per_pt / 2) continue;
      const Eigen::Vector3d pc = pose[c].Rotation().Rotate(pt[j]);
      const double X = pc[0] + pose[c].Data()[4], Y = pc[1] + pose[c].Data()[5], Z = pc[2] + pose[c].Data()[6];
      const double px = -X / Z, py = -Y / Z, r2 = px * px + py * py;
      const double r = 1 + intr[c][1] * r2 + intr[c][2] * r2 * r2, f = intr[c][0];
      factors.push_back(sym::Factord::Hessian(
          sym::SnavelyReprojectionFactor<double>,
          {CAM_T_WORLD.WithSuper(c), INTRINSICS.WithSuper(c), POINT.WithSuper(j), PIXEL.WithSuper(obs), EPSILON},
          {CAM_T_WORLD.WithSuper(c), INTRINSICS.WithSuper(c), POINT.WithSuper(j)}));
      values.Set(PIXEL.WithSuper(obs), Eigen::Vector2d(f * r * px + 0.5 * n(gen), f * r * py + 0.5 * n(gen)));
      ++obs;
    }
  Eigen::Matrix<double, 6, 6> sqrt_info = Eigen::Matrix<double, 6, 6>::Identity() * 10.0;
  values.Set(SQRT_INFO, sqrt_info);
  for (int i = 0; i < cams; i++) {
    values.Set(PRIOR.WithSuper(i), pose[i]);
    values.Set(CAM_T_WORLD.WithSuper(i), pose[i].Retract(sym::Vector6d(0.01 * n(gen), 0.01 * n(gen), 0.01 * n(gen),
                                                                       0.05 * n(gen), 0.05 * n(gen), 0.05 * n(gen))));
    values.Set(INTRINSICS.WithSuper(i), intr[i]);
    factors.push_back(sym::Factord::Hessian(sym::PriorFactorPose3<double>,
                                            {CAM_T_WORLD.WithSuper(i), PRIOR.WithSuper(i), SQRT_INFO, EPSILON},
                                            {CAM_T_WORLD.WithSuper(i)}));
  }
  for (int j = 0; j < pts; j++)
    values.Set(POINT.WithSuper(j), Eigen::Vector3d(pt[j][0] + 0.1 * n(gen), pt[j][1] + 0.1 * n(gen), pt[j][2] + 0.1 * n(gen)));
  values.Set(EPSILON, sym::kDefaultEpsilond);

  auto params = sym::DefaultOptimizerParams();
  params.lambda_update_type = sym::lambda_update_type_t::DYNAMIC;
  std::vector<sym::Key> cam_keys;
  for (int i = 0; i < cams; i++) cam_keys.push_back(CAM_T_WORLD.WithSuper(i));
  for (int i = 0; i < cams; i++) cam_keys.push_back(INTRINSICS.WithSuper(i));

  // A: Schur problem (AUTO picks the trailing points)
  sym::Valuesd va = values;
  sym::Optimizerd opt_a{params, factors};
  const auto stats_a = opt_a.Optimize(va);
  CovMap cov_a;
  const auto info_a = opt_a.ComputeCovariances(opt_a.Linearize(va), cam_keys, cov_a);

  // B, C: the same problem solved without Schur elimination
  sym::GpuSolverOptions chol;
  chol.solver = sym::GpuSolverOptions::CHOLESKY;
  sym::Valuesd vb = values;
  sym::Optimizerd opt_b{params, factors, "chol", {}, sym::kDefaultEpsilond, chol};
  const auto stats_b = opt_b.Optimize(vb);
  const auto lin_b = opt_b.Linearize(vb);
  CovMap cov_b, cov_c;
  const auto info_b = opt_b.ComputeCovariances(lin_b, cam_keys, cov_b);
  opt_b.ComputeAllCovariances(lin_b, cov_c);

  // D, E: the general-C branch (internal/covariance_utils.h:41-103, 142-145)
  CovMap cov_d, cov_e;
  const auto info_d = opt_a.ComputeCovariances(opt_a.Linearize(va), cam_keys, cov_d, /*c_is_block_diagonal=*/false);
  const std::vector<sym::Key> three_poses(cam_keys.begin(), cam_keys.begin() + 3);
  const auto info_e = opt_b.ComputeCovariances(lin_b, three_poses, cov_e, /*c_is_block_diagonal=*/false);
  const double d_ad = MaxRelDiff(cov_a, cov_d, cam_keys), d_ce = MaxRelDiff(cov_e, cov_c, three_poses);

  const double final_a = stats_a.iterations[stats_a.best_index].new_error;
  const double final_b = stats_b.iterations[stats_b.best_index].new_error;
  const double d_ab = MaxRelDiff(cov_a, cov_b, cam_keys), d_bc = MaxRelDiff(cov_b, cov_c, cam_keys);
  std::printf("final error schur %.12e cholesky %.12e\n", final_a, final_b);
  std::printf("covariance keys: A %zu B %zu C %zu (of %zu optimized keys)\n", cov_a.size(), cov_b.size(), cov_c.size(),
              opt_b.Keys().size());
  std::printf("max rel diff A vs B: %.3e\nmax rel diff B vs C: %.3e\n", d_ab, d_bc);
  std::printf("max rel diff A vs D: %.3e\nmax rel diff E vs C: %.3e\n", d_ad, d_ce);
  std::printf("sigma(f) of camera 0: %.6e\n", std::sqrt(cov_a.at(INTRINSICS.WithSuper(0))(0, 0)));
  const bool ok = info_a == sym::kSuccess && info_b == sym::kSuccess && cov_a.size() == cam_keys.size() &&
                  cov_c.size() == opt_b.Keys().size() && d_ab < 1e-6 && d_bc < 1e-6 &&
                  info_d == sym::kSuccess && info_e == sym::kSuccess && cov_d.size() == cam_keys.size() &&
                  cov_e.size() == three_poses.size() && d_ad < 1e-6 && d_ce < 1e-6 &&
                  std::fabs(final_a - final_b) < 1e-8 * final_b;
  std::printf(ok ? "COVARIANCE_OK\n" : "COVARIANCE_FAIL\n");
  return ok ? 0 : 1;
}
