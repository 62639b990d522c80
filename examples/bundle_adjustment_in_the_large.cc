// Bundle-Adjustment-in-the-Large on the GPU path, written against the sym:: API like the reference
// example (symforce/examples/bundle_adjustment_in_the_large/bundle_adjustment_in_the_large.cc:27-140).
//   bal_example <problem.txt>                     read a BAL file (https://grail.cs.washington.edu/projects/bal/)
//   bal_example --synthetic <cams> <pts> <obs/pt>  generate a BAL-shaped problem (datasets are not in this image)
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <algorithm>
#include <array>
#include <random>
#include <string>

#include <sym/sym.h>

static const sym::Key CAM_T_WORLD = 'c';
static const sym::Key INTRINSICS = 'i';
static const sym::Key POINT = 'p';
static const sym::Key PIXEL = 'P';
static const sym::Key EPSILON = 'e';

static sym::Factord MakeFactor(int camera, int point, int pixel) {
  return sym::Factord::Hessian(sym::SnavelyReprojectionFactor<double>,
                               {CAM_T_WORLD.WithSuper(camera), INTRINSICS.WithSuper(camera), POINT.WithSuper(point),
                                PIXEL.WithSuper(pixel), EPSILON},
                               {CAM_T_WORLD.WithSuper(camera), INTRINSICS.WithSuper(camera), POINT.WithSuper(point)});
}

struct Problem {
  std::vector<sym::Factord> factors;
  sym::Valuesd values;
  int num_cameras = 0, num_points = 0, num_observations = 0;
};

static Problem ReadProblem(const std::string& filename) {
  std::ifstream file(filename);
  if (!file) throw std::runtime_error("cannot open " + filename);
  Problem p;
  file >> p.num_cameras >> p.num_points >> p.num_observations;
  for (int i = 0; i < p.num_observations; i++) {
    int camera, point;
    double px, py;
    file >> camera >> point >> px >> py;
    p.factors.push_back(MakeFactor(camera, point, i));
    p.values.Set(PIXEL.WithSuper(i), Eigen::Vector2d(px, py));
  }
  for (int i = 0; i < p.num_cameras; i++) {
    double rx, ry, rz, tx, ty, tz, f, k1, k2;
    file >> rx >> ry >> rz >> tx >> ty >> tz >> f >> k1 >> k2;
    p.values.Set(CAM_T_WORLD.WithSuper(i),
                 sym::Pose3d(sym::Rot3d::FromTangent(Eigen::Vector3d(rx, ry, rz)), Eigen::Vector3d(tx, ty, tz)));
    p.values.Set(INTRINSICS.WithSuper(i), Eigen::Vector3d(f, k1, k2));
  }
  for (int i = 0; i < p.num_points; i++) {
    double x, y, z;
    file >> x >> y >> z;
    p.values.Set(POINT.WithSuper(i), Eigen::Vector3d(x, y, z));
  }
  p.values.Set(EPSILON, sym::kDefaultEpsilond);
  return p;
}

// Cameras on a circle looking at a point cloud (BAL convention: camera looks down -z), every point
// seen by `per_pt` neighbouring cameras, 0.5 px noise, perturbed initial guess.
static Problem SyntheticProblem(int cams, int pts, int per_pt) {
  Problem p;
  p.num_cameras = cams;
  p.num_points = pts;
  std::mt19937_64 gen(0xBA1);
  std::uniform_real_distribution<double> u(-1, 1);
  std::normal_distribution<double> n(0, 1);
  std::vector<sym::Pose3d> truth_pose;
  std::vector<Eigen::Vector3d> truth_intr, truth_pt;
  const double kPi = 3.14159265358979323846;
  for (int i = 0; i < cams; i++) {
    const double th = 2 * kPi * i / cams;
    // world -> camera: yaw about z so that cameras differ, then flip so the cloud (z ~ 20) has negative camera z
    const sym::Rot3d flip = sym::Rot3d::FromTangent(Eigen::Vector3d(kPi, 0, 0));
    const sym::Rot3d yaw = sym::Rot3d::FromTangent(Eigen::Vector3d(0, 0, 0.3 * std::sin(th)));
    const sym::Rot3d R = flip.Compose(yaw);
    const Eigen::Vector3d C(5 * std::cos(th), 5 * std::sin(th), 0);
    const Eigen::Vector3d RC = R.Rotate(C);
    truth_pose.emplace_back(R, Eigen::Vector3d(-RC[0], -RC[1], -RC[2]));
    truth_intr.emplace_back(1000.0 + 200 * u(gen), 1e-3 * n(gen), 1e-5 * n(gen));
  }
  for (int j = 0; j < pts; j++) truth_pt.emplace_back(10 * u(gen), 10 * u(gen), 20 + 10 * u(gen));
  int obs = 0;
  std::vector<std::array<int, 2>> pairs;
  // camera-major list of (camera, point) with |camera - centre(point)| <= per_pt / 2 on the ring of cameras, points in
  // ascending order per camera; built from per-centre buckets (O(observations), not O(cameras x points))
  {
    std::vector<std::vector<int>> by_centre(cams);
    for (int j = 0; j < pts; j++) by_centre[(int)((long long)j * cams / pts)].push_back(j);
    const int half = per_pt / 2;
    for (int c = 0; c < cams; c++) {
      std::vector<int> mine;
      for (int dd = -half; dd <= half; dd++) {
        const int centre = ((c + dd) % cams + cams) % cams;
        int d = std::abs(c - centre);
        d = std::min(d, cams - d);
        if (d > half) continue;  // (tiny rings: the same centre must not be taken twice)
        bool dup = false;
        for (int e = -half; e < dd; e++) dup = dup || (((c + e) % cams + cams) % cams) == centre;
        if (!dup) mine.insert(mine.end(), by_centre[centre].begin(), by_centre[centre].end());
      }
      std::sort(mine.begin(), mine.end());
      for (int j : mine) pairs.push_back({c, j});
    }
  }
  for (auto& cj : pairs) {
    const int c = cj[0], j = cj[1];
    const Eigen::Vector3d pc = truth_pose[c].Rotation().Rotate(truth_pt[j]);
    const double X = pc[0] + truth_pose[c].Data()[4], Y = pc[1] + truth_pose[c].Data()[5], Z = pc[2] + truth_pose[c].Data()[6];
    const double px = -X / Z, py = -Y / Z, r2 = px * px + py * py;
    const double r = 1 + truth_intr[c][1] * r2 + truth_intr[c][2] * r2 * r2, f = truth_intr[c][0];
    p.factors.push_back(MakeFactor(c, j, obs));
    p.values.Set(PIXEL.WithSuper(obs), Eigen::Vector2d(f * r * px + 0.5 * n(gen), f * r * py + 0.5 * n(gen)));
    ++obs;
  }
  p.num_observations = obs;
  for (int i = 0; i < cams; i++) {
    p.values.Set(CAM_T_WORLD.WithSuper(i),
                 truth_pose[i].Retract(sym::Vector6d(0.01 * n(gen), 0.01 * n(gen), 0.01 * n(gen), 0.05 * n(gen),
                                                     0.05 * n(gen), 0.05 * n(gen))));
    p.values.Set(INTRINSICS.WithSuper(i), truth_intr[i]);
  }
  for (int j = 0; j < pts; j++)
    p.values.Set(POINT.WithSuper(j), Eigen::Vector3d(truth_pt[j][0] + 0.1 * n(gen), truth_pt[j][1] + 0.1 * n(gen),
                                                     truth_pt[j][2] + 0.1 * n(gen)));
  p.values.Set(EPSILON, sym::kDefaultEpsilond);
  return p;
}

int main(int argc, char** argv) {
  const auto t_start = std::chrono::steady_clock::now();
  auto seconds_since = [](std::chrono::steady_clock::time_point t0) {
    return std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
  };
  Problem problem;
  if (argc == 5 && std::string(argv[1]) == "--synthetic")
    problem = SyntheticProblem(std::atoi(argv[2]), std::atoi(argv[3]), std::atoi(argv[4]));
  else if (argc == 2)
    problem = ReadProblem(argv[1]);
  else {
    std::fprintf(stderr, "usage: %s <problem.txt> | --synthetic <cams> <pts> <obs/pt>\n", argv[0]);
    return 2;
  }
  std::printf("Created problem with %d cameras, %d points, %d observations\n", problem.num_cameras, problem.num_points,
              problem.num_observations);
  std::fflush(stdout);
  sym::Valuesd optimized_values = problem.values;
  auto params = sym::DefaultOptimizerParams();
  params.lambda_update_type = sym::lambda_update_type_t::DYNAMIC;
  // keys c.., i.., p.. (lexical): the trailing points are eliminated by the GPU Schur path (AUTO)
  const double build_s = seconds_since(t_start);
  std::fprintf(stderr, "[host] factor list + Values built in %.2f s\n", build_s);
  const auto t_ctor = std::chrono::steady_clock::now();
  sym::Optimizerd optimizer{params, std::move(problem.factors)};
  std::fprintf(stderr, "[host] Optimizer constructed (keys to optimize collected and ordered) in %.2f s\n", seconds_since(t_ctor));
  const auto t_opt = std::chrono::steady_clock::now();
  const auto stats = optimizer.Optimize(optimized_values);
  const double optimize_s = seconds_since(t_opt);
  sfx_timings tm{};
  sfx_get_timings(optimizer.Handle(), &tm);
  // host-side ingestion (factor / Values construction), first-call initialisation (indexing, structural analysis,
  // METIS, symbolic factorization, upload) and the device time of the LM iterations themselves
  std::printf("Timing: build problem %.2f s, Optimize() wall %.2f s (first call includes Initialize), device %.1f ms for %d "
              "iterations = %.2f ms/iteration\n",
              build_s, optimize_s, tm.total_ms, tm.iterations_run, tm.iterations_run ? tm.total_ms / tm.iterations_run : 0.0);
  for (const auto& it : stats.iterations)
    std::printf("[iter %4d] lambda: %.3e, error: %.9e, rel reduction: %.5e, accepted: %d\n", it.iteration,
                it.current_lambda, it.new_error, it.relative_reduction, (int)it.update_accepted);
  std::printf("Finished in %zu iterations, status %d, best error %.9e\n", stats.iterations.size(), (int)stats.status,
              stats.iterations[stats.best_index].new_error);
  return stats.status == sym::optimization_status_t::SUCCESS ? 0 : 1;
}
