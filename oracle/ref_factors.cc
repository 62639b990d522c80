/*
 * ref_factors.cc -- compiles the REFERENCE's own generated factor headers, where they lie under
 * /root/reference, behind a C ABI, so tests can pin oracle/gen/factors_gen.h (and through it the
 * CUDA kernels) to the reference's arithmetic.  TEST INFRASTRUCTURE ONLY; built into
 * oracle/_ref/libref_factors.so (git-ignored) by oracle/Makefile; only buildable where
 * /root/reference exists.  The few Eigen / sym types those headers need are provided by
 * oracle/ref_shim (Eigen is not installed in this image).
 */
#include <Eigen/Core>
#include <sym/linear_camera_cal.h>
#include <sym/pose3.h>
#include <sym/rot3.h>

#include <snavely_reprojection_factor.h>               // symforce/examples/bundle_adjustment_in_the_large/gen
#include <matching_factor.h>                           // symforce/examples/robot_3d_localization/gen
#include <odometry_factor.h>
#include <sym/factors/between_factor_pose3.h>          // gen/cpp
#include <sym/factors/between_factor_rot3.h>
#include <sym/factors/inverse_range_landmark_linear_gnc_factor.h>
#include <sym/factors/inverse_range_landmark_prior_factor.h>
#include <sym/factors/prior_factor_pose3.h>
#include <sym/factors/prior_factor_rot3.h>
#include <symforce/gnc_factors/barron_factor.h>        // test/symforce_function_codegen_test_data/symengine/gnc_test_data/cpp

template <int R, int C>
using M = Eigen::Matrix<double, R, C>;

template <int R, int T>
static void out(const M<R, 1>& res, const M<R, T>& J, const M<T, T>& H, const M<T, 1>& rhs, double* res_o,
                double* J_o, double* H_o, double* rhs_o) {
  for (int i = 0; i < R; ++i) res_o[i] = res[i];
  for (int i = 0; i < R * T; ++i) J_o[i] = J.data()[i];
  for (int i = 0; i < T * T; ++i) H_o[i] = H.data()[i];
  for (int i = 0; i < T; ++i) rhs_o[i] = rhs[i];
}

extern "C" int ref_eval_factor(int kind, const double* const* a, double* res, double* J, double* H, double* rhs) {
  switch (kind) {
    case 0: {
      M<2, 1> r; M<2, 12> j; M<12, 12> h; M<12, 1> g;
      sym::SnavelyReprojectionFactor<double>(sym::Pose3<double>(a[0]), M<3, 1>(a[1]), M<3, 1>(a[2]), M<2, 1>(a[3]),
                                             a[4][0], &r, &j, &h, &g);
      out<2, 12>(r, j, h, g, res, J, H, rhs);
      return 0;
    }
    case 1: {
      M<6, 1> r; M<6, 12> j; M<12, 12> h; M<12, 1> g;
      sym::BetweenFactorPose3<double>(sym::Pose3<double>(a[0]), sym::Pose3<double>(a[1]), sym::Pose3<double>(a[2]),
                                      M<6, 6>(a[3]), a[4][0], &r, &j, &h, &g);
      out<6, 12>(r, j, h, g, res, J, H, rhs);
      return 0;
    }
    case 2: {
      M<6, 1> r; M<6, 6> j; M<6, 6> h; M<6, 1> g;
      sym::PriorFactorPose3<double>(sym::Pose3<double>(a[0]), sym::Pose3<double>(a[1]), M<6, 6>(a[2]), a[3][0], &r,
                                    &j, &h, &g);
      out<6, 6>(r, j, h, g, res, J, H, rhs);
      return 0;
    }
    case 3: {
      M<3, 1> r; M<3, 6> j; M<6, 6> h; M<6, 1> g;
      sym::MatchingFactor<double>(sym::Pose3<double>(a[0]), M<3, 1>(a[1]), M<3, 1>(a[2]), a[3][0], &r, &j, &h, &g);
      out<3, 6>(r, j, h, g, res, J, H, rhs);
      return 0;
    }
    case 4: {
      M<6, 1> r; M<6, 12> j; M<12, 12> h; M<12, 1> g;
      sym::OdometryFactor<double>(sym::Pose3<double>(a[0]), sym::Pose3<double>(a[1]), sym::Pose3<double>(a[2]),
                                  M<6, 1>(a[3]), a[4][0], &r, &j, &h, &g);
      out<6, 12>(r, j, h, g, res, J, H, rhs);
      return 0;
    }
    case 5: {
      M<2, 1> r; M<2, 13> j; M<13, 13> h; M<13, 1> g;
      sym::InverseRangeLandmarkLinearGncFactor<double>(
          sym::Pose3<double>(a[0]), sym::LinearCameraCal<double>(a[1]), sym::Pose3<double>(a[2]),
          sym::LinearCameraCal<double>(a[3]), a[4][0], M<2, 1>(a[5]), M<2, 1>(a[6]), a[7][0], a[8][0], a[9][0],
          a[10][0], &r, &j, &h, &g);
      out<2, 13>(r, j, h, g, res, J, H, rhs);
      return 0;
    }
    case 6: {
      M<1, 1> r; M<1, 1> j; M<1, 1> h; M<1, 1> g;
      sym::InverseRangeLandmarkPriorFactor<double>(a[0][0], a[1][0], a[2][0], a[3][0], a[4][0], &r, &j, &h, &g);
      out<1, 1>(r, j, h, g, res, J, H, rhs);
      return 0;
    }
    case 7: {
      M<3, 1> r; M<3, 6> j; M<6, 6> h; M<6, 1> g;
      sym::BetweenFactorRot3<double>(sym::Rot3<double>(a[0]), sym::Rot3<double>(a[1]), sym::Rot3<double>(a[2]),
                                     M<3, 3>(a[3]), a[4][0], &r, &j, &h, &g);
      out<3, 6>(r, j, h, g, res, J, H, rhs);
      return 0;
    }
    case 8: {
      M<3, 1> r; M<3, 3> j; M<3, 3> h; M<3, 1> g;
      sym::PriorFactorRot3<double>(sym::Rot3<double>(a[0]), sym::Rot3<double>(a[1]), M<3, 3>(a[2]), a[3][0], &r, &j,
                                   &h, &g);
      out<3, 3>(r, j, h, g, res, J, H, rhs);
      return 0;
    }
    case 9: {
      M<5, 1> r; M<5, 5> j; M<5, 5> h; M<5, 1> g;
      gnc_factors::BarronFactor<double>(M<5, 1>(a[0]), M<5, 1>(a[1]), a[2][0], a[3][0], &r, &j, &h, &g);
      out<5, 5>(r, j, h, g, res, J, H, rhs);
      return 0;
    }
  }
  return 1;
}
