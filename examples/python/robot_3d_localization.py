"""
The reference's Python robot_3d_localization example (symforce/examples/robot_3d_localization/
robot_3d_localization.py:28-260) on the GPU path: same Values keys, same factor keys, same optimizer
parameters; the two residual functions are the device kinds generated from the reference's symbolic
definitions (robot_3d_localization.py:119-150) instead of Python callables.

    python examples/python/robot_3d_localization.py
"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))

from symforce_b200.opt import Factor, Optimizer, Pose3, Values, residuals  # noqa: E402
from symforce_b200.geo import K_DEFAULT_EPSILON  # noqa: E402

NUM_POSES = 5
NUM_LANDMARKS = 20


def build_values(num_poses):
    """robot_3d_localization.py:28-103 (np.random.seed(42): the data of gen/measurements.cc)."""
    np.random.seed(42)
    values = Values()

    gt_world_T_body = []
    for i in range(num_poses):
        t = i / num_poses
        tangent_vec = np.array([-1 * t, -2 * t, -3 * t, 8 * np.sin(t * np.pi / 1.3), 9 * np.sin(t * np.pi / 2),
                                5 * np.sin(t * np.pi / 1.8)])
        gt_world_T_body.append(Pose3.from_tangent(tangent_vec, epsilon=K_DEFAULT_EPSILON))

    values["world_T_body"] = [Pose3.identity() for _ in range(num_poses)]
    values["world_t_landmark"] = [np.random.uniform(low=0.0, high=10.0, size=3) for _ in range(NUM_LANDMARKS)]
    num_landmarks = len(values["world_t_landmark"])

    values["odometry_diagonal_sigmas"] = np.array([0.05, 0.05, 0.05, 0.2, 0.2, 0.2])
    values["odometry_relative_pose_measurements"] = []
    for i in range(num_poses - 1):
        gt_relative_pose = gt_world_T_body[i].inverse() * gt_world_T_body[i + 1]
        tangent_perturbation = np.random.normal(size=6) * values["odometry_diagonal_sigmas"]
        values["odometry_relative_pose_measurements"].append(
            gt_relative_pose.retract(tangent_perturbation, epsilon=K_DEFAULT_EPSILON))

    values["matching_sigma"] = 0.1
    meas = np.zeros((num_poses, num_landmarks, 3))
    for i in range(num_poses):
        for j in range(num_landmarks):
            gt_body_t_landmark = gt_world_T_body[i].inverse() * values["world_t_landmark"][j]
            meas[i, j, :] = gt_body_t_landmark + np.random.normal(scale=values["matching_sigma"], size=3)
    values["body_t_landmark_measurements"] = [list(m) for m in meas]

    values["epsilon"] = K_DEFAULT_EPSILON
    return values, num_landmarks


def build_factors(num_poses, num_landmarks):
    """robot_3d_localization.py:159-186"""
    for i in range(num_poses):
        for j in range(num_landmarks):
            yield Factor(
                residual=residuals.matching_residual,
                keys=[f"world_T_body[{i}]", f"world_t_landmark[{j}]", f"body_t_landmark_measurements[{i}][{j}]",
                      "matching_sigma"],
            )
    for i in range(num_poses - 1):
        yield Factor(
            residual=residuals.odometry_residual,
            keys=[f"world_T_body[{i}]", f"world_T_body[{i + 1}]", f"odometry_relative_pose_measurements[{i}]",
                  "odometry_diagonal_sigmas", "epsilon"],
        )


def make_optimizer(num_poses=NUM_POSES, num_landmarks=NUM_LANDMARKS, **kwargs):
    """robot_3d_localization.py:229-243"""
    return Optimizer(
        factors=build_factors(num_poses, num_landmarks),
        optimized_keys=[f"world_T_body[{i}]" for i in range(num_poses)],
        params=Optimizer.Params(verbose=False, initial_lambda=1e4, lambda_down_factor=1 / 2.0, debug_stats=True),
        **kwargs,
    )


def main():
    values, num_landmarks = build_values(NUM_POSES)
    optimizer = make_optimizer(NUM_POSES, num_landmarks)
    result = optimizer.optimize(values)
    print(f"Num iterations: {len(result.iterations) - 1}")
    print(f"Final error: {result.error():.6f}")
    print(f"Status: {result.status.name}")
    for i, pose in enumerate(result.optimized_values["world_T_body"]):
        print(f"Pose {i}: t = {pose.position()}, heading = {pose.rotation().to_tangent()}")
    cov = optimizer.compute_all_covariances(result.optimized_values)
    print("trace of the pose covariances:", [float(np.trace(cov[f"world_T_body[{i}]"])) for i in range(NUM_POSES)])


if __name__ == "__main__":
    main()
