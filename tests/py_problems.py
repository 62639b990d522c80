"""Problems of the reference's Python tests expressed with symforce_b200.opt (shared by the CPU and GPU tests)."""
import os
import sys

import numpy as np

from symforce_b200.geo import K_DEFAULT_EPSILON
from symforce_b200.opt import Factor, Optimizer, Pose3, Rot3, Values, residuals

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "examples", "python"))
import robot_3d_localization as robot3d  # noqa: E402,F401


def rotation_smoothing(num_samples=10, **kwargs):
    """
    test/symforce_py_optimizer_test.py:37-83.  The reference's two lambdas are x.local_coordinates(y) and
    x.local_coordinates(x_prior); the device kinds are the generated between / prior factors
    (symforce/codegen/geo_factors_codegen.py), which compute the same residuals with a_T_b = identity and
    sqrt_info = I (the prior with the opposite sign: same error, Hessian and rhs).
    """
    xs = [f"x{i}" for i in range(num_samples)]
    x_priors = [f"x_prior{i}" for i in range(num_samples)]
    factors = []
    for i in range(num_samples - 1):
        factors.append(Factor(keys=[xs[i], xs[i + 1], "identity", "sqrt_info", "epsilon"],
                              residual=residuals.between_factor_rot3))
    for i in range(num_samples):
        factors.append(Factor(keys=[xs[i], x_priors[i], "sqrt_info", "epsilon"], name="prior",
                              residual=residuals.prior_factor_rot3))
    optimizer = Optimizer(factors=factors, optimized_keys=xs, **kwargs)
    initial_values = Values(epsilon=K_DEFAULT_EPSILON, identity=Rot3.identity(), sqrt_info=np.eye(3))
    for i in range(num_samples):
        initial_values[xs[i]] = Rot3.from_yaw_pitch_roll(yaw=0.0, pitch=0.1 * i, roll=0.0)
    for i in range(num_samples):
        initial_values[x_priors[i]] = Rot3.from_yaw_pitch_roll(roll=0.1 * i)
    return optimizer, initial_values


def bal_front(flat):
    """A flat BAL problem (problems.bal_problem / read_bal) re-expressed with the front: Values c[j], i[j], p[k], P[n], e
    and one Snavely factor per observation, keys as the reference example orders them
    (bundle_adjustment_in_the_large.cc:61-118).  Returns (values, factors, optimized_keys)."""
    m = flat.meta
    nc, npt = m["n_cams"], m["n_pts"]
    v = flat.values
    values = Values()
    values["c"] = [Pose3.from_storage(v[o:o + 7]) for o in flat.keys[:nc, 1]]
    values["i"] = [v[o:o + 3].copy() for o in flat.keys[nc:2 * nc, 1]]
    values["p"] = [v[o:o + 3].copy() for o in flat.keys[2 * nc:, 1]]
    values["P"] = [v[o:o + 2].copy() for o in flat.batches[0][1][3]]
    values["e"] = K_DEFAULT_EPSILON
    factors = [Factor(keys=[f"c[{c}]", f"i[{c}]", f"p[{p}]", f"P[{n}]", "e"], residual=residuals.snavely)
               for n, (c, p) in enumerate(zip(m["cam"], m["pt"]))]
    keys = [f"c[{j}]" for j in range(nc)] + [f"i[{j}]" for j in range(nc)] + [f"p[{k}]" for k in range(npt)]
    return values, factors, keys
