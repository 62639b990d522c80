"""
Pins the generated factor arithmetic (oracle/gen/factors_gen.h == symforce_b200/csrc/gen/factors_gen.cuh,
same DAG) against the REFERENCE's own generated headers compiled in place (oracle/_ref), and against
golden vectors recorded from that comparison (tests/golden/factor_vectors.json) so the check also
runs where /root/reference is absent.
"""
import json
import os

import numpy as np
import pytest

from symforce_b200 import desc as D
from tests import oracle_capi as O

GOLDEN = os.path.join(os.path.dirname(__file__), "golden", "factor_vectors.json")


def random_args(kind, rng):
    meta = D.KINDS[kind]
    args = []
    for name, dim in zip(meta["arg_names"], meta["arg_dims"]):
        if name == "epsilon":
            a = np.array([D.K_DEFAULT_EPSILON])
        elif name == "eps":  # barron: kEpsilon of test/symforce_gnc_test.cc:23
            a = np.array([1e-12])
        elif name == "mu":
            a = np.array([rng.uniform(0.0, 0.99)])
        elif dim == 7:  # Pose3
            q = rng.normal(size=4)
            q /= np.linalg.norm(q)
            a = np.concatenate([q, rng.normal(size=3) * 2])
        elif dim == 4 and "calibration" not in name:  # Rot3
            a = rng.normal(size=4)
            a /= np.linalg.norm(a)
        elif "calibration" in name:
            a = np.array([740.0, 740.0, 639.5, 359.5]) + rng.normal(size=4)
        elif name in ("sigma", "weight", "gnc_scale"):
            a = np.array([rng.uniform(0.5, 2.0)])
        elif name == "gnc_mu":
            a = np.array([rng.uniform(0.0, 0.9)])
        elif name == "diagonal_sigmas":
            a = rng.uniform(0.05, 0.5, size=dim)
        elif name in ("source_inverse_range", "landmark_inverse_range", "inverse_range_prior"):
            a = np.array([rng.uniform(0.05, 0.4)])
        elif "pixel" in name and kind == D.KIND_IRL_LINEAR_GNC:
            a = np.array([rng.uniform(100, 1100), rng.uniform(100, 600)])
        elif name == "intrinsics":
            a = np.array([rng.uniform(500, 1500), rng.normal() * 1e-3, rng.normal() * 1e-5])
        elif name == "point":
            a = np.array([rng.normal(), rng.normal(), rng.uniform(-30, -10)])  # in front (camera looks down -z)
        else:
            a = rng.normal(size=dim)
        args.append(a)
    if kind == D.KIND_SNAVELY:
        args[0] = np.array([0, 0, 0, 1.0, 0, 0, 0]) + np.concatenate([rng.normal(size=3) * 0.05, [0], rng.normal(size=3) * 0.3])
        args[0][:4] /= np.linalg.norm(args[0][:4])
    if kind == D.KIND_IRL_LINEAR_GNC:
        # keep both cameras roughly facing +z with small relative motion so the warp is valid
        for i in (0, 2):
            q = np.array([0, 0, 0, 1.0]) + np.concatenate([rng.normal(size=3) * 0.03, [0]])
            args[i] = np.concatenate([q / np.linalg.norm(q), rng.normal(size=3) * 0.2])
    return args


def rel_err(a, b):
    scale = max(np.max(np.abs(a)), np.max(np.abs(b)), 1e-300)
    return np.max(np.abs(a - b)) / scale


@pytest.mark.parametrize("kind", range(len(D.KINDS)))
def test_oracle_factor_matches_reference_headers(kind):
    ref = O.load_ref()
    if ref is None:
        pytest.skip("oracle/_ref not built (reference absent)")
    lib = O.load()
    rng = np.random.default_rng(1234 + kind)
    for _ in range(200):
        args = random_args(kind, rng)
        r0, J0, H0, g0 = O.eval_factor(ref, "ref_eval_factor", kind, args)
        r1, J1, H1, g1 = O.eval_factor(lib, "orc_eval_factor", kind, args)
        assert rel_err(r0, r1) < 1e-12
        assert rel_err(J0, J1) < 1e-11
        # H = J^T J and rhs = J^T r are compared against their natural scales |J|^2 and |J||r|
        # (entries can cancel far below that, e.g. for a saturated robust loss)
        jn, rn = np.linalg.norm(J0), np.linalg.norm(r0)
        assert np.max(np.abs(np.tril(H0) - np.tril(H1))) < 1e-11 * jn * jn + 1e-300
        assert np.max(np.abs(g0 - g1)) < 1e-11 * jn * rn + 1e-300
        # the reference writes only the lower triangle of H (upper left as zero)
        assert np.all(np.triu(H0, 1) == 0)


def test_oracle_factor_matches_golden_vectors():
    """Golden vectors were produced by ref_eval_factor (tests/golden/make_factor_vectors.py)."""
    lib = O.load()
    with open(GOLDEN) as f:
        g = json.load(f)
    for entry in g["vectors"]:
        kind = entry["kind"]
        args = [np.array(a) for a in entry["args"]]
        r1, J1, H1, g1 = O.eval_factor(lib, "orc_eval_factor", kind, args)
        T, R = D.KINDS[kind]["tan_dim"], D.KINDS[kind]["res_dim"]
        J0 = np.array(entry["J"]).reshape(R, T)
        r0 = np.array(entry["res"])
        jn, rn = np.linalg.norm(J0), np.linalg.norm(r0)
        assert rel_err(r0, r1) < 1e-12
        assert rel_err(J0, J1) < 1e-11
        assert np.max(np.abs(np.array(entry["H"]).reshape(T, T) - np.tril(H1))) < 1e-11 * jn * jn + 1e-300
        assert np.max(np.abs(np.array(entry["rhs"]) - g1)) < 1e-11 * jn * rn + 1e-300
