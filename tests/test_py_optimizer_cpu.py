"""
CPU tests of the Python front (symforce_b200/opt.py, geo.py): geometry against vectors generated from the
reference's own numeric package, Values flattening, the error behaviour of symforce/opt/optimizer.py, and the
lowering of the reference's Python examples -- checked by handing the lowered problem to the CPU oracle and
reading the known answers of test/symforce_examples_robot_3d_localization_test.py:49-51 and
test/symforce_py_optimizer_test.py:85-104.  Nothing here touches a GPU; Optimizer.optimize itself is covered by
tests/test_gpu_py_optimizer.py.
"""
import json
import os

import numpy as np
import pytest

from symforce_b200 import desc as D, problems as P
from symforce_b200.geo import K_DEFAULT_EPSILON, Pose3, Rot3
from symforce_b200.opt import Factor, Optimizer, Values, index_entry_t, residuals, type_t
from tests import oracle_capi as O
from tests import py_problems as PP

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def test_geo_matches_reference_vectors():
    with open(os.path.join(GOLDEN, "geo_vectors.json")) as f:
        g = json.load(f)
    eps = g["epsilon"]
    assert eps == K_DEFAULT_EPSILON
    for c in g["cases"]:
        a = Pose3.from_tangent(c["v"], eps)
        b = Pose3.from_tangent(c["w"], eps)
        tol = dict(rtol=0, atol=2e-14)
        assert np.allclose(a.to_storage(), c["a"], **tol)
        assert np.allclose(a.inverse().to_storage(), c["a_inv"], **tol)
        assert np.allclose((a * b).to_storage(), c["ab"], **tol)
        assert np.allclose(a.retract(c["w"], eps).to_storage(), c["a_retract_w"], **tol)
        assert np.allclose(a.local_coordinates(b, eps), c["a_local_b"], **tol)
        assert np.allclose(a.to_tangent(eps), c["a_tangent"], **tol)
        assert np.allclose(a * np.array(c["pt"]), c["a_pt"], **tol)
        assert np.allclose(Rot3.from_yaw_pitch_roll(*c["ypr"]).to_storage(), c["rot_ypr"], **tol)
        assert np.allclose(a.R.to_rotation_matrix().reshape(-1), c["rot_matrix"], **tol)
        assert np.allclose(a.R.local_coordinates(b.R, eps), c["rot_local"], **tol)
        q = Rot3.from_rotation_matrix(a.R.to_rotation_matrix()).data
        assert min(np.abs(q - a.R.data).max(), np.abs(q + a.R.data).max()) < 1e-14
        # group identities
        assert np.allclose((a * a.inverse()).to_storage(), Pose3.identity().to_storage(), atol=1e-14)
        b2 = np.array(a.retract(a.local_coordinates(b, eps), eps).to_storage())
        b2[:4] *= np.sign(b2[:4] @ np.array(c["b"][:4]))  # q and -q are the same rotation
        assert np.allclose(b2, c["b"], atol=1e-13)


def test_values_flattening_and_round_trip():
    v = Values(a=1.5, m=np.arange(6.0).reshape(2, 3))
    v["poses"] = [Pose3.identity(), Pose3.from_tangent([0.1, 0.2, 0.3, 1, 2, 3])]
    v["grid"] = [[np.array([1.0, 2.0]), np.array([3.0, 4.0])], [np.array([5.0, 6.0]), np.array([7.0, 8.0])]]
    v["sub"] = dict(r=Rot3.from_yaw_pitch_roll(0.3, 0.2, 0.1), s=2.0)
    assert v.keys_recursive() == ["a", "m", "poses[0]", "poses[1]", "grid[0][0]", "grid[0][1]", "grid[1][0]",
                                  "grid[1][1]", "sub.r", "sub.s"]
    assert v["grid[1][0]"].tolist() == [5.0, 6.0] and v["sub.s"] == 2.0 and "poses[1]" in v and "poses[2]" not in v
    st = v.to_storage()
    assert len(st) == 1 + 6 + 14 + 8 + 4 + 1
    assert st[1:7] == [0.0, 3.0, 1.0, 4.0, 2.0, 5.0]  # matrices are stored column-major like Eigen's
    v["poses[0]"] = Pose3.from_tangent([0, 0, 0, 9, 9, 9])
    assert v["poses"][0].t.tolist() == [9.0, 9.0, 9.0]
    with pytest.raises(TypeError):
        Values(x="a string").to_storage()


def test_constructor_errors_follow_the_reference():
    """test/symforce_py_optimizer_test.py:125-143 and the device-kind restrictions."""
    f1 = Factor(["present_key", "prior", "w", "sigma", "epsilon"], residuals.inverse_range_landmark_prior_factor)
    f2 = Factor(["absent_key", "prior", "w", "sigma", "epsilon"], residuals.inverse_range_landmark_prior_factor)
    optimized_keys = ["present_key", "other_present_key"]
    with pytest.raises(ValueError) as e:
        Optimizer([f1, f2], optimized_keys)
    assert str(f2.keys) in str(e.value) and str(optimized_keys) in str(e.value)
    with pytest.raises(ValueError, match="must specify `optimized_keys`"):
        Optimizer([f1])
    with pytest.raises(ValueError, match="not a device factor kind"):
        Factor(["x"], lambda x: x)
    with pytest.raises(ValueError, match="takes 4 keys"):
        Factor(["a", "b"], residuals.matching_residual)
    # the matching kind is differentiated with respect to the pose only
    with pytest.raises(ValueError, match="only differentiated"):
        Optimizer([Factor(["T", "l", "m", "s"], residuals.matching_residual)], optimized_keys=["T", "l"])
    with pytest.raises(AssertionError, match="Duplicates"):
        Optimizer([f1], ["present_key", "present_key"])
    # NumericFactor analogue: optimized keys collected from the factors (optimizer.py:206-213)
    o = Optimizer([Factor(["T", "l", "m", "s"], residuals.matching_residual, optimized_keys=["T"])])
    assert o.optimized_keys == ["T"]
    assert Optimizer([f1], ["present_key"]).params.verbose is True  # optimizer.py:216-219


def test_params_defaults_match_the_c_abi():
    p = Optimizer.Params().to_c()
    d = D.default_params()
    for name, _ in D.Params._fields_:
        assert getattr(p, name) == getattr(d, name), name
    q = Optimizer.Params(lambda_update_type=2, iterations=7, use_diagonal_damping=True, initial_lambda=1e4).to_c()
    assert (q.lambda_update_type, q.iterations, q.use_diagonal_damping, q.initial_lambda) == (2, 7, 1, 1e4)


def test_robot_3d_example_lowers_to_the_reference_problem():
    """examples/python/robot_3d_localization.py: the np.random.seed(42) data reproduces the reference's
    gen/measurements.cc fixture, the lowering equals the hand-built flat problem, and the oracle on it gives the
    known answers of test/symforce_examples_robot_3d_localization_test.py:49-51."""
    values, num_landmarks = PP.robot3d.build_values(PP.robot3d.NUM_POSES)
    optimizer = PP.robot3d.make_optimizer(PP.robot3d.NUM_POSES, num_landmarks)
    prob = optimizer.problem(values)
    flat = P.robot_3d_localization()
    assert np.array_equal(prob.keys, flat.keys)
    assert len(prob.batches) == len(flat.batches)
    for a, b in zip(prob.batches, flat.batches):
        assert a[0] == b[0] and all(np.array_equal(x, y) for x, y in zip(a[1:], b[1:]))
    assert np.allclose(prob.values, flat.values, rtol=0, atol=1e-11)
    assert prob.solver == D.SOLVER_CHOLESKY and prob.params.initial_lambda == 1e4
    o = O.OracleProblem(prob)
    st = o.optimize()
    its = o.iterations()
    assert abs(its[0].new_error - 463700.5576620833) < 1e-7  # assertAlmostEqual (7 places)
    assert its[st.best_index].new_error < 140
    assert st.status == Optimizer.Status.SUCCESS
    e = optimizer.linearization_index()["world_T_body[2]"]
    assert e == index_entry_t(key="world_T_body[2]", type=type_t.POSE3, offset=12, storage_dim=7, tangent_dim=6)


def test_rotation_smoothing_known_answers():
    """test/symforce_py_optimizer_test.py:85-111"""
    optimizer, initial_values = PP.rotation_smoothing()
    o = O.OracleProblem(optimizer.problem(initial_values))
    st = o.optimize()
    its = o.iterations()
    assert len(its) == 7
    assert round(its[st.best_index].new_error - 0.039, 3) == 0
    assert st.status == Optimizer.Status.SUCCESS
    assert Optimizer.FailureReason(st.failure_reason) == Optimizer.FailureReason.INVALID
    index_entry = optimizer.linearization_index()["x1"]
    assert index_entry == index_entry_t(key="x1", type=type_t.ROT3, offset=3, storage_dim=4, tangent_dim=3)
    assert optimizer.linearization_index_entry("x1") == index_entry


def test_bal_lowering_picks_the_schur_solver():
    """A BAL-shaped problem through the front: keys c_j, i_j, p_k as the reference example orders them
    (bundle_adjustment_in_the_large.cc:61-118) -> the trailing points become the Schur block."""
    flat = P.bal_problem("tiny", solver=D.SOLVER_SCHUR)
    npt = flat.meta["n_pts"]
    kind, ao, ok, fi = flat.batches[0]
    values, factors, keys = PP.bal_front(flat)
    opt = Optimizer(factors, keys, params=Optimizer.Params(lambda_update_type=2))
    prob = opt.problem(values)
    assert prob.solver == D.SOLVER_SCHUR and prob.schur_num_keys == npt
    assert np.array_equal(prob.keys[:, [0, 2, 3]], flat.keys[:, [0, 2, 3]])
    assert np.array_equal(prob.batches[0][2], ok)
    # same problem, different storage order: identical linearization in the oracle
    r1, g1, H1 = O.OracleProblem(prob).linearize()
    r0, g0, H0 = O.OracleProblem(flat).linearize()
    assert np.array_equal(r1, r0) and np.array_equal(g1, g0) and np.array_equal(H1, H0)
    assert Optimizer(factors, keys, solver="cholesky").problem(values).solver == D.SOLVER_CHOLESKY
