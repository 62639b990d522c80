"""
GPU tests of the Python front (symforce_b200/opt.py) written after the reference's Python tests:
test/symforce_examples_robot_3d_localization_test.py:17-51 and test/symforce_py_optimizer_test.py:37-118, plus parity of
what the front returns (iterations, optimized Values, linearization, covariances) with the CPU oracle on the problem
the front lowers to.  Tolerances as in test_gpu_parity.py: 1e-9 relative on H / rhs, 1e-8 on errors and covariances.
"""
import numpy as np
import pytest

from oracle import covariance_ref as R
from symforce_b200 import capi, desc as D, problems as P
from symforce_b200.opt import Optimizer, Pose3
from tests import oracle_capi as O
from tests import py_problems as PP

pytestmark = pytest.mark.gpu


def test_robot_3d_localization_python_example():
    values, num_landmarks = PP.robot3d.build_values(PP.robot3d.NUM_POSES)
    optimizer = PP.robot3d.make_optimizer(PP.robot3d.NUM_POSES, num_landmarks)
    result = optimizer.optimize(values)

    # the reference's assertions
    assert abs(result.iterations[0].new_error - 463700.5576620833) < 1e-7
    assert result.error() < 140
    assert result.status == Optimizer.Status.SUCCESS

    # parity with the oracle on the lowered problem
    o = O.OracleProblem(optimizer.problem(values))
    st = o.optimize()
    its = o.iterations()
    assert len(result.iterations) == len(its) and result.best_index == st.best_index
    for a, b in zip(result.iterations, its):
        assert a.iteration == b.iteration and a.update_accepted == bool(b.update_accepted)
        assert abs(a.new_error - b.new_error) <= 1e-8 * abs(b.new_error)
    best = o.best_values()
    got = np.array(result.optimized_values.to_storage())
    # Python Values order == storage order here (optimized keys come first in both)
    assert np.allclose(got, best, rtol=0, atol=1e-9)
    # the input Values are not modified; the result holds new objects of the same types
    assert values["world_T_body"][0].to_storage() == Pose3.identity().to_storage()
    assert isinstance(result.optimized_values["world_T_body"][3], Pose3)
    assert result.optimized_values["matching_sigma"] == 0.1

    # debug_stats payloads: the best iteration's values are the optimized values
    rec = result.iterations[result.best_index]
    assert rec.values is not None and rec.residual is not None
    loaded = optimizer.load_iteration_values(rec.values)
    assert np.allclose(loaded.to_storage(), got, rtol=0, atol=0)
    assert abs(0.5 * rec.residual @ rec.residual - rec.new_error) <= 1e-9 * rec.new_error
    assert sorted(result.linear_solver_ordering.tolist()) == list(range(30))
    assert rec.update.shape == (30,) and result.iterations[0].update.shape == (0,)
    assert rec.jacobian_values is None and result.jacobian_sparsity.shape == ()  # include_jacobians is off
    with pytest.raises(ValueError, match="include_jacobians"):
        result.jacobian_view(rec)

    # debug_stats + include_jacobians: OptimizationStats::JacobianView of a record (optimization_stats.h:67-75) is the
    # Jacobian Optimizer::Linearize yields at that record's values
    optimizer.params.include_jacobians = True
    with_j = optimizer.optimize(values)
    rec_j = with_j.iterations[with_j.best_index]
    J = with_j.jacobian_view(rec_j)
    assert J.shape == (rec_j.residual.shape[0], 30) and with_j.jacobian_sparsity.shape == J.shape
    J_lin = optimizer.linearize(with_j.optimized_values).jacobian
    assert (J != J_lin).nnz == 0
    # rhs = J^T r of the same record
    lin_best = optimizer.linearize(with_j.optimized_values)
    assert np.allclose(J.T @ rec_j.residual, lin_best.rhs, rtol=0, atol=1e-9 * np.abs(lin_best.rhs).max())
    # check_derivatives (optimizer.tcc:261-272): asserted at the initial values and at every record's values
    ok, err = optimizer.check_derivatives(values)
    assert ok and err["jacobian"] < 1e-6 and err["hessian"] < 1e-12
    optimizer.params.check_derivatives = True
    checked = optimizer.optimize(values)
    assert len(checked.iterations) == len(result.iterations)
    optimizer.params.check_derivatives = False
    optimizer.params.include_jacobians = False

    # a second optimize on the same optimizer starts over from the given Values
    again = optimizer.optimize(values)
    assert len(again.iterations) == len(result.iterations)
    assert abs(again.error() - result.error()) <= 1e-12 * result.error()  # atomic accumulation order varies run to run
    optimizer.close()


def test_rotation_smoothing_python_kat():
    optimizer, initial_values = PP.rotation_smoothing()
    result = optimizer.optimize(initial_values)
    assert len(result.iterations) == 7
    assert round(result.error() - 0.039, 3) == 0
    assert result.status == Optimizer.Status.SUCCESS
    assert result.failure_reason == Optimizer.FailureReason.INVALID
    optimizer.close()


def test_linearize_and_covariances_through_the_front():
    values, num_landmarks = PP.robot3d.build_values(PP.robot3d.NUM_POSES)
    optimizer = PP.robot3d.make_optimizer(PP.robot3d.NUM_POSES, num_landmarks)
    result = optimizer.optimize(values, populate_best_linearization=True)
    lin = optimizer.linearize(result.optimized_values)
    o = O.OracleProblem(optimizer.problem(result.optimized_values))
    res, rhs, H = o.linearize()
    assert np.allclose(lin.residual, res, rtol=0, atol=1e-9 * np.abs(res).max())
    # at the optimum rhs = J^T r cancels to ~1e-5 from terms of order 1e2: the scale of the sum is the terms'
    rhs_scale = max(1.0, np.abs(rhs).max())
    assert np.allclose(lin.rhs, rhs, rtol=0, atol=1e-9 * rhs_scale)
    assert np.allclose(lin.hessian_lower.data, H, rtol=0, atol=1e-9 * np.abs(H).max())
    assert abs(lin.error() - result.error()) <= 1e-9 * result.error()
    assert np.allclose(result.best_linearization.rhs, lin.rhs, rtol=0, atol=1e-9 * rhs_scale)
    # linear_error(0) is the error; a small step along -H^-1 rhs lowers the linear model
    assert lin.linear_error(np.zeros(30)) == lin.error()

    Hd = lin.hessian_lower.toarray()
    Hd = Hd + Hd.T - np.diag(np.diag(Hd))
    want = R.full_covariance(Hd, optimizer.epsilon)
    full = optimizer.compute_full_covariance(result.optimized_values)
    assert np.max(np.abs(full - want)) <= 1e-8 * np.max(np.abs(want))
    by_key = optimizer.compute_all_covariances(result.optimized_values)
    assert list(by_key) == optimizer.optimized_keys
    for k, e in optimizer.linearization_index().items():
        blk = want[e.offset:e.offset + e.tangent_dim, e.offset:e.offset + e.tangent_dim]
        assert by_key[k].shape == (6, 6) and np.max(np.abs(by_key[k] - blk)) <= 1e-8 * np.max(np.abs(want))
    # compute_covariances with every key is compute_all_covariances; a strict prefix needs a Schur-eliminable tail
    allk = optimizer.compute_covariances(result.optimized_values, optimizer.optimized_keys)
    # two separate linearizations on the device: the Hessian is accumulated with fp64 atomics, whose order (and so
    # the last bits) differs from run to run -- same wiring, not the same bits
    for k in by_key:
        assert np.max(np.abs(allk[k] - by_key[k])) <= 1e-11 * np.max(np.abs(want)), k
    with pytest.raises(ValueError, match="first optimized keys"):
        optimizer.compute_covariances(result.optimized_values, optimizer.optimized_keys[1:3])
    # a strict prefix whose tail is NOT Schur-eliminable (6-dim poses tied by odometry): the general-C branch
    # (c_is_block_diagonal = False, covariance_utils.h:41-103) = the leading block of the inverse damped on C only
    first_two = optimizer.compute_covariances(result.optimized_values, optimizer.optimized_keys[:2],
                                              c_is_block_diagonal=False)
    want_two = R.covariance_block(Hd, 12, optimizer.epsilon, c_is_block_diagonal=False)
    assert list(first_two) == optimizer.optimized_keys[:2]
    for i, k in enumerate(first_two):
        blk = want_two[6 * i:6 * i + 6, 6 * i:6 * i + 6]
        assert np.max(np.abs(first_two[k] - blk)) <= 1e-8 * np.max(np.abs(want_two)), k
    optimizer.close()


def test_bal_through_the_front_uses_the_schur_path():
    """BAL keys c, i, p through the front: the trailing points are eliminated (solver 'auto'), and the run is the one
    the flat problem gives through the C ABI."""
    flat = P.bal_problem("small", solver=D.SOLVER_SCHUR)
    values, factors, keys = PP.bal_front(flat)
    optimizer = Optimizer(factors, keys, params=Optimizer.Params(lambda_update_type=Optimizer.Params().lambda_update_type.DYNAMIC))
    result = optimizer.optimize(values)
    assert optimizer.problem(values).solver == D.SOLVER_SCHUR
    g = capi.SfxProblem(flat)
    st = g.optimize()
    its = g.iterations()
    assert len(result.iterations) == len(its) and result.best_index == st.best_index
    assert int(result.status) == st.status
    for a, b in zip(result.iterations, its):
        assert a.update_accepted == bool(b.update_accepted)
        assert abs(a.new_error - b.new_error) <= 1e-8 * abs(b.new_error)
    assert result.error() < 0.5 * result.iterations[0].new_error
    pts = np.array([x for x in result.optimized_values["p"]])
    assert pts.shape == (flat.meta["n_pts"], 3) and np.isfinite(pts).all()
    g.close()
    optimizer.close()
