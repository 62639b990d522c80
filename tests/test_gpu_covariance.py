"""
GPU tests of the marginal-covariance path: sfx_compute_covariance through the C ABI
against the numpy restatement of the reference's covariance_utils.h (oracle/covariance_ref.py) on the SAME
linearization, exported from the device in the reference's CSC layout.  Tolerance: 1e-8 relative on the whole block
(the covariance is an inverse; H itself matches the oracle to 1e-9, see test_gpu_parity.py).
"""
import numpy as np
import pytest

from oracle import covariance_ref as R
from symforce_b200 import capi, desc as D, problems as P

pytestmark = pytest.mark.gpu

COV_TOL = 1e-8


def _dense_best(g):
    N, _, _ = g.dims()
    outer, inner = g.hessian_pattern()
    _, _, Hv = g.best_linearization()
    return R.dense_from_csc_lower(N, outer, inner, Hv), (outer, inner, Hv)


def relerr(a, b):
    return np.max(np.abs(a - b)) / np.max(np.abs(b))


def test_full_covariance_ba_example():
    # config A (solved with the Cholesky solver like the reference): ComputeAllCovariances / ComputeFullCovariance
    prob = P.ba_example()
    g = capi.SfxProblem(prob)
    g.optimize()
    H, (outer, inner, Hv) = _dense_best(g)
    N = H.shape[0]
    want = R.full_covariance(H, prob.epsilon)
    got_best = g.compute_covariance(N)                       # best linearization on the device
    got_host = g.compute_covariance(N, hessian_values=Hv)    # caller-provided Linearization::hessian_lower
    assert relerr(got_best, want) < COV_TOL
    assert relerr(got_host, want) < COV_TOL
    assert np.allclose(got_best, got_best.T, rtol=1e-9, atol=1e-14)
    # the optimizer state is untouched: the best linearization is still what it was
    _, _, Hv2 = g.best_linearization()
    assert np.array_equal(Hv, Hv2)
    g.close()


def test_full_covariance_robot3d():
    # config B: Cholesky problem, ComputeFullCovariance = (H + eps I)^-1
    prob = P.robot_3d_localization()
    g = capi.SfxProblem(prob)
    g.optimize()
    H, _ = _dense_best(g)
    N = H.shape[0]
    want = R.full_covariance(H, prob.epsilon)
    got = g.compute_covariance(N)
    assert relerr(got, want) < COV_TOL
    g.close()


@pytest.mark.parametrize("name, blocks", [("robot3d", (6, 12, 29)), ("pose_graph", (6, 60)), ("bal_small_chol", (9, 45))])
def test_general_c_covariance_block(name, blocks):
    """ComputeCovariances with c_is_block_diagonal = false (optimizer.tcc:177-199 -> covariance_utils.h:124-147 ->
    FromSparseC :41-103): a leading block of a problem solved without Schur elimination, C of any structure (6-dim poses
    that share factors; cameras and points).  Against the numpy restatement on the same exported linearization."""
    if name == "robot3d":
        prob = P.robot_3d_localization()
    elif name == "pose_graph":
        prob = P.pose_graph_problem(n_poses=40, n_loops=12)
    else:
        prob = P.bal_problem("small", solver=D.SOLVER_CHOLESKY)
    g = capi.SfxProblem(prob)
    g.optimize()
    N, _, _ = g.dims()
    H, (outer, inner, Hv_best) = _dense_best(g)
    Hv = Hv_best.copy()
    if name == "bal_small_chol":  # gauge freedom: a unit prior on every diagonal entry, as in the Schur test below
        for c in range(N):
            Hv[outer[c]] += 1.0
        H = H + np.eye(N)
    for b in blocks:
        want = R.covariance_block(H, b, prob.epsilon, c_is_block_diagonal=False)
        got = g.compute_covariance(b, hessian_values=Hv)
        assert got.shape == (b, b)
        e = relerr(got, want)
        print(f"GENERALC {name} block={b} of {N} relerr={e:.2e}")
        assert e < COV_TOL
        assert np.allclose(got, got.T, rtol=1e-8, atol=1e-14 * np.abs(got).max())
    # the whole system is still the damped inverse, and the optimizer state is untouched
    if name != "bal_small_chol":
        assert relerr(g.compute_covariance(N), R.full_covariance(H, prob.epsilon)) < COV_TOL
    _, _, Hv2 = g.best_linearization()
    assert np.array_equal(Hv2, Hv_best)
    g.close()


def test_schur_covariance_block_bal_fast_path():
    # BAL shape (9-dim cameras, 3-dim points: W / s9 kernels); the gauge freedom makes the undamped S singular, so
    # the linearization handed in is the exported one with a prior of weight 1 on every diagonal entry
    prob = P.bal_problem("small", solver=D.SOLVER_SCHUR)
    g = capi.SfxProblem(prob)
    g.optimize()
    N, _, _ = g.dims()
    H, (outer, inner, Hv) = _dense_best(g)
    Hv = Hv.copy()
    for c in range(N):
        assert inner[outer[c]] == c  # explicit diagonal first in every column (linearizer.cc:173-178)
        Hv[outer[c]] += 1.0
    H = H + np.eye(N)
    b = g.info()["reduced_dim"]
    want = R.covariance_block_schur(H, b, prob.epsilon)
    got = g.compute_covariance(b, hessian_values=Hv)
    assert relerr(got, want) < COV_TOL
    g.close()


def test_covariance_error_codes():
    prob = P.bal_problem("tiny", solver=D.SOLVER_SCHUR)
    g = capi.SfxProblem(prob)
    g.optimize()
    b = g.info()["reduced_dim"]
    with pytest.raises(RuntimeError, match="rc=3"):  # SFX_ERR_UNSUPPORTED: not the block the solver factors
        g.compute_covariance(b - 1)
    with pytest.raises(RuntimeError, match="rc=6"):  # SFX_ERR_NUMERICAL: gauge freedom, S is singular
        g.compute_covariance(b)
    # the problem is still usable afterwards
    g.set_values(prob.values)
    st = g.optimize()
    assert st.status in (1, 2)
    g.close()
