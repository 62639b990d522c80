"""
GPU test of the Jacobian export (optimizer_params_t::include_jacobians; Linearization::jacobian of
symforce/opt/linearization.h:58-60, built by linearizer.cc:252-259, 297-313): sfx_get_jacobian_pattern /
sfx_linearize_jacobian through the C ABI against the oracle's triplet-built CSC.  Pattern bit-exact; values within 1e-9
relative (same bar as H and rhs); J^T J and J^T r reproduce the device's own Hessian and rhs.
"""
import numpy as np
import pytest
import scipy.sparse as sp

from symforce_b200 import capi, desc as D, problems as P
from tests import oracle_capi as O

pytestmark = pytest.mark.gpu

PROBLEMS = {
    "robot3d": lambda: P.robot_3d_localization(),
    "ba_example": lambda: P.ba_example(),
    "frozen_keys": lambda: P.frozen_keys(),
    "gnc_test": lambda: P.gnc_test(),
    "bal_small_schur": lambda: P.bal_problem("small", solver=D.SOLVER_SCHUR),
    "bal_small_chol": lambda: P.bal_problem("small", solver=D.SOLVER_CHOLESKY),
    "pose_graph": lambda: P.pose_graph_problem(n_poses=200, n_loops=40),
}


@pytest.mark.parametrize("name", list(PROBLEMS))
def test_jacobian_matches_oracle(name):
    prob = PROBLEMS[name]()
    g, o = capi.SfxProblem(prob), O.OracleProblem(prob)
    outer, inner, val = g.jacobian()
    o_outer, o_inner, o_val = o.jacobian()
    assert np.array_equal(outer, o_outer) and np.array_equal(inner, o_inner)
    assert np.max(np.abs(val - o_val)) <= 1e-9 * np.max(np.abs(o_val))
    N, M, _ = g.dims()
    J = sp.csc_matrix((val, inner, outer), shape=(M, N))
    res, rhs, Hv = g.linearize()
    ho, hi = g.hessian_pattern()
    H = sp.csc_matrix((Hv, hi, ho), shape=(N, N)).toarray()
    assert np.allclose(np.tril((J.T @ J).toarray()), H, rtol=0, atol=1e-9 * np.abs(H).max())
    assert np.allclose(J.T @ res, rhs, rtol=0, atol=1e-9 * max(1.0, np.abs(rhs).max()))
    g.close()


def test_jacobian_export_leaves_the_optimizer_state_alone():
    prob = P.robot_3d_localization()
    g = capi.SfxProblem(prob)
    st = g.optimize()
    best = g.best_values()
    _, _, H0 = g.best_linearization()
    g.set_values(best)
    _, _, val = g.jacobian()
    assert np.isfinite(val).all()
    assert np.array_equal(g.best_values(), best)
    _, _, H1 = g.best_linearization()
    assert np.array_equal(H0, H1)
    assert len(g.iterations()) == st.n_iterations
    g.close()


def test_python_front_fills_the_jacobian():
    from tests import py_problems as PP

    values, num_landmarks = PP.robot3d.build_values(PP.robot3d.NUM_POSES)
    optimizer = PP.robot3d.make_optimizer(PP.robot3d.NUM_POSES, num_landmarks)
    lin = optimizer.linearize(values)
    with pytest.raises(ValueError, match="include_jacobians"):
        lin.jacobian
    optimizer.params.include_jacobians = True
    lin = optimizer.linearize(values)
    J = lin.jacobian
    assert J.shape == (lin.residual.shape[0], 30)
    assert np.allclose(J.T @ lin.residual, lin.rhs, rtol=0, atol=1e-9 * np.abs(lin.rhs).max())
    optimizer.close()
