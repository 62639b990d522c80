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


@pytest.mark.parametrize("name", ["robot3d", "ba_example", "frozen_keys", "bal_small_schur", "pose_graph"])
def test_check_derivatives(name):
    """optimizer_params_t::check_derivatives (optimizer.tcc:261-272 -> internal/derivative_checker.h:32-123):
    sfx_check_derivatives compares the device's linearization with central differences of the device's residual (step
    sqrt(epsilon), tolerance 10 sqrt(epsilon)) and with J^T J / J^T r (tolerance sqrt(epsilon)).  The numerical Jacobian is
    also held against the ORACLE's analytic Jacobian, which ties the check to something the device did not compute."""
    prob = PROBLEMS[name]()
    g, o = capi.SfxProblem(prob), O.OracleProblem(prob)
    ok, err, num_j = g.check_derivatives(want_numerical_jacobian=True)
    print(f"CHECKDERIV {name} " + " ".join(f"{k}={v:.2e}" for k, v in err.items()))
    tol = np.sqrt(prob.epsilon)
    assert ok and err["jacobian"] <= 10 * tol and err["hessian"] <= tol and err["rhs"] <= tol
    assert err["hessian"] < 1e-12 and err["rhs"] < 1e-9  # these two are exact up to the summation order
    outer, inner, val = o.jacobian()
    N, M, _ = g.dims()
    J = sp.csc_matrix((val, inner, outer), shape=(M, N)).toarray()
    assert num_j.shape == (M, N)
    assert np.linalg.norm(num_j - J) <= 10 * tol * min(np.linalg.norm(J), np.linalg.norm(num_j))
    # like linearize(), the check leaves a reset optimizer: the next optimize starts over and matches a fresh problem
    st = g.optimize()
    g2 = capi.SfxProblem(prob)
    st2 = g2.optimize()
    assert st.n_iterations == st2.n_iterations and st.best_index == st2.best_index
    assert g.iterations()[st.best_index].new_error == pytest.approx(g2.iterations()[st2.best_index].new_error, rel=1e-9)
    g.close()
    g2.close()


def test_check_derivatives_is_a_small_problem_tool():
    g = capi.SfxProblem(P.bal_problem("ladybug", solver=D.SOLVER_SCHUR))
    with pytest.raises(RuntimeError, match="rc=3"):  # SFX_ERR_UNSUPPORTED: N = 23,769
        g.check_derivatives()
    st = g.optimize(2)  # still usable
    assert st.n_iterations == 3
    g.close()
