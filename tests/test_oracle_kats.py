"""
Pins the CPU oracle against the known-answer tests the reference holds for this path
(SURVEY.md section 8c).  All CPU; the GPU parity tests then compare against this oracle.
"""
import numpy as np
import pytest
import scipy.sparse as sp

from symforce_b200 import desc as D
from symforce_b200 import problems as P
from tests import oracle_capi as O


def test_pose_smoothing_static():
    # test/symforce_optimizer_test.cc:141-166: iteration == 12, lambda ~ 0.0039 (10%), error ~ 7.801 (1e-3)
    params = D.default_params()
    params.iterations = 50
    params.early_exit_min_reduction = 0.0001
    o = O.OracleProblem(P.pose_smoothing(params))
    st = o.optimize()
    its = o.iterations()
    last = its[-1]
    assert st.status == D.STATUS_SUCCESS
    assert last.iteration == 12
    assert last.current_lambda == pytest.approx(0.0039, rel=1e-1)
    assert last.new_error == pytest.approx(7.801, rel=1e-3)


def test_pose_smoothing_dynamic_and_best_linearization():
    # test/symforce_optimizer_test.cc:461-506: 27 stats entries; residual 66, H 60x60 with 534 nonzeros
    params = D.default_params()
    params.lambda_update_type = D.LAMBDA_DYNAMIC
    o = O.OracleProblem(P.pose_smoothing(params))
    st = o.optimize()
    assert st.status == D.STATUS_SUCCESS
    assert st.failure_reason == 0
    assert st.n_iterations == 27
    N, M, nnz = o.dims()
    assert (N, M, nnz) == (60, 66, 534)
    res, rhs, H = o.best_linearization()
    assert np.all(np.isfinite(res)) and np.all(np.isfinite(rhs)) and np.all(np.isfinite(H))


def test_rotation_smoothing():
    # test/symforce_optimizer_test.cc:238-256: iteration == 6, lambda ~ 2.4e-4, error ~ 2.174
    params = D.default_params()
    params.iterations = 50
    params.early_exit_min_reduction = 0.0001
    o = O.OracleProblem(P.rotation_smoothing(params))
    st = o.optimize()
    last = o.iterations()[-1]
    assert st.status == D.STATUS_SUCCESS
    assert last.iteration == 6
    assert last.current_lambda == pytest.approx(2.4e-4, rel=1e-1)
    assert last.new_error == pytest.approx(2.174, rel=1e-3)


def test_frozen_out_of_order_keys():
    # test/symforce_optimizer_test.cc:314-338: iteration == 5, lambda < 1e-3, error < 1e-15
    params = D.default_params()
    params.iterations = 50
    params.early_exit_min_reduction = 0.0001
    o = O.OracleProblem(P.frozen_keys(params))
    st = o.optimize()
    last = o.iterations()[-1]
    assert st.status == D.STATUS_SUCCESS
    assert last.iteration == 5
    assert last.current_lambda < 1e-3
    assert last.new_error < 1e-15


def test_robot_3d_localization():
    # test/symforce_examples_robot_3d_localization_test.py:49-51: initial error 463700.5576620833
    # (constants printed to 12 decimals -> ~1e-9 relative), final < 140, SUCCESS
    o = O.OracleProblem(P.robot_3d_localization())
    st = o.optimize()
    its = o.iterations()
    assert st.status == D.STATUS_SUCCESS
    assert its[0].new_error == pytest.approx(463700.5576620833, rel=1e-8)
    assert its[st.best_index].new_error < 140


def _csc(o, H):
    N, M, nnz = o.dims()
    outer, inner = o.hessian_pattern()
    return sp.csc_matrix((H, inner, outer), shape=(N, N))


def test_linearizer_identities_and_repeatability():
    # test/symforce_linearizer_test.cc:113-125: first vs subsequent relinearize identical;
    # H symmetric-consistent with rhs: check H == tril(J^T J) through a dense numerical Jacobian
    # of the residual in the tangent space (CheckLinearError-style, 1e-6).
    prob = P.pose_smoothing()
    o = O.OracleProblem(prob)
    res1, rhs1, H1 = o.linearize()
    res2, rhs2, H2 = o.linearize()
    assert np.array_equal(res1, res2) and np.array_equal(rhs1, rhs2) and np.array_equal(H1, H2)
    # numerical Jacobian through retract
    N = 60
    J = np.zeros((66, N))
    h = 1e-6
    import ctypes as C
    lib = o.lib
    for i in range(N):
        for sgn in (+1, -1):
            v = prob.values.copy()
            d = np.zeros(N)
            d[i] = sgn * h
            for k in range(10):
                st = v[prob.keys[k, 1]:prob.keys[k, 1] + 7].copy()
                dd = np.ascontiguousarray(d[6 * k:6 * k + 6])
                lib.orc_retract(C.c_int(D.TYPE_POSE3), C.c_int(6), st.ctypes.data_as(C.POINTER(C.c_double)),
                                dd.ctypes.data_as(C.POINTER(C.c_double)), C.c_double(D.K_DEFAULT_EPSILON))
                v[prob.keys[k, 1]:prob.keys[k, 1] + 7] = st
            o.set_values(v)
            r, _, _ = o.linearize()
            J[:, i] += sgn * r / (2 * h)
    A = _csc(o, H1).toarray()
    JtJ = J.T @ J
    assert np.allclose(np.tril(A), np.tril(JtJ), rtol=1e-5, atol=1e-5 * np.abs(JtJ).max())
    assert np.allclose(rhs1, J.T @ res1, rtol=1e-5, atol=1e-5 * np.abs(rhs1).max())


def _solve_oracle_ldlt(A_lower, b, ordering=D.ORDERING_METIS_SCALAR):
    import ctypes as C
    lib = O.load()
    A = sp.csc_matrix(A_lower)
    A.sort_indices()
    n = A.shape[0]
    x = np.zeros(n)
    outer = A.indptr.astype(np.int32)
    inner = A.indices.astype(np.int32)
    val = A.data.astype(np.float64)
    pi, pd = C.POINTER(C.c_int32), C.POINTER(C.c_double)
    rc = lib.orc_ldlt_solve(C.c_int(n), outer.ctypes.data_as(pi), inner.ctypes.data_as(pi), val.ctypes.data_as(pd),
                            C.c_int(ordering), np.ascontiguousarray(b).ctypes.data_as(pd), x.ctypes.data_as(pd),
                            None, None)
    assert rc == 0, lib.orc_last_error()
    return x


def _solve_oracle_schur(A_lower, C_dim, b):
    import ctypes as C
    lib = O.load()
    A = sp.csc_matrix(A_lower)
    A.sort_indices()
    n = A.shape[0]
    x = np.zeros(n)
    outer = A.indptr.astype(np.int32)
    inner = A.indices.astype(np.int32)
    val = A.data.astype(np.float64)
    pi, pd = C.POINTER(C.c_int32), C.POINTER(C.c_double)
    rc = lib.orc_schur_solve(C.c_int(n), outer.ctypes.data_as(pi), inner.ctypes.data_as(pi), val.ctypes.data_as(pd),
                             C.c_int(C_dim), np.ascontiguousarray(b).ctypes.data_as(pd), x.ctypes.data_as(pd))
    assert rc == 0, lib.orc_last_error()
    return x


def test_sparse_ldlt_random_spd():
    # test/sparse_cholesky_solver_test.cc:59-185: random 300x300 SPD patterns, several numeric
    # refactorizations, against an independent solver, isApprox 1e-5
    rng = np.random.default_rng(0)
    for trial in range(10):
        n = 300
        M = sp.random(n, n, density=0.01, random_state=rng.integers(1 << 30), format="csc")
        A = (M @ M.T + sp.identity(n) * (1.0 + rng.random())).tocsc()
        b = rng.normal(size=n)
        for ordering in (D.ORDERING_METIS_SCALAR, D.ORDERING_NATURAL):
            x = _solve_oracle_ldlt(sp.tril(A), b, ordering)
            x_ref = np.linalg.solve(A.toarray(), b)
            assert np.allclose(x, x_ref, rtol=1e-5, atol=1e-8)


def build_small_schur_matrix():
    """test/sparse_schur_solver_test.cc:30-57 BuildSmallMatrix: 10 pose dims + 30 landmark dims in
    2x2 blocks... restated: a Jacobian with a dense pose part and block-diagonal landmark part, filled
    with 1, 2, 3, ..., A = J^T J + I."""
    n_pose, n_lm, lm_dim = 10, 15, 2
    rows = n_lm * 2
    J = np.zeros((rows, n_pose + n_lm * lm_dim))
    v = 1.0
    for l in range(n_lm):
        for r in range(2):
            row = l * 2 + r
            for c in range(n_pose):
                J[row, c] = v
                v += 1
            for c in range(lm_dim):
                J[row, n_pose + l * lm_dim + c] = v
                v += 1
    J /= v
    A = J.T @ J + np.eye(J.shape[1])
    return A, n_lm * lm_dim


def test_schur_solver_small_matrix():
    # test/sparse_schur_solver_test.cc:103-250: Schur solution vs sparse Cholesky, tol 1e-3 (double),
    # plus 5 random diagonal rescalings (seed 12345, U[1,5])
    A, C_dim = build_small_schur_matrix()
    rng = np.random.default_rng(12345)
    n = A.shape[0]
    for k in range(6):
        Ak = A.copy()
        if k > 0:
            s = rng.uniform(1, 5, n)
            Ak = Ak * np.sqrt(np.outer(s, s))
        b = rng.normal(size=n)
        x_s = _solve_oracle_schur(sp.tril(sp.csc_matrix(Ak)), C_dim, b)
        x_c = _solve_oracle_ldlt(sp.tril(sp.csc_matrix(Ak)), b)
        x_d = np.linalg.solve(Ak, b)
        assert np.allclose(x_s, x_c, rtol=1e-3, atol=1e-9)
        assert np.allclose(x_s, x_d, rtol=1e-8, atol=1e-10)


def test_bal_schur_vs_full_cholesky_same_iterates():
    """North star wiring: LM with the Schur linear solver must follow the same iterates as LM with the
    default full-H SparseCholeskySolver (exact elimination)."""
    pf = P.bal_problem("small", solver=D.SOLVER_CHOLESKY)
    ps = P.bal_problem("small", solver=D.SOLVER_SCHUR)
    of, os_ = O.OracleProblem(pf), O.OracleProblem(ps)
    sf, ss = of.optimize(), os_.optimize()
    itf, its = of.iterations(), os_.iterations()
    assert sf.status == ss.status == D.STATUS_SUCCESS
    assert len(itf) == len(its)
    for a, b in zip(itf, its):
        assert a.new_error == pytest.approx(b.new_error, rel=1e-8)
        assert a.update_accepted == b.update_accepted
    assert its[-1].new_error < 0.05 * its[0].new_error
    np.testing.assert_allclose(of.best_values(), os_.best_values(), rtol=1e-6, atol=1e-8)


def test_status_codes():
    # test/symforce_optimizer_test.cc:367-430: HIT_ITERATION_LIMIT and LAMBDA_OUT_OF_BOUNDS semantics
    params = D.default_params()
    params.iterations = 2
    o = O.OracleProblem(P.pose_smoothing(params))
    st = o.optimize()
    assert st.status == D.STATUS_HIT_ITERATION_LIMIT
    assert st.n_iterations == 3
