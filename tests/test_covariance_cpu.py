"""Pins the numpy restatement of the covariance path (oracle/covariance_ref.py) with the property the
reference's own test uses (test/symforce_covariance_utils_test.cc:100-140: the Schur-complement block equals the
top-left block of the dense inverse)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import covariance_ref as R  # noqa: E402


def _arrowhead(rng, b, n_lm, d):
    n = b + n_lm * d
    J = rng.normal(size=(3 * n, n))
    H = J.T @ J
    # block-diagonal C: zero the coupling between different landmarks
    for i in range(n_lm):
        for j in range(n_lm):
            if i != j:
                H[b + i * d:b + (i + 1) * d, b + j * d:b + (j + 1) * d] = 0.0
    return H + 5.0 * np.eye(n)


def test_schur_block_equals_block_of_dense_inverse():
    rng = np.random.default_rng(7)
    for b, n_lm, d in [(6, 10, 1), (12, 20, 3), (30, 5, 2)]:
        H = _arrowhead(rng, b, n_lm, d)
        eps = 1e-9
        Hd = H.copy()
        idx = np.arange(b, H.shape[0])
        Hd[idx, idx] += eps
        want = np.linalg.inv(Hd)[:b, :b]
        got = R.covariance_block_schur(H, b, eps)
        assert np.allclose(got, want, rtol=1e-9, atol=1e-12)


def _golden(name):
    z = np.load(os.path.join(ROOT, "tests", "golden", "covariance_test_matrix.npz"))
    return R.dense_from_csc_lower(1000, z[name + "_outer"], z[name + "_inner"], z[name + "_values"])


def _is_approx(a, b, prec):
    """Eigen's isApprox: ||a - b||_F <= prec * min(||a||_F, ||b||_F)."""
    return np.linalg.norm(a - b) <= prec * min(np.linalg.norm(a), np.linalg.norm(b))


def test_sparse_c_block_on_the_reference_fixture_singular_c():
    """test/symforce_covariance_utils_test.cc:85-106 ("Test covariance is correct for singular C"): bottom-right corner of
    the reference's fixture matrix, block of 30, default epsilon, against the pseudo-inverse at the reference's 1e-2."""
    H = _golden("bottom_right")
    want = np.linalg.pinv(H)[:30, :30]
    got = R.covariance_block_sparse_c(H, 30)
    assert _is_approx(want, got, 1e-2)


def test_sparse_c_block_on_the_reference_fixture():
    """test/symforce_covariance_utils_test.cc:108-128 ("Test covariance is correct"): top-left corner, epsilon 0, 1e-1."""
    H = _golden("top_left")
    want = np.linalg.pinv(H)[:30, :30]
    got = R.covariance_block_sparse_c(H, 30, epsilon=0.0)
    assert _is_approx(want, got, 1e-1)


def test_general_c_block_equals_block_of_dense_inverse():
    """The c_is_block_diagonal = false branch (covariance_utils.h:142-145) on matrices whose C is NOT block diagonal."""
    rng = np.random.default_rng(11)
    for b, n in [(6, 30), (12, 72), (1, 9)]:
        J = rng.normal(size=(3 * n, n))
        H = J.T @ J + 2.0 * np.eye(n)
        eps = 1e-9
        Hd = H.copy()
        idx = np.arange(b, n)
        Hd[idx, idx] += eps
        want = np.linalg.inv(Hd)[:b, :b]
        assert np.allclose(R.covariance_block(H, b, eps, False), want, rtol=1e-9, atol=1e-12)
        # with a block-diagonal C both branches agree
    H = _arrowhead(rng, 12, 20, 3)
    assert np.allclose(R.covariance_block(H, 12, 1e-9, False), R.covariance_block(H, 12, 1e-9, True), rtol=1e-9, atol=1e-12)


def test_full_covariance_is_damped_inverse():
    rng = np.random.default_rng(8)
    H = _arrowhead(rng, 8, 4, 3)
    cov = R.full_covariance(H, 1e-9)
    assert np.allclose(cov @ (H + 1e-9 * np.eye(H.shape[0])), np.eye(H.shape[0]), atol=1e-9)


def test_dense_from_csc_lower_roundtrip():
    rng = np.random.default_rng(9)
    H = _arrowhead(rng, 4, 3, 2)
    n = H.shape[0]
    outer, inner, vals = [0], [], []
    for c in range(n):
        for r in range(c, n):
            if H[r, c] != 0.0:
                inner.append(r)
                vals.append(H[r, c])
        outer.append(len(inner))
    assert np.allclose(R.dense_from_csc_lower(n, outer, inner, vals), H)
