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
