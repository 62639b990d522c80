"""
TEST INFRASTRUCTURE (checker only; never imported by the product path).

numpy restatement of the reference's marginal-covariance path:

* covariance_block_schur  -- Optimizer::ComputeCovariances with c_is_block_diagonal = true
  (symforce/opt/optimizer.tcc:177-199) -> internal::ComputeCovarianceBlockWithSchurComplement
  (symforce/opt/internal/covariance_utils.h:124-147): damp the C diagonal with epsilon (:131),
  S = B - E C^-1 E^T (SparseSchurSolver::Factorize, sparse_schur_solver.tcc:101-138),
  covariance = S^-1 I (SInvInPlace, :165-170).
* covariance_block_sparse_c -- internal::ComputeCovarianceBlockWithSchurComplementFromSparseC
  (symforce/opt/internal/covariance_utils.h:41-103), the c_is_block_diagonal = false branch of the same entry point
  (:142-145; the epsilon damping of :131 is applied by the caller, covariance_block): diagonal entries of C at or below
  `epsilon` are clamped to it (:65-69), C is factored (:72-76), S = B - E C^-1 E^T is accumulated column by column
  (:79-91) and inverted (:94-95).
* full_covariance -- LevenbergMarquardtSolver::ComputeCovariance
  (symforce/opt/levenberg_marquardt_solver.tcc:345-356): (H + epsilon I)^-1.

Pinned by the property the reference's own test checks (test/symforce_covariance_utils_test.cc:
the Schur block equals the top-left block of the dense (pseudo-)inverse), on random arrowhead matrices and on the
reference's own fixture matrix (tests/golden/covariance_test_matrix.npz, both cases of the reference test with its
tolerances): tests/test_covariance_cpu.py.
"""
import numpy as np


def dense_from_csc_lower(n, outer, inner, values):
    H = np.zeros((n, n))
    for c in range(n):
        for q in range(outer[c], outer[c + 1]):
            H[inner[q], c] = values[q]
    return H + np.tril(H, -1).T


def covariance_block_schur(H, block_dim, epsilon):
    H = np.array(H, dtype=np.float64, copy=True)
    n = H.shape[0]
    idx = np.arange(block_dim, n)
    H[idx, idx] += epsilon  # covariance_utils.h:131
    B = H[:block_dim, :block_dim]
    E = H[:block_dim, block_dim:]
    C = H[block_dim:, block_dim:]
    S = B - E @ np.linalg.solve(C, E.T)
    return np.linalg.solve(S, np.eye(block_dim))


def covariance_block_sparse_c(H, block_dim, epsilon=np.finfo(np.float64).eps):
    """covariance_utils.h:41-103 on a dense symmetric copy of A (the reference reads the lower triangle only)."""
    H = np.array(H, dtype=np.float64, copy=True)
    n = H.shape[0]
    B = H[:block_dim, :block_dim].copy()
    E_T = H[block_dim:, :block_dim]
    C = H[block_dim:, block_dim:].copy()
    d = np.arange(n - block_dim)
    C[d, d] = np.where(C[d, d] <= epsilon, epsilon, C[d, d])  # :65-69
    S = B
    X = np.linalg.solve(C, E_T)  # :72-76 one factorization of C, :82-87 one solve per column of E^T
    for j in range(block_dim):  # :89-90
        S[:, j] -= E_T.T @ X[:, j]
    return np.linalg.solve(S, np.eye(block_dim))  # :94-95


def covariance_block(H, block_dim, epsilon, c_is_block_diagonal):
    """internal::ComputeCovarianceBlockWithSchurComplement (covariance_utils.h:124-147)."""
    if c_is_block_diagonal:
        return covariance_block_schur(H, block_dim, epsilon)
    H = np.array(H, dtype=np.float64, copy=True)
    idx = np.arange(block_dim, H.shape[0])
    H[idx, idx] += epsilon  # :131
    return covariance_block_sparse_c(H, block_dim)


def full_covariance(H, epsilon):
    n = H.shape[0]
    return np.linalg.solve(np.array(H, dtype=np.float64) + epsilon * np.eye(n), np.eye(n))
