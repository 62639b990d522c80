"""
TEST INFRASTRUCTURE (checker only; never imported by the product path).

numpy restatement of the reference's marginal-covariance path:

* covariance_block_schur  -- Optimizer::ComputeCovariances with c_is_block_diagonal = true
  (symforce/opt/optimizer.tcc:177-199) -> internal::ComputeCovarianceBlockWithSchurComplement
  (symforce/opt/internal/covariance_utils.h:124-147): damp the C diagonal with epsilon (:131),
  S = B - E C^-1 E^T (SparseSchurSolver::Factorize, sparse_schur_solver.tcc:101-138),
  covariance = S^-1 I (SInvInPlace, :165-170).
* full_covariance -- LevenbergMarquardtSolver::ComputeCovariance
  (symforce/opt/levenberg_marquardt_solver.tcc:345-356): (H + epsilon I)^-1.

Pinned by the property the reference's own test checks (test/symforce_covariance_utils_test.cc:
the Schur block equals the top-left block of the dense inverse): tests/test_covariance_cpu.py.
The reference holds no numeric known-answer vector for this path.
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


def full_covariance(H, epsilon):
    n = H.shape[0]
    return np.linalg.solve(np.array(H, dtype=np.float64) + epsilon * np.eye(n), np.eye(n))
