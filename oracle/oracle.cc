/*
 * oracle.cc -- CPU restatement of the reference's sparse Levenberg-Marquardt path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing in the product (symforce_b200/, include/) links, includes or
 * calls this file; only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline/reference leg
 * may load liboracle.so, and there only as the checker / reported CPU baseline.
 *
 * It is an Eigen-free, single-threaded C++17 restatement (the reference's symforce/opt has no
 * threading) of, function by function:
 *   Linearizer::BuildInitialLinearization / Relinearize / UpdateFromLinearizedDenseFactorIntoSparse
 *       symforce/opt/linearizer.cc:55-120, 149-356, 359-433; internal/linearizer_utils.h:219-286
 *   LevenbergMarquardtSolver::Iterate / DampHessian + the 3-block state machine
 *       symforce/opt/levenberg_marquardt_solver.tcc:23-54, 139-343;
 *       symforce/opt/internal/levenberg_marquardt_state.h:96-158, 217-261
 *   OptimizeImpl / IterateToConvergenceImpl   symforce/opt/internal/optimizer_utils.h:36-104
 *   Linearization::Error / LinearDeltaError    symforce/opt/linearization.h:51-67
 *   Values::Retract (Pose3 / Rot3 / vector)    symforce/opt/values.cc:294-327,
 *       gen/cpp/sym/ops/pose3/lie_group_ops.cc:70-103, gen/cpp/sym/ops/rot3/lie_group_ops.cc:63-92,
 *       ctor normalisation gen/cpp/sym/pose3.h:66-69, gen/cpp/sym/rot3.h:46-47
 *   SparseCholeskySolver (simplicial up-looking LDL^T + METIS ordering)
 *       symforce/opt/sparse_cholesky/sparse_cholesky_solver.tcc:13-259
 *   SparseSchurSolver                          symforce/opt/sparse_schur_solver.tcc:16-170
 * Factor arithmetic: oracle/gen/factors_gen.h (generated from the reference's symbolic residuals by
 * tools/gen_factors.py; validated against the reference's own generated headers compiled in
 * place -> oracle/_ref, see oracle/ref_factors.cc and tests/test_oracle_factors.py).
 *
 * Third-party pieces not present under /root/reference, restated from their published behaviour:
 *   Eigen 3.4.0 setFromTriplets (sorted CSC, duplicates summed), twistedBy, MetisOrdering
 *   (adjacency of pattern(A)+pattern(A^T) without diagonal -> METIS_NodeND, default options),
 *   dense LLT for the C blocks; METIS 5.1.0 METIS_NodeND (here: the image's 64-bit-idx_t static
 *   library, so the permutation itself is NOT pinned to upstream's 32-bit build -- "ordering
 *   parity unpinned"; solution-level parity is pinned by the reference's KATs, see tests/).
 */
#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <functional>
#include <map>
#include <numeric>
#include <stdexcept>
#include <string>
#include <unordered_map>
#include <vector>

#include "../include/sfx.h"
#include "gen/factors_gen.h"
#include "gen/kinds_gen.h"

extern "C" int METIS_NodeND(int64_t* nvtxs, int64_t* xadj, int64_t* adjncy, int64_t* vwgt,
                            int64_t* options, int64_t* perm, int64_t* iperm);

namespace orc {

using Clock = std::chrono::steady_clock;
static double now_s() {
  return std::chrono::duration<double>(Clock::now().time_since_epoch()).count();
}

#define ORC_ASSERT(cond, msg)                                                          \
  do {                                                                                 \
    if (!(cond)) throw std::runtime_error(std::string("ORC_ASSERT: " #cond " -- ") + (msg)); \
  } while (0)

// ------------------------------------------------------------------------------------------------
// Sparse containers
// ------------------------------------------------------------------------------------------------
struct Csc {
  int n = 0;
  std::vector<int> outer;  // n+1
  std::vector<int> inner;  // nnz
  std::vector<double> val; // nnz
  int64_t nnz() const { return (int64_t)inner.size(); }
};

// ------------------------------------------------------------------------------------------------
// Ordering: Eigen::MetisOrdering semantics.  Returns perm (new -> old) and iperm (old -> new).
// ------------------------------------------------------------------------------------------------
static void metis_ordering_lower(const Csc& A_lower, std::vector<int>& perm, std::vector<int>& iperm) {
  const int n = A_lower.n;
  perm.resize(n);
  iperm.resize(n);
  if (n == 0) return;
  // Symmetrise: adjacency of column j = rows of A(:,j) and of A^T(:,j), diagonal excluded.
  // A_selfadjoint is the full symmetric matrix in the reference (sparse_cholesky_solver.tcc:17),
  // so each column's adjacency is just its sorted off-diagonal index set.
  std::vector<int64_t> deg(n, 0);
  for (int j = 0; j < n; ++j)
    for (int p = A_lower.outer[j]; p < A_lower.outer[j + 1]; ++p) {
      const int i = A_lower.inner[p];
      if (i != j) {
        deg[i]++;
        deg[j]++;
      }
    }
  std::vector<int64_t> xadj(n + 1, 0);
  for (int j = 0; j < n; ++j) xadj[j + 1] = xadj[j] + deg[j];
  std::vector<int64_t> adj(xadj[n]);
  std::vector<int64_t> fill(xadj.begin(), xadj.end() - 1);
  // Upper part of column j (rows < j) come from A^T: entries (j, c) with c < j; traversing columns
  // in increasing order emits them in increasing c, then the lower part (rows > j) in CSC order:
  // the result is sorted by index.
  for (int j = 0; j < n; ++j)
    for (int p = A_lower.outer[j]; p < A_lower.outer[j + 1]; ++p) {
      const int i = A_lower.inner[p];
      if (i != j) adj[fill[i]++] = j;  // (i, j), i > j: j is an "upper" neighbour of i
    }
  for (int j = 0; j < n; ++j)
    for (int p = A_lower.outer[j]; p < A_lower.outer[j + 1]; ++p) {
      const int i = A_lower.inner[p];
      if (i != j) adj[fill[j]++] = i;
    }
  int64_t nv = n;
  std::vector<int64_t> p64(n), ip64(n);
  bool any_edge = xadj[n] > 0;
  if (!any_edge) {
    for (int i = 0; i < n; ++i) p64[i] = ip64[i] = i;
  } else {
    int rc = METIS_NodeND(&nv, xadj.data(), adj.data(), nullptr, nullptr, p64.data(), ip64.data());
    ORC_ASSERT(rc == 1, "METIS_NodeND failed");
  }
  for (int i = 0; i < n; ++i) {
    perm[i] = (int)p64[i];
    iperm[i] = (int)ip64[i];
  }
}

// ------------------------------------------------------------------------------------------------
// Simplicial up-looking LDL^T (sparse_cholesky_solver.tcc)
// ------------------------------------------------------------------------------------------------
struct Ldlt {
  int n = 0;
  bool initialized = false;
  int ordering = SFX_ORDERING_METIS_SCALAR;
  std::vector<int> perm, iperm;       // perm: new->old (Eigen inv_permutation_), iperm: old->new
  Csc Ap;                             // upper triangle of twisted A
  std::vector<int> ap_src;            // Ap.val[k] = A.val[ap_src[k]]
  std::vector<int> parent, visited, nnz_per_col, pattern;
  std::vector<int> L_outer, L_inner;
  std::vector<double> L_val, D, D_agg;

  void twist(const Csc& A) {
    for (size_t k = 0; k < ap_src.size(); ++k) Ap.val[k] = A.val[ap_src[k]];
  }

  void compute_symbolic(const Csc& A) {
    n = A.n;
    if (ordering == SFX_ORDERING_NATURAL) {
      perm.resize(n);
      iperm.resize(n);
      std::iota(perm.begin(), perm.end(), 0);
      std::iota(iperm.begin(), iperm.end(), 0);
    } else {
      metis_ordering_lower(A, perm, iperm);
    }
    // Build twisted upper-triangular pattern, columns sorted by row.
    std::vector<int> cnt(n + 1, 0);
    const int64_t nnz = A.nnz();
    std::vector<int> tr(nnz), tc(nnz);
    for (int j = 0; j < n; ++j)
      for (int p = A.outer[j]; p < A.outer[j + 1]; ++p) {
        int a = iperm[A.inner[p]], b = iperm[j];
        int r = std::min(a, b), c = std::max(a, b);
        tr[p] = r;
        tc[p] = c;
        cnt[c + 1]++;
      }
    Ap.n = n;
    Ap.outer.assign(n + 1, 0);
    for (int j = 0; j < n; ++j) Ap.outer[j + 1] = Ap.outer[j] + cnt[j + 1];
    Ap.inner.resize(nnz);
    Ap.val.resize(nnz);
    ap_src.resize(nnz);
    // counting sort by (col, row): first bucket by row then stable by col
    std::vector<int> order(nnz);
    {
      std::vector<int> rc(n + 1, 0);
      for (int64_t p = 0; p < nnz; ++p) rc[tr[p] + 1]++;
      for (int i = 0; i < n; ++i) rc[i + 1] += rc[i];
      std::vector<int> byrow(nnz);
      for (int64_t p = 0; p < nnz; ++p) byrow[rc[tr[p]]++] = (int)p;
      std::vector<int> pos(Ap.outer.begin(), Ap.outer.end() - 1);
      for (int64_t k = 0; k < nnz; ++k) {
        int p = byrow[k];
        int dst = pos[tc[p]]++;
        Ap.inner[dst] = tr[p];
        ap_src[dst] = p;
      }
    }
    visited.assign(n, -1);
    parent.assign(n, -1);
    nnz_per_col.assign(n, 0);
    for (int k = 0; k < n; ++k) {
      visited[k] = k;
      for (int p = Ap.outer[k]; p < Ap.outer[k + 1]; ++p) {
        int i = Ap.inner[p];
        if (i >= k) continue;
        while (visited[i] != k) {
          if (parent[i] == -1) parent[i] = k;
          nnz_per_col[i]++;
          visited[i] = k;
          i = parent[i];
        }
      }
    }
    L_outer.assign(n + 1, 0);
    for (int k = 0; k < n; ++k) L_outer[k + 1] = L_outer[k] + nnz_per_col[k];
    L_inner.assign(L_outer[n], 0);
    L_val.assign(L_outer[n], 0.0);
    D.assign(n, 0.0);
    pattern.assign(n, 0);
    D_agg.assign(n, 0.0);
    initialized = true;
  }

  bool factorize(const Csc& A) {
    ORC_ASSERT(A.n == n, "size");
    twist(A);
    std::fill(nnz_per_col.begin(), nnz_per_col.end(), 0);
    std::fill(D_agg.begin(), D_agg.end(), 0.0);
    for (int k = 0; k < n; ++k) {
      visited[k] = k;
      int top = n;
      for (int p = Ap.outer[k]; p < Ap.outer[k + 1]; ++p) {
        int i = Ap.inner[p];
        if (i > k) continue;
        D_agg[i] += Ap.val[p];
        int depth = 0;
        while (visited[i] != k) {
          pattern[depth] = i;
          visited[i] = k;
          i = parent[i];
          depth++;
        }
        while (depth > 0) {
          top--;
          depth--;
          pattern[top] = pattern[depth];
        }
      }
      double Dk = D_agg[k];
      D_agg[k] = 0.0;
      for (; top < n; ++top) {
        const int i = pattern[top];
        const double Dagg_i = D_agg[i];
        const double Lki = Dagg_i / D[i];
        const int ps = L_outer[i];
        const int pe = ps + nnz_per_col[i];
        D_agg[i] = 0.0;
        int p;
        for (p = ps; p < pe; ++p) D_agg[L_inner[p]] -= L_val[p] * Dagg_i;
        L_inner[p] = k;
        L_val[p] = Lki;
        Dk -= Lki * Dagg_i;
        nnz_per_col[i]++;
      }
      D[k] = Dk;
    }
    return true;
  }

  // in-place solve of one right-hand side (sparse_cholesky_solver.tcc:232-259)
  void solve_in_place(double* b) const {
    std::vector<double> x(n);
    for (int i = 0; i < n; ++i) x[iperm[i]] = b[i];
    for (int j = 0; j < n; ++j) {  // unit-lower forward
      const double xj = x[j];
      for (int p = L_outer[j]; p < L_outer[j + 1]; ++p) x[L_inner[p]] -= L_val[p] * xj;
    }
    for (int j = 0; j < n; ++j) x[j] /= D[j];
    for (int j = n - 1; j >= 0; --j) {  // L^T backward
      double s = x[j];
      for (int p = L_outer[j]; p < L_outer[j + 1]; ++p) s -= L_val[p] * x[L_inner[p]];
      x[j] = s;
    }
    for (int k = 0; k < n; ++k) b[perm[k]] = x[k];
  }
};

// ------------------------------------------------------------------------------------------------
// Schur complement solver (sparse_schur_solver.tcc)
// ------------------------------------------------------------------------------------------------
struct Schur {
  int total = 0, B_dim = 0, C_dim = 0;
  struct CBlock {
    int start, dim;
  };
  std::vector<CBlock> cblocks;
  std::vector<int> row_block;  // row (>=B_dim) -> block id
  // For every C block: list of (column c < B_dim, position of the block's first row in A's column c)
  std::vector<int> e_ptr;      // per block range into e_col/e_pos
  std::vector<int> e_col, e_pos;
  std::vector<double> Cinv;    // per block dense dim*dim (col-major), offset cinv_off
  std::vector<int> cinv_off;
  // runs of B columns (one run per optimized key when known; otherwise each column)
  std::vector<int> run_start;  // size nruns+1
  std::vector<int> col_run;    // B column -> run
  Csc S;
  // S assembly maps
  std::vector<int> b_src, b_dst;            // S.val[b_dst] += A.val[b_src]
  std::vector<int64_t> pair_ptr;            // per block: range of run pairs
  std::vector<int> pair_dst;                // for each (block, run pair): S position of the top-left entry per column, flattened
  Ldlt s_solver;
  bool initialized = false;

  void compute_symbolic(const Csc& A, int Cd, const std::vector<int>* key_runs) {
    total = A.n;
    C_dim = Cd;
    B_dim = total - Cd;
    cblocks.clear();
    bool in_block = false;
    for (int col = B_dim; col < total; ++col) {
      int start_row = -1, prev_row = -1;
      for (int p = A.outer[col]; p < A.outer[col + 1]; ++p) {
        int r = A.inner[p];
        if (start_row == -1) start_row = r;
        if (prev_row != -1)
          ORC_ASSERT(r == prev_row + 1, "Submatrix C of A is not block diagonal");
        prev_row = r;
      }
      ORC_ASSERT(start_row != -1 && start_row == col, "C column empty or diagonal missing");
      int nz = prev_row - start_row + 1;
      if (in_block) {
        const CBlock& b = cblocks.back();
        int off = col - b.start;
        ORC_ASSERT(nz == b.dim - off, "C block not dense");
        if (off == b.dim - 1) in_block = false;
      } else {
        cblocks.push_back({col, nz});
        if (nz > 1) in_block = true;
      }
    }
    row_block.assign(C_dim, -1);
    cinv_off.assign(cblocks.size() + 1, 0);
    for (size_t b = 0; b < cblocks.size(); ++b) {
      for (int r = 0; r < cblocks[b].dim; ++r) row_block[cblocks[b].start - B_dim + r] = (int)b;
      cinv_off[b + 1] = cinv_off[b] + cblocks[b].dim * cblocks[b].dim;
    }
    Cinv.assign(cinv_off.back(), 0.0);
    // runs
    run_start.clear();
    if (key_runs && !key_runs->empty()) {
      for (int s : *key_runs)
        if (s < B_dim) run_start.push_back(s);
      run_start.push_back(B_dim);
    } else {
      for (int c = 0; c <= B_dim; ++c) run_start.push_back(c);
    }
    const int nruns = (int)run_start.size() - 1;
    col_run.assign(B_dim, 0);
    for (int r = 0; r < nruns; ++r)
      for (int c = run_start[r]; c < run_start[r + 1]; ++c) col_run[c] = r;

    // E^T structure: per block list of columns
    std::vector<int> cnt(cblocks.size() + 1, 0);
    for (int c = 0; c < B_dim; ++c)
      for (int p = A.outer[c]; p < A.outer[c + 1]; ++p) {
        int r = A.inner[p];
        if (r < B_dim) continue;
        int b = row_block[r - B_dim];
        if (r == cblocks[b].start) cnt[b + 1]++;
      }
    e_ptr.assign(cblocks.size() + 1, 0);
    for (size_t b = 0; b < cblocks.size(); ++b) e_ptr[b + 1] = e_ptr[b] + cnt[b + 1];
    e_col.resize(e_ptr.back());
    e_pos.resize(e_ptr.back());
    std::vector<int> fillp(e_ptr.begin(), e_ptr.end() - 1);
    for (int c = 0; c < B_dim; ++c)
      for (int p = A.outer[c]; p < A.outer[c + 1]; ++p) {
        int r = A.inner[p];
        if (r < B_dim) continue;
        int b = row_block[r - B_dim];
        if (r == cblocks[b].start) {
          // all rows of the block must be present, contiguously
          ORC_ASSERT(p + cblocks[b].dim <= A.outer[c + 1] &&
                         A.inner[p + cblocks[b].dim - 1] == r + cblocks[b].dim - 1,
                     "E block not dense");
          e_col[fillp[b]] = c;
          e_pos[fillp[b]] = p;
          fillp[b]++;
        } else {
          ORC_ASSERT(p > A.outer[c] && A.inner[p - 1] == r - 1, "E block not dense (2)");
        }
      }
    // S pattern at run-block level: set of (run_i >= run_j)
    std::vector<std::vector<int>> runrows(nruns);
    for (int c = 0; c < B_dim; ++c)
      for (int p = A.outer[c]; p < A.outer[c + 1]; ++p) {
        int r = A.inner[p];
        if (r >= B_dim) break;
        runrows[col_run[c]].push_back(col_run[r]);
      }
    for (size_t b = 0; b < cblocks.size(); ++b) {
      // distinct runs touched by this block (columns sorted -> runs sorted)
      int last = -1;
      std::vector<int> rs;
      for (int k = e_ptr[b]; k < e_ptr[b + 1]; ++k) {
        int r = col_run[e_col[k]];
        if (r != last) rs.push_back(r), last = r;
      }
      for (size_t j = 0; j < rs.size(); ++j)
        for (size_t i = j; i < rs.size(); ++i) runrows[rs[j]].push_back(rs[i]);
    }
    for (auto& v : runrows) {
      std::sort(v.begin(), v.end());
      v.erase(std::unique(v.begin(), v.end()), v.end());
    }
    // Expand to scalar CSC (lower).  NB: within a run pair the block is treated as dense, which is
    // a superset of Eigen's product pattern whenever a key's E block is dense (always, here).
    S.n = B_dim;
    S.outer.assign(B_dim + 1, 0);
    for (int c = 0; c < B_dim; ++c) {
      int rj = col_run[c];
      int64_t n = run_start[rj + 1] - c;  // diag run: rows c..end of run
      for (int ri : runrows[rj])
        if (ri != rj) n += run_start[ri + 1] - run_start[ri];
      S.outer[c + 1] = S.outer[c] + (int)n;
    }
    S.inner.resize(S.outer[B_dim]);
    S.val.assign(S.outer[B_dim], 0.0);
    // rowpos[run pair] offsets: map (rj, ri) -> offset of run ri's first row inside column c,
    // relative to the column start, for c the first column of run rj minus the diag shrink.
    std::vector<std::unordered_map<int, int>> run_off(nruns);  // off-diagonal: offset after diag part
    for (int rj = 0; rj < nruns; ++rj) {
      int off = 0;
      for (int ri : runrows[rj]) {
        if (ri == rj) continue;
        run_off[rj][ri] = off;
        off += run_start[ri + 1] - run_start[ri];
      }
    }
    for (int c = 0; c < B_dim; ++c) {
      int rj = col_run[c];
      int p = S.outer[c];
      for (int r = c; r < run_start[rj + 1]; ++r) S.inner[p++] = r;
      for (int ri : runrows[rj]) {
        if (ri == rj) continue;
        for (int r = run_start[ri]; r < run_start[ri + 1]; ++r) S.inner[p++] = r;
      }
    }
    auto s_pos = [&](int r, int c) -> int {  // r >= c
      int rj = col_run[c], ri = col_run[r];
      if (ri == rj) return S.outer[c] + (r - c);
      return S.outer[c] + (run_start[rj + 1] - c) + run_off[rj].at(ri) + (r - run_start[ri]);
    };
    // B copy map
    b_src.clear();
    b_dst.clear();
    for (int c = 0; c < B_dim; ++c)
      for (int p = A.outer[c]; p < A.outer[c + 1]; ++p) {
        int r = A.inner[p];
        if (r >= B_dim) break;
        b_src.push_back(p);
        b_dst.push_back(s_pos(r, c));
      }
    // pair destination map: for each block, for each ordered pair (i >= j) of its columns: S position
    pair_ptr.assign(cblocks.size() + 1, 0);
    for (size_t b = 0; b < cblocks.size(); ++b) {
      int64_t k = e_ptr[b + 1] - e_ptr[b];
      pair_ptr[b + 1] = pair_ptr[b] + k * (k + 1) / 2;
    }
    pair_dst.resize(pair_ptr.back());
    for (size_t b = 0; b < cblocks.size(); ++b) {
      int64_t q = pair_ptr[b];
      for (int j = e_ptr[b]; j < e_ptr[b + 1]; ++j)
        for (int i = j; i < e_ptr[b + 1]; ++i) pair_dst[q++] = s_pos(e_col[i], e_col[j]);
    }
    initialized = true;
  }

  void factorize(const Csc& A) {
    // C^-1 per block via dense LLT solve of the identity (sparse_schur_solver.tcc:105-119)
    for (size_t b = 0; b < cblocks.size(); ++b) {
      const int d = cblocks[b].dim, s = cblocks[b].start;
      double Lm[36 * 36];
      ORC_ASSERT(d <= 36, "C block too large for oracle");
      // gather lower block
      for (int c = 0; c < d; ++c) {
        int p = A.outer[s + c];
        for (int r = c; r < d; ++r) Lm[r + c * d] = A.val[p + (r - c)];
      }
      // cholesky (lower)
      for (int j = 0; j < d; ++j) {
        double x = Lm[j + j * d];
        for (int k = 0; k < j; ++k) x -= Lm[j + k * d] * Lm[j + k * d];
        x = std::sqrt(x);
        Lm[j + j * d] = x;
        for (int i = j + 1; i < d; ++i) {
          double y = Lm[i + j * d];
          for (int k = 0; k < j; ++k) y -= Lm[i + k * d] * Lm[j + k * d];
          Lm[i + j * d] = y / x;
        }
      }
      double* Ci = &Cinv[cinv_off[b]];
      for (int e = 0; e < d; ++e) {
        double y[36];
        for (int i = 0; i < d; ++i) {
          double v = (i == e) ? 1.0 : 0.0;
          for (int k = 0; k < i; ++k) v -= Lm[i + k * d] * y[k];
          y[i] = v / Lm[i + i * d];
        }
        for (int i = d - 1; i >= 0; --i) {
          double v = y[i];
          for (int k = i + 1; k < d; ++k) v -= Lm[k + i * d] * y[k];
          y[i] = v / Lm[i + i * d];
        }
        for (int i = 0; i < d; ++i) Ci[i + e * d] = y[i];
      }
      // the reference stores only the lower part and reads it back as selfadjoint: symmetrise
      for (int c = 0; c < d; ++c)
        for (int r = 0; r < c; ++r) Ci[r + c * d] = Ci[c + r * d];
    }
    // S = B - E C^-1 E^T (lower)
    std::fill(S.val.begin(), S.val.end(), 0.0);
    for (size_t k = 0; k < b_src.size(); ++k) S.val[b_dst[k]] += A.val[b_src[k]];
    std::vector<double> t;
    for (size_t b = 0; b < cblocks.size(); ++b) {
      const int d = cblocks[b].dim;
      const double* Ci = &Cinv[cinv_off[b]];
      const int k0 = e_ptr[b], k1 = e_ptr[b + 1];
      t.resize((size_t)(k1 - k0) * d);
      for (int j = k0; j < k1; ++j) {
        const double* ej = &A.val[e_pos[j]];
        for (int a = 0; a < d; ++a) {
          double s = 0;
          for (int bb = 0; bb < d; ++bb) s += Ci[a + bb * d] * ej[bb];
          t[(size_t)(j - k0) * d + a] = s;
        }
      }
      int64_t q = pair_ptr[b];
      for (int j = k0; j < k1; ++j) {
        const double* tj = &t[(size_t)(j - k0) * d];
        for (int i = j; i < k1; ++i) {
          const double* ei = &A.val[e_pos[i]];
          double s = 0;
          for (int a = 0; a < d; ++a) s += ei[a] * tj[a];
          S.val[pair_dst[q++]] -= s;
        }
      }
    }
    if (!s_solver.initialized) s_solver.compute_symbolic(S);
    s_solver.factorize(S);
  }

  // rhs (length total) -> solution in place
  void solve_in_place(const Csc& A, double* x) const {
    std::vector<double> w(x + B_dim, x + total);
    std::vector<double> cw(C_dim);
    for (size_t b = 0; b < cblocks.size(); ++b) {
      const int d = cblocks[b].dim, s = cblocks[b].start - B_dim;
      const double* Ci = &Cinv[cinv_off[b]];
      for (int a = 0; a < d; ++a) {
        double v = 0;
        for (int bb = 0; bb < d; ++bb) v += Ci[a + bb * d] * w[s + bb];
        cw[s + a] = v;
      }
    }
    // schur_rhs = v - E * C_inv * w
    std::vector<double> y(x, x + B_dim);
    for (size_t b = 0; b < cblocks.size(); ++b) {
      const int d = cblocks[b].dim, s = cblocks[b].start - B_dim;
      for (int k = e_ptr[b]; k < e_ptr[b + 1]; ++k) {
        const double* e = &A.val[e_pos[k]];
        double v = 0;
        for (int a = 0; a < d; ++a) v += e[a] * cw[s + a];
        y[e_col[k]] -= v;
      }
    }
    s_solver.solve_in_place(y.data());
    // z = C_inv * (w - E^T y)
    std::vector<double> u(w);
    for (size_t b = 0; b < cblocks.size(); ++b) {
      const int d = cblocks[b].dim, s = cblocks[b].start - B_dim;
      for (int k = e_ptr[b]; k < e_ptr[b + 1]; ++k) {
        const double* e = &A.val[e_pos[k]];
        const double yc = y[e_col[k]];
        for (int a = 0; a < d; ++a) u[s + a] -= e[a] * yc;
      }
    }
    for (int i = 0; i < B_dim; ++i) x[i] = y[i];
    for (size_t b = 0; b < cblocks.size(); ++b) {
      const int d = cblocks[b].dim, s = cblocks[b].start - B_dim;
      const double* Ci = &Cinv[cinv_off[b]];
      for (int a = 0; a < d; ++a) {
        double v = 0;
        for (int bb = 0; bb < d; ++bb) v += Ci[a + bb * d] * u[s + bb];
        x[B_dim + s + a] = v;
      }
    }
  }
};

// ------------------------------------------------------------------------------------------------
// Factor evaluation
// ------------------------------------------------------------------------------------------------
static inline void eval_factor(int kind, const double* const* a, double* res, double* J) {
  switch (kind) {
    case SFX_KIND_SNAVELY: sfx_factor_snavely(a[0], a[1], a[2], a[3], a[4], res, J); break;
    case SFX_KIND_BETWEEN_POSE3: sfx_factor_between_pose3(a[0], a[1], a[2], a[3], a[4], res, J); break;
    case SFX_KIND_PRIOR_POSE3: sfx_factor_prior_pose3(a[0], a[1], a[2], a[3], res, J); break;
    case SFX_KIND_MATCHING: sfx_factor_matching(a[0], a[1], a[2], a[3], res, J); break;
    case SFX_KIND_ODOMETRY: sfx_factor_odometry(a[0], a[1], a[2], a[3], a[4], res, J); break;
    case SFX_KIND_IRL_LINEAR_GNC:
      sfx_factor_irl_linear_gnc(a[0], a[1], a[2], a[3], a[4], a[5], a[6], a[7], a[8], a[9], a[10], res, J);
      break;
    case SFX_KIND_IRL_PRIOR: sfx_factor_irl_prior(a[0], a[1], a[2], a[3], a[4], res, J); break;
    case SFX_KIND_BETWEEN_ROT3: sfx_factor_between_rot3(a[0], a[1], a[2], a[3], a[4], res, J); break;
    case SFX_KIND_PRIOR_ROT3: sfx_factor_prior_rot3(a[0], a[1], a[2], a[3], res, J); break;
    case SFX_KIND_BARRON: sfx_factor_barron(a[0], a[1], a[2], a[3], res, J); break;
    default: throw std::runtime_error("unknown factor kind");
  }
}

// ------------------------------------------------------------------------------------------------
// Retract (values.cc:294-327 + generated LieGroupOps)
// ------------------------------------------------------------------------------------------------
static void retract_rot3(double* a, const double* v, double eps) {
  const double t0 = std::sqrt(eps * eps + v[0] * v[0] + v[1] * v[1] + v[2] * v[2]);
  const double t1 = 0.5 * t0;
  const double s = std::sin(t1) / t0;
  const double c = std::cos(t1);
  // quaternion product a * [s*v, c]
  const double x = a[0], y = a[1], z = a[2], w = a[3];
  const double bx = s * v[0], by = s * v[1], bz = s * v[2];
  double r[4];
  r[0] = x * c + y * bz - z * by + w * bx;
  r[1] = -x * bz + y * c + z * bx + w * by;
  r[2] = z * c + w * bz + x * by - y * bx;
  r[3] = -z * bz + w * c - x * bx - y * by;
  const double n2 = r[0] * r[0] + r[1] * r[1] + r[2] * r[2] + r[3] * r[3];
  if (n2 > 0) {
    const double n = std::sqrt(n2);
    for (int i = 0; i < 4; ++i) r[i] /= n;
  }
  for (int i = 0; i < 4; ++i) a[i] = r[i];
}

static void retract_key(int type, int tangent_dim, double* a, const double* v, double eps) {
  switch (type) {
    case SFX_TYPE_VECTOR:
      for (int i = 0; i < tangent_dim; ++i) a[i] += v[i];
      break;
    case SFX_TYPE_ROT3: retract_rot3(a, v, eps); break;
    case SFX_TYPE_POSE3:
      retract_rot3(a, v, eps);
      a[4] += v[3];
      a[5] += v[4];
      a[6] += v[5];
      break;
    default: throw std::runtime_error("unknown key type");
  }
}

// ------------------------------------------------------------------------------------------------
// Problem: linearizer + LM
// ------------------------------------------------------------------------------------------------
struct Linearization {
  std::vector<double> residual, rhs, H;
  bool initialized = false;
  double error() const {
    double s = 0;
    for (double r : residual) s += r * r;
    return 0.5 * s;
  }
};

struct KeyHelper {
  int factor_offset, tangent_dim, combined_offset;
  std::vector<int> hcol_starts;
};

struct FactorHelper {
  int kind;
  int res_off, res_dim;
  int arg_off[ORC_MAX_ARGS];
  std::vector<KeyHelper> keys;
};

struct StateBlock {
  std::vector<double> values;
  Linearization lin;
  bool have_err = false;
  double err = 0;
  double error() {
    if (!have_err) {
      err = lin.error();
      have_err = true;
    }
    return err;
  }
};

struct Timings {
  double linearize = 0, factorize = 0, solve = 0, total = 0, setup = 0;
  int n_lin = 0, n_fac = 0, iters = 0;
};

struct Problem {
  sfx_params p;
  double eps;
  int64_t n_values;
  std::vector<sfx_key_entry> keys;
  std::vector<int> key_toff;  // tangent offsets
  int N = 0, M = 0;
  std::vector<FactorHelper> factors;  // in factor_index order
  Csc Hpat;                           // pattern only (val unused)
  int solver, schur_num_keys, ordering;
  int C_dim = 0;
  std::vector<int> key_runs;
  // LM state
  StateBlock blocks[3];
  int init_idx = 0, new_idx = 1, best_idx = 0, free_idx = 2;
  bool best_valid = false;
  double lambda = 0, nu = 0;
  int iteration = -1;
  bool have_max_diag = false, have_last_update = false, solver_analyzed = false;
  std::vector<double> max_diag, update, last_update, damping, undamped;
  Ldlt ldlt;
  Schur schur;
  std::vector<sfx_iteration> iters;
  sfx_stats stats{};
  std::vector<double> cur_values;  // values given by set_values
  Linearization cur_lin;           // for orc_linearize / solve_step
  std::string err;
  Timings tm;

  StateBlock& Init() { return blocks[init_idx]; }
  StateBlock& New() { return blocks[new_idx]; }
  StateBlock& Best() { return blocks[best_idx]; }

  void build(const sfx_problem_desc& d);
  void relinearize(const std::vector<double>& values, Linearization& lin);
  void jacobian(const std::vector<double>& values, Csc& jac) const;
  void damp(std::vector<double>& H, double lam);
  void undamp(std::vector<double>& H);
  void analyze(const Csc& A);
  void factorize(const Csc& A);
  void solve(const Csc& A, std::vector<double>& x);
  int iterate();
  void optimize(int num_iterations);
  void optimize_continue(int num_iterations);
  void run_iterations(int num_iterations);
};

void Problem::build(const sfx_problem_desc& d) {
  double t0 = now_s();
  p = d.params;
  eps = d.epsilon;
  n_values = d.n_values;
  solver = d.solver;
  schur_num_keys = d.schur_num_keys;
  ordering = d.ordering;
  keys.assign(d.keys, d.keys + d.n_keys);
  key_toff.resize(d.n_keys + 1);
  key_toff[0] = 0;
  for (int k = 0; k < d.n_keys; ++k) key_toff[k + 1] = key_toff[k] + keys[k].tangent_dim;
  N = key_toff[d.n_keys];
  key_runs.assign(key_toff.begin(), key_toff.end() - 1);
  if (solver == SFX_SOLVER_SCHUR) {
    ORC_ASSERT(schur_num_keys > 0 && schur_num_keys < d.n_keys, "schur_num_keys");
    C_dim = N - key_toff[d.n_keys - schur_num_keys];
  }
  // factors in caller order
  factors.resize(d.n_factors);
  std::vector<char> seen(d.n_factors, 0);
  for (int b = 0; b < d.n_batches; ++b) {
    const sfx_factor_batch& fb = d.batches[b];
    ORC_ASSERT(fb.kind >= 0 && fb.kind < ORC_NUM_KINDS, "kind");
    const orc_kind_meta& km = ORC_KIND_META[fb.kind];
    for (int f = 0; f < fb.n; ++f) {
      int fi = fb.factor_index[f];
      ORC_ASSERT(fi >= 0 && fi < d.n_factors && !seen[fi], "factor_index must be a permutation");
      seen[fi] = 1;
      FactorHelper& h = factors[fi];
      h.kind = fb.kind;
      h.res_dim = km.res_dim;
      for (int a = 0; a < km.n_args; ++a) h.arg_off[a] = fb.arg_offsets[(int64_t)a * fb.n + f];
      int foff = 0;
      for (int o = 0; o < km.n_opt; ++o) {
        int key = fb.opt_keys[(int64_t)o * fb.n + f];
        if (key >= 0) {
          ORC_ASSERT(key < d.n_keys, "opt key index");
          ORC_ASSERT(keys[key].tangent_dim == km.opt_dims[o], "tangent dim mismatch");
          KeyHelper kh;
          kh.factor_offset = foff;
          kh.tangent_dim = km.opt_dims[o];
          kh.combined_offset = key_toff[key];
          h.keys.push_back(kh);
        }
        foff += km.opt_dims[o];
      }
    }
  }
  int roff = 0;
  for (auto& h : factors) {
    h.res_off = roff;
    roff += h.res_dim;
  }
  M = roff;
  // "Key ... is in the state vector but is not optimized by any factor" (linearizer.cc:277-284)
  {
    std::vector<char> touched(d.n_keys, 0);
    std::unordered_map<int, int> off2key;
    for (int k = 0; k < d.n_keys; ++k) off2key[key_toff[k]] = k;
    for (auto& h : factors)
      for (auto& kh : h.keys) touched[off2key[kh.combined_offset]] = 1;
    for (int k = 0; k < d.n_keys; ++k)
      if (!touched[k])
        throw std::runtime_error("Key " + std::to_string(k) +
                                 " is in the state vector but is not optimized by any factor.");
  }
  // ---- Hessian pattern at key-block level (same pattern setFromTriplets gives, linearizer.cc:324)
  const int nk = d.n_keys;
  std::unordered_map<int, int> off2key;
  for (int k = 0; k < nk; ++k) off2key[key_toff[k]] = k;
  std::vector<std::vector<int>> colrows(nk);  // off-diagonal row keys per column key
  for (auto& h : factors) {
    for (size_t i = 0; i < h.keys.size(); ++i)
      for (size_t j = 0; j < i; ++j) {
        int ki = off2key[h.keys[i].combined_offset], kj = off2key[h.keys[j].combined_offset];
        if (ki == kj) continue;
        int lo = std::min(ki, kj), hi = std::max(ki, kj);
        colrows[lo].push_back(hi);
      }
  }
  for (auto& v : colrows) {
    std::sort(v.begin(), v.end());
    v.erase(std::unique(v.begin(), v.end()), v.end());
  }
  Hpat.n = N;
  Hpat.outer.assign(N + 1, 0);
  std::vector<std::vector<int>> rowprefix(nk);
  for (int k = 0; k < nk; ++k) {
    int off = 0;
    rowprefix[k].resize(colrows[k].size());
    for (size_t q = 0; q < colrows[k].size(); ++q) {
      rowprefix[k][q] = off;
      off += keys[colrows[k][q]].tangent_dim;
    }
    const int dk = keys[k].tangent_dim;
    for (int c = 0; c < dk; ++c) {
      int64_t cnt = (int64_t)(dk - c) + off;
      int col = key_toff[k] + c;
      ORC_ASSERT((int64_t)Hpat.outer[col] + cnt < (int64_t)INT32_MAX, "nnz overflow");
      Hpat.outer[col + 1] = Hpat.outer[col] + (int)cnt;
    }
  }
  Hpat.inner.resize(Hpat.outer[N]);
  for (int k = 0; k < nk; ++k) {
    const int dk = keys[k].tangent_dim;
    for (int c = 0; c < dk; ++c) {
      int col = key_toff[k] + c;
      int pz = Hpat.outer[col];
      for (int r = c; r < dk; ++r) Hpat.inner[pz++] = key_toff[k] + r;
      for (int rk : colrows[k])
        for (int r = 0; r < keys[rk].tangent_dim; ++r) Hpat.inner[pz++] = key_toff[rk] + r;
    }
  }
  auto entry_pos = [&](int row_key, int row_local, int col_key, int col_local) -> int {
    // position in CSC of (key_toff[row_key]+row_local, key_toff[col_key]+col_local), row_key > col_key
    const int dk = keys[col_key].tangent_dim;
    int col = key_toff[col_key] + col_local;
    auto& v = colrows[col_key];
    size_t q = std::lower_bound(v.begin(), v.end(), row_key) - v.begin();
    return Hpat.outer[col] + (dk - col_local) + rowprefix[col_key][q] + row_local;
  };
  // per-factor column starts (internal/linearizer_utils.h:99-143 semantics)
  for (auto& h : factors) {
    for (size_t i = 0; i < h.keys.size(); ++i) {
      KeyHelper& ki = h.keys[i];
      int keyi = off2key[ki.combined_offset];
      for (int c = 0; c < ki.tangent_dim; ++c)
        ki.hcol_starts.push_back(Hpat.outer[ki.combined_offset + c]);  // diagonal entry is first in its column
      for (size_t j = 0; j < i; ++j) {
        const KeyHelper& kj = h.keys[j];
        int keyj = off2key[kj.combined_offset];
        if (kj.combined_offset < ki.combined_offset) {
          for (int c = 0; c < kj.tangent_dim; ++c) ki.hcol_starts.push_back(entry_pos(keyi, 0, keyj, c));
        } else if (kj.combined_offset > ki.combined_offset) {
          for (int c = 0; c < ki.tangent_dim; ++c) ki.hcol_starts.push_back(entry_pos(keyj, 0, keyi, c));
        } else {
          // same key twice in one factor: reference would add both into the diagonal block;
          // not produced by any supported kind
          throw std::runtime_error("factor references the same optimized key twice");
        }
      }
    }
  }
  for (auto& b : blocks) {
    b.lin.initialized = false;
  }
  tm.setup = now_s() - t0;
}

void Problem::relinearize(const std::vector<double>& values, Linearization& lin) {
  double t0 = now_s();
  lin.residual.resize(M);
  lin.rhs.assign(N, 0.0);
  lin.H.assign(Hpat.nnz(), 0.0);
  double res[8], J[8 * 16], Hf[16 * 16], rf[16];
  const double* a[ORC_MAX_ARGS];
  for (const FactorHelper& h : factors) {
    const orc_kind_meta& km = ORC_KIND_META[h.kind];
    for (int k = 0; k < km.n_args; ++k) a[k] = values.data() + h.arg_off[k];
    eval_factor(h.kind, a, res, J);
    const int R = km.res_dim, T = km.tan_dim;
    // Gauss-Newton blocks: H = J^T J (lower), rhs = J^T r  (codegen.py:796-807)
    for (int c = 0; c < T; ++c) {
      for (int r = c; r < T; ++r) {
        double s = 0;
        for (int q = 0; q < R; ++q) s += J[q + r * R] * J[q + c * R];
        Hf[r + c * T] = s;
      }
      double s = 0;
      for (int q = 0; q < R; ++q) s += J[q + c * R] * res[q];
      rf[c] = s;
    }
    for (int q = 0; q < R; ++q) lin.residual[h.res_off + q] = res[q];
    // UpdateFromLinearizedDenseFactorIntoSparse (linearizer.cc:359-433)
    for (size_t i = 0; i < h.keys.size(); ++i) {
      const KeyHelper& ki = h.keys[i];
      for (int r = 0; r < ki.tangent_dim; ++r) lin.rhs[ki.combined_offset + r] += rf[ki.factor_offset + r];
      size_t it = 0;
      for (int c = 0; c < ki.tangent_dim; ++c) {
        int cs = ki.hcol_starts[it++];
        for (int r = 0; r < ki.tangent_dim - c; ++r)
          lin.H[cs + r] += Hf[(ki.factor_offset + c + r) + (ki.factor_offset + c) * T];
      }
      for (size_t j = 0; j < i; ++j) {
        const KeyHelper& kj = h.keys[j];
        if (kj.combined_offset < ki.combined_offset) {
          for (int c = 0; c < kj.tangent_dim; ++c) {
            int cs = ki.hcol_starts[it++];
            for (int r = 0; r < ki.tangent_dim; ++r)
              lin.H[cs + r] += Hf[(ki.factor_offset + r) + (kj.factor_offset + c) * T];
          }
        } else {
          for (int c = 0; c < ki.tangent_dim; ++c) {
            int cs = ki.hcol_starts[it++];
            for (int r = 0; r < kj.tangent_dim; ++r)
              lin.H[cs + r] += Hf[(ki.factor_offset + c) + (kj.factor_offset + r) * T];
          }
        }
      }
    }
  }
  lin.initialized = true;
  tm.linearize += now_s() - t0;
  tm.n_lin++;
}

// Linearization::jacobian with include_jacobians (linearizer.cc:252-259, 297-313): one triplet per entry of every
// factor's dense Jacobian block over its optimized keys, in factor order, compressed like Eigen's setFromTriplets
// (column-major, rows ascending, duplicates summed) into an M x N CSC matrix.
void Problem::jacobian(const std::vector<double>& values, Csc& jac) const {
  struct Trip {
    int row, col;
    double v;
  };
  std::vector<Trip> trips;
  double res[8], J[8 * 16];
  const double* a[ORC_MAX_ARGS];
  for (const FactorHelper& h : factors) {
    const orc_kind_meta& km = ORC_KIND_META[h.kind];
    for (int k = 0; k < km.n_args; ++k) a[k] = values.data() + h.arg_off[k];
    eval_factor(h.kind, a, res, J);
    const int R = km.res_dim;
    for (const KeyHelper& kh : h.keys)
      for (int c = 0; c < kh.tangent_dim; ++c)
        for (int r = 0; r < R; ++r)
          trips.push_back(Trip{h.res_off + r, kh.combined_offset + c, J[r + (kh.factor_offset + c) * R]});
  }
  std::stable_sort(trips.begin(), trips.end(),
                   [](const Trip& x, const Trip& y) { return x.col != y.col ? x.col < y.col : x.row < y.row; });
  jac.n = N;
  jac.outer.assign(N + 1, 0);
  jac.inner.clear();
  jac.val.clear();
  for (size_t i = 0; i < trips.size(); ++i) {
    if (i > 0 && trips[i].col == trips[i - 1].col && trips[i].row == trips[i - 1].row) {
      jac.val.back() += trips[i].v;
      continue;
    }
    jac.inner.push_back(trips[i].row);
    jac.val.push_back(trips[i].v);
    jac.outer[trips[i].col + 1]++;
  }
  for (int j = 0; j < N; ++j) jac.outer[j + 1] += jac.outer[j];
}

// DampHessian (levenberg_marquardt_solver.tcc:23-54)
void Problem::damp(std::vector<double>& H, double lam) {
  undamped.resize(N);
  damping.resize(N);
  for (int i = 0; i < N; ++i) undamped[i] = H[Hpat.outer[i]];
  if (p.use_diagonal_damping) {
    if (p.keep_max_diagonal_damping) {
      if (!have_max_diag) {
        max_diag.resize(N);
        for (int i = 0; i < N; ++i) max_diag[i] = std::max(undamped[i], p.diagonal_damping_min);
      } else {
        for (int i = 0; i < N; ++i) max_diag[i] = std::max(max_diag[i], undamped[i]);
      }
      have_max_diag = true;
      for (int i = 0; i < N; ++i) damping[i] = max_diag[i] * lam;
    } else {
      for (int i = 0; i < N; ++i) damping[i] = std::max(undamped[i], p.diagonal_damping_min) * lam;
    }
  } else {
    std::fill(damping.begin(), damping.end(), 0.0);
  }
  if (p.use_unit_damping)
    for (int i = 0; i < N; ++i) damping[i] += lam;
  for (int i = 0; i < N; ++i) H[Hpat.outer[i]] += damping[i];
}

void Problem::undamp(std::vector<double>& H) {
  for (int i = 0; i < N; ++i) H[Hpat.outer[i]] = undamped[i];
}

void Problem::analyze(const Csc& A) {
  if (solver == SFX_SOLVER_SCHUR) {
    schur.s_solver.ordering = ordering == SFX_ORDERING_NATURAL ? SFX_ORDERING_NATURAL : SFX_ORDERING_METIS_SCALAR;
    schur.compute_symbolic(A, C_dim, &key_runs);
  } else {
    ldlt.ordering = ordering == SFX_ORDERING_NATURAL ? SFX_ORDERING_NATURAL : SFX_ORDERING_METIS_SCALAR;
    ldlt.compute_symbolic(A);
  }
}

void Problem::factorize(const Csc& A) {
  double t0 = now_s();
  if (solver == SFX_SOLVER_SCHUR)
    schur.factorize(A);
  else
    ldlt.factorize(A);
  tm.factorize += now_s() - t0;
  tm.n_fac++;
}

void Problem::solve(const Csc& A, std::vector<double>& x) {
  double t0 = now_s();
  if (solver == SFX_SOLVER_SCHUR)
    schur.solve_in_place(A, x.data());
  else
    ldlt.solve_in_place(x.data());
  tm.solve += now_s() - t0;
}

// Shares pattern arrays with Hpat but carries the numeric values of a linearization.
static Csc view_with_values(const Csc& pat, const std::vector<double>& v) {
  Csc A;
  A.n = pat.n;
  A.outer = pat.outer;
  A.inner = pat.inner;
  A.val = v;
  return A;
}

// returns 0 = continue, else optimization_status_t; failure reason in stats.failure_reason
int Problem::iterate() {
  // state_.Step(); iteration_++
  std::swap(init_idx, new_idx);
  iteration++;
  if (!Init().lin.initialized) {
    relinearize(Init().values, Init().lin);
    Init().have_err = false;
    // SetBestToInit
    best_valid = true;
    if (best_idx != init_idx) {
      if (best_idx != new_idx) free_idx = best_idx;
      best_idx = init_idx;
    }
  }
  if (iteration == 0) {
    sfx_iteration it{};
    it.iteration = -1;
    it.new_error = Init().error();
    it.current_lambda = lambda;
    iters.push_back(it);
    if (!std::isfinite(Init().error())) {
      stats.failure_reason = 2;
      return 3;
    }
  }
  static thread_local Csc A;  // reused buffer
  if (A.n != Hpat.n || A.inner.size() != Hpat.inner.size()) {
    A.n = Hpat.n;
    A.outer = Hpat.outer;
    A.inner = Hpat.inner;
  }
  A.val.swap(Init().lin.H);
  if (!solver_analyzed) {
    double t0 = now_s();
    analyze(A);
    tm.setup += now_s() - t0;
    solver_analyzed = true;
  }
  damp(A.val, lambda);
  factorize(A);
  update.assign(Init().lin.rhs.begin(), Init().lin.rhs.end());
  solve(A, update);
  for (double& u : update) u = -u;
  undamp(A.val);
  A.val.swap(Init().lin.H);

  // UpdateNewFromInit: copy + retract (levenberg_marquardt_state.h:217-239)
  New().values = Init().values;
  for (size_t k = 0; k < keys.size(); ++k)
    retract_key(keys[k].type, keys[k].tangent_dim, New().values.data() + keys[k].offset,
                update.data() + key_toff[k], eps);
  relinearize(New().values, New().lin);
  New().have_err = false;

  const double new_error = New().error();
  const double init_error = Init().error();
  const double relative_reduction = (init_error - new_error) / (init_error + eps);
  // LinearDeltaError (linearization.h:62-67)
  double lde = 0;
  for (int i = 0; i < N; ++i) lde += update[i] * (Init().lin.rhs[i] - damping[i] * update[i]);
  const double new_error_linear = init_error + 0.5 * lde;
  const double gain_ratio = (init_error - new_error) / (init_error - new_error_linear);

  sfx_iteration it{};
  it.iteration = iteration;
  it.current_lambda = lambda;
  it.new_error = new_error;
  it.new_error_linear = new_error_linear;
  it.relative_reduction = relative_reduction;

  int status = 0;
  if (relative_reduction > -p.early_exit_min_reduction / 10 && relative_reduction < p.early_exit_min_reduction) {
    status = 1;
  } else if (new_error < p.early_exit_min_absolute_error) {
    status = 1;
  }
  bool accept = relative_reduction > 0;
  double angle = 0;
  if (p.enable_bold_updates && have_last_update && !accept) {
    double nl = 0, nu2 = 0, dot = 0;
    for (int i = 0; i < N; ++i) {
      nl += last_update[i] * last_update[i];
      nu2 += update[i] * update[i];
      dot += last_update[i] * update[i];
    }
    angle = dot / (std::sqrt(nl) * std::sqrt(nu2));
    accept = ((1 - angle) * (1 - angle) * new_error) <= Best().error();
  }
  if (!accept && lambda >= p.lambda_upper_bound) {
    status = 3;
    stats.failure_reason = 1;
  }
  if (!accept) {
    if (p.lambda_update_type == 1) {
      lambda *= p.lambda_up_factor;
    } else {
      lambda *= nu;
      nu *= 2;
    }
    std::swap(init_idx, new_idx);
  } else {
    if (p.lambda_update_type == 1) {
      lambda *= p.lambda_down_factor;
    } else {
      lambda *= std::max(1.0 / p.dynamic_lambda_update_gamma,
                         1.0 - (p.dynamic_lambda_update_beta - 1) *
                                   std::pow(2 * gain_ratio - 1, (double)p.dynamic_lambda_update_p));
      nu = 2;
    }
    have_last_update = true;
    last_update = update;
    if (New().error() <= Best().error()) {
      // SetBestToNew
      best_valid = true;
      if (best_idx != new_idx) {
        if (best_idx != init_idx) free_idx = best_idx;
        best_idx = new_idx;
      }
      stats.best_index = (int)iters.size();  // index of the entry about to be pushed
    }
    // SetInitToNotBest
    if (best_idx == init_idx) {
      init_idx = free_idx;
      free_idx = best_idx;
    }
  }
  lambda = std::min(std::max(lambda, p.lambda_lower_bound), p.lambda_upper_bound);
  it.update_angle_change = angle;
  it.update_accepted = accept ? 1 : 0;
  iters.push_back(it);
  return status;
}

void Problem::optimize(int num_iterations) {
  if (num_iterations < 0) num_iterations = p.iterations;
  ORC_ASSERT(num_iterations > 0, "num_iterations must be positive");
  double t0 = now_s();
  // Reset (levenberg_marquardt_solver.h:163-182; state Reset :81-89, ResetValues :250-261)
  lambda = p.initial_lambda;
  nu = p.dynamic_lambda_update_beta;
  iteration = -1;
  have_max_diag = false;
  have_last_update = false;
  New().values = cur_values;
  for (auto& b : blocks) {
    b.lin.initialized = false;
    b.have_err = false;
  }
  best_valid = false;
  iters.clear();
  stats = sfx_stats{};
  run_iterations(num_iterations);
  tm.total += now_s() - t0;
}

// OptimizeContinue (symforce/opt/gnc_optimizer.h:133-142): nonlinear_solver_.ResetState(values)
// (levenberg_marquardt_solver.h:178-183: have_max_diagonal_ = have_last_update_ = false, state_.Reset(values):
// internal/levenberg_marquardt_state.h:81-89, 250-261) then IterateToConvergence; lambda, nu, the iteration
// counter and the stats keep accumulating.
void Problem::optimize_continue(int num_iterations) {
  ORC_ASSERT(num_iterations > 0, "num_iterations must be positive");
  double t0 = now_s();
  have_max_diag = false;
  have_last_update = false;
  New().values = cur_values;
  for (auto& b : blocks) {
    b.lin.initialized = false;
    b.have_err = false;
  }
  best_valid = false;
  stats.status = 0;
  stats.failure_reason = 0;
  run_iterations(num_iterations);
  tm.total += now_s() - t0;
}

// IterateToConvergenceImpl (symforce/opt/internal/optimizer_utils.h:36-73)
void Problem::run_iterations(int num_iterations) {
  int i;
  for (i = 0; i < num_iterations; ++i) {
    int st = iterate();
    if (st) {
      stats.status = st;
      if (st != 3) stats.failure_reason = 0;
      break;
    }
  }
  if (i == num_iterations) {
    stats.status = 2;
    stats.failure_reason = 0;
  }
  stats.n_iterations = (int)iters.size();
  tm.iters += std::min(i + 1, num_iterations);
}

}  // namespace orc

// ------------------------------------------------------------------------------------------------
// C ABI (ctypes)
// ------------------------------------------------------------------------------------------------
using orc::Problem;
static thread_local std::string g_err;

#define ORC_TRY(p, ...)                  \
  try {                                  \
    __VA_ARGS__;                         \
    return 0;                              \
  } catch (const std::exception& e) {    \
    g_err = e.what();                    \
    if (p) ((Problem*)p)->err = e.what(); \
    return 1;                            \
  }

extern "C" {

const char* orc_last_error() { return g_err.c_str(); }

int orc_create(const sfx_problem_desc* d, void** out) {
  Problem* p = nullptr;
  try {
    p = new Problem();
    p->build(*d);
    *out = p;
    return 0;
  } catch (const std::exception& e) {
    g_err = e.what();
    delete p;
    *out = nullptr;
    return 1;
  }
}

void orc_destroy(void* p) { delete (Problem*)p; }

int orc_update_params(void* p, const sfx_params* params) { ORC_TRY(p, ((Problem*)p)->p = *params); }

int orc_set_values(void* p, const double* v, int64_t n) {
  ORC_TRY(p, {
    Problem* P = (Problem*)p;
    ORC_ASSERT(n == P->n_values, "values length");
    P->cur_values.assign(v, v + n);
    P->cur_lin.initialized = false;
  });
}

int orc_optimize(void* p, int num_iterations, sfx_stats* stats) {
  ORC_TRY(p, {
    Problem* P = (Problem*)p;
    P->optimize(num_iterations);
    if (stats) *stats = P->stats;
  });
}

int orc_optimize_continue(void* p, int num_iterations, sfx_stats* stats) {
  ORC_TRY(p, {
    ((Problem*)p)->optimize_continue(num_iterations);
    if (stats) *stats = ((Problem*)p)->stats;
  });
}

// RelaxDampingToInitial (levenberg_marquardt_solver.h:171-174)
int orc_relax_damping_to_initial(void* p) {
  ORC_TRY(p, {
    Problem* q = (Problem*)p;
    q->lambda = std::min(q->lambda, q->p.initial_lambda);
    q->nu = q->p.dynamic_lambda_update_beta;
  });
}

int orc_get_best_values(void* p, double* v, int64_t n) {
  ORC_TRY(p, {
    Problem* P = (Problem*)p;
    ORC_ASSERT(P->best_valid && n == P->n_values, "best values");
    std::copy(P->Best().values.begin(), P->Best().values.end(), v);
  });
}

int orc_get_iterations(void* p, sfx_iteration* buf, int cap, int* n) {
  ORC_TRY(p, {
    Problem* P = (Problem*)p;
    *n = (int)P->iters.size();
    for (int i = 0; i < std::min(cap, *n); ++i) buf[i] = P->iters[i];
  });
}

int orc_get_dims(void* p, int32_t* N, int32_t* M, int64_t* nnz) {
  ORC_TRY(p, {
    Problem* P = (Problem*)p;
    *N = P->N;
    *M = P->M;
    *nnz = P->Hpat.nnz();
  });
}

int orc_get_hessian_pattern(void* p, int32_t* outer, int32_t* inner) {
  ORC_TRY(p, {
    Problem* P = (Problem*)p;
    std::copy(P->Hpat.outer.begin(), P->Hpat.outer.end(), outer);
    std::copy(P->Hpat.inner.begin(), P->Hpat.inner.end(), inner);
  });
}

int orc_linearize(void* p, double* residual, double* rhs, double* H) {
  ORC_TRY(p, {
    Problem* P = (Problem*)p;
    P->relinearize(P->cur_values, P->cur_lin);
    if (residual) std::copy(P->cur_lin.residual.begin(), P->cur_lin.residual.end(), residual);
    if (rhs) std::copy(P->cur_lin.rhs.begin(), P->cur_lin.rhs.end(), rhs);
    if (H) std::copy(P->cur_lin.H.begin(), P->cur_lin.H.end(), H);
  });
}

// jacobian at the values last set: pass outer == NULL to query nnz only; any output may be NULL
int orc_linearize_jacobian(void* p, int64_t* nnz, int32_t* outer, int32_t* inner, double* values) {
  ORC_TRY(p, {
    Problem* P = (Problem*)p;
    orc::Csc jac;
    P->jacobian(P->cur_values, jac);
    if (nnz) *nnz = (int64_t)jac.inner.size();
    if (outer) std::copy(jac.outer.begin(), jac.outer.end(), outer);
    if (inner) std::copy(jac.inner.begin(), jac.inner.end(), inner);
    if (values) std::copy(jac.val.begin(), jac.val.end(), values);
  });
}

int orc_get_best_linearization(void* p, double* residual, double* rhs, double* H) {
  ORC_TRY(p, {
    Problem* P = (Problem*)p;
    ORC_ASSERT(P->best_valid && P->Best().lin.initialized, "best linearization");
    const orc::Linearization& L = P->Best().lin;
    if (residual) std::copy(L.residual.begin(), L.residual.end(), residual);
    if (rhs) std::copy(L.rhs.begin(), L.rhs.end(), rhs);
    if (H) std::copy(L.H.begin(), L.H.end(), H);
  });
}

int orc_solve_step(void* p, double lambda, double* update) {
  ORC_TRY(p, {
    Problem* P = (Problem*)p;
    if (!P->cur_lin.initialized) P->relinearize(P->cur_values, P->cur_lin);
    orc::Csc A = orc::view_with_values(P->Hpat, P->cur_lin.H);
    if (!P->solver_analyzed) {
      P->analyze(A);
      P->solver_analyzed = true;
    }
    bool save = P->have_max_diag;
    P->damp(A.val, lambda);
    P->have_max_diag = save;
    P->factorize(A);
    std::vector<double> x(P->cur_lin.rhs);
    P->solve(A, x);
    for (int i = 0; i < P->N; ++i) update[i] = -x[i];
  });
}

int orc_get_ordering(void* p, int32_t* perm, int cap, int* n) {
  ORC_TRY(p, {
    Problem* P = (Problem*)p;
    const std::vector<int>& pm = P->solver == SFX_SOLVER_SCHUR ? P->schur.s_solver.perm : P->ldlt.perm;
    *n = (int)pm.size();
    for (int i = 0; i < std::min(cap, *n); ++i) perm[i] = pm[i];
  });
}

// timings: [setup, linearize, factorize, solve, total, n_lin, n_fac, iters]
int orc_get_timings(void* p, double* out8) {
  ORC_TRY(p, {
    Problem* P = (Problem*)p;
    out8[0] = P->tm.setup;
    out8[1] = P->tm.linearize;
    out8[2] = P->tm.factorize;
    out8[3] = P->tm.solve;
    out8[4] = P->tm.total;
    out8[5] = P->tm.n_lin;
    out8[6] = P->tm.n_fac;
    out8[7] = P->tm.iters;
  });
}

int orc_reset_timings(void* p) {
  ORC_TRY(p, {
    double s = ((Problem*)p)->tm.setup;
    ((Problem*)p)->tm = orc::Timings{};
    ((Problem*)p)->tm.setup = s;
  });
}

// ---- standalone solver KAT helpers (lower-triangular CSC in, one rhs) --------------------------
int orc_ldlt_solve(int n, const int32_t* outer, const int32_t* inner, const double* val, int ordering,
                   const double* b, double* x, int32_t* perm_out, int64_t* nnz_L) {
  ORC_TRY(nullptr, {
    orc::Csc A;
    A.n = n;
    A.outer.assign(outer, outer + n + 1);
    A.inner.assign(inner, inner + outer[n]);
    A.val.assign(val, val + outer[n]);
    orc::Ldlt s;
    s.ordering = ordering;
    s.compute_symbolic(A);
    s.factorize(A);
    std::copy(b, b + n, x);
    s.solve_in_place(x);
    if (perm_out) std::copy(s.perm.begin(), s.perm.end(), perm_out);
    if (nnz_L) *nnz_L = s.L_outer[n];
  });
}

int orc_schur_solve(int n, const int32_t* outer, const int32_t* inner, const double* val, int C_dim,
                    const double* b, double* x) {
  ORC_TRY(nullptr, {
    orc::Csc A;
    A.n = n;
    A.outer.assign(outer, outer + n + 1);
    A.inner.assign(inner, inner + outer[n]);
    A.val.assign(val, val + outer[n]);
    orc::Schur s;
    s.compute_symbolic(A, C_dim, nullptr);
    s.factorize(A);
    std::copy(b, b + n, x);
    s.solve_in_place(A, x);
  });
}

// Evaluate one factor (residual, column-major J, lower H = J^T J, rhs = J^T r); used to pin the
// generated arithmetic against the reference's own generated headers (oracle/_ref).
int orc_eval_factor(int kind, const double* const* args, double* res, double* J, double* H, double* rhs) {
  ORC_TRY(nullptr, {
    ORC_ASSERT(kind >= 0 && kind < ORC_NUM_KINDS, "kind");
    const orc_kind_meta& km = ORC_KIND_META[kind];
    double r_[8], J_[8 * 16];
    orc::eval_factor(kind, args, r_, J_);
    const int R = km.res_dim, T = km.tan_dim;
    for (int i = 0; i < R; ++i) res[i] = r_[i];
    for (int i = 0; i < R * T; ++i) J[i] = J_[i];
    for (int c = 0; c < T; ++c) {
      for (int r = 0; r < T; ++r) {
        double s = 0;
        for (int q = 0; q < R; ++q) s += J_[q + r * R] * J_[q + c * R];
        H[r + c * T] = (r >= c) ? s : 0.0;
      }
      double s = 0;
      for (int q = 0; q < R; ++q) s += J_[q + c * R] * r_[q];
      rhs[c] = s;
    }
  });
}

int orc_retract(int type, int tangent_dim, double* storage, const double* delta, double eps) {
  ORC_TRY(nullptr, orc::retract_key(type, tangent_dim, storage, delta, eps));
}

int orc_kind_info(int kind, int* n_args, int* arg_dims, int* n_opt, int* opt_args, int* opt_dims,
                  int* res_dim, int* tan_dim) {
  ORC_TRY(nullptr, {
    ORC_ASSERT(kind >= 0 && kind < ORC_NUM_KINDS, "kind");
    const orc_kind_meta& km = ORC_KIND_META[kind];
    *n_args = km.n_args;
    for (int i = 0; i < km.n_args; ++i) arg_dims[i] = km.arg_dims[i];
    *n_opt = km.n_opt;
    for (int i = 0; i < km.n_opt; ++i) {
      opt_args[i] = km.opt_args[i];
      opt_dims[i] = km.opt_dims[i];
    }
    *res_dim = km.res_dim;
    *tan_dim = km.tan_dim;
  });
}

}  // extern "C"
