// Internal host-side structures of libsfx (not part of the C ABI).
//
// Vocabulary
//   key    : an optimized variable of the caller (sym::Key + index_entry_t), reference order
//   node   : one or more keys that always occur together in factors (e.g. BAL camera pose +
//            intrinsics) treated as one block variable; nodes are the rows/cols of the
//            block-sparse Hessian.  Landmark (Schur-eliminated) keys are always their own node.
//   block  : dense (dim(row node) x dim(col node)) column-major tile of the lower Hessian
//   front  : dense frontal matrix of one supernode of the multifrontal Cholesky
#pragma once
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../include/sfx.h"
#include "gen/kinds_gen.h"

namespace sfx {

struct Error : std::runtime_error {
  sfx_status code;
  Error(sfx_status c, const std::string& m) : std::runtime_error(m), code(c) {}
};

#define SFX_CHECK(cond, code, msg)                                 \
  do {                                                             \
    if (!(cond)) throw ::sfx::Error(code, std::string(msg) + " [" #cond "]"); \
  } while (0)

constexpr int kMaxNodeDim = 16;
constexpr int kMaxLandmarkDim = 3;
constexpr int kMaxGroups = 3;

struct KeyInfo {
  int type, voff, sdim, tdim;
  int ref_toff;  // tangent offset in reference (keys_) order
  int node;      // node id
  int sub;       // offset of this key inside its node
};

struct NodeInfo {
  int dim;
  int toff;  // internal tangent offset
  int first_key, n_keys;
};

// Block-sparse symmetric matrix (lower), blocks sorted by (col, row); CSC-of-blocks.
struct BlockMatrix {
  int n_nodes = 0;
  std::vector<int> node_dim;
  std::vector<int> node_off;   // scalar offsets (n_nodes + 1)
  std::vector<int> col_ptr;    // n_nodes + 1, into row_idx/blk_off
  std::vector<int> row_idx;    // row node of each block (first entry of every column is the diagonal)
  std::vector<int64_t> blk_off;  // value offset of each block (col-major, ld = dim(row))
  int64_t n_values = 0;
  int find(int row, int col) const;  // block id or -1
};

// One device launch batch: factors of one kind with identical grouping / fixed-key pattern.
struct BatchPlan {
  int kind = 0;
  int n = 0;
  int n_used_args = 0;
  int used_args[SFX_MAX_ARGS];
  int n_opt = 0;
  int key_group[SFX_MAX_OPT];  // local group of each optimized arg; -1 when the key is fixed
  int key_sub[SFX_MAX_OPT];    // offset inside the group's node
  int n_groups = 0;
  int group_dim[kMaxGroups];   // node dim of each local group
  // host copies of the SoA index arrays (uploaded by the problem)
  std::vector<int32_t> arg_off;   // [n_used_args][n]
  std::vector<int32_t> res_off;   // [n]
  std::vector<int32_t> rhs_off;   // [n_groups][n] internal tangent offset of the group's node
  std::vector<int32_t> diag_off;  // [n_groups][n] value offset of the node's diagonal block
  std::vector<uint32_t> off_off;  // [n_groups*(n_groups-1)/2][n] off-diagonal block offset | flags
  std::vector<int32_t> factor_index;  // [n] caller's factor index of each slot
  // Jacobian export (build_jacobian_csc): per optimized arg and slot, the position of entry (row 0, column 0) of the
  // factor's Jacobian block in the CSC value array (-1: key is fixed) and the entry count of that key's columns
  std::vector<int32_t> jac_base;    // [n_opt][n]
  std::vector<int32_t> jac_colnnz;  // [n_opt][n]
};
constexpr uint32_t kOffExclusive = 0x80000000u;  // block has exactly one contributor: plain store
constexpr uint32_t kOffTransposed = 0x40000000u; // node(g) < node(h): write the transpose
constexpr uint32_t kOffMask = 0x3fffffffu;

struct SchurPlan {
  int n_landmarks = 0;        // landmarks of THIS rank
  int n_landmarks_total = 0;
  int lm_begin = 0;           // first own landmark (relative to first_lm_node)
  int first_lm_node = 0;
  int reduced_dim = 0;  // scalar dim of the reduced system
  // per landmark
  std::vector<int32_t> lm_dim, lm_cdiag_off /* H value offset of C block */, lm_toff /* internal tangent off */;
  std::vector<int32_t> lm_e_ptr;                 // n_landmarks + 1
  std::vector<int32_t> lm_e_off, lm_e_node;      // E block (landmark, node) H-value offset, reduced node
  // reduced system blocks
  BlockMatrix S;
  std::vector<int32_t> s_b_src;                  // per S block: H value offset of the B block or -1
  std::vector<int64_t> s_m_ptr;                  // per S block: match list range
  std::vector<int32_t> m_eoff_i, m_eoff_j, m_lm; // matches
  // reduced rhs: per reduced node the E blocks in its column
  std::vector<int32_t> r_ptr, r_eoff, r_lm;
};

// Multifrontal symbolic factorization of a BlockMatrix.
struct FrontPlan {
  int n = 0;                       // scalar dimension
  int n_fronts = 0;
  std::vector<int> perm_nodes;     // elimination position -> node
  std::vector<int> scalar_perm;    // elimination scalar position -> system scalar index (length n)
  // per front
  std::vector<int> f_w, f_u;       // pivot width, update height (scalars)
  std::vector<int> f_parent;       // parent front or -1
  std::vector<int> f_level;
  std::vector<int64_t> f_off;      // offset of the (w+u)^2 column-major front in the front buffer
  std::vector<int> f_piv;          // first pivot scalar position (elimination order)
  std::vector<int> f_rows_ptr;     // into f_rows: update rows as elimination scalar positions
  std::vector<int> f_rows;
  std::vector<int> f_rel_ptr;      // per front: into rel: position of each update row inside the PARENT front
  std::vector<int> f_rel;
  std::vector<int> f_child_ptr, f_child;  // children lists
  std::vector<int> f_toff;         // offset of this front's update vector in the solve workspace
  // assembly of the system matrix into fronts: per front a list of block copies
  struct Copy {
    int64_t src;      // value offset in the system matrix
    int32_t rows, cols, src_ld;
    int32_t dst_row, dst_col;  // top-left position inside the front
    int32_t transposed;        // 1: front(dst_row + c, dst_col + r) = src(r, c)
    int32_t lower_only;        // diagonal block: copy r >= c only
  };
  std::vector<int> f_copy_ptr;
  std::vector<Copy> copies;
  // level schedule
  int n_levels = 0;
  std::vector<int> level_ptr;      // fronts sorted by level: level_ptr[l]..level_ptr[l+1] into level_fronts
  std::vector<int> level_fronts;
  int64_t front_values = 0;
  int64_t nnz_L = 0;
  double flops = 0;
  int max_front = 0;
  int64_t solve_ws = 0;
};

struct Analysis {
  // sizes
  int n_keys = 0, N = 0, M = 0, n_factors = 0;
  int64_t n_values = 0;
  std::vector<KeyInfo> keys;
  std::vector<NodeInfo> nodes;
  std::vector<int> ref2int;   // reference tangent index -> internal tangent index
  BlockMatrix H;
  int64_t h_accum_values = 0; // prefix of H values that is accumulated (must be zeroed)
  int64_t b_values = 0;       // prefix holding every block among reduced nodes (B); summed across ranks
  int rank = 0, world = 1;
  std::vector<int> lm_rank_begin;  // multi-GPU: first landmark key of every rank (world + 1 entries, identical on all ranks)
  std::vector<int32_t> diag_pos;  // per internal scalar: H value offset of its diagonal entry
  std::vector<BatchPlan> batches;
  bool schur = false;
  SchurPlan sp;
  FrontPlan fp;
  // CSC export (reference layout); built lazily
  bool csc_built = false;
  std::vector<int32_t> csc_outer, csc_inner;
  std::vector<int32_t> csc_src;  // per CSC entry: H value offset
  int64_t nnz = 0;
  // Jacobian CSC (reference layout of Linearization::jacobian, M x N); built lazily
  bool jac_built = false;
  std::vector<int32_t> jac_outer, jac_inner;
  int64_t jac_nnz = 0;
};

void analyze_problem(const sfx_problem_desc& d, Analysis& a);
void build_csc(Analysis& a);
void build_jacobian_csc(Analysis& a);
void build_point_lists(const BatchPlan& bp, std::vector<int32_t>& order, std::vector<int32_t>& ptr,
                       std::vector<int32_t>& diag, std::vector<int32_t>& rhs);
// nd_depth >= 0 replaces METIS_NodeND by "dissect to that depth, then sweep" (symbolic.cc); `relax` is the share of
// explicit zeros relaxed amalgamation accepts in a merged front -- per merge, or over everything the merged supernode
// has absorbed when `cumulative`.
struct PlanOptions {
  int nd_depth = -1;
  double relax = 0.08;
  bool cumulative = false;
  int max_merge_w = 0;  // > 0: no merge produces a supernode wider than this (a merged front is ONE diagonal chain for
                        // the tile-DAG kernel: siblings that would have run side by side get serialised)
};
void build_front_plan(const BlockMatrix& A, int ordering, const std::vector<int>& ref_scalar_of_sys /* may be empty */,
                      FrontPlan& fp, const PlanOptions& opt = PlanOptions());

// SFX_TIMING=1: wall time of the host analysis phases on stderr (setup cost is outside the LM metric)
struct PhaseClock {
  std::chrono::steady_clock::time_point t0 = std::chrono::steady_clock::now();
  const bool on = getenv("SFX_TIMING") != nullptr;
  void lap(const char* what) {
    if (!on) return;
    const auto t1 = std::chrono::steady_clock::now();
    std::fprintf(stderr, "[sfx analysis] %-28s %8.1f ms\n", what, std::chrono::duration<double, std::milli>(t1 - t0).count());
    t0 = t1;
  }
};

}  // namespace sfx
