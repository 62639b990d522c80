// Device-side declarations shared between kernels.cu and the host driver (sfx_api.cu).
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

#include "../../include/sfx.h"

namespace sfx {

constexpr int kMaxIterations = 1024;   // stats capacity per sfx_optimize call
constexpr int kReducePartials = 1 << 16;

// Device-resident LM control block: everything LevenbergMarquardtSolver keeps in host members
// (levenberg_marquardt_solver.h:228-262) plus the 3-block state indices
// (internal/levenberg_marquardt_state.h:262-268).  Written only by single-thread kernels.
struct Ctrl {
  sfx_params p;
  double epsilon;
  double lambda, nu;
  double err[3];          // cached 0.5*|r|^2 per state block
  int lin_valid[3];       // linearization initialized flags
  int init_idx, new_idx, best_idx, free_idx;
  int best_valid;
  int iteration;          // LM iteration counter (starts at -1)
  int have_max_diag, have_last_update;
  int done;               // 0 running, else optimization_status_t
  int failure_reason;
  int best_index;
  int n_iters;            // entries in iters[]
  int lin_target;         // state block the next linearize writes
  int skip_linearize;     // first-iteration linearize of Init only when not valid
  double red[8];          // reduction results: 0: new |r|^2 sum, 1: upd.(rhs - D.upd), 2: last.upd, 3: |last|^2, 4: |upd|^2
  int chol_fail;          // non-positive pivot seen
  int fail_where;         // tile-DAG path: (large front + 1) << 16 | pivot tile of the first failing POTRF (0: none)
  // diagnostics of the run (debug_checks / verbose, levenberg_marquardt_solver.tcc:57-90, 225-227)
  int n_chol_fail;        // iterations whose factorization met a non-positive pivot (the step is NaN and gets rejected)
  int n_nonfinite_update; // iterations with a non-finite update vector
  int n_zero_diag;        // damped diagonal entries below epsilon seen by the damping pass (debug_checks)
  int zero_diag_idx[15];  // the first of them (internal tangent order)
  sfx_iteration iters[kMaxIterations + 1];
};

struct StatePtrs {
  double* values[3];
  double* H[3];
  double* rhs[3];
  double* res[3];
};

struct LinBatch {
  int kind, n;
  const int32_t* arg_off;
  const int32_t* res_off;
  const int32_t* rhs_off;
  const int32_t* diag_off;
  const uint32_t* off_off;
  int key_group[3], key_sub[3];
  int group_dim[3];
  int n_groups;
  int partial_base;  // offset of this batch's per-CTA partial sums
  int bal_fast;      // Snavely batch with (pose+intrinsics) 9-node and 3-dim point node: linearize_bal_kernel
  // bal_fast: per-observation point contributions [slots rounded up to 128][9] and the point -> slots lists
  double* pbuf;      // nullptr: point blocks are accumulated with atomics instead
  const int32_t *pf_ptr, *pf_slot, *pf_diag, *pf_rhs;
  int n_pf;
  int pf_exclusive;  // the point blocks receive contributions from this batch only
};

struct SchurDev {
  int n_landmarks, n_reduced_nodes, reduced_dim;
  int add_b;  // 1: this rank adds B, damping and the camera rhs (rank 0 / single GPU)
  const int32_t *lm_dim, *lm_cdiag_off, *lm_toff, *lm_e_ptr, *lm_e_off, *lm_e_node;
  const int32_t* node_toff;   // reduced node -> internal tangent offset
  const int32_t* node_dim;
  int n_sblocks;
  const int32_t *s_row, *s_col;  // node ids of each S block
  const int64_t* s_off;          // S value offsets
  const int32_t* s_b_src;
  const int64_t* s_m_ptr;
  const int32_t *m_eoff_i, *m_eoff_j, *m_lm;
  // work items of schur_s_kernel: (block, first match, count, flags: bit0 first chunk, bit1 only chunk)
  int n_items;
  const int32_t *item_blk, *item_m0, *item_cnt, *item_flags;
  int64_t s_values;
  const int32_t *r_ptr, *r_eoff, *r_lm;
  // fast path (all landmarks dim 3, reduced nodes dim <= 16): entry -> reduced node, G = C^-1 E buffer
  int fast3, n_entries;
  const int32_t* r_node;
  double* G;   // same offsets as the E blocks in H (v1: C^-1 E; v2: W = L^-1 E)
  double* wl;  // v2: [n_landmarks][9] = L^-1 (i00 i10 i11 i20 i21 i22), u = L^-1 w; nullptr selects v1
  const double* zeros;  // 8 zero doubles (padding lanes of the DMMA fragments load from here)
  // s9 with the diagonal blocks taken out: S_II -= sum_l W_Il^T W_Il is accumulated by schur_w_rhs_kernel from the
  // blocks it has just whitened (a third of all matches are (I, I) pairs); items3 / pm_* then list the off-diagonal
  // blocks only (n_items3 of them) and schur_diag_init_kernel stores B_II + damping before the accumulation
  const int64_t* s_diag_off;   // per reduced node: offset of S_II in S (nullptr: diagonal blocks are s9 items)
  const int32_t* s_diag_bsrc;  // ... offset of B_II in H (-1: none)
  int n_items3;
  const void* items3;  // s9: 16-byte headers {s_off lo, s_off hi, bsrc, flags|dI<<8|dJ<<16|diag<<24|cnt<<25}; nullptr: use items2
  const int32_t *pm_i, *pm_j;  // s9: match offsets, 64 slots per item, unused slots -> zero block behind W
  const int32_t* item_toI;     // v3: tangent offset of the row node (damping of diagonal blocks)
  const void* items2;  // v2: packed 32-byte item headers (SItem2 in kernels.cu)
  // slot view of the camera columns of H (BAL shape: every block there is 27 doubles or the 81 = 3 x 27 doubles of a
  // camera's diagonal block, so the region is an array of 27-double slots that bulk copies stream tile by tile)
  const int32_t* slot_lm;    // per slot: landmark of the E block in it, -1 for the slots of a diagonal block
  const int32_t* slot_node;  // per slot: reduced node (camera) of the column it belongs to
  int n_slots;               // 0: the blocks are not laid out that way, per-entry kernels are used
  int64_t slot_base;         // offset of slot 0 in H (even: bulk copies need 16-byte aligned sources)
  double* sl;  // [n_landmarks][3] back-substitution accumulators
  double* cinv;   // [n_landmarks][9]
  double* tl;     // [n_landmarks][3]
  double* S;      // S values
  double* rhs_red;  // reduced rhs (reduced_dim)
};

struct FrontCopy {
  int64_t src;
  int32_t rows, cols, src_ld, dst_row, dst_col, transposed, lower_only, pad;
};

struct FrontDev {
  int n_fronts, n;
  const int32_t *f_w, *f_u, *f_piv, *f_rows_ptr, *f_rows, *f_rel, *f_child_ptr, *f_child, *f_toff, *f_copy_ptr;
  const int64_t* f_off;
  const FrontCopy* copies;
  const int32_t* level_fronts;
  const int32_t* scalar_perm;
  double* fronts;   // front_values
  double* twork;    // solve_ws
  double* ywork;    // n (elimination order)
};

// ---- large fronts: tile-DAG Cholesky (chol_large.cu) ---------------------------------------------
struct LargeFront {
  int64_t off;       // offset of the front in FrontDev::fronts
  int64_t linv_off;  // offset of this front's L_kk^-1 tiles (wt * 64*64 doubles)
  int m, w;
  int wt, nt;        // pivot tiles, total tiles
  int cnt_off;       // offset of the nt*nt tile version counters
  int front;         // front id
  int flag_off;      // solve flags: [wt] forward published, [wt] backward contribution counters
  int parent_lf;     // fused schedule: large-front index of the parent (-1: root, or the parent is assembled between launches)
  int64_t contrib_off;  // backward-solve contribution slots: wt * nt * 64 doubles
  int n_ea;          // fused schedule: extend-add tasks (type 6) that assemble this front inside the factor kernel
  int asm_off;       // ... and the counter (in LargeDev::counters) they bump; DIAG(0) waits for n_ea
  // fused forward substitution (tasks 8-10): the front's right-hand side b (m doubles at LargeDev::fwd_b + fb_off) and
  // its counters in LargeDev::counters at vc_off: [nt] updates applied to tile row i | [wt] y_k published |
  // [wt] L_kk^-1 ready | [1] children's update vectors added
  int fb_off, vc_off;
  int n_vch;         // fused children whose update vector is added by a VEC-EXTEND-ADD task
  int sticky;        // fused schedule: the CTA that claims DIAG(0) runs the whole diagonal chain DIAG(0..wt-1); the
                     // updated tile (k+1,k+1) stays in shared memory between the steps (no DIAG(k > 0) list entries)
};
struct LargeTask {
  int lf;
  short type, k, i, j;  // type: 1 TRSM(i,k), 2 UPDATE(i,j,k), 3 DIAG(k), 4 UPDATE(i,j,[k,k1)), 5 INV(k),
                        //       6 EXTEND-ADD of update tile (i,j) of front `lf` into its parent front,
                        //       8 y_k = L_kk^-1 b_k, 9 b_i -= L(i,k) y_k, 10 update part of b -> the parent's b
  short k1, pad;
};
struct LargeJob {
  int lf, type, idx, c0, c1;  // type 0: copy idx; 1: child front idx, columns [c0,c1); 2: damping rows [c0,c1)
};
struct LargeLevel {
  int lf0, n_lf;   // range of large fronts of this level
  int t0, t1;      // task range
  int j0, j1;      // assembly job range
  int max_m;
  int max_nt;
  int solve_p;     // CTAs per front in the cooperative solves (grid = n_lf * solve_p <= #SMs)
};
struct LargeDev {
  const LargeFront* lf;
  const LargeTask* tasks;
  const LargeJob* jobs;
  int* counters;
  int* queue;  // one head per level
  double* linv;
  const int4* zero_jobs;  // large_zero_kernel: {large front, first column, end column, 0}
  double* binv;     // inverses of the eight 8x8 diagonal blocks of every L_kk (512 doubles per pivot tile): DIAG -> TRSMs
  int* sflags;      // solve flags / counters (zeroed before every solve)
  double* contrib;  // backward-solve contribution slots (v1 solves)
  uint4* ll_y;       // v2 solves: LL slots of the forward solution (elimination order)
  uint4* ll_contrib; // v2 solves: LL slots of the backward contributions (same indexing as contrib)
  double* fwd_b;     // fused forward substitution: right-hand sides of the fused fronts
};
void launch_large_level(cudaStream_t st, Ctrl* ctrl, const FrontDev& fd, const LargeDev& ld, const LargeLevel& lv,
                        int level, const double* sys_static, StatePtrs sp, int use_state_H, const double* dvec,
                        int grid_cap);
void launch_large_solve_fwd(cudaStream_t st, const Ctrl* ctrl, const FrontDev& fd, const LargeDev& ld,
                            const LargeLevel& lv, const double* rhs_static, StatePtrs sp, int use_state_rhs,
                            unsigned epoch);
void launch_large_solve_bwd(cudaStream_t st, const Ctrl* ctrl, const FrontDev& fd, const LargeDev& ld,
                            const LargeLevel& lv, unsigned epoch);
void launch_large_preassemble(cudaStream_t st, const Ctrl* ctrl, const FrontDev& fd, const LargeDev& ld,
                              const double* sys_static, StatePtrs sp, int use_state_H, const double* dvec, int pre_j0,
                              int pre_j1, int damp_j0, int damp_j1);
void launch_large_zero(cudaStream_t st, const Ctrl* ctrl, const FrontDev& fd, const LargeDev& ld, int n_jobs);
cudaError_t configure_large_kernels();
int large_factor_resident_ctas();  // CTAs of large_factor_kernel the device keeps resident (occupancy x SM count)
void launch_large_fused(cudaStream_t st, Ctrl* ctrl, const FrontDev& fd, const LargeDev& ld, int t0, int t1, int j0,
                        int j1, int queue_slot, const double* sys_static, StatePtrs sp, int use_state_H,
                        const double* dvec, int fwd);
void launch_large_fwd_init(cudaStream_t st, const Ctrl* ctrl, const FrontDev& fd, const LargeDev& ld, int lf0, int n_lf,
                           int first_fused_level, const int32_t* f_level, const double* rhs_static, StatePtrs sp,
                           int use_state_rhs);

// launchers (all asynchronous on `st`)
void launch_zero_lin(cudaStream_t st, const Ctrl* ctrl, StatePtrs sp, int mode, int64_t n_h, int n_rhs);
void launch_linearize(cudaStream_t st, const Ctrl* ctrl, StatePtrs sp, int mode, const LinBatch& b, double* partials);
void launch_jacobian(cudaStream_t st, const double* values, const LinBatch& b, const int32_t* jac_base,
                     const int32_t* jac_colnnz, double* out);
void launch_finish_error(cudaStream_t st, Ctrl* ctrl, int mode, const double* partials, int n_partials,
                         double* eager_err = nullptr);
// eager linearization of freshly uploaded values (mode 2, sfx_set_values): the BAL batch in CTA ranges, then the point
// sums and the error; launch_adopt_eager hands the result to the Init block of the next Optimize
constexpr int kLinModeEager = 2;
// the state block the eager linearization is written to: the Init block of the first iteration after a reset
// (reset_ctrl: init 0, new 1; lm_begin_kernel swaps the two at the start of every iteration)
constexpr int kEagerBlock = 1;
void launch_linearize_bal_range(cudaStream_t st, const Ctrl* ctrl, StatePtrs sp, int mode, const LinBatch& b, double* partials,
                                int block0, int block1);
void launch_bal_point_finalize(cudaStream_t st, const Ctrl* ctrl, StatePtrs sp, int mode, const LinBatch& b);
int linearize_bal_blocks(const LinBatch& b);
void launch_adopt_eager(cudaStream_t st, Ctrl* ctrl, const double* eager_err);
void launch_damping(cudaStream_t st, Ctrl* ctrl, StatePtrs sp, const int32_t* diag_pos, int N, double* dvec,
                    double* max_diag);
void launch_schur(cudaStream_t st, const Ctrl* ctrl, StatePtrs sp, const SchurDev& sd, const double* dvec);
void launch_schur_back(cudaStream_t st, const Ctrl* ctrl, StatePtrs sp, const SchurDev& sd, const double* y,
                       double* upd);
void launch_front_factor(cudaStream_t st, Ctrl* ctrl, const FrontDev& fd, const double* sysvals_static,
                         StatePtrs sp, int use_state_H, const double* dvec, int lvl_begin, int lvl_count,
                         int smem_m_max);
void launch_front_solve_fwd(cudaStream_t st, const Ctrl* ctrl, const FrontDev& fd, const double* rhs_static,
                            StatePtrs sp, int use_state_rhs, int lvl_begin, int lvl_count, int smem_bytes);
void launch_front_solve_bwd(cudaStream_t st, const Ctrl* ctrl, const FrontDev& fd, int lvl_begin, int lvl_count,
                            int smem_bytes);
void launch_unpermute(cudaStream_t st, const Ctrl* ctrl, const FrontDev& fd, double* out, double scale);
void launch_retract(cudaStream_t st, const Ctrl* ctrl, StatePtrs sp, const int32_t* key_type, const int32_t* key_voff,
                    const int32_t* key_sdim, const int32_t* key_tdim, const int32_t* key_itoff, int n_keys,
                    const double* upd);
void launch_step_reduce(cudaStream_t st, Ctrl* ctrl, StatePtrs sp, const double* upd, const double* dvec,
                        const double* last_upd, int N, double* partials, int i0);
void launch_pack_b(cudaStream_t st, Ctrl* ctrl, StatePtrs sp, int mode, int64_t nb, int nr, double* stage, int unpack);
void launch_commit_error(cudaStream_t st, Ctrl* ctrl, int mode);
void launch_mask_values(cudaStream_t st, const double* v, const unsigned char* mask, int64_t n, double* out);
void launch_lm_begin(cudaStream_t st, Ctrl* ctrl);
void launch_lm_after_first_linearize(cudaStream_t st, Ctrl* ctrl);
void launch_lm_end(cudaStream_t st, Ctrl* ctrl, const double* upd, double* last_upd, int N, int* host_done);
void launch_debug_snapshot(cudaStream_t st, const Ctrl* ctrl, StatePtrs sp, int first, int64_t n_values, int M, int cap,
                           double* dv, double* dr, const double* upd, const int32_t* ref2int, int N, double* du);
void launch_copy_values(cudaStream_t st, double* dst, const double* src, int64_t n);
void launch_export_csc(cudaStream_t st, const double* Hvals, const int32_t* csc_src, int64_t nnz, double* out);
cudaError_t configure_front_kernels(int smem_m_max, int max_front);
constexpr int kWarpFrontRows = 32;  // small fronts up to this many rows are factored by one warp each (kernels.cu)
void launch_import_csc(cudaStream_t st, const double* in, const int32_t* csc_src, int64_t nnz, double* Hvals);
void launch_set_unit(cudaStream_t st, double* v, int j, int prev);
void launch_set_entry(cudaStream_t st, double* v, int j, double value, int prev);  // v[prev] = 0 (prev >= 0), v[j] = value
void launch_permute_vec(cudaStream_t st, const double* in, const int32_t* ref2int, int N, double* out);

}  // namespace sfx
