// libsfx: problem object, device residency, LM driver and the C ABI of include/sfx.h.
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <nccl.h>

#include <algorithm>
#include <chrono>
#include <cmath>
#include <array>
#include <cstdio>
#include <map>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <limits>
#include <memory>
#include <queue>
#include <string>
#include <vector>

#include "kernels.cuh"
#include "sfx_internal.h"

namespace sfx {

extern int64_t g_launches;
extern int g_lin_skip;
void set_factor_trace(unsigned long long* buf);
void get_diag_stamps(unsigned long long* out);

#define CUDA_OK(expr)                                                                                   \
  do {                                                                                                  \
    cudaError_t e__ = (expr);                                                                           \
    if (e__ != cudaSuccess)                                                                             \
      throw Error(SFX_ERR_CUDA, std::string(#expr) + ": " + cudaGetErrorString(e__));                   \
  } while (0)

// Device allocations of one problem.  Every upload / memset is issued on the problem's own stream `st` and
// completed before returning: the library's streams are cudaStreamNonBlocking, so work on the legacy
// default stream (a plain cudaMemcpy from pageable memory returns once the data is STAGED, cudaMemset on
// device memory is asynchronous) is not ordered against the kernels that consume these buffers.
struct DevPool {
  std::vector<void*> ptrs;
  int64_t bytes = 0;
  cudaStream_t st = nullptr;
  template <typename T>
  T* alloc(size_t n) {
    void* p = nullptr;
    size_t b = std::max<size_t>(n, 1) * sizeof(T);
    CUDA_OK(cudaMalloc(&p, b));
    ptrs.push_back(p);
    bytes += (int64_t)b;
    return (T*)p;
  }
  template <typename T>
  T* upload(const std::vector<T>& v) {
    T* p = alloc<T>(v.size());
    if (!v.empty()) {
      CUDA_OK(cudaMemcpyAsync(p, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice, st));
      CUDA_OK(cudaStreamSynchronize(st));  // `v` may be a temporary; the DMA has landed when this returns
    }
    return p;
  }
  void zero(void* p, size_t nbytes) { CUDA_OK(cudaMemsetAsync(p, 0, nbytes, st)); }
  ~DevPool() {
    for (void* p : ptrs) cudaFree(p);
  }
};

enum Phase { PH_LIN = 0, PH_SCHUR, PH_FACTOR, PH_SOLVE, PH_UPDATE, PH_COUNT };

// NCCL is resolved at run time (dlopen) so that libsfx.so has no link-time dependency on it and
// shares the copy torch already loaded in a torchrun process.
struct NcclApi {
  void* h = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*Reduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*Broadcast)(const void*, void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*GroupStart)() = nullptr;
  ncclResult_t (*GroupEnd)() = nullptr;
  const char* (*GetErrorString)(ncclResult_t) = nullptr;
};
static NcclApi& nccl() {
  static NcclApi api;
  if (!api.h) {
    const char* names[] = {"libnccl.so.2", "libnccl.so"};
    for (const char* n : names) {
      api.h = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
      if (api.h) break;
    }
    if (!api.h) throw Error(SFX_ERR_NCCL, "cannot dlopen libnccl.so.2");
    auto sym = [&](const char* n) {
      void* f = dlsym(api.h, n);
      if (!f) throw Error(SFX_ERR_NCCL, std::string("NCCL symbol missing: ") + n);
      return f;
    };
    api.GetUniqueId = (decltype(api.GetUniqueId))sym("ncclGetUniqueId");
    api.CommInitRank = (decltype(api.CommInitRank))sym("ncclCommInitRank");
    api.CommDestroy = (decltype(api.CommDestroy))sym("ncclCommDestroy");
    api.AllReduce = (decltype(api.AllReduce))sym("ncclAllReduce");
    api.Reduce = (decltype(api.Reduce))sym("ncclReduce");
    api.Broadcast = (decltype(api.Broadcast))sym("ncclBroadcast");
    api.AllGather = (decltype(api.AllGather))sym("ncclAllGather");
    api.GroupStart = (decltype(api.GroupStart))sym("ncclGroupStart");
    api.GroupEnd = (decltype(api.GroupEnd))sym("ncclGroupEnd");
    api.GetErrorString = (decltype(api.GetErrorString))sym("ncclGetErrorString");
  }
  return api;
}
#define NCCL_OK(expr)                                                                          \
  do {                                                                                         \
    ncclResult_t r__ = (expr);                                                                 \
    if (r__ != ncclSuccess) throw Error(SFX_ERR_NCCL, std::string(#expr) + ": " + nccl().GetErrorString(r__)); \
  } while (0)

}  // namespace sfx

using namespace sfx;

struct sfx_comm {
  ncclComm_t comm = nullptr;
  int rank = 0, world = 1, device = 0;
};

struct sfx_problem {
  sfx_comm* comm = nullptr;  // not owned
  double* d_stage = nullptr;        // multi-GPU staging: [B | reduced rhs | err] / masked values
  unsigned char* d_vmask = nullptr; // which entries of the values buffer this rank contributes
  Analysis a;
  sfx_params params;
  double epsilon = 0;
  int device = 0;
  int ordering = 0;
  std::string err;
  DevPool pool;
  cudaStream_t st = nullptr;
  cudaStream_t st2 = nullptr;  // side stream: front zeroing overlaps damping + Schur
  cudaEvent_t ev_fork = nullptr, ev_join = nullptr, ev_fork2 = nullptr, ev_join2 = nullptr;
  int sm_count = 148;  // multiprocessors of the device (cudaDevAttrMultiProcessorCount at create)
  int plan_nd_depth = -1;  // plan chosen by choose_front_plan: -1 METIS_NodeND as the reference orders, 0 METIS_NodeND
                           // with capped cumulative amalgamation, d >= 2 dissect to depth d + sweep
  double ref_plan_flops = 0.0;  // factorization flops of the METIS_NodeND plan
  int fused_T0 = -1;  // first level of the fused top of the elimination tree (-1: none)
  int fused_t0 = 0, fused_t1 = 0, fused_j0 = 0, fused_j1 = 0;
  bool fused_fwd = false;       // the forward substitution of the fused fronts rides inside the factor kernel
  int32_t* d_f_level = nullptr; // front -> level (large_fwd_init_kernel)
  int n_large_fronts = 0;
  int n_zero_jobs = 0;
  unsigned solve_epoch = 0;
  // debug_stats: per-record snapshots of values / residual (allocated on first use)
  double *dbg_values = nullptr, *dbg_res = nullptr, *dbg_upd = nullptr;
  int dbg_cap = 0;
  bool dbg_valid = false;
  bool can_continue = false;  // the control block is the one the last sfx_optimize[_continue] left
  std::vector<std::pair<int64_t, int64_t>> opt_ranges;  // merged storage ranges of the optimized keys
  // multi-GPU: storage range [first, second) of the landmarks every rank owns, when each is one contiguous run of the
  // values buffer (BAL: points are stored in key order); empty otherwise (masked all-reduce fallback)
  std::vector<std::pair<int64_t, int64_t>> rank_lm_range;
  int64_t values_chunk = 0;  // multi-GPU: doubles of the values buffer every rank uploads (ceil(n_values / world))
  int pre_j0 = 0, pre_j1 = 0, damp_j0 = 0, damp_j1 = 0;  // assembly jobs run before level 0 (copies, damping)
  Ctrl* d_ctrl = nullptr;
  Ctrl* h_ctrl = nullptr;  // pinned
  int* h_done = nullptr;   // pinned + mapped
  int* d_done = nullptr;
  StatePtrs sp{};
  double* d_cur_values = nullptr;
  double *d_dvec = nullptr, *d_maxdiag = nullptr, *d_upd = nullptr, *d_last = nullptr, *d_y = nullptr;
  double* d_partials = nullptr;
  int n_partials = 0;
  int32_t* d_diag_pos = nullptr;
  int32_t *d_key_type = nullptr, *d_key_voff = nullptr, *d_key_sdim = nullptr, *d_key_tdim = nullptr,
          *d_key_itoff = nullptr;
  int32_t* d_ref2int = nullptr;
  std::vector<LinBatch> lin;
  SchurDev sd{};
  FrontDev fd{};
  std::vector<int> lvl_max_m;      // largest SMALL front per level
  std::vector<int> lvl_small_cnt;  // small fronts per level (listed first in level_fronts)
  std::vector<int> lvl_tiny_cnt;   // ... of which the first ones have at most 32 rows (one warp per front)
  std::vector<int> lvl_rest_max_m; // largest small front per level that is not tiny
  std::vector<LargeLevel> lvl_large;
  LargeDev ld{};
  int64_t n_counters = 0, n_sflags = 0;
  int small_max_m = 64;  // fronts with more rows go to the tile-DAG path
  int smem_cap_m = 168;
  // csc export
  int32_t* d_csc_src = nullptr;
  double* d_export = nullptr;
  int64_t export_cap = 0;
  std::vector<const int32_t*> d_jac_base, d_jac_colnnz;  // per batch, uploaded by ensure_jacobian
  // timing
  std::vector<cudaEvent_t> ev;
  std::vector<int> ev_phase;  // phase ending at event i (event 0 = start)
  sfx_timings tm{};
  sfx_stats last_stats{};
  bool values_set = false;
  // eager linearization (single GPU, one BAL fast-path batch whose pixel arguments are one ascending run of the values
  // buffer): sfx_set_values uploads the pixels in chunks and linearizes every chunk of observations as it lands -- the
  // Init linearization of the next sfx_optimize hides behind the upload, which then adopts it (adopt_eager_kernel)
  bool eager_ok = false;
  int64_t eager_pix_base = 0;
  double* d_eager_err = nullptr;
  bool eager_valid = false;     // block kEagerBlock holds the linearization of d_cur_values, untouched since
  bool lin0_clobbered = false;  // ... and therefore no longer the one the last Optimize left there
  std::vector<cudaEvent_t> ev_up;

  ~sfx_problem() {
    for (auto e : ev) cudaEventDestroy(e);
    for (auto e : ev_up) cudaEventDestroy(e);
    if (st) cudaStreamDestroy(st);
    if (st2) cudaStreamDestroy(st2);
    if (dbg_values) cudaFree(dbg_values);
    if (dbg_res) cudaFree(dbg_res);
    if (dbg_upd) cudaFree(dbg_upd);
    if (ev_fork) cudaEventDestroy(ev_fork);
    if (ev_join) cudaEventDestroy(ev_join);
    if (ev_fork2) cudaEventDestroy(ev_fork2);
    if (ev_join2) cudaEventDestroy(ev_join2);
    if (h_ctrl) cudaFreeHost(h_ctrl);
    if (h_done) cudaFreeHost(h_done);
  }
};

static thread_local std::string g_create_err;

namespace {

// Tile task list of one large front (see chol_large.cu for the task types).
void build_front_tasks(const LargeFront& x, int li, int Kc, std::vector<LargeTask>& tl) {
  const int wt = x.wt, nt = x.nt;
  // Tile (i, j), i >= j, needs the updates k in [0, min(j, wt)).  Those with k >= i - 2 are on or
  // next to the critical path and stay single-step tasks: DIAG(k) = POTRF(k) + TRSM(k+1,k) +
  // UPDATE(k+1,k+1,k) fused on one CTA, and the priority tasks TRSM(k+2,k), UPDATE(k+2,k+1,k),
  // UPDATE(k+2,k+2,k), with DIAG(k+1) hoisted in front of the bulk of step k.  All other
  // updates are range tasks UPDATE(i,j,[a,e)) with the accumulator kept in registers: cut at
  // multiples of Kc (panel boundaries) and at e = min(j, i-2, wt), emitted once TRSM(.,e-1) is.
  auto T = [&](int i, int k) { tl.push_back(LargeTask{li, 1, (short)k, (short)i, (short)k, 0, 0}); };
  auto U = [&](int i, int j, int k) { tl.push_back(LargeTask{li, 2, (short)k, (short)i, (short)j, 0, 0}); };
  auto R = [&](int i, int j, int a, int e) {
    if (e - a == 1)
      U(i, j, a);
    else
      tl.push_back(LargeTask{li, 4, (short)a, (short)i, (short)j, (short)e, 0});
  };
  auto cend = [&](int i, int j) { return std::min(std::min(j, i - 2), wt); };  // end of the coarse range
  struct Def {
    int need, i, j, a, e;
  };
  std::vector<Def> deferred;
  size_t dpos = 0;
  int d_steps_left = 0;
  tl.push_back(LargeTask{li, 3, 0, 0, 0, 0, 0});
  for (int k = 0; k < wt; ++k) {
    if (k + 2 < nt) {
      T(k + 2, k);
      U(k + 2, k + 1, k);
      U(k + 2, k + 2, k);
    }
    if (k + 1 < wt) tl.push_back(LargeTask{li, 3, (short)(k + 1), (short)(k + 1), (short)(k + 1), 0, 0});
    for (int i = k + 3; i < nt; ++i) T(i, k);
    tl.push_back(LargeTask{li, 5, (short)k, (short)k, (short)k, 0, 0});  // INV(k): L_kk^-1 for the solves
    // deferred range tasks of the last panel boundary: everything needed by this step, plus a share
    if (dpos < deferred.size()) {
      size_t upto = dpos + (deferred.size() - dpos + d_steps_left - 1) / std::max(1, d_steps_left);
      while (upto < deferred.size() && deferred[upto].need <= k) ++upto;  // sorted by need
      for (; dpos < upto; ++dpos) R(deferred[dpos].i, deferred[dpos].j, deferred[dpos].a, deferred[dpos].e);
      if (d_steps_left > 1) --d_steps_left;
    }
    const int a0 = (k / Kc) * Kc;
    const int e = k + 1;
    // final pieces: column k+1 (rows >= k+3) and row k+3
    if (k + 1 < nt)
      for (int i = k + 3; i < nt; ++i)
        if (cend(i, k + 1) == e) R(i, k + 1, a0, e);
    if (k + 3 < nt)
      for (int j = k + 2; j <= k + 3; ++j)
        if (cend(k + 3, j) == e) R(k + 3, j, a0, e);
    // panel boundary (or last pivot step): one piece [a0, e) for every other tile that has not
    // reached the end of its coarse range (at the last step these are final pieces)
    const bool last = k == wt - 1;
    if (e % Kc == 0 || last) {
      std::vector<Def> batch;
      for (int j = k + 2; j < nt; ++j)
        for (int i = std::max(j, k + 4); i < nt; ++i) batch.push_back(Def{std::min(j - 1, i - 3), i, j, a0, e});
      // anything still deferred from the previous boundary goes first (same tiles, earlier pieces)
      for (; dpos < deferred.size(); ++dpos)
        R(deferred[dpos].i, deferred[dpos].j, deferred[dpos].a, deferred[dpos].e);
      deferred.clear();
      dpos = 0;
      if (last) {
        for (auto& d : batch) R(d.i, d.j, d.a, d.e);
      } else {
        std::stable_sort(batch.begin(), batch.end(), [](const Def& p, const Def& q) { return p.need < q.need; });
        deferred.swap(batch);
        d_steps_left = Kc;
      }
    }
  }
  for (; dpos < deferred.size(); ++dpos) R(deferred[dpos].i, deferred[dpos].j, deferred[dpos].a, deferred[dpos].e);
}

// Replays one front's task list sequentially against the tile version counters: every wait condition
// of chol_large.cu must already hold when its task comes up (so in-order claiming cannot deadlock) and
// every tile must end up final.
void verify_task_list(const LargeFront& x, const std::vector<LargeTask>& tl) {
  const int nt = x.nt, wt = x.wt;
  std::vector<int> cnt((size_t)nt * nt, 0);
  auto C = [&](int i, int j) -> int& { return cnt[(size_t)i * nt + j]; };
  auto fail = [&](const LargeTask& t, const char* why) {
    throw Error(SFX_ERR_INVALID_ARG, std::string("internal: tile task list invalid (") + why + ") type " +
                                         std::to_string(t.type) + " k " + std::to_string(t.k) + " i " +
                                         std::to_string(t.i) + " j " + std::to_string(t.j) + " k1 " + std::to_string(t.k1));
  };
  for (const LargeTask& t : tl) {
    const int k = t.k, i = t.i, j = t.j;
    if (t.type == 3) {
      if (C(k, k) != k) fail(t, "diag not ready");
      C(k, k) = k + 1;
      if (k + 1 < nt) {
        if (C(k + 1, k) != k) fail(t, "diag trsm");
        C(k + 1, k) = k + 1;
        if (C(k + 1, k + 1) != k) fail(t, "diag update");
        C(k + 1, k + 1) = k + 1;
      }
    } else if (t.type == 5) {
      if (C(k, k) < k + 1) fail(t, "inv");
    } else if (t.type == 1) {
      if (C(k, k) < k + 1 || C(i, k) != k) fail(t, "trsm");
      C(i, k) = k + 1;
    } else if (t.type == 2) {
      if (C(i, k) < k + 1 || C(j, k) < k + 1 || C(i, j) != k) fail(t, "update");
      C(i, j) = k + 1;
    } else if (t.type == 4) {
      for (int kk = k; kk < t.k1; ++kk)
        if (C(i, kk) < kk + 1 || C(j, kk) < kk + 1) fail(t, "range operands");
      if (C(i, j) != k) fail(t, "range target");
      C(i, j) = t.k1;
    } else {
      fail(t, "type");
    }
  }
  for (int j = 0; j < nt; ++j)
    for (int i = j; i < nt; ++i) {
      const int want = j < wt ? j + 1 : wt;
      if (C(i, j) != want)
        throw Error(SFX_ERR_INVALID_ARG, "internal: tile (" + std::to_string(i) + "," + std::to_string(j) +
                                             ") ends at version " + std::to_string(C(i, j)) + ", expected " +
                                             std::to_string(want));
    }
}

// ---- fused schedule of the top of the elimination tree -------------------------------------------------------
// All large fronts from level T0 up run in ONE launch of the tile-DAG kernel: the children's update matrices are
// added by EXTEND-ADD tasks (type 6, one per update tile) and a front's DIAG(0) waits for them, so a parent's
// diagonal chain overlaps the trailing updates of its siblings' subtrees instead of idling behind a level barrier.
// The task order is the start order of a list schedule (critical-path priority, `workers` CTAs, task durations
// from the task trace of round 1) simulated here once; CTAs claim tasks in that order and spin on tile versions,
// which cannot deadlock because every dependency starts -- hence sits -- earlier in the list.
struct FusedDur {
  // us, from the task trace of round 1 (profiles/r01_results.md); DIAG = POTRF | TRSM(k+1,k) | UPDATE(k+1,k+1,k).
  // (The round-2 trace has shorter tasks -- TRSM 6.8, UPDATE 7.9, RANGE step 5.3, INV 12.7, EXTEND-ADD 5.5 -- but the
  // list these priorities produce with the round-2 numbers ran slower: METIS plan 4.13 vs 3.94 ms, measured.)
  double potrf = 14.0, diag_trsm = 5.0, diag_syrk = 5.0, trsm = 7.7, update = 9.0, range_step = 6.0, range_fix = 2.0,
         inv = 16.0, ea = 3.0, vsolve = 2.5, gemv = 2.5, veav = 3.0, sticky_gain = 5.0;
};
// Returns the modelled span of the launch (us).
double build_fused_schedule(const std::vector<LargeFront>& lfs, int lf_begin, int lf_end, int Kc, int workers, bool with_fwd,
                            std::vector<LargeTask>& out) {
  struct Edge {
    int to;
    double lag;  // the successor may start `lag` after this task started (output offset - input offset)
  };
  struct Node {
    LargeTask t;
    double dur, bl, ready;
    std::vector<Edge> succ;
    int indeg;
    int chain = 0;  // sticky chain: 1 = DIAG(0) (takes a CTA and keeps it), 2 = DIAG(k > 0) (runs on that CTA, no list
                    // entry), +4 = last step of the chain (its end frees the CTA)
  };
  struct Writer {
    int id = -1;
    double out = 0.0;  // offset into the writer task at which the tile is published
  };
  FusedDur D;
  if (const char* e = getenv("SFX_SIM_POTRF_US")) D.potrf = atof(e);
  if (const char* e = getenv("SFX_SIM_DIAG_TAIL_US")) D.diag_trsm = D.diag_syrk = 0.5 * atof(e);
  if (const char* e = getenv("SFX_SIM_RANGE_STEP_US")) D.range_step = atof(e);
  std::vector<Node> g;
  std::vector<std::vector<int>> ea_of(lfs.size());  // per parent large front: its extend-add tasks
  std::vector<std::vector<int>> veav_of(lfs.size());  // ... and the vector extend-adds of the forward substitution
  auto add_dep = [&](const Writer& w, int to, double in_off) {
    if (w.id < 0) return;
    g[w.id].succ.push_back(Edge{to, w.out - in_off});
    g[to].indeg++;
  };
  for (int li = lf_begin; li < lf_end; ++li) {
    const LargeFront& x = lfs[li];
    std::vector<LargeTask> tl;
    build_front_tasks(x, li, Kc, tl);
    verify_task_list(x, tl);
    const int nt = x.nt, wt = x.wt;
    std::vector<Writer> writer((size_t)nt * nt);
    auto Wr = [&](int i, int j) -> Writer& { return writer[(size_t)i * nt + j]; };
    for (const LargeTask& t : tl) {
      const int id = (int)g.size();
      const int k = t.k, i = t.i, j = t.j;
      g.push_back(Node{t, 0.0, 0.0, 0.0, {}, 0});
      switch (t.type) {
        case 3: {
          const bool more = k + 1 < nt;
          g[id].dur = D.potrf + (more ? D.diag_trsm + D.diag_syrk : 0.0);
          if (x.sticky) {
            g[id].chain = (k == 0 ? 1 : 2) | (k == wt - 1 ? 4 : 0);
            if (k > 0) g[id].dur -= D.sticky_gain;  // no store + publish + acquire + reload of tile (k,k)
          }
          if (k == 0)
            for (int e : ea_of[li]) add_dep(Writer{e, g[e].dur}, id, 0.0);
          add_dep(Wr(k, k), id, 0.0);
          Wr(k, k) = Writer{id, D.potrf};
          if (more) {
            add_dep(Wr(k + 1, k), id, D.potrf);
            add_dep(Wr(k + 1, k + 1), id, D.potrf + D.diag_trsm);
            Wr(k + 1, k) = Writer{id, D.potrf + D.diag_trsm};
            Wr(k + 1, k + 1) = Writer{id, g[id].dur};
          }
          break;
        }
        case 5:
          g[id].dur = D.inv;
          add_dep(Wr(k, k), id, 0.0);
          break;
        case 1:
          g[id].dur = D.trsm;
          add_dep(Wr(k, k), id, 0.0);
          add_dep(Wr(i, k), id, 0.0);
          Wr(i, k) = Writer{id, g[id].dur};
          break;
        case 2:
          g[id].dur = D.update;
          add_dep(Wr(i, k), id, 0.0);
          if (j != i) add_dep(Wr(j, k), id, 0.0);
          add_dep(Wr(i, j), id, D.update - D.range_fix);
          Wr(i, j) = Writer{id, g[id].dur};
          break;
        case 4:
          g[id].dur = D.range_fix + D.range_step * (t.k1 - k);
          for (int kk = k; kk < t.k1; ++kk) {  // operands are consumed one pivot step at a time
            add_dep(Wr(i, kk), id, D.range_step * (kk - k));
            if (j != i) add_dep(Wr(j, kk), id, D.range_step * (kk - k));
          }
          add_dep(Wr(i, j), id, g[id].dur - D.range_fix);
          Wr(i, j) = Writer{id, g[id].dur};
          break;
        default: throw Error(SFX_ERR_INVALID_ARG, "internal: unknown tile task type");
      }
    }
    if (x.parent_lf >= 0)
      for (int j = wt; j < nt; ++j)
        for (int i = j; i < nt; ++i) {
          const int id = (int)g.size();
          g.push_back(Node{LargeTask{li, 6, 0, (short)i, (short)j, 0, 0}, D.ea, 0.0, 0.0, {}, 0});
          add_dep(Wr(i, j), id, 0.0);
          ea_of[x.parent_lf].push_back(id);
        }
    if (with_fwd) {
      // forward substitution riding behind the factorization: y_k = L_kk^-1 b_k (8), b_i -= L(i,k) y_k (9), update
      // rows of b into the parent (10)
      std::vector<int> inv_of(wt, -1);
      for (int id = 0; id < (int)g.size(); ++id)
        if (g[id].t.lf == li && g[id].t.type == 5) inv_of[g[id].t.k] = id;
      std::vector<std::vector<int>> into(nt);  // GEMV tasks that update row tile i
      std::vector<int> gemv_update;            // ... those of the update rows
      for (int k = 0; k < wt; ++k) {
        const int vs = (int)g.size();
        g.push_back(Node{LargeTask{li, 8, (short)k, (short)k, (short)k, 0, 0}, D.vsolve, 0.0, 0.0, {}, 0});
        add_dep(Writer{inv_of[k], D.inv}, vs, 0.0);
        for (int e : into[k]) add_dep(Writer{e, D.gemv}, vs, 0.0);
        if (k == 0)
          for (int e : veav_of[li]) add_dep(Writer{e, D.veav}, vs, 0.0);
        for (int i = k + 1; i < nt; ++i) {
          const int id = (int)g.size();
          g.push_back(Node{LargeTask{li, 9, (short)k, (short)i, (short)k, 0, 0}, D.gemv, 0.0, 0.0, {}, 0});
          add_dep(Writer{vs, D.vsolve}, id, 0.0);
          add_dep(Wr(i, k), id, 0.0);  // the final writer of L(i,k): TRSM(i,k) or DIAG(k)
          into[i].push_back(id);
          if (i >= wt) gemv_update.push_back(id);
        }
      }
      if (x.parent_lf >= 0) {
        const int id = (int)g.size();
        g.push_back(Node{LargeTask{li, 10, 0, 0, 0, 0, 0}, D.veav, 0.0, 0.0, {}, 0});
        for (int e : gemv_update) add_dep(Writer{e, D.gemv}, id, 0.0);
        if (gemv_update.empty() && wt > 0) add_dep(Writer{id - 1, D.vsolve}, id, 0.0);
        veav_of[x.parent_lf].push_back(id);
      }
    }
  }
  // bottom levels (construction order is topological: fronts by ascending level, per-front lists valid)
  const int n = (int)g.size();
  double cp = 0.0;
  for (int id = n - 1; id >= 0; --id) {
    double m = g[id].dur;
    for (const Edge& e : g[id].succ) m = std::max(m, e.lag + g[e.to].bl);
    g[id].bl = m;
    cp = std::max(cp, m);
  }
  // list scheduling: a free CTA takes the ready task with the longest path to the end
  using Ev = std::pair<double, int>;
  std::priority_queue<Ev, std::vector<Ev>, std::greater<Ev>> running;  // (finish time, task)
  std::priority_queue<Ev, std::vector<Ev>, std::greater<Ev>> avail;    // (ready time, task): every predecessor started
  std::priority_queue<Ev> ready;                                        // (bottom level, task): ready now
  for (int id = 0; id < n; ++id)
    if (g[id].indeg == 0) avail.push({0.0, id});
  int free_w = std::max(1, workers);
  double now = 0.0;
  out.reserve(out.size() + n);
  int started = 0;
  auto start_task = [&](int id, double at) {
    ++started;
    for (const Edge& e : g[id].succ) {
      g[e.to].ready = std::max(g[e.to].ready, at + e.lag);
      if (--g[e.to].indeg == 0) avail.push({g[e.to].ready, e.to});
    }
  };
  std::priority_queue<Ev, std::vector<Ev>, std::greater<Ev>> chain_end;  // ends of the last steps of sticky chains
  const bool util_bins = getenv("SFX_TIMING") && atoi(getenv("SFX_TIMING")) >= 2;  // modelled utilisation per 100 us
  std::vector<double> util;
  std::vector<std::array<double, 11>> util_type;
  std::vector<std::map<int, double>> util_lf;
  while (started < n) {
    bool moved = false;
    while (!avail.empty() && avail.top().first <= now) {
      const int id = avail.top().second;
      avail.pop();
      if (g[id].chain & 2) {
        // a later step of a sticky chain: runs on the chain's own CTA as soon as it is ready, no list entry
        start_task(id, now);
        if (g[id].chain & 4) chain_end.push({now + g[id].dur, id});
        moved = true;
      } else {
        ready.push({g[id].bl, id});
      }
    }
    while (!chain_end.empty() && chain_end.top().first <= now) {
      chain_end.pop();
      ++free_w;
      moved = true;
    }
    if (moved) continue;
    if (free_w > 0 && !ready.empty()) {
      const int id = ready.top().second;
      ready.pop();
      out.push_back(g[id].t);
      if (util_bins) {
        for (double t = now; t < now + g[id].dur; t += 10.0) {
          const size_t b = (size_t)(t / 100.0);
          if (b >= util.size()) util.resize(b + 1, 0.0);
          util[b] += std::min(10.0, now + g[id].dur - t);
          if (b >= util_type.size()) util_type.resize(b + 1);
          util_type[b][g[id].t.type] += std::min(10.0, now + g[id].dur - t);
          if (b >= util_lf.size()) util_lf.resize(b + 1);
          util_lf[b][g[id].t.lf] += std::min(10.0, now + g[id].dur - t);
        }
      }
      // DIAG(0) of a sticky chain keeps its CTA until the chain's last step ends (a one-step chain: like any task)
      if ((g[id].chain & 1) && !(g[id].chain & 4))
        ;  // the CTA is released by chain_end
      else
        running.push({now + g[id].dur, id});
      --free_w;
      start_task(id, now);
      continue;
    }
    double next = 1e300;
    if (!running.empty() && free_w == 0) next = running.top().first;
    if (!avail.empty()) next = std::min(next, avail.top().first);
    if (!chain_end.empty()) next = std::min(next, chain_end.top().first);
    if (free_w > 0 && !running.empty() && avail.empty()) next = std::min(next, running.top().first);
    if (next >= 1e300) throw Error(SFX_ERR_INVALID_ARG, "internal: fused schedule has a dependency cycle");
    // workers whose task finished by `next` become free
    now = std::max(now, next);
    while (!running.empty() && running.top().first <= now) {
      running.pop();
      ++free_w;
    }
  }
  while (!running.empty()) {
    now = std::max(now, running.top().first);
    running.pop();
  }
  while (!chain_end.empty()) {
    now = std::max(now, chain_end.top().first);
    chain_end.pop();
  }
  if (util_bins) {
    std::fprintf(stderr, "[sfx analysis] modelled busy CTAs per 100 us:");
    for (double u : util) std::fprintf(stderr, " %.0f", u / 100.0);
    std::fprintf(stderr, "\n");
    if (atoi(getenv("SFX_TIMING")) >= 3)
      for (size_t b = 0; b < util_type.size(); ++b) {
        std::fprintf(stderr, "[sfx analysis]   bin %2zu: by type", b);
        for (int ty = 1; ty < 11; ++ty)
          if (util_type[b][ty] > 50.0) std::fprintf(stderr, " t%d=%.0f", ty, util_type[b][ty] / 100.0);
        std::fprintf(stderr, " | by front");
        for (auto& kv : util_lf[b])
          if (kv.second > 300.0) std::fprintf(stderr, " f%d(w%d)=%.0f", kv.first, lfs[kv.first].w, kv.second / 100.0);
        std::fprintf(stderr, "\n");
      }
  }
  if (getenv("SFX_TIMING")) {
    double busy[11] = {0};
    int cnt[11] = {0};
    for (int id = 0; id < n; ++id) {
      busy[g[id].t.type] += g[id].dur;
      cnt[g[id].t.type]++;
    }
    std::fprintf(stderr,
                 "[sfx analysis] fused schedule: %d tasks on %d CTAs, modelled span %.0f us (critical path %.0f us); CTA-ms busy: "
                 "TRSM %d/%.0f UPDATE %d/%.0f DIAG %d/%.0f RANGE %d/%.0f INV %d/%.0f EA %d/%.0f FWD %d/%.0f\n",
                 n, workers, now, cp, cnt[1], busy[1] / 1e3, cnt[2], busy[2] / 1e3, cnt[3], busy[3] / 1e3, cnt[4], busy[4] / 1e3,
                 cnt[5], busy[5] / 1e3, cnt[6], busy[6] / 1e3, cnt[8] + cnt[9] + cnt[10], (busy[8] + busy[9] + busy[10]) / 1e3);
  }
  return now;
}

// Replays a fused list against the tile version counters and the assembly counters: every wait condition of
// chol_large.cu must already hold when its task comes up, and every tile must end up final.
void verify_fused_list(const std::vector<LargeFront>& lfs, int lf_begin, int lf_end, const std::vector<LargeTask>& tl,
                       size_t t_begin) {
  std::vector<std::vector<int>> cnt(lfs.size());
  std::vector<int> assembled(lfs.size(), 0);
  // forward substitution: per front updates applied to row tile i, y_k published, L_kk^-1 ready, children vectors added
  std::vector<std::vector<int>> bcount(lfs.size()), ypub(lfs.size()), invd(lfs.size());
  std::vector<int> vasm(lfs.size(), 0), n_fwd(lfs.size(), 0);
  for (int li = lf_begin; li < lf_end; ++li) {
    bcount[li].assign(lfs[li].nt, 0);
    ypub[li].assign(lfs[li].wt, 0);
    invd[li].assign(lfs[li].wt, 0);
  }
  for (int li = lf_begin; li < lf_end; ++li) cnt[li].assign((size_t)lfs[li].nt * lfs[li].nt, 0);
  auto fail = [&](const LargeTask& t, const char* why) {
    throw Error(SFX_ERR_INVALID_ARG, std::string("internal: fused task list invalid (") + why + ") front " +
                                         std::to_string(t.lf) + " type " + std::to_string(t.type) + " k " +
                                         std::to_string(t.k) + " i " + std::to_string(t.i) + " j " + std::to_string(t.j));
  };
  // sticky chains: DIAG(0) activates the chain; its later steps run as soon as their inputs are there (in the kernel:
  // on the CTA that holds the chain, which spins), i.e. after whichever list task produced the last input
  std::vector<int> chain_k(lfs.size(), -1);
  std::vector<int> active;
  auto advance = [&](int li) {
    const LargeFront& x = lfs[li];
    const int nt = x.nt;
    auto C = [&](int i, int j) -> int& { return cnt[li][(size_t)i * nt + j]; };
    bool any = false;
    while (chain_k[li] >= 0 && chain_k[li] < x.wt) {
      const int kd = chain_k[li];
      if (kd == 0 && assembled[li] != x.n_ea) break;
      if (C(kd, kd) != kd) break;
      if (kd + 1 < nt && (C(kd + 1, kd) != kd || C(kd + 1, kd + 1) != kd)) break;
      C(kd, kd) = kd + 1;
      if (kd + 1 < nt) C(kd + 1, kd) = C(kd + 1, kd + 1) = kd + 1;
      chain_k[li] = kd + 1;
      any = true;
    }
    return any;
  };
  for (size_t q = t_begin; q < tl.size(); ++q) {
    const LargeTask& t = tl[q];
    if (t.lf < lf_begin || t.lf >= lf_end) fail(t, "front out of range");
    const LargeFront& x = lfs[t.lf];
    const int nt = x.nt;
    auto C = [&](int i, int j) -> int& { return cnt[t.lf][(size_t)i * nt + j]; };
    const int k = t.k, i = t.i, j = t.j;
    switch (t.type) {
      case 3:
        if (x.sticky) {
          if (k != 0) fail(t, "sticky front with a DIAG(k > 0) list entry");
          chain_k[t.lf] = 0;
          active.push_back(t.lf);
          if (!advance(t.lf)) fail(t, "sticky chain cannot start");
          break;
        }
        if (k == 0 && assembled[t.lf] != x.n_ea) fail(t, "front not assembled");
        if (C(k, k) != k) fail(t, "diag not ready");
        C(k, k) = k + 1;
        if (k + 1 < nt) {
          if (C(k + 1, k) != k) fail(t, "diag trsm");
          C(k + 1, k) = k + 1;
          if (C(k + 1, k + 1) != k) fail(t, "diag update");
          C(k + 1, k + 1) = k + 1;
        }
        break;
      case 5:
        if (C(k, k) < k + 1) fail(t, "inv");
        invd[t.lf][k] = 1;
        break;
      case 8:
        if (k == 0 && vasm[t.lf] != x.n_vch) fail(t, "forward: children's vectors missing");
        if (!invd[t.lf][k] || bcount[t.lf][k] != k) fail(t, "forward: y_k not ready");
        ypub[t.lf][k] = 1;
        n_fwd[t.lf]++;
        break;
      case 9:
        if (!ypub[t.lf][k] || C(i, k) < k + 1) fail(t, "forward: gemv operands");
        bcount[t.lf][i]++;
        n_fwd[t.lf]++;
        break;
      case 10:
        for (int q = x.wt; q < nt; ++q)
          if (bcount[t.lf][q] != x.wt) fail(t, "forward: update rows incomplete");
        if (x.parent_lf < lf_begin || x.parent_lf >= lf_end) fail(t, "forward: no fused parent");
        vasm[x.parent_lf]++;
        break;
      case 1:
        if (C(k, k) < k + 1 || C(i, k) != k) fail(t, "trsm");
        C(i, k) = k + 1;
        break;
      case 2:
        if (C(i, k) < k + 1 || C(j, k) < k + 1 || C(i, j) != k) fail(t, "update");
        C(i, j) = k + 1;
        break;
      case 4:
        for (int kk = k; kk < t.k1; ++kk)
          if (C(i, kk) < kk + 1 || C(j, kk) < kk + 1) fail(t, "range operands");
        if (C(i, j) != k) fail(t, "range target");
        C(i, j) = t.k1;
        break;
      case 6:
        if (x.parent_lf < lf_begin || x.parent_lf >= lf_end) fail(t, "extend-add without a fused parent");
        if (j < x.wt || C(i, j) != x.wt) fail(t, "extend-add of a tile that is not final");
        assembled[x.parent_lf]++;
        break;
      default: fail(t, "type");
    }
    // chain CTAs run concurrently with the list: let every active chain take the steps that became possible
    for (bool again = true; again;) {
      again = false;
      for (size_t a = 0; a < active.size(); ++a)
        if (advance(active[a])) again = true;
    }
    for (size_t a = 0; a < active.size();)
      if (chain_k[active[a]] >= lfs[active[a]].wt) {
        active[a] = active.back();
        active.pop_back();
      } else {
        ++a;
      }
  }
  if (!active.empty()) throw Error(SFX_ERR_INVALID_ARG, "internal: a sticky diagonal chain cannot finish");
  for (int li = lf_begin; li < lf_end; ++li) {
    const LargeFront& x = lfs[li];
    if (assembled[li] != x.n_ea) throw Error(SFX_ERR_INVALID_ARG, "internal: fused front misses extend-add tasks");
    if (n_fwd[li] > 0) {
      if (vasm[li] != x.n_vch) throw Error(SFX_ERR_INVALID_ARG, "internal: fused forward substitution misses a child vector");
      for (int k2 = 0; k2 < x.wt; ++k2)
        if (!ypub[li][k2]) throw Error(SFX_ERR_INVALID_ARG, "internal: fused forward substitution leaves a pivot tile unsolved");
    }
    for (int j = 0; j < x.nt; ++j)
      for (int i = j; i < x.nt; ++i)
        if (cnt[li][(size_t)i * x.nt + j] != (j < x.wt ? j + 1 : x.wt))
          throw Error(SFX_ERR_INVALID_ARG, "internal: fused list leaves a tile unfinished");
  }
}

// Host-side plan of the large-front path: which fronts go to the tile-DAG kernel, their tile task lists (per level,
// or ONE fused list from level fused_T0 up), assembly jobs and workspace offsets.  No device calls: `workers` is the
// number of CTAs of large_factor_kernel the device keeps resident (input of the list schedule).
struct LargeHostPlan {
  std::vector<int> lvl_fronts;
  std::vector<LargeFront> lfs;
  std::vector<LargeTask> tasks;
  std::vector<LargeJob> jobs, pre_jobs, damp_jobs;
  int64_t linv_off = 0, cnt_off = 0, flag_off = 0, contrib_off = 0, fwd_b_size = 0;
  double model_span_us = -1.0;  // modelled span of the fused launch (-1: no fused launch)
};
void plan_large_fronts(sfx_problem* p, int workers, LargeHostPlan& hp) {
  FrontPlan& f = p->a.fp;
  // ---- split every level into small fronts (one CTA each, in shared memory) and large fronts
  //      (tile-DAG kernel); small ones first in level_fronts
  if (const char* e = getenv("SFX_SMALL_MAX")) p->small_max_m = std::min(atoi(e), p->smem_cap_m);
  const int T = 64;
  std::vector<int>& lvl_fronts = hp.lvl_fronts;
  lvl_fronts = f.level_fronts;
  p->lvl_max_m.assign(f.n_levels, 0);
  p->lvl_small_cnt.assign(f.n_levels, 0);
  p->lvl_tiny_cnt.assign(f.n_levels, 0);
  p->lvl_rest_max_m.assign(f.n_levels, 0);
  p->lvl_large.assign(f.n_levels, LargeLevel{});
  std::vector<LargeFront>& lfs = hp.lfs;
  std::vector<LargeTask>& tasks = hp.tasks;
  std::vector<LargeJob>&jobs = hp.jobs, &pre_jobs = hp.pre_jobs, &damp_jobs = hp.damp_jobs;
  int64_t &linv_off = hp.linv_off, &cnt_off = hp.cnt_off, &flag_off = hp.flag_off, &contrib_off = hp.contrib_off;
  auto fits_small = [&](int s) { return f.f_w[s] + f.f_u[s] <= p->small_max_m; };
  // First level of the fused top (see build_fused_schedule).  Every front from there up runs as a tile-DAG front, also
  // the few that would fit the one-CTA-per-front kernel: a one-tile front costs ~45 us of CTA time there, so a budget
  // of them is cheaper than the five launches per level they would otherwise keep alive (the 100 k pose graph: levels
  // 3-17 fused with 567 such fronts instead of levels 8-17).
  int T0 = f.n_levels;
  if (!getenv("SFX_NO_FUSE")) {
    int budget = getenv("SFX_FUSE_SMALL_BUDGET") ? atoi(getenv("SFX_FUSE_SMALL_BUDGET")) : 600;
    while (T0 > 0) {
      int n_small = 0;
      for (int q = f.level_ptr[T0 - 1]; q < f.level_ptr[T0]; ++q) n_small += fits_small(f.level_fronts[q]) ? 1 : 0;
      if (n_small > budget) break;
      budget -= n_small;
      --T0;
    }
    if (T0 >= f.n_levels - 1) T0 = f.n_levels;  // a single level gains nothing
  }
  p->fused_T0 = T0 < f.n_levels ? T0 : -1;
  auto is_small = [&](int s) { return fits_small(s) && f.f_level[s] < T0; };
  std::vector<int> lf_of_front(f.n_fronts, -1);
  for (int l = 0; l < f.n_levels; ++l) {
    int* b = lvl_fronts.data() + f.level_ptr[l];
    int* e = lvl_fronts.data() + f.level_ptr[l + 1];
    const bool fused = l >= T0;
    int* mid = std::stable_partition(b, e, is_small);
    p->lvl_small_cnt[l] = (int)(mid - b);
    for (int* q = b; q < mid; ++q) p->lvl_max_m[l] = std::max(p->lvl_max_m[l], f.f_w[*q] + f.f_u[*q]);
    int* tiny_end = std::stable_partition(b, mid, [&](int s) { return f.f_w[s] + f.f_u[s] <= kWarpFrontRows; });
    p->lvl_tiny_cnt[l] = (int)(tiny_end - b);
    for (int* q = tiny_end; q < mid; ++q) p->lvl_rest_max_m[l] = std::max(p->lvl_rest_max_m[l], f.f_w[*q] + f.f_u[*q]);
    LargeLevel& lv = p->lvl_large[l];
    lv.lf0 = (int)lfs.size();
    lv.n_lf = (int)(e - mid);
    lv.t0 = (int)tasks.size();
    lv.j0 = (int)jobs.size();
    lv.max_m = 0;
    lv.max_nt = 0;
    int max_wt = 0;
    for (int* q = mid; q < e; ++q) {
      const int s = *q;
      LargeFront x{};
      x.off = f.f_off[s];
      x.m = f.f_w[s] + f.f_u[s];
      x.w = f.f_w[s];
      x.wt = (x.w + T - 1) / T;
      x.nt = x.wt + (f.f_u[s] + T - 1) / T;
      x.linv_off = linv_off;
      x.cnt_off = (int)cnt_off;
      x.front = s;
      x.flag_off = (int)flag_off;
      x.contrib_off = contrib_off;
      x.parent_lf = -1;
      x.n_ea = 0;
      x.asm_off = 0;
      x.fb_off = x.vc_off = x.n_vch = x.sticky = 0;
      lf_of_front[s] = (int)lfs.size();
      flag_off += 2 * x.wt;
      contrib_off += (int64_t)x.wt * x.nt * T;
      lv.max_nt = std::max(lv.max_nt, x.nt);
      linv_off += (int64_t)x.wt * T * T;
      cnt_off += (int64_t)x.nt * x.nt;
      SFX_CHECK(cnt_off < (int64_t)INT32_MAX, SFX_ERR_UNSUPPORTED, "too many tiles");
      lv.max_m = std::max(lv.max_m, x.m);
      max_wt = std::max(max_wt, x.wt);
      const int li = (int)lfs.size();
      lfs.push_back(x);
      // assembly jobs
      // system-matrix block copies and damping do not depend on the children: they run for all levels
      // at once before level 0 (copies as plain stores: every entry has one source block); only the
      // extend-add of the children stays between the levels
      for (int c = f.f_copy_ptr[s]; c < f.f_copy_ptr[s + 1]; ++c) pre_jobs.push_back(LargeJob{li, 0, c, 0, 0});
      for (int r = 0; r < x.w; r += 1024) damp_jobs.push_back(LargeJob{li, 2, 0, r, std::min(x.w, r + 1024)});
      for (int ci = f.f_child_ptr[s]; ci < f.f_child_ptr[s + 1]; ++ci) {
        const int c = f.f_child[ci];
        const int uc = f.f_u[c];
        if (fused && f.f_level[c] >= T0) continue;  // assembled by EXTEND-ADD tasks inside the fused launch
        int c0 = 0;
        while (c0 < uc) {
          int c1 = c0;
          int64_t el = 0;
          while (c1 < uc && el < 1024) {
            el += uc - c1;
            ++c1;
          }
          jobs.push_back(LargeJob{li, 1, c, c0, c1});
          c0 = c1;
        }
      }
    }
    // tasks: per front in an order that keeps every dependency earlier in the list (see
    // chol_large.cu), then merged across the fronts of the level in proportion to their work so that
    // the window of tasks in flight always spans all fronts (one front's critical path hides behind
    // the others' trailing updates)
    if (!fused) {
      const int Kc = getenv("SFX_KC") ? std::max(1, atoi(getenv("SFX_KC"))) : 6;
      std::vector<std::vector<LargeTask>> per(lv.n_lf);
      for (int q = 0; q < lv.n_lf; ++q) {
        // (shorter panels for narrow fronts were measured: slightly slower, 4.87 vs 4.82 ms at Final-shape)
        build_front_tasks(lfs[lv.lf0 + q], lv.lf0 + q, Kc, per[q]);
        verify_task_list(lfs[lv.lf0 + q], per[q]);
      }
      // proportional merge by cumulative cost
      auto cost = [](const LargeTask& t) { return t.type == 4 ? (double)(t.k1 - t.k) : t.type == 3 ? 3.0 : 1.0; };
      std::vector<double> tot(lv.n_lf, 0.0);
      double longest = 0;
      for (int q = 0; q < lv.n_lf; ++q) {
        for (auto& t : per[q]) tot[q] += cost(t);
        longest = std::max(longest, tot[q]);
      }
      std::vector<size_t> pos(lv.n_lf, 0);
      std::vector<double> done(lv.n_lf, 0.0);
      const int nsteps = 4096;
      for (int step = 1; step <= nsteps; ++step)
        for (int q = 0; q < lv.n_lf; ++q) {
          const double upto = tot[q] * step / nsteps;
          while (pos[q] < per[q].size() && done[q] < upto) {
            done[q] += cost(per[q][pos[q]]);
            tasks.push_back(per[q][pos[q]++]);
          }
        }
      for (int q = 0; q < lv.n_lf; ++q)
        while (pos[q] < per[q].size()) tasks.push_back(per[q][pos[q]++]);
      (void)longest;
    }
    lv.t1 = (int)tasks.size();
    lv.j1 = (int)jobs.size();
    // cooperative solves: P CTAs per front, one CTA per SM (their spin-waits need every CTA of a launch resident);
    // the SM count comes from the device, a few SMs are left free
    lv.solve_p = lv.n_lf > 0 ? std::max(1, std::min(lv.max_nt, std::max(1, p->sm_count - 4) / lv.n_lf)) : 1;
    // (P == 1: a front's only CTA waits for nobody, so levels with more fronts than SMs are fine)
    SFX_CHECK(lv.solve_p == 1 || lv.n_lf * lv.solve_p <= std::max(1, p->sm_count), SFX_ERR_UNSUPPORTED,
              "internal: cooperative solve grid exceeds the SM count");
  }
  if (T0 < f.n_levels) {
    // fused top: parents, extend-add counts and assembly counters, then the list schedule
    const int lf_begin = p->lvl_large[T0].lf0, lf_end = (int)lfs.size();
    for (int li = lf_begin; li < lf_end; ++li) {
      const int par = f.f_parent[lfs[li].front];
      if (par < 0) continue;
      const int pl = lf_of_front[par];
      SFX_CHECK(pl >= lf_begin, SFX_ERR_INVALID_ARG, "internal: parent of a fused front is not fused");
      lfs[li].parent_lf = pl;
      const int ntu = lfs[li].nt - lfs[li].wt;
      lfs[pl].n_ea += ntu * (ntu + 1) / 2;
    }
    for (int li = lf_begin; li < lf_end; ++li) lfs[li].asm_off = (int)cnt_off++;
    // fused forward substitution: right-hand sides and counters of the fused fronts
    p->fused_fwd = !getenv("SFX_NO_FUSED_FWD");
    int64_t fb = 0;
    for (int li = lf_begin; li < lf_end; ++li) {
      lfs[li].fb_off = (int)fb;
      fb += lfs[li].m;
      lfs[li].vc_off = (int)cnt_off;
      cnt_off += lfs[li].nt + 2 * lfs[li].wt + 1;
      if (p->fused_fwd && lfs[li].parent_lf >= 0) lfs[lfs[li].parent_lf].n_vch++;
    }
    SFX_CHECK(fb < (int64_t)INT32_MAX, SFX_ERR_UNSUPPORTED, "fused fronts too large");
    hp.fwd_b_size = fb;
    SFX_CHECK(cnt_off < (int64_t)INT32_MAX, SFX_ERR_UNSUPPORTED, "too many tiles");
    // sticky diagonal chains: the fronts with the longest chains, at most a quarter of the resident CTAs (a chain CTA
    // waits for tasks that come later in the list, so enough other CTAs must stay free to claim them)
    if (!getenv("SFX_NO_STICKY")) {
      std::vector<int> cand;
      for (int li = lf_begin; li < lf_end; ++li)
        if (lfs[li].wt >= 6) cand.push_back(li);
      std::sort(cand.begin(), cand.end(), [&](int x, int y) { return lfs[x].wt != lfs[y].wt ? lfs[x].wt > lfs[y].wt : x < y; });
      const size_t cap = (size_t)std::max(1, workers / 4);
      for (size_t q = 0; q < cand.size() && q < cap; ++q) lfs[cand[q]].sticky = 1;
    }
    const int Kc = getenv("SFX_KC") ? std::max(1, atoi(getenv("SFX_KC"))) : 6;
    p->fused_t0 = (int)tasks.size();
    hp.model_span_us = build_fused_schedule(lfs, lf_begin, lf_end, Kc, workers, p->fused_fwd, tasks);
    verify_fused_list(lfs, lf_begin, lf_end, tasks, (size_t)p->fused_t0);
    p->fused_t1 = (int)tasks.size();
    p->fused_j0 = p->lvl_large[T0].j0;
    p->fused_j1 = (int)jobs.size();
  }
}

// Ordering + front plan of the system the Cholesky factors.  The reference's ordering (METIS_NodeND, symbolic.cc) is
// always planned; for systems of a few thousand block variables whose whole elimination tree runs as ONE fused
// tile-DAG launch, "dissect to depth d, then sweep" candidates are planned as well and the plan with the shortest
// modelled launch (the list-schedule model of build_fused_schedule, calibrated against task traces) is kept.  Any
// fill-reducing permutation gives the same factorization up to rounding; only the time differs.  SFX_ORDERING_SEARCH=0
// keeps the reference's ordering.
void choose_front_plan(sfx_problem* p, const BlockMatrix& sys, int ordering, const std::vector<int>& sys2ref, int workers) {
  FrontPlan& fp = p->a.fp;
  if (sys.n_nodes > 8192 && !getenv("SFX_RELAX")) {
    // beyond the range of the search below: the reference's ordering with the cumulative amalgamation criterion
    // (explicit zeros <= 10 % of the merged front over everything a supernode has absorbed; measured on the 100 k pose
    // graph: 7.39 instead of 7.87 GFLOP, 4.91 instead of 5.08 ms per iteration)
    PlanOptions opt;
    opt.cumulative = true;
    opt.relax = 0.10;
    build_front_plan(sys, ordering, sys2ref, fp, opt);
  } else {
    build_front_plan(sys, ordering, sys2ref, fp);
  }
  p->plan_nd_depth = -1;
  p->ref_plan_flops = fp.flops;
  const char* e = getenv("SFX_ORDERING_SEARCH");
  if ((e && atoi(e) == 0) || ordering == SFX_ORDERING_NATURAL || getenv("SFX_ND_DEPTH")) return;
  if (sys.n_nodes < 256 || sys.n_nodes > 8192 || workers <= 0) return;
  auto model = [&](FrontPlan& f, double& span) {
    sfx_problem tmp;
    tmp.sm_count = p->sm_count;
    tmp.small_max_m = p->small_max_m;
    tmp.smem_cap_m = p->smem_cap_m;
    tmp.a.fp = std::move(f);
    LargeHostPlan hp;
    plan_large_fronts(&tmp, workers, hp);
    f = std::move(tmp.a.fp);
    span = hp.model_span_us;
    return tmp.fused_T0 == 0 && span > 0.0;  // the model covers a fully fused tree only
  };
  double best = 0.0;
  if (!model(fp, best)) return;
  const double base = best;
  for (int depth = 1; depth <= 4; ++depth) {
    PlanOptions opt;
    opt.nd_depth = depth == 1 ? -1 : depth;  // first candidate: METIS_NodeND with the amalgamation rules of the others
    opt.relax = 0.10;
    opt.cumulative = true;
    opt.max_merge_w = 1024;
    FrontPlan cand;
    build_front_plan(sys, ordering, sys2ref, cand, opt);
    double span = 0.0;
    if (model(cand, span) && span < 0.97 * best) {
      best = span;
      fp = std::move(cand);
      p->plan_nd_depth = opt.nd_depth < 0 ? 0 : depth;
    }
  }
  if (getenv("SFX_TIMING"))
    std::fprintf(stderr, "[sfx analysis] ordering: modelled factor launch %.0f us with METIS_NodeND, %.0f us chosen (%s)\n",
                 base, best, p->plan_nd_depth < 0    ? "METIS_NodeND"
                 : p->plan_nd_depth == 0 ? "METIS_NodeND, capped amalgamation"
                                         : ("dissect to depth " + std::to_string(p->plan_nd_depth) + ", then sweep").c_str());
}

void upload_structures(sfx_problem* p) {
  Analysis& a = p->a;
  DevPool& P = p->pool;
  {
    // storage ranges of the optimized keys, merged (BAL: cameras and points are a handful of runs); used by
    // sfx_update_best_values
    std::vector<std::pair<int64_t, int64_t>> r;
    for (const auto& k : a.keys) r.emplace_back((int64_t)k.voff, (int64_t)k.voff + k.sdim);
    std::sort(r.begin(), r.end());
    for (const auto& x : r) {
      if (!p->opt_ranges.empty() && x.first <= p->opt_ranges.back().second)
        p->opt_ranges.back().second = std::max(p->opt_ranges.back().second, x.second);
      else
        p->opt_ranges.push_back(x);
    }
  }
  // keys
  std::vector<int32_t> kt, kv, ks, kd, ki;
  for (auto& k : a.keys) {
    kt.push_back(k.type);
    kv.push_back(k.voff);
    ks.push_back(k.sdim);
    kd.push_back(k.tdim);
    ki.push_back(a.nodes[k.node].toff + k.sub);
  }
  p->d_key_type = P.upload(kt);
  p->d_key_voff = P.upload(kv);
  p->d_key_sdim = P.upload(ks);
  p->d_key_tdim = P.upload(kd);
  p->d_key_itoff = P.upload(ki);
  p->d_ref2int = P.upload(std::vector<int32_t>(a.ref2int.begin(), a.ref2int.end()));
  p->d_diag_pos = P.upload(a.diag_pos);
  // batches
  int partial_base = 0;
  for (auto& bp : a.batches) {
    LinBatch lb{};
    lb.kind = bp.kind;
    lb.n = bp.n;
    lb.arg_off = P.upload(bp.arg_off);
    lb.res_off = P.upload(bp.res_off);
    lb.rhs_off = P.upload(bp.rhs_off);
    lb.diag_off = P.upload(bp.diag_off);
    lb.off_off = P.upload(bp.off_off);
    for (int i = 0; i < 3; ++i) {
      lb.key_group[i] = i < bp.n_opt ? bp.key_group[i] : -1;
      lb.key_sub[i] = i < bp.n_opt ? bp.key_sub[i] : 0;
      lb.group_dim[i] = i < bp.n_groups ? bp.group_dim[i] : 0;
    }
    lb.n_groups = bp.n_groups;
    // BAL fast path: camera pose + intrinsics merged into one 9-dim node, point node of dim 3,
    // every point-camera block stored untransposed (landmark nodes trail the camera nodes)
    lb.bal_fast = 0;
    if (bp.kind == SFX_KIND_SNAVELY && bp.n_groups == 2 && bp.key_group[0] == 0 && bp.key_group[1] == 0 &&
        bp.key_group[2] == 1 && bp.key_sub[0] == 0 && bp.key_sub[1] == 6 && bp.key_sub[2] == 0 &&
        bp.group_dim[0] == 9 && bp.group_dim[1] == 3 && !getenv("SFX_NO_BAL_FAST")) {
      bool ok = true;
      for (uint32_t v : bp.off_off)
        if (v & kOffTransposed) {
          ok = false;
          break;
        }
      lb.bal_fast = ok ? 1 : 0;
      if (ok && !getenv("SFX_POINT_ATOMICS")) {
        // point -> observation slots (points identified by the offset of their diagonal block)
        std::vector<int32_t> order, ptr, diag, rhs;
        build_point_lists(bp, order, ptr, diag, rhs);
        lb.n_pf = (int)diag.size();
        lb.pf_exclusive = (a.batches.size() == 1 && getenv("SFX_PF_STORES")) ? 1 : 0;  // measured: plain stores are ~4% slower than REDs here
        lb.pf_ptr = P.upload(ptr);
        lb.pf_slot = P.upload(order);
        lb.pf_diag = P.upload(diag);
        lb.pf_rhs = P.upload(rhs);
        lb.pbuf = P.alloc<double>((size_t)((bp.n + 127) / 128) * 128 * 8);  // kPbufStride (kernels.cu); 256-byte aligned
      }
    }
    lb.partial_base = partial_base;
    partial_base += (bp.n + 127) / 128;
    p->lin.push_back(lb);
  }
  p->n_partials = partial_base;
  p->d_partials = P.alloc<double>(std::max(partial_base, 4 * 296));
  // state
  for (int b = 0; b < 3; ++b) {
    p->sp.values[b] = P.alloc<double>(a.n_values);
    p->sp.H[b] = P.alloc<double>(a.H.n_values + 2);  // (+2: a bulk copy of an odd tail reads 8 bytes further)
    p->sp.rhs[b] = P.alloc<double>(a.N);
    p->sp.res[b] = P.alloc<double>(a.M);
  }
  p->values_chunk = (a.n_values + a.world - 1) / a.world;
  p->d_cur_values = P.alloc<double>(p->values_chunk * a.world);
  if (a.world > 1 && (int)a.lm_rank_begin.size() == a.world + 1) {
    bool ok = true;
    for (int r = 0; r < a.world && ok; ++r) {
      int64_t lo = INT64_MAX, hi = -1, tot = 0;
      for (int k = a.lm_rank_begin[r]; k < a.lm_rank_begin[r + 1]; ++k) {
        lo = std::min<int64_t>(lo, a.keys[k].voff);
        hi = std::max<int64_t>(hi, (int64_t)a.keys[k].voff + a.keys[k].sdim);
        tot += a.keys[k].sdim;
      }
      if (tot == 0) lo = hi = 0;
      ok = tot == hi - lo;  // contiguous: nothing else lives inside the run
      p->rank_lm_range.emplace_back(lo, hi);
    }
    if (!ok) p->rank_lm_range.clear();
  }
  p->d_dvec = P.alloc<double>(a.N);
  p->d_maxdiag = P.alloc<double>(a.N);
  p->d_upd = P.alloc<double>(a.N);
  p->d_last = P.alloc<double>(a.N);
  p->d_y = P.alloc<double>(a.N);
  P.zero(p->d_last, sizeof(double) * a.N);
  P.zero(p->d_upd, sizeof(double) * a.N);  // entries of other ranks' landmarks stay 0
  if (a.world > 1) {
    p->d_stage = P.alloc<double>(std::max<int64_t>(a.n_values, a.b_values + a.sp.reduced_dim + 1));
    std::vector<unsigned char> mask(a.n_values, a.rank == 0 ? 1 : 0);
    const int first_lm_key = (int)a.keys.size() - a.sp.n_landmarks_total;
    for (int k = first_lm_key; k < (int)a.keys.size(); ++k) {
      const int l = a.keys[k].node - a.sp.first_lm_node - a.sp.lm_begin;
      const unsigned char own = (l >= 0 && l < a.sp.n_landmarks) ? 1 : 0;
      for (int q = 0; q < a.keys[k].sdim; ++q) mask[a.keys[k].voff + q] = own;
    }
    p->d_vmask = P.upload(mask);
  }
  P.zero(p->d_maxdiag, sizeof(double) * a.N);
  p->d_ctrl = P.alloc<Ctrl>(1);
  P.zero(p->d_ctrl, sizeof(Ctrl));  // the whole block (iteration records included) is read back after every run
  CUDA_OK(cudaMallocHost(&p->h_ctrl, sizeof(Ctrl)));
  std::memset(p->h_ctrl, 0, sizeof(Ctrl));  // entry points that come before any sfx_optimize read best_valid / n_iters
  CUDA_OK(cudaHostAlloc(&p->h_done, sizeof(int) * (kMaxIterations + 2), cudaHostAllocMapped));
  CUDA_OK(cudaHostGetDevicePointer(&p->d_done, p->h_done, 0));
  std::memset(p->h_done, 0, sizeof(int) * (kMaxIterations + 2));

  // Schur
  if (a.schur) {
    SchurPlan& s = a.sp;
    SchurDev& d = p->sd;
    d.n_landmarks = s.n_landmarks;
    d.add_b = a.rank == 0 ? 1 : 0;
    d.n_reduced_nodes = s.first_lm_node;
    d.reduced_dim = s.reduced_dim;
    d.lm_dim = P.upload(s.lm_dim);
    d.lm_cdiag_off = P.upload(s.lm_cdiag_off);
    d.lm_toff = P.upload(s.lm_toff);
    d.lm_e_ptr = P.upload(s.lm_e_ptr);
    d.lm_e_off = P.upload(s.lm_e_off);
    d.lm_e_node = P.upload(s.lm_e_node);
    std::vector<int32_t> nto, ndm;
    for (int i = 0; i < s.first_lm_node; ++i) {
      nto.push_back(a.nodes[i].toff);
      ndm.push_back(a.nodes[i].dim);
    }
    d.node_toff = P.upload(nto);
    d.node_dim = P.upload(ndm);
    d.n_sblocks = (int)s.S.row_idx.size();
    std::vector<int32_t> srow(s.S.row_idx.begin(), s.S.row_idx.end()), scol(srow.size());
    for (int j = 0; j < s.S.n_nodes; ++j)
      for (int q = s.S.col_ptr[j]; q < s.S.col_ptr[j + 1]; ++q) scol[q] = j;
    d.s_row = P.upload(srow);
    d.s_col = P.upload(scol);
    d.s_off = P.upload(s.S.blk_off);
    d.s_b_src = P.upload(s.s_b_src);
    d.s_m_ptr = P.upload(s.s_m_ptr);
    d.m_eoff_i = P.upload(s.m_eoff_i);
    d.m_eoff_j = P.upload(s.m_eoff_j);
    d.m_lm = P.upload(s.m_lm);
    {
      bool v3 = !getenv("SFX_SCHUR_V2") && !getenv("SFX_SCHUR_V1") && !getenv("SFX_NO_SCHUR_FAST");
      for (int l = 0; l < s.n_landmarks && v3; ++l) v3 = s.lm_dim[l] == 3;
      for (int i = 0; i < s.first_lm_node && v3; ++i) v3 = a.nodes[i].dim == 9;  // schur_s9_kernel: BAL shape only
      SFX_CHECK(a.H.n_values + 64 < (int64_t)INT32_MAX, SFX_ERR_UNSUPPORTED, "Hessian too large for int32 offsets");
      const int32_t zero_block = (int32_t)a.H.n_values;  // 32 zero doubles behind the W buffer
      const int chunk = v3 ? 64 : getenv("SFX_SCHUR_CHUNK") ? atoi(getenv("SFX_SCHUR_CHUNK")) : 64;
      std::vector<int32_t> ib, im, ic, ifl;
      SFX_CHECK(s.m_lm.size() < (size_t)INT32_MAX, SFX_ERR_UNSUPPORTED, "too many Schur matches");
      for (int b = 0; b < d.n_sblocks; ++b) {
        const int64_t m0 = s.s_m_ptr[b], m1 = s.s_m_ptr[b + 1];
        const int64_t n = m1 - m0;
        const int nch = n == 0 ? 1 : (int)((n + chunk - 1) / chunk);
        for (int c = 0; c < nch; ++c) {
          ib.push_back(b);
          im.push_back((int32_t)(m0 + (int64_t)c * chunk));
          ic.push_back((int32_t)std::min<int64_t>(chunk, n - (int64_t)c * chunk));
          ifl.push_back((c == 0 ? 1 : 0) | (nch == 1 ? 2 : 0));
        }
      }
      {
        // packed headers for schur_s2_kernel
        std::vector<int32_t> hdr(ib.size() * 8);
        for (size_t q = 0; q < ib.size(); ++q) {
          const int b = ib[q];
          const int I = srow[b], J = scol[b];
          const int dI = a.nodes[I].dim, dJ = a.nodes[J].dim;
          const int64_t so = s.S.blk_off[b];
          int32_t* h = &hdr[q * 8];
          h[0] = im[q];
          h[1] = ic[q];
          h[2] = ifl[q] | (dI << 8) | (dJ << 16) | ((I == J ? 1 : 0) << 24);
          h[3] = a.nodes[I].toff;
          h[4] = (int32_t)(uint32_t)(so & 0xffffffffll);
          h[5] = (int32_t)(so >> 32);
          h[6] = s.s_b_src[b];
          h[7] = 0;
        }
        d.items2 = P.upload(hdr);
        if (v3) {
          // padded per-item match offsets + 16-byte headers for the persistent kernel
          // the diagonal blocks are accumulated by schur_w_rhs_kernel (every reduced node needs its S_II block for that)
          std::vector<int64_t> diag_off(s.first_lm_node, -1);
          std::vector<int32_t> diag_bsrc(s.first_lm_node, -1);
          for (int b = 0; b < d.n_sblocks; ++b)
            if (srow[b] == scol[b]) {
              diag_off[srow[b]] = s.S.blk_off[b];
              diag_bsrc[srow[b]] = s.s_b_src[b];
            }
          // (with few cameras the REDs of all warps meet on a handful of blocks: Ladybug-shape, 49 cameras, is 0.023 ms
          // per iteration slower with the fusion than without)
          bool fuse_diag = !getenv("SFX_S9_DIAG_ITEMS") && (s.first_lm_node >= 256 || getenv("SFX_S9_DIAG_FUSE"));
          for (int i = 0; i < s.first_lm_node && fuse_diag; ++i) fuse_diag = diag_off[i] >= 0;
          std::vector<size_t> keep;  // items of the persistent kernel
          for (size_t q = 0; q < ib.size(); ++q)
            if (!fuse_diag || srow[ib[q]] != scol[ib[q]]) keep.push_back(q);
          std::vector<int32_t> h3(keep.size() * 4), pmi(keep.size() * 64, zero_block), pmj(keep.size() * 64, zero_block),
              toI(keep.size());
          for (size_t o = 0; o < keep.size(); ++o) {
            const size_t q = keep[o];
            const int32_t* h = &hdr[q * 8];
            h3[o * 4 + 0] = h[4];
            h3[o * 4 + 1] = h[5];
            h3[o * 4 + 2] = h[6];
            h3[o * 4 + 3] = h[2] | (ic[q] << 25);
            toI[o] = h[3];
            for (int c = 0; c < ic[q]; ++c) {
              pmi[o * 64 + c] = s.m_eoff_i[im[q] + c];
              pmj[o * 64 + c] = s.m_eoff_j[im[q] + c];
            }
          }
          d.n_items3 = (int)keep.size();
          if (fuse_diag) {
            d.s_diag_off = P.upload(diag_off);
            d.s_diag_bsrc = P.upload(diag_bsrc);
          }
          d.items3 = P.upload(h3);
          d.pm_i = P.upload(pmi);
          d.pm_j = P.upload(pmj);
          d.item_toI = P.upload(toI);
        }
      }
      d.n_items = (int)ib.size();
      d.item_blk = P.upload(ib);
      d.item_m0 = P.upload(im);
      d.item_cnt = P.upload(ic);
      d.item_flags = P.upload(ifl);
      d.s_values = s.S.n_values;
    }
    d.r_ptr = P.upload(s.r_ptr);
    d.r_eoff = P.upload(s.r_eoff);
    d.r_lm = P.upload(s.r_lm);
    d.cinv = P.alloc<double>((size_t)s.n_landmarks * 9);
    d.tl = P.alloc<double>((size_t)s.n_landmarks * 3);
    d.S = P.alloc<double>(s.S.n_values);
    d.rhs_red = P.alloc<double>(s.reduced_dim);
    {
      bool fast = !getenv("SFX_NO_SCHUR_FAST");
      for (int l = 0; l < s.n_landmarks && fast; ++l) fast = s.lm_dim[l] == 3;
      for (int i = 0; i < s.first_lm_node && fast; ++i) fast = a.nodes[i].dim <= 16;
      d.fast3 = fast ? 1 : 0;
      d.n_entries = (int)s.r_eoff.size();
      std::vector<int32_t> rnode(s.r_eoff.size());
      for (int j = 0; j < s.first_lm_node; ++j)
        for (int q = s.r_ptr[j]; q < s.r_ptr[j + 1]; ++q) rnode[q] = j;
      d.r_node = P.upload(rnode);
      d.G = fast ? P.alloc<double>(a.H.n_values + 32) : nullptr;
      if (fast) P.zero(d.G + a.H.n_values, 32 * sizeof(double));
      d.wl = (fast && !getenv("SFX_SCHUR_V1")) ? P.alloc<double>((size_t)s.n_landmarks * 9) : nullptr;
      d.zeros = P.upload(std::vector<double>(8, 0.0));
      d.sl = P.alloc<double>((size_t)s.n_landmarks * 3);
      // slot view for the TMA-streamed kernels: every E block of an own landmark on a multiple of 27 doubles, every
      // reduced node 9-dimensional, and nothing but such blocks and the 81-double camera diagonal blocks before the
      // last of them
      d.n_slots = 0;
      d.slot_base = 0;
      bool slots = fast && !getenv("SFX_NO_TMA") && !s.r_eoff.empty();
      for (int i = 0; i < s.first_lm_node && slots; ++i) slots = a.nodes[i].dim == 9;
      int64_t first = INT64_MAX, last = 0;
      for (size_t q = 0; q < s.r_eoff.size(); ++q) {
        first = std::min<int64_t>(first, s.r_eoff[q]);
        last = std::max<int64_t>(last, s.r_eoff[q]);
      }
      slots = slots && first % 2 == 0;  // 16-byte aligned source of the bulk copies
      for (size_t q = 0; q < s.r_eoff.size() && slots; ++q) slots = (s.r_eoff[q] - first) % 27 == 0;
      if (slots) {
        const int64_t ns = (last - first) / 27 + 1;
        // (an odd number of slots makes the last bulk copy read 8 bytes past the last block: keep them inside H)
        slots = ns < (int64_t)INT32_MAX / 32 && first + ns * 27 <= a.H.n_values + 1 && ns <= 2 * (int64_t)s.r_eoff.size();
        if (slots) {
          std::vector<int32_t> slm((size_t)ns, -1), snode((size_t)ns, 0);
          for (int j = 0; j < s.first_lm_node; ++j)
            for (int q = s.r_ptr[j]; q < s.r_ptr[j + 1]; ++q) {
              slm[(s.r_eoff[q] - first) / 27] = s.r_lm[q];
              snode[(s.r_eoff[q] - first) / 27] = j;
            }
          d.slot_lm = P.upload(slm);
          d.slot_node = P.upload(snode);
          d.n_slots = (int)ns;
          d.slot_base = first;
        }
      }
    }
  }
  // fronts
  {
    FrontPlan& f = a.fp;
    FrontDev& d = p->fd;
    d.n_fronts = f.n_fronts;
    d.n = f.n;
    auto up32 = [&](const std::vector<int>& v) { return P.upload(std::vector<int32_t>(v.begin(), v.end())); };
    d.f_w = up32(f.f_w);
    d.f_u = up32(f.f_u);
    d.f_piv = up32(f.f_piv);
    d.f_rows_ptr = up32(f.f_rows_ptr);
    d.f_rows = up32(f.f_rows);
    d.f_rel = up32(f.f_rel);
    d.f_child_ptr = up32(f.f_child_ptr);
    d.f_child = up32(f.f_child);
    d.f_toff = up32(f.f_toff);
    d.f_copy_ptr = up32(f.f_copy_ptr);
    d.f_off = P.upload(f.f_off);
    std::vector<FrontCopy> cp(f.copies.size());
    for (size_t i = 0; i < cp.size(); ++i) {
      const auto& c = f.copies[i];
      cp[i] = FrontCopy{c.src, c.rows, c.cols, c.src_ld, c.dst_row, c.dst_col, c.transposed, c.lower_only, 0};
    }
    d.copies = P.upload(cp);
    d.scalar_perm = up32(f.scalar_perm);
    d.fronts = P.alloc<double>(f.front_values);
    d.twork = P.alloc<double>(f.solve_ws);
    d.ywork = P.alloc<double>(f.n);
    CUDA_OK(configure_large_kernels());
    LargeHostPlan hp;
    plan_large_fronts(p, large_factor_resident_ctas(), hp);
    std::vector<int>& lvl_fronts = hp.lvl_fronts;
    std::vector<LargeFront>& lfs = hp.lfs;
    std::vector<LargeTask>& tasks = hp.tasks;
    std::vector<LargeJob>&jobs = hp.jobs, &pre_jobs = hp.pre_jobs, &damp_jobs = hp.damp_jobs;
    const int64_t linv_off = hp.linv_off, cnt_off = hp.cnt_off, flag_off = hp.flag_off, contrib_off = hp.contrib_off;
    d.level_fronts = up32(lvl_fronts);
    p->n_large_fronts = (int)lfs.size();
    {
      // zeroing jobs: runs of columns with ~8k entries each (small fronts: one job; the 2844-row root: ~500)
      std::vector<int4> zj;
      for (size_t q = 0; q < lfs.size(); ++q) {
        const int m = lfs[q].m;
        int c0 = 0;
        while (c0 < m) {
          int c1 = c0;
          int64_t el = 0;
          while (c1 < m && el < 8192) {
            el += m - std::max(0, c1 - 63);
            ++c1;
          }
          zj.push_back(make_int4((int)q, c0, c1, 0));
          c0 = c1;
        }
      }
      p->n_zero_jobs = (int)zj.size();
      p->ld.zero_jobs = P.upload(zj);
    }
    p->ld.fwd_b = P.alloc<double>(std::max<int64_t>(hp.fwd_b_size, 1));
    p->d_f_level = up32(f.f_level);
    p->ld.lf = P.upload(lfs);
    p->ld.tasks = P.upload(tasks);
    p->pre_j0 = (int)jobs.size();
    jobs.insert(jobs.end(), pre_jobs.begin(), pre_jobs.end());
    p->pre_j1 = p->damp_j0 = (int)jobs.size();
    jobs.insert(jobs.end(), damp_jobs.begin(), damp_jobs.end());
    p->damp_j1 = (int)jobs.size();
    p->ld.jobs = P.upload(jobs);
    p->n_counters = cnt_off;
    p->ld.counters = P.alloc<int>(cnt_off);
    p->ld.queue = P.alloc<int>(f.n_levels);
    p->ld.linv = P.alloc<double>(linv_off);
    p->ld.binv = P.alloc<double>(linv_off / 8 + 512);
    p->n_sflags = flag_off;
    p->ld.sflags = P.alloc<int>(flag_off);
    p->ld.contrib = P.alloc<double>(contrib_off);
    p->ld.ll_y = P.alloc<uint4>(f.n);
    p->ld.ll_contrib = P.alloc<uint4>(contrib_off);
    P.zero(p->ld.ll_y, sizeof(uint4) * std::max<int64_t>(f.n, 1));
    P.zero(p->ld.ll_contrib, sizeof(uint4) * std::max<int64_t>(contrib_off, 1));
    CUDA_OK(configure_front_kernels(p->smem_cap_m, f.max_front));
    CUDA_OK(configure_large_kernels());
  }
}

// optimizer_params_t sanity (the reference SYM_ASSERTs on lambda_update_type INVALID,
// levenberg_marquardt_solver.tcc:322-339; a zero-initialised struct must not silently run another algorithm)
void validate_params(const sfx_params& q) {
  SFX_CHECK(q.lambda_update_type == 1 || q.lambda_update_type == 2, SFX_ERR_INVALID_ARG,
            "lambda_update_type must be STATIC (1) or DYNAMIC (2)");
  SFX_CHECK(q.iterations > 0, SFX_ERR_INVALID_ARG, "iterations must be positive");
  SFX_CHECK(q.initial_lambda >= 0 && q.lambda_lower_bound >= 0 && q.lambda_lower_bound <= q.lambda_upper_bound,
            SFX_ERR_INVALID_ARG, "need 0 <= lambda_lower_bound <= lambda_upper_bound and initial_lambda >= 0");
  SFX_CHECK(q.lambda_up_factor > 0 && q.lambda_down_factor > 0, SFX_ERR_INVALID_ARG, "lambda factors must be positive");
  if (q.lambda_update_type == 2)
    SFX_CHECK(q.dynamic_lambda_update_beta > 0 && q.dynamic_lambda_update_gamma > 0 && q.dynamic_lambda_update_p > 0,
              SFX_ERR_INVALID_ARG, "dynamic lambda update needs beta, gamma, p > 0");
}

void reset_ctrl(sfx_problem* p) {
  Ctrl* c = p->h_ctrl;
  p->can_continue = false;
  p->eager_valid = false;  // every caller goes on to overwrite the state blocks
  p->lin0_clobbered = false;
  std::memset(c, 0, offsetof(Ctrl, iters));
  c->p = p->params;
  c->epsilon = p->epsilon;
  c->lambda = p->params.initial_lambda;
  c->nu = p->params.dynamic_lambda_update_beta;
  c->init_idx = 0;
  c->new_idx = 1;
  c->best_idx = 0;
  c->free_idx = 2;
  c->iteration = -1;
  CUDA_OK(cudaMemcpyAsync(p->d_ctrl, c, offsetof(Ctrl, iters), cudaMemcpyHostToDevice, p->st));
  std::memset(p->h_done, 0, sizeof(int) * (kMaxIterations + 2));
}

// LevenbergMarquardtSolver::ResetState (levenberg_marquardt_solver.h:178-183) on top of the control block left
// by the previous Optimize: linearizations and Best become invalid, max-diagonal / last-update memories are
// dropped; lambda, nu, the iteration counter, the iteration records and the state-block indices stay.
void reset_ctrl_continue(sfx_problem* p) {
  Ctrl* c = p->h_ctrl;
  c->p = p->params;
  c->have_max_diag = 0;
  c->have_last_update = 0;
  c->lin_valid[0] = c->lin_valid[1] = c->lin_valid[2] = 0;
  c->best_valid = 0;
  c->done = 0;
  c->failure_reason = 0;
  c->chol_fail = 0;
  c->fail_where = 0;
  c->n_chol_fail = c->n_nonfinite_update = c->n_zero_diag = 0;
  CUDA_OK(cudaMemcpyAsync(p->d_ctrl, c, offsetof(Ctrl, iters), cudaMemcpyHostToDevice, p->st));
  std::memset(p->h_done, 0, sizeof(int) * (kMaxIterations + 2));
}

void enqueue_linearize(sfx_problem* p, int mode) {
  launch_zero_lin(p->st, p->d_ctrl, p->sp, mode, p->a.h_accum_values, p->a.N);
  for (auto& lb : p->lin) launch_linearize(p->st, p->d_ctrl, p->sp, mode, lb, p->d_partials);
  launch_finish_error(p->st, p->d_ctrl, mode, p->d_partials, p->n_partials);
  if (p->a.world > 1) {
    // one all-reduce per linearization: B blocks, camera rhs and the error partial
    const int64_t nb = p->a.b_values;
    const int nr = p->a.sp.reduced_dim;
    launch_pack_b(p->st, p->d_ctrl, p->sp, mode, nb, nr, p->d_stage, 0);
    NCCL_OK(nccl().AllReduce(p->d_stage, p->d_stage, (size_t)(nb + nr + 1), ncclDouble, ncclSum, p->comm->comm, p->st));
    launch_pack_b(p->st, p->d_ctrl, p->sp, mode, nb, nr, p->d_stage, 1);
  }
  launch_commit_error(p->st, p->d_ctrl, mode);
}

// zeroes the large fronts on the side stream (joined by enqueue_factorize); the previous solve, the
// last reader of the fronts, precedes the fork event on the main stream
void enqueue_zero_fork(sfx_problem* p) {
  if (p->n_large_fronts == 0) return;
  CUDA_OK(cudaEventRecord(p->ev_fork, p->st));
  CUDA_OK(cudaStreamWaitEvent(p->st2, p->ev_fork, 0));
  launch_large_zero(p->st2, p->d_ctrl, p->fd, p->ld, p->n_zero_jobs);
  CUDA_OK(cudaEventRecord(p->ev_join, p->st2));
}

// multifrontal Cholesky of the damped system: S (Schur problems; damping already inside) or
// H[init_idx] + diag(d_dvec)
// First level of the chain-bound top of the elimination tree: every level from there up is one large front and no
// small ones (0 = none).  While those levels factor (their diagonal chain leaves most SMs idle, and one CTA per SM
// is as fast as two), the forward substitution of everything below can already run on the side stream.
int chain_top_level(const sfx_problem* p) {
  const FrontPlan& f = p->a.fp;
  // measured at Final-shape: the solve phase drops 1.01 -> 0.86 ms but the chain-bound levels slow down by the same
  // amount (the substitution CTAs share SMs with the diagonal-chain CTAs), so this stays opt-in
  if (!getenv("SFX_SOLVE_OVERLAP") || getenv("SFX_SOLVE_V1") || p->fused_T0 >= 0) return 0;
  int T = f.n_levels;
  while (T > 0 && p->lvl_large[T - 1].n_lf == 1 && p->lvl_small_cnt[T - 1] == 0) --T;
  if (T >= f.n_levels || T == 0) return 0;
  return T;
}

void begin_tri_solves(sfx_problem* p) {
  if (p->n_sflags > 0) CUDA_OK(cudaMemsetAsync(p->ld.sflags, 0, sizeof(int) * p->n_sflags, p->st));
  if (++p->solve_epoch == 0) p->solve_epoch = 1;  // LL slots of the v2 solves: 0 means "never written"
}

void enqueue_fwd_levels(sfx_problem* p, cudaStream_t st, int l0, int l1, const double* rhs_static, int use_state_rhs) {
  const FrontPlan& f = p->a.fp;
  for (int l = l0; l < l1; ++l) {
    if (p->lvl_small_cnt[l] > 0)
      launch_front_solve_fwd(st, p->d_ctrl, p->fd, rhs_static, p->sp, use_state_rhs, f.level_ptr[l],
                             p->lvl_small_cnt[l], p->lvl_max_m[l] * 8);
    launch_large_solve_fwd(st, p->d_ctrl, p->fd, p->ld, p->lvl_large[l], rhs_static, p->sp, use_state_rhs, p->solve_epoch);
  }
}

// overlap_T > 0: begins the triangular solves too -- the forward substitution of the levels below overlap_T runs on
// the side stream while the levels from overlap_T up factor (with one CTA per SM); joined before returning
// with_fwd: the forward substitution of (rhs_static | the state right-hand side) is part of this call -- levels below
// the fused top by their own kernels, the fused fronts by tasks of the factor kernel; returns true when it was, and
// enqueue_tri_solves must then be told that every level is done (fwd_done_below = n_levels)
bool enqueue_factorize(sfx_problem* p, int overlap_T = 0, const double* rhs_static = nullptr, int use_state_rhs = 0,
                       bool with_fwd = false) {
  Analysis& a = p->a;
  const FrontPlan& f = a.fp;
  const double* sys = a.schur ? p->sd.S : nullptr;
  const int use_H = a.schur ? 0 : 1;
  const double* dv = a.schur ? nullptr : p->d_dvec;
  if (p->n_large_fronts > 0) CUDA_OK(cudaStreamWaitEvent(p->st, p->ev_join, 0));
  if (p->n_counters > 0) {
    CUDA_OK(cudaMemsetAsync(p->ld.counters, 0, sizeof(int) * p->n_counters, p->st));
    CUDA_OK(cudaMemsetAsync(p->ld.queue, 0, sizeof(int) * f.n_levels, p->st));
  }
  launch_large_preassemble(p->st, p->d_ctrl, p->fd, p->ld, sys, p->sp, use_H, dv, p->pre_j0, p->pre_j1, p->damp_j0,
                           p->damp_j1);
  for (int l = 0; l < f.n_levels; ++l) {
    if (l == p->fused_T0) {
      const bool fwd = with_fwd && p->fused_fwd;
      if (fwd) {
        begin_tri_solves(p);
        enqueue_fwd_levels(p, p->st, 0, l, rhs_static, use_state_rhs);
        launch_large_fwd_init(p->st, p->d_ctrl, p->fd, p->ld, p->lvl_large[l].lf0, p->n_large_fronts - p->lvl_large[l].lf0, l,
                              p->d_f_level, rhs_static, p->sp, use_state_rhs);
      }
      launch_large_fused(p->st, p->d_ctrl, p->fd, p->ld, p->fused_t0, p->fused_t1, p->fused_j0, p->fused_j1, l, sys, p->sp,
                         use_H, dv, fwd ? 1 : 0);
      return fwd;
    }
    if (overlap_T > 0 && l == overlap_T) {
      begin_tri_solves(p);
      CUDA_OK(cudaEventRecord(p->ev_fork2, p->st));
      CUDA_OK(cudaStreamWaitEvent(p->st2, p->ev_fork2, 0));
      enqueue_fwd_levels(p, p->st2, 0, overlap_T, rhs_static, use_state_rhs);
      CUDA_OK(cudaEventRecord(p->ev_join2, p->st2));
    }
    // small fronts: the ones with at most 32 rows one warp each, the others one CTA each
    if (p->lvl_tiny_cnt[l] > 0)
      launch_front_factor(p->st, p->d_ctrl, p->fd, sys, p->sp, use_H, dv, f.level_ptr[l], p->lvl_tiny_cnt[l], kWarpFrontRows);
    if (p->lvl_small_cnt[l] > p->lvl_tiny_cnt[l])
      launch_front_factor(p->st, p->d_ctrl, p->fd, sys, p->sp, use_H, dv, f.level_ptr[l] + p->lvl_tiny_cnt[l],
                          p->lvl_small_cnt[l] - p->lvl_tiny_cnt[l], p->lvl_rest_max_m[l]);
    launch_large_level(p->st, p->d_ctrl, p->fd, p->ld, p->lvl_large[l], l, sys, p->sp, use_H, dv,
                       (overlap_T > 0 && l >= overlap_T) ? p->sm_count : 0);
  }
  if (overlap_T > 0) CUDA_OK(cudaStreamWaitEvent(p->st, p->ev_join2, 0));
  return false;
}

// forward + backward substitution with the current factor; rhs in system scalar order (rhs_static, or
// the rhs of state block init_idx); the solution is left in fd.ywork (elimination order)
// fwd_done_below > 0: the forward substitution of the levels below was already enqueued by enqueue_factorize
void enqueue_tri_solves(sfx_problem* p, const double* rhs_static, int use_state_rhs, int fwd_done_below = 0) {
  const FrontPlan& f = p->a.fp;
  if (fwd_done_below == 0) begin_tri_solves(p);
  enqueue_fwd_levels(p, p->st, fwd_done_below, f.n_levels, rhs_static, use_state_rhs);
  for (int l = f.n_levels - 1; l >= 0; --l) {
    if (p->lvl_small_cnt[l] > 0)
      launch_front_solve_bwd(p->st, p->d_ctrl, p->fd, f.level_ptr[l], p->lvl_small_cnt[l], p->lvl_max_m[l] * 8);
    launch_large_solve_bwd(p->st, p->d_ctrl, p->fd, p->ld, p->lvl_large[l], p->solve_epoch);
  }
}

// damping + [Schur] + factorize + solve -> d_upd (internal order) = -H_damped^-1 rhs
void enqueue_solve(sfx_problem* p, const std::function<void(int)>& mark) {
  Analysis& a = p->a;
  const bool factors_here = !(a.world > 1 && a.rank != 0);
  if (factors_here) enqueue_zero_fork(p);  // overlaps damping and the Schur complement
  launch_damping(p->st, p->d_ctrl, p->sp, p->d_diag_pos, a.N, p->d_dvec, p->d_maxdiag);
  if (a.schur) launch_schur(p->st, p->d_ctrl, p->sp, p->sd, p->d_dvec);
  const bool mg = a.world > 1;
  if (mg) {
    // the block-sparse S (identical pattern on every rank) and the reduced rhs are summed onto rank 0
    NCCL_OK(nccl().Reduce(p->sd.S, p->sd.S, (size_t)a.sp.S.n_values, ncclDouble, ncclSum, 0, p->comm->comm, p->st));
    NCCL_OK(nccl().Reduce(p->sd.rhs_red, p->sd.rhs_red, (size_t)a.sp.reduced_dim, ncclDouble, ncclSum, 0,
                          p->comm->comm, p->st));
  }
  mark(PH_SCHUR);
  const FrontPlan& f = a.fp;
  if (mg && a.rank != 0) {
    // rank 0 factors and solves; everyone receives the camera step
    mark(PH_FACTOR);
    NCCL_OK(nccl().Broadcast(p->d_y, p->d_y, (size_t)a.sp.reduced_dim, ncclDouble, 0, p->comm->comm, p->st));
    launch_schur_back(p->st, p->d_ctrl, p->sp, p->sd, p->d_y, p->d_upd);
    mark(PH_SOLVE);
    return;
  }
  (void)f;
  const int T = chain_top_level(p);
  const bool fwd_done = enqueue_factorize(p, T, a.schur ? p->sd.rhs_red : nullptr, a.schur ? 0 : 1, /*with_fwd=*/true);
  mark(PH_FACTOR);
  enqueue_tri_solves(p, a.schur ? p->sd.rhs_red : nullptr, a.schur ? 0 : 1, fwd_done ? a.fp.n_levels : T);
  if (a.schur) {
    launch_unpermute(p->st, p->d_ctrl, p->fd, p->d_y, 1.0);
    if (mg) NCCL_OK(nccl().Broadcast(p->d_y, p->d_y, (size_t)a.sp.reduced_dim, ncclDouble, 0, p->comm->comm, p->st));
    launch_schur_back(p->st, p->d_ctrl, p->sp, p->sd, p->d_y, p->d_upd);
  } else {
    launch_unpermute(p->st, p->d_ctrl, p->fd, p->d_upd, -1.0);
  }
  mark(PH_SOLVE);
}

void ensure_csc(sfx_problem* p) {
  build_csc(p->a);
  if (!p->d_csc_src) p->d_csc_src = p->pool.upload(p->a.csc_src);
  if (p->export_cap < p->a.nnz) {
    p->d_export = p->pool.alloc<double>(p->a.nnz);
    p->export_cap = p->a.nnz;
  }
}

// index maps of the Jacobian export, built and uploaded on first use
void ensure_jacobian(sfx_problem* p) {
  build_jacobian_csc(p->a);
  if (p->d_jac_base.empty())
    for (const auto& bp : p->a.batches) {
      p->d_jac_base.push_back(p->pool.upload(bp.jac_base));
      p->d_jac_colnnz.push_back(p->pool.upload(bp.jac_colnnz));
    }
  if (p->export_cap < p->a.jac_nnz) {
    p->d_export = p->pool.alloc<double>(p->a.jac_nnz);
    p->export_cap = p->a.jac_nnz;
  }
}

// multi-GPU: `src` holds the replicated keys and this rank's landmarks; returns a buffer with every rank's landmarks.
// Contiguous landmark runs travel as one broadcast per owner over NVLink; other layouts as a masked all-reduce.
const double* gather_sharded_values(sfx_problem* p, const double* src, int64_t n) {
  if (!p->rank_lm_range.empty()) {
    CUDA_OK(cudaMemcpyAsync(p->d_stage, src, sizeof(double) * n, cudaMemcpyDeviceToDevice, p->st));
    NCCL_OK(nccl().GroupStart());
    for (int r = 0; r < p->a.world; ++r) {
      const auto& x = p->rank_lm_range[r];
      if (x.second > x.first)
        NCCL_OK(nccl().Broadcast(p->d_stage + x.first, p->d_stage + x.first, (size_t)(x.second - x.first), ncclDouble, r,
                                 p->comm->comm, p->st));
    }
    NCCL_OK(nccl().GroupEnd());
    return p->d_stage;
  }
  launch_mask_values(p->st, src, p->d_vmask, n, p->d_stage);
  NCCL_OK(nccl().AllReduce(p->d_stage, p->d_stage, (size_t)n, ncclDouble, ncclSum, p->comm->comm, p->st));
  return p->d_stage;
}

void export_linearization(sfx_problem* p, int blk, double* residual, double* rhs, double* Hv) {
  Analysis& a = p->a;
  if (residual)
    CUDA_OK(cudaMemcpyAsync(residual, p->sp.res[blk], sizeof(double) * a.M, cudaMemcpyDeviceToHost, p->st));
  if (rhs) {
    launch_permute_vec(p->st, p->sp.rhs[blk], p->d_ref2int, a.N, p->d_y);
    CUDA_OK(cudaMemcpyAsync(rhs, p->d_y, sizeof(double) * a.N, cudaMemcpyDeviceToHost, p->st));
  }
  if (Hv) {
    ensure_csc(p);
    launch_export_csc(p->st, p->sp.H[blk], p->d_csc_src, a.nnz, p->d_export);
    CUDA_OK(cudaMemcpyAsync(Hv, p->d_export, sizeof(double) * a.nnz, cudaMemcpyDeviceToHost, p->st));
  }
  CUDA_OK(cudaStreamSynchronize(p->st));
}

}  // namespace

#define SFX_API_BEGIN try {
#define SFX_API_END(p)                                  \
  return SFX_OK;                                        \
  }                                                     \
  catch (const sfx::Error& e) {                         \
    if (p) (p)->err = e.what();                         \
    g_create_err = e.what();                            \
    return e.code;                                      \
  }                                                     \
  catch (const std::exception& e) {                     \
    if (p) (p)->err = e.what();                         \
    g_create_err = e.what();                            \
    return SFX_ERR_INVALID_ARG;                         \
  }

extern "C" {

sfx_status sfx_default_params(sfx_params* out) {
  if (!out) return SFX_ERR_INVALID_ARG;
  sfx_params p{};
  p.initial_lambda = 1.0;
  p.lambda_lower_bound = 0.0;
  p.lambda_upper_bound = 1000000.0;
  p.lambda_update_type = 1;
  p.lambda_up_factor = 4.0;
  p.lambda_down_factor = 1 / 4.0;
  p.dynamic_lambda_update_beta = 2.0;
  p.dynamic_lambda_update_gamma = 3.0;
  p.dynamic_lambda_update_p = 3;
  p.use_diagonal_damping = 0;
  p.use_unit_damping = 1;
  p.keep_max_diagonal_damping = 0;
  p.diagonal_damping_min = 1e-6;
  p.iterations = 50;
  p.early_exit_min_reduction = 1e-6;
  p.early_exit_min_absolute_error = 0.0;
  p.enable_bold_updates = 0;
  *out = p;
  return SFX_OK;
}

const char* sfx_last_error(const sfx_problem* p) { return p ? p->err.c_str() : g_create_err.c_str(); }

sfx_status sfx_problem_create(const sfx_problem_desc* desc, sfx_problem** out) {
  sfx_problem* p = nullptr;
  SFX_API_BEGIN
  SFX_CHECK(desc && out, SFX_ERR_INVALID_ARG, "null argument");
  *out = nullptr;
  PhaseClock clk;  // SFX_TIMING=1: where the setup time goes
  int ndev = 0;
  cudaError_t ce = cudaGetDeviceCount(&ndev);
  if (ce != cudaSuccess || ndev == 0)
    throw Error(SFX_ERR_CUDA, "no CUDA device: libsfx has no CPU fallback");
  SFX_CHECK(desc->device >= 0 && desc->device < ndev, SFX_ERR_INVALID_ARG, "device ordinal out of range");
  CUDA_OK(cudaSetDevice(desc->device));
  std::unique_ptr<sfx_problem> up(new sfx_problem());
  p = up.get();
  p->device = desc->device;
  CUDA_OK(cudaDeviceGetAttribute(&p->sm_count, cudaDevAttrMultiProcessorCount, desc->device));
  validate_params(desc->params);
  p->params = desc->params;
  p->epsilon = desc->epsilon;
  p->ordering = desc->ordering;
  p->comm = (sfx_comm*)desc->comm;
  if (p->comm) {
    SFX_CHECK(p->comm->rank == desc->rank && p->comm->world == desc->world, SFX_ERR_INVALID_ARG,
              "rank/world of the descriptor and the communicator differ");
    SFX_CHECK(p->comm->device == desc->device, SFX_ERR_INVALID_ARG, "communicator was created for another device");
  }
  clk.lap("create: CUDA context");
  analyze_problem(*desc, p->a);
  clk.lap("create: analysis (total)");
  {
    Analysis& a = p->a;
    const BlockMatrix& sys = a.schur ? a.sp.S : a.H;
    std::vector<int> sys2ref;
    if (desc->ordering == SFX_ORDERING_METIS_SCALAR) {
      std::vector<int> int2ref(a.N);
      for (int r = 0; r < a.N; ++r) int2ref[a.ref2int[r]] = r;
      sys2ref.assign(int2ref.begin(), int2ref.begin() + sys.node_off[sys.n_nodes]);
    }
    choose_front_plan(p, sys, desc->ordering, sys2ref, large_factor_resident_ctas());
  }
  clk.lap("create: ordering + front plan");
  CUDA_OK(cudaStreamCreateWithFlags(&p->st, cudaStreamNonBlocking));
  CUDA_OK(cudaStreamCreateWithFlags(&p->st2, cudaStreamNonBlocking));
  CUDA_OK(cudaEventCreateWithFlags(&p->ev_fork, cudaEventDisableTiming));
  CUDA_OK(cudaEventCreateWithFlags(&p->ev_join, cudaEventDisableTiming));
  CUDA_OK(cudaEventCreateWithFlags(&p->ev_fork2, cudaEventDisableTiming));
  CUDA_OK(cudaEventCreateWithFlags(&p->ev_join2, cudaEventDisableTiming));
  p->pool.st = p->st;
  upload_structures(p);
  {
    // eager linearization at sfx_set_values: one BAL fast-path batch whose pixel argument of observation s sits at
    // base + 2 s of the values buffer (the reference example's layout: measurements stored in observation order)
    const Analysis& a = p->a;
    bool ok = a.world == 1 && p->lin.size() == 1 && p->lin[0].bal_fast && !getenv("SFX_NO_EAGER_LINEARIZE");
    if (ok) {
      const auto& ao = a.batches[0].arg_off;
      const int64_t nb = p->lin[0].n;
      ok = nb > 0 && (int64_t)ao.size() >= 4 * nb;
      const int64_t base = ok ? ao[3 * nb] : 0;
      for (int64_t q = 0; q < nb && ok; ++q) ok = ao[3 * nb + q] == base + 2 * q;
      ok = ok && base >= 0 && base + 2 * nb <= a.n_values;
      // no other argument may live inside the pixel run (it is uploaded after the first kernels have started)
      for (int k = 0; k < 3 && ok; ++k)
        for (int64_t q = 0; q < nb && ok; ++q) ok = ao[k * nb + q] + 16 <= base || ao[k * nb + q] >= base + 2 * nb;
      if (ok) {
        p->eager_ok = true;
        p->eager_pix_base = base;
        p->d_eager_err = p->pool.alloc<double>(1);
      }
    }
  }
  CUDA_OK(cudaDeviceSynchronize());
  clk.lap("create: device structures");
  *out = up.release();
  p = nullptr;
  SFX_API_END(p)
}

void sfx_problem_destroy(sfx_problem* p) {
  if (!p) return;
  cudaSetDevice(p->device);
  delete p;
}

sfx_status sfx_update_params(sfx_problem* p, const sfx_params* params) {
  SFX_API_BEGIN
  SFX_CHECK(p && params, SFX_ERR_INVALID_ARG, "null argument");
  validate_params(*params);
  p->params = *params;
  SFX_API_END(p)
}

sfx_status sfx_set_values(sfx_problem* p, const double* values, int64_t n) {
  SFX_API_BEGIN
  SFX_CHECK(p && values, SFX_ERR_INVALID_ARG, "null argument");
  SFX_CHECK(n == p->a.n_values, SFX_ERR_INVALID_ARG, "values length mismatch");
  CUDA_OK(cudaSetDevice(p->device));
  if (p->a.world > 1 && !getenv("SFX_FULL_UPLOAD")) {
    // every rank holds the same Values (the sharded solve is SPMD): each uploads one slice over its own PCIe link and
    // the slices are exchanged over NVLink -- n / world doubles per rank through PCIe instead of n
    const int64_t c = p->values_chunk, lo = std::min<int64_t>(n, c * p->a.rank), hi = std::min<int64_t>(n, lo + c);
    if (hi > lo)
      CUDA_OK(cudaMemcpyAsync(p->d_cur_values + lo, values + lo, sizeof(double) * (hi - lo), cudaMemcpyHostToDevice, p->st));
    NCCL_OK(nccl().AllGather(p->d_cur_values + c * p->a.rank, p->d_cur_values, (size_t)c, ncclDouble, p->comm->comm, p->st));
  } else if (p->eager_ok) {
    // everything but the pixels first, then the pixels in chunks on the side stream; the main stream linearizes the
    // observations of a chunk (into state block kEagerBlock, reading the upload buffer) as soon as its event fires
    const LinBatch& lb = p->lin[0];
    const int blocks = linearize_bal_blocks(lb);
    const int n_chunks = std::max(1, std::min(8, blocks / 1024));
    while ((int)p->ev_up.size() < n_chunks + 1) {
      cudaEvent_t e;
      CUDA_OK(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
      p->ev_up.push_back(e);
    }
    const int64_t pb = p->eager_pix_base, pe = pb + 2 * (int64_t)lb.n;
    p->eager_valid = false;
    if (pb > 0) CUDA_OK(cudaMemcpyAsync(p->d_cur_values, values, sizeof(double) * pb, cudaMemcpyHostToDevice, p->st2));
    if (pe < n)
      CUDA_OK(cudaMemcpyAsync(p->d_cur_values + pe, values + pe, sizeof(double) * (n - pe), cudaMemcpyHostToDevice, p->st2));
    CUDA_OK(cudaEventRecord(p->ev_up[0], p->st2));
    StatePtrs spx = p->sp;
    spx.values[kEagerBlock] = p->d_cur_values;
    CUDA_OK(cudaStreamWaitEvent(p->st, p->ev_up[0], 0));
    launch_zero_lin(p->st, p->d_ctrl, spx, kLinModeEager, p->a.h_accum_values, p->a.N);
    for (int c = 0; c < n_chunks; ++c) {
      const int b0 = (int)((int64_t)blocks * c / n_chunks), b1 = (int)((int64_t)blocks * (c + 1) / n_chunks);
      const int64_t v0 = pb + 2 * 128 * (int64_t)b0, v1 = std::min(pe, pb + 2 * 128 * (int64_t)b1);
      CUDA_OK(cudaMemcpyAsync(p->d_cur_values + v0, values + v0, sizeof(double) * (v1 - v0), cudaMemcpyHostToDevice, p->st2));
      CUDA_OK(cudaEventRecord(p->ev_up[c + 1], p->st2));
      CUDA_OK(cudaStreamWaitEvent(p->st, p->ev_up[c + 1], 0));
      launch_linearize_bal_range(p->st, p->d_ctrl, spx, kLinModeEager, lb, p->d_partials, b0, b1);
    }
    launch_bal_point_finalize(p->st, p->d_ctrl, spx, kLinModeEager, lb);
    launch_finish_error(p->st, p->d_ctrl, kLinModeEager, p->d_partials, p->n_partials, p->d_eager_err);
    CUDA_OK(cudaStreamSynchronize(p->st2));
    CUDA_OK(cudaStreamSynchronize(p->st));
    p->eager_valid = true;
    p->lin0_clobbered = true;
    p->can_continue = false;  // (that state block no longer holds what the last Optimize left)
    p->values_set = true;
    return SFX_OK;
  } else {
    CUDA_OK(cudaMemcpyAsync(p->d_cur_values, values, sizeof(double) * n, cudaMemcpyHostToDevice, p->st));
  }
  CUDA_OK(cudaStreamSynchronize(p->st));
  p->eager_valid = false;
  p->values_set = true;
  SFX_API_END(p)
}

static sfx_status optimize_impl(sfx_problem* p, int32_t num_iterations, sfx_stats* stats, bool cont) {
  SFX_API_BEGIN
  SFX_CHECK(p, SFX_ERR_INVALID_ARG, "null problem");
  SFX_CHECK(p->values_set, SFX_ERR_INVALID_ARG, "sfx_set_values must be called first");
  if (num_iterations < 0) num_iterations = p->params.iterations;
  SFX_CHECK(num_iterations > 0, SFX_ERR_INVALID_ARG, "num_iterations must be positive");
  SFX_CHECK(num_iterations <= kMaxIterations, SFX_ERR_INVALID_ARG, "num_iterations exceeds stats capacity");
  CUDA_OK(cudaSetDevice(p->device));
  Analysis& a = p->a;
  const int64_t launches0 = g_launches;
  const int n_iters_before = cont ? p->h_ctrl->n_iters : 0;
  if (cont) {
    SFX_CHECK(p->can_continue, SFX_ERR_INVALID_ARG,
              "sfx_optimize_continue must directly follow sfx_optimize / sfx_optimize_continue (SYM_ASSERT: IsInitialized())");
    SFX_CHECK(p->h_ctrl->n_iters + num_iterations <= kMaxIterations, SFX_ERR_INVALID_ARG,
              "num_iterations exceeds stats capacity");
  }
  // debug_stats (optimizer_params_t, lcmtypes/symforce.lcm): keep values and residual of every record
  const bool dbg = p->params.debug_stats != 0;
  if (dbg) {
    SFX_CHECK(a.world == 1, SFX_ERR_UNSUPPORTED, "debug_stats snapshots are single-GPU only");
    const int need = (cont ? p->h_ctrl->n_iters : 0) + num_iterations + 1;
    if (need > p->dbg_cap) {
      double *nv = nullptr, *nr = nullptr, *nu = nullptr;
      if (cudaMalloc(&nv, sizeof(double) * (size_t)need * a.n_values) != cudaSuccess ||
          cudaMalloc(&nr, sizeof(double) * (size_t)need * std::max(a.M, 1)) != cudaSuccess ||
          cudaMalloc(&nu, sizeof(double) * (size_t)need * std::max(a.N, 1)) != cudaSuccess) {
        if (nv) cudaFree(nv);
        if (nr) cudaFree(nr);
        cudaGetLastError();
        throw Error(SFX_ERR_CUDA, "out of device memory for the debug_stats snapshots (iterations x Values)");
      }
      if (cont && p->dbg_valid) {  // keep the records of the stages before
        CUDA_OK(cudaMemcpyAsync(nv, p->dbg_values, sizeof(double) * (size_t)p->h_ctrl->n_iters * a.n_values,
                                cudaMemcpyDeviceToDevice, p->st));
        CUDA_OK(cudaMemcpyAsync(nr, p->dbg_res, sizeof(double) * (size_t)p->h_ctrl->n_iters * a.M, cudaMemcpyDeviceToDevice,
                                p->st));
        CUDA_OK(cudaMemcpyAsync(nu, p->dbg_upd, sizeof(double) * (size_t)p->h_ctrl->n_iters * a.N, cudaMemcpyDeviceToDevice,
                                p->st));
        CUDA_OK(cudaStreamSynchronize(p->st));
      }
      if (p->dbg_values) cudaFree(p->dbg_values);
      if (p->dbg_res) cudaFree(p->dbg_res);
      if (p->dbg_upd) cudaFree(p->dbg_upd);
      p->dbg_values = nv;
      p->dbg_res = nr;
      p->dbg_upd = nu;
      p->dbg_cap = need;
    }
  }
  p->dbg_valid = dbg && (cont ? p->dbg_valid || p->h_ctrl->n_iters == 0 : true);
  // the linearization sfx_set_values computed while uploading is the Init linearization of this run
  const bool adopt = !cont && p->eager_valid && !dbg;
  if (cont) {
    reset_ctrl_continue(p);
  } else {
    // Reset(values): all three state blocks hold the full values buffer; optimized keys are
    // overwritten by retract (levenberg_marquardt_solver.h:163-182, state ResetValues)
    reset_ctrl(p);
  }
  for (int b = 0; b < 3; ++b) launch_copy_values(p->st, p->sp.values[b], p->d_cur_values, a.n_values);
  // events
  const size_t need = (size_t)num_iterations * (PH_COUNT + 1) + 8;
  while (p->ev.size() < need) {
    cudaEvent_t e;
    CUDA_OK(cudaEventCreate(&e));
    p->ev.push_back(e);
  }
  p->ev_phase.clear();
  size_t evi = 0;
  auto mark = [&](int phase) {
    CUDA_OK(cudaEventRecord(p->ev[evi++], p->st));
    p->ev_phase.push_back(phase);
  };
  std::vector<size_t> iter_end_ev;
  mark(-1);
  int enq = 0;
  for (int i = 0; i < num_iterations; ++i) {
    if (i >= 2) {  // bound the run-ahead; the device always has >= 1 iteration queued
      // the status slot of iteration i-2 is written exactly once, by that iteration: every rank of a
      // sharded run reads the same value here and enqueues the same collectives
      CUDA_OK(cudaEventSynchronize(p->ev[iter_end_ev[i - 2]]));
      if (((volatile int*)p->h_done)[i - 2]) break;
    }
    launch_lm_begin(p->st, p->d_ctrl);
    if (i == 0) {
      if (adopt)
        launch_adopt_eager(p->st, p->d_ctrl, p->d_eager_err);
      else
        enqueue_linearize(p, /*mode=*/0);
      launch_lm_after_first_linearize(p->st, p->d_ctrl);
      if (dbg) launch_debug_snapshot(p->st, p->d_ctrl, p->sp, 1, a.n_values, a.M, p->dbg_cap, p->dbg_values, p->dbg_res,
                                     p->d_upd, p->d_ref2int, a.N, p->dbg_upd);
      mark(PH_LIN);
    }
    enqueue_solve(p, mark);
    launch_retract(p->st, p->d_ctrl, p->sp, p->d_key_type, p->d_key_voff, p->d_key_sdim, p->d_key_tdim,
                   p->d_key_itoff, a.n_keys, p->d_upd);
    mark(PH_UPDATE);
    enqueue_linearize(p, /*mode=*/1);
    if (dbg) launch_debug_snapshot(p->st, p->d_ctrl, p->sp, 0, a.n_values, a.M, p->dbg_cap, p->dbg_values, p->dbg_res,
                                     p->d_upd, p->d_ref2int, a.N, p->dbg_upd);
    mark(PH_LIN);
    launch_step_reduce(p->st, p->d_ctrl, p->sp, p->d_upd, p->d_dvec, p->d_last, a.N, p->d_partials,
                       (a.world > 1 && a.rank != 0) ? a.sp.reduced_dim : 0);
    if (a.world > 1) {
      double* red = (double*)((char*)p->d_ctrl + offsetof(Ctrl, red)) + 1;
      NCCL_OK(nccl().AllReduce(red, red, 4, ncclDouble, ncclSum, p->comm->comm, p->st));
    }
    launch_lm_end(p->st, p->d_ctrl, p->d_upd, p->d_last, a.N, p->d_done + i);
    mark(PH_UPDATE);
    iter_end_ev.push_back(evi - 1);
    enq++;
  }
  CUDA_OK(cudaMemcpyAsync(p->h_ctrl, p->d_ctrl, sizeof(Ctrl), cudaMemcpyDeviceToHost, p->st));
  CUDA_OK(cudaStreamSynchronize(p->st));
  {
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) throw Error(SFX_ERR_CUDA, std::string("kernel failure: ") + cudaGetErrorString(e));
  }
  const Ctrl& c = *p->h_ctrl;
  // optimizer_params_t::verbose / debug_checks: the reference logs from inside Iterate (spdlog, stdout); the records
  // live on the device until here, so the same lines come out after the run, on stderr
  if ((p->params.verbose || p->params.debug_checks) && a.rank == 0) {
    const int first = cont ? std::max(n_iters_before, 1) : 1;
    if (p->params.verbose) {
      double prev = c.n_iters > 0 ? c.iters[0].new_error : 0.0;
      for (int i = 1; i < c.n_iters; ++i) {
        const sfx_iteration& it = c.iters[i];
        const double gain = (prev - it.new_error) / (prev - it.new_error_linear);
        if (i >= first)
          std::fprintf(stderr,
                       "LM<sfx> [iter %4d] lambda: %.3e, error prev/linear/new: %.3e/%.3e/%.3e, rel reduction: %.5e, "
                       "gain ratio: %.5e\n",
                       it.iteration, it.current_lambda, prev, it.new_error_linear, it.new_error, it.relative_reduction, gain);
        if (it.update_accepted) prev = it.new_error;
      }
    }
    if (c.n_chol_fail > 0)
      std::fprintf(stderr, "LM<sfx> the Cholesky factorization met a non-positive pivot in %d iteration(s): those steps are "
                           "non-finite and were rejected (lambda = %.2e at the end)\n", c.n_chol_fail, c.lambda);
    if (p->params.debug_checks) {
      if (c.n_nonfinite_update > 0)
        std::fprintf(stderr, "LM<sfx> Encountered non-finite values in the update vector in %d iteration(s)\n",
                     c.n_nonfinite_update);
      if (c.n_zero_diag > 0) {
        std::string idx;
        for (int i = 0; i < std::min(c.n_zero_diag, 15); ++i) idx += (i ? ", " : "") + std::to_string(c.zero_diag_idx[i]);
        std::fprintf(stderr, "LM<sfx> Zero on diagonal after damping (epsilon = %.2e) at internal indices: [%s%s]\n", c.epsilon,
                     idx.c_str(), c.n_zero_diag > 15 ? (", ... (" + std::to_string(c.n_zero_diag - 15) + " omitted)").c_str() : "");
      }
    }
  }
  sfx_stats s{};
  s.status = c.done ? c.done : 2;  // HIT_ITERATION_LIMIT
  s.failure_reason = c.done == 3 ? c.failure_reason : 0;
  s.best_index = c.best_index;
  s.n_iterations = c.n_iters;
  p->last_stats = s;
  if (stats) *stats = s;
  // timings
  sfx_timings tm{};
  double acc[PH_COUNT] = {0, 0, 0, 0, 0};
  for (size_t i = 1; i < evi; ++i) {
    float ms = 0;
    if (cudaEventElapsedTime(&ms, p->ev[i - 1], p->ev[i]) == cudaSuccess && p->ev_phase[i] >= 0)
      acc[p->ev_phase[i]] += ms;
  }
  float tot = 0;
  cudaEventElapsedTime(&tot, p->ev[0], p->ev[evi - 1]);
  tm.total_ms = tot;
  tm.linearize_ms = acc[PH_LIN];
  tm.schur_ms = acc[PH_SCHUR];
  tm.factorize_ms = acc[PH_FACTOR];
  tm.solve_ms = acc[PH_SOLVE];
  tm.update_ms = acc[PH_UPDATE];
  tm.iterations_run = c.n_iters > 0 ? c.n_iters - 1 : 0;
  tm.n_linearize = tm.iterations_run + 1;
  tm.n_factorize = tm.iterations_run;
  tm.kernel_launches = (int32_t)(g_launches - launches0);
  p->tm = tm;
  p->can_continue = true;
  SFX_API_END(p)
}

sfx_status sfx_optimize(sfx_problem* p, int32_t num_iterations, sfx_stats* stats) {
  return optimize_impl(p, num_iterations, stats, false);
}

sfx_status sfx_optimize_continue(sfx_problem* p, int32_t num_iterations, sfx_stats* stats) {
  return optimize_impl(p, num_iterations, stats, true);
}

sfx_status sfx_relax_damping_to_initial(sfx_problem* p) {
  SFX_API_BEGIN
  SFX_CHECK(p, SFX_ERR_INVALID_ARG, "null problem");
  SFX_CHECK(p->can_continue, SFX_ERR_INVALID_ARG, "no optimization to continue (SYM_ASSERT: IsInitialized())");
  Ctrl* c = p->h_ctrl;  // uploaded by the next sfx_optimize_continue
  c->lambda = std::min(c->lambda, p->params.initial_lambda);
  c->nu = p->params.dynamic_lambda_update_beta;
  SFX_API_END(p)
}

sfx_status sfx_get_best_values(sfx_problem* p, double* values, int64_t n) {
  SFX_API_BEGIN
  SFX_CHECK(p && values, SFX_ERR_INVALID_ARG, "null argument");
  SFX_CHECK(n == p->a.n_values, SFX_ERR_INVALID_ARG, "values length mismatch");
  SFX_CHECK(p->h_ctrl->best_valid, SFX_ERR_INVALID_ARG, "SYM_ASSERT: state_.BestIsValid()");
  CUDA_OK(cudaSetDevice(p->device));
  const double* src = p->sp.values[p->h_ctrl->best_idx];
  if (p->a.world > 1) src = gather_sharded_values(p, src, n);
  CUDA_OK(cudaMemcpyAsync(values, src, sizeof(double) * n, cudaMemcpyDeviceToHost, p->st));
  CUDA_OK(cudaStreamSynchronize(p->st));
  SFX_API_END(p)
}

sfx_status sfx_update_best_values(sfx_problem* p, double* values, int64_t n, int64_t* bytes_copied) {
  SFX_API_BEGIN
  SFX_CHECK(p && values, SFX_ERR_INVALID_ARG, "null argument");
  SFX_CHECK(n == p->a.n_values, SFX_ERR_INVALID_ARG, "values length mismatch");
  SFX_CHECK(p->h_ctrl->best_valid, SFX_ERR_INVALID_ARG, "SYM_ASSERT: state_.BestIsValid()");
  CUDA_OK(cudaSetDevice(p->device));
  Analysis& a = p->a;
  if (p->opt_ranges.size() > 256) {
    // scattered keys: the full buffer is cheaper than thousands of small copies
    if (bytes_copied) *bytes_copied = (int64_t)sizeof(double) * n;
    return sfx_get_best_values(p, values, n);
  }
  const double* src = p->sp.values[p->h_ctrl->best_idx];
  if (a.world > 1) {
    if (!p->rank_lm_range.empty()) {
      src = gather_sharded_values(p, src, n);
    } else {
      // every rank holds the cameras and its own landmarks: a masked all-reduce per optimized range assembles them
      for (const auto& x : p->opt_ranges) {
        const int64_t len = x.second - x.first;
        launch_mask_values(p->st, src + x.first, p->d_vmask + x.first, len, p->d_stage + x.first);
        NCCL_OK(nccl().AllReduce(p->d_stage + x.first, p->d_stage + x.first, (size_t)len, ncclDouble, ncclSum,
                                 p->comm->comm, p->st));
      }
      src = p->d_stage;
    }
  }
  int64_t total = 0;
  for (const auto& x : p->opt_ranges) {
    CUDA_OK(cudaMemcpyAsync(values + x.first, src + x.first, sizeof(double) * (x.second - x.first),
                            cudaMemcpyDeviceToHost, p->st));
    total += (int64_t)sizeof(double) * (x.second - x.first);
  }
  CUDA_OK(cudaStreamSynchronize(p->st));
  if (bytes_copied) *bytes_copied = total;
  SFX_API_END(p)
}

sfx_status sfx_get_iteration_debug(sfx_problem* p, int32_t record, double* values, double* residual) {
  SFX_API_BEGIN
  SFX_CHECK(p, SFX_ERR_INVALID_ARG, "null problem");
  SFX_CHECK(p->dbg_valid, SFX_ERR_INVALID_ARG, "the last optimization did not run with optimizer_params_t::debug_stats");
  SFX_CHECK(record >= 0 && record < p->h_ctrl->n_iters && record < p->dbg_cap, SFX_ERR_INVALID_ARG, "no such iteration record");
  CUDA_OK(cudaSetDevice(p->device));
  const Analysis& a = p->a;
  if (values)
    CUDA_OK(cudaMemcpyAsync(values, p->dbg_values + (size_t)record * a.n_values, sizeof(double) * a.n_values,
                            cudaMemcpyDeviceToHost, p->st));
  if (residual)
    CUDA_OK(cudaMemcpyAsync(residual, p->dbg_res + (size_t)record * a.M, sizeof(double) * a.M, cudaMemcpyDeviceToHost, p->st));
  CUDA_OK(cudaStreamSynchronize(p->st));
  SFX_API_END(p)
}

sfx_status sfx_get_iteration_update(sfx_problem* p, int32_t record, double* update) {
  SFX_API_BEGIN
  SFX_CHECK(p && update, SFX_ERR_INVALID_ARG, "null argument");
  SFX_CHECK(p->dbg_valid, SFX_ERR_INVALID_ARG, "the last optimization did not run with optimizer_params_t::debug_stats");
  SFX_CHECK(record >= 0 && record < p->h_ctrl->n_iters && record < p->dbg_cap, SFX_ERR_INVALID_ARG, "no such iteration record");
  CUDA_OK(cudaSetDevice(p->device));
  CUDA_OK(cudaMemcpyAsync(update, p->dbg_upd + (size_t)record * p->a.N, sizeof(double) * p->a.N, cudaMemcpyDeviceToHost,
                          p->st));
  CUDA_OK(cudaStreamSynchronize(p->st));
  SFX_API_END(p)
}

sfx_status sfx_get_iteration_jacobian(sfx_problem* p, int32_t record, double* jacobian_values) {
  SFX_API_BEGIN
  SFX_CHECK(p && jacobian_values, SFX_ERR_INVALID_ARG, "null argument");
  SFX_CHECK(p->dbg_valid, SFX_ERR_INVALID_ARG, "the last optimization did not run with optimizer_params_t::debug_stats");
  SFX_CHECK(record >= 0 && record < p->h_ctrl->n_iters && record < p->dbg_cap, SFX_ERR_INVALID_ARG, "no such iteration record");
  CUDA_OK(cudaSetDevice(p->device));
  ensure_jacobian(p);
  // J is a pure function of the values: evaluated from the record's snapshot of the Values buffer, by the kernel
  // sfx_linearize_jacobian runs (the LM loop itself never forms J); the optimizer state is not touched
  const double* v = p->dbg_values + (size_t)record * p->a.n_values;
  for (size_t b = 0; b < p->lin.size(); ++b)
    launch_jacobian(p->st, v, p->lin[b], p->d_jac_base[b], p->d_jac_colnnz[b], p->d_export);
  CUDA_OK(cudaMemcpyAsync(jacobian_values, p->d_export, sizeof(double) * p->a.jac_nnz, cudaMemcpyDeviceToHost, p->st));
  CUDA_OK(cudaStreamSynchronize(p->st));
  SFX_API_END(p)
}

sfx_status sfx_get_iterations(sfx_problem* p, sfx_iteration* buf, int32_t capacity, int32_t* n) {
  SFX_API_BEGIN
  SFX_CHECK(p && n, SFX_ERR_INVALID_ARG, "null argument");
  *n = p->h_ctrl->n_iters;
  for (int i = 0; i < std::min(capacity, *n); ++i) buf[i] = p->h_ctrl->iters[i];
  SFX_API_END(p)
}

sfx_status sfx_get_dims(sfx_problem* p, int32_t* N, int32_t* M, int64_t* nnz) {
  SFX_API_BEGIN
  SFX_CHECK(p, SFX_ERR_INVALID_ARG, "null problem");
  if (N) *N = p->a.N;
  if (M) *M = p->a.M;
  if (nnz) {
    build_csc(p->a);
    *nnz = p->a.nnz;
  }
  SFX_API_END(p)
}

sfx_status sfx_get_hessian_pattern(sfx_problem* p, int32_t* outer, int32_t* inner) {
  SFX_API_BEGIN
  SFX_CHECK(p && outer && inner, SFX_ERR_INVALID_ARG, "null argument");
  build_csc(p->a);
  std::copy(p->a.csc_outer.begin(), p->a.csc_outer.end(), outer);
  std::copy(p->a.csc_inner.begin(), p->a.csc_inner.end(), inner);
  SFX_API_END(p)
}

sfx_status sfx_linearize(sfx_problem* p, double* residual, double* rhs, double* hessian_values) {
  SFX_API_BEGIN
  SFX_CHECK(p, SFX_ERR_INVALID_ARG, "null problem");
  SFX_CHECK(p->values_set, SFX_ERR_INVALID_ARG, "sfx_set_values must be called first");
  CUDA_OK(cudaSetDevice(p->device));
  reset_ctrl(p);
  launch_copy_values(p->st, p->sp.values[0], p->d_cur_values, p->a.n_values);
  enqueue_linearize(p, 0);
  export_linearization(p, 0, residual, rhs, hessian_values);
  SFX_API_END(p)
}

sfx_status sfx_get_jacobian_pattern(sfx_problem* p, int64_t* nnz, int32_t* outer, int32_t* inner) {
  SFX_API_BEGIN
  SFX_CHECK(p, SFX_ERR_INVALID_ARG, "null problem");
  build_jacobian_csc(p->a);
  if (nnz) *nnz = p->a.jac_nnz;
  if (outer) std::copy(p->a.jac_outer.begin(), p->a.jac_outer.end(), outer);
  if (inner) std::copy(p->a.jac_inner.begin(), p->a.jac_inner.end(), inner);
  SFX_API_END(p)
}

sfx_status sfx_linearize_jacobian(sfx_problem* p, double* jacobian_values) {
  SFX_API_BEGIN
  SFX_CHECK(p && jacobian_values, SFX_ERR_INVALID_ARG, "null argument");
  SFX_CHECK(p->values_set, SFX_ERR_INVALID_ARG, "sfx_set_values must be called first");
  CUDA_OK(cudaSetDevice(p->device));
  ensure_jacobian(p);
  // a pure function of the values last set: the LM state (control block, linearizations) is not touched
  for (size_t b = 0; b < p->lin.size(); ++b)
    launch_jacobian(p->st, p->d_cur_values, p->lin[b], p->d_jac_base[b], p->d_jac_colnnz[b], p->d_export);
  CUDA_OK(cudaMemcpyAsync(jacobian_values, p->d_export, sizeof(double) * p->a.jac_nnz, cudaMemcpyDeviceToHost, p->st));
  CUDA_OK(cudaStreamSynchronize(p->st));
  SFX_API_END(p)
}

sfx_status sfx_check_derivatives(sfx_problem* p, double* rel_errors, int32_t* ok, double* numerical_jacobian) {
  SFX_API_BEGIN
  SFX_CHECK(p && ok, SFX_ERR_INVALID_ARG, "null argument");
  SFX_CHECK(p->values_set, SFX_ERR_INVALID_ARG, "sfx_set_values must be called first");
  Analysis& a = p->a;
  SFX_CHECK(a.world == 1, SFX_ERR_UNSUPPORTED, "derivatives are checked on one GPU");
  const int N = a.N, M = a.M;
  SFX_CHECK((int64_t)M * N <= (int64_t(1) << 24) && N <= 4096, SFX_ERR_UNSUPPORTED,
            "check_derivatives forms the dense M x N Jacobian and N x N Hessian on the host and relinearizes 2 N times: "
            "a debugging aid for small problems (M * N <= 2^24, N <= 4096)");
  CUDA_OK(cudaSetDevice(p->device));
  ensure_jacobian(p);
  // the linearization to check: at the values last set, like sfx_linearize
  reset_ctrl(p);
  // Init and New hold the whole Values buffer: retract only writes the optimized keys of New
  for (int b = 0; b < 2; ++b) launch_copy_values(p->st, p->sp.values[b], p->d_cur_values, a.n_values);
  enqueue_linearize(p, 0);
  std::vector<double> res(M), rhs(N), Hv;
  build_csc(a);
  Hv.resize(a.nnz);
  export_linearization(p, 0, res.data(), rhs.data(), Hv.data());
  std::vector<double> Jv(a.jac_nnz);
  for (size_t b = 0; b < p->lin.size(); ++b)
    launch_jacobian(p->st, p->d_cur_values, p->lin[b], p->d_jac_base[b], p->d_jac_colnnz[b], p->d_export);
  CUDA_OK(cudaMemcpyAsync(Jv.data(), p->d_export, sizeof(double) * a.jac_nnz, cudaMemcpyDeviceToHost, p->st));
  // numerical Jacobian by central differences in the tangent space (util.h:97-127 with delta = sqrt(epsilon), as
  // derivative_checker.h:54-56 calls it): New = Init (+) (+-delta e_c), residual of New
  const double delta = std::sqrt(p->epsilon);
  std::vector<double> numJ((size_t)M * N), rp(M), rm(M);
  CUDA_OK(cudaMemsetAsync(p->d_upd, 0, sizeof(double) * N, p->st));
  int prev = -1;
  for (int c = 0; c < N; ++c) {
    const int ic = a.ref2int[c];
    for (int sgn = 0; sgn < 2; ++sgn) {
      launch_set_entry(p->st, p->d_upd, ic, sgn == 0 ? delta : -delta, prev);
      prev = ic;
      launch_retract(p->st, p->d_ctrl, p->sp, p->d_key_type, p->d_key_voff, p->d_key_sdim, p->d_key_tdim, p->d_key_itoff,
                     a.n_keys, p->d_upd);
      enqueue_linearize(p, 1);
      CUDA_OK(cudaMemcpyAsync(sgn == 0 ? rp.data() : rm.data(), p->sp.res[1], sizeof(double) * M, cudaMemcpyDeviceToHost,
                              p->st));
    }
    CUDA_OK(cudaStreamSynchronize(p->st));
    double* col = numJ.data() + (size_t)c * M;
    for (int r = 0; r < M; ++r) col[r] = ((rp[r] - res[r]) - (rm[r] - res[r])) / (2.0 * delta);
  }
  CUDA_OK(cudaMemsetAsync(p->d_upd, 0, sizeof(double) * N, p->st));
  // the state blocks hold perturbed values now: back to a clean slate (as after sfx_linearize)
  reset_ctrl(p);
  CUDA_OK(cudaStreamSynchronize(p->st));
  // Eigen's isApprox: |x - y|_F <= prec * min(|x|_F, |y|_F)
  auto rel = [](double diff2, double na2, double nb2) {
    const double m = std::sqrt(std::min(na2, nb2));
    return diff2 == 0.0 ? 0.0 : (m > 0.0 ? std::sqrt(diff2) / m : std::numeric_limits<double>::infinity());
  };
  // (1) numerical vs analytic Jacobian, 10 sqrt(epsilon) (derivative_checker.h:58-59)
  double dj = 0, nj = 0, nn = 0;
  for (int c = 0; c < N; ++c) {
    const double* col = numJ.data() + (size_t)c * M;
    int q = a.jac_outer[c];
    const int q1 = a.jac_outer[c + 1];
    for (int r = 0; r < M; ++r) {
      double ja = 0.0;
      if (q < q1 && a.jac_inner[q] == r) ja = Jv[q++];
      dj += (col[r] - ja) * (col[r] - ja);
      nj += ja * ja;
      nn += col[r] * col[r];
    }
  }
  const double e_j = rel(dj, nj, nn);
  // (2) hessian_lower (symmetrized) vs J^T J, sqrt(epsilon) (:78-98); (3) rhs vs J^T r, sqrt(epsilon) (:101-117)
  std::vector<std::vector<std::pair<int, double>>> rows(M);
  for (int c = 0; c < N; ++c)
    for (int q = a.jac_outer[c]; q < a.jac_outer[c + 1]; ++q) rows[a.jac_inner[q]].push_back({c, Jv[q]});
  std::vector<double> JtJ((size_t)N * N, 0.0), Jtr(N, 0.0);
  for (int r = 0; r < M; ++r)
    for (const auto& x : rows[r]) {
      Jtr[x.first] += x.second * res[r];
      for (const auto& y : rows[r])
        if (y.first >= x.first) JtJ[(size_t)y.first + (size_t)x.first * N] += x.second * y.second;  // lower triangle
    }
  double dh = 0, nh = 0, nq = 0;
  {
    std::vector<double> Hd((size_t)N * N, 0.0);
    for (int c = 0; c < N; ++c)
      for (int q = a.csc_outer[c]; q < a.csc_outer[c + 1]; ++q) Hd[(size_t)a.csc_inner[q] + (size_t)c * N] = Hv[q];
    for (int c = 0; c < N; ++c)
      for (int r = c; r < N; ++r) {
        const double x = Hd[(size_t)r + (size_t)c * N], y = JtJ[(size_t)r + (size_t)c * N];
        const double w = r == c ? 1.0 : 2.0;  // both triangles of the full matrices
        dh += w * (x - y) * (x - y);
        nh += w * x * x;
        nq += w * y * y;
      }
  }
  const double e_h = rel(dh, nh, nq);
  double dr = 0, nr = 0, nt = 0;
  for (int c = 0; c < N; ++c) {
    dr += (rhs[c] - Jtr[c]) * (rhs[c] - Jtr[c]);
    nr += rhs[c] * rhs[c];
    nt += Jtr[c] * Jtr[c];
  }
  const double e_r = rel(dr, nr, nt);
  if (rel_errors) {
    rel_errors[0] = e_j;
    rel_errors[1] = e_h;
    rel_errors[2] = e_r;
  }
  *ok = (e_j <= 10.0 * delta && e_h <= delta && e_r <= delta) ? 1 : 0;
  if (numerical_jacobian) std::copy(numJ.begin(), numJ.end(), numerical_jacobian);
  if (!*ok && p->params.verbose)
    std::fprintf(stderr, "[sfx] derivative check failed: |J_num - J| %.3e (tol %.3e), |H - J^T J| %.3e, |rhs - J^T r| %.3e (tol %.3e)\n",
                 e_j, 10.0 * delta, e_h, e_r, delta);
  SFX_API_END(p)
}

sfx_status sfx_get_best_linearization(sfx_problem* p, double* residual, double* rhs, double* hessian_values) {
  SFX_API_BEGIN
  SFX_CHECK(p, SFX_ERR_INVALID_ARG, "null problem");
  const Ctrl& c = *p->h_ctrl;
  SFX_CHECK(c.best_valid && c.lin_valid[c.best_idx] && !(p->lin0_clobbered && c.best_idx == kEagerBlock), SFX_ERR_INVALID_ARG,
            "SYM_ASSERT: state_.BestIsValid() && Best().GetLinearization().IsInitialized()");
  CUDA_OK(cudaSetDevice(p->device));
  export_linearization(p, c.best_idx, residual, rhs, hessian_values);
  SFX_API_END(p)
}

sfx_status sfx_compute_covariance(sfx_problem* p, const double* hessian_values, int32_t block_dim,
                                  double* covariance) {
  SFX_API_BEGIN
  SFX_CHECK(p && covariance, SFX_ERR_INVALID_ARG, "null argument");
  Analysis& a = p->a;
  SFX_CHECK(a.world == 1, SFX_ERR_UNSUPPORTED, "covariances are computed on one GPU");
  const int sys_dim = a.schur ? a.sp.reduced_dim : a.N;
  // a Schur problem inverts its reduced system (block-diagonal C, SparseSchurSolver::SInvInPlace); a problem solved
  // without Schur elimination yields any leading block: the general-C path of internal/covariance_utils.h:41-103
  // (S = B - E C^-1 E^T, covariance = S^-1), computed as the leading block_dim x block_dim block of
  // (H + epsilon on the diagonal of C)^-1 from block_dim solves with the sparse factor of the whole matrix
  SFX_CHECK(a.schur ? block_dim == sys_dim : (block_dim >= 1 && block_dim <= sys_dim), SFX_ERR_UNSUPPORTED,
            "covariance block must be the block the linear solver factors (all keys before the Schur-eliminated "
            "landmarks), or a leading block of a problem solved without Schur elimination");
  const int nb = block_dim;
  CUDA_OK(cudaSetDevice(p->device));
  Ctrl* c = p->h_ctrl;
  int blk;
  if (hessian_values != nullptr) {
    // a caller-provided Linearization::hessian_lower: scattered into a state block that is not Best
    blk = c->best_valid ? (c->best_idx + 1) % 3 : 0;
    ensure_csc(p);
    CUDA_OK(cudaMemcpyAsync(p->d_export, hessian_values, sizeof(double) * a.nnz, cudaMemcpyHostToDevice, p->st));
    CUDA_OK(cudaMemsetAsync(p->sp.H[blk], 0, sizeof(double) * a.H.n_values, p->st));
    launch_import_csc(p->st, p->d_export, p->d_csc_src, a.nnz, p->sp.H[blk]);
    c->lin_valid[blk] = 0;
  } else {
    SFX_CHECK(c->best_valid && c->lin_valid[c->best_idx], SFX_ERR_INVALID_ARG,
              "SYM_ASSERT: state_.BestIsValid() && Best().GetLinearization().IsInitialized()");
    blk = c->best_idx;
  }
  // run the solver kernels on that block: they select it through ctrl->init_idx and stop on ctrl->done
  const int saved_init = c->init_idx, saved_done = c->done, saved_fail = c->chol_fail;
  c->init_idx = blk;
  c->done = 0;
  c->chol_fail = 0;
  CUDA_OK(cudaMemcpyAsync(p->d_ctrl, c, offsetof(Ctrl, iters), cudaMemcpyHostToDevice, p->st));
  // damping of internal/covariance_utils.h:131-135 (epsilon on C only) resp.
  // LevenbergMarquardtSolver::ComputeCovariance (levenberg_marquardt_solver.tcc:345-356: epsilon everywhere)
  {
    std::vector<double> dv(a.N, p->epsilon);
    if (a.schur) std::fill(dv.begin(), dv.begin() + sys_dim, 0.0);
    if (!a.schur && nb < sys_dim)  // covariance_utils.h:131-135: only the marginalized block is damped
      for (int r = 0; r < nb; ++r) dv[a.ref2int[r]] = 0.0;
    CUDA_OK(cudaMemcpyAsync(p->d_dvec, dv.data(), sizeof(double) * a.N, cudaMemcpyHostToDevice, p->st));
    CUDA_OK(cudaStreamSynchronize(p->st));
  }
  enqueue_zero_fork(p);
  if (a.schur) launch_schur(p->st, p->d_ctrl, p->sp, p->sd, p->d_dvec);
  enqueue_factorize(p);
  // S^-1 = solves against the identity (SparseSchurSolver::SInvInPlace, sparse_schur_solver.tcc:165-170), column by column
  double *d_unit = nullptr, *d_cov = nullptr;
  CUDA_OK(cudaMalloc(&d_unit, sizeof(double) * sys_dim));
  if (cudaMalloc(&d_cov, sizeof(double) * (size_t)sys_dim * nb) != cudaSuccess) {
    cudaFree(d_unit);
    cudaGetLastError();
    throw Error(SFX_ERR_CUDA, "out of device memory for the covariance block");
  }
  CUDA_OK(cudaMemsetAsync(d_unit, 0, sizeof(double) * sys_dim, p->st));
  for (int r = 0; r < nb; ++r) {  // column r of the inverse, in the internal row order
    launch_set_unit(p->st, d_unit, a.ref2int[r], r > 0 ? a.ref2int[r - 1] : -1);
    enqueue_tri_solves(p, d_unit, 0);
    launch_unpermute(p->st, p->d_ctrl, p->fd, d_cov + (size_t)r * sys_dim, 1.0);
  }
  std::vector<double> cov_int((size_t)sys_dim * nb);
  CUDA_OK(cudaMemcpyAsync(cov_int.data(), d_cov, sizeof(double) * cov_int.size(), cudaMemcpyDeviceToHost, p->st));
  int fail = 0;
  CUDA_OK(cudaMemcpyAsync(&fail, (char*)p->d_ctrl + offsetof(Ctrl, chol_fail), sizeof(int), cudaMemcpyDeviceToHost, p->st));
  // restore the control block
  c->init_idx = saved_init;
  c->done = saved_done;
  c->chol_fail = saved_fail;
  CUDA_OK(cudaMemcpyAsync(p->d_ctrl, c, offsetof(Ctrl, iters), cudaMemcpyHostToDevice, p->st));
  CUDA_OK(cudaStreamSynchronize(p->st));
  cudaFree(d_unit);
  cudaFree(d_cov);
  SFX_CHECK(!fail, SFX_ERR_NUMERICAL, "the matrix to invert is not positive definite");
  // internal tangent order -> keys_ order
  for (int cj = 0; cj < nb; ++cj)
    for (int ri = 0; ri < nb; ++ri) covariance[ri + (size_t)cj * nb] = cov_int[a.ref2int[ri] + (size_t)cj * sys_dim];
  SFX_API_END(p)
}

sfx_status sfx_solve_step(sfx_problem* p, double lambda, double* update) {
  SFX_API_BEGIN
  SFX_CHECK(p && update, SFX_ERR_INVALID_ARG, "null argument");
  SFX_CHECK(p->values_set, SFX_ERR_INVALID_ARG, "sfx_set_values must be called first");
  CUDA_OK(cudaSetDevice(p->device));
  reset_ctrl(p);
  p->h_ctrl->lambda = lambda;
  CUDA_OK(cudaMemcpyAsync(p->d_ctrl, p->h_ctrl, offsetof(Ctrl, iters), cudaMemcpyHostToDevice, p->st));
  launch_copy_values(p->st, p->sp.values[0], p->d_cur_values, p->a.n_values);
  enqueue_linearize(p, 0);
  enqueue_solve(p, [](int) {});
  launch_permute_vec(p->st, p->d_upd, p->d_ref2int, p->a.N, p->d_y);
  CUDA_OK(cudaMemcpyAsync(update, p->d_y, sizeof(double) * p->a.N, cudaMemcpyDeviceToHost, p->st));
  CUDA_OK(cudaStreamSynchronize(p->st));
  SFX_API_END(p)
}

sfx_status sfx_get_ordering(sfx_problem* p, int32_t* perm, int32_t capacity, int32_t* n) {
  SFX_API_BEGIN
  SFX_CHECK(p && n, SFX_ERR_INVALID_ARG, "null argument");
  // elimination scalar position -> reference scalar index of the factored system
  const Analysis& a = p->a;
  std::vector<int> int2ref(a.N);
  for (int r = 0; r < a.N; ++r) int2ref[a.ref2int[r]] = r;
  *n = a.fp.n;
  for (int i = 0; i < std::min(capacity, *n); ++i) perm[i] = int2ref[a.fp.scalar_perm[i]];
  SFX_API_END(p)
}

sfx_status sfx_get_timings(sfx_problem* p, sfx_timings* out) {
  SFX_API_BEGIN
  SFX_CHECK(p && out, SFX_ERR_INVALID_ARG, "null argument");
  *out = p->tm;
  SFX_API_END(p)
}

// debug (host only): build and verify the tile task list of a front with `wt` pivot tiles out of `nt`;
// returns the number of tasks, or -1 with sfx_last_error(NULL) set
int32_t sfx_debug_verify_tasks(int32_t wt, int32_t nt, int32_t kc) {
  try {
    LargeFront x{};
    x.wt = wt;
    x.nt = nt;
    std::vector<LargeTask> tl;
    build_front_tasks(x, 0, kc, tl);
    verify_task_list(x, tl);
    return (int32_t)tl.size();
  } catch (const std::exception& e) {
    g_create_err = e.what();
    return -1;
  }
}

// debug (host only): the tile task list of such a front as int16 records {type, k, i, j, k1}; returns the number of
// tasks (call with out == NULL to size the buffer).  Input of tools/simulate_tile_dag.py.
int32_t sfx_debug_front_tasks(int32_t wt, int32_t nt, int32_t kc, int16_t* out, int32_t capacity) {
  try {
    LargeFront x{};
    x.wt = wt;
    x.nt = nt;
    std::vector<LargeTask> tl;
    build_front_tasks(x, 0, kc, tl);
    if (out)
      for (size_t q = 0; q < tl.size() && (int32_t)q < capacity; ++q) {
        out[5 * q + 0] = (int16_t)tl[q].type;
        out[5 * q + 1] = tl[q].k;
        out[5 * q + 2] = tl[q].i;
        out[5 * q + 3] = tl[q].j;
        out[5 * q + 4] = tl[q].k1;
      }
    return (int32_t)tl.size();
  } catch (const std::exception& e) {
    g_create_err = e.what();
    return -1;
  }
}

// debug: the reduced camera system S (block values, n from the first call with out == NULL) as the last Schur pass left it
sfx_status sfx_debug_read_S(sfx_problem* p, double* out, int64_t* n) {
  SFX_API_BEGIN
  SFX_CHECK(p && n && p->a.schur, SFX_ERR_INVALID_ARG, "not a Schur problem");
  *n = p->a.sp.S.n_values;
  if (out) {
    CUDA_OK(cudaSetDevice(p->device));
    CUDA_OK(cudaMemcpyAsync(out, p->sd.S, sizeof(double) * (size_t)*n, cudaMemcpyDeviceToHost, p->st));
    CUDA_OK(cudaStreamSynchronize(p->st));
  }
  SFX_API_END(p)
}

// debug: the frontal matrices as the last factorization left them; *n = number of doubles; geometry of the large fronts
// as int64 records {off, m, w, wt, nt} in `lf_info` (capacity lf_cap records, *n_lf = count)
sfx_status sfx_debug_read_fronts(sfx_problem* p, double* out, int64_t* n, int64_t* lf_info, int32_t lf_cap, int32_t* n_lf) {
  SFX_API_BEGIN
  SFX_CHECK(p && n, SFX_ERR_INVALID_ARG, "null argument");
  *n = p->a.fp.front_values;
  CUDA_OK(cudaSetDevice(p->device));
  if (out) {
    CUDA_OK(cudaMemcpyAsync(out, p->fd.fronts, sizeof(double) * (size_t)*n, cudaMemcpyDeviceToHost, p->st));
    CUDA_OK(cudaStreamSynchronize(p->st));
  }
  if (lf_info && n_lf) {
    std::vector<LargeFront> lfs(p->n_large_fronts);
    CUDA_OK(cudaMemcpy(lfs.data(), p->ld.lf, sizeof(LargeFront) * lfs.size(), cudaMemcpyDeviceToHost));
    *n_lf = (int32_t)lfs.size();
    for (int i = 0; i < std::min<int>(lf_cap, *n_lf); ++i) {
      lf_info[5 * i + 0] = lfs[i].off;
      lf_info[5 * i + 1] = lfs[i].m;
      lf_info[5 * i + 2] = lfs[i].w;
      lf_info[5 * i + 3] = lfs[i].wt;
      lf_info[5 * i + 4] = lfs[i].nt;
    }
  }
  SFX_API_END(p)
}

// debug: {chol_fail, fail_where} of the control block on the device right now
sfx_status sfx_debug_chol_fail(sfx_problem* p, int32_t out[2]) {
  SFX_API_BEGIN
  SFX_CHECK(p && out, SFX_ERR_INVALID_ARG, "null argument");
  CUDA_OK(cudaSetDevice(p->device));
  CUDA_OK(cudaMemcpyAsync(out, (char*)p->d_ctrl + offsetof(Ctrl, chol_fail), 2 * sizeof(int), cudaMemcpyDeviceToHost, p->st));
  CUDA_OK(cudaStreamSynchronize(p->st));
  SFX_API_END(p)
}

// debug (host only): plans the large-front path of a problem for `workers` resident CTAs and reports
// "fused_T0 n_large_fronts n_tasks n_fused_tasks n_ea_tasks n_jobs"; the fused list has passed verify_fused_list.
// Returns 0, or 1 with sfx_last_error(NULL) set.
int32_t sfx_debug_large_plan(const sfx_problem_desc* desc, int32_t workers, int64_t out[6]) {
  try {
    sfx_problem pr;
    pr.params = desc->params;
    analyze_problem(*desc, pr.a);
    Analysis& a = pr.a;
    const BlockMatrix& sys = a.schur ? a.sp.S : a.H;
    std::vector<int> sys2ref;
    if (desc->ordering == SFX_ORDERING_METIS_SCALAR) {
      std::vector<int> int2ref(a.N);
      for (int r = 0; r < a.N; ++r) int2ref[a.ref2int[r]] = r;
      sys2ref.assign(int2ref.begin(), int2ref.begin() + sys.node_off[sys.n_nodes]);
    }
    choose_front_plan(&pr, sys, desc->ordering, sys2ref, workers);
    LargeHostPlan hp;
    plan_large_fronts(&pr, workers, hp);
    int64_t n_ea = 0;
    for (int q = pr.fused_t0; q < pr.fused_t1; ++q) n_ea += hp.tasks[q].type == 6;
    out[0] = pr.fused_T0;
    out[1] = (int64_t)hp.lfs.size();
    out[2] = (int64_t)hp.tasks.size();
    out[3] = pr.fused_t1 - pr.fused_t0;
    out[4] = n_ea;
    out[5] = (int64_t)hp.jobs.size();
    return 0;
  } catch (const std::exception& e) {
    g_create_err = e.what();
    return 1;
  }
}

// debug: average time of `reps` linearizations of state block 0 (zero + kernels + error reduce), with
// parts of the BAL kernel left out when skip != 0 (timing experiment; leaves an invalid linearization)
sfx_status sfx_debug_time_linearize(sfx_problem* p, int32_t skip, int32_t reps, float* ms) {
  SFX_API_BEGIN
  SFX_CHECK(p && ms && reps > 0, SFX_ERR_INVALID_ARG, "bad argument");
  CUDA_OK(cudaSetDevice(p->device));
  reset_ctrl(p);
  for (int b = 0; b < 3; ++b) launch_copy_values(p->st, p->sp.values[b], p->d_cur_values, p->a.n_values);
  cudaEvent_t e0, e1;
  CUDA_OK(cudaEventCreate(&e0));
  CUDA_OK(cudaEventCreate(&e1));
  sfx::g_lin_skip = skip;
  enqueue_linearize(p, 1);
  CUDA_OK(cudaEventRecord(e0, p->st));
  for (int r = 0; r < reps; ++r) enqueue_linearize(p, 1);
  CUDA_OK(cudaEventRecord(e1, p->st));
  sfx::g_lin_skip = 0;
  CUDA_OK(cudaStreamSynchronize(p->st));
  float t = 0;
  CUDA_OK(cudaEventElapsedTime(&t, e0, e1));
  *ms = t / reps;
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  SFX_API_END(p)
}

// debug: trace the tile-DAG tasks of the next factorizations (buffer of n_tasks*4 u64, device)
sfx_status sfx_debug_diag_stamps(unsigned long long* out) {
  sfx::get_diag_stamps(out);
  return SFX_OK;
}
sfx_status sfx_debug_trace_tasks(sfx_problem* p, unsigned long long* host_out, int32_t* n_tasks, int16_t* task_info) {
  SFX_API_BEGIN
  static unsigned long long* dbuf = nullptr;
  static int64_t ntask = 0;
  if (!host_out) {  // arm
    ntask = 0;
    for (auto& lv : p->lvl_large) ntask = std::max<int64_t>(ntask, lv.t1);
    ntask = std::max<int64_t>(ntask, p->fused_t1);
    CUDA_OK(cudaMalloc(&dbuf, sizeof(unsigned long long) * 4 * ntask));
    CUDA_OK(cudaMemset(dbuf, 0, sizeof(unsigned long long) * 4 * ntask));
    CUDA_OK(cudaDeviceSynchronize());  // legacy-stream memset vs the non-blocking problem streams
    sfx::set_factor_trace(dbuf);
    *n_tasks = (int32_t)ntask;
  } else {
    CUDA_OK(cudaMemcpy(host_out, dbuf, sizeof(unsigned long long) * 4 * ntask, cudaMemcpyDeviceToHost));
    std::vector<LargeTask> tk(ntask);
    CUDA_OK(cudaMemcpy(tk.data(), p->ld.tasks, sizeof(LargeTask) * ntask, cudaMemcpyDeviceToHost));
    for (int64_t i = 0; i < ntask; ++i) {
      task_info[i * 5 + 0] = (int16_t)tk[i].lf;
      task_info[i * 5 + 1] = tk[i].type;
      task_info[i * 5 + 2] = tk[i].k;
      task_info[i * 5 + 3] = tk[i].i;
      task_info[i * 5 + 4] = tk[i].j;
    }
    sfx::set_factor_trace(nullptr);
  }
  SFX_API_END(p)
}

sfx_status sfx_get_info(sfx_problem* p, int64_t* out, int32_t capacity) {
  SFX_API_BEGIN
  SFX_CHECK(p && out, SFX_ERR_INVALID_ARG, "null argument");
  const Analysis& a = p->a;
  int64_t v[SFX_INFO_COUNT] = {0};
  v[SFX_INFO_N] = a.N;
  v[SFX_INFO_M] = a.M;
  v[SFX_INFO_NNZ_H] = a.csc_built ? a.nnz : -1;
  v[SFX_INFO_NUM_NODES] = (int64_t)a.nodes.size();
  v[SFX_INFO_REDUCED_DIM] = a.fp.n;
  v[SFX_INFO_NNZ_L] = a.fp.nnz_L;
  v[SFX_INFO_NUM_SUPERNODES] = a.fp.n_fronts;
  v[SFX_INFO_NUM_LEVELS] = a.fp.n_levels;
  v[SFX_INFO_FACTOR_FLOPS] = (int64_t)a.fp.flops;
  v[SFX_INFO_S_BLOCKS] = a.schur ? (int64_t)a.sp.S.row_idx.size() : 0;
  v[SFX_INFO_SCHUR_PAIRS] = a.schur ? (int64_t)a.sp.m_lm.size() : 0;
  v[SFX_INFO_MAX_FRONT] = a.fp.max_front;
  v[SFX_INFO_DEVICE_BYTES] = p->pool.bytes;
  v[SFX_INFO_CHOL_FAILURES] = p->h_ctrl->n_chol_fail;
  v[SFX_INFO_NONFINITE_UPDATES] = p->h_ctrl->n_nonfinite_update;
  v[SFX_INFO_ZERO_DIAGONAL] = p->h_ctrl->n_zero_diag;
  v[SFX_INFO_PLAN] = p->plan_nd_depth;
  v[SFX_INFO_REF_ORDERING_FLOPS] = (int64_t)p->ref_plan_flops;
  for (int i = 0; i < std::min<int>(capacity, SFX_INFO_COUNT); ++i) out[i] = v[i];
  SFX_API_END(p)
}

sfx_status sfx_comm_unique_id(char id_out[128]) {
  sfx_problem* p = nullptr;
  SFX_API_BEGIN
  SFX_CHECK(id_out, SFX_ERR_INVALID_ARG, "null argument");
  static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId size");
  ncclUniqueId id;
  NCCL_OK(nccl().GetUniqueId(&id));
  std::memcpy(id_out, &id, 128);
  SFX_API_END(p)
}

sfx_status sfx_comm_create(const char id[128], int32_t rank, int32_t world, int32_t device, sfx_comm** out) {
  sfx_problem* p = nullptr;
  SFX_API_BEGIN
  SFX_CHECK(id && out && world >= 1 && rank >= 0 && rank < world, SFX_ERR_INVALID_ARG, "bad communicator arguments");
  CUDA_OK(cudaSetDevice(device));
  ncclUniqueId uid;
  std::memcpy(&uid, id, 128);
  std::unique_ptr<sfx_comm> c(new sfx_comm());
  c->rank = rank;
  c->world = world;
  c->device = device;
  NCCL_OK(nccl().CommInitRank(&c->comm, world, uid, rank));
  *out = c.release();
  SFX_API_END(p)
}

void sfx_comm_destroy(sfx_comm* c) {
  if (!c) return;
  if (c->comm) nccl().CommDestroy(c->comm);
  delete c;
}

}  // extern "C"
