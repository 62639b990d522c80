// Host-side structural analysis, done ONCE per problem (replaces the index building of
// sym::Linearizer::BuildInitialLinearization, symforce/opt/linearizer.cc:149-356, and of
// SparseSchurSolver::ComputeSymbolicSparsity, symforce/opt/sparse_schur_solver.tcc:16-97):
//   keys -> nodes, block-sparse Hessian layout in HBM, per-factor scatter indices,
//   Schur match lists, CSC export map (reference layout of Linearization::hessian_lower).
#include <algorithm>
#include <atomic>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <exception>
#include <thread>
#include <map>
#include <memory>
#include <numeric>
#include <type_traits>
#include <unordered_map>

#include "sfx_internal.h"

namespace sfx {

int BlockMatrix::find(int row, int col) const {
  const int* b = row_idx.data() + col_ptr[col];
  const int* e = row_idx.data() + col_ptr[col + 1];
  const int* it = std::lower_bound(b, e, row);
  if (it == e || *it != row) return -1;
  return (int)(it - row_idx.data());
}

// order[] = stable ascending order of keys[]: LSD radix sort with 16-bit digits (digits on which all keys agree are
// skipped).  Replaces std::stable_sort with an indirect comparator on the 5 M contributions / 16.6 M Schur matches of
// a Final-shape BAL problem (2.9 s -> 0.5 s of the host analysis).
static void stable_order_by_key(const uint64_t* keys, size_t n, std::vector<uint32_t>& order) {
  order.resize(n);
  std::iota(order.begin(), order.end(), 0u);
  if (n < 2 || std::is_sorted(keys, keys + n)) return;  // BAL files list observations by (camera, point)
  uint64_t all_or = 0, all_and = ~0ull;
  for (size_t i = 0; i < n; ++i) {
    all_or |= keys[i];
    all_and &= keys[i];
  }
  std::vector<uint32_t> tmp(n);
  std::vector<size_t> cnt(1 << 16);
  for (int shift = 0; shift < 64; shift += 16) {
    if ((((all_or ^ all_and) >> shift) & 0xffff) == 0) continue;  // same digit everywhere
    std::fill(cnt.begin(), cnt.end(), 0);
    for (size_t i = 0; i < n; ++i) cnt[(keys[i] >> shift) & 0xffff]++;
    size_t run = 0;
    for (size_t b = 0; b < cnt.size(); ++b) {
      const size_t c = cnt[b];
      cnt[b] = run;
      run += c;
    }
    for (size_t i = 0; i < n; ++i) {
      const uint32_t o = order[i];
      tmp[cnt[(keys[o] >> shift) & 0xffff]++] = o;
    }
    order.swap(tmp);
  }
}

// Host threads for `n` work items of which a thread should get at least `min_chunk`: up to 8 (the index building of a
// 5 M-factor problem is bound by cache misses on the caller's arrays), SFX_HOST_THREADS overrides.
static int host_threads(int64_t n, int64_t min_chunk) {
  int nt = (int)std::min<int64_t>(std::min<unsigned>(8u, std::max(1u, std::thread::hardware_concurrency())),
                                  (n + min_chunk - 1) / min_chunk);
  if (getenv("SFX_HOST_THREADS")) nt = atoi(getenv("SFX_HOST_THREADS"));
  return std::max(1, nt);
}

// fn(t) for t in [0, nt) on nt threads (the caller's thread runs t = 0); exceptions thrown by a thread are rethrown
// on the calling thread
template <typename Fn>
static void run_threads(int nt, const Fn& fn) {
  if (nt <= 1) {
    fn(0);
    return;
  }
  std::vector<std::thread> th;
  std::vector<std::exception_ptr> err(nt);
  auto body = [&](int t) {
    try {
      fn(t);
    } catch (...) {
      err[t] = std::current_exception();
    }
  };
  for (int t = 1; t < nt; ++t) th.emplace_back(body, t);
  body(0);
  for (auto& x : th) x.join();
  for (auto& e : err)
    if (e) std::rethrow_exception(e);
}

// fn(begin, end) over contiguous chunks of [0, n)
template <typename Fn>
static void parallel_chunks(int64_t n, const Fn& fn) {
  const int nt = host_threads(n, 1 << 16);
  run_threads(nt, [&](int t) { fn(n * t / nt, n * (t + 1) / nt); });
}

// Uninitialised array of a trivial type: std::vector would zero-fill hundreds of MB on one thread before the host
// threads overwrite every element (and take the page faults in parallel)
template <typename T>
struct RawBuf {
  static_assert(std::is_trivially_copyable<T>::value, "RawBuf holds trivial types");
  std::unique_ptr<T[]> p;
  size_t n = 0;
  RawBuf() = default;
  explicit RawBuf(size_t n_) : p(new T[n_]), n(n_) {}
  T* data() { return p.get(); }
  const T* data() const { return p.get(); }
  size_t size() const { return n; }
  T& operator[](size_t i) { return p[i]; }
  const T& operator[](size_t i) const { return p[i]; }
  void reset() {
    p.reset();
    n = 0;
  }
};

static inline uint64_t mix64(uint64_t x) {
  x += 0x9e3779b97f4a7c15ull;
  x = (x ^ (x >> 30)) * 0xbf58476d1ce4e5b9ull;
  x = (x ^ (x >> 27)) * 0x94d049bb133111ebull;
  return x ^ (x >> 31);
}

struct FactorRef {
  int batch, idx;  // input batch / position
};

void analyze_problem(const sfx_problem_desc& d, Analysis& a) {
  PhaseClock clk;
  SFX_CHECK(d.abi_version == SFX_ABI_VERSION, SFX_ERR_INVALID_ARG, "ABI version mismatch");
  SFX_CHECK(d.n_keys > 0 && d.keys, SFX_ERR_INVALID_ARG, "no optimized keys");
  SFX_CHECK(d.n_batches > 0 && d.batches, SFX_ERR_INVALID_ARG, "no factors");
  a.n_keys = d.n_keys;
  a.n_values = d.n_values;
  a.n_factors = d.n_factors;
  a.schur = d.solver == SFX_SOLVER_SCHUR;
  const int nk = d.n_keys;
  const int n_lm_keys = a.schur ? d.schur_num_keys : 0;
  SFX_CHECK(!a.schur || (n_lm_keys > 0 && n_lm_keys < nk), SFX_ERR_INVALID_ARG,
            "schur_num_keys must be in (0, n_keys)");
  const int first_lm_key = nk - n_lm_keys;

  a.keys.resize(nk);
  int toff = 0;
  for (int k = 0; k < nk; ++k) {
    const sfx_key_entry& e = d.keys[k];
    SFX_CHECK(e.type == SFX_TYPE_VECTOR || e.type == SFX_TYPE_ROT3 || e.type == SFX_TYPE_POSE3, SFX_ERR_UNSUPPORTED,
              "key type has no device retract");
    SFX_CHECK(e.tangent_dim > 0 && e.tangent_dim <= kMaxNodeDim, SFX_ERR_UNSUPPORTED, "tangent dim out of range");
    SFX_CHECK(e.offset >= 0 && (int64_t)e.offset + e.storage_dim <= d.n_values, SFX_ERR_INVALID_ARG,
              "key storage outside the values buffer");
    if (e.type == SFX_TYPE_ROT3) SFX_CHECK(e.storage_dim == 4 && e.tangent_dim == 3, SFX_ERR_INVALID_ARG, "Rot3 dims");
    if (e.type == SFX_TYPE_POSE3) SFX_CHECK(e.storage_dim == 7 && e.tangent_dim == 6, SFX_ERR_INVALID_ARG, "Pose3 dims");
    if (e.type == SFX_TYPE_VECTOR) SFX_CHECK(e.storage_dim == e.tangent_dim, SFX_ERR_INVALID_ARG, "vector dims");
    a.keys[k] = KeyInfo{e.type, e.offset, e.storage_dim, e.tangent_dim, toff, -1, 0};
    toff += e.tangent_dim;
  }
  a.N = toff;

  // ---- factor list in caller order; validate; key signatures ------------------------------------
  std::vector<FactorRef> fref(d.n_factors, FactorRef{-1, -1});
  int64_t total = 0;
  for (int b = 0; b < d.n_batches; ++b) {
    const sfx_factor_batch& fb = d.batches[b];
    SFX_CHECK(fb.kind >= 0 && fb.kind < SFX_NUM_KINDS, SFX_ERR_UNSUPPORTED, "factor kind has no device implementation");
    total += fb.n;
    for (int f = 0; f < fb.n; ++f) {
      int fi = fb.factor_index[f];
      SFX_CHECK(fi >= 0 && fi < d.n_factors && fref[fi].batch < 0, SFX_ERR_INVALID_ARG,
                "factor_index must be a permutation of [0, n_factors)");
      fref[fi] = FactorRef{b, f};
    }
  }
  SFX_CHECK(total == d.n_factors, SFX_ERR_INVALID_ARG, "n_factors != sum of batch sizes");

  // per key: order-independent signature of the set of factors that optimize it (sums of hashes, accumulated with
  // relaxed atomic adds from all host threads: integer addition commutes, so the result does not depend on the schedule)
  std::vector<uint64_t> h1(nk, 0), h2(nk, 0);
  std::vector<int> cnt(nk, 0);
  for (int b = 0; b < d.n_batches; ++b) {
    const sfx_factor_batch& fb = d.batches[b];
    const sfx_kind_meta& km = SFX_KIND_META[fb.kind];
    parallel_chunks(fb.n, [&](int64_t f0, int64_t f1) {
      for (int o = 0; o < km.n_opt; ++o)
        for (int64_t f = f0; f < f1; ++f) {
          int key = fb.opt_keys[(int64_t)o * fb.n + f];
          if (key < 0) continue;
          SFX_CHECK(key < nk, SFX_ERR_INVALID_ARG, "opt key index out of range");
          SFX_CHECK(a.keys[key].tdim == km.opt_dims[o], SFX_ERR_INVALID_ARG, "tangent dim of key does not match factor");
          uint64_t fi = (uint64_t)fb.factor_index[f];
          __atomic_fetch_add(&h1[key], mix64(fi), __ATOMIC_RELAXED);
          __atomic_fetch_add(&h2[key], mix64(fi ^ 0x5bd1e995deadbeefull), __ATOMIC_RELAXED);
          __atomic_fetch_add(&cnt[key], 1, __ATOMIC_RELAXED);
        }
      for (int ar = 0; ar < km.n_args; ++ar)
        if (km.arg_used[ar])
          for (int64_t f = f0; f < f1; ++f) {
            int off = fb.arg_offsets[(int64_t)ar * fb.n + f];
            SFX_CHECK(off >= 0 && (int64_t)off + km.arg_dims[ar] <= d.n_values, SFX_ERR_INVALID_ARG,
                      "factor argument outside the values buffer");
          }
    });
  }
  for (int k = 0; k < nk; ++k)
    if (cnt[k] == 0)
      throw Error(SFX_ERR_STRUCTURE,
                  "Key #" + std::to_string(k) + " is in the state vector but is not optimized by any factor.");

  clk.lap("factor list + signatures");
  // ---- nodes: merge non-landmark keys with identical factor sets ---------------------------------
  {
    struct Sig {
      uint64_t a, b;
      int c;
      bool operator<(const Sig& o) const { return std::tie(a, b, c) < std::tie(o.a, o.b, o.c); }
    };
    std::map<Sig, int> sig2node;
    a.nodes.clear();
    for (int k = 0; k < nk; ++k) {
      int node = -1;
      if (k < first_lm_key) {
        Sig s{h1[k], h2[k], cnt[k]};
        auto it = sig2node.find(s);
        if (it != sig2node.end() && a.nodes[it->second].dim + a.keys[k].tdim <= kMaxNodeDim) {
          node = it->second;
        } else {
          node = (int)a.nodes.size();
          a.nodes.push_back(NodeInfo{0, 0, k, 0});
          sig2node[s] = node;
        }
      } else {
        SFX_CHECK(a.keys[k].tdim <= kMaxLandmarkDim, SFX_ERR_UNSUPPORTED, "Schur landmark dim > 3");
        node = (int)a.nodes.size();
        a.nodes.push_back(NodeInfo{0, 0, k, 0});
      }
      a.keys[k].node = node;
      a.keys[k].sub = a.nodes[node].dim;
      a.nodes[node].dim += a.keys[k].tdim;
      a.nodes[node].n_keys++;
    }
    int off = 0;
    for (auto& n : a.nodes) {
      n.toff = off;
      off += n.dim;
    }
    a.ref2int.resize(a.N);
    for (int k = 0; k < nk; ++k)
      for (int i = 0; i < a.keys[k].tdim; ++i)
        a.ref2int[a.keys[k].ref_toff + i] = a.nodes[a.keys[k].node].toff + a.keys[k].sub + i;
  }
  const int nn = (int)a.nodes.size();
  int first_lm_node = nn;
  if (a.schur) first_lm_node = a.keys[first_lm_key].node;

  clk.lap("nodes");
  // ---- batch plans: split by (kind, grouping pattern) -------------------------------------------
  struct PatternKey {
    int kind;
    int grp[SFX_MAX_OPT], sub[SFX_MAX_OPT];
    bool operator<(const PatternKey& o) const {
      return std::tie(kind, grp[0], grp[1], grp[2], sub[0], sub[1], sub[2]) <
             std::tie(o.kind, o.grp[0], o.grp[1], o.grp[2], o.sub[0], o.sub[1], o.sub[2]);
    }
  };
  std::vector<std::vector<FactorRef>> plan_factors;
  std::vector<PatternKey> plan_pat;
  // residual offsets in caller order
  std::vector<int32_t> res_off_of_factor(d.n_factors);
  {
    int r = 0;
    for (int fi = 0; fi < d.n_factors; ++fi) {
      res_off_of_factor[fi] = r;
      r += SFX_KIND_META[d.batches[fref[fi].batch].kind].res_dim;
    }
    a.M = r;
  }
  // pattern of one factor: which of its optimized keys fall into the same node (group) and where inside it
  auto pattern_of = [&](int fi, PatternKey& pk) {
    const sfx_factor_batch& fb = d.batches[fref[fi].batch];
    const sfx_kind_meta& km = SFX_KIND_META[fb.kind];
    const int f = fref[fi].idx;
    pk.kind = fb.kind;
    int nodes_seen[SFX_MAX_OPT];
    int ng = 0;
    for (int o = 0; o < SFX_MAX_OPT; ++o) {
      pk.grp[o] = -2;
      pk.sub[o] = 0;
    }
    for (int o = 0; o < km.n_opt; ++o) {
      int key = fb.opt_keys[(int64_t)o * fb.n + f];
      if (key < 0) {
        pk.grp[o] = -1;
        continue;
      }
      int node = a.keys[key].node;
      int g = -1;
      for (int q = 0; q < ng; ++q)
        if (nodes_seen[q] == node) g = q;
      if (g < 0) {
        g = ng;
        nodes_seen[ng++] = node;
      } else {
        // the same key twice in one factor is not representable
        for (int o2 = 0; o2 < o; ++o2)
          SFX_CHECK(fb.opt_keys[(int64_t)o2 * fb.n + f] != key, SFX_ERR_UNSUPPORTED,
                    "factor references the same optimized key twice");
      }
      pk.grp[o] = g;
      pk.sub[o] = a.keys[key].sub;
    }
  };
  // (1) packed pattern of every factor, on all host threads (the key -> node lookups miss the cache);
  // (2) serial pass in caller order: plans are numbered by first appearance, slots keep the caller's order
  static_assert(SFX_MAX_OPT == 3 && kMaxNodeDim <= 255 && SFX_NUM_KINDS <= 255, "pattern packing");
  RawBuf<uint64_t> fpat(d.n_factors);
  parallel_chunks(d.n_factors, [&](int64_t f0, int64_t f1) {
    PatternKey pk;
    for (int64_t fi = f0; fi < f1; ++fi) {
      pattern_of((int)fi, pk);
      uint64_t v = (uint64_t)pk.kind;
      for (int o = 0; o < SFX_MAX_OPT; ++o) v = (v << 11) | ((uint64_t)(pk.grp[o] + 2) << 8) | (uint64_t)pk.sub[o];
      fpat[fi] = v;
    }
  });
  {
    std::map<uint64_t, int> packed2plan;
    uint64_t last_pat = ~0ull;
    int last_plan = -1;
    for (int fi = 0; fi < d.n_factors; ++fi) {
      if (fpat[fi] != last_pat) {
        last_pat = fpat[fi];
        auto it = packed2plan.find(last_pat);
        if (it == packed2plan.end()) {
          last_plan = (int)plan_factors.size();
          packed2plan[last_pat] = last_plan;
          plan_factors.emplace_back();
          PatternKey pk;
          pattern_of(fi, pk);
          plan_pat.push_back(pk);
        } else {
          last_plan = it->second;
        }
      }
      plan_factors[last_plan].push_back(fref[fi]);
    }
  }

  clk.lap("batch plans");
  // ---- multi-GPU: landmark ranges and factor ownership -----------------------------
  // Landmarks are split into `world` contiguous ranges balanced on the Schur work k(k+1)/2; a
  // factor belongs to the rank of its landmark (factors without landmark: rank 0).  Structure
  // (blocks, S pattern, front plan) is built from ALL factors so that it is identical on every
  // rank; only the slots, landmarks and matches of this rank are kept for the device.
  a.rank = d.comm ? d.rank : 0;
  a.world = d.comm ? d.world : 1;
  SFX_CHECK(a.world >= 1 && a.rank >= 0 && a.rank < a.world, SFX_ERR_INVALID_ARG, "rank/world");
  SFX_CHECK(a.world == 1 || a.schur, SFX_ERR_UNSUPPORTED,
            "multi-GPU sharding needs the Schur solver (pose-graph problems run as replicas)");
  std::vector<int> lm_rank_begin(a.world + 1, nk);
  if (a.schur) {
    double total_cost = 0;
    for (int k = first_lm_key; k < nk; ++k) total_cost += 0.5 * cnt[k] * (cnt[k] + 1.0) + cnt[k];
    double acc_cost = 0;
    int r = 0;
    lm_rank_begin[0] = first_lm_key;
    for (int k = first_lm_key; k < nk; ++k) {
      while (r + 1 < a.world && acc_cost >= total_cost * (r + 1) / a.world) lm_rank_begin[++r] = k;
      acc_cost += 0.5 * cnt[k] * (cnt[k] + 1.0) + cnt[k];
    }
    while (r + 1 < a.world) lm_rank_begin[++r] = nk;
    lm_rank_begin[a.world] = nk;
  }
  a.lm_rank_begin = lm_rank_begin;
  auto owner_of = [&](const FactorRef& fr) -> int {
    if (a.world == 1) return 0;
    const sfx_factor_batch& fb = d.batches[fr.batch];
    const sfx_kind_meta& km = SFX_KIND_META[fb.kind];
    for (int o = 0; o < km.n_opt; ++o) {
      int key = fb.opt_keys[(int64_t)o * fb.n + fr.idx];
      if (key >= first_lm_key) {
        int r = (int)(std::upper_bound(lm_rank_begin.begin(), lm_rank_begin.end(), key) - lm_rank_begin.begin()) - 1;
        return std::min(std::max(r, 0), a.world - 1);
      }
    }
    return 0;
  };
  std::vector<std::vector<int>> plan_rank_begin(plan_factors.size(), std::vector<int>(a.world + 1, 0));
  if (a.world > 1)
    for (size_t pl = 0; pl < plan_factors.size(); ++pl) {
      auto& v = plan_factors[pl];
      RawBuf<int> own(v.size());
      parallel_chunks((int64_t)v.size(), [&](int64_t i0, int64_t i1) {
        for (int64_t i = i0; i < i1; ++i) own[i] = owner_of(v[i]);
      });
      // stable counting sort of the slots by owner rank
      for (size_t i = 0; i < v.size(); ++i) plan_rank_begin[pl][own[i] + 1]++;
      std::vector<size_t> cursor(a.world, 0);
      for (int r = 1; r < a.world; ++r) cursor[r] = cursor[r - 1] + plan_rank_begin[pl][r];
      std::vector<FactorRef> sorted(v.size());
      for (size_t i = 0; i < v.size(); ++i) sorted[cursor[own[i]]++] = v[i];
      for (int r = 0; r < a.world; ++r) plan_rank_begin[pl][r + 1] += plan_rank_begin[pl][r];
      v.swap(sorted);
    }

  clk.lap("ownership");
  // ---- Hessian blocks ---------------------------------------------------------------------------
  // contributions to off-diagonal blocks: (col node, row node) keyed, with (plan, slot, pair)
  struct Contrib {
    uint64_t key;
    int plan, slot, pair;
    uint32_t transposed;
  };
  size_t n_contribs = 0;
  for (size_t pl = 0; pl < plan_factors.size(); ++pl) {
    int ng = 0;
    for (int o = 0; o < SFX_MAX_OPT; ++o) ng = std::max(ng, plan_pat[pl].grp[o] + 1);
    n_contribs += plan_factors[pl].size() * (size_t)(ng * (ng - 1) / 2);
  }
  RawBuf<Contrib> contribs(n_contribs);  // (plan, slot, pair) order, filled by the chunks below
  size_t c_next = 0;
  a.batches.assign(plan_factors.size(), BatchPlan{});
  for (size_t pl = 0; pl < plan_factors.size(); ++pl) {
    BatchPlan& bp = a.batches[pl];
    const PatternKey& pk = plan_pat[pl];
    const sfx_kind_meta& km = SFX_KIND_META[pk.kind];
    bp.kind = pk.kind;
    bp.n = (int)plan_factors[pl].size();
    bp.n_opt = km.n_opt;
    bp.n_used_args = 0;
    for (int ar = 0; ar < km.n_args; ++ar)
      if (km.arg_used[ar]) bp.used_args[bp.n_used_args++] = ar;
    bp.n_groups = 0;
    for (int o = 0; o < km.n_opt; ++o) {
      bp.key_group[o] = pk.grp[o];
      bp.key_sub[o] = pk.sub[o];
      if (pk.grp[o] >= 0) bp.n_groups = std::max(bp.n_groups, pk.grp[o] + 1);
    }
    const int n = bp.n, ng = bp.n_groups;
    const int npairs = ng * (ng - 1) / 2;
    bp.arg_off.resize((size_t)bp.n_used_args * n);
    bp.res_off.resize(n);
    bp.rhs_off.resize((size_t)ng * n);
    bp.diag_off.resize((size_t)ng * n);
    bp.off_off.resize((size_t)npairs * n);
    bp.factor_index.resize(n);
    // group dims from the first slot (checked against every slot below)
    if (n > 0) {
      const FactorRef& fr = plan_factors[pl][0];
      const sfx_factor_batch& fb = d.batches[fr.batch];
      for (int o = 0; o < km.n_opt; ++o)
        if (pk.grp[o] >= 0) bp.group_dim[pk.grp[o]] = a.nodes[a.keys[fb.opt_keys[(int64_t)o * fb.n + fr.idx]].node].dim;
    }
    const size_t c_base = c_next;
    c_next += (size_t)n * npairs;
    SFX_CHECK(c_next <= n_contribs, SFX_ERR_STRUCTURE, "contribution count");
    parallel_chunks(n, [&](int64_t s_begin, int64_t s_end) {
      for (int s = (int)s_begin; s < (int)s_end; ++s) {
        const FactorRef& fr = plan_factors[pl][s];
        const sfx_factor_batch& fb = d.batches[fr.batch];
        const int f = fr.idx;
        const int fi = fb.factor_index[f];
        bp.factor_index[s] = fi;
        bp.res_off[s] = res_off_of_factor[fi];
        for (int u = 0; u < bp.n_used_args; ++u)
          bp.arg_off[(size_t)u * n + s] = fb.arg_offsets[(int64_t)bp.used_args[u] * fb.n + f];
        int gnode[kMaxGroups];
        for (int o = 0; o < km.n_opt; ++o)
          if (pk.grp[o] >= 0) gnode[pk.grp[o]] = a.keys[fb.opt_keys[(int64_t)o * fb.n + f]].node;
        for (int g = 0; g < ng; ++g) {
          SFX_CHECK(bp.group_dim[g] == a.nodes[gnode[g]].dim, SFX_ERR_STRUCTURE, "non-uniform node dims inside a batch");
          bp.rhs_off[(size_t)g * n + s] = a.nodes[gnode[g]].toff;
        }
        int pair = 0;
        for (int g = 1; g < ng; ++g)
          for (int h = 0; h < g; ++h, ++pair) {
            int I = gnode[g], J = gnode[h];
            uint32_t tr = 0;
            if (I < J) {
              std::swap(I, J);
              tr = 1;
            }
            contribs[c_base + (size_t)s * npairs + pair] = Contrib{((uint64_t)J << 32) | (uint32_t)I, (int)pl, s, pair, tr};
          }
      }
    });
  }
  clk.lap("  contribs built");
  // sort contributions by block; stable so that slot order is kept inside a block
  std::vector<uint32_t> order;
  {
    RawBuf<uint64_t> ck(contribs.size());
    parallel_chunks((int64_t)contribs.size(), [&](int64_t c0, int64_t c1) {
      for (int64_t i = c0; i < c1; ++i) ck[i] = contribs[i].key;
    });
    stable_order_by_key(ck.data(), ck.size(), order);
  }
  clk.lap("  contribs sorted");
  // unique blocks with contributor counts
  std::vector<uint64_t> blk_key;
  std::vector<int> blk_cnt;
  RawBuf<int> contrib_blk(contribs.size());
  for (size_t i = 0; i < order.size(); ++i) {
    const Contrib& c = contribs[order[i]];
    if (blk_key.empty() || blk_key.back() != c.key) {
      blk_key.push_back(c.key);
      blk_cnt.push_back(0);
    }
    blk_cnt.back()++;
    contrib_blk[order[i]] = (int)blk_key.size() - 1;
  }
  clk.lap("  unique blocks");
  // Block matrix structure: per column, diagonal first then sorted off-diagonal rows
  BlockMatrix& H = a.H;
  H.n_nodes = nn;
  H.node_dim.resize(nn);
  H.node_off.resize(nn + 1);
  for (int i = 0; i < nn; ++i) {
    H.node_dim[i] = a.nodes[i].dim;
    H.node_off[i] = a.nodes[i].toff;
  }
  H.node_off[nn] = a.N;
  H.col_ptr.assign(nn + 1, 0);
  for (uint64_t k : blk_key) H.col_ptr[(k >> 32) + 1]++;
  for (int j = 0; j < nn; ++j) H.col_ptr[j + 1] += H.col_ptr[j] + 1;  // +1: diagonal block
  const int nblk = H.col_ptr[nn];
  H.row_idx.resize(nblk);
  H.blk_off.assign(nblk, -1);
  std::vector<int> offdiag_id(blk_key.size());
  {
    std::vector<int> fill(nn);
    for (int j = 0; j < nn; ++j) {
      H.row_idx[H.col_ptr[j]] = j;
      fill[j] = H.col_ptr[j] + 1;
    }
    for (size_t b = 0; b < blk_key.size(); ++b) {  // blk_key is sorted by (col,row)
      int col = (int)(blk_key[b] >> 32), row = (int)(blk_key[b] & 0xffffffffu);
      offdiag_id[b] = fill[col];
      H.row_idx[fill[col]++] = row;
    }
  }
  if (a.schur) {
    // C must be block diagonal: no block between two landmark nodes
    for (size_t b = 0; b < blk_key.size(); ++b) {
      int col = (int)(blk_key[b] >> 32);
      SFX_CHECK(col < first_lm_node, SFX_ERR_STRUCTURE,
                "Submatrix C of A is not block diagonal, cannot use a Schur complement solver");
    }
  }
  clk.lap("  block matrix");
  // value offsets: [reduced diag | all reduced-reduced off-diag] (= B, summed across ranks in the
  // multi-GPU path) | landmark diag | shared landmark-reduced blocks || exclusive blocks in
  // (plan, slot, pair) order so that a thread's stores are contiguous.  Everything before `||`
  // is zeroed before each linearization.
  int64_t off = 0;
  for (int j = 0; j < first_lm_node; ++j) {
    H.blk_off[H.col_ptr[j]] = off;
    off += (int64_t)a.nodes[j].dim * a.nodes[j].dim;
  }
  for (size_t b = 0; b < blk_key.size(); ++b) {
    int col = (int)(blk_key[b] >> 32), row = (int)(blk_key[b] & 0xffffffffu);
    if (row < first_lm_node) {
      H.blk_off[offdiag_id[b]] = off;
      off += (int64_t)a.nodes[row].dim * a.nodes[col].dim;
    }
  }
  a.b_values = off;
  for (int j = first_lm_node; j < nn; ++j) {
    H.blk_off[H.col_ptr[j]] = off;
    off += (int64_t)a.nodes[j].dim * a.nodes[j].dim;
  }
  for (size_t b = 0; b < blk_key.size(); ++b) {
    int col = (int)(blk_key[b] >> 32), row = (int)(blk_key[b] & 0xffffffffu);
    if (row >= first_lm_node && blk_cnt[b] > 1) {
      H.blk_off[offdiag_id[b]] = off;
      off += (int64_t)a.nodes[row].dim * a.nodes[col].dim;
    }
  }
  off = (off + 15) & ~(int64_t)15;  // the exclusive blocks start on a 128-byte boundary (bulk copies stream them)
  a.h_accum_values = off;
  for (size_t c = 0; c < contribs.size(); ++c) {  // contribs are in (plan, slot, pair) order
    int b = contrib_blk[c];
    int col = (int)(blk_key[b] >> 32), row = (int)(blk_key[b] & 0xffffffffu);
    if (row >= first_lm_node && blk_cnt[b] == 1) {
      H.blk_off[offdiag_id[b]] = off;
      off += (int64_t)a.nodes[row].dim * a.nodes[col].dim;
    }
  }
  H.n_values = off;
  SFX_CHECK(off < (int64_t)kOffMask, SFX_ERR_UNSUPPORTED, "Hessian has more than 2^30 block values");
  clk.lap("Hessian blocks");
  // fill per-factor scatter indices
  parallel_chunks((int64_t)contribs.size(), [&](int64_t c0, int64_t c1) {
    for (int64_t c = c0; c < c1; ++c) {
      const Contrib& ct = contribs[c];
      BatchPlan& bp = a.batches[ct.plan];
      int b = contrib_blk[c];
      uint32_t v = (uint32_t)H.blk_off[offdiag_id[b]];
      if (blk_cnt[b] == 1) v |= kOffExclusive;
      if (ct.transposed) v |= kOffTransposed;
      bp.off_off[(size_t)ct.pair * bp.n + ct.slot] = v;
    }
  });
  {
    std::vector<int> toff2node(a.N + 1, -1);  // tangent offset of a node -> node
    for (int i = 0; i < nn; ++i) toff2node[a.nodes[i].toff] = i;
    for (auto& bp : a.batches)
      parallel_chunks((int64_t)bp.rhs_off.size(), [&](int64_t i0, int64_t i1) {
        for (int64_t i = i0; i < i1; ++i) bp.diag_off[i] = (int32_t)H.blk_off[H.col_ptr[toff2node[bp.rhs_off[i]]]];
      });
  }
  // keep only this rank's slots (contiguous thanks to the owner-sorted slot order)
  if (a.world > 1)
    for (size_t pl = 0; pl < a.batches.size(); ++pl) {
      BatchPlan& bp = a.batches[pl];
      const int s0 = plan_rank_begin[pl][a.rank], s1 = plan_rank_begin[pl][a.rank + 1];
      const int n = bp.n, m = s1 - s0;
      auto trim = [&](auto& v, int rows) {
        typename std::remove_reference<decltype(v)>::type w((size_t)rows * m);
        for (int r = 0; r < rows; ++r)
          for (int q = 0; q < m; ++q) w[(size_t)r * m + q] = v[(size_t)r * n + s0 + q];
        v.swap(w);
      };
      trim(bp.arg_off, bp.n_used_args);
      trim(bp.res_off, 1);
      trim(bp.rhs_off, bp.n_groups);
      trim(bp.diag_off, bp.n_groups);
      trim(bp.off_off, bp.n_groups * (bp.n_groups - 1) / 2);
      trim(bp.factor_index, 1);
      bp.n = m;
    }
  a.batches.erase(std::remove_if(a.batches.begin(), a.batches.end(), [](const BatchPlan& b) { return b.n == 0; }),
                  a.batches.end());
  a.diag_pos.resize(a.N);
  for (int i = 0; i < nn; ++i) {
    const int dmn = a.nodes[i].dim;
    for (int r = 0; r < dmn; ++r) a.diag_pos[a.nodes[i].toff + r] = (int32_t)(H.blk_off[H.col_ptr[i]] + r + (int64_t)r * dmn);
  }

  clk.lap("scatter indices");
  // ---- Schur plan -------------------------------------------------------------------------------
  if (a.schur) {
    SchurPlan& sp = a.sp;
    sp.first_lm_node = first_lm_node;
    sp.n_landmarks_total = nn - first_lm_node;
    // landmark nodes of this rank: [lm0, lm1) relative to first_lm_node
    const int lm0 = a.keys[std::min(lm_rank_begin[a.rank], nk - 1)].node - first_lm_node +
                    (lm_rank_begin[a.rank] >= nk ? 1 : 0);
    const int lm1 = lm_rank_begin[a.rank + 1] >= nk ? sp.n_landmarks_total
                                                    : a.keys[lm_rank_begin[a.rank + 1]].node - first_lm_node;
    sp.lm_begin = lm0;
    sp.n_landmarks = std::max(0, lm1 - lm0);
    sp.reduced_dim = a.nodes[first_lm_node].toff;
    const int nl = sp.n_landmarks, nr = first_lm_node, nlt = sp.n_landmarks_total;
    sp.lm_dim.resize(nl);
    sp.lm_cdiag_off.resize(nl);
    sp.lm_toff.resize(nl);
    for (int l = 0; l < nl; ++l) {
      int node = first_lm_node + lm0 + l;
      sp.lm_dim[l] = a.nodes[node].dim;
      sp.lm_cdiag_off[l] = (int32_t)H.blk_off[H.col_ptr[node]];
      sp.lm_toff[l] = a.nodes[node].toff;
    }
    // E blocks of ALL landmarks (for the S pattern): blocks (row = landmark node, col = reduced node)
    // Every host thread takes a contiguous range of landmarks and, in each column, the (sorted) rows inside it: the
    // per-landmark lists are written without conflicts and come out ordered by reduced node.
    std::vector<int32_t> all_ptr(nlt + 1, 0), all_off, all_node;
    const int nt_e = host_threads(H.col_ptr[nr], 1 << 18);
    auto lm_rows_of = [&](int t, int j, int& p0, int& p1) {  // blocks of column j whose row is a landmark of thread t
      const int* rb = H.row_idx.data() + H.col_ptr[j] + 1;
      const int* re = H.row_idx.data() + H.col_ptr[j + 1];
      const int lo = first_lm_node + (int)((int64_t)nlt * t / nt_e), hi = first_lm_node + (int)((int64_t)nlt * (t + 1) / nt_e);
      p0 = (int)(std::lower_bound(rb, re, lo) - H.row_idx.data());
      p1 = (int)(std::lower_bound(rb, re, hi) - H.row_idx.data());
    };
    run_threads(nt_e, [&](int t) {
      for (int j = 0; j < nr; ++j) {
        int p0, p1;
        lm_rows_of(t, j, p0, p1);
        for (int p = p0; p < p1; ++p) all_ptr[H.row_idx[p] - first_lm_node + 1]++;
      }
    });
    for (int l = 0; l < nlt; ++l) all_ptr[l + 1] += all_ptr[l];
    all_off.resize(all_ptr[nlt]);
    all_node.resize(all_ptr[nlt]);
    {
      std::vector<int> fill(all_ptr.begin(), all_ptr.end() - 1);
      run_threads(nt_e, [&](int t) {
        for (int j = 0; j < nr; ++j) {  // increasing j -> each landmark's list sorted by node
          int p0, p1;
          lm_rows_of(t, j, p0, p1);
          for (int p = p0; p < p1; ++p) {
            const int l = H.row_idx[p] - first_lm_node;
            all_off[fill[l]] = (int32_t)H.blk_off[p];
            all_node[fill[l]] = j;
            fill[l]++;
          }
        }
      });
    }
    sp.lm_e_ptr.assign(nl + 1, 0);
    for (int l = 0; l < nl; ++l) sp.lm_e_ptr[l + 1] = sp.lm_e_ptr[l] + (all_ptr[lm0 + l + 1] - all_ptr[lm0 + l]);
    sp.lm_e_off.assign(all_off.begin() + all_ptr[lm0], all_off.begin() + all_ptr[lm0 + nl]);
    sp.lm_e_node.assign(all_node.begin() + all_ptr[lm0], all_node.begin() + all_ptr[lm0 + nl]);
  clk.lap("  E lists");
    // reduced rhs lists (per reduced node: E blocks of its column that belong to own landmarks)
    sp.r_ptr.assign(nr + 1, 0);
    for (int j = 0; j < nr; ++j) {
      int c = 0;
      for (int p = H.col_ptr[j] + 1; p < H.col_ptr[j + 1]; ++p) {
        int l = H.row_idx[p] - first_lm_node - lm0;
        if (H.row_idx[p] >= first_lm_node && l >= 0 && l < nl) c++;
      }
      sp.r_ptr[j + 1] = sp.r_ptr[j] + c;
    }
    sp.r_eoff.resize(sp.r_ptr[nr]);
    sp.r_lm.resize(sp.r_ptr[nr]);
    for (int j = 0; j < nr; ++j) {
      int q = sp.r_ptr[j];
      for (int p = H.col_ptr[j] + 1; p < H.col_ptr[j + 1]; ++p) {
        int l = H.row_idx[p] - first_lm_node - lm0;
        if (H.row_idx[p] >= first_lm_node && l >= 0 && l < nl) {
          sp.r_eoff[q] = (int32_t)H.blk_off[p];
          sp.r_lm[q] = l;
          q++;
        }
      }
    }
  clk.lap("  rhs lists");
    // S pattern: B blocks + all pairs (I >= J) of reduced nodes adjacent to a common landmark (ANY
    // rank's landmark, so the pattern is the same everywhere); matches only for own landmarks.
    // Match order: by S block (column J, then row I), inside a block by landmark.  Built without a global sort:
    // (1) per-thread counts of the matches each column receives, (2) every thread scatters the matches of its
    // landmark range into the column buckets (thread ranges are in landmark order, so a bucket stays landmark-ordered),
    // (3) each column is counting-sorted by row on its own (cache resident; columns are handed out dynamically) and
    // written straight into the final arrays.
    struct ColMatch {
      int32_t row, ei, ej, lm;
    };
    std::vector<uint64_t> skeys;
    if (nl < nlt) {  // other ranks' landmarks only contribute to the pattern
      const int64_t n_words = ((int64_t)nr * nr + 63) / 64;
      if (n_words <= (int64_t)(1 << 23)) {
        // nr x nr bitmap (<= 64 MB) of the (column, row) pairs, set with relaxed atomic ORs from all host threads
        std::vector<uint64_t> bits((size_t)n_words, 0);
        const int nt_p = host_threads(all_ptr[nlt], 1 << 16);
        run_threads(nt_p, [&](int t) {
          for (int l = (int)((int64_t)nlt * t / nt_p); l < (int)((int64_t)nlt * (t + 1) / nt_p); ++l) {
            if (l >= lm0 && l < lm0 + nl) continue;
            for (int q = all_ptr[l]; q < all_ptr[l + 1]; ++q)
              for (int p = q; p < all_ptr[l + 1]; ++p) {
                const int64_t bit = (int64_t)all_node[q] * nr + all_node[p];
                const uint64_t m = 1ull << (bit & 63);
                if (!(__atomic_load_n(&bits[bit >> 6], __ATOMIC_RELAXED) & m))
                  __atomic_fetch_or(&bits[bit >> 6], m, __ATOMIC_RELAXED);
              }
          }
        });
        for (int64_t w = 0; w < n_words; ++w)
          for (uint64_t x = bits[w]; x; x &= x - 1) {
            const int64_t bit = w * 64 + __builtin_ctzll(x);
            skeys.push_back(((uint64_t)(bit / nr) << 32) | (uint32_t)(bit % nr));
          }
      } else {
        for (int l = 0; l < nlt; ++l) {
          if (l >= lm0 && l < lm0 + nl) continue;
          for (int q = all_ptr[l]; q < all_ptr[l + 1]; ++q)
            for (int p = q; p < all_ptr[l + 1]; ++p) skeys.push_back(((uint64_t)all_node[q] << 32) | (uint32_t)all_node[p]);
          if (skeys.size() > (size_t)(1 << 24)) {  // keep the temporary bounded
            std::sort(skeys.begin(), skeys.end());
            skeys.erase(std::unique(skeys.begin(), skeys.end()), skeys.end());
          }
        }
      }
    }
    std::vector<int64_t> m_prefix(nl + 1, 0);  // matches of own landmarks [0, l)
    for (int l = 0; l < nl; ++l) {
      const int64_t k = sp.lm_e_ptr[l + 1] - sp.lm_e_ptr[l];
      m_prefix[l + 1] = m_prefix[l] + k * (k + 1) / 2;
    }
    const int64_t n_matches = m_prefix[nl];
    SFX_CHECK(n_matches < (int64_t)UINT32_MAX, SFX_ERR_UNSUPPORTED, "too many Schur matches");
    const int nt = host_threads(n_matches, 1 << 18);
    std::vector<int> t_begin(nt + 1, nl);  // landmark ranges with equal numbers of matches
    for (int t = 0; t <= nt; ++t)
      t_begin[t] = (int)(std::lower_bound(m_prefix.begin(), m_prefix.end(), n_matches * t / nt) - m_prefix.begin());
    t_begin[0] = 0;
    t_begin[nt] = nl;
    std::vector<std::vector<int64_t>> t_cnt(nt, std::vector<int64_t>(nr, 0));
    run_threads(nt, [&](int t) {
      auto& c = t_cnt[t];
      for (int l = t_begin[t]; l < t_begin[t + 1]; ++l) {
        const int e0 = all_ptr[lm0 + l], e1 = all_ptr[lm0 + l + 1];
        for (int q = e0; q < e1; ++q) c[all_node[q]] += e1 - q;
      }
    });
    std::vector<int64_t> col_start(nr + 1, 0);
    for (int j = 0; j < nr; ++j) {
      int64_t run = col_start[j];
      for (int t = 0; t < nt; ++t) {
        const int64_t c = t_cnt[t][j];
        t_cnt[t][j] = run;  // becomes this thread's write cursor in column j
        run += c;
      }
      col_start[j + 1] = run;
    }
    RawBuf<ColMatch> bucket((size_t)n_matches);
    run_threads(nt, [&](int t) {
      auto& cur = t_cnt[t];
      for (int l = t_begin[t]; l < t_begin[t + 1]; ++l) {
        const int e0 = all_ptr[lm0 + l], e1 = all_ptr[lm0 + l + 1];
        for (int q = e0; q < e1; ++q) {
          ColMatch* dst = bucket.data() + cur[all_node[q]];
          for (int p = q; p < e1; ++p) *dst++ = ColMatch{all_node[p], all_off[p], all_off[q], l};  // node[p] >= node[q]
          cur[all_node[q]] += e1 - q;
        }
      }
    });
    clk.lap("  matches bucketed");
    sp.m_eoff_i.resize((size_t)n_matches);
    sp.m_eoff_j.resize((size_t)n_matches);
    sp.m_lm.resize((size_t)n_matches);
    // per column: the S blocks that receive matches (row, count), in row order
    std::vector<std::vector<std::pair<int32_t, int32_t>>> col_runs(nr);
    {
      std::atomic<int> next_col{0};
      run_threads(nt, [&](int) {
        std::vector<int32_t> start(nr + 1);
        for (;;) {
          const int j = next_col.fetch_add(1, std::memory_order_relaxed);
          if (j >= nr) break;
          const int64_t b0 = col_start[j], b1 = col_start[j + 1];
          if (b0 == b1) continue;
          SFX_CHECK(b1 - b0 < (int64_t)INT32_MAX, SFX_ERR_UNSUPPORTED, "too many Schur matches in one column");
          std::fill(start.begin(), start.end(), 0);
          for (int64_t i = b0; i < b1; ++i) start[bucket[i].row + 1]++;
          auto& runs = col_runs[j];
          for (int r = 0; r < nr; ++r) {
            if (start[r + 1]) runs.emplace_back(r, start[r + 1]);
            start[r + 1] += start[r];
          }
          for (int64_t i = b0; i < b1; ++i) {  // stable: landmark order is kept inside a block
            const ColMatch& m = bucket[i];
            const int64_t o = b0 + start[m.row]++;
            sp.m_eoff_i[o] = m.ei;
            sp.m_eoff_j[o] = m.ej;
            sp.m_lm[o] = m.lm;
          }
        }
      });
    }
    bucket.reset();
    clk.lap("  matches sorted");
    // merged column structure
    for (int j = 0; j < nr; ++j)
      for (int p = H.col_ptr[j]; p < H.col_ptr[j + 1]; ++p)
        if (H.row_idx[p] < first_lm_node) skeys.push_back(((uint64_t)j << 32) | (uint32_t)H.row_idx[p]);
    for (int j = 0; j < nr; ++j)
      for (const auto& rc : col_runs[j]) skeys.push_back(((uint64_t)j << 32) | (uint32_t)rc.first);
    std::sort(skeys.begin(), skeys.end());
    skeys.erase(std::unique(skeys.begin(), skeys.end()), skeys.end());
    clk.lap("  skeys");
    BlockMatrix& S = sp.S;
    S.n_nodes = nr;
    S.node_dim.assign(H.node_dim.begin(), H.node_dim.begin() + nr);
    S.node_off.assign(H.node_off.begin(), H.node_off.begin() + nr + 1);
    S.col_ptr.assign(nr + 1, 0);
    for (uint64_t k : skeys) S.col_ptr[(k >> 32) + 1]++;
    for (int j = 0; j < nr; ++j) S.col_ptr[j + 1] += S.col_ptr[j];
    S.row_idx.resize(skeys.size());
    S.blk_off.resize(skeys.size());
    int64_t soff = 0;
    for (size_t b = 0; b < skeys.size(); ++b) {
      int col = (int)(skeys[b] >> 32), row = (int)(skeys[b] & 0xffffffffu);
      S.row_idx[b] = row;
      S.blk_off[b] = soff;
      soff += (int64_t)S.node_dim[row] * S.node_dim[col];
    }
    S.n_values = soff;
    sp.s_b_src.assign(skeys.size(), -1);
    for (int j = 0; j < nr; ++j)
      for (int p = H.col_ptr[j]; p < H.col_ptr[j + 1]; ++p)
        if (H.row_idx[p] < first_lm_node) sp.s_b_src[S.find(H.row_idx[p], j)] = (int32_t)H.blk_off[p];
    clk.lap("  S structure");
    sp.s_m_ptr.assign(skeys.size() + 1, 0);
    {
      size_t b = 0;
      for (int j = 0; j < nr; ++j)
        for (const auto& rc : col_runs[j]) {
          const uint64_t key = ((uint64_t)j << 32) | (uint32_t)rc.first;
          while (skeys[b] != key) ++b;
          sp.s_m_ptr[b + 1] = rc.second;
        }
      for (size_t q = 0; q < skeys.size(); ++q) sp.s_m_ptr[q + 1] += sp.s_m_ptr[q];
    }
  }
  clk.lap("Schur plan");
}

// Reference CSC layout of Linearization::hessian_lower (key order, lower incl. explicit diagonal)
// and, per entry, where it lives in the block storage.
void build_csc(Analysis& a) {
  if (a.csc_built) return;
  const int nk = a.n_keys;
  const BlockMatrix& H = a.H;
  // key-level off-diagonal adjacency: keys ki > kj whose nodes share a block (or the same node)
  // Per node column J: row nodes I (incl. J).  Keys of node: contiguous? not necessarily in key
  // order, so go through a per-key list.
  std::vector<std::vector<int>> node_keys(a.nodes.size());
  for (int k = 0; k < nk; ++k) node_keys[a.keys[k].node].push_back(k);
  a.csc_outer.assign(a.N + 1, 0);
  // count per reference column
  std::vector<std::vector<int>> rows_of_key(nk);  // off-diagonal row keys (ref order > key)
  for (int J = 0; J < H.n_nodes; ++J)
    for (int p = H.col_ptr[J]; p < H.col_ptr[J + 1]; ++p) {
      int I = H.row_idx[p];
      for (int kj : node_keys[J])
        for (int ki : node_keys[I]) {
          if (ki == kj) continue;
          if (I == J && ki < kj) continue;  // handled when the roles are swapped
          int lo = std::min(ki, kj), hi = std::max(ki, kj);
          rows_of_key[lo].push_back(hi);
        }
    }
  int64_t nnz = 0;
  for (int k = 0; k < nk; ++k) {
    auto& v = rows_of_key[k];
    std::sort(v.begin(), v.end());
    v.erase(std::unique(v.begin(), v.end()), v.end());
    int offd = 0;
    for (int r : v) offd += a.keys[r].tdim;
    for (int c = 0; c < a.keys[k].tdim; ++c) {
      int64_t n = (a.keys[k].tdim - c) + offd;
      nnz += n;
      SFX_CHECK(nnz < (int64_t)INT32_MAX, SFX_ERR_UNSUPPORTED,
                "hessian_lower has >= 2^31 nonzeros (reference limit, linearizer.cc:317-323)");
      a.csc_outer[a.keys[k].ref_toff + c + 1] = (int32_t)nnz;
    }
  }
  a.nnz = nnz;
  a.csc_inner.resize(nnz);
  a.csc_src.resize(nnz);
  auto value_pos = [&](int krow, int r, int kcol, int c) -> int32_t {
    // H value offset of entry (key krow row r, key kcol col c), reference-lower (krow >= kcol)
    int I = a.keys[krow].node, J = a.keys[kcol].node;
    int rr = a.keys[krow].sub + r, cc = a.keys[kcol].sub + c;
    if (I < J || (I == J && rr < cc)) {
      std::swap(I, J);
      std::swap(rr, cc);
    }
    int b = H.find(I, J);
    return (int32_t)(H.blk_off[b] + rr + (int64_t)cc * H.node_dim[I]);
  };
  for (int k = 0; k < nk; ++k) {
    const int dk = a.keys[k].tdim;
    for (int c = 0; c < dk; ++c) {
      int64_t p = a.csc_outer[a.keys[k].ref_toff + c];
      for (int r = c; r < dk; ++r) {
        a.csc_inner[p] = a.keys[k].ref_toff + r;
        a.csc_src[p] = value_pos(k, r, k, c);
        ++p;
      }
      for (int rk : rows_of_key[k])
        for (int r = 0; r < a.keys[rk].tdim; ++r) {
          a.csc_inner[p] = a.keys[rk].ref_toff + r;
          a.csc_src[p] = value_pos(rk, r, k, c);
          ++p;
        }
    }
  }
  a.csc_built = true;
}

// Reference layout of Linearization::jacobian (include_jacobians; linearizer.cc:252-259, 297-313): M x N CSC, one
// entry per (residual row of a factor, tangent column of one of its optimized keys), rows ascending inside a column,
// i.e. factors in residual order.  All columns of a key hold the same rows, so a factor's R x dim(key) block sits at
// base + c * colnnz(key) + r: `jac_base` / `jac_colnnz` per (optimized arg, slot) are what jacobian_kernel scatters with.
void build_jacobian_csc(Analysis& a) {
  if (a.jac_built) return;
  SFX_CHECK(a.world == 1, SFX_ERR_UNSUPPORTED, "the Jacobian export is single-GPU");
  std::vector<int> int2ref(a.N);
  for (int r = 0; r < a.N; ++r) int2ref[a.ref2int[r]] = r;
  std::vector<int> key_of_toff(a.N, -1);  // reference tangent offset of a key's first scalar -> key
  for (int k = 0; k < a.n_keys; ++k) key_of_toff[a.keys[k].ref_toff] = k;
  struct Ent {
    int key, res_off, res_dim, batch, opt, slot;
  };
  std::vector<Ent> ents;
  std::vector<uint64_t> ekey;
  for (size_t b = 0; b < a.batches.size(); ++b) {
    BatchPlan& bp = a.batches[b];
    const sfx_kind_meta& km = SFX_KIND_META[bp.kind];
    bp.jac_base.assign((size_t)bp.n_opt * bp.n, -1);
    bp.jac_colnnz.assign((size_t)bp.n_opt * bp.n, 0);
    for (int o = 0; o < bp.n_opt; ++o) {
      const int g = bp.key_group[o];
      if (g < 0) continue;
      for (int s = 0; s < bp.n; ++s) {
        const int key = key_of_toff[int2ref[bp.rhs_off[(size_t)g * bp.n + s] + bp.key_sub[o]]];
        SFX_CHECK(key >= 0 && a.keys[key].tdim == km.opt_dims[o], SFX_ERR_STRUCTURE, "Jacobian index: key lookup");
        ents.push_back(Ent{key, bp.res_off[s], km.res_dim, (int)b, o, s});
        ekey.push_back(((uint64_t)key << 32) | (uint32_t)bp.res_off[s]);
      }
    }
  }
  std::vector<uint32_t> order;
  stable_order_by_key(ekey.data(), ekey.size(), order);  // by (key, residual offset)
  a.jac_outer.assign(a.N + 1, 0);
  int64_t nnz = 0;
  size_t i = 0;
  std::vector<int64_t> key_begin(a.n_keys + 1, 0);  // CSC position of column 0 of each key
  std::vector<int> key_colnnz(a.n_keys, 0);
  for (int k = 0; k < a.n_keys; ++k) {  // keys are in reference (state vector) order
    key_begin[k] = nnz;
    int rows = 0;
    while (i < order.size() && ents[order[i]].key == k) {
      const Ent& e = ents[order[i]];
      BatchPlan& bp = a.batches[e.batch];
      SFX_CHECK(nnz + rows < (int64_t)INT32_MAX, SFX_ERR_UNSUPPORTED,
                "jacobian has >= 2^31 nonzeros (reference limit, linearizer.cc:303-311)");
      bp.jac_base[(size_t)e.opt * bp.n + e.slot] = (int32_t)(nnz + rows);
      rows += e.res_dim;
      ++i;
    }
    key_colnnz[k] = rows;
    for (int c = 0; c < a.keys[k].tdim; ++c) {
      nnz += rows;
      SFX_CHECK(nnz < (int64_t)INT32_MAX, SFX_ERR_UNSUPPORTED,
                "jacobian has >= 2^31 nonzeros (reference limit, linearizer.cc:303-311)");
      a.jac_outer[a.keys[k].ref_toff + c + 1] = (int32_t)nnz;
    }
  }
  key_begin[a.n_keys] = nnz;
  a.jac_nnz = nnz;
  a.jac_inner.resize(nnz);
  for (size_t q = 0; q < order.size(); ++q) {
    const Ent& e = ents[order[q]];
    BatchPlan& bp = a.batches[e.batch];
    const int cn = key_colnnz[e.key];
    bp.jac_colnnz[(size_t)e.opt * bp.n + e.slot] = cn;
    const int64_t base = bp.jac_base[(size_t)e.opt * bp.n + e.slot];
    for (int c = 0; c < a.keys[e.key].tdim; ++c)
      for (int r = 0; r < e.res_dim; ++r) a.jac_inner[base + (int64_t)c * cn + r] = e.res_off + r;
  }
  a.jac_built = true;
}

// BAL fast path: point -> observation slots of a Snavely batch (points identified by the value offset of their
// diagonal block, group 1): `order` lists the slots grouped by point in ascending block offset, slot order kept inside
// a point; ptr / diag / rhs describe the groups.  Consumed by bal_point_finalize_kernel.
void build_point_lists(const BatchPlan& bp, std::vector<int32_t>& order, std::vector<int32_t>& ptr,
                       std::vector<int32_t>& diag, std::vector<int32_t>& rhs) {
  const int n = bp.n;
  const int32_t* pd = bp.diag_off.data() + n;
  const int32_t* pr = bp.rhs_off.data() + n;
  {
    RawBuf<uint64_t> key(n);  // stable radix order instead of a comparison sort of 5 M slots
    parallel_chunks(n, [&](int64_t i0, int64_t i1) {
      for (int64_t i = i0; i < i1; ++i) key[i] = (uint64_t)(uint32_t)pd[i];
    });
    std::vector<uint32_t> o;
    stable_order_by_key(key.data(), (size_t)n, o);
    order.assign(o.begin(), o.end());
  }
  ptr.clear();
  diag.clear();
  rhs.clear();
  for (int i = 0; i < n; ++i)
    if (i == 0 || pd[order[i]] != pd[order[i - 1]]) {
      ptr.push_back(i);
      diag.push_back(pd[order[i]]);
      rhs.push_back(pr[order[i]]);
    }
  ptr.push_back(n);
}

}  // namespace sfx
