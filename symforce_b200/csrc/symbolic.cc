// Symbolic analysis for the multifrontal supernodal Cholesky (replaces
// SparseCholeskySolver::ComputePermutationMatrix / ComputeSymbolicSparsity,
// symforce/opt/sparse_cholesky/sparse_cholesky_solver.tcc:13-106).  Host, once per problem.
//
// Ordering: METIS_NodeND on the scalar pattern in the reference's scalar numbering -- the same
// call on the same graph Eigen::MetisOrdering makes for the reference -- coarsened to nodes
// (a node's scalars have identical closed adjacency, so METIS' own compression keeps them
// together); or METIS on the node quotient graph; or natural order.  Then elimination tree,
// postorder, node-level symbolic factorization, fundamental + relaxed supernodes, frontal
// index lists, extend-add maps and the level schedule the GPU kernels run.
#include <algorithm>
#include <numeric>

#include "sfx_internal.h"

extern "C" int METIS_NodeND(int64_t* nvtxs, int64_t* xadj, int64_t* adjncy, int64_t* vwgt, int64_t* options,
                            int64_t* perm, int64_t* iperm);

namespace sfx {

static void metis(std::vector<int64_t>& xadj, std::vector<int64_t>& adj, std::vector<int64_t>& iperm) {
  int64_t n = (int64_t)xadj.size() - 1;
  iperm.resize(n);
  std::vector<int64_t> perm(n);
  if (xadj[n] == 0) {
    std::iota(iperm.begin(), iperm.end(), 0);
    return;
  }
  int rc = METIS_NodeND(&n, xadj.data(), adj.data(), nullptr, nullptr, perm.data(), iperm.data());
  SFX_CHECK(rc == 1, SFX_ERR_STRUCTURE, "METIS_NodeND failed");
}

extern "C" int METIS_ComputeVertexSeparator(int64_t* nvtxs, int64_t* xadj, int64_t* adjncy, int64_t* vwgt,
                                            int64_t* options, int64_t* sepsize, int64_t* part);

// ---- "dissect, then sweep" ordering -------------------------------------------------------------------------------
// Nested dissection only as deep as the device needs independent subtrees (2^depth of them), then every subdomain is
// eliminated in one sweep (Cuthill-McKee from the vertex farthest from the subdomain's separators).  On banded systems
// (camera chains, trajectories; the synthetic BAL ring is one) a swept subdomain carries the band and ONE border in
// its fronts until the sweep reaches the other end, where full nested dissection carries both borders everywhere:
// about half the factorization flops of the subdomains, and the sweep's column-by-column dependency is exactly what
// the tile-DAG kernel pipelines.  Which depth (or plain METIS_NodeND) to use is decided by the caller from the
// modelled factorization time.
namespace {
struct Dissector {
  const std::vector<std::vector<int>>& adj;  // symmetric node adjacency (no self loops)
  const std::vector<int>& dim;
  std::vector<int> loc;        // scratch: vertex -> local index (-1 outside the current subgraph)
  std::vector<char> is_border; // vertices adjacent to an enclosing separator
  std::vector<int> out;        // elimination order

  void induced(const std::vector<int>& vs, std::vector<int64_t>& xadj, std::vector<int64_t>& ad) {
    for (size_t i = 0; i < vs.size(); ++i) loc[vs[i]] = (int)i;
    xadj.assign(vs.size() + 1, 0);
    ad.clear();
    for (size_t i = 0; i < vs.size(); ++i) {
      for (int u : adj[vs[i]])
        if (loc[u] >= 0) ad.push_back(loc[u]);
      xadj[i + 1] = (int64_t)ad.size();
    }
    for (int v : vs) loc[v] = -1;
  }
  // Cuthill-McKee sweep of every component; starts at the vertex farthest from the border vertices of the component
  // (a pseudo-peripheral vertex when there are none), so the border joins the fronts last.
  void sweep(const std::vector<int>& vs) {
    std::vector<int64_t> xadj, ad;
    induced(vs, xadj, ad);
    const int n = (int)vs.size();
    std::vector<int> dist(n), comp, q;
    std::vector<char> done(n, 0);
    auto bfs = [&](const std::vector<int>& src, const std::vector<int>& within) {
      for (int v : within) dist[v] = -1;
      q.clear();
      for (int v : src) {
        dist[v] = 0;
        q.push_back(v);
      }
      for (size_t h = 0; h < q.size(); ++h)
        for (int64_t e = xadj[q[h]]; e < xadj[q[h] + 1]; ++e)
          if (dist[ad[e]] < 0) {
            dist[ad[e]] = dist[q[h]] + 1;
            q.push_back((int)ad[e]);
          }
      return q.back();
    };
    for (int s0 = 0; s0 < n; ++s0) {
      if (done[s0]) continue;
      // component of s0
      std::fill(dist.begin(), dist.end(), -1);
      bfs({s0}, {});
      comp = q;
      std::vector<int> border;
      for (int v : comp)
        if (is_border[vs[v]]) border.push_back(v);
      int start;
      if (!border.empty()) {
        start = bfs(border, comp);
      } else {
        start = bfs({s0}, comp);
        start = bfs({start}, comp);
      }
      // Cuthill-McKee from `start`: BFS, neighbours by increasing degree
      for (int v : comp) dist[v] = -1;
      std::vector<int> ord{start};
      dist[start] = 0;
      std::vector<int> nb;
      for (size_t h = 0; h < ord.size(); ++h) {
        nb.clear();
        for (int64_t e = xadj[ord[h]]; e < xadj[ord[h] + 1]; ++e)
          if (dist[ad[e]] < 0) {
            dist[ad[e]] = 0;
            nb.push_back((int)ad[e]);
          }
        std::sort(nb.begin(), nb.end(), [&](int a, int b) {
          const int64_t da = xadj[a + 1] - xadj[a], db = xadj[b + 1] - xadj[b];
          return da != db ? da < db : a < b;
        });
        for (int u : nb) ord.push_back(u);
      }
      for (int v : ord) {
        done[v] = 1;
        out.push_back(vs[v]);
      }
    }
  }
  void dissect(const std::vector<int>& vs, int depth) {
    if (vs.empty()) return;
    if (depth <= 0 || vs.size() < 8) {
      sweep(vs);
      return;
    }
    std::vector<int64_t> xadj, ad;
    induced(vs, xadj, ad);
    int64_t n = (int64_t)vs.size(), sep = 0;
    if (ad.empty()) {
      for (int v : vs) out.push_back(v);
      return;
    }
    std::vector<int64_t> vw(n), part(n);
    for (int64_t i = 0; i < n; ++i) vw[i] = dim[vs[i]];
    const int rc = METIS_ComputeVertexSeparator(&n, xadj.data(), ad.data(), vw.data(), nullptr, &sep, part.data());
    SFX_CHECK(rc == 1, SFX_ERR_STRUCTURE, "METIS_ComputeVertexSeparator failed");
    std::vector<int> part_of[3];
    for (int64_t i = 0; i < n; ++i) part_of[part[i] < 0 || part[i] > 2 ? 2 : part[i]].push_back(vs[i]);
    if (part_of[0].empty() || part_of[1].empty()) {
      sweep(vs);
      return;
    }
    // vertices next to the new separator become border vertices of their side
    std::vector<int> marked;
    for (int v : part_of[2])
      for (int u : adj[v])
        if (!is_border[u]) {
          is_border[u] = 1;
          marked.push_back(u);
        }
    dissect(part_of[0], depth - 1);
    dissect(part_of[1], depth - 1);
    for (int u : marked) is_border[u] = 0;
    sweep(part_of[2]);
  }
};
}  // namespace

void dissect_then_sweep(const std::vector<std::vector<int>>& adj, const std::vector<int>& dim, int depth,
                        std::vector<int>& order) {
  const int nn = (int)adj.size();
  Dissector d{adj, dim, std::vector<int>(nn, -1), std::vector<char>(nn, 0), {}};
  d.out.reserve(nn);
  std::vector<int> all(nn);
  std::iota(all.begin(), all.end(), 0);
  d.dissect(all, depth);
  SFX_CHECK((int)d.out.size() == nn, SFX_ERR_STRUCTURE, "internal: dissection lost vertices");
  order = d.out;
}

void build_front_plan(const BlockMatrix& A, int ordering, const std::vector<int>& sys2ref, FrontPlan& fp,
                      const PlanOptions& opt_in) {
  PlanOptions opt = opt_in;
  // experiment knobs (override the caller's choice)
  if (getenv("SFX_ND_DEPTH")) opt.nd_depth = atoi(getenv("SFX_ND_DEPTH"));
  if (getenv("SFX_RELAX")) opt.relax = atof(getenv("SFX_RELAX"));
  if (getenv("SFX_RELAX_CUM")) opt.cumulative = atoi(getenv("SFX_RELAX_CUM")) != 0;
  if (getenv("SFX_MAX_MERGE_W")) opt.max_merge_w = atoi(getenv("SFX_MAX_MERGE_W"));
  const int nn = A.n_nodes;
  const int n = A.node_off[nn];
  fp = FrontPlan{};
  fp.n = n;
  // ---- symmetric node adjacency -----------------------------------------------------------------
  std::vector<std::vector<int>> adj(nn);
  for (int j = 0; j < nn; ++j)
    for (int p = A.col_ptr[j]; p < A.col_ptr[j + 1]; ++p) {
      int i = A.row_idx[p];
      if (i == j) continue;
      adj[i].push_back(j);
      adj[j].push_back(i);
    }
  // ---- ordering -> pos_of[node] ------------------------------------------------------------------
  std::vector<int> order(nn);  // elimination position -> node
  std::iota(order.begin(), order.end(), 0);
  const int nd_depth = opt.nd_depth;
  if (nd_depth >= 0 && ordering != SFX_ORDERING_NATURAL) {
    for (int i = 0; i < nn; ++i) {
      std::sort(adj[i].begin(), adj[i].end());
      adj[i].erase(std::unique(adj[i].begin(), adj[i].end()), adj[i].end());
    }
    dissect_then_sweep(adj, A.node_dim, nd_depth, order);
  } else if (ordering == SFX_ORDERING_METIS_BLOCK) {
    std::vector<int64_t> xadj(nn + 1, 0), ad, iperm;
    for (int i = 0; i < nn; ++i) {
      std::sort(adj[i].begin(), adj[i].end());
      xadj[i + 1] = xadj[i] + (int64_t)adj[i].size();
    }
    ad.reserve(xadj[nn]);
    for (int i = 0; i < nn; ++i)
      for (int v : adj[i]) ad.push_back(v);
    metis(xadj, ad, iperm);
    for (int i = 0; i < nn; ++i) order[iperm[i]] = i;
  } else if (ordering == SFX_ORDERING_METIS_SCALAR) {
    // scalar graph in reference numbering
    SFX_CHECK((int)sys2ref.size() == n, SFX_ERR_INVALID_ARG, "scalar ordering needs the reference numbering");
    std::vector<int64_t> deg(n, 0);
    for (int i = 0; i < nn; ++i) {
      int64_t dsum = A.node_dim[i] - 1;
      for (int v : adj[i]) dsum += A.node_dim[v];
      for (int r = 0; r < A.node_dim[i]; ++r) deg[sys2ref[A.node_off[i] + r]] = dsum;
    }
    std::vector<int64_t> xadj(n + 1, 0);
    for (int v = 0; v < n; ++v) xadj[v + 1] = xadj[v] + deg[v];
    std::vector<int64_t> ad(xadj[n]);
    std::vector<int> tmp;
    for (int i = 0; i < nn; ++i) {
      tmp.clear();
      for (int r = 0; r < A.node_dim[i]; ++r) tmp.push_back(sys2ref[A.node_off[i] + r]);
      for (int v : adj[i])
        for (int r = 0; r < A.node_dim[v]; ++r) tmp.push_back(sys2ref[A.node_off[v] + r]);
      std::sort(tmp.begin(), tmp.end());
      for (int r = 0; r < A.node_dim[i]; ++r) {
        int me = sys2ref[A.node_off[i] + r];
        int64_t q = xadj[me];
        for (int v : tmp)
          if (v != me) ad[q++] = v;
      }
    }
    std::vector<int64_t> iperm;
    metis(xadj, ad, iperm);
    std::vector<int64_t> key(nn);
    for (int i = 0; i < nn; ++i) {
      int64_t m = INT64_MAX;
      for (int r = 0; r < A.node_dim[i]; ++r) m = std::min(m, iperm[sys2ref[A.node_off[i] + r]]);
      key[i] = m;
    }
    std::sort(order.begin(), order.end(), [&](int x, int y) { return key[x] < key[y]; });
  }
  std::vector<int> pos(nn);
  for (int p = 0; p < nn; ++p) pos[order[p]] = p;

  // ---- elimination tree (Liu) on positions -------------------------------------------------------
  std::vector<int> parent(nn, -1), anc(nn, -1);
  for (int p = 0; p < nn; ++p) {
    int node = order[p];
    for (int v : adj[node]) {
      int i = pos[v];
      while (i != -1 && i < p) {
        int nxt = anc[i];
        anc[i] = p;
        if (nxt == -1) parent[i] = p;
        i = nxt;
      }
    }
  }
  // ---- postorder ---------------------------------------------------------------------------------
  {
    std::vector<int> head(nn, -1), next(nn, -1);
    for (int p = nn - 1; p >= 0; --p)
      if (parent[p] >= 0) {
        next[p] = head[parent[p]];
        head[parent[p]] = p;
      }
    std::vector<int> post;
    post.reserve(nn);
    std::vector<int> stack;
    for (int r = 0; r < nn; ++r) {
      if (parent[r] != -1) continue;
      stack.push_back(r);
      while (!stack.empty()) {
        int v = stack.back();
        int c = head[v];
        if (c == -1) {
          post.push_back(v);
          stack.pop_back();
        } else {
          head[v] = next[c];
          stack.push_back(c);
        }
      }
    }
    std::vector<int> newpos(nn);
    for (int q = 0; q < nn; ++q) newpos[post[q]] = q;
    std::vector<int> order2(nn), parent2(nn, -1);
    for (int p = 0; p < nn; ++p) {
      order2[newpos[p]] = order[p];
      parent2[newpos[p]] = parent[p] < 0 ? -1 : newpos[parent[p]];
    }
    order.swap(order2);
    parent.swap(parent2);
    for (int p = 0; p < nn; ++p) pos[order[p]] = p;
  }
  // ---- node-level symbolic factorization ---------------------------------------------------------
  std::vector<std::vector<int>> st(nn);  // struct of each column position (rows > p), sorted
  {
    std::vector<std::vector<int>> children(nn);
    for (int p = 0; p < nn; ++p)
      if (parent[p] >= 0) children[parent[p]].push_back(p);
    std::vector<int> mark(nn, -1);
    for (int p = 0; p < nn; ++p) {
      std::vector<int>& s = st[p];
      mark[p] = p;
      for (int v : adj[order[p]]) {
        int i = pos[v];
        if (i > p && mark[i] != p) {
          mark[i] = p;
          s.push_back(i);
        }
      }
      for (int c : children[p])
        for (int i : st[c])
          if (i > p && mark[i] != p) {
            mark[i] = p;
            s.push_back(i);
          }
      std::sort(s.begin(), s.end());
    }
  }
  // ---- supernodes ---------------------------------------------------------------------------------
  std::vector<int> dim_pos(nn);
  for (int p = 0; p < nn; ++p) dim_pos[p] = A.node_dim[order[p]];
  std::vector<int> sn_first, sn_last;  // position ranges
  {
    std::vector<int> n_children(nn, 0);
    for (int p = 0; p < nn; ++p)
      if (parent[p] >= 0) n_children[parent[p]]++;
    int cur_w = 0;
    for (int p = 0; p < nn; ++p) {
      bool merge = p > 0 && parent[p - 1] == p && st[p - 1].size() == st[p].size() + 1;
      // a column with several children starts its own supernode when the run before it is already wide (a quarter of
      // max_merge_w): that run and its sibling subtrees then stay separate fronts the tile-DAG kernel works on side by
      // side, instead of one diagonal chain
      if (merge && opt.max_merge_w > 0 && n_children[p] > 1 && 4 * cur_w >= opt.max_merge_w) merge = false;
      if (merge) {
        sn_last.back() = p;
        cur_w += dim_pos[p];
      } else {
        sn_first.push_back(p);
        sn_last.push_back(p);
        cur_w = dim_pos[p];
      }
    }
  }
  // relaxed amalgamation: merge a supernode into its parent when contiguous and cheap
  {
    int ns = (int)sn_first.size();
    std::vector<int> sn_of(nn);
    for (int s = 0; s < ns; ++s)
      for (int p = sn_first[s]; p <= sn_last[s]; ++p) sn_of[p] = s;
    std::vector<int> alive(ns, 1);
    auto width = [&](int s) {
      int w = 0;
      for (int p = sn_first[s]; p <= sn_last[s]; ++p) w += dim_pos[p];
      return w;
    };
    auto uheight = [&](int s) {
      int u = 0;
      for (int i : st[sn_last[s]]) u += dim_pos[i];
      return u;
    };
    std::vector<int> sw(ns), su(ns);
    std::vector<double> sz(ns, 0.0);  // explicit zeros a supernode has accumulated through merges
    for (int s = 0; s < ns; ++s) {
      sw[s] = width(s);
      su[s] = uheight(s);
    }
    // process parents in increasing order; the candidate child is the supernode ending at first-1
    for (int s = 0; s < ns; ++s) {
      if (!alive[s]) continue;
      while (sn_first[s] > 0) {
        int c = sn_of[sn_first[s] - 1];
        if (parent[sn_last[c]] != sn_first[s]) break;  // not a child of this supernode's first column
        const double wc = sw[c], uc = su[c], wp = sw[s], up = su[s];
        const double mp = wp + up;
        const double zeros = wc * (mp - uc);  // rows of the parent front absent from the child
        const double merged = (wc + wp) * (wc + wp + up);
        const double relax_big = opt.relax;
        const double ztot = opt.cumulative ? zeros + sz[c] + sz[s] : zeros;
        bool ok = zeros <= 0.0 || (wc + wp <= 48 && zeros <= 0.35 * merged) || ztot <= relax_big * merged;
        if (opt.max_merge_w > 0 && wc + wp > opt.max_merge_w) ok = false;
        if (!ok) break;
        // merge c into s
        sz[s] += sz[c] + zeros;
        alive[c] = 0;
        sn_first[s] = sn_first[c];
        for (int p = sn_first[c]; p <= sn_last[c]; ++p) sn_of[p] = s;
        sw[s] += sw[c];
      }
    }
    std::vector<int> f2, l2;
    for (int s = 0; s < ns; ++s)
      if (alive[s]) {
        f2.push_back(sn_first[s]);
        l2.push_back(sn_last[s]);
      }
    sn_first.swap(f2);
    sn_last.swap(l2);
  }
  const int nf = (int)sn_first.size();
  fp.n_fronts = nf;
  std::vector<int> sn_of(nn);
  for (int s = 0; s < nf; ++s)
    for (int p = sn_first[s]; p <= sn_last[s]; ++p) sn_of[p] = s;
  // scalar positions
  std::vector<int> spos(nn + 1, 0);
  for (int p = 0; p < nn; ++p) spos[p + 1] = spos[p] + dim_pos[p];
  fp.perm_nodes = order;
  fp.scalar_perm.resize(n);
  for (int p = 0; p < nn; ++p)
    for (int r = 0; r < dim_pos[p]; ++r) fp.scalar_perm[spos[p] + r] = A.node_off[order[p]] + r;

  fp.f_w.resize(nf);
  fp.f_u.resize(nf);
  fp.f_parent.assign(nf, -1);
  fp.f_level.assign(nf, 0);
  fp.f_off.resize(nf);
  fp.f_piv.resize(nf);
  fp.f_rows_ptr.assign(nf + 1, 0);
  fp.f_toff.resize(nf);
  // a merged supernode's update rows: struct of its last column (merging keeps rows of parent)
  for (int s = 0; s < nf; ++s) {
    fp.f_piv[s] = spos[sn_first[s]];
    fp.f_w[s] = spos[sn_last[s] + 1] - spos[sn_first[s]];
    int u = 0;
    for (int i : st[sn_last[s]]) u += dim_pos[i];
    fp.f_u[s] = u;
    fp.f_rows_ptr[s + 1] = fp.f_rows_ptr[s] + u;
    int pl = parent[sn_last[s]];
    fp.f_parent[s] = pl < 0 ? -1 : sn_of[pl];
  }
  fp.f_rows.resize(fp.f_rows_ptr[nf]);
  for (int s = 0; s < nf; ++s) {
    int q = fp.f_rows_ptr[s];
    for (int i : st[sn_last[s]])
      for (int r = 0; r < dim_pos[i]; ++r) fp.f_rows[q++] = spos[i] + r;
  }
  // levels, children, offsets
  fp.f_child_ptr.assign(nf + 1, 0);
  for (int s = 0; s < nf; ++s)
    if (fp.f_parent[s] >= 0) fp.f_child_ptr[fp.f_parent[s] + 1]++;
  for (int s = 0; s < nf; ++s) fp.f_child_ptr[s + 1] += fp.f_child_ptr[s];
  fp.f_child.resize(fp.f_child_ptr[nf]);
  {
    std::vector<int> fill(fp.f_child_ptr.begin(), fp.f_child_ptr.end() - 1);
    for (int s = 0; s < nf; ++s)
      if (fp.f_parent[s] >= 0) fp.f_child[fill[fp.f_parent[s]]++] = s;
  }
  for (int s = 0; s < nf; ++s)  // children precede parents (postorder)
    if (fp.f_parent[s] >= 0) fp.f_level[fp.f_parent[s]] = std::max(fp.f_level[fp.f_parent[s]], fp.f_level[s] + 1);
  fp.n_levels = 0;
  for (int s = 0; s < nf; ++s) fp.n_levels = std::max(fp.n_levels, fp.f_level[s] + 1);
  fp.level_ptr.assign(fp.n_levels + 1, 0);
  for (int s = 0; s < nf; ++s) fp.level_ptr[fp.f_level[s] + 1]++;
  for (int l = 0; l < fp.n_levels; ++l) fp.level_ptr[l + 1] += fp.level_ptr[l];
  fp.level_fronts.resize(nf);
  {
    std::vector<int> fill(fp.level_ptr.begin(), fp.level_ptr.end() - 1);
    for (int s = 0; s < nf; ++s) fp.level_fronts[fill[fp.f_level[s]]++] = s;
  }
  int64_t off = 0, toff = 0;
  for (int s = 0; s < nf; ++s) {
    const int64_t m = fp.f_w[s] + fp.f_u[s];
    fp.f_off[s] = off;
    off += m * m;
    fp.f_toff[s] = (int)toff;
    toff += fp.f_u[s];
    fp.max_front = std::max(fp.max_front, (int)m);
    const double w = fp.f_w[s], u = fp.f_u[s];
    fp.nnz_L += (int64_t)(w * (w + 1) / 2 + u * w);
    for (int k = 0; k < fp.f_w[s]; ++k) fp.flops += (double)(m - k) * (m - k);
  }
  fp.front_values = off;
  fp.solve_ws = toff;
  // relative indices into the parent front
  fp.f_rel_ptr = fp.f_rows_ptr;
  fp.f_rel.resize(fp.f_rows.size());
  {
    std::vector<int> where(n, -1);
    for (int s = 0; s < nf; ++s) {
      // children of s need `where` of s
      if (fp.f_child_ptr[s] == fp.f_child_ptr[s + 1]) continue;
      for (int r = 0; r < fp.f_w[s]; ++r) where[fp.f_piv[s] + r] = r;
      for (int q = fp.f_rows_ptr[s]; q < fp.f_rows_ptr[s + 1]; ++q) where[fp.f_rows[q]] = fp.f_w[s] + (q - fp.f_rows_ptr[s]);
      for (int ci = fp.f_child_ptr[s]; ci < fp.f_child_ptr[s + 1]; ++ci) {
        int c = fp.f_child[ci];
        for (int q = fp.f_rows_ptr[c]; q < fp.f_rows_ptr[c + 1]; ++q) {
          SFX_CHECK(where[fp.f_rows[q]] >= 0, SFX_ERR_STRUCTURE, "internal: child row missing in parent front");
          fp.f_rel[q] = where[fp.f_rows[q]];
        }
      }
      for (int r = 0; r < fp.f_w[s]; ++r) where[fp.f_piv[s] + r] = -1;
      for (int q = fp.f_rows_ptr[s]; q < fp.f_rows_ptr[s + 1]; ++q) where[fp.f_rows[q]] = -1;
    }
  }
  // assembly copies
  {
    std::vector<std::vector<FrontPlan::Copy>> per_front(nf);
    std::vector<int> where(n, -1);
    // group blocks by owning front: owning column position = min(pos[I], pos[J])
    std::vector<std::vector<int>> blocks_of_front(nf);
    std::vector<int> blk_col(A.row_idx.size());
    for (int j = 0; j < nn; ++j)
      for (int p = A.col_ptr[j]; p < A.col_ptr[j + 1]; ++p) {
        blk_col[p] = j;
        int cp = std::min(pos[A.row_idx[p]], pos[j]);
        blocks_of_front[sn_of[cp]].push_back(p);
      }
    for (int s = 0; s < nf; ++s) {
      for (int r = 0; r < fp.f_w[s]; ++r) where[fp.f_piv[s] + r] = r;
      for (int q = fp.f_rows_ptr[s]; q < fp.f_rows_ptr[s + 1]; ++q) where[fp.f_rows[q]] = fp.f_w[s] + (q - fp.f_rows_ptr[s]);
      for (int p : blocks_of_front[s]) {
        int I = A.row_idx[p], J = blk_col[p];
        int pi = pos[I], pj = pos[J];
        FrontPlan::Copy c;
        c.src = A.blk_off[p];
        c.rows = A.node_dim[I];
        c.cols = A.node_dim[J];
        c.src_ld = A.node_dim[I];
        c.lower_only = (I == J);
        if (pi >= pj) {
          c.transposed = 0;
          c.dst_col = spos[pj] - fp.f_piv[s];
          c.dst_row = where[spos[pi]];
        } else {
          c.transposed = 1;
          c.dst_col = spos[pi] - fp.f_piv[s];
          c.dst_row = where[spos[pj]];
        }
        SFX_CHECK(c.dst_row >= 0, SFX_ERR_STRUCTURE, "internal: block row missing in front");
        per_front[s].push_back(c);
      }
      for (int r = 0; r < fp.f_w[s]; ++r) where[fp.f_piv[s] + r] = -1;
      for (int q = fp.f_rows_ptr[s]; q < fp.f_rows_ptr[s + 1]; ++q) where[fp.f_rows[q]] = -1;
    }
    fp.f_copy_ptr.assign(nf + 1, 0);
    for (int s = 0; s < nf; ++s) fp.f_copy_ptr[s + 1] = fp.f_copy_ptr[s] + (int)per_front[s].size();
    fp.copies.reserve(fp.f_copy_ptr[nf]);
    for (int s = 0; s < nf; ++s)
      for (auto& c : per_front[s]) fp.copies.push_back(c);
  }
}

}  // namespace sfx
