// Host-only introspection of the structural analysis (no CUDA calls): lets the CPU test-suite
// replay the index maps, Schur match lists and the multifrontal plan in numpy and check them
// against the oracle before any kernel runs.  Small problems only (JSON text).
#include <cstring>
#include <sstream>

#include "sfx_internal.h"

namespace {
template <typename T>
void arr(std::ostringstream& o, const char* name, const std::vector<T>& v, bool comma = true) {
  o << '"' << name << "\":[";
  for (size_t i = 0; i < v.size(); ++i) {
    if (i) o << ',';
    o << (long long)v[i];
  }
  o << ']';
  if (comma) o << ',';
}
void blockmatrix(std::ostringstream& o, const char* name, const sfx::BlockMatrix& B) {
  o << '"' << name << "\":{";
  arr(o, "node_dim", B.node_dim);
  arr(o, "node_off", B.node_off);
  arr(o, "col_ptr", B.col_ptr);
  arr(o, "row_idx", B.row_idx);
  arr(o, "blk_off", B.blk_off);
  o << "\"n_values\":" << (long long)B.n_values << "},";
}
}  // namespace

extern "C" int sfx_debug_analysis_json(const sfx_problem_desc* d, char** out) {
  using namespace sfx;
  static thread_local std::string buf;
  try {
    Analysis a;
    analyze_problem(*d, a);
    build_csc(a);
    const BlockMatrix& sys = a.schur ? a.sp.S : a.H;
    std::vector<int> sys2ref;
    if (d->ordering == SFX_ORDERING_METIS_SCALAR) {
      std::vector<int> int2ref(a.N);
      for (int r = 0; r < a.N; ++r) int2ref[a.ref2int[r]] = r;
      sys2ref.assign(int2ref.begin(), int2ref.begin() + sys.node_off[sys.n_nodes]);
    }
    build_front_plan(sys, d->ordering, sys2ref, a.fp);
    std::ostringstream o;
    o << "{\"N\":" << a.N << ",\"M\":" << a.M << ",\"schur\":" << (a.schur ? 1 : 0) << ",";
    o << "\"h_accum_values\":" << (long long)a.h_accum_values << ",";
    std::vector<int> v1, v2, v3;
    for (auto& k : a.keys) {
      v1.push_back(k.node);
      v2.push_back(k.sub);
      v3.push_back(k.ref_toff);
    }
    arr(o, "key_node", v1);
    arr(o, "key_sub", v2);
    arr(o, "key_ref_toff", v3);
    arr(o, "ref2int", a.ref2int);
    arr(o, "diag_pos", a.diag_pos);
    blockmatrix(o, "H", a.H);
    arr(o, "csc_outer", a.csc_outer);
    arr(o, "csc_inner", a.csc_inner);
    arr(o, "csc_src", a.csc_src);
    if (a.world == 1) {
      build_jacobian_csc(a);
      arr(o, "jac_outer", a.jac_outer);
      arr(o, "jac_inner", a.jac_inner);
    }
    o << "\"batches\":[";
    for (size_t b = 0; b < a.batches.size(); ++b) {
      const BatchPlan& bp = a.batches[b];
      if (b) o << ',';
      o << "{\"kind\":" << bp.kind << ",\"n\":" << bp.n << ",\"n_groups\":" << bp.n_groups << ",";
      arr(o, "used_args", std::vector<int>(bp.used_args, bp.used_args + bp.n_used_args));
      arr(o, "key_group", std::vector<int>(bp.key_group, bp.key_group + bp.n_opt));
      arr(o, "key_sub", std::vector<int>(bp.key_sub, bp.key_sub + bp.n_opt));
      arr(o, "group_dim", std::vector<int>(bp.group_dim, bp.group_dim + bp.n_groups));
      arr(o, "arg_off", bp.arg_off);
      arr(o, "res_off", bp.res_off);
      arr(o, "rhs_off", bp.rhs_off);
      arr(o, "diag_off", bp.diag_off);
      arr(o, "off_off", bp.off_off);
      arr(o, "jac_base", bp.jac_base);
      arr(o, "jac_colnnz", bp.jac_colnnz);
      arr(o, "factor_index", bp.factor_index, false);
      o << '}';
    }
    o << "],";
    if (a.schur) {
      const SchurPlan& s = a.sp;
      o << "\"schur_plan\":{\"n_landmarks\":" << s.n_landmarks << ",\"first_lm_node\":" << s.first_lm_node
        << ",\"reduced_dim\":" << s.reduced_dim << ",";
      arr(o, "lm_dim", s.lm_dim);
      arr(o, "lm_cdiag_off", s.lm_cdiag_off);
      arr(o, "lm_toff", s.lm_toff);
      arr(o, "lm_e_ptr", s.lm_e_ptr);
      arr(o, "lm_e_off", s.lm_e_off);
      arr(o, "lm_e_node", s.lm_e_node);
      blockmatrix(o, "S", s.S);
      arr(o, "s_b_src", s.s_b_src);
      arr(o, "s_m_ptr", s.s_m_ptr);
      arr(o, "m_eoff_i", s.m_eoff_i);
      arr(o, "m_eoff_j", s.m_eoff_j);
      arr(o, "m_lm", s.m_lm);
      arr(o, "r_ptr", s.r_ptr);
      arr(o, "r_eoff", s.r_eoff);
      arr(o, "r_lm", s.r_lm, false);
      o << "},";
    }
    const FrontPlan& f = a.fp;
    o << "\"fronts\":{\"n\":" << f.n << ",\"n_fronts\":" << f.n_fronts << ",\"n_levels\":" << f.n_levels
      << ",\"front_values\":" << (long long)f.front_values << ",\"nnz_L\":" << (long long)f.nnz_L
      << ",\"max_front\":" << f.max_front << ",\"solve_ws\":" << (long long)f.solve_ws << ",";
    arr(o, "perm_nodes", f.perm_nodes);
    arr(o, "scalar_perm", f.scalar_perm);
    arr(o, "f_w", f.f_w);
    arr(o, "f_u", f.f_u);
    arr(o, "f_parent", f.f_parent);
    arr(o, "f_level", f.f_level);
    arr(o, "f_off", f.f_off);
    arr(o, "f_piv", f.f_piv);
    arr(o, "f_rows_ptr", f.f_rows_ptr);
    arr(o, "f_rows", f.f_rows);
    arr(o, "f_rel", f.f_rel);
    arr(o, "f_child_ptr", f.f_child_ptr);
    arr(o, "f_child", f.f_child);
    arr(o, "f_toff", f.f_toff);
    arr(o, "f_copy_ptr", f.f_copy_ptr);
    arr(o, "level_ptr", f.level_ptr);
    arr(o, "level_fronts", f.level_fronts);
    o << "\"copies\":[";
    for (size_t i = 0; i < f.copies.size(); ++i) {
      const auto& c = f.copies[i];
      if (i) o << ',';
      o << '[' << (long long)c.src << ',' << c.rows << ',' << c.cols << ',' << c.src_ld << ',' << c.dst_row << ','
        << c.dst_col << ',' << c.transposed << ',' << c.lower_only << ']';
    }
    o << "]}}";
    buf = o.str();
    *out = const_cast<char*>(buf.c_str());
    return 0;
  } catch (const std::exception& e) {
    buf = e.what();
    *out = const_cast<char*>(buf.c_str());
    return 1;
  }
}

// 64-bit digests of every array of the analysis (no plan of the fronts), one "name hash size" line each: lets a change of
// the host analysis be checked for identical output on problems far too large for the JSON dump (tools/analysis_digest.py).
namespace {
template <typename T>
void digest(std::ostringstream& o, const std::string& name, const std::vector<T>& v) {
  uint64_t h = 0xcbf29ce484222325ull;
  const unsigned char* p = reinterpret_cast<const unsigned char*>(v.data());
  const size_t nb = v.size() * sizeof(T);
  size_t i = 0;
  for (; i + 8 <= nb; i += 8) {
    uint64_t w;
    memcpy(&w, p + i, 8);
    h = (h ^ w) * 0x100000001b3ull;
    h ^= h >> 29;
  }
  for (; i < nb; ++i) h = (h ^ p[i]) * 0x100000001b3ull;
  o << name << ' ' << std::hex << h << std::dec << ' ' << v.size() << "\n";
}
void digest_bm(std::ostringstream& o, const std::string& name, const sfx::BlockMatrix& B) {
  digest(o, name + ".node_dim", B.node_dim);
  digest(o, name + ".node_off", B.node_off);
  digest(o, name + ".col_ptr", B.col_ptr);
  digest(o, name + ".row_idx", B.row_idx);
  digest(o, name + ".blk_off", B.blk_off);
  o << name << ".n_values " << (long long)B.n_values << "\n";
}
}  // namespace

extern "C" int sfx_debug_analysis_digest(const sfx_problem_desc* d, char** out) {
  using namespace sfx;
  static thread_local std::string buf;
  try {
    Analysis a;
    analyze_problem(*d, a);
    std::ostringstream o;
    o << "N " << a.N << " M " << a.M << " b_values " << (long long)a.b_values << " h_accum_values "
      << (long long)a.h_accum_values << "\n";
    std::vector<int> v1, v2;
    for (auto& k : a.keys) {
      v1.push_back(k.node);
      v2.push_back(k.sub);
    }
    digest(o, "key_node", v1);
    digest(o, "key_sub", v2);
    digest(o, "ref2int", a.ref2int);
    digest(o, "diag_pos", a.diag_pos);
    digest_bm(o, "H", a.H);
    for (size_t b = 0; b < a.batches.size(); ++b) {
      const BatchPlan& bp = a.batches[b];
      const std::string n = "batch" + std::to_string(b);
      o << n << " kind " << bp.kind << " n " << bp.n << " groups " << bp.n_groups << "\n";
      digest(o, n + ".arg_off", bp.arg_off);
      digest(o, n + ".res_off", bp.res_off);
      digest(o, n + ".rhs_off", bp.rhs_off);
      digest(o, n + ".diag_off", bp.diag_off);
      digest(o, n + ".off_off", bp.off_off);
      digest(o, n + ".factor_index", bp.factor_index);
      if (bp.kind == SFX_KIND_SNAVELY && bp.n_groups == 2) {
        std::vector<int32_t> order, ptr, dg, rh;
        PhaseClock clk;
        build_point_lists(bp, order, ptr, dg, rh);
        clk.lap("point lists");
        digest(o, n + ".pf_slot", order);
        digest(o, n + ".pf_ptr", ptr);
        digest(o, n + ".pf_diag", dg);
        digest(o, n + ".pf_rhs", rh);
      }
    }
    if (a.schur) {
      const SchurPlan& s = a.sp;
      o << "schur n_landmarks " << s.n_landmarks << " lm_begin " << s.lm_begin << " reduced_dim " << s.reduced_dim << "\n";
      digest(o, "lm_dim", s.lm_dim);
      digest(o, "lm_cdiag_off", s.lm_cdiag_off);
      digest(o, "lm_toff", s.lm_toff);
      digest(o, "lm_e_ptr", s.lm_e_ptr);
      digest(o, "lm_e_off", s.lm_e_off);
      digest(o, "lm_e_node", s.lm_e_node);
      digest_bm(o, "S", s.S);
      digest(o, "s_b_src", s.s_b_src);
      digest(o, "s_m_ptr", s.s_m_ptr);
      digest(o, "m_eoff_i", s.m_eoff_i);
      digest(o, "m_eoff_j", s.m_eoff_j);
      digest(o, "m_lm", s.m_lm);
      digest(o, "r_ptr", s.r_ptr);
      digest(o, "r_eoff", s.r_eoff);
      digest(o, "r_lm", s.r_lm);
    }
    buf = o.str();
    *out = const_cast<char*>(buf.c_str());
    return 0;
  } catch (const std::exception& e) {
    buf = e.what();
    *out = const_cast<char*>(buf.c_str());
    return 1;
  }
}

// Per-front summary (level, w, u) of the multifrontal plan as text; cheap even for large problems.
extern "C" int sfx_debug_front_summary(const sfx_problem_desc* d, char** out) {
  using namespace sfx;
  static thread_local std::string buf;
  try {
    Analysis a;
    analyze_problem(*d, a);
    const BlockMatrix& sys = a.schur ? a.sp.S : a.H;
    std::vector<int> sys2ref;
    if (d->ordering == SFX_ORDERING_METIS_SCALAR) {
      std::vector<int> int2ref(a.N);
      for (int r = 0; r < a.N; ++r) int2ref[a.ref2int[r]] = r;
      sys2ref.assign(int2ref.begin(), int2ref.begin() + sys.node_off[sys.n_nodes]);
    }
    build_front_plan(sys, d->ordering, sys2ref, a.fp);
    std::ostringstream o;
    o << "n=" << a.fp.n << " fronts=" << a.fp.n_fronts << " levels=" << a.fp.n_levels << " nnzL=" << a.fp.nnz_L
      << " flops=" << a.fp.flops << " sblocks=" << (a.schur ? a.sp.S.row_idx.size() : 0)
      << " matches=" << (a.schur ? a.sp.m_lm.size() : 0) << "\n";
    for (int s = 0; s < a.fp.n_fronts; ++s)
      o << a.fp.f_level[s] << " " << a.fp.f_w[s] << " " << a.fp.f_u[s] << " " << a.fp.f_parent[s] << "\n";
    buf = o.str();
    *out = const_cast<char*>(buf.c_str());
    return 0;
  } catch (const std::exception& e) {
    buf = e.what();
    *out = const_cast<char*>(buf.c_str());
    return 1;
  }
}
