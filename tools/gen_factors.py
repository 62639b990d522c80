#!/usr/bin/env python
"""
Factor-function generator for symforce_b200.

Runs ONLY in the development container (needs /root/reference on PYTHONPATH, SymPy backend):

    SYMFORCE_SYMBOLIC_API=sympy PYTHONPATH=/root/reference:/root/reference/gen/python \
        python tools/gen_factors.py

It imports the reference's *symbolic* residual definitions (the authoring-time layer, which is
not part of the product), differentiates them in the tangent space with the reference's own
`jacobian_helpers.tangent_jacobians`, runs SymPy CSE, and prints residual + stacked Jacobian as
straight-line fp64 code with OUR printer and OUR calling convention (raw storage pointers).

Residual definitions used (reference file:line):
  snavely                 symforce/examples/bundle_adjustment_in_the_large/bundle_adjustment_in_the_large.py:18-63
  between / prior         symforce/codegen/geo_factors_codegen.py:53-86
  matching / odometry     symforce/examples/robot_3d_localization/robot_3d_localization.py:115-150
  inverse-range landmark  symforce/codegen/slam_factors_codegen.py:28-50, 162-229
  barron (GNC test)       test/symforce_gnc_codegen_test.py:24-33

Two files are written, from the same expression DAG but as independent translation units:
  symforce_b200/csrc/gen/factors_gen.cuh   (__device__ functions used by the CUDA linearize kernels)
  oracle/gen/factors_gen.h                 (plain C++ used by the CPU oracle)
The Gauss-Newton H = J^T J (lower) and rhs = J^T r are formed by the callers, which is the
algebraic definition the reference codegen uses (symforce/codegen/codegen.py:796-807).
"""

import os
import sys
import time
from pathlib import Path

os.environ.setdefault("SYMFORCE_SYMBOLIC_API", "sympy")
sys.path.insert(0, "/root/reference")
sys.path.insert(0, "/root/reference/gen/python")

import symforce

symforce.set_epsilon_to_symbol()

import sympy  # noqa: E402
import symforce.symbolic as sf  # noqa: E402
from symforce import jacobian_helpers  # noqa: E402
from symforce import ops  # noqa: E402
from symforce.codegen import geo_factors_codegen, slam_factors_codegen  # noqa: E402
from symforce.examples.bundle_adjustment_in_the_large import (  # noqa: E402
    bundle_adjustment_in_the_large as bal,
)

ROOT = Path(__file__).resolve().parent.parent


def _robot3d_residuals():
    """
    The robot_3d_localization module cannot be imported here (it pulls in symforce.opt.factor ->
    graphviz / cc_sym at import time), so its two residuals are restated from their definition
    (symforce/examples/robot_3d_localization/robot_3d_localization.py:115-150):
      matching:  (world_T_body^-1 * world_t_landmark - body_t_landmark) / sigma
      odometry:  diag(sigmas)^-1 * local_coordinates(world_T_a^-1 * world_T_b, a_T_b)
    """

    def matching_residual(world_T_body, world_t_landmark, body_t_landmark, sigma):
        predicted = world_T_body.inverse() * world_t_landmark
        return (predicted - body_t_landmark) / sigma

    def odometry_residual(world_T_a, world_T_b, a_T_b, diagonal_sigmas, epsilon):
        predicted = world_T_a.inverse() * world_T_b
        err = sf.V6(predicted.local_coordinates(a_T_b, epsilon=epsilon))
        return sf.V6([err[i] / diagonal_sigmas[i] for i in range(6)])

    return matching_residual, odometry_residual


# ----------------------------------------------------------------------------------------------
# Kind table.  arg spec: (name, type, storage_dim).  opt: names of optimized args, J column order.
# ----------------------------------------------------------------------------------------------


def M66():
    return sf.M66.symbolic("sqrt_info")


def M33(name):
    return sf.M33.symbolic(name)


def build_kinds():
    matching_residual, odometry_residual = _robot3d_residuals()
    eps = sf.Symbol("epsilon")
    kinds = []

    def add(name, args, opt, res):
        kinds.append(dict(name=name, args=args, opt=opt, res=res))

    # 0: Snavely reprojection
    a = [
        ("cam_T_world", sf.Pose3.symbolic("cam_T_world")),
        ("intrinsics", sf.V3.symbolic("intrinsics")),
        ("point", sf.V3.symbolic("point")),
        ("pixel", sf.V2.symbolic("pixel")),
        ("epsilon", eps),
    ]
    add(
        "snavely",
        a,
        ["cam_T_world", "intrinsics", "point"],
        bal.snavely_reprojection_residual(*[x[1] for x in a]),
    )

    # 1: Between Pose3
    a = [
        ("a", sf.Pose3.symbolic("a")),
        ("b", sf.Pose3.symbolic("b")),
        ("a_T_b", sf.Pose3.symbolic("a_T_b")),
        ("sqrt_info", sf.M66.symbolic("sqrt_info")),
        ("epsilon", eps),
    ]
    add("between_pose3", a, ["a", "b"], geo_factors_codegen.between_factor(*[x[1] for x in a]))

    # 2: Prior Pose3
    a = [
        ("value", sf.Pose3.symbolic("value")),
        ("prior", sf.Pose3.symbolic("prior")),
        ("sqrt_info", sf.M66.symbolic("sqrt_info")),
        ("epsilon", eps),
    ]
    add("prior_pose3", a, ["value"], geo_factors_codegen.prior_factor(*[x[1] for x in a]))

    # 3: Matching (robot_3d_localization)
    a = [
        ("world_T_body", sf.Pose3.symbolic("world_T_body")),
        ("world_t_landmark", sf.V3.symbolic("world_t_landmark")),
        ("body_t_landmark", sf.V3.symbolic("body_t_landmark")),
        ("sigma", sf.Symbol("sigma")),
    ]
    add("matching", a, ["world_T_body"], matching_residual(*[x[1] for x in a]))

    # 4: Odometry (robot_3d_localization)
    a = [
        ("world_T_a", sf.Pose3.symbolic("world_T_a")),
        ("world_T_b", sf.Pose3.symbolic("world_T_b")),
        ("a_T_b", sf.Pose3.symbolic("a_T_b")),
        ("diagonal_sigmas", sf.V6.symbolic("diagonal_sigmas")),
        ("epsilon", eps),
    ]
    add("odometry", a, ["world_T_a", "world_T_b"], odometry_residual(*[x[1] for x in a]))

    # 5: Inverse range landmark, linear camera, GNC (Barron) noise
    a = [
        ("source_pose", sf.Pose3.symbolic("source_pose")),
        ("source_calibration", sf.LinearCameraCal.symbolic("source_calibration")),
        ("target_pose", sf.Pose3.symbolic("target_pose")),
        ("target_calibration", sf.LinearCameraCal.symbolic("target_calibration")),
        ("source_inverse_range", sf.Symbol("source_inverse_range")),
        ("source_pixel", sf.V2.symbolic("source_pixel")),
        ("target_pixel", sf.V2.symbolic("target_pixel")),
        ("weight", sf.Symbol("weight")),
        ("gnc_mu", sf.Symbol("gnc_mu")),
        ("gnc_scale", sf.Symbol("gnc_scale")),
        ("epsilon", eps),
    ]
    add(
        "irl_linear_gnc",
        a,
        ["source_pose", "target_pose", "source_inverse_range"],
        slam_factors_codegen.inverse_range_landmark_gnc_residual(*[x[1] for x in a]),
    )

    # 6: Inverse range landmark prior
    a = [
        ("landmark_inverse_range", sf.Symbol("landmark_inverse_range")),
        ("inverse_range_prior", sf.Symbol("inverse_range_prior")),
        ("weight", sf.Symbol("weight")),
        ("sigma", sf.Symbol("sigma")),
        ("epsilon", eps),
    ]
    add(
        "irl_prior",
        a,
        ["landmark_inverse_range"],
        slam_factors_codegen.inverse_range_landmark_prior_residual(*[x[1] for x in a]),
    )

    # 7: Between Rot3
    a = [
        ("a", sf.Rot3.symbolic("a")),
        ("b", sf.Rot3.symbolic("b")),
        ("a_T_b", sf.Rot3.symbolic("a_T_b")),
        ("sqrt_info", sf.M33.symbolic("sqrt_info")),
        ("epsilon", eps),
    ]
    add("between_rot3", a, ["a", "b"], geo_factors_codegen.between_factor(*[x[1] for x in a]))

    # 8: Prior Rot3
    a = [
        ("value", sf.Rot3.symbolic("value")),
        ("prior", sf.Rot3.symbolic("prior")),
        ("sqrt_info", sf.M33.symbolic("sqrt_info")),
        ("epsilon", eps),
    ]
    add("prior_rot3", a, ["value"], geo_factors_codegen.prior_factor(*[x[1] for x in a]))

    # 9: Barron-robust difference of two 5-vectors, the factor of the reference's GNC test
    # (test/symforce_gnc_codegen_test.py:24-33, generated as gnc_factors::BarronFactor and used by
    # test/symforce_gnc_test.cc:44-48): whiten(x - y) with alpha = compute_alpha_from_mu(mu, eps), delta = 1
    from symforce.opt.noise_models import BarronNoiseModel  # noqa: PLC0415

    a = [
        ("x", sf.V5.symbolic("x")),
        ("y", sf.V5.symbolic("y")),
        ("mu", sf.Symbol("mu")),
        ("eps", sf.Symbol("eps")),
    ]
    alpha = BarronNoiseModel.compute_alpha_from_mu(a[2][1], a[3][1])
    noise_model = BarronNoiseModel(alpha=alpha, delta=1, scalar_information=1, x_epsilon=a[3][1])
    add("barron", a, ["x"], noise_model.whiten(a[0][1] - a[1][1]))
    return kinds


# ----------------------------------------------------------------------------------------------
# Printer
# ----------------------------------------------------------------------------------------------
from sympy.printing.c import C99CodePrinter  # noqa: E402


class Printer(C99CodePrinter):
    """fp64 straight-line printer: integer powers expanded, rationals as double literals."""

    def _print_Pow(self, expr):
        base, exp = expr.as_base_exp()
        if exp.is_Integer:
            n = int(exp)
            b = self.parenthesize(base, 100)
            if n == -1:
                return f"(1.0 / {b})"
            if 2 <= abs(n) <= 4:
                prod = " * ".join([b] * abs(n))
                return f"({prod})" if n > 0 else f"(1.0 / ({prod}))"
        if exp == sympy.Rational(1, 2):
            return f"sqrt({self._print(base)})"
        if exp == sympy.Rational(-1, 2):
            return f"(1.0 / sqrt({self._print(base)}))"
        if exp == sympy.Rational(3, 2):
            b = self._print(base)
            return f"(({b}) * sqrt({b}))"
        if exp == sympy.Rational(-3, 2):
            b = self._print(base)
            return f"(1.0 / (({b}) * sqrt({b})))"
        return f"pow({self._print(base)}, {self._print(exp)})"

    def _print_Rational(self, expr):
        return f"({int(expr.p)}.0 / {int(expr.q)}.0)"

    def _print_Integer(self, expr):
        return f"{int(expr)}.0" if abs(int(expr)) < (1 << 52) else super()._print_Integer(expr)

    def _print_sign(self, expr):
        a = self._print(expr.args[0])
        return f"SFX_SIGN({a})"

    def _print_Max(self, expr):
        args = [self._print(a) for a in expr.args]
        out = args[0]
        for a in args[1:]:
            out = f"fmax({out}, {a})"
        return out

    def _print_Min(self, expr):
        args = [self._print(a) for a in expr.args]
        out = args[0]
        for a in args[1:]:
            out = f"fmin({out}, {a})"
        return out

    def _print_Abs(self, expr):
        return f"fabs({self._print(expr.args[0])})"

    # same semantics the reference C++ printer gives these two
    # (symforce/codegen/backends/cpp/cpp_code_printer.py:166-173)
    def _print_SignNoZero(self, expr):
        return f"copysign(1.0, {self._print(expr.args[0])})"

    def _print_CopysignNoZero(self, expr):
        a, b = (self._print(x) for x in expr.args)
        return f"copysign({a}, {b})"

    def _print_Function(self, expr):
        name = expr.func.__name__
        if name == "copysign_no_zero":
            a, b = (self._print(x) for x in expr.args)
            return f"copysign({a}, {b})"
        if name == "sign_no_zero":
            return f"SFX_SIGN_NO_ZERO({self._print(expr.args[0])})"
        return super()._print_Function(expr)


def storage_syms(x):
    if isinstance(x, (sf.Symbol, sympy.Symbol)):
        return [x]
    return list(ops.StorageOps.to_storage(x))


def tangent_dim(x):
    if isinstance(x, (sf.Symbol, sympy.Symbol)):
        return 1
    return ops.LieGroupOps.tangent_dim(x)


def gen_kind(kind):
    t0 = time.time()
    name = kind["name"]
    args = kind["args"]
    res = kind["res"]
    if not isinstance(res, sf.Matrix):
        res = sf.M(res)
    opt_elems = [dict(args)[n] for n in kind["opt"]]
    jac_blocks = jacobian_helpers.tangent_jacobians(res, opt_elems)
    J = sf.Matrix.block_matrix([jac_blocks])
    R = res.shape[0]
    Tn = J.shape[1]
    assert J.shape[0] == R

    outs = [res[i, 0] for i in range(R)]
    # column-major J
    outs += [J[r, c] for c in range(Tn) for r in range(R)]
    outs = [sympy.sympify(o) for o in outs]

    # Replace storage symbols by indexed names arg{k}[i]
    sub = {}
    arg_dims = []
    for k, (aname, a) in enumerate(args):
        syms = storage_syms(a)
        arg_dims.append(len(syms))
        for i, s in enumerate(syms):
            sub[s] = sympy.Symbol(f"a{k}_{i}")
    outs = [o.xreplace(sub) for o in outs]

    repl, reduced = sympy.cse(outs, symbols=sympy.numbered_symbols("t"), order="none")
    used = set()
    for _, e in repl:
        used |= {str(s) for s in e.free_symbols}
    for e in reduced:
        used |= {str(s) for s in e.free_symbols}
    arg_used = [any(f"a{k}_{i}" in used for i in range(arg_dims[k])) for k in range(len(args))]

    pr = Printer()
    n_ops = sum(sympy.count_ops(e) for _, e in repl) + sum(sympy.count_ops(e) for e in reduced)
    body = []
    for k, (aname, a) in enumerate(args):
        for i in range(arg_dims[k]):
            if f"a{k}_{i}" in used:
                body.append(f"  const double a{k}_{i} = a{k}[{i}];")
    for s, e in repl:
        body.append(f"  const double {s} = {pr.doprint(e)};")
    for i in range(R):
        body.append(f"  res[{i}] = {pr.doprint(reduced[i])};")
    for j in range(R * Tn):
        body.append(f"  J[{j}] = {pr.doprint(reduced[R + j])};")

    sig_args = ", ".join(f"const double* __restrict__ a{k}" for k in range(len(args)))
    doc = [
        f"// kind '{name}': residual dim {R}, tangent dim {Tn}; J is column-major [{R} x {Tn}]",
        "// args: " + ", ".join(f"a{k}={aname}[{arg_dims[k]}]" for k, (aname, _) in enumerate(args)),
        "// optimized args (J column order): "
        + ", ".join(f"{n}({tangent_dim(dict(args)[n])})" for n in kind["opt"]),
        f"// ~{n_ops} ops, {len(repl)} temporaries",
    ]
    fn = "\n".join(doc) + "\n"
    fn += f"SFX_FACTOR_FN void sfx_factor_{name}({sig_args}, double* __restrict__ res, double* __restrict__ J) {{\n"
    for k in range(len(args)):
        if not arg_used[k]:
            fn += f"  (void)a{k};\n"
    fn += "\n".join(body) + "\n}\n"

    meta = dict(
        name=name,
        n_args=len(args),
        arg_names=[a[0] for a in args],
        arg_dims=arg_dims,
        arg_used=arg_used,
        opt_args=[[a[0] for a in args].index(n) for n in kind["opt"]],
        opt_dims=[tangent_dim(dict(args)[n]) for n in kind["opt"]],
        res_dim=R,
        tan_dim=Tn,
        n_ops=int(n_ops),
    )
    print(f"[gen] {name}: R={R} T={Tn} ops~{n_ops} temps={len(repl)}  ({time.time() - t0:.1f}s)")
    return fn, meta


HEADER = """// GENERATED by tools/gen_factors.py -- do not edit.
// Straight-line fp64 residual + tangent-space Jacobian for each supported factor kind, derived
// from the reference's symbolic residual definitions (see tools/gen_factors.py for file:line).
#pragma once
#include <math.h>
#ifndef SFX_SIGN
#define SFX_SIGN(x) ((double)(((x) > 0.0) - ((x) < 0.0)))
#define SFX_SIGN_NO_ZERO(x) (((x) >= 0.0) ? 1.0 : -1.0)
#endif
"""


def meta_table(metas):
    lines = []
    lines.append(f"#define SFX_NUM_KINDS {len(metas)}")
    lines.append("#define SFX_MAX_ARGS 11")
    lines.append("#define SFX_MAX_OPT 3")
    lines.append("struct sfx_kind_meta { const char* name; int n_args; int arg_dims[SFX_MAX_ARGS]; "
                 "int arg_used[SFX_MAX_ARGS]; int n_opt; int opt_args[SFX_MAX_OPT]; "
                 "int opt_dims[SFX_MAX_OPT]; int res_dim; int tan_dim; int n_ops; };")
    lines.append("static const sfx_kind_meta SFX_KIND_META[SFX_NUM_KINDS] = {")
    for m in metas:
        ad = ", ".join(str(x) for x in m["arg_dims"] + [0] * (11 - len(m["arg_dims"])))
        au = ", ".join(str(int(x)) for x in m["arg_used"] + [0] * (11 - len(m["arg_used"])))
        oa = ", ".join(str(x) for x in m["opt_args"] + [-1] * (3 - len(m["opt_args"])))
        od = ", ".join(str(x) for x in m["opt_dims"] + [0] * (3 - len(m["opt_dims"])))
        lines.append(
            f'  {{"{m["name"]}", {m["n_args"]}, {{{ad}}}, {{{au}}}, {len(m["opt_args"])}, '
            f'{{{oa}}}, {{{od}}}, {m["res_dim"]}, {m["tan_dim"]}, {m["n_ops"]}}},'
        )
    lines.append("};")
    return "\n".join(lines) + "\n"


def main():
    kinds = build_kinds()
    fns, metas = [], []
    for k in kinds:
        fn, meta = gen_kind(k)
        fns.append(fn)
        metas.append(meta)

    dev = ROOT / "symforce_b200" / "csrc" / "gen"
    cpu = ROOT / "oracle" / "gen"
    dev.mkdir(parents=True, exist_ok=True)
    cpu.mkdir(parents=True, exist_ok=True)

    (dev / "factors_gen.cuh").write_text(
        HEADER + "#define SFX_FACTOR_FN static __device__ __forceinline__\n\n" + "\n".join(fns)
    )
    (dev / "kinds_gen.h").write_text(
        "// GENERATED by tools/gen_factors.py -- do not edit.\n#pragma once\n" + meta_table(metas)
    )
    (cpu / "factors_gen.h").write_text(
        HEADER + "#define SFX_FACTOR_FN static inline\n\n" + "\n".join(fns)
    )
    (cpu / "kinds_gen.h").write_text(
        "// GENERATED by tools/gen_factors.py -- do not edit.\n#pragma once\n"
        + meta_table(metas).replace("sfx_kind_meta", "orc_kind_meta").replace("SFX_", "ORC_")
    )
    import json

    (ROOT / "symforce_b200" / "kinds.json").write_text(json.dumps(metas, indent=1))


if __name__ == "__main__":
    main()
