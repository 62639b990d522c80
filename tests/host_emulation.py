"""
numpy replay of what the CUDA kernels do with the host-built index maps (symforce_b200/csrc/
analysis.cc + symbolic.cc): scatter of per-factor J^T J / J^T r blocks, Schur complement from the
match lists, multifrontal Cholesky + solves from the front plan.  Executable specification used by
the CPU tests to validate the host logic against the oracle before any kernel runs.
"""
import numpy as np

from symforce_b200 import desc as D
from tests import oracle_capi as O


def factor_J(lib, kind, args):
    res, J, _, _ = O.eval_factor(lib, "orc_eval_factor", kind, args)
    return res, J


def emulate_linearize(problem, A):
    lib = O.load()
    vals = problem.values
    Hv = np.zeros(A["H"]["n_values"])
    rhs = np.zeros(A["N"])
    res = np.zeros(A["M"])
    for b in A["batches"]:
        meta = D.KINDS[b["kind"]]
        n = b["n"]
        used = b["used_args"]
        arg_off = np.array(b["arg_off"]).reshape(len(used), n)
        ng = b["n_groups"]
        rhs_off = np.array(b["rhs_off"]).reshape(ng, n) if ng else None
        diag_off = np.array(b["diag_off"]).reshape(ng, n) if ng else None
        npairs = ng * (ng - 1) // 2
        off_off = np.array(b["off_off"], dtype=np.uint64).reshape(npairs, n) if npairs else None
        cols = np.concatenate([[0], np.cumsum(meta["opt_dims"])])
        for s in range(n):
            args = []
            for ai, dim in enumerate(meta["arg_dims"]):
                if ai in used:
                    o = arg_off[used.index(ai), s]
                    args.append(vals[o:o + dim])
                else:
                    args.append(np.zeros(dim))
            r, J = factor_J(lib, b["kind"], args)
            ro = b["res_off"][s]
            res[ro:ro + len(r)] = r
            for ka in range(len(meta["opt_dims"])):
                g = b["key_group"][ka]
                if g < 0:
                    continue
                sa = b["key_sub"][ka]
                Ja = J[:, cols[ka]:cols[ka + 1]]
                rhs[rhs_off[g, s] + sa: rhs_off[g, s] + sa + Ja.shape[1]] += Ja.T @ r
                for kb in range(ka + 1):
                    h = b["key_group"][kb]
                    if h < 0:
                        continue
                    sb = b["key_sub"][kb]
                    Jb = J[:, cols[kb]:cols[kb + 1]]
                    blk = Ja.T @ Jb  # rows a, cols b
                    if g == h:
                        ld = b["group_dim"][g]
                        base = diag_off[g, s]
                        for c in range(blk.shape[1]):
                            for rr in range(blk.shape[0]):
                                if ka == kb and rr < c:
                                    continue
                                if ka == kb or sa > sb:
                                    o = base + (sa + rr) + (sb + c) * ld
                                else:
                                    o = base + (sb + c) + (sa + rr) * ld
                                Hv[o] += blk[rr, c]
                    else:
                        G, Hh = max(g, h), min(g, h)
                        ob = int(off_off[G * (G - 1) // 2 + Hh, s])
                        off = ob & 0x3FFFFFFF
                        excl = (ob >> 31) & 1
                        tr = (ob >> 30) & 1
                        a_is_row = (g == G) != bool(tr)
                        ld = b["group_dim"][g] if a_is_row else b["group_dim"][h]
                        for c in range(blk.shape[1]):
                            for rr in range(blk.shape[0]):
                                o = off + ((sa + rr) + (sb + c) * ld if a_is_row else (sb + c) + (sa + rr) * ld)
                                if excl:
                                    Hv[o] = blk[rr, c]
                                else:
                                    Hv[o] += blk[rr, c]
    return res, rhs, Hv


def damping_vector(A, Hv, lam, params):
    N = A["N"]
    d = np.zeros(N)
    if params.use_diagonal_damping:
        diag = Hv[np.array(A["diag_pos"])]
        d = np.maximum(diag, params.diagonal_damping_min) * lam
    if params.use_unit_damping:
        d = d + lam
    return d


def emulate_schur(A, Hv, rhs, dvec):
    sp = A["schur_plan"]
    nl = sp["n_landmarks"]
    S = sp["S"]
    cinv = np.zeros((nl, 3, 3))
    tl = np.zeros((nl, 3))
    for l in range(nl):
        d = sp["lm_dim"][l]
        to = sp["lm_toff"][l]
        Cb = Hv[sp["lm_cdiag_off"][l]: sp["lm_cdiag_off"][l] + d * d].reshape(d, d, order="F")
        Cb = np.tril(Cb) + np.tril(Cb, -1).T + np.diag(dvec[to:to + d])
        ci = np.linalg.inv(Cb)
        cinv[l, :d, :d] = ci
        tl[l, :d] = ci @ rhs[to:to + d]
    Sv = np.zeros(S["n_values"])
    nd = S["node_dim"]
    nto = S["node_off"]
    scol = np.zeros(len(S["row_idx"]), dtype=int)
    for j in range(len(nd)):
        scol[S["col_ptr"][j]:S["col_ptr"][j + 1]] = j
    for b in range(len(S["row_idx"])):
        I, J = S["row_idx"][b], scol[b]
        dI, dJ = nd[I], nd[J]
        acc = np.zeros((dI, dJ))
        for m in range(sp["s_m_ptr"][b], sp["s_m_ptr"][b + 1]):
            l = sp["m_lm"][m]
            dl = sp["lm_dim"][l]
            EI = Hv[sp["m_eoff_i"][m]: sp["m_eoff_i"][m] + dl * dI].reshape(dl, dI, order="F")
            EJ = Hv[sp["m_eoff_j"][m]: sp["m_eoff_j"][m] + dl * dJ].reshape(dl, dJ, order="F")
            acc += EI.T @ cinv[l, :dl, :dl] @ EJ
        out = -acc
        if sp["s_b_src"][b] >= 0:
            out += Hv[sp["s_b_src"][b]: sp["s_b_src"][b] + dI * dJ].reshape(dI, dJ, order="F")
        if I == J:
            out += np.diag(dvec[nto[I]:nto[I] + dI])
        Sv[S["blk_off"][b]: S["blk_off"][b] + dI * dJ] = out.reshape(-1, order="F")
    rhs_red = np.zeros(sp["reduced_dim"])
    for I in range(len(nd)):
        acc = np.zeros(nd[I])
        for q in range(sp["r_ptr"][I], sp["r_ptr"][I + 1]):
            l = sp["r_lm"][q]
            dl = sp["lm_dim"][l]
            E = Hv[sp["r_eoff"][q]: sp["r_eoff"][q] + dl * nd[I]].reshape(dl, nd[I], order="F")
            acc += E.T @ tl[l, :dl]
        rhs_red[nto[I]:nto[I] + nd[I]] = rhs[nto[I]:nto[I] + nd[I]] - acc
    return Sv, rhs_red, cinv, tl


def emulate_schur_back(A, Hv, cinv, tl, y):
    sp = A["schur_plan"]
    S = sp["S"]
    upd = np.zeros(A["N"])
    upd[: sp["reduced_dim"]] = -y
    for l in range(sp["n_landmarks"]):
        dl = sp["lm_dim"][l]
        s = np.zeros(dl)
        for q in range(sp["lm_e_ptr"][l], sp["lm_e_ptr"][l + 1]):
            J = sp["lm_e_node"][q]
            dJ = S["node_dim"][J]
            E = Hv[sp["lm_e_off"][q]: sp["lm_e_off"][q] + dl * dJ].reshape(dl, dJ, order="F")
            s += E @ y[S["node_off"][J]: S["node_off"][J] + dJ]
        z = tl[l, :dl] - cinv[l, :dl, :dl] @ s
        upd[sp["lm_toff"][l]: sp["lm_toff"][l] + dl] = -z
    return upd


def emulate_fronts_solve(A, sysvals, rhs_sys, dvec=None):
    """Multifrontal Cholesky + solves following the front plan. Returns x in system scalar order."""
    f = A["fronts"]
    nf = f["n_fronts"]
    fronts = [None] * nf
    sperm = np.array(f["scalar_perm"])
    order = np.argsort(np.array(f["f_level"]), kind="stable")
    for s in order:
        w, u = f["f_w"][s], f["f_u"][s]
        m = w + u
        F = np.zeros((m, m))
        for ci in range(f["f_copy_ptr"][s], f["f_copy_ptr"][s + 1]):
            src, rows, cols, ld, dr, dc, tr, lo = f["copies"][ci]
            blk = sysvals[src: src + ld * cols].reshape(ld, cols, order="F")[:rows, :]
            for c in range(cols):
                for r in range(rows):
                    if lo and r < c:
                        continue
                    if tr:
                        F[dr + c, dc + r] += blk[r, c]
                    else:
                        F[dr + r, dc + c] += blk[r, c]
        if dvec is not None:
            for r in range(w):
                F[r, r] += dvec[sperm[f["f_piv"][s] + r]]
        for ci in range(f["f_child_ptr"][s], f["f_child_ptr"][s + 1]):
            c = f["f_child"][ci]
            wc, uc = f["f_w"][c], f["f_u"][c]
            U = fronts[c][wc:, wc:]
            rel = f["f_rel"][f["f_rows_ptr"][c]: f["f_rows_ptr"][c + 1]]
            for j in range(uc):
                for i in range(j, uc):
                    assert rel[i] >= rel[j]
                    F[rel[i], rel[j]] += U[i, j]
        # partial cholesky on lower
        Fs = np.tril(F) + np.tril(F, -1).T
        L11 = np.linalg.cholesky(Fs[:w, :w])
        L21 = np.linalg.solve(L11, Fs[:w, w:]).T if u else np.zeros((0, w))
        U = Fs[w:, w:] - L21 @ L21.T
        out = np.zeros((m, m))
        out[:w, :w] = L11
        out[w:, :w] = L21
        out[w:, w:] = np.tril(U)
        fronts[s] = out
    n = f["n"]
    ywork = np.zeros(n)
    tw = [None] * nf
    for s in order:
        w, u = f["f_w"][s], f["f_u"][s]
        fv = np.zeros(w + u)
        fv[:w] = rhs_sys[sperm[f["f_piv"][s]: f["f_piv"][s] + w]]
        for ci in range(f["f_child_ptr"][s], f["f_child_ptr"][s + 1]):
            c = f["f_child"][ci]
            rel = f["f_rel"][f["f_rows_ptr"][c]: f["f_rows_ptr"][c + 1]]
            np.add.at(fv, rel, tw[c])
        L = fronts[s]
        y1 = np.linalg.solve(L[:w, :w], fv[:w])
        tw[s] = fv[w:] - L[w:, :w] @ y1
        ywork[f["f_piv"][s]: f["f_piv"][s] + w] = y1
    for s in order[::-1]:
        w, u = f["f_w"][s], f["f_u"][s]
        L = fronts[s]
        rows = f["f_rows"][f["f_rows_ptr"][s]: f["f_rows_ptr"][s + 1]]
        g = ywork[f["f_piv"][s]: f["f_piv"][s] + w] - L[w:, :w].T @ ywork[rows]
        ywork[f["f_piv"][s]: f["f_piv"][s] + w] = np.linalg.solve(L[:w, :w].T, g)
    x = np.zeros(n)
    x[sperm] = ywork
    return x


def emulate_solve_step(problem, A, lam):
    """Full replay: linearize -> damping -> [Schur] -> multifrontal solve -> update (reference order)."""
    res, rhs, Hv = emulate_linearize(problem, A)
    dvec = damping_vector(A, Hv, lam, problem.params)
    if A["schur"]:
        Sv, rhs_red, cinv, tl = emulate_schur(A, Hv, rhs, dvec)
        y = emulate_fronts_solve(A, Sv, rhs_red)
        upd = emulate_schur_back(A, Hv, cinv, tl, y)
    else:
        upd = -emulate_fronts_solve(A, Hv, rhs, dvec)
    return upd[np.array(A["ref2int"])], (res, rhs[np.array(A["ref2int"])], Hv[np.array(A["csc_src"])])


def emulate_jacobian(problem, A):
    """What jacobian_kernel writes: every factor's R x dim(key) Jacobian block at jac_base + c * jac_colnnz + r."""
    lib = O.load()
    vals = problem.values
    out = np.full(len(A["jac_inner"]), np.nan)
    for b in A["batches"]:
        meta = D.KINDS[b["kind"]]
        n = b["n"]
        used = b["used_args"]
        arg_off = np.array(b["arg_off"]).reshape(len(used), n)
        n_opt = len(meta["opt_dims"])
        base = np.array(b["jac_base"]).reshape(n_opt, n)
        coln = np.array(b["jac_colnnz"]).reshape(n_opt, n)
        cols = np.concatenate([[0], np.cumsum(meta["opt_dims"])])
        for s in range(n):
            args = []
            for ai, dim in enumerate(meta["arg_dims"]):
                if ai in used:
                    o = arg_off[used.index(ai), s]
                    args.append(vals[o:o + dim])
                else:
                    args.append(np.zeros(dim))
            r, J = factor_J(lib, b["kind"], args)
            for ka in range(n_opt):
                if base[ka, s] < 0:
                    assert b["key_group"][ka] < 0
                    continue
                for c in range(meta["opt_dims"][ka]):
                    p0 = base[ka, s] + c * coln[ka, s]
                    assert np.all(np.isnan(out[p0:p0 + len(r)])), "two factors write the same Jacobian entries"
                    out[p0:p0 + len(r)] = J[:, cols[ka] + c]
    return out
