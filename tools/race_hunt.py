"""Debug: repeats sfx_solve_step on fixed values and reports deviations between repetitions (race detector)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from symforce_b200 import capi, desc as D, problems as P
shape = sys.argv[1] if len(sys.argv) > 1 else "final"
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 20
lam = float(sys.argv[3]) if len(sys.argv) > 3 else 1e-3
prob = P.bal_problem(shape, solver=D.SOLVER_SCHUR)
gpu = capi.SfxProblem(prob, device=0)
ref = None
nbad = 0
import ctypes as C
nS = C.c_int64()
gpu.lib.sfx_debug_read_S(gpu.h, None, C.byref(nS))
Sref = None
Fref = None
nF = C.c_int64()
lfi = np.zeros((64, 5), dtype=np.int64)
nlf = C.c_int32()
gpu.lib.sfx_debug_read_fronts(gpu.h, None, C.byref(nF), lfi.ctypes.data_as(C.POINTER(C.c_int64)), 64, C.byref(nlf))
def read_F():
    out = np.empty(nF.value)
    gpu.lib.sfx_debug_read_fronts(gpu.h, out.ctypes.data_as(C.POINTER(C.c_double)), C.byref(nF), None, 0, None)
    return out
def tiles(m, w, wt, nt):
    st = [t * 64 if t < wt else w + (t - wt) * 64 for t in range(nt)]
    sz = [min(64, w - t * 64) if t < wt else min(64, m - st[t]) for t in range(nt)]
    return st, sz
def compare_fronts(Fb):
    for li in range(nlf.value):
        off, m, w, wt, nt = [int(x) for x in lfi[li]]
        A = Fref[off:off + m * m].reshape(m, m).T  # column-major
        B = Fb[off:off + m * m].reshape(m, m).T
        st, sz = tiles(m, w, wt, nt)
        scale = np.nanmax(np.abs(np.tril(A)))
        badt = []
        for j in range(nt):
            for i in range(j, nt):
                a = A[st[i]:st[i] + sz[i], st[j]:st[j] + sz[j]]
                b = B[st[i]:st[i] + sz[i], st[j]:st[j] + sz[j]]
                if i == j:
                    a, b = np.tril(a), np.tril(b)
                bad = (~np.isfinite(b)) | (np.abs(a - b) > 1e-9 * scale)
                if bad.any():
                    cols = np.flatnonzero(bad.any(axis=0)); rows = np.flatnonzero(bad.any(axis=1))
                    badt.append((i, j, int(bad.sum()), int((~np.isfinite(b)).sum()), int((~np.isfinite(a)).sum()),
                                 (int(rows[0]), int(rows[-1])), (int(cols[0]), int(cols[-1]))))
        if badt:
            badt.sort(key=lambda x: (x[1], x[0]))
            print(f"  front {li} (m {m} w {w} wt {wt} nt {nt}): {len(badt)} tiles differ; (i, j, n_bad, n_nan, n_nan_ref, rows, cols): {badt[:5]}")
            return
def read_S():
    out = np.empty(nS.value)
    gpu.lib.sfx_debug_read_S(gpu.h, out.ctypes.data_as(C.POINTER(C.c_double)), C.byref(nS))
    return out
for rep in range(reps):
    u = gpu.solve_step(lam)
    if ref is None:
        ref = u.copy()
        Sref = read_S()
        Fref = read_F()
    d = np.abs(u - ref)
    bad = ~np.isfinite(u)
    rel = np.nanmax(d) / np.max(np.abs(ref))
    if bad.any() or rel > 1e-7:
        nbad += 1
        idx = np.flatnonzero(bad | (d > 1e-7 * np.max(np.abs(ref))))
        import ctypes as C
        cf = (C.c_int32 * 2)()
        gpu.lib.sfx_debug_chol_fail(gpu.h, cf)
        Sb = read_S()
        dS = np.abs(Sb - Sref)
        print(f"  S: nonfinite {(~np.isfinite(Sb)).sum()} max abs diff {np.nanmax(dS):.3e} (max |S| {np.max(np.abs(Sref)):.3e}) entries off by >1e-6*max: {(dS > 1e-6 * np.max(np.abs(Sref))).sum()}")
        compare_fronts(read_F())
        print(f"  chol_fail {cf[0]} first failing front {(cf[1] >> 16) - 1} pivot tile {cf[1] & 0xffff}")
        print(f"rep {rep}: nonfinite {bad.sum()} rel {rel:.2e} first bad idx {idx[:5]} last {idx[-5:]} count {idx.size} of {u.size}", flush=True)
print(f"{shape} lam {lam}: {nbad} bad of {reps}")
