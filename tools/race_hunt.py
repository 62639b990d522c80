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
for rep in range(reps):
    u = gpu.solve_step(lam)
    if ref is None:
        ref = u.copy()
    d = np.abs(u - ref)
    bad = ~np.isfinite(u)
    rel = np.nanmax(d) / np.max(np.abs(ref))
    if bad.any() or rel > 1e-7:
        nbad += 1
        idx = np.flatnonzero(bad | (d > 1e-7 * np.max(np.abs(ref))))
        import ctypes as C
        cf = (C.c_int32 * 2)()
        gpu.lib.sfx_debug_chol_fail(gpu.h, cf)
        print(f"  chol_fail {cf[0]} first failing front {(cf[1] >> 16) - 1} pivot tile {cf[1] & 0xffff}")
        print(f"rep {rep}: nonfinite {bad.sum()} rel {rel:.2e} first bad idx {idx[:5]} last {idx[-5:]} count {idx.size} of {u.size}", flush=True)
print(f"{shape} lam {lam}: {nbad} bad of {reps}")
