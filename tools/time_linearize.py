"""Debug: time linearize_bal_kernel with parts left out (sfx_debug_time_linearize) -- cost breakdown."""
import ctypes as C, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from symforce_b200 import capi, desc as D, problems as P
wl = sys.argv[1] if len(sys.argv) > 1 else "final"
prob = P.bal_problem(wl, solver=D.SOLVER_SCHUR)
g = capi.SfxProblem(prob)
names = {0: "full", 1: "no camera block", 2: "no point block", 4: "no E block", 7: "factor + residual only",
         8: "no factor arithmetic", 15: "loads + residual store only"}
for skip in [0, 1, 2, 4, 7, 8, 15, 0]:
    ms = C.c_float()
    rc = g.lib.sfx_debug_time_linearize(g.h, C.c_int32(skip), C.c_int32(10), C.byref(ms))
    print(f"skip={skip:2d} {names[skip]:32s} rc={rc} {ms.value:.4f} ms", flush=True)
