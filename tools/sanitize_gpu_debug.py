"""Small GPU run for compute-sanitizer (memcheck / initcheck) of the debug_stats readers and the general-C covariance
block: robot_3d_localization with debug_stats (per-record update / values / residual snapshots, the Jacobian of a record
re-evaluated from its snapshot), then leading covariance blocks of that problem (solved without Schur elimination)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from symforce_b200 import capi, problems as P
prob = P.robot_3d_localization()
prob.params.debug_stats = 1
g = capi.SfxProblem(prob)
st = g.optimize()
n = len(g.iterations())
upd = [g.iteration_update(j) for j in range(n)]
jac = [g.iteration_jacobian(j) for j in range(n)]
print("records", n, "|update|", float(np.linalg.norm(upd[1])), "|J|", float(np.linalg.norm(jac[-1])), flush=True)
for b in (6, 12, 30):
    c = g.compute_covariance(b)
    print("covariance block", b, "trace", float(np.trace(c)), flush=True)
g.close()
