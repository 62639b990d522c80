"""Debug: per-task timeline of the tile-DAG factorization (final-shape BAL), summarised per level."""
import ctypes as C, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from symforce_b200 import capi, desc as D, problems as P
wl = sys.argv[1] if len(sys.argv) > 1 else "final"
prob = P.bal_problem(wl, solver=D.SOLVER_SCHUR)
g = capi.SfxProblem(prob)
g.solve_step(1.0)
n = C.c_int32()
g.lib.sfx_debug_trace_tasks(g.h, None, C.byref(n), None)
g.solve_step(1.0)
buf = np.zeros((n.value, 4), dtype=np.uint64)
info = np.zeros((n.value, 5), dtype=np.int16)
g.lib.sfx_debug_trace_tasks(g.h, buf.ctypes.data_as(C.POINTER(C.c_uint64)), C.byref(n), info.ctypes.data_as(C.POINTER(C.c_int16)))
np.savez_compressed("gpurun_out/factor_trace.npz", t=buf, info=info)
t = buf.astype(np.int64)
ok = t[:, 0] > 0
t0 = t[ok, 0].min()
claim, ready, done, potrf = (t[:, 0] - t0) / 1e3, (t[:, 1] - t0) / 1e3, (t[:, 2] - t0) / 1e3, (t[:, 3] - t0) / 1e3
for ty, name in [(3, "DIAG"), (1, "TRSM"), (2, "UPDATE"), (4, "RANGE"), (5, "INV"), (6, "EXTEND-ADD")]:
    if not (ok & (info[:, 1] == ty)).any():
        continue
    m = ok & (info[:, 1] == ty)
    print(f"{name}: n={m.sum()} wait(claim->ready) mean {np.mean(ready[m]-claim[m]):.1f} us  exec(ready->done) mean {np.mean(done[m]-ready[m]):.1f} us  p90 {np.percentile(done[m]-ready[m],90):.1f}")
m = ok & (info[:, 1] == 3)
print("DIAG potrf part (ready->potrf published) mean us:", np.mean(potrf[m] - ready[m]))
print(f"all tasks: span {done[ok].max() - claim[ok].min():.0f} us, busy {np.sum(done[ok] - ready[ok]) / 1e3:.1f} CTA-ms, "
      f"waiting {np.sum(ready[ok] - claim[ok]) / 1e3:.1f} CTA-ms")
# critical path of the biggest front: DIAG done times by k
lf = np.bincount(info[m, 0]).argmax()
mm = m & (info[:, 0] == lf)
order = np.argsort(info[mm, 2])
print("front", lf, "DIAG ready times (us) by k:", np.round(ready[mm][order][:12], 1), "done:", np.round(done[mm][order][:12], 1))

st = np.zeros((512, 8), dtype=np.uint64)
g.lib.sfx_debug_diag_stamps(st.ctypes.data_as(C.POINTER(C.c_uint64)))
st = st.astype(np.int64)
v = st[:40]
print("DIAG phases (us, last writer per k): potrf", np.round(np.mean((v[:, 1] - v[:, 0]) / 1e3), 2), "store+sync",
      np.round(np.mean((v[:, 2] - v[:, 1]) / 1e3), 2), "trsm(k+1,k)", np.round(np.mean((v[:, 3] - v[:, 2]) / 1e3), 2))

# per large front: span, number of tasks, busy time
for lf in np.unique(info[ok, 0]):
    mm = ok & (info[:, 0] == lf)
    md = mm & (info[:, 1] == 3)
    print(f"front {lf}: tasks {mm.sum()} diag {md.sum()} span {claim[mm].min():.0f}..{done[mm].max():.0f} us  "
          f"busy(sum exec) {np.sum(done[mm]-ready[mm])/1e3:.2f} ms  wait(sum) {np.sum(ready[mm]-claim[mm])/1e3:.2f} ms  "
          f"diag exec {np.mean(done[md]-ready[md]):.1f} us, diag step {(np.diff(np.sort(ready[md])).mean() if md.sum()>1 else 0):.1f} us")
