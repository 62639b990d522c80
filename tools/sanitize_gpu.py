"""Small GPU run for compute-sanitizer (memcheck / racecheck): a few LM iterations of BAL ladybug (tile-DAG Cholesky with
several tiles per front, range updates, the s9 Schur kernel with multi-chunk blocks, v2 solves), of a pose graph and of a
600-camera BAL problem whose 25 fronts run as one fused launch."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from symforce_b200 import capi, desc as D, problems as P
# bal_600: 25 tile-DAG fronts in one fused launch (extend-add tasks, sticky chains, forward substitution inside the kernel)
for name, prob in [("ladybug", P.bal_problem("ladybug", solver=D.SOLVER_SCHUR)), ("pose_graph_300", P.pose_graph_problem(300, 60)),
                   ("bal_600", P.bal_problem(n_cams=600, n_pts=30000, n_obs=150000, window=8))]:
    g = capi.SfxProblem(prob)
    st = g.optimize(3)
    print(name, "status", st.status, "error", g.iterations()[st.best_index].new_error, flush=True)
    g.close()
