"""Debug: repeated Final-shape (or other BAL shape) solves on the GPU against the oracle fixture; prints where a run departs."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from symforce_b200 import capi, desc as D, problems as P
shape = sys.argv[1] if len(sys.argv) > 1 else "final"
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 5
gold = json.load(open(f"tests/golden/{shape}_shape_history.json"))
params = D.default_params(); params.lambda_update_type = D.LAMBDA_DYNAMIC; params.iterations = 50
prob = P.bal_problem(shape, solver=D.SOLVER_SCHUR, params=params)
gpu = capi.SfxProblem(prob, device=0)
for rep in range(reps):
    gpu.set_values(prob.values)
    st = gpu.optimize()
    its = gpu.iterations()
    bad = None
    worst = 0.0
    for r, g in zip(its, gold["records"]):
        rel = abs(r.new_error - g["new_error"]) / g["new_error"]
        if not (rel <= 1e-8) or r.update_accepted != g["update_accepted"]:
            bad = (r.iteration, r.new_error, g["new_error"], r.update_accepted, g["update_accepted"])
            break
        worst = max(worst, rel)
    print(f"rep {rep}: status {st.status} records {len(its)} (gold {gold['n_records']}) worst rel {worst:.2e} first departure {bad}", flush=True)
