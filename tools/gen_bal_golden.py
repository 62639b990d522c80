#!/usr/bin/env python
"""
Generates tests/golden/<shape>_shape_history.json: the LM iteration history of the CPU oracle (oracle/oracle.cc,
the restatement of symforce/opt/levenberg_marquardt_solver.tcc:139-343 pinned to the reference's KATs) on a synthetic
BAL shape with the reference's BAL parameters (DefaultOptimizerParams + DYNAMIC lambda,
symforce/examples/bundle_adjustment_in_the_large/bundle_adjustment_in_the_large.cc:133-136).

    python tools/gen_bal_golden.py final      # ~35 s per LM iteration on one core (scalar LDLT of the 16,002^2 S)
    python tools/gen_bal_golden.py ladybug

The fixture holds the iteration records, status, best index and a fingerprint of the best values (strided samples per
key class + norms); tests/test_gpu_golden.py replays the same problem on the GPU and compares.  This script is test
infrastructure: it is the only writer of those fixtures and the only place that runs the oracle at Final-shape.
"""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

from symforce_b200 import desc as D, problems as P  # noqa: E402
from tests import oracle_capi as O  # noqa: E402


def fingerprint(values, prob):
    """Strided samples of the optimized storage + norms: a cheap, tolerance-comparable stand-in for a hash."""
    v = np.asarray(values)
    n = v.shape[0]
    idx = np.unique(np.linspace(0, n - 1, 4096).astype(np.int64))
    return {"n": int(n), "l2": float(np.linalg.norm(v)), "sum": float(v.sum()), "sample_idx": idx.tolist(),
            "sample": v[idx].tolist()}


def main():
    shape = sys.argv[1] if len(sys.argv) > 1 else "final"
    iters = int(sys.argv[2]) if len(sys.argv) > 2 else 50
    params = D.default_params()
    params.iterations = iters
    t0 = time.time()
    if shape == "pose_graph":
        # BASELINE.json configs[4]: 100k Pose3 poses + 20k loop closures, DefaultOptimizerParams (STATIC lambda)
        prob = P.pose_graph_problem(100000, 20000, params=params)
        return run(prob, "pose_graph_100k", {"n_poses": 100000, "n_loops": 20000}, {"seed": 0x51A4},
                   {"lambda_update_type": "STATIC", "iterations": iters, "rest": "DefaultOptimizerParams (optimizer.cc:8-56)"}, t0)
    params.lambda_update_type = D.LAMBDA_DYNAMIC
    prob = P.bal_problem(shape, solver=D.SOLVER_SCHUR, params=params)
    s = P.BAL_SHAPES[shape]
    return run(prob, f"{shape}_shape", {k: s[k] for k in ("n_cams", "n_pts", "n_obs", "window")},
               {"structure": 0xBA1, "noise": 0xBA2},
               {"lambda_update_type": "DYNAMIC", "iterations": iters, "rest": "DefaultOptimizerParams (optimizer.cc:8-56)"}, t0)


def run(prob, name, shape_desc, seeds, params_desc, t0):
    shape = name
    print(f"problem built in {time.time() - t0:.1f} s", flush=True)
    t0 = time.time()
    o = O.OracleProblem(prob)
    print(f"oracle created in {time.time() - t0:.1f} s", flush=True)
    t0 = time.time()
    st = o.optimize()
    wall = time.time() - t0
    its = o.iterations()
    tm = o.timings()
    best = o.best_values()
    out = {
        "generator": "tools/gen_bal_golden.py " + shape,
        "oracle": "oracle/oracle.cc (CPU restatement of levenberg_marquardt_solver.tcc:139-343 + SparseSchurSolver + LDLT)",
        "shape": shape_desc,
        "seeds": seeds,
        "params": params_desc,
        "status": int(st.status), "failure_reason": int(st.failure_reason), "best_index": int(st.best_index),
        "n_records": len(its),
        "records": [{"iteration": int(r.iteration), "current_lambda": float(r.current_lambda),
                     "new_error": float(r.new_error), "relative_reduction": float(r.relative_reduction),
                     "update_accepted": int(r.update_accepted)} for r in its],
        "final_error": float(its[st.best_index].new_error),
        "best_values": fingerprint(best, prob),
        "oracle_wall_s": wall,
        "oracle_timings": tm,
    }
    path = os.path.join(ROOT, "tests", "golden", f"{name}_history.json")
    with open(path, "w") as f:
        json.dump(out, f)
    print(f"wrote {path}: {len(its)} records, status {st.status}, best {st.best_index}, final error "
          f"{out['final_error']:.12g}, {wall:.0f} s", flush=True)


if __name__ == "__main__":
    main()
