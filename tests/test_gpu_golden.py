"""
Parity at the benchmarked configuration and at N > 1 (run on the B200 box).

* Golden histories: tests/golden/{ladybug,final}_shape_history.json were written by tools/gen_bal_golden.py from the
  CPU oracle (reference loop: symforce/opt/levenberg_marquardt_solver.tcc:139-343) with the reference's BAL parameters
  (bundle_adjustment_in_the_large.cc:133-136).  The GPU must reproduce them: same status, same number of records, same
  accept/reject sequence, every record's error within 1e-8 relative (the north star's final-cost tolerance), lambdas within
  a tolerance that follows the conditioning of the gain ratio (see _compare), best values within the fingerprint tolerance.
* Multi-GPU: tests/mgpu_check.py under torchrun on 2 GPUs when the box has them (iteration-history identity with 1 GPU).
"""
import json
import os
import subprocess
import sys

import numpy as np
import pytest

from symforce_b200 import capi, desc as D, problems as P

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
COST_TOL = 1e-8  # north star: final cost within 1e-8 relative, same iteration count


def _replay(shape):
    path = os.path.join(ROOT, "tests", "golden", f"{shape}_shape_history.json")
    if not os.path.exists(path):
        pytest.skip(f"{path} not generated (tools/gen_bal_golden.py {shape})")
    gold = json.load(open(path))
    s = P.BAL_SHAPES[shape]
    assert {k: s[k] for k in ("n_cams", "n_pts", "n_obs", "window")} == gold["shape"]
    params = D.default_params()
    params.lambda_update_type = D.LAMBDA_DYNAMIC
    params.iterations = gold["params"]["iterations"]
    prob = P.bal_problem(shape, solver=D.SOLVER_SCHUR, params=params)
    gpu = capi.SfxProblem(prob, device=0)
    st = gpu.optimize()
    its = gpu.iterations()
    best = gpu.best_values()
    gpu.close()
    return gold, st, its, best


def _compare(gold, st, its, best, cost_tol=COST_TOL, values_tol=1e-5):
    assert st.status == gold["status"], (st.status, gold["status"])
    assert len(its) == gold["n_records"], (len(its), gold["n_records"])  # same iteration count
    assert st.best_index == gold["best_index"]
    worst = 0.0
    # lambda is a derived control variable: the DYNAMIC update multiplies it by a function of the gain ratio
    # (actual / predicted reduction), whose relative error is (error noise) / (relative reduction) -- 1e-10 / 1e-6 near
    # convergence -- and the factors compound from one iteration to the next
    lam_tol = 1e-6
    for r, g in zip(its, gold["records"]):
        if g["iteration"] >= 0:
            lam_tol += 2e-9 / max(abs(g["relative_reduction"]), 1e-9)
        assert r.iteration == g["iteration"]
        assert r.update_accepted == g["update_accepted"], (r.iteration, "accept/reject differs")
        rel = abs(r.new_error - g["new_error"]) / abs(g["new_error"])
        worst = max(worst, rel)
        assert rel <= cost_tol, (r.iteration, r.new_error, g["new_error"], rel)
        assert abs(r.current_lambda - g["current_lambda"]) <= lam_tol * abs(g["current_lambda"]), (
            r.iteration, r.current_lambda, g["current_lambda"], lam_tol)
    final = its[st.best_index].new_error
    assert abs(final - gold["final_error"]) <= cost_tol * abs(gold["final_error"]), (final, gold["final_error"])
    fp = gold["best_values"]
    idx = np.array(fp["sample_idx"])
    want = np.array(fp["sample"])
    assert best.shape[0] == fp["n"]
    # best values: the optimum is flat along the gauge directions, so storage agrees less tightly than the cost
    assert np.max(np.abs(best[idx] - want)) <= values_tol * max(1.0, np.max(np.abs(want)))
    assert abs(np.linalg.norm(best) - fp["l2"]) <= 1e-7 * fp["l2"]
    return worst


def test_ladybug_shape_history_matches_the_oracle_fixture():
    gold, st, its, best = _replay("ladybug")
    worst = _compare(gold, st, its, best)
    print(f"ladybug: {len(its) - 1} iterations, worst relative error deviation {worst:.2e}")


def test_final_shape_history_matches_the_oracle_fixture():
    """The north star's acceptance line: the 1,778 / 993,923 / 5,001,946 problem converges to the CPU oracle's cost
    (1e-8 relative) with the same iteration count."""
    gold, st, its, best = _replay("final")
    worst = _compare(gold, st, its, best)
    print(f"final: {len(its) - 1} iterations, worst relative error deviation {worst:.2e}")


def test_pose_graph_100k_history_matches_the_oracle_fixture():
    """BASELINE.json configs[4] (SparseCholeskySolver path, 1 GPU): the first 10 LM iterations of the 100k-pose graph
    -- two of them rejected -- against the oracle fixture.  The costs fall from 4.6e11 to 3.2e8 through steps with large
    rotations; iteration costs agree to 1e-7 relative (measured ~1e-9), the tolerance round 1 established for this path."""
    path = os.path.join(ROOT, "tests", "golden", "pose_graph_100k_history.json")
    gold = json.load(open(path))
    params = D.default_params()
    params.iterations = gold["params"]["iterations"]
    prob = P.pose_graph_problem(gold["shape"]["n_poses"], gold["shape"]["n_loops"], params=params)
    gpu = capi.SfxProblem(prob, device=0)
    st = gpu.optimize()
    its = gpu.iterations()
    best = gpu.best_values()
    gpu.close()
    worst = _compare(gold, st, its, best, cost_tol=1e-7, values_tol=1e-4)
    print(f"pose graph: {len(its) - 1} iterations, worst relative error deviation {worst:.2e}")


def _n_gpus():
    try:
        import torch

        return torch.cuda.device_count()
    except Exception:
        return 0


@pytest.mark.parametrize("world", [2])
def test_sharded_solve_matches_single_gpu(world):
    """landmarks + observations sharded over `world` ranks, NCCL reduce of S; the iteration history and
    the assembled best values must be those of the single-GPU solve (tests/mgpu_check.py does the comparison)."""
    if _n_gpus() < world:
        pytest.skip(f"needs {world} GPUs on the box")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
           "--master-addr", "127.0.0.1", "--master-port", "29533", os.path.join(ROOT, "tests", "mgpu_check.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=ROOT)
    sys.stdout.write(r.stdout[-4000:])
    assert r.returncode == 0 and "MGPU_OK" in r.stdout, r.stdout[-2000:] + r.stderr[-4000:]
