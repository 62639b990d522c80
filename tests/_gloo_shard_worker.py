"""Worker of test_two_process_gloo_allreduce_of_schur_contributions (gloo backend, CPU only)."""
import sys

import numpy as np
import torch
import torch.distributed as dist

from symforce_b200 import capi, desc as D, problems as P
from tests import host_emulation as E
from tests import oracle_capi as O

rank, world = int(sys.argv[1]), int(sys.argv[2])
dist.init_process_group("gloo", rank=rank, world_size=world)
prob = P.bal_problem("tiny", solver=D.SOLVER_SCHUR)
lam = 0.25
A = capi.analysis_json(prob, rank=rank, world=world)
res, rhs, Hv = E.emulate_linearize(prob, A)
red = A["schur_plan"]["reduced_dim"]
b_end = max(A["diag_pos"][:red]) + 1


def allreduce(x):
    t = torch.from_numpy(np.ascontiguousarray(x))
    dist.all_reduce(t)
    return t.numpy()


Hv[:b_end] = allreduce(Hv[:b_end])
rhs[:red] = allreduce(rhs[:red])
err = allreduce(np.array([0.5 * np.sum(res * res)]))[0]
dvec = E.damping_vector(A, Hv, lam, prob.params)
if rank == 0:
    Sv, rhs_red, cinv, tl = E.emulate_schur(A, Hv, rhs, dvec)
else:
    _, _, cinv, tl = E.emulate_schur(A, Hv, rhs, dvec)
    Hz = Hv.copy()
    Hz[:b_end] = 0
    Sv, rhs_red, _, _ = E.emulate_schur(A, Hz, np.concatenate([np.zeros(red), rhs[red:]]),
                                        np.concatenate([np.zeros(red), dvec[red:]]))
Sv = allreduce(Sv)
rhs_red = allreduce(rhs_red)
y = E.emulate_fronts_solve(A, Sv, rhs_red)
u = E.emulate_schur_back(A, Hv, cinv, tl, y)
upd = np.zeros(A["N"])
sp = A["schur_plan"]
if rank == 0:
    upd[:red] = u[:red]
for l in range(sp["n_landmarks"]):
    to, d = sp["lm_toff"][l], sp["lm_dim"][l]
    upd[to:to + d] = u[to:to + d]
upd = allreduce(upd)
if rank == 0:
    o = O.OracleProblem(prob)
    assert abs(err - o.linearize()[0] @ o.linearize()[0] * 0.5) < 1e-9 * err
    upd_o = o.solve_step(lam)
    assert np.allclose(upd[np.array(A["ref2int"])], upd_o, rtol=1e-8, atol=1e-10 * np.abs(upd_o).max())
    print("step matches oracle")
dist.destroy_process_group()
