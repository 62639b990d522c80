"""
Multi-GPU parity check, run under torchrun on N GPUs of one box:
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29511 tests/mgpu_check.py
Every rank solves its landmark shard; the iteration history and final values must match the
single-GPU solve of the same problem (which rank 0 also runs) to the north star's tolerances.
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np
import torch
import torch.distributed as dist

from symforce_b200 import capi, desc as D, problems as P

rank = int(os.environ["RANK"])
world = int(os.environ["WORLD_SIZE"])
local = int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
comm = capi.Comm(rank, world, local)
ok = True
for name in ["small", "ladybug"]:
    prob = P.bal_problem(name, solver=D.SOLVER_SCHUR)
    g = capi.SfxProblem(prob, device=local, rank=rank, world=world, comm=comm)
    st = g.optimize()
    its = g.iterations()
    vals = g.best_values()
    upd = np.array(prob.values, dtype=np.float64, copy=True)
    g.update_best_values(upd)  # Values::Update semantics must assemble the same buffer on every rank
    assert np.array_equal(upd, vals), "update_best_values != best_values on rank %d" % rank
    if rank == 0:
        g1 = capi.SfxProblem(prob, device=local)
        st1 = g1.optimize()
        its1 = g1.iterations()
        v1 = g1.best_values()
        same = (st.status == st1.status and len(its) == len(its1) and
                all(abs(a.new_error - b.new_error) <= 1e-8 * abs(b.new_error) and a.update_accepted == b.update_accepted
                    for a, b in zip(its, its1)) and np.allclose(vals, v1, rtol=1e-7, atol=1e-9))
        print(f"[mgpu] {name}: world={world} iterations={len(its) - 1} final={its[st.best_index].new_error:.9g} "
              f"single={its1[st1.best_index].new_error:.9g} match={same}", flush=True)
        ok = ok and same
        g1.close()
    g.close()
comm.close()
dist.destroy_process_group()
if rank == 0:
    print("MGPU_OK" if ok else "MGPU_FAIL", flush=True)
    sys.exit(0 if ok else 1)
