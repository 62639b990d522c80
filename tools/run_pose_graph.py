#!/usr/bin/env python
"""Config E driver: synthetic Pose3 pose graph through the C ABI, optional oracle comparison."""
import argparse
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from symforce_b200 import capi, desc as D, problems as P  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--poses", type=int, default=100000)
ap.add_argument("--loops", type=int, default=20000)
ap.add_argument("--oracle", type=int, default=0)
ap.add_argument("--ordering", type=int, default=D.ORDERING_METIS_SCALAR)
args = ap.parse_args()
t0 = time.time()
prob = P.pose_graph_problem(args.poses, args.loops, ordering=args.ordering)
print(f"generated {prob.meta} in {time.time() - t0:.1f}s", flush=True)
t0 = time.time()
g = capi.SfxProblem(prob)
print(f"create {time.time() - t0:.2f}s info {g.info()}", flush=True)
g.optimize(2)  # warm-up
g.set_values(prob.values)
t0 = time.time()
st = g.optimize()
wall = time.time() - t0
its = g.iterations()
tm = g.timings()
print(f"status {st.status} iterations {len(its) - 1} initial {its[0].new_error:.6g} final {its[st.best_index].new_error:.9g} "
      f"wall {wall * 1e3:.1f} ms  device {tm['total_ms']:.1f} ms  per-iteration {tm['total_ms'] / max(tm['iterations_run'], 1):.2f} ms")
print({k: round(v, 3) if isinstance(v, float) else v for k, v in tm.items()})
if args.oracle:
    from tests import oracle_capi as O
    t0 = time.time()
    o = O.OracleProblem(prob)
    so = o.optimize()
    ito = o.iterations()
    print(f"oracle: status {so.status} iterations {len(ito) - 1} final {ito[so.best_index].new_error:.9g} "
          f"in {time.time() - t0:.1f}s; timings {o.timings()}")
    same = len(ito) == len(its) and all(abs(a.new_error - b.new_error) <= 1e-7 * abs(b.new_error) for a, b in zip(its, ito))
    print("PARITY", same)
