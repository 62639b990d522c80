"""
Discrete-event model of the persistent tile-DAG Cholesky kernel (large_factor_kernel, chol_large.cu) on ONE front:
CTAs claim tasks in list order (atomic counter), a CTA that claimed a task spins until the task's tiles have the versions
it needs (the rules of verify_task_list, sfx_api.cu), runs it for a duration taken from the measured task trace
(profiles/r01_factor_trace.npz) and publishes the new versions.  Host only; used to size scheduling changes before
they are written as CUDA.

    python tools/simulate_tile_dag.py [wt nt [workers]]        default: the Final-shape root front, 45 / 45 / 296
"""
import ctypes as C
import heapq
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from symforce_b200 import capi  # noqa: E402

# measured on the root front of Final-shape (us): see profiles/r01_results.md
DUR = dict(trsm=7.5, update=8.7, diag_potrf=14.0, diag_rest=9.4, inv=15.5, range_fixed=5.0, range_per_step=5.1, claim=0.4)


def task_list(wt, nt, kc):
    lib = capi.load()
    n = lib.sfx_debug_front_tasks(wt, nt, kc, None, 0)
    buf = np.zeros((n, 5), dtype=np.int16)
    lib.sfx_debug_front_tasks(wt, nt, kc, buf.ctypes.data_as(C.POINTER(C.c_int16)), n)
    return buf


def simulate(tasks, nt, workers, dur=DUR, fused_row2=False):
    """Returns (makespan, busy, waiting).  ready[(i, j)][v] = time tile (i, j) reached version v.
    fused_row2: TRSM(k+2,k) + UPDATE(k+2,k+1,k) + UPDATE(k+2,k+2,k) cost one load/store round instead of three."""
    INF = float("inf")
    ver_time = {}

    def t_of(i, j, v):  # time at which tile (i, j) has version >= v
        if v <= 0:
            return 0.0
        return ver_time.get((i, j, v), INF)

    free = [0.0] * workers
    heapq.heapify(free)
    busy = wait = 0.0
    end = 0.0
    for ty, k, i, j, k1 in tasks:
        ty, k, i, j, k1 = int(ty), int(k), int(i), int(j), int(k1)
        claim = heapq.heappop(free) + dur["claim"]
        if ty == 3:
            # POTRF starts as soon as tile (k, k) is final; the TRSM / SYRK parts wait for their own tiles
            start = max(claim, t_of(k, k, k))
            t_potrf = start + dur["diag_potrf"]
            ver_time[(k, k, k + 1)] = t_potrf
            done = t_potrf
            if k + 1 < nt:
                t_trsm = max(t_potrf, t_of(k + 1, k, k)) + 0.5 * dur["diag_rest"]
                ver_time[(k + 1, k, k + 1)] = t_trsm
                done = max(t_trsm, t_of(k + 1, k + 1, k)) + 0.5 * dur["diag_rest"]
                ver_time[(k + 1, k + 1, k + 1)] = done
        elif ty == 5:
            start = max(claim, t_of(k, k, k + 1))
            done = start + dur["inv"]
        elif ty == 1:
            start = max(claim, t_of(k, k, k + 1), t_of(i, k, k))
            d = dur["trsm"]
            done = start + d
            ver_time[(i, k, k + 1)] = done
        elif ty == 2:
            start = max(claim, t_of(i, k, k + 1), t_of(j, k, k + 1), t_of(i, j, k))
            d = dur["update"]
            if fused_row2 and i == k + 2:
                d *= 0.55  # operands stay in shared memory
            done = start + d
            ver_time[(i, j, k + 1)] = done
        elif ty == 4:
            # the range task takes its operand tiles one pivot step at a time and stalls on each as needed
            start = max(claim, t_of(i, j, k))
            tcur = start + 0.5 * dur["range_fixed"]
            for kk in range(k, k1):
                tcur = max(tcur, t_of(i, kk, kk + 1), t_of(j, kk, kk + 1)) + dur["range_per_step"]
            done = tcur + 0.5 * dur["range_fixed"]
            ver_time[(i, j, k1)] = done
        else:
            raise ValueError(ty)
        assert start < INF, ("task waits for a version nobody produces", ty, k, i, j, k1)
        busy += done - start
        wait += start - claim
        end = max(end, done)
        heapq.heappush(free, done)
    return end, busy, wait


def main():
    wt, nt = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (45, 45)
    workers = int(sys.argv[3]) if len(sys.argv) > 3 else 296
    tasks = task_list(wt, nt, 6)
    print(f"front wt={wt} nt={nt}: {len(tasks)} tasks, {workers} CTAs")
    base, busy, wait = simulate(tasks, nt, workers)
    print(f"  model of today's kernel      : span {base:7.0f} us   busy {busy / 1e3:6.1f} ms  waiting {wait / 1e3:6.1f} ms")
    for name, d, fused in [
        ("TRSM/SYRK of row k+1 off the chain (band-continuous diagonal task)", dict(DUR, diag_rest=0.8), False),
        ("  + fused row-(k+2) priority tasks", dict(DUR, diag_rest=0.8, trsm=5.0), True),
        ("  + POTRF 14.0 -> 10.0 us", dict(DUR, diag_rest=0.8, trsm=5.0, diag_potrf=10.0), True),
        ("RANGE tasks 20 % faster only", dict(DUR, range_per_step=4.1), False),
    ]:
        span, b, w = simulate(tasks, nt, workers, d, fused)
        print(f"  {name:<68s}: span {span:7.0f} us ({100 * span / base:5.1f} %)")


if __name__ == "__main__":
    main()
