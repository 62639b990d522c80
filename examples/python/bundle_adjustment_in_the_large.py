"""
Bundle Adjustment in the Large from Python: reads a BAL text file (https://grail.cs.washington.edu/projects/bal/) or
generates a BAL-shaped synthetic problem, and runs the sparse LM loop on the GPU through the C ABI.  Same problem
statement and optimizer parameters as the reference example
(symforce/examples/bundle_adjustment_in_the_large/bundle_adjustment_in_the_large.cc:123-140: DefaultOptimizerParams +
DYNAMIC lambda); the trailing points are eliminated with the Schur complement.

    python examples/python/bundle_adjustment_in_the_large.py problem-49-7776-pre.txt
    python examples/python/bundle_adjustment_in_the_large.py --synthetic ladybug        # or: final, small, tiny
"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))

from symforce_b200 import capi, desc as D, problems as P  # noqa: E402


def main():
    if len(sys.argv) == 3 and sys.argv[1] == "--synthetic":
        problem = P.bal_problem(sys.argv[2], solver=D.SOLVER_SCHUR)
    elif len(sys.argv) == 2:
        t0 = time.time()
        problem = P.read_bal(sys.argv[1], solver=D.SOLVER_SCHUR)
        print(f"read {sys.argv[1]} in {time.time() - t0:.2f} s")
    else:
        print(__doc__)
        return 2
    m = problem.meta
    print(f"Created problem with {m['n_cams']} cameras, {m['n_pts']} points, {m['n_obs']} observations")
    t0 = time.time()
    gpu = capi.SfxProblem(problem)  # structural analysis + upload, once per problem
    print(f"setup {time.time() - t0:.2f} s")
    stats = gpu.optimize()
    for it in gpu.iterations():
        print(f"[iter {it.iteration:4d}] lambda: {it.current_lambda:.3e}, error: {it.new_error:.9e}, "
              f"rel reduction: {it.relative_reduction:.5e}, accepted: {it.update_accepted}")
    tm = gpu.timings()
    print(f"status {stats.status}, best iteration record {stats.best_index}; device time {tm['total_ms']:.1f} ms for "
          f"{tm['iterations_run']} iterations (linearize {tm['linearize_ms']:.1f}, Schur {tm['schur_ms']:.1f}, "
          f"factorize {tm['factorize_ms']:.1f}, solve {tm['solve_ms']:.1f}, update {tm['update_ms']:.1f})")
    gpu.close()
    return 0


if __name__ == "__main__":
    sys.exit(main())
