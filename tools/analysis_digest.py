"""
Digests of the host analysis (sfx_debug_analysis_digest) for a set of problems, to check that a change of
symforce_b200/csrc/analysis.cc leaves every index array identical on problems too large for the JSON replay of
tests/test_host_analysis.py.  No GPU needed.

    python tools/analysis_digest.py out.txt [--lib other_libsfx.so] [--small]
    SFX_HOST_THREADS=1 python tools/analysis_digest.py a.txt; python tools/analysis_digest.py b.txt; diff a.txt b.txt
"""
import ctypes as C
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    out_path = sys.argv[1]
    if "--lib" in sys.argv:
        os.environ["SFX_LIB"] = sys.argv[sys.argv.index("--lib") + 1]
    from symforce_b200 import capi, desc as D, problems as P

    lib = capi.load()
    cases = [
        ("tiny", lambda: P.bal_problem("tiny"), [(0, 1), (0, 2), (1, 2)]),
        ("small", lambda: P.bal_problem("small"), [(0, 1), (1, 3)]),
        ("small_chol", lambda: P.bal_problem("small", solver=D.SOLVER_CHOLESKY), [(0, 1)]),
        ("robot3d", lambda: P.robot_3d_localization(), [(0, 1)]),
        ("ba_example", lambda: P.ba_example(), [(0, 1)]),
        ("pose_graph_2k", lambda: P.pose_graph_problem(n_poses=2000, n_loops=400), [(0, 1)]),
        ("ladybug", lambda: P.bal_problem("ladybug"), [(0, 1), (0, 2), (1, 2), (3, 8)]),
    ]
    if "--small" not in sys.argv:
        cases += [
            ("pose_graph_100k", lambda: P.pose_graph_problem(), [(0, 1)]),
            ("final", lambda: P.bal_problem("final"), [(0, 1), (1, 2), (5, 8)]),
        ]
    with open(out_path, "w") as f:
        for name, make, shards in cases:
            prob = make()
            for rank, world in shards:
                d, keep = prob.desc(rank=rank, world=world, comm=(1 if world > 1 else None))
                out = C.c_char_p()
                t = time.time()
                rc = lib.sfx_debug_analysis_digest(C.byref(d), C.byref(out))
                dt = time.time() - t
                print(f"{name} rank {rank}/{world}: rc={rc} {dt:.2f} s", flush=True)
                f.write(f"== {name} rank {rank}/{world} rc={rc}\n{out.value.decode()}\n")


if __name__ == "__main__":
    main()
