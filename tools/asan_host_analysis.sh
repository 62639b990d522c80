#!/bin/bash
# Host-side structural analysis (analysis.cc, symbolic.cc, debug.cc: plain C++) under AddressSanitizer + UBSan, on the
# problem set of the CPU tests plus Final-shape; no GPU needed.   bash tools/asan_host_analysis.sh
set -e
cd "$(dirname "$0")/.."
g++ -std=c++17 -O1 -g -fsanitize=address,undefined -fPIC -shared -I/usr/local/cuda/include -o /tmp/libsfx_asan.so \
  symforce_b200/csrc/analysis.cc symforce_b200/csrc/symbolic.cc symforce_b200/csrc/debug.cc \
  /usr/local/cuda/targets/x86_64-linux/lib/libmetis_static.a -lpthread
cat > /tmp/asan_run.py <<'PY'
import ctypes as C, sys
sys.path.insert(0, '.')
from symforce_b200 import desc as D, problems as P
lib = C.CDLL('/tmp/libsfx_asan.so')
def run(prob, world=1):
    for rank in range(world):
        d, keep = prob.desc(rank=rank, world=world, comm=(1 if world > 1 else None))
        out = C.c_char_p()
        assert lib.sfx_debug_analysis_json(C.byref(d), C.byref(out)) == 0, out.value
        assert lib.sfx_debug_front_summary(C.byref(d), C.byref(out)) == 0, out.value
for name, prob in [("tiny", P.bal_problem("tiny", solver=D.SOLVER_SCHUR)), ("ladybug", P.bal_problem("ladybug", solver=D.SOLVER_SCHUR)),
                   ("robot3d", P.robot_3d_localization()), ("ba_example", P.ba_example()), ("pose_graph", P.pose_graph_problem(300, 60)),
                   ("mid", P.bal_problem(n_cams=64, n_pts=20000, n_obs=90000, window=8, solver=D.SOLVER_SCHUR))]:
    run(prob)
    if name in ("ladybug", "mid"):
        run(prob, world=2)
    print("ok", name, flush=True)
if len(sys.argv) > 1 and sys.argv[1] == "final":
    d, keep = P.bal_problem("final", solver=D.SOLVER_SCHUR).desc()
    out = C.c_char_p()
    assert lib.sfx_debug_front_summary(C.byref(d), C.byref(out)) == 0
    print("ok final", out.value.decode().split("\n")[0])
PY
LD_PRELOAD=$(gcc -print-file-name=libasan.so):$(gcc -print-file-name=libubsan.so) ASAN_OPTIONS=detect_leaks=0 SFX_HOST_THREADS=4 \
  python /tmp/asan_run.py "$@"
