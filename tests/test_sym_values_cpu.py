"""
Host-side sym::Values / Rot3 / Pose3 of include/sym/sym.h (symforce/opt/values.h:31-324, values.cc:45-350; rows a7/a8 of
SURVEY.md section 8) without a GPU: tests/cpp/values_api.cc is compiled against the header and its output is recomputed
here with symforce_b200/geo.py, which is itself pinned to the reference's numeric package (tests/golden/geo_vectors.json).
"""
import os
import subprocess

import numpy as np

from symforce_b200.geo import K_DEFAULT_EPSILON as EPS, Pose3, Rot3

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(tmp_path):
    exe = str(tmp_path / "values_api")
    subprocess.check_call(["g++", "-std=c++17", "-O1", "-I" + os.path.join(ROOT, "include"), "-o", exe,
                           os.path.join(ROOT, "tests", "cpp", "values_api.cc"),
                           "-L" + os.path.join(ROOT, "symforce_b200", "lib"), "-lsfx",
                           "-Wl,-rpath," + os.path.join(ROOT, "symforce_b200", "lib")])
    out = subprocess.run([exe], capture_output=True, text=True, timeout=60)
    assert out.returncode == 0, out.stderr
    rec = {}
    for line in out.stdout.splitlines():
        name, *rest = line.split()
        rec[name] = rest
    return rec


def _f(rec, name):
    return np.array([float(x) for x in rec[name]])


def test_values_api_against_geo(tmp_path):
    rec = _run(tmp_path)
    a = Pose3(Rot3.from_tangent([0.3, -0.2, 0.5], EPS), [1.0, 2.0, 3.0])
    b = Pose3(Rot3.from_tangent([-0.7, 0.1, 0.9], EPS), [-0.5, 0.25, 4.0])
    tol = dict(rtol=0, atol=5e-15)
    assert np.allclose(_f(rec, "a"), a.data, **tol) and np.allclose(_f(rec, "b"), b.data, **tol)
    assert np.allclose(_f(rec, "a_inv"), a.inverse().data, **tol)
    assert np.allclose(_f(rec, "ab"), (a * b).data, **tol)
    assert np.allclose(_f(rec, "a_local_b"), a.local_coordinates(b, EPS), **tol)
    assert np.allclose(_f(rec, "rot_tangent"), [-0.7, 0.1, 0.9], rtol=0, atol=1e-15)

    # layout: p (7) | r (4) | x_1 (3) | s (1); tangent 6 + 3 + 3 + 1
    assert rec["keys"] == ["p,r,x_1,s,"] and rec["index"] == ["15", "13", "4"]
    v = np.concatenate([a.data, b.R.data, [1.0, -1.0, 0.5], [2.5]])
    w = np.concatenate([b.data, a.R.data, [0.0, 4.0, 0.25], [-1.0]])
    delta = np.concatenate([a.local_coordinates(b, EPS), b.R.local_coordinates(a.R, EPS), w[11:14] - v[11:14], [-3.5]])
    assert np.allclose(_f(rec, "delta"), delta, **tol)
    assert np.allclose(_f(rec, "target"), w, **tol)
    assert np.allclose(_f(rec, "retracted"), w, rtol=0, atol=1e-14)  # retract(local_coordinates) lands on the target
    upd = v.copy()
    upd[7:11] = w[7:11]
    upd[14] = w[14]
    assert np.allclose(_f(rec, "updated"), upd, **tol)
    # Remove('r') + Cleanup(): 4 scalars freed, the rest compacted in order; UpdateOrSet overwrites s and appends r
    assert rec["remove"] == ["1", "0", "4", "11", "0"]
    assert np.allclose(_f(rec, "compacted"), np.concatenate([v[:7], v[11:]]), **tol)
    assert np.allclose(_f(rec, "update_or_set"), np.concatenate([v[:7], v[11:14], [-1.0], w[7:11]]), **tol)
    assert rec["entry"] == ["1"] and _f(rec, "at_entry").tolist() == [1.0, -1.0, 0.5]
    assert rec["set_entry"] == ["7"] and rec["misc"] == ["1", "1"]
    # Factor::Jacobian: same kind and optimized keys as Factor::Hessian, 5 keys / 2 optimized, lambdas and reordered
    # keys_to_optimize rejected
    assert rec["factor_jacobian"] == ["1", "1", "5", "2", "1", "1"]


def test_cpp_callers_build_and_refuse_to_run_without_a_gpu():
    """Every C++17 caller of include/sym/sym.h under examples/ compiles and links against libsfx.so on a CPU-only box
    (all entry points the header layer binds are exported), and on a box without a CUDA device the optimizer throws
    instead of computing anything on the host: the product path has no CPU fallback."""
    import torch

    subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "examples")])
    for name in ("robot_3d_localization", "bundle_adjustment_in_the_large", "covariance_check", "bundle_adjustment", "gnc_test"):
        assert os.path.exists(os.path.join(ROOT, "examples", "_build", name)), name
    if torch.cuda.is_available():
        return
    out = subprocess.run([os.path.join(ROOT, "examples", "_build", "robot_3d_localization")], capture_output=True, text=True,
                         timeout=60)
    assert out.returncode != 0
    assert "no CUDA device: libsfx has no CPU fallback" in out.stderr
