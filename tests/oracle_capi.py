"""
ctypes binding of oracle/_build/liboracle.so (CPU restatement of the reference; the CHECKER).
Only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs may import this module.
"""
import ctypes as C
import os
import subprocess

import numpy as np

from symforce_b200 import desc as D

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_LIB = os.path.join(ROOT, "oracle", "_build", "liboracle.so")
_REF = os.path.join(ROOT, "oracle", "_ref", "libref_factors.so")


def build():
    subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "oracle")])


def load():
    if not os.path.exists(_LIB):
        build()
    lib = C.CDLL(_LIB)
    lib.orc_last_error.restype = C.c_char_p
    return lib


def load_ref():
    """The reference's own generated factor headers compiled in place (oracle/_ref)."""
    if not os.path.exists(_REF):
        return None
    return C.CDLL(_REF)


class OracleProblem(D._LibProblem):
    prefix = "orc_"

    def __init__(self, problem: D.Problem):
        self.lib = load()
        self.problem = problem
        self.n_values = problem.values.shape[0]
        d, keep = problem.desc()
        self._keep = keep
        h = C.c_void_p()
        rc = self.lib.orc_create(C.byref(d), C.byref(h))
        if rc != 0:
            raise RuntimeError("orc_create failed: " + self.lib.orc_last_error().decode())
        self.h = h
        self.set_values(problem.values)

    def _last_error(self):
        return self.lib.orc_last_error().decode()

    def jacobian(self):
        """Linearization::jacobian at the values last set, as (outer, inner, values) of the M x N CSC matrix."""
        nnz = C.c_int64()
        self._check(self.lib.orc_linearize_jacobian(self.h, C.byref(nnz), None, None, None), "linearize_jacobian")
        N, _, _ = self.dims()
        outer = np.empty(N + 1, dtype=np.int32)
        inner = np.empty(nnz.value, dtype=np.int32)
        val = np.empty(nnz.value)
        self._check(self.lib.orc_linearize_jacobian(self.h, None, outer.ctypes.data_as(C.POINTER(C.c_int32)),
                                                    inner.ctypes.data_as(C.POINTER(C.c_int32)),
                                                    val.ctypes.data_as(C.POINTER(C.c_double))), "linearize_jacobian")
        return outer, inner, val

    def timings(self):
        out = (C.c_double * 8)()
        self.lib.orc_get_timings(self.h, out)
        names = ["setup_s", "linearize_s", "factorize_s", "solve_s", "total_s", "n_linearize", "n_factorize", "iters"]
        return dict(zip(names, list(out)))

    def reset_timings(self):
        self.lib.orc_reset_timings(self.h)

    def __del__(self):
        try:
            if getattr(self, "h", None):
                self.lib.orc_destroy(self.h)
                self.h = None
        except Exception:
            pass


def eval_factor(lib, fn, kind, args):
    """args: list of 1-D float64 arrays. Returns res, J (col-major RxT), H (TxT col-major), rhs."""
    meta = D.KINDS[kind]
    R, T = meta["res_dim"], meta["tan_dim"]
    arrs = [np.ascontiguousarray(a, dtype=np.float64) for a in args]
    ptrs = (C.POINTER(C.c_double) * len(arrs))(*[a.ctypes.data_as(C.POINTER(C.c_double)) for a in arrs])
    res = np.zeros(R)
    J = np.zeros(R * T)
    H = np.zeros(T * T)
    rhs = np.zeros(T)
    p = C.POINTER(C.c_double)
    rc = getattr(lib, fn)(C.c_int(kind), ptrs, res.ctypes.data_as(p), J.ctypes.data_as(p), H.ctypes.data_as(p),
                          rhs.ctypes.data_as(p))
    assert rc == 0
    return res, J.reshape(T, R).T, H.reshape(T, T).T, rhs
