"""
ctypes binding of symforce_b200/lib/libsfx.so -- the product: CUDA (sm_100a) sparse LM behind the
C ABI of include/sfx.h.  There is NO CPU fallback: if the library is missing this module raises,
and without a CUDA device sfx_problem_create fails with SFX_ERR_CUDA.
"""
import ctypes as C
import json
import os

import numpy as np

from . import desc as D

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("SFX_LIB") or os.path.join(_HERE, "lib", "libsfx.so")  # SFX_LIB: experiment builds

EXPORTS = [
    "sfx_default_params", "sfx_problem_create", "sfx_problem_destroy", "sfx_last_error", "sfx_update_params",
    "sfx_set_values", "sfx_optimize", "sfx_optimize_continue", "sfx_relax_damping_to_initial", "sfx_get_best_values",
    "sfx_update_best_values", "sfx_get_iteration_debug", "sfx_get_iteration_update", "sfx_get_iteration_jacobian",
    "sfx_get_iterations", "sfx_get_dims",
    "sfx_get_hessian_pattern", "sfx_linearize", "sfx_get_jacobian_pattern", "sfx_linearize_jacobian", "sfx_check_derivatives",
    "sfx_get_best_linearization", "sfx_solve_step",
    "sfx_compute_covariance", "sfx_get_ordering", "sfx_get_timings", "sfx_get_info", "sfx_comm_unique_id", "sfx_comm_create",
    "sfx_comm_destroy",
]

INFO_NAMES = ["N", "M", "nnz_H", "num_nodes", "reduced_dim", "nnz_L", "num_supernodes", "num_levels",
              "factor_flops", "s_blocks", "schur_pairs", "max_front", "device_bytes", "chol_failures",
              "nonfinite_updates", "zero_diagonal", "plan", "ref_ordering_flops"]

_lib = None


def load():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "(symforce_b200 has no CPU fallback)")
        _lib = C.CDLL(os.environ.get("SFX_LIB", LIB_PATH))  # SFX_LIB: A/B runs against another build
        _lib.sfx_last_error.restype = C.c_char_p
        _lib.sfx_last_error.argtypes = [C.c_void_p]
    return _lib


def analysis_json(problem: D.Problem, rank=0, world=1):
    """Host-only structural analysis (no CUDA), for the CPU test-suite.  world > 1 analyses the
    shard of `rank` (a dummy non-NULL communicator pointer enables sharding; it is never used)."""
    lib = load()
    d, keep = problem.desc(rank=rank, world=world, comm=(1 if world > 1 else None))
    out = C.c_char_p()
    rc = lib.sfx_debug_analysis_json(C.byref(d), C.byref(out))
    if rc != 0:
        raise RuntimeError("analysis failed: " + out.value.decode())
    return json.loads(out.value.decode())


class Comm:
    """One NCCL communicator per rank (sfx_comm_*); the unique id travels over torch.distributed."""

    def __init__(self, rank, world, device):
        import torch
        import torch.distributed as dist

        self.lib = load()
        ident = C.create_string_buffer(128)
        if rank == 0:
            rc = self.lib.sfx_comm_unique_id(ident)
            if rc != 0:
                raise RuntimeError("sfx_comm_unique_id failed: " + self.lib.sfx_last_error(None).decode())
        t = torch.tensor(list(ident.raw), dtype=torch.uint8, device=f"cuda:{device}")
        dist.broadcast(t, src=0)
        ident = C.create_string_buffer(bytes(t.cpu().tolist()), 128)
        h = C.c_void_p()
        rc = self.lib.sfx_comm_create(ident, C.c_int32(rank), C.c_int32(world), C.c_int32(device), C.byref(h))
        if rc != 0:
            raise RuntimeError("sfx_comm_create failed: " + self.lib.sfx_last_error(None).decode())
        self.h = h
        self.rank, self.world, self.device = rank, world, device

    def close(self):
        if getattr(self, "h", None):
            self.lib.sfx_comm_destroy(self.h)
            self.h = None


class SfxProblem(D._LibProblem):
    prefix = "sfx_"

    def __init__(self, problem: D.Problem, device=0, rank=0, world=1, comm=None):
        self.lib = load()
        self.problem = problem
        self.n_values = problem.values.shape[0]
        d, keep = problem.desc(device=device, rank=rank, world=world, comm=(comm.h if comm is not None else None))
        self._keep = keep
        h = C.c_void_p()
        rc = self.lib.sfx_problem_create(C.byref(d), C.byref(h))
        if rc != 0:
            raise RuntimeError(f"sfx_problem_create failed (rc={rc}): " + self.lib.sfx_last_error(None).decode())
        self.h = h
        self.set_values(problem.values)

    def _last_error(self):
        return self.lib.sfx_last_error(self.h).decode()

    def timings(self):
        t = D.Timings()
        self._check(self.lib.sfx_get_timings(self.h, C.byref(t)), "get_timings")
        return {f[0]: getattr(t, f[0]) for f in D.Timings._fields_}

    def info(self):
        out = (C.c_int64 * len(INFO_NAMES))()
        self._check(self.lib.sfx_get_info(self.h, out, C.c_int32(len(INFO_NAMES))), "get_info")
        return dict(zip(INFO_NAMES, list(out)))

    def update_best_values(self, values):
        """Values::Update with the best values: overwrites the optimized keys' storage in `values` (the buffer that was
        given to set_values); returns the number of bytes moved device -> host."""
        assert values.dtype == np.float64 and values.flags["C_CONTIGUOUS"] and values.shape[0] == self.n_values
        nb = C.c_int64(0)
        self._check(self.lib.sfx_update_best_values(self.h, values.ctypes.data_as(C.POINTER(C.c_double)),
                                                    C.c_int64(values.shape[0]), C.byref(nb)), "update_best_values")
        return nb.value

    def iteration_debug(self, record):
        """debug_stats payload of an iteration record: (values data, residual)."""
        _, M, _ = self.dims()
        v = np.empty(self.n_values)
        r = np.empty(M)
        p = C.POINTER(C.c_double)
        self._check(self.lib.sfx_get_iteration_debug(self.h, C.c_int32(record), v.ctypes.data_as(p), r.ctypes.data_as(p)),
                    "get_iteration_debug")
        return v, r

    def iteration_update(self, record):
        """debug_stats: optimization_iteration_t::update of a record (reference tangent order; zeros for record 0)."""
        N, _, _ = self.dims()
        u = np.empty(N)
        self._check(self.lib.sfx_get_iteration_update(self.h, C.c_int32(record), u.ctypes.data_as(C.POINTER(C.c_double))),
                    "get_iteration_update")
        return u

    def iteration_jacobian(self, record):
        """debug_stats: optimization_iteration_t::jacobian_values of a record, in the CSC order of jacobian()."""
        nnz = C.c_int64()
        self._check(self.lib.sfx_get_jacobian_pattern(self.h, C.byref(nnz), None, None), "get_jacobian_pattern")
        val = np.empty(nnz.value)
        self._check(self.lib.sfx_get_iteration_jacobian(self.h, C.c_int32(record), val.ctypes.data_as(C.POINTER(C.c_double))),
                    "get_iteration_jacobian")
        return val

    def jacobian_pattern(self):
        """CSC pattern of Linearization::jacobian: (column pointers [N + 1], row indices [nnz])."""
        nnz = C.c_int64()
        self._check(self.lib.sfx_get_jacobian_pattern(self.h, C.byref(nnz), None, None), "get_jacobian_pattern")
        N, _, _ = self.dims()
        outer = np.empty(N + 1, dtype=np.int32)
        inner = np.empty(nnz.value, dtype=np.int32)
        pi = C.POINTER(C.c_int32)
        self._check(self.lib.sfx_get_jacobian_pattern(self.h, None, outer.ctypes.data_as(pi), inner.ctypes.data_as(pi)),
                    "get_jacobian_pattern")
        return outer, inner

    def jacobian(self):
        """Linearization::jacobian (include_jacobians) at the values last set: (outer, inner, values) of the M x N CSC."""
        outer, inner = self.jacobian_pattern()
        val = np.empty(inner.shape[0])
        self._check(self.lib.sfx_linearize_jacobian(self.h, val.ctypes.data_as(C.POINTER(C.c_double))),
                    "linearize_jacobian")
        return outer, inner, val

    def check_derivatives(self, want_numerical_jacobian=False):
        """internal::CheckDerivatives (derivative_checker.h:32-123) at the values last set: (ok, {jacobian, hessian, rhs
        relative errors}[, dense M x N numerical Jacobian]).  Resets the optimizer state like linearize()."""
        N, M, _ = self.dims()
        err = (C.c_double * 3)()
        ok = C.c_int32(0)
        nj = np.empty((M, N), order="F") if want_numerical_jacobian else None
        self._check(self.lib.sfx_check_derivatives(self.h, err, C.byref(ok),
                                                   nj.ctypes.data_as(C.POINTER(C.c_double)) if nj is not None else None),
                    "check_derivatives")
        out = (bool(ok.value), {"jacobian": err[0], "hessian": err[1], "rhs": err[2]})
        return out + (nj,) if want_numerical_jacobian else out

    def compute_covariance(self, block_dim, hessian_values=None):
        """Optimizer::ComputeCovariances / ComputeFullCovariance: dense block_dim x block_dim covariance in keys_
        order, from the given Linearization::hessian_lower values (CSC order) or the best linearization."""
        cov = np.empty((block_dim, block_dim), dtype=np.float64, order="F")
        hv = None
        if hessian_values is not None:
            hessian_values = np.ascontiguousarray(hessian_values, dtype=np.float64)
            hv = hessian_values.ctypes.data_as(C.POINTER(C.c_double))
        self._check(self.lib.sfx_compute_covariance(self.h, hv, C.c_int32(block_dim),
                                                    cov.ctypes.data_as(C.POINTER(C.c_double))), "compute_covariance")
        return cov

    def close(self):
        if getattr(self, "h", None):
            self.lib.sfx_problem_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
