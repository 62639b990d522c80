"""
ctypes mirror of include/sfx.h (structs + enums) and a small builder that lowers a flat
"values + keys + factor batches" problem into an `sfx_problem_desc`.

This is the plumbing underneath symforce_b200/opt.py (the Python `Optimizer`), the tests and bench.py; the drop-in for
C++ callers is the `sym::` header layer under include/sym/ which produces the same descriptor.
"""
import ctypes as C
import json
import os

import numpy as np

SFX_ABI_VERSION = 1

# sfx_type
TYPE_VECTOR, TYPE_ROT3, TYPE_POSE3 = 0, 1, 2
# sfx_factor_kind
(
    KIND_SNAVELY,
    KIND_BETWEEN_POSE3,
    KIND_PRIOR_POSE3,
    KIND_MATCHING,
    KIND_ODOMETRY,
    KIND_IRL_LINEAR_GNC,
    KIND_IRL_PRIOR,
    KIND_BETWEEN_ROT3,
    KIND_PRIOR_ROT3,
    KIND_BARRON,
) = range(10)
SOLVER_CHOLESKY, SOLVER_SCHUR = 0, 1
ORDERING_METIS_SCALAR, ORDERING_METIS_BLOCK, ORDERING_NATURAL = 0, 1, 2
STATUS_SUCCESS, STATUS_HIT_ITERATION_LIMIT, STATUS_FAILED = 1, 2, 3
LAMBDA_STATIC, LAMBDA_DYNAMIC = 1, 2

_HERE = os.path.dirname(os.path.abspath(__file__))
with open(os.path.join(_HERE, "kinds.json")) as _f:
    KINDS = json.load(_f)


class Params(C.Structure):
    _fields_ = [
        ("verbose", C.c_int32),
        ("debug_stats", C.c_int32),
        ("check_derivatives", C.c_int32),
        ("include_jacobians", C.c_int32),
        ("debug_checks", C.c_int32),
        ("initial_lambda", C.c_double),
        ("lambda_lower_bound", C.c_double),
        ("lambda_upper_bound", C.c_double),
        ("lambda_update_type", C.c_int32),
        ("lambda_up_factor", C.c_double),
        ("lambda_down_factor", C.c_double),
        ("dynamic_lambda_update_beta", C.c_double),
        ("dynamic_lambda_update_gamma", C.c_double),
        ("dynamic_lambda_update_p", C.c_int32),
        ("use_diagonal_damping", C.c_int32),
        ("use_unit_damping", C.c_int32),
        ("keep_max_diagonal_damping", C.c_int32),
        ("diagonal_damping_min", C.c_double),
        ("iterations", C.c_int32),
        ("early_exit_min_reduction", C.c_double),
        ("early_exit_min_absolute_error", C.c_double),
        ("enable_bold_updates", C.c_int32),
    ]


def default_params() -> Params:
    """sym::DefaultOptimizerParams() -- symforce/opt/optimizer.cc:8-56"""
    return Params(
        verbose=0,
        debug_stats=0,
        check_derivatives=0,
        include_jacobians=0,
        debug_checks=0,
        initial_lambda=1.0,
        lambda_lower_bound=0.0,
        lambda_upper_bound=1000000.0,
        lambda_update_type=LAMBDA_STATIC,
        lambda_up_factor=4.0,
        lambda_down_factor=0.25,
        dynamic_lambda_update_beta=2.0,
        dynamic_lambda_update_gamma=3.0,
        dynamic_lambda_update_p=3,
        use_diagonal_damping=0,
        use_unit_damping=1,
        keep_max_diagonal_damping=0,
        diagonal_damping_min=1e-6,
        iterations=50,
        early_exit_min_reduction=1e-6,
        early_exit_min_absolute_error=0.0,
        enable_bold_updates=0,
    )


class KeyEntry(C.Structure):
    _fields_ = [
        ("type", C.c_int32),
        ("offset", C.c_int32),
        ("storage_dim", C.c_int32),
        ("tangent_dim", C.c_int32),
    ]


class FactorBatch(C.Structure):
    _fields_ = [
        ("kind", C.c_int32),
        ("n", C.c_int32),
        ("arg_offsets", C.POINTER(C.c_int32)),
        ("opt_keys", C.POINTER(C.c_int32)),
        ("factor_index", C.POINTER(C.c_int32)),
    ]


class ProblemDesc(C.Structure):
    _fields_ = [
        ("abi_version", C.c_int32),
        ("params", Params),
        ("epsilon", C.c_double),
        ("n_values", C.c_int64),
        ("n_keys", C.c_int32),
        ("keys", C.POINTER(KeyEntry)),
        ("n_batches", C.c_int32),
        ("batches", C.POINTER(FactorBatch)),
        ("n_factors", C.c_int32),
        ("solver", C.c_int32),
        ("schur_num_keys", C.c_int32),
        ("ordering", C.c_int32),
        ("device", C.c_int32),
        ("rank", C.c_int32),
        ("world", C.c_int32),
        ("comm", C.c_void_p),
    ]


class Iteration(C.Structure):
    _fields_ = [
        ("iteration", C.c_int32),
        ("update_accepted", C.c_int32),
        ("current_lambda", C.c_double),
        ("new_error_linear", C.c_double),
        ("new_error", C.c_double),
        ("relative_reduction", C.c_double),
        ("update_angle_change", C.c_double),
    ]


class Stats(C.Structure):
    _fields_ = [
        ("status", C.c_int32),
        ("failure_reason", C.c_int32),
        ("best_index", C.c_int32),
        ("n_iterations", C.c_int32),
    ]


class Timings(C.Structure):
    _fields_ = [
        ("total_ms", C.c_double),
        ("linearize_ms", C.c_double),
        ("schur_ms", C.c_double),
        ("factorize_ms", C.c_double),
        ("solve_ms", C.c_double),
        ("update_ms", C.c_double),
        ("n_linearize", C.c_int32),
        ("n_factorize", C.c_int32),
        ("kernel_launches", C.c_int32),
        ("iterations_run", C.c_int32),
    ]


K_DEFAULT_EPSILON = 10 * np.finfo(np.float64).eps  # sym::kDefaultEpsilon<double>, gen/cpp/sym/util/epsilon.h:31


class Problem:
    """
    Flat problem: `values` (float64 buffer == sym::Values::data_), optimized `keys`
    [(type, offset, storage_dim, tangent_dim)] in keys_ order, and factor batches
    (kind, arg_offsets[n_args, n], opt_keys[n_opt, n], factor_index[n]).
    """

    def __init__(self, values, keys, batches, solver=SOLVER_CHOLESKY, schur_num_keys=0,
                 params=None, epsilon=K_DEFAULT_EPSILON, ordering=ORDERING_METIS_SCALAR):
        self.values = np.ascontiguousarray(values, dtype=np.float64)
        self.keys = np.ascontiguousarray(keys, dtype=np.int32).reshape(-1, 4)
        self.batches = []
        for kind, arg_offsets, opt_keys, factor_index in batches:
            meta = KINDS[kind]
            ao = np.ascontiguousarray(arg_offsets, dtype=np.int32)
            ok = np.ascontiguousarray(opt_keys, dtype=np.int32)
            fi = np.ascontiguousarray(factor_index, dtype=np.int32)
            n = fi.shape[0]
            assert ao.shape == (meta["n_args"], n), (ao.shape, meta["n_args"], n)
            assert ok.shape == (len(meta["opt_args"]), n)
            self.batches.append((kind, ao, ok, fi))
        self.n_factors = int(sum(b[3].shape[0] for b in self.batches))
        self.solver = solver
        self.schur_num_keys = schur_num_keys
        self.params = params if params is not None else default_params()
        self.epsilon = float(epsilon)
        self.ordering = ordering

    @property
    def tangent_dim(self):
        return int(self.keys[:, 3].sum())

    def desc(self, device=0, rank=0, world=1, comm=None):
        """Returns (ProblemDesc, keepalive) -- keepalive must outlive the create call."""
        nk = self.keys.shape[0]
        keys_arr = (KeyEntry * nk)()
        C.memmove(keys_arr, self.keys.ctypes.data, self.keys.nbytes)
        nb = len(self.batches)
        b_arr = (FactorBatch * nb)()
        for i, (kind, ao, ok, fi) in enumerate(self.batches):
            b_arr[i].kind = kind
            b_arr[i].n = fi.shape[0]
            b_arr[i].arg_offsets = ao.ctypes.data_as(C.POINTER(C.c_int32))
            b_arr[i].opt_keys = ok.ctypes.data_as(C.POINTER(C.c_int32))
            b_arr[i].factor_index = fi.ctypes.data_as(C.POINTER(C.c_int32))
        d = ProblemDesc()
        d.abi_version = SFX_ABI_VERSION
        d.params = self.params
        d.epsilon = self.epsilon
        d.n_values = self.values.shape[0]
        d.n_keys = nk
        d.keys = keys_arr
        d.n_batches = nb
        d.batches = b_arr
        d.n_factors = self.n_factors
        d.solver = self.solver
        d.schur_num_keys = self.schur_num_keys
        d.ordering = self.ordering
        d.device = device
        d.rank = rank
        d.world = world
        d.comm = comm
        return d, (keys_arr, b_arr, self)


class _LibProblem:
    """Common ctypes wrapper over a library exporting the sfx_/orc_ entry points."""

    prefix = None
    lib = None

    def _fn(self, name):
        return getattr(self.lib, self.prefix + name)

    def _check(self, rc, what):
        if rc != 0:
            raise RuntimeError(f"{self.prefix}{what} failed (rc={rc}): {self._last_error()}")

    def set_values(self, values):
        v = np.ascontiguousarray(values, dtype=np.float64)
        self._check(self._fn("set_values")(self.h, v.ctypes.data_as(C.POINTER(C.c_double)),
                                           C.c_int64(v.shape[0])), "set_values")

    def update_params(self, params):
        self._check(self._fn("update_params")(self.h, C.byref(params)), "update_params")

    def optimize(self, num_iterations=-1):
        st = Stats()
        self._check(self._fn("optimize")(self.h, C.c_int32(num_iterations), C.byref(st)), "optimize")
        return st

    def optimize_continue(self, num_iterations):
        """OptimizeContinue of gnc_optimizer.h:133-142: ResetState(values last set) + IterateToConvergence."""
        st = Stats()
        self._check(self._fn("optimize_continue")(self.h, C.c_int32(num_iterations), C.byref(st)), "optimize_continue")
        return st

    def relax_damping_to_initial(self):
        self._check(self._fn("relax_damping_to_initial")(self.h), "relax_damping_to_initial")

    def best_values(self, out=None):
        """values = GetBestValues(); `out` (float64, contiguous, e.g. a pinned buffer) avoids the allocation."""
        if out is None:
            out = np.empty(self.n_values, dtype=np.float64)
        assert out.dtype == np.float64 and out.flags["C_CONTIGUOUS"] and out.shape[0] == self.n_values
        self._check(self._fn("get_best_values")(self.h, out.ctypes.data_as(C.POINTER(C.c_double)),
                                                C.c_int64(out.shape[0])), "get_best_values")
        return out

    def iterations(self):
        cap = 4096
        buf = (Iteration * cap)()
        n = C.c_int32(0)
        self._check(self._fn("get_iterations")(self.h, buf, C.c_int32(cap), C.byref(n)), "get_iterations")
        return [buf[i] for i in range(n.value)]

    def dims(self):
        N, M, nnz = C.c_int32(), C.c_int32(), C.c_int64()
        self._check(self._fn("get_dims")(self.h, C.byref(N), C.byref(M), C.byref(nnz)), "get_dims")
        return N.value, M.value, nnz.value

    def hessian_pattern(self):
        N, M, nnz = self.dims()
        outer = np.empty(N + 1, dtype=np.int32)
        inner = np.empty(nnz, dtype=np.int32)
        self._check(self._fn("get_hessian_pattern")(self.h, outer.ctypes.data_as(C.POINTER(C.c_int32)),
                                                    inner.ctypes.data_as(C.POINTER(C.c_int32))), "get_hessian_pattern")
        return outer, inner

    def _lin(self, fn):
        N, M, nnz = self.dims()
        res = np.empty(M)
        rhs = np.empty(N)
        H = np.empty(nnz)
        p = C.POINTER(C.c_double)
        self._check(self._fn(fn)(self.h, res.ctypes.data_as(p), rhs.ctypes.data_as(p), H.ctypes.data_as(p)), fn)
        return res, rhs, H

    def linearize(self):
        return self._lin("linearize")

    def best_linearization(self):
        return self._lin("get_best_linearization")

    def solve_step(self, lam):
        N, _, _ = self.dims()
        upd = np.empty(N)
        self._check(self._fn("solve_step")(self.h, C.c_double(lam), upd.ctypes.data_as(C.POINTER(C.c_double))),
                    "solve_step")
        return upd

    def ordering(self):
        N, _, _ = self.dims()
        perm = np.empty(N, dtype=np.int32)
        n = C.c_int32()
        self._check(self._fn("get_ordering")(self.h, perm.ctypes.data_as(C.POINTER(C.c_int32)), C.c_int32(N),
                                             C.byref(n)), "get_ordering")
        return perm[: n.value]
