"""
Python front of the GPU path with the interface of the reference's `symforce.opt.optimizer.Optimizer`
(symforce/opt/optimizer.py:30-420, which wraps cc_sym.Optimizer of symforce/pybind/cc_optimizer.cc:29-169):

    optimizer = Optimizer(factors=[Factor(keys=[...], residual=residuals.matching_residual), ...],
                          optimized_keys=[...], params=Optimizer.Params(...))
    result = optimizer.optimize(Values(...))      # result.optimized_values, .iterations, .status, .error()
    optimizer.linearize(values); optimizer.compute_all_covariances(values); optimizer.linearization_index()

What differs from the reference, and why: a reference `Factor` carries a symbolic residual that is code-generated
into a host function; here the residual names one of the device factor kinds compiled into libsfx.so
(symforce_b200/kinds.json, generated from the same symbolic definitions by tools/gen_factors.py), because the
linearization runs on the GPU.  A residual that is not a device kind, or an optimized key at an argument the kind
does not differentiate, raises ValueError at construction -- there is no host fallback.

Everything numeric goes through the C ABI (capi.SfxProblem -> include/sfx.h); this module only flattens `Values`
into the `sym::Values::data_` layout, indexes the factors and rebuilds Python objects from the results.
"""
from __future__ import annotations

import copy
import enum
from dataclasses import dataclass, fields
from functools import cached_property

import numpy as np

from . import capi
from . import desc as D
from .geo import K_DEFAULT_EPSILON, Pose3, Rot3

__all__ = ["Values", "Factor", "Optimizer", "OptimizerParams", "Rot3", "Pose3", "residuals", "Linearization",
           "index_entry_t", "type_t", "optimization_status_t", "levenberg_marquardt_solver_failure_reason_t",
           "lambda_update_type_t"]


# ----------------------------------------------------------------------------------------------------------------
# enums / small messages (lcmtypes/symforce.lcm:117-131, 229-262, 279-299; lcmtypes/symforce_types.lcm:12-50)
# ----------------------------------------------------------------------------------------------------------------
class lambda_update_type_t(enum.IntEnum):
    INVALID = 0
    STATIC = 1
    DYNAMIC = 2


class optimization_status_t(enum.IntEnum):
    INVALID = 0
    SUCCESS = 1
    HIT_ITERATION_LIMIT = 2
    FAILED = 3


class levenberg_marquardt_solver_failure_reason_t(enum.IntEnum):
    INVALID = 0
    LAMBDA_OUT_OF_BOUNDS = 1
    INITIAL_ERROR_NOT_FINITE = 2


class type_t(enum.IntEnum):
    INVALID = 0
    SCALAR = 1
    ROT3 = 3
    POSE3 = 5
    VECTORX = 10


@dataclass(frozen=True)
class index_entry_t:
    """lcmtypes/symforce_types.lcm index_entry_t; in a linearization index `offset` counts tangent scalars."""
    key: str
    type: type_t
    offset: int
    storage_dim: int
    tangent_dim: int


@dataclass
class optimization_iteration_t:
    iteration: int
    current_lambda: float
    new_error_linear: float
    new_error: float
    relative_reduction: float
    update_accepted: bool
    update_angle_change: float
    # debug_stats payloads (empty otherwise), levenberg_marquardt_solver.tcc:115-122, 166-177
    values: np.ndarray = None
    residual: np.ndarray = None
    update: np.ndarray = None  # empty in the record of iteration -1
    jacobian_values: np.ndarray = None  # with include_jacobians: in the order of Result.jacobian_sparsity


@dataclass
class sparse_matrix_structure_t:
    """lcmtypes/symforce.lcm:268-277"""
    row_indices: np.ndarray = None
    column_pointers: np.ndarray = None
    shape: tuple = ()


# ----------------------------------------------------------------------------------------------------------------
# OptimizerParams (symforce/opt/optimizer_params.py:12-48; defaults == sym::DefaultOptimizerParams, optimizer.cc:8-56)
# ----------------------------------------------------------------------------------------------------------------
@dataclass
class OptimizerParams:
    verbose: bool = False
    debug_stats: bool = False
    check_derivatives: bool = False
    include_jacobians: bool = False
    debug_checks: bool = False
    initial_lambda: float = 1.0
    lambda_lower_bound: float = 0.0
    lambda_upper_bound: float = 1000000.0
    lambda_update_type: lambda_update_type_t = lambda_update_type_t.STATIC
    lambda_up_factor: float = 4.0
    lambda_down_factor: float = 1 / 4.0
    dynamic_lambda_update_beta: float = 2.0
    dynamic_lambda_update_gamma: float = 3.0
    dynamic_lambda_update_p: int = 3
    use_diagonal_damping: bool = False
    use_unit_damping: bool = True
    keep_max_diagonal_damping: bool = False
    diagonal_damping_min: float = 1e-6
    iterations: int = 50
    early_exit_min_reduction: float = 1e-6
    early_exit_min_absolute_error: float = 0.0
    enable_bold_updates: bool = False

    def to_c(self) -> D.Params:
        """The `sfx_optimizer_params` of include/sfx.h (the reference's `to_lcm`)."""
        p = D.Params()
        for f in fields(self):
            setattr(p, f.name, type(getattr(p, f.name))(getattr(self, f.name)))
        return p


# ----------------------------------------------------------------------------------------------------------------
# Values (symforce/values/values.py: ordered, nested; keys_recursive / items_recursive / to_storage)
# ----------------------------------------------------------------------------------------------------------------
def _leaf_storage(v):
    """(type, storage as a 1-D float64 array, tangent_dim) of one leaf, as cc_sym.Values.set stores it (sym::Values::Set,
    values.h:95-140: scalars as 1 double, Eigen matrices column-major, geo types by StorageOps)."""
    if isinstance(v, Pose3):
        return D.TYPE_POSE3, v.data, Pose3.TANGENT_DIM
    if isinstance(v, Rot3):
        return D.TYPE_ROT3, v.data, Rot3.TANGENT_DIM
    if isinstance(v, (bool, int, float, np.integer, np.floating)):
        return D.TYPE_VECTOR, np.array([v], dtype=np.float64), 1
    if isinstance(v, np.ndarray):
        if v.ndim > 2:
            raise TypeError(f"arrays with {v.ndim} axes are not a Values leaf; use nested lists for the leading axes")
        flat = np.asarray(v, dtype=np.float64).reshape(-1, order="F")
        return D.TYPE_VECTOR, flat, flat.shape[0]
    raise TypeError(f"unsupported Values leaf type {type(v).__name__}")


def _leaf_from_storage(template, data):
    if isinstance(template, Pose3):
        return Pose3.from_storage(data)
    if isinstance(template, Rot3):
        return Rot3.from_storage(data)
    if isinstance(template, np.ndarray):
        return np.asarray(data, dtype=np.float64).reshape(template.shape, order="F").copy()
    return float(data[0])


class Values:
    """Ordered key -> value container; lists and nested Values flatten to `name[i]` / `name.sub` keys."""

    def __init__(self, **kwargs):
        self._d = {}
        for k, v in kwargs.items():
            self[k] = v

    # -- mapping ---------------------------------------------------------------------------------------------------
    def __setitem__(self, key, value):
        if isinstance(value, dict):
            value = Values(**value)
        elif isinstance(value, tuple):
            value = list(value)
        if key in self._d or "." not in key and "[" not in key:
            self._d[key] = value
            return
        parent, last = self._resolve_parent(key)
        parent[last] = value

    def __getitem__(self, key):
        if key in self._d:
            return self._d[key]
        parent, last = self._resolve_parent(key)
        return parent[last]

    def __contains__(self, key):
        try:
            self[key]
            return True
        except (KeyError, IndexError, TypeError):
            return False

    def __len__(self):
        return len(self._d)

    def keys(self):
        return self._d.keys()

    def items(self):
        return self._d.items()

    def _resolve_parent(self, key):
        """Walks a flattened key ('a.b[2][0]') to its parent container and final index."""
        import re

        tokens = []
        for part in key.split("."):
            m = re.fullmatch(r"([^\[\]]+)((\[\d+\])*)", part)
            if not m:
                raise KeyError(key)
            tokens.append(m.group(1))
            tokens.extend(int(i) for i in re.findall(r"\[(\d+)\]", m.group(2)))
        node = self._d
        for t in tokens[:-1]:
            node = node._d[t] if isinstance(node, Values) else node[t]
        if isinstance(node, Values):
            node = node._d
        if isinstance(tokens[-1], str) and tokens[-1] not in node and len(tokens) > 1:
            raise KeyError(key)
        return node, tokens[-1]

    # -- flattening ------------------------------------------------------------------------------------------------
    @staticmethod
    def _walk(prefix, v, out):
        if isinstance(v, Values):
            for k, x in v._d.items():
                Values._walk(f"{prefix}.{k}" if prefix else k, x, out)
        elif isinstance(v, (list, tuple)):
            for i, x in enumerate(v):
                Values._walk(f"{prefix}[{i}]", x, out)
        else:
            out.append((prefix, v))

    def items_recursive(self):
        out = []
        Values._walk("", self, out)
        return out

    def keys_recursive(self):
        return [k for k, _ in self.items_recursive()]

    def values_recursive(self):
        return [v for _, v in self.items_recursive()]

    def to_storage(self):
        out = []
        for _, v in self.items_recursive():
            out.extend(float(x) for x in _leaf_storage(v)[1])
        return out

    def to_numerical(self):
        return self

    def dataclasses_to_values(self):
        return self

    def copy(self):
        return copy.deepcopy(self)

    def __repr__(self):
        return "Values(\n" + "".join(f"  {k}: {v!r}\n" for k, v in self.items_recursive()) + ")"


# ----------------------------------------------------------------------------------------------------------------
# Residuals = device factor kinds, and Factor (symforce/opt/factor.py:33-120, numeric_factor.py:20-60)
# ----------------------------------------------------------------------------------------------------------------
class Residual:
    """A residual function the GPU library implements: `kind` indexes symforce_b200/kinds.json."""

    def __init__(self, kind: int):
        self.kind = kind
        self.meta = D.KINDS[kind]
        self.__name__ = self.meta["name"]

    def __repr__(self):
        return f"<device residual {self.meta['name']}({', '.join(self.meta['arg_names'])})>"


class _Residuals:
    """Namespace of the device kinds under the names of the reference functions they were generated from."""

    def __init__(self):
        self.by_name = {}
        for kind, meta in enumerate(D.KINDS):
            r = Residual(kind)
            self.by_name[meta["name"]] = r
            setattr(self, meta["name"], r)
        # reference names: examples/robot_3d_localization/robot_3d_localization.py:119-150,
        # examples/bundle_adjustment_in_the_large/bundle_adjustment_in_the_large.py:18-63,
        # symforce/codegen/slam_factors_codegen.py (inverse_range_landmark_*), geo_factors_codegen.py (between/prior)
        self.matching_residual = self.matching
        self.odometry_residual = self.odometry
        self.snavely_reprojection_residual = self.snavely
        self.between_factor_pose3 = self.between_pose3
        self.prior_factor_pose3 = self.prior_pose3
        self.between_factor_rot3 = self.between_rot3
        self.prior_factor_rot3 = self.prior_rot3
        self.inverse_range_landmark_linear_gnc_factor = self.irl_linear_gnc
        self.inverse_range_landmark_prior_factor = self.irl_prior
        self.barron_factor = self.barron  # test/symforce_gnc_codegen_test.py:24-33

    def get(self, r):
        if isinstance(r, Residual):
            return r
        if isinstance(r, str) and r in self.by_name:
            return self.by_name[r]
        if isinstance(r, str) and isinstance(getattr(self, r, None), Residual):
            return getattr(self, r)
        raise ValueError(
            f"residual {r!r} is not a device factor kind ({', '.join(self.by_name)}): the GPU path linearizes "
            "compiled kinds only and has no host fallback for Python residual functions")


residuals = _Residuals()


class Factor:
    """
    A residual over Values keys.  `keys` are in the argument order of the residual, as for the reference's Factor;
    `optimized_keys`, if given, makes this the analogue of a NumericFactor (the Optimizer then collects the
    optimized keys from its factors, optimizer.py:206-213).
    """

    def __init__(self, keys, residual, name=None, optimized_keys=None):
        self.keys = list(keys)
        self.residual = residuals.get(residual)
        self.name = name if name is not None else self.residual.meta["name"]
        meta = self.residual.meta
        if len(self.keys) != meta["n_args"]:
            raise ValueError(f"Factor {self.name} takes {meta['n_args']} keys ({', '.join(meta['arg_names'])}), "
                             f"got {len(self.keys)}: {self.keys}")
        self.optimized_keys = None if optimized_keys is None else list(optimized_keys)
        if self.optimized_keys is not None:
            for k in self.optimized_keys:
                if k not in self.keys:
                    raise ValueError(f"optimized key {k} is not an argument of factor {self.name} (keys: {self.keys})")


# ----------------------------------------------------------------------------------------------------------------
# Linearization (cc_sym.Linearization, symforce/opt/linearization.h:22-93)
# ----------------------------------------------------------------------------------------------------------------
class Linearization:
    def __init__(self, residual, rhs, outer, inner, hvalues, jac=None):
        self.residual = residual
        self.rhs = rhs
        self._outer, self._inner, self._hvalues = outer, inner, hvalues
        self._jac = jac  # (outer, inner, values) when OptimizerParams.include_jacobians

    @cached_property
    def hessian_lower(self):
        import scipy.sparse as sp

        n = self.rhs.shape[0]
        return sp.csc_matrix((self._hvalues, self._inner, self._outer), shape=(n, n))

    @cached_property
    def jacobian(self):
        """M x N sparse Jacobian (linearization.h:58-60); exported from the device when include_jacobians is set."""
        if self._jac is None:
            raise ValueError("jacobian is filled out when OptimizerParams.include_jacobians is True")
        import scipy.sparse as sp

        outer, inner, values = self._jac
        return sp.csc_matrix((values, inner, outer), shape=(self.residual.shape[0], self.rhs.shape[0]))

    def error(self):
        return 0.5 * float(self.residual @ self.residual)

    def linear_error(self, x_update):
        """linearization.h:62-67 without J: 0.5 |r - J dx|^2 = error - dx.rhs + 0.5 dx.H.dx  (rhs = J^T r)."""
        H = self.hessian_lower
        Hx = H @ x_update + H.T @ x_update - H.diagonal() * x_update
        return self.error() - float(x_update @ self.rhs) + 0.5 * float(x_update @ Hx)


class _Stats:
    """cc_sym.OptimizationStats (lcmtypes/symforce.lcm:302-330) read back through the C ABI."""

    def __init__(self, st, iterations, best_linearization, ordering, jacobian_sparsity=None):
        self.jacobian_sparsity = jacobian_sparsity if jacobian_sparsity is not None else sparse_matrix_structure_t()
        # not exposed by this linear solver (supernodal LL^T in dense fronts): default constructed, as the reference
        # leaves it for solvers without L() (optimization_stats.h:55-60)
        self.cholesky_factor_sparsity = sparse_matrix_structure_t()
        self.status = optimization_status_t(st.status)
        self.failure_reason = int(st.failure_reason)
        self.best_index = int(st.best_index)
        self.iterations = iterations
        self.best_linearization = best_linearization
        self.linear_solver_ordering = ordering


# ----------------------------------------------------------------------------------------------------------------
# Optimizer
# ----------------------------------------------------------------------------------------------------------------
class Optimizer:
    """See the module docstring; argument meaning and error behaviour follow symforce/opt/optimizer.py."""

    Params = OptimizerParams
    Status = optimization_status_t
    FailureReason = levenberg_marquardt_solver_failure_reason_t

    @dataclass
    class Result:
        """optimizer.py:93-176"""

        initial_values: Values
        optimized_values: Values
        _stats: _Stats

        @property
        def iterations(self):
            return self._stats.iterations

        @property
        def best_index(self):
            return self._stats.best_index

        @property
        def status(self):
            return self._stats.status

        @property
        def failure_reason(self):
            return Optimizer.FailureReason(self._stats.failure_reason)

        @property
        def best_linearization(self):
            return self._stats.best_linearization

        @property
        def linear_solver_ordering(self):
            return self._stats.linear_solver_ordering

        @property
        def jacobian_sparsity(self):
            return self._stats.jacobian_sparsity

        @property
        def cholesky_factor_sparsity(self):
            return self._stats.cholesky_factor_sparsity

        def jacobian_view(self, iteration):
            """OptimizationStats::JacobianView (optimization_stats.h:67-75): the M x N Jacobian of a debug_stats record."""
            sparsity = self._stats.jacobian_sparsity
            if len(sparsity.shape) != 2:
                raise ValueError("Jacobian sparsity is empty, did you set debug_stats = true and include_jacobians = true?")
            import scipy.sparse as sp

            return sp.csc_matrix((iteration.jacobian_values, sparsity.row_indices, sparsity.column_pointers),
                                 shape=sparsity.shape)

        def error(self):
            return self.iterations[self.best_index].new_error

    def __init__(self, factors, optimized_keys=None, params=None, *, epsilon=K_DEFAULT_EPSILON, solver="auto",
                 ordering=D.ORDERING_METIS_SCALAR, device=0):
        if optimized_keys is None:
            self.optimized_keys = []
        else:
            self.optimized_keys = list(optimized_keys)
            assert len(optimized_keys) == len(set(optimized_keys)), f"Duplicates in optimized keys: {optimized_keys}"
        optimized_keys_set = set(self.optimized_keys)

        self.factors = []
        self._factor_opt_keys = []
        for factor in factors:
            if factor.optimized_keys is None:
                if optimized_keys is None:
                    raise ValueError("You must specify `optimized_keys` when passing symbolic factors.")
                factor_opt_keys = [k for k in factor.keys if k in optimized_keys_set]
                if not factor_opt_keys:
                    raise ValueError(f"Factor {factor.name} has no arguments (keys: {factor.keys}) in "
                                     f"optimized_keys ({optimized_keys}).")
            else:
                factor_opt_keys = factor.optimized_keys
                for k in factor_opt_keys:
                    if k not in optimized_keys_set:
                        optimized_keys_set.add(k)
                        self.optimized_keys.append(k)
            meta = factor.residual.meta
            for k in factor_opt_keys:
                if factor.keys.index(k) not in meta["opt_args"]:
                    raise ValueError(
                        f"Factor {factor.name}: key {k} is argument `{meta['arg_names'][factor.keys.index(k)]}` of the "
                        f"device kind `{meta['name']}`, which is only differentiated with respect to "
                        f"{[meta['arg_names'][a] for a in meta['opt_args']]}")
            self.factors.append(factor)
            self._factor_opt_keys.append(factor_opt_keys)

        self.params = OptimizerParams(verbose=True) if params is None else params
        # SYM_ASSERT(!params.check_derivatives || params.include_jacobians) (optimizer.tcc:39, 62)
        assert not self.params.check_derivatives or self.params.include_jacobians, \
            "check_derivatives needs include_jacobians"
        self.epsilon = float(epsilon)
        if solver not in ("auto", "cholesky", "schur") and not isinstance(solver, int):
            raise ValueError("solver must be 'auto', 'cholesky', 'schur' or a number of trailing keys to eliminate")
        self._solver = solver
        self._ordering = ordering
        self._device = device
        self._initialized = False
        self.values_keys_ordered = None
        self._problem = None  # desc.Problem
        self._gpu = None  # capi.SfxProblem
        self._cov_gpu = None
        self._cov_split = None
        self._check_gpu = None  # check_derivatives: same structure and solver, its own LM state

    # -- lowering to the flat problem of include/sfx.h --------------------------------------------------------------
    def _initialize(self, values: Values):
        """
        Fixes the storage layout and indexes the factors.  Storage order = the order cc_sym.Values is filled in
        by the reference (optimizer.py:262-271 iterates `_cc_keys_map`: optimized keys first, then the remaining
        keys of the Values in keys_recursive order); state-vector order = `optimized_keys` (cc keys x_0, x_1, ...
        sort to that order in sym::ComputeKeysToOptimize, factor.h:424-449).
        """
        leaves = dict(values.items_recursive())
        order = list(self.optimized_keys)
        seen = set(order)
        for k in leaves:
            if k not in seen:
                order.append(k)
                seen.add(k)
        layout = {}
        off = 0
        for k in order:
            if k not in leaves:
                raise KeyError(f"optimized key {k} is not in the Values")
            typ, storage, tdim = _leaf_storage(leaves[k])
            layout[k] = (typ, off, len(storage), tdim)
            off += len(storage)
        self._layout = layout
        self._storage_order = order
        self._n_values = off
        self.values_keys_ordered = list(leaves)
        self._template = values.copy()

        key_index = {k: i for i, k in enumerate(self.optimized_keys)}
        keys = [layout[k] for k in self.optimized_keys]
        by_kind = {}
        for fi, (f, fopt) in enumerate(zip(self.factors, self._factor_opt_keys)):
            meta = f.residual.meta
            args = []
            for a, k in enumerate(f.keys):
                if k not in layout:
                    raise KeyError(f"Factor {f.name}: key {k} is not in the Values")
                if layout[k][2] != meta["arg_dims"][a]:
                    raise ValueError(f"Factor {f.name}: key {k} has {layout[k][2]} storage elements, argument "
                                     f"`{meta['arg_names'][a]}` takes {meta['arg_dims'][a]}")
                args.append(layout[k][1])
            fset = set(fopt)
            opt = [key_index[f.keys[a]] if f.keys[a] in fset else -1 for a in meta["opt_args"]]
            b = by_kind.setdefault(f.residual.kind, ([], [], []))
            b[0].append(args)
            b[1].append(opt)
            b[2].append(fi)
        batches = [(kind, np.array(a, dtype=np.int32).T, np.array(o, dtype=np.int32).T, np.array(i, dtype=np.int32))
                   for kind, (a, o, i) in sorted(by_kind.items())]

        n_elim = self._schur_keys(keys)
        self._problem = D.Problem(np.zeros(self._n_values), keys, batches,
                                  solver=D.SOLVER_SCHUR if n_elim > 0 else D.SOLVER_CHOLESKY, schur_num_keys=n_elim,
                                  params=self.params.to_c(), epsilon=self.epsilon, ordering=self._ordering)
        self._initialized = True

    def _schur_keys(self, keys):
        """Number of trailing keys eliminated per block (sym::SparseSchurSolver's C).  'auto': the longest trailing
        run of vector keys of dimension <= 3 no two of which share a factor, taken when it is at least half of
        the keys -- the rule of include/sym/sym.h AutoSchurKeys."""
        nk = len(keys)
        if self._solver == "cholesky":
            return 0
        if isinstance(self._solver, int):
            return int(self._solver)
        first = nk
        while first > 0 and keys[first - 1][0] == D.TYPE_VECTOR and keys[first - 1][3] <= 3:
            first -= 1
        if first < nk:
            key_index = {k: i for i, k in enumerate(self.optimized_keys)}
            for fopt in self._factor_opt_keys:
                idx = sorted((key_index[k] for k in fopt if key_index[k] >= first), reverse=True)
                if len(idx) >= 2:
                    first = max(first, idx[1] + 1)
        n = nk - first
        if first == 0 or n <= 0:
            n = 0
        elif self._solver == "auto" and n < nk // 2:
            n = 0
        if self._solver == "schur" and n == 0:
            raise ValueError("solver='schur': the optimized keys do not end in a run of independent vector keys")
        return n

    def problem(self, values: Values) -> D.Problem:
        """The flat problem (values buffer + key entries + factor batches) `values` lowers to; what
        sfx_problem_create receives."""
        data = self._storage(values)
        p = copy.copy(self._problem)
        p.values = data
        return p

    def _storage(self, values: Values) -> np.ndarray:
        if not self._initialized:
            self._initialize(values)
        leaves = dict(values.items_recursive())
        out = np.empty(self._n_values)
        for k in self._storage_order:
            typ, off, sdim, _ = self._layout[k]
            if k not in leaves:
                raise KeyError(f"key {k} is missing from the Values")
            t2, storage, _ = _leaf_storage(leaves[k])
            if t2 != typ or len(storage) != sdim:
                raise ValueError(f"key {k} changed type or size since the first call")
            out[off:off + sdim] = storage
        return out

    def _values_from_storage(self, template: Values, data) -> Values:
        out = template.copy()
        for k, v in template.items_recursive():
            _, off, sdim, _ = self._layout[k]
            out[k] = _leaf_from_storage(v, data[off:off + sdim])
        return out

    def _device_problem(self, values: Values):
        data = self._storage(values)
        if self._gpu is None:
            self._problem.values = data
            self._gpu = capi.SfxProblem(self._problem, device=self._device)  # raises without a CUDA device
        else:
            self._gpu.set_values(data)
        return self._gpu

    # -- the reference interface -----------------------------------------------------------------------------------
    def optimize(self, initial_guess: Values, num_iterations: int = -1,
                 populate_best_linearization: bool = False) -> "Optimizer.Result":
        """optimizer.py:322-358 -> Optimizer::Optimize (optimizer.tcc:59-83)."""
        gpu = self._device_problem(initial_guess)
        # check_derivatives (optimizer.tcc:261-272): the reference asserts at the values of every linearization; here on a
        # sibling problem at the initial values and, after the run, at every record's values (debug_stats) or the best ones
        if self.params.check_derivatives:
            self._assert_derivatives(self._storage(initial_guess))
        st = gpu.optimize(num_iterations)
        best = gpu.best_values()
        if self.params.check_derivatives and not self.params.debug_stats:
            self._assert_derivatives(best)
        optimized_values = self._values_from_storage(initial_guess, best)
        its = []
        for r, it in enumerate(gpu.iterations()):
            rec = optimization_iteration_t(
                iteration=int(it.iteration), current_lambda=it.current_lambda, new_error_linear=it.new_error_linear,
                new_error=it.new_error, relative_reduction=it.relative_reduction,
                update_accepted=bool(it.update_accepted), update_angle_change=it.update_angle_change)
            if self.params.debug_stats:
                rec.values, rec.residual = gpu.iteration_debug(r)
                rec.update = gpu.iteration_update(r) if it.iteration >= 0 else np.zeros(0)
                if self.params.include_jacobians:
                    rec.jacobian_values = gpu.iteration_jacobian(r)
                if self.params.check_derivatives and r > 0:
                    self._assert_derivatives(rec.values)
            its.append(rec)
        jac_sparsity = None
        if self.params.debug_stats and self.params.include_jacobians:  # levenberg_marquardt_solver.tcc:172-175
            n, m, _ = gpu.dims()
            j_outer, j_inner = gpu.jacobian_pattern()
            jac_sparsity = sparse_matrix_structure_t(row_indices=j_inner, column_pointers=j_outer, shape=(m, n))
        best_lin = None
        if populate_best_linearization:
            res, rhs, H = gpu.best_linearization()
            outer, inner = gpu.hessian_pattern()
            best_lin = Linearization(res, rhs, outer, inner, H)
        ordering = gpu.ordering() if self.params.debug_stats else np.zeros(0, dtype=np.int32)
        return Optimizer.Result(initial_values=initial_guess, optimized_values=optimized_values,
                                _stats=_Stats(st, its, best_lin, ordering, jac_sparsity))

    def linearize(self, values: Values) -> Linearization:
        """optimizer.py:360-364 -> Optimizer::Linearize (optimizer.tcc:85-92)."""
        gpu = self._device_problem(values)
        res, rhs, H = gpu.linearize()
        outer, inner = gpu.hessian_pattern()
        jac = gpu.jacobian() if self.params.include_jacobians else None
        if self.params.check_derivatives:
            self._assert_derivatives(self._storage(values))
        return Linearization(res, rhs, outer, inner, H, jac)

    def check_derivatives(self, values: Values):
        """internal::CheckDerivatives (internal/derivative_checker.h:32-123) at `values`: (ok, relative errors of the
        Jacobian against central differences, of hessian_lower against J^T J, of rhs against J^T r)."""
        self._device_problem(values)
        return self._check_derivatives(self._storage(values))

    def _check_derivatives(self, data):
        if self._check_gpu is None:
            p = copy.copy(self._problem)
            p.values = np.ascontiguousarray(data, dtype=np.float64)
            self._check_gpu = capi.SfxProblem(p, device=self._device)
        else:
            self._check_gpu.set_values(np.ascontiguousarray(data, dtype=np.float64))
        return self._check_gpu.check_derivatives()

    def _assert_derivatives(self, data):
        ok, err = self._check_derivatives(data)
        if not ok:
            raise RuntimeError("SYM_ASSERT: internal::CheckDerivatives(linearizer_, values, index_, linearization, epsilon_): "
                               f"{err}")

    def load_iteration_values(self, values_data) -> Values:
        """optimizer.py:366-381: a debug_stats iteration's values (flat storage) as a Python Values."""
        assert self._initialized, "load_iteration_values is available after the first optimize / linearize"
        return self._values_from_storage(self._template, np.asarray(values_data))

    def linearization_index(self):
        """optimizer.py:383-392"""
        return {k: self.linearization_index_entry(k) for k in self.optimized_keys}

    def linearization_index_entry(self, key: str) -> index_entry_t:
        """optimizer.py:394-405 -> Linearizer state index: offsets in tangent scalars, keys_ order."""
        assert self._initialized, "linearization_index is available after the first optimize / linearize"
        off = 0
        for k in self.optimized_keys:
            typ, _, sdim, tdim = self._layout[k]
            if k == key:
                t = {D.TYPE_POSE3: type_t.POSE3, D.TYPE_ROT3: type_t.ROT3}.get(
                    typ, type_t.SCALAR if sdim == 1 else type_t.VECTORX)
                return index_entry_t(key=k, type=t, offset=off, storage_dim=sdim, tangent_dim=tdim)
            off += tdim
        raise KeyError(key)

    # -- covariances (optimizer.py:273-320 -> optimizer.tcc:113-121, 177-206) -------------------------------------------
    def _cov_problem(self, n_elim):
        """Device problem whose linear solver eliminates the trailing n_elim keys (0: none)."""
        if (self._problem.schur_num_keys if self._problem.solver == D.SOLVER_SCHUR else 0) == n_elim:
            return self._gpu
        if self._cov_gpu is not None and self._cov_split != n_elim:
            self._cov_gpu.close()
            self._cov_gpu = None
        if self._cov_gpu is None:
            p = copy.copy(self._problem)
            p.solver = D.SOLVER_SCHUR if n_elim > 0 else D.SOLVER_CHOLESKY
            p.schur_num_keys = n_elim
            self._cov_gpu = capi.SfxProblem(p, device=self._device)
            self._cov_split = n_elim
        return self._cov_gpu

    def _split_by_key(self, cov, keys):
        out, off = {}, 0
        for k in keys:
            d = self._layout[k][3]
            out[k] = cov[off:off + d, off:off + d].copy()
            off += d
        return out

    def compute_full_covariance(self, optimized_value: Values) -> np.ndarray:
        lin = self.linearize(optimized_value)
        n = lin.rhs.shape[0]
        return np.array(self._cov_problem(0).compute_covariance(n, lin._hvalues))

    def compute_all_covariances(self, optimized_value: Values):
        return self._split_by_key(self.compute_full_covariance(optimized_value), self.optimized_keys)

    def compute_covariances(self, optimized_value: Values, keys, c_is_block_diagonal=True):
        """optimizer.py:294-321 -> Optimizer::ComputeCovariances (optimizer.tcc:177-199).  c_is_block_diagonal (the C++
        argument, optimizer.h:207-220): True eliminates the later keys with the block-diagonal Schur solver, False allows
        any structure of C (covariance_utils.h:41-103)."""
        keys = list(keys)
        if keys != self.optimized_keys[:len(keys)] or not keys:
            raise ValueError("keys must be the first optimized keys, in order (CheckKeyOrderMatchesLinearizerKeysStart)")
        if len(keys) == len(self.optimized_keys):
            return self.compute_all_covariances(optimized_value)
        lin = self.linearize(optimized_value)
        dim = sum(self._layout[k][3] for k in keys)
        n_elim = len(self.optimized_keys) - len(keys) if c_is_block_diagonal else 0
        cov = self._cov_problem(n_elim).compute_covariance(dim, lin._hvalues)
        return self._split_by_key(np.array(cov), keys)

    def close(self):
        for g in (self._gpu, self._cov_gpu, self._check_gpu):
            if g is not None:
                g.close()
        self._gpu = self._cov_gpu = self._check_gpu = None
