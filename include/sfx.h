/*
 * sfx.h -- C ABI of symforce_b200: the sparse Levenberg-Marquardt inner loop of SymForce
 * (linearize -> [Schur] -> Cholesky -> retract -> accept/reject) running entirely on one B200.
 *
 * This is the drop-in boundary.  Every entry point cites the reference
 * interface it replaces (paths relative to the reference checkout).  The C++17 `sym::` header
 * layer under include/sym/ lowers sym::Optimizer / sym::Factor / sym::Values onto these calls
 * and turns return codes back into std::runtime_error / optimization_status_t.
 *
 * Conventions
 *   - plain pointers and sizes only; all pointers are caller-owned HOST memory unless the
 *     function name ends in `_device`;
 *   - the handle owns all device memory, streams, events and CUDA graphs; one handle == one
 *     host thread ("Not thread safe! Create one per thread", symforce/opt/optimizer.h:23);
 *   - every function returns sfx_status; the message for the last failure is
 *     sfx_last_error(handle) (handle may be NULL for a failed sfx_problem_create);
 *   - there is NO CPU fallback: without a CUDA device sfx_problem_create fails with
 *     SFX_ERR_CUDA.
 */
#ifndef SFX_H_
#define SFX_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SFX_ABI_VERSION 1

typedef enum {
  SFX_OK = 0,
  SFX_ERR_INVALID_ARG = 1, /* maps to SYM_ASSERT -> std::runtime_error (symforce/opt/assert.h:40-92) */
  SFX_ERR_CUDA = 2,        /* no device / CUDA runtime failure */
  SFX_ERR_UNSUPPORTED = 3, /* factor kind / key type that has no device implementation */
  SFX_ERR_STRUCTURE = 4,   /* e.g. "Key ... is in the state vector but is not optimized by any
                              factor" (symforce/opt/linearizer.cc:277-284), non block-diagonal C */
  SFX_ERR_NCCL = 5,
  SFX_ERR_NUMERICAL = 6    /* Eigen::NumericalIssue of ComputeCovariances (optimizer.h:214-217) */
} sfx_status;

/* Storage/tangent semantics of an optimized key; subset of sym::type_t
 * (lcmtypes/symforce_types.lcm) that the configs of BASELINE.json use. */
typedef enum {
  SFX_TYPE_VECTOR = 0, /* scalar, VectorN, MatrixNM: storage == tangent, retract is += */
  SFX_TYPE_ROT3 = 1,   /* storage 4 (qx,qy,qz,qw), tangent 3; gen/cpp/sym/ops/rot3/lie_group_ops.cc:63-92 */
  SFX_TYPE_POSE3 = 2   /* storage 7 (q, t), tangent 6 [rot, trans]; gen/cpp/sym/ops/pose3/lie_group_ops.cc:70-103 */
} sfx_type;

/* Factor kinds with a device implementation (order == tools/gen_factors.py == kinds_gen.h). */
typedef enum {
  SFX_KIND_SNAVELY = 0,        /* symforce/examples/bundle_adjustment_in_the_large/gen/snavely_reprojection_factor.h:33 */
  SFX_KIND_BETWEEN_POSE3 = 1,  /* gen/cpp/sym/factors/between_factor_pose3.h:34 */
  SFX_KIND_PRIOR_POSE3 = 2,    /* gen/cpp/sym/factors/prior_factor_pose3.h:32 */
  SFX_KIND_MATCHING = 3,       /* symforce/examples/robot_3d_localization/gen/matching_factor.h:27 */
  SFX_KIND_ODOMETRY = 4,       /* symforce/examples/robot_3d_localization/gen/odometry_factor.h:28 */
  SFX_KIND_IRL_LINEAR_GNC = 5, /* gen/cpp/sym/factors/inverse_range_landmark_linear_gnc_factor.h:58 */
  SFX_KIND_IRL_PRIOR = 6,      /* gen/cpp/sym/factors/inverse_range_landmark_prior_factor.h:30 */
  SFX_KIND_BETWEEN_ROT3 = 7,   /* gen/cpp/sym/factors/between_factor_rot3.h */
  SFX_KIND_PRIOR_ROT3 = 8,     /* gen/cpp/sym/factors/prior_factor_rot3.h */
  SFX_KIND_BARRON = 9,         /* test/symforce_function_codegen_test_data/symengine/gnc_test_data/cpp/symforce/gnc_factors/barron_factor.h:33
                                  (the factor of test/symforce_gnc_test.cc) */
  SFX_KIND_COUNT = 10
} sfx_factor_kind;

/* Mirrors sym::optimizer_params_t field for field (lcmtypes/symforce.lcm:134-200; defaults in
 * symforce/opt/optimizer.cc:8-56 are returned by sfx_default_params). */
typedef struct {
  int32_t verbose;
  int32_t debug_stats;
  int32_t check_derivatives;
  int32_t include_jacobians;
  int32_t debug_checks;
  double initial_lambda;
  double lambda_lower_bound;
  double lambda_upper_bound;
  int32_t lambda_update_type; /* 1 = STATIC, 2 = DYNAMIC (lcmtypes/symforce.lcm:126-131) */
  double lambda_up_factor;
  double lambda_down_factor;
  double dynamic_lambda_update_beta;
  double dynamic_lambda_update_gamma;
  int32_t dynamic_lambda_update_p;
  int32_t use_diagonal_damping;
  int32_t use_unit_damping;
  int32_t keep_max_diagonal_damping;
  double diagonal_damping_min;
  int32_t iterations;
  double early_exit_min_reduction;
  double early_exit_min_absolute_error;
  int32_t enable_bold_updates;
} sfx_params;

/* One optimized key, in `keys_` order (symforce/opt/optimizer.h:241-268 Keys());
 * mirrors sym::index_entry_t minus the key itself (lcmtypes/symforce.lcm:3-26). */
typedef struct {
  int32_t type;        /* sfx_type */
  int32_t offset;      /* into the flat values buffer (Values::data_, symforce/opt/values.h:313) */
  int32_t storage_dim;
  int32_t tangent_dim;
} sfx_key_entry;

/* All factors of one kind.  Arrays are SoA over the n factors of the batch. */
typedef struct {
  int32_t kind;                 /* sfx_factor_kind */
  int32_t n;                    /* number of factors */
  const int32_t* arg_offsets;   /* [n_args(kind)][n]: values-buffer offset of every function
                                   argument, in the generated function's argument order
                                   (what Factor::AllKeys()+index entries give the reference,
                                   symforce/opt/linearizer.cc:194) */
  const int32_t* opt_keys;      /* [n_opt(kind)][n]: index into sfx_problem_desc.keys of every
                                   linearized argument, or -1 when that argument is not optimized
                                   (dropped like FactorOffsets does, internal/linearizer_utils.h:226-232) */
  const int32_t* factor_index;  /* [n]: position of the factor in the caller's factor list; fixes
                                   residual row order and the summation order the CPU reference
                                   uses (symforce/opt/linearizer.cc:71-110) */
} sfx_factor_batch;

typedef enum {
  SFX_SOLVER_CHOLESKY = 0, /* sym::SparseCholeskySolver on the full Hessian (LM default,
                              symforce/opt/levenberg_marquardt_solver.h:131-134) */
  SFX_SOLVER_SCHUR = 1     /* sym::SparseSchurSolver (symforce/opt/sparse_schur_solver.h:62-135):
                              trailing `schur_num_keys` keys are eliminated per block */
} sfx_solver;

typedef enum {
  SFX_ORDERING_METIS_SCALAR = 0, /* METIS_NodeND on the scalar pattern, as Eigen::MetisOrdering does
                                    for the reference (sparse_cholesky_solver.h:57) */
  SFX_ORDERING_METIS_BLOCK = 1,  /* METIS_NodeND on the key-block quotient graph */
  SFX_ORDERING_NATURAL = 2       /* Eigen::NaturalOrdering analogue (test/symforce_optimizer_test.cc:362) */
} sfx_ordering;

typedef struct {
  int32_t abi_version;          /* SFX_ABI_VERSION */
  sfx_params params;
  double epsilon;               /* sym::kDefaultEpsilon<double> unless overridden (optimizer.h:92-94) */
  int64_t n_values;             /* length of the flat values buffer in doubles */
  int32_t n_keys;
  const sfx_key_entry* keys;    /* optimized keys in keys_ order (default: LexicalLessThan order,
                                   symforce/opt/factor.h:424-449) */
  int32_t n_batches;
  const sfx_factor_batch* batches;
  int32_t n_factors;            /* total factors == sum of batch n; factor_index is a permutation of [0,n_factors) */
  int32_t solver;               /* sfx_solver */
  int32_t schur_num_keys;       /* SFX_SOLVER_SCHUR: number of trailing keys forming C */
  int32_t ordering;             /* sfx_ordering */
  int32_t device;               /* CUDA device ordinal */
  /* multi-GPU (landmark sharding): this rank's factors are the ones given; keys and
   * values are replicated.  comm == NULL -> single GPU. */
  int32_t rank;
  int32_t world;
  void* comm;                   /* sfx_comm* from sfx_comm_create, or NULL */
} sfx_problem_desc;

/* Mirrors sym::optimization_iteration_t (lcmtypes/symforce.lcm:229-262) without debug payloads. */
typedef struct {
  int32_t iteration; /* -1 for the initial entry */
  int32_t update_accepted;
  double current_lambda;
  double new_error_linear;
  double new_error;
  double relative_reduction;
  double update_angle_change;
} sfx_iteration;

/* Mirrors sym::OptimizationStats (symforce/opt/optimization_stats.h:22-94). */
typedef struct {
  int32_t status;         /* optimization_status_t: 1 SUCCESS, 2 HIT_ITERATION_LIMIT, 3 FAILED */
  int32_t failure_reason; /* levenberg_marquardt_solver_failure_reason_t: 1 LAMBDA_OUT_OF_BOUNDS,
                             2 INITIAL_ERROR_NOT_FINITE */
  int32_t best_index;
  int32_t n_iterations;   /* number of sfx_iteration entries (incl. the initial one) */
} sfx_stats;

/* Device-time breakdown of the last sfx_optimize (CUDA events on the solver stream), ms. */
typedef struct {
  double total_ms;
  double linearize_ms;
  double schur_ms;
  double factorize_ms;
  double solve_ms;
  double update_ms; /* retract + reductions + accept/reject */
  int32_t n_linearize;
  int32_t n_factorize;
  int32_t kernel_launches;
  int32_t iterations_run;
} sfx_timings;

typedef struct sfx_problem sfx_problem;
typedef struct sfx_comm sfx_comm;

/* sym::DefaultOptimizerParams() -- symforce/opt/optimizer.cc:8-56 */
sfx_status sfx_default_params(sfx_params* out);

/* sym::Optimizer<double>::Optimizer(params, factors, name, keys, epsilon) + the lazy
 * Initialize()/BuildInitialLinearization()/AnalyzeSparsityPattern() of the first Optimize call
 * (symforce/opt/optimizer.tcc:21-40, 273-278; linearizer.cc:149-356;
 * levenberg_marquardt_solver.tcc:189-193): builds index maps, block structure, ordering,
 * symbolic factorization and all device buffers. */
sfx_status sfx_problem_create(const sfx_problem_desc* desc, sfx_problem** out);
void sfx_problem_destroy(sfx_problem* p);
const char* sfx_last_error(const sfx_problem* p);

/* Optimizer::UpdateParams (symforce/opt/optimizer.h:259-263) */
sfx_status sfx_update_params(sfx_problem* p, const sfx_params* params);

/* Copy the caller's Values::data_ to the device (what nonlinear_solver.Reset(values) does on the
 * host, symforce/opt/internal/optimizer_utils.h:96). */
sfx_status sfx_set_values(sfx_problem* p, const double* values, int64_t n);

/* Optimizer::Optimize(values, num_iterations, ...) (symforce/opt/optimizer.tcc:79-89 ->
 * internal/optimizer_utils.h:36-104 -> LevenbergMarquardtSolver::Iterate): runs up to
 * num_iterations (<0: params.iterations) LM iterations with no host round trip per iteration. */
sfx_status sfx_optimize(sfx_problem* p, int32_t num_iterations, sfx_stats* stats);

/* values = nonlinear_solver.GetBestValues() (internal/optimizer_utils.h:69) */
sfx_status sfx_get_best_values(sfx_problem* p, double* values, int64_t n);

/* optimizer_params_t::debug_stats (lcmtypes/symforce.lcm:236-246, levenberg_marquardt_solver.tcc:166-171,245-250):
 * optimization_iteration_t::values (the data of the Values buffer, whose index the caller already has) and
 * ::residual of iteration record `record` (0 = the record of iteration -1) of the last sfx_optimize[_continue],
 * which must have run with debug_stats set.  Either output may be NULL.  Costs iterations x (Values + residual +
 * update) of device memory; single GPU. */
sfx_status sfx_get_iteration_debug(sfx_problem* p, int32_t record, double* values, double* residual);

/* optimization_iteration_t::update of the same record (levenberg_marquardt_solver.tcc:116, update_ of :217): the N
 * entries of the step this iteration tried, in the reference's tangent order; zeros for record 0 (the record of
 * iteration -1, where the reference leaves the vector empty). */
sfx_status sfx_get_iteration_update(sfx_problem* p, int32_t record, double* update);

/* optimization_iteration_t::jacobian_values of the same record (levenberg_marquardt_solver.tcc:120-121, :172-175;
 * JacobianValues of linearization.h:180): the nnz values of the Jacobian at the record's values, in the CSC order of
 * sfx_get_jacobian_pattern (= OptimizationStats::jacobian_sparsity).  The LM loop never forms J, so nothing is kept per
 * iteration: J is re-evaluated on request from the record's snapshot of the Values buffer by the kernel behind
 * sfx_linearize_jacobian (bit-identical to evaluating it inside the iteration).  Does not touch the optimizer state. */
sfx_status sfx_get_iteration_jacobian(sfx_problem* p, int32_t record, double* jacobian_values);

/* GncOptimizer::Optimize outer loop (symforce/opt/gnc_optimizer.h:53-130, OptimizeContinue :133-142):
 * sfx_optimize_continue = LevenbergMarquardtSolver::ResetState(values) (levenberg_marquardt_solver.h:178-183) with
 * the values last given to sfx_set_values + IterateToConvergence for up to num_iterations more iterations: lambda,
 * nu, the iteration counter and the iteration records continue from the preceding sfx_optimize[_continue] (stats
 * report the accumulated records).  sfx_relax_damping_to_initial = RelaxDampingToInitial (:171-174).  The caller
 * changes the convexity parameter in its Values and the early-exit threshold with sfx_update_params in between. */
sfx_status sfx_optimize_continue(sfx_problem* p, int32_t num_iterations, sfx_stats* stats);
sfx_status sfx_relax_damping_to_initial(sfx_problem* p);

/* Values::Update(index, other) (symforce/opt/values.cc:269-275) with other = GetBestValues(): overwrites only the
 * storage of the optimized keys in `values`, which must be the buffer last given to sfx_set_values (every other
 * entry of the best values is identical to it by construction: only optimized keys are retracted).  Moves
 * 8 * (optimized storage) bytes over PCIe instead of the whole buffer; *bytes_copied (may be NULL) reports it. */
sfx_status sfx_update_best_values(sfx_problem* p, double* values, int64_t n, int64_t* bytes_copied);

/* stats.iterations (symforce/opt/optimization_stats.h:30) */
sfx_status sfx_get_iterations(sfx_problem* p, sfx_iteration* buf, int32_t capacity, int32_t* n);

/* Problem dimensions: N = tangent dim of the state, M = residual dim, nnz = nonzeros of the
 * lower-triangular CSC Hessian (incl. explicit diagonal, linearizer.cc:173-178). */
sfx_status sfx_get_dims(sfx_problem* p, int32_t* N, int32_t* M, int64_t* nnz);

/* Sparsity of Linearization::hessian_lower (symforce/opt/linearization.h:70-73): CSC column
 * pointers [N+1] and row indices [nnz], identical to what setFromTriplets yields in the
 * reference (linearizer.cc:324). */
sfx_status sfx_get_hessian_pattern(sfx_problem* p, int32_t* outer, int32_t* inner);

/* Optimizer::Linearize(values) (symforce/opt/optimizer.h:177): linearize at the values last given
 * to sfx_set_values.  Any output may be NULL.  hessian_values is in the CSC order of
 * sfx_get_hessian_pattern. */
sfx_status sfx_linearize(sfx_problem* p, double* residual, double* rhs, double* hessian_values);

/* Linearization::jacobian with optimizer_params_t::include_jacobians (symforce/opt/linearization.h:58-60;
 * built by linearizer.cc:252-259, 297-313): the M x N Jacobian in CSC form, rows ascending in every column, one
 * entry per (residual row of a factor, tangent column of one of its optimized keys).  The LM loop itself never
 * forms J (it assembles J^T J and J^T r directly), so this is an export evaluated on request.
 * sfx_get_jacobian_pattern: nnz, column pointers [N+1], row indices [nnz]; any output may be NULL.
 * sfx_linearize_jacobian: values [nnz] at the values last given to sfx_set_values; leaves the optimizer state
 * untouched.  Single GPU. */
sfx_status sfx_get_jacobian_pattern(sfx_problem* p, int64_t* nnz, int32_t* outer, int32_t* inner);
sfx_status sfx_linearize_jacobian(sfx_problem* p, double* jacobian_values);

/* optimizer_params_t::check_derivatives (symforce/opt/optimizer.tcc:39, 261-272 -> internal::CheckDerivatives,
 * internal/derivative_checker.h:32-123) at the values last given to sfx_set_values: the linearization is compared with
 *   (0) the numerical Jacobian of the residual -- central differences in the tangent space with step sqrt(epsilon)
 *       (util.h:97-127), every perturbed residual evaluated by the device's retract + linearize -- tolerance
 *       10 sqrt(epsilon);  (1) hessian_lower against J^T J and (2) rhs against J^T r, tolerance sqrt(epsilon);
 * all three as Eigen's isApprox does (|x - y|_F <= tol * min(|x|_F, |y|_F)).  rel_errors[3] (may be NULL) receives the
 * three ratios, *ok whether all are within tolerance (the reference SYM_ASSERTs on it), numerical_jacobian (may be
 * NULL) the dense M x N column-major numerical Jacobian.  Relinearizes 2 N times and forms dense matrices on the host:
 * M * N <= 2^24 and N <= 4096, SFX_ERR_UNSUPPORTED beyond.  Resets the optimizer state like sfx_linearize.  Single GPU. */
sfx_status sfx_check_derivatives(sfx_problem* p, double* rel_errors, int32_t* ok, double* numerical_jacobian);

/* stats.best_linearization (populate_best_linearization, internal/optimizer_utils.h:71-76) */
sfx_status sfx_get_best_linearization(sfx_problem* p, double* residual, double* rhs,
                                      double* hessian_values);

/* Optimizer::ComputeCovariances(linearization, keys, ..., c_is_block_diagonal = true)
 * (symforce/opt/optimizer.tcc:177-199 -> internal/covariance_utils.h:124-147 ->
 * SparseSchurSolver::Factorize + SInvInPlace, sparse_schur_solver.tcc:101-138,165-170) for a problem
 * created with the Schur solver: `keys` are the keys in front of the eliminated landmarks,
 * block_dim = their tangent dimension, covariance = (B - E (C + eps I)^-1 E^T)^-1, column-major
 * block_dim x block_dim in keys_ order.  For a problem created with the Cholesky solver:
 * block_dim = N is Optimizer::ComputeFullCovariance / ComputeAllCovariances (optimizer.tcc:113-121, 201-206 ->
 * LevenbergMarquardtSolver::ComputeCovariance, levenberg_marquardt_solver.tcc:345-356): covariance = (H + eps I)^-1;
 * block_dim < N is ComputeCovariances with c_is_block_diagonal = false, C of any structure
 * (internal/covariance_utils.h:41-103, 142-145: S = B - E C^-1 E^T with eps on the diagonal of C, covariance = S^-1),
 * computed as the leading block of the inverse of the matrix damped on C, from block_dim solves with the sparse
 * Cholesky factor of the whole matrix.  hessian_values: Linearization::hessian_lower values
 * in the CSC order of sfx_get_hessian_pattern, or NULL for the best linearization of the last
 * sfx_optimize.  Other blocks of a Schur problem: SFX_ERR_UNSUPPORTED; a non-positive pivot: SFX_ERR_NUMERICAL. */
sfx_status sfx_compute_covariance(sfx_problem* p, const double* hessian_values, int32_t block_dim,
                                  double* covariance);

/* Parity hook for one linear solve: DampHessian + Factorize + Solve of
 * LevenbergMarquardtSolver::Iterate (levenberg_marquardt_solver.tcc:195-218) at the
 * linearization of the values last set; writes update = -H_damped^{-1} rhs in keys_ order. */
sfx_status sfx_solve_step(sfx_problem* p, double lambda, double* update);

/* stats.linear_solver_ordering under debug_stats (levenberg_marquardt_solver.tcc:209-212):
 * scalar permutation used by the linear solver (length = dimension of the factored system). */
sfx_status sfx_get_ordering(sfx_problem* p, int32_t* perm, int32_t capacity, int32_t* n);

/* Observability: per-phase device times of the last sfx_optimize (the SYM_TIME_SCOPE names of
 * levenberg_marquardt_solver.tcc:27,142,201,216,230,235 backed by CUDA events). */
sfx_status sfx_get_timings(sfx_problem* p, sfx_timings* out);

/* Structure report (host-side analysis results), for DESIGN/bench: fills up to `capacity`
 * int64 counters, see SFX_INFO_* indices. */
enum {
  SFX_INFO_N = 0,
  SFX_INFO_M = 1,
  SFX_INFO_NNZ_H = 2,
  SFX_INFO_NUM_NODES = 3,
  SFX_INFO_REDUCED_DIM = 4,     /* dimension of the factored system (S or H) */
  SFX_INFO_NNZ_L = 5,           /* scalar nonzeros of the supernodal factor */
  SFX_INFO_NUM_SUPERNODES = 6,
  SFX_INFO_NUM_LEVELS = 7,
  SFX_INFO_FACTOR_FLOPS = 8,
  SFX_INFO_S_BLOCKS = 9,
  SFX_INFO_SCHUR_PAIRS = 10,
  SFX_INFO_MAX_FRONT = 11,
  SFX_INFO_DEVICE_BYTES = 12,
  SFX_INFO_CHOL_FAILURES = 13,  /* iterations of the last sfx_optimize whose LLT met a non-positive pivot (step rejected) */
  SFX_INFO_NONFINITE_UPDATES = 14, /* iterations of the last sfx_optimize with a non-finite update vector */
  SFX_INFO_ZERO_DIAGONAL = 15,  /* debug_checks: damped diagonal entries below epsilon (CheckHessianDiagonal) */
  SFX_INFO_PLAN = 16,  /* front plan the library chose by modelled factorization time: -1 METIS_NodeND as the reference
                          orders (Eigen::MetisOrdering, sparse_cholesky_solver.tcc:13-30), 0 the same ordering with capped
                          amalgamation, d >= 2 nested dissection to depth d + a sweep of every subdomain */
  SFX_INFO_REF_ORDERING_FLOPS = 17, /* factorization flops of the METIS_NodeND plan (SFX_INFO_FACTOR_FLOPS: chosen plan) */
  SFX_INFO_COUNT = 18
};
sfx_status sfx_get_info(sfx_problem* p, int64_t* out, int32_t capacity);

/* Multi-GPU plumbing: one communicator per rank over NCCL (NVLink/NVSwitch). */
sfx_status sfx_comm_unique_id(char id_out[128]);
sfx_status sfx_comm_create(const char id[128], int32_t rank, int32_t world, int32_t device,
                           sfx_comm** out);
void sfx_comm_destroy(sfx_comm* c);

#ifdef __cplusplus
}
#endif
#endif /* SFX_H_ */
