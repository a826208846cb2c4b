/*
 * include/cddp_b200.h — C ABI of the B200-native batched CLDDP/iLQR engine.
 *
 * This is the drop-in boundary for the ONE hot path this repository accelerates: a DDP
 * iteration of cddp-cpp's CLDDP solver (backward Riccati sweep + forward rollout / line search),
 * batched over independent problem instances.  Everything below is plain C: pointers, sizes and
 * PODs — no C++/torch types.  The C++ shim that mirrors the reference API (cddp::CDDP,
 * ISolverAlgorithm, DynamicalSystem, ...) lives in cddp-cpp_b200/host/ and calls only these
 * entry points; INTEGRATION.md shows the binding a cddp-cpp maintainer would add.
 *
 * Citations are file:line in astomodynamics/cddp-cpp @ f71fa80 (v0.5.2).  The reference has no
 * FFI and no batch API; each entry point names the reference interface it
 * replaces for a batch of B instances.
 *
 * Conventions: row-major doubles; batch outermost.  x0,xref: [B][n]; ref_traj: [B][N+1][n];
 * X: [B][N+1][n]; U: [B][N][m]; K: [B][N][m][n]; k: [B][N][m].  Every function returns 0 on
 * success or a CDDP_B200_ERR_* code; cddp_b200_error_string() describes it.  Solve OUTCOMES are
 * never errors — they come back as per-instance status codes that map to the reference's
 * status_message strings (cddp_solver_base.cpp:69,82,162; clddp_solver.cpp:209,270,274).
 * There is no CPU fallback: if no CUDA device is usable every entry point that computes fails
 * with CDDP_B200_ERR_CUDA.
 */
#ifndef CDDP_B200_H
#define CDDP_B200_H

#ifndef __CUDACC_RTC__
#include <stddef.h>
#endif

#ifdef __cplusplus
extern "C" {
#endif

#if defined(__GNUC__)
#define CDDP_B200_API __attribute__((visibility("default")))
#else
#define CDDP_B200_API
#endif

#define CDDP_B200_ABI_VERSION 1
#define CDDP_B200_MAX_N 16      /* state dimension cap of the in-kernel dense blocks */
#define CDDP_B200_MAX_M 8       /* control dimension cap */
#define CDDP_B200_MAX_ALPHAS 32 /* line-search candidates evaluated in parallel (one lane each) */

/* error codes */
enum {
  CDDP_B200_OK = 0,
  CDDP_B200_ERR_INVALID_ARGUMENT = 1, /* null pointer, bad dimension, unsupported option */
  CDDP_B200_ERR_UNSUPPORTED_MODEL = 2, /* host-only DynamicalSystem: no device dynamics (no CPU fallback) */
  CDDP_B200_ERR_CUDA = 3,              /* CUDA runtime error / no device; see cddp_b200_last_cuda_error */
  CDDP_B200_ERR_OUT_OF_MEMORY = 4,
  CDDP_B200_ERR_STATE = 5,             /* call order violated (e.g. backward before linearize) */
  CDDP_B200_ERR_USER_MODEL = 6         /* the user-supplied dynamics source failed to compile; see cddp_b200_last_compile_log */
};

/* Device-resident dynamics models (src/dynamics_model/<name>.cpp).  model_params layout:
 *   PENDULUM  : length, mass, damping                                   (pendulum.cpp:24-27)
 *   CARTPOLE  : cart_mass, pole_mass, pole_length, gravity, damping     (cartpole.cpp:27-36)
 *   UNICYCLE  : (none)                                                  (unicycle.cpp:24-26)
 *   QUADROTOR : mass, inertia 3x3 row-major (9), arm_length             (quadrotor.cpp:25-31)
 *   LTI       : (none) — discrete A_d, B_d passed in lti_A, lti_B       (lti_system.cpp:71-92)
 *   USER      : free — passed to the user's device function as `p`      (see cddp_b200_create_ex)
 */
enum {
  CDDP_B200_MODEL_PENDULUM = 0,
  CDDP_B200_MODEL_CARTPOLE = 1,
  CDDP_B200_MODEL_UNICYCLE = 2,
  CDDP_B200_MODEL_QUADROTOR = 3,
  CDDP_B200_MODEL_LTI = 4,
  CDDP_B200_MODEL_USER = 5 /* dynamics supplied as CUDA source, compiled at create time (cddp_b200_create_ex) */
};

/* DynamicalSystem integration_type (dynamical_system.cpp:67-83) */
enum { CDDP_B200_EULER = 0, CDDP_B200_HEUN = 1, CDDP_B200_RK3 = 2, CDDP_B200_RK4 = 3 };

/* per-instance outcome -> CDDPSolution::status_message */
enum {
  CDDP_B200_STATUS_RUNNING = 0,        /* "Running" (only before/within a solve) */
  CDDP_B200_STATUS_OPTIMAL = 1,        /* "OptimalSolutionFound" */
  CDDP_B200_STATUS_ACCEPTABLE = 2,     /* "AcceptableSolutionFound" */
  CDDP_B200_STATUS_MAX_ITERATIONS = 3, /* "MaxIterationsReached" */
  CDDP_B200_STATUS_REG_LIMIT = 4,      /* "RegularizationLimitReached_NotConverged" */
  CDDP_B200_STATUS_MAX_CPU_TIME = 5    /* "MaxCpuTimeReached" */
};

/* The subset of cddp::CDDPOptions the CLDDP path reads (include/cddp-cpp/cddp_core/options.hpp:
 * :41-50 line search, :58-66 regularization, :93-105 filter.armijo_constant, :208-251 main;
 * boxqp.hpp:30-41 BoxQPOptions).  Defaults from cddp_b200_default_options() equal the
 * reference's member initialisers. */
typedef struct {
  double tolerance;                      /* 1e-5 */
  double acceptable_tolerance;           /* 1e-6 */
  int max_iterations;                    /* 1 */
  int enable_parallel;                   /* 0: first accepted alpha wins (default, cddp_solver_base.cpp:255-263);
                                            1: lowest cost among accepted alphas (enable_parallel=true, :264-285) */
  double max_cpu_time;                   /* seconds, 0 = unlimited (checked between batched iterations) */
  double termination_scaling_max_factor; /* 100 */
  int ls_max_iterations;                 /* line_search.max_iterations 11 */
  int reserved1;
  double ls_initial_step_size;           /* 1 */
  double ls_min_step_size;               /* 1e-8 */
  double ls_step_reduction_factor;       /* 0.5 */
  double reg_initial_value;              /* regularization.initial_value 1e-6 */
  double reg_update_factor;              /* 10 */
  double reg_max_value;                  /* 1e7 */
  double reg_min_value;                  /* 1e-10 */
  int qp_max_iterations;                 /* box_qp.max_iterations 100 */
  int reserved2;
  double qp_min_gradient_norm;           /* 1e-8 */
  double qp_min_relative_improvement;    /* 1e-8 */
  double qp_step_decrease_factor;        /* 0.6 */
  double qp_min_step_size;               /* 1e-22 */
  double qp_armijo_constant;             /* 0.1 */
  double armijo_constant;                /* filter.armijo_constant 1e-4 (clddp_solver.cpp:256) */
} cddp_b200_options;

/* Batch-shared problem definition = what a cddp::CDDP object holds besides per-instance data:
 * system (model id + parameters), QuadraticObjective weights (objective.cpp:30-64; Q and R are
 * passed UNscaled, the engine applies the reference's Q*dt, R*dt), horizon, timestep, and the
 * optional ControlConstraint bounds (constraint.hpp:144-251; looked up by the literal name
 * "ControlConstraint", clddp_solver.cpp:85-86). */
typedef struct {
  int model;
  int n;       /* state_dim */
  int m;       /* control_dim */
  int horizon; /* N */
  double dt;   /* timestep */
  int integrator;
  int has_control_box;
  double model_params[16];
  const double *lti_A; /* [n][n] (LTI only, else NULL) */
  const double *lti_B; /* [n][m] */
  const double *Q;     /* [n][n] */
  const double *R;     /* [m][m] */
  const double *Qf;    /* [n][n] */
  const double *lb;    /* [m] rawLowerBound (NULL if !has_control_box) */
  const double *ub;    /* [m] rawUpperBound */
} cddp_b200_problem;

/* ---- IPDDP: interior-point DDP with path inequality constraints (src/cddp_core/ipddp_solver.cpp) ----
 * One entry of the path-constraint set = one addPathConstraint(name, constraint) call (cddp_core.cpp:156-161).
 * Pass the entries in the order the reference iterates them: its std::map is keyed by constraint NAME, so
 * alphabetical ("BallConstraint" < "ControlConstraint" < "LinearConstraint" < "StateConstraint").
 *   CONTROL_BOX / STATE_BOX: p0 = lower [m|n], p1 = upper [m|n]; dual dim 2m | 2n   (constraint.hpp:144-251)
 *   BALL  : rows = dimension of the centre, p0 = centre [rows], p1 = &radius; dual dim 1  (constraint.hpp:320-440)
 *   LINEAR: rows, p0 = A [rows][n], p1 = b [rows]; dual dim rows; scale ignored as in the reference (:253-318) */
enum { CDDP_B200_CON_CONTROL_BOX = 0, CDDP_B200_CON_STATE_BOX = 1, CDDP_B200_CON_BALL = 2, CDDP_B200_CON_LINEAR = 3 };
typedef struct {
  int type;
  int rows;
  double scale; /* scale_factor */
  const double *p0;
  const double *p1;
} cddp_b200_constraint;

/* BarrierStrategy (options.hpp:28-33); MONOTONIC and IPOPT share one update rule in IPDDP (ipddp_solver.cpp:2601-2613) */
enum { CDDP_B200_BARRIER_ADAPTIVE = 0, CDDP_B200_BARRIER_MONOTONIC = 1, CDDP_B200_BARRIER_IPOPT = 2 };

/* cddp::CDDPOptions::ipddp + ::filter (options.hpp:75-104, :148-186); defaults from
 * cddp_b200_ipddp_default_options() equal the reference's member initialisers.  filter.armijo_constant is
 * cddp_b200_options::armijo_constant. */
typedef struct {
  double dual_var_init_scale;            /* 1e-1 */
  double slack_var_init_scale;           /* 1e-2 */
  double barrier_tol_mult;               /* 0.1 */
  double barrier_update_dual_weight;     /* 0.01 */
  double mu_kappa_epsilon;               /* 10 */
  double theta_0_floor;                  /* 1 */
  double mu_initial;                     /* barrier.mu_initial 1 */
  double mu_min_value;                   /* 1e-10 */
  double mu_update_factor;               /* 0.5 */
  double mu_update_power;                /* 1.2 */
  double min_fraction_to_boundary;       /* 0.99 */
  double merit_acceptance_threshold;     /* filter 1e-6 */
  double violation_acceptance_threshold; /* 1e-6 */
  double max_violation_threshold;        /* 1e4 */
  double min_violation_for_armijo_check; /* 1e-7 */
  double jacobian_regularization_value;    /* ipddp 1e-8: reduced-system floor of the terminal-equality solve (options.hpp:181-184) */
  double jacobian_regularization_exponent; /* 0.25 */
  int theta_norm_l2;                     /* theta_norm: 0 = "l1" (default), 1 = "l2" */
  int max_filter_size;                   /* 5 (at most 7) */
  int barrier_strategy;                  /* CDDP_B200_BARRIER_ADAPTIVE */
  int terminal_equality;                 /* 1 = addTerminalConstraint("TerminalEqualityConstraint", TerminalEqualityConstraint(
                                            reference state)) (terminal_constraint.hpp:61-117): h(x_N) = x_N - xref[b] = 0, solved by
                                            the terminal-equality branch (ipddp_solver.cpp:1120-1353, :484-639) */
} cddp_b200_ipddp_options;

/* CUDA-event timings accumulated since the last cddp_b200_reset_timing(); one launch of each
 * kernel per batched DDP iteration. */
typedef struct {
  double linearize_ms;
  double backward_ms;
  double forward_ms;
  long long linearize_launches;
  long long backward_launches;
  long long forward_launches;
  long long other_launches; /* init / cost / bookkeeping kernels */
} cddp_b200_timing;

typedef struct cddp_b200_solver cddp_b200_solver; /* opaque: owns all device buffers */

/* ---- library ---- */
CDDP_B200_API int cddp_b200_abi_version(void);
CDDP_B200_API const char *cddp_b200_error_string(int err);
CDDP_B200_API const char *cddp_b200_last_cuda_error(void);
CDDP_B200_API const char *cddp_b200_status_string(int status);
CDDP_B200_API int cddp_b200_device_count(int *count);

/* CDDPOptions() defaults (options.hpp) and detail::buildLineSearchAlphas
 * (cddp_context_utils.cpp:37-57).  Returns the number of alphas via *count. */
CDDP_B200_API void cddp_b200_default_options(cddp_b200_options *opts);
CDDP_B200_API int cddp_b200_build_alphas(const cddp_b200_options *opts, double *alphas, int capacity, int *count);

/* ---- lifetime: replaces constructing B cddp::CDDP objects (cddp_core.cpp:37-61) plus
 * addPathConstraint("ControlConstraint", ...) (cddp_core.cpp:148-153) ---- */
CDDP_B200_API int cddp_b200_create(const cddp_b200_problem *problem, const cddp_b200_options *opts, int batch, int device,
                     cddp_b200_solver **out);
CDDP_B200_API int cddp_b200_destroy(cddp_b200_solver *s);
/* ---- user-defined dynamics: the device counterpart of subclassing cddp::DynamicalSystem
 * (include/cddp-cpp/cddp_core/dynamical_system.hpp:33-152).  The reference's virtuals (getContinuousDynamics,
 * getContinuousDynamicsAutodiff, getStateJacobian, ...) are host-only Eigen calls that cannot run inside a kernel, so a
 * plugin model hands over CUDA C++ source instead: problem->model = CDDP_B200_MODEL_USER, any n <= CDDP_B200_MAX_N,
 * m <= CDDP_B200_MAX_M, and `model_source` defining at global scope
 *
 *     template <class T>
 *     __device__ void cddp_user_dynamics(const T *x, const T *u, const double *p, T *xdot);   // p = model_params
 *
 * — ONE text, instantiated with T = double for the rollouts (getContinuousDynamics) and with a forward-mode dual number
 * for the Jacobians, which is how the reference itself obtains them (autodiff of getContinuousDynamicsAutodiff,
 * src/cddp_core/dynamical_system.cpp:102-133).  Optionally `#define CDDP_USER_HAS_JACOBIAN` and
 *     __device__ void cddp_user_jacobian(const double *x, const double *u, const double *p, double *Fx, double *Fu);
 * (row-major continuous-time Jacobians [n][n], [n][m]; = overriding getStateJacobian / getControlJacobian).
 * The source is compiled for sm_100a with NVRTC together with the engine's own rollout / linearisation kernels (the same
 * kernel text the built-in models use); integrators, costs, line search, constraints are the engine's.  A source that
 * does not compile is CDDP_B200_ERR_USER_MODEL with the compiler output in cddp_b200_last_compile_log().
 * model_source == NULL makes both _ex entry points identical to the plain ones. ---- */
CDDP_B200_API int cddp_b200_create_ex(const cddp_b200_problem *problem, const cddp_b200_options *opts, const char *model_source,
                                      int batch, int device, cddp_b200_solver **out);
/* compile only — needs libnvrtc but no GPU/driver: validates a model source for dimensions (n, m) */
CDDP_B200_API int cddp_b200_compile_user_model(const char *model_source, int n, int m, size_t *cubin_bytes);
CDDP_B200_API const char *cddp_b200_last_compile_log(void);

/* kernels are launched on this cudaStream_t (default: the legacy default stream) */
CDDP_B200_API int cddp_b200_set_stream(cddp_b200_solver *s, void *cuda_stream);
/* CDDP::setOptions (cddp_core.cpp:109-113): rebuilds the alpha schedule */
CDDP_B200_API int cddp_b200_set_options(cddp_b200_solver *s, const cddp_b200_options *opts);

/* HBM layout of the per-timestep linearisation records the backward sweep streams (no reference
 * counterpart: the reference recomputes A = I + dt*Fx, B = dt*Fu inside the loop, clddp_solver.cpp:113-118).
 * DENSE = stacked n*n + n*m Jacobians (any model; required for cddp_b200_set_linearization, which
 * switches to it).  STRUCTURED = only the structural non-zeros of the built-in model's Jacobians
 * (default when the model has a pattern; the sweep kernel unrolls over the pattern). */
enum { CDDP_B200_RECORDS_DENSE = 0, CDDP_B200_RECORDS_STRUCTURED = 1 };
CDDP_B200_API int cddp_b200_set_record_layout(cddp_b200_solver *s, int layout);
CDDP_B200_API int cddp_b200_get_record_layout(cddp_b200_solver *s, int *layout, int *record_bytes);

/* ---- per-instance data: CDDP::setInitialState / setReferenceState(s) / setInitialTrajectory
 * (cddp_core.cpp:68-100,127-142).  Host pointers; copied H2D on the solver's stream.
 * ref_traj may be NULL (single reference state, objective.cpp:84-88).  X0 may be NULL: the state
 * trajectory is then zero except X[0] = x0 (ensureTrajectoryShape, cddp_context_utils.cpp:97-107).
 * U0 may be NULL (zeros). ---- */
CDDP_B200_API int cddp_b200_set_instances(cddp_b200_solver *s, const double *x0, const double *xref, const double *ref_traj,
                            const double *X0, const double *U0);
/* same, device pointers (device-to-device; used to time the path with inputs resident in HBM) */
CDDP_B200_API int cddp_b200_set_instances_device(cddp_b200_solver *s, const double *x0, const double *xref, const double *ref_traj,
                                   const double *X0, const double *U0);

/* ---- solver steps (the reference's template-method hooks, one call = whole batch) ---- */
/* initializeProblemIfNecessary (cddp_core.cpp:272-306) + CLDDPSolver::initialize cold start
 * (clddp_solver.cpp:28-75): X[0]=x0, gains zeroed, cost from the given X,U, reg=initial. */
CDDP_B200_API int cddp_b200_initialize(cddp_b200_solver *s);
/* A_t = I + dt*Fx, B_t = dt*Fu and cost gradients for every t (clddp_solver.cpp:113-122) */
CDDP_B200_API int cddp_b200_linearize(cddp_b200_solver *s);
/* one CLDDPSolver::backwardPass per instance at its current regularization, no retry
 * (clddp_solver.cpp:79-204); per-instance success flag readable via cddp_b200_get_sweep */
CDDP_B200_API int cddp_b200_backward_pass(cddp_b200_solver *s);
/* CDDPSolverBase::performForwardPass, sequential semantics = first accepted alpha
 * (cddp_solver_base.cpp:248-263) with every alpha rolled out in parallel; does not apply */
CDDP_B200_API int cddp_b200_forward_pass(cddp_b200_solver *s);
/* `iterations` passes of the CDDPSolverBase::solve loop body (cddp_solver_base.cpp:74-154) for
 * every instance still running: backward with regularization retry, early convergence test,
 * line search, accept / reject, regularization schedule, convergence test. */
CDDP_B200_API int cddp_b200_iterate(cddp_b200_solver *s, int iterations);
/* CDDP::solve("CLDDP") (cddp_core.cpp:235-270): initialize + iterate until every instance has
 * stopped or max_iterations; instances still running are marked MaxIterationsReached. */
CDDP_B200_API int cddp_b200_solve(cddp_b200_solver *s);
CDDP_B200_API int cddp_b200_num_running(cddp_b200_solver *s, int *running);
CDDP_B200_API int cddp_b200_synchronize(cddp_b200_solver *s);

/* ---- results: CDDPSolution fields (cddp_core.hpp:54-103) per instance.  Any pointer may be
 * NULL.  Synchronises the stream. ---- */
CDDP_B200_API int cddp_b200_get_solution(cddp_b200_solver *s, double *X, double *U, double *K, double *final_objective,
                           int *iterations_completed, int *status, double *final_step_length,
                           double *final_regularization, double *inf_du);
/* Asynchronous serving (no reference counterpart; the reference's stated use is MPC, README.md:11): with
 * cddp_b200_set_poll_interval(s, 0) cddp_b200_solve only ENQUEUES work on the solver's stream (no host
 * synchronisation; instances that finish early are masked on the device), and cddp_b200_get_solution_async enqueues
 * the device-to-host copies without waiting — call cddp_b200_synchronize before reading the host buffers.  Host
 * buffers must be pinned and stay valid until then.  Two or three handles on their own streams give a pipeline in
 * which call k+1's upload + solve overlaps call k's download.  interval -1 = automatic (every iteration from the second
 * on when batch x horizon >= 32768, else at iterations 4, 8, 12, 16, 24, ...), k > 0 = poll every k iterations.  At every
 * poll the still-running instances are compacted into the work list of the per-iteration kernels, so the launches that
 * follow cover only those (a batch whose instances converge at different iterations stops paying for the finished ones). */
CDDP_B200_API int cddp_b200_set_poll_interval(cddp_b200_solver *s, int interval);
/* Windowed line search (default OFF; sequential rule with 9..16 candidates only).  performForwardPass takes the FIRST
 * accepted alpha (cddp_solver_base.cpp:255-263), so the batched rollout can look at alphas_[0..7] with 8 lanes per
 * trajectory first and run the full width only for the instances none of them settled.  Decisions and trajectories are
 * identical with and without it (tests).  Measured on B200 for the headline batch it is SLOWER (0.464 ms against
 * 0.418 ms): the rollout is bound by the latency of its 100 dependent RK4 steps, not by FP64 throughput, and half the
 * warps hide less of it.  Kept as an option for throughput-bound models. */
CDDP_B200_API int cddp_b200_set_line_search_window(cddp_b200_solver *s, int enable);
/* User-model handles, sequential rule: try alphas_[0] on one lane per instance before the 16-wide line search (same
 * decisions and trajectories).  mode 1 = always, 0 = never, -1 (default) = only while the full search is throughput-bound
 * (it saves 15 of 16 rollouts per settled instance but adds one rollout of latency). */
CDDP_B200_API int cddp_b200_set_first_alpha_speculation(cddp_b200_solver *s, int mode);
/* Fused linearisation (default off; CLDDP, quadrotor, structured records): A = I + dt Fx, B = dt Fu are formed inside the
 * backward sweep, as CLDDPSolver::backwardPass does (clddp_solver.cpp:113-118), by the sweep's otherwise idle QP warp, and
 * cddp_b200_iterate / _solve skip the separate linearisation launch.  Results are identical; measured on B200 it is slower
 * (iteration 1.061 -> 1.114 ms at the headline batch), hence opt-in. */
CDDP_B200_API int cddp_b200_set_fused_linearization(cddp_b200_solver *s, int enable);
CDDP_B200_API int cddp_b200_get_solution_async(cddp_b200_solver *s, double *X, double *U, double *K, double *final_objective,
                                 int *iterations_completed, int *status, double *final_step_length,
                                 double *final_regularization, double *inf_du);
/* Receding-horizon (MPC) loop on a persistent handle — the reference's stated use case (README.md:11), where it
 * re-creates the solver and re-sends everything on every CDDP::solve (cddp_core.cpp:241) with options.warm_start
 * keeping X_/U_ (cddp_core.cpp:287-292).  cddp_b200_mpc_advance shifts every instance's nominal trajectory left by
 * `steps` (X[t] <- X[t+steps], U[t] <- U[t+steps], tail = last state / last control) ON THE DEVICE and uploads only
 * the new measured initial states x0_new [B][n] (and optionally new references); the next cddp_b200_solve then starts
 * from the shifted trajectory with X[0] = x0_new, zero gains and the initial regularisation, exactly like a
 * warm-started CDDP::solve.  cddp_b200_get_first_controls_async returns U[0] [B][m] (+ cost, status), the only
 * thing an MPC loop needs back; it does not synchronise. */
CDDP_B200_API int cddp_b200_mpc_advance(cddp_b200_solver *s, int steps, const double *x0_new, const double *xref_new);
CDDP_B200_API int cddp_b200_get_first_controls_async(cddp_b200_solver *s, double *u0, double *final_objective, int *status);
/* optional History (cddp_core.hpp:77-102) when recorded: [B][max_iterations+1][4] =
 * {objective, step_length_primal, dual_infeasibility, regularization}; lens [B] */
CDDP_B200_API int cddp_b200_enable_history(cddp_b200_solver *s, int enable);
CDDP_B200_API int cddp_b200_get_history(cddp_b200_solver *s, double *history, int *lens);
/* Decision trace (audit / parity instrumentation, no reference counterpart; CLDDP and IPDDP handles).  One int per entry of the
 * main loop of CDDPSolverBase::solve (cddp_solver_base.cpp:74) and instance, [B][cap], cap = max_iterations at enable
 * time: (backward-pass failures of this iteration << 8) | code, code = 1 + index of the accepted alpha
 * (performForwardPass, :248-263), 0 = no alpha accepted (handleForwardPassFailure, :206-218), 0xff = early convergence
 * exit (clddp_solver.cpp:206-213), 0xfe = regularisation limit during the backward retry (:95-109).  The parity tests
 * feed it to the CPU oracle, which then follows the same decision sequence. */
CDDP_B200_API int cddp_b200_enable_trace(cddp_b200_solver *s, int enable);
CDDP_B200_API int cddp_b200_get_trace(cddp_b200_solver *s, int *trace, int *cap);

/* ---- white-box access for parity tests (the reference's tests use a friend shim the same way,
 * tests/cddp_core/test_ipddp_solver.cpp:30-134) ---- */
CDDP_B200_API int cddp_b200_get_feedforward(cddp_b200_solver *s, double *k);                  /* k_u_ */
CDDP_B200_API int cddp_b200_set_gains(cddp_b200_solver *s, const double *K, const double *k); /* warm start k_u_/K_u_ */
CDDP_B200_API int cddp_b200_set_regularization(cddp_b200_solver *s, const double *reg);       /* [B] */
CDDP_B200_API int cddp_b200_set_cost(cddp_b200_solver *s, const double *cost);                /* [B] cost_ */
CDDP_B200_API int cddp_b200_get_linearization(cddp_b200_solver *s, double *A, double *B);     /* [B][N][n][n], [B][N][n][m] */
CDDP_B200_API int cddp_b200_set_linearization(cddp_b200_solver *s, const double *A, const double *B);
/* after backward_pass: dV [B][2], ok [B], inf_du [B], Vx0 [B][n], Vxx0 [B][n][n] (value function at t=0) */
CDDP_B200_API int cddp_b200_get_sweep(cddp_b200_solver *s, double *dV, int *ok, double *inf_du, double *Vx0, double *Vxx0);
/* after forward_pass: costs [B][num_alphas], accepted alpha index [B] (-1 = none), candidate X,U of
 * the accepted alpha (unchanged nominal if none) */
CDDP_B200_API int cddp_b200_get_forward(cddp_b200_solver *s, double *costs, int *accepted, double *Xnew, double *Unew);

/* ---- IPDDP handle: replaces B x { CDDP::addPathConstraint(...); CDDP::solve("IPDDP") } (cddp_core.cpp:156-161,
 * :235-270; IPDDPSolver, src/cddp_core/ipddp_solver.cpp).  Scope: cold start (options.warm_start = false),
 * use_ilqr = true, path inequality constraints of the four kinds above, optionally a TerminalEqualityConstraint on the
 * reference state (ipddp_opts->terminal_equality; no terminal INequalities), built-in models except LTI; problem->has_control_box is ignored (a ControlConstraint is an entry of `constraints`).  The handle is
 * driven by the same entry points as a CLDDP handle: cddp_b200_set_instances (X0 is ignored: IPDDP re-rolls the
 * state trajectory out from the given controls, ipddp_solver.cpp:876-882), cddp_b200_initialize / _linearize /
 * _backward_pass / _forward_pass / _iterate / _solve, cddp_b200_get_solution (inf_du = unscaled max |Q_u|). ---- */
CDDP_B200_API void cddp_b200_ipddp_default_options(cddp_b200_ipddp_options *opts);
CDDP_B200_API int cddp_b200_ipddp_create(const cddp_b200_problem *problem, const cddp_b200_options *opts,
                                         const cddp_b200_ipddp_options *ipddp_opts, const cddp_b200_constraint *constraints,
                                         int num_constraints, int batch, int device, cddp_b200_solver **out);
/* same with a user-defined dynamics model (see cddp_b200_create_ex) */
CDDP_B200_API int cddp_b200_ipddp_create_ex(const cddp_b200_problem *problem, const cddp_b200_options *opts,
                                            const cddp_b200_ipddp_options *ipddp_opts, const cddp_b200_constraint *constraints,
                                            int num_constraints, const char *model_source, int batch, int device,
                                            cddp_b200_solver **out);
/* total dual dimension d of the handle's constraint set (getTotalDualDim, ipddp_solver.cpp:2134-2143); 0 for CLDDP handles */
CDDP_B200_API int cddp_b200_ipddp_dual_dim(cddp_b200_solver *s, int *d);
/* CDDPSolution interior-point fields (cddp_core.hpp:54-103; populateSolverSpecificSolution, ipddp_solver.cpp:2090-2097) and
 * the dual / slack / constraint-value trajectories Y, S, G [B][N][d].  scalars [B][8] = final_barrier_mu, merit,
 * inf_pr, inf_comp, step_norm, alpha_du, alpha_pr_max, alpha_du_max.  Any pointer may be NULL. */
CDDP_B200_API int cddp_b200_ipddp_get_solution(cddp_b200_solver *s, double *Y, double *S, double *G, double *scalars);
/* white box: slack/dual gains of the last backward pass, k_y,k_s [B][N][d], K_y,K_s [B][N][d][n]; line-search table
 * [B][num_alphas][4] = accepted?, cost, barrier merit, theta of every alpha of the last forward pass */
/* white box (parity tests): what IPDDPSolver carries from one iteration to the next besides the trajectories —
 * Lambda_T_eq_ [B][n] (zeros without a terminal equality), the filter points [B][8][2] (merit, theta) and their count [B]
 * (ipddp_solver.hpp filter_), scalars [B][3] = {filter theta of the nominal, sum log s of the nominal, Lambda_T . h_T} */
CDDP_B200_API int cddp_b200_ipddp_get_iteration_state(cddp_b200_solver *s, double *lamT, double *filter, int *filter_size, double *scalars);
CDDP_B200_API int cddp_b200_ipddp_get_gains(cddp_b200_solver *s, double *ky, double *Ky, double *ks, double *Ks);
CDDP_B200_API int cddp_b200_ipddp_get_line_search(cddp_b200_solver *s, double *table);
/* History with interior-point columns: [B][max_iterations+1][9] = objective, merit, alpha_pr, alpha_du, inf_du, inf_pr,
 * inf_comp, regularization, barrier_mu (cddp_b200_enable_history must have been called); lens [B] */
CDDP_B200_API int cddp_b200_ipddp_get_history(cddp_b200_solver *s, double *history, int *lens);

/* ---- measurement ---- */
CDDP_B200_API int cddp_b200_reset_timing(cddp_b200_solver *s);
CDDP_B200_API int cddp_b200_get_timing(cddp_b200_solver *s, cddp_b200_timing *t);
CDDP_B200_API int cddp_b200_enable_timing(cddp_b200_solver *s, int enable);
/* algorithmic HBM bytes of one backward sweep of the whole batch: 8*(n^2+2nm+n+3m)*N*B */
CDDP_B200_API int cddp_b200_backward_algorithmic_bytes(cddp_b200_solver *s, double *bytes);

/* ---- one-shot convenience: the call a user of the batched API makes.  Host buffers in, host
 * buffers out; H2D and D2H copies are inside.  X,U are in/out. ---- */
CDDP_B200_API int cddp_b200_solve_host(const cddp_b200_problem *problem, const cddp_b200_options *opts, int batch, int device,
                         const double *x0, const double *xref, const double *ref_traj, double *X, double *U, double *K,
                         double *final_objective, int *iterations_completed, int *status, double *final_step_length,
                         double *final_regularization, double *inf_du);

#ifdef __cplusplus
}
#endif
#endif /* CDDP_B200_H */
