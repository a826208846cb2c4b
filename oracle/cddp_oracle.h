/*
 * oracle/cddp_oracle.h — C API of the CPU oracle.
 *
 * TEST INFRASTRUCTURE ONLY.  This is a plain-C++17 CPU restatement of the reference's
 * CLDDP hot path (astomodynamics/cddp-cpp @ f71fa80, v0.5.2).  Only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load it.
 * The product (cddp-cpp_b200/) never includes, links or calls anything in this directory.
 *
 * PARITY STATUS: "parity unpinned".  The reference cannot be compiled in this image
 * (Eigen 3.4.0 and autodiff v1.1.2 are fetched by CMake FetchContent, CMakeLists.txt:65-97,
 * :120-125; no network), and the reference's own tests pin no gains / trajectories / costs on
 * this path.  The oracle is therefore pinned by (i) the closed-form
 * known-answer tests the reference does hold (test_objective.cpp, test_constraint.cpp,
 * test_finite_difference.cpp, test_quadrotor.cpp hover + Jacobian-vs-FD), (ii) an independent
 * numpy restatement (oracle/np_oracle.py) that must agree with it, and (iii) LQR closed forms.
 *
 * All matrices are row-major doubles.  Trajectories are [t][dim].
 */
#ifndef CDDP_ORACLE_H
#define CDDP_ORACLE_H

#ifdef __cplusplus
extern "C" {
#endif

#define ORACLE_MAX_N 16
#define ORACLE_MAX_M 8
#define ORACLE_MAX_ALPHAS 64

/* model ids (src/dynamics_model/<name>.cpp) */
enum { ORACLE_PENDULUM = 0, ORACLE_CARTPOLE = 1, ORACLE_UNICYCLE = 2, ORACLE_QUADROTOR = 3, ORACLE_LTI = 4,
       ORACLE_BICYCLE = 5 /* bicycle.cpp */, ORACLE_CHAIN7 = 6 /* test plugin model, not in the reference */,
       ORACLE_MANIP7 = 7 /* 7-DOF generalisation of manipulator.cpp (BASELINE config #5) */ };
/* integrators (src/cddp_core/dynamical_system.cpp:28-83) */
enum { ORACLE_EULER = 0, ORACLE_HEUN = 1, ORACLE_RK3 = 2, ORACLE_RK4 = 3 };
/* BoxQP status (include/cddp-cpp/cddp_core/boxqp.hpp:47-55) */
enum {
  ORACLE_QP_HESSIAN_NOT_PD = -1, ORACLE_QP_NO_DESCENT = 0, ORACLE_QP_MAX_ITER_EXCEEDED = 1,
  ORACLE_QP_MAX_LS_EXCEEDED = 2, ORACLE_QP_NO_BOUNDS = 3, ORACLE_QP_SUCCESS = 4, ORACLE_QP_ALL_CLAMPED = 5
};
/* solve status -> reference status_message strings (cddp_solver_base.cpp:69,82,162; clddp_solver.cpp:209,270,274) */
enum {
  ORACLE_RUNNING = 0, ORACLE_OPTIMAL = 1, ORACLE_ACCEPTABLE = 2, ORACLE_MAX_ITERATIONS = 3,
  ORACLE_REG_LIMIT = 4, ORACLE_MAX_CPU_TIME = 5
};

/* Field-for-field the same layout as cddp_b200_options (include/cddp_b200.h) so that the tests
 * can feed the oracle and the product from one ctypes Structure.  Defaults: options.hpp:41-66,
 * :93-105, :208-251; boxqp.hpp:30-41. */
typedef struct {
  double tolerance;                       /* 1e-5 */
  double acceptable_tolerance;            /* 1e-6 */
  int max_iterations;                     /* 1 */
  int enable_parallel;                    /* 0 */
  double max_cpu_time;                    /* 0 = unlimited */
  double termination_scaling_max_factor;  /* 100 */
  int ls_max_iterations;                  /* 11 */
  int reserved1;
  double ls_initial_step_size;            /* 1 */
  double ls_min_step_size;                /* 1e-8 */
  double ls_step_reduction_factor;        /* 0.5 */
  double reg_initial_value;               /* 1e-6 */
  double reg_update_factor;               /* 10 */
  double reg_max_value;                   /* 1e7 */
  double reg_min_value;                   /* 1e-10 */
  int qp_max_iterations;                  /* 100 */
  int reserved2;
  double qp_min_gradient_norm;            /* 1e-8 */
  double qp_min_relative_improvement;     /* 1e-8 */
  double qp_step_decrease_factor;         /* 0.6 */
  double qp_min_step_size;                /* 1e-22 */
  double qp_armijo_constant;              /* 0.1 */
  double armijo_constant;                 /* filter.armijo_constant 1e-4 */
} oracle_options;

/* Batch-shared problem description; same layout as cddp_b200_problem. */
typedef struct {
  int model;
  int n;
  int m;
  int horizon;
  double dt;
  int integrator;
  int has_control_box;
  double model_params[16];
  const double *lti_A; /* n x n discrete A_d (LTI only) */
  const double *lti_B; /* n x m discrete B_d (LTI only) */
  const double *Q;     /* n x n, UNscaled: the objective multiplies by dt (objective.cpp:38-39) */
  const double *R;     /* m x m, UNscaled */
  const double *Qf;    /* n x n */
  const double *lb;    /* m, ControlConstraint raw lower bound (constraint.hpp:222) */
  const double *ub;    /* m */
} oracle_problem;

void oracle_default_options(oracle_options *o);
int oracle_build_alphas(const oracle_options *o, double *alphas /* [ORACLE_MAX_ALPHAS] */);

/* L2 plugin surface restated: dynamics, integrators, Jacobians, objective */
void oracle_continuous_dynamics(const oracle_problem *p, const double *x, const double *u, double t, double *xdot);
void oracle_discrete_dynamics(const oracle_problem *p, const double *x, const double *u, double t, double *xnext);
void oracle_jacobians(const oracle_problem *p, const double *x, const double *u, double t, double *Fx, double *Fu);
double oracle_running_cost(const oracle_problem *p, const double *x, const double *u, const double *ref);
double oracle_terminal_cost(const oracle_problem *p, const double *x, const double *ref);
double oracle_trajectory_cost(const oracle_problem *p, const double *X, const double *U, const double *xref,
                              const double *ref_traj);

/* BoxQPSolver::solve (boxqp.cpp:25-182).  Kfree_rhs/Kfree_out (optional, n x ncols): on return
 * Kfree_out[free rows] = Hfree.solve(Kfree_rhs[free rows]) with the factor the solver ended with. */
int oracle_boxqp(const oracle_options *o, int n, const double *H, const double *g, const double *lower,
                 const double *upper, const double *x0 /* may be NULL */, double *x, int *free_mask, int *iterations,
                 int *factorizations, double *final_value, double *final_grad_norm, int ncols, const double *Kfree_rhs,
                 double *Kfree_out);

/* CLDDPSolver::backwardPass (clddp_solver.cpp:79-204).  k is in/out (warm start + result).
 * Vx_dbg [N+1][n], Vxx_dbg [N+1][n*n] optional.  Returns 1 on success, 0 on failure;
 * fail_t (optional) receives the failing timestep. */
int oracle_backward_pass(const oracle_problem *p, const oracle_options *o, const double *X, const double *U,
                         const double *xref, const double *ref_traj, double reg, double *K, double *k, double *dV,
                         double *inf_du, double *Vx_dbg, double *Vxx_dbg, int *fail_t);

/* Same sweep on caller-supplied stacked discrete Jacobians A[t] (n x n), B[t] (n x m). */
int oracle_backward_pass_AB(const oracle_problem *p, const oracle_options *o, const double *A, const double *B,
                            const double *X, const double *U, const double *xref, const double *ref_traj, double reg,
                            double *K, double *k, double *dV, double *inf_du, double *Vx_dbg, double *Vxx_dbg,
                            int *fail_t);

/* clddp_solver.cpp:113-118: A = I + dt*Fx, B = dt*Fu for every t. */
void oracle_linearize(const oracle_problem *p, const double *X, const double *U, double *A, double *B);

/* CLDDPSolver::forwardPass (clddp_solver.cpp:215-262). Returns success flag. */
int oracle_forward_pass(const oracle_problem *p, const oracle_options *o, const double *x0, const double *X,
                        const double *U, const double *xref, const double *ref_traj, const double *K, const double *k,
                        const double *dV, double cost, double alpha, double *Xn, double *Un, double *Jn);

/* CDDP::solve("CLDDP") for one instance (cddp_core.cpp:235-270 + cddp_solver_base.cpp:29-186).
 * X,U in/out.  history (optional) [max_iterations+1][4] = {objective, alpha, inf_du, reg}. */
typedef struct {
  double final_objective;
  double final_step_length;
  double final_regularization;
  double inf_du;
  int iterations;
  int status;
  int history_len;
  int reserved;
} oracle_result;

void oracle_solve(const oracle_problem *p, const oracle_options *o, const double *x0, const double *xref,
                  const double *ref_traj, double *X, double *U, double *K, double *k, oracle_result *res,
                  double *history);

/* Batch over independent instances, statically partitioned over nthreads std::threads.
 * x0 [B][n], xref [B][n], ref_traj [B][N+1][n] or NULL, X [B][N+1][n], U [B][N][m], K [B][N][m][n], k [B][N][m]. */
void oracle_solve_batch(const oracle_problem *p, const oracle_options *o, int batch, int nthreads, const double *x0,
                        const double *xref, const double *ref_traj, double *X, double *U, double *K, double *k,
                        oracle_result *res);

/* Decision trace / decision replay (test instrumentation; DESIGN.md "Parity procedure").
 * trace: [B][max_iterations] ints, one per entry of the main loop (cddp_solver_base.cpp:74):
 *   (backward-pass failures in this iteration << 8) | code,
 *   code = 1 + index of the accepted alpha, or one of ORACLE_TRACE_*.
 * oracle_solve_batch_traced with replay_trace == NULL runs the normal solve and (trace_out != NULL) records its
 * decisions.  With replay_trace / replay_iterations / replay_status (e.g. downloaded from the CUDA path) the oracle's
 * arithmetic FOLLOWS those decisions instead of taking its own: accept / reject of every line-search candidate
 * (clddp_solver.cpp:251-257), backward failure + regularisation bump (cddp_solver_base.cpp:93-111), early and late
 * convergence (clddp_solver.cpp:206-213,264-277).  rep[b] says how often its own verdict differed and the largest
 * relative distance to the decision threshold among those decisions.  history: [B][max_iterations+1][4] or NULL. */
enum { ORACLE_TRACE_LS_FAILED = 0, ORACLE_TRACE_EARLY_EXIT = 0xff, ORACLE_TRACE_BW_LIMIT = 0xfe };
typedef struct {
  int n_disagree;          /* threshold decisions (Armijo ratio, inf_du < tolerance, dJ < acceptable_tolerance) */
  int n_backward_disagree; /* backward-pass success / failure verdicts */
  int infeasible;          /* the recorded sequence could not be followed (own backward pass failed where it must succeed) */
  int reserved;
  double max_margin;       /* max over the n_disagree decisions of the distance to the threshold, in the units roundoff
                              enters the test: |dJ - c expected| / |cost| (Armijo), |inf_du - tol| / tol */
} oracle_replay_report;

void oracle_solve_batch_traced(const oracle_problem *p, const oracle_options *o, int batch, int nthreads, const double *x0,
                               const double *xref, const double *ref_traj, double *X, double *U, double *K, double *k,
                               oracle_result *res, double *history, int *trace_out, const int *replay_trace,
                               const int *replay_iterations, const int *replay_status, oracle_replay_report *rep);

/* ONE entry of the main loop (cddp_solver_base.cpp:74-170) for every instance of a batch, from a caller-supplied solver
 * state: X, U (nominal trajectory), k (BoxQP warm start k_u_), reg, cost, alpha, inf_du are in/out, K and dV out.
 * follow == NULL: the oracle takes its own decisions and reports them in code[b] (trace encoding) and status[b]
 * (ORACLE_RUNNING = not terminated).  follow / follow_status != NULL: it follows those decisions and rep[b] reports
 * how its own verdicts differed.  The lock-step parity tests feed it the CUDA path's state before every iteration. */
void oracle_iterate_batch(const oracle_problem *p, const oracle_options *o, int batch, int nthreads, const double *x0,
                          const double *xref, const double *ref_traj, double *X, double *U, double *K, double *k,
                          double *reg, double *cost, double *alpha, double *inf_du, double *dV, const int *follow,
                          const int *follow_status, int *code, int *status, oracle_replay_report *rep);

/* ---------------------------------------------------------------------------------------------
 * IPDDP (src/cddp_core/ipddp_solver.cpp): cold start, use_ilqr = true, path inequality constraints, no
 * terminal constraints.  Same parity status as above ("parity unpinned").
 * ------------------------------------------------------------------------------------------- */
/* path-constraint kinds (include/cddp-cpp/cddp_core/constraint.hpp) */
enum { ORACLE_CON_CONTROL_BOX = 0, ORACLE_CON_STATE_BOX = 1, ORACLE_CON_BALL = 2, ORACLE_CON_LINEAR = 3 };
enum { ORACLE_BARRIER_ADAPTIVE = 0, ORACLE_BARRIER_MONOTONIC = 1, ORACLE_BARRIER_IPOPT = 2 };
#define ORACLE_IPDDP_HISTORY_COLS 9 /* objective, merit, alpha_pr, alpha_du, inf_du, inf_pr, inf_comp, reg, mu */

/* One entry of the path-constraint set; pass the entries in the order the reference iterates them (a
 * std::map keyed by constraint name: alphabetical).  Same layout as cddp_b200_constraint.
 *   CONTROL_BOX / STATE_BOX: p0 = lower [m|n], p1 = upper [m|n]            (constraint.hpp:144-251)
 *   BALL: rows = dim of the centre, p0 = centre [rows], p1 = &radius       (:320-440)
 *   LINEAR: rows, p0 = A [rows][n], p1 = b [rows]; scale is ignored        (:253-318) */
typedef struct {
  int type;
  int rows;
  double scale; /* scale_factor */
  const double *p0;
  const double *p1;
} oracle_constraint;

/* options.hpp:75-104 (barrier, filter) and :148-186 (IPDDPAlgorithmOptions); same layout as cddp_b200_ipddp_options */
typedef struct {
  double dual_var_init_scale;            /* 1e-1 */
  double slack_var_init_scale;           /* 1e-2 */
  double barrier_tol_mult;               /* 0.1 */
  double barrier_update_dual_weight;     /* 0.01 */
  double mu_kappa_epsilon;               /* 10 */
  double theta_0_floor;                  /* 1 */
  double mu_initial;                     /* 1 */
  double mu_min_value;                   /* 1e-10 */
  double mu_update_factor;               /* 0.5 */
  double mu_update_power;                /* 1.2 */
  double min_fraction_to_boundary;       /* 0.99 */
  double merit_acceptance_threshold;     /* filter 1e-6 */
  double violation_acceptance_threshold; /* 1e-6 */
  double max_violation_threshold;        /* 1e4 */
  double min_violation_for_armijo_check; /* 1e-7 */
  double jacobian_regularization_value;    /* ipddp 1e-8  (terminal-equality reduced system, options.hpp:181-184) */
  double jacobian_regularization_exponent; /* 0.25 */
  int theta_norm_l2;                     /* 0 = "l1" */
  int max_filter_size;                   /* 5 */
  int barrier_strategy;                  /* ADAPTIVE */
  int terminal_equality;                 /* 1: addTerminalConstraint(TerminalEqualityConstraint(reference state)) */
} oracle_ipddp_options;

typedef struct {
  double final_objective;
  double final_step_length; /* alpha_pr_ */
  double final_regularization;
  double inf_du;
  double inf_pr;
  double inf_comp;
  double mu;
  double merit;
  double decision_margin; /* test instrumentation: smallest relative margin of any line-search accept/reject decision */
  int iterations;
  int status;
  int history_len;
  int dual_dim;
} oracle_ipddp_result;

void oracle_ipddp_default_options(oracle_ipddp_options *io);
int oracle_total_dual_dim(const oracle_problem *p, const oracle_constraint *cs, int nc);
/* g(x,u) - upper [d], Gx [d][n], Gu [d][m] of the stacked constraint set */
void oracle_eval_constraints(const oracle_problem *p, const oracle_constraint *cs, int nc, const double *x,
                             const double *u, double *g, double *Gx, double *Gu);
/* CDDP::solve("IPDDP") for one instance.  U in = initial controls (X is re-rolled out, :876-882), X,U out;
 * K [N][m][n]; Y,S [N][d] (optional); history [max_iterations+1][ORACLE_IPDDP_HISTORY_COLS] (optional). */
void oracle_ipddp_solve(const oracle_problem *p, const oracle_options *o, const oracle_ipddp_options *io,
                        const oracle_constraint *cs, int nc, const double *x0, const double *xref,
                        const double *ref_traj, double *X, double *U, double *K, double *Y, double *S,
                        oracle_ipddp_result *res, double *history);
void oracle_ipddp_solve_batch(const oracle_problem *p, const oracle_options *o, const oracle_ipddp_options *io,
                              const oracle_constraint *cs, int nc, int batch, int nthreads, const double *x0,
                              const double *xref, const double *ref_traj, double *X, double *U, double *K, double *Y,
                              double *S, oracle_ipddp_result *res);
/* white-box single steps for per-step parity: state in/out through flat arrays.
 * oracle_ipddp_step runs initialize + `iters` full iterations and then ONE backward pass, returning every
 * intermediate the CUDA path exposes. */
/* ONE IPDDP main-loop entry per instance from a caller-supplied solver state (see cddp_oracle.cpp); follow / follow_status
 * as for oracle_iterate_batch. */
void oracle_ipddp_iterate_batch(const oracle_problem *p, const oracle_options *o, const oracle_ipddp_options *io,
                                const oracle_constraint *cs, int nc, int batch, int nthreads, const double *x0, const double *xref,
                                const double *ref_traj, double *X, double *U, double *Y, double *S, double *G, double *lamT,
                                double *filter, int *filter_size, double *scalars, const int *follow, const int *follow_status,
                                int *code, int *status, oracle_replay_report *rep, double *trial_table /* [B][ORACLE_MAX_ALPHAS][6] or NULL */);
void oracle_ipddp_probe(const oracle_problem *p, const oracle_options *o, const oracle_ipddp_options *io,
                        const oracle_constraint *cs, int nc, const double *x0, const double *xref,
                        const double *ref_traj, const double *U0, int iters, double *X, double *U, double *Y, double *S,
                        double *G, double *ku, double *Ku, double *ky, double *Ky, double *ks, double *Ks, double *dS,
                        double *dY, double *scalars /* [16]: mu, cost, merit, inf_pr, inf_du, inf_comp, step_norm, reg,
                        dV0, dV1, alpha_pr_max, alpha_du_max, filter_theta, theta, filter_size, bw_ok */,
                        double *trial_costs /* [num_alphas][4]: success, cost, merit, theta of every alpha */);

void oracle_debug_qp_stats(int enable, long long *iters32, long long *trials64, long long *facts16, long long *calls);
int oracle_hardware_threads(void);
const char *oracle_status_string(int status);

#ifdef __cplusplus
}
#endif
#endif
