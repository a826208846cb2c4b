// Marshals cddp::CDDP problem objects into the C ABI (include/cddp_b200.h).  See b200_solver.hpp.
#include "cddp_b200/b200_solver.hpp"

#include <chrono>
#include <cstring>

#include "../../../include/cddp_b200.h"

namespace cddp {
namespace b200 {

namespace {

struct Shared {  // the batch-shared part of a problem, in C-ABI form
  cddp_b200_problem p{};
  cddp_b200_options o{};
  std::vector<double> Q, R, Qf, lb, ub, ltiA, ltiB;
  bool has_ref_traj = false;
};

int integrator_id(const std::string &s) {
  if (s == "euler") return CDDP_B200_EULER;
  if (s == "heun") return CDDP_B200_HEUN;
  if (s == "rk3") return CDDP_B200_RK3;
  if (s == "rk4") return CDDP_B200_RK4;
  throw std::runtime_error("B200 CLDDP: unknown integration type '" + s + "'");
}

void flatten(const Eigen::MatrixXd &M, std::vector<double> &out) {
  out.resize((size_t)(M.rows() * M.cols()));
  for (long i = 0; i < M.rows(); ++i)
    for (long j = 0; j < M.cols(); ++j) out[(size_t)(i * M.cols() + j)] = M(i, j);
}

// Reads one CDDP into the shared C-ABI description; throws std::runtime_error for anything the device path
// cannot run (there is no CPU fallback to hand it to).
void describe(CDDP &ctx, Shared &s) {
  if (!ctx.hasSystem()) throw std::runtime_error("Dynamical system must be set before solving.");
  if (!ctx.hasObjective()) throw std::runtime_error("Objective function must be set before solving.");
  const DynamicalSystem &sys = ctx.getSystem();
  DeviceModelDescriptor dm;
  if (!sys.getDeviceModel(dm))
    throw std::runtime_error("B200 CLDDP: this DynamicalSystem has no device implementation (getDeviceModel() returned false); "
                             "there is no CPU fallback");
  const auto *obj = dynamic_cast<const QuadraticObjective *>(&ctx.getObjective());
  if (!obj) throw std::runtime_error("B200 CLDDP: only QuadraticObjective is device-resident; there is no CPU fallback");
  const int n = sys.getStateDim(), m = sys.getControlDim();
  if (n > CDDP_B200_MAX_N || m > CDDP_B200_MAX_M) throw std::runtime_error("B200 CLDDP: state/control dimension exceeds the kernel caps");
  s.p.model = dm.model;
  s.p.n = n;
  s.p.m = m;
  s.p.horizon = ctx.getHorizon();
  s.p.dt = ctx.getTimestep();
  s.p.integrator = integrator_id(sys.getIntegrationType());
  std::memcpy(s.p.model_params, dm.params, sizeof(dm.params));
  s.ltiA = dm.lti_A;
  s.ltiB = dm.lti_B;
  // QuadraticObjective scales by ITS OWN timestep (objective.cpp:38-39); the C ABI scales by the problem timestep
  const double ratio = obj->getTimestep() / ctx.getTimestep();
  flatten(obj->unscaledQ() * ratio, s.Q);
  flatten(obj->unscaledR() * ratio, s.R);
  flatten(obj->getQf(), s.Qf);
  // CLDDP looks the box up by the literal name and exact type (clddp_solver.cpp:85-86); other constraints are ignored by it
  if (auto *cc = ctx.getConstraint<ControlConstraint>("ControlConstraint")) {
    s.p.has_control_box = 1;
    s.lb.assign(cc->rawLowerBound().data(), cc->rawLowerBound().data() + m);
    s.ub.assign(cc->rawUpperBound().data(), cc->rawUpperBound().data() + m);
  }
  s.has_ref_traj = ctx.getObjective().getReferenceStates().size() == (size_t)(ctx.getHorizon() + 1);
  const CDDPOptions &o = ctx.getOptions();
  cddp_b200_default_options(&s.o);
  s.o.tolerance = o.tolerance;
  s.o.acceptable_tolerance = o.acceptable_tolerance;
  s.o.max_iterations = o.max_iterations;
  s.o.enable_parallel = o.enable_parallel ? 1 : 0;
  s.o.max_cpu_time = o.max_cpu_time;
  s.o.termination_scaling_max_factor = o.termination_scaling_max_factor;
  s.o.ls_max_iterations = o.line_search.max_iterations;
  s.o.ls_initial_step_size = o.line_search.initial_step_size;
  s.o.ls_min_step_size = o.line_search.min_step_size;
  s.o.ls_step_reduction_factor = o.line_search.step_reduction_factor;
  s.o.reg_initial_value = o.regularization.initial_value;
  s.o.reg_update_factor = o.regularization.update_factor;
  s.o.reg_max_value = o.regularization.max_value;
  s.o.reg_min_value = o.regularization.min_value;
  s.o.qp_max_iterations = o.box_qp.max_iterations;
  s.o.qp_min_gradient_norm = o.box_qp.min_gradient_norm;
  s.o.qp_min_relative_improvement = o.box_qp.min_relative_improvement;
  s.o.qp_step_decrease_factor = o.box_qp.step_decrease_factor;
  s.o.qp_min_step_size = o.box_qp.min_step_size;
  s.o.qp_armijo_constant = o.box_qp.armijo_constant;
  s.o.armijo_constant = o.filter.armijo_constant;
}

void bind_pointers(Shared &s) {
  s.p.Q = s.Q.data();
  s.p.R = s.R.data();
  s.p.Qf = s.Qf.data();
  s.p.lb = s.p.has_control_box ? s.lb.data() : nullptr;
  s.p.ub = s.p.has_control_box ? s.ub.data() : nullptr;
  s.p.lti_A = s.ltiA.empty() ? nullptr : s.ltiA.data();
  s.p.lti_B = s.ltiB.empty() ? nullptr : s.ltiB.data();
}

bool same_shared(const Shared &a, const Shared &b) {
  return a.p.model == b.p.model && a.p.n == b.p.n && a.p.m == b.p.m && a.p.horizon == b.p.horizon && a.p.dt == b.p.dt &&
         a.p.integrator == b.p.integrator && a.p.has_control_box == b.p.has_control_box &&
         std::memcmp(a.p.model_params, b.p.model_params, sizeof(a.p.model_params)) == 0 && a.Q == b.Q && a.R == b.R && a.Qf == b.Qf &&
         a.lb == b.lb && a.ub == b.ub && a.ltiA == b.ltiA && a.ltiB == b.ltiB && a.has_ref_traj == b.has_ref_traj &&
         std::memcmp(&a.o, &b.o, sizeof(a.o)) == 0;
}

void check(int rc, const char *what) {
  if (rc == CDDP_B200_OK) return;
  std::string msg = std::string("B200 CLDDP: ") + what + ": " + cddp_b200_error_string(rc);
  if (rc == CDDP_B200_ERR_CUDA || rc == CDDP_B200_ERR_OUT_OF_MEMORY) msg += std::string(" — ") + cddp_b200_last_cuda_error();
  throw std::runtime_error(msg);
}

}  // namespace

std::vector<CDDPSolution> solveBatch(const std::vector<CDDP *> &problems, int device) {
  if (problems.empty()) return {};
  const auto t_start = std::chrono::high_resolution_clock::now();
  const int B = (int)problems.size();
  Shared sh;
  describe(*problems[0], sh);
  for (int b = 1; b < B; ++b) {
    Shared other;
    describe(*problems[b], other);
    if (!same_shared(sh, other))
      throw std::runtime_error("B200 CLDDP: solveBatch needs structurally identical problems (model, weights, horizon, timestep, "
                               "options, control box); instance " + std::to_string(b) + " differs from instance 0");
  }
  bind_pointers(sh);
  const int n = sh.p.n, m = sh.p.m, N = sh.p.horizon;
  std::vector<double> x0((size_t)B * n), xref((size_t)B * n), X((size_t)B * (N + 1) * n), U((size_t)B * N * m), rt;
  if (sh.has_ref_traj) rt.resize((size_t)B * (N + 1) * n);
  for (int b = 0; b < B; ++b) {
    CDDP &c = *problems[b];
    c.initializeProblemIfNecessary();  // sizes X_/U_, X_[0] = x0, cost = inf, reg = initial (cddp_core.cpp:272-306)
    const Eigen::VectorXd ref = c.getObjective().getReferenceState();
    if ((int)ref.size() != n) throw std::runtime_error("B200 CLDDP: reference state has the wrong dimension");
    for (int i = 0; i < n; ++i) {
      x0[(size_t)b * n + i] = c.getInitialState()[i];
      xref[(size_t)b * n + i] = ref[i];
    }
    for (int t = 0; t <= N; ++t)
      for (int i = 0; i < n; ++i) X[((size_t)b * (N + 1) + t) * n + i] = c.X_[(size_t)t][i];
    for (int t = 0; t < N; ++t)
      for (int i = 0; i < m; ++i) U[((size_t)b * N + t) * m + i] = c.U_[(size_t)t][i];
    if (sh.has_ref_traj) {
      const auto refs = c.getObjective().getReferenceStates();
      for (int t = 0; t <= N; ++t)
        for (int i = 0; i < n; ++i) rt[((size_t)b * (N + 1) + t) * n + i] = refs[(size_t)t][i];
    }
  }
  std::vector<double> K((size_t)B * N * m * n), cost((size_t)B), alpha((size_t)B), reg((size_t)B), inf_du((size_t)B);
  std::vector<int> iters((size_t)B), status((size_t)B);
  const bool want_hist = problems[0]->getOptions().return_iteration_info;
  std::vector<double> hist;
  std::vector<int> hist_len;

  cddp_b200_solver *h = nullptr;
  check(cddp_b200_create(&sh.p, &sh.o, B, device, &h), "create");
  struct Guard {
    cddp_b200_solver *h;
    ~Guard() { cddp_b200_destroy(h); }
  } guard{h};
  if (want_hist) check(cddp_b200_enable_history(h, 1), "enable_history");
  check(cddp_b200_set_instances(h, x0.data(), xref.data(), sh.has_ref_traj ? rt.data() : nullptr, X.data(), U.data()), "set_instances");
  check(cddp_b200_solve(h), "solve");
  check(cddp_b200_get_solution(h, X.data(), U.data(), K.data(), cost.data(), iters.data(), status.data(), alpha.data(), reg.data(),
                               inf_du.data()), "get_solution");
  if (want_hist) {
    hist.resize((size_t)B * (sh.o.max_iterations + 1) * 4);
    hist_len.resize((size_t)B);
    check(cddp_b200_get_history(h, hist.data(), hist_len.data()), "get_history");
  }
  const double ms = std::chrono::duration<double, std::milli>(std::chrono::high_resolution_clock::now() - t_start).count();

  std::vector<CDDPSolution> out((size_t)B);
  for (int b = 0; b < B; ++b) {
    CDDP &c = *problems[b];
    CDDPSolution &s = out[(size_t)b];
    s.solver_name = "CLDDP";
    s.status_message = cddp_b200_status_string(status[(size_t)b]);
    s.iterations_completed = iters[(size_t)b];
    s.solve_time_ms = ms;  // wall time of the whole batched call
    s.final_objective = cost[(size_t)b];
    s.final_step_length = alpha[(size_t)b];
    s.final_regularization = reg[(size_t)b];
    s.final_dual_infeasibility = inf_du[(size_t)b];
    s.time_points.resize((size_t)N + 1);
    s.state_trajectory.assign((size_t)N + 1, Eigen::VectorXd::Zero(n));
    s.control_trajectory.assign((size_t)N, Eigen::VectorXd::Zero(m));
    s.feedback_gains.assign((size_t)N, Eigen::MatrixXd::Zero(m, n));
    for (int t = 0; t <= N; ++t) {
      s.time_points[(size_t)t] = t * sh.p.dt;
      for (int i = 0; i < n; ++i) s.state_trajectory[(size_t)t][i] = X[((size_t)b * (N + 1) + t) * n + i];
    }
    for (int t = 0; t < N; ++t) {
      for (int i = 0; i < m; ++i) {
        s.control_trajectory[(size_t)t][i] = U[((size_t)b * N + t) * m + i];
        for (int j = 0; j < n; ++j) s.feedback_gains[(size_t)t](i, j) = K[(((size_t)b * N + t) * m + i) * n + j];
      }
    }
    if (want_hist) {
      const int cap = sh.o.max_iterations + 1;
      for (int k = 0; k < hist_len[(size_t)b]; ++k) {
        const double *row = &hist[((size_t)b * cap + k) * 4];
        s.history.objective.push_back(row[0]);
        s.history.merit_function.push_back(row[0]);
        s.history.step_length_primal.push_back(row[1]);
        s.history.dual_infeasibility.push_back(row[2]);
        s.history.regularization.push_back(row[3]);
      }
    }
    // leave the context as CDDPSolverBase::solve leaves it (callers and warm starts read these fields)
    c.X_ = s.state_trajectory;
    c.U_ = s.control_trajectory;
    c.cost_ = c.merit_function_ = s.final_objective;
    c.inf_du_ = inf_du[(size_t)b];
    c.alpha_pr_ = s.final_step_length;
    c.regularization_ = s.final_regularization;
  }
  return out;
}

void CLDDPSolver::initialize(CDDP &context) {
  Shared s;
  describe(context, s);  // throws for problems the device path cannot run
}

CDDPSolution CLDDPSolver::solve(CDDP &context) {
  std::vector<CDDP *> one{&context};
  return solveBatch(one, device_)[0];
}

void registerSolvers(int device) {
  auto factory = [device]() { return std::unique_ptr<ISolverAlgorithm>(new CLDDPSolver(device)); };
  CDDP::registerSolver("CLDDP", factory);
  CDDP::registerSolver("CLDDP_B200", factory);
}

}  // namespace b200
}  // namespace cddp
