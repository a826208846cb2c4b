// Marshals cddp::CDDP problem objects into the C ABI (include/cddp_b200.h).  See b200_solver.hpp.
#include "cddp_b200/b200_solver.hpp"

#include <chrono>
#include <exception>
#include <thread>
#include <cstring>

#include "../../../include/cddp_b200.h"

namespace cddp {
namespace b200 {

namespace {

struct Shared {  // the batch-shared part of a problem, in C-ABI form
  cddp_b200_problem p{};
  cddp_b200_options o{};
  std::vector<double> Q, R, Qf, lb, ub, ltiA, ltiB;
  std::string source;  // user-defined dynamics (CDDP_B200_MODEL_USER)
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

// The effect of CDDP::initializeProblemIfNecessary (private, cddp_core.cpp:272-306) that the batched facade needs, through
// the PUBLIC members of the reference class only (X_, U_ are public fields, cddp_core.hpp:323-326): size the nominal
// trajectory if the caller gave none, and pin its first state to the initial state.
void ensure_trajectory(CDDP &c, int n, int m, int N) {
  auto fits = [](const std::vector<Eigen::VectorXd> &v, int len, int dim) {
    if ((int)v.size() != len) return false;
    for (const auto &e : v)
      if ((int)e.size() != dim) return false;
    return true;
  };
  if (!fits(c.X_, N + 1, n)) c.X_.assign((size_t)(N + 1), Eigen::VectorXd::Zero(n));
  if (!fits(c.U_, N, m)) c.U_.assign((size_t)N, Eigen::VectorXd::Zero(m));
  c.X_[0] = c.getInitialState();
}

// Reads one CDDP into the shared C-ABI description; throws std::runtime_error for anything the device path
// cannot run (there is no CPU fallback to hand it to).
void describe(CDDP &ctx, Shared &s) {
  // (a missing system / objective is rejected by CDDP::solve before a plugin is called, cddp_core.cpp:277-282; solveBatch
  // documents complete problems as its precondition — the reference has no public "is it set" query)
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
  s.source = dm.source;
  if (dm.model == CDDP_B200_MODEL_USER && dm.source.empty())
    throw std::runtime_error("B200: a CDDP_B200_MODEL_USER DynamicalSystem must return its device source from getDeviceModel()");
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
         a.lb == b.lb && a.ub == b.ub && a.ltiA == b.ltiA && a.ltiB == b.ltiB && a.source == b.source && a.has_ref_traj == b.has_ref_traj &&
         std::memcmp(&a.o, &b.o, sizeof(a.o)) == 0;
}

void check(int rc, const char *what) {
  if (rc == CDDP_B200_OK) return;
  std::string msg = std::string("B200 CLDDP: ") + what + ": " + cddp_b200_error_string(rc);
  if (rc == CDDP_B200_ERR_CUDA || rc == CDDP_B200_ERR_OUT_OF_MEMORY) msg += std::string(" — ") + cddp_b200_last_cuda_error();
  if (rc == CDDP_B200_ERR_USER_MODEL) msg += std::string("\n") + cddp_b200_last_compile_log();
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
    ensure_trajectory(c, n, m, N);
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
  check(cddp_b200_create_ex(&sh.p, &sh.o, sh.source.empty() ? nullptr : sh.source.c_str(), B, device, &h), "create");
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

// ------------------------------------------------------------------------------------------------ several devices
namespace {
template <class SolveOne>
std::vector<CDDPSolution> solve_sharded(const std::vector<CDDP *> &problems, const std::vector<int> &devices, SolveOne solve_one) {
  if (devices.empty()) throw std::runtime_error("B200: solveBatch needs at least one device");
  const size_t B = problems.size(), G = devices.size();
  if (B == 0) return {};
  const size_t per = (B + G - 1) / G;  // B_g = ceil(B / G), trailing shards may be short or empty
  std::vector<std::vector<CDDPSolution>> parts(G);
  std::vector<std::exception_ptr> errors(G);
  std::vector<std::thread> threads;
  for (size_t g = 0; g < G; ++g) {
    const size_t lo = std::min(g * per, B), hi = std::min((g + 1) * per, B);
    if (lo == hi) continue;
    threads.emplace_back([&, g, lo, hi]() {
      try {
        std::vector<CDDP *> shard(problems.begin() + (long)lo, problems.begin() + (long)hi);
        parts[g] = solve_one(shard, devices[g]);
      } catch (...) {
        errors[g] = std::current_exception();
      }
    });
  }
  for (auto &t : threads) t.join();
  for (auto &e : errors)
    if (e) std::rethrow_exception(e);
  std::vector<CDDPSolution> out;
  out.reserve(B);
  for (auto &p : parts)
    for (auto &sol : p) out.push_back(std::move(sol));
  return out;
}
}  // namespace

std::vector<CDDPSolution> solveBatch(const std::vector<CDDP *> &problems, const std::vector<int> &devices) {
  if (!problems.empty()) {  // structural identity across shards (each shard checks its own members against its first)
    Shared first;
    describe(*problems[0], first);
    const size_t per = (problems.size() + devices.size() - 1) / std::max<size_t>(devices.size(), 1);
    for (size_t lo = per; lo < problems.size(); lo += per) {
      Shared other;
      describe(*problems[lo], other);
      if (!same_shared(first, other))
        throw std::runtime_error("B200 CLDDP: solveBatch needs structurally identical problems; instance " + std::to_string(lo) +
                                 " differs from instance 0");
    }
  }
  return solve_sharded(problems, devices, [](const std::vector<CDDP *> &shard, int dev) { return solveBatch(shard, dev); });
}

std::vector<CDDPSolution> solveBatchIPDDP(const std::vector<CDDP *> &problems, const std::vector<int> &devices) {
  return solve_sharded(problems, devices, [](const std::vector<CDDP *> &shard, int dev) { return solveBatchIPDDP(shard, dev); });
}

std::vector<int> availableDevices() {
  int count = 0;
  if (cddp_b200_device_count(&count) != CDDP_B200_OK) count = 0;
  std::vector<int> d((size_t)std::max(count, 0));
  for (int i = 0; i < count; ++i) d[(size_t)i] = i;
  return d;
}

void CLDDPSolver::initialize(CDDP &context) {
  Shared s;
  describe(context, s);  // throws for problems the device path cannot run
}

CDDPSolution CLDDPSolver::solve(CDDP &context) {
  std::vector<CDDP *> one{&context};
  return solveBatch(one, device_)[0];
}

// ------------------------------------------------------------------------------------------------ IPDDP
namespace {

struct ConDesc {  // one path constraint in C-ABI form, with its backing storage
  int type = 0, rows = 0;
  double scale = 1.0;
  std::vector<double> p0, p1;
  bool operator==(const ConDesc &o) const { return type == o.type && rows == o.rows && scale == o.scale && p0 == o.p0 && p1 == o.p1; }
};

struct IpShared {
  cddp_b200_ipddp_options io{};
  std::vector<ConDesc> cons;
};

void describe_ipddp(CDDP &ctx, IpShared &s) {
  const CDDPOptions &o = ctx.getOptions();
  if (o.warm_start) throw std::runtime_error("B200 IPDDP: options.warm_start is not supported by the device path");
  if (!o.use_ilqr) throw std::runtime_error("B200 IPDDP: use_ilqr = false (second-order dynamics terms) is not supported by the device path");
  // terminal constraints: one TerminalEqualityConstraint whose target is the objective's reference state runs on the device
  // (terminal-equality branch, ipddp_solver.cpp:1120-1353); anything else is a setup error
  bool terminal_equality = false;
  for (const auto &kv : ctx.getTerminalConstraintSet()) {
    const auto *te = dynamic_cast<const TerminalEqualityConstraint *>(kv.second.get());
    if (!te || terminal_equality)
      throw std::runtime_error("B200 IPDDP: terminal constraint '" + kv.first + "' is not supported by the device path (supported: "
                               "a single TerminalEqualityConstraint on the reference state)");
    const Eigen::VectorXd ref = ctx.getObjective().getReferenceState();
    if (te->getTargetState().size() != ref.size())
      throw std::runtime_error("B200 IPDDP: TerminalEqualityConstraint target has the wrong dimension");
    for (long i = 0; i < ref.size(); ++i)
      if (te->getTargetState()[i] != ref[i])
        throw std::runtime_error("B200 IPDDP: the device path supports a TerminalEqualityConstraint only on the objective's reference state");
    terminal_equality = true;
  }
  if (o.ipddp.check_state_stationarity || o.ipddp.warmstart_repair)
    throw std::runtime_error("B200 IPDDP: check_state_stationarity / warmstart_repair are not supported by the device path");
  cddp_b200_ipddp_default_options(&s.io);
  s.io.dual_var_init_scale = o.ipddp.dual_var_init_scale;
  s.io.slack_var_init_scale = o.ipddp.slack_var_init_scale;
  s.io.barrier_tol_mult = o.ipddp.barrier_tol_mult;
  s.io.barrier_update_dual_weight = o.ipddp.barrier_update_dual_weight;
  s.io.mu_kappa_epsilon = o.ipddp.mu_kappa_epsilon;
  s.io.theta_0_floor = o.ipddp.theta_0_floor;
  s.io.mu_initial = o.ipddp.barrier.mu_initial;
  s.io.mu_min_value = o.ipddp.barrier.mu_min_value;
  s.io.mu_update_factor = o.ipddp.barrier.mu_update_factor;
  s.io.mu_update_power = o.ipddp.barrier.mu_update_power;
  s.io.min_fraction_to_boundary = o.ipddp.barrier.min_fraction_to_boundary;
  s.io.merit_acceptance_threshold = o.filter.merit_acceptance_threshold;
  s.io.violation_acceptance_threshold = o.filter.violation_acceptance_threshold;
  s.io.max_violation_threshold = o.filter.max_violation_threshold;
  s.io.min_violation_for_armijo_check = o.filter.min_violation_for_armijo_check;
  s.io.theta_norm_l2 = o.ipddp.theta_norm == "l2" ? 1 : 0;
  s.io.max_filter_size = o.ipddp.max_filter_size;
  s.io.jacobian_regularization_value = o.ipddp.jacobian_regularization_value;
  s.io.jacobian_regularization_exponent = o.ipddp.jacobian_regularization_exponent;
  s.io.terminal_equality = terminal_equality ? 1 : 0;
  s.io.barrier_strategy = o.ipddp.barrier.strategy == BarrierStrategy::ADAPTIVE ? CDDP_B200_BARRIER_ADAPTIVE
                          : o.ipddp.barrier.strategy == BarrierStrategy::MONOTONIC ? CDDP_B200_BARRIER_MONOTONIC : CDDP_B200_BARRIER_IPOPT;
  const int n = ctx.getStateDim();
  // std::map iteration = the reference's constraint order (ipddp_solver.cpp:1375-1390)
  for (const auto &kv : ctx.getConstraintSet()) {
    ConDesc c;
    const Constraint *base = kv.second.get();
    if (auto *cc = dynamic_cast<const ControlConstraint *>(base)) {
      c.type = CDDP_B200_CON_CONTROL_BOX;
      c.rows = (int)cc->rawLowerBound().size();
      c.scale = cc->getScaleFactor();
      c.p0.assign(cc->rawLowerBound().data(), cc->rawLowerBound().data() + c.rows);
      c.p1.assign(cc->rawUpperBound().data(), cc->rawUpperBound().data() + c.rows);
    } else if (auto *sc = dynamic_cast<const StateConstraint *>(base)) {
      c.type = CDDP_B200_CON_STATE_BOX;
      c.rows = (int)sc->rawLowerBound().size();
      c.scale = sc->getScaleFactor();
      c.p0.assign(sc->rawLowerBound().data(), sc->rawLowerBound().data() + c.rows);
      c.p1.assign(sc->rawUpperBound().data(), sc->rawUpperBound().data() + c.rows);
    } else if (auto *bc = dynamic_cast<const BallConstraint *>(base)) {
      c.type = CDDP_B200_CON_BALL;
      c.rows = (int)bc->getCenter().size();
      c.scale = bc->getScaleFactor();
      c.p0.assign(bc->getCenter().data(), bc->getCenter().data() + c.rows);
      c.p1.assign(1, bc->getRadius());
    } else if (auto *lc = dynamic_cast<const LinearConstraint *>(base)) {
      c.type = CDDP_B200_CON_LINEAR;
      c.rows = (int)lc->getUpperBound().size();
      c.scale = lc->getScaleFactor();
      if (lc->getA().cols() != n) throw std::runtime_error("B200 IPDDP: LinearConstraint A has the wrong number of columns");
      flatten(lc->getA(), c.p0);
      const Eigen::VectorXd b = lc->getUpperBound();
      c.p1.assign(b.data(), b.data() + c.rows);
    } else {
      throw std::runtime_error("B200 IPDDP: path constraint '" + kv.first + "' has no device implementation (supported: ControlConstraint, "
                               "StateConstraint, LinearConstraint, BallConstraint); there is no CPU fallback");
    }
    s.cons.push_back(std::move(c));
  }
}

}  // namespace

std::vector<CDDPSolution> solveBatchIPDDP(const std::vector<CDDP *> &problems, int device) {
  if (problems.empty()) return {};
  const auto t_start = std::chrono::high_resolution_clock::now();
  const int B = (int)problems.size();
  Shared sh;
  IpShared ish;
  describe(*problems[0], sh);
  describe_ipddp(*problems[0], ish);
  if (sh.has_ref_traj) throw std::runtime_error("B200 IPDDP: per-timestep reference states are not supported by the device path");
  for (int b = 1; b < B; ++b) {
    Shared other;
    IpShared iother;
    describe(*problems[b], other);
    describe_ipddp(*problems[b], iother);
    if (!same_shared(sh, other) || std::memcmp(&ish.io, &iother.io, sizeof(ish.io)) != 0 || !(ish.cons == iother.cons))
      throw std::runtime_error("B200 IPDDP: solveBatchIPDDP needs structurally identical problems (model, weights, horizon, timestep, "
                               "options, constraint set); instance " + std::to_string(b) + " differs from instance 0");
  }
  sh.p.has_control_box = 0;  // a ControlConstraint is a row pair of the constraint set here, never a clamp
  bind_pointers(sh);
  std::vector<cddp_b200_constraint> cs(ish.cons.size());
  for (size_t i = 0; i < cs.size(); ++i) {
    cs[i].type = ish.cons[i].type;
    cs[i].rows = ish.cons[i].rows;
    cs[i].scale = ish.cons[i].scale;
    cs[i].p0 = ish.cons[i].p0.data();
    cs[i].p1 = ish.cons[i].p1.data();
  }
  const int n = sh.p.n, m = sh.p.m, N = sh.p.horizon;
  std::vector<double> x0((size_t)B * n), xref((size_t)B * n), X((size_t)B * (N + 1) * n), U((size_t)B * N * m);
  for (int b = 0; b < B; ++b) {
    CDDP &c = *problems[b];
    ensure_trajectory(c, n, m, N);
    const Eigen::VectorXd ref = c.getObjective().getReferenceState();
    if ((int)ref.size() != n) throw std::runtime_error("B200 IPDDP: reference state has the wrong dimension");
    for (int i = 0; i < n; ++i) {
      x0[(size_t)b * n + i] = c.getInitialState()[i];
      xref[(size_t)b * n + i] = ref[i];
    }
    for (int t = 0; t < N; ++t)
      for (int i = 0; i < m; ++i) U[((size_t)b * N + t) * m + i] = c.U_[(size_t)t][i];
  }
  std::vector<double> K((size_t)B * N * m * n), cost((size_t)B), alpha((size_t)B), reg((size_t)B), inf_du((size_t)B), sc((size_t)B * 8);
  std::vector<int> iters((size_t)B), status((size_t)B);
  const bool want_hist = problems[0]->getOptions().return_iteration_info;
  std::vector<double> hist;
  std::vector<int> hist_len;
  cddp_b200_solver *h = nullptr;
  check(cddp_b200_ipddp_create_ex(&sh.p, &sh.o, &ish.io, cs.empty() ? nullptr : cs.data(), (int)cs.size(),
                                  sh.source.empty() ? nullptr : sh.source.c_str(), B, device, &h), "ipddp_create");
  struct Guard {
    cddp_b200_solver *h;
    ~Guard() { cddp_b200_destroy(h); }
  } guard{h};
  if (want_hist) check(cddp_b200_enable_history(h, 1), "enable_history");
  check(cddp_b200_set_instances(h, x0.data(), xref.data(), nullptr, nullptr, U.data()), "set_instances");
  check(cddp_b200_solve(h), "solve");
  check(cddp_b200_get_solution(h, X.data(), U.data(), K.data(), cost.data(), iters.data(), status.data(), alpha.data(), reg.data(),
                               inf_du.data()), "get_solution");
  check(cddp_b200_ipddp_get_solution(h, nullptr, nullptr, nullptr, sc.data()), "ipddp_get_solution");
  const int cap = sh.o.max_iterations + 1;
  if (want_hist) {
    hist.resize((size_t)B * cap * 9);
    hist_len.resize((size_t)B);
    check(cddp_b200_ipddp_get_history(h, hist.data(), hist_len.data()), "ipddp_get_history");
  }
  const double ms = std::chrono::duration<double, std::milli>(std::chrono::high_resolution_clock::now() - t_start).count();
  std::vector<CDDPSolution> out((size_t)B);
  for (int b = 0; b < B; ++b) {
    CDDP &c = *problems[b];
    CDDPSolution &s = out[(size_t)b];
    const double *q = &sc[(size_t)b * 8];  // mu, merit, inf_pr, inf_comp, step_norm, alpha_du, alpha_pr_max, alpha_du_max
    s.solver_name = "IPDDP";
    s.status_message = cddp_b200_status_string(status[(size_t)b]);
    s.iterations_completed = iters[(size_t)b];
    s.solve_time_ms = ms;
    s.final_objective = cost[(size_t)b];
    s.final_step_length = alpha[(size_t)b];
    s.final_regularization = reg[(size_t)b];
    s.final_barrier_mu = q[0];  // populateSolverSpecificSolution (ipddp_solver.cpp:2090-2097)
    s.final_primal_infeasibility = q[2];
    s.final_dual_infeasibility = inf_du[(size_t)b];
    s.final_complementary_infeasibility = q[3];
    s.time_points.resize((size_t)N + 1);
    s.state_trajectory.assign((size_t)N + 1, Eigen::VectorXd::Zero(n));
    s.control_trajectory.assign((size_t)N, Eigen::VectorXd::Zero(m));
    s.feedback_gains.assign((size_t)N, Eigen::MatrixXd::Zero(m, n));
    for (int t = 0; t <= N; ++t) {
      s.time_points[(size_t)t] = t * sh.p.dt;
      for (int i = 0; i < n; ++i) s.state_trajectory[(size_t)t][i] = X[((size_t)b * (N + 1) + t) * n + i];
    }
    for (int t = 0; t < N; ++t)
      for (int i = 0; i < m; ++i) {
        s.control_trajectory[(size_t)t][i] = U[((size_t)b * N + t) * m + i];
        for (int j = 0; j < n; ++j) s.feedback_gains[(size_t)t](i, j) = K[(((size_t)b * N + t) * m + i) * n + j];
      }
    if (want_hist)
      for (int k = 0; k < hist_len[(size_t)b]; ++k) {
        const double *row = &hist[((size_t)b * cap + k) * 9];
        s.history.objective.push_back(row[0]);
        s.history.merit_function.push_back(row[1]);
        s.history.step_length_primal.push_back(row[2]);
        s.history.step_length_dual.push_back(row[3]);
        s.history.dual_infeasibility.push_back(row[4]);
        s.history.primal_infeasibility.push_back(row[5]);
        s.history.complementary_infeasibility.push_back(row[6]);
        s.history.regularization.push_back(row[7]);
        s.history.barrier_mu.push_back(row[8]);
      }
    c.X_ = s.state_trajectory;
    c.U_ = s.control_trajectory;
    c.cost_ = s.final_objective;
    c.merit_function_ = q[1];
    c.inf_pr_ = q[2];
    c.inf_du_ = inf_du[(size_t)b];
    c.inf_comp_ = q[3];
    c.step_norm_ = q[4];
    c.alpha_pr_ = s.final_step_length;
    c.alpha_du_ = q[5];
    c.regularization_ = s.final_regularization;
  }
  return out;
}

void IPDDPSolver::initialize(CDDP &context) {
  Shared s;
  IpShared is;
  describe(context, s);
  describe_ipddp(context, is);
}

CDDPSolution IPDDPSolver::solve(CDDP &context) {
  std::vector<CDDP *> one{&context};
  return solveBatchIPDDP(one, device_)[0];
}

void registerSolvers(int device) {
  auto factory = [device]() { return std::unique_ptr<ISolverAlgorithm>(new CLDDPSolver(device)); };
  CDDP::registerSolver("CLDDP", factory);
  CDDP::registerSolver("CLDDP_B200", factory);
  auto ip_factory = [device]() { return std::unique_ptr<ISolverAlgorithm>(new IPDDPSolver(device)); };
  CDDP::registerSolver("IPDDP", ip_factory);
  CDDP::registerSolver("IPDDP_B200", ip_factory);
}

}  // namespace b200
}  // namespace cddp
