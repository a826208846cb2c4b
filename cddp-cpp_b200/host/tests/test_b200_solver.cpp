// GPU test of the drop-in route: cddp::CDDP::solve("CLDDP") resolved through the registry to the B200 solver, and
// the batched facade, written the way the reference's own CLDDP tests are (tests/cddp_core/test_clddp_solver.cpp:
// SolvePendulum :28-118, SolveUnicycle :231-290, SolveQuadrotor pattern :570-710) — plus what those tests do NOT
// assert: agreement of cost / trajectory / gains with the CPU oracle (linked here ONLY as the checker).
#include <cstring>

#include "cddp_b200/b200_solver.hpp"
#include "../../../include/cddp_b200.h"  // CDDP_B200_MODEL_USER
#include "cddp_oracle.h"
#include "check.hpp"

using namespace cddp;

namespace {
Eigen::VectorXd vec(std::initializer_list<double> l) {
  Eigen::VectorXd v((long)l.size());
  long i = 0;
  for (double x : l) v[i++] = x;
  return v;
}
Eigen::MatrixXd diag(std::initializer_list<double> l) {
  Eigen::MatrixXd M = Eigen::MatrixXd::Zero((long)l.size(), (long)l.size());
  long i = 0;
  for (double x : l) { M(i, i) = x; ++i; }
  return M;
}

void SolvePendulum() {
  const int N = 500;
  const double dt = 0.05;
  CDDPOptions o;
  o.max_iterations = 100; o.tolerance = 1e-3; o.acceptable_tolerance = 1e-4; o.regularization.initial_value = 1e-6; o.verbose = false;
  o.return_iteration_info = true;
  const auto x0 = vec({M_PI, 0.0}), goal = vec({0.0, 0.0});
  CDDP c(x0, goal, N, dt, std::make_unique<Pendulum>(dt, 1.0, 1.0, 0.0, "euler"),
         std::make_unique<QuadraticObjective>(Eigen::MatrixXd::Zero(2, 2), diag({0.1}), diag({100, 100}), goal,
                                              std::vector<Eigen::VectorXd>(), dt), o);
  c.addPathConstraint("ControlConstraint", std::make_unique<ControlConstraint>(vec({-10.0}), vec({10.0})));
  std::vector<Eigen::VectorXd> X((size_t)N + 1, x0), U((size_t)N, vec({0.0}));
  c.setInitialTrajectory(X, U);
  const double J0 = c.getObjective().evaluate(X, U);
  CDDPSolution s = c.solve(SolverType::CLDDP);
  // what the reference test asserts (:105-118, :149-151)
  CHECK(s.solver_name == "CLDDP");
  CHECK(s.status_message == "OptimalSolutionFound" || s.status_message == "AcceptableSolutionFound");
  CHECK(s.iterations_completed > 0 && s.iterations_completed <= o.max_iterations);
  CHECK(s.final_objective < J0);
  CHECK(s.state_trajectory.size() == (size_t)N + 1 && s.control_trajectory.size() == (size_t)N && s.feedback_gains.size() == (size_t)N);
  CHECK(s.time_points.size() == (size_t)N + 1 && std::fabs(s.time_points.back() - N * dt) < 1e-12);
  CHECK(s.feedback_gains[0].rows() == 1 && s.feedback_gains[0].cols() == 2);
  CHECK(!s.history.objective.empty() && s.history.objective.front() == J0);
  // context left as CDDPSolverBase::solve leaves it
  CHECK(c.cost_ == s.final_objective && c.X_.size() == (size_t)N + 1 && c.regularization_ == s.final_regularization);
  // oracle parity
  const double Q[4] = {0, 0, 0, 0}, R[1] = {0.1}, Qf[4] = {100, 0, 0, 100}, lb[1] = {-10}, ub[1] = {10};
  oracle_problem p{};
  p.model = ORACLE_PENDULUM; p.n = 2; p.m = 1; p.horizon = N; p.dt = dt; p.integrator = ORACLE_EULER; p.has_control_box = 1;
  p.model_params[0] = 1.0; p.model_params[1] = 1.0; p.Q = Q; p.R = R; p.Qf = Qf; p.lb = lb; p.ub = ub;
  oracle_options oo;
  oracle_default_options(&oo);
  oo.max_iterations = 100; oo.tolerance = 1e-3; oo.acceptable_tolerance = 1e-4;
  std::vector<double> Xo((size_t)(N + 1) * 2), Uo((size_t)N, 0.0), Ko((size_t)N * 2), ko((size_t)N);
  for (int t = 0; t <= N; ++t) { Xo[(size_t)t * 2] = M_PI; Xo[(size_t)t * 2 + 1] = 0.0; }
  const double x0a[2] = {M_PI, 0.0}, xr[2] = {0, 0};
  oracle_result res;
  oracle_solve(&p, &oo, x0a, xr, nullptr, Xo.data(), Uo.data(), Ko.data(), ko.data(), &res, nullptr);
  CHECK(res.iterations == s.iterations_completed);
  CHECK(std::string(oracle_status_string(res.status)) == s.status_message);
  CHECK(std::fabs(res.final_objective - s.final_objective) <= 1e-6 * std::fabs(res.final_objective));
  double dx = 0.0, dk = 0.0, kmax = 0.0;
  for (int t = 0; t <= N; ++t)
    for (int i = 0; i < 2; ++i) dx = std::max(dx, std::fabs(Xo[(size_t)t * 2 + i] - s.state_trajectory[(size_t)t][i]));
  for (int t = 0; t < N; ++t)
    for (int j = 0; j < 2; ++j) {
      dk = std::max(dk, std::fabs(Ko[(size_t)t * 2 + j] - s.feedback_gains[(size_t)t](0, j)));
      kmax = std::max(kmax, std::fabs(Ko[(size_t)t * 2 + j]));
    }
  CHECK(dx < 1e-6);
  CHECK(dk < 1e-6 * kmax);
  std::printf("  pendulum: %d iterations, %s, J=%.9f (oracle %.9f), |dX|max=%.2e\n", s.iterations_completed, s.status_message.c_str(),
              s.final_objective, res.final_objective, dx);
}

void SolveQuadrotorBatch() {
  const int N = 60, B = 12;
  const double dt = 0.02, hover = 9.81 / 4.0;
  CDDPOptions o;
  o.max_iterations = 15; o.verbose = false; o.regularization.initial_value = 1e-4; o.line_search.max_iterations = 15;
  Eigen::MatrixXd I = Eigen::MatrixXd::Zero(3, 3);
  I(0, 0) = 0.01; I(1, 1) = 0.01; I(2, 2) = 0.02;
  Eigen::MatrixXd Q = Eigen::MatrixXd::Zero(13, 13);
  Q(4, 4) = Q(5, 5) = Q(6, 6) = 0.1;
  const Eigen::MatrixXd R = diag({0.1, 0.1, 0.1, 0.1}), Qf = diag({500, 500, 500, 1, 1, 1, 1, 10, 10, 10, 0, 0, 0});
  std::vector<std::unique_ptr<CDDP>> owners;
  std::vector<CDDP *> batch;
  for (int b = 0; b < B; ++b) {
    Eigen::VectorXd x0 = Eigen::VectorXd::Zero(13), goal = Eigen::VectorXd::Zero(13);
    x0[3] = 1.0;
    x0[0] = 0.05 * b;
    goal[0] = 3.0 - 0.1 * b; goal[2] = 2.0 + 0.05 * b; goal[3] = 1.0;
    auto c = std::make_unique<CDDP>(x0, goal, N, dt, std::make_unique<Quadrotor>(dt, 1.0, I, 0.2, "rk4"),
                                    std::make_unique<QuadraticObjective>(Q, R, Qf, goal, std::vector<Eigen::VectorXd>(), dt), o);
    c->addPathConstraint("ControlConstraint", std::make_unique<ControlConstraint>(Eigen::VectorXd::Zero(4), Eigen::VectorXd::Constant(4, 5.0)));
    std::vector<Eigen::VectorXd> X((size_t)N + 1, x0), U((size_t)N, Eigen::VectorXd::Constant(4, hover));
    c->setInitialTrajectory(X, U);
    batch.push_back(c.get());
    owners.push_back(std::move(c));
  }
  auto sols = b200::solveBatch(batch);
  CHECK(sols.size() == (size_t)B);
  // the same problems one by one through the plugin route must give the same answers (batch-independence)
  for (int b : {0, 5, 11}) {
    Eigen::VectorXd x0 = Eigen::VectorXd::Zero(13), goal = Eigen::VectorXd::Zero(13);
    x0[3] = 1.0; x0[0] = 0.05 * b;
    goal[0] = 3.0 - 0.1 * b; goal[2] = 2.0 + 0.05 * b; goal[3] = 1.0;
    CDDP c(x0, goal, N, dt, std::make_unique<Quadrotor>(dt, 1.0, I, 0.2, "rk4"),
           std::make_unique<QuadraticObjective>(Q, R, Qf, goal, std::vector<Eigen::VectorXd>(), dt), o);
    c.addPathConstraint("ControlConstraint", std::make_unique<ControlConstraint>(Eigen::VectorXd::Zero(4), Eigen::VectorXd::Constant(4, 5.0)));
    std::vector<Eigen::VectorXd> X((size_t)N + 1, x0), U((size_t)N, Eigen::VectorXd::Constant(4, hover));
    c.setInitialTrajectory(X, U);
    auto s = c.solve("CLDDP_B200");
    CHECK(s.final_objective == sols[(size_t)b].final_objective);
    CHECK(s.iterations_completed == sols[(size_t)b].iterations_completed);
    CHECK((s.state_trajectory.back() - sols[(size_t)b].state_trajectory.back()).norm() == 0.0);
  }
  // oracle on every instance
  std::vector<double> Qa(169, 0.0), Ra(16, 0.0), Qfa(169, 0.0);
  for (int i = 0; i < 13; ++i) { Qa[(size_t)i * 14] = Q(i, i); Qfa[(size_t)i * 14] = Qf(i, i); }
  for (int i = 0; i < 4; ++i) Ra[(size_t)i * 5] = 0.1;
  const double lb[4] = {0, 0, 0, 0}, ub[4] = {5, 5, 5, 5};
  oracle_problem p{};
  p.model = ORACLE_QUADROTOR; p.n = 13; p.m = 4; p.horizon = N; p.dt = dt; p.integrator = ORACLE_RK4; p.has_control_box = 1;
  const double prm[11] = {1.0, 0.01, 0, 0, 0, 0.01, 0, 0, 0, 0.02, 0.2};
  std::memcpy(p.model_params, prm, sizeof(prm));
  p.Q = Qa.data(); p.R = Ra.data(); p.Qf = Qfa.data(); p.lb = lb; p.ub = ub;
  oracle_options oo;
  oracle_default_options(&oo);
  oo.max_iterations = 15; oo.reg_initial_value = 1e-4; oo.ls_max_iterations = 15;
  double worst = 0.0;
  for (int b = 0; b < B; ++b) {
    std::vector<double> Xo((size_t)(N + 1) * 13, 0.0), Uo((size_t)N * 4, hover), Ko((size_t)N * 52), ko((size_t)N * 4);
    double x0a[13] = {0}, xr[13] = {0};
    x0a[3] = 1.0; x0a[0] = 0.05 * b;
    xr[0] = 3.0 - 0.1 * b; xr[2] = 2.0 + 0.05 * b; xr[3] = 1.0;
    for (int t = 0; t <= N; ++t) std::memcpy(&Xo[(size_t)t * 13], x0a, sizeof(x0a));
    oracle_result res;
    oracle_solve(&p, &oo, x0a, xr, nullptr, Xo.data(), Uo.data(), Ko.data(), ko.data(), &res, nullptr);
    CHECK(res.iterations == sols[(size_t)b].iterations_completed);
    worst = std::max(worst, std::fabs(res.final_objective - sols[(size_t)b].final_objective) / std::fabs(res.final_objective));
    for (int t = 0; t < N; ++t)
      for (int i = 0; i < 4; ++i) CHECK(sols[(size_t)b].control_trajectory[(size_t)t][i] >= 0.0 && sols[(size_t)b].control_trajectory[(size_t)t][i] <= 5.0);
  }
  CHECK(worst < 1e-6);
  std::printf("  quadrotor batch of %d: worst final-cost rel err vs oracle %.2e\n", B, worst);
  // multi-device entry point: the batch cut into shards, one host thread + handle per shard.  Every available device is
  // used (a one-GPU box runs three shards on device 0); results must be bitwise those of the single-handle call,
  // in problem order.
  std::vector<int> devs = b200::availableDevices();
  CHECK(!devs.empty());
  if (devs.size() == 1) devs = {0, 0, 0};
  for (int b = 0; b < B; ++b) {  // the first call left its solution in the contexts: back to the original nominal
    std::vector<Eigen::VectorXd> X((size_t)N + 1, batch[(size_t)b]->getInitialState()), U((size_t)N, Eigen::VectorXd::Constant(4, hover));
    batch[(size_t)b]->setInitialTrajectory(X, U);
  }
  auto sharded = b200::solveBatch(batch, devs);
  CHECK(sharded.size() == (size_t)B);
  for (int b = 0; b < B; ++b) {
    CHECK(sharded[(size_t)b].final_objective == sols[(size_t)b].final_objective);
    CHECK(sharded[(size_t)b].iterations_completed == sols[(size_t)b].iterations_completed);
    CHECK((sharded[(size_t)b].control_trajectory[3] - sols[(size_t)b].control_trajectory[3]).norm() == 0.0);
    for (int i = 0; i < 4; ++i)
      for (int j = 0; j < 13; ++j) CHECK(sharded[(size_t)b].feedback_gains[7](i, j) == sols[(size_t)b].feedback_gains[7](i, j));
  }
  std::printf("  multi-device solveBatch over %zu shard(s): identical to the single-device call\n", devs.size());
  // mismatched batch is a setup error (exception), not an outcome
  CDDPOptions o2 = o;
  o2.max_iterations = 3;
  batch[1]->setOptions(o2);
  bool threw = false;
  try {
    b200::solveBatch(batch);
  } catch (const std::runtime_error &) {
    threw = true;
  }
  CHECK(threw);
}

// IPDDP through the registry, written like the reference's IPDDP tests (tests/cddp_core/test_ipddp_solver.cpp:552-620
// SolveUnicycle: status in {Optimal, Acceptable}) on the obstacle-avoidance problem of examples/python_portfolio_lib.py:
// 374-473 (ControlConstraint + BallConstraint), single solve and batched facade, with oracle parity.
void SolveUnicycleObstacleIPDDP() {
  const int N = 100, B = 6;
  const double dt = 0.03;
  CDDPOptions o;
  o.max_iterations = 150; o.tolerance = 1e-4; o.acceptable_tolerance = 1e-6; o.regularization.initial_value = 1e-6; o.verbose = false;
  o.return_iteration_info = true;
  const Eigen::MatrixXd Q = Eigen::MatrixXd::Zero(3, 3), R = diag({0.05, 0.05}), Qf = diag({100, 100, 50});
  std::vector<std::unique_ptr<CDDP>> owners;
  std::vector<CDDP *> batch;
  for (int b = 0; b < B; ++b) {
    const auto x0 = vec({0.0, 0.0, 0.0}), goal = vec({2.0 + 0.03 * b, 2.0 - 0.02 * b, M_PI / 2.0});
    auto c = std::make_unique<CDDP>(x0, goal, N, dt, std::make_unique<Unicycle>(dt, "euler"),
                                    std::make_unique<QuadraticObjective>(Q, R, Qf, goal, std::vector<Eigen::VectorXd>(), dt), o);
    // inserted in non-alphabetical order on purpose: the std::map orders them Ball < Control like the reference
    c->addPathConstraint("ControlConstraint", std::make_unique<ControlConstraint>(vec({-1.1, -M_PI}), vec({1.1, M_PI})));
    c->addPathConstraint("BallConstraint", std::make_unique<BallConstraint>(0.4, vec({1.0, 1.0})));
    std::vector<Eigen::VectorXd> X((size_t)N + 1, x0), U((size_t)N, vec({0.0, 0.0}));
    c->setInitialTrajectory(X, U);
    batch.push_back(c.get());
    owners.push_back(std::move(c));
  }
  CHECK(batch[0]->getTotalDualDim() == 5);
  CDDPSolution s0 = batch[0]->solve(SolverType::IPDDP);
  CHECK(s0.solver_name == "IPDDP");
  CHECK(s0.status_message == "OptimalSolutionFound" || s0.status_message == "AcceptableSolutionFound");
  CHECK(s0.final_barrier_mu > 0.0 && s0.final_barrier_mu < 1e-3);
  CHECK(s0.final_primal_infeasibility < 1e-3 && s0.final_complementary_infeasibility < 1e-3);
  CHECK(s0.history.barrier_mu.size() == s0.history.objective.size() && !s0.history.barrier_mu.empty());
  CHECK(batch[0]->cost_ == s0.final_objective && batch[0]->inf_pr_ == s0.final_primal_infeasibility);
  double mind = 1e9;
  for (int t = 0; t <= N; ++t) mind = std::min(mind, std::hypot(s0.state_trajectory[(size_t)t][0] - 1.0, s0.state_trajectory[(size_t)t][1] - 1.0));
  CHECK(mind > 0.4 - 1e-3);
  // fresh problems for the batched call (solve() left instance 0 at its solution)
  for (int b = 0; b < B; ++b) {
    std::vector<Eigen::VectorXd> X((size_t)N + 1, batch[(size_t)b]->getInitialState()), U((size_t)N, vec({0.0, 0.0}));
    batch[(size_t)b]->setInitialTrajectory(X, U);
  }
  std::vector<CDDPSolution> sols = b200::solveBatchIPDDP(batch);
  CHECK(sols[0].iterations_completed == s0.iterations_completed && sols[0].final_objective == s0.final_objective);
  // oracle parity
  const double Qa[9] = {0}, Ra[4] = {0.05, 0, 0, 0.05}, Qfa[9] = {100, 0, 0, 0, 100, 0, 0, 0, 50};
  oracle_problem p{};
  p.model = ORACLE_UNICYCLE; p.n = 3; p.m = 2; p.horizon = N; p.dt = dt; p.integrator = ORACLE_EULER; p.Q = Qa; p.R = Ra; p.Qf = Qfa;
  oracle_options oo;
  oracle_default_options(&oo);
  oo.max_iterations = 150; oo.tolerance = 1e-4; oo.acceptable_tolerance = 1e-6; oo.reg_initial_value = 1e-6;
  oracle_ipddp_options oi;
  oracle_ipddp_default_options(&oi);
  const double center[2] = {1.0, 1.0}, radius[1] = {0.4}, lb[2] = {-1.1, -M_PI}, ub[2] = {1.1, M_PI};
  oracle_constraint cs[2] = {{ORACLE_CON_BALL, 2, 1.0, center, radius}, {ORACLE_CON_CONTROL_BOX, 2, 1.0, lb, ub}};
  double worst = 0.0;
  int compared = 0;
  for (int b = 0; b < B; ++b) {
    std::vector<double> Xo((size_t)(N + 1) * 3), Uo((size_t)N * 2, 0.0), Ko((size_t)N * 6);
    const double x0a[3] = {0, 0, 0}, xr[3] = {2.0 + 0.03 * b, 2.0 - 0.02 * b, M_PI / 2.0};
    oracle_ipddp_result res;
    oracle_ipddp_solve(&p, &oo, &oi, cs, 2, x0a, xr, nullptr, Xo.data(), Uo.data(), Ko.data(), nullptr, nullptr, &res, nullptr);
    if (res.decision_margin < 1e-9) continue;  // roundoff-decided line search (tests/test_gpu_ipddp.py)
    ++compared;
    CHECK(res.iterations == sols[(size_t)b].iterations_completed);
    CHECK(std::string(oracle_status_string(res.status)) == sols[(size_t)b].status_message);
    worst = std::max(worst, std::fabs(res.final_objective - sols[(size_t)b].final_objective) / std::fabs(res.final_objective));
    CHECK(std::fabs(res.mu - sols[(size_t)b].final_barrier_mu) <= 1e-9 * res.mu);
  }
  CHECK(compared >= B / 2);
  CHECK(worst < 1e-6);
  std::printf("  unicycle obstacle IPDDP: %d iterations, %s, J=%.9f, mu=%.3e, min distance %.4f; batch of %d: worst cost rel err vs oracle %.2e\n",
              s0.iterations_completed, s0.status_message.c_str(), s0.final_objective, s0.final_barrier_mu, mind, B, worst);
  // BASELINE config #4 in full: + TerminalEqualityConstraint(goal) (test_ipddp_solver.cpp:1417-1419 pattern) -> terminal-equality branch
  {
    std::vector<Eigen::VectorXd> X((size_t)N + 1, batch[0]->getInitialState()), U((size_t)N, vec({0.0, 0.0}));
    batch[0]->setInitialTrajectory(X, U);
    const Eigen::VectorXd goal = batch[0]->getObjective().getReferenceState();
    batch[0]->addTerminalConstraint("TerminalEqualityConstraint", std::make_unique<TerminalEqualityConstraint>(goal));
    CDDPSolution st = batch[0]->solve("IPDDP");
    CHECK(st.status_message == "OptimalSolutionFound" || st.status_message == "AcceptableSolutionFound");
    double term = 0.0;
    for (int i = 0; i < 3; ++i) term = std::max(term, std::fabs(st.state_trajectory.back()[i] - goal[i]));
    CHECK(term < 1e-4);
    oracle_ipddp_options oit = oi;
    oit.terminal_equality = 1;
    std::vector<double> Xo((size_t)(N + 1) * 3), Uo((size_t)N * 2, 0.0), Ko((size_t)N * 6);
    const double x0a[3] = {0, 0, 0}, xr[3] = {goal[0], goal[1], goal[2]};
    oracle_ipddp_result res;
    oracle_ipddp_solve(&p, &oo, &oit, cs, 2, x0a, xr, nullptr, Xo.data(), Uo.data(), Ko.data(), nullptr, nullptr, &res, nullptr);
    if (res.decision_margin > 1e-9) {
      CHECK(res.iterations == st.iterations_completed);
      CHECK(std::fabs(res.final_objective - st.final_objective) <= 1e-6 * std::fabs(res.final_objective));
    }
    std::printf("  + terminal equality: %d iterations, %s, J=%.9f (oracle %.9f), |x_N - goal| = %.2e\n", st.iterations_completed,
                st.status_message.c_str(), st.final_objective, res.final_objective, term);
  }
  // setup errors: a terminal equality on another target / unsupported constraint classes are exceptions, never a CPU fallback
  bool threw = false;
  batch[1]->addTerminalConstraint("TerminalEqualityConstraint", std::make_unique<TerminalEqualityConstraint>(vec({9.0, 9.0, 0.0})));
  try {
    batch[1]->solve("IPDDP");
  } catch (const std::runtime_error &) {
    threw = true;
  }
  CHECK(threw);
}

// A user-defined DynamicalSystem — the reference's Bicycle (src/dynamics_model/bicycle.cpp:29-47), which is NOT one of the
// engine's built-in models — plugged in the way a cddp-cpp user would: subclass, host dynamics for the host-side API, plus
// getDeviceModel() returning the CUDA twin of getContinuousDynamics.  Solved through the registry; oracle parity against
// the oracle's native Bicycle.
class UserBicycle : public DynamicalSystem {
 public:
  UserBicycle(double dt, double wheelbase, std::string integ) : DynamicalSystem(4, 2, dt, std::move(integ)), wheelbase_(wheelbase) {}
  Eigen::VectorXd getContinuousDynamics(const Eigen::VectorXd &x, const Eigen::VectorXd &u, double) const override {
    Eigen::VectorXd xd(4);
    xd[0] = x[3] * std::cos(x[2]);
    xd[1] = x[3] * std::sin(x[2]);
    xd[2] = (x[3] / wheelbase_) * std::tan(u[1]);
    xd[3] = u[0];
    return xd;
  }
  bool getDeviceModel(DeviceModelDescriptor &d) const override {
    d.model = CDDP_B200_MODEL_USER;
    d.params[0] = wheelbase_;
    d.source =
        "template <class T> __device__ void cddp_user_dynamics(const T *x, const T *u, const double *p, T *xd) {\n"
        "  xd[0] = x[3] * cos(x[2]); xd[1] = x[3] * sin(x[2]); xd[2] = (x[3] / p[0]) * tan(u[1]); xd[3] = u[0];\n}\n";
    return true;
  }

 private:
  double wheelbase_;
};

void SolveUserDefinedBicycle() {
  const int N = 80;
  const double dt = 0.05;
  CDDPOptions o;
  o.max_iterations = 60; o.tolerance = 1e-5; o.acceptable_tolerance = 1e-7; o.regularization.initial_value = 1e-5; o.verbose = false;
  const auto x0 = vec({0.0, 0.0, 0.0, 0.0}), goal = vec({4.0, 3.0, M_PI / 2.0, 0.0});
  const Eigen::MatrixXd Q = diag({0, 0, 0, 0.01}), R = diag({0.1, 0.5}), Qf = diag({100, 100, 50, 10});
  CDDP c(x0, goal, N, dt, std::make_unique<UserBicycle>(dt, 2.0, "rk4"),
         std::make_unique<QuadraticObjective>(Q, R, Qf, goal, std::vector<Eigen::VectorXd>(), dt), o);
  c.addPathConstraint("ControlConstraint", std::make_unique<ControlConstraint>(vec({-2.0, -0.6}), vec({2.0, 0.6})));
  std::vector<Eigen::VectorXd> X((size_t)N + 1, x0), U((size_t)N, vec({0.2, 0.0}));
  c.setInitialTrajectory(X, U);
  CDDPSolution s = c.solve("CLDDP");
  CHECK(s.status_message == "OptimalSolutionFound" || s.status_message == "AcceptableSolutionFound");
  const double Qa[16] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0.01}, Ra[4] = {0.1, 0, 0, 0.5},
               Qfa[16] = {100, 0, 0, 0, 0, 100, 0, 0, 0, 0, 50, 0, 0, 0, 0, 10}, lb[2] = {-2.0, -0.6}, ub[2] = {2.0, 0.6};
  oracle_problem p{};
  p.model = ORACLE_BICYCLE; p.n = 4; p.m = 2; p.horizon = N; p.dt = dt; p.integrator = ORACLE_RK4; p.has_control_box = 1;
  p.model_params[0] = 2.0; p.Q = Qa; p.R = Ra; p.Qf = Qfa; p.lb = lb; p.ub = ub;
  oracle_options oo;
  oracle_default_options(&oo);
  oo.max_iterations = 60; oo.tolerance = 1e-5; oo.acceptable_tolerance = 1e-7; oo.reg_initial_value = 1e-5;
  std::vector<double> Xo((size_t)(N + 1) * 4, 0.0), Uo((size_t)N * 2, 0.0), Ko((size_t)N * 8), ko((size_t)N * 2);
  for (int t = 0; t < N; ++t) Uo[(size_t)t * 2] = 0.2;
  const double x0a[4] = {0, 0, 0, 0}, xr[4] = {4.0, 3.0, M_PI / 2.0, 0.0};
  oracle_result res;
  oracle_solve(&p, &oo, x0a, xr, nullptr, Xo.data(), Uo.data(), Ko.data(), ko.data(), &res, nullptr);
  CHECK(res.iterations == s.iterations_completed);
  CHECK(std::fabs(res.final_objective - s.final_objective) <= 1e-6 * std::fabs(res.final_objective));
  std::printf("  user-defined bicycle: %d iterations, %s, J=%.9f (oracle %.9f)\n", s.iterations_completed, s.status_message.c_str(),
              s.final_objective, res.final_objective);
  // a source that does not compile is a setup error carrying the compiler log
  class Broken : public UserBicycle {
   public:
    using UserBicycle::UserBicycle;
    bool getDeviceModel(DeviceModelDescriptor &d) const override {
      d.model = CDDP_B200_MODEL_USER;
      d.source = "template <class T> __device__ void cddp_user_dynamics(const T *x, const T *u, const double *p, T *xd) { xd[0] = oops; }";
      return true;
    }
  };
  CDDP c2(x0, goal, N, dt, std::make_unique<Broken>(dt, 2.0, "rk4"),
          std::make_unique<QuadraticObjective>(Q, R, Qf, goal, std::vector<Eigen::VectorXd>(), dt), o);
  bool threw = false;
  try {
    c2.solve("CLDDP");
  } catch (const std::runtime_error &e) {
    threw = std::string(e.what()).find("oops") != std::string::npos;
  }
  CHECK(threw);
}
}  // namespace

int main() {
  b200::registerSolvers();
  RUN(SolvePendulum);
  RUN(SolveQuadrotorBatch);
  RUN(SolveUnicycleObstacleIPDDP);
  RUN(SolveUserDefinedBicycle);
  return finish();
}
