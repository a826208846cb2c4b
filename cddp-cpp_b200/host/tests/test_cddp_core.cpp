// Facade / plugin-API contract of the host mirror, re-stating what the reference's own test file asserts
// (tests/cddp_core/test_cddp_core.cpp: registration :316-344, use :347-369, unknown solver :393-412, precedence
// :463-483, stale-trajectory re-init :547-577, reference-state plumbing :579-635, dual-dim bookkeeping :637-676) plus
// the objective/constraint unit facts (test_objective.cpp:39-128, test_constraint.cpp:22-69).  Needs NO GPU: the
// solver used is a mock, exactly like the reference's MockExternalSolver.
#include <algorithm>

#include "cddp_b200/b200_solver.hpp"
#include "check.hpp"

using namespace cddp;

namespace {

class MockExternalSolver : public ISolverAlgorithm {
 public:
  void initialize(CDDP &) override {}
  CDDPSolution solve(CDDP &context) override {
    CDDPSolution s;
    s.solver_name = getSolverName();
    s.status_message = "OptimalSolutionFound";
    s.iterations_completed = 5;
    s.solve_time_ms = 100.0;
    s.final_objective = 1.23;
    s.final_step_length = 1.0;
    for (int t = 0; t <= context.getHorizon(); ++t) s.time_points.push_back(t * context.getTimestep());
    s.state_trajectory.assign((size_t)context.getHorizon() + 1, Eigen::VectorXd::Zero(context.getStateDim()));
    s.control_trajectory.assign((size_t)context.getHorizon(), Eigen::VectorXd::Zero(context.getControlDim()));
    return s;
  }
  std::string getSolverName() const override { return "MockExternalSolver"; }
};
std::unique_ptr<ISolverAlgorithm> createMock() { return std::make_unique<MockExternalSolver>(); }

const int state_dim = 3, control_dim = 2, horizon = 10;
const double timestep = 0.1;

Eigen::VectorXd vec(std::initializer_list<double> l) {
  Eigen::VectorXd v((long)l.size());
  long i = 0;
  for (double x : l) v[i++] = x;
  return v;
}
Eigen::MatrixXd eye(int n, double s = 1.0) { return Eigen::MatrixXd::Identity(n, n) * s; }

std::unique_ptr<CDDP> make(const Eigen::VectorXd &x0, const Eigen::VectorXd &goal) {
  CDDPOptions o;
  o.max_iterations = 5;
  o.verbose = false;
  return std::make_unique<CDDP>(x0, goal, horizon, timestep, std::make_unique<Unicycle>(timestep, "euler"),
                                std::make_unique<QuadraticObjective>(eye(state_dim), eye(control_dim), eye(state_dim, 10.0), goal,
                                                                     std::vector<Eigen::VectorXd>(), timestep),
                                o);
}
bool contains(const std::vector<std::string> &v, const std::string &s) { return std::find(v.begin(), v.end(), s) != v.end(); }

void ExternalSolverRegistration() {
  CDDP::registerSolver("MockExternalSolver", createMock);
  CHECK(CDDP::isSolverRegistered("MockExternalSolver"));
  CHECK(!CDDP::isSolverRegistered("NonExistentSolver"));
  CHECK(contains(CDDP::getRegisteredSolvers(), "MockExternalSolver"));
  CDDP::registerSolver("MockSolver2", createMock);
  CHECK(CDDP::getRegisteredSolvers().size() >= 2);
}

void UseRegisteredExternalSolver() {
  auto c = make(vec({0, 0, 0}), vec({2, 2, M_PI / 2}));
  auto s = c->solve("MockExternalSolver");
  CHECK(s.solver_name == "MockExternalSolver");
  CHECK(s.status_message == "OptimalSolutionFound");
  CHECK(s.iterations_completed == 5);
  CHECK(s.final_objective == 1.23);
}

void UnknownSolverErrorHandling() {
  auto c = make(vec({0, 0, 0}), vec({2, 2, M_PI / 2}));
  auto s = c->solve("NonExistentSolver");  // an OUTCOME, not an exception (cddp_core.cpp:243-265)
  CHECK(s.solver_name == "NonExistentSolver");
  CHECK(s.status_message.find("UnknownSolver") != std::string::npos);
  CHECK(s.status_message.find("NonExistentSolver") != std::string::npos);
  CHECK(s.iterations_completed == 0);
  CHECK(s.final_step_length == 1.0);
}

void SolverPrecedence() {
  // the external registry is consulted FIRST, so registering "CLDDP" shadows whatever would otherwise serve it:
  // this is the drop-in route cddp::b200::registerSolvers() uses
  CDDP::registerSolver("CLDDP", createMock);
  auto c = make(vec({0, 0, 0}), vec({2, 2, M_PI / 2}));
  auto s = c->solve("CLDDP");
  CHECK(s.solver_name == "MockExternalSolver");
  CHECK(s.final_objective == 1.23);
  auto s2 = c->solve(SolverType::CLDDP);  // enum route maps to the same string
  CHECK(s2.solver_name == "MockExternalSolver");
  b200::registerSolvers();  // replaces the mock
  CHECK(CDDP::isSolverRegistered("CLDDP") && CDDP::isSolverRegistered("CLDDP_B200"));
}

void SolveReinitializesStaleTrajectoryDimensions() {
  auto c = make(vec({0.5, -0.5, 0.1}), vec({2, 2, M_PI / 2}));
  c->X_.assign((size_t)horizon + 1, Eigen::VectorXd::Zero(state_dim + 2));
  c->U_.assign((size_t)horizon, Eigen::VectorXd::Zero(control_dim + 1));
  auto s = c->solve("MockExternalSolver");
  CHECK(s.solver_name == "MockExternalSolver");
  CHECK(c->X_.size() == (size_t)horizon + 1 && c->U_.size() == (size_t)horizon);
  for (auto &x : c->X_) CHECK(x.size() == state_dim);
  for (auto &u : c->U_) CHECK(u.size() == control_dim);
  CHECK((c->X_.front() - vec({0.5, -0.5, 0.1})).norm() < 1e-15);
  CHECK(std::isinf(c->cost_));  // initializeProblemIfNecessary: cost_ = inf (cddp_core.cpp:297)
  CHECK(c->regularization_ == c->getOptions().regularization.initial_value);
}

void ReferenceStatePlumbing() {
  auto c = make(vec({0, 0, 0}), vec({2, 2, M_PI / 2}));
  std::vector<Eigen::VectorXd> refs;
  for (int t = 0; t <= horizon; ++t) refs.push_back(vec({0.1 * t, 0.2 * t, 0.0}));
  c->setReferenceStates(refs);
  CHECK((c->getReferenceState() - refs.back()).norm() < 1e-15);
  CHECK_NEAR(c->getObjective().running_cost(refs.front(), Eigen::VectorXd::Zero(control_dim), 0), 0.0, 1e-12);
  CHECK_NEAR(c->getObjective().terminal_cost(refs.back()), 0.0, 1e-12);
  c->setObjective(std::make_unique<QuadraticObjective>(eye(state_dim), eye(control_dim), eye(state_dim, 10.0), vec({9, 9, 9}),
                                                       std::vector<Eigen::VectorXd>(), timestep));
  CHECK_NEAR(c->getObjective().running_cost(refs[3], Eigen::VectorXd::Zero(control_dim), 3), 0.0, 1e-12);
  CHECK_NEAR(c->getObjective().terminal_cost(refs.back()), 0.0, 1e-12);
  // setInitialTrajectory overwrites the initial state with X[0] (cddp_core.cpp:139-141)
  std::vector<Eigen::VectorXd> X((size_t)horizon + 1, vec({7, 8, 9})), U((size_t)horizon, Eigen::VectorXd::Zero(control_dim));
  c->setInitialTrajectory(X, U);
  CHECK((c->getInitialState() - vec({7, 8, 9})).norm() == 0.0);
}

void DualDimBookkeepingAndNullConstraint() {
  auto c = make(vec({0, 0, 0}), vec({2, 2, M_PI / 2}));
  c->addPathConstraint("ControlConstraint", std::make_unique<ControlConstraint>(vec({-1, -2}), vec({1, 2})));
  CHECK(c->getTotalDualDim() == 2 * control_dim);
  c->addPathConstraint("ControlConstraint", std::make_unique<ControlConstraint>(vec({1, 2})));  // replace, not add
  CHECK(c->getTotalDualDim() == 2 * control_dim);
  CHECK(c->getConstraint<ControlConstraint>("ControlConstraint") != nullptr);
  CHECK(c->getConstraint<ControlConstraint>("SomethingElse") == nullptr);
  // terminal constraints share the dual-dimension bookkeeping (test_cddp_core.cpp:637-676)
  c->addTerminalConstraint("RepeatedTerminalConstraint", std::make_unique<TerminalEqualityConstraint>(vec({2, 2, 0})));
  CHECK(c->getTotalDualDim() == 2 * control_dim + state_dim);
  Eigen::MatrixXd A1(1, state_dim);
  A1(0, 0) = 1.0;
  c->addTerminalConstraint("RepeatedTerminalConstraint", std::make_unique<TerminalInequalityConstraint>(A1, vec({2.5})));
  CHECK(c->getTotalDualDim() == 2 * control_dim + 1);
  CHECK(c->getTerminalConstraint<TerminalInequalityConstraint>("RepeatedTerminalConstraint") != nullptr);
  CHECK(c->getTerminalConstraint<TerminalEqualityConstraint>("RepeatedTerminalConstraint") == nullptr);
  CHECK(c->getTerminalConstraintSet().size() == 1);
  CHECK_NEAR(c->getTerminalConstraint<TerminalInequalityConstraint>("RepeatedTerminalConstraint")->evaluate(vec({3, 0, 0}), vec({0, 0}))[0], 0.5, 1e-15);
  CHECK(c->removeTerminalConstraint("RepeatedTerminalConstraint"));
  CHECK(!c->removeTerminalConstraint("RepeatedTerminalConstraint"));
  CHECK(c->getTotalDualDim() == 2 * control_dim);
  TerminalEqualityConstraint te(vec({1, 2, 3}));
  CHECK((te.evaluate(vec({1.5, 2, 2}), vec({0, 0})) - vec({0.5, 0, -1})).norm() < 1e-15 && te.getName() == "TerminalEqualityConstraint");
  CHECK(c->removePathConstraint("ControlConstraint"));
  CHECK(!c->removePathConstraint("ControlConstraint"));
  CHECK(c->getTotalDualDim() == 0);
  bool threw = false;
  try {
    c->addPathConstraint("Null", nullptr);
  } catch (const std::runtime_error &) {
    threw = true;
  }
  CHECK(threw);
}

void SetupErrorsThrow() {
  CDDPOptions o;
  CDDP c(vec({0, 0, 0}), vec({1, 1, 0}), horizon, timestep, nullptr, nullptr, o);
  bool threw = false;
  try {
    c.solve("MockExternalSolver");
  } catch (const std::runtime_error &e) {
    threw = std::string(e.what()).find("Dynamical system must be set") != std::string::npos;
  }
  CHECK(threw);
  threw = false;
  try {
    c.getStateDim();
  } catch (const std::runtime_error &) {
    threw = true;
  }
  CHECK(threw);
}

void OptionsAndAlphas() {
  CDDPOptions o;  // defaults = the reference's member initialisers (options.hpp)
  CHECK(o.tolerance == 1e-5 && o.acceptable_tolerance == 1e-6 && o.max_iterations == 1 && !o.enable_parallel);
  CHECK(o.line_search.max_iterations == 11 && o.regularization.initial_value == 1e-6 && o.regularization.max_value == 1e7);
  CHECK(o.box_qp.max_iterations == 100 && o.box_qp.armijo_constant == 0.1 && o.filter.armijo_constant == 1e-4);
  auto c = make(vec({0, 0, 0}), vec({1, 1, 0}));
  CHECK(c->alphas_.size() == 11);
  for (size_t i = 0; i < c->alphas_.size(); ++i) CHECK(c->alphas_[i] == std::ldexp(1.0, -(int)i));
  o.line_search.max_iterations = 4;
  o.line_search.initial_step_size = 0.5;
  c->setOptions(o);  // rebuilds the schedule and resets alpha_pr_ (cddp_core.cpp:109-113)
  CHECK(c->alphas_.size() == 4 && c->alphas_[0] == 0.5 && c->alpha_pr_ == 0.5);
  c->regularization_ = 1e6;
  c->increaseRegularization();
  CHECK(c->regularization_ == 1e7 && c->isRegularizationLimitReached());
  c->regularization_ = 1e-10;
  c->decreaseRegularization();
  CHECK(c->regularization_ == 1e-10);
}

void QuadraticObjectiveIdentities() {  // test_objective.cpp:39-128
  const Eigen::VectorXd goal = vec({1.1, 0.6, 0.3});
  QuadraticObjective obj(eye(3), eye(2, 0.1), eye(3, 2.0), goal, {}, 0.1);
  std::vector<Eigen::VectorXd> X, U;
  for (int i = 0; i <= 5; ++i) X.push_back(vec({1.0 + 0.1 * i, 0.5 + 0.1 * i, 0.2 + 0.1 * i}));
  for (int i = 0; i < 5; ++i) U.push_back(vec({0.8, 0.5}));
  double expected = 0.0;
  for (int i = 0; i < 5; ++i) {
    const auto e = X[(size_t)i] - goal;
    expected += e.dot(e) * 0.1 + 0.1 * U[(size_t)i].dot(U[(size_t)i]) * 0.1;
  }
  expected += 2.0 * (X.back() - goal).dot(X.back() - goal);
  CHECK_NEAR(obj.evaluate(X, U), expected, 1e-12);
  auto [lx, lu] = obj.getRunningCostGradients(X[0], U[0], 0);
  CHECK(((X[0] - goal) * (2.0 * 0.1) - lx).norm() < 1e-15 && (U[0] * (2.0 * 0.1 * 0.1) - lu).norm() < 1e-15);
  auto [lxx, luu, lux] = obj.getRunningCostHessians(X[0], U[0], 0);
  CHECK_NEAR(lxx(1, 1), 0.2, 1e-15);
  CHECK_NEAR(luu(0, 0), 0.02, 1e-15);
  CHECK(lux.rows() == 2 && lux.cols() == 3 && lux(1, 2) == 0.0);
  CHECK_NEAR(obj.getFinalCostHessian(X[0])(2, 2), 4.0, 1e-15);
  CHECK((obj.getFinalCostGradient(X[0]) - (X[0] - goal) * 4.0).norm() < 1e-15);
}

void ControlConstraintFacts() {  // test_constraint.cpp:22-69
  ControlConstraint cc(vec({-1, -2}), vec({1, 2}));
  CHECK(cc.getName() == "ControlConstraint");
  const auto g = cc.evaluate(vec({0.5, 1.0}), vec({1.5, -2.5}));
  CHECK(g.size() == 4 && g[0] == -1.5 && g[1] == 2.5 && g[2] == 1.5 && g[3] == -2.5);
  CHECK((cc.rawLowerBound() - vec({-1, -2})).norm() == 0 && (cc.rawUpperBound() - vec({1, 2})).norm() == 0);
  CHECK((cc.clamp(vec({0.5, 1.0})) - vec({0.5, 1.0})).norm() == 0);
  CHECK((cc.clamp(vec({1.5, -2.5})) - vec({1.0, -2.0})).norm() == 0);
}

void ModelsHostSide() {
  Pendulum p(0.05, 1.0, 1.0, 0.0, "euler");  // test_finite_difference.cpp:27-69
  const auto A = p.getStateJacobian(vec({0.1, 0.0}), vec({0.0}), 0.0);
  CHECK_NEAR(A(1, 0), 9.81 * std::cos(0.1), 1e-12);
  CHECK(A(0, 1) == 1.0 && p.getControlJacobian(vec({0.1, 0.0}), vec({0.0}), 0.0)(1, 0) == 1.0);
  Eigen::MatrixXd I(3, 3);
  I(0, 0) = 0.01; I(1, 1) = 0.01; I(2, 2) = 0.02;
  Quadrotor q(0.01, 1.0, I, 0.2, "rk4");  // hover => f = 0 (test_quadrotor.cpp:166-212)
  Eigen::VectorXd x = Eigen::VectorXd::Zero(13);
  x[2] = 1.0; x[3] = 1.0;
  const auto xd = q.getContinuousDynamics(x, Eigen::VectorXd::Constant(4, 9.81 / 4.0), 0.0);
  for (int i = 0; i < 13; ++i) CHECK_NEAR(xd[i], 0.0, 1e-10);
  const auto xn = q.getDiscreteDynamics(x, Eigen::VectorXd::Constant(4, 9.81 / 4.0), 0.0);
  CHECK((xn - x).norm() < 1e-12);
  DeviceModelDescriptor d;
  CHECK(q.getDeviceModel(d) && d.model == 3 && d.params[0] == 1.0 && d.params[10] == 0.2 && d.params[9] == 0.02);
  struct HostOnly : DynamicalSystem {
    HostOnly() : DynamicalSystem(1, 1, 0.1, "euler") {}
    Eigen::VectorXd getContinuousDynamics(const Eigen::VectorXd &x, const Eigen::VectorXd &u, double) const override { return x * -1.0 + u; }
  } h;
  CHECK(!h.getDeviceModel(d));  // default: not device-capable
  CHECK_NEAR(h.getStateJacobian(vec({0.3}), vec({0.1}), 0.0)(0, 0), -1.0, 1e-9);  // FD default
}

void HostOnlyDynamicsIsRejectedNotFallenBack() {
  struct HostOnly : DynamicalSystem {
    HostOnly() : DynamicalSystem(1, 1, 0.1, "euler") {}
    Eigen::VectorXd getContinuousDynamics(const Eigen::VectorXd &x, const Eigen::VectorXd &u, double) const override { return x * -1.0 + u; }
  };
  b200::registerSolvers();
  CDDPOptions o;
  CDDP c(vec({1.0}), vec({0.0}), 5, 0.1, std::make_unique<HostOnly>(),
         std::make_unique<QuadraticObjective>(eye(1), eye(1), eye(1), vec({0.0}), std::vector<Eigen::VectorXd>(), 0.1), o);
  bool threw = false;
  try {
    c.solve("CLDDP");
  } catch (const std::runtime_error &e) {
    threw = std::string(e.what()).find("no CPU fallback") != std::string::npos;
  }
  CHECK(threw);
}

}  // namespace

int main() {
  RUN(ExternalSolverRegistration);
  RUN(UseRegisteredExternalSolver);
  RUN(UnknownSolverErrorHandling);
  RUN(SolverPrecedence);
  RUN(SolveReinitializesStaleTrajectoryDimensions);
  RUN(ReferenceStatePlumbing);
  RUN(DualDimBookkeepingAndNullConstraint);
  RUN(SetupErrorsThrow);
  RUN(OptionsAndAlphas);
  RUN(QuadraticObjectiveIdentities);
  RUN(ControlConstraintFacts);
  RUN(ModelsHostSide);
  RUN(HostOnlyDynamicsIsRejectedNotFallenBack);
  return finish();
}
