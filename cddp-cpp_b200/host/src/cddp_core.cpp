// Host mirror of the cddp-cpp facade, plugin base classes and the five device-resident models (see cddp.hpp for the
// reference file:line of every class).  Plain C++17; the only arithmetic here is what the reference's host-side API
// promises to callers (cost evaluation, dynamics of a single state, clamp) — solving happens in libcddp_b200.so.
#include <algorithm>
#include <cmath>
#include <iostream>
#include <limits>

#include "cddp_b200/cddp.hpp"
#include "../../../include/cddp_b200.h"

namespace cddp {

namespace detail {
std::vector<double> buildLineSearchAlphas(const LineSearchOptions &o) {
  // one implementation of the schedule for host and device: the C ABI's (cddp_context_utils.cpp:37-57)
  cddp_b200_options c;
  cddp_b200_default_options(&c);
  c.ls_max_iterations = o.max_iterations;
  c.ls_initial_step_size = o.initial_step_size;
  c.ls_min_step_size = o.min_step_size;
  c.ls_step_reduction_factor = o.step_reduction_factor;
  std::vector<double> a((size_t)std::max(2, o.max_iterations + 1));
  int cnt = 0;
  cddp_b200_build_alphas(&c, a.data(), (int)a.size(), &cnt);
  a.resize((size_t)cnt);
  return a;
}

static bool compatible(const std::vector<Eigen::VectorXd> &t, int size, int dim) {
  if ((int)t.size() != size) return false;
  for (const auto &v : t)
    if ((int)v.size() != dim) return false;
  return true;
}
}  // namespace detail

// ------------------------------------------------------------------------------------------------ DynamicalSystem
Eigen::VectorXd DynamicalSystem::getContinuousDynamics(const Eigen::VectorXd &, const Eigen::VectorXd &, double) const {
  throw std::logic_error("getContinuousDynamics must be overridden in the derived class.");
}

Eigen::VectorXd DynamicalSystem::getDiscreteDynamics(const Eigen::VectorXd &x, const Eigen::VectorXd &u, double t) const {
  const double dt = timestep_;
  auto axpy = [](const Eigen::VectorXd &a, double s, const Eigen::VectorXd &b) {
    Eigen::VectorXd r(a.size());
    for (long i = 0; i < a.size(); ++i) r[i] = a[i] + s * b[i];
    return r;
  };
  if (integration_type_ == "euler") return axpy(x, dt, getContinuousDynamics(x, u, t));
  if (integration_type_ == "heun") {
    const auto k1 = getContinuousDynamics(x, u, t);
    const auto k2 = getContinuousDynamics(axpy(x, dt, k1), u, t + dt);
    Eigen::VectorXd r(x.size());
    for (long i = 0; i < x.size(); ++i) r[i] = x[i] + 0.5 * dt * (k1[i] + k2[i]);
    return r;
  }
  if (integration_type_ == "rk3") {
    const auto k1 = getContinuousDynamics(x, u, t);
    const auto k2 = getContinuousDynamics(axpy(x, 0.5 * dt, k1), u, t + 0.5 * dt);
    Eigen::VectorXd xt(x.size());
    for (long i = 0; i < x.size(); ++i) xt[i] = x[i] - dt * k1[i] + 2.0 * dt * k2[i];
    const auto k3 = getContinuousDynamics(xt, u, t + dt);
    Eigen::VectorXd r(x.size());
    for (long i = 0; i < x.size(); ++i) r[i] = x[i] + (dt / 6.0) * (k1[i] + 4.0 * k2[i] + k3[i]);
    return r;
  }
  if (integration_type_ == "rk4") {
    const auto k1 = getContinuousDynamics(x, u, t);
    const auto k2 = getContinuousDynamics(axpy(x, 0.5 * dt, k1), u, t + 0.5 * dt);
    const auto k3 = getContinuousDynamics(axpy(x, 0.5 * dt, k2), u, t + 0.5 * dt);
    const auto k4 = getContinuousDynamics(axpy(x, dt, k3), u, t + dt);
    Eigen::VectorXd r(x.size());
    for (long i = 0; i < x.size(); ++i) r[i] = x[i] + (dt / 6.0) * (k1[i] + 2.0 * k2[i] + 2.0 * k3[i] + k4[i]);
    return r;
  }
  throw std::invalid_argument("Unknown integration type: " + integration_type_);
}

static Eigen::MatrixXd fd_jacobian(const std::function<Eigen::VectorXd(const Eigen::VectorXd &)> &f, const Eigen::VectorXd &z, int rows) {
  const double h = 2e-5;
  Eigen::MatrixXd J(rows, z.size());
  for (long j = 0; j < z.size(); ++j) {
    Eigen::VectorXd zp = z, zm = z;
    zp[j] += h;
    zm[j] -= h;
    const auto fp = f(zp), fm = f(zm);
    for (int i = 0; i < rows; ++i) J(i, j) = (fp[i] - fm[i]) / (2.0 * h);
  }
  return J;
}
Eigen::MatrixXd DynamicalSystem::getStateJacobian(const Eigen::VectorXd &x, const Eigen::VectorXd &u, double t) const {
  return fd_jacobian([&](const Eigen::VectorXd &z) { return getContinuousDynamics(z, u, t); }, x, state_dim_);
}
Eigen::MatrixXd DynamicalSystem::getControlJacobian(const Eigen::VectorXd &x, const Eigen::VectorXd &u, double t) const {
  return fd_jacobian([&](const Eigen::VectorXd &z) { return getContinuousDynamics(x, z, t); }, u, state_dim_);
}

// ------------------------------------------------------------------------------------------------ models
Pendulum::Pendulum(double timestep, double length, double mass, double damping, std::string integration_type)
    : DynamicalSystem(2, 1, timestep, std::move(integration_type)), length_(length), mass_(mass), damping_(damping) {}
Eigen::VectorXd Pendulum::getContinuousDynamics(const Eigen::VectorXd &x, const Eigen::VectorXd &u, double) const {
  Eigen::VectorXd xd(2);
  const double g = 9.81, inertia = mass_ * length_ * length_;
  xd[0] = x[1];
  xd[1] = (u[0] - damping_ * x[1] + mass_ * g * length_ * std::sin(x[0])) / inertia;
  return xd;
}
Eigen::MatrixXd Pendulum::getStateJacobian(const Eigen::VectorXd &x, const Eigen::VectorXd &, double) const {
  Eigen::MatrixXd A = Eigen::MatrixXd::Zero(2, 2);
  A(0, 1) = 1.0;
  A(1, 0) = (9.81 / length_) * std::cos(x[0]);
  A(1, 1) = -damping_ / (mass_ * length_ * length_);
  return A;
}
Eigen::MatrixXd Pendulum::getControlJacobian(const Eigen::VectorXd &, const Eigen::VectorXd &, double) const {
  Eigen::MatrixXd B = Eigen::MatrixXd::Zero(2, 1);
  B(1, 0) = 1.0 / (mass_ * length_ * length_);
  return B;
}
bool Pendulum::getDeviceModel(DeviceModelDescriptor &d) const {
  d.model = CDDP_B200_MODEL_PENDULUM;
  d.params[0] = length_; d.params[1] = mass_; d.params[2] = damping_;
  return true;
}

CartPole::CartPole(double timestep, std::string integration_type, double cart_mass, double pole_mass, double pole_length,
                   double gravity, double damping)
    : DynamicalSystem(4, 1, timestep, std::move(integration_type)), cart_mass_(cart_mass), pole_mass_(pole_mass),
      pole_length_(pole_length), gravity_(gravity), damping_(damping) {}
Eigen::VectorXd CartPole::getContinuousDynamics(const Eigen::VectorXd &x, const Eigen::VectorXd &u, double) const {
  Eigen::VectorXd xd(4);
  const double s = std::sin(x[1]), c = std::cos(x[1]), w = x[3], F = u[0];
  const double den = cart_mass_ + pole_mass_ * s * s;
  xd[0] = x[2];
  xd[1] = w;
  xd[2] = (F + pole_mass_ * s * (pole_length_ * w * w + gravity_ * c)) / den;
  xd[3] = (-F * c - pole_mass_ * pole_length_ * w * w * c * s - (cart_mass_ + pole_mass_) * gravity_ * s) / (pole_length_ * den);
  return xd;
}
bool CartPole::getDeviceModel(DeviceModelDescriptor &d) const {
  d.model = CDDP_B200_MODEL_CARTPOLE;
  d.params[0] = cart_mass_; d.params[1] = pole_mass_; d.params[2] = pole_length_; d.params[3] = gravity_; d.params[4] = damping_;
  return true;
}

Unicycle::Unicycle(double timestep, std::string integration_type) : DynamicalSystem(3, 2, timestep, std::move(integration_type)) {}
Eigen::VectorXd Unicycle::getContinuousDynamics(const Eigen::VectorXd &x, const Eigen::VectorXd &u, double) const {
  Eigen::VectorXd xd(3);
  xd[0] = u[0] * std::cos(x[2]);
  xd[1] = u[0] * std::sin(x[2]);
  xd[2] = u[1];
  return xd;
}
bool Unicycle::getDeviceModel(DeviceModelDescriptor &d) const {
  d.model = CDDP_B200_MODEL_UNICYCLE;
  return true;
}

Quadrotor::Quadrotor(double timestep, double mass, const Eigen::Matrix3d &I, double arm_length, std::string integration_type)
    : DynamicalSystem(13, 4, timestep, std::move(integration_type)), mass_(mass), arm_length_(arm_length) {
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) inertia_[i * 3 + j] = I(i, j);
}
Eigen::VectorXd Quadrotor::getContinuousDynamics(const Eigen::VectorXd &x, const Eigen::VectorXd &u, double) const {
  Eigen::VectorXd xd(13);
  double qw = x[3], qx = x[4], qy = x[5], qz = x[6];
  const double norm = std::sqrt(qw * qw + qx * qx + qy * qy + qz * qz);
  if (norm > 1e-6) { qw /= norm; qx /= norm; qy /= norm; qz /= norm; } else { qw = 1; qx = qy = qz = 0; }
  const double wx = x[10], wy = x[11], wz = x[12];
  xd[0] = x[7]; xd[1] = x[8]; xd[2] = x[9];
  xd[3] = -0.5 * (qx * wx + qy * wy + qz * wz);
  xd[4] = 0.5 * (qw * wx + qy * wz - qz * wy);
  xd[5] = 0.5 * (qw * wy - qx * wz + qz * wx);
  xd[6] = 0.5 * (qw * wz + qx * wy - qy * wx);
  const double thrust = u[0] + u[1] + u[2] + u[3];
  const double tau[3] = {arm_length_ * (u[0] - u[2]), arm_length_ * (u[1] - u[3]), 0.1 * (u[0] - u[1] + u[2] - u[3])};
  xd[7] = 2.0 * (qx * qz + qy * qw) * thrust / mass_;
  xd[8] = 2.0 * (qy * qz - qx * qw) * thrust / mass_;
  xd[9] = (1.0 - 2.0 * (qx * qx + qy * qy)) * thrust / mass_ - 9.81;
  const double *I = inertia_;
  const double Iw[3] = {I[0] * wx + I[1] * wy + I[2] * wz, I[3] * wx + I[4] * wy + I[5] * wz, I[6] * wx + I[7] * wy + I[8] * wz};
  const double r[3] = {tau[0] - (wy * Iw[2] - wz * Iw[1]), tau[1] - (wz * Iw[0] - wx * Iw[2]), tau[2] - (wx * Iw[1] - wy * Iw[0])};
  const double c00 = I[4] * I[8] - I[5] * I[7], c01 = I[2] * I[7] - I[1] * I[8], c02 = I[1] * I[5] - I[2] * I[4];
  const double c10 = I[5] * I[6] - I[3] * I[8], c11 = I[0] * I[8] - I[2] * I[6], c12 = I[2] * I[3] - I[0] * I[5];
  const double c20 = I[3] * I[7] - I[4] * I[6], c21 = I[1] * I[6] - I[0] * I[7], c22 = I[0] * I[4] - I[1] * I[3];
  const double id = 1.0 / (I[0] * c00 + I[1] * c10 + I[2] * c20);
  xd[10] = id * (c00 * r[0] + c01 * r[1] + c02 * r[2]);
  xd[11] = id * (c10 * r[0] + c11 * r[1] + c12 * r[2]);
  xd[12] = id * (c20 * r[0] + c21 * r[1] + c22 * r[2]);
  return xd;
}
bool Quadrotor::getDeviceModel(DeviceModelDescriptor &d) const {
  d.model = CDDP_B200_MODEL_QUADROTOR;
  d.params[0] = mass_;
  for (int i = 0; i < 9; ++i) d.params[1 + i] = inertia_[i];
  d.params[10] = arm_length_;
  return true;
}

LTISystem::LTISystem(const Eigen::MatrixXd &A, const Eigen::MatrixXd &B, double timestep, std::string integration_type)
    : DynamicalSystem((int)A.rows(), (int)B.cols(), timestep, std::move(integration_type)), A_(A), B_(B) {}
Eigen::VectorXd LTISystem::getDiscreteDynamics(const Eigen::VectorXd &x, const Eigen::VectorXd &u, double) const {
  Eigen::VectorXd r(state_dim_);
  for (int i = 0; i < state_dim_; ++i) {
    double s = 0.0, s2 = 0.0;
    for (int j = 0; j < state_dim_; ++j) s += A_(i, j) * x[j];
    for (int j = 0; j < control_dim_; ++j) s2 += B_(i, j) * u[j];
    r[i] = s + s2;
  }
  return r;
}
Eigen::VectorXd LTISystem::getContinuousDynamics(const Eigen::VectorXd &x, const Eigen::VectorXd &u, double t) const {
  const auto xn = getDiscreteDynamics(x, u, t);  // (x+ - x)/dt, lti_system.cpp:61-69
  Eigen::VectorXd r(state_dim_);
  for (int i = 0; i < state_dim_; ++i) r[i] = (xn[i] - x[i]) / timestep_;
  return r;
}
Eigen::MatrixXd LTISystem::getStateJacobian(const Eigen::VectorXd &, const Eigen::VectorXd &, double) const {
  Eigen::MatrixXd F(state_dim_, state_dim_);  // (A_d - I)/dt, lti_system.cpp:78-84
  for (int i = 0; i < state_dim_; ++i)
    for (int j = 0; j < state_dim_; ++j) F(i, j) = (A_(i, j) - (i == j ? 1.0 : 0.0)) / timestep_;
  return F;
}
Eigen::MatrixXd LTISystem::getControlJacobian(const Eigen::VectorXd &, const Eigen::VectorXd &, double) const {
  Eigen::MatrixXd F(state_dim_, control_dim_);
  for (int i = 0; i < state_dim_; ++i)
    for (int j = 0; j < control_dim_; ++j) F(i, j) = B_(i, j) / timestep_;
  return F;
}
bool LTISystem::getDeviceModel(DeviceModelDescriptor &d) const {
  d.model = CDDP_B200_MODEL_LTI;
  d.lti_A.resize((size_t)state_dim_ * state_dim_);
  d.lti_B.resize((size_t)state_dim_ * control_dim_);
  for (int i = 0; i < state_dim_; ++i) {
    for (int j = 0; j < state_dim_; ++j) d.lti_A[(size_t)i * state_dim_ + j] = A_(i, j);
    for (int j = 0; j < control_dim_; ++j) d.lti_B[(size_t)i * control_dim_ + j] = B_(i, j);
  }
  return true;
}

// ------------------------------------------------------------------------------------------------ QuadraticObjective
QuadraticObjective::QuadraticObjective(const Eigen::MatrixXd &Q, const Eigen::MatrixXd &R, const Eigen::MatrixXd &Qf,
                                       const Eigen::VectorXd &reference_state, const std::vector<Eigen::VectorXd> &reference_states,
                                       double timestep)
    : Q0_(Q), R0_(R), Q_(Q * timestep), R_(R * timestep), Qf_(Qf), timestep_(timestep) {  // objective.cpp:38-39
  reference_state_ = reference_state;
  reference_states_ = reference_states;
}
const Eigen::VectorXd &QuadraticObjective::refAt(int index) const {  // objective.cpp:84-88
  if (!reference_states_.empty() && index >= 0 && index < (int)reference_states_.size()) return reference_states_[(size_t)index];
  return reference_state_;
}
static double quad(const Eigen::MatrixXd &M, const Eigen::VectorXd &e) {
  double s = 0.0;
  for (long j = 0; j < e.size(); ++j) {
    double r = 0.0;
    for (long i = 0; i < e.size(); ++i) r += e[i] * M(i, j);
    s += r * e[j];
  }
  return s;
}
double QuadraticObjective::running_cost(const Eigen::VectorXd &x, const Eigen::VectorXd &u, int index) const {
  return quad(Q_, x - refAt(index)) + quad(R_, u);
}
double QuadraticObjective::terminal_cost(const Eigen::VectorXd &x) const { return quad(Qf_, x - reference_state_); }
double QuadraticObjective::evaluate(const std::vector<Eigen::VectorXd> &X, const std::vector<Eigen::VectorXd> &U) const {
  double J = 0.0;
  for (size_t t = 0; t + 1 < X.size(); ++t) J += running_cost(X[t], U[t], (int)t);
  return J + terminal_cost(X.back());
}
std::tuple<Eigen::VectorXd, Eigen::VectorXd> QuadraticObjective::getRunningCostGradients(const Eigen::VectorXd &x, const Eigen::VectorXd &u,
                                                                                        int index) const {
  return {(Q_ * (x - refAt(index))) * 2.0, (R_ * u) * 2.0};
}
std::tuple<Eigen::MatrixXd, Eigen::MatrixXd, Eigen::MatrixXd> QuadraticObjective::getRunningCostHessians(const Eigen::VectorXd &x,
                                                                                                       const Eigen::VectorXd &u, int) const {
  return {Q_ * 2.0, R_ * 2.0, Eigen::MatrixXd::Zero(u.size(), x.size())};
}
Eigen::VectorXd QuadraticObjective::getFinalCostGradient(const Eigen::VectorXd &x) const { return (Qf_ * (x - reference_state_)) * 2.0; }
Eigen::MatrixXd QuadraticObjective::getFinalCostHessian(const Eigen::VectorXd &) const { return Qf_ * 2.0; }

// ------------------------------------------------------------------------------------------------ ControlConstraint
ControlConstraint::ControlConstraint(const Eigen::VectorXd &lower_bound, const Eigen::VectorXd &upper_bound, double scale_factor)
    : Constraint("ControlConstraint"), lower_bound_(lower_bound), upper_bound_(upper_bound), scale_factor_(scale_factor) {}
ControlConstraint::ControlConstraint(const Eigen::VectorXd &upper_bound)
    : Constraint("ControlConstraint"), lower_bound_(upper_bound * -1.0), upper_bound_(upper_bound) {}
Eigen::VectorXd ControlConstraint::evaluate(const Eigen::VectorXd &, const Eigen::VectorXd &u, int) const {
  Eigen::VectorXd g(2 * u.size());
  for (long i = 0; i < u.size(); ++i) { g[i] = -u[i] * scale_factor_; g[u.size() + i] = u[i] * scale_factor_; }  // (:163-172)
  return g;
}
Eigen::VectorXd ControlConstraint::getLowerBound() const { return Eigen::VectorXd::Constant(2 * upper_bound_.size(), -std::numeric_limits<double>::infinity()); }
Eigen::VectorXd ControlConstraint::getUpperBound() const {
  Eigen::VectorXd r(2 * upper_bound_.size());
  for (long i = 0; i < upper_bound_.size(); ++i) {  // [-lb; ub] * scale (:156-159)
    r[i] = -lower_bound_[i] * scale_factor_;
    r[upper_bound_.size() + i] = upper_bound_[i] * scale_factor_;
  }
  return r;
}
Eigen::VectorXd ControlConstraint::clamp(const Eigen::VectorXd &v) const {
  Eigen::VectorXd r(v.size());
  for (long i = 0; i < v.size(); ++i) r[i] = std::min(std::max(v[i], lower_bound_[i]), upper_bound_[i]);
  return r;
}

// ------------------------------------------------------------------------------------------------ CDDP facade
static std::map<std::string, std::function<std::unique_ptr<ISolverAlgorithm>()>> &registry() {
  static std::map<std::string, std::function<std::unique_ptr<ISolverAlgorithm>()>> r;  // unguarded, like the reference's (cddp_core.cpp:34-35)
  return r;
}

CDDP::CDDP(const Eigen::VectorXd &initial_state, const Eigen::VectorXd &reference_state, int horizon, double timestep,
           std::unique_ptr<DynamicalSystem> system, std::unique_ptr<Objective> objective, const CDDPOptions &options)
    : alpha_pr_(options.line_search.initial_step_size), regularization_(options.regularization.initial_value),
      terminal_regularization_(options.regularization.initial_value), system_(std::move(system)), objective_(std::move(objective)),
      initial_state_(initial_state), reference_state_(reference_state), horizon_(horizon), timestep_(timestep), options_(options) {
  if (objective_ && reference_state.size() > 0 && !reference_state.isZero()) objective_->setReferenceState(reference_state_);
  alphas_ = detail::buildLineSearchAlphas(options_.line_search);
}
int CDDP::getStateDim() const {
  if (!system_) throw std::runtime_error("Dynamical system not set.");
  return system_->getStateDim();
}
int CDDP::getControlDim() const {
  if (!system_) throw std::runtime_error("Dynamical system not set.");
  return system_->getControlDim();
}
void CDDP::setDynamicalSystem(std::unique_ptr<DynamicalSystem> system) {
  system_ = std::move(system);
  initialized_ = false;
}
void CDDP::setInitialState(const Eigen::VectorXd &initial_state) {
  initial_state_ = initial_state;
  if (!X_.empty() && X_[0].size() == initial_state.size()) X_[0] = initial_state_;
}
void CDDP::setReferenceState(const Eigen::VectorXd &reference_state) {
  reference_state_ = reference_state;
  if (objective_) objective_->setReferenceState(reference_state_);
  reference_states_.clear();
  reference_states_.push_back(reference_state_);
}
void CDDP::setReferenceStates(const std::vector<Eigen::VectorXd> &reference_states) {
  reference_states_ = reference_states;
  if (!reference_states_.empty()) reference_state_ = reference_states_.back();
  if (objective_) {
    if (!reference_states_.empty()) objective_->setReferenceState(reference_state_);
    objective_->setReferenceStates(reference_states_);
  }
}
void CDDP::setHorizon(int horizon) {
  horizon_ = horizon;
  initialized_ = false;
}
void CDDP::setOptions(const CDDPOptions &options) {
  options_ = options;
  alphas_ = detail::buildLineSearchAlphas(options_.line_search);
  alpha_pr_ = options_.line_search.initial_step_size;
}
void CDDP::setObjective(std::unique_ptr<Objective> objective) {
  objective_ = std::move(objective);
  if (objective_ && !reference_states_.empty()) {
    objective_->setReferenceState(reference_state_);
    objective_->setReferenceStates(reference_states_);
  } else if (objective_ && reference_state_.size() > 0 && !reference_state_.isZero()) {
    objective_->setReferenceState(reference_state_);
  }
}
void CDDP::setInitialTrajectory(const std::vector<Eigen::VectorXd> &X, const std::vector<Eigen::VectorXd> &U) {
  if (X.size() != (size_t)(horizon_ + 1) || U.size() != (size_t)horizon_)
    std::cerr << "Warning: Provided initial trajectory dimensions do not match horizon." << std::endl;
  X_ = X;
  U_ = U;
  if (!X_.empty()) initial_state_ = X_[0];  // quirk kept: the trajectory's first state becomes the initial state (cddp_core.cpp:139-141)
}
void CDDP::addPathConstraint(std::string name, std::unique_ptr<Constraint> constraint) {
  if (!constraint) throw std::runtime_error("Cannot add null constraint.");  // cddp_context_utils.cpp:82-84
  const int dual = constraint->getDualDim();
  auto it = path_constraint_set_.find(name);
  if (it != path_constraint_set_.end()) total_dual_dim_ -= it->second->getDualDim();
  path_constraint_set_[name] = std::move(constraint);
  total_dual_dim_ += dual;
  initialized_ = false;
}
bool CDDP::removePathConstraint(const std::string &name) {
  auto it = path_constraint_set_.find(name);
  if (it == path_constraint_set_.end()) return false;
  total_dual_dim_ -= it->second->getDualDim();
  path_constraint_set_.erase(it);
  initialized_ = false;
  return true;
}

void CDDP::addTerminalConstraint(std::string name, std::unique_ptr<Constraint> constraint) {
  if (!constraint) throw std::runtime_error("Cannot add null constraint.");
  const int dual = constraint->getDualDim();
  auto it = terminal_constraint_set_.find(name);
  if (it != terminal_constraint_set_.end()) total_dual_dim_ -= it->second->getDualDim();
  terminal_constraint_set_[name] = std::move(constraint);
  total_dual_dim_ += dual;
  initialized_ = false;
}
bool CDDP::removeTerminalConstraint(const std::string &name) {
  auto it = terminal_constraint_set_.find(name);
  if (it == terminal_constraint_set_.end()) return false;
  total_dual_dim_ -= it->second->getDualDim();
  terminal_constraint_set_.erase(it);
  initialized_ = false;
  return true;
}
Eigen::VectorXd TerminalInequalityConstraint::getLowerBound() const {
  return Eigen::VectorXd::Constant(b_.size(), -std::numeric_limits<double>::infinity());
}

static const char *solverTypeToString(SolverType t) {
  switch (t) {
    case SolverType::CLDDP: return "CLDDP";
    case SolverType::LogDDP: return "LogDDP";
    case SolverType::IPDDP: return "IPDDP";
    case SolverType::MSIPDDP: return "MSIPDDP";
  }
  return "CLDDP";
}
CDDPSolution CDDP::solve(SolverType t) { return solve(std::string(solverTypeToString(t))); }

std::unique_ptr<ISolverAlgorithm> CDDP::createSolver(const std::string &solver_type) {
  auto it = registry().find(solver_type);  // external registry FIRST (cddp_core.cpp:216-219)
  if (it != registry().end()) return it->second();
  return nullptr;  // no built-in CPU solvers in this repository (no CPU fallback)
}

CDDPSolution CDDP::solve(const std::string &solver_type) {  // cddp_core.cpp:235-270
  initializeProblemIfNecessary();
  solver_ = createSolver(solver_type);
  if (!solver_) {
    CDDPSolution s;
    s.solver_name = solver_type;
    s.status_message = "UnknownSolver - No solver registered for '" + solver_type + "'";
    s.iterations_completed = 0;
    s.solve_time_ms = 0.0;
    s.final_objective = 0.0;
    s.final_step_length = 1.0;
    return s;
  }
  solver_->initialize(*this);
  return solver_->solve(*this);
}

void CDDP::initializeProblemIfNecessary() {  // cddp_core.cpp:272-306
  if (initialized_) return;
  if (!system_) throw std::runtime_error("Dynamical system must be set before solving.");
  if (!objective_) throw std::runtime_error("Objective function must be set before solving.");
  const int n = system_->getStateDim(), m = system_->getControlDim();
  const bool keep = options_.warm_start && detail::compatible(X_, horizon_ + 1, n) && detail::compatible(U_, horizon_, m);
  if (!keep) {
    if (!detail::compatible(X_, horizon_ + 1, n)) X_.assign((size_t)(horizon_ + 1), Eigen::VectorXd::Zero(n));
    if (!detail::compatible(U_, horizon_, m)) U_.assign((size_t)horizon_, Eigen::VectorXd::Zero(m));
  }
  X_[0] = initial_state_;
  const double inf = std::numeric_limits<double>::infinity();
  cost_ = merit_function_ = inf_pr_ = inf_du_ = inf_comp_ = inf;
  regularization_ = terminal_regularization_ = options_.regularization.initial_value;
  initialized_ = true;
}
void CDDP::increaseRegularization() { regularization_ = std::min(regularization_ * options_.regularization.update_factor, options_.regularization.max_value); }
void CDDP::decreaseRegularization() { regularization_ = std::max(regularization_ / options_.regularization.update_factor, options_.regularization.min_value); }
bool CDDP::isRegularizationLimitReached() const { return regularization_ >= options_.regularization.max_value; }

void CDDP::registerSolver(const std::string &name, std::function<std::unique_ptr<ISolverAlgorithm>()> factory) { registry()[name] = std::move(factory); }
bool CDDP::isSolverRegistered(const std::string &name) { return registry().find(name) != registry().end(); }
std::vector<std::string> CDDP::getRegisteredSolvers() {
  std::vector<std::string> names;
  for (const auto &p : registry()) names.push_back(p.first);
  return names;
}

}  // namespace cddp
