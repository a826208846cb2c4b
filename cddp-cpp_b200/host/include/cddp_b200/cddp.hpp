// cddp.hpp — C++ host mirror of the part of the cddp-cpp API that the CLDDP hot path touches, written from
// scratch over the C ABI in include/cddp_b200.h.  Same namespace, class names, method names, argument meaning and
// error behaviour as the reference so that code (and tests) written against cddp-cpp read the same here:
//
//   reference (astomodynamics/cddp-cpp @ f71fa80)                      here
//   include/cddp-cpp/cddp_core/options.hpp:41-66,93-105,208-251        LineSearchOptions, RegularizationOptions, BoxQPOptions,
//                                                                      SolverSpecificFilterOptions, CDDPOptions
//   include/cddp-cpp/cddp_core/cddp_core.hpp:54-103                    CDDPSolution (+ History)
//   include/cddp-cpp/cddp_core/cddp_core.hpp:186-210                   ISolverAlgorithm
//   include/cddp-cpp/cddp_core/cddp_core.hpp:212-442                   CDDP (+ static solver registry, src/cddp_core/cddp_core.cpp)
//   include/cddp-cpp/cddp_core/dynamical_system.hpp:33-152             DynamicalSystem (host virtuals) + getDeviceModel() hook [ADDITION]
//   include/cddp-cpp/cddp_core/objective.hpp:23-201                    Objective, QuadraticObjective
//   include/cddp-cpp/cddp_core/constraint.hpp:31-251                   Constraint, ControlConstraint
//   include/cddp-cpp/dynamics_model/{pendulum,cartpole,unicycle,quadrotor,lti_system}.hpp   the five device-resident models
//
// What is NOT mirrored (out of scope): the IPDDP/LogDDP/MSIPDDP solvers, autodiff, Hessians of the
// dynamics (CLDDP never reads them, clddp_solver.cpp:79-204), the other constraint / model classes, printing.
// There is no built-in CPU solver: CDDP::solve("CLDDP") resolves through the registry to the B200 solver
// (b200_solver.hpp) after cddp::b200::registerSolvers(); without it the reference's own "UnknownSolver" outcome is
// returned (cddp_core.cpp:243-265).
#pragma once

#if __has_include(<Eigen/Dense>)
#include <Eigen/Dense>
#else
#include "eigen_standin.hpp"
#endif

#include <functional>
#include <map>
#include <memory>
#include <stdexcept>
#include <string>
#include <tuple>
#include <vector>

namespace cddp {

// ------------------------------------------------------------------------------------------------ options
struct LineSearchOptions {  // options.hpp:41-50
  int max_iterations = 11;
  double initial_step_size = 1.0;
  double min_step_size = 1e-8;
  double step_reduction_factor = 0.5;
};
struct RegularizationOptions {  // options.hpp:58-66
  double initial_value = 1e-6;
  double update_factor = 10.0;
  double max_value = 1e7;
  double min_value = 1e-10;
  double step_initial_value = 1.0;
};
struct BoxQPOptions {  // boxqp.hpp:30-41
  int max_iterations = 100;
  double min_gradient_norm = 1e-8;
  double min_relative_improvement = 1e-8;
  double step_decrease_factor = 0.6;
  double min_step_size = 1e-22;
  double armijo_constant = 0.1;
  bool verbose = false;
};
struct SolverSpecificFilterOptions {  // options.hpp:93-105 (CLDDP reads armijo_constant only)
  double merit_acceptance_threshold = 1e-6;
  double violation_acceptance_threshold = 1e-6;
  double max_violation_threshold = 1e4;
  double min_violation_for_armijo_check = 1e-7;
  double armijo_constant = 1e-4;
};
enum class BarrierStrategy { ADAPTIVE, MONOTONIC, IPOPT };  // options.hpp:28-33
struct SolverSpecificBarrierOptions {  // options.hpp:75-87
  double mu_initial = 1e-0;
  double mu_min_value = 1e-10;
  double mu_update_factor = 0.5;
  double mu_update_power = 1.2;
  double min_fraction_to_boundary = 0.99;
  BarrierStrategy strategy = BarrierStrategy::ADAPTIVE;
};
struct IPDDPAlgorithmOptions {  // options.hpp:148-186 (the members the cold-start path-constraint path reads)
  double dual_var_init_scale = 1e-1;
  double slack_var_init_scale = 1e-2;
  double barrier_tol_mult = 0.1;
  double barrier_update_dual_weight = 0.01;
  double mu_kappa_epsilon = 10.0;
  bool check_state_stationarity = false;
  std::string theta_norm = "l1";
  int max_filter_size = 5;
  double theta_0_floor = 1.0;
  bool warmstart_repair = false;
  double jacobian_regularization_value = 1e-8;
  double jacobian_regularization_exponent = 0.25;
  SolverSpecificBarrierOptions barrier;
};
struct CDDPOptions {  // options.hpp:208-251
  double tolerance = 1e-5;
  double acceptable_tolerance = 1e-6;
  int max_iterations = 1;
  double max_cpu_time = 0.0;
  bool verbose = true;
  bool debug = false;
  bool print_solver_header = true;
  bool print_solver_options = false;
  bool use_ilqr = true;
  bool enable_parallel = false;
  int num_threads = 1;
  bool return_iteration_info = false;
  bool warm_start = false;
  double termination_scaling_max_factor = 100.0;
  LineSearchOptions line_search;
  RegularizationOptions regularization;
  BoxQPOptions box_qp;
  SolverSpecificFilterOptions filter;
  IPDDPAlgorithmOptions ipddp;
};

namespace detail {
std::vector<double> buildLineSearchAlphas(const LineSearchOptions &options);  // cddp_context_utils.cpp:37-57
}

// ------------------------------------------------------------------------------------------------ plugins
enum class SolverType { CLDDP, LogDDP, IPDDP, MSIPDDP };  // cddp_core.hpp:43-48

// [ADDITION to the reference surface] what a DynamicalSystem must provide to run inside the sm_100a kernels.
// The reference's virtuals are host-only Eigen calls; a model that cannot fill this is rejected by the B200
// solver with std::runtime_error (no CPU fallback).  model = CDDP_B200_MODEL_* (include/cddp_b200.h).
// A user-defined DynamicalSystem runs on the device by returning model = CDDP_B200_MODEL_USER and `source` = CUDA C++
// defining `template <class T> __device__ void cddp_user_dynamics(const T *x, const T *u, const double *p, T *xdot)`
// (include/cddp_b200.h, cddp_b200_create_ex): the device twin of its getContinuousDynamics / getContinuousDynamicsAutodiff.
struct DeviceModelDescriptor {
  int model = -1;
  double params[16] = {0};
  std::vector<double> lti_A, lti_B;  // row-major, LTI only
  std::string source;                // CDDP_B200_MODEL_USER only
};

class DynamicalSystem {  // dynamical_system.hpp:33-152
 public:
  DynamicalSystem(int state_dim, int control_dim, double timestep, std::string integration_type)
      : state_dim_(state_dim), control_dim_(control_dim), timestep_(timestep), integration_type_(std::move(integration_type)) {}
  virtual ~DynamicalSystem() {}
  virtual Eigen::VectorXd getContinuousDynamics(const Eigen::VectorXd &state, const Eigen::VectorXd &control, double time) const;
  // euler / heun / rk3 / rk4 on getContinuousDynamics (dynamical_system.cpp:28-83)
  virtual Eigen::VectorXd getDiscreteDynamics(const Eigen::VectorXd &state, const Eigen::VectorXd &control, double time) const;
  // continuous-time Jacobians; the base implementation is central finite differences (h = 2e-5, helper.hpp:96-147) —
  // the reference's base uses autodiff, which is unavailable here; every device model overrides with closed forms
  virtual Eigen::MatrixXd getStateJacobian(const Eigen::VectorXd &state, const Eigen::VectorXd &control, double time) const;
  virtual Eigen::MatrixXd getControlJacobian(const Eigen::VectorXd &state, const Eigen::VectorXd &control, double time) const;
  virtual std::tuple<Eigen::MatrixXd, Eigen::MatrixXd> getJacobians(const Eigen::VectorXd &state, const Eigen::VectorXd &control,
                                                                   double time) const {
    return {getStateJacobian(state, control, time), getControlJacobian(state, control, time)};
  }
  virtual bool getDeviceModel(DeviceModelDescriptor &) const { return false; }
  int getStateDim() const { return state_dim_; }
  int getControlDim() const { return control_dim_; }
  double getTimestep() const { return timestep_; }
  const std::string &getIntegrationType() const { return integration_type_; }

 protected:
  int state_dim_, control_dim_;
  double timestep_;
  std::string integration_type_;
};

class Pendulum : public DynamicalSystem {  // pendulum.hpp:38-42, pendulum.cpp:29-66
 public:
  Pendulum(double timestep, double length = 1.0, double mass = 1.0, double damping = 0.0, std::string integration_type = "euler");
  Eigen::VectorXd getContinuousDynamics(const Eigen::VectorXd &, const Eigen::VectorXd &, double) const override;
  Eigen::MatrixXd getStateJacobian(const Eigen::VectorXd &, const Eigen::VectorXd &, double) const override;
  Eigen::MatrixXd getControlJacobian(const Eigen::VectorXd &, const Eigen::VectorXd &, double) const override;
  bool getDeviceModel(DeviceModelDescriptor &) const override;

 private:
  double length_, mass_, damping_;
};

class CartPole : public DynamicalSystem {  // cartpole.hpp:49-55, cartpole.cpp:38-103
 public:
  CartPole(double timestep, std::string integration_type = "rk4", double cart_mass = 1.0, double pole_mass = 0.2,
           double pole_length = 0.5, double gravity = 9.81, double damping = 0.0);
  Eigen::VectorXd getContinuousDynamics(const Eigen::VectorXd &, const Eigen::VectorXd &, double) const override;
  bool getDeviceModel(DeviceModelDescriptor &) const override;

 private:
  double cart_mass_, pole_mass_, pole_length_, gravity_, damping_;
};

class Unicycle : public DynamicalSystem {  // unicycle.hpp:39-40, unicycle.cpp:28-66
 public:
  explicit Unicycle(double timestep, std::string integration_type = "euler");
  Eigen::VectorXd getContinuousDynamics(const Eigen::VectorXd &, const Eigen::VectorXd &, double) const override;
  bool getDeviceModel(DeviceModelDescriptor &) const override;
};

class Quadrotor : public DynamicalSystem {  // quadrotor.hpp:34-35, quadrotor.cpp:33-96
 public:
  Quadrotor(double timestep, double mass, const Eigen::Matrix3d &inertia_matrix, double arm_length,
            std::string integration_type = "euler");
  Eigen::VectorXd getContinuousDynamics(const Eigen::VectorXd &, const Eigen::VectorXd &, double) const override;
  bool getDeviceModel(DeviceModelDescriptor &) const override;

 private:
  double mass_, arm_length_, inertia_[9];
};

class LTISystem : public DynamicalSystem {  // lti_system.hpp:52-55, lti_system.cpp:71-92 (A, B are the DISCRETE matrices)
 public:
  LTISystem(const Eigen::MatrixXd &A, const Eigen::MatrixXd &B, double timestep, std::string integration_type = "euler");
  Eigen::VectorXd getContinuousDynamics(const Eigen::VectorXd &, const Eigen::VectorXd &, double) const override;
  Eigen::VectorXd getDiscreteDynamics(const Eigen::VectorXd &, const Eigen::VectorXd &, double) const override;
  Eigen::MatrixXd getStateJacobian(const Eigen::VectorXd &, const Eigen::VectorXd &, double) const override;
  Eigen::MatrixXd getControlJacobian(const Eigen::VectorXd &, const Eigen::VectorXd &, double) const override;
  bool getDeviceModel(DeviceModelDescriptor &) const override;

 private:
  Eigen::MatrixXd A_, B_;
};

class Objective {  // objective.hpp:23-120
 public:
  virtual ~Objective() = default;
  virtual double evaluate(const std::vector<Eigen::VectorXd> &states, const std::vector<Eigen::VectorXd> &controls) const = 0;
  virtual double running_cost(const Eigen::VectorXd &state, const Eigen::VectorXd &control, int index) const = 0;
  virtual double terminal_cost(const Eigen::VectorXd &final_state) const = 0;
  virtual std::tuple<Eigen::VectorXd, Eigen::VectorXd> getRunningCostGradients(const Eigen::VectorXd &state,
                                                                             const Eigen::VectorXd &control, int index) const = 0;
  virtual std::tuple<Eigen::MatrixXd, Eigen::MatrixXd, Eigen::MatrixXd> getRunningCostHessians(const Eigen::VectorXd &state,
                                                                                            const Eigen::VectorXd &control,
                                                                                            int index) const = 0;
  virtual Eigen::VectorXd getFinalCostGradient(const Eigen::VectorXd &final_state) const = 0;
  virtual Eigen::MatrixXd getFinalCostHessian(const Eigen::VectorXd &final_state) const = 0;
  virtual Eigen::VectorXd getReferenceState() const { return reference_state_; }
  virtual std::vector<Eigen::VectorXd> getReferenceStates() const { return reference_states_; }
  virtual void setReferenceState(const Eigen::VectorXd &reference_state) { reference_state_ = reference_state; }
  virtual void setReferenceStates(const std::vector<Eigen::VectorXd> &reference_states) { reference_states_ = reference_states; }

 protected:
  Eigen::VectorXd reference_state_;
  std::vector<Eigen::VectorXd> reference_states_;
};

class QuadraticObjective : public Objective {  // objective.hpp:122-201, objective.cpp:30-154
 public:
  QuadraticObjective(const Eigen::MatrixXd &Q, const Eigen::MatrixXd &R, const Eigen::MatrixXd &Qf,
                     const Eigen::VectorXd &reference_state = Eigen::VectorXd::Zero(0),
                     const std::vector<Eigen::VectorXd> &reference_states = std::vector<Eigen::VectorXd>(), double timestep = 0.1);
  double evaluate(const std::vector<Eigen::VectorXd> &states, const std::vector<Eigen::VectorXd> &controls) const override;
  double running_cost(const Eigen::VectorXd &state, const Eigen::VectorXd &control, int index) const override;
  double terminal_cost(const Eigen::VectorXd &final_state) const override;
  std::tuple<Eigen::VectorXd, Eigen::VectorXd> getRunningCostGradients(const Eigen::VectorXd &, const Eigen::VectorXd &, int) const override;
  std::tuple<Eigen::MatrixXd, Eigen::MatrixXd, Eigen::MatrixXd> getRunningCostHessians(const Eigen::VectorXd &, const Eigen::VectorXd &,
                                                                                    int) const override;
  Eigen::VectorXd getFinalCostGradient(const Eigen::VectorXd &final_state) const override;
  Eigen::MatrixXd getFinalCostHessian(const Eigen::VectorXd &final_state) const override;
  // Q_ = Q*dt, R_ = R*dt are what the reference stores (objective.cpp:38-39); the UNscaled inputs are kept for the C ABI
  const Eigen::MatrixXd &getQ() const { return Q_; }
  const Eigen::MatrixXd &getR() const { return R_; }
  const Eigen::MatrixXd &getQf() const { return Qf_; }
  const Eigen::MatrixXd &unscaledQ() const { return Q0_; }
  const Eigen::MatrixXd &unscaledR() const { return R0_; }
  double getTimestep() const { return timestep_; }

 private:
  const Eigen::VectorXd &refAt(int index) const;
  Eigen::MatrixXd Q0_, R0_, Q_, R_, Qf_;
  double timestep_;
};

class Constraint {  // constraint.hpp:31-138 (the members CLDDP and the facade use)
 public:
  explicit Constraint(std::string name) : name_(std::move(name)) {}
  virtual ~Constraint() = default;
  const std::string &getName() const { return name_; }
  virtual int getDualDim() const { return 0; }
  virtual Eigen::VectorXd evaluate(const Eigen::VectorXd &state, const Eigen::VectorXd &control, int index = 0) const = 0;
  virtual Eigen::VectorXd getLowerBound() const = 0;
  virtual Eigen::VectorXd getUpperBound() const = 0;

 private:
  std::string name_;
};

class ControlConstraint : public Constraint {  // constraint.hpp:144-251 (BoxConstraint<Control>)
 public:
  // scale_factor multiplies evaluate(), the IP upper bound and the Jacobians (constraint.hpp:147-218); the raw bounds
  // and clamp() — what CLDDP's BoxQP uses — are unscaled
  ControlConstraint(const Eigen::VectorXd &lower_bound, const Eigen::VectorXd &upper_bound, double scale_factor = 1.0);
  explicit ControlConstraint(const Eigen::VectorXd &upper_bound);  // symmetric: lower = -upper
  int getDualDim() const override { return 2 * (int)upper_bound_.size(); }
  Eigen::VectorXd evaluate(const Eigen::VectorXd &state, const Eigen::VectorXd &control, int index = 0) const override;  // [-u; u]
  Eigen::VectorXd getLowerBound() const override;
  Eigen::VectorXd getUpperBound() const override;  // [-lb; ub]
  const Eigen::VectorXd &rawLowerBound() const { return lower_bound_; }  // :222
  const Eigen::VectorXd &rawUpperBound() const { return upper_bound_; }  // :223
  Eigen::VectorXd clamp(const Eigen::VectorXd &v) const;                 // :225-228
  double getScaleFactor() const { return scale_factor_; }

 private:
  Eigen::VectorXd lower_bound_, upper_bound_;
  double scale_factor_ = 1.0;
};

class StateConstraint : public Constraint {  // constraint.hpp:144-251 (BoxConstraint<State>)
 public:
  StateConstraint(const Eigen::VectorXd &lower_bound, const Eigen::VectorXd &upper_bound, double scale_factor = 1.0)
      : Constraint("StateConstraint"), lower_bound_(lower_bound), upper_bound_(upper_bound), scale_factor_(scale_factor) {}
  int getDualDim() const override { return 2 * (int)upper_bound_.size(); }
  Eigen::VectorXd evaluate(const Eigen::VectorXd &state, const Eigen::VectorXd &, int = 0) const override {  // [-x; x] * scale
    const long k = state.size();
    Eigen::VectorXd g(2 * k);
    for (long i = 0; i < k; ++i) { g[i] = -state[i] * scale_factor_; g[k + i] = state[i] * scale_factor_; }
    return g;
  }
  Eigen::VectorXd getLowerBound() const override { return Eigen::VectorXd::Constant(getDualDim(), -std::numeric_limits<double>::infinity()); }
  Eigen::VectorXd getUpperBound() const override {  // [-lb; ub] * scale (:156-159)
    const long k = upper_bound_.size();
    Eigen::VectorXd g(2 * k);
    for (long i = 0; i < k; ++i) { g[i] = -lower_bound_[i] * scale_factor_; g[k + i] = upper_bound_[i] * scale_factor_; }
    return g;
  }
  const Eigen::VectorXd &rawLowerBound() const { return lower_bound_; }
  const Eigen::VectorXd &rawUpperBound() const { return upper_bound_; }
  double getScaleFactor() const { return scale_factor_; }

 private:
  Eigen::VectorXd lower_bound_, upper_bound_;
  double scale_factor_;
};

class LinearConstraint : public Constraint {  // constraint.hpp:253-318: A x <= b (scale_factor is stored, never applied)
 public:
  LinearConstraint(const Eigen::MatrixXd &A, const Eigen::VectorXd &b, double scale_factor = 1.0)
      : Constraint("LinearConstraint"), A_(A), b_(b), scale_factor_(scale_factor) {}
  int getDualDim() const override { return (int)b_.size(); }
  Eigen::VectorXd evaluate(const Eigen::VectorXd &state, const Eigen::VectorXd &, int = 0) const override { return A_ * state; }
  Eigen::VectorXd getLowerBound() const override { return Eigen::VectorXd::Constant(b_.size(), -std::numeric_limits<double>::infinity()); }
  Eigen::VectorXd getUpperBound() const override { return b_; }
  const Eigen::MatrixXd &getA() const { return A_; }
  double getScaleFactor() const { return scale_factor_; }

 private:
  Eigen::MatrixXd A_;
  Eigen::VectorXd b_;
  double scale_factor_;
};

class BallConstraint : public Constraint {  // constraint.hpp:320-440: -scale |x[0:dim] - center|^2 <= -scale r^2
 public:
  BallConstraint(double radius, const Eigen::VectorXd &center, double scale_factor = 1.0)
      : Constraint("BallConstraint"), radius_(radius), center_(center), scale_factor_(scale_factor) {}
  int getDualDim() const override { return 1; }
  Eigen::VectorXd evaluate(const Eigen::VectorXd &state, const Eigen::VectorXd &, int = 0) const override {
    double sq = 0.0;
    for (long i = 0; i < center_.size(); ++i) sq += (state[i] - center_[i]) * (state[i] - center_[i]);
    return Eigen::VectorXd::Constant(1, -(scale_factor_ * sq));
  }
  Eigen::VectorXd getLowerBound() const override { return Eigen::VectorXd::Constant(1, -std::numeric_limits<double>::infinity()); }
  Eigen::VectorXd getUpperBound() const override { return Eigen::VectorXd::Constant(1, -(radius_ * radius_) * scale_factor_); }
  const Eigen::VectorXd &getCenter() const { return center_; }
  double getRadius() const { return radius_; }
  double getScaleFactor() const { return scale_factor_; }

 private:
  double radius_;
  Eigen::VectorXd center_;
  double scale_factor_;
};

// Terminal constraints (terminal_constraint.hpp:29-158).  CLDDP never reads them (clddp_solver.cpp looks up
// "ControlConstraint" only); they are mirrored so that CDDP::addTerminalConstraint and the dual-dimension bookkeeping
// behave as in the reference (test_cddp_core.cpp:637-676).
class TerminalConstraint : public Constraint {
 public:
  explicit TerminalConstraint(const std::string &name) : Constraint(name) {}
};
class TerminalEqualityConstraint : public TerminalConstraint {  // x_N - target = 0
 public:
  explicit TerminalEqualityConstraint(const Eigen::VectorXd &target_state, const std::string &name = "TerminalEqualityConstraint")
      : TerminalConstraint(name), target_state_(target_state) {}
  int getDualDim() const override { return (int)target_state_.size(); }
  Eigen::VectorXd evaluate(const Eigen::VectorXd &final_state, const Eigen::VectorXd &, int = 0) const override {
    if (final_state.size() != target_state_.size())
      throw std::invalid_argument("TerminalEqualityConstraint: final_state dimension mismatch.");
    return final_state - target_state_;
  }
  Eigen::VectorXd getLowerBound() const override { return Eigen::VectorXd::Zero(target_state_.size()); }
  Eigen::VectorXd getUpperBound() const override { return Eigen::VectorXd::Zero(target_state_.size()); }
  const Eigen::VectorXd &getTargetState() const { return target_state_; }

 private:
  Eigen::VectorXd target_state_;
};
class TerminalInequalityConstraint : public TerminalConstraint {  // A x_N - b <= 0
 public:
  TerminalInequalityConstraint(const Eigen::MatrixXd &A, const Eigen::VectorXd &b, const std::string &name = "TerminalInequalityConstraint")
      : TerminalConstraint(name), A_(A), b_(b) {}
  int getDualDim() const override { return (int)b_.size(); }
  Eigen::VectorXd evaluate(const Eigen::VectorXd &final_state, const Eigen::VectorXd &, int = 0) const override {
    return (A_ * final_state) - b_;
  }
  Eigen::VectorXd getLowerBound() const override;
  Eigen::VectorXd getUpperBound() const override { return Eigen::VectorXd::Zero(b_.size()); }

 private:
  Eigen::MatrixXd A_;
  Eigen::VectorXd b_;
};

// ------------------------------------------------------------------------------------------------ solution / solver / facade
struct CDDPSolution {  // cddp_core.hpp:54-103
  std::string solver_name;
  std::string status_message = "Running";
  int iterations_completed = 0;
  double solve_time_ms = 0.0;
  double final_objective = 0.0;
  double final_step_length = 0.0;
  double final_regularization = 0.0;
  std::vector<double> time_points;
  std::vector<Eigen::VectorXd> state_trajectory;
  std::vector<Eigen::VectorXd> control_trajectory;
  std::vector<Eigen::MatrixXd> feedback_gains;
  double final_primal_infeasibility = 0.0;
  double final_dual_infeasibility = 0.0;
  double final_complementary_infeasibility = 0.0;
  double final_barrier_mu = 0.0;
  struct History {
    std::vector<double> objective, merit_function, step_length_primal, step_length_dual, dual_infeasibility, primal_infeasibility,
        complementary_infeasibility, barrier_mu, regularization;
  } history;
};

class CDDP;

class ISolverAlgorithm {  // cddp_core.hpp:186-210
 public:
  virtual ~ISolverAlgorithm() = default;
  virtual void initialize(CDDP &context) = 0;
  virtual CDDPSolution solve(CDDP &context) = 0;
  virtual std::string getSolverName() const = 0;
};

class CDDP {  // cddp_core.hpp:212-442, cddp_core.cpp
 public:
  CDDP(const Eigen::VectorXd &initial_state, const Eigen::VectorXd &reference_state, int horizon, double timestep,
       std::unique_ptr<DynamicalSystem> system = nullptr, std::unique_ptr<Objective> objective = nullptr,
       const CDDPOptions &options = CDDPOptions());
  virtual ~CDDP() = default;

  const DynamicalSystem &getSystem() const { return *system_; }
  const Objective &getObjective() const { return *objective_; }
  const Eigen::VectorXd &getInitialState() const { return initial_state_; }
  const Eigen::VectorXd &getReferenceState() const { return reference_state_; }
  const std::vector<Eigen::VectorXd> &getReferenceStates() const { return reference_states_; }
  int getHorizon() const { return horizon_; }
  double getTimestep() const { return timestep_; }
  int getStateDim() const;
  int getControlDim() const;
  int getTotalDualDim() const { return total_dual_dim_; }
  const CDDPOptions &getOptions() const { return options_; }
  const std::map<std::string, std::unique_ptr<Constraint>> &getConstraintSet() const { return path_constraint_set_; }
  const std::map<std::string, std::unique_ptr<Constraint>> &getTerminalConstraintSet() const { return terminal_constraint_set_; }

  void setDynamicalSystem(std::unique_ptr<DynamicalSystem> system);
  void setInitialState(const Eigen::VectorXd &initial_state);
  void setReferenceState(const Eigen::VectorXd &reference_state);
  void setReferenceStates(const std::vector<Eigen::VectorXd> &reference_states);
  void setHorizon(int horizon);
  void setTimestep(double timestep) { timestep_ = timestep; }
  void setOptions(const CDDPOptions &options);
  void setObjective(std::unique_ptr<Objective> objective);
  void setInitialTrajectory(const std::vector<Eigen::VectorXd> &X, const std::vector<Eigen::VectorXd> &U);
  void addPathConstraint(std::string constraint_name, std::unique_ptr<Constraint> constraint);
  bool removePathConstraint(const std::string &constraint_name);
  void addTerminalConstraint(std::string constraint_name, std::unique_ptr<Constraint> constraint);
  bool removeTerminalConstraint(const std::string &constraint_name);

  template <typename T>
  T *getConstraint(const std::string &name) const {  // exact name AND dynamic type (clddp_solver.cpp:85-86)
    auto it = path_constraint_set_.find(name);
    if (it == path_constraint_set_.end()) return nullptr;
    return dynamic_cast<T *>(it->second.get());
  }

  template <typename T>
  T *getTerminalConstraint(const std::string &name) const {
    auto it = terminal_constraint_set_.find(name);
    if (it == terminal_constraint_set_.end()) return nullptr;
    return dynamic_cast<T *>(it->second.get());
  }

  CDDPSolution solve(SolverType solver_type = SolverType::CLDDP);
  CDDPSolution solve(const std::string &solver_type);

  static void registerSolver(const std::string &solver_name, std::function<std::unique_ptr<ISolverAlgorithm>()> factory);
  static bool isSolverRegistered(const std::string &solver_name);
  static std::vector<std::string> getRegisteredSolvers();

  // public iterate state shared with solver strategies (cddp_core.hpp:323-342)
  std::vector<Eigen::VectorXd> X_, U_;
  double cost_ = 0.0, merit_function_ = 0.0, inf_pr_ = 0.0, inf_du_ = 0.0, inf_comp_ = 0.0, step_norm_ = 0.0;
  bool initialized_ = false;
  std::vector<double> alphas_;
  double alpha_pr_ = 1.0, alpha_du_ = 0.0;
  double regularization_ = 0.0, terminal_regularization_ = 0.0;

  double getCurrentCost() const { return cost_; }
  double getCurrentRegularization() const { return regularization_; }
  void increaseRegularization();
  void decreaseRegularization();
  bool isRegularizationLimitReached() const;

 protected:
  virtual std::unique_ptr<ISolverAlgorithm> createSolver(const std::string &solver_type);

 private:
  void initializeProblemIfNecessary();  // cddp_core.cpp:272-306 (private in the reference too, cddp_core.hpp:441)
  std::unique_ptr<DynamicalSystem> system_;
  std::unique_ptr<Objective> objective_;
  std::map<std::string, std::unique_ptr<Constraint>> path_constraint_set_;
  std::map<std::string, std::unique_ptr<Constraint>> terminal_constraint_set_;
  Eigen::VectorXd initial_state_, reference_state_;
  std::vector<Eigen::VectorXd> reference_states_;
  int horizon_;
  double timestep_;
  CDDPOptions options_;
  int total_dual_dim_ = 0;
  std::unique_ptr<ISolverAlgorithm> solver_;
};

}  // namespace cddp
