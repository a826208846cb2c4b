// B200 CLDDP solver behind the reference's plugin point (ISolverAlgorithm + CDDP::registerSolver,
// include/cddp-cpp/cddp_core/cddp_core.hpp:186-210,305-319; registry consulted before the built-ins,
// src/cddp_core/cddp_core.cpp:213-233) and the batched facade the reference lacks.
// Both marshal cddp::CDDP objects into the C ABI (include/cddp_b200.h); no solver arithmetic on the host.
#pragma once
#include "cddp.hpp"

namespace cddp {
namespace b200 {

// Drop-in for cddp::CLDDPSolver (src/cddp_core/clddp_solver.cpp): one CDDP = a batch of one.
class CLDDPSolver : public ISolverAlgorithm {
 public:
  explicit CLDDPSolver(int device = 0) : device_(device) {}
  void initialize(CDDP &context) override;   // validates that the problem can run on the device (throws std::runtime_error)
  CDDPSolution solve(CDDP &context) override;
  std::string getSolverName() const override { return "CLDDP"; }

 private:
  int device_;
};

// Drop-in for cddp::IPDDPSolver (src/cddp_core/ipddp_solver.cpp): cold start, use_ilqr = true, path inequality
// constraints of the kinds ControlConstraint / StateConstraint / LinearConstraint / BallConstraint, no terminal
// constraints.  Anything else (terminal constraints, warm_start, use_ilqr = false, other constraint classes, LTISystem)
// is a setup error: std::runtime_error, never a CPU fallback.
class IPDDPSolver : public ISolverAlgorithm {
 public:
  explicit IPDDPSolver(int device = 0) : device_(device) {}
  void initialize(CDDP &context) override;
  CDDPSolution solve(CDDP &context) override;
  std::string getSolverName() const override { return "IPDDP"; }

 private:
  int device_;
};

// Batched IPDDP facade: same contract as solveBatch; the constraint sets must be identical across the batch.
std::vector<CDDPSolution> solveBatchIPDDP(const std::vector<CDDP *> &problems, int device = 0);

// Registers the B200 solver under "CLDDP" (shadows the name the reference's built-in uses — SolverPrecedence
// semantics, tests/cddp_core/test_cddp_core.cpp:463-483) and under "CLDDP_B200"; likewise "IPDDP" / "IPDDP_B200".
void registerSolvers(int device = 0);

// Batched facade: B structurally identical problems (same model + parameters, objective weights, horizon, timestep,
// options, control box) that differ in initial state, reference state(s) and initial trajectory, solved in ONE
// launch sequence.  Updates every context's X_, U_, cost_, inf_du_, alpha_pr_, regularization_ exactly as
// CDDP::solve would, and returns one CDDPSolution per problem.  Throws std::runtime_error on setup errors
// (mismatched problems, host-only dynamics, non-quadratic objective, CUDA failure); solve OUTCOMES are status
// strings, never exceptions (cddp_solver_base.cpp:69,82,162).
std::vector<CDDPSolution> solveBatch(const std::vector<CDDP *> &problems, int device = 0);

// The same facade over SEVERAL devices of one node: the batch is cut into contiguous shards of
// ceil(B / G) problems, shard g goes to devices[g] (an entry may repeat), every shard is driven by its own host thread
// with its own C-ABI handle and CUDA stream, and nothing is exchanged between devices — the instances share only
// read-only constants, which every handle uploads for itself.  Solutions come back in problem order.  An exception on
// any shard is rethrown after all threads have joined.  availableDevices() = {0, ..., cudaGetDeviceCount() - 1}.
std::vector<CDDPSolution> solveBatch(const std::vector<CDDP *> &problems, const std::vector<int> &devices);
std::vector<CDDPSolution> solveBatchIPDDP(const std::vector<CDDP *> &problems, const std::vector<int> &devices);
std::vector<int> availableDevices();

}  // namespace b200
}  // namespace cddp
