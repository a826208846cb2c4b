// Batched CLDDP forward rollout + line search — one warp per trajectory, one lane per alpha.
//
// Reference behaviour followed:
//   CLDDPSolver::forwardPass          src/cddp_core/clddp_solver.cpp:215-262
//   CDDPSolverBase::performForwardPass src/cddp_core/cddp_solver_base.cpp:248-263 (sequential
//       semantics: FIRST accepted alpha wins — every alpha is rolled out in parallel here, and the
//       lowest-index accepted lane is selected, which is the same decision)
//   accept / reject bookkeeping       src/cddp_core/cddp_solver_base.cpp:124-139, 206-218
//   regularisation schedule           src/cddp_core/cddp_core.cpp:308-326
//   CLDDPSolver::checkConvergence     src/cddp_core/clddp_solver.cpp:264-277
//   QuadraticObjective costs          src/cddp_core/objective.cpp:80-98
//   ControlConstraint::clamp          include/cddp-cpp/cddp_core/constraint.hpp:225-228
//
// Lane 0 (alpha_0, the full step) writes its rollout straight into the candidate buffers; if a
// smaller alpha is the first accepted one the warp replays that alpha once (all lanes in lockstep,
// lane 0 writing).  Accepting = flipping the instance's nominal/candidate buffer index.
#include "engine.h"
#include "kernels_forward.cuh"
#include "user_model_host.h"

namespace cddp_b200 {

namespace {

using namespace kern;

// LTI: runtime dimensions, discrete dynamics x+ = A_d x + B_d u (lti_system.cpp:71-76)
__global__ void __launch_bounds__(kWarpsPerCta * 32) forward_lti_kernel(Constants c, DeviceState d, int mode) {
  constexpr int MS = CDDP_B200_MAX_N, MC = CDDP_B200_MAX_M;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int b = slot_instance(d, blockIdx.x * kWarpsPerCta + warp);
  if (b >= d.B) return;
  if (mode == FW_ITERATE && d.status[b] != CDDP_B200_STATUS_RUNNING) return;
  const int n = d.n, m = d.m, N = d.N, na = c.num_alphas;
  const int cur = d.cur[b];
  const double *Xn = d.X[cur] + (size_t)b * (N + 1) * n;
  const double *Un = d.U[cur] + (size_t)b * N * m;
  double *Xc = d.X[cur ^ 1] + (size_t)b * (N + 1) * n;
  double *Uc = d.U[cur ^ 1] + (size_t)b * N * m;
  const double *gK = d.K + (size_t)b * N * m * n;
  const double *gk = d.kff + (size_t)b * N * m;
  const double *xref = d.xref + (size_t)b * n;
  const double *rtraj = d.ref_traj ? d.ref_traj + (size_t)b * (N + 1) * n : nullptr;

  auto rollout = [&](double alpha, bool writer) -> double {
    double x[MS], xn[MS], u[MC];
    for (int i = 0; i < n; ++i) x[i] = d.x0[(size_t)b * n + i];
    double J = 0.0;
    for (int t = 0; t < N; ++t) {
      for (int i = 0; i < m; ++i) {
        double acc = 0.0;
        for (int j = 0; j < n; ++j) acc += gK[((size_t)t * m + i) * n + j] * (x[j] - Xn[(size_t)t * n + j]);
        u[i] = Un[(size_t)t * m + i] + alpha * gk[(size_t)t * m + i] + acc;
      }
      if (c.has_box)
        for (int i = 0; i < m; ++i) u[i] = clamp_box(u[i], c.lb[i], c.ub[i]);
      const double *ref = rtraj ? rtraj + (size_t)t * n : xref;
      double sx = 0.0, su = 0.0;
      for (int j = 0; j < n; ++j) {
        double r = 0.0;
        for (int i = 0; i < n; ++i) r += (x[i] - ref[i]) * (0.5 * c.Qdt2[i * n + j]);
        sx += r * (x[j] - ref[j]);
      }
      for (int j = 0; j < m; ++j) {
        double r = 0.0;
        for (int i = 0; i < m; ++i) r += u[i] * (0.5 * c.Rdt2[i * m + j]);
        su += r * u[j];
      }
      J += sx + su;
      if (writer) {
        for (int i = 0; i < n; ++i) Xc[(size_t)t * n + i] = x[i];
        for (int i = 0; i < m; ++i) Uc[(size_t)t * m + i] = u[i];
      }
      for (int i = 0; i < n; ++i) {
        double s = 0.0, s2 = 0.0;
        for (int j = 0; j < n; ++j) s += c.mp.lti_A[i * n + j] * x[j];
        for (int j = 0; j < m; ++j) s2 += c.mp.lti_B[i * m + j] * u[j];
        xn[i] = s + s2;
      }
      for (int i = 0; i < n; ++i) x[i] = xn[i];
    }
    double sx = 0.0;
    for (int j = 0; j < n; ++j) {
      double r = 0.0;
      for (int i = 0; i < n; ++i) r += (x[i] - xref[i]) * (0.5 * c.Qf2[i * n + j]);
      sx += r * (x[j] - xref[j]);
    }
    J += sx;
    if (writer)
      for (int i = 0; i < n; ++i) Xc[(size_t)N * n + i] = x[i];
    return J;
  };

  const bool active = lane < na;
  const double alpha = c.alphas[active ? lane : (na - 1)];
  const double J = rollout(alpha, lane == 0);
  const double dJ = d.cost[b] - J;
  const double expected = -alpha * (d.dV[2 * b] + 0.5 * alpha * d.dV[2 * b + 1]);
  const double ratio = expected > 0.0 ? dJ / expected : copysign(1.0, dJ);
  const bool success = active && (ratio > c.opt.armijo_constant);
  const int first = select_alpha<32>(success, J, lane, 0, c.opt.enable_parallel);
  if (active) d.ls_cost[(size_t)b * CDDP_B200_MAX_ALPHAS + lane] = J;
  double Jacc = __shfl_sync(0xffffffffu, J, first >= 0 ? first : 0);
  if (first > 0) {
    __syncwarp();
    (void)rollout(c.alphas[first], lane == 0);
  }
  if (lane == 0) finish_line_search(c, d, b, mode, first, Jacc);
}

}  // namespace

template <int MODEL>
cudaError_t launch_forward_model(const Constants &c, const DeviceState &d, int mode, cudaStream_t st) {
  const int threads = kWarpsPerCta * 32;
  const bool diag = c.cost_diag != 0;
  if (c.num_alphas <= 16) {
    if (mode == FW_ITERATE && c.num_alphas > 8 && !c.opt.enable_parallel && c.ls_window) {
      // windowed line search (kernels_forward.cuh): the first 8 candidates with 8 lanes per trajectory, then the full
      // width for the instances none of them settled
      const int wblocks = (d.n_slots + kWarpsPerCta * 4 - 1) / (kWarpsPerCta * 4);
      if (diag) forward_kernel<MODEL, 8, true, true><<<wblocks, threads, 0, st>>>(c, d, mode);
      else forward_kernel<MODEL, 8, false, true><<<wblocks, threads, 0, st>>>(c, d, mode);
    }
    const int blocks = (d.n_slots + kWarpsPerCta * 2 - 1) / (kWarpsPerCta * 2);
    if (diag) forward_kernel<MODEL, 16, true><<<blocks, threads, 0, st>>>(c, d, mode);
    else forward_kernel<MODEL, 16, false><<<blocks, threads, 0, st>>>(c, d, mode);
  } else {
    const int blocks = (d.n_slots + kWarpsPerCta - 1) / kWarpsPerCta;
    if (diag) forward_kernel<MODEL, 32, true><<<blocks, threads, 0, st>>>(c, d, mode);
    else forward_kernel<MODEL, 32, false><<<blocks, threads, 0, st>>>(c, d, mode);
  }
  return cudaGetLastError();
}

cudaError_t launch_forward(const Constants &c, const DeviceState &d, int mode, cudaStream_t st) {
  switch (c.model) {
    case CDDP_B200_MODEL_PENDULUM: return launch_forward_model<CDDP_B200_MODEL_PENDULUM>(c, d, mode, st);
    case CDDP_B200_MODEL_CARTPOLE: return launch_forward_model<CDDP_B200_MODEL_CARTPOLE>(c, d, mode, st);
    case CDDP_B200_MODEL_UNICYCLE: return launch_forward_model<CDDP_B200_MODEL_UNICYCLE>(c, d, mode, st);
    case CDDP_B200_MODEL_QUADROTOR: return launch_forward_model<CDDP_B200_MODEL_QUADROTOR>(c, d, mode, st);
    case CDDP_B200_MODEL_LTI: {
      const int blocks = (d.n_slots + kWarpsPerCta - 1) / kWarpsPerCta;
      forward_lti_kernel<<<blocks, kWarpsPerCta * 32, 0, st>>>(c, d, mode);
      return cudaGetLastError();
    }
    case CDDP_B200_MODEL_USER: return launch_user_forward(c, d, mode, st);
    default: return cudaErrorInvalidValue;
  }
}

}  // namespace cddp_b200
