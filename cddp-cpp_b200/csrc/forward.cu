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

namespace cddp_b200 {

namespace {

constexpr int kWarpsPerCta = 4;

__device__ __forceinline__ double pos_inf() { return __longlong_as_double(0x7ff0000000000000LL); }


// Which alpha is applied.  Sequential rule (enable_parallel=false, cddp_solver_base.cpp:255-263): the
// first accepted alpha.  Parallel rule (enable_parallel=true, :264-285): the accepted alpha with the
// strictly lowest cost, scanning in alpha order (so ties keep the earlier one; a non-finite cost can
// never beat the initial +inf).
__device__ __forceinline__ int select_alpha(bool success, double J, int lane, int enable_parallel) {
  if (!enable_parallel) {
    const unsigned ballot = __ballot_sync(0xffffffffu, success);
    return ballot ? (__ffs(ballot) - 1) : -1;
  }
  const bool cand = success && (J < pos_inf());
  double Jm = cand ? J : pos_inf();
  int idx = cand ? lane : 64;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const double Jo = __shfl_xor_sync(0xffffffffu, Jm, o);
    const int io = __shfl_xor_sync(0xffffffffu, idx, o);
    if (Jo < Jm || (Jo == Jm && io < idx)) {
      Jm = Jo;
      idx = io;
    }
  }
  return idx < 64 ? idx : -1;
}

// per-instance bookkeeping after the line search (lane 0 only)
__device__ void finish_line_search(const Constants &c, const DeviceState &d, int b, int mode, int first, double Jacc) {
  d.accepted[b] = first;
  if (mode != FW_ITERATE) return;
  double reg = d.reg[b];
  int status = CDDP_B200_STATUS_RUNNING;
  if (first >= 0) {
    const double dJ = d.cost[b] - Jacc;  // cddp_solver_base.cpp:129
    d.cost[b] = Jacc;                    // applyForwardPassResult :190-198
    d.alpha[b] = c.alphas[first];
    d.cur[b] ^= 1;
    d.lin_valid[b] = 0;
    if (d.history) {  // recordIterationHistory BEFORE decreaseRegularization (:132-135)
      const int hl = d.history_len[b];
      if (hl < d.history_cap) {
        double *h = d.history + ((size_t)b * d.history_cap + hl) * 4;
        h[0] = Jacc;
        h[1] = c.alphas[first];
        h[2] = d.inf_du[b];
        h[3] = reg;
        d.history_len[b] = hl + 1;
      }
    }
    reg = fmax(reg / c.opt.reg_update_factor, c.opt.reg_min_value);  // cddp_core.cpp:316-322
    if (d.inf_du[b] < c.opt.tolerance)                               // clddp_solver.cpp:268-271
      status = CDDP_B200_STATUS_OPTIMAL;
    else if (dJ > 0.0 && dJ < c.opt.acceptable_tolerance)            // :272-275
      status = CDDP_B200_STATUS_ACCEPTABLE;
  } else {  // handleForwardPassFailure, cddp_solver_base.cpp:206-218
    reg = fmin(reg * c.opt.reg_update_factor, c.opt.reg_max_value);
    if (reg >= c.opt.reg_max_value) status = CDDP_B200_STATUS_REG_LIMIT;
  }
  d.reg[b] = reg;
  if (status != CDDP_B200_STATUS_RUNNING) d.status[b] = status;
}

template <int MODEL>
__global__ void __launch_bounds__(kWarpsPerCta * 32) forward_kernel(Constants c, DeviceState d, int mode) {
  constexpr int NS = Model<MODEL>::NS, NC = Model<MODEL>::NC;
  constexpr int STEP = NS + 2 * NC + NC * NS;  // x_nom | u_nom | k | K
  constexpr int PF = (STEP + 31) / 32;
  __shared__ double sQ[NS * NS], sR[NC * NC], sQf[NS * NS];
  __shared__ double stage[kWarpsPerCta][2][STEP + 1];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < NS * NS; i += blockDim.x) {
    sQ[i] = 0.5 * c.Qdt2[i];   // Q_ = Q*dt
    sQf[i] = 0.5 * c.Qf2[i];
  }
  for (int i = threadIdx.x; i < NC * NC; i += blockDim.x) sR[i] = 0.5 * c.Rdt2[i];
  __syncthreads();
  const int b = blockIdx.x * kWarpsPerCta + warp;
  if (b >= d.B) return;
  if (mode == FW_ITERATE && d.status[b] != CDDP_B200_STATUS_RUNNING) return;

  const int N = d.N, na = c.num_alphas;
  const int cur = d.cur[b];
  const double *Xn = d.X[cur] + (size_t)b * (N + 1) * NS;
  const double *Un = d.U[cur] + (size_t)b * N * NC;
  double *Xc = d.X[cur ^ 1] + (size_t)b * (N + 1) * NS;
  double *Uc = d.U[cur ^ 1] + (size_t)b * N * NC;
  const double *gK = d.K + (size_t)b * N * NC * NS;
  const double *gk = d.kff + (size_t)b * N * NC;
  const double *xref = d.xref + (size_t)b * NS;
  const double *rtraj = d.ref_traj ? d.ref_traj + (size_t)b * (N + 1) * NS : nullptr;

  auto load_step = [&](int t, double *pf) {
#pragma unroll
    for (int q = 0; q < PF; ++q) {
      const int i = lane + 32 * q;
      double v = 0.0;
      if (i < NS) v = Xn[(size_t)t * NS + i];
      else if (i < NS + NC) v = Un[(size_t)t * NC + (i - NS)];
      else if (i < NS + 2 * NC) v = gk[(size_t)t * NC + (i - NS - NC)];
      else if (i < STEP) v = gK[(size_t)t * NC * NS + (i - NS - 2 * NC)];
      pf[q] = v;
    }
  };
  auto store_step = [&](double *dst, const double *pf) {
#pragma unroll
    for (int q = 0; q < PF; ++q) {
      const int i = lane + 32 * q;
      if (i < STEP) dst[i] = pf[q];
    }
  };

  // one rollout of every lane's alpha; returns the lane's total cost
  auto rollout = [&](double alpha, bool writer) -> double {
    double x[NS];
#pragma unroll
    for (int i = 0; i < NS; ++i) x[i] = d.x0[(size_t)b * NS + i];  // getInitialState(), :224
    double J = 0.0;
    double pf[PF];
    load_step(0, pf);
    store_step(stage[warp][0], pf);
    __syncwarp();
    for (int t = 0; t < N; ++t) {
      const double *s = stage[warp][t & 1];
      if (t + 1 < N) load_step(t + 1, pf);
      double u[NC];
#pragma unroll
      for (int i = 0; i < NC; ++i) {  // u' = u + alpha k + K (x' - x)  (:229-233)
        double acc = 0.0;
#pragma unroll
        for (int j = 0; j < NS; ++j) acc += s[NS + 2 * NC + i * NS + j] * (x[j] - s[j]);
        u[i] = s[NS + i] + alpha * s[NS + NC + i] + acc;
      }
      if (c.has_box) {
#pragma unroll
        for (int i = 0; i < NC; ++i) u[i] = fmin(fmax(u[i], c.lb[i]), c.ub[i]);  // clamp (:235-238)
      }
      {  // running cost (e^T Q_) e + (u^T R_) u  (:240-241)
        const double *ref = rtraj ? rtraj + (size_t)t * NS : xref;
        double e[NS];
#pragma unroll
        for (int i = 0; i < NS; ++i) e[i] = x[i] - ref[i];
        double sx = 0.0, su = 0.0;
#pragma unroll
        for (int j = 0; j < NS; ++j) {
          double r = 0.0;
#pragma unroll
          for (int i = 0; i < NS; ++i) r += e[i] * sQ[i * NS + j];
          sx += r * e[j];
        }
#pragma unroll
        for (int j = 0; j < NC; ++j) {
          double r = 0.0;
#pragma unroll
          for (int i = 0; i < NC; ++i) r += u[i] * sR[i * NC + j];
          su += r * u[j];
        }
        J += sx + su;
      }
      if (writer) {
#pragma unroll
        for (int i = 0; i < NS; ++i) Xc[(size_t)t * NS + i] = x[i];
#pragma unroll
        for (int i = 0; i < NC; ++i) Uc[(size_t)t * NC + i] = u[i];
      }
      double xn[NS];
      discrete_step<MODEL>(c.mp, c.integrator, c.dt, x, u, xn);  // (:243-244)
#pragma unroll
      for (int i = 0; i < NS; ++i) x[i] = xn[i];
      if (t + 1 < N) store_step(stage[warp][(t + 1) & 1], pf);
      __syncwarp();
    }
    {  // terminal cost (:247)
      double e[NS];
#pragma unroll
      for (int i = 0; i < NS; ++i) e[i] = x[i] - xref[i];
      double sx = 0.0;
#pragma unroll
      for (int j = 0; j < NS; ++j) {
        double r = 0.0;
#pragma unroll
        for (int i = 0; i < NS; ++i) r += e[i] * sQf[i * NS + j];
        sx += r * e[j];
      }
      J += sx;
    }
    if (writer) {
#pragma unroll
      for (int i = 0; i < NS; ++i) Xc[(size_t)N * NS + i] = x[i];
    }
    return J;
  };

  const bool active = lane < na;
  const double alpha = c.alphas[active ? lane : (na - 1)];
  const double J = rollout(alpha, lane == 0);
  const double cost = d.cost[b];
  const double dJ = cost - J;  // (:249-253)
  const double expected = -alpha * (d.dV[2 * b] + 0.5 * alpha * d.dV[2 * b + 1]);
  const double ratio = expected > 0.0 ? dJ / expected : copysign(1.0, dJ);
  const bool success = active && (ratio > c.opt.armijo_constant);
  const int first = select_alpha(success, J, lane, c.opt.enable_parallel);
  if (active) d.ls_cost[(size_t)b * CDDP_B200_MAX_ALPHAS + lane] = J;
  double Jacc = __shfl_sync(0xffffffffu, J, first >= 0 ? first : 0);
  if (first > 0) {
    __syncwarp();
    (void)rollout(c.alphas[first], lane == 0);  // replay the accepted alpha; identical arithmetic => identical J
  }
  if (lane == 0) finish_line_search(c, d, b, mode, first, Jacc);
}

// LTI: runtime dimensions, discrete dynamics x+ = A_d x + B_d u (lti_system.cpp:71-76)
__global__ void __launch_bounds__(kWarpsPerCta * 32) forward_lti_kernel(Constants c, DeviceState d, int mode) {
  constexpr int MS = CDDP_B200_MAX_N, MC = CDDP_B200_MAX_M;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int b = blockIdx.x * kWarpsPerCta + warp;
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
        for (int i = 0; i < m; ++i) u[i] = fmin(fmax(u[i], c.lb[i]), c.ub[i]);
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
  const int first = select_alpha(success, J, lane, c.opt.enable_parallel);
  if (active) d.ls_cost[(size_t)b * CDDP_B200_MAX_ALPHAS + lane] = J;
  double Jacc = __shfl_sync(0xffffffffu, J, first >= 0 ? first : 0);
  if (first > 0) {
    __syncwarp();
    (void)rollout(c.alphas[first], lane == 0);
  }
  if (lane == 0) finish_line_search(c, d, b, mode, first, Jacc);
}

}  // namespace

cudaError_t launch_forward(const Constants &c, const DeviceState &d, int mode, cudaStream_t st) {
  const int blocks = (d.B + kWarpsPerCta - 1) / kWarpsPerCta;
  const int threads = kWarpsPerCta * 32;
  switch (c.model) {
    case CDDP_B200_MODEL_PENDULUM: forward_kernel<CDDP_B200_MODEL_PENDULUM><<<blocks, threads, 0, st>>>(c, d, mode); break;
    case CDDP_B200_MODEL_CARTPOLE: forward_kernel<CDDP_B200_MODEL_CARTPOLE><<<blocks, threads, 0, st>>>(c, d, mode); break;
    case CDDP_B200_MODEL_UNICYCLE: forward_kernel<CDDP_B200_MODEL_UNICYCLE><<<blocks, threads, 0, st>>>(c, d, mode); break;
    case CDDP_B200_MODEL_QUADROTOR: forward_kernel<CDDP_B200_MODEL_QUADROTOR><<<blocks, threads, 0, st>>>(c, d, mode); break;
    case CDDP_B200_MODEL_LTI: forward_lti_kernel<<<blocks, threads, 0, st>>>(c, d, mode); break;
    default: return cudaErrorInvalidValue;
  }
  return cudaGetLastError();
}

}  // namespace cddp_b200
