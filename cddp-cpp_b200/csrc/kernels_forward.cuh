// forward_kernel: batched CLDDP forward rollout + line search (see forward.cu for the reference citations).
// Header-only so that the SAME kernel is compiled ahead of time for the built-in models (forward.cu) and at run time by
// NVRTC for a user-supplied model (user_model_host.cu).
#pragma once
#include "kernel_common.cuh"

namespace cddp_b200 {
namespace kern {

// Two warps per CTA and at most 146 registers: 7 CTAs = 14 warps = 28 trajectories per SM, so that the headline batch
// (4096 trajectories = 27.7 per SM) is ONE wave.  With 4-warp CTAs at 168 registers only 12 warps fitted and the launch
// ran 1.15 waves (ncu launch__waves_per_multiprocessor), i.e. a second, nearly empty pass of the whole rollout.
constexpr int kWarpsPerCta = 2;
constexpr int kMinCtasPerSm = 7;


// Which alpha is applied, within a lane group of LG lanes (one lane per alpha).  Sequential rule
// (enable_parallel=false, cddp_solver_base.cpp:255-263): the first accepted alpha.  Parallel rule
// (enable_parallel=true, :264-285): the accepted alpha with the strictly lowest cost, scanning in alpha order
// (so ties keep the earlier one; a non-finite cost can never beat the initial +inf).
template <int LG>
__device__ __forceinline__ int select_alpha(bool success, double J, int al, int grp, int enable_parallel) {
  if (!enable_parallel) {
    unsigned ballot = __ballot_sync(0xffffffffu, success);
    if (LG < 32) ballot = (ballot >> (grp * LG)) & ((1u << (LG & 31)) - 1u);
    return ballot ? (__ffs(ballot) - 1) : -1;
  }
  const bool cand = success && (J < pos_inf());
  double Jm = cand ? J : pos_inf();
  int idx = cand ? al : 64;
#pragma unroll
  for (int o = LG / 2; o > 0; o >>= 1) {
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
__device__ inline void finish_line_search(const Constants &c, const DeviceState &d, int b, int mode, int first, double Jacc) {
  d.accepted[b] = first;
  if (mode != FW_ITERATE) return;
  trace_line_search(d, b, first);
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

// Quadratic costs.  DIAG: Q, R, Qf diagonal (every shipped workload) -> O(n) per evaluation; the dense form is
// the literal (e^T Q) e of objective.cpp:80-98.  Both accumulate in the same index order.
template <int NS, bool DIAG>
__device__ __forceinline__ double quad_form(const double *M, const double *e) {
  double s = 0.0;
  if (DIAG) {
#pragma unroll
    for (int j = 0; j < NS; ++j) s += (e[j] * M[j]) * e[j];
  } else {
#pragma unroll
    for (int j = 0; j < NS; ++j) {
      double r = 0.0;
#pragma unroll
      for (int i = 0; i < NS; ++i) r += e[i] * M[i * NS + j];
      s += r * e[j];
    }
  }
  return s;
}

// One lane per alpha, LG = 16 or 32 lanes per trajectory (32/LG trajectories per warp).
//   pass 1: every lane rolls its alpha out (no trajectory writes), saving its state every `seg` steps;
//   select: first accepted alpha (or lowest cost with enable_parallel);
//   pass 2: the LG lanes re-run the accepted alpha's rollout in parallel, one `seg`-step segment each, from
//           the saved states, and write the candidate trajectory (1/LG of a rollout in latency instead of
//           a full sequential replay).
//
// WINDOW (sequential rule only, FW_ITERATE): the kernel looks at the first LG candidates alphas_[0..LG-1] of a longer
// list.  The first accepted alpha wins (cddp_solver_base.cpp:255-263), so if one of them passes the Armijo test the
// remaining candidates are never looked at and the instance is settled here (fw_done = 1); otherwise it is left
// untouched (fw_done = 0) for the full-width launch that follows.  Same decisions, same trajectories; a bang-bang
// workload whose accepted index is 3..6 (the headline quadrotor batch: alphas_[0] is accepted in 0 % of the
// instance-iterations, index <= 7 in all of them) pays 8 rollouts per trajectory instead of 16.
template <int MODEL, int LG, bool DIAG, bool WINDOW = false>
__global__ void __launch_bounds__(kWarpsPerCta * 32, LG == 8 ? 4 : kMinCtasPerSm) forward_kernel(Constants c, DeviceState d, int mode) {
  constexpr int NS = Model<MODEL>::NS, NC = Model<MODEL>::NC;
  constexpr int STEP = NS + 2 * NC + NC * NS;  // x_nom | u_nom | k | K
  constexpr int STEPP = (STEP + 1) & ~1;
  constexpr int TPW = 32 / LG;
  constexpr int PF = (STEP + LG - 1) / LG;
  constexpr int QN = DIAG ? NS : NS * NS, RN = DIAG ? NC : NC * NC;
  __shared__ double sQ[QN], sR[RN], sQf[QN];
  __shared__ double stage[kWarpsPerCta][TPW][2][STEPP];
  __shared__ double sref[kWarpsPerCta][TPW][NS];  // the instance's reference state (read every timestep by every lane)
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < QN; i += blockDim.x) {
    const int src = DIAG ? i * NS + i : i;
    sQ[i] = 0.5 * c.Qdt2[src];  // Q_ = Q*dt
    sQf[i] = 0.5 * c.Qf2[src];
  }
  for (int i = threadIdx.x; i < RN; i += blockDim.x) sR[i] = 0.5 * c.Rdt2[DIAG ? i * NC + i : i];
  __syncthreads();
  const int grp = lane / LG, al = lane % LG;
  const int b = slot_instance(d, (blockIdx.x * kWarpsPerCta + warp) * TPW + grp);
  // (a WINDOW launch is the first of its iteration: it ignores, then rewrites, the flags of the previous one)
  const bool alive = b < d.B && !(mode == FW_ITERATE && (d.status[b] != CDDP_B200_STATUS_RUNNING || (!WINDOW && d.fw_done[b])));
  // the full-width launch consumes the flags: nothing stale is left for a later launch with other options
  if (!WINDOW && mode == FW_ITERATE && b < d.B && al == 0 && d.fw_done[b]) d.fw_done[b] = 0;
  if (!__any_sync(0xffffffffu, alive)) return;
  const int bb = alive ? b : 0;

  const int N = d.N, na = WINDOW ? min(c.num_alphas, LG) : c.num_alphas;
  const int seg = (N + LG - 1) / LG;
  const int cur = d.cur[bb];
  const double *Xn = d.X[cur] + (size_t)bb * (N + 1) * NS;
  const double *Un = d.U[cur] + (size_t)bb * N * NC;
  double *Xc = d.X[cur ^ 1] + (size_t)bb * (N + 1) * NS;
  double *Uc = d.U[cur ^ 1] + (size_t)bb * N * NC;
  const double *gK = d.K + (size_t)bb * N * NC * NS;
  const double *gk = d.kff + (size_t)bb * N * NC;
  for (int i = al; i < NS; i += LG) sref[warp][grp][i] = d.xref[(size_t)bb * NS + i];
  __syncwarp();
  const double *xref = sref[warp][grp];
  const double *rtraj = d.ref_traj ? d.ref_traj + (size_t)bb * (N + 1) * NS : nullptr;
  double *ck = d.ckpt + ((size_t)bb * LG + al) * LG * NS;  // this lane's saved states [LG][NS]
  double(*st)[STEPP] = stage[warp][grp];

  auto load_step = [&](int t, double *pf) {
#pragma unroll
    for (int q = 0; q < PF; ++q) {
      const int i = al + LG * q;
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
      const int i = al + LG * q;
      if (i < STEP) dst[i] = pf[q];
    }
  };
  // one timestep of CLDDPSolver::forwardPass (clddp_solver.cpp:229-246): s = x_nom | u_nom | k | K of step t
  auto advance = [&](const double *s, int t, double alpha, double *x, double *u, double &J) {
#pragma unroll
    for (int i = 0; i < NC; ++i) {  // u' = u + alpha k + K (x' - x)  (:229-233)
      double acc = 0.0;
#pragma unroll
      for (int j = 0; j < NS; ++j) acc += s[NS + 2 * NC + i * NS + j] * (x[j] - s[j]);
      u[i] = s[NS + i] + alpha * s[NS + NC + i] + acc;
    }
    if (c.has_box) {
#pragma unroll
      for (int i = 0; i < NC; ++i) u[i] = clamp_box(u[i], c.lb[i], c.ub[i]);  // clamp (:235-238)
    }
    {  // running cost (e^T Q_) e + (u^T R_) u  (:240-241)
      const double *ref = rtraj ? rtraj + (size_t)t * NS : xref;
      double e[NS];
#pragma unroll
      for (int i = 0; i < NS; ++i) e[i] = x[i] - ref[i];
      J += quad_form<NS, DIAG>(sQ, e) + quad_form<NC, DIAG>(sR, u);
    }
  };
  auto terminal = [&](const double *x) {  // (:247)
    double e[NS];
#pragma unroll
    for (int i = 0; i < NS; ++i) e[i] = x[i] - xref[i];
    return quad_form<NS, DIAG>(sQf, e);
  };

  // ---------------- pass 1: all alphas ----------------
  const bool active = al < na;
  const double alpha = c.alphas[active ? al : (na - 1)];
  double J = 0.0;
  {
    double x[NS], u[NC], xn[NS];
#pragma unroll
    for (int i = 0; i < NS; ++i) x[i] = d.x0[(size_t)bb * NS + i];  // getInitialState(), :224
    double pf[PF];
    load_step(0, pf);
    store_step(st[0], pf);
    __syncwarp();
    int next_ck = 0;
    for (int t = 0; t < N; ++t) {
      if (t + 1 < N) load_step(t + 1, pf);
      if (t == next_ck) {  // save the state entering segment t/seg
        if (alive) {
          double *dst = ck + (size_t)(t / seg) * NS;
#pragma unroll
          for (int i = 0; i < NS; ++i) dst[i] = x[i];
        }
        next_ck += seg;
      }
      advance(st[t & 1], t, alpha, x, u, J);
      discrete_step<MODEL>(c.mp, c.integrator, c.dt, x, u, xn);  // (:243-244)
#pragma unroll
      for (int i = 0; i < NS; ++i) x[i] = xn[i];
      if (t + 1 < N) store_step(st[(t + 1) & 1], pf);
      __syncwarp();
    }
    J += terminal(x);
  }
  const double cost = d.cost[bb];
  const double dJ = cost - J;  // (:249-253)
  const double expected = -alpha * (d.dV[2 * bb] + 0.5 * alpha * d.dV[2 * bb + 1]);
  const double ratio = expected > 0.0 ? dJ / expected : copysign(1.0, dJ);
  const bool success = alive && active && (ratio > c.opt.armijo_constant);
  const int first = select_alpha<LG>(success, J, al, grp, WINDOW ? 0 : c.opt.enable_parallel);
  if (alive && active) d.ls_cost[(size_t)b * CDDP_B200_MAX_ALPHAS + al] = J;
  if (WINDOW) {
    if (alive && al == 0) d.fw_done[b] = first >= 0 ? 1 : 0;
    if (!__any_sync(0xffffffffu, alive && first >= 0)) return;
  }
  const double Jacc = __shfl_sync(0xffffffffu, J, grp * LG + (first >= 0 ? first : 0));

  // ---------------- pass 2: write the accepted rollout, one segment per lane ----------------
  if (alive && first >= 0) {
    const int t0 = al * seg;
    if (t0 < N) {
      const double a1 = c.alphas[first];
      const double *src = d.ckpt + (((size_t)b * LG + first) * LG + al) * NS;
      double x[NS], u[NC], xn[NS], Jd = 0.0;
#pragma unroll
      for (int i = 0; i < NS; ++i) x[i] = src[i];
      const int t1 = min(t0 + seg, N);
      for (int t = t0; t < t1; ++t) {
        double s[STEP];
#pragma unroll
        for (int i = 0; i < NS; ++i) s[i] = Xn[(size_t)t * NS + i];
#pragma unroll
        for (int i = 0; i < NC; ++i) {
          s[NS + i] = Un[(size_t)t * NC + i];
          s[NS + NC + i] = gk[(size_t)t * NC + i];
        }
#pragma unroll
        for (int i = 0; i < NC * NS; ++i) s[NS + 2 * NC + i] = gK[(size_t)t * NC * NS + i];
        advance(s, t, a1, x, u, Jd);
#pragma unroll
        for (int i = 0; i < NS; ++i) Xc[(size_t)t * NS + i] = x[i];
#pragma unroll
        for (int i = 0; i < NC; ++i) Uc[(size_t)t * NC + i] = u[i];
        discrete_step<MODEL>(c.mp, c.integrator, c.dt, x, u, xn);
#pragma unroll
        for (int i = 0; i < NS; ++i) x[i] = xn[i];
      }
      if (t1 == N) {
#pragma unroll
        for (int i = 0; i < NS; ++i) Xc[(size_t)N * NS + i] = x[i];
      }
    }
  }
  if (alive && al == 0 && !(WINDOW && first < 0)) finish_line_search(c, d, b, mode, first, Jacc);
}

// Speculative first step of the sequential line search (cddp_solver_base.cpp:255-263: the FIRST accepted alpha wins, so
// when alphas_[0] is accepted the other candidates are never looked at).  One LANE per trajectory rolls alphas_[0] out,
// writes the candidate trajectory and, if the Armijo test passes, settles the instance (fw_done = 1); forward_kernel
// then runs only for the instances that are left (same decisions, same trajectories: the candidate it would have picked is
// the one written here).  Worth it when the rollout is throughput-bound (a user model with
// transcendental-heavy dynamics at a large batch: 16 lock-step lanes per trajectory cost 16 rollouts, of which a
// well-conditioned problem needs one); launched for CDDP_B200_MODEL_USER handles with enable_parallel = false.
template <int MODEL, bool DIAG>
__global__ void __launch_bounds__(64) forward_first_kernel(Constants c, DeviceState d, int mode) {
  constexpr int NS = Model<MODEL>::NS, NC = Model<MODEL>::NC;
  constexpr int QN = DIAG ? NS : NS * NS, RN = DIAG ? NC : NC * NC;
  __shared__ double sQ[QN], sR[RN], sQf[QN];
  for (int i = threadIdx.x; i < QN; i += blockDim.x) {
    const int src = DIAG ? i * NS + i : i;
    sQ[i] = 0.5 * c.Qdt2[src];
    sQf[i] = 0.5 * c.Qf2[src];
  }
  for (int i = threadIdx.x; i < RN; i += blockDim.x) sR[i] = 0.5 * c.Rdt2[DIAG ? i * NC + i : i];
  __syncthreads();
  const int b = slot_instance(d, blockIdx.x * blockDim.x + threadIdx.x);
  if (b >= d.B) return;
  d.fw_done[b] = 0;
  if (d.status[b] != CDDP_B200_STATUS_RUNNING) return;
  // speculate only where it is likely to pay: the instance's previous line search accepted alphas_[0] (or this is its
  // first iteration).  After a backtracked or failed line search the full kernel runs alone until alphas_[0] wins again.
  if (!(d.iter[b] <= 1 || d.accepted[b] == 0)) return;
  const int N = d.N, cur = d.cur[b];
  const double *Xn = d.X[cur] + (size_t)b * (N + 1) * NS;
  const double *Un = d.U[cur] + (size_t)b * N * NC;
  double *Xc = d.X[cur ^ 1] + (size_t)b * (N + 1) * NS;
  double *Uc = d.U[cur ^ 1] + (size_t)b * N * NC;
  const double *gK = d.K + (size_t)b * N * NC * NS;
  const double *gk = d.kff + (size_t)b * N * NC;
  const double *xref = d.xref + (size_t)b * NS;
  const double *rtraj = d.ref_traj ? d.ref_traj + (size_t)b * (N + 1) * NS : nullptr;
  const double alpha = c.alphas[0];
  double x[NS], u[NC], xn[NS], J = 0.0;
#pragma unroll
  for (int i = 0; i < NS; ++i) x[i] = d.x0[(size_t)b * NS + i];
  for (int t = 0; t < N; ++t) {
    // same arithmetic, in the same order, as forward_kernel's advance()
#pragma unroll
    for (int i = 0; i < NC; ++i) {
      double acc = 0.0;
#pragma unroll
      for (int j = 0; j < NS; ++j) acc += gK[((size_t)t * NC + i) * NS + j] * (x[j] - Xn[(size_t)t * NS + j]);
      u[i] = Un[(size_t)t * NC + i] + alpha * gk[(size_t)t * NC + i] + acc;
    }
    if (c.has_box) {
#pragma unroll
      for (int i = 0; i < NC; ++i) u[i] = clamp_box(u[i], c.lb[i], c.ub[i]);
    }
    {
      const double *ref = rtraj ? rtraj + (size_t)t * NS : xref;
      double e[NS];
#pragma unroll
      for (int i = 0; i < NS; ++i) e[i] = x[i] - ref[i];
      J += quad_form<NS, DIAG>(sQ, e) + quad_form<NC, DIAG>(sR, u);
    }
#pragma unroll
    for (int i = 0; i < NS; ++i) Xc[(size_t)t * NS + i] = x[i];
#pragma unroll
    for (int i = 0; i < NC; ++i) Uc[(size_t)t * NC + i] = u[i];
    discrete_step<MODEL>(c.mp, c.integrator, c.dt, x, u, xn);
#pragma unroll
    for (int i = 0; i < NS; ++i) x[i] = xn[i];
  }
  {
    double e[NS];
#pragma unroll
    for (int i = 0; i < NS; ++i) {
      Xc[(size_t)N * NS + i] = x[i];
      e[i] = x[i] - xref[i];
    }
    J += quad_form<NS, DIAG>(sQf, e);
  }
  const double dJ = d.cost[b] - J;
  const double expected = -alpha * (d.dV[2 * b] + 0.5 * alpha * d.dV[2 * b + 1]);
  const double ratio = expected > 0.0 ? dJ / expected : copysign(1.0, dJ);
  d.ls_cost[(size_t)b * CDDP_B200_MAX_ALPHAS] = J;
  if (ratio > c.opt.armijo_constant) {
    finish_line_search(c, d, b, mode, 0, J);
    d.fw_done[b] = 1;
  }
}

}  // namespace kern
}  // namespace cddp_b200
