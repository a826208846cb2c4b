// Backward Riccati sweep for ONE control (m = 1: pendulum, cartpole — BASELINE configs 1 and 2): the control-space
// subproblem is a scalar (PD test = a sign, BoxQP = a clamp with the reference's iteration around it), so it is solved
// INLINE by every lane of the trajectory's group, redundantly, and the step needs no QP warp, no hand-off through shared
// memory and NO CTA barrier — the two barriers and the hand-off were ~1.5 k of the 2.1 k cycles per step of the
// warp-specialised kernel (backward_fast.cu) at n = 4, where the matrix work is a few dozen FMAs.
// Included by backward_fast.cu inside its anonymous namespace (shares SweepCfg's lane geometry, the mbarrier-free record
// staging below and sweep_epilogue).
//
// Reference behaviour followed: CLDDPSolver::backwardPass, src/cddp_core/clddp_solver.cpp:79-204 (same statements, same
// order of operations as sweep_kernel: the per-step parity tests hold both to the oracle); regularisation retry
// cddp_solver_base.cpp:93-111; BoxQPSolver::solve via SmallQP<1> (boxqp_small.cuh).
#pragma once

template <int NS, class PAT, int W>
struct InlineCfg {
  using L = RecordLayout<NS, 1, PAT>;
  static constexpr int G = group_size(NS);
  static constexpr int TPW = 32 / G;
  static constexpr int T = W * TPW;
  static constexpr int RS = L::stride;
  static constexpr int ev(int x) { return (x + 1) & ~1; }
  static constexpr int oRec = 0;                       // [2][RS] double-buffered record
  static constexpr int oPA = oRec + 2 * RS;            // [NS+1][NS] P_A rows (+ row NS = V_x^T A); reused for the V' transpose
  static constexpr int oPB = oPA + ev((NS + 1) * NS);  // [NS+1] P_B = V B (+ B^T V_x)
  static constexpr int oQux = oPB + ev(NS + 1);        // [NS] Q_ux
  static constexpr int oVx = oQux + ev(NS);            // [NS] new V_x
  static constexpr int raw = oVx + ev(NS);
  static constexpr int ST = raw | 1;  // odd stride: the groups of a warp start in different banks
  static constexpr size_t smemBytes = sizeof(double) * (size_t)(NS * NS + 2 + T * ST);
  static constexpr int threads = W * 32;
};

template <int NS, class PAT, int W>
__global__ void __launch_bounds__(W * 32) sweep_inline_kernel(Constants c, DeviceState d, int mode) {
  using Cfg = InlineCfg<NS, PAT, W>;
  using L = typename Cfg::L;
  constexpr int G = Cfg::G, TPW = Cfg::TPW, T = Cfg::T, RS = Cfg::RS, ST = Cfg::ST, NC = 1;
  constexpr int PF = (RS + G - 1) / G;  // record entries fetched per lane
  extern __shared__ __align__(16) double smem[];
  double *sQ = smem;  // l_xx = 2 Q dt
  double *traj0 = smem + NS * NS + 2;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int N = d.N;
  for (int i = threadIdx.x; i < NS * NS; i += blockDim.x) sQ[i] = c.Qdt2[i];
  const double Rdt2 = c.Rdt2[0];
  const int hw = lane / G, r = lane % G;
  const int q = warp * TPW + hw;
  const int b = slot_instance(d, blockIdx.x * T + q);
  double *S = traj0 + q * ST;
  const volatile double *Sv = S;
  const int row = r, rr = r < NS ? r : NS - 1;
  const bool isrow = r < NS, isvx = r == NS, has = r <= NS;
  const bool alive = b < d.B && !(mode == BW_ITERATE && d.status[b] != CDDP_B200_STATUS_RUNNING);
  const int bb = alive ? b : 0;
  const double *grec = d.rec + (size_t)bb * N * RS;
  double *gK = d.K + (size_t)bb * N * NS;
  double *gk = d.kff + (size_t)bb * N;
  const double *vterm = d.vterm + (size_t)bb * NS;
  __syncthreads();

  double V[NS];        // a row of V_xx, or V_x^T
  double pf[PF];       // this lane's share of the NEXT record, in flight from HBM
  double kprev = 0.0;  // BoxQP warm start k_u_[t] of the step about to be processed (every lane)
  double reg = alive ? d.reg[bb] : 0.0;
  if (alive && mode == BW_ITERATE && r == 0) d.iter[b] += 1;  // ++iter, cddp_solver_base.cpp:75
  double dV0 = 0.0, dV1 = 0.0, Qu_err = 0.0, nrm = 0.0;
  int t = N - 1, buf = 0, status = CDDP_B200_STATUS_RUNNING, failures = 0;
  bool run = alive, ok = false;

  auto fetch = [&](int tt) {
#pragma unroll
    for (int k = 0; k < PF; ++k) {
      const int i = r + k * G;
      pf[k] = (i < RS && tt >= 0) ? grec[(size_t)tt * RS + i] : 0.0;
    }
  };
  auto put = [&](int x) {
#pragma unroll
    for (int k = 0; k < PF; ++k) {
      const int i = r + k * G;
      if (i < RS) S[Cfg::oRec + x * RS + i] = pf[k];
    }
  };
  auto init_sweep = [&]() {  // V_xx = 2 Qf, V_x = 2 Qf (x_N - ref)  (clddp_solver.cpp:89-92)
#pragma unroll
    for (int j = 0; j < NS; ++j) V[j] = isrow ? c.Qf2[rr * NS + j] : vterm[j];
    t = N - 1;
    buf = 0;
    nrm = 0.0;
    dV0 = dV1 = Qu_err = 0.0;
    fetch(N - 1);
    put(0);
    kprev = gk[N - 1];
    fetch(N - 2);
  };
  if (run) init_sweep();
  __syncwarp();
  const double qd = sQ[rr * NS + rr];

  while (__any_sync(0xffffffffu, run)) {
    const volatile double *rc = S + Cfg::oRec + buf * RS;
    // ------------------------------------------------------------ P_B = V B, P_A = V A                        (:124-125)
    double PBr = 0.0, PAr[NS];
#pragma unroll
    for (int j = 0; j < NS; ++j) PAr[j] = 0.0;
    static_for<0, NS>([&](auto lc) {
      constexpr int l = decltype(lc)::value;
      if constexpr (PAT::brow(l)) PBr = fma(V[l], rc[L::idxB(l, 0)], PBr);
      static_for<0, NS>([&](auto jc) {
        constexpr int j = decltype(jc)::value;
        if constexpr (PAT::a(l, j)) PAr[j] = fma(V[l], rc[L::idxA(l, j)], PAr[j]);
      });
    });
    if (has) {
      S[Cfg::oPB + row] = PBr;
#pragma unroll
      for (int j = 0; j < NS; ++j) S[Cfg::oPA + row * NS + j] = PAr[j];
    }
    {
      double l1 = 0.0;  // ||V_x||_1 of the value function entering this step (:194)
#pragma unroll
      for (int j = 0; j < NS; ++j) l1 += fabs(V[j]);
      nrm += isvx ? l1 : 0.0;
    }
    __syncwarp();
    // ------------------------------------------------------------ Q_uu, Q_u, Q_xx, Q_xu, Q_x                  (:126-128)
    double Quu = Rdt2;
    static_for<0, NS>([&](auto lc) {
      constexpr int l = decltype(lc)::value;
      if constexpr (PAT::brow(l)) Quu = fma(rc[L::idxB(l, 0)], Sv[Cfg::oPB + l], Quu);
    });
    const double Qu = rc[L::offLu] + Sv[Cfg::oPB + NS];
    const double un = rc[L::offU];
    double col[NS], Qxx[NS], Qxu = 0.0;
#pragma unroll
    for (int l = 0; l < NS; ++l) col[l] = S[Cfg::oPA + l * NS + rr];
    if (c.q_diag) {
#pragma unroll
      for (int j = 0; j < NS; ++j) Qxx[j] = (j == rr) ? qd : 0.0;
    } else {
#pragma unroll
      for (int j = 0; j < NS; ++j) Qxx[j] = sQ[rr * NS + j];
    }
    static_for<0, NS>([&](auto lc) {
      constexpr int l = decltype(lc)::value;
      static_for<0, NS>([&](auto jc) {
        constexpr int j = decltype(jc)::value;
        if constexpr (PAT::a(l, j)) Qxx[j] = fma(col[l], rc[L::idxA(l, j)], Qxx[j]);
      });
      if constexpr (PAT::brow(l)) Qxu = fma(col[l], rc[L::idxB(l, 0)], Qxu);
    });
    const double Qx = rc[L::offLx + rr] + S[Cfg::oPA + NS * NS + rr];  // Q_x = l_x + A^T V_x
    if (isrow) S[Cfg::oQux + row] = Qxu;
    // ------------------------------------------------------------ control-space subproblem, every lane   (:130-178)
    // (after the loads above and with no barrier in between, so that its divide chain overlaps the Q_xx FMAs)
    double H[1], g[1], kk[1], Hk[1], Hinv[1];
    H[0] = Quu + reg;  // Q_uu_reg (:130-131)
    g[0] = Qu;
    kk[0] = 0.0;
    Hk[0] = 0.0;
    Hinv[0] = 0.0;
    unsigned free_mask = 1u;
    bool good;
    if (c.has_box) {  // (:147-159)
      good = SymPD<1>::run(H);  // PD test (:133-140)
      if (good) {
        double lo[1], hi[1];
        lo[0] = c.lb[0] - un;
        hi[0] = c.ub[0] - un;
        kk[0] = kprev;
        const int qs = SmallQP<1>::solve(c.opt, H, g, lo, hi, kk, free_mask, Hinv, Hk);
        good = !(qs == QP_HESSIAN_NOT_PD || qs == QP_NO_DESCENT);
        if (good && c.opt.qp_max_iterations <= 0) SmallQP<1>::masked_inverse_raw(H, free_mask, Hinv);
        SmallQP<1>::zero_clamped(free_mask, Hinv);  // ALL_CLAMPED: free_mask == 0 -> K = 0 (:163)
      }
    } else {  // k = -H^-1 Q_u (:142-144)
      good = SymInverse<1>::run(H, Hinv);
      kk[0] = -fma(Hinv[0], g[0], 0.0);
      Hk[0] = fma(H[0], kk[0], 0.0);
    }
    good = good && run;
    const double sq = fma(-reg, kk[0], Hk[0]);  // (Q_uu k) = (Q_uu_reg k) - reg k
    const double wq = sq + g[0];                // w = Q_uu k + Q_u
    __syncwarp();  // Q_ux complete; every lane is done with this step's record
    // next step's record: the fetched one goes to the other buffer, the one after is requested
    if (run && t > 0) put(buf ^ 1);
    const double kprev_next = (run && t > 0) ? gk[t - 1] : 0.0;
    if (run) fetch(t - 2);
    // ------------------------------------------------------------ value update                            (:184-192)
    //   V_xx' = Q_xx + Q_ux^T T,  T = 2 K - Ht Q_uu K,  K = -Ht Q_ux;  V_x' = Q_x + Q_ux^T (k - Ht w)
    const double Quu_ = Quu;  // unregularised
    const double Ht = Hinv[0];
    const double z = kk[0] - fma(Ht, wq, 0.0);
    const double Kc = -fma(Ht, Qxu, 0.0);
    const double QK = fma(Quu_, Kc, 0.0);
    const double Tt = fma(-Ht, QK, 2.0 * Kc);
    double Vn[NS];
    const double vxn = fma(Qxu, z, Qx);
#pragma unroll
    for (int j = 0; j < NS; ++j) Vn[j] = fma(Tt, Sv[Cfg::oQux + j], Qxx[j]);
    if (good && isrow) {
#pragma unroll
      for (int j = 0; j < NS; ++j) S[Cfg::oPA + row * NS + j] = Vn[j];
      S[Cfg::oVx + row] = vxn;
    }
    __syncwarp();
#pragma unroll
    for (int j = 0; j < NS; ++j)  // V_xx = (V' + V'^T)/2 (:192); the V_x slot takes the new V_x^T
      V[j] = isrow ? 0.5 * (Vn[j] + S[Cfg::oPA + j * NS + rr]) : Sv[Cfg::oVx + j];
    if (good) {
      if (isrow) gK[(size_t)t * NS + row] = Kc;  // K_u_[t] (:182)
      if (r == 0) gk[t] = kk[0];                 // k_u_[t] (:181)
      dV0 += fma(g[0], kk[0], 0.0);              // dV += (Q_u.k, 0.5 k^T Q_uu k) (:184-186)
      dV1 += 0.5 * fma(kk[0], sq, 0.0);
      Qu_err = max_ref(Qu_err, max_ref(0.0, fabs(g[0])));  // (:195)
      kprev = kprev_next;
      --t;
      buf ^= 1;
      if (t < 0) {  // sweep finished: white-box value function at t = 0, and the V_x(0) term of the norm
        run = false;
        ok = true;
        double l1 = 0.0;
#pragma unroll
        for (int j = 0; j < NS; ++j) {
          l1 += fabs(V[j]);
          if (isrow) d.Vxx0[((size_t)b * NS + row) * NS + j] = V[j];
          if (isvx) d.Vx0[(size_t)b * NS + j] = V[j];
        }
        if (isvx) S[Cfg::oVx] = nrm + l1;  // handed to lane 0 for the epilogue
      }
    } else if (run) {
      if (mode == BW_SINGLE) {
        run = false;
      } else {  // increaseRegularization + limit test (cddp_solver_base.cpp:95-109, cddp_core.cpp:308-326)
        reg = fmin(reg * c.opt.reg_update_factor, c.opt.reg_max_value);
        ++failures;
        if (reg >= c.opt.reg_max_value) {
          status = CDDP_B200_STATUS_REG_LIMIT;
          run = false;
        } else {
          init_sweep();  // sweep again at the increased regularisation
        }
      }
    }
    __syncwarp();  // all reads of the transpose buffer done before the next step overwrites it
  }
  __syncwarp();
  if (alive && r == 0) sweep_epilogue<NS>(c, d, mode, b, ok, Sv[Cfg::oVx], N, reg, dV0, dV1, Qu_err, status, failures);
}

template <int NS, class PAT, int W>
cudaError_t launch_sweep_inline(const Constants &c, const DeviceState &d, int mode, cudaStream_t st) {
  using Cfg = InlineCfg<NS, PAT, W>;
  static_assert(NS + 1 <= Cfg::G, "one row per lane, V_x rides as an extra row");
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(sweep_inline_kernel<NS, PAT, W>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Cfg::smemBytes);
    if (e != cudaSuccess) return e;
    configured = true;
  }
  const int blocks = (d.n_slots + Cfg::T - 1) / Cfg::T;
  sweep_inline_kernel<NS, PAT, W><<<blocks, Cfg::threads, Cfg::smemBytes, st>>>(c, d, mode);
  return cudaGetLastError();
}
