// IPDDP backward pass with a TerminalEqualityConstraint on the reference state (h(x_N) = x_N - x_ref, H_T = I):
// IPDDPSolver::backwardPass, terminal-equality branch (src/cddp_core/ipddp_solver.cpp:1120-1353) with
// solveTerminalEqualityLQR (:484-639) and solveSequentialLQR (:411-482).
//
// The reference runs p+1 = n+1 complete sequential-LQR sweeps that differ only in the terminal costate q_N (+ H_T row v),
// then p+1 linear rollouts, a small regularised least-squares problem for the multiplier step, and a final rollout.  The
// value Hessians P_t and the feedback gains K_t do not depend on q, so here ONE sweep carries P, K and the p+1 costate /
// feed-forward columns (p_v, k_v) side by side, followed by one forward pass that rolls all p+1 variants
// out together, the small solve on one lane, a time-parallel combination pass, and the final rollout that also forms the
// slack / dual gains and the fraction-to-boundary step caps.  dV_ stays zero on this branch, as in the reference.
#include "engine.h"
#include "kernels_ipddp.cuh"
#include "ldlt_small.cuh"

#include <cstdlib>
#include <string>

namespace cddp_b200 {

namespace {

using namespace kern;

constexpr int kTeqThreads = 128;

#include "ipddp_teq_small.cuh"

// CDDP_B200_TEQ_KERNEL=shared keeps the shared-memory kernel where the register-resident one applies (A-B timing and the
// both-kernels test; read per launch)
static bool teq_reg_enabled() {
  const char *e = std::getenv("CDDP_B200_TEQ_KERNEL");
  return !(e && std::string(e) == "shared");
}

__host__ __device__ inline int teq_stage_doubles(int n, int m, int D, int rs) {
  const int sweep = rs + n + 3 * D;                                 // record | x | y | s | g
  const int roll = rs + m * n + (n + 1) * m;                        // record | K | k_v (variants rollout)
  const int fin = rs + n + 3 * D + m * n + m;                       // record | x | y | s | g | K | k (final rollout)
  int b = sweep > roll ? sweep : roll;
  b = b > fin ? b : fin;
  return (b + 1) & ~1;
}
__host__ __device__ inline int teq_group_doubles(int n, int m, int D, int rs) {
  const int nv = n + 1;
  int c = 2 * teq_stage_doubles(n, m, D, rs);
  c += 3 * n * n;            // P, PA, Qt
  c += 2 * m * n;            // BtP, Qux
  c += 3 * m * m;            // Quu, Qf (factor), Rt
  c += n * m;                // Mt
  c += n + m;                // qt, rt
  c += m * (2 * n + 1);      // RHS: [Qux | Qu_v] -> [K | k_v]
  c += nv * n * 2;           // pv, Qxv
  c += nv * m;               // Quv (copy of the right-hand sides before the solve)
  c += m * n + nv * m;       // QuuK, Quuk_v
  c += D * n + D * m + 6 * D;  // Gx, Gu, ssafe, YS, prim, rhat, w, (spare)
  c += nv * n * 2;           // dx_v, dxn_v
  c += 4 * n * n + 4 * n;    // small solve: As, AtA, shifted, scratch vectors
  c += 2 * CDDP_B200_MAX_N;  // transpositions + flags (ints)
  return (c + 1) & ~1;
}

template <int G, int NS, int NC, int DC = 0>
__global__ void __launch_bounds__(kTeqThreads) ip_backward_teq_kernel(Constants c, DeviceState d, IpConstants ic, IpDevice ip, int mode) {
  extern __shared__ double smem[];
  const int n = NS ? NS : d.n, m = NC ? NC : d.m, N = d.N, rs = d.rec_stride, D = DC ? DC : ic.d;
  const int nv = n + 1, nc2 = 2 * n + 1;
  constexpr int GPC = kTeqThreads / G;
  double *sQ = smem;
  double *sR = sQ + n * n;
  double *tGx = sR + m * m + ((n * n + m * m) & 1);
  double *tGu = tGx + D * n;
  double *tOff = tGu + D * m;
  double *tScale = tOff + D;
  int *tType = reinterpret_cast<int *>(tScale + D);
  int *tBdim = tType + D;
  for (int i = threadIdx.x; i < n * n; i += blockDim.x) sQ[i] = c.Qdt2[i];
  for (int i = threadIdx.x; i < m * m; i += blockDim.x) sR[i] = c.Rdt2[i];
  for (int i = threadIdx.x; i < D * n; i += blockDim.x) tGx[i] = ic.Gx[i];
  for (int i = threadIdx.x; i < D * m; i += blockDim.x) tGu[i] = ic.Gu[i];
  for (int i = threadIdx.x; i < D; i += blockDim.x) {
    tOff[i] = ic.off[i];
    tScale[i] = ic.scale[i];
    tType[i] = ic.row_type[i];
    tBdim[i] = ic.row_bdim[i];
  }
  __syncthreads();
  const int grp = threadIdx.x / G, r = threadIdx.x % G;
  const int b = slot_instance(d, blockIdx.x * GPC + grp);
  const bool alive = b < d.B && !(mode == BW_ITERATE && d.status[b] != CDDP_B200_STATUS_RUNNING);
  const int bb = alive ? b : 0;
  const int blk = teq_stage_doubles(n, m, D, rs);
  const int tab = (D * n + D * m + 3 * D + 1) & ~1;
  double *w = tGx + tab + (size_t)grp * teq_group_doubles(n, m, D, rs);
  double *stg = w;   w += 2 * blk;
  double *P = w;     w += n * n;
  double *PA = w;    w += n * n;
  double *Qt = w;    w += n * n;
  double *BtP = w;   w += m * n;
  double *Qux = w;   w += m * n;
  double *Quu = w;   w += m * m;
  double *Qf = w;    w += m * m;
  double *Rt = w;    w += m * m;
  double *Mt = w;    w += n * m;
  double *qt = w;    w += n;
  double *rt = w;    w += m;
  double *RHS = w;   w += m * nc2;
  double *pv = w;    w += nv * n;
  double *Qxv = w;   w += nv * n;
  double *Quv = w;   w += nv * m;
  double *QuuK = w;  w += m * n;
  double *Quuk = w;  w += nv * m;
  double *Gx = w;    w += D * n;
  double *Gu = w;    w += D * m;
  double *ssafe = w; w += D;
  double *YS = w;    w += D;
  double *prim = w;  w += D;
  double *rhat = w;  w += D;
  double *wv = w;    w += 2 * D;
  double *dxv = w;   w += nv * n;
  double *dxn = w;   w += nv * n;
  double *As = w;    w += n * n;
  double *AtA = w;   w += n * n;
  double *Sh = w;    w += n * n;
  double *vec = w;   w += n * n + 4 * n;  // rhs | Atb | lam | best | eigen scratch
  int *tr = reinterpret_cast<int *>(w);

  const int cur = d.cur[bb];
  const double *grec = d.rec + (size_t)bb * N * rs;
  const double *gX = d.X[cur] + (size_t)bb * (N + 1) * n;
  const double *gY = ip.Y[cur] + (size_t)bb * N * D, *gS = ip.S[cur] + (size_t)bb * N * D, *gG = ip.G[cur] + (size_t)bb * N * D;
  double *gK = d.K + (size_t)bb * N * m * n, *gk = d.kff + (size_t)bb * N * m;
  double *gky = ip.ky + (size_t)bb * N * D, *gks = ip.ks + (size_t)bb * N * D;
  double *gKy = ip.Ky + (size_t)bb * N * D * n, *gKs = ip.Ks + (size_t)bb * N * D * n;
  double *kvar = ip.kvar + (size_t)bb * nv * N * m;        // [v][t][m]
  double *pvar = ip.pvar + (size_t)bb * nv * (N + 1) * n;  // [v][t][n]
  double *rvar = ip.rvar + (size_t)bb * N * m;             // [t][m]
  const double *lamT = ip.lamT + (size_t)bb * n;
  const double *xref = d.xref + (size_t)bb * n;

  const double mu = alive ? ip.mu[bb] : 1.0;
  double reg = alive ? d.reg[bb] : 0.0;
  if (alive && mode == BW_ITERATE && r == 0) d.iter[b] += 1;
  int status = CDDP_B200_STATUS_RUNNING;
  bool need = alive, ok = false;
  int failures = 0;  // backward-pass failures of this iteration (decision trace)
  double inf_du = 0.0, inf_pr = 0.0, inf_comp = 0.0, step_norm = 0.0, apm = 1.0, adm = 1.0;

  auto issue_sweep = [&](int tt, int x) {
    double *dst = stg + x * blk;
    for (int i = r; i < rs; i += G) cp_async8(dst + i, grec + (size_t)tt * rs + i);
    for (int i = r; i < n; i += G) cp_async8(dst + rs + i, gX + (size_t)tt * n + i);
    for (int i = r; i < D; i += G) {
      cp_async8(dst + rs + n + i, gY + (size_t)tt * D + i);
      cp_async8(dst + rs + n + D + i, gS + (size_t)tt * D + i);
      cp_async8(dst + rs + n + 2 * D + i, gG + (size_t)tt * D + i);
    }
  };
  // constraint Jacobians and barrier terms of one timestep from its staged x, y, s, g (:1181-1220)
  auto barrier_terms = [&](const double *xs, const double *ys, const double *ss, const double *gs) {
    for (int i = r; i < D * n; i += G) {
      const int row = i / n, j = i - row * n;
      const int ty = tType[row];
      double v = 0.0;
      if (ty == IP_ROW_STATE) v = tGx[i];
      else if (ty == IP_ROW_BALL && j < tBdim[row]) v = -2.0 * tScale[row] * (xs[j] - tGx[i]);
      Gx[i] = v;
    }
    for (int i = r; i < D * m; i += G) Gu[i] = (tType[i / m] == IP_ROW_CONTROL) ? tGu[i] : 0.0;
    for (int i = r; i < D; i += G) {
      const double sf = fmax(ss[i], fmax(mu * 1e-3, EPS_SLACK));
      ssafe[i] = sf;
      YS[i] = clip_pos(ys[i], sf);
      const double pr = gs[i] + ss[i];
      prim[i] = pr;
      const double rh = ys[i] * pr - (ys[i] * ss[i] - mu);
      rhat[i] = rh;
      wv[i] = ys[i] + clip_signed(rh, sf);
    }
  };

  while (__any_sync(0xffffffffu, need)) {
    bool act = need;
    // ---------------------------------------------------------------- sweep: P, K shared; p_v, k_v per variant
    if (act) {
      for (int i = r; i < n * n; i += G) {  // P_N = sym(sym(2 Qf)) (:990, :438)
        const int a = i / n, e = i - a * n;
        P[i] = 0.5 * (c.Qf2[a * n + e] + c.Qf2[e * n + a]);
      }
      for (int i = r; i < nv * n; i += G) {  // p_v[N] = V_x + lambda_prev (+ e_{v-1})   (:520-526, :548-553)
        const int v = i / n, j = i - v * n;
        const double val = (d.vterm[(size_t)bb * n + j] + lamT[j]) + ((v > 0 && v - 1 == j) ? 1.0 : 0.0);
        pv[i] = val;
        pvar[((size_t)v * (N + 1) + N) * n + j] = val;
      }
      issue_sweep(N - 1, 0);
    }
    inf_pr = inf_comp = 0.0;
    if (act)
      for (int j = 0; j < n; ++j) inf_pr = fmax(inf_pr, fabs(gX[(size_t)N * n + j] - xref[j]));  // |h_T| (:1041)
    cp_async_wait_all();
    __syncwarp();
    for (int t = N - 1; t >= 0; --t) {
      const int bi = (N - 1 - t) & 1;
      if (act && t > 0) issue_sweep(t - 1, bi ^ 1);
      const double *rec = stg + bi * blk, *xs = rec + rs, *ys = xs + n, *ss = ys + D, *gs = ss + D;
      const double *A = rec, *Bm = rec + n * n, *lx = rec + d.offLx, *lu = rec + d.offLu;
      if (r == 0) tr[CDDP_B200_MAX_N + 1] = 0;  // non-finite flag of this step (read several barriers later)
      if (act) {
        barrier_terms(xs, ys, ss, gs);
        for (int idx = r; idx < m * n; idx += G) {  // BtP = B^T P
          const int i = idx / n, j = idx - i * n;
          double s = 0.0;
          for (int l = 0; l < n; ++l) s += Bm[l * m + i] * P[l * n + j];
          BtP[idx] = s;
        }
        for (int idx = r; idx < n * n; idx += G) {  // PA = P A
          const int i = idx / n, j = idx - i * n;
          double s = 0.0;
          for (int l = 0; l < n; ++l) s += P[i * n + l] * A[l * n + j];
          PA[idx] = s;
        }
      }
      __syncwarp();
      if (act) {
        for (int q = 0; q < D; ++q) {
          inf_pr = fmax(inf_pr, fabs(prim[q]));
          inf_comp = fmax(inf_comp, fabs(ys[q] * ss[q] - mu));
        }
        // condensed stage cost (:1143-1254): Q_t, q_t, R_t (+ reg), r_t, M_t
        for (int idx = r; idx < n * n + m * m + n * m + n + m; idx += G) {
          int e = idx;
          if (e < n * n) {
            const int i = e / n, j = e - i * n;
            double a1 = 0.0, a2 = 0.0;
            for (int q = 0; q < D; ++q) {
              a1 += Gx[q * n + i] * (YS[q] * Gx[q * n + j]);
              a2 += Gx[q * n + j] * (YS[q] * Gx[q * n + i]);
            }
            const double base = 0.5 * (sQ[i * n + j] + sQ[j * n + i]);
            Qt[e] = D ? 0.5 * ((base + a1) + (base + a2)) : base;
            continue;
          }
          e -= n * n;
          if (e < m * m) {
            const int i = e / m, j = e - i * m;
            double a1 = 0.0, a2 = 0.0;
            for (int q = 0; q < D; ++q) {
              a1 += Gu[q * m + i] * (YS[q] * Gu[q * m + j]);
              a2 += Gu[q * m + j] * (YS[q] * Gu[q * m + i]);
            }
            const double base = 0.5 * (sR[i * m + j] + sR[j * m + i]);
            Rt[e] = (D ? 0.5 * ((base + a1) + (base + a2)) : base) + (i == j ? reg : 0.0);
            continue;
          }
          e -= m * m;
          if (e < n * m) {
            const int i = e / m, j = e - i * m;
            double a = 0.0;
            for (int q = 0; q < D; ++q) a += Gu[q * m + j] * (YS[q] * Gx[q * n + i]);
            Mt[e] = a;
            continue;
          }
          e -= n * m;
          if (e < n) {
            double a = 0.0;
            for (int q = 0; q < D; ++q) a += Gx[q * n + e] * wv[q];
            qt[e] = D ? lx[e] + a : lx[e];
            continue;
          }
          e -= n;
          {
            double a = 0.0;
            for (int q = 0; q < D; ++q) a += Gu[q * m + e] * wv[q];
            const double v = D ? lu[e] + a : lu[e];
            rt[e] = v;
            rvar[(size_t)t * m + e] = v;
          }
        }
      }
      __syncwarp();
      if (act) {
        // Q_uu = 0.5 (R + BtP B + R^T + B^T P^T B), Q_ux = BtP A + M^T, Q_x_v = q + A^T p_v, Q_u_v = r + B^T p_v  (:446-455)
        for (int idx = r; idx < m * m + m * n + nv * n + nv * m; idx += G) {
          int e = idx;
          if (e < m * m) {
            const int i = e / m, j = e - i * m;
            double a1 = 0.0, a2 = 0.0;
            for (int l = 0; l < n; ++l) a1 += BtP[i * n + l] * Bm[l * m + j];
            for (int l = 0; l < n; ++l) {
              double cc = 0.0;
              for (int q = 0; q < n; ++q) cc += Bm[q * m + i] * P[l * n + q];
              a2 += cc * Bm[l * m + j];
            }
            const double v = 0.5 * (((Rt[i * m + j] + a1) + Rt[j * m + i]) + a2);
            Quu[e] = v;
            Qf[e] = v;
            continue;
          }
          e -= m * m;
          if (e < m * n) {
            const int i = e / n, j = e - i * n;
            double a = 0.0;
            for (int l = 0; l < n; ++l) a += BtP[i * n + l] * A[l * n + j];
            const double v = a + Mt[j * m + i];
            Qux[e] = v;
            RHS[i * nc2 + j] = v;
            continue;
          }
          e -= m * n;
          if (e < nv * n) {
            const int v = e / n, i = e - v * n;
            double a = 0.0;
            for (int l = 0; l < n; ++l) a += A[l * n + i] * pv[v * n + l];
            Qxv[e] = qt[i] + a;
            continue;
          }
          e -= nv * n;
          {
            const int v = e / m, i = e - v * m;
            double a = 0.0;
            for (int l = 0; l < n; ++l) a += Bm[l * m + i] * pv[v * n + l];
            const double val = rt[i] + a;
            Quv[e] = val;
            RHS[i * nc2 + n + v] = val;
          }
        }
      }
      __syncwarp();
      bool fail = false;
      if (act && r == 0) tr[CDDP_B200_MAX_N] = ldlt_small_t<NC>(Qf, tr, m) ? 1 : 0;  // Eigen::LDLT(Q_uu) (:457-461)
      __syncwarp();
      if (act) {
        fail = tr[CDDP_B200_MAX_N] == 0;
        if (!fail)
          for (int col = r; col < nc2; col += G) {  // K = -solve(Q_ux), k_v = -solve(Q_u_v) (:463-464)
            ldlt_solve_t<NC>(Qf, tr, m, RHS + col, nc2);
            for (int i = 0; i < m; ++i) RHS[i * nc2 + col] = -RHS[i * nc2 + col];
          }
      }
      __syncwarp();
      if (act && !fail) {
        for (int idx = r; idx < m * n + nv * m; idx += G) {  // Q_uu K, Q_uu k_v
          if (idx < m * n) {
            const int i = idx / n, j = idx - i * n;
            double a = 0.0;
            for (int l = 0; l < m; ++l) a += Quu[i * m + l] * RHS[l * nc2 + j];
            QuuK[idx] = a;
            gK[((size_t)t * m + i) * n + j] = RHS[i * nc2 + j];
          } else {
            const int e = idx - m * n, v = e / m, i = e - v * m;
            double a = 0.0;
            for (int l = 0; l < m; ++l) a += Quu[i * m + l] * RHS[l * nc2 + n + v];
            Quuk[e] = a;
            kvar[((size_t)v * N + t) * m + i] = RHS[i * nc2 + n + v];
          }
        }
      }
      __syncwarp();
      if (act && !fail) {
        // P = Q + A^T P A + Q_xu K + K^T Q_ux + K^T Q_uu K (:465-467) -> Qt (in place) ; p_v (:468-469) -> Qxv (in place)
        for (int idx = r; idx < n * n + nv * n; idx += G) {
          if (idx < n * n) {
            const int i = idx / n, j = idx - i * n;
            double apa = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
            for (int l = 0; l < n; ++l) apa += A[l * n + i] * PA[l * n + j];
            for (int l = 0; l < m; ++l) {
              a1 += Qux[l * n + i] * RHS[l * nc2 + j];
              a2 += RHS[l * nc2 + i] * Qux[l * n + j];
              a3 += RHS[l * nc2 + i] * QuuK[l * n + j];
            }
            Sh[idx] = (((Qt[idx] + apa) + a1) + a2) + a3;
          } else {
            const int e = idx - n * n, v = e / n, i = e - v * n;
            double a1 = 0.0, a2 = 0.0, a3 = 0.0;
            for (int l = 0; l < m; ++l) {
              a1 += Qux[l * n + i] * RHS[l * nc2 + n + v];
              a2 += RHS[l * nc2 + i] * Quv[v * m + l];
              a3 += RHS[l * nc2 + i] * Quuk[v * m + l];
            }
            dxv[e] = ((Qxv[e] + a1) + a2) + a3;
          }
        }
      }
      __syncwarp();
      if (act && !fail) {
        bool fin = true;
        for (int idx = r; idx < n * n; idx += G) {
          const int i = idx / n, j = idx - i * n;
          const double v = 0.5 * (Sh[i * n + j] + Sh[j * n + i]);
          P[idx] = v;
          fin = fin && finite_d(v);
        }
        for (int e = r; e < nv * n; e += G) {
          const int v = e / n, i = e - v * n;
          pv[e] = dxv[e];
          pvar[((size_t)v * (N + 1) + t) * n + i] = dxv[e];
          fin = fin && finite_d(dxv[e]);
        }
        for (int i = r; i < m * nc2; i += G) fin = fin && finite_d(RHS[i]);
        if (!fin) tr[CDDP_B200_MAX_N + 1] = 1;
      }
      __syncwarp();
      if (act && !fail && tr[CDDP_B200_MAX_N + 1] == 1) fail = true;  // !allFinite() (:470-474)
      if (act && fail) act = false;
      cp_async_wait_all();
      __syncwarp();
    }
    if (need) {
      if (act) {  // the sweep reached t = 0
        need = false;
        ok = true;
      } else if (mode == BW_SINGLE) {  // sequential LQR failed: the backward pass fails
        need = false;
      } else {  // regularisation retry (cddp_solver_base.cpp:93-111, cddp_core.cpp:308-326)
        reg = fmin(reg * c.opt.reg_update_factor, c.opt.reg_max_value);
        ++failures;
        if (reg >= c.opt.reg_max_value) {
          status = CDDP_B200_STATUS_REG_LIMIT;
          need = false;
        }
      }
    }
    __syncwarp();
  }

  if (__any_sync(0xffffffffu, ok)) {
    // ---------------------------------------------------------------- rollouts of the p+1 variants (:540-546)
    const int oK = rs, okv = oK + m * n;
    auto issue_roll = [&](int tt, int x) {
      double *dst = stg + x * blk;
      for (int i = r; i < rs; i += G) cp_async8(dst + i, grec + (size_t)tt * rs + i);
      for (int i = r; i < m * n; i += G) cp_async8(dst + oK + i, gK + (size_t)tt * m * n + i);
      for (int i = r; i < nv * m; i += G) {
        const int v = i / m, j = i - v * m;
        cp_async8(dst + okv + i, kvar + ((size_t)v * N + tt) * m + j);
      }
    };
    if (ok) {
      for (int i = r; i < nv * n; i += G) dxv[i] = 0.0;
      issue_roll(0, 0);
    }
    cp_async_wait_all();
    __syncwarp();
    for (int t = 0; t < N; ++t) {
      const int bi = t & 1;
      if (ok && t + 1 < N) issue_roll(t + 1, bi ^ 1);
      const double *st_ = stg + bi * blk;
      const double *A = st_, *Bm = st_ + n * n;
      if (ok)
        for (int e = r; e < nv * m; e += G) {  // du_v = k_v + K dx_v
          const int v = e / m, i = e - v * m;
          double a = 0.0;
          for (int j = 0; j < n; ++j) a += st_[oK + i * n + j] * dxv[v * n + j];
          Quuk[e] = st_[okv + e] + a;
        }
      __syncwarp();
      if (ok)
        for (int e = r; e < nv * n; e += G) {
          const int v = e / n, i = e - v * n;
          double a1 = 0.0, a2 = 0.0;
          for (int j = 0; j < n; ++j) a1 += A[i * n + j] * dxv[v * n + j];
          for (int j = 0; j < m; ++j) a2 += Bm[i * m + j] * Quuk[v * m + j];
          dxn[e] = (a1 + a2) + 0.0;
        }
      __syncwarp();
      if (ok)
        for (int e = r; e < nv * n; e += G) dxv[e] = dxn[e];
      cp_async_wait_all();
      __syncwarp();
    }
    // ---------------------------------------------------------------- multiplier step: regularised least squares (:548-623)
    double *rhs = vec, *Atb = vec + n, *lam = vec + 2 * n, *best = vec + 3 * n, *ev = vec + 4 * n;
    if (ok && r == 0) {
      for (int i = 0; i < n; ++i) {
        for (int j = 0; j < n; ++j) As[i * n + j] = dxv[(j + 1) * n + i] - dxv[i];          // A_small = H_T S = S
        rhs[i] = -(gX[(size_t)N * n + i] - xref[i]) - dxv[i];                                // b_T - H_T xT_0, b_T = -h_T
      }
      double trace = 0.0, rn = 0.0;
      for (int i = 0; i < n; ++i) {
        for (int j = 0; j < n; ++j) {
          double a = 0.0;
          for (int l = 0; l < n; ++l) a += As[l * n + i] * As[l * n + j];
          AtA[i * n + j] = a;
        }
        double a = 0.0;
        for (int l = 0; l < n; ++l) a += As[l * n + i] * rhs[l];
        Atb[i] = a;
        trace += AtA[i * n + i];
        rn += rhs[i] * rhs[i];
        best[i] = 0.0;
      }
      const double trace_term = trace > 1.0 ? trace / (double)n : 1.0;
      const double base_floor = fmax(1e-10, ic.io.jacobian_regularization_value * pow(fmax(mu, 0.0), ic.io.jacobian_regularization_exponent));
      const double regq = fmax(base_floor, 1e-6 * trace_term);
      // singular values of A_small = sqrt(eig(A^T A)): cyclic Jacobi on a copy (stands in for Eigen::JacobiSVD, :557-560)
      for (int i = 0; i < n * n; ++i) Sh[i] = 0.5 * (AtA[i] + AtA[(i % n) * n + i / n]);
      for (int sweep = 0; sweep < 64; ++sweep) {
        double off = 0.0;
        for (int i = 0; i < n; ++i)
          for (int j = i + 1; j < n; ++j) off += Sh[i * n + j] * Sh[i * n + j];
        if (off < 1e-300) break;
        for (int p_ = 0; p_ < n; ++p_)
          for (int q = p_ + 1; q < n; ++q) {
            const double apq = Sh[p_ * n + q];
            if (apq == 0.0) continue;
            const double th = (Sh[q * n + q] - Sh[p_ * n + p_]) / (2.0 * apq);
            const double tt = (th >= 0 ? 1.0 : -1.0) / (fabs(th) + sqrt(th * th + 1.0));
            const double cs = 1.0 / sqrt(tt * tt + 1.0), sn = tt * cs;
            for (int k = 0; k < n; ++k) {
              const double akp = Sh[k * n + p_], akq = Sh[k * n + q];
              Sh[k * n + p_] = cs * akp - sn * akq;
              Sh[k * n + q] = sn * akp + cs * akq;
            }
            for (int k = 0; k < n; ++k) {
              const double apk = Sh[p_ * n + k], aqk = Sh[q * n + k];
              Sh[p_ * n + k] = cs * apk - sn * aqk;
              Sh[q * n + k] = sn * apk + cs * aqk;
            }
          }
      }
      double smax = 0.0, smin = pos_inf();
      for (int i = 0; i < n; ++i) {
        ev[i] = sqrt(fmax(Sh[i * n + i], 0.0));
        smax = fmax(smax, ev[i]);
        smin = fmin(smin, ev[i]);
      }
      const double reg_base = fmax(regq, fmax(1e-8 * smax - smin, 0.0));
      const double cap = 100.0 * (1.0 + sqrt(rn));
      const double scales[5] = {1.0, 10.0, 100.0, 1e3, 1e4};
      double best_res = pos_inf();
      bool found = false;
      for (int si = 0; si < 5; ++si) {
        const double reg_i = fmax(reg_base * scales[si], 1e-12);
        for (int i = 0; i < n * n; ++i) Sh[i] = AtA[i];
        for (int i = 0; i < n; ++i) Sh[i * n + i] += reg_i;
        if (!ldlt_small(Sh, tr, n)) continue;
        for (int i = 0; i < n; ++i) lam[i] = Atb[i];
        ldlt_solve(Sh, tr, n, lam, 1);
        bool fin = true;
        double ln = 0.0;
        for (int i = 0; i < n; ++i) {
          fin = fin && finite_d(lam[i]);
          ln += lam[i] * lam[i];
        }
        if (!fin) continue;
        ln = sqrt(ln);
        if (ln > cap)
          for (int i = 0; i < n; ++i) lam[i] *= cap / fmax(ln, 1e-12);
        double res = 0.0;
        for (int i = 0; i < n; ++i) {
          double a = 0.0;
          for (int j = 0; j < n; ++j) a += As[i * n + j] * lam[j];
          res += (a - rhs[i]) * (a - rhs[i]);
        }
        res = sqrt(res);
        if (!finite_d(res)) continue;
        if (!found || res < best_res) {
          for (int i = 0; i < n; ++i) best[i] = lam[i];
          best_res = res;
          found = true;
        }
      }
      for (int i = 0; i < n; ++i) ip.dlamT[(size_t)b * n + i] = best[i];  // dLambda_T_eq_ = lambda_delta (:1267)
    }
    __syncwarp();
    // ---------------------------------------------------------------- combination (:625-636) + inf_du, step_norm (:1268-1274)
    if (ok) {
      for (int t = r; t < N; t += G) {  // time-parallel: no recursion here
        for (int i = 0; i < m; ++i) {
          const double k0 = kvar[(size_t)t * m + i];
          double kk = k0;
          for (int v = 0; v < n; ++v) kk += best[v] * (kvar[((size_t)(v + 1) * N + t) * m + i] - k0);
          gk[(size_t)t * m + i] = kk;
          step_norm = fmax(step_norm, fabs(kk));
        }
        const double *Bg = grec + (size_t)t * rs + n * n;
        for (int i = 0; i < m; ++i) {
          double a = 0.0;
          for (int l = 0; l < n; ++l) {
            const double p0 = pvar[(size_t)(t + 1) * n + l];
            double pl = p0;
            for (int v = 0; v < n; ++v) pl += best[v] * (pvar[((size_t)(v + 1) * (N + 1) + t + 1) * n + l] - p0);
            a += Bg[l * m + i] * pl;
          }
          inf_du = fmax(inf_du, fabs(rvar[(size_t)t * m + i] + a));
        }
      }
    }
#pragma unroll
    for (int o = G / 2; o > 0; o >>= 1) {
      inf_du = fmax(inf_du, __shfl_xor_sync(0xffffffffu, inf_du, o));
      step_norm = fmax(step_norm, __shfl_xor_sync(0xffffffffu, step_norm, o));
    }
    __syncwarp();
    // ---------------------------------------------------------------- final rollout: slack / dual gains, dS, dY, step caps (:1276-1320)
    const int fX = rs, fY = fX + n, fS = fY + D, fG = fS + D, fK = fG + D, fk = fK + m * n;
    auto issue_fin = [&](int tt, int x) {
      double *dst = stg + x * blk;
      for (int i = r; i < rs; i += G) cp_async8(dst + i, grec + (size_t)tt * rs + i);
      for (int i = r; i < n; i += G) cp_async8(dst + fX + i, gX + (size_t)tt * n + i);
      for (int i = r; i < D; i += G) {
        cp_async8(dst + fY + i, gY + (size_t)tt * D + i);
        cp_async8(dst + fS + i, gS + (size_t)tt * D + i);
        cp_async8(dst + fG + i, gG + (size_t)tt * D + i);
      }
      for (int i = r; i < m * n; i += G) cp_async8(dst + fK + i, gK + (size_t)tt * m * n + i);
      for (int i = r; i < m; i += G) cp_async8(dst + fk + i, gk + (size_t)tt * m + i);
    };
    const double tau_b = fmax(ic.io.min_fraction_to_boundary, 1.0 - mu);
    if (ok) {
      for (int i = r; i < n; i += G) dxv[i] = 0.0;
      issue_fin(0, 0);
    }
    cp_async_wait_all();
    __syncwarp();
    for (int t = 0; t < N; ++t) {
      const int bi = t & 1;
      if (ok && t + 1 < N) issue_fin(t + 1, bi ^ 1);
      const double *st_ = stg + bi * blk;
      const double *A = st_, *Bm = st_ + n * n, *Kt = st_ + fK, *kt = st_ + fk;
      if (ok && D) barrier_terms(st_ + fX, st_ + fY, st_ + fS, st_ + fG);
      __syncwarp();
      if (ok) {
        for (int q = r; q < D; q += G) {
          double temp = 0.0;
          for (int i = 0; i < m; ++i) temp += Gu[q * m + i] * kt[i];
          const size_t e = (size_t)t * D + q;
          const double kyq = clip_signed(rhat[q] + st_[fY + q] * temp, ssafe[q]);
          const double ksq = -prim[q] - temp;
          gky[e] = kyq;
          gks[e] = ksq;
          double a1 = 0.0, a2 = 0.0;
          for (int j = 0; j < n; ++j) {
            double gkk = 0.0;
            for (int i = 0; i < m; ++i) gkk += Gu[q * m + i] * Kt[i * n + j];
            const double qq = Gx[q * n + j] + gkk;
            const double Kyq = clampd(YS[q] * qq, -MAX_BARRIER_RATIO, MAX_BARRIER_RATIO);
            const double Ksq = -Gx[q * n + j] - gkk;
            gKy[e * n + j] = Kyq;
            gKs[e * n + j] = Ksq;
            a1 += Ksq * dxv[j];
            a2 += Kyq * dxv[j];
          }
          const double ds = __dadd_rn(ksq, a1);
          const double dy = clampd(__dadd_rn(kyq, a2), -MAX_BARRIER_RATIO, MAX_BARRIER_RATIO);
          if (ds < 0.0) apm = fmin(apm, __ddiv_rn(__dmul_rn(-tau_b, st_[fS + q]), ds));
          if (dy < 0.0) adm = fmin(adm, __ddiv_rn(__dmul_rn(-tau_b, st_[fY + q]), dy));
        }
        for (int i = r; i < m; i += G) {
          double a = 0.0;
          for (int j = 0; j < n; ++j) a += Kt[i * n + j] * dxv[j];
          Quuk[i] = kt[i] + a;
        }
      }
      __syncwarp();
      if (ok)
        for (int i = r; i < n; i += G) {
          double a1 = 0.0, a2 = 0.0;
          for (int j = 0; j < n; ++j) a1 += A[i * n + j] * dxv[j];
          for (int j = 0; j < m; ++j) a2 += Bm[i * m + j] * Quuk[j];
          dxn[i] = (a1 + a2) + 0.0;
        }
      __syncwarp();
      if (ok)
        for (int i = r; i < n; i += G) dxv[i] = dxn[i];
      cp_async_wait_all();
      __syncwarp();
    }
#pragma unroll
    for (int o = G / 2; o > 0; o >>= 1) {
      apm = fmin(apm, __shfl_xor_sync(0xffffffffu, apm, o));
      adm = fmin(adm, __shfl_xor_sync(0xffffffffu, adm, o));
    }
    apm = clampd(apm, 0.0, 1.0);
    adm = clampd(adm, 0.0, 1.0);
  }

  if (alive && r == 0) {
    d.bw_ok[b] = ok ? 1 : 0;
    d.lin_valid[b] = 1;
    if (ok) {
      d.dV[2 * b] = 0.0;  // dV_ is not accumulated on this branch
      d.dV[2 * b + 1] = 0.0;
      d.inf_du[b] = inf_du;
      ip.step_norm[b] = step_norm;
      ip.inf_pr[b] = inf_pr;
      ip.inf_comp[b] = inf_comp;
      ip.apm[b] = apm;
      ip.adm[b] = adm;
    }
    if (mode == BW_ITERATE) {
      d.reg[b] = reg;
      if (ok) {  // checkEarlyConvergence (:925-958); a terminal equality alone needs no barrier
        bool early;
        if (ic.nc == 0) {
          early = inf_pr < c.opt.tolerance && inf_du < c.opt.tolerance;
        } else {
          const double tol = fmax(c.opt.tolerance, ic.io.barrier_tol_mult * mu);
          early = inf_pr < tol && inf_du < tol && inf_comp < tol && fabs(d.alpha[b]) * step_norm < c.opt.tolerance * 10.0;
        }
        if (early) {
          status = CDDP_B200_STATUS_OPTIMAL;
          ip_record_history(d, ip, b);
        }
      }
      if (status != CDDP_B200_STATUS_RUNNING) d.status[b] = status;
      trace_backward(d, b, d.iter[b], failures, status == CDDP_B200_STATUS_OPTIMAL ? 0xff : status == CDDP_B200_STATUS_REG_LIMIT ? 0xfe : 0);
    }
  }
}

template <int G, int NS, int NC, int DC = 0>
cudaError_t launch_teq_g(const Constants &c, const DeviceState &d, const IpConstants &ic, const IpDevice &ip, int mode, cudaStream_t st) {
  const int n = d.n, m = d.m, D = ic.d;
  const int gpc = kTeqThreads / G;
  const int tab = (D * n + D * m + 3 * D + 1) & ~1;
  const size_t shm = sizeof(double) * ((size_t)n * n + m * m + ((n * n + m * m) & 1) + tab + (size_t)gpc * teq_group_doubles(n, m, D, d.rec_stride));
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(ip_backward_teq_kernel<G, NS, NC, DC>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    if (e != cudaSuccess) return e;
    configured = true;
  }
  if (shm > 200 * 1024) return cudaErrorInvalidValue;
  ip_backward_teq_kernel<G, NS, NC, DC><<<(d.n_slots + gpc - 1) / gpc, kTeqThreads, shm, st>>>(c, d, ic, ip, mode);
  return cudaGetLastError();
}

}  // namespace

cudaError_t launch_ip_backward_teq(const Constants &c, const DeviceState &d, const IpConstants &ic, const IpDevice &ip, int mode,
                                   cudaStream_t st) {
  if (d.n == 2 && d.m == 1) return launch_teq_g<8, 2, 1>(c, d, ic, ip, mode, st);
  static_assert(TeqStage<3, 2>::stride == 26, "engine.h teq_stage_stride");
  if (teq_small_case(d.n, d.m, ic.d))  // BASELINE config #4
    return teq_reg_enabled() ? launch_teq_reg<3, 2, 5>(c, d, ic, ip, mode, st) : launch_teq_g<16, 3, 2, 5>(c, d, ic, ip, mode, st);
  if (d.n == 3 && d.m == 2) return launch_teq_g<16, 3, 2>(c, d, ic, ip, mode, st);
  if (d.n == 4 && d.m == 1) return launch_teq_g<16, 4, 1>(c, d, ic, ip, mode, st);
  if (d.n * d.n <= 64) return launch_teq_g<16, 0, 0>(c, d, ic, ip, mode, st);
  return launch_teq_g<32, 0, 0>(c, d, ic, ip, mode, st);
}

}  // namespace cddp_b200
