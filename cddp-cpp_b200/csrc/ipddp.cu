// Batched IPDDP (interior-point DDP) on the device: backward sweep with the primal-dual condensation of the path
// inequality constraints, forward rollout with parallel line-search alphas and the filter acceptance test, barrier
// update and convergence tests — one DDP iteration = linearize_kernel (linearize.cu, shared with CLDDP) +
// ip_backward_kernel + ip_forward_kernel.
//
// Reference behaviour followed (astomodynamics/cddp-cpp @ f71fa80), options.warm_start = false, use_ilqr = true
// (the defaults, options.hpp:223,230), path inequality constraints only (no terminal constraints):
//   IPDDPSolver::initialize (cold start)              src/cddp_core/ipddp_solver.cpp:818-913
//   evaluateTrajectory / initializeDualSlackVariables  :2252-2296, :2428-2482
//   backwardPass, unconstrained branch                 :1055-1118
//   backwardPass, path-constraint branch               :1355-1569 (+ rolloutLinearPolicy :368-392)
//   computeMaxStepSizes                                :2939-2988
//   forwardPass (+ filter acceptance)                  :1571-1876
//   applyForwardPassResult / updateBarrierParameters   :1878-1951, :2548-2660
//   checkEarlyConvergence / checkConvergence           :925-958, :1953-2025
//   handleForwardPassFailure                           :2037-2082
//   computeTheta / computeBarrierMerit / computePrimalAndComplementarity  :2778-2937
//   filter helpers                                     src/cddp_core/interior_point_utils.cpp:81-141
//   constraints                                        include/cddp-cpp/cddp_core/constraint.hpp:144-440
//   outer loop                                         src/cddp_core/cddp_solver_base.cpp:29-186
// The costate bookkeeping (Lambda_, k_lambda_, K_lambda_) only feeds an allFinite() test in the reference
// (:1612-1618, :1665-1671) and is not carried.
//
// Deliberate numerical differences (DESIGN.md "Numerics"): the barrier merit and theta sums run over (t, row) in
// time-major order (the reference sums constraint-major over a std::map of trajectories) and the merit after a
// barrier update is formed as cost - mu * sum(log s) from the carried log-sum; both are reorderings of the same
// sums.  Eigen::LDLT (diagonal pivoting) is followed literally (ldlt_small below), because unlike a Cholesky
// factorisation it does not reject indefinite matrices and the reference relies on that (:1431-1435).
#include "engine.h"
#include "kernels_ipddp.cuh"
#include "ldlt_small.cuh"
#include "user_model_host.h"

namespace cddp_b200 {

namespace {

using namespace kern;

// ---------------------------------------------------------------------------------------------- backward sweep
// A group of G lanes owns one trajectory; its dense blocks live in the group's slice of shared memory and the lanes
// split the output entries of every product (runtime dimensions).  The m x m pivoted LDLT runs on one lane; the
// 1 + n right-hand sides [Q_u | Q_ux] are solved one column per lane.
constexpr int kBwThreads = 128;

__host__ __device__ inline int ip_stage_doubles(int n, int m, int D, int rs) {
  const int sweep = rs + n + 3 * D;                          // record | x | y | s | g of one timestep
  const int roll = rs + m * n + m;                           // record | K | k (linear rollout of dx)
  const int b = sweep > roll ? sweep : roll;
  return (b + 1) & ~1;
}
__host__ __device__ inline int ip_group_doubles(int n, int m, int D, int rs) {
  int c = 2 * ip_stage_doubles(n, m, D, rs);  // double-buffered staging block
  c += n * n + n;                    // V, vx
  c += n * n + n * m;                // PA, PB
  c += n * n + m * n + 2 * m * m;    // Qxx, Qux, Quu, Qr
  c += n + m;                        // Qx, Qu
  c += m * (n + 1);                  // RHS / kK
  c += m * (n > m ? n : m);          // Q_uu K scratch (also stages the m x m condensed Q_uu)
  c += D * n + D * m;                // Gx, Gu
  c += 5 * D;                        // ssafe, YS, prim, rhat, Sir
  c += 3 * n + 2 * m;                // dx, dxn, scratch
  c += CDDP_B200_MAX_M;              // transpositions (ints, generously)
  return (c + 1) & ~1;
}
__host__ __device__ inline int ip_table_doubles(int n, int m, int D) {  // constraint table staged in shared memory
  // Gx | Gu | off | scale | (row_type, row_bdim) packed as ints in D doubles | sparsity masks (64-bit each): rows per
  // state column, rows per control column, controls per row
  return D * n + D * m + 2 * D + D + (n + m + D);
}

// NS, NC: compile-time state / control dimensions (0 = runtime): with constants the inner products unroll and the
// index arithmetic (idx / n, idx % n) strength-reduces — about half of the generic kernel's instruction stream.
template <int G, int NS, int NC, int DC = 0>
__global__ void __launch_bounds__(kBwThreads) ip_backward_kernel(Constants c, DeviceState d, IpConstants ic, IpDevice ip,
                                                                 int mode) {
  extern __shared__ double smem[];
  const int n = NS ? NS : d.n, m = NC ? NC : d.m, N = d.N, rs = d.rec_stride, D = DC ? DC : ic.d;
  constexpr int GPC = kBwThreads / G;  // groups per CTA
  double *sQ = smem;
  double *sR = sQ + n * n;
  double *tGx = sR + m * m + ((n * n + m * m) & 1);  // constraint table (batch-shared)
  double *tGu = tGx + D * n;
  double *tOff = tGu + D * m;
  double *tScale = tOff + D;
  int *tType = reinterpret_cast<int *>(tScale + D);
  int *tBdim = tType + D;
  // Structural sparsity of the constraint Jacobians (batch-shared): every supported row touches EITHER the state
  // (StateConstraint / LinearConstraint / BallConstraint rows, Q_yu row = 0) OR the control (ControlConstraint rows,
  // Q_yx row = 0), and a box row has a single +-1.  The reference multiplies the dense d x n / d x m blocks
  // (ipddp_solver.cpp:1427-1447, :1461-1497); here every sum over rows visits only the rows whose entry is
  // structurally non-zero, in ascending row order — the same sum without its exact-zero terms.
  static_assert(IP_MAX_DUAL <= 64, "row masks are 64-bit");
  unsigned long long *mColX = reinterpret_cast<unsigned long long *>(tScale + 2 * D);  // [n] rows with Gx(q, j) != 0
  unsigned long long *mColU = mColX + n;                                                // [m] rows with Gu(q, a) != 0
  unsigned long long *mRowU = mColU + m;                                                // [D] controls with Gu(q, a) != 0
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
  for (int j = threadIdx.x; j < n + m + D; j += blockDim.x) {
    unsigned long long mk = 0ull;
    if (j < n) {
      for (int q = 0; q < D; ++q) {
        const int ty = ic.row_type[q];
        const bool nz = (ty == IP_ROW_STATE && ic.Gx[q * n + j] != 0.0) || (ty == IP_ROW_BALL && j < ic.row_bdim[q]);
        if (nz) mk |= 1ull << q;
      }
      mColX[j] = mk;
    } else if (j < n + m) {
      const int a = j - n;
      for (int q = 0; q < D; ++q)
        if (ic.row_type[q] == IP_ROW_CONTROL && ic.Gu[q * m + a] != 0.0) mk |= 1ull << q;
      mColU[a] = mk;
    } else {
      const int q = j - n - m;
      for (int a = 0; a < m; ++a)
        if (ic.row_type[q] == IP_ROW_CONTROL && ic.Gu[q * m + a] != 0.0) mk |= 1ull << a;
      mRowU[q] = mk;
    }
  }
  __syncthreads();
  // sum over the rows (or controls) in a mask, ascending; with a compile-time dual dimension the dense unrolled loop
  // over all `cnt` indices is kept (tiny d: the loop overhead of the sparse walk would cost more than the zeros)
  auto for_bits = [&](unsigned long long mk, int cnt, auto &&f) {
    if constexpr (DC != 0) {
      for (int q = 0; q < cnt; ++q) f(q);
    } else {
      unsigned lo = (unsigned)mk, hi = (unsigned)(mk >> 32);  // two 32-bit walks: a 64-bit find-first-set is emulated
      while (lo) {
        const int q = __ffs((int)lo) - 1;
        lo &= lo - 1;
        f(q);
      }
      while (hi) {
        const int q = 31 + __ffs((int)hi);
        hi &= hi - 1;
        f(q);
      }
    }
  };
  const int grp = threadIdx.x / G, r = threadIdx.x % G;
  const int b = slot_instance(d, blockIdx.x * GPC + grp);
  const bool alive = b < d.B && !(mode == BW_ITERATE && d.status[b] != CDDP_B200_STATUS_RUNNING);
  const int bb = alive ? b : 0;
  const int blk = ip_stage_doubles(n, m, D, rs);
  bool any_ball = false;
  for (int q = 0; q < D; ++q) any_ball = any_ball || tType[q] == IP_ROW_BALL;
  double *w = tGx + ((ip_table_doubles(n, m, D) + 1) & ~1) + (size_t)grp * ip_group_doubles(n, m, D, rs);
  double *stg = w;  w += 2 * blk;
  double *V = w;    w += n * n;
  double *vx = w;   w += n;
  double *PA = w;   w += n * n;
  double *PB = w;   w += n * m;
  double *Qxx = w;  w += n * n;
  double *Qux = w;  w += m * n;
  double *Quu = w;  w += m * m;
  double *Qr = w;   w += m * m;
  double *Qx = w;   w += n;
  double *Qu = w;   w += m;
  double *RHS = w;  w += m * (n + 1);  // [m][1+n]: column 0 -> k, columns 1..n -> K
  double *QK = w;   w += m * (n > m ? n : m);
  double *Gx = w;   w += D * n;
  double *Gu = w;   w += D * m;
  double *ssafe = w; w += D;
  double *YS = w;   w += D;
  double *prim = w; w += D;
  double *rhat = w; w += D;
  double *Sir = w;  w += D;
  double *dx = w;   w += n;
  double *dxn = w;  w += n;
  double *scr = w;  w += n + 2 * m;
  int *tr = reinterpret_cast<int *>(w);

  const int cur = d.cur[bb];
  const double *grec = d.rec + (size_t)bb * N * rs;
  const double *gX = d.X[cur] + (size_t)bb * (N + 1) * n;
  const double *gY = ip.Y[cur] + (size_t)bb * N * D, *gS = ip.S[cur] + (size_t)bb * N * D, *gG = ip.G[cur] + (size_t)bb * N * D;
  double *gK = d.K + (size_t)bb * N * m * n, *gk = d.kff + (size_t)bb * N * m;
  double *gky = ip.ky + (size_t)bb * N * D, *gks = ip.ks + (size_t)bb * N * D;
  double *gKy = ip.Ky + (size_t)bb * N * D * n, *gKs = ip.Ks + (size_t)bb * N * D * n;
  const int nc1 = n + 1;
  // stage the operands of timestep tt into staging buffer `x`: record | x_t | y_t | s_t | g_t
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

  const double mu = alive ? ip.mu[bb] : 1.0;
  double reg = alive ? d.reg[bb] : 0.0;
  if (alive && mode == BW_ITERATE && r == 0) d.iter[b] += 1;  // ++iter, cddp_solver_base.cpp:75
  int status = CDDP_B200_STATUS_RUNNING;
  bool need = alive, ok = false;
  int failures = 0;  // backward-pass failures of this iteration (decision trace)
  double dV0 = 0.0, dV1 = 0.0, inf_du = 0.0, inf_pr = 0.0, inf_comp = 0.0, step_norm = 0.0;

  while (__any_sync(0xffffffffu, need)) {
    bool act = need;  // this group sweeps in this round
    if (act) {
      // V_xx = sym(2 Qf), V_x = 2 Qf (x_N - ref)   (:983-990)
      for (int i = r; i < n * n; i += G) {
        const int a = i / n, e = i - a * n;
        V[i] = 0.5 * (c.Qf2[a * n + e] + c.Qf2[e * n + a]);
      }
      for (int i = r; i < n; i += G) vx[i] = d.vterm[(size_t)bb * n + i];
    }
    dV0 = dV1 = inf_du = inf_pr = inf_comp = step_norm = 0.0;
    if (act) {  // constraint Jacobians (precomputeConstraintGradients :2145-2250): the state-independent rows
      for (int i = r; i < D * n; i += G) Gx[i] = (tType[i / n] == IP_ROW_STATE) ? tGx[i] : 0.0;
      for (int i = r; i < D * m; i += G) Gu[i] = (tType[i / m] == IP_ROW_CONTROL) ? tGu[i] : 0.0;
    }
    if (act) issue_sweep(N - 1, 0);
    cp_async_wait_all();
    __syncwarp();
    for (int t = N - 1; t >= 0; --t) {
      const int bi = (N - 1 - t) & 1;
      if (act && t > 0) issue_sweep(t - 1, bi ^ 1);  // consumed one step from now
      const double *rec = stg + bi * blk, *xs = rec + rs, *ys = xs + n, *ss = ys + D, *gs = ss + D;
      const double *A = rec, *Bm = rec + n * n, *lx = rec + d.offLx, *lu = rec + d.offLu;
      if (act) {
        // constraint Jacobians of this step (precomputeConstraintGradients :2145-2250) and the barrier terms (:1413-1443)
        if (any_ball) {  // only BallConstraint rows depend on x_t; the others were written once before the sweep
          for (int i = r; i < D * n; i += G) {
            const int row = i / n, j = i - row * n;
            if (tType[row] == IP_ROW_BALL) Gx[i] = (j < tBdim[row]) ? -2.0 * tScale[row] * (xs[j] - tGx[i]) : 0.0;
          }
        }
        for (int i = r; i < D; i += G) {
          const double sf = fmax(ss[i], fmax(mu * 1e-3, EPS_SLACK));
          ssafe[i] = sf;
          YS[i] = clip_pos(ys[i], sf);
          const double pr = gs[i] + ss[i];
          const double cp = ys[i] * ss[i] - mu;
          prim[i] = pr;
          const double rh = ys[i] * pr - cp;
          rhat[i] = rh;
          Sir[i] = clip_signed(rh, sf);
        }
        // P = V [A|B]
        for (int idx = r; idx < n * (n + m); idx += G) {
          const int i = idx / (n + m), j = idx - i * (n + m);
          double s = 0.0;
          if (j < n) {
            for (int l = 0; l < n; ++l) s += V[i * n + l] * A[l * n + j];
            PA[i * n + j] = s;
          } else {
            for (int l = 0; l < n; ++l) s += V[i * n + l] * Bm[l * m + (j - n)];
            PB[i * m + (j - n)] = s;
          }
        }
      }
      __syncwarp();
      if (act) {
        // Q_x = l_x + Q_yx^T y + A^T V_x ; Q_u = l_u + Q_yu^T y + B^T V_x   (:1393-1394)
        for (int j = r; j < n + m; j += G) {
          if (j < n) {
            double gy = 0.0, av = 0.0;
            for_bits(mColX[j], D, [&](int q) { gy += Gx[q * n + j] * ys[q]; });
            for (int l = 0; l < n; ++l) av += A[l * n + j] * vx[l];
            Qx[j] = D ? (lx[j] + gy) + av : lx[j] + av;
          } else {
            const int a = j - n;
            double gy = 0.0, bv = 0.0;
            for_bits(mColU[a], D, [&](int q) { gy += Gu[q * m + a] * ys[q]; });
            for (int l = 0; l < n; ++l) bv += Bm[l * m + a] * vx[l];
            Qu[a] = D ? (lu[a] + gy) + bv : lu[a] + bv;
          }
        }
        // Q_xx = l_xx + A^T P_A ; Q_ux = B^T P_A ; Q_uu = l_uu + B^T P_B   (:1395-1397)
        for (int idx = r; idx < n * n + m * n + m * m; idx += G) {
          double s = 0.0;
          if (idx < n * n) {
            const int i = idx / n, j = idx - i * n;
            for (int l = 0; l < n; ++l) s += A[l * n + i] * PA[l * n + j];
            Qxx[idx] = sQ[idx] + s;
          } else if (idx < n * n + m * n) {
            const int e = idx - n * n, i = e / n, j = e - i * n;
            for (int l = 0; l < n; ++l) s += Bm[l * m + i] * PA[l * n + j];
            Qux[e] = s;
          } else {
            const int e = idx - n * n - m * n, i = e / m, j = e - i * m;
            for (int l = 0; l < n; ++l) s += Bm[l * m + i] * PB[l * m + j];
            Quu[e] = sR[e] + s;
          }
        }
      }
      __syncwarp();
      if (act) {
        // Q_uu_reg = sym(Q_uu) + Q_yu^T YSinv Q_yu + reg I (:1427-1429; unconstrained :1083-1084) and
        // bigRHS = [Q_u + Q_yu^T S^-1 rhat | Q_ux + Q_yu^T YSinv Q_yx]   (:1444-1447)
        for (int idx = r; idx < m * m + m * nc1; idx += G) {
          if (idx < m * m) {
            const int i = idx / m, j = idx - i * m;
            double acc = 0.0;
            for_bits(mColU[i] & mColU[j], D, [&](int q) { acc += Gu[q * m + i] * (YS[q] * Gu[q * m + j]); });
            Qr[idx] = 0.5 * (Quu[i * m + j] + Quu[j * m + i]) + acc + (i == j ? reg : 0.0);
          } else {
            const int e = idx - m * m, i = e / nc1, j = e - i * nc1;
            double acc = 0.0;
            if (j == 0) {
              for_bits(mColU[i], D, [&](int q) { acc += Gu[q * m + i] * Sir[q]; });
              RHS[e] = D ? Qu[i] + acc : Qu[i];
            } else {
              for_bits(mColU[i] & mColX[j - 1], D, [&](int q) { acc += Gu[q * m + i] * (YS[q] * Gx[q * n + (j - 1)]); });
              RHS[e] = D ? Qux[i * n + (j - 1)] + acc : Qux[i * n + (j - 1)];
            }
          }
        }
      }
      __syncwarp();
      // condensed Q terms that need the pre-solve RHS (:1493-1497), before RHS is overwritten by the solve
      if (act) {
        for (int idx = r; idx < m + m * n + n + n * n + m * m; idx += G) {
          int e = idx;
          if (e < m) {
            Qu[e] = RHS[e * nc1];
            continue;
          }
          e -= m;
          if (e < m * n) {
            const int i = e / n, j = e - i * n;
            Qux[e] = RHS[i * nc1 + 1 + j];
            continue;
          }
          e -= m * n;
          if (e < n) {
            double acc = 0.0;
            for_bits(mColX[e], D, [&](int q) { acc += Gx[q * n + e] * Sir[q]; });
            if (D) Qx[e] += acc;
            continue;
          }
          e -= n;
          if (e < n * n) {
            const int i = e / n, j = e - i * n;
            double acc = 0.0;
            for_bits(mColX[i] & mColX[j], D, [&](int q) { acc += Gx[q * n + i] * (YS[q] * Gx[q * n + j]); });
            if (D) Qxx[e] += acc;
            continue;
          }
          e -= n * n;
          {
            const int i = e / m, j = e - i * m;
            if (D) {
              double acc = 0.0;
              for_bits(mColU[i] & mColU[j], D, [&](int q) { acc += Gu[q * m + i] * (YS[q] * Gu[q * m + j]); });
              QK[e] = Quu[e] + acc;  // staged: Quu is still being read by other lanes' Qr (done above) — safe after sync
            } else {
              QK[e] = Qr[e];  // unconstrained branch: the symmetrised + regularised Q_uu enters V and dV (:1083-1084)
            }
          }
        }
      }
      __syncwarp();
      bool fail = false;
      if (act) {
        for (int i = r; i < m * m; i += G) Quu[i] = QK[i];
        if (r == 0) {
          const bool good = ldlt_small_t<NC>(Qr, tr, m);  // Eigen::LDLT(Q_uu_reg) (:1431)
          tr[CDDP_B200_MAX_M] = good ? 1 : 0;
        }
      }
      __syncwarp();
      if (act) {
        fail = tr[CDDP_B200_MAX_M] == 0;  // ldlt.info() != Success -> backward pass fails (:1432-1435)
        if (!fail) {
          for (int col = r; col < nc1; col += G) {  // kK = -ldlt.solve(bigRHS) (:1449)
            ldlt_solve_t<NC>(Qr, tr, m, RHS + col, nc1);
            for (int i = 0; i < m; ++i) RHS[i * nc1 + col] = -RHS[i * nc1 + col];
          }
        }
      }
      __syncwarp();
      if (act && !fail) {
        // gains to HBM (:1458-1459)
        for (int i = r; i < m * nc1; i += G) {
          const int a = i / nc1, j = i - a * nc1;
          if (j == 0) gk[(size_t)t * m + a] = RHS[i];
          else gK[((size_t)t * m + a) * n + (j - 1)] = RHS[i];
        }
        // k_y, K_y, k_s, K_s (:1461-1491)
        for (int idx = r; idx < D * nc1; idx += G) {
          const int q = idx / nc1, j = idx - q * nc1;
          if (j == 0) {
            double temp = 0.0;
            for_bits(mRowU[q], m, [&](int i) { temp += Gu[q * m + i] * RHS[i * nc1]; });
            gky[(size_t)t * D + q] = clip_signed(rhat[q] + ys[q] * temp, ssafe[q]);
            gks[(size_t)t * D + q] = -prim[q] - temp;
          } else {
            double gkk = 0.0;
            for_bits(mRowU[q], m, [&](int i) { gkk += Gu[q * m + i] * RHS[i * nc1 + j]; });
            const double qq = Gx[q * n + (j - 1)] + gkk;
            gKy[((size_t)t * D + q) * n + (j - 1)] = clampd(YS[q] * qq, -MAX_BARRIER_RATIO, MAX_BARRIER_RATIO);
            gKs[((size_t)t * D + q) * n + (j - 1)] = -Gx[q * n + (j - 1)] - gkk;
          }
        }
        // Q_uu K, Q_uu k
        for (int idx = r; idx < m * n + m; idx += G) {
          if (idx < m * n) {
            const int i = idx / n, j = idx - i * n;
            double acc = 0.0;
            for (int l = 0; l < m; ++l) acc += Quu[i * m + l] * RHS[l * nc1 + 1 + j];
            QK[idx] = acc;
          } else {
            const int i = idx - m * n;
            double acc = 0.0;
            for (int l = 0; l < m; ++l) acc += Quu[i * m + l] * RHS[l * nc1];
            scr[i] = acc;
          }
        }
      }
      __syncwarp();
      if (act && !fail) {
        // dV (:1499-1500), infeasibility measures (:1510-1513): every lane redundantly (tiny)
        double d0 = 0.0, d1 = 0.0;
        for (int i = 0; i < m; ++i) {
          const double ki = RHS[i * nc1];
          d0 += ki * Qu[i];
          d1 += ki * scr[i];
          inf_du = fmax(inf_du, fabs(Qu[i]));
          step_norm = fmax(step_norm, fabs(ki));
        }
        dV0 += d0;
        dV1 += 0.5 * d1;
        for (int q = r; q < D; q += G) {  // this lane's rows; the group maximum is formed after the sweep
          inf_pr = fmax(inf_pr, fabs(prim[q]));
          inf_comp = fmax(inf_comp, fabs(ys[q] * ss[q] - mu));
        }
        // V_x = Q_x + K^T Q_u + Q_ux^T k + K^T Q_uu k (:1502-1503) -> dxn (scratch) ; V_xx (:1504-1506) -> PA (scratch)
        for (int idx = r; idx < n + n * n; idx += G) {
          if (idx < n) {
            const int i = idx;
            double a1 = 0.0, a2 = 0.0, a3 = 0.0;
            for (int l = 0; l < m; ++l) {
              const double Kli = RHS[l * nc1 + 1 + i];
              a1 += Kli * Qu[l];
              a2 += Qux[l * n + i] * RHS[l * nc1];
              a3 += Kli * scr[l];
            }
            dxn[i] = ((Qx[i] + a1) + a2) + a3;
          } else {
            const int e = idx - n, i = e / n, j = e - i * n;
            double a1 = 0.0, a2 = 0.0, a3 = 0.0;
            for (int l = 0; l < m; ++l) {
              const double Kli = RHS[l * nc1 + 1 + i];
              a1 += Kli * Qux[l * n + j];
              a2 += Qux[l * n + i] * RHS[l * nc1 + 1 + j];
              a3 += Kli * QK[l * n + j];
            }
            PA[e] = ((Qxx[e] + a1) + a2) + a3;
          }
        }
      }
      __syncwarp();
      if (act && !fail) {
        for (int i = r; i < n; i += G) vx[i] = dxn[i];
        for (int e = r; e < n * n; e += G) {
          const int i = e / n, j = e - i * n;
          V[e] = 0.5 * (PA[i * n + j] + PA[j * n + i]);
        }
      }
      if (act && fail) act = false;  // idle until the end of this sweep; `need` stays set -> retry
      else if (act && t == 0) {
        need = false;
        ok = true;
      }
      cp_async_wait_all();  // the next step's operands have landed
      __syncwarp();
    }
    if (need && alive) {  // this group's sweep failed
      if (mode == BW_SINGLE) {
        need = false;
      } else {  // increaseRegularization + limit test (cddp_solver_base.cpp:95-109, cddp_core.cpp:308-326)
        reg = fmin(reg * c.opt.reg_update_factor, c.opt.reg_max_value);
        ++failures;
        if (reg >= c.opt.reg_max_value) {
          status = CDDP_B200_STATUS_REG_LIMIT;
          need = false;
        }
      }
    }
  }

#pragma unroll
  for (int o = G / 2; o > 0; o >>= 1) {
    inf_pr = fmax(inf_pr, __shfl_xor_sync(0xffffffffu, inf_pr, o));
    inf_comp = fmax(inf_comp, __shfl_xor_sync(0xffffffffu, inf_comp, o));
  }
  // linearised slack / dual steps and the fraction-to-boundary step caps: rolloutLinearPolicy from dx0 = 0 (:368-392),
  // dS = k_s + K_s dX, dY = clamp(k_y + K_y dX) (:1516-1538), computeMaxStepSizes (:2939-2988)
  double apm = 1.0, adm = 1.0;
  const bool roll = ok && D > 0;
  if (__any_sync(0xffffffffu, roll)) {
    const double tau_b = fmax(ic.io.min_fraction_to_boundary, 1.0 - mu);
    // Phase 1 (sequential in t, tiny): dx_{t+1} = A dx_t + B (k + K dx_t), staged record | K | k one step ahead; every
    // dx_t goes to a scratch row in HBM.  The scratch is the instance's CANDIDATE state buffer X[cur ^ 1]: nothing reads it
    // between this sweep and the forward pass, which rewrites it.
    // Phase 2 (parallel over (t, row), streaming): the slack / dual steps and their fraction-to-boundary caps.  K_s, K_y
    // (2 d n doubles per timestep) are read straight from HBM by consecutive lanes instead of being staged per timestep —
    // staging them made the group's shared-memory block 1.8x larger than the sweep itself needs.
    const int oK = rs, ok_ = oK + m * n;
    double *dxs = d.X[cur ^ 1] + (size_t)bb * (N + 1) * n;
    auto issue_roll = [&](int tt, int x) {
      double *dst = stg + x * blk;
      for (int i = r; i < rs; i += G) cp_async8(dst + i, grec + (size_t)tt * rs + i);
      for (int i = r; i < m * n; i += G) cp_async8(dst + oK + i, gK + (size_t)tt * m * n + i);
      for (int i = r; i < m; i += G) cp_async8(dst + ok_ + i, gk + (size_t)tt * m + i);
    };
    if (roll) {
      for (int i = r; i < n; i += G) dx[i] = 0.0;
      issue_roll(0, 0);
    }
    cp_async_wait_all();
    __syncwarp();
    for (int t = 0; t < N; ++t) {
      const int bi = t & 1;
      if (roll && t + 1 < N) issue_roll(t + 1, bi ^ 1);
      const double *st_ = stg + bi * blk;
      const double *A = st_, *Bm = st_ + n * n;
      if (roll) {
        for (int i = r; i < n; i += G) dxs[(size_t)t * n + i] = dx[i];
        for (int i = r; i < m; i += G) {
          double acc = 0.0;
          for (int j = 0; j < n; ++j) acc += st_[oK + i * n + j] * dx[j];
          scr[i] = st_[ok_ + i] + acc;
        }
      }
      __syncwarp();
      if (roll)
        for (int i = r; i < n; i += G) {
          double a1 = 0.0, a2 = 0.0;
          for (int j = 0; j < n; ++j) a1 += A[i * n + j] * dx[j];
          for (int j = 0; j < m; ++j) a2 += Bm[i * m + j] * scr[j];
          dxn[i] = (a1 + a2) + 0.0;
        }
      __syncwarp();
      if (roll)
        for (int i = r; i < n; i += G) dx[i] = dxn[i];
      cp_async_wait_all();
      __syncwarp();
    }
    if (roll) {
      for (int idx = r; idx < N * D; idx += G) {
        const int t = idx / D;
        const double *ksr = gKs + (size_t)idx * n, *kyr = gKy + (size_t)idx * n, *dxt = dxs + (size_t)t * n;
        double a1 = 0.0, a2 = 0.0;
        for (int j = 0; j < n; ++j) {
          a1 += ksr[j] * dxt[j];
          a2 += kyr[j] * dxt[j];
        }
        const double ds = __dadd_rn(gks[idx], a1);
        const double dy = clampd(__dadd_rn(gky[idx], a2), -MAX_BARRIER_RATIO, MAX_BARRIER_RATIO);
        if (ds < 0.0) apm = fmin(apm, __ddiv_rn(__dmul_rn(-tau_b, gS[idx]), ds));
        if (dy < 0.0) adm = fmin(adm, __ddiv_rn(__dmul_rn(-tau_b, gY[idx]), dy));
      }
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
      d.dV[2 * b] = dV0;
      d.dV[2 * b + 1] = dV1;
      d.inf_du[b] = inf_du;
      ip.step_norm[b] = step_norm;
      ip.inf_pr[b] = D ? inf_pr : 0.0;  // (:1113-1116, :1565-1568)
      ip.inf_comp[b] = D ? inf_comp : 0.0;
      ip.apm[b] = apm;
      ip.adm[b] = adm;
    }
    if (mode == BW_ITERATE) {
      d.reg[b] = reg;
      if (ok) {  // checkEarlyConvergence (:925-958)
        bool early;
        const double ipr = D ? inf_pr : 0.0, icp = D ? inf_comp : 0.0;
        if (ic.nc == 0) {
          early = ipr < c.opt.tolerance && inf_du < c.opt.tolerance;
        } else {
          const double tol = fmax(c.opt.tolerance, ic.io.barrier_tol_mult * mu);
          early = ipr < tol && inf_du < tol && icp < tol && fabs(d.alpha[b]) * step_norm < c.opt.tolerance * 10.0;
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

template <int MODEL, int DC>
cudaError_t launch_ip_forward_d(const Constants &c, const DeviceState &d, const IpConstants &ic, const IpDevice &ip, int mode,
                                cudaStream_t st) {
  const int per_cta = kFwThreads / 16;
  const size_t shm = sizeof(double) * (size_t)ip_fw_smem_doubles(d.n, d.m, ic.d);
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(ip_forward_kernel<MODEL, DC>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    if (e != cudaSuccess) return e;
    configured = true;
  }
  if (shm > 200 * 1024) return cudaErrorInvalidValue;
  ip_forward_kernel<MODEL, DC><<<(d.n_slots + per_cta - 1) / per_cta, kFwThreads, shm, st>>>(c, d, ic, ip, mode);
  return cudaGetLastError();
}

template <int MODEL>
cudaError_t launch_ip_forward_model(const Constants &c, const DeviceState &d, const IpConstants &ic, const IpDevice &ip, int mode,
                                    cudaStream_t st) {
  // compile-time dual dimension for the BASELINE config #4 constraint set (control box 2m = 4 + ball 1)
  if (MODEL == CDDP_B200_MODEL_UNICYCLE && ic.d == 5) return launch_ip_forward_d<MODEL, 5>(c, d, ic, ip, mode, st);
  return launch_ip_forward_d<MODEL, 0>(c, d, ic, ip, mode, st);
}

template <int MODEL>
cudaError_t launch_ip_init_model(const Constants &c, const DeviceState &d, const IpConstants &ic, const IpDevice &ip,
                                 cudaStream_t st) {
  ip_initialize_kernel<MODEL><<<(d.B + 63) / 64, 64, 0, st>>>(c, d, ic, ip);
  return cudaGetLastError();
}

template <int G, int NS, int NC, int DC = 0>
cudaError_t launch_ip_backward_g(const Constants &c, const DeviceState &d, const IpConstants &ic, const IpDevice &ip, int mode,
                                 cudaStream_t st) {
  const int n = d.n, m = d.m;
  const int gpc = kBwThreads / G;
  const size_t shm = sizeof(double) * ((size_t)n * n + m * m + ((n * n + m * m) & 1) + ((ip_table_doubles(n, m, ic.d) + 1) & ~1) +
                                       (size_t)gpc * ip_group_doubles(n, m, ic.d, d.rec_stride));
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(ip_backward_kernel<G, NS, NC, DC>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    if (e != cudaSuccess) return e;
    configured = true;
  }
  if (shm > 200 * 1024) return cudaErrorInvalidValue;
  ip_backward_kernel<G, NS, NC, DC><<<(d.n_slots + gpc - 1) / gpc, kBwThreads, shm, st>>>(c, d, ic, ip, mode);
  return cudaGetLastError();
}

}  // namespace

cudaError_t launch_ip_initialize(const Constants &c, const DeviceState &d, const IpConstants &ic, const IpDevice &ip,
                                 cudaStream_t st) {
  switch (c.model) {
    case CDDP_B200_MODEL_PENDULUM: return launch_ip_init_model<CDDP_B200_MODEL_PENDULUM>(c, d, ic, ip, st);
    case CDDP_B200_MODEL_CARTPOLE: return launch_ip_init_model<CDDP_B200_MODEL_CARTPOLE>(c, d, ic, ip, st);
    case CDDP_B200_MODEL_UNICYCLE: return launch_ip_init_model<CDDP_B200_MODEL_UNICYCLE>(c, d, ic, ip, st);
    case CDDP_B200_MODEL_QUADROTOR: return launch_ip_init_model<CDDP_B200_MODEL_QUADROTOR>(c, d, ic, ip, st);
    case CDDP_B200_MODEL_USER: return launch_user_ip_initialize(c, d, ic, ip, st);
    default: return cudaErrorInvalidValue;
  }
}

cudaError_t launch_ip_backward(const Constants &c, const DeviceState &d, const IpConstants &ic, const IpDevice &ip, int mode,
                               cudaStream_t st) {
  if (ic.teq) return launch_ip_backward_teq(c, d, ic, ip, mode, st);  // terminal-equality branch (ipddp_teq.cu)
  // lanes per trajectory: the widest per-step loop has n*(n+m) entries
  if (d.n == 2 && d.m == 1) return launch_ip_backward_g<8, 2, 1>(c, d, ic, ip, mode, st);
  if (d.n == 3 && d.m == 2 && ic.d == 5) return launch_ip_backward_g<16, 3, 2, 5>(c, d, ic, ip, mode, st);
  if (d.n == 3 && d.m == 2) return launch_ip_backward_g<16, 3, 2>(c, d, ic, ip, mode, st);
  if (d.n == 4 && d.m == 1) return launch_ip_backward_g<8, 4, 1>(c, d, ic, ip, mode, st);
  if (d.n == 13 && d.m == 4) return launch_ip_backward_g<32, 13, 4>(c, d, ic, ip, mode, st);
  if (d.n == 14 && d.m == 7) return launch_ip_backward_g<32, 14, 7>(c, d, ic, ip, mode, st);
  if (d.n * (d.n + d.m) <= 24) return launch_ip_backward_g<8, 0, 0>(c, d, ic, ip, mode, st);
  if (d.n * (d.n + d.m) <= 64) return launch_ip_backward_g<16, 0, 0>(c, d, ic, ip, mode, st);
  return launch_ip_backward_g<32, 0, 0>(c, d, ic, ip, mode, st);
}

cudaError_t launch_ip_forward(const Constants &c, const DeviceState &d, const IpConstants &ic, const IpDevice &ip, int mode,
                              cudaStream_t st) {
  switch (c.model) {
    case CDDP_B200_MODEL_PENDULUM: return launch_ip_forward_model<CDDP_B200_MODEL_PENDULUM>(c, d, ic, ip, mode, st);
    case CDDP_B200_MODEL_CARTPOLE: return launch_ip_forward_model<CDDP_B200_MODEL_CARTPOLE>(c, d, ic, ip, mode, st);
    case CDDP_B200_MODEL_UNICYCLE: return launch_ip_forward_model<CDDP_B200_MODEL_UNICYCLE>(c, d, ic, ip, mode, st);
    case CDDP_B200_MODEL_QUADROTOR: return launch_ip_forward_model<CDDP_B200_MODEL_QUADROTOR>(c, d, ic, ip, mode, st);
    case CDDP_B200_MODEL_USER: return launch_user_ip_forward(c, d, ic, ip, mode, st);
    default: return cudaErrorInvalidValue;
  }
}

}  // namespace cddp_b200
