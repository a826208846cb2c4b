// Batched CLDDP backward Riccati sweep — one warp per trajectory.
//
// Reference behaviour followed: CLDDPSolver::backwardPass, src/cddp_core/clddp_solver.cpp:79-204
// (Gauss-Newton Q-function assembly :124-128, control-space regularisation :130-131, PD test
// :133-140, unconstrained gains :142-145, BoxQP + free-set feedback :147-178, dV :184-186, value
// recursion with the UNregularised Q_uu and symmetrisation :188-192, inf_du scaling :194-201) and
// the regularisation-retry loop of CDDPSolverBase::solve, src/cddp_core/cddp_solver_base.cpp:93-111
// with CDDP::increaseRegularization / isRegularizationLimitReached, src/cddp_core/cddp_core.cpp:308-326.
//
// This translation unit holds the GENERIC kernel (any n <= 16, m <= 8; dimensions may be runtime).
// The dense blocks of one trajectory live in the warp's slice of shared memory; lanes split the
// output entries of each product.  backward_fast.cu holds the register-tiled specialisations.
#include "boxqp.cuh"
#include "engine.h"

namespace cddp_b200 {

cudaError_t launch_backward_fast(const Constants &c, const DeviceState &d, int mode, cudaStream_t st, bool *handled);

namespace {

constexpr int kWarpsPerCta = 4;

template <int NS, int NC>
struct Dims {
  static constexpr int MS = NS ? NS : CDDP_B200_MAX_N;  // capacity
  static constexpr int MC = NC ? NC : CDDP_B200_MAX_M;
};

__host__ __device__ inline int per_warp_doubles(int n, int m, int rs) {
  // rec[2] | V | vx | PA | PB | Qxx | Qux | Quu | Qx | Qu | K | k | M
  return 2 * rs + n * n + n + n * n + n * m + n * n + m * n + m * m + n + m + m * n + m + m * n + 8;
}

template <int NS, int NC>
__global__ void __launch_bounds__(kWarpsPerCta * 32) backward_kernel(Constants c, DeviceState d, int mode) {
  constexpr int MC = Dims<NS, NC>::MC;
  extern __shared__ double smem[];
  const int n = NS ? NS : d.n, m = NC ? NC : d.m, N = d.N, rs = d.rec_stride;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // CTA-shared constants: 2*Q*dt, 2*R*dt
  double *sQ = smem;
  double *sR = sQ + n * n;
  for (int i = threadIdx.x; i < n * n; i += blockDim.x) sQ[i] = c.Qdt2[i];
  for (int i = threadIdx.x; i < m * m; i += blockDim.x) sR[i] = c.Rdt2[i];
  __syncthreads();
  const int b = slot_instance(d, blockIdx.x * kWarpsPerCta + warp);
  if (b >= d.B) return;
  if (mode == BW_ITERATE && d.status[b] != CDDP_B200_STATUS_RUNNING) return;

  double *w = sR + m * m + ((n * n + m * m) & 1) + (size_t)warp * per_warp_doubles(n, m, rs);
  double *rec0 = w;            w += rs;
  double *rec1 = w;            w += rs;
  double *V = w;               w += n * n;
  double *vx = w;              w += n;
  double *PA = w;              w += n * n;
  double *PB = w;              w += n * m;
  double *Qxx = w;             w += n * n;
  double *Qux = w;             w += m * n;
  double *Quu = w;             w += m * m;
  double *Qx = w;              w += n;
  double *Qu = w;              w += m;
  double *Kt = w;              w += m * n;
  double *kt = w;              w += m;
  double *Mt = w;              w += m * n;

  const double *grec = d.rec + (size_t)b * N * rs;
  double *gK = d.K + (size_t)b * N * m * n;
  double *gk = d.kff + (size_t)b * N * m;

  double reg = d.reg[b];
  if (mode == BW_ITERATE && lane == 0) d.iter[b] += 1;  // ++iter, cddp_solver_base.cpp:75

  bool ok = false;
  double dV0 = 0.0, dV1 = 0.0, inf_du = 0.0;
  int status = CDDP_B200_STATUS_RUNNING, failures = 0;

  while (true) {
    // ---- one sweep at regularisation `reg` ----
    for (int i = lane; i < n * n; i += 32) V[i] = c.Qf2[i];                     // V_xx = 2 Qf
    for (int i = lane; i < n; i += 32) vx[i] = d.vterm[(size_t)b * n + i];       // V_x  = 2 Qf (x_N - ref)
    for (int i = lane; i < rs; i += 32) rec0[i] = grec[(size_t)(N - 1) * rs + i];
    __syncwarp();
    double norm_Vx = 0.0, Qu_err = 0.0;
    for (int i = 0; i < n; ++i) norm_Vx += fabs(vx[i]);
    dV0 = 0.0;
    dV1 = 0.0;
    ok = true;

    for (int t = N - 1; t >= 0; --t) {
      double *rec = ((N - 1 - t) & 1) ? rec1 : rec0;
      double *nxt = ((N - 1 - t) & 1) ? rec0 : rec1;
      // register-staged prefetch of the next record (consumed at the end of this step)
      constexpr int PF = NS ? ((NS * NS + NS * NC + NS + 2 * NC + 1 + 31) / 32) : 13;
      double pf[PF];
      if (t > 0) {
        const double *g = grec + (size_t)(t - 1) * rs;
#pragma unroll
        for (int q = 0; q < PF; ++q) {
          const int i = lane + 32 * q;
          pf[q] = (i < rs) ? g[i] : 0.0;
        }
      }
      const double *A = rec, *Bm = rec + n * n, *lx = Bm + n * m, *lu = lx + n, *un = lu + m;

      // P = V [A|B] ; Q_x = l_x + A^T V_x ; Q_u = l_u + B^T V_x        (:124-125)
      const int nm = n + m;
      for (int idx = lane; idx < n * nm + nm; idx += 32) {
        if (idx < n * nm) {
          const int i = idx / nm, j = idx - i * nm;
          double s = 0.0;
          if (j < n) {
            for (int l = 0; l < n; ++l) s += V[i * n + l] * A[l * n + j];
            PA[i * n + j] = s;
          } else {
            for (int l = 0; l < n; ++l) s += V[i * n + l] * Bm[l * m + (j - n)];
            PB[i * m + (j - n)] = s;
          }
        } else {
          const int j = idx - n * nm;
          double s = 0.0;
          if (j < n) {
            for (int l = 0; l < n; ++l) s += A[l * n + j] * vx[l];
            Qx[j] = lx[j] + s;
          } else {
            for (int l = 0; l < n; ++l) s += Bm[l * m + (j - n)] * vx[l];
            Qu[j - n] = lu[j - n] + s;
          }
        }
      }
      __syncwarp();
      // Q_xx = l_xx + A^T P_A ; Q_ux = B^T P_A ; Q_uu = l_uu + B^T P_B   (:126-128)
      for (int idx = lane; idx < n * n + m * n + m * m; idx += 32) {
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
      __syncwarp();

      // ---- control-space subproblem: every lane redundantly, in registers ----
      double H[MC * MC], g[MC], L[MC * MC], kk[MC];
#pragma unroll
      for (int i = 0; i < MC; ++i)
        if (i < m) {
          g[i] = Qu[i];
#pragma unroll
          for (int j = 0; j < MC; ++j)
            if (j < m) H[i * MC + j] = Quu[i * m + j] + (i == j ? reg : 0.0);  // Q_uu_reg (:130-131)
        }
      const unsigned all = (1u << m) - 1u;
      unsigned free_mask = all;
      // PD test (:133-140)
      if (!SmallMat<MC>::masked_cholesky(m, H, all, L)) {
        ok = false;
        break;
      }
      if (!c.has_box) {  // (:142-145)
#pragma unroll
        for (int i = 0; i < MC; ++i)
          if (i < m) kk[i] = -g[i];
        SmallMat<MC>::chol_solve(m, L, kk);
      } else {  // (:147-159)
        double lo[MC], hi[MC];
#pragma unroll
        for (int i = 0; i < MC; ++i)
          if (i < m) {
            lo[i] = c.lb[i] - un[i];
            hi[i] = c.ub[i] - un[i];
            kk[i] = gk[(size_t)t * m + i];  // warm start x0 = k_u_[t]
          }
        const int qs = SmallMat<MC>::boxqp(c.opt, m, H, g, lo, hi, kk, free_mask, L, true);
        if (qs == QP_HESSIAN_NOT_PD || qs == QP_NO_DESCENT) {
          ok = false;
          break;
        }
        if (c.opt.qp_max_iterations <= 0) SmallMat<MC>::masked_cholesky(m, H, free_mask, L);
      }
      // K = -H_free^{-1} Q_ux[free,:], clamped rows zero (:161-178); lane c owns column c
      for (int col = lane; col < n; col += 32) {
        double rhs[MC];
#pragma unroll
        for (int i = 0; i < MC; ++i)
          if (i < m) rhs[i] = ((free_mask >> i) & 1u) ? -Qux[i * n + col] : 0.0;
        if (free_mask) SmallMat<MC>::chol_solve(m, L, rhs);
#pragma unroll
        for (int i = 0; i < MC; ++i)
          if (i < m) {
            const double v = ((free_mask >> i) & 1u) ? rhs[i] : 0.0;
            Kt[i * n + col] = v;
            gK[((size_t)t * m + i) * n + col] = v;  // K_u_[t] (:182)
          }
      }
      __syncwarp();  // all lanes have read the warm start before k_u_[t] is overwritten
      if (lane < m) {
        double v = 0.0;
#pragma unroll
        for (int i = 0; i < MC; ++i)
          if (i == lane) v = kk[i];
        kt[lane] = v;
        gk[(size_t)t * m + lane] = v;  // k_u_[t] (:181)
      }
      // dV += (Q_u.k, 0.5 k^T Q_uu k) with the UNregularised Q_uu (:184-186)
      double Quuk[MC];
      {
        double d0 = 0.0, d1 = 0.0;
#pragma unroll
        for (int i = 0; i < MC; ++i)
          if (i < m) {
            double s = 0.0;
#pragma unroll
            for (int j = 0; j < MC; ++j)
              if (j < m) s += Quu[i * m + j] * kk[j];
            Quuk[i] = s;
            d0 += g[i] * kk[i];
            d1 += kk[i] * s;
          }
        dV0 += d0;
        dV1 += 0.5 * d1;
      }
      __syncwarp();
      // M = Q_uu K + Q_ux
      for (int idx = lane; idx < m * n; idx += 32) {
        const int i = idx / n, j = idx - i * n;
        double s = Qux[idx];
        for (int a = 0; a < m; ++a) s += Quu[i * m + a] * Kt[a * n + j];
        Mt[idx] = s;
      }
      __syncwarp();
      // V_x = Q_x + K^T Q_uu k + Q_ux^T k + K^T Q_u (:188-189)
      for (int i = lane; i < n; i += 32) {
        double s = Qx[i];
#pragma unroll
        for (int a = 0; a < MC; ++a)
          if (a < m) s += Kt[a * n + i] * Quuk[a] + Qux[a * n + i] * kk[a] + Kt[a * n + i] * g[a];
        vx[i] = s;
      }
      // V_xx = sym(Q_xx + K^T Q_uu K + Q_ux^T K + K^T Q_ux) (:190-192)
      for (int idx = lane; idx < n * n; idx += 32) {
        const int i = idx / n, j = idx - i * n;
        double sij = Qxx[i * n + j], sji = Qxx[j * n + i];
        for (int a = 0; a < m; ++a) {
          sij += Kt[a * n + i] * Mt[a * n + j] + Qux[a * n + i] * Kt[a * n + j];
          sji += Kt[a * n + j] * Mt[a * n + i] + Qux[a * n + j] * Kt[a * n + i];
        }
        V[idx] = 0.5 * (sij + sji);
      }
      // stage the prefetched record
      if (t > 0) {
#pragma unroll
        for (int q = 0; q < PF; ++q) {
          const int i = lane + 32 * q;
          if (i < rs) nxt[i] = pf[q];
        }
      }
      __syncwarp();
      double l1 = 0.0, linf = 0.0;  // (:194-195)
      for (int i = 0; i < n; ++i) l1 += fabs(vx[i]);
#pragma unroll
      for (int i = 0; i < MC; ++i)
        if (i < m) linf = fmax(linf, fabs(g[i]));
      norm_Vx += l1;
      Qu_err = fmax(Qu_err, linf);
    }

    if (ok) {
      double sf = c.opt.termination_scaling_max_factor;  // (:197-201)
      sf = fmax(sf, norm_Vx / (double)(N * n)) / sf;
      inf_du = Qu_err / sf;
      break;
    }
    if (mode == BW_SINGLE) break;
    // backward failure: increaseRegularization + limit test (cddp_solver_base.cpp:95-109)
    reg = fmin(reg * c.opt.reg_update_factor, c.opt.reg_max_value);
    ++failures;
    if (reg >= c.opt.reg_max_value) {
      status = CDDP_B200_STATUS_REG_LIMIT;
      break;
    }
    __syncwarp();
  }

  // white-box: value function at t = 0
  if (ok) {
    for (int i = lane; i < n; i += 32) d.Vx0[(size_t)b * n + i] = vx[i];
    for (int i = lane; i < n * n; i += 32) d.Vxx0[(size_t)b * n * n + i] = V[i];
  }
  if (lane == 0) {
    d.bw_ok[b] = ok ? 1 : 0;
    d.lin_valid[b] = 1;
    if (ok) {
      d.dV[2 * b] = dV0;
      d.dV[2 * b + 1] = dV1;
      d.inf_du[b] = inf_du;
    }
    if (mode == BW_ITERATE) {
      d.reg[b] = reg;
      if (ok && inf_du < c.opt.tolerance) {  // checkEarlyConvergence, clddp_solver.cpp:206-213
        status = CDDP_B200_STATUS_OPTIMAL;
        if (d.history) {  // recordIterationHistory, cddp_solver_base.cpp:116-118
          const int hl = d.history_len[b];
          if (hl < d.history_cap) {
            double *h = d.history + ((size_t)b * d.history_cap + hl) * 4;
            h[0] = d.cost[b];
            h[1] = d.alpha[b];
            h[2] = inf_du;
            h[3] = reg;
            d.history_len[b] = hl + 1;
          }
        }
      }
      if (status != CDDP_B200_STATUS_RUNNING) d.status[b] = status;
      trace_backward(d, b, d.iter[b], failures, status == CDDP_B200_STATUS_OPTIMAL ? 0xff : status == CDDP_B200_STATUS_REG_LIMIT ? 0xfe : 0);
    }
  }
}

template <int NS, int NC>
cudaError_t launch_t(const Constants &c, const DeviceState &d, int mode, cudaStream_t st) {
  const int n = d.n, m = d.m;
  const size_t shm =
      sizeof(double) * ((size_t)n * n + m * m + ((n * n + m * m) & 1) + (size_t)kWarpsPerCta * per_warp_doubles(n, m, d.rec_stride));
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(backward_kernel<NS, NC>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    if (e != cudaSuccess) return e;
    configured = true;
  }
  const int blocks = (d.n_slots + kWarpsPerCta - 1) / kWarpsPerCta;
  backward_kernel<NS, NC><<<blocks, kWarpsPerCta * 32, shm, st>>>(c, d, mode);
  return cudaGetLastError();
}

}  // namespace

cudaError_t launch_backward_generic(const Constants &c, const DeviceState &d, int mode, cudaStream_t st) {
  if (d.n == 2 && d.m == 1) return launch_t<2, 1>(c, d, mode, st);
  if (d.n == 3 && d.m == 2) return launch_t<3, 2>(c, d, mode, st);
  if (d.n == 4 && d.m == 1) return launch_t<4, 1>(c, d, mode, st);
  if (d.n == 13 && d.m == 4) return launch_t<13, 4>(c, d, mode, st);
  return launch_t<0, 0>(c, d, mode, st);
}

cudaError_t launch_backward(const Constants &c, const DeviceState &d, int mode, cudaStream_t st) {
  bool handled = false;
  cudaError_t e = launch_backward_fast(c, d, mode, st, &handled);
  if (handled) return e;
  return launch_backward_generic(c, d, mode, st);
}

}  // namespace cddp_b200
