// In-kernel dense helpers for the control-space (m x m) subproblem of the backward sweep:
// Cholesky-based positive-definiteness test, gains, and the Tassa-2014 projected-Newton BoxQP.
//
// Reference behaviour followed (astomodynamics/cddp-cpp @ f71fa80):
//   BoxQPSolver::solve        src/cddp_core/boxqp.cpp:25-182
//   BoxQPSolver::lineSearch   src/cddp_core/boxqp.cpp:207-233
//   initializeX / project     src/cddp_core/boxqp.cpp:184-205, 241-250
//   PD test + gains           src/cddp_core/clddp_solver.cpp:130-178
//
// Deliberate numerical substitutions (documented in DESIGN.md "Numerics"):
//   * Eigen::EigenSolver min-eigenvalue<=0 test  -> Cholesky pivot<=0 test (same verdict for a
//     symmetric matrix except within roundoff of singular).
//   * Eigen::LDLT (pivoted) on the gathered free block -> unpivoted Cholesky of the MASKED matrix
//     (clamped rows/columns replaced by identity), which factors exactly the free block without a
//     gather/scatter and keeps every lane on fixed-size m x m code.
//   * MatrixXd::inverse() then multiply -> Cholesky solve.
//   * The factor keeps 1/l_jj on its diagonal (rsqrt of the pivot) and the substitutions multiply by it: an FP64
//     division is a ~17-instruction dependent sequence and a 7 x 7 inverse from the factor used to spend 112 of
//     them per timestep on the sweep's critical path.
// Every lane of the warp executes this code redundantly on identical data (one problem per warp,
// so there is no intra-warp divergence); results live in registers of all lanes.
#pragma once
#include "engine.h"

namespace cddp_b200 {

enum {
  QP_HESSIAN_NOT_PD = -1, QP_NO_DESCENT = 0, QP_MAX_ITER_EXCEEDED = 1, QP_MAX_LS_EXCEEDED = 2,
  QP_NO_BOUNDS = 3, QP_SUCCESS = 4, QP_ALL_CLAMPED = 5
};

// M = compile-time capacity, m = runtime size (== M for specialised kernels)
template <int M>
struct SmallMat {
  // Cholesky of the masked matrix; returns false if a pivot is not > 0 (also for NaN).  L holds the strict lower
  // triangle of the factor and, ON THE DIAGONAL, the reciprocals 1/l_jj.
  __device__ __forceinline__ static bool masked_cholesky(int m, const double *H, unsigned free_mask, double *L) {
#pragma unroll
    for (int j = 0; j < M; ++j) {
      if (j < m) {
        const bool fj = (free_mask >> j) & 1u;
        double s = fj ? H[j * M + j] : 1.0;
#pragma unroll
        for (int k = 0; k < M; ++k)
          if (k < j) s -= L[j * M + k] * L[j * M + k];
        if (!(s > 0.0)) return false;
        const double inv = rsqrt(s);
        L[j * M + j] = inv;
#pragma unroll
        for (int i = 0; i < M; ++i) {
          if (i > j && i < m) {
            const bool fi = (free_mask >> i) & 1u;
            double v = (fi && fj) ? H[i * M + j] : 0.0;
#pragma unroll
            for (int k = 0; k < M; ++k)
              if (k < j) v -= L[i * M + k] * L[j * M + k];
            L[i * M + j] = v * inv;
          }
        }
      }
    }
    return true;
  }

  // solves (L L^T) y = b in place
  __device__ __forceinline__ static void chol_solve(int m, const double *L, double *b) {
#pragma unroll
    for (int i = 0; i < M; ++i) {
      if (i < m) {
        double s = b[i];
#pragma unroll
        for (int k = 0; k < M; ++k)
          if (k < i) s -= L[i * M + k] * b[k];
        b[i] = s * L[i * M + i];
      }
    }
#pragma unroll
    for (int ii = 0; ii < M; ++ii) {
      const int i = M - 1 - ii;
      if (i < m) {
        double s = b[i];
#pragma unroll
        for (int k = 0; k < M; ++k)
          if (k > i && k < m) s -= L[k * M + i] * b[k];
        b[i] = s * L[i * M + i];
      }
    }
  }

  // column `col` of (L L^T)^-1: the same substitutions with b = e_col, skipping the leading zeros of the forward pass
  // (call with a compile-time col from an unrolled loop)
  __device__ __forceinline__ static void chol_inverse_column(int m, const double *L, int col, double *b) {
#pragma unroll
    for (int i = 0; i < M; ++i) {
      if (i < col || i >= m) {
        b[i] = 0.0;
      } else if (i == col) {
        b[i] = L[i * M + i];
      } else {
        double s = 0.0;
#pragma unroll
        for (int k = 0; k < M; ++k)
          if (k >= col && k < i) s -= L[i * M + k] * b[k];
        b[i] = s * L[i * M + i];
      }
    }
#pragma unroll
    for (int ii = 0; ii < M; ++ii) {
      const int i = M - 1 - ii;
      if (i < m) {
        double s = b[i];
#pragma unroll
        for (int k = 0; k < M; ++k)
          if (k > i && k < m) s -= L[k * M + i] * b[k];
        b[i] = s * L[i * M + i];
      }
    }
  }

  __device__ __forceinline__ static double qp_value(int m, const double *H, const double *g, const double *x) {
    double a = 0.0, bb = 0.0;
#pragma unroll
    for (int i = 0; i < M; ++i) {
      if (i < m) {
        double hx = 0.0;
#pragma unroll
        for (int j = 0; j < M; ++j)
          if (j < m) hx += H[i * M + j] * x[j];
        a += x[i] * hx;
        bb += g[i] * x[i];
      }
    }
    return 0.5 * a + bb;
  }

  // BoxQPSolver::solve.  x: in = warm start x0 (boxqp.cpp:184-187, x0 always supplied by CLDDP),
  // out = solution.  free_mask/L: final free set and its (masked) Cholesky factor.  L_is_full: on entry L already
  // holds the factor of the unmasked H (the caller's PD test), which is what iteration 0 needs when nothing is clamped.
  __device__ static int boxqp(const cddp_b200_options &o, int m, const double *H, const double *g, const double *lo,
                              const double *hi, double *x, unsigned &free_mask, double *L, bool L_is_full = false) {
    const unsigned all = (m >= 32) ? 0xffffffffu : ((1u << m) - 1u);
    int status = QP_MAX_ITER_EXCEEDED;
#pragma unroll
    for (int i = 0; i < M; ++i)
      if (i < m) x[i] = clamp_box(x[i], lo[i], hi[i]);
    unsigned clamped = 0u;
    free_mask = all;
    double value = qp_value(m, H, g, x);
    double old_value = __longlong_as_double(0x7ff0000000000000LL);
    const double gtol2 = o.qp_min_gradient_norm * o.qp_min_gradient_norm;
    for (int iter = 0; iter < o.qp_max_iterations; ++iter) {
      if (iter > 0 && fabs(old_value - value) < o.qp_min_relative_improvement * fabs(old_value)) {
        status = QP_SUCCESS;
        break;
      }
      old_value = value;
      double grad[M];
#pragma unroll
      for (int i = 0; i < M; ++i) {
        if (i < m) {
          double s = 0.0;
#pragma unroll
          for (int j = 0; j < M; ++j)
            if (j < m) s += H[i * M + j] * x[j];
          grad[i] = g[i] + s;
        }
      }
      const unsigned old_clamped = clamped;
      clamped = 0u;
#pragma unroll
      for (int i = 0; i < M; ++i)
        if (i < m)
          if ((x[i] == lo[i] && grad[i] > 0.0) || (x[i] == hi[i] && grad[i] < 0.0)) clamped |= (1u << i);
      free_mask = all & ~clamped;
      if (clamped == all) {
        status = QP_ALL_CLAMPED;
        break;
      }
      if ((iter == 0 && !(L_is_full && clamped == 0u)) || (iter > 0 && clamped != old_clamped)) {
        if (!masked_cholesky(m, H, free_mask, L)) {
          status = QP_HESSIAN_NOT_PD;
          break;
        }
      }
      double gn = 0.0;
#pragma unroll
      for (int i = 0; i < M; ++i)
        if (i < m && ((free_mask >> i) & 1u)) gn += grad[i] * grad[i];
      if (gn < gtol2) {  // sqrt(gn) < min_gradient_norm
        status = QP_SUCCESS;
        break;
      }
      double rhs[M];
#pragma unroll
      for (int j = 0; j < M; ++j) {
        if (j < m) {
          double s = g[j];
#pragma unroll
          for (int i = 0; i < M; ++i)
            if (i < m && ((clamped >> i) & 1u)) s += H[j * M + i] * x[i];
          rhs[j] = ((free_mask >> j) & 1u) ? s : 0.0;
        }
      }
      chol_solve(m, L, rhs);
      double search[M];
      double sdotg = 0.0;
#pragma unroll
      for (int i = 0; i < M; ++i) {
        if (i < m) {
          search[i] = ((free_mask >> i) & 1u) ? (-rhs[i] - x[i]) : 0.0;
          sdotg += search[i] * grad[i];
        }
      }
      if (sdotg >= 0.0) {
        status = QP_NO_DESCENT;
        break;
      }
      double step = 1.0, vn = 0.0;
      bool ls_ok = false;
      double xn[M];
      while (step > o.qp_min_step_size) {
#pragma unroll
        for (int i = 0; i < M; ++i)
          if (i < m) xn[i] = clamp_box(x[i] + step * search[i], lo[i], hi[i]);
        vn = qp_value(m, H, g, xn);
        if ((vn - value) <= o.qp_armijo_constant * step * sdotg) {
          ls_ok = true;
          break;
        }
        step *= o.qp_step_decrease_factor;
      }
      if (!ls_ok) {
        status = QP_MAX_LS_EXCEEDED;
        break;
      }
#pragma unroll
      for (int i = 0; i < M; ++i)
        if (i < m) x[i] = xn[i];
      value = qp_value(m, H, g, x);
    }
    return status;
  }
};

}  // namespace cddp_b200
