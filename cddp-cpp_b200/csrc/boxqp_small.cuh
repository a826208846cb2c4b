// Low-latency control-space subproblem for m <= 4, one trajectory per LANE (the QP warp of the sweep
// kernel).  The backward sweep is a chain of N dependent timesteps, so its run time is set by the
// LATENCY of one step; the m x m subproblem sits in the middle of that chain.  FP64 division and
// square root are ~10-instruction dependent sequences on the GPU, and a Cholesky factorisation is
// m of them back to back, so here the (masked) inverse of the symmetric m x m matrix is formed in
// closed form from 2x2 minors (adjugate / determinant): a dependency depth of ~6 FMAs plus ONE
// reciprocal, with all minors independent.  Positive-definiteness is decided by Sylvester's
// criterion on the leading principal minors, which fall out of the same minors.
//
// Reference behaviour followed (astomodynamics/cddp-cpp @ f71fa80): BoxQPSolver::solve,
// src/cddp_core/boxqp.cpp:25-182 (+ lineSearch :207-233, initializeX :184-205); same iteration
// structure, exit tests and status codes as SmallMat<M>::boxqp in boxqp.cuh.
//
// Numerical substitutions (DESIGN.md "Numerics"): Eigen::LDLT on the gathered free block -> closed-form
// inverse of the MASKED matrix (clamped rows/columns replaced by identity; the free block of the
// result is the inverse of the free block); `sqrt(|grad_free|^2) < min_gradient_norm` is evaluated
// as `|grad_free|^2 < min_gradient_norm^2`.
#pragma once
#include "boxqp.cuh"

namespace cddp_b200 {

// Hm = masked symmetric matrix (full storage, Hm[i*M+j] == Hm[j*M+i]).  Writes Inv (full, symmetric) and
// returns true iff all leading principal minors are > 0 (false also for NaN).
template <int M>
struct SymInverse;

template <>
struct SymInverse<1> {
  __device__ __forceinline__ static bool run(const double *a, double *b) {
    const bool pd = a[0] > 0.0;
    b[0] = 1.0 / a[0];
    return pd;
  }
};

template <>
struct SymInverse<2> {
  __device__ __forceinline__ static bool run(const double *a, double *b) {
    const double det = a[0] * a[3] - a[1] * a[1];
    const double id = 1.0 / det;
    b[0] = a[3] * id;
    b[1] = b[2] = -a[1] * id;
    b[3] = a[0] * id;
    return (a[0] > 0.0) && (det > 0.0);
  }
};

template <>
struct SymInverse<3> {
  __device__ __forceinline__ static bool run(const double *a, double *b) {
    const double c00 = a[4] * a[8] - a[5] * a[5];
    const double c01 = a[2] * a[5] - a[1] * a[8];
    const double c02 = a[1] * a[5] - a[2] * a[4];
    const double c11 = a[0] * a[8] - a[2] * a[2];
    const double c12 = a[1] * a[2] - a[0] * a[5];
    const double c22 = a[0] * a[4] - a[1] * a[1];  // leading 2x2 minor
    const double det = a[0] * c00 + a[1] * c01 + a[2] * c02;
    const double id = 1.0 / det;
    b[0] = c00 * id;
    b[1] = b[3] = c01 * id;
    b[2] = b[6] = c02 * id;
    b[4] = c11 * id;
    b[5] = b[7] = c12 * id;
    b[8] = c22 * id;
    return (a[0] > 0.0) && (c22 > 0.0) && (det > 0.0);
  }
};

template <>
struct SymInverse<4> {
  __device__ __forceinline__ static bool run(const double *a, double *b) {
    // symmetric: a(i,j) == a(j,i); index helper
#define A_(i, j) a[(i) * 4 + (j)]
    const double s0 = A_(0, 0) * A_(1, 1) - A_(0, 1) * A_(0, 1);
    const double s1 = A_(0, 0) * A_(1, 2) - A_(0, 1) * A_(0, 2);
    const double s2 = A_(0, 0) * A_(1, 3) - A_(0, 1) * A_(0, 3);
    const double s3 = A_(0, 1) * A_(1, 2) - A_(1, 1) * A_(0, 2);
    const double s4 = A_(0, 1) * A_(1, 3) - A_(1, 1) * A_(0, 3);
    const double s5 = A_(0, 2) * A_(1, 3) - A_(1, 2) * A_(0, 3);
    const double c5 = A_(2, 2) * A_(3, 3) - A_(2, 3) * A_(2, 3);
    const double c4 = A_(1, 2) * A_(3, 3) - A_(1, 3) * A_(2, 3);
    const double c3 = A_(1, 2) * A_(2, 3) - A_(1, 3) * A_(2, 2);
    const double c2 = A_(0, 2) * A_(3, 3) - A_(0, 3) * A_(2, 3);
    const double c1 = A_(0, 2) * A_(2, 3) - A_(0, 3) * A_(2, 2);
    const double c0 = A_(0, 2) * A_(1, 3) - A_(0, 3) * A_(1, 2);
    const double det = (s0 * c5 - s1 * c4) + (s2 * c3 + s3 * c2) + (s5 * c0 - s4 * c1);
    const double id = 1.0 / det;
    const double n00 = A_(1, 1) * c5 - A_(1, 2) * c4 + A_(1, 3) * c3;
    const double n01 = -A_(0, 1) * c5 + A_(0, 2) * c4 - A_(0, 3) * c3;
    const double n02 = A_(1, 3) * s5 - A_(2, 3) * s4 + A_(3, 3) * s3;
    const double n03 = -A_(1, 2) * s5 + A_(2, 2) * s4 - A_(2, 3) * s3;
    const double n11 = A_(0, 0) * c5 - A_(0, 2) * c2 + A_(0, 3) * c1;
    const double n12 = -A_(0, 3) * s5 + A_(2, 3) * s2 - A_(3, 3) * s1;
    const double n13 = A_(0, 2) * s5 - A_(2, 2) * s2 + A_(2, 3) * s1;
    const double n22 = A_(0, 3) * s4 - A_(1, 3) * s2 + A_(3, 3) * s0;
    const double n23 = -A_(0, 2) * s4 + A_(1, 2) * s2 - A_(2, 3) * s0;
    const double n33 = A_(0, 2) * s3 - A_(1, 2) * s1 + A_(2, 2) * s0;  // leading 3x3 minor
#undef A_
    b[0] = n00 * id;
    b[1] = b[4] = n01 * id;
    b[2] = b[8] = n02 * id;
    b[3] = b[12] = n03 * id;
    b[5] = n11 * id;
    b[6] = b[9] = n12 * id;
    b[7] = b[13] = n13 * id;
    b[10] = n22 * id;
    b[11] = b[14] = n23 * id;
    b[15] = n33 * id;
    return (a[0] > 0.0) && (s0 > 0.0) && (n33 > 0.0) && (det > 0.0);
  }
};

// Sylvester's criterion on the FULL matrix only (no inverse): the PD test of clddp_solver.cpp:133-140.
template <int M>
struct SymPD;
template <>
struct SymPD<1> {
  __device__ __forceinline__ static bool run(const double *a) { return a[0] > 0.0; }
};
template <>
struct SymPD<2> {
  __device__ __forceinline__ static bool run(const double *a) { return (a[0] > 0.0) && (a[0] * a[3] - a[1] * a[1] > 0.0); }
};
template <>
struct SymPD<3> {
  __device__ __forceinline__ static bool run(const double *a) {
    const double c00 = a[4] * a[8] - a[5] * a[5], c01 = a[2] * a[5] - a[1] * a[8], c02 = a[1] * a[5] - a[2] * a[4];
    return (a[0] > 0.0) && (a[0] * a[4] - a[1] * a[1] > 0.0) && (a[0] * c00 + a[1] * c01 + a[2] * c02 > 0.0);
  }
};
template <>
struct SymPD<4> {
  __device__ __forceinline__ static bool run(const double *a) {
#define A_(i, j) a[(i) * 4 + (j)]
    const double s0 = A_(0, 0) * A_(1, 1) - A_(0, 1) * A_(0, 1);
    const double s1 = A_(0, 0) * A_(1, 2) - A_(0, 1) * A_(0, 2);
    const double s2 = A_(0, 0) * A_(1, 3) - A_(0, 1) * A_(0, 3);
    const double s3 = A_(0, 1) * A_(1, 2) - A_(1, 1) * A_(0, 2);
    const double s4 = A_(0, 1) * A_(1, 3) - A_(1, 1) * A_(0, 3);
    const double s5 = A_(0, 2) * A_(1, 3) - A_(1, 2) * A_(0, 3);
    const double c5 = A_(2, 2) * A_(3, 3) - A_(2, 3) * A_(2, 3);
    const double c4 = A_(1, 2) * A_(3, 3) - A_(1, 3) * A_(2, 3);
    const double c3 = A_(1, 2) * A_(2, 3) - A_(1, 3) * A_(2, 2);
    const double c2 = A_(0, 2) * A_(3, 3) - A_(0, 3) * A_(2, 3);
    const double c1 = A_(0, 2) * A_(2, 3) - A_(0, 3) * A_(2, 2);
    const double c0 = A_(0, 2) * A_(1, 3) - A_(0, 3) * A_(1, 2);
    const double det = (s0 * c5 - s1 * c4) + (s2 * c3 + s3 * c2) + (s5 * c0 - s4 * c1);
    const double n33 = A_(0, 2) * s3 - A_(1, 2) * s1 + A_(2, 2) * s0;
#undef A_
    return (a[0] > 0.0) && (s0 > 0.0) && (n33 > 0.0) && (det > 0.0);
  }
};

template <int M>
struct SmallQP {
  // inverse of the masked matrix (clamped rows/columns replaced by identity).  The result is block diagonal: its
  // free block is the inverse of the free block of H; clamped rows/columns come back as identity rows — callers
  // mask their right-hand sides / results instead of paying 2*M*M selects here.  H must be symmetric-stored.
  __device__ __forceinline__ static bool masked_inverse_raw(const double *H, unsigned free_mask, double *Inv) {
    double Hm[M * M];  // symmetric: the M (M + 1) / 2 distinct entries are selected once and mirrored
#pragma unroll
    for (int i = 0; i < M; ++i)
#pragma unroll
      for (int j = i; j < M; ++j) {
        const bool f = ((free_mask >> i) & 1u) && ((free_mask >> j) & 1u);
        const double h = f ? H[i * M + j] : (i == j ? 1.0 : 0.0);
        Hm[i * M + j] = h;
        Hm[j * M + i] = h;
      }
    return SymInverse<M>::run(Hm, Inv);
  }
  // same, with clamped rows/columns zeroed (the form K = -Hinv Q_ux wants)
  __device__ __forceinline__ static bool masked_inverse(const double *H, unsigned free_mask, double *Hinv) {
    const bool pd = masked_inverse_raw(H, free_mask, Hinv);
    zero_clamped(free_mask, Hinv);
    return pd;
  }
  __device__ __forceinline__ static void zero_clamped(unsigned free_mask, double *Hinv) {
#pragma unroll
    for (int i = 0; i < M; ++i)
#pragma unroll
      for (int j = 0; j < M; ++j) {
        const bool f = ((free_mask >> i) & 1u) && ((free_mask >> j) & 1u);
        Hinv[i * M + j] = f ? Hinv[i * M + j] : 0.0;
      }
  }

  // BoxQPSolver::solve.  x: in = warm start, out = solution.  On return free_mask / Hinv describe the free set of
  // the LAST factorisation (boxqp.cpp:89-111): Hinv is the raw masked inverse (use zero_clamped before forming K);
  // for QP_ALL_CLAMPED free_mask == 0 and Hinv is whatever was last computed (the caller uses K = 0, clddp_solver.cpp:163).
  // H x is carried between iterations: the product evaluated for the Armijo test of the accepted candidate IS the
  // one the next iteration's gradient needs (same arithmetic as recomputing it, boxqp.cpp:61 / :170-172).  On return
  // Hx = H x for the returned x.
  __device__ static int solve(const cddp_b200_options &o, const double *H, const double *g, const double *lo,
                              const double *hi, double *x, unsigned &free_mask, double *Hinv, double *Hx) {
    constexpr unsigned all = (1u << M) - 1u;
    int status = QP_MAX_ITER_EXCEEDED;
#pragma unroll
    for (int i = 0; i < M; ++i) x[i] = clamp_box(x[i], lo[i], hi[i]);
    auto matvec_value = [&](const double *z, double *Hz) {  // 0.5 z^T H z + g^T z  (boxqp.cpp:235-239)
      double a = 0.0, bb = 0.0;
#pragma unroll
      for (int i = 0; i < M; ++i) {
        double s = 0.0;
#pragma unroll
        for (int j = 0; j < M; ++j) s += H[i * M + j] * z[j];
        Hz[i] = s;
        a += z[i] * s;
        bb += g[i] * z[i];
      }
      return 0.5 * a + bb;
    };
    unsigned clamped = 0u;
    free_mask = all;
    double value_ = matvec_value(x, Hx);
    double old_value = __longlong_as_double(0x7ff0000000000000LL);
    const double gtol2 = o.qp_min_gradient_norm * o.qp_min_gradient_norm;
    for (int iter = 0; iter < o.qp_max_iterations; ++iter) {
      if (iter > 0 && fabs(old_value - value_) < o.qp_min_relative_improvement * fabs(old_value)) {
        status = QP_SUCCESS;
        break;
      }
      old_value = value_;
      double grad[M];
#pragma unroll
      for (int i = 0; i < M; ++i) grad[i] = g[i] + Hx[i];
      const unsigned old_clamped = clamped;
      clamped = 0u;
#pragma unroll
      for (int i = 0; i < M; ++i)
        if ((x[i] == lo[i] && grad[i] > 0.0) || (x[i] == hi[i] && grad[i] < 0.0)) clamped |= (1u << i);
      free_mask = all & ~clamped;
      if (clamped == all) {
        status = QP_ALL_CLAMPED;
        break;
      }
      if (iter == 0 || clamped != old_clamped) {
        if (!masked_inverse_raw(H, free_mask, Hinv)) {
          status = QP_HESSIAN_NOT_PD;
          break;
        }
      }
      double gn = 0.0;
#pragma unroll
      for (int i = 0; i < M; ++i)
        if ((free_mask >> i) & 1u) gn += grad[i] * grad[i];
      if (gn < gtol2) {
        status = QP_SUCCESS;
        break;
      }
      double rhs[M], xc[M];  // xc = x on the clamped entries, 0 elsewhere: g + sum over the clamped i of H[:, i] x_i (:128-146)
#pragma unroll                //  as M unpredicated FMAs per row (adding H * 0 changes nothing)
      for (int i = 0; i < M; ++i) xc[i] = ((clamped >> i) & 1u) ? x[i] : 0.0;
#pragma unroll
      for (int j = 0; j < M; ++j) {
        double s = g[j];
#pragma unroll
        for (int i = 0; i < M; ++i) s += H[j * M + i] * xc[i];
        rhs[j] = ((free_mask >> j) & 1u) ? s : 0.0;  // masked rhs: the identity rows of the raw inverse then contribute nothing
      }
      double search[M];
      double sdotg = 0.0;
#pragma unroll
      for (int i = 0; i < M; ++i) {
        double y = 0.0;
#pragma unroll
        for (int j = 0; j < M; ++j) y += Hinv[i * M + j] * rhs[j];
        search[i] = ((free_mask >> i) & 1u) ? (-y - x[i]) : 0.0;
        sdotg += search[i] * grad[i];
      }
      if (sdotg >= 0.0) {
        status = QP_NO_DESCENT;
        break;
      }
      double step = 1.0, vn = 0.0;
      bool ls_ok = false;
      double xn[M], Hxn[M];
      while (step > o.qp_min_step_size) {
#pragma unroll
        for (int i = 0; i < M; ++i) xn[i] = clamp_box(x[i] + step * search[i], lo[i], hi[i]);
        vn = matvec_value(xn, Hxn);
        if ((vn - value_) <= o.qp_armijo_constant * step * sdotg) {
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
      for (int i = 0; i < M; ++i) {
        x[i] = xn[i];
        Hx[i] = Hxn[i];
      }
      value_ = vn;
    }
    return status;
  }
};

}  // namespace cddp_b200
