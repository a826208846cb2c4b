// Control-space subproblem for m = 4 with FOUR LANES per trajectory (one row of the Hessian each): the QP warps of the
// warp-specialised sweep kernel (backward_fast.cu, SweepCfg::QPQ).  Eight trajectories share a warp, so the projected-
// Newton iterations run in lock step over 8 subproblems instead of the 28 of the one-lane-per-trajectory QP warp, and a
// lane carries ~170 instructions per iteration instead of ~430.
//
// Reference behaviour followed (astomodynamics/cddp-cpp @ f71fa80): BoxQPSolver::solve, src/cddp_core/boxqp.cpp:25-182
// (+ lineSearch :207-233, initializeX :184-205) — the same iteration structure, exit tests and status codes as
// SmallQP<4>::solve (boxqp_small.cuh), statement by statement; CLDDPSolver::backwardPass's PD test
// (clddp_solver.cpp:133-140, Sylvester's criterion as in SymPD<4>).
//
// Layout.  Lane i (0..3) of a quad owns row i: x_i, g_i, lo_i, hi_i, (H x)_i and row i of the masked inverse.  Vectors are
// exchanged with quad-wide shuffles; scalar sums (objective value, |grad_free|^2, search . grad) are butterfly sums
// over the quad — (t_i + t_{i^1}) + (t_{i^2} + t_{i^3}) — which are bit-identical in the four lanes (floating-point
// addition commutes), so every lane of a quad takes the same branch.  The masked inverse (clamped rows / columns replaced
// by identity, boxqp_small.cuh) is formed by ROTATED cofactor expansion: lane i relabels the indices j -> (j - i) mod 4
// and expands along ITS row, which is row 0 of the relabelled matrix — uniform code in the four lanes, 6 minors + 4
// cofactors + one reciprocal each, no lane computes another lane's row.  The leading principal minors Sylvester's
// criterion needs fall out of the same numbers: a_00 in lane 0, the {0,1} minor in lane 2 (rows 2,3 of its relabelling),
// the {0,1,2} minor in lane 3 (its C_00), the determinant everywhere.
//
// All lanes of the warp execute every instruction (finished quads are predicated off), so the shuffles use the full mask.
#pragma once
#include "boxqp.cuh"
#include "engine.h"

namespace cddp_b200 {

struct QuadQP {
  static constexpr unsigned kFull = 0xffffffffu;

  __device__ __forceinline__ static double quad_sum(double t) {
    t += __shfl_xor_sync(kFull, t, 1);
    t += __shfl_xor_sync(kFull, t, 2);
    return t;
  }
  __device__ __forceinline__ static double quad_max(double t) {
    t = max_ref(t, __shfl_xor_sync(kFull, t, 1));
    t = max_ref(t, __shfl_xor_sync(kFull, t, 2));
    return t;
  }
  // v of the quad's lanes in ORIGINAL index order
  __device__ __forceinline__ static void gather(double v, int qb, double *out) {
#pragma unroll
    for (int j = 0; j < 4; ++j) out[j] = __shfl_sync(kFull, v, qb + j);
  }
  // v of the quad's lanes in this lane's ROTATED order: out[b] = v of lane (i + b) mod 4
  __device__ __forceinline__ static void gather_rot(double v, int qb, int i, double *out) {
#pragma unroll
    for (int b = 0; b < 4; ++b) out[b] = __shfl_sync(kFull, v, qb + ((i + b) & 3));
  }
  __device__ __forceinline__ static unsigned quad_bits(bool p, int qb) { return (__ballot_sync(kFull, p) >> qb) & 15u; }

  // Row i of the inverse of the masked matrix, in rotated order (inv[b] <-> column (i + b) mod 4), and Sylvester's
  // criterion on the masked matrix (uniform over the quad).  Hr = full symmetric matrix in this lane's rotated indices,
  // fr = free mask rotated the same way (bit a <-> index (i + a) mod 4).
  // Quu: the trajectory's UNREGULARISED Q_uu in shared memory (row-major 4 x 4, upper triangle read — the same entries
  // the one-lane QP warp reads); it is re-read at every factorisation instead of living in 20 registers across the
  // iteration (the QP warps run at 72 registers per thread, see SweepCfg).
  __device__ __forceinline__ static bool inverse_row(const volatile double *Quu, double reg, unsigned fr, int i, int qb, double *inv) {
    double M[4][4];
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
      for (int b = a; b < 4; ++b) {
        const int ia = (i + a) & 3, ib = (i + b) & 3;
        const double h = Quu[ia < ib ? ia * 4 + ib : ib * 4 + ia] + (a == b ? reg : 0.0);  // Q_uu_reg (:130-131)
        const bool f = ((fr >> a) & 1u) && ((fr >> b) & 1u);
        M[a][b] = M[b][a] = f ? h : (a == b ? 1.0 : 0.0);
      }
    const double m01 = M[2][0] * M[3][1] - M[2][1] * M[3][0];
    const double m02 = M[2][0] * M[3][2] - M[2][2] * M[3][0];
    const double m03 = M[2][0] * M[3][3] - M[2][3] * M[3][0];
    const double m12 = M[2][1] * M[3][2] - M[2][2] * M[3][1];
    const double m13 = M[2][1] * M[3][3] - M[2][3] * M[3][1];
    const double m23 = M[2][2] * M[3][3] - M[2][3] * M[3][2];
    const double C0 = M[1][1] * m23 - M[1][2] * m13 + M[1][3] * m12;
    const double C1 = -(M[1][0] * m23 - M[1][2] * m03 + M[1][3] * m02);
    const double C2 = M[1][0] * m13 - M[1][1] * m03 + M[1][3] * m01;
    const double C3 = -(M[1][0] * m12 - M[1][1] * m02 + M[1][2] * m01);
    const double det = (M[0][0] * C0 + M[0][1] * C1) + (M[0][2] * C2 + M[0][3] * C3);
    const double id = 1.0 / det;
    inv[0] = C0 * id;
    inv[1] = C1 * id;
    inv[2] = C2 * id;
    inv[3] = C3 * id;
    // leading principal minors of the ORIGINAL ordering: a_00 (lane 0), {0,1} (lane 2: its rows/columns 2,3),
    // {0,1,2} (lane 3: its rows/columns 1,2,3), and the determinant (every lane, its own expansion)
    const double lead = i == 0 ? M[0][0] : (i == 2 ? m23 : (i == 3 ? C0 : 1.0));
    return quad_bits((lead > 0.0) && (det > 0.0), qb) == 15u;
  }

  // One trajectory-step of the control-space subproblem, executed by all 32 lanes (8 quads).
  //   run   : this quad has a subproblem this round (uniform over the quad)
  //   Quu, reg : unregularised Q_uu in shared memory + regularisation; Ho: row i of Q_uu_reg in original order
  //   g, lo, hi, x(in: warm start; out: k_i) : this lane's entries
  //   Hx    : out, (Q_uu_reg k)_i
  //   fm    : out, free mask of the LAST factorisation (original order); inv: row i of its masked inverse (rotated, raw)
  // Returns the BoxQP status (boxqp.cuh) of the quad, or QP_HESSIAN_NOT_PD if Q_uu_reg itself fails the PD test.
  __device__ static int solve_box(const cddp_b200_options &o, bool run, const volatile double *Quu, double reg,
                                  const double (&Ho)[4], double g, double lo, double hi, double &x, double &Hx, unsigned &fm,
                                  double *inv, int i, int qb) {
    // PD test of the full matrix (clddp_solver.cpp:133-140); its inverse row is reused by iteration 0 when nothing is clamped
    bool have_full = false;
    const bool pd_full = inverse_row(Quu, reg, 15u, i, qb, inv);
    have_full = true;
    bool act = run && pd_full;
    int status = pd_full ? QP_MAX_ITER_EXCEEDED : QP_HESSIAN_NOT_PD;
    x = clamp_box(x, lo, hi);  // initializeX (:184-205)
    double v4[4];
    auto row_dot = [&](const double *z) {  // (H z)_i, summed in original index order
      double s = 0.0;
#pragma unroll
      for (int j = 0; j < 4; ++j) s += Ho[j] * z[j];
      return s;
    };
    gather(x, qb, v4);
    Hx = row_dot(v4);
    double value = quad_sum(0.5 * (x * Hx) + g * x);  // 0.5 x^T H x + g^T x (:235-239)
    double old_value = __longlong_as_double(0x7ff0000000000000LL);
    unsigned clamped = 0u;
    fm = 15u;
    const double gtol2 = o.qp_min_gradient_norm * o.qp_min_gradient_norm;
    for (int iter = 0; iter < o.qp_max_iterations; ++iter) {
      if (!__any_sync(kFull, act)) break;
      if (act && iter > 0 && fabs(old_value - value) < o.qp_min_relative_improvement * fabs(old_value)) {  // (:52-57)
        status = QP_SUCCESS;
        act = false;
      }
      old_value = value;
      const double grad = g + Hx;  // (:58-61)
      const unsigned newcl = quad_bits((x == lo && grad > 0.0) || (x == hi && grad < 0.0), qb);  // (:67-72)
      const bool changed = iter == 0 || newcl != clamped;
      if (act) {
        clamped = newcl;
        fm = 15u & ~clamped;
        if (clamped == 15u) {  // (:73-79)
          status = QP_ALL_CLAMPED;
          act = false;
        }
      }
      const bool need = act && changed && !(have_full && clamped == 0u);  // (:89-111)
      if (__any_sync(kFull, need)) {
        double t[4];
        const unsigned fr = ((fm | (fm << 4)) >> i) & 15u;
        const bool pd = inverse_row(Quu, reg, fr, i, qb, t);
        if (need) {
#pragma unroll
          for (int b = 0; b < 4; ++b) inv[b] = t[b];
          if (!pd) {
            status = QP_HESSIAN_NOT_PD;
            act = false;
          }
        }
      }
      if (act && changed) have_full = clamped == 0u;
      const bool fi = (fm >> i) & 1u;
      const double gn = quad_sum(fi ? grad * grad : 0.0);  // (:114-125)
      if (act && gn < gtol2) {
        status = QP_SUCCESS;
        act = false;
      }
      gather(fi ? 0.0 : x, qb, v4);  // g + sum over the clamped i of H[:, i] x_i (:128-146)
      double rhs = g;
#pragma unroll
      for (int j = 0; j < 4; ++j) rhs += Ho[j] * v4[j];
      gather_rot(fi ? rhs : 0.0, qb, i, v4);  // masked rhs: the identity rows of the raw inverse then contribute nothing
      double y = 0.0;
#pragma unroll
      for (int b = 0; b < 4; ++b) y += inv[b] * v4[b];
      const double search = fi ? (-y - x) : 0.0;  // (:147-152)
      const double sdotg = quad_sum(search * grad);
      if (act && sdotg >= 0.0) {  // (:155-159)
        status = QP_NO_DESCENT;
        act = false;
      }
      double step = 1.0, xn = x, Hxn = Hx, vn = value;  // lineSearch (:207-233)
      bool ls = act, found = false;
      while (__any_sync(kFull, ls)) {
        const double xt = clamp_box(x + step * search, lo, hi);
        gather(xt, qb, v4);
        const double Hxt = row_dot(v4);
        const double vt = quad_sum(0.5 * (xt * Hxt) + g * xt);
        if (ls) {
          if ((vt - value) <= o.qp_armijo_constant * step * sdotg) {
            xn = xt;
            Hxn = Hxt;
            vn = vt;
            found = true;
            ls = false;
          } else {
            step *= o.qp_step_decrease_factor;
            if (!(step > o.qp_min_step_size)) ls = false;
          }
        }
      }
      if (act && !found) {  // (:162-167)
        status = QP_MAX_LS_EXCEEDED;
        act = false;
      }
      if (act) {  // (:170-172)
        x = xn;
        Hx = Hxn;
        value = vn;
      }
    }
    return status;
  }
};

}  // namespace cddp_b200
