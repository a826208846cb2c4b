// Device side of the user-model plugin (compiled ONLY by NVRTC, together with the user's CUDA source):
// the reference's DynamicalSystem extension point (include/cddp-cpp/cddp_core/dynamical_system.hpp:33-152 —
// getContinuousDynamics, getContinuousDynamicsAutodiff, getStateJacobian/getControlJacobian) restated for code that
// has to run inside the sm_100a kernels.
//
// The user source must define, at global scope,
//
//     template <class T>
//     __device__ void cddp_user_dynamics(const T *x, const T *u, const double *p, T *xdot);
//
// (the counterpart of getContinuousDynamics / getContinuousDynamicsAutodiff: one text, instantiated with T = double for
// the rollout and with T = cddp_b200::Dual for the Jacobians, exactly how the reference obtains Jacobians by forward-mode
// autodiff of the *Autodiff variant, dynamical_system.cpp:102-133), where p = cddp_b200_problem::model_params.  It may
// additionally define CDDP_USER_HAS_JACOBIAN and
//
//     __device__ void cddp_user_jacobian(const double *x, const double *u, const double *p, double *Fx, double *Fu);
//
// (row-major Fx [n][n], Fu [n][m]; the counterpart of overriding getStateJacobian / getControlJacobian).
#pragma once
#include "engine.h"

#ifndef CDDP_USER_NS
#error "CDDP_USER_NS / CDDP_USER_NC must be defined by the host before this header is compiled"
#endif

namespace cddp_b200 {
namespace ad {  // its own namespace: the math overloads below are found by argument-dependent lookup only, so that they never
                // shadow ::sqrt / ::pow / ... for plain doubles inside the engine's kernels

// forward-mode dual number carrying CDDP_DUAL_WIDTH tangent directions at once (stands in for autodiff::dual; first
// derivatives only: use_ilqr = true).  One direction per pass — what autodiff::jacobian does — re-evaluates every
// sin/cos/exp of the model n+m times; W directions per pass evaluate them ceil((n+m)/W) times and pay only W
// multiply-adds per operation for the tangents.  W = 3 keeps the live duals of a 14-state / 7-control model in registers.
#ifndef CDDP_DUAL_WIDTH
#define CDDP_DUAL_WIDTH 3
#endif
struct Dual {
  static constexpr int W = CDDP_DUAL_WIDTH;
  double v, d[W];
  __device__ Dual() : v(0.0) {
#pragma unroll
    for (int k = 0; k < W; ++k) d[k] = 0.0;
  }
  __device__ Dual(double v_) : v(v_) {
#pragma unroll
    for (int k = 0; k < W; ++k) d[k] = 0.0;
  }
};
// f(a) with derivative fp: value fv, tangent fp * a.d
__device__ __forceinline__ Dual chain(const Dual &a, double fv, double fp) {
  Dual r;
  r.v = fv;
#pragma unroll
  for (int k = 0; k < Dual::W; ++k) r.d[k] = fp * a.d[k];
  return r;
}
__device__ inline Dual operator+(const Dual &a, const Dual &b) {
  Dual r;
  r.v = a.v + b.v;
#pragma unroll
  for (int k = 0; k < Dual::W; ++k) r.d[k] = a.d[k] + b.d[k];
  return r;
}
__device__ inline Dual operator-(const Dual &a, const Dual &b) {
  Dual r;
  r.v = a.v - b.v;
#pragma unroll
  for (int k = 0; k < Dual::W; ++k) r.d[k] = a.d[k] - b.d[k];
  return r;
}
__device__ inline Dual operator-(const Dual &a) { return chain(a, -a.v, -1.0); }
__device__ inline Dual operator+(const Dual &a) { return a; }
__device__ inline Dual operator*(const Dual &a, const Dual &b) {
  Dual r;
  r.v = a.v * b.v;
#pragma unroll
  for (int k = 0; k < Dual::W; ++k) r.d[k] = a.d[k] * b.v + a.v * b.d[k];
  return r;
}
__device__ inline Dual operator/(const Dual &a, const Dual &b) {
  Dual r;
  const double ib = 1.0 / b.v, q = a.v * ib;
  r.v = a.v / b.v;
#pragma unroll
  for (int k = 0; k < Dual::W; ++k) r.d[k] = (a.d[k] - q * b.d[k]) * ib;
  return r;
}
// mixed double / Dual arithmetic without promoting the scalar to a dual with W zero tangents
__device__ inline Dual operator+(const Dual &a, double b) { Dual r = a; r.v += b; return r; }
__device__ inline Dual operator+(double a, const Dual &b) { Dual r = b; r.v += a; return r; }
__device__ inline Dual operator-(const Dual &a, double b) { Dual r = a; r.v -= b; return r; }
__device__ inline Dual operator-(double a, const Dual &b) { return chain(b, a - b.v, -1.0); }
__device__ inline Dual operator*(const Dual &a, double b) { return chain(a, a.v * b, b); }
__device__ inline Dual operator*(double a, const Dual &b) { return chain(b, a * b.v, a); }
__device__ inline Dual operator/(const Dual &a, double b) {
  Dual r = chain(a, 0.0, 1.0 / b);
  r.v = a.v / b;
  return r;
}
__device__ inline Dual operator/(double a, const Dual &b) {
  const double q = a / b.v;
  return chain(b, q, -q / b.v);
}
__device__ inline Dual &operator+=(Dual &a, const Dual &b) { a = a + b; return a; }
__device__ inline Dual &operator-=(Dual &a, const Dual &b) { a = a - b; return a; }
__device__ inline Dual &operator*=(Dual &a, const Dual &b) { a = a * b; return a; }
__device__ inline Dual &operator/=(Dual &a, const Dual &b) { a = a / b; return a; }
__device__ inline Dual &operator+=(Dual &a, double b) { a.v += b; return a; }
__device__ inline Dual &operator-=(Dual &a, double b) { a.v -= b; return a; }
__device__ inline Dual &operator*=(Dual &a, double b) { a = a * b; return a; }
__device__ inline Dual &operator/=(Dual &a, double b) { a = a / b; return a; }
__device__ inline bool operator<(const Dual &a, const Dual &b) { return a.v < b.v; }
__device__ inline bool operator>(const Dual &a, const Dual &b) { return a.v > b.v; }
__device__ inline bool operator<=(const Dual &a, const Dual &b) { return a.v <= b.v; }
__device__ inline bool operator>=(const Dual &a, const Dual &b) { return a.v >= b.v; }
__device__ inline bool operator<(const Dual &a, double b) { return a.v < b; }
__device__ inline bool operator>(const Dual &a, double b) { return a.v > b; }
__device__ inline bool operator<=(const Dual &a, double b) { return a.v <= b; }
__device__ inline bool operator>=(const Dual &a, double b) { return a.v >= b; }
__device__ inline bool operator<(double a, const Dual &b) { return a < b.v; }
__device__ inline bool operator>(double a, const Dual &b) { return a > b.v; }
__device__ inline bool operator<=(double a, const Dual &b) { return a <= b.v; }
__device__ inline bool operator>=(double a, const Dual &b) { return a >= b.v; }
__device__ inline Dual sin(const Dual &a) {
  double sn, cs;
  ::sincos(a.v, &sn, &cs);
  return chain(a, sn, cs);
}
__device__ inline Dual cos(const Dual &a) {
  double sn, cs;
  ::sincos(a.v, &sn, &cs);
  return chain(a, cs, -sn);
}
__device__ inline Dual tan(const Dual &a) {
  const double t = ::tan(a.v);
  return chain(a, t, 1.0 + t * t);
}
__device__ inline Dual sqrt(const Dual &a) {
  const double r = ::sqrt(a.v);
  return chain(a, r, 1.0 / (2.0 * r));
}
__device__ inline Dual exp(const Dual &a) {
  const double e = ::exp(a.v);
  return chain(a, e, e);
}
__device__ inline Dual log(const Dual &a) { return chain(a, ::log(a.v), 1.0 / a.v); }
__device__ inline Dual tanh(const Dual &a) {
  const double t = ::tanh(a.v);
  return chain(a, t, 1.0 - t * t);
}
__device__ inline Dual atan(const Dual &a) { return chain(a, ::atan(a.v), 1.0 / (1.0 + a.v * a.v)); }
__device__ inline Dual asin(const Dual &a) { return chain(a, ::asin(a.v), 1.0 / ::sqrt(1.0 - a.v * a.v)); }
__device__ inline Dual acos(const Dual &a) { return chain(a, ::acos(a.v), -1.0 / ::sqrt(1.0 - a.v * a.v)); }
__device__ inline Dual atan2(const Dual &y, const Dual &x) {
  const double ir2 = 1.0 / (x.v * x.v + y.v * y.v);
  Dual r;
  r.v = ::atan2(y.v, x.v);
#pragma unroll
  for (int k = 0; k < Dual::W; ++k) r.d[k] = (x.v * y.d[k] - y.v * x.d[k]) * ir2;
  return r;
}
__device__ inline Dual pow(const Dual &a, double e) { return chain(a, ::pow(a.v, e), e * ::pow(a.v, e - 1.0)); }
__device__ inline Dual fabs(const Dual &a) { return a.v < 0.0 ? -a : a; }

}  // namespace ad
using ad::Dual;
}  // namespace cddp_b200

using cddp_b200::ad::Dual;

template <class T>
__device__ void cddp_user_dynamics(const T *x, const T *u, const double *p, T *xdot);
#ifdef CDDP_USER_HAS_JACOBIAN
__device__ void cddp_user_jacobian(const double *x, const double *u, const double *p, double *Fx, double *Fu);
#endif

namespace cddp_b200 {

template <>
struct Model<CDDP_B200_MODEL_USER> {
  static constexpr int NS = CDDP_USER_NS, NC = CDDP_USER_NC;
  __device__ __forceinline__ static void f(const ModelParams &P, const double *x, const double *u, double *xd) {
    cddp_user_dynamics<double>(x, u, P.p, xd);
  }
  // continuous-time Jacobians Fx [NS][NS], Fu [NS][NC], fully written
  __device__ __forceinline__ static void jac(const ModelParams &P, const double *x, const double *u, double *Fx, double *Fu) {
#ifdef CDDP_USER_HAS_JACOBIAN
    cddp_user_jacobian(x, u, P.p, Fx, Fu);
#else
    constexpr int W = Dual::W;
#pragma unroll 1
    for (int base = 0; base < NS + NC; base += W) {  // W tangent directions per pass (autodiff::jacobian, dynamical_system.cpp:102-133)
      Dual xs[NS], us[NC], xd[NS];
#pragma unroll
      for (int i = 0; i < NS; ++i) {
        xs[i].v = x[i];
#pragma unroll
        for (int k = 0; k < W; ++k) xs[i].d[k] = (i == base + k) ? 1.0 : 0.0;
      }
#pragma unroll
      for (int i = 0; i < NC; ++i) {
        us[i].v = u[i];
#pragma unroll
        for (int k = 0; k < W; ++k) us[i].d[k] = (NS + i == base + k) ? 1.0 : 0.0;
      }
      cddp_user_dynamics<Dual>(xs, us, P.p, xd);
#pragma unroll
      for (int k = 0; k < W; ++k) {
        const int dir = base + k;
        if (dir < NS) {
          for (int i = 0; i < NS; ++i) Fx[i * NS + dir] = xd[i].d[k];
        } else if (dir < NS + NC) {
          for (int i = 0; i < NS; ++i) Fu[i * NC + (dir - NS)] = xd[i].d[k];
        }
      }
    }
#endif
  }
};

}  // namespace cddp_b200
