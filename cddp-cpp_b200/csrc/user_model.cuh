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

// forward-mode dual number, one tangent direction (stands in for autodiff::dual; first derivatives only: use_ilqr = true)
struct Dual {
  double v, d;
  __device__ Dual() : v(0.0), d(0.0) {}
  __device__ Dual(double v_) : v(v_), d(0.0) {}
  __device__ Dual(double v_, double d_) : v(v_), d(d_) {}
};
__device__ inline Dual operator+(Dual a, Dual b) { return Dual(a.v + b.v, a.d + b.d); }
__device__ inline Dual operator-(Dual a, Dual b) { return Dual(a.v - b.v, a.d - b.d); }
__device__ inline Dual operator-(Dual a) { return Dual(-a.v, -a.d); }
__device__ inline Dual operator+(Dual a) { return a; }
__device__ inline Dual operator*(Dual a, Dual b) { return Dual(a.v * b.v, a.d * b.v + a.v * b.d); }
__device__ inline Dual operator/(Dual a, Dual b) {
  const double q = a.v / b.v;
  return Dual(q, (a.d - q * b.d) / b.v);
}
__device__ inline Dual &operator+=(Dual &a, Dual b) { a = a + b; return a; }
__device__ inline Dual &operator-=(Dual &a, Dual b) { a = a - b; return a; }
__device__ inline Dual &operator*=(Dual &a, Dual b) { a = a * b; return a; }
__device__ inline Dual &operator/=(Dual &a, Dual b) { a = a / b; return a; }
__device__ inline bool operator<(Dual a, Dual b) { return a.v < b.v; }
__device__ inline bool operator>(Dual a, Dual b) { return a.v > b.v; }
__device__ inline bool operator<=(Dual a, Dual b) { return a.v <= b.v; }
__device__ inline bool operator>=(Dual a, Dual b) { return a.v >= b.v; }
__device__ inline Dual sin(Dual a) { return Dual(::sin(a.v), ::cos(a.v) * a.d); }
__device__ inline Dual cos(Dual a) { return Dual(::cos(a.v), -::sin(a.v) * a.d); }
__device__ inline Dual tan(Dual a) {
  const double t = ::tan(a.v);
  return Dual(t, (1.0 + t * t) * a.d);
}
__device__ inline Dual sqrt(Dual a) {
  const double r = ::sqrt(a.v);
  return Dual(r, a.d / (2.0 * r));
}
__device__ inline Dual exp(Dual a) {
  const double e = ::exp(a.v);
  return Dual(e, e * a.d);
}
__device__ inline Dual log(Dual a) { return Dual(::log(a.v), a.d / a.v); }
__device__ inline Dual tanh(Dual a) {
  const double t = ::tanh(a.v);
  return Dual(t, (1.0 - t * t) * a.d);
}
__device__ inline Dual atan(Dual a) { return Dual(::atan(a.v), a.d / (1.0 + a.v * a.v)); }
__device__ inline Dual asin(Dual a) { return Dual(::asin(a.v), a.d / ::sqrt(1.0 - a.v * a.v)); }
__device__ inline Dual acos(Dual a) { return Dual(::acos(a.v), -a.d / ::sqrt(1.0 - a.v * a.v)); }
__device__ inline Dual atan2(Dual y, Dual x) {
  const double r2 = x.v * x.v + y.v * y.v;
  return Dual(::atan2(y.v, x.v), (x.v * y.d - y.v * x.d) / r2);
}
__device__ inline Dual pow(Dual a, double e) { return Dual(::pow(a.v, e), e * ::pow(a.v, e - 1.0) * a.d); }
__device__ inline Dual fabs(Dual a) { return a.v < 0.0 ? -a : a; }

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
    Dual xs[NS], us[NC], xd[NS];
#pragma unroll
    for (int i = 0; i < NS; ++i) xs[i] = Dual(x[i]);
#pragma unroll
    for (int i = 0; i < NC; ++i) us[i] = Dual(u[i]);
    for (int dir = 0; dir < NS + NC; ++dir) {  // one tangent direction per pass (autodiff::jacobian, dynamical_system.cpp:102-133)
      if (dir < NS) xs[dir].d = 1.0;
      else us[dir - NS].d = 1.0;
      cddp_user_dynamics<Dual>(xs, us, P.p, xd);
      if (dir < NS) {
        xs[dir].d = 0.0;
        for (int i = 0; i < NS; ++i) Fx[i * NS + dir] = xd[i].d;
      } else {
        us[dir - NS].d = 0.0;
        for (int i = 0; i < NS; ++i) Fu[i * NC + (dir - NS)] = xd[i].d;
      }
    }
#endif
  }
};

}  // namespace cddp_b200
