// Device-resident dynamics models: continuous dynamics f(x,u), closed-form continuous Jacobians
// (Fx, Fu) and the reference's fixed-step integrators.
//
// Reference behaviour followed (file:line in astomodynamics/cddp-cpp @ f71fa80):
//   integrators            src/cddp_core/dynamical_system.cpp:28-83
//   Pendulum               src/dynamics_model/pendulum.cpp:29-66
//   CartPole               src/dynamics_model/cartpole.cpp:38-103  (Jacobian = autodiff of the
//                          dual version, which includes the damping term, :90)
//   Unicycle               src/dynamics_model/unicycle.cpp:28-66
//   Quadrotor              src/dynamics_model/quadrotor.cpp:33-140 (Jacobian = autodiff incl. the
//                          derivative of the in-dynamics quaternion normalisation, :45-55)
//   LTISystem              src/dynamics_model/lti_system.cpp:71-92
// The reference obtains cartpole/quadrotor Jacobians by forward-mode AD; here they are derived by
// hand (no AD on the device) and parity-checked against the oracle's dual-number Jacobians.
#pragma once
#ifndef __CUDACC_RTC__
#include <cuda_runtime.h>
#endif

#include "../../include/cddp_b200.h"

namespace cddp_b200 {

// p[0..15] = cddp_b200_problem.model_params; p[16..] = derived constants filled by prepare_model_params()
// (quadrotor: p[16..24] = inertia^-1 row-major, p[25] = 1/mass — the reference recomputes inertia_.inverse() in
// every dynamics call, quadrotor.cpp:92).
struct ModelParams {
  double p[32];
  const double *lti_A;  // device, [n][n]
  const double *lti_B;  // device, [n][m]
  int n, m;
};

template <int MODEL>
struct Model;

// ---------------------------------------------------------------- Pendulum (theta, theta_dot)
template <>
struct Model<CDDP_B200_MODEL_PENDULUM> {
  static constexpr int NS = 2, NC = 1;
  __device__ __forceinline__ static void f(const ModelParams &P, const double *x, const double *u, double *xd) {
    const double length = P.p[0], mass = P.p[1], damping = P.p[2], gravity = 9.81;
    const double inertia = mass * length * length;
    xd[0] = x[1];
    xd[1] = (u[0] - damping * x[1] + mass * gravity * length * sin(x[0])) / inertia;
  }
  __device__ __forceinline__ static void jac(const ModelParams &P, const double *x, const double *, double *Fx,
                                             double *Fu) {
    const double length = P.p[0], mass = P.p[1], damping = P.p[2], gravity = 9.81;
    Fx[0] = 0.0;
    Fx[1] = 1.0;
    Fx[2] = (gravity / length) * cos(x[0]);
    Fx[3] = -damping / (mass * length * length);
    Fu[0] = 0.0;
    Fu[1] = 1.0 / (mass * length * length);
  }
};

// ---------------------------------------------------------------- CartPole (x, theta, x_dot, theta_dot)
template <>
struct Model<CDDP_B200_MODEL_CARTPOLE> {
  static constexpr int NS = 4, NC = 1;
  __device__ __forceinline__ static void f(const ModelParams &P, const double *x, const double *u, double *xd) {
    const double mc = P.p[0], mp = P.p[1], l = P.p[2], g = P.p[3];
    double s, c;
    sincos(x[1], &s, &c);
    const double w = x[3], F = u[0];
    const double den = mc + mp * s * s;
    xd[0] = x[2];
    xd[1] = w;
    xd[2] = (F + mp * s * (l * w * w + g * c)) / den;
    xd[3] = (-F * c - mp * l * w * w * c * s - (mc + mp) * g * s) / (l * den);  // no damping in f (cartpole.cpp:60)
  }
  __device__ __forceinline__ static void jac(const ModelParams &P, const double *x, const double *u, double *Fx,
                                             double *Fu) {
    const double mc = P.p[0], mp = P.p[1], l = P.p[2], g = P.p[3], d = P.p[4];
    double s, c;
    sincos(x[1], &s, &c);
    const double w = x[3], F = u[0];
    const double den = mc + mp * s * s;
    const double dden = 2.0 * mp * s * c;
    const double num3 = F + mp * s * (l * w * w + g * c);
    const double dnum3_dth = mp * c * (l * w * w + g * c) - mp * g * s * s;
    const double dnum3_dw = 2.0 * mp * s * l * w;
    const double num4 = -F * c - mp * l * w * w * c * s - (mc + mp) * g * s - d * w;  // damping enters (cartpole.cpp:90)
    const double dnum4_dth = F * s - mp * l * w * w * (c * c - s * s) - (mc + mp) * g * c;
    const double dnum4_dw = -2.0 * mp * l * w * c * s - d;
#pragma unroll
    for (int i = 0; i < 16; ++i) Fx[i] = 0.0;
    Fx[0 * 4 + 2] = 1.0;
    Fx[1 * 4 + 3] = 1.0;
    Fx[2 * 4 + 1] = (dnum3_dth * den - num3 * dden) / (den * den);
    Fx[2 * 4 + 3] = dnum3_dw / den;
    Fx[3 * 4 + 1] = (dnum4_dth * den - num4 * dden) / (l * den * den);
    Fx[3 * 4 + 3] = dnum4_dw / (l * den);
    Fu[0] = 0.0;
    Fu[1] = 0.0;
    Fu[2] = 1.0 / den;
    Fu[3] = -c / (l * den);
  }
};

// ---------------------------------------------------------------- Unicycle (x, y, theta; v, omega)
template <>
struct Model<CDDP_B200_MODEL_UNICYCLE> {
  static constexpr int NS = 3, NC = 2;
  __device__ __forceinline__ static void f(const ModelParams &, const double *x, const double *u, double *xd) {
    double s, c;
    sincos(x[2], &s, &c);
    xd[0] = u[0] * c;
    xd[1] = u[0] * s;
    xd[2] = u[1];
  }
  __device__ __forceinline__ static void jac(const ModelParams &, const double *x, const double *u, double *Fx,
                                             double *Fu) {
    double s, c;
    sincos(x[2], &s, &c);
#pragma unroll
    for (int i = 0; i < 9; ++i) Fx[i] = 0.0;
    Fx[0 * 3 + 2] = -u[0] * s;
    Fx[1 * 3 + 2] = u[0] * c;
    Fu[0] = c;
    Fu[1] = 0.0;
    Fu[2] = s;
    Fu[3] = 0.0;
    Fu[4] = 0.0;
    Fu[5] = 1.0;
  }
};

// ---------------------------------------------------------------- Quadrotor
// state [p(3), q_wxyz(4), v(3), omega(3)], control f1..f4
template <>
struct Model<CDDP_B200_MODEL_QUADROTOR> {
  static constexpr int NS = 13, NC = 4;

  struct Inertia {
    double I[9], Iinv[9];
  };
  __host__ __device__ __forceinline__ static void inertia(const ModelParams &P, Inertia &J) {
#pragma unroll
    for (int i = 0; i < 9; ++i) {
      J.I[i] = P.p[1 + i];
      J.Iinv[i] = P.p[16 + i];
    }
  }
#ifndef __CUDACC_RTC__
  // host-side: 3x3 inverse by cofactors, stored in p[16..24]; p[25] = 1/mass
  __host__ static void prepare(ModelParams &P) {
    const double *I = P.p + 1;
    const double c00 = I[4] * I[8] - I[5] * I[7], c01 = I[2] * I[7] - I[1] * I[8], c02 = I[1] * I[5] - I[2] * I[4];
    const double c10 = I[5] * I[6] - I[3] * I[8], c11 = I[0] * I[8] - I[2] * I[6], c12 = I[2] * I[3] - I[0] * I[5];
    const double c20 = I[3] * I[7] - I[4] * I[6], c21 = I[1] * I[6] - I[0] * I[7], c22 = I[0] * I[4] - I[1] * I[3];
    const double id = 1.0 / (I[0] * c00 + I[1] * c10 + I[2] * c20);
    double *o = P.p + 16;
    o[0] = c00 * id; o[1] = c01 * id; o[2] = c02 * id;
    o[3] = c10 * id; o[4] = c11 * id; o[5] = c12 * id;
    o[6] = c20 * id; o[7] = c21 * id; o[8] = c22 * id;
    P.p[25] = 1.0 / P.p[0];
  }
#endif

  __device__ __forceinline__ static void f(const ModelParams &P, const double *x, const double *u, double *xd) {
    const double L = P.p[10], gravity = 9.81;
    Inertia J;
    inertia(P, J);
    xd[0] = x[7];
    xd[1] = x[8];
    xd[2] = x[9];
    double qw = x[3], qx = x[4], qy = x[5], qz = x[6];
    const double n2 = qw * qw + qx * qx + qy * qy + qz * qz;
    if (n2 > 1e-12) {  // ||q|| > 1e-6 (quadrotor.cpp:45-55); q/||q|| as q * rsqrt(|q|^2): one reciprocal-sqrt, no divisions
      const double inv = rsqrt(n2);
      qw *= inv; qx *= inv; qy *= inv; qz *= inv;
    } else {
      qw = 1.0; qx = 0.0; qy = 0.0; qz = 0.0;
    }
    const double wx = x[10], wy = x[11], wz = x[12];
    xd[3] = -0.5 * (qx * wx + qy * wy + qz * wz);
    xd[4] = 0.5 * (qw * wx + qy * wz - qz * wy);
    xd[5] = 0.5 * (qw * wy - qx * wz + qz * wx);
    xd[6] = 0.5 * (qw * wz + qx * wy - qy * wx);
    const double thrust = u[0] + u[1] + u[2] + u[3];
    const double tx = L * (u[0] - u[2]), ty = L * (u[1] - u[3]), tz = 0.1 * (u[0] - u[1] + u[2] - u[3]);
    const double im = P.p[25];
    xd[7] = im * (2.0 * (qx * qz + qy * qw) * thrust);
    xd[8] = im * (2.0 * (qy * qz - qx * qw) * thrust);
    xd[9] = im * ((1.0 - 2.0 * (qx * qx + qy * qy)) * thrust) - gravity;
    const double Iw0 = J.I[0] * wx + J.I[1] * wy + J.I[2] * wz;
    const double Iw1 = J.I[3] * wx + J.I[4] * wy + J.I[5] * wz;
    const double Iw2 = J.I[6] * wx + J.I[7] * wy + J.I[8] * wz;
    const double r0 = tx - (wy * Iw2 - wz * Iw1), r1 = ty - (wz * Iw0 - wx * Iw2), r2 = tz - (wx * Iw1 - wy * Iw0);
    xd[10] = J.Iinv[0] * r0 + J.Iinv[1] * r1 + J.Iinv[2] * r2;
    xd[11] = J.Iinv[3] * r0 + J.Iinv[4] * r1 + J.Iinv[5] * r2;
    xd[12] = J.Iinv[6] * r0 + J.Iinv[7] * r1 + J.Iinv[8] * r2;
  }

  // Fx [13][13], Fu [13][4], row-major, fully written.
  __device__ __forceinline__ static void jac(const ModelParams &P, const double *x, const double *u, double *Fx,
                                             double *Fu) {
    const double L = P.p[10];
    Inertia J;
    inertia(P, J);
#pragma unroll
    for (int i = 0; i < 169; ++i) Fx[i] = 0.0;
#pragma unroll
    for (int i = 0; i < 52; ++i) Fu[i] = 0.0;
    // d p_dot / d v
    Fx[0 * 13 + 7] = 1.0;
    Fx[1 * 13 + 8] = 1.0;
    Fx[2 * 13 + 9] = 1.0;
    double qw = x[3], qx = x[4], qy = x[5], qz = x[6];
    // 1/|q| as rsqrt(|q|^2) and products instead of a square root and six divisions (each an FP64 division is a
    // ~25-instruction dependent sequence; the Jacobian is a latency chain wherever it runs), as f() does
    const double n2 = qw * qw + qx * qx + qy * qy + qz * qz;
    const bool reg = n2 > 1e-12;  // |q| > 1e-6 (quadrotor.cpp:45-55)
    double inv_norm = 0.0;
    if (reg) {
      inv_norm = rsqrt(n2);
      qw *= inv_norm; qx *= inv_norm; qy *= inv_norm; qz *= inv_norm;
    } else {
      qw = 1.0; qx = 0.0; qy = 0.0; qz = 0.0;
    }
    const double wx = x[10], wy = x[11], wz = x[12];
    const double thrust = u[0] + u[1] + u[2] + u[3];
    const double tm = thrust * P.p[25];  // p[25] = 1 / mass (prepare())
    // G = d(q_dot, v_dot)/d(qn): 7 rows x 4 cols
    double G[7][4] = {
        {0.0, -0.5 * wx, -0.5 * wy, -0.5 * wz},
        {0.5 * wx, 0.0, 0.5 * wz, -0.5 * wy},
        {0.5 * wy, -0.5 * wz, 0.0, 0.5 * wx},
        {0.5 * wz, 0.5 * wy, -0.5 * wx, 0.0},
        {tm * 2.0 * qy, tm * 2.0 * qz, tm * 2.0 * qw, tm * 2.0 * qx},
        {-tm * 2.0 * qx, -tm * 2.0 * qw, tm * 2.0 * qz, tm * 2.0 * qy},
        {0.0, -tm * 4.0 * qx, -tm * 4.0 * qy, 0.0}};
    // d qn / d q = (I - qn qn^T)/|q|  (zero in the degenerate branch)
    const double qn[4] = {qw, qx, qy, qz};
#pragma unroll
    for (int r = 0; r < 7; ++r) {
      const double gq = G[r][0] * qn[0] + G[r][1] * qn[1] + G[r][2] * qn[2] + G[r][3] * qn[3];
      const int row = (r < 4) ? (3 + r) : (7 + (r - 4));
#pragma unroll
      for (int c = 0; c < 4; ++c) Fx[row * 13 + 3 + c] = reg ? (G[r][c] - gq * qn[c]) * inv_norm : 0.0;
    }
    // d q_dot / d omega
    Fx[3 * 13 + 10] = -0.5 * qx; Fx[3 * 13 + 11] = -0.5 * qy; Fx[3 * 13 + 12] = -0.5 * qz;
    Fx[4 * 13 + 10] = 0.5 * qw;  Fx[4 * 13 + 11] = -0.5 * qz; Fx[4 * 13 + 12] = 0.5 * qy;
    Fx[5 * 13 + 10] = 0.5 * qz;  Fx[5 * 13 + 11] = 0.5 * qw;  Fx[5 * 13 + 12] = -0.5 * qx;
    Fx[6 * 13 + 10] = -0.5 * qy; Fx[6 * 13 + 11] = 0.5 * qx;  Fx[6 * 13 + 12] = 0.5 * qw;
    // d omega_dot / d omega = -Iinv * ([w]x I - [Iw]x)
    const double *I = J.I;
    const double Iw0 = I[0] * wx + I[1] * wy + I[2] * wz;
    const double Iw1 = I[3] * wx + I[4] * wy + I[5] * wz;
    const double Iw2 = I[6] * wx + I[7] * wy + I[8] * wz;
    // D = [w]x I - [Iw]x ; [a]x = [[0,-az,ay],[az,0,-ax],[-ay,ax,0]]
    double D[9];
    D[0] = (-wz * I[3] + wy * I[6]);
    D[1] = (-wz * I[4] + wy * I[7]) + Iw2;
    D[2] = (-wz * I[5] + wy * I[8]) - Iw1;
    D[3] = (wz * I[0] - wx * I[6]) - Iw2;
    D[4] = (wz * I[1] - wx * I[7]);
    D[5] = (wz * I[2] - wx * I[8]) + Iw0;
    D[6] = (-wy * I[0] + wx * I[3]) + Iw1;
    D[7] = (-wy * I[1] + wx * I[4]) - Iw0;
    D[8] = (-wy * I[2] + wx * I[5]);
#pragma unroll
    for (int r = 0; r < 3; ++r)
#pragma unroll
      for (int c = 0; c < 3; ++c)
        Fx[(10 + r) * 13 + 10 + c] =
            -(J.Iinv[r * 3 + 0] * D[0 * 3 + c] + J.Iinv[r * 3 + 1] * D[1 * 3 + c] + J.Iinv[r * 3 + 2] * D[2 * 3 + c]);
    // Fu: v_dot rows = r3/m for every rotor; omega_dot rows = Iinv * dtau/df
    const double im = P.p[25];
    const double r3x = 2.0 * (qx * qz + qy * qw), r3y = 2.0 * (qy * qz - qx * qw), r3z = 1.0 - 2.0 * (qx * qx + qy * qy);
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      Fu[7 * 4 + c] = im * r3x;
      Fu[8 * 4 + c] = im * r3y;
      Fu[9 * 4 + c] = im * r3z;
    }
    const double T[3][4] = {{L, 0.0, -L, 0.0}, {0.0, L, 0.0, -L}, {0.1, -0.1, 0.1, -0.1}};
#pragma unroll
    for (int r = 0; r < 3; ++r)
#pragma unroll
      for (int c = 0; c < 4; ++c)
        Fu[(10 + r) * 4 + c] = J.Iinv[r * 3 + 0] * T[0][c] + J.Iinv[r * 3 + 1] * T[1][c] + J.Iinv[r * 3 + 2] * T[2][c];
  }
};

// ---------------------------------------------------------------- LTI (runtime dims, discrete)
template <>
struct Model<CDDP_B200_MODEL_LTI> {
  static constexpr int NS = 0, NC = 0;  // runtime
};

// Integrators (dynamical_system.cpp:28-65), zero-order-hold control.  xn may alias neither x nor u.
template <int MODEL>
__device__ __forceinline__ void discrete_step(const ModelParams &P, int integrator, double dt, const double *x,
                                              const double *u, double *xn) {
  constexpr int NS = Model<MODEL>::NS;
  double k[NS], acc[NS], xt[NS];
  Model<MODEL>::f(P, x, u, k);
  if (integrator == CDDP_B200_EULER) {
#pragma unroll
    for (int i = 0; i < NS; ++i) xn[i] = x[i] + dt * k[i];
  } else if (integrator == CDDP_B200_HEUN) {
#pragma unroll
    for (int i = 0; i < NS; ++i) { acc[i] = k[i]; xt[i] = x[i] + dt * k[i]; }
    Model<MODEL>::f(P, xt, u, k);
#pragma unroll
    for (int i = 0; i < NS; ++i) xn[i] = x[i] + 0.5 * dt * (acc[i] + k[i]);
  } else if (integrator == CDDP_B200_RK3) {
    double k1[NS];
#pragma unroll
    for (int i = 0; i < NS; ++i) { k1[i] = k[i]; xt[i] = x[i] + 0.5 * dt * k[i]; }
    Model<MODEL>::f(P, xt, u, k);
#pragma unroll
    for (int i = 0; i < NS; ++i) { acc[i] = k1[i] + 4.0 * k[i]; xt[i] = x[i] - dt * k1[i] + 2.0 * dt * k[i]; }
    Model<MODEL>::f(P, xt, u, k);
#pragma unroll
    for (int i = 0; i < NS; ++i) xn[i] = x[i] + (dt / 6.0) * (acc[i] + k[i]);
  } else {
#pragma unroll
    for (int i = 0; i < NS; ++i) { acc[i] = k[i]; xt[i] = x[i] + 0.5 * dt * k[i]; }
    Model<MODEL>::f(P, xt, u, k);
#pragma unroll
    for (int i = 0; i < NS; ++i) { acc[i] += 2.0 * k[i]; xt[i] = x[i] + 0.5 * dt * k[i]; }
    Model<MODEL>::f(P, xt, u, k);
#pragma unroll
    for (int i = 0; i < NS; ++i) { acc[i] += 2.0 * k[i]; xt[i] = x[i] + dt * k[i]; }
    Model<MODEL>::f(P, xt, u, k);
#pragma unroll
    for (int i = 0; i < NS; ++i) xn[i] = x[i] + (dt / 6.0) * (acc[i] + k[i]);
  }
}

}  // namespace cddp_b200
