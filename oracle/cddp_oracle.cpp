/*
 * oracle/cddp_oracle.cpp — CPU restatement of the reference CLDDP hot path.
 *
 * TEST INFRASTRUCTURE ONLY (see cddp_oracle.h).  PARITY STATUS: "parity unpinned" — the
 * reference needs Eigen 3.4.0 + autodiff v1.1.2 (network FetchContent) and cannot be built here;
 * this file restates the algorithm from the cited reference lines and the published algorithms
 * of the Eigen routines the reference calls (PartialPivLU inverse, pivoted LDLT, eigenvalue PD test).
 *
 * Every function cites the reference file:line (relative to /root/reference) it follows.
 */
#include "cddp_oracle.h"

#include <algorithm>
#include <chrono>
#include <cmath>
#include <functional>
#include <cstring>
#include <limits>
#include <thread>
#include <vector>

namespace {

constexpr int MAXN = ORACLE_MAX_N;
constexpr int MAXM = 16; /* capacity (>= ORACLE_MAX_M): the reference BoxQP fixture of test_boxqp.cpp:125-220 is 15x15 */
constexpr int MAXD = MAXN + MAXM;

/* ------------------------------------------------------------------------------------------
 * Forward-mode dual number: stands in for autodiff::dual2nd + autodiff::jacobian
 * (dynamical_system.cpp:102-133, quadrotor.cpp:116-140).  First derivatives only — CLDDP never
 * consults Hessians (clddp_solver.cpp has no use_ilqr branch).
 * ---------------------------------------------------------------------------------------- */
struct Dual {
  double v;
  double d[MAXD];
  int nd;
};

inline Dual dconst(double c, int nd) {
  Dual r;
  r.v = c;
  r.nd = nd;
  for (int i = 0; i < nd; ++i) r.d[i] = 0.0;
  return r;
}
inline Dual operator+(const Dual &a, const Dual &b) {
  Dual r;
  r.nd = a.nd;
  r.v = a.v + b.v;
  for (int i = 0; i < a.nd; ++i) r.d[i] = a.d[i] + b.d[i];
  return r;
}
inline Dual operator-(const Dual &a, const Dual &b) {
  Dual r;
  r.nd = a.nd;
  r.v = a.v - b.v;
  for (int i = 0; i < a.nd; ++i) r.d[i] = a.d[i] - b.d[i];
  return r;
}
inline Dual operator-(const Dual &a) {
  Dual r;
  r.nd = a.nd;
  r.v = -a.v;
  for (int i = 0; i < a.nd; ++i) r.d[i] = -a.d[i];
  return r;
}
inline Dual operator*(const Dual &a, const Dual &b) {
  Dual r;
  r.nd = a.nd;
  r.v = a.v * b.v;
  for (int i = 0; i < a.nd; ++i) r.d[i] = a.d[i] * b.v + a.v * b.d[i];
  return r;
}
inline Dual operator/(const Dual &a, const Dual &b) {
  Dual r;
  r.nd = a.nd;
  r.v = a.v / b.v;
  for (int i = 0; i < a.nd; ++i) r.d[i] = (a.d[i] - r.v * b.d[i]) / b.v;
  return r;
}
inline Dual operator+(const Dual &a, double c) { Dual r = a; r.v += c; return r; }
inline Dual operator+(double c, const Dual &a) { return a + c; }
inline Dual operator-(const Dual &a, double c) { Dual r = a; r.v -= c; return r; }
inline Dual operator-(double c, const Dual &a) { return (-a) + c; }
inline Dual operator*(const Dual &a, double c) {
  Dual r = a;
  r.v *= c;
  for (int i = 0; i < a.nd; ++i) r.d[i] *= c;
  return r;
}
inline Dual operator*(double c, const Dual &a) { return a * c; }
inline Dual operator/(const Dual &a, double c) { return a * (1.0 / c); }
inline Dual sqrt(const Dual &a) {
  Dual r;
  r.nd = a.nd;
  r.v = std::sqrt(a.v);
  for (int i = 0; i < a.nd; ++i) r.d[i] = a.d[i] / (2.0 * r.v);
  return r;
}
inline Dual sin(const Dual &a) {
  Dual r;
  r.nd = a.nd;
  r.v = std::sin(a.v);
  const double c = std::cos(a.v);
  for (int i = 0; i < a.nd; ++i) r.d[i] = c * a.d[i];
  return r;
}
inline Dual cos(const Dual &a) {
  Dual r;
  r.nd = a.nd;
  r.v = std::cos(a.v);
  const double s = -std::sin(a.v);
  for (int i = 0; i < a.nd; ++i) r.d[i] = s * a.d[i];
  return r;
}
inline Dual tan(const Dual &a) {
  Dual r;
  r.nd = a.nd;
  r.v = std::tan(a.v);
  const double c = 1.0 + r.v * r.v;
  for (int i = 0; i < a.nd; ++i) r.d[i] = c * a.d[i];
  return r;
}
inline double val(double a) { return a; }
inline double val(const Dual &a) { return a.v; }
using std::cos;
using std::sin;
using std::sqrt;
using std::tan;

/* ------------------------------------------------------------------------------------------
 * Model dynamics, templated on the scalar so the same text serves f and its AD Jacobian.
 * `jac` selects the *Autodiff variant of the reference where the two differ (cartpole damping,
 * cartpole.cpp:90; pendulum gravity sign, pendulum.cpp:97 — the latter is never used by CLDDP
 * because Pendulum overrides the Jacobians analytically, pendulum.cpp:45-66).
 * ---------------------------------------------------------------------------------------- */

/* quadrotor.cpp:33-96 (double) and :162-221 (dual).  params: mass, Ixx,Ixy,Ixz,Iyx,...,Izz (9), arm_length */
template <typename T>
void quadrotor_f(const double *P, const T *x, const T *u, T *xd) {
  const double mass = P[0];
  const double *I = P + 1; /* row-major 3x3 */
  const double L = P[10];
  const double gravity = 9.81; /* quadrotor.hpp gravity_ */
  xd[0] = x[7];
  xd[1] = x[8];
  xd[2] = x[9];
  T qw = x[3], qx = x[4], qy = x[5], qz = x[6];
  T norm = sqrt(qw * qw + qx * qx + qy * qy + qz * qz);
  if (val(norm) > 1e-6) { /* quadrotor.cpp:45-55 */
    qw = qw / norm;
    qx = qx / norm;
    qy = qy / norm;
    qz = qz / norm;
  } else {
    qw = qw * 0.0 + 1.0;
    qx = qx * 0.0;
    qy = qy * 0.0;
    qz = qz * 0.0;
  }
  T wx = x[10], wy = x[11], wz = x[12];
  xd[3] = -0.5 * (qx * wx + qy * wy + qz * wz); /* :65-68 */
  xd[4] = 0.5 * (qw * wx + qy * wz - qz * wy);
  xd[5] = 0.5 * (qw * wy - qx * wz + qz * wx);
  xd[6] = 0.5 * (qw * wz + qx * wy - qy * wx);
  T f1 = u[0], f2 = u[1], f3 = u[2], f4 = u[3];
  T thrust = f1 + f2 + f3 + f4; /* :75-81 */
  T tau_x = L * (f1 - f3);
  T tau_y = L * (f2 - f4);
  T tau_z = 0.1 * (f1 - f2 + f3 - f4);
  /* rotation matrix third column (thrust along body z), :99-113 */
  T R02 = 2.0 * (qx * qz + qy * qw);
  T R12 = 2.0 * (qy * qz - qx * qw);
  T R22 = 1.0 - 2.0 * (qx * qx + qy * qy);
  const double im = 1.0 / mass;
  xd[7] = im * (R02 * thrust); /* :86-88 */
  xd[8] = im * (R12 * thrust);
  xd[9] = im * (R22 * thrust) - gravity;
  /* angular_acc = I^{-1} (tau - w x (I w)), :90-93; 3x3 cofactor inverse (Eigen fixed-size inverse) */
  T Iw0 = I[0] * wx + I[1] * wy + I[2] * wz;
  T Iw1 = I[3] * wx + I[4] * wy + I[5] * wz;
  T Iw2 = I[6] * wx + I[7] * wy + I[8] * wz;
  T r0 = tau_x - (wy * Iw2 - wz * Iw1);
  T r1 = tau_y - (wz * Iw0 - wx * Iw2);
  T r2 = tau_z - (wx * Iw1 - wy * Iw0);
  const double c00 = I[4] * I[8] - I[5] * I[7], c01 = I[2] * I[7] - I[1] * I[8], c02 = I[1] * I[5] - I[2] * I[4];
  const double c10 = I[5] * I[6] - I[3] * I[8], c11 = I[0] * I[8] - I[2] * I[6], c12 = I[2] * I[3] - I[0] * I[5];
  const double c20 = I[3] * I[7] - I[4] * I[6], c21 = I[1] * I[6] - I[0] * I[7], c22 = I[0] * I[4] - I[1] * I[3];
  const double det = I[0] * c00 + I[1] * c10 + I[2] * c20;
  const double id = 1.0 / det;
  xd[10] = (c00 * id) * r0 + (c01 * id) * r1 + (c02 * id) * r2;
  xd[11] = (c10 * id) * r0 + (c11 * id) * r1 + (c12 * id) * r2;
  xd[12] = (c20 * id) * r0 + (c21 * id) * r1 + (c22 * id) * r2;
}

/* cartpole.cpp:38-62 (double) and :64-93 (dual, adds -damping*theta_dot).
 * params: cart_mass, pole_mass, pole_length, gravity, damping */
template <typename T>
void cartpole_f(const double *P, const T *x, const T *u, T *xd, bool jac) {
  const double mc = P[0], mp = P[1], l = P[2], g = P[3], damping = P[4];
  T theta = x[1], x_dot = x[2], theta_dot = x[3], force = u[0];
  T s = sin(theta), c = cos(theta);
  const double total_mass = mc + mp;
  T den = mc + mp * s * s;
  xd[0] = x_dot;
  xd[1] = theta_dot;
  xd[2] = (force + mp * s * (l * theta_dot * theta_dot + g * c)) / den;
  T num = -force * c - mp * l * theta_dot * theta_dot * c * s - total_mass * g * s;
  if (jac) num = num - damping * theta_dot; /* cartpole.cpp:90 */
  xd[3] = num / (l * den);
}

/* pendulum.cpp:29-43.  params: length, mass, damping */
/* bicycle.cpp:29-47 (double) / :49-64 (dual): state (x, y, theta, v), control (a, delta); params: wheelbase */
template <typename T>
void bicycle_f(const double *P, const T *x, const T *u, T *xd) {
  xd[0] = x[3] * cos(x[2]);
  xd[1] = x[3] * sin(x[2]);
  xd[2] = (x[3] / P[0]) * tan(u[1]);
  xd[3] = u[0];
}

/* NOT a reference model: a 7-joint chain with gravity, viscous friction and nearest-neighbour elastic coupling,
 * q_i'' = (tau_i - g_i sin q_i - c q_i' - k (sin(q_i - q_{i-1}) + sin(q_i - q_{i+1}))) / I_i, state (q, q'), n = 14, m = 7.
 * It stands in for BASELINE config #5's "7-DOF manipulator" (no such model exists in the reference) and
 * exercises the user-model plugin at n = 14, m = 7.  params: g, c, k, I_1..I_7. */
template <typename T>
void chain7_f(const double *P, const T *x, const T *u, T *xd) {
  const double g = P[0], c = P[1], k = P[2];
  for (int i = 0; i < 7; ++i) {
    xd[i] = x[7 + i];
    T acc = u[i] - g * sin(x[i]) - c * x[7 + i];
    if (i > 0) acc = acc - k * sin(x[i] - x[i - 1]);
    if (i < 6) acc = acc - k * sin(x[i] - x[i + 1]);
    xd[7 + i] = acc / P[3 + i];
  }
}

/* 7-DOF serial manipulator: the n-joint generalisation of the reference's 3-DOF Manipulator (manipulator.cpp:29-51,
 * :174-208 — point masses at the link ends, base joint about the vertical, M_ij = mu_max(i,j) l_i l_j cos(q_{i+1}+...+q_j),
 * G_k = -sum_{j>=k} mu_j g l_j cos(q_1+...+q_j)) with the Coriolis / centrifugal vector of that M(q) and viscous friction:
 * M(q) qdd + h(q,qd) + G(q) + b qd = tau.  The native twin of the plugin model MANIP7_SOURCE (cddp-cpp_b200/problems.py);
 * BASELINE config #5.  params: g, b, m_0..m_6, l_0..l_6.  Written independently of the plugin text: explicit h_k sums,
 * Gaussian elimination with the mass matrix kept whole. */
template <typename T>
void manip7_f(const double *P, const T *x, const T *u, T *xd) {
  const double g = P[0], bv = P[1];
  const double *mass = P + 2, *len = P + 9;
  double mu[7];
  for (int j = 0; j < 7; ++j) {
    mu[j] = 0.0;
    for (int k = j; k < 7; ++k) mu[j] += mass[k];
  }
  const T zero = x[0] * 0.0;
  T sig[7], w[7];
  sig[0] = zero;
  w[0] = zero;
  for (int j = 1; j < 7; ++j) {
    sig[j] = sig[j - 1] + x[j];
    w[j] = w[j - 1] + x[7 + j];
  }
  T M[7][7], S[7][7], rhs[7];
  for (int i = 0; i < 7; ++i)
    for (int j = 0; j < 7; ++j) {
      const int lo = i < j ? i : j, hi = i < j ? j : i;
      const double a = mu[hi] * len[lo] * len[hi];
      M[i][j] = a * cos(sig[hi] - sig[lo]);
      S[i][j] = a * sin(sig[hi] - sig[lo]); /* symmetric: S_ij = S_ji = a sin(sigma_hi - sigma_lo) */
    }
  for (int k = 0; k < 7; ++k) {
    T h = zero; /* h_k = sum_j Mdot_kj qd_j - 1/2 d/dq_k (qd^T M qd) */
    for (int j = 0; j < 7; ++j) {
      if (j == k) continue;
      const int lo = k < j ? k : j, hi = k < j ? j : k;
      h = h - S[k][j] * (w[hi] - w[lo]) * x[7 + j];
    }
    for (int i = 0; i < k; ++i)
      for (int j = k; j < 7; ++j) h = h + S[i][j] * x[7 + i] * x[7 + j];
    T G = zero;
    if (k >= 1)
      for (int j = k; j < 7; ++j) G = G - (mu[j] * g * len[j]) * cos(sig[j]);
    rhs[k] = u[k] - h - G - bv * x[7 + k];
  }
  /* qdd = M^-1 rhs: Gaussian elimination without pivoting (M is symmetric positive definite) */
  for (int c = 0; c < 7; ++c)
    for (int r = c + 1; r < 7; ++r) {
      const T f = M[r][c] / M[c][c];
      for (int k = c; k < 7; ++k) M[r][k] = M[r][k] - f * M[c][k];
      rhs[r] = rhs[r] - f * rhs[c];
    }
  for (int r = 6; r >= 0; --r) {
    T acc = rhs[r];
    for (int k = r + 1; k < 7; ++k) acc = acc - M[r][k] * rhs[k];
    rhs[r] = acc / M[r][r];
  }
  for (int i = 0; i < 7; ++i) {
    xd[i] = x[7 + i];
    xd[7 + i] = rhs[i];
  }
}

void pendulum_f(const double *P, const double *x, const double *u, double *xd) {
  const double length = P[0], mass = P[1], damping = P[2], gravity = 9.81;
  const double inertia = mass * length * length;
  xd[0] = x[1];
  xd[1] = (u[0] - damping * x[1] + mass * gravity * length * std::sin(x[0])) / inertia;
}

/* unicycle.cpp:28-41 */
void unicycle_f(const double *x, const double *u, double *xd) {
  xd[0] = u[0] * std::cos(x[2]);
  xd[1] = u[0] * std::sin(x[2]);
  xd[2] = u[1];
}

void lti_step(const oracle_problem *p, const double *x, const double *u, double *xn) {
  /* lti_system.cpp:71-76 */
  const int n = p->n, m = p->m;
  for (int i = 0; i < n; ++i) {
    double s = 0.0;
    for (int j = 0; j < n; ++j) s += p->lti_A[i * n + j] * x[j];
    double s2 = 0.0;
    for (int j = 0; j < m; ++j) s2 += p->lti_B[i * m + j] * u[j];
    xn[i] = s + s2;
  }
}

void continuous_dynamics(const oracle_problem *p, const double *x, const double *u, double t, double *xd);

/* dynamical_system.cpp:28-83 */
void discrete_dynamics(const oracle_problem *p, const double *x, const double *u, double t, double *xn) {
  const int n = p->n;
  const double dt = p->dt;
  if (p->model == ORACLE_LTI) {
    lti_step(p, x, u, xn);
    return;
  }
  double k1[MAXN], k2[MAXN], k3[MAXN], k4[MAXN], tmp[MAXN];
  switch (p->integrator) {
    case ORACLE_EULER:
      continuous_dynamics(p, x, u, t, k1);
      for (int i = 0; i < n; ++i) xn[i] = x[i] + dt * k1[i];
      break;
    case ORACLE_HEUN:
      continuous_dynamics(p, x, u, t, k1);
      for (int i = 0; i < n; ++i) tmp[i] = x[i] + dt * k1[i];
      continuous_dynamics(p, tmp, u, t + dt, k2);
      for (int i = 0; i < n; ++i) xn[i] = x[i] + 0.5 * dt * (k1[i] + k2[i]);
      break;
    case ORACLE_RK3:
      continuous_dynamics(p, x, u, t, k1);
      for (int i = 0; i < n; ++i) tmp[i] = x[i] + 0.5 * dt * k1[i];
      continuous_dynamics(p, tmp, u, t + 0.5 * dt, k2);
      for (int i = 0; i < n; ++i) tmp[i] = x[i] - dt * k1[i] + 2 * dt * k2[i];
      continuous_dynamics(p, tmp, u, t + dt, k3);
      for (int i = 0; i < n; ++i) xn[i] = x[i] + (dt / 6) * (k1[i] + 4 * k2[i] + k3[i]);
      break;
    default: /* rk4 */
      continuous_dynamics(p, x, u, t, k1);
      for (int i = 0; i < n; ++i) tmp[i] = x[i] + 0.5 * dt * k1[i];
      continuous_dynamics(p, tmp, u, t + 0.5 * dt, k2);
      for (int i = 0; i < n; ++i) tmp[i] = x[i] + 0.5 * dt * k2[i];
      continuous_dynamics(p, tmp, u, t + 0.5 * dt, k3);
      for (int i = 0; i < n; ++i) tmp[i] = x[i] + dt * k3[i];
      continuous_dynamics(p, tmp, u, t + dt, k4);
      for (int i = 0; i < n; ++i) xn[i] = x[i] + (dt / 6) * (k1[i] + 2 * k2[i] + 2 * k3[i] + k4[i]);
      break;
  }
}

void continuous_dynamics(const oracle_problem *p, const double *x, const double *u, double t, double *xd) {
  switch (p->model) {
    case ORACLE_PENDULUM: pendulum_f(p->model_params, x, u, xd); break;
    case ORACLE_CARTPOLE: cartpole_f<double>(p->model_params, x, u, xd, false); break;
    case ORACLE_UNICYCLE: unicycle_f(x, u, xd); break;
    case ORACLE_QUADROTOR: quadrotor_f<double>(p->model_params, x, u, xd); break;
    case ORACLE_BICYCLE: bicycle_f<double>(p->model_params, x, u, xd); break;
    case ORACLE_CHAIN7: chain7_f<double>(p->model_params, x, u, xd); break;
    case ORACLE_MANIP7: manip7_f<double>(p->model_params, x, u, xd); break;
    case ORACLE_LTI: {
      /* base-class fallback dynamical_system.cpp:85-98: (x_next - x)/dt */
      double xn[MAXN];
      lti_step(p, x, u, xn);
      for (int i = 0; i < p->n; ++i) xd[i] = (xn[i] - x[i]) / p->dt;
      break;
    }
    default:
      for (int i = 0; i < p->n; ++i) xd[i] = 0.0;
  }
  (void)t;
}

/* getJacobians: continuous-time Fx (n x n), Fu (n x m). */
void jacobians(const oracle_problem *p, const double *x, const double *u, double t, double *Fx, double *Fu) {
  const int n = p->n, m = p->m;
  std::fill(Fx, Fx + n * n, 0.0);
  std::fill(Fu, Fu + n * m, 0.0);
  const double *P = p->model_params;
  switch (p->model) {
    case ORACLE_PENDULUM: { /* pendulum.cpp:45-66 */
      const double length = P[0], mass = P[1], damping = P[2], gravity = 9.81;
      Fx[0 * 2 + 1] = 1.0;
      Fx[1 * 2 + 0] = (gravity / length) * std::cos(x[0]);
      Fx[1 * 2 + 1] = -damping / (mass * length * length);
      Fu[1] = 1.0 / (mass * length * length);
      break;
    }
    case ORACLE_UNICYCLE: { /* unicycle.cpp:43-66 */
      Fx[0 * 3 + 2] = -u[0] * std::sin(x[2]);
      Fx[1 * 3 + 2] = u[0] * std::cos(x[2]);
      Fu[0 * 2 + 0] = std::cos(x[2]);
      Fu[1 * 2 + 0] = std::sin(x[2]);
      Fu[2 * 2 + 1] = 1.0;
      break;
    }
    case ORACLE_LTI: { /* lti_system.cpp:78-92 */
      for (int i = 0; i < n; ++i)
        for (int j = 0; j < n; ++j) Fx[i * n + j] = (p->lti_A[i * n + j] - (i == j ? 1.0 : 0.0)) / p->dt;
      for (int i = 0; i < n * m; ++i) Fu[i] = p->lti_B[i] / p->dt;
      break;
    }
    case ORACLE_CARTPOLE:
    case ORACLE_BICYCLE: /* analytic in the reference (bicycle.cpp:66-113) = the derivative of the same expressions */
    case ORACLE_CHAIN7:
    case ORACLE_MANIP7:
    case ORACLE_QUADROTOR: { /* autodiff::jacobian: cartpole.cpp:95-103, quadrotor.cpp:116-140 */
      const int nd = n + m;
      Dual xs[MAXN], us[MAXM], xd[MAXN];
      for (int i = 0; i < n; ++i) {
        xs[i] = dconst(x[i], nd);
        xs[i].d[i] = 1.0;
      }
      for (int j = 0; j < m; ++j) {
        us[j] = dconst(u[j], nd);
        us[j].d[n + j] = 1.0;
      }
      for (int i = 0; i < n; ++i) xd[i] = dconst(0.0, nd);
      if (p->model == ORACLE_CARTPOLE)
        cartpole_f<Dual>(P, xs, us, xd, true);
      else if (p->model == ORACLE_BICYCLE)
        bicycle_f<Dual>(P, xs, us, xd);
      else if (p->model == ORACLE_CHAIN7)
        chain7_f<Dual>(P, xs, us, xd);
      else if (p->model == ORACLE_MANIP7)
        manip7_f<Dual>(P, xs, us, xd);
      else
        quadrotor_f<Dual>(P, xs, us, xd);
      for (int i = 0; i < n; ++i) {
        for (int j = 0; j < n; ++j) Fx[i * n + j] = xd[i].d[j];
        for (int j = 0; j < m; ++j) Fu[i * m + j] = xd[i].d[n + j];
      }
      break;
    }
    default: break;
  }
  (void)t;
}

/* ------------------------------------------------------------------------------------------
 * QuadraticObjective (objective.cpp:30-154): Q_ = Q*dt, R_ = R*dt, no 1/2 factors.
 * ---------------------------------------------------------------------------------------- */
double running_cost(const oracle_problem *p, const double *x, const double *u, const double *ref) {
  const int n = p->n, m = p->m;
  const double dt = p->dt;
  double e[MAXN];
  for (int i = 0; i < n; ++i) e[i] = x[i] - ref[i];
  /* (e^T Q_) e  — left-to-right as Eigen evaluates the product chain (objective.cpp:89-90) */
  double sx = 0.0;
  for (int j = 0; j < n; ++j) {
    double r = 0.0;
    for (int i = 0; i < n; ++i) r += e[i] * (p->Q[i * n + j] * dt);
    sx += r * e[j];
  }
  double su = 0.0;
  for (int j = 0; j < m; ++j) {
    double r = 0.0;
    for (int i = 0; i < m; ++i) r += u[i] * (p->R[i * m + j] * dt);
    su += r * u[j];
  }
  return sx + su;
}

double terminal_cost(const oracle_problem *p, const double *x, const double *ref) { /* objective.cpp:94-98 */
  const int n = p->n;
  double e[MAXN];
  for (int i = 0; i < n; ++i) e[i] = x[i] - ref[i];
  double s = 0.0;
  for (int j = 0; j < n; ++j) {
    double r = 0.0;
    for (int i = 0; i < n; ++i) r += e[i] * p->Qf[i * n + j];
    s += r * e[j];
  }
  return s;
}

inline const double *ref_at(const oracle_problem *p, const double *xref, const double *ref_traj, int t) {
  /* objective.cpp:84-88: reference_states_[index] if a trajectory was given else reference_state_ */
  return ref_traj ? ref_traj + (size_t)t * p->n : xref;
}

/* cddp_solver_base.cpp:416-424 */
double trajectory_cost(const oracle_problem *p, const double *X, const double *U, const double *xref,
                       const double *ref_traj) {
  const int n = p->n, m = p->m, N = p->horizon;
  double J = 0.0;
  for (int t = 0; t < N; ++t) J += running_cost(p, X + (size_t)t * n, U + (size_t)t * m, ref_at(p, xref, ref_traj, t));
  J += terminal_cost(p, X + (size_t)N * n, xref);
  return J;
}

/* ------------------------------------------------------------------------------------------
 * Small dense numerics standing in for the Eigen 3.4.0 routines the reference calls.
 * ---------------------------------------------------------------------------------------- */

/* Eigen::EigenSolver(M).eigenvalues().real().minCoeff() for a (numerically) symmetric M
 * (clddp_solver.cpp:133-134): cyclic Jacobi on the symmetric part. */
double min_eigenvalue_sym(const double *M, int n) {
  double a[MAXM * MAXM];
  for (int i = 0; i < n; ++i)
    for (int j = 0; j < n; ++j) a[i * n + j] = 0.5 * (M[i * n + j] + M[j * n + i]);
  for (int sweep = 0; sweep < 64; ++sweep) {
    double off = 0.0;
    for (int i = 0; i < n; ++i)
      for (int j = i + 1; j < n; ++j) off += a[i * n + j] * a[i * n + j];
    if (off < 1e-300) break;
    for (int p_ = 0; p_ < n; ++p_)
      for (int q = p_ + 1; q < n; ++q) {
        const double apq = a[p_ * n + q];
        if (apq == 0.0) continue;
        const double theta = (a[q * n + q] - a[p_ * n + p_]) / (2.0 * apq);
        const double tt = (theta >= 0 ? 1.0 : -1.0) / (std::fabs(theta) + std::sqrt(theta * theta + 1.0));
        const double c = 1.0 / std::sqrt(tt * tt + 1.0), s = tt * c;
        for (int k = 0; k < n; ++k) {
          const double akp = a[k * n + p_], akq = a[k * n + q];
          a[k * n + p_] = c * akp - s * akq;
          a[k * n + q] = s * akp + c * akq;
        }
        for (int k = 0; k < n; ++k) {
          const double apk = a[p_ * n + k], aqk = a[q * n + k];
          a[p_ * n + k] = c * apk - s * aqk;
          a[q * n + k] = s * apk + c * aqk;
        }
      }
  }
  double mn = a[0];
  for (int i = 1; i < n; ++i) mn = std::min(mn, a[i * n + i]);
  return mn;
}

/* all eigenvalues of a symmetric matrix (cyclic Jacobi), used for the singular values of a small square matrix:
 * sigma_i = sqrt(eig_i(A^T A)) stands in for Eigen::JacobiSVD(A).singularValues() (ipddp_solver.cpp:557-560) */
void eigenvalues_sym(const double *M, int n, double *ev) {
  double a[MAXM * MAXM];
  for (int i = 0; i < n; ++i)
    for (int j = 0; j < n; ++j) a[i * n + j] = 0.5 * (M[i * n + j] + M[j * n + i]);
  for (int sweep = 0; sweep < 64; ++sweep) {
    double off = 0.0;
    for (int i = 0; i < n; ++i)
      for (int j = i + 1; j < n; ++j) off += a[i * n + j] * a[i * n + j];
    if (off < 1e-300) break;
    for (int p_ = 0; p_ < n; ++p_)
      for (int q = p_ + 1; q < n; ++q) {
        const double apq = a[p_ * n + q];
        if (apq == 0.0) continue;
        const double theta = (a[q * n + q] - a[p_ * n + p_]) / (2.0 * apq);
        const double tt = (theta >= 0 ? 1.0 : -1.0) / (std::fabs(theta) + std::sqrt(theta * theta + 1.0));
        const double c = 1.0 / std::sqrt(tt * tt + 1.0), s = tt * c;
        for (int k = 0; k < n; ++k) {
          const double akp = a[k * n + p_], akq = a[k * n + q];
          a[k * n + p_] = c * akp - s * akq;
          a[k * n + q] = s * akp + c * akq;
        }
        for (int k = 0; k < n; ++k) {
          const double apk = a[p_ * n + k], aqk = a[q * n + k];
          a[p_ * n + k] = c * apk - s * aqk;
          a[q * n + k] = s * apk + c * aqk;
        }
      }
  }
  for (int i = 0; i < n; ++i) ev[i] = a[i * n + i];
}

/* MatrixXd::inverse() for a dynamic matrix = PartialPivLU().inverse() (clddp_solver.cpp:143). */
void inverse_lu(const double *M, int n, double *Inv) {
  double a[MAXM * MAXM];
  int perm[MAXM];
  std::memcpy(a, M, sizeof(double) * n * n);
  for (int i = 0; i < n; ++i) perm[i] = i;
  for (int k = 0; k < n; ++k) {
    int piv = k;
    double best = std::fabs(a[k * n + k]);
    for (int i = k + 1; i < n; ++i)
      if (std::fabs(a[i * n + k]) > best) {
        best = std::fabs(a[i * n + k]);
        piv = i;
      }
    if (piv != k) {
      for (int j = 0; j < n; ++j) std::swap(a[k * n + j], a[piv * n + j]);
      std::swap(perm[k], perm[piv]);
    }
    for (int i = k + 1; i < n; ++i) {
      a[i * n + k] /= a[k * n + k];
      for (int j = k + 1; j < n; ++j) a[i * n + j] -= a[i * n + k] * a[k * n + j];
    }
  }
  for (int c = 0; c < n; ++c) {
    double y[MAXM];
    for (int i = 0; i < n; ++i) {
      double s = (perm[i] == c) ? 1.0 : 0.0;
      for (int j = 0; j < i; ++j) s -= a[i * n + j] * y[j];
      y[i] = s;
    }
    for (int i = n - 1; i >= 0; --i) {
      double s = y[i];
      for (int j = i + 1; j < n; ++j) s -= a[i * n + j] * y[j];
      y[i] = s / a[i * n + i];
    }
    for (int i = 0; i < n; ++i) Inv[i * n + c] = y[i];
  }
}

/* Eigen::LDLT<MatrixXd> (lower, diagonal pivoting) — the published algorithm of Eigen 3.4.0
 * LDLT.h ldlt_inplace<Lower>::unblocked + solve (boxqp.hpp:62, boxqp.cpp:105,147,
 * clddp_solver.cpp:174). */
struct LDLT {
  int n = 0;
  double a[MAXM * MAXM]; /* L strictly below the diagonal, D on the diagonal */
  int tr[MAXM];          /* transpositions */
  bool ok = true;

  void compute(const double *M, int n_) {
    n = n_;
    ok = true;
    std::memcpy(a, M, sizeof(double) * n * n);
    bool found_zero_pivot = false;
    for (int k = 0; k < n; ++k) {
      int big = k;
      double best = std::fabs(a[k * n + k]);
      for (int i = k + 1; i < n; ++i)
        if (std::fabs(a[i * n + i]) > best) {
          best = std::fabs(a[i * n + i]);
          big = i;
        }
      tr[k] = big;
      if (big != k) {
        /* symmetric row/column interchange on the lower triangle */
        const int s = n - big - 1;
        for (int j = 0; j < k; ++j) std::swap(a[k * n + j], a[big * n + j]);
        for (int i = 0; i < s; ++i) std::swap(a[(big + 1 + i) * n + k], a[(big + 1 + i) * n + big]);
        std::swap(a[k * n + k], a[big * n + big]);
        for (int i = k + 1; i < big; ++i) std::swap(a[i * n + k], a[big * n + i]);
      }
      const int rs = n - k - 1;
      if (k > 0) {
        double temp[MAXM];
        for (int j = 0; j < k; ++j) temp[j] = a[j * n + j] * a[k * n + j];
        double s = 0.0;
        for (int j = 0; j < k; ++j) s += a[k * n + j] * temp[j];
        a[k * n + k] -= s;
        for (int i = 0; i < rs; ++i) {
          double s2 = 0.0;
          for (int j = 0; j < k; ++j) s2 += a[(k + 1 + i) * n + j] * temp[j];
          a[(k + 1 + i) * n + k] -= s2;
        }
      }
      const double akk = a[k * n + k];
      const bool pivot_is_valid = std::fabs(akk) > 0.0;
      if (k == 0 && !pivot_is_valid) {
        for (int j = 0; j < n; ++j) tr[j] = j;
        return;
      }
      if (rs > 0 && pivot_is_valid) {
        for (int i = 0; i < rs; ++i) a[(k + 1 + i) * n + k] /= akk;
      } else if (rs > 0) {
        for (int i = 0; i < rs; ++i)
          if (a[(k + 1 + i) * n + k] != 0.0) ok = false;
      }
      if (found_zero_pivot && pivot_is_valid)
        ok = false;
      else if (!pivot_is_valid)
        found_zero_pivot = true;
    }
  }

  void solve_inplace(double *b) const {
    for (int k = 0; k < n; ++k)
      if (tr[k] != k) std::swap(b[k], b[tr[k]]);
    for (int i = 0; i < n; ++i)
      for (int j = 0; j < i; ++j) b[i] -= a[i * n + j] * b[j];
    const double tol = std::numeric_limits<double>::min();
    for (int i = 0; i < n; ++i) {
      if (std::fabs(a[i * n + i]) > tol)
        b[i] /= a[i * n + i];
      else
        b[i] = 0.0;
    }
    for (int i = n - 1; i >= 0; --i)
      for (int j = i + 1; j < n; ++j) b[i] -= a[j * n + i] * b[j];
    for (int k = n - 1; k >= 0; --k)
      if (tr[k] != k) std::swap(b[k], b[tr[k]]);
  }
};

/* ------------------------------------------------------------------------------------------
 * BoxQPSolver (boxqp.cpp:25-250)
 * ---------------------------------------------------------------------------------------- */
struct BoxQPOut {
  double x[MAXM];
  int free_mask[MAXM];
  int status;
  int iterations;
  int factorizations;
  double final_value;
  double final_grad_norm;
  LDLT Hfree;
  int nfree_factor; /* size of the factor currently held */
};

inline double qp_value(const double *x, const double *H, const double *g, int n) { /* boxqp.cpp:235-239 */
  double Hx[MAXM];
  for (int i = 0; i < n; ++i) {
    double s = 0.0;
    for (int j = 0; j < n; ++j) s += H[i * n + j] * x[j];
    Hx[i] = s;
  }
  double a = 0.0, b = 0.0;
  for (int i = 0; i < n; ++i) {
    a += x[i] * Hx[i];
    b += g[i] * x[i];
  }
  return 0.5 * a + b;
}

/* test/profiling instrumentation: histogram of (iterations, line-search trials, factorizations) per BoxQP call */
struct QpStats {
  bool on = false;
  long long iters[32] = {0}, trials[64] = {0}, facts[16] = {0}, calls = 0;
};
QpStats g_qp_stats;

void boxqp_solve(const oracle_options *o, int n, const double *H, const double *g, const double *lower,
                 const double *upper, const double *x0, BoxQPOut *r) {
  int dbg_trials = 0;
  struct StatGuard {
    BoxQPOut *r;
    int *trials;
    ~StatGuard() {
      if (!g_qp_stats.on) return;
      g_qp_stats.calls++;
      g_qp_stats.iters[std::min(r->iterations, 31)]++;
      g_qp_stats.trials[std::min(*trials, 63)]++;
      g_qp_stats.facts[std::min(r->factorizations, 15)]++;
    }
  } stat_guard{r, &dbg_trials};
  r->status = ORACLE_QP_MAX_ITER_EXCEEDED;
  r->iterations = 0;
  r->factorizations = 0;
  r->final_grad_norm = 0.0;
  r->nfree_factor = 0;
  r->Hfree.n = 0;
  /* initializeX, boxqp.cpp:184-205 */
  if (x0) {
    for (int i = 0; i < n; ++i) r->x[i] = std::min(std::max(x0[i], lower[i]), upper[i]);
  } else {
    for (int i = 0; i < n; ++i) {
      if (std::isfinite(lower[i]) && std::isfinite(upper[i]))
        r->x[i] = 0.5 * (lower[i] + upper[i]);
      else if (std::isfinite(lower[i]))
        r->x[i] = lower[i];
      else if (std::isfinite(upper[i]))
        r->x[i] = upper[i];
      else
        r->x[i] = 0.0;
    }
  }
  int clamped[MAXM], old_clamped[MAXM];
  for (int i = 0; i < n; ++i) {
    clamped[i] = 0;
    r->free_mask[i] = 1;
  }
  double value = qp_value(r->x, H, g, n);
  double old_value = std::numeric_limits<double>::infinity();

  for (int iter = 0; iter < o->qp_max_iterations; ++iter) {
    r->iterations = iter + 1;
    if (iter > 0 && std::fabs(old_value - value) < o->qp_min_relative_improvement * std::fabs(old_value)) {
      r->status = ORACLE_QP_SUCCESS; /* :52-57 */
      break;
    }
    old_value = value;
    double grad[MAXM];
    for (int i = 0; i < n; ++i) { /* :61 */
      double s = 0.0;
      for (int j = 0; j < n; ++j) s += H[i * n + j] * r->x[j];
      grad[i] = g[i] + s;
    }
    int nclamped = 0;
    for (int i = 0; i < n; ++i) { /* :64-73 */
      old_clamped[i] = clamped[i];
      clamped[i] = ((r->x[i] == lower[i] && grad[i] > 0) || (r->x[i] == upper[i] && grad[i] < 0)) ? 1 : 0;
      r->free_mask[i] = 1 - clamped[i];
      nclamped += clamped[i];
    }
    if (nclamped == n) { /* :76-79 */
      r->status = ORACLE_QP_ALL_CLAMPED;
      break;
    }
    bool any_different = false;
    for (int i = 0; i < n; ++i)
      if (old_clamped[i] != clamped[i]) any_different = true;
    int free_idx[MAXM], nf = 0;
    for (int i = 0; i < n; ++i)
      if (!clamped[i]) free_idx[nf++] = i;
    if (iter == 0 || any_different) { /* :89-111 */
      double Hf[MAXM * MAXM];
      for (int i = 0; i < nf; ++i)
        for (int j = 0; j < nf; ++j) Hf[i * nf + j] = H[free_idx[i] * n + free_idx[j]];
      r->Hfree.compute(Hf, nf);
      r->nfree_factor = nf;
      if (!r->Hfree.ok) {
        r->status = ORACLE_QP_HESSIAN_NOT_PD;
        break;
      }
      r->factorizations++;
    }
    double gn = 0.0; /* :114-125 */
    for (int i = 0; i < n; ++i)
      if (!clamped[i]) gn += grad[i] * grad[i];
    gn = std::sqrt(gn);
    r->final_grad_norm = gn;
    if (gn < o->qp_min_gradient_norm) {
      r->status = ORACLE_QP_SUCCESS;
      break;
    }
    double search[MAXM], gc[MAXM]; /* :128-152 */
    for (int i = 0; i < n; ++i) {
      search[i] = 0.0;
      gc[i] = g[i];
    }
    for (int i = 0; i < n; ++i)
      if (clamped[i])
        for (int j = 0; j < n; ++j) gc[j] += H[j * n + i] * r->x[i];
    double gf[MAXM];
    for (int i = 0; i < nf; ++i) gf[i] = gc[free_idx[i]];
    r->Hfree.solve_inplace(gf);
    for (int i = 0; i < nf; ++i) search[free_idx[i]] = -gf[i] - r->x[free_idx[i]];
    double sdotg = 0.0; /* :155-159 */
    for (int i = 0; i < n; ++i) sdotg += search[i] * grad[i];
    if (sdotg >= 0) {
      r->status = ORACLE_QP_NO_DESCENT;
      break;
    }
    /* lineSearch, :207-233 */
    double step = 1.0;
    bool ls_ok = false;
    double xn[MAXM];
    while (step > o->qp_min_step_size) {
      ++dbg_trials;
      for (int i = 0; i < n; ++i) xn[i] = std::min(std::max(r->x[i] + step * search[i], lower[i]), upper[i]);
      const double vn = qp_value(xn, H, g, n);
      if ((vn - value) <= o->qp_armijo_constant * step * sdotg) {
        ls_ok = true;
        break;
      }
      step *= o->qp_step_decrease_factor;
    }
    if (!ls_ok) {
      r->status = ORACLE_QP_MAX_LS_EXCEEDED;
      break;
    }
    for (int i = 0; i < n; ++i) r->x[i] = xn[i];
    value = qp_value(r->x, H, g, n); /* :170-172 */
  }
  r->final_value = value;
}

/* ------------------------------------------------------------------------------------------
 * CLDDPSolver::backwardPass (clddp_solver.cpp:79-204) on stacked A,B (or computed on the fly).
 * ---------------------------------------------------------------------------------------- */
int backward_pass(const oracle_problem *p, const oracle_options *o, const double *Astack, const double *Bstack,
                  const double *X, const double *U, const double *xref, const double *ref_traj, double reg, double *K,
                  double *k, double *dV, double *inf_du, double *Vx_dbg, double *Vxx_dbg, int *fail_t) {
  const int n = p->n, m = p->m, N = p->horizon;
  const double dt = p->dt;
  double Vx[MAXN], Vxx[MAXN * MAXN];
  /* terminal derivatives, objective.cpp:122-126,151-154 */
  {
    const double *xN = X + (size_t)N * n;
    double e[MAXN];
    for (int i = 0; i < n; ++i) e[i] = xN[i] - xref[i];
    for (int i = 0; i < n; ++i) {
      double s = 0.0;
      for (int j = 0; j < n; ++j) s += (2.0 * p->Qf[i * n + j]) * e[j];
      Vx[i] = s;
      for (int j = 0; j < n; ++j) Vxx[i * n + j] = 2.0 * p->Qf[i * n + j];
    }
  }
  if (Vx_dbg) std::memcpy(Vx_dbg + (size_t)N * n, Vx, sizeof(double) * n);
  if (Vxx_dbg) std::memcpy(Vxx_dbg + (size_t)N * n * n, Vxx, sizeof(double) * n * n);
  dV[0] = dV[1] = 0.0;
  double norm_Vx = 0.0;
  for (int i = 0; i < n; ++i) norm_Vx += std::fabs(Vx[i]);
  double Qu_error = 0.0;
  if (fail_t) *fail_t = -1;

  double A[MAXN * MAXN], B[MAXN * MAXM], Fx[MAXN * MAXN], Fu[MAXN * MAXM];
  double Qx[MAXN], Qu[MAXM], Qxx[MAXN * MAXN], Qux[MAXM * MAXN], Quu[MAXM * MAXM], Quu_reg[MAXM * MAXM];
  double VA[MAXN * MAXN], VB[MAXN * MAXM];
  double kk[MAXM], KK[MAXM * MAXN];

  for (int t = N - 1; t >= 0; --t) {
    const double *x = X + (size_t)t * n;
    const double *u = U + (size_t)t * m;
    if (Astack) {
      std::memcpy(A, Astack + (size_t)t * n * n, sizeof(double) * n * n);
      std::memcpy(B, Bstack + (size_t)t * n * m, sizeof(double) * n * m);
    } else {
      jacobians(p, x, u, t * dt, Fx, Fu); /* :113-118 */
      for (int i = 0; i < n * n; ++i) A[i] = dt * Fx[i];
      for (int i = 0; i < n; ++i) A[i * n + i] += 1.0;
      for (int i = 0; i < n * m; ++i) B[i] = dt * Fu[i];
    }
    /* cost derivatives (objective.cpp:100-149): l_x = 2 Q_ e, l_u = 2 R_ u, l_xx = 2 Q_, l_uu = 2 R_, l_ux = 0 */
    const double *ref = ref_at(p, xref, ref_traj, t);
    double e[MAXN];
    for (int i = 0; i < n; ++i) e[i] = x[i] - ref[i];
    for (int i = 0; i < n; ++i) { /* Q_x = l_x + A^T V_x (:124) */
      double lx = 0.0;
      for (int j = 0; j < n; ++j) lx += (2.0 * (p->Q[i * n + j] * dt)) * e[j];
      double s = 0.0;
      for (int j = 0; j < n; ++j) s += A[j * n + i] * Vx[j];
      Qx[i] = lx + s;
    }
    for (int i = 0; i < m; ++i) { /* Q_u = l_u + B^T V_x (:125) */
      double lu = 0.0;
      for (int j = 0; j < m; ++j) lu += (2.0 * (p->R[i * m + j] * dt)) * u[j];
      double s = 0.0;
      for (int j = 0; j < n; ++j) s += B[j * m + i] * Vx[j];
      Qu[i] = lu + s;
    }
    for (int i = 0; i < n; ++i) /* V_xx A, V_xx B */
      for (int j = 0; j < n; ++j) {
        double s = 0.0;
        for (int l = 0; l < n; ++l) s += Vxx[i * n + l] * A[l * n + j];
        VA[i * n + j] = s;
      }
    for (int i = 0; i < n; ++i)
      for (int j = 0; j < m; ++j) {
        double s = 0.0;
        for (int l = 0; l < n; ++l) s += Vxx[i * n + l] * B[l * m + j];
        VB[i * m + j] = s;
      }
    for (int i = 0; i < n; ++i) /* Q_xx = l_xx + A^T V_xx A (:126) */
      for (int j = 0; j < n; ++j) {
        double s = 0.0;
        for (int l = 0; l < n; ++l) s += A[l * n + i] * VA[l * n + j];
        Qxx[i * n + j] = 2.0 * (p->Q[i * n + j] * dt) + s;
      }
    for (int i = 0; i < m; ++i) /* Q_ux = l_ux + B^T V_xx A (:127) */
      for (int j = 0; j < n; ++j) {
        double s = 0.0;
        for (int l = 0; l < n; ++l) s += B[l * m + i] * VA[l * n + j];
        Qux[i * n + j] = s;
      }
    for (int i = 0; i < m; ++i) /* Q_uu = l_uu + B^T V_xx B (:128) */
      for (int j = 0; j < m; ++j) {
        double s = 0.0;
        for (int l = 0; l < n; ++l) s += B[l * m + i] * VB[l * m + j];
        Quu[i * m + j] = 2.0 * (p->R[i * m + j] * dt) + s;
      }
    std::memcpy(Quu_reg, Quu, sizeof(double) * m * m); /* :130-131 */
    for (int i = 0; i < m; ++i) Quu_reg[i * m + i] += reg;

    if (!(min_eigenvalue_sym(Quu_reg, m) > 0.0)) { /* :133-140 (NaN also fails) */
      if (fail_t) *fail_t = t;
      return 0;
    }

    if (!p->has_control_box) { /* :142-145 */
      double H[MAXM * MAXM];
      inverse_lu(Quu_reg, m, H);
      for (int i = 0; i < m; ++i) {
        double s = 0.0;
        for (int j = 0; j < m; ++j) s += H[i * m + j] * Qu[j];
        kk[i] = -s;
        for (int c = 0; c < n; ++c) {
          double s2 = 0.0;
          for (int j = 0; j < m; ++j) s2 += H[i * m + j] * Qux[j * n + c];
          KK[i * n + c] = -s2;
        }
      }
    } else { /* :147-178 */
      double lb[MAXM], ub[MAXM];
      for (int i = 0; i < m; ++i) {
        lb[i] = p->lb[i] - u[i];
        ub[i] = p->ub[i] - u[i];
      }
      BoxQPOut qp;
      boxqp_solve(o, m, Quu_reg, Qu, lb, ub, k + (size_t)t * m, &qp);
      if (qp.status == ORACLE_QP_HESSIAN_NOT_PD || qp.status == ORACLE_QP_NO_DESCENT) {
        if (fail_t) *fail_t = t;
        return 0;
      }
      for (int i = 0; i < m; ++i) kk[i] = qp.x[i];
      std::fill(KK, KK + m * n, 0.0);
      int free_idx[MAXM], nf = 0;
      for (int i = 0; i < m; ++i)
        if (qp.free_mask[i]) free_idx[nf++] = i;
      if (nf > 0) {
        /* K_free = -Hfree.solve(Q_ux[free,:]) (:174).  Every exit of BoxQPSolver::solve that leaves
         * a non-empty free set leaves the factor of exactly that set (the set only changes together
         * with a refactorisation, boxqp.cpp:81-111; the relative-improvement exit at :52-57 runs
         * before the set is touched).  The size guard below only protects the restatement. */
        if (qp.nfree_factor != nf) {
          double Hf[MAXM * MAXM];
          for (int i = 0; i < nf; ++i)
            for (int j = 0; j < nf; ++j) Hf[i * nf + j] = Quu_reg[free_idx[i] * m + free_idx[j]];
          qp.Hfree.compute(Hf, nf);
        }
        for (int c = 0; c < n; ++c) {
          double col[MAXM];
          for (int i = 0; i < nf; ++i) col[i] = Qux[free_idx[i] * n + c];
          qp.Hfree.solve_inplace(col);
          for (int i = 0; i < nf; ++i) KK[free_idx[i] * n + c] = -col[i];
        }
      }
    }
    std::memcpy(k + (size_t)t * m, kk, sizeof(double) * m); /* :181-182 */
    std::memcpy(K + (size_t)t * m * n, KK, sizeof(double) * m * n);

    /* dV (:184-186), unregularised Q_uu */
    double Quuk[MAXM];
    for (int i = 0; i < m; ++i) {
      double s = 0.0;
      for (int j = 0; j < m; ++j) s += Quu[i * m + j] * kk[j];
      Quuk[i] = s;
    }
    double d0 = 0.0, d1 = 0.0;
    for (int i = 0; i < m; ++i) {
      d0 += Qu[i] * kk[i];
      d1 += kk[i] * Quuk[i];
    }
    dV[0] += d0;
    dV[1] += 0.5 * d1;

    /* V_x = Q_x + K^T Q_uu k + Q_ux^T k + K^T Q_u (:188-189) */
    double nVx[MAXN];
    for (int i = 0; i < n; ++i) {
      double a = 0.0, b = 0.0, c = 0.0;
      for (int j = 0; j < m; ++j) {
        a += KK[j * n + i] * Quuk[j];
        b += Qux[j * n + i] * kk[j];
        c += KK[j * n + i] * Qu[j];
      }
      nVx[i] = Qx[i] + a + b + c;
    }
    /* V_xx = Q_xx + K^T Q_uu K + Q_ux^T K + K^T Q_ux, then symmetrise (:190-192) */
    double QuuK[MAXM * MAXN];
    for (int i = 0; i < m; ++i)
      for (int c = 0; c < n; ++c) {
        double s = 0.0;
        for (int j = 0; j < m; ++j) s += Quu[i * m + j] * KK[j * n + c];
        QuuK[i * n + c] = s;
      }
    double nV[MAXN * MAXN];
    for (int i = 0; i < n; ++i)
      for (int j = 0; j < n; ++j) {
        double a = 0.0, b = 0.0, c = 0.0;
        for (int l = 0; l < m; ++l) {
          a += KK[l * n + i] * QuuK[l * n + j];
          b += Qux[l * n + i] * KK[l * n + j];
          c += KK[l * n + i] * Qux[l * n + j];
        }
        nV[i * n + j] = Qxx[i * n + j] + a + b + c;
      }
    for (int i = 0; i < n; ++i) {
      Vx[i] = nVx[i];
      for (int j = 0; j < n; ++j) Vxx[i * n + j] = 0.5 * (nV[i * n + j] + nV[j * n + i]);
    }
    if (Vx_dbg) std::memcpy(Vx_dbg + (size_t)t * n, Vx, sizeof(double) * n);
    if (Vxx_dbg) std::memcpy(Vxx_dbg + (size_t)t * n * n, Vxx, sizeof(double) * n * n);
    double l1 = 0.0, linf = 0.0; /* :194-195 */
    for (int i = 0; i < n; ++i) l1 += std::fabs(Vx[i]);
    for (int i = 0; i < m; ++i) linf = std::max(linf, std::fabs(Qu[i]));
    norm_Vx += l1;
    Qu_error = std::max(Qu_error, linf);
  }
  double sf = o->termination_scaling_max_factor; /* :197-201 */
  sf = std::max(sf, norm_Vx / (double)(N * n)) / sf;
  *inf_du = Qu_error / sf;
  return 1;
}

/* CLDDPSolver::forwardPass (clddp_solver.cpp:215-262) */
int forward_pass(const oracle_problem *p, const oracle_options *o, const double *x0, const double *X, const double *U,
                 const double *xref, const double *ref_traj, const double *K, const double *k, const double *dV,
                 double cost, double alpha, double *Xn, double *Un, double *Jn, double *ratio_out = nullptr) {
  const int n = p->n, m = p->m, N = p->horizon;
  std::memcpy(Xn, x0, sizeof(double) * n); /* :224 */
  double J = 0.0;
  for (int t = 0; t < N; ++t) {
    const double *x = Xn + (size_t)t * n;
    double dx[MAXN];
    for (int i = 0; i < n; ++i) dx[i] = x[i] - X[(size_t)t * n + i];
    double *u = Un + (size_t)t * m;
    for (int i = 0; i < m; ++i) { /* :232-233 */
      double s = 0.0;
      for (int j = 0; j < n; ++j) s += K[((size_t)t * m + i) * n + j] * dx[j];
      u[i] = U[(size_t)t * m + i] + alpha * k[(size_t)t * m + i] + s;
    }
    if (p->has_control_box) /* :235-238, constraint.hpp:225-228 */
      for (int i = 0; i < m; ++i) u[i] = std::min(std::max(u[i], p->lb[i]), p->ub[i]);
    J += running_cost(p, x, u, ref_at(p, xref, ref_traj, t)); /* :240-241 */
    discrete_dynamics(p, x, u, t * p->dt, Xn + (size_t)(t + 1) * n);
  }
  J += terminal_cost(p, Xn + (size_t)N * n, xref);
  const double dJ = cost - J; /* :249-257 */
  const double expected = -alpha * (dV[0] + 0.5 * alpha * dV[1]);
  const double ratio = expected > 0.0 ? dJ / expected : std::copysign(1.0, dJ);
  *Jn = J;
  if (ratio_out) *ratio_out = ratio;
  return ratio > o->armijo_constant ? 1 : 0;
}

int build_alphas(const oracle_options *o, double *alphas) { /* cddp_context_utils.cpp:37-57 */
  int cnt = 0;
  double a = o->ls_initial_step_size;
  for (int i = 0; i < o->ls_max_iterations && cnt < ORACLE_MAX_ALPHAS - 1; ++i) {
    alphas[cnt++] = a;
    a *= o->ls_step_reduction_factor;
    if (a < o->ls_min_step_size && i < o->ls_max_iterations - 1) {
      alphas[cnt++] = o->ls_min_step_size;
      break;
    }
  }
  if (cnt == 0) alphas[cnt++] = o->ls_initial_step_size;
  return cnt;
}

/* CDDP::solve("CLDDP"): cddp_core.cpp:235-306 + clddp_solver.cpp:28-75 + cddp_solver_base.cpp:29-186 */
struct Replay;
void solve_one_replay(const oracle_problem *p, const oracle_options *o, const double *x0, const double *xref,
                      const double *ref_traj, double *X, double *U, double *K, double *k, oracle_result *res,
                      double *history, const Replay &rp, oracle_replay_report *rep);

/* Decision trace (test instrumentation, no reference counterpart): one int per entry of the main loop,
 * (backward failures in this iteration << 8) | code, code = ORACLE_TRACE_* or 1 + index of the accepted alpha.
 * With `replay` the solve FOLLOWS a recorded decision sequence instead of taking its own accept / reject / converged
 * decisions (all arithmetic is still the oracle's); `rep` then reports where its own verdict would have differed
 * and by how much it missed the threshold. */
struct Replay {
  const int *trace;
  int iterations;
  int status;
};

void solve_one(const oracle_problem *p, const oracle_options *o, const double *x0, const double *xref,
               const double *ref_traj, double *X, double *U, double *K, double *k, oracle_result *res,
               double *history, int *trace_out = nullptr, const Replay *replay = nullptr,
               oracle_replay_report *rep = nullptr) {
  if (replay) {
    solve_one_replay(p, o, x0, xref, ref_traj, X, U, K, k, res, history, *replay, rep);
    return;
  }
  const int n = p->n, m = p->m, N = p->horizon;
  double alphas[ORACLE_MAX_ALPHAS];
  const int na = build_alphas(o, alphas);
  /* initializeProblemIfNecessary (cddp_core.cpp:272-306): X_[0] = initial_state_, reg = initial */
  std::memcpy(X, x0, sizeof(double) * n);
  double reg = o->reg_initial_value;
  double alpha_pr = o->ls_initial_step_size;
  /* CLDDPSolver::initialize cold start (clddp_solver.cpp:68-74) */
  std::fill(K, K + (size_t)N * m * n, 0.0);
  std::fill(k, k + (size_t)N * m, 0.0);
  double cost = trajectory_cost(p, X, U, xref, ref_traj);
  double inf_du = std::numeric_limits<double>::infinity();
  double dV[2] = {0.0, 0.0};
  std::vector<double> Xn((size_t)(N + 1) * n), Un((size_t)N * m);
  int hl = 0;
  auto record = [&]() {
    if (history) {
      history[hl * 4 + 0] = cost;
      history[hl * 4 + 1] = alpha_pr;
      history[hl * 4 + 2] = inf_du;
      history[hl * 4 + 3] = reg;
      ++hl;
    }
  };
  record();
  const auto start = std::chrono::high_resolution_clock::now();
  int iter = 0;
  int status = ORACLE_MAX_ITERATIONS;
  bool converged = false;
  while (iter < o->max_iterations) {
    ++iter;
    if (o->max_cpu_time > 0) { /* cddp_solver_base.cpp:77-90 */
      auto el = std::chrono::duration_cast<std::chrono::milliseconds>(std::chrono::high_resolution_clock::now() - start);
      if (el.count() > o->max_cpu_time * 1000) {
        status = ORACLE_MAX_CPU_TIME;
        break;
      }
    }
    bool backward_ok = false; /* :93-111 */
    int bw_fail = 0;
    while (!backward_ok) {
      backward_ok = backward_pass(p, o, nullptr, nullptr, X, U, xref, ref_traj, reg, K, k, dV, &inf_du, nullptr,
                                  nullptr, nullptr) != 0;
      if (!backward_ok) {
        ++bw_fail;
        reg = std::min(reg * o->reg_update_factor, o->reg_max_value); /* cddp_core.cpp:308-314 */
        if (reg >= o->reg_max_value) {
          status = ORACLE_REG_LIMIT;
          break;
        }
      }
    }
    if (trace_out) trace_out[iter - 1] = (bw_fail << 8) | (backward_ok ? ORACLE_TRACE_LS_FAILED : ORACLE_TRACE_BW_LIMIT);
    if (!backward_ok) break;
    if (inf_du < o->tolerance) { /* clddp_solver.cpp:206-213 */
      status = ORACLE_OPTIMAL;
      converged = true;
      record();
      if (trace_out) trace_out[iter - 1] = (bw_fail << 8) | ORACLE_TRACE_EARLY_EXIT;
      break;
    }
    bool fp_success = false; /* cddp_solver_base.cpp:255-263: first success wins */
    double Jn = 0.0, a_acc = 0.0;
    int a_idx = -1;
    if (!o->enable_parallel) {
      for (int ai = 0; ai < na; ++ai) {
        if (forward_pass(p, o, x0, X, U, xref, ref_traj, K, k, dV, cost, alphas[ai], Xn.data(), Un.data(), &Jn)) {
          fp_success = true;
          a_acc = alphas[ai];
          a_idx = ai;
          break;
        }
      }
    } else { /* enable_parallel: lowest merit among successes, :264-285 */
      std::vector<double> Xt((size_t)(N + 1) * n), Ut((size_t)N * m);
      double best = std::numeric_limits<double>::infinity();
      for (int ai = 0; ai < na; ++ai) {
        double Jt = 0.0;
        if (forward_pass(p, o, x0, X, U, xref, ref_traj, K, k, dV, cost, alphas[ai], Xt.data(), Ut.data(), &Jt) &&
            Jt < best) {
          best = Jt;
          fp_success = true;
          a_acc = alphas[ai];
          a_idx = ai;
          Jn = Jt;
          Xn = Xt;
          Un = Ut;
        }
      }
    }
    if (trace_out && fp_success) trace_out[iter - 1] = (bw_fail << 8) | (1 + a_idx);
    if (fp_success) { /* :129-139 */
      const double dJ = cost - Jn;
      std::memcpy(X, Xn.data(), sizeof(double) * (N + 1) * n);
      std::memcpy(U, Un.data(), sizeof(double) * N * m);
      cost = Jn;
      alpha_pr = a_acc;
      record();
      reg = std::max(reg / o->reg_update_factor, o->reg_min_value); /* cddp_core.cpp:316-322 */
      if (inf_du < o->tolerance) { /* clddp_solver.cpp:264-277 */
        status = ORACLE_OPTIMAL;
        converged = true;
      } else if (dJ > 0.0 && dJ < o->acceptable_tolerance) {
        status = ORACLE_ACCEPTABLE;
        converged = true;
      }
    } else { /* cddp_solver_base.cpp:206-218 */
      reg = std::min(reg * o->reg_update_factor, o->reg_max_value);
      if (reg >= o->reg_max_value) {
        status = ORACLE_REG_LIMIT;
        break;
      }
    }
    if (converged) break;
  }
  res->final_objective = cost;
  res->final_step_length = alpha_pr;
  res->final_regularization = reg;
  res->inf_du = inf_du;
  res->iterations = iter;
  res->status = status;
  res->history_len = hl;
  res->reserved = 0;
}

/* ONE entry of the main loop of CDDPSolverBase::solve (cddp_solver_base.cpp:74-170) from a given solver state —
 * test instrumentation: the same statements as the loop body of solve_one, with two additions.
 *   (1) `follow` >= 0: a recorded decision ((backward failures << 8) | code, see cddp_oracle.h) and, in
 *       `follow_status`, the recorded status after this iteration (ORACLE_RUNNING = the solve went on).  Every `if`
 *       on a line-search / convergence / backward-failure verdict then takes the RECORDED branch; all arithmetic is
 *       still the oracle's.
 *   (2) own verdicts are always evaluated; where one differs from the recorded branch, rep counts it and keeps the
 *       largest MARGIN: how far the tested quantity was from its threshold, measured in the units in which roundoff
 *       enters it — |dJ - c*expected| / |cost| for the Armijo test (clddp_solver.cpp:251-257), the same for
 *       0 < dJ < acceptable_tolerance (:272-275), |inf_du - tol| / tol for the inf_du tests (:206-213, :268-271).
 * Returns the decision taken in *code_out and the status after the iteration in *status_out. */
struct IterState {
  double *X, *U, *K, *k;
  double reg, cost, alpha_pr, inf_du;
  double dV[2];
};

void iterate_once(const oracle_problem *p, const oracle_options *o, const double *alphas, int na, const double *x0,
                  const double *xref, const double *ref_traj, IterState &st, int follow, int follow_status,
                  int *code_out, int *status_out, double *reg_used_out, oracle_replay_report &R, std::vector<double> &Xn,
                  std::vector<double> &Un, std::vector<double> &Xt, std::vector<double> &Ut) {
  const int n = p->n, m = p->m, N = p->horizon;
  const bool rec = follow >= 0;
  const int rcode = rec ? (follow & 0xff) : 0, rfail = rec ? (follow >> 8) : 0;
  auto margin = [&](double dist, double scale) {
    ++R.n_disagree;
    const double mg = std::fabs(dist) / std::max(std::fabs(scale), 1e-300);
    if (!(mg <= R.max_margin)) R.max_margin = mg; /* NaN sticks */
  };
  *status_out = ORACLE_RUNNING;
  *reg_used_out = st.reg;
  /* backward pass with regularisation retry (:93-111) */
  bool ok = false;
  int bw_fail = 0;
  while (true) {
    if (rec && bw_fail == rfail && rcode == ORACLE_TRACE_BW_LIMIT) break;
    ok = backward_pass(p, o, nullptr, nullptr, st.X, st.U, xref, ref_traj, st.reg, st.K, st.k, st.dV, &st.inf_du, nullptr,
                       nullptr, nullptr) != 0;
    if (rec) {
      if (bw_fail < rfail) {
        if (ok) ++R.n_backward_disagree; /* recorded: failure; own: success */
        ok = false;
      } else if (!ok) {
        ++R.n_backward_disagree; /* recorded: success; own: failure -> the sequence cannot be followed */
        R.infeasible = 1;
        break;
      }
    }
    if (ok) break;
    ++bw_fail;
    st.reg = std::min(st.reg * o->reg_update_factor, o->reg_max_value); /* cddp_core.cpp:308-314 */
    if (!rec && st.reg >= o->reg_max_value) break;
  }
  if (!ok) {
    *code_out = (bw_fail << 8) | ORACLE_TRACE_BW_LIMIT;
    *status_out = ORACLE_REG_LIMIT;
    return;
  }
  *reg_used_out = st.reg; /* the regularisation of the successful sweep (what recordIterationHistory sees) */
  const bool own_early = st.inf_du < o->tolerance; /* clddp_solver.cpp:206-213 */
  const bool early = rec ? rcode == ORACLE_TRACE_EARLY_EXIT : own_early;
  if (own_early != early) margin(st.inf_du - o->tolerance, o->tolerance);
  if (early) {
    *code_out = (bw_fail << 8) | ORACLE_TRACE_EARLY_EXIT;
    *status_out = ORACLE_OPTIMAL;
    return;
  }
  /* line search (cddp_solver_base.cpp:255-285) */
  int acc = -1;
  double Jn = 0.0;
  auto armijo_distance = [&](double Jt, double alpha) { /* distance of dJ to the acceptance threshold */
    const double dJ = st.cost - Jt, expected = -alpha * (st.dV[0] + 0.5 * alpha * st.dV[1]);
    return expected > 0.0 ? dJ - o->armijo_constant * expected : dJ;
  };
  if (!o->enable_parallel) {
    const int racc = rcode - 1;
    const int upto = rec ? (racc >= 0 ? racc : na - 1) : na - 1;
    for (int ai = 0; ai <= upto; ++ai) {
      double Jt = 0.0;
      const bool s = forward_pass(p, o, x0, st.X, st.U, xref, ref_traj, st.K, st.k, st.dV, st.cost, alphas[ai], Xt.data(),
                                  Ut.data(), &Jt) != 0;
      const bool take = rec ? ai == racc : s;
      if (s != take) margin(armijo_distance(Jt, alphas[ai]), st.cost);
      if (take) {
        acc = ai;
        Jn = Jt;
        Xn.swap(Xt);
        Un.swap(Ut);
        break;
      }
    }
  } else { /* lowest cost among the accepted candidates (:264-285) */
    double best = std::numeric_limits<double>::infinity();
    int own = -1;
    const int racc = rcode - 1;
    for (int ai = 0; ai < na; ++ai) {
      double Jt = 0.0;
      const bool s = forward_pass(p, o, x0, st.X, st.U, xref, ref_traj, st.K, st.k, st.dV, st.cost, alphas[ai], Xt.data(),
                                  Ut.data(), &Jt) != 0;
      if (s && Jt < best) {
        best = Jt;
        own = ai;
      }
      if (rec ? ai == racc : (own == ai)) {
        acc = ai;
        Jn = Jt;
        Xn.swap(Xt);
        Un.swap(Ut);
      }
    }
    if (rec && own != racc) margin(racc >= 0 && own >= 0 ? Jn - best : st.cost, st.cost);
  }
  *code_out = (bw_fail << 8) | (acc >= 0 ? 1 + acc : ORACLE_TRACE_LS_FAILED);
  if (acc >= 0) { /* :129-139 */
    const double dJ = st.cost - Jn;
    std::memcpy(st.X, Xn.data(), sizeof(double) * (N + 1) * n);
    std::memcpy(st.U, Un.data(), sizeof(double) * N * m);
    st.cost = Jn;
    st.alpha_pr = alphas[acc];
    st.reg = std::max(st.reg / o->reg_update_factor, o->reg_min_value); /* cddp_core.cpp:316-322 */
    const bool own_opt = st.inf_du < o->tolerance; /* clddp_solver.cpp:264-277 */
    const bool own_acc = !own_opt && dJ > 0.0 && dJ < o->acceptable_tolerance;
    const bool r_opt = rec ? follow_status == ORACLE_OPTIMAL : own_opt;
    const bool r_acc = rec ? follow_status == ORACLE_ACCEPTABLE : own_acc;
    if (own_opt != r_opt) margin(st.inf_du - o->tolerance, o->tolerance);
    else if (own_acc != r_acc)
      margin(std::min(std::fabs(dJ), std::fabs(dJ - o->acceptable_tolerance)), st.cost);
    if (r_opt) *status_out = ORACLE_OPTIMAL;
    else if (r_acc) *status_out = ORACLE_ACCEPTABLE;
  } else { /* handleForwardPassFailure, cddp_solver_base.cpp:206-218 */
    st.reg = std::min(st.reg * o->reg_update_factor, o->reg_max_value);
    if (rec ? follow_status == ORACLE_REG_LIMIT : st.reg >= o->reg_max_value) *status_out = ORACLE_REG_LIMIT;
  }
}

/* CDDP::solve("CLDDP") driven by a recorded decision sequence: initialisation as solve_one, then iterate_once per
 * recorded entry. */
void solve_one_replay(const oracle_problem *p, const oracle_options *o, const double *x0, const double *xref,
                      const double *ref_traj, double *X, double *U, double *K, double *k, oracle_result *res,
                      double *history, const Replay &rp, oracle_replay_report *rep) {
  const int n = p->n, m = p->m, N = p->horizon;
  double alphas[ORACLE_MAX_ALPHAS];
  const int na = build_alphas(o, alphas);
  std::memcpy(X, x0, sizeof(double) * n);
  std::fill(K, K + (size_t)N * m * n, 0.0);
  std::fill(k, k + (size_t)N * m, 0.0);
  IterState st{X, U, K, k, o->reg_initial_value, 0.0, o->ls_initial_step_size, std::numeric_limits<double>::infinity(), {0.0, 0.0}};
  st.cost = trajectory_cost(p, X, U, xref, ref_traj);
  std::vector<double> Xn((size_t)(N + 1) * n), Un((size_t)N * m), Xt((size_t)(N + 1) * n), Ut((size_t)N * m);
  oracle_replay_report R;
  std::memset(&R, 0, sizeof(R));
  int hl = 0;
  auto record = [&](double reg) {
    if (history) {
      history[hl * 4 + 0] = st.cost;
      history[hl * 4 + 1] = st.alpha_pr;
      history[hl * 4 + 2] = st.inf_du;
      history[hl * 4 + 3] = reg;
      ++hl;
    }
  };
  record(st.reg);
  int iter = 0;
  while (iter < rp.iterations && !R.infeasible) {
    ++iter;
    const bool last = iter == rp.iterations;
    if (last && rp.status == ORACLE_MAX_CPU_TIME) break;
    int code = 0, status = 0;
    double reg_used = st.reg;
    iterate_once(p, o, alphas, na, x0, xref, ref_traj, st, rp.trace[iter - 1], last ? rp.status : ORACLE_RUNNING, &code,
                 &status, &reg_used, R, Xn, Un, Xt, Ut);
    const int c = code & 0xff;
    if (c == ORACLE_TRACE_EARLY_EXIT || (c >= 1 && c < ORACLE_TRACE_BW_LIMIT)) record(reg_used); /* :116-118, :132-135 */
  }
  res->final_objective = st.cost;
  res->final_step_length = st.alpha_pr;
  res->final_regularization = st.reg;
  res->inf_du = st.inf_du;
  res->iterations = iter;
  res->status = rp.status;
  res->history_len = hl;
  res->reserved = 0;
  if (rep) *rep = R;
}

/* ==========================================================================================
 * IPDDP (src/cddp_core/ipddp_solver.cpp) — cold start (options.warm_start = false), use_ilqr = true
 * (the default, options.hpp:223: no dynamics-Hessian terms), path INEQUALITY constraints of the types
 * ControlConstraint / StateConstraint (constraint.hpp:144-251), LinearConstraint (:253-318) and
 * BallConstraint (:320-440); no terminal constraints (branches 1 and 3 of IPDDPSolver::backwardPass:
 * :1055-1118 unconstrained, :1355-1569 path constraints).  The costate bookkeeping (Lambda_, k_lambda_,
 * K_lambda_) only feeds an allFinite() test in the reference (:1612-1618, :1665-1671) and is not carried.
 * ======================================================================================== */
constexpr int MAXDUAL = 64; /* total dual dimension capacity */
constexpr double kSlackInteriorOffset = 1e-4; /* ipddp_solver.cpp:35-38 */
constexpr double EPS_SLACK = 1e-10;
constexpr double MAX_BARRIER_RATIO = 1e6;

inline double clampd(double v, double lo, double hi) { return v < lo ? lo : (hi < v ? hi : v); } /* std::clamp */
inline double clip_pos(double num, double den) { return clampd(num / den, 0.0, MAX_BARRIER_RATIO); }      /* :222-225 */
inline double clip_signed(double num, double den) { return clampd(num / den, -MAX_BARRIER_RATIO, MAX_BARRIER_RATIO); } /* :227-231 */

int constraint_dim(const oracle_problem *p, const oracle_constraint &c) {
  switch (c.type) {
    case ORACLE_CON_CONTROL_BOX: return 2 * p->m; /* constraint.hpp:156 */
    case ORACLE_CON_STATE_BOX: return 2 * p->n;
    case ORACLE_CON_BALL: return 1;               /* :324 */
    default: return c.rows;                       /* LinearConstraint :261 */
  }
}
int total_dual_dim(const oracle_problem *p, const oracle_constraint *cs, int nc) { /* ipddp_solver.cpp:2134-2143 */
  int d = 0;
  for (int i = 0; i < nc; ++i) d += constraint_dim(p, cs[i]);
  return d;
}

/* g = constraint.evaluate(x,u) - constraint.getUpperBound() for the stacked set, in the caller's order (the
 * reference iterates a std::map keyed by constraint name, i.e. alphabetical order) (:2273-2278) */
void eval_constraints(const oracle_problem *p, const oracle_constraint *cs, int nc, const double *x, const double *u,
                      double *g) {
  int o = 0;
  for (int c = 0; c < nc; ++c) {
    const oracle_constraint &C = cs[c];
    if (C.type == ORACLE_CON_CONTROL_BOX || C.type == ORACLE_CON_STATE_BOX) { /* constraint.hpp:156-180 */
      const int k = C.type == ORACLE_CON_CONTROL_BOX ? p->m : p->n;
      const double *v = C.type == ORACLE_CON_CONTROL_BOX ? u : x;
      for (int i = 0; i < k; ++i) {
        g[o + i] = (-v[i]) * C.scale - (-C.p0[i] * C.scale);
        g[o + k + i] = v[i] * C.scale - C.p1[i] * C.scale;
      }
      o += 2 * k;
    } else if (C.type == ORACLE_CON_BALL) { /* :326-343 */
      double sq = 0.0;
      for (int i = 0; i < C.rows; ++i) sq += (x[i] - C.p0[i]) * (x[i] - C.p0[i]);
      g[o] = -(C.scale * sq) - (-(C.p1[0] * C.p1[0]) * C.scale);
      o += 1;
    } else { /* LinearConstraint: A x - b (scale_factor is NOT applied by evaluate, :263-268) */
      for (int r = 0; r < C.rows; ++r) {
        double s = 0.0;
        for (int j = 0; j < p->n; ++j) s += C.p0[r * p->n + j] * x[j];
        g[o + r] = s - C.p1[r];
      }
      o += C.rows;
    }
  }
}

/* getStateJacobian / getControlJacobian of the stacked set: Gx [d][n], Gu [d][m] */
void constraint_jacobians(const oracle_problem *p, const oracle_constraint *cs, int nc, const double *x, double *Gx,
                          double *Gu) {
  const int n = p->n, m = p->m;
  int o = 0;
  for (int c = 0; c < nc; ++c) {
    const oracle_constraint &C = cs[c];
    const int dim = constraint_dim(p, C);
    for (int r = 0; r < dim; ++r) {
      for (int j = 0; j < n; ++j) Gx[(o + r) * n + j] = 0.0;
      for (int j = 0; j < m; ++j) Gu[(o + r) * m + j] = 0.0;
    }
    if (C.type == ORACLE_CON_CONTROL_BOX) { /* constraint.hpp:201-217 */
      for (int i = 0; i < m; ++i) {
        Gu[(o + i) * m + i] = -1.0 * C.scale;
        Gu[(o + m + i) * m + i] = 1.0 * C.scale;
      }
    } else if (C.type == ORACLE_CON_STATE_BOX) { /* :183-199 */
      for (int i = 0; i < n; ++i) {
        Gx[(o + i) * n + i] = -1.0 * C.scale;
        Gx[(o + n + i) * n + i] = 1.0 * C.scale;
      }
    } else if (C.type == ORACLE_CON_BALL) { /* :360-373 */
      for (int i = 0; i < C.rows; ++i) Gx[o * n + i] = -2.0 * C.scale * (x[i] - C.p0[i]);
    } else { /* :278-284 */
      for (int r = 0; r < C.rows; ++r)
        for (int j = 0; j < n; ++j) Gx[(o + r) * n + j] = C.p0[r * n + j];
    }
    o += dim;
  }
}

struct FilterPt { double merit, theta; };
inline bool dominates(const FilterPt &a, const FilterPt &b) { return a.merit <= b.merit && a.theta <= b.theta; } /* cddp_core.hpp:171-174 */
bool accept_filter_entry(std::vector<FilterPt> &f, double merit, double theta) { /* interior_point_utils.cpp:81-97 */
  const FilterPt cand{merit, theta};
  for (const auto &q : f)
    if (dominates(q, cand)) return false;
  f.erase(std::remove_if(f.begin(), f.end(), [&](const FilterPt &q) { return dominates(cand, q); }), f.end());
  f.push_back(cand);
  return true;
}
void prune_filter(std::vector<FilterPt> &f) { /* interior_point_utils.cpp:116-141 */
  if (f.empty()) return;
  FilterPt bv = f[0], bm = f[0];
  for (const auto &q : f) { /* std::min_element: first minimal element */
    if (q.theta < bv.theta) bv = q;
    if (q.merit < bm.merit) bm = q;
  }
  f.clear();
  f.push_back(bv);
  if (std::fabs(bm.theta - bv.theta) > 1e-12 || std::fabs(bm.merit - bv.merit) > 1e-12) f.push_back(bm);
}

struct IpState {
  int n, m, N, d, nc;
  const oracle_problem *p;
  const oracle_options *o;
  const oracle_ipddp_options *io;
  const oracle_constraint *cs;
  const double *x0, *xref, *ref_traj;
  std::vector<double> X, U, Y, S, G;                     /* nominal trajectory, duals, slacks, constraint values */
  std::vector<double> ku, Ku, ky, Ky, ks, Ks, dS, dY;   /* gains and linearised steps */
  std::vector<double> A, B;                              /* F_x_, F_u_ (cddp_solver_base.cpp:319-345) */
  double mu = 0, cost = 0, merit = 0, phi = 0, theta = 0, filter_theta = 0;
  double inf_pr = 0, inf_du = 0, inf_comp = 0, step_norm = 0, reg = 0, alpha_pr = 1, alpha_du = 1;
  double dV[2] = {0, 0};
  std::vector<FilterPt> filter;
  /* TerminalEqualityConstraint(target = reference state): h(x_N) = x_N - xref, Jacobian I (terminal_constraint.hpp:61-117) */
  bool teq = false;
  std::vector<double> lamT, dlamT; /* Lambda_T_eq_, dLambda_T_eq_ */
  std::vector<double> kvar, pvar;  /* k / p of the p+1 sequential-LQR variants */
};

double ip_theta(const IpState &s, const double *G, const double *S, const double *hT = nullptr) { /* computeTheta :2778-2848 */
  const bool l2 = s.io->theta_norm_l2 != 0;
  double total = 0.0, max_entry = 0.0;
  for (int c = 0, o = 0; c < s.nc; ++c) { /* per constraint, per t, as the reference's map-of-trajectories loops */
    const int dim = constraint_dim(s.p, s.cs[c]);
    for (int t = 0; t < s.N; ++t) {
      double acc = 0.0, mx = 0.0;
      for (int i = 0; i < dim; ++i) {
        const double r = G[(size_t)t * s.d + o + i] + S[(size_t)t * s.d + o + i];
        acc += l2 ? r * r : std::fabs(r);
        mx = std::max(mx, std::fabs(r));
      }
      total += acc;
      max_entry = std::max(max_entry, mx);
    }
    o += dim;
  }
  if (hT) { /* :2833-2845 */
    double acc = 0.0, mx = 0.0;
    for (int i = 0; i < s.n; ++i) {
      acc += l2 ? hT[i] * hT[i] : std::fabs(hT[i]);
      mx = std::max(mx, std::fabs(hT[i]));
    }
    total += acc;
    max_entry = std::max(max_entry, mx);
  }
  const double th = l2 ? std::sqrt(total) : total;
  return std::max(th, max_entry);
}
double ip_merit(const IpState &s, const double *S, double cost, const double *lamT = nullptr, const double *hT = nullptr) { /* computeBarrierMerit :2850-2880 */
  double merit = cost;
  for (int c = 0, o = 0; c < s.nc; ++c) {
    const int dim = constraint_dim(s.p, s.cs[c]);
    for (int t = 0; t < s.N; ++t) {
      double acc = 0.0;
      for (int i = 0; i < dim; ++i) acc += std::log(std::max(S[(size_t)t * s.d + o + i], EPS_SLACK));
      merit -= s.mu * acc;
    }
    o += dim;
  }
  if (lamT && hT) { /* :2872-2878 */
    double dot = 0.0;
    for (int i = 0; i < s.n; ++i) dot += lamT[i] * hT[i];
    merit += dot;
  }
  return merit;
}
void ip_primal_comp(const IpState &s, const double *G, const double *S, const double *Y, double mu, double *inf_pr,
                    double *inf_comp, const double *hT = nullptr) { /* computePrimalAndComplementarity :2882-2937 */
  double ip = 0.0, ic = 0.0;
  for (size_t i = 0; i < (size_t)s.N * s.d; ++i) {
    ip = std::max(ip, std::fabs(G[i] + S[i]));
    ic = std::max(ic, std::fabs(Y[i] * S[i] - mu));
  }
  if (hT)
    for (int i = 0; i < s.n; ++i) ip = std::max(ip, std::fabs(hT[i]));
  *inf_pr = ip;
  *inf_comp = ic;
}
inline void ip_terminal_residual(const IpState &s, const double *X, double *hT) { /* evaluateTerminalEqualityResidual :150-176 */
  for (int i = 0; i < s.n; ++i) hT[i] = X[(size_t)s.N * s.n + i] - s.xref[i];
}
void ip_reset_filter(IpState &s) { /* resetBarrierFilter :2484-2517 */
  double hT[MAXN];
  if (s.teq) ip_terminal_residual(s, s.X.data(), hT);
  const double *h = s.teq ? hT : nullptr;
  ip_primal_comp(s, s.G.data(), s.S.data(), s.Y.data(), s.mu, &s.inf_pr, &s.inf_comp, h);
  s.merit = ip_merit(s, s.S.data(), s.cost, s.teq ? s.lamT.data() : nullptr, h);
  s.phi = s.merit;
  s.filter_theta = std::max(ip_theta(s, s.G.data(), s.S.data(), h), 1e-8);
  s.theta = std::max(s.filter_theta, std::max(s.io->theta_0_floor, 1e-8));
  s.filter.clear();
  if (s.teq) accept_filter_entry(s.filter, s.phi, s.filter_theta); /* :2513-2516 */
}

void ip_initialize(IpState &s, const double *U0) { /* IPDDPSolver::initialize cold start :818-913 */
  const int n = s.n, m = s.m, N = s.N, d = s.d;
  s.X.assign((size_t)(N + 1) * n, 0.0);
  s.U.assign((size_t)N * m, 0.0);
  if (U0) std::memcpy(s.U.data(), U0, sizeof(double) * N * m);
  s.ku.assign((size_t)N * m, 0.0);
  s.Ku.assign((size_t)N * m * n, 0.0);
  s.Y.assign((size_t)N * d, 0.0); s.S.assign((size_t)N * d, 0.0); s.G.assign((size_t)N * d, 0.0);
  s.ky.assign((size_t)N * d, 0.0); s.Ky.assign((size_t)N * d * n, 0.0);
  s.ks.assign((size_t)N * d, 0.0); s.Ks.assign((size_t)N * d * n, 0.0);
  s.dS.assign((size_t)N * d, 0.0); s.dY.assign((size_t)N * d, 0.0);
  s.A.assign((size_t)N * n * n, 0.0); s.B.assign((size_t)N * n * m, 0.0);
  std::memcpy(s.X.data(), s.x0, sizeof(double) * n);
  for (int t = 0; t < N; ++t) /* :876-882 rollout of the given controls */
    discrete_dynamics(s.p, &s.X[(size_t)t * n], &s.U[(size_t)t * m], t * s.p->dt, &s.X[(size_t)(t + 1) * n]);
  s.teq = s.io->terminal_equality != 0;
  s.lamT.assign(n, 0.0);
  s.dlamT.assign(n, 0.0);
  s.mu = (s.nc == 0 && !s.teq) ? std::max(s.o->tolerance / 10.0, s.io->mu_min_value) : s.io->mu_initial; /* :884-887 */
  s.reg = s.o->reg_initial_value;
  s.step_norm = 0.0;
  s.alpha_pr = 1.0;
  s.alpha_du = 1.0;
  /* evaluateTrajectory (:2252-2296) + initializeDualSlackVariables (:2428-2482) */
  for (int t = 0; t < N; ++t) {
    double *g = &s.G[(size_t)t * d];
    eval_constraints(s.p, s.cs, s.nc, &s.X[(size_t)t * n], &s.U[(size_t)t * m], g);
    for (int i = 0; i < d; ++i) {
      const double si = std::max(s.io->slack_var_init_scale, -g[i] + kSlackInteriorOffset);
      s.S[(size_t)t * d + i] = si;
      s.Y[(size_t)t * d + i] = (s.mu * s.io->dual_var_init_scale) / std::max(si, EPS_SLACK);
    }
  }
  s.cost = trajectory_cost(s.p, s.X.data(), s.U.data(), s.xref, s.ref_traj);
  ip_reset_filter(s);
  s.inf_du = 0.0;
  s.dV[0] = s.dV[1] = 0.0;
}

/* solveSequentialLQR (ipddp_solver.cpp:411-482) for ONE right-hand side: Q,q (N+1), R,r,M (N), d = 0.  K [N][m][n] and
 * P [N+1][n][n] do not depend on q/r, but the reference recomputes them per variant; so does this restatement. */
bool sequential_lqr(int n, int m, int N, const std::vector<double> &Q, const std::vector<double> &q, const std::vector<double> &R,
                    const std::vector<double> &r, const std::vector<double> &M, const std::vector<double> &A,
                    const std::vector<double> &B, std::vector<double> &K, std::vector<double> &k, std::vector<double> &P,
                    std::vector<double> &pp) {
  K.assign((size_t)N * m * n, 0.0);
  k.assign((size_t)N * m, 0.0);
  P.assign((size_t)(N + 1) * n * n, 0.0);
  pp.assign((size_t)(N + 1) * n, 0.0);
  for (int i = 0; i < n; ++i) {
    for (int j = 0; j < n; ++j) P[(size_t)N * n * n + i * n + j] = 0.5 * (Q[(size_t)N * n * n + i * n + j] + Q[(size_t)N * n * n + j * n + i]);
    pp[(size_t)N * n + i] = q[(size_t)N * n + i];
  }
  LDLT ldlt;
  for (int t = N - 1; t >= 0; --t) {
    const double *Pn = &P[(size_t)(t + 1) * n * n], *pn = &pp[(size_t)(t + 1) * n];
    const double *At = &A[(size_t)t * n * n], *Bt = &B[(size_t)t * n * m];
    const double *Qt = &Q[(size_t)t * n * n], *Rt = &R[(size_t)t * m * m], *Mt = &M[(size_t)t * n * m];
    double BtP[MAXM * MAXN], Quu[MAXM * MAXM], Qux[MAXM * MAXN], Qx[MAXN], Qu[MAXM];
    for (int i = 0; i < m; ++i)
      for (int j = 0; j < n; ++j) {
        double a = 0.0;
        for (int l = 0; l < n; ++l) a += Bt[l * m + i] * Pn[l * n + j];
        BtP[i * n + j] = a;
      }
    for (int i = 0; i < m; ++i)
      for (int j = 0; j < m; ++j) {
        double a1 = 0.0, a2 = 0.0; /* BtP B and B^T P^T B */
        for (int l = 0; l < n; ++l) a1 += BtP[i * n + l] * Bt[l * m + j];
        for (int l = 0; l < n; ++l) {
          double c = 0.0;
          for (int q2 = 0; q2 < n; ++q2) c += Bt[q2 * m + i] * Pn[l * n + q2];
          a2 += c * Bt[l * m + j];
        }
        Quu[i * m + j] = 0.5 * (((Rt[i * m + j] + a1) + Rt[j * m + i]) + a2);
      }
    for (int i = 0; i < m; ++i)
      for (int j = 0; j < n; ++j) {
        double a = 0.0;
        for (int l = 0; l < n; ++l) a += BtP[i * n + l] * At[l * n + j];
        Qux[i * n + j] = a + Mt[j * m + i];
      }
    for (int i = 0; i < n; ++i) { /* drift = p_next (d = 0) */
      double a = 0.0;
      for (int l = 0; l < n; ++l) a += At[l * n + i] * pn[l];
      Qx[i] = q[(size_t)t * n + i] + a;
    }
    for (int i = 0; i < m; ++i) {
      double a = 0.0;
      for (int l = 0; l < n; ++l) a += Bt[l * m + i] * pn[l];
      Qu[i] = r[(size_t)t * m + i] + a;
    }
    ldlt.compute(Quu, m);
    if (!ldlt.ok) return false;
    double *Kt = &K[(size_t)t * m * n], *kt = &k[(size_t)t * m];
    double col[MAXM];
    for (int j = 0; j < n; ++j) {
      for (int i = 0; i < m; ++i) col[i] = Qux[i * n + j];
      ldlt.solve_inplace(col);
      for (int i = 0; i < m; ++i) Kt[i * n + j] = -col[i];
    }
    for (int i = 0; i < m; ++i) col[i] = Qu[i];
    ldlt.solve_inplace(col);
    for (int i = 0; i < m; ++i) kt[i] = -col[i];
    double APA[MAXN * MAXN], PA[MAXN * MAXN], QuuK[MAXM * MAXN], Quuk[MAXM];
    for (int i = 0; i < n; ++i)
      for (int j = 0; j < n; ++j) {
        double a = 0.0;
        for (int l = 0; l < n; ++l) a += Pn[i * n + l] * At[l * n + j];
        PA[i * n + j] = a;
      }
    for (int i = 0; i < n; ++i)
      for (int j = 0; j < n; ++j) {
        double a = 0.0;
        for (int l = 0; l < n; ++l) a += At[l * n + i] * PA[l * n + j];
        APA[i * n + j] = a;
      }
    for (int i = 0; i < m; ++i) {
      for (int j = 0; j < n; ++j) {
        double a = 0.0;
        for (int l = 0; l < m; ++l) a += Quu[i * m + l] * Kt[l * n + j];
        QuuK[i * n + j] = a;
      }
      double a = 0.0;
      for (int l = 0; l < m; ++l) a += Quu[i * m + l] * kt[l];
      Quuk[i] = a;
    }
    double Pt[MAXN * MAXN];
    for (int i = 0; i < n; ++i)
      for (int j = 0; j < n; ++j) { /* P = Q + A^T P A + Q_xu K + K^T Q_ux + K^T Q_uu K */
        double a1 = 0.0, a2 = 0.0, a3 = 0.0;
        for (int l = 0; l < m; ++l) {
          a1 += Qux[l * n + i] * Kt[l * n + j];
          a2 += Kt[l * n + i] * Qux[l * n + j];
          a3 += Kt[l * n + i] * QuuK[l * n + j];
        }
        Pt[i * n + j] = (((Qt[i * n + j] + APA[i * n + j]) + a1) + a2) + a3;
      }
    bool finite = true;
    for (int i = 0; i < n; ++i) {
      for (int j = 0; j < n; ++j) {
        const double v = 0.5 * (Pt[i * n + j] + Pt[j * n + i]);
        P[(size_t)t * n * n + i * n + j] = v;
        finite = finite && std::isfinite(v);
      }
      double a1 = 0.0, a2 = 0.0, a3 = 0.0; /* p = Q_x + Q_xu k + K^T Q_u + K^T Q_uu k */
      for (int l = 0; l < m; ++l) {
        a1 += Qux[l * n + i] * kt[l];
        a2 += Kt[l * n + i] * Qu[l];
        a3 += Kt[l * n + i] * Quuk[l];
      }
      pp[(size_t)t * n + i] = ((Qx[i] + a1) + a2) + a3;
      finite = finite && std::isfinite(pp[(size_t)t * n + i]);
    }
    for (int i = 0; i < m * n; ++i) finite = finite && std::isfinite(Kt[i]);
    for (int i = 0; i < m; ++i) finite = finite && std::isfinite(kt[i]);
    if (!finite) return false;
  }
  return true;
}

/* IPDDPSolver::backwardPass, terminal-equality branch (:1120-1353) with solveTerminalEqualityLQR (:484-639); H_T = I,
 * b_T = -h_T, dx0 = 0, d = 0. */
bool ip_backward_teq(IpState &s, const double *Vx, const double *Vxx) {
  const int n = s.n, m = s.m, N = s.N, d = s.d, pd = s.n;
  const oracle_problem *p = s.p;
  const double dt = p->dt;
  std::vector<double> Q((size_t)(N + 1) * n * n, 0.0), q((size_t)(N + 1) * n, 0.0), R((size_t)N * m * m, 0.0), r((size_t)N * m, 0.0),
      M((size_t)N * n * m, 0.0);
  for (int i = 0; i < n * n; ++i) Q[(size_t)N * n * n + i] = Vxx[i];
  for (int i = 0; i < n; ++i) q[(size_t)N * n + i] = Vx[i];
  double hT[MAXN];
  ip_terminal_residual(s, s.X.data(), hT);
  double inf_pr = 0.0, inf_comp = 0.0, inf_du = 0.0, step_norm = 0.0;
  for (int i = 0; i < n; ++i) inf_pr = std::max(inf_pr, std::fabs(hT[i])); /* :1041 */
  std::vector<double> Gx((size_t)std::max(d, 1) * n), Gu((size_t)std::max(d, 1) * m);
  std::vector<double> YSall((size_t)N * std::max(d, 1)), rhat_all((size_t)N * std::max(d, 1)), prim_all((size_t)N * std::max(d, 1)),
      ssafe_all((size_t)N * std::max(d, 1));
  for (int t = 0; t < N; ++t) {
    const double *x = &s.X[(size_t)t * n], *u = &s.U[(size_t)t * m];
    const double *ref = ref_at(p, s.xref, s.ref_traj, t);
    double *Qt = &Q[(size_t)t * n * n], *qt = &q[(size_t)t * n], *Rt = &R[(size_t)t * m * m], *rt = &r[(size_t)t * m], *Mt = &M[(size_t)t * n * m];
    for (int i = 0; i < n; ++i) {
      double lx = 0.0;
      for (int j = 0; j < n; ++j) lx += (2.0 * (p->Q[i * n + j] * dt)) * (x[j] - ref[j]);
      qt[i] = lx;
      for (int j = 0; j < n; ++j) Qt[i * n + j] = 0.5 * (2.0 * (p->Q[i * n + j] * dt) + 2.0 * (p->Q[j * n + i] * dt)); /* :1149 */
    }
    for (int i = 0; i < m; ++i) {
      double lu = 0.0;
      for (int j = 0; j < m; ++j) lu += (2.0 * (p->R[i * m + j] * dt)) * u[j];
      rt[i] = lu;
      for (int j = 0; j < m; ++j) Rt[i * m + j] = 0.5 * (2.0 * (p->R[i * m + j] * dt) + 2.0 * (p->R[j * m + i] * dt));
    }
    if (d) { /* :1181-1252 */
      const double *y = &s.Y[(size_t)t * d], *sl = &s.S[(size_t)t * d], *g = &s.G[(size_t)t * d];
      constraint_jacobians(p, s.cs, s.nc, x, Gx.data(), Gu.data());
      double w[MAXDUAL];
      double *YS = &YSall[(size_t)t * d], *rh = &rhat_all[(size_t)t * d], *pr = &prim_all[(size_t)t * d], *sf = &ssafe_all[(size_t)t * d];
      for (int i = 0; i < d; ++i) {
        sf[i] = std::max(sl[i], std::max(s.mu * 1e-3, EPS_SLACK));
        YS[i] = clip_pos(y[i], sf[i]);
        pr[i] = g[i] + sl[i];
        const double comp = y[i] * sl[i] - s.mu;
        rh[i] = y[i] * pr[i] - comp;
        w[i] = y[i] + clip_signed(rh[i], sf[i]);
        inf_pr = std::max(inf_pr, std::fabs(pr[i]));
        inf_comp = std::max(inf_comp, std::fabs(comp));
      }
      for (int i = 0; i < n; ++i) {
        double a = 0.0;
        for (int k2 = 0; k2 < d; ++k2) a += Gx[k2 * n + i] * w[k2];
        qt[i] += a;
      }
      for (int i = 0; i < m; ++i) {
        double a = 0.0;
        for (int k2 = 0; k2 < d; ++k2) a += Gu[k2 * m + i] * w[k2];
        rt[i] += a;
      }
      double Qn[MAXN * MAXN], Rn[MAXM * MAXM];
      for (int i = 0; i < n; ++i)
        for (int j = 0; j < n; ++j) {
          double a = 0.0;
          for (int k2 = 0; k2 < d; ++k2) a += Gx[k2 * n + i] * (YS[k2] * Gx[k2 * n + j]);
          Qn[i * n + j] = Qt[i * n + j] + a;
        }
      for (int i = 0; i < n; ++i)
        for (int j = 0; j < m; ++j) { /* M += (Q_yu^T YSinv Q_yx)^T */
          double a = 0.0;
          for (int k2 = 0; k2 < d; ++k2) a += Gu[k2 * m + j] * (YS[k2] * Gx[k2 * n + i]);
          Mt[i * m + j] += a;
        }
      for (int i = 0; i < m; ++i)
        for (int j = 0; j < m; ++j) {
          double a = 0.0;
          for (int k2 = 0; k2 < d; ++k2) a += Gu[k2 * m + i] * (YS[k2] * Gu[k2 * m + j]);
          Rn[i * m + j] = Rt[i * m + j] + a;
        }
      for (int i = 0; i < n; ++i)
        for (int j = 0; j < n; ++j) Qt[i * n + j] = 0.5 * (Qn[i * n + j] + Qn[j * n + i]);
      for (int i = 0; i < m; ++i)
        for (int j = 0; j < m; ++j) Rt[i * m + j] = 0.5 * (Rn[i * m + j] + Rn[j * m + i]);
    }
    for (int i = 0; i < m; ++i) Rt[i * m + i] += s.reg; /* :1254 */
  }
  /* solveTerminalEqualityLQR */
  std::vector<double> qb = q;
  for (int i = 0; i < n; ++i) qb[(size_t)N * n + i] += s.lamT[i]; /* H_T^T lambda_prev, H_T = I (:520-526) */
  std::vector<std::vector<double>> Kv(pd + 1), kv(pd + 1), Pv(pd + 1), pv(pd + 1);
  std::vector<double> xT((size_t)(pd + 1) * n, 0.0);
  for (int v = 0; v <= pd; ++v) {
    std::vector<double> qv = qb;
    if (v > 0) qv[(size_t)N * n + (v - 1)] += 1.0; /* H_T.row(v-1)^T */
    if (!sequential_lqr(n, m, N, Q, qv, R, r, M, s.A, s.B, Kv[v], kv[v], Pv[v], pv[v])) return false;
    double dx[MAXN], dxn[MAXN], du[MAXM];
    for (int i = 0; i < n; ++i) dx[i] = 0.0;
    for (int t = 0; t < N; ++t) { /* rolloutLinearPolicy (:368-392) */
      for (int i = 0; i < m; ++i) {
        double a = 0.0;
        for (int j = 0; j < n; ++j) a += Kv[v][((size_t)t * m + i) * n + j] * dx[j];
        du[i] = kv[v][(size_t)t * m + i] + a;
      }
      for (int i = 0; i < n; ++i) {
        double a1 = 0.0, a2 = 0.0;
        for (int j = 0; j < n; ++j) a1 += s.A[(size_t)t * n * n + i * n + j] * dx[j];
        for (int j = 0; j < m; ++j) a2 += s.B[(size_t)t * n * m + i * m + j] * du[j];
        dxn[i] = (a1 + a2) + 0.0;
      }
      for (int i = 0; i < n; ++i) dx[i] = dxn[i];
    }
    for (int i = 0; i < n; ++i) xT[(size_t)v * n + i] = dx[i];
  }
  double As[MAXN * MAXN], rhs[MAXN], AtA[MAXN * MAXN], Atb[MAXN];
  for (int i = 0; i < n; ++i) {
    for (int j = 0; j < pd; ++j) As[i * pd + j] = xT[(size_t)(j + 1) * n + i] - xT[i]; /* A_small = H_T S_mat = S_mat */
    rhs[i] = -hT[i] - xT[i];                                                              /* b_T - H_T xT_0 */
  }
  double trace = 0.0;
  for (int i = 0; i < pd; ++i) {
    for (int j = 0; j < pd; ++j) {
      double a = 0.0;
      for (int l = 0; l < n; ++l) a += As[l * pd + i] * As[l * pd + j];
      AtA[i * pd + j] = a;
    }
    double a = 0.0;
    for (int l = 0; l < n; ++l) a += As[l * pd + i] * rhs[l];
    Atb[i] = a;
    trace += AtA[i * pd + i];
  }
  const double trace_term = trace > 1.0 ? trace / std::max(pd, 1) : 1.0;
  const double base_floor = std::max(1e-10, s.io->jacobian_regularization_value * std::pow(std::max(s.mu, 0.0), s.io->jacobian_regularization_exponent));
  const double regq = std::max(base_floor, 1e-6 * trace_term);
  double ev[MAXN];
  eigenvalues_sym(AtA, pd, ev);
  double smax = 0.0, smin = std::numeric_limits<double>::infinity();
  for (int i = 0; i < pd; ++i) {
    const double sg = std::sqrt(std::max(ev[i], 0.0));
    smax = std::max(smax, sg);
    smin = std::min(smin, sg);
  }
  const double svd_reg = std::max(1e-8 * smax - smin, 0.0);
  const double reg_base = std::max(regq, svd_reg);
  double rn = 0.0;
  for (int i = 0; i < n; ++i) rn += rhs[i] * rhs[i];
  const double cap = 100.0 * (1.0 + std::sqrt(rn));
  const double scales[5] = {1.0, 10.0, 100.0, 1e3, 1e4};
  double best[MAXN], best_res = std::numeric_limits<double>::infinity();
  bool found = false;
  for (int i = 0; i < pd; ++i) best[i] = 0.0;
  LDLT ldlt;
  for (int si = 0; si < 5; ++si) {
    const double reg_i = std::max(reg_base * scales[si], 1e-12);
    double sh[MAXN * MAXN], lam[MAXN];
    for (int i = 0; i < pd * pd; ++i) sh[i] = AtA[i];
    for (int i = 0; i < pd; ++i) sh[i * pd + i] += reg_i;
    ldlt.compute(sh, pd);
    if (!ldlt.ok) continue;
    for (int i = 0; i < pd; ++i) lam[i] = Atb[i];
    ldlt.solve_inplace(lam);
    bool fin = true;
    double ln = 0.0;
    for (int i = 0; i < pd; ++i) {
      fin = fin && std::isfinite(lam[i]);
      ln += lam[i] * lam[i];
    }
    if (!fin) continue;
    ln = std::sqrt(ln);
    if (ln > cap)
      for (int i = 0; i < pd; ++i) lam[i] *= cap / std::max(ln, 1e-12);
    double res = 0.0;
    for (int i = 0; i < n; ++i) {
      double a = 0.0;
      for (int j = 0; j < pd; ++j) a += As[i * pd + j] * lam[j];
      res += (a - rhs[i]) * (a - rhs[i]);
    }
    res = std::sqrt(res);
    if (!std::isfinite(res)) continue;
    if (!found || res < best_res) {
      for (int i = 0; i < pd; ++i) best[i] = lam[i];
      best_res = res;
      found = true;
    }
  }
  s.Ku = Kv[0];
  s.ku = kv[0];
  std::vector<double> pout = pv[0];
  for (int v = 0; v < pd; ++v) {
    const double coeff = best[v];
    for (size_t i = 0; i < s.ku.size(); ++i) s.ku[i] += coeff * (kv[v + 1][i] - kv[0][i]);
    for (size_t i = 0; i < pout.size(); ++i) pout[i] += coeff * (pv[v + 1][i] - pv[0][i]);
  }
  for (int i = 0; i < pd; ++i) s.dlamT[i] = best[i]; /* dLambda_T_eq_ = lambda_delta (:1267) */
  for (int t = 0; t < N; ++t) { /* :1268-1274 */
    for (int i = 0; i < m; ++i) {
      double a = 0.0;
      for (int l = 0; l < n; ++l) a += s.B[(size_t)t * n * m + l * m + i] * pout[(size_t)(t + 1) * n + l];
      inf_du = std::max(inf_du, std::fabs(r[(size_t)t * m + i] + a));
      step_norm = std::max(step_norm, std::fabs(s.ku[(size_t)t * m + i]));
    }
  }
  /* rolloutLinearPolicy with the final gains + slack / dual gains (:1276-1320) */
  {
    double dx[MAXN], dxn[MAXN], du[MAXM];
    for (int i = 0; i < n; ++i) dx[i] = 0.0;
    for (int t = 0; t < N; ++t) {
      const double *x = &s.X[(size_t)t * n];
      const double *Ku = &s.Ku[(size_t)t * m * n], *ku = &s.ku[(size_t)t * m];
      if (d) {
        constraint_jacobians(p, s.cs, s.nc, x, Gx.data(), Gu.data());
        const double *y = &s.Y[(size_t)t * d];
        const double *YS = &YSall[(size_t)t * d], *rh = &rhat_all[(size_t)t * d], *pr = &prim_all[(size_t)t * d], *sf = &ssafe_all[(size_t)t * d];
        for (int k2 = 0; k2 < d; ++k2) {
          double temp = 0.0;
          for (int i = 0; i < m; ++i) temp += Gu[k2 * m + i] * ku[i];
          const size_t e = (size_t)t * d + k2;
          s.ky[e] = clip_signed(rh[k2] + y[k2] * temp, sf[k2]);
          s.ks[e] = -pr[k2] - temp;
          double a1 = 0.0, a2 = 0.0;
          for (int j = 0; j < n; ++j) {
            double gk = 0.0;
            for (int i = 0; i < m; ++i) gk += Gu[k2 * m + i] * Ku[i * n + j];
            const double qq = Gx[k2 * n + j] + gk;
            s.Ky[e * n + j] = clampd(YS[k2] * qq, -MAX_BARRIER_RATIO, MAX_BARRIER_RATIO);
            s.Ks[e * n + j] = -Gx[k2 * n + j] - gk;
            a1 += s.Ks[e * n + j] * dx[j];
            a2 += s.Ky[e * n + j] * dx[j];
          }
          s.dS[e] = s.ks[e] + a1;
          s.dY[e] = clampd(s.ky[e] + a2, -MAX_BARRIER_RATIO, MAX_BARRIER_RATIO);
        }
      }
      for (int i = 0; i < m; ++i) {
        double a = 0.0;
        for (int j = 0; j < n; ++j) a += Ku[i * n + j] * dx[j];
        du[i] = ku[i] + a;
      }
      for (int i = 0; i < n; ++i) {
        double a1 = 0.0, a2 = 0.0;
        for (int j = 0; j < n; ++j) a1 += s.A[(size_t)t * n * n + i * n + j] * dx[j];
        for (int j = 0; j < m; ++j) a2 += s.B[(size_t)t * n * m + i * m + j] * du[j];
        dxn[i] = (a1 + a2) + 0.0;
      }
      for (int i = 0; i < n; ++i) dx[i] = dxn[i];
    }
  }
  s.inf_pr = inf_pr;
  s.inf_du = inf_du;
  s.inf_comp = inf_comp;
  s.step_norm = step_norm;
  return true;
}

/* IPDDPSolver::backwardPass (:960-1569), branches 1 and 3 */
bool ip_backward(IpState &s) {
  const int n = s.n, m = s.m, N = s.N, d = s.d;
  const oracle_problem *p = s.p;
  const double dt = p->dt;
  /* precomputeDynamicsDerivatives: F_x = dt*Fx with +1 on the diagonal, F_u = dt*Fu (cddp_solver_base.cpp:340-344) */
  {
    double Fx[MAXN * MAXN], Fu[MAXN * MAXM];
    for (int t = 0; t < N; ++t) {
      jacobians(p, &s.X[(size_t)t * n], &s.U[(size_t)t * m], t * dt, Fx, Fu);
      double *A = &s.A[(size_t)t * n * n], *B = &s.B[(size_t)t * n * m];
      for (int i = 0; i < n * n; ++i) A[i] = dt * Fx[i];
      for (int i = 0; i < n; ++i) A[i * n + i] += 1.0;
      for (int i = 0; i < n * m; ++i) B[i] = dt * Fu[i];
    }
  }
  double Vx[MAXN], Vxx[MAXN * MAXN];
  {
    const double *xN = &s.X[(size_t)N * n];
    double e[MAXN];
    for (int i = 0; i < n; ++i) e[i] = xN[i] - s.xref[i];
    for (int i = 0; i < n; ++i) {
      double acc = 0.0;
      for (int j = 0; j < n; ++j) acc += (2.0 * p->Qf[i * n + j]) * e[j];
      Vx[i] = acc;
    }
    for (int i = 0; i < n; ++i)
      for (int j = 0; j < n; ++j) Vxx[i * n + j] = 0.5 * (2.0 * p->Qf[i * n + j] + 2.0 * p->Qf[j * n + i]); /* :990 */
  }
  s.dV[0] = s.dV[1] = 0.0;
  if (s.teq) return ip_backward_teq(s, Vx, Vxx); /* :1120-1353; dV_ stays zero on this branch */
  double inf_du = 0.0, inf_pr = 0.0, inf_comp = 0.0, step_norm = 0.0;
  std::vector<double> Gx((size_t)std::max(d, 1) * n), Gu((size_t)std::max(d, 1) * m);
  double Qx[MAXN], Qu[MAXM], Qxx[MAXN * MAXN], Qux[MAXM * MAXN], Quu[MAXM * MAXM], Qr[MAXM * MAXM];
  double VA[MAXN * MAXN], VB[MAXN * MAXM];
  double YS[MAXDUAL], rhat[MAXDUAL], Sir[MAXDUAL], prim[MAXDUAL], comp[MAXDUAL], ssafe[MAXDUAL];
  LDLT ldlt;
  for (int t = N - 1; t >= 0; --t) {
    const double *x = &s.X[(size_t)t * n], *u = &s.U[(size_t)t * m];
    const double *A = &s.A[(size_t)t * n * n], *B = &s.B[(size_t)t * n * m];
    const double *y = d ? &s.Y[(size_t)t * d] : nullptr, *sl = d ? &s.S[(size_t)t * d] : nullptr;
    const double *g = d ? &s.G[(size_t)t * d] : nullptr;
    if (d) constraint_jacobians(p, s.cs, s.nc, x, Gx.data(), Gu.data()); /* precomputeConstraintGradients :2145-2250 */
    const double *ref = ref_at(p, s.xref, s.ref_traj, t);
    double e[MAXN];
    for (int i = 0; i < n; ++i) e[i] = x[i] - ref[i];
    for (int i = 0; i < n; ++i) { /* Q_x = l_x + Q_yx^T y + A^T V_x (:1393) */
      double lx = 0.0;
      for (int j = 0; j < n; ++j) lx += (2.0 * (p->Q[i * n + j] * dt)) * e[j];
      double gy = 0.0;
      for (int r = 0; r < d; ++r) gy += Gx[r * n + i] * y[r];
      double av = 0.0;
      for (int j = 0; j < n; ++j) av += A[j * n + i] * Vx[j];
      Qx[i] = d ? (lx + gy) + av : lx + av;
    }
    for (int i = 0; i < m; ++i) { /* Q_u = l_u + Q_yu^T y + B^T V_x (:1394) */
      double lu = 0.0;
      for (int j = 0; j < m; ++j) lu += (2.0 * (p->R[i * m + j] * dt)) * u[j];
      double gy = 0.0;
      for (int r = 0; r < d; ++r) gy += Gu[r * m + i] * y[r];
      double bv = 0.0;
      for (int j = 0; j < n; ++j) bv += B[j * m + i] * Vx[j];
      Qu[i] = d ? (lu + gy) + bv : lu + bv;
    }
    for (int i = 0; i < n; ++i)
      for (int j = 0; j < n; ++j) {
        double acc = 0.0;
        for (int l = 0; l < n; ++l) acc += Vxx[i * n + l] * A[l * n + j];
        VA[i * n + j] = acc;
      }
    for (int i = 0; i < n; ++i)
      for (int j = 0; j < m; ++j) {
        double acc = 0.0;
        for (int l = 0; l < n; ++l) acc += Vxx[i * n + l] * B[l * m + j];
        VB[i * m + j] = acc;
      }
    for (int i = 0; i < n; ++i) /* Q_xx = l_xx + A^T V_xx A (:1395) */
      for (int j = 0; j < n; ++j) {
        double acc = 0.0;
        for (int l = 0; l < n; ++l) acc += A[l * n + i] * VA[l * n + j];
        Qxx[i * n + j] = 2.0 * (p->Q[i * n + j] * dt) + acc;
      }
    for (int i = 0; i < m; ++i) /* Q_ux = l_ux + B^T V_xx A (:1396), l_ux = 0 */
      for (int j = 0; j < n; ++j) {
        double acc = 0.0;
        for (int l = 0; l < n; ++l) acc += B[l * m + i] * VA[l * n + j];
        Qux[i * n + j] = acc;
      }
    for (int i = 0; i < m; ++i) /* Q_uu = l_uu + B^T V_xx B (:1397) */
      for (int j = 0; j < m; ++j) {
        double acc = 0.0;
        for (int l = 0; l < n; ++l) acc += B[l * m + i] * VB[l * m + j];
        Quu[i * m + j] = 2.0 * (p->R[i * m + j] * dt) + acc;
      }
    for (int i = 0; i < d; ++i) { /* :1413-1425 */
      ssafe[i] = std::max(sl[i], std::max(s.mu * 1e-3, EPS_SLACK));
      YS[i] = clip_pos(y[i], ssafe[i]);
      prim[i] = g[i] + sl[i];
      comp[i] = y[i] * sl[i] - s.mu;
      rhat[i] = y[i] * prim[i] - comp[i];
      Sir[i] = clip_signed(rhat[i], ssafe[i]); /* :1437-1443 */
    }
    for (int i = 0; i < m; ++i) /* Q_uu_reg = sym(Q_uu) + Q_yu^T YSinv Q_yu + reg I (:1427-1429; branch 1: :1083-1084) */
      for (int j = 0; j < m; ++j) {
        double acc = 0.0;
        for (int r = 0; r < d; ++r) acc += Gu[r * m + i] * (YS[r] * Gu[r * m + j]);
        Qr[i * m + j] = 0.5 * (Quu[i * m + j] + Quu[j * m + i]) + acc;
      }
    for (int i = 0; i < m; ++i) Qr[i * m + i] += s.reg;
    ldlt.compute(Qr, m);
    if (!ldlt.ok) return false; /* :1431-1435 */
    double rhs0[MAXM], rhsK[MAXM * MAXN];
    for (int i = 0; i < m; ++i) { /* bigRHS (:1444-1447) */
      double acc = 0.0;
      for (int r = 0; r < d; ++r) acc += Gu[r * m + i] * Sir[r];
      rhs0[i] = d ? Qu[i] + acc : Qu[i];
      for (int j = 0; j < n; ++j) {
        double a2 = 0.0;
        for (int r = 0; r < d; ++r) a2 += Gu[r * m + i] * (YS[r] * Gx[r * n + j]);
        rhsK[i * n + j] = d ? Qux[i * n + j] + a2 : Qux[i * n + j];
      }
    }
    double *ku = &s.ku[(size_t)t * m], *Ku = &s.Ku[(size_t)t * m * n];
    {
      double col[MAXM];
      for (int i = 0; i < m; ++i) col[i] = rhs0[i];
      ldlt.solve_inplace(col);
      for (int i = 0; i < m; ++i) ku[i] = -col[i];
      for (int j = 0; j < n; ++j) {
        for (int i = 0; i < m; ++i) col[i] = rhsK[i * n + j];
        ldlt.solve_inplace(col);
        for (int i = 0; i < m; ++i) Ku[i * n + j] = -col[i];
      }
    }
    if (d) { /* :1463-1491 */
      double *ky = &s.ky[(size_t)t * d], *Ky = &s.Ky[(size_t)t * d * n];
      double *ks = &s.ks[(size_t)t * d], *Ks = &s.Ks[(size_t)t * d * n];
      for (int r = 0; r < d; ++r) {
        double temp = 0.0;
        for (int i = 0; i < m; ++i) temp += Gu[r * m + i] * ku[i];
        ky[r] = clip_signed(rhat[r] + y[r] * temp, ssafe[r]);
        ks[r] = -prim[r] - temp;
        for (int j = 0; j < n; ++j) {
          double gk = 0.0;
          for (int i = 0; i < m; ++i) gk += Gu[r * m + i] * Ku[i * n + j];
          const double q = Gx[r * n + j] + gk;
          Ky[r * n + j] = clampd(YS[r] * q, -MAX_BARRIER_RATIO, MAX_BARRIER_RATIO);
          Ks[r * n + j] = -Gx[r * n + j] - gk;
        }
      }
      /* condensed Q-terms (:1493-1497) */
      for (int i = 0; i < m; ++i) Qu[i] = rhs0[i];
      for (int i = 0; i < n; ++i) {
        double acc = 0.0;
        for (int r = 0; r < d; ++r) acc += Gx[r * n + i] * Sir[r];
        Qx[i] += acc;
      }
      for (int i = 0; i < n; ++i)
        for (int j = 0; j < n; ++j) {
          double acc = 0.0;
          for (int r = 0; r < d; ++r) acc += Gx[r * n + i] * (YS[r] * Gx[r * n + j]);
          Qxx[i * n + j] += acc;
        }
      for (int i = 0; i < m; ++i) {
        for (int j = 0; j < n; ++j) Qux[i * n + j] = rhsK[i * n + j];
        for (int j = 0; j < m; ++j) {
          double acc = 0.0;
          for (int r = 0; r < d; ++r) acc += Gu[r * m + i] * (YS[r] * Gu[r * m + j]);
          Quu[i * m + j] += acc;
        }
      }
    } else { /* branch 1 (:1083-1084): the regularised, symmetrised Q_uu is what enters V and dV */
      for (int i = 0; i < m * m; ++i) Quu[i] = Qr[i];
    }
    double Quuk[MAXM];
    for (int i = 0; i < m; ++i) {
      double acc = 0.0;
      for (int j = 0; j < m; ++j) acc += Quu[i * m + j] * ku[j];
      Quuk[i] = acc;
    }
    double d0 = 0.0, d1 = 0.0;
    for (int i = 0; i < m; ++i) { /* :1499-1500 */
      d0 += ku[i] * Qu[i];
      d1 += ku[i] * Quuk[i];
    }
    s.dV[0] += d0;
    s.dV[1] += 0.5 * d1;
    double Vxn[MAXN], Vxxn[MAXN * MAXN], QuuK[MAXM * MAXN];
    for (int i = 0; i < m; ++i)
      for (int j = 0; j < n; ++j) {
        double acc = 0.0;
        for (int l = 0; l < m; ++l) acc += Quu[i * m + l] * Ku[l * n + j];
        QuuK[i * n + j] = acc;
      }
    for (int i = 0; i < n; ++i) { /* V_x = Q_x + K^T Q_u + Q_ux^T k + K^T Q_uu k (:1502-1503) */
      double a1 = 0.0, a2 = 0.0, a3 = 0.0;
      for (int l = 0; l < m; ++l) {
        a1 += Ku[l * n + i] * Qu[l];
        a2 += Qux[l * n + i] * ku[l];
        a3 += Ku[l * n + i] * Quuk[l];
      }
      Vxn[i] = ((Qx[i] + a1) + a2) + a3;
    }
    for (int i = 0; i < n; ++i) /* V_xx = Q_xx + K^T Q_ux + Q_ux^T K + K^T Q_uu K (:1504-1505) */
      for (int j = 0; j < n; ++j) {
        double a1 = 0.0, a2 = 0.0, a3 = 0.0;
        for (int l = 0; l < m; ++l) {
          a1 += Ku[l * n + i] * Qux[l * n + j];
          a2 += Qux[l * n + i] * Ku[l * n + j];
          a3 += Ku[l * n + i] * QuuK[l * n + j];
        }
        Vxxn[i * n + j] = ((Qxx[i * n + j] + a1) + a2) + a3;
      }
    for (int i = 0; i < n; ++i) {
      Vx[i] = Vxn[i];
      for (int j = 0; j < n; ++j) Vxx[i * n + j] = 0.5 * (Vxxn[i * n + j] + Vxxn[j * n + i]);
    }
    for (int i = 0; i < m; ++i) { /* :1510-1513 */
      inf_du = std::max(inf_du, std::fabs(Qu[i]));
      step_norm = std::max(step_norm, std::fabs(ku[i]));
    }
    for (int i = 0; i < d; ++i) {
      inf_pr = std::max(inf_pr, std::fabs(prim[i]));
      inf_comp = std::max(inf_comp, std::fabs(comp[i]));
    }
  }
  if (d) { /* rolloutLinearPolicy from dx0 = 0 (:368-392) and the linearised slack/dual steps (:1516-1538) */
    double dx[MAXN], dxn[MAXN], du[MAXM];
    for (int i = 0; i < n; ++i) dx[i] = 0.0;
    for (int t = 0; t < N; ++t) {
      const double *A = &s.A[(size_t)t * n * n], *B = &s.B[(size_t)t * n * m];
      for (int r = 0; r < d; ++r) {
        double a1 = 0.0, a2 = 0.0;
        for (int j = 0; j < n; ++j) {
          a1 += s.Ks[((size_t)t * d + r) * n + j] * dx[j];
          a2 += s.Ky[((size_t)t * d + r) * n + j] * dx[j];
        }
        s.dS[(size_t)t * d + r] = s.ks[(size_t)t * d + r] + a1;
        s.dY[(size_t)t * d + r] = clampd(s.ky[(size_t)t * d + r] + a2, -MAX_BARRIER_RATIO, MAX_BARRIER_RATIO);
      }
      for (int i = 0; i < m; ++i) {
        double acc = 0.0;
        for (int j = 0; j < n; ++j) acc += s.Ku[((size_t)t * m + i) * n + j] * dx[j];
        du[i] = s.ku[(size_t)t * m + i] + acc;
      }
      for (int i = 0; i < n; ++i) {
        double a1 = 0.0, a2 = 0.0;
        for (int j = 0; j < n; ++j) a1 += A[i * n + j] * dx[j];
        for (int j = 0; j < m; ++j) a2 += B[i * m + j] * du[j];
        dxn[i] = (a1 + a2) + 0.0;
      }
      for (int i = 0; i < n; ++i) dx[i] = dxn[i];
    }
    s.inf_pr = inf_pr;
    s.inf_comp = inf_comp;
  } else {
    s.inf_pr = 0.0; /* :1113-1116 */
    s.inf_comp = 0.0;
  }
  s.inf_du = inf_du;
  s.step_norm = step_norm;
  return true;
}

struct IpTrial {
  bool success = false;
  double cost = 0, merit = 0, theta = 0, inf_pr = 0, inf_comp = 0, alpha_pr = 0, alpha_du = 0;
  /* Test instrumentation (not in the reference): relative margin by which this trial's accept/reject decision was
   * taken.  With alpha_pr at its fraction-to-boundary cap the slack test at t = 0 compares two numbers that are equal
   * in exact arithmetic, so the reference's own decision is then made by roundoff; tests use the smallest margin of a
   * solve to tell roundoff-decided instances from robust ones. */
  double margin = 0;
  std::vector<double> X, U, Y, S, G, lamT;
};

/* IPDDPSolver::forwardPass (:1571-1876) */
/* force (decision-following instrumentation): a fraction-to-boundary violation is noted (own verdict = reject, margin =
 * that violation's) but the trial is still evaluated to the end, so that a caller following a RECORDED acceptance gets the
 * complete candidate; without it the function returns at the first violation as the reference does (:1623-1630). */
void ip_forward(const IpState &s, double alpha, IpTrial &r, bool force = false) {
  const int n = s.n, m = s.m, N = s.N, d = s.d;
  const oracle_problem *p = s.p;
  r.success = false;
  r.margin = 1.0;
  double feas_margin = 1.0;
  bool own_reject = false;
  double reject_margin = 1.0;
  /* computeMaxStepSizes (:2939-2988) */
  const double tau_b = std::max(s.io->min_fraction_to_boundary, 1.0 - s.mu);
  double apm = 1.0, adm = 1.0;
  for (int c = 0, o = 0; c < s.nc; ++c) {
    const int dim = constraint_dim(p, s.cs[c]);
    for (int t = 0; t < N; ++t)
      for (int i = 0; i < dim; ++i) {
        const size_t q = (size_t)t * d + o + i;
        if (s.dS[q] < 0.0) apm = std::min(apm, -tau_b * s.S[q] / s.dS[q]);
        if (s.dY[q] < 0.0) adm = std::min(adm, -tau_b * s.Y[q] / s.dY[q]);
      }
    o += dim;
  }
  apm = clampd(apm, 0.0, 1.0);
  adm = clampd(adm, 0.0, 1.0);
  const double tau = s.nc == 0 ? 1.0 : std::max(s.io->min_fraction_to_boundary, 1.0 - s.mu); /* :1585-1588 */
  const double alpha_pr = std::min(alpha, apm), alpha_du = std::min(alpha, adm);
  r.alpha_pr = alpha_pr;
  r.alpha_du = alpha_du;
  r.X.assign((size_t)(N + 1) * n, 0.0);
  r.U.assign((size_t)N * m, 0.0);
  r.Y = s.Y;
  r.S = s.S;
  r.G.assign((size_t)N * d, 0.0);
  std::memcpy(r.X.data(), s.x0, sizeof(double) * n);
  auto finite_vec = [](const double *v, int k) {
    for (int i = 0; i < k; ++i)
      if (!std::isfinite(v[i])) return false;
    return true;
  };
  for (int t = 0; t < N; ++t) {
    double dx[MAXN];
    const double *xt = &r.X[(size_t)t * n];
    for (int i = 0; i < n; ++i) dx[i] = xt[i] - s.X[(size_t)t * n + i];
    for (int c = 0, o = 0; c < s.nc; ++c) { /* :1620-1647, per constraint */
      const int dim = constraint_dim(p, s.cs[c]);
      double sn[MAXDUAL], yn[MAXDUAL], sscale[MAXDUAL], yscale[MAXDUAL];
      for (int i = 0; i < dim; ++i) {
        const size_t q = (size_t)t * d + o + i;
        double a1 = 0.0, a2 = 0.0;
        for (int j = 0; j < n; ++j) {
          a1 += s.Ks[q * n + j] * dx[j];
          a2 += s.Ky[q * n + j] * dx[j];
        }
        sn[i] = (s.S[q] + alpha_pr * s.ks[q]) + a1;
        yn[i] = (s.Y[q] + alpha_du * s.ky[q]) + a2;
        /* instrumentation: the magnitude roundoff is relative to — dx = x' - x carries the absolute roundoff of the rolled-out
         * state (~eps |x|), which the gains K_s, K_y amplify; a threshold (1 - tau) s at the scale of a collapsed slack
         * (1e-14) can be far below it */
        double b1 = std::fabs(s.S[q]) + std::fabs(alpha_pr * s.ks[q]), b2 = std::fabs(s.Y[q]) + std::fabs(alpha_du * s.ky[q]);
        for (int j = 0; j < n; ++j) {
          const double xm = std::fabs(xt[j]) + std::fabs(s.X[(size_t)t * n + j]);
          b1 += std::fabs(s.Ks[q * n + j]) * xm;
          b2 += std::fabs(s.Ky[q * n + j]) * xm;
        }
        sscale[i] = b1;
        yscale[i] = b2;
      }
      for (int i = 0; i < dim; ++i) {
        const size_t q = (size_t)t * d + o + i;
        const double smin = (1.0 - tau) * s.S[q], ymin = (1.0 - tau) * s.Y[q];
        const double ms = std::fabs(sn[i] - smin) / std::max(sscale[i], 1e-300);
        const double my = std::fabs(yn[i] - ymin) / std::max(yscale[i], 1e-300);
        if (sn[i] < smin || yn[i] < ymin) {
          /* rejected: the decision is as robust as this violation (the scan below looks for a clearer one) */
          double best = std::max(sn[i] < smin ? ms : 0.0, yn[i] < ymin ? my : 0.0);
          r.margin = best;
          if (!force) return;
          if (!own_reject) reject_margin = best;
          own_reject = true;
          continue;
        }
        feas_margin = std::min(feas_margin, std::min(ms, my));
      }
      if (!finite_vec(sn, dim) || !finite_vec(yn, dim)) return;
      for (int i = 0; i < dim; ++i) {
        r.S[(size_t)t * d + o + i] = sn[i];
        r.Y[(size_t)t * d + o + i] = yn[i];
      }
      o += dim;
    }
    double *ut = &r.U[(size_t)t * m];
    for (int i = 0; i < m; ++i) { /* :1650-1651 (no clamp) */
      double acc = 0.0;
      for (int j = 0; j < n; ++j) acc += s.Ku[((size_t)t * m + i) * n + j] * dx[j];
      ut[i] = (s.U[(size_t)t * m + i] + alpha_pr * s.ku[(size_t)t * m + i]) + acc;
    }
    discrete_dynamics(p, xt, ut, t * p->dt, &r.X[(size_t)(t + 1) * n]);
    if (!finite_vec(&r.X[(size_t)(t + 1) * n], n) || !finite_vec(ut, m)) return;
  }
  double cost_new = 0.0; /* :1735-1751 */
  for (int t = 0; t < N; ++t) {
    cost_new += running_cost(p, &r.X[(size_t)t * n], &r.U[(size_t)t * m], ref_at(p, s.xref, s.ref_traj, t));
    if (d) eval_constraints(p, s.cs, s.nc, &r.X[(size_t)t * n], &r.U[(size_t)t * m], &r.G[(size_t)t * d]);
  }
  cost_new += terminal_cost(p, &r.X[(size_t)N * n], s.xref);
  double hTn[MAXN];
  r.lamT = s.lamT;
  if (s.teq) { /* :1716-1723, :1756-1760 */
    for (int i = 0; i < n; ++i) {
      r.lamT[i] = s.lamT[i] + alpha_pr * s.dlamT[i];
      if (!std::isfinite(r.lamT[i])) return;
    }
    ip_terminal_residual(s, r.X.data(), hTn);
  }
  const double *hn = s.teq ? hTn : nullptr;
  const double phi_new = ip_merit(s, r.S.data(), cost_new, s.teq ? r.lamT.data() : nullptr, hn);
  const double theta_new = ip_theta(s, r.G.data(), r.S.data(), hn);
  double ipn = 0.0, icn = 0.0;
  ip_primal_comp(s, r.G.data(), r.S.data(), r.Y.data(), s.mu, &ipn, &icn, hn);
  if (!std::isfinite(phi_new) || !std::isfinite(theta_new) || !std::isfinite(ipn) || !std::isfinite(icn)) return;
  bool accept = false;
  double acc_margin = 1.0;
  auto rm = [](double a, double b) { return std::fabs(a - b) / std::max(std::max(std::fabs(a), std::fabs(b)), 1e-300); };
  if (s.nc == 0 && !s.teq) { /* :1785-1794: hard-coded 1e-6 */
    const double dJ = s.cost - cost_new;
    const double expected = -alpha_pr * (s.dV[0] + 0.5 * alpha_pr * s.dV[1]);
    const double ratio = expected > 0.0 ? dJ / expected : std::copysign(1.0, dJ);
    accept = ratio > 1e-6;
    acc_margin = std::fabs(dJ) / std::max(std::fabs(s.cost), 1e-300);
  } else { /* :1796-1839 */
    const double expected_improvement = alpha_pr * s.dV[0];
    const double cv_old = s.filter.empty() ? 0.0 : s.filter.back().theta;
    const double high_ref = s.filter.empty() ? s.filter_theta : cv_old;
    const double merit_old = s.merit;
    if (theta_new > s.io->max_violation_threshold) {
      const double rhs = (1 - s.io->violation_acceptance_threshold) * high_ref;
      if (theta_new < rhs) accept = true;
      acc_margin = rm(theta_new, rhs);
    } else if (std::max(theta_new, cv_old) < s.io->min_violation_for_armijo_check && expected_improvement < 0) {
      const double rhs = merit_old + s.o->armijo_constant * expected_improvement;
      if (phi_new < rhs) accept = true;
      acc_margin = rm(phi_new, rhs);
    } else {
      const double r1 = merit_old - s.io->merit_acceptance_threshold * theta_new;
      const double r2 = (1 - s.io->violation_acceptance_threshold) * cv_old;
      const bool c1 = phi_new < r1, c2 = theta_new < r2;
      if (c1 || c2) accept = true;
      const double m1 = rm(phi_new, r1), m2 = rm(theta_new, r2);
      acc_margin = accept ? std::max(c1 ? m1 : 0.0, c2 ? m2 : 0.0) : std::min(m1, m2);
    }
  }
  r.margin = accept ? std::min(feas_margin, acc_margin) : acc_margin;
  if (own_reject) r.margin = reject_margin;
  if (!accept && !force) return;
  r.success = accept && !own_reject;
  r.cost = cost_new;
  r.merit = phi_new;
  r.theta = theta_new;
  r.inf_pr = ipn;
  r.inf_comp = icn;
}

void ip_update_barrier(IpState &s) { /* updateBarrierParameters(context, true) (:2548-2660) */
  const bool no_barrier = s.nc == 0;
  const double scaled_inf_du = s.inf_du; /* computeScaledDualInfeasibility, check_state_stationarity = false (:2725-2733) */
  const double mu_old = s.mu;
  if (no_barrier) {
    s.mu = mu_old;
  } else if (s.io->barrier_strategy == ORACLE_BARRIER_ADAPTIVE) {
    const double kkt = std::max(std::max(s.inf_pr, scaled_inf_du), s.inf_comp);
    const double threshold = std::max(s.io->mu_update_factor * s.mu, 2.0 * s.mu);
    if (kkt <= threshold) {
      double factor = s.io->mu_update_factor;
      if (s.mu > 1e-20) {
        const double ratio = kkt / std::max(s.mu, 1e-20);
        if (ratio < 0.01) factor = 0.1 * s.io->mu_update_factor;
        else if (ratio < 0.1) factor = 0.3 * s.io->mu_update_factor;
        else if (ratio < 0.5) factor = 0.6 * s.io->mu_update_factor;
      }
      const double linear = factor * s.mu;
      const double superlinear = std::pow(s.mu, s.io->mu_update_power);
      s.mu = std::max(std::min(linear, superlinear), std::max(s.io->mu_min_value, s.o->tolerance / 100.0));
    }
  } else {
    const double kkt = std::max(std::max(s.inf_pr, scaled_inf_du * s.io->barrier_update_dual_weight), s.inf_comp);
    if (kkt <= s.io->mu_kappa_epsilon * s.mu) {
      const double linear = s.io->mu_update_factor * s.mu;
      const double superlinear = std::pow(s.mu, s.io->mu_update_power);
      s.mu = std::max(s.io->mu_min_value, std::min(linear, superlinear));
    }
  }
  double hT[MAXN];
  if (s.teq) ip_terminal_residual(s, s.X.data(), hT);
  const double *h = s.teq ? hT : nullptr;
  const double filter_theta = std::max(ip_theta(s, s.G.data(), s.S.data(), h), 1e-8);
  const bool reset_filter = (s.mu < mu_old) && (s.mu > 0.0);
  if (reset_filter) {
    s.filter.clear();
    if (s.teq) accept_filter_entry(s.filter, s.phi, filter_theta); /* :2633-2636 */
  } else {
    accept_filter_entry(s.filter, s.phi, filter_theta);
    if ((int)s.filter.size() > s.io->max_filter_size) prune_filter(s.filter);
  }
  ip_primal_comp(s, s.G.data(), s.S.data(), s.Y.data(), s.mu, &s.inf_pr, &s.inf_comp, h);
  s.merit = ip_merit(s, s.S.data(), s.cost, s.teq ? s.lamT.data() : nullptr, h);
  s.phi = s.merit;
  s.filter_theta = filter_theta;
  s.theta = std::max(filter_theta, std::max(s.io->theta_0_floor, 1e-8));
}

/* ONE entry of the main loop of CDDPSolverBase::solve (cddp_solver_base.cpp:74-170) with the IPDDP hooks, on the solver
 * state `s`.  follow < 0: the oracle takes its own decisions (this IS the loop body of ipddp_solve_one).  follow >= 0
 * (test instrumentation, see iterate_once above): a recorded decision ((backward failures << 8) | code) and the recorded
 * status after the iteration; every branch on a backward-failure / line-search / convergence verdict takes the RECORDED
 * side while all arithmetic stays the oracle's, and R reports where its own verdict differed and how close to its
 * threshold the tested quantity was (the line-search margin is the trial's own accept/reject margin).
 * Returns the decision in *code_out, the status after the iteration in *status_out (ORACLE_RUNNING = goes on). */
thread_local double *g_ip_trial_table = nullptr; /* debug: [na][6] success, cost, merit, theta, margin, alpha_pr of every evaluated candidate */

void ip_iterate_once(IpState &s, int iter, const double *alphas, int na, int follow, int follow_status, int *code_out,
                     int *status_out, double *min_margin, oracle_replay_report &R, const std::function<void()> &record) {
  const oracle_options *o = s.o;
  const oracle_ipddp_options *io = s.io;
  const bool no_barrier = s.nc == 0;
  const bool rec = follow >= 0;
  const int rcode = rec ? (follow & 0xff) : 0, rfail = rec ? (follow >> 8) : 0;
  auto note = [&](double mg, int kind = 0) { /* kind (reserved): 1 early exit, 2 line search, 3 convergence, 4 failure status */
    ++R.n_disagree;
    if (!(mg <= R.max_margin)) {
      R.max_margin = mg;
      R.reserved = kind;
    }
  };
  auto near = [](double a, double b) { return std::fabs(a - b) / std::max(std::max(std::fabs(a), std::fabs(b)), 1e-300); };
  *status_out = ORACLE_RUNNING;
  bool backward_ok = false; /* cddp_solver_base.cpp:93-111 */
  int bw_fail = 0;
  while (true) {
    if (rec && bw_fail == rfail && rcode == ORACLE_TRACE_BW_LIMIT) break;
    backward_ok = ip_backward(s);
    if (rec) {
      if (bw_fail < rfail) {
        if (backward_ok) ++R.n_backward_disagree;
        backward_ok = false;
      } else if (!backward_ok) {
        ++R.n_backward_disagree;
        R.infeasible = 1;
        break;
      }
    }
    if (backward_ok) break;
    ++bw_fail;
    s.reg = std::min(s.reg * o->reg_update_factor, o->reg_max_value);
    if (!rec && s.reg >= o->reg_max_value) break;
  }
  if (!backward_ok) {
    *code_out = (bw_fail << 8) | ORACLE_TRACE_BW_LIMIT;
    *status_out = ORACLE_REG_LIMIT;
    return;
  }
  { /* checkEarlyConvergence (:925-958) */
    bool own;
    double mg;
    if (no_barrier) {
      own = s.inf_pr < o->tolerance && s.inf_du < o->tolerance;
      mg = std::min(near(s.inf_pr, o->tolerance), near(s.inf_du, o->tolerance));
    } else {
      const double tol = std::max(o->tolerance, io->barrier_tol_mult * s.mu);
      own = s.inf_pr < tol && s.inf_du < tol && s.inf_comp < tol && std::fabs(s.alpha_pr) * s.step_norm < o->tolerance * 10.0;
      mg = std::min(std::min(near(s.inf_pr, tol), near(s.inf_du, tol)),
                    std::min(near(s.inf_comp, tol), near(std::fabs(s.alpha_pr) * s.step_norm, o->tolerance * 10.0)));
    }
    const bool early = rec ? rcode == ORACLE_TRACE_EARLY_EXIT : own;
    if (own != early) note(mg, 1);
    if (early) {
      record();
      *code_out = (bw_fail << 8) | ORACLE_TRACE_EARLY_EXIT;
      *status_out = ORACLE_OPTIMAL;
      return;
    }
  }
  IpTrial trial, cand;
  bool fp = false; /* performForwardPass: sequential = first success (cddp_solver_base.cpp:255-263) */
  int acc = -1;
  const int racc = rcode - 1;
  if (!o->enable_parallel) {
    const int upto = rec ? (racc >= 0 ? racc : na - 1) : na - 1;
    for (int ai = 0; ai <= upto; ++ai) {
      ip_forward(s, alphas[ai], trial, rec && ai == racc);
      *min_margin = std::min(*min_margin, trial.margin);
      if (g_ip_trial_table) {
        double *row = g_ip_trial_table + (size_t)ai * 6;
        row[0] = trial.success ? 1.0 : 0.0; row[1] = trial.cost; row[2] = trial.merit; row[3] = trial.theta; row[4] = trial.margin;
        row[5] = trial.alpha_pr;
      }
      const bool take = rec ? ai == racc : trial.success;
      if (trial.success != take) note(trial.margin, 2 + 16 * ai);
      if (take) {
        fp = true;
        acc = ai;
        break;
      }
    }
  } else { /* enable_parallel: the success with the strictly lowest merit, scanning in alpha order (:264-285) */
    double best_merit = std::numeric_limits<double>::infinity();
    int own = -1;
    for (int ai = 0; ai < na; ++ai) {
      ip_forward(s, alphas[ai], cand, rec && ai == racc);
      *min_margin = std::min(*min_margin, cand.margin);
      if (cand.success && cand.merit < best_merit) {
        best_merit = cand.merit;
        own = ai;
        if (!rec) trial = cand;
      }
      if (rec && ai == racc) trial = cand;
    }
    acc = rec ? racc : own;
    fp = acc >= 0;
    if (rec && own != racc) note(racc >= 0 && own >= 0 ? near(trial.merit, best_merit) : 1.0, 2);
  }
  *code_out = (bw_fail << 8) | (fp ? 1 + acc : ORACLE_TRACE_LS_FAILED);
  if (fp) {
    const double dJ = s.cost - trial.cost;
    /* applyForwardPassResult (:1878-1951) */
    s.X = trial.X; s.U = trial.U; s.cost = trial.cost; s.merit = trial.merit;
    s.alpha_pr = trial.alpha_pr; s.alpha_du = trial.alpha_du;
    s.Y = trial.Y; s.S = trial.S; s.G = trial.G; s.lamT = trial.lamT;
    s.inf_pr = trial.inf_pr; s.inf_comp = trial.inf_comp;
    s.phi = trial.merit; s.filter_theta = trial.theta; s.theta = trial.theta;
    ip_update_barrier(s);
    record();
    s.reg = std::max(s.reg / o->reg_update_factor, o->reg_min_value);
    /* checkConvergence (:1953-2025) */
    int own = ORACLE_RUNNING;
    double mg = 1.0;
    if (no_barrier) {
      mg = std::min(near(s.inf_pr, o->tolerance), near(s.inf_du, o->tolerance));
      if (s.inf_pr < o->tolerance && s.inf_du < o->tolerance) {
        own = ORACLE_OPTIMAL;
      } else if (o->acceptable_tolerance > 0.0) {
        const double sq = std::sqrt(o->acceptable_tolerance);
        bool a = s.inf_pr < sq && s.inf_du < sq && iter > 50;
        if (dJ > 0.0) a = a || (dJ < o->acceptable_tolerance && iter > 50 && s.inf_pr < sq && s.inf_du < sq);
        if (a) own = ORACLE_ACCEPTABLE;
        mg = std::min(mg, std::min(near(s.inf_pr, sq), near(s.inf_du, sq)));
      }
    } else {
      const double tol = std::max(o->tolerance, io->barrier_tol_mult * s.mu);
      mg = std::min(std::min(near(s.inf_pr, tol), near(s.inf_du, tol)), std::min(near(s.inf_comp, tol), near(s.step_norm, o->tolerance * 10.0)));
      if (s.inf_pr < tol && s.inf_du < tol && s.inf_comp < tol && s.step_norm < o->tolerance * 10.0) {
        own = ORACLE_OPTIMAL;
      } else if (o->acceptable_tolerance > 0.0) {
        const double at = std::sqrt(o->acceptable_tolerance);
        const double bat = std::max(io->mu_min_value * 100.0, o->tolerance / 10.0);
        const bool kkt = s.inf_pr < at && s.inf_du < at && s.inf_comp < at;
        const bool done = s.mu <= bat;
        bool a = kkt && done && iter > 10 && std::fabs(dJ) < o->acceptable_tolerance;
        a = a || (kkt && done && iter >= 1 && s.step_norm < o->tolerance * 10.0 && s.inf_pr < 1e-4);
        if (a) own = ORACLE_ACCEPTABLE;
        mg = std::min(mg, std::min(std::min(near(s.inf_pr, at), near(s.inf_du, at)), std::min(near(s.inf_comp, at), near(s.mu, bat))));
        mg = std::min(mg, std::min(near(std::fabs(dJ), o->acceptable_tolerance), near(s.inf_pr, 1e-4)));
      }
    }
    const int taken = rec ? ((follow_status == ORACLE_OPTIMAL || follow_status == ORACLE_ACCEPTABLE) ? follow_status : ORACLE_RUNNING) : own;
    if (taken != own) note(mg, 3);
    *status_out = taken;
  } else { /* handleForwardPassFailure (:2037-2082) */
    s.reg = std::min(s.reg * o->reg_update_factor, o->reg_max_value);
    if (!no_barrier && s.teq) s.reg = std::min(s.reg * o->reg_update_factor, o->reg_max_value); /* :2047-2051 */
    if (s.reg >= o->reg_max_value) {
      const double base = std::sqrt(std::max(o->acceptable_tolerance, o->tolerance));
      const double at = no_barrier ? base : std::max(base, io->barrier_tol_mult * s.mu);
      const bool a = o->acceptable_tolerance > 0.0 && s.inf_pr < at && s.inf_du < at && (no_barrier || s.inf_comp < at);
      const int own = a ? ORACLE_ACCEPTABLE : ORACLE_REG_LIMIT;
      const int taken = rec && (follow_status == ORACLE_ACCEPTABLE || follow_status == ORACLE_REG_LIMIT) ? follow_status : own;
      if (taken != own) note(std::min(std::min(near(s.inf_pr, at), near(s.inf_du, at)), near(s.inf_comp, at)), 4);
      *status_out = taken;
    }
  }
}

/* CDDP::solve("IPDDP"): CDDPSolverBase::solve (cddp_solver_base.cpp:29-186) with the IPDDP hooks */
void ipddp_solve_one(const oracle_problem *p, const oracle_options *o, const oracle_ipddp_options *io,
                     const oracle_constraint *cs, int nc, const double *x0, const double *xref, const double *ref_traj,
                     double *X, double *U, double *K, double *Yout, double *Sout, oracle_ipddp_result *res,
                     double *history) {
  IpState s;
  s.p = p; s.o = o; s.io = io; s.cs = cs; s.nc = nc;
  s.n = p->n; s.m = p->m; s.N = p->horizon; s.d = total_dual_dim(p, cs, nc);
  s.x0 = x0; s.xref = xref; s.ref_traj = ref_traj;
  double alphas[ORACLE_MAX_ALPHAS];
  const int na = build_alphas(o, alphas);
  ip_initialize(s, U);
  int hl = 0;
  std::function<void()> record = [&]() { /* recordIterationHistory (cddp_solver_base.cpp:220-232 + ipddp :2084-2088) */
    if (!history) return;
    double *h = history + (size_t)hl * ORACLE_IPDDP_HISTORY_COLS;
    h[0] = s.cost; h[1] = s.merit; h[2] = s.alpha_pr; h[3] = s.alpha_du; h[4] = s.inf_du;
    h[5] = s.inf_pr; h[6] = s.inf_comp; h[7] = s.reg; h[8] = s.mu;
    ++hl;
  };
  record();
  int iter = 0, status = ORACLE_MAX_ITERATIONS;
  double min_margin = 1.0;
  oracle_replay_report R;
  std::memset(&R, 0, sizeof(R));
  while (iter < o->max_iterations) {
    ++iter;
    int code = 0, st = ORACLE_RUNNING;
    ip_iterate_once(s, iter, alphas, na, -1, ORACLE_RUNNING, &code, &st, &min_margin, R, record);
    if (st != ORACLE_RUNNING) {
      status = st;
      break;
    }
  }
  std::memcpy(X, s.X.data(), sizeof(double) * s.X.size());
  std::memcpy(U, s.U.data(), sizeof(double) * s.U.size());
  if (K) std::memcpy(K, s.Ku.data(), sizeof(double) * s.Ku.size());
  if (Yout && s.d) std::memcpy(Yout, s.Y.data(), sizeof(double) * s.Y.size());
  if (Sout && s.d) std::memcpy(Sout, s.S.data(), sizeof(double) * s.S.size());
  res->final_objective = s.cost;
  res->final_step_length = s.alpha_pr;
  res->final_regularization = s.reg;
  res->inf_du = s.inf_du;
  res->inf_pr = s.inf_pr;
  res->inf_comp = s.inf_comp;
  res->mu = s.mu;
  res->merit = s.merit;
  res->decision_margin = min_margin;
  res->iterations = iter;
  res->status = status;
  res->history_len = hl;
  res->dual_dim = s.d;
}

} /* namespace */

extern "C" {

void oracle_default_options(oracle_options *o) { /* options.hpp:41-66,93-105,208-251; boxqp.hpp:30-41 */
  std::memset(o, 0, sizeof(*o));
  o->tolerance = 1e-5;
  o->acceptable_tolerance = 1e-6;
  o->max_iterations = 1;
  o->max_cpu_time = 0.0;
  o->termination_scaling_max_factor = 100.0;
  o->ls_max_iterations = 11;
  o->ls_initial_step_size = 1.0;
  o->ls_min_step_size = 1e-8;
  o->ls_step_reduction_factor = 0.5;
  o->reg_initial_value = 1e-6;
  o->reg_update_factor = 10.0;
  o->reg_max_value = 1e7;
  o->reg_min_value = 1e-10;
  o->qp_max_iterations = 100;
  o->qp_min_gradient_norm = 1e-8;
  o->qp_min_relative_improvement = 1e-8;
  o->qp_step_decrease_factor = 0.6;
  o->qp_min_step_size = 1e-22;
  o->qp_armijo_constant = 0.1;
  o->armijo_constant = 1e-4;
}

int oracle_build_alphas(const oracle_options *o, double *alphas) { return build_alphas(o, alphas); }

void oracle_continuous_dynamics(const oracle_problem *p, const double *x, const double *u, double t, double *xdot) {
  continuous_dynamics(p, x, u, t, xdot);
}
void oracle_discrete_dynamics(const oracle_problem *p, const double *x, const double *u, double t, double *xnext) {
  discrete_dynamics(p, x, u, t, xnext);
}
void oracle_jacobians(const oracle_problem *p, const double *x, const double *u, double t, double *Fx, double *Fu) {
  jacobians(p, x, u, t, Fx, Fu);
}
double oracle_running_cost(const oracle_problem *p, const double *x, const double *u, const double *ref) {
  return running_cost(p, x, u, ref);
}
double oracle_terminal_cost(const oracle_problem *p, const double *x, const double *ref) {
  return terminal_cost(p, x, ref);
}
double oracle_trajectory_cost(const oracle_problem *p, const double *X, const double *U, const double *xref,
                              const double *ref_traj) {
  return trajectory_cost(p, X, U, xref, ref_traj);
}

int oracle_boxqp(const oracle_options *o, int n, const double *H, const double *g, const double *lower,
                 const double *upper, const double *x0, double *x, int *free_mask, int *iterations,
                 int *factorizations, double *final_value, double *final_grad_norm, int ncols, const double *Kfree_rhs,
                 double *Kfree_out) {
  BoxQPOut r;
  boxqp_solve(o, n, H, g, lower, upper, x0, &r);
  for (int i = 0; i < n; ++i) {
    x[i] = r.x[i];
    free_mask[i] = r.free_mask[i];
  }
  if (iterations) *iterations = r.iterations;
  if (factorizations) *factorizations = r.factorizations;
  if (final_value) *final_value = r.final_value;
  if (final_grad_norm) *final_grad_norm = r.final_grad_norm;
  if (Kfree_rhs && Kfree_out && ncols > 0) {
    int free_idx[MAXM], nf = 0;
    for (int i = 0; i < n; ++i)
      if (r.free_mask[i]) free_idx[nf++] = i;
    std::fill(Kfree_out, Kfree_out + n * ncols, 0.0);
    if (nf > 0) {
      double Hf[MAXM * MAXM];
      for (int i = 0; i < nf; ++i)
        for (int j = 0; j < nf; ++j) Hf[i * nf + j] = H[free_idx[i] * n + free_idx[j]];
      LDLT f;
      f.compute(Hf, nf);
      for (int c = 0; c < ncols; ++c) {
        double col[MAXM];
        for (int i = 0; i < nf; ++i) col[i] = Kfree_rhs[free_idx[i] * ncols + c];
        f.solve_inplace(col);
        for (int i = 0; i < nf; ++i) Kfree_out[free_idx[i] * ncols + c] = col[i];
      }
    }
  }
  return r.status;
}

int oracle_backward_pass(const oracle_problem *p, const oracle_options *o, const double *X, const double *U,
                         const double *xref, const double *ref_traj, double reg, double *K, double *k, double *dV,
                         double *inf_du, double *Vx_dbg, double *Vxx_dbg, int *fail_t) {
  return backward_pass(p, o, nullptr, nullptr, X, U, xref, ref_traj, reg, K, k, dV, inf_du, Vx_dbg, Vxx_dbg, fail_t);
}

int oracle_backward_pass_AB(const oracle_problem *p, const oracle_options *o, const double *A, const double *B,
                            const double *X, const double *U, const double *xref, const double *ref_traj, double reg,
                            double *K, double *k, double *dV, double *inf_du, double *Vx_dbg, double *Vxx_dbg,
                            int *fail_t) {
  return backward_pass(p, o, A, B, X, U, xref, ref_traj, reg, K, k, dV, inf_du, Vx_dbg, Vxx_dbg, fail_t);
}

void oracle_linearize(const oracle_problem *p, const double *X, const double *U, double *A, double *B) {
  const int n = p->n, m = p->m, N = p->horizon;
  double Fx[MAXN * MAXN], Fu[MAXN * MAXM];
  for (int t = 0; t < N; ++t) {
    jacobians(p, X + (size_t)t * n, U + (size_t)t * m, t * p->dt, Fx, Fu);
    double *At = A + (size_t)t * n * n, *Bt = B + (size_t)t * n * m;
    for (int i = 0; i < n * n; ++i) At[i] = p->dt * Fx[i];
    for (int i = 0; i < n; ++i) At[i * n + i] += 1.0;
    for (int i = 0; i < n * m; ++i) Bt[i] = p->dt * Fu[i];
  }
}

int oracle_forward_pass(const oracle_problem *p, const oracle_options *o, const double *x0, const double *X,
                        const double *U, const double *xref, const double *ref_traj, const double *K, const double *k,
                        const double *dV, double cost, double alpha, double *Xn, double *Un, double *Jn) {
  return forward_pass(p, o, x0, X, U, xref, ref_traj, K, k, dV, cost, alpha, Xn, Un, Jn);
}

void oracle_solve(const oracle_problem *p, const oracle_options *o, const double *x0, const double *xref,
                  const double *ref_traj, double *X, double *U, double *K, double *k, oracle_result *res,
                  double *history) {
  solve_one(p, o, x0, xref, ref_traj, X, U, K, k, res, history);
}

void oracle_solve_batch(const oracle_problem *p, const oracle_options *o, int batch, int nthreads, const double *x0,
                        const double *xref, const double *ref_traj, double *X, double *U, double *K, double *k,
                        oracle_result *res) {
  oracle_solve_batch_traced(p, o, batch, nthreads, x0, xref, ref_traj, X, U, K, k, res, nullptr, nullptr, nullptr, nullptr,
                            nullptr, nullptr);
}

void oracle_solve_batch_traced(const oracle_problem *p, const oracle_options *o, int batch, int nthreads, const double *x0,
                               const double *xref, const double *ref_traj, double *X, double *U, double *K, double *k,
                               oracle_result *res, double *history, int *trace_out, const int *replay_trace,
                               const int *replay_iterations, const int *replay_status, oracle_replay_report *rep) {
  const int n = p->n, m = p->m, N = p->horizon;
  const size_t hstride = (size_t)(o->max_iterations + 1) * 4, tstride = (size_t)(o->max_iterations > 0 ? o->max_iterations : 1);
  if (nthreads < 1) nthreads = 1;
  if (nthreads > batch) nthreads = batch > 0 ? batch : 1;
  auto work = [&](int tid) {
    const int lo = (int)((long long)batch * tid / nthreads), hi = (int)((long long)batch * (tid + 1) / nthreads);
    for (int b = lo; b < hi; ++b) {
      Replay rp{nullptr, 0, 0};
      if (replay_trace) rp = Replay{replay_trace + (size_t)b * tstride, replay_iterations[b], replay_status[b]};
      solve_one(p, o, x0 + (size_t)b * n, xref + (size_t)b * n,
                ref_traj ? ref_traj + (size_t)b * (N + 1) * n : nullptr, X + (size_t)b * (N + 1) * n,
                U + (size_t)b * N * m, K + (size_t)b * N * m * n, k + (size_t)b * N * m, res + b,
                history ? history + (size_t)b * hstride : nullptr, trace_out ? trace_out + (size_t)b * tstride : nullptr,
                replay_trace ? &rp : nullptr, rep ? rep + b : nullptr);
    }
  };
  if (nthreads == 1) {
    work(0);
    return;
  }
  std::vector<std::thread> th;
  for (int t = 0; t < nthreads; ++t) th.emplace_back(work, t);
  for (auto &t : th) t.join();
}

void oracle_iterate_batch(const oracle_problem *p, const oracle_options *o, int batch, int nthreads, const double *x0,
                          const double *xref, const double *ref_traj, double *X, double *U, double *K, double *k,
                          double *reg, double *cost, double *alpha, double *inf_du, double *dV, const int *follow,
                          const int *follow_status, int *code, int *status, oracle_replay_report *rep) {
  const int n = p->n, m = p->m, N = p->horizon;
  double alphas[ORACLE_MAX_ALPHAS];
  const int na = build_alphas(o, alphas);
  if (nthreads < 1) nthreads = 1;
  if (nthreads > batch) nthreads = batch > 0 ? batch : 1;
  auto work = [&](int tid) {
    const int lo = (int)((long long)batch * tid / nthreads), hi = (int)((long long)batch * (tid + 1) / nthreads);
    std::vector<double> Xn((size_t)(N + 1) * n), Un((size_t)N * m), Xt((size_t)(N + 1) * n), Ut((size_t)N * m);
    for (int b = lo; b < hi; ++b) {
      IterState st{X + (size_t)b * (N + 1) * n, U + (size_t)b * N * m, K + (size_t)b * N * m * n, k + (size_t)b * N * m,
                   reg[b], cost[b], alpha[b], inf_du[b], {0.0, 0.0}};
      oracle_replay_report R;
      std::memset(&R, 0, sizeof(R));
      double reg_used = 0.0;
      iterate_once(p, o, alphas, na, x0 + (size_t)b * n, xref + (size_t)b * n,
                   ref_traj ? ref_traj + (size_t)b * (N + 1) * n : nullptr, st, follow ? follow[b] : -1,
                   follow_status ? follow_status[b] : ORACLE_RUNNING, code + b, status + b, &reg_used, R, Xn, Un, Xt, Ut);
      reg[b] = st.reg;
      cost[b] = st.cost;
      alpha[b] = st.alpha_pr;
      inf_du[b] = st.inf_du;
      if (dV) {
        dV[2 * b] = st.dV[0];
        dV[2 * b + 1] = st.dV[1];
      }
      if (rep) rep[b] = R;
    }
  };
  if (nthreads == 1) {
    work(0);
    return;
  }
  std::vector<std::thread> th;
  for (int t = 0; t < nthreads; ++t) th.emplace_back(work, t);
  for (auto &t : th) t.join();
}

void oracle_ipddp_default_options(oracle_ipddp_options *io) { /* options.hpp:75-104,148-186 */
  std::memset(io, 0, sizeof(*io));
  io->dual_var_init_scale = 1e-1;
  io->slack_var_init_scale = 1e-2;
  io->barrier_tol_mult = 0.1;
  io->barrier_update_dual_weight = 0.01;
  io->mu_kappa_epsilon = 10.0;
  io->theta_0_floor = 1.0;
  io->mu_initial = 1.0;
  io->mu_min_value = 1e-10;
  io->mu_update_factor = 0.5;
  io->mu_update_power = 1.2;
  io->min_fraction_to_boundary = 0.99;
  io->merit_acceptance_threshold = 1e-6;
  io->violation_acceptance_threshold = 1e-6;
  io->max_violation_threshold = 1e4;
  io->min_violation_for_armijo_check = 1e-7;
  io->theta_norm_l2 = 0;
  io->max_filter_size = 5;
  io->barrier_strategy = ORACLE_BARRIER_ADAPTIVE;
  io->jacobian_regularization_value = 1e-8;
  io->jacobian_regularization_exponent = 0.25;
  io->terminal_equality = 0;
}

int oracle_total_dual_dim(const oracle_problem *p, const oracle_constraint *cs, int nc) { return total_dual_dim(p, cs, nc); }

void oracle_eval_constraints(const oracle_problem *p, const oracle_constraint *cs, int nc, const double *x,
                             const double *u, double *g, double *Gx, double *Gu) {
  if (g) eval_constraints(p, cs, nc, x, u, g);
  if (Gx && Gu) constraint_jacobians(p, cs, nc, x, Gx, Gu);
}

void oracle_ipddp_solve(const oracle_problem *p, const oracle_options *o, const oracle_ipddp_options *io,
                        const oracle_constraint *cs, int nc, const double *x0, const double *xref,
                        const double *ref_traj, double *X, double *U, double *K, double *Y, double *S,
                        oracle_ipddp_result *res, double *history) {
  ipddp_solve_one(p, o, io, cs, nc, x0, xref, ref_traj, X, U, K, Y, S, res, history);
}

void oracle_ipddp_solve_batch(const oracle_problem *p, const oracle_options *o, const oracle_ipddp_options *io,
                              const oracle_constraint *cs, int nc, int batch, int nthreads, const double *x0,
                              const double *xref, const double *ref_traj, double *X, double *U, double *K, double *Y,
                              double *S, oracle_ipddp_result *res) {
  const int n = p->n, m = p->m, N = p->horizon, d = total_dual_dim(p, cs, nc);
  if (nthreads < 1) nthreads = 1;
  if (nthreads > batch) nthreads = batch;
  auto work = [&](int tid) {
    const int lo = (int)((long long)batch * tid / nthreads), hi = (int)((long long)batch * (tid + 1) / nthreads);
    for (int b = lo; b < hi; ++b)
      ipddp_solve_one(p, o, io, cs, nc, x0 + (size_t)b * n, xref + (size_t)b * n,
                      ref_traj ? ref_traj + (size_t)b * (N + 1) * n : nullptr, X + (size_t)b * (N + 1) * n,
                      U + (size_t)b * N * m, K ? K + (size_t)b * N * m * n : nullptr,
                      Y ? Y + (size_t)b * N * d : nullptr, S ? S + (size_t)b * N * d : nullptr, res + b, nullptr);
  };
  std::vector<std::thread> th;
  for (int t = 1; t < nthreads; ++t) th.emplace_back(work, t);
  work(0);
  for (auto &t : th) t.join();
}

void oracle_ipddp_probe(const oracle_problem *p, const oracle_options *o, const oracle_ipddp_options *io,
                        const oracle_constraint *cs, int nc, const double *x0, const double *xref,
                        const double *ref_traj, const double *U0, int iters, double *X, double *U, double *Y, double *S,
                        double *G, double *ku, double *Ku, double *ky, double *Ky, double *ks, double *Ks, double *dS,
                        double *dY, double *scalars, double *trial_costs) {
  IpState s;
  s.p = p; s.o = o; s.io = io; s.cs = cs; s.nc = nc;
  s.n = p->n; s.m = p->m; s.N = p->horizon; s.d = total_dual_dim(p, cs, nc);
  s.x0 = x0; s.xref = xref; s.ref_traj = ref_traj;
  double alphas[ORACLE_MAX_ALPHAS];
  const int na = build_alphas(o, alphas);
  ip_initialize(s, U0);
  IpTrial trial;
  bool ok = true;
  for (int it = 0; it <= iters && ok; ++it) {
    ok = ip_backward(s);
    if (!ok || it == iters) break;
    bool fp = false;
    for (int ai = 0; ai < na; ++ai) {
      ip_forward(s, alphas[ai], trial);
      if (trial.success) { fp = true; break; }
    }
    if (fp) {
      s.X = trial.X; s.U = trial.U; s.cost = trial.cost; s.merit = trial.merit;
      s.alpha_pr = trial.alpha_pr; s.alpha_du = trial.alpha_du;
      s.Y = trial.Y; s.S = trial.S; s.G = trial.G; s.lamT = trial.lamT;
      s.inf_pr = trial.inf_pr; s.inf_comp = trial.inf_comp;
      s.phi = trial.merit; s.filter_theta = trial.theta; s.theta = trial.theta;
      ip_update_barrier(s);
      s.reg = std::max(s.reg / o->reg_update_factor, o->reg_min_value);
    } else {
      s.reg = std::min(s.reg * o->reg_update_factor, o->reg_max_value);
      if (s.nc != 0 && s.teq) s.reg = std::min(s.reg * o->reg_update_factor, o->reg_max_value);
    }
  }
  auto cp = [](double *dst, const std::vector<double> &v) { if (dst && !v.empty()) std::memcpy(dst, v.data(), sizeof(double) * v.size()); };
  cp(X, s.X); cp(U, s.U); cp(Y, s.Y); cp(S, s.S); cp(G, s.G); cp(ku, s.ku); cp(Ku, s.Ku);
  cp(ky, s.ky); cp(Ky, s.Ky); cp(ks, s.ks); cp(Ks, s.Ks); cp(dS, s.dS); cp(dY, s.dY);
  if (trial_costs) {
    for (int ai = 0; ai < na; ++ai) {
      ip_forward(s, alphas[ai], trial);
      trial_costs[ai * 4 + 0] = trial.success ? 1.0 : 0.0;
      trial_costs[ai * 4 + 1] = trial.cost;
      trial_costs[ai * 4 + 2] = trial.merit;
      trial_costs[ai * 4 + 3] = trial.theta;
    }
  }
  if (scalars) {
    const double tau_b = std::max(io->min_fraction_to_boundary, 1.0 - s.mu);
    double apm = 1.0, adm = 1.0;
    for (size_t q = 0; q < s.dS.size(); ++q) {
      if (s.dS[q] < 0.0) apm = std::min(apm, -tau_b * s.S[q] / s.dS[q]);
      if (s.dY[q] < 0.0) adm = std::min(adm, -tau_b * s.Y[q] / s.dY[q]);
    }
    scalars[0] = s.mu; scalars[1] = s.cost; scalars[2] = s.merit; scalars[3] = s.inf_pr; scalars[4] = s.inf_du;
    scalars[5] = s.inf_comp; scalars[6] = s.step_norm; scalars[7] = s.reg; scalars[8] = s.dV[0]; scalars[9] = s.dV[1];
    scalars[10] = clampd(apm, 0.0, 1.0); scalars[11] = clampd(adm, 0.0, 1.0); scalars[12] = s.filter_theta;
    scalars[13] = s.theta; scalars[14] = (double)s.filter.size(); scalars[15] = ok ? 1.0 : 0.0;
  }
}

/* ONE IPDDP main-loop entry per instance from a caller-supplied solver state (lock-step parity tests).  In / out, per
 * instance: X [N+1][n], U [N][m], Y, S, G [N][d], lamT [n], filter [8][2] (merit, theta) + filter_size, scalars [12] =
 * {mu, cost, merit, filter_theta, inf_pr, inf_comp, reg, alpha_pr, alpha_du, step_norm, inf_du, iteration index (1-based, in)}.
 * follow / follow_status == NULL: own decisions. */
void oracle_ipddp_iterate_batch(const oracle_problem *p, const oracle_options *o, const oracle_ipddp_options *io,
                                const oracle_constraint *cs, int nc, int batch, int nthreads, const double *x0, const double *xref,
                                const double *ref_traj, double *X, double *U, double *Y, double *S, double *G, double *lamT,
                                double *filter, int *filter_size, double *scalars, const int *follow, const int *follow_status,
                                int *code, int *status, oracle_replay_report *rep, double *trial_table) {
  const int n = p->n, m = p->m, N = p->horizon, d = total_dual_dim(p, cs, nc);
  double alphas[ORACLE_MAX_ALPHAS];
  const int na = build_alphas(o, alphas);
  auto work = [&](int lo, int hi) {
    for (int b = lo; b < hi; ++b) {
      IpState s;
      s.p = p; s.o = o; s.io = io; s.cs = cs; s.nc = nc;
      s.n = n; s.m = m; s.N = N; s.d = d;
      s.x0 = x0 + (size_t)b * n; s.xref = xref + (size_t)b * n;
      s.ref_traj = ref_traj ? ref_traj + (size_t)b * (N + 1) * n : nullptr;
      ip_initialize(s, U + (size_t)b * N * m); /* sizes every array and sets the flags; the state proper is overwritten below */
      double *sc = scalars + (size_t)b * 12;
      s.X.assign(X + (size_t)b * (N + 1) * n, X + (size_t)(b + 1) * (N + 1) * n);
      s.U.assign(U + (size_t)b * N * m, U + (size_t)(b + 1) * N * m);
      if (d) {
        s.Y.assign(Y + (size_t)b * N * d, Y + (size_t)(b + 1) * N * d);
        s.S.assign(S + (size_t)b * N * d, S + (size_t)(b + 1) * N * d);
        s.G.assign(G + (size_t)b * N * d, G + (size_t)(b + 1) * N * d);
      }
      if (s.teq) s.lamT.assign(lamT + (size_t)b * n, lamT + (size_t)(b + 1) * n);
      s.filter.clear();
      for (int k = 0; k < filter_size[b]; ++k) {
        FilterPt fpnt;
        fpnt.merit = filter[((size_t)b * 8 + k) * 2];
        fpnt.theta = filter[((size_t)b * 8 + k) * 2 + 1];
        s.filter.push_back(fpnt);
      }
      s.mu = sc[0]; s.cost = sc[1]; s.merit = sc[2]; s.phi = sc[2]; s.filter_theta = sc[3];
      s.theta = std::max(s.filter_theta, std::max(io->theta_0_floor, 1e-8));
      s.inf_pr = sc[4]; s.inf_comp = sc[5]; s.reg = sc[6]; s.alpha_pr = sc[7]; s.alpha_du = sc[8]; s.step_norm = sc[9]; s.inf_du = sc[10];
      const int iter = (int)sc[11];
      oracle_replay_report R;
      std::memset(&R, 0, sizeof(R));
      double mm = 1.0;
      std::function<void()> norecord = []() {};
      g_ip_trial_table = trial_table ? trial_table + (size_t)b * ORACLE_MAX_ALPHAS * 6 : nullptr;
      ip_iterate_once(s, iter, alphas, na, follow ? follow[b] : -1, follow_status ? follow_status[b] : ORACLE_RUNNING, &code[b], &status[b],
                      &mm, R, norecord);
      g_ip_trial_table = nullptr;
      std::memcpy(X + (size_t)b * (N + 1) * n, s.X.data(), sizeof(double) * (N + 1) * n);
      std::memcpy(U + (size_t)b * N * m, s.U.data(), sizeof(double) * N * m);
      if (d) {
        std::memcpy(Y + (size_t)b * N * d, s.Y.data(), sizeof(double) * N * d);
        std::memcpy(S + (size_t)b * N * d, s.S.data(), sizeof(double) * N * d);
        std::memcpy(G + (size_t)b * N * d, s.G.data(), sizeof(double) * N * d);
      }
      if (s.teq) std::memcpy(lamT + (size_t)b * n, s.lamT.data(), sizeof(double) * n);
      filter_size[b] = (int)s.filter.size();
      for (int k = 0; k < (int)s.filter.size() && k < 8; ++k) {
        filter[((size_t)b * 8 + k) * 2] = s.filter[k].merit;
        filter[((size_t)b * 8 + k) * 2 + 1] = s.filter[k].theta;
      }
      sc[0] = s.mu; sc[1] = s.cost; sc[2] = s.merit; sc[3] = s.filter_theta; sc[4] = s.inf_pr; sc[5] = s.inf_comp; sc[6] = s.reg;
      sc[7] = s.alpha_pr; sc[8] = s.alpha_du; sc[9] = s.step_norm; sc[10] = s.inf_du;
      if (rep) rep[b] = R;
    }
  };
  const int nt = std::max(1, std::min(nthreads, batch));
  std::vector<std::thread> th;
  const int per = (batch + nt - 1) / nt;
  for (int t = 0; t < nt; ++t) {
    const int lo = t * per, hi = std::min(batch, (t + 1) * per);
    if (lo < hi) th.emplace_back(work, lo, hi);
  }
  for (auto &t : th) t.join();
}

/* profiling aid (single-threaded use): enable/reset and read the BoxQP histograms */
void oracle_debug_qp_stats(int enable, long long *iters32, long long *trials64, long long *facts16, long long *calls) {
  if (iters32) std::memcpy(iters32, g_qp_stats.iters, sizeof(g_qp_stats.iters));
  if (trials64) std::memcpy(trials64, g_qp_stats.trials, sizeof(g_qp_stats.trials));
  if (facts16) std::memcpy(facts16, g_qp_stats.facts, sizeof(g_qp_stats.facts));
  if (calls) *calls = g_qp_stats.calls;
  g_qp_stats = QpStats();
  g_qp_stats.on = enable != 0;
}

int oracle_hardware_threads(void) {
  unsigned h = std::thread::hardware_concurrency();
  return h ? (int)h : 1;
}

const char *oracle_status_string(int status) {
  switch (status) {
    case ORACLE_OPTIMAL: return "OptimalSolutionFound";
    case ORACLE_ACCEPTABLE: return "AcceptableSolutionFound";
    case ORACLE_MAX_ITERATIONS: return "MaxIterationsReached";
    case ORACLE_REG_LIMIT: return "RegularizationLimitReached_NotConverged";
    case ORACLE_MAX_CPU_TIME: return "MaxCpuTimeReached";
    default: return "Running";
  }
}

} /* extern "C" */
