// Per-timestep linearisation RECORD: the unit the backward sweep streams from HBM.
//
//   record(b, t) = [ A entries | B rows | l_x (n) | l_u (m) | u (m) | pad to an even count ]
//
// DENSE layout (any model, runtime or compile-time dims; what cddp_b200_set_linearization feeds):
//   A = all n*n entries row-major, B = all n*m entries row-major  ->  the "stacked Jacobians in HBM"
//   of the north star, 8*(n^2+nm+n+2m) bytes per record.
// STRUCTURED layout (built-in models only): only the STRUCTURAL non-zeros of A = I + dt*Fx and the
//   structurally non-zero ROWS of B = dt*Fu are stored, in row-major order of the pattern.  The
//   pattern is a compile-time property of the model's continuous dynamics (which state derivatives
//   depend on which states), so the sweep kernel unrolls over it and never multiplies by a known
//   zero.  Quadrotor: 58 of 169 A entries, 6 of 13 B rows -> 832-byte records instead of 2384.
//
// Reference: A = I + dt*Fx, B = dt*Fu (src/cddp_core/clddp_solver.cpp:113-118); the Jacobian
// structure follows src/dynamics_model/{quadrotor,cartpole,unicycle,pendulum}.cpp.
#pragma once
#include "../../include/cddp_b200.h"

namespace cddp_b200 {

struct DensePattern {
  __host__ __device__ static constexpr bool a(int, int) { return true; }
  __host__ __device__ static constexpr bool brow(int) { return true; }
};

template <int MODEL>
struct ModelPattern : DensePattern {};

// Quadrotor, state [p(0-2) q(3-6) v(7-9) w(10-12)] (quadrotor.cpp:33-96):
//   p_dot = v;  q_dot = f(q, w);  v_dot = f(q, u);  w_dot = f(w, u)
template <>
struct ModelPattern<CDDP_B200_MODEL_QUADROTOR> {
  __host__ __device__ static constexpr bool a(int l, int j) {
    if (l == j) return true;
    if (l < 3) return j == l + 7;
    if (l < 7) return (j >= 3 && j < 7) || j >= 10;
    if (l < 10) return j >= 3 && j < 7;
    return j >= 10;
  }
  __host__ __device__ static constexpr bool brow(int l) { return l >= 7; }
};

// CartPole, state (x, theta, x_dot, theta_dot) (cartpole.cpp:38-93)
template <>
struct ModelPattern<CDDP_B200_MODEL_CARTPOLE> {
  __host__ __device__ static constexpr bool a(int l, int j) {
    if (l == j) return true;
    if (l == 0) return j == 2;
    if (l == 1) return j == 3;
    return j == 1 || j == 3;
  }
  __host__ __device__ static constexpr bool brow(int l) { return l >= 2; }
};

// Unicycle (x, y, theta; v, omega) (unicycle.cpp:28-66)
template <>
struct ModelPattern<CDDP_B200_MODEL_UNICYCLE> {
  __host__ __device__ static constexpr bool a(int l, int j) { return l == j || (l < 2 && j == 2); }
  __host__ __device__ static constexpr bool brow(int) { return true; }
};

// Compile-time packed layout of one record.
template <int NS, int NC, class PAT>
struct RecordLayout {
  __host__ __device__ static constexpr int idxA(int l, int j) {  // valid only where PAT::a(l,j)
    int c = 0;
    for (int ll = 0; ll < NS; ++ll)
      for (int jj = 0; jj < NS; ++jj) {
        if (ll == l && jj == j) return c;
        if (PAT::a(ll, jj)) ++c;
      }
    return c;
  }
  static constexpr int nA = idxA(NS, 0);
  __host__ __device__ static constexpr int browrank(int l) {
    int c = 0;
    for (int ll = 0; ll < l; ++ll)
      if (PAT::brow(ll)) ++c;
    return c;
  }
  static constexpr int nBrows = browrank(NS);
  static constexpr int offB = nA;
  __host__ __device__ static constexpr int idxB(int l, int a) { return offB + browrank(l) * NC + a; }
  static constexpr int offLx = offB + nBrows * NC;
  static constexpr int offLu = offLx + NS;
  static constexpr int offU = offLu + NC;
  static constexpr int count = offU + NC;
  static constexpr int stride = (count + 1) & ~1;  // even => 16-byte multiple => one cp.async.bulk per record
};

// Runtime description of a layout (host side + generic kernels): index tables into the record.
struct RecordMap {
  int stride, offLx, offLu, offU;
  int idxA[CDDP_B200_MAX_N * CDDP_B200_MAX_N];  // -1 = structurally zero (diagonal is never -1)
  int idxB[CDDP_B200_MAX_N * CDDP_B200_MAX_M];
};

template <int NS, int NC, class PAT>
inline void fill_record_map(RecordMap &m) {
  using L = RecordLayout<NS, NC, PAT>;
  m.stride = L::stride; m.offLx = L::offLx; m.offLu = L::offLu; m.offU = L::offU;
  for (int l = 0; l < NS; ++l) {
    for (int j = 0; j < NS; ++j) m.idxA[l * NS + j] = PAT::a(l, j) ? L::idxA(l, j) : -1;
    for (int a = 0; a < NC; ++a) m.idxB[l * NC + a] = PAT::brow(l) ? L::idxB(l, a) : -1;
  }
}

inline void fill_record_map_dense(RecordMap &m, int n, int mm) {
  const int c = n * n + n * mm + n + 2 * mm;
  m.stride = (c + 1) & ~1; m.offLx = n * n + n * mm; m.offLu = m.offLx + n; m.offU = m.offLu + mm;
  for (int i = 0; i < n * n; ++i) m.idxA[i] = i;
  for (int i = 0; i < n * mm; ++i) m.idxB[i] = n * n + i;
}

// which models have a structured layout
inline bool model_has_structured_layout(int model) {
  return model == CDDP_B200_MODEL_QUADROTOR || model == CDDP_B200_MODEL_CARTPOLE || model == CDDP_B200_MODEL_UNICYCLE;
}

inline void fill_record_map_for(RecordMap &m, int model, int n, int mm, bool structured) {
  if (structured && model == CDDP_B200_MODEL_QUADROTOR)
    fill_record_map<13, 4, ModelPattern<CDDP_B200_MODEL_QUADROTOR>>(m);
  else if (structured && model == CDDP_B200_MODEL_CARTPOLE)
    fill_record_map<4, 1, ModelPattern<CDDP_B200_MODEL_CARTPOLE>>(m);
  else if (structured && model == CDDP_B200_MODEL_UNICYCLE)
    fill_record_map<3, 2, ModelPattern<CDDP_B200_MODEL_UNICYCLE>>(m);
  else
    fill_record_map_dense(m, n, mm);
}

}  // namespace cddp_b200
