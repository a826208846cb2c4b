// Eigen::LDLT restated for tiny matrices held in shared memory (one lane executes it).
#pragma once
#include "engine.h"

namespace cddp_b200 {
namespace kern {

// Eigen 3.4.0 LDLT (lower, diagonal pivoting), ldlt_inplace<Lower>::unblocked — executed by ONE lane on the m x m
// matrix in shared memory.  a: in = matrix, out = L strictly below the diagonal and D on it; tr: transpositions.
// Returns info() == Success.
// NN: compile-time size (0 = the run-time n_rt): with a constant size the loops unroll and the addressing is immediate.
template <int NN>
__device__ __forceinline__ bool ldlt_small_t(double *a, int *tr, int n_rt) {
  const int n = NN ? NN : n_rt;
  bool ok = true, found_zero_pivot = false;
  for (int k = 0; k < n; ++k) {
    int big = k;
    double best = fabs(a[k * n + k]);
    for (int i = k + 1; i < n; ++i)
      if (fabs(a[i * n + i]) > best) {
        best = fabs(a[i * n + i]);
        big = i;
      }
    tr[k] = big;
    if (big != k) {
      const int s = n - big - 1;
      for (int j = 0; j < k; ++j) { const double t = a[k * n + j]; a[k * n + j] = a[big * n + j]; a[big * n + j] = t; }
      for (int i = 0; i < s; ++i) {
        const double t = a[(big + 1 + i) * n + k];
        a[(big + 1 + i) * n + k] = a[(big + 1 + i) * n + big];
        a[(big + 1 + i) * n + big] = t;
      }
      { const double t = a[k * n + k]; a[k * n + k] = a[big * n + big]; a[big * n + big] = t; }
      for (int i = k + 1; i < big; ++i) { const double t = a[i * n + k]; a[i * n + k] = a[big * n + i]; a[big * n + i] = t; }
    }
    const int rs = n - k - 1;
    if (k > 0) {
      double temp[NN ? NN : CDDP_B200_MAX_N];  // capacity: also used for the n x n terminal-multiplier system
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
    const bool valid = fabs(akk) > 0.0;
    if (k == 0 && !valid) {
      for (int j = 0; j < n; ++j) tr[j] = j;
      return ok;
    }
    if (rs > 0 && valid) {
      for (int i = 0; i < rs; ++i) a[(k + 1 + i) * n + k] /= akk;
    } else if (rs > 0) {
      for (int i = 0; i < rs; ++i)
        if (a[(k + 1 + i) * n + k] != 0.0) ok = false;
    }
    if (found_zero_pivot && valid) ok = false;
    else if (!valid) found_zero_pivot = true;
  }
  return ok;
}

__device__ inline bool ldlt_small(double *a, int *tr, int n) { return ldlt_small_t<0>(a, tr, n); }

// LDLT::solve for one right-hand side held in b[0..n) with stride `st`
template <int NN>
__device__ __forceinline__ void ldlt_solve_t(const double *a, const int *tr, int n_rt, double *b, int st) {
  const int n = NN ? NN : n_rt;
  for (int k = 0; k < n; ++k)
    if (tr[k] != k) { const double t = b[k * st]; b[k * st] = b[tr[k] * st]; b[tr[k] * st] = t; }
  for (int i = 0; i < n; ++i) {  // running value in a register: the same subtractions in the same order, without a
    double s = b[i * st];        // shared-memory store -> load round trip between them
    for (int j = 0; j < i; ++j) s -= a[i * n + j] * b[j * st];
    b[i * st] = s;
  }
  const double tol = 2.2250738585072014e-308;  // numeric_limits<double>::min()
  for (int i = 0; i < n; ++i) {
    if (fabs(a[i * n + i]) > tol) b[i * st] /= a[i * n + i];
    else b[i * st] = 0.0;
  }
  for (int i = n - 1; i >= 0; --i) {
    double s = b[i * st];
    for (int j = i + 1; j < n; ++j) s -= a[j * n + i] * b[j * st];
    b[i * st] = s;
  }
  for (int k = n - 1; k >= 0; --k)
    if (tr[k] != k) { const double t = b[k * st]; b[k * st] = b[tr[k] * st]; b[tr[k] * st] = t; }
}
__device__ inline void ldlt_solve(const double *a, const int *tr, int n, double *b, int st) { ldlt_solve_t<0>(a, tr, n, b, st); }

}  // namespace kern
}  // namespace cddp_b200
