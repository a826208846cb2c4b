// Internal (non-ABI) definitions shared by the kernels and the C-ABI layer.
#pragma once
#ifndef __CUDACC_RTC__
#include <cuda_runtime.h>
#endif

#include "../../include/cddp_b200.h"
#include "models.cuh"
#include "records.cuh"

namespace cddp_b200 {

// ---- HBM layout (see DESIGN.md "Data layout") ----
// All per-instance arrays are batch-outermost and contiguous per instance so that one warp streams
// one trajectory.
//
//  X[2]     : [B][N+1][n]   double-buffered nominal / candidate state trajectory (cur[b] selects)
//  U[2]     : [B][N][m]
//  rec      : [B][N][rec_stride]  per-timestep linearisation record consumed by the backward sweep
//               (records.cuh): A entries | B rows | lx (n) | lu (m) | u (m) | pad to an even count
//               (16-byte multiple so that one cp.async.bulk moves one record).  `layout` says whether
//               A,B are stored densely or only at the model's structural non-zeros.
//  vterm    : [B][n]        terminal gradient V_x(N) = 2 Qf (x_N - ref)
//  K        : [B][N][m][n]  feedback gains K_u_
//  kff      : [B][N][m]     feed-forward k_u_ (updated in place: doubles as BoxQP warm start)
struct DeviceState {
  int B, n, m, N;
  int rec_stride;  // doubles per record
  int layout;      // RECORDS_DENSE / RECORDS_STRUCTURED
  const int *idxA; // device [n*n] record index of A(l,j), -1 = structural zero   (generic pack/unpack kernels)
  const int *idxB; // device [n*m]
  int offLx, offLu, offU;
  int num_alphas;
  double *X[2];
  double *U[2];
  double *rec;
  double *vterm;
  double *K;
  double *kff;
  double *x0;        // [B][n]
  double *xref;      // [B][n]
  double *ref_traj;  // [B][N+1][n] or nullptr
  // per-instance scalars
  int *cur;        // which X/U buffer holds the nominal trajectory
  int *status;     // CDDP_B200_STATUS_*
  int *iter;       // iterations_completed
  int *lin_valid;  // rec matches the nominal trajectory
  int *bw_ok;      // last backward sweep succeeded
  int *accepted;   // index of the accepted alpha in the last line search (-1 none)
  int *fw_done;    // FW_ITERATE only: this forward call's line search was already settled by forward_first_kernel
  // Work list of the per-iteration kernels: slot s of a launch works on instance order[s] (identity when order is null);
  // n_slots <= B slots are launched.  solve() compacts the still-running instances into the list whenever it polls the
  // running counter, so a batch whose instances converge at different iterations stops paying for the finished ones.
  const int *order;
  int n_slots;
  double *reg;     // regularization_
  double *cost;    // cost_
  double *alpha;   // alpha_pr_
  double *inf_du;  // inf_du_
  double *dV;      // [B][2]
  double *ls_cost; // [B][CDDP_B200_MAX_ALPHAS] cost of every line-search candidate
  double *Vx0;     // [B][n]   value gradient at t=0 after the last sweep (white-box)
  double *Vxx0;    // [B][n][n]
  double *ckpt;    // [B][LG][LG][n] line-search scratch: state of every alpha at the start of every segment
  double *history; // [B][max_iterations+1][4] or nullptr
  int *history_len;
  int history_cap;
  // decision trace (test / audit instrumentation, cddp_b200_enable_trace): one int per entry of the main loop,
  // (backward failures << 8) | code, code = 1 + index of the accepted alpha, 0 = line search failed,
  // 0xff = early convergence exit, 0xfe = regularisation limit in the backward retry
  int *trace;  // [B][trace_cap] or nullptr
  int trace_cap;
  int *num_running;  // device counter
  void *user;        // HOST pointer (never dereferenced on the device): NVRTC-compiled kernels of a CDDP_B200_MODEL_USER handle
};

// linearize_kernel stages a record of `rs` doubles through shared memory in lin_parts(rs) slices of lin_part_width(rs)
// doubles (kernels_linearize.cuh); shared by the kernel and by the host launchers that size its shared memory.
#ifdef __CUDACC__
#define CDDP_B200_HD __host__ __device__
#else
#define CDDP_B200_HD
#endif
CDDP_B200_HD constexpr int lin_parts(int rs) { return (rs + 55) / 56; }
CDDP_B200_HD constexpr int lin_part_width(int rs) { return (rs + lin_parts(rs) - 1) / lin_parts(rs); }
// shared-memory geometry of ip_forward_kernel (kernels_ipddp.cuh): the constraint table staged once per CTA, and one
// timestep's operands x | u | k | K | S | Y | k_s | k_y | K_s | K_y of a trajectory
CDDP_B200_HD constexpr int con_table_doubles(int n, int m, int D) { return (D * n + D * m + 3 * D + 1) & ~1; }
CDDP_B200_HD constexpr int ip_fw_step_doubles(int n, int m, int D) { return (n + 2 * m + m * n + 4 * D + 2 * D * n + 1) & ~1; }
// ... preceded by the halved cost matrices Q dt | R dt | Qf (the rollout's running / terminal cost, read every timestep)
CDDP_B200_HD constexpr int ip_fw_cost_doubles(int n, int m) { return (2 * n * n + m * m + 1) & ~1; }

#ifdef __CUDACC__
// std::min(std::max(v, lo), hi) as the reference writes it (boxqp.cpp:241-250; Eigen's cwiseMax / cwiseMin in
// ControlConstraint::clamp, constraint.hpp:225-228): (v < lo) ? lo : v, then (hi < v) ? hi : v.  Two compares and two
// 64-bit selects; CUDA's fmax / fmin add NaN canonicalisation and come to 15 instructions per clamp, which made the
// projection the most expensive part of a BoxQP line-search trial on the sweep's critical path.
__device__ __forceinline__ double clamp_box(double v, double lo, double hi) {
  v = (v < lo) ? lo : v;
  return (hi < v) ? hi : v;
}
// std::max / std::min as compare-select (same reason): (a < b) ? b : a and (b < a) ? b : a
__device__ __forceinline__ double max_ref(double a, double b) { return (a < b) ? b : a; }
__device__ __forceinline__ double min_ref(double a, double b) { return (b < a) ? b : a; }
// decision trace: the sweep stores (failures << 8) | code for iteration `iter` (1-based), the line search adds its code
__device__ __forceinline__ void trace_backward(const DeviceState &d, int b, int iter, int failures, int code) {
  if (d.trace && iter >= 1 && iter <= d.trace_cap) d.trace[(size_t)b * d.trace_cap + iter - 1] = (failures << 8) | code;
}
__device__ __forceinline__ void trace_line_search(const DeviceState &d, int b, int accepted) {
  const int iter = d.iter[b];
  if (d.trace && accepted >= 0 && iter >= 1 && iter <= d.trace_cap) d.trace[(size_t)b * d.trace_cap + iter - 1] |= 1 + accepted;
}
// instance handled by work-list slot `slot` (d.B = none)
__device__ __forceinline__ int slot_instance(const DeviceState &d, int slot) {
  return slot < d.n_slots ? (d.order ? d.order[slot] : slot) : d.B;
}
#endif

// batch-shared constants, passed by value to kernels (fits the 4 KB parameter space comfortably)
struct Constants {
  int model, n, m, N, integrator, has_box;
  int q_diag;     // Q is diagonal (l_xx touches only the diagonal of Q_xx)
  int cost_diag;  // Q, R and Qf are all diagonal
  double dt;
  ModelParams mp;
  const double *Qdt2;  // device [n][n] = 2*Q*dt  (l_xx, objective.cpp:130-134)
  const double *Rdt2;  // device [m][m] = 2*R*dt  (l_uu)
  const double *Qf2;   // device [n][n] = 2*Qf    (phi_xx)
  double lb[CDDP_B200_MAX_M], ub[CDDP_B200_MAX_M];
  double alphas[CDDP_B200_MAX_ALPHAS];
  int num_alphas;
  int ls_window;  // windowed line search (kernels_forward.cuh): first 8 candidates at 8 lanes per trajectory, then the rest
  int fuse_lin;   // CLDDP, quadrotor structured records: the sweep's QP warp forms the linearisation records itself (opt-in)
  int speculate;  // user models: speculative alphas_[0] pass; -1 = automatic (only while the full search is throughput-bound)
  cddp_b200_options opt;
};

enum RecordLayoutKind { RECORDS_DENSE = 0, RECORDS_STRUCTURED = 1 };

// ---- IPDDP (ipddp.cu) ----
// The path-constraint set flattened to rows: g_r(x,u) = a_r . x - off_r (STATE rows: StateConstraint, LinearConstraint),
// a_r . u - off_r (CONTROL rows: ControlConstraint) or -scale |x[0:bdim] - centre|^2 + scale radius^2 (BALL rows).
enum { IP_ROW_STATE = 0, IP_ROW_CONTROL = 1, IP_ROW_BALL = 2 };
constexpr int IP_FILTER_CAP = 8;
constexpr int IP_HISTORY_COLS = 9;  // objective, merit, alpha_pr, alpha_du, inf_du, inf_pr, inf_comp, reg, mu
constexpr int IP_MAX_DUAL = 48;

struct IpConstants {
  int d;   // total dual dimension (rows)
  int nc;  // number of constraints in the set (0 = unconstrained branch)
  int teq; // TerminalEqualityConstraint on the reference state (terminal-equality branch of the backward pass)
  const int *row_type;   // device [d]
  const int *row_bdim;   // device [d]  ball: dimension of the centre
  const double *Gx;      // device [d][n]  STATE rows: dg/dx; BALL rows: centre in the first bdim entries
  const double *Gu;      // device [d][m]  CONTROL rows: dg/du
  const double *off;     // device [d]  affine rows: upper bound; BALL rows: radius
  const double *scale;   // device [d]
  cddp_b200_ipddp_options io;
};

// doubles per stage record of the small terminal-equality kernels: Q_t | R_t | M_t | q_t | r_t | two residual maxima
inline int teq_stage_stride(int n, int m) { return (n * n + m * m + n * m + n + m + 2 + 1) & ~1; }
// (n, m, dual dimension) for which launch_ip_backward_teq takes the register-resident kernels
inline bool teq_small_case(int n, int m, int dd) { return n == 3 && m == 2 && dd == 5; }

struct IpDevice {
  double *Y[2], *S[2], *G[2];  // [B][N][d] duals, slacks, constraint values; double-buffered like X/U (cur[b] selects)
  double *ky, *ks;             // [B][N][d]
  double *Ky, *Ks;             // [B][N][d][n]
  // per-instance scalars
  double *mu, *merit, *logsum, *filter_theta, *inf_pr, *inf_comp, *step_norm, *alpha_du, *apm, *adm;
  double *filter;    // [B][IP_FILTER_CAP][2] (merit, theta)
  int *filter_size;  // [B]
  double *ls_stats;  // [B][CDDP_B200_MAX_ALPHAS][4] success, cost, merit, theta of every line-search candidate
  // terminal equality (ipddp_teq.cu)
  double *lamT, *dlamT;  // [B][n] Lambda_T_eq_, dLambda_T_eq_
  double *lamh;          // [B] Lambda_T_eq_ . h_T of the nominal trajectory (the merit's multiplier term)
  double *kvar;          // [B][n+1][N][m]   feed-forward of the p+1 sequential-LQR variants
  double *pvar;          // [B][n+1][N+1][n] costate of the variants
  double *rvar;          // [B][N][m]        condensed control gradient r_t
  double *stage;         // [B][N][teq_stage_stride(n,m)] condensed stage cost of the register-resident kernels (ipddp_teq_small.cuh)
  int *teq_ran;          // [B] the sweep kernel ran and succeeded for this instance in the current launch sequence
};

enum BackwardMode { BW_SINGLE = 0 /* one sweep, no retry, no iteration bookkeeping */, BW_ITERATE = 1 };
enum ForwardMode { FW_EVALUATE = 0 /* do not apply */, FW_ITERATE = 1 };

#ifndef __CUDACC_RTC__
// kernel launchers (one per translation unit)
cudaError_t launch_linearize(const Constants &c, const DeviceState &d, bool force, cudaStream_t st);
cudaError_t launch_initialize(const Constants &c, const DeviceState &d, cudaStream_t st);
cudaError_t launch_backward(const Constants &c, const DeviceState &d, int mode, cudaStream_t st);
bool backward_fuses_linearization(const Constants &c, const DeviceState &d);
cudaError_t launch_forward(const Constants &c, const DeviceState &d, int mode, cudaStream_t st);
cudaError_t launch_finalize(const Constants &c, const DeviceState &d, int final_status, cudaStream_t st);
cudaError_t launch_count_running(const DeviceState &d, cudaStream_t st);
cudaError_t launch_compact_running(const DeviceState &d, int *order, cudaStream_t st);
cudaError_t launch_shift(const DeviceState &d, int k, cudaStream_t st);
cudaError_t launch_gather_by_cur(const DeviceState &d, const double *buf0, const double *buf1, size_t per_instance, double *out,
                                 cudaStream_t st);
cudaError_t launch_first_controls(const DeviceState &d, double *u0, cudaStream_t st);
cudaError_t launch_unpack_linearization(const Constants &c, const DeviceState &d, double *A, double *Bm,
                                        cudaStream_t st);
cudaError_t launch_pack_linearization(const Constants &c, const DeviceState &d, const double *A, const double *Bm,
                                      cudaStream_t st);
cudaError_t launch_ip_initialize(const Constants &c, const DeviceState &d, const IpConstants &ic, const IpDevice &ip,
                                 cudaStream_t st);
cudaError_t launch_ip_backward(const Constants &c, const DeviceState &d, const IpConstants &ic, const IpDevice &ip, int mode,
                               cudaStream_t st);
cudaError_t launch_ip_forward(const Constants &c, const DeviceState &d, const IpConstants &ic, const IpDevice &ip, int mode,
                              cudaStream_t st);
cudaError_t launch_ip_backward_teq(const Constants &c, const DeviceState &d, const IpConstants &ic, const IpDevice &ip, int mode,
                                   cudaStream_t st);
cudaError_t launch_gather_current(const Constants &c, const DeviceState &d, double *X, double *U, int which,
                                  cudaStream_t st);

#endif  // !__CUDACC_RTC__

}  // namespace cddp_b200
