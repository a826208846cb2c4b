// Linearisation + bookkeeping kernels.
//
//  linearize : for every (instance, t) build the backward-sweep record
//                A = I + dt*Fx, B = dt*Fu                       (clddp_solver.cpp:113-118)
//                lx = 2 Q_ (x - ref_t), lu = 2 R_ u             (objective.cpp:100-120; Q_=Q*dt, R_=R*dt)
//              and the terminal gradient V_x(N) = 2 Qf (x_N - ref) (objective.cpp:122-126).
//  initialize: initializeProblemIfNecessary + CLDDPSolver::initialize cold start
//              (cddp_core.cpp:272-306, clddp_solver.cpp:68-74, cddp_solver_base.cpp:416-424).
#include "engine.h"
#include "kernels_linearize.cuh"
#include "user_model_host.h"

namespace cddp_b200 {

namespace {

using kern::linearize_kernel;
using kern::ref_ptr;

// LTI: runtime dimensions; Fx = (A_d - I)/dt, Fu = B_d/dt (lti_system.cpp:78-92) then the solver's
// A = I + dt*Fx, B = dt*Fu — reproduced literally so that roundoff matches the reference's path.
__global__ void __launch_bounds__(128) linearize_lti_kernel(Constants c, DeviceState d, int force) {
  const long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const int per = d.N + 1, n = d.n, m = d.m;
  if (gid >= (long long)d.B * per) return;
  const int b = (int)(gid / per), t = (int)(gid % per);
  if (!force && (d.status[b] != CDDP_B200_STATUS_RUNNING || d.lin_valid[b])) return;
  const int cur = d.cur[b];
  const double *x = d.X[cur] + ((size_t)b * (d.N + 1) + t) * n;
  if (t == d.N) {
    const double *ref = d.xref + (size_t)b * n;
    for (int i = 0; i < n; ++i) {
      double s = 0.0;
      for (int j = 0; j < n; ++j) s += c.Qf2[i * n + j] * (x[j] - ref[j]);
      d.vterm[(size_t)b * n + i] = s;
    }
    return;
  }
  const double *u = d.U[cur] + ((size_t)b * d.N + t) * m;
  double *r = d.rec + ((size_t)b * d.N + t) * d.rec_stride;
  for (int i = 0; i < n; ++i)
    for (int j = 0; j < n; ++j) {
      const double fx = (c.mp.lti_A[i * n + j] - (i == j ? 1.0 : 0.0)) / c.dt;
      r[i * n + j] = c.dt * fx + (i == j ? 1.0 : 0.0);
    }
  r += n * n;
  for (int i = 0; i < n * m; ++i) r[i] = c.dt * (c.mp.lti_B[i] / c.dt);
  r += n * m;
  const double *ref = ref_ptr(d, b, t);
  for (int i = 0; i < n; ++i) {
    double s = 0.0;
    for (int j = 0; j < n; ++j) s += c.Qdt2[i * n + j] * (x[j] - ref[j]);
    r[i] = s;
  }
  r += n;
  for (int i = 0; i < m; ++i) {
    double s = 0.0;
    for (int j = 0; j < m; ++j) s += c.Rdt2[i * m + j] * u[j];
    r[i] = s;
  }
  r += m;
  for (int i = 0; i < m; ++i) r[i] = u[i];
}

// One warp per instance: X[0] = x0, zero gains, initial cost of the GIVEN trajectory (no re-rollout,
// clddp_solver.cpp:72-74), reset per-instance solver state.
__global__ void __launch_bounds__(128) initialize_kernel(Constants c, DeviceState d) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (warp >= d.B) return;
  const int b = warp, n = d.n, m = d.m, N = d.N;
  const int cur = d.cur[b];
  double *X = d.X[cur] + (size_t)b * (N + 1) * n;
  const double *U = d.U[cur] + (size_t)b * N * m;
  for (int i = lane; i < n; i += 32) X[i] = d.x0[(size_t)b * n + i];  // cddp_core.cpp:294
  __syncwarp();
  for (size_t i = lane; i < (size_t)N * m * n; i += 32) d.K[(size_t)b * N * m * n + i] = 0.0;
  for (size_t i = lane; i < (size_t)N * m; i += 32) d.kff[(size_t)b * N * m + i] = 0.0;
  // running cost: (e^T Q_) e + (u^T R_) u with Q_ = Q*dt (Qdt2 = 2*Q_), objective.cpp:80-92
  double J = 0.0;
  for (int t = lane; t <= N; t += 32) {
    const double *x = X + (size_t)t * n;
    if (t == N) {
      const double *ref = d.xref + (size_t)b * n;
      double s = 0.0;
      for (int j = 0; j < n; ++j) {
        double r = 0.0;
        for (int i = 0; i < n; ++i) r += (x[i] - ref[i]) * (0.5 * c.Qf2[i * n + j]);
        s += r * (x[j] - ref[j]);
      }
      J += s;
    } else {
      const double *ref = ref_ptr(d, b, t);
      const double *u = U + (size_t)t * m;
      double sx = 0.0, su = 0.0;
      for (int j = 0; j < n; ++j) {
        double r = 0.0;
        for (int i = 0; i < n; ++i) r += (x[i] - ref[i]) * (0.5 * c.Qdt2[i * n + j]);
        sx += r * (x[j] - ref[j]);
      }
      for (int j = 0; j < m; ++j) {
        double r = 0.0;
        for (int i = 0; i < m; ++i) r += u[i] * (0.5 * c.Rdt2[i * m + j]);
        su += r * u[j];
      }
      J += sx + su;
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) J += __shfl_xor_sync(0xffffffffu, J, o);
  if (lane == 0) {
    d.cost[b] = J;
    d.reg[b] = c.opt.reg_initial_value;
    d.alpha[b] = c.opt.ls_initial_step_size;
    d.inf_du[b] = __longlong_as_double(0x7ff0000000000000LL);
    d.status[b] = CDDP_B200_STATUS_RUNNING;
    d.iter[b] = 0;
    d.lin_valid[b] = 0;
    d.bw_ok[b] = 0;
    d.accepted[b] = -1;
    d.dV[2 * b] = 0.0;
    d.dV[2 * b + 1] = 0.0;
    if (d.history) {
      double *h = d.history + (size_t)b * d.history_cap * 4;
      h[0] = J;
      h[1] = c.opt.ls_initial_step_size;
      h[2] = __longlong_as_double(0x7ff0000000000000LL);
      h[3] = c.opt.reg_initial_value;
      d.history_len[b] = 1;
    }
  }
}

__global__ void finalize_kernel(DeviceState d, int final_status) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= d.B) return;
  if (d.status[b] == CDDP_B200_STATUS_RUNNING) d.status[b] = final_status;
}

__global__ void count_running_kernel(DeviceState d) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  const bool run = b < d.B && d.status[b] == CDDP_B200_STATUS_RUNNING;
  const unsigned mask = __ballot_sync(0xffffffffu, run);
  if ((threadIdx.x & 31) == 0 && mask) atomicAdd(d.num_running, __popc(mask));
}

// Stable compaction of the running instances into `order` (one CTA: every thread owns a contiguous chunk of the batch,
// block-wide exclusive scan of the chunk counts); also writes the count to d.num_running.
__global__ void __launch_bounds__(1024) compact_running_kernel(DeviceState d, int *order) {
  __shared__ int part[1024];
  const int chunk = (d.B + 1023) / 1024;
  const int lo = min(threadIdx.x * chunk, d.B), hi = min(lo + chunk, d.B);
  int cnt = 0;
  for (int b = lo; b < hi; ++b) cnt += d.status[b] == CDDP_B200_STATUS_RUNNING;
  part[threadIdx.x] = cnt;
  __syncthreads();
  for (int o = 1; o < 1024; o <<= 1) {  // Hillis-Steele inclusive scan
    const int v = threadIdx.x >= o ? part[threadIdx.x - o] : 0;
    __syncthreads();
    part[threadIdx.x] += v;
    __syncthreads();
  }
  int pos = part[threadIdx.x] - cnt;
  for (int b = lo; b < hi; ++b)
    if (d.status[b] == CDDP_B200_STATUS_RUNNING) order[pos++] = b;
  if (threadIdx.x == 1023) *d.num_running = part[1023];
}

__global__ void unpack_lin_kernel(DeviceState d, double *A, double *Bm) {
  const long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (gid >= (long long)d.B * d.N) return;
  const int n = d.n, m = d.m;
  const double *r = d.rec + (size_t)gid * d.rec_stride;
  for (int i = 0; i < n * n; ++i) A[(size_t)gid * n * n + i] = d.idxA[i] >= 0 ? r[d.idxA[i]] : 0.0;
  for (int i = 0; i < n * m; ++i) Bm[(size_t)gid * n * m + i] = d.idxB[i] >= 0 ? r[d.idxB[i]] : 0.0;
}

// dense layout only (capi switches the solver to RECORDS_DENSE before packing caller-supplied Jacobians)
__global__ void pack_lin_kernel(DeviceState d, const double *A, const double *Bm) {
  const long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (gid >= (long long)d.B * d.N) return;
  const int n = d.n, m = d.m;
  double *r = d.rec + (size_t)gid * d.rec_stride;
  for (int i = 0; i < n * n; ++i)
    if (d.idxA[i] >= 0) r[d.idxA[i]] = A[(size_t)gid * n * n + i];
  for (int i = 0; i < n * m; ++i)
    if (d.idxB[i] >= 0) r[d.idxB[i]] = Bm[(size_t)gid * n * m + i];
}

// copy the nominal (which=0) or candidate (which=1) trajectory of every instance to contiguous output
__global__ void gather_current_kernel(DeviceState d, double *X, double *U, int which) {
  const int b = blockIdx.x;
  const int buf = d.cur[b] ^ which;
  const size_t nx = (size_t)(d.N + 1) * d.n, nu = (size_t)d.N * d.m;
  if (X)
    for (size_t i = threadIdx.x; i < nx; i += blockDim.x) X[b * nx + i] = d.X[buf][b * nx + i];
  if (U)
    for (size_t i = threadIdx.x; i < nu; i += blockDim.x) U[b * nu + i] = d.U[buf][b * nu + i];
}

// gather a double-buffered per-instance array (IPDDP duals / slacks / constraint values) of the CURRENT nominal
__global__ void gather_by_cur_kernel(DeviceState d, const double *buf0, const double *buf1, size_t per_instance, double *out) {
  const int b = blockIdx.x;
  const double *src = (d.cur[b] ? buf1 : buf0) + (size_t)b * per_instance;
  for (size_t i = threadIdx.x; i < per_instance; i += blockDim.x) out[(size_t)b * per_instance + i] = src[i];
}

// Receding-horizon shift of the nominal trajectory (MPC warm start): X[t] <- X[t+k], U[t] <- U[t+k]; the tail repeats the
// last control and the last state.  Written into the candidate buffer, then the buffers are swapped.
__global__ void shift_kernel(DeviceState d, int k) {
  const int b = blockIdx.x;
  const int n = d.n, m = d.m, N = d.N;
  const int cur = d.cur[b];
  const double *Xs = d.X[cur] + (size_t)b * (N + 1) * n;
  const double *Us = d.U[cur] + (size_t)b * N * m;
  double *Xd = d.X[cur ^ 1] + (size_t)b * (N + 1) * n;
  double *Ud = d.U[cur ^ 1] + (size_t)b * N * m;
  for (int i = threadIdx.x; i < (N + 1) * n; i += blockDim.x) {
    const int t = i / n, j = i - t * n;
    Xd[i] = Xs[(size_t)min(t + k, N) * n + j];
  }
  for (int i = threadIdx.x; i < N * m; i += blockDim.x) {
    const int t = i / m, j = i - t * m;
    Ud[i] = Us[(size_t)min(t + k, N - 1) * m + j];
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    d.cur[b] = cur ^ 1;
    d.lin_valid[b] = 0;
  }
}

__global__ void first_controls_kernel(DeviceState d, double *u0) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= d.B * d.m) return;
  const int b = i / d.m, j = i - b * d.m;
  u0[i] = d.U[d.cur[b]][(size_t)b * d.N * d.m + j];
}

}  // namespace

cudaError_t launch_gather_by_cur(const DeviceState &d, const double *buf0, const double *buf1, size_t per_instance, double *out,
                                 cudaStream_t st) {
  gather_by_cur_kernel<<<d.B, 128, 0, st>>>(d, buf0, buf1, per_instance, out);
  return cudaGetLastError();
}

cudaError_t launch_shift(const DeviceState &d, int k, cudaStream_t st) {
  shift_kernel<<<d.B, 128, 0, st>>>(d, k);
  return cudaGetLastError();
}

cudaError_t launch_first_controls(const DeviceState &d, double *u0, cudaStream_t st) {
  first_controls_kernel<<<(d.B * d.m + 127) / 128, 128, 0, st>>>(d, u0);
  return cudaGetLastError();
}

template <int MODEL, class PAT>
cudaError_t launch_linearize_model(const Constants &c, const DeviceState &d, bool force, cudaStream_t stream) {
  constexpr int NS = Model<MODEL>::NS, NC = Model<MODEL>::NC;
  using L = RecordLayout<NS, NC, PAT>;
  constexpr int PSH = lin_part_width(L::stride) | 1;
  constexpr size_t per_warp = sizeof(double) * 32 * PSH;
  constexpr int wpc = 4;
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(linearize_kernel<MODEL, PAT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(per_warp * wpc));
    if (e != cudaSuccess) return e;
    configured = true;
  }
  const long long warps = (long long)d.n_slots * ((d.N + 1 + 31) / 32);
  const int blocks = (int)((warps + wpc - 1) / wpc);
  linearize_kernel<MODEL, PAT><<<blocks, 128, per_warp * wpc, stream>>>(c, d, force ? 1 : 0, wpc);
  return cudaGetLastError();
}

cudaError_t launch_linearize(const Constants &c, const DeviceState &d, bool force, cudaStream_t stream) {
  const bool st = d.layout == RECORDS_STRUCTURED;
  switch (c.model) {
    case CDDP_B200_MODEL_PENDULUM: return launch_linearize_model<CDDP_B200_MODEL_PENDULUM, DensePattern>(c, d, force, stream);
    case CDDP_B200_MODEL_CARTPOLE:
      if (st) return launch_linearize_model<CDDP_B200_MODEL_CARTPOLE, ModelPattern<CDDP_B200_MODEL_CARTPOLE>>(c, d, force, stream);
      return launch_linearize_model<CDDP_B200_MODEL_CARTPOLE, DensePattern>(c, d, force, stream);
    case CDDP_B200_MODEL_UNICYCLE:
      if (st) return launch_linearize_model<CDDP_B200_MODEL_UNICYCLE, ModelPattern<CDDP_B200_MODEL_UNICYCLE>>(c, d, force, stream);
      return launch_linearize_model<CDDP_B200_MODEL_UNICYCLE, DensePattern>(c, d, force, stream);
    case CDDP_B200_MODEL_QUADROTOR:
      if (st) return launch_linearize_model<CDDP_B200_MODEL_QUADROTOR, ModelPattern<CDDP_B200_MODEL_QUADROTOR>>(c, d, force, stream);
      return launch_linearize_model<CDDP_B200_MODEL_QUADROTOR, DensePattern>(c, d, force, stream);
    case CDDP_B200_MODEL_LTI: {
      const long long total = (long long)d.B * (d.N + 1);
      linearize_lti_kernel<<<(int)((total + 127) / 128), 128, 0, stream>>>(c, d, force);
      return cudaGetLastError();
    }
    case CDDP_B200_MODEL_USER: return launch_user_linearize(c, d, force, stream);
    default: return cudaErrorInvalidValue;
  }
}

cudaError_t launch_initialize(const Constants &c, const DeviceState &d, cudaStream_t st) {
  const int threads = 128;
  const int blocks = (d.B * 32 + threads - 1) / threads;
  initialize_kernel<<<blocks, threads, 0, st>>>(c, d);
  return cudaGetLastError();
}

cudaError_t launch_finalize(const Constants &, const DeviceState &d, int final_status, cudaStream_t st) {
  finalize_kernel<<<(d.B + 127) / 128, 128, 0, st>>>(d, final_status);
  return cudaGetLastError();
}

cudaError_t launch_count_running(const DeviceState &d, cudaStream_t st) {
  cudaError_t e = cudaMemsetAsync(d.num_running, 0, sizeof(int), st);
  if (e != cudaSuccess) return e;
  count_running_kernel<<<(d.B + 127) / 128, 128, 0, st>>>(d);
  return cudaGetLastError();
}

cudaError_t launch_compact_running(const DeviceState &d, int *order, cudaStream_t st) {
  compact_running_kernel<<<1, 1024, 0, st>>>(d, order);
  return cudaGetLastError();
}

cudaError_t launch_unpack_linearization(const Constants &, const DeviceState &d, double *A, double *Bm,
                                        cudaStream_t st) {
  const long long total = (long long)d.B * d.N;
  unpack_lin_kernel<<<(int)((total + 127) / 128), 128, 0, st>>>(d, A, Bm);
  return cudaGetLastError();
}

cudaError_t launch_pack_linearization(const Constants &, const DeviceState &d, const double *A, const double *Bm,
                                      cudaStream_t st) {
  const long long total = (long long)d.B * d.N;
  pack_lin_kernel<<<(int)((total + 127) / 128), 128, 0, st>>>(d, A, Bm);
  return cudaGetLastError();
}

cudaError_t launch_gather_current(const Constants &, const DeviceState &d, double *X, double *U, int which,
                                  cudaStream_t st) {
  gather_current_kernel<<<d.B, 128, 0, st>>>(d, X, U, which);
  return cudaGetLastError();
}

}  // namespace cddp_b200
