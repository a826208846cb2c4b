// C ABI of the B200 batched CLDDP engine (include/cddp_b200.h): handle, HBM buffers, the host
// side of the per-iteration launch sequence.  No solver arithmetic happens on the host: the host
// only launches fixed-shape kernels and polls a running-instance counter
// (CDDPSolverBase::solve outer loop, src/cddp_core/cddp_solver_base.cpp:74-154, lives on the
// device as a per-instance state machine — see backward.cu / forward.cu).
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <new>
#include <string>
#include <vector>

#include "engine.h"
#include "user_model_host.h"

using namespace cddp_b200;

namespace {

thread_local std::string g_last_cuda_error;
thread_local std::string g_last_compile_log;

int cuda_fail(cudaError_t e, const char *what) {
  g_last_cuda_error = std::string(what) + ": " + cudaGetErrorString(e);
  return (e == cudaErrorMemoryAllocation) ? CDDP_B200_ERR_OUT_OF_MEMORY : CDDP_B200_ERR_CUDA;
}

#define CU(call)                                   \
  do {                                             \
    cudaError_t e__ = (call);                      \
    if (e__ != cudaSuccess) return cuda_fail(e__, #call); \
  } while (0)

int build_alphas(const cddp_b200_options &o, double *alphas, int cap) {
  // detail::buildLineSearchAlphas, src/cddp_core/cddp_context_utils.cpp:37-57
  int cnt = 0;
  double a = o.ls_initial_step_size;
  for (int i = 0; i < o.ls_max_iterations && cnt < cap; ++i) {
    alphas[cnt++] = a;
    a *= o.ls_step_reduction_factor;
    if (a < o.ls_min_step_size && i < o.ls_max_iterations - 1) {
      if (cnt < cap) alphas[cnt++] = o.ls_min_step_size;
      break;
    }
  }
  if (cnt == 0) alphas[cnt++] = o.ls_initial_step_size;
  return cnt;
}

int count_alphas(const cddp_b200_options &o) {
  double tmp[256];
  return build_alphas(o, tmp, 256);
}

}  // namespace

struct cddp_b200_solver {
  int device = 0;
  cudaStream_t stream = nullptr;
  Constants c{};
  DeviceState d{};
  std::vector<void *> allocs;
  double *dQdt2 = nullptr, *dRdt2 = nullptr, *dQf2 = nullptr, *dltiA = nullptr, *dltiB = nullptr;
  int *didxA = nullptr, *didxB = nullptr;
  int *order_buf = nullptr;  // work list of the running instances (solve(), see DeviceState::order)
  UserKernels *user_kernels = nullptr;  // CDDP_B200_MODEL_USER: the NVRTC-compiled module
  int kind = 0;      // 0 = CLDDP, 1 = IPDDP
  IpConstants ic{};  // IPDDP: flattened constraint rows + options
  IpDevice ip{};     // IPDDP: duals, slacks, gains, per-instance barrier/filter state
  int ckpt_lg = 16;  // lanes per trajectory the line-search scratch was sized for
  int poll_interval = -1;  // -1: automatic (default: every iteration for heavy batches, else a widening stride); 0: never poll (fully asynchronous solve); k > 0: every k iterations
  int speculate = -1;     // cddp_b200_set_first_alpha_speculation
  int fuse_lin = 0;       // cddp_b200_set_fused_linearization
  int ls_window = 0;      // windowed line search (cddp_b200_set_line_search_window); off: measured slower on B200
  double *rec_by_layout[2] = {nullptr, nullptr};  // record buffers are allocated lazily per layout
  int *h_running = nullptr;  // pinned
  bool initialized = false;
  bool have_instances = false;
  bool timing_enabled = false;
  cddp_b200_timing timing{};
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;
  double *history_buf = nullptr;  // the allocation behind d.history (which is null while history is disabled)
  int *trace_buf = nullptr;       // same for d.trace
  // scratch for host-pointer setters / getters
  double *scratch = nullptr;
  size_t scratch_bytes = 0;

  template <typename T>
  int alloc(T **p, size_t count) {
    void *q = nullptr;
    cudaError_t e = cudaMalloc(&q, count * sizeof(T) ? count * sizeof(T) : sizeof(T));
    if (e != cudaSuccess) return cuda_fail(e, "cudaMalloc");
    allocs.push_back(q);
    *p = static_cast<T *>(q);
    return 0;
  }
  int ensure_scratch(size_t bytes) {
    if (bytes <= scratch_bytes) return 0;
    if (scratch) cudaFree(scratch);
    scratch = nullptr;
    scratch_bytes = 0;
    cudaError_t e = cudaMalloc((void **)&scratch, bytes);
    if (e != cudaSuccess) return cuda_fail(e, "cudaMalloc(scratch)");
    scratch_bytes = bytes;
    return 0;
  }
};

namespace {

struct DeviceGuard {
  int prev = -1;
  bool ok = true;
  explicit DeviceGuard(int dev) {
    if (cudaGetDevice(&prev) != cudaSuccess) { ok = false; return; }
    if (prev != dev && cudaSetDevice(dev) != cudaSuccess) ok = false;
  }
  ~DeviceGuard() {
    int now = -1;
    if (prev >= 0 && cudaGetDevice(&now) == cudaSuccess && now != prev) cudaSetDevice(prev);
  }
};

// option checks shared by create and cddp_b200_set_options; kind 1 = IPDDP (one lane per alpha, 16 lanes per trajectory)
int validate_options(const cddp_b200_options &o, int kind) {
  if (o.max_iterations < 0 || o.ls_max_iterations < 0) return CDDP_B200_ERR_INVALID_ARGUMENT;
  if (count_alphas(o) > (kind == 1 ? 16 : CDDP_B200_MAX_ALPHAS)) return CDDP_B200_ERR_INVALID_ARGUMENT;
  return 0;
}

int validate(const cddp_b200_problem *p, const cddp_b200_options *o, int batch, bool has_user_source = false) {
  if (!p || !o) return CDDP_B200_ERR_INVALID_ARGUMENT;
  if (batch < 1 || p->horizon < 1 || !(p->dt > 0.0)) return CDDP_B200_ERR_INVALID_ARGUMENT;
  if (p->n < 1 || p->n > CDDP_B200_MAX_N || p->m < 1 || p->m > CDDP_B200_MAX_M) return CDDP_B200_ERR_INVALID_ARGUMENT;
  if (!p->Q || !p->R || !p->Qf) return CDDP_B200_ERR_INVALID_ARGUMENT;
  if (p->has_control_box && (!p->lb || !p->ub)) return CDDP_B200_ERR_INVALID_ARGUMENT;
  if (p->integrator < CDDP_B200_EULER || p->integrator > CDDP_B200_RK4) return CDDP_B200_ERR_INVALID_ARGUMENT;
  switch (p->model) {
    case CDDP_B200_MODEL_PENDULUM: if (p->n != 2 || p->m != 1) return CDDP_B200_ERR_INVALID_ARGUMENT; break;
    case CDDP_B200_MODEL_CARTPOLE: if (p->n != 4 || p->m != 1) return CDDP_B200_ERR_INVALID_ARGUMENT; break;
    case CDDP_B200_MODEL_UNICYCLE: if (p->n != 3 || p->m != 2) return CDDP_B200_ERR_INVALID_ARGUMENT; break;
    case CDDP_B200_MODEL_QUADROTOR: if (p->n != 13 || p->m != 4) return CDDP_B200_ERR_INVALID_ARGUMENT; break;
    case CDDP_B200_MODEL_LTI: if (!p->lti_A || !p->lti_B) return CDDP_B200_ERR_INVALID_ARGUMENT; break;
    case CDDP_B200_MODEL_USER: if (!has_user_source) return CDDP_B200_ERR_UNSUPPORTED_MODEL; break;  // no source, no device dynamics
    default: return CDDP_B200_ERR_UNSUPPORTED_MODEL;
  }
  return validate_options(*o, 0);
}

void set_options(cddp_b200_solver *s, const cddp_b200_options &o) {
  s->c.opt = o;
  s->c.num_alphas = build_alphas(o, s->c.alphas, CDDP_B200_MAX_ALPHAS);
  s->d.num_alphas = s->c.num_alphas;
  s->c.ls_window = s->ls_window;
  s->c.speculate = s->speculate;
  s->c.fuse_lin = s->fuse_lin;
}

// Selects the record layout (records.cuh): allocates the record buffer of that layout on first use and
// uploads the index tables the generic pack/unpack kernels use.
int apply_layout(cddp_b200_solver *s, int layout) {
  DeviceState &d = s->d;
  if (layout == RECORDS_STRUCTURED && !model_has_structured_layout(s->c.model)) layout = RECORDS_DENSE;
  RecordMap map;
  fill_record_map_for(map, s->c.model, d.n, d.m, layout == RECORDS_STRUCTURED);
  if (!s->rec_by_layout[layout]) {
    double *p = nullptr;
    const size_t cnt = (size_t)d.B * d.N * map.stride;
    int r = s->alloc(&p, cnt);
    if (r) return r;
    CU(cudaMemsetAsync(p, 0, cnt * sizeof(double), s->stream));
    s->rec_by_layout[layout] = p;
  }
  if (!s->didxA) {
    int r;
    if ((r = s->alloc(&s->didxA, (size_t)CDDP_B200_MAX_N * CDDP_B200_MAX_N))) return r;
    if ((r = s->alloc(&s->didxB, (size_t)CDDP_B200_MAX_N * CDDP_B200_MAX_M))) return r;
  }
  CU(cudaMemcpyAsync(s->didxA, map.idxA, sizeof(map.idxA), cudaMemcpyHostToDevice, s->stream));
  CU(cudaMemcpyAsync(s->didxB, map.idxB, sizeof(map.idxB), cudaMemcpyHostToDevice, s->stream));
  CU(cudaStreamSynchronize(s->stream));  // `map` is a stack object
  d.rec = s->rec_by_layout[layout];
  d.rec_stride = map.stride;
  d.layout = layout;
  d.idxA = s->didxA;
  d.idxB = s->didxB;
  d.offLx = map.offLx; d.offLu = map.offLu; d.offU = map.offU;
  return 0;
}

struct KernelTimer {
  cddp_b200_solver *s;
  double *acc;
  bool on;
  KernelTimer(cddp_b200_solver *s_, double *acc_) : s(s_), acc(acc_), on(s_->timing_enabled) {
    if (on) cudaEventRecord(s->ev0, s->stream);
  }
  void stop() {
    if (!on) return;
    cudaEventRecord(s->ev1, s->stream);
    cudaEventSynchronize(s->ev1);
    float ms = 0.f;
    cudaEventElapsedTime(&ms, s->ev0, s->ev1);
    *acc += ms;
  }
};

int do_linearize(cddp_b200_solver *s, bool force) {
  KernelTimer t(s, &s->timing.linearize_ms);
  CU(launch_linearize(s->c, s->d, force, s->stream));
  t.stop();
  s->timing.linearize_launches++;
  return 0;
}
int do_backward(cddp_b200_solver *s, int mode) {
  KernelTimer t(s, &s->timing.backward_ms);
  if (s->kind == 1) CU(launch_ip_backward(s->c, s->d, s->ic, s->ip, mode, s->stream));
  else CU(launch_backward(s->c, s->d, mode, s->stream));
  t.stop();
  s->timing.backward_launches++;
  return 0;
}
int do_forward(cddp_b200_solver *s, int mode) {
  KernelTimer t(s, &s->timing.forward_ms);
  if (s->kind == 1) CU(launch_ip_forward(s->c, s->d, s->ic, s->ip, mode, s->stream));
  else CU(launch_forward(s->c, s->d, mode, s->stream));
  t.stop();
  s->timing.forward_launches++;
  return 0;
}

int one_iteration(cddp_b200_solver *s) {
  int r;
  // CLDDP sweeps that form their own linearisation records (fused linearisation, backward_fast.cu) need no linearize launch
  if (!(s->kind == 0 && backward_fuses_linearization(s->c, s->d)) && (r = do_linearize(s, false))) return r;
  if ((r = do_backward(s, BW_ITERATE))) return r;
  if ((r = do_forward(s, FW_ITERATE))) return r;
  return 0;
}

int read_running(cddp_b200_solver *s, int *running) {
  CU(launch_count_running(s->d, s->stream));
  s->timing.other_launches++;
  CU(cudaMemcpyAsync(s->h_running, s->d.num_running, sizeof(int), cudaMemcpyDeviceToHost, s->stream));
  CU(cudaStreamSynchronize(s->stream));
  *running = *s->h_running;
  return 0;
}

int upload(cddp_b200_solver *s, double *dst, const double *src, size_t count, bool device_src) {
  if (!src) {
    CU(cudaMemsetAsync(dst, 0, count * sizeof(double), s->stream));
    return 0;
  }
  CU(cudaMemcpyAsync(dst, src, count * sizeof(double), device_src ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice,
                     s->stream));
  return 0;
}

int set_instances_impl(cddp_b200_solver *s, const double *x0, const double *xref, const double *ref_traj,
                       const double *X0, const double *U0, bool dev) {
  if (!s || !x0 || !xref) return CDDP_B200_ERR_INVALID_ARGUMENT;
  DeviceGuard g(s->device);
  if (!g.ok) return cuda_fail(cudaErrorInvalidDevice, "cudaSetDevice");
  const DeviceState &d = s->d;
  const size_t B = d.B, n = d.n, m = d.m, N = d.N;
  int r;
  if ((r = upload(s, d.x0, x0, B * n, dev))) return r;
  if ((r = upload(s, d.xref, xref, B * n, dev))) return r;
  if (ref_traj) {
    if (!s->d.ref_traj) {
      double *p = nullptr;
      if ((r = s->alloc(&p, B * (N + 1) * n))) return r;
      s->d.ref_traj = p;
    }
    if ((r = upload(s, s->d.ref_traj, ref_traj, B * (N + 1) * n, dev))) return r;
  } else {
    s->d.ref_traj = nullptr;  // (buffer, if any, stays owned by allocs)
  }
  CU(cudaMemsetAsync(d.cur, 0, B * sizeof(int), s->stream));
  if ((r = upload(s, d.X[0], X0, B * (N + 1) * n, dev))) return r;
  if ((r = upload(s, d.U[0], U0, B * N * m, dev))) return r;
  s->have_instances = true;
  s->initialized = false;
  return 0;
}

}  // namespace

extern "C" {

int cddp_b200_abi_version(void) { return CDDP_B200_ABI_VERSION; }

const char *cddp_b200_error_string(int err) {
  switch (err) {
    case CDDP_B200_OK: return "ok";
    case CDDP_B200_ERR_INVALID_ARGUMENT: return "invalid argument";
    case CDDP_B200_ERR_UNSUPPORTED_MODEL:
      return "unsupported model: the dynamics have no device implementation (there is no CPU fallback)";
    case CDDP_B200_ERR_CUDA: return "CUDA error (see cddp_b200_last_cuda_error)";
    case CDDP_B200_ERR_OUT_OF_MEMORY: return "out of device memory";
    case CDDP_B200_ERR_STATE: return "invalid call order";
    case CDDP_B200_ERR_USER_MODEL: return "the user-supplied dynamics source did not compile (see cddp_b200_last_compile_log)";
    default: return "unknown error";
  }
}

const char *cddp_b200_last_cuda_error(void) { return g_last_cuda_error.c_str(); }

const char *cddp_b200_status_string(int status) {
  switch (status) {
    case CDDP_B200_STATUS_OPTIMAL: return "OptimalSolutionFound";
    case CDDP_B200_STATUS_ACCEPTABLE: return "AcceptableSolutionFound";
    case CDDP_B200_STATUS_MAX_ITERATIONS: return "MaxIterationsReached";
    case CDDP_B200_STATUS_REG_LIMIT: return "RegularizationLimitReached_NotConverged";
    case CDDP_B200_STATUS_MAX_CPU_TIME: return "MaxCpuTimeReached";
    default: return "Running";
  }
}

int cddp_b200_device_count(int *count) {
  if (!count) return CDDP_B200_ERR_INVALID_ARGUMENT;
  *count = 0;
  CU(cudaGetDeviceCount(count));
  return 0;
}

void cddp_b200_default_options(cddp_b200_options *o) {
  if (!o) return;
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

int cddp_b200_build_alphas(const cddp_b200_options *opts, double *alphas, int capacity, int *count) {
  if (!opts || !alphas || !count || capacity < 1) return CDDP_B200_ERR_INVALID_ARGUMENT;
  *count = build_alphas(*opts, alphas, capacity);
  return 0;
}

const char *cddp_b200_last_compile_log(void) { return g_last_compile_log.c_str(); }

int cddp_b200_compile_user_model(const char *model_source, int n, int m, size_t *cubin_bytes) {
  if (!model_source || n < 1 || n > CDDP_B200_MAX_N || m < 1 || m > CDDP_B200_MAX_M) return CDDP_B200_ERR_INVALID_ARGUMENT;
  g_last_compile_log.clear();
  return user_model_compile_only(model_source, n, m, g_last_compile_log, cubin_bytes);
}

int cddp_b200_create(const cddp_b200_problem *p, const cddp_b200_options *o, int batch, int device, cddp_b200_solver **out) {
  return cddp_b200_create_ex(p, o, nullptr, batch, device, out);
}

int cddp_b200_create_ex(const cddp_b200_problem *p, const cddp_b200_options *o, const char *model_source, int batch, int device,
                        cddp_b200_solver **out) {
  if (!out) return CDDP_B200_ERR_INVALID_ARGUMENT;
  *out = nullptr;
  int r = validate(p, o, batch, model_source != nullptr);
  if (r) return r;
  if (model_source && p->model != CDDP_B200_MODEL_USER) return CDDP_B200_ERR_INVALID_ARGUMENT;
  int ndev = 0;
  CU(cudaGetDeviceCount(&ndev));
  if (device < 0 || device >= ndev) return cuda_fail(cudaErrorInvalidDevice, "device index");
  DeviceGuard g(device);
  if (!g.ok) return cuda_fail(cudaErrorInvalidDevice, "cudaSetDevice");
  cddp_b200_solver *s = new (std::nothrow) cddp_b200_solver();
  if (!s) return CDDP_B200_ERR_OUT_OF_MEMORY;
  s->device = device;
  const int n = p->n, m = p->m, N = p->horizon;
  const size_t B = batch;
  Constants &c = s->c;
  DeviceState &d = s->d;
  c.model = p->model; c.n = n; c.m = m; c.N = N; c.integrator = p->integrator; c.has_box = p->has_control_box ? 1 : 0;
  c.dt = p->dt;
  std::memset(c.mp.p, 0, sizeof(c.mp.p));
  std::memcpy(c.mp.p, p->model_params, sizeof(p->model_params));
  if (p->model == CDDP_B200_MODEL_QUADROTOR) Model<CDDP_B200_MODEL_QUADROTOR>::prepare(c.mp);
  c.mp.n = n; c.mp.m = m;
  for (int i = 0; i < CDDP_B200_MAX_M; ++i) {
    c.lb[i] = (p->has_control_box && i < m) ? p->lb[i] : -INFINITY;
    c.ub[i] = (p->has_control_box && i < m) ? p->ub[i] : INFINITY;
  }
  d.B = batch; d.n = n; d.m = m; d.N = N;
  d.order = nullptr; d.n_slots = batch;
  c.q_diag = 1;
  for (int i = 0; i < n; ++i)
    for (int j = 0; j < n; ++j)
      if (i != j && p->Q[i * n + j] != 0.0) c.q_diag = 0;
  c.cost_diag = c.q_diag;
  for (int i = 0; i < n; ++i)
    for (int j = 0; j < n; ++j)
      if (i != j && p->Qf[i * n + j] != 0.0) c.cost_diag = 0;
  for (int i = 0; i < m; ++i)
    for (int j = 0; j < m; ++j)
      if (i != j && p->R[i * m + j] != 0.0) c.cost_diag = 0;
  set_options(s, *o);

#define AL(ptr, cnt)                       \
  do {                                     \
    if ((r = s->alloc(&(ptr), (cnt)))) {   \
      cddp_b200_destroy(s);                \
      return r;                            \
    }                                      \
  } while (0)
  // batch-shared constants: l_xx = 2 Q dt, l_uu = 2 R dt, phi_xx = 2 Qf (objective.cpp:38-39,130-154)
  std::vector<double> hQ((size_t)n * n), hR((size_t)m * m), hQf((size_t)n * n);
  for (int i = 0; i < n * n; ++i) { hQ[i] = 2.0 * (p->Q[i] * p->dt); hQf[i] = 2.0 * p->Qf[i]; }
  for (int i = 0; i < m * m; ++i) hR[i] = 2.0 * (p->R[i] * p->dt);
  AL(s->dQdt2, (size_t)n * n); AL(s->dRdt2, (size_t)m * m); AL(s->dQf2, (size_t)n * n);
  cudaError_t e;
  e = cudaMemcpy(s->dQdt2, hQ.data(), hQ.size() * sizeof(double), cudaMemcpyHostToDevice);
  if (e == cudaSuccess) e = cudaMemcpy(s->dRdt2, hR.data(), hR.size() * sizeof(double), cudaMemcpyHostToDevice);
  if (e == cudaSuccess) e = cudaMemcpy(s->dQf2, hQf.data(), hQf.size() * sizeof(double), cudaMemcpyHostToDevice);
  if (e == cudaSuccess && p->model == CDDP_B200_MODEL_LTI) {
    AL(s->dltiA, (size_t)n * n); AL(s->dltiB, (size_t)n * m);
    e = cudaMemcpy(s->dltiA, p->lti_A, (size_t)n * n * sizeof(double), cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = cudaMemcpy(s->dltiB, p->lti_B, (size_t)n * m * sizeof(double), cudaMemcpyHostToDevice);
  }
  if (e != cudaSuccess) { cddp_b200_destroy(s); return cuda_fail(e, "cudaMemcpy(constants)"); }
  c.Qdt2 = s->dQdt2; c.Rdt2 = s->dRdt2; c.Qf2 = s->dQf2;
  c.mp.lti_A = s->dltiA; c.mp.lti_B = s->dltiB;

  AL(d.X[0], B * (N + 1) * n); AL(d.X[1], B * (N + 1) * n);
  AL(d.U[0], B * N * m); AL(d.U[1], B * N * m);
  if ((r = apply_layout(s, RECORDS_STRUCTURED))) {  // falls back to dense for models without a pattern
    cddp_b200_destroy(s);
    return r;
  }
  AL(d.vterm, B * n);
  AL(d.K, B * N * m * n); AL(d.kff, B * N * m);
  AL(d.x0, B * n); AL(d.xref, B * n);
  d.ref_traj = nullptr;
  AL(d.cur, B); AL(d.status, B); AL(d.iter, B); AL(d.lin_valid, B); AL(d.bw_ok, B); AL(d.accepted, B); AL(d.fw_done, B); AL(s->order_buf, B);
  AL(d.reg, B); AL(d.cost, B); AL(d.alpha, B); AL(d.inf_du, B); AL(d.dV, 2 * B);
  AL(d.ls_cost, B * CDDP_B200_MAX_ALPHAS);
  AL(d.Vx0, B * n); AL(d.Vxx0, B * n * n);
  s->ckpt_lg = s->c.num_alphas > 16 ? 32 : 16;
  AL(d.ckpt, B * s->ckpt_lg * s->ckpt_lg * n);
  AL(d.num_running, 1);
  d.history = nullptr; d.history_len = nullptr; d.history_cap = 0;
  d.trace = nullptr; d.trace_cap = 0;
#undef AL
  e = cudaMemset(d.cur, 0, B * sizeof(int));
  if (e == cudaSuccess) e = cudaMemset(d.status, 0, B * sizeof(int));
  if (e == cudaSuccess) e = cudaMemset(d.fw_done, 0, B * sizeof(int));
  if (e == cudaSuccess) e = cudaMemset(d.K, 0, B * N * m * n * sizeof(double));
  if (e == cudaSuccess) e = cudaMemset(d.kff, 0, B * N * m * sizeof(double));
  if (e == cudaSuccess) e = cudaMemset(d.ls_cost, 0, B * CDDP_B200_MAX_ALPHAS * sizeof(double));
  if (e == cudaSuccess) e = cudaMallocHost((void **)&s->h_running, sizeof(int));
  if (e == cudaSuccess) e = cudaEventCreate(&s->ev0);
  if (e == cudaSuccess) e = cudaEventCreate(&s->ev1);
  if (e != cudaSuccess) { cddp_b200_destroy(s); return cuda_fail(e, "solver setup"); }
  d.user = nullptr;
  if (model_source) {  // compile the model-dependent kernels around the user's dynamics and load them on this device
    g_last_compile_log.clear();
    r = user_model_build(model_source, n, m, c.cost_diag != 0, &s->user_kernels, g_last_compile_log);
    if (r) {
      if (r == CDDP_B200_ERR_CUDA) g_last_cuda_error = g_last_compile_log;
      cddp_b200_destroy(s);
      return r;
    }
    d.user = s->user_kernels;
  }
  *out = s;
  return 0;
}

int cddp_b200_destroy(cddp_b200_solver *s) {
  if (!s) return 0;
  DeviceGuard g(s->device);
  cudaDeviceSynchronize();
  for (void *p : s->allocs) cudaFree(p);
  if (s->user_kernels) user_model_destroy(s->user_kernels);
  if (s->scratch) cudaFree(s->scratch);
  if (s->h_running) cudaFreeHost(s->h_running);
  if (s->ev0) cudaEventDestroy(s->ev0);
  if (s->ev1) cudaEventDestroy(s->ev1);
  delete s;
  return 0;
}

int cddp_b200_set_stream(cddp_b200_solver *s, void *stream) {
  if (!s) return CDDP_B200_ERR_INVALID_ARGUMENT;
  s->stream = static_cast<cudaStream_t>(stream);
  return 0;
}

int cddp_b200_set_options(cddp_b200_solver *s, const cddp_b200_options *o) {
  if (!s || !o) return CDDP_B200_ERR_INVALID_ARGUMENT;
  if (int r = validate_options(*o, s->kind)) return r;
  if (s->d.history && o->max_iterations + 1 > s->d.history_cap) return CDDP_B200_ERR_STATE;
  if (s->d.trace && o->max_iterations > s->d.trace_cap) return CDDP_B200_ERR_STATE;
  if (count_alphas(*o) > 16 && s->ckpt_lg < 32) {  // more than 16 alphas: one trajectory per warp, bigger scratch
    DeviceGuard g(s->device);
    double *p = nullptr;
    int r = s->alloc(&p, (size_t)s->d.B * 32 * 32 * s->d.n);
    if (r) return r;
    s->d.ckpt = p;
    s->ckpt_lg = 32;
  }
  set_options(s, *o);
  return 0;
}

int cddp_b200_set_record_layout(cddp_b200_solver *s, int layout) {
  if (!s || (layout != CDDP_B200_RECORDS_DENSE && layout != CDDP_B200_RECORDS_STRUCTURED)) return CDDP_B200_ERR_INVALID_ARGUMENT;
  if (s->kind == 1 && layout != CDDP_B200_RECORDS_DENSE) return CDDP_B200_ERR_INVALID_ARGUMENT;  // the IPDDP sweep reads dense records
  DeviceGuard g(s->device);
  const bool had = s->initialized;
  int r = apply_layout(s, layout);
  if (r) return r;
  if (had) {  // re-derive the records of the current nominal trajectory in the new layout
    CU(launch_linearize(s->c, s->d, true, s->stream));
    s->timing.other_launches++;
  }
  return 0;
}

int cddp_b200_get_record_layout(cddp_b200_solver *s, int *layout, int *record_bytes) {
  if (!s) return CDDP_B200_ERR_INVALID_ARGUMENT;
  if (layout) *layout = s->d.layout;
  if (record_bytes) *record_bytes = s->d.rec_stride * (int)sizeof(double);
  return 0;
}

int cddp_b200_set_instances(cddp_b200_solver *s, const double *x0, const double *xref, const double *ref_traj,
                            const double *X0, const double *U0) {
  return set_instances_impl(s, x0, xref, ref_traj, X0, U0, false);
}

int cddp_b200_set_instances_device(cddp_b200_solver *s, const double *x0, const double *xref, const double *ref_traj,
                                   const double *X0, const double *U0) {
  return set_instances_impl(s, x0, xref, ref_traj, X0, U0, true);
}

int cddp_b200_initialize(cddp_b200_solver *s) {
  if (!s) return CDDP_B200_ERR_INVALID_ARGUMENT;
  if (!s->have_instances) return CDDP_B200_ERR_STATE;
  DeviceGuard g(s->device);
  s->d.order = nullptr;  // every instance runs again
  s->d.n_slots = s->d.B;
  if (s->kind == 1) CU(launch_ip_initialize(s->c, s->d, s->ic, s->ip, s->stream));
  else CU(launch_initialize(s->c, s->d, s->stream));
  s->timing.other_launches++;
  s->initialized = true;
  return 0;
}

int cddp_b200_linearize(cddp_b200_solver *s) {
  if (!s) return CDDP_B200_ERR_INVALID_ARGUMENT;
  if (!s->initialized) return CDDP_B200_ERR_STATE;
  DeviceGuard g(s->device);
  return do_linearize(s, true);
}

int cddp_b200_backward_pass(cddp_b200_solver *s) {
  if (!s) return CDDP_B200_ERR_INVALID_ARGUMENT;
  if (!s->initialized) return CDDP_B200_ERR_STATE;
  DeviceGuard g(s->device);
  return do_backward(s, BW_SINGLE);
}

int cddp_b200_forward_pass(cddp_b200_solver *s) {
  if (!s) return CDDP_B200_ERR_INVALID_ARGUMENT;
  if (!s->initialized) return CDDP_B200_ERR_STATE;
  DeviceGuard g(s->device);
  return do_forward(s, FW_EVALUATE);
}

int cddp_b200_iterate(cddp_b200_solver *s, int iterations) {
  if (!s || iterations < 0) return CDDP_B200_ERR_INVALID_ARGUMENT;
  if (!s->initialized) return CDDP_B200_ERR_STATE;
  DeviceGuard g(s->device);
  for (int i = 0; i < iterations; ++i) {
    int r = one_iteration(s);
    if (r) return r;
  }
  return 0;
}

int cddp_b200_solve(cddp_b200_solver *s) {
  if (!s) return CDDP_B200_ERR_INVALID_ARGUMENT;
  if (!s->have_instances) return CDDP_B200_ERR_STATE;
  DeviceGuard g(s->device);
  int r = cddp_b200_initialize(s);
  if (r) return r;
  const auto t0 = std::chrono::steady_clock::now();
  const int max_it = s->c.opt.max_iterations;
  int final_status = CDDP_B200_STATUS_MAX_ITERATIONS;
  struct WorkListReset {  // single-step entry points and the next solve address the whole batch, also after an error return
    cddp_b200_solver *s;
    ~WorkListReset() {
      s->d.order = nullptr;
      s->d.n_slots = s->d.B;
    }
  } work_list_reset{s};
  // poll the running counter with a widening stride: cheap for short solves, rare for long ones.  With
  // poll_interval == 0 the solve is enqueued without any host synchronisation (finished instances are masked on
  // the device, so running the remaining launches is correct, just not free).
  // A poll is one small kernel + one stream synchronisation (~20 us): for a batch whose iteration takes a millisecond or
  // more it is taken every iteration from the second on, so that the work list (DeviceState::order) shrinks as soon as
  // instances finish; small batches keep the widening stride.
  const bool heavy = (long long)s->d.B * s->d.N >= 32768;
  int next_poll = s->poll_interval > 0 ? s->poll_interval : (heavy ? 2 : 4);
  for (int it = 0; it < max_it; ++it) {
    if (s->c.opt.max_cpu_time > 0.0) {  // cddp_solver_base.cpp:77-90 (wall clock, checked per batched iteration)
      CU(cudaStreamSynchronize(s->stream));
      const double el = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
      if (el > s->c.opt.max_cpu_time) {
        final_status = CDDP_B200_STATUS_MAX_CPU_TIME;
        break;
      }
    }
    if ((r = one_iteration(s))) return r;
    if (s->poll_interval != 0 && it + 1 == next_poll && it + 1 < max_it) {
      // count the running instances and compact them into the work list: the launches that follow cover only those
      int running = 0;
      CU(launch_compact_running(s->d, s->order_buf, s->stream));
      s->timing.other_launches++;
      CU(cudaMemcpyAsync(s->h_running, s->d.num_running, sizeof(int), cudaMemcpyDeviceToHost, s->stream));
      CU(cudaStreamSynchronize(s->stream));
      running = *s->h_running;
      if (running == 0) break;
      s->d.order = s->order_buf;
      s->d.n_slots = running;
      next_poll += s->poll_interval > 0 ? s->poll_interval : (heavy ? 1 : ((next_poll < 16) ? 4 : 8));
    }
  }
  s->d.order = nullptr;  // single-step entry points and the next solve address the whole batch
  s->d.n_slots = s->d.B;
  CU(launch_finalize(s->c, s->d, final_status, s->stream));
  s->timing.other_launches++;
  return 0;
}

int cddp_b200_num_running(cddp_b200_solver *s, int *running) {
  if (!s || !running) return CDDP_B200_ERR_INVALID_ARGUMENT;
  DeviceGuard g(s->device);
  return read_running(s, running);
}

int cddp_b200_synchronize(cddp_b200_solver *s) {
  if (!s) return CDDP_B200_ERR_INVALID_ARGUMENT;
  DeviceGuard g(s->device);
  CU(cudaStreamSynchronize(s->stream));
  return 0;
}

static int download(cddp_b200_solver *s, void *dst, const void *src, size_t bytes) {
  if (!dst) return 0;
  CU(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, s->stream));
  return 0;
}

static int get_solution_impl(cddp_b200_solver *s, double *X, double *U, double *K, double *final_objective,
                             int *iterations_completed, int *status, double *final_step_length, double *final_regularization,
                             double *inf_du, bool sync) {
  if (!s) return CDDP_B200_ERR_INVALID_ARGUMENT;
  DeviceGuard g(s->device);
  const DeviceState &d = s->d;
  const size_t B = d.B, n = d.n, m = d.m, N = d.N;
  int r;
  if (X || U) {
    const size_t nx = B * (N + 1) * n, nu = B * N * m;
    if ((r = s->ensure_scratch((nx + nu) * sizeof(double)))) return r;
    CU(launch_gather_current(s->c, d, X ? s->scratch : nullptr, U ? s->scratch + nx : nullptr, 0, s->stream));
    s->timing.other_launches++;
    if ((r = download(s, X, s->scratch, nx * sizeof(double)))) return r;
    if ((r = download(s, U, s->scratch + nx, nu * sizeof(double)))) return r;
  }
  if ((r = download(s, K, d.K, B * N * m * n * sizeof(double)))) return r;
  if ((r = download(s, final_objective, d.cost, B * sizeof(double)))) return r;
  if ((r = download(s, iterations_completed, d.iter, B * sizeof(int)))) return r;
  if ((r = download(s, status, d.status, B * sizeof(int)))) return r;
  if ((r = download(s, final_step_length, d.alpha, B * sizeof(double)))) return r;
  if ((r = download(s, final_regularization, d.reg, B * sizeof(double)))) return r;
  if ((r = download(s, inf_du, d.inf_du, B * sizeof(double)))) return r;
  if (sync) CU(cudaStreamSynchronize(s->stream));
  return 0;
}

int cddp_b200_get_solution(cddp_b200_solver *s, double *X, double *U, double *K, double *final_objective,
                           int *iterations_completed, int *status, double *final_step_length,
                           double *final_regularization, double *inf_du) {
  return get_solution_impl(s, X, U, K, final_objective, iterations_completed, status, final_step_length, final_regularization,
                           inf_du, true);
}

int cddp_b200_get_solution_async(cddp_b200_solver *s, double *X, double *U, double *K, double *final_objective,
                                 int *iterations_completed, int *status, double *final_step_length,
                                 double *final_regularization, double *inf_du) {
  return get_solution_impl(s, X, U, K, final_objective, iterations_completed, status, final_step_length, final_regularization,
                           inf_du, false);
}

int cddp_b200_mpc_advance(cddp_b200_solver *s, int steps, const double *x0_new, const double *xref_new) {
  if (!s || steps < 0 || steps > s->d.N) return CDDP_B200_ERR_INVALID_ARGUMENT;
  if (!s->have_instances) return CDDP_B200_ERR_STATE;
  DeviceGuard g(s->device);
  const DeviceState &d = s->d;
  if (steps > 0) {
    CU(launch_shift(d, steps, s->stream));
    s->timing.other_launches++;
  }
  if (x0_new) CU(cudaMemcpyAsync(d.x0, x0_new, (size_t)d.B * d.n * sizeof(double), cudaMemcpyHostToDevice, s->stream));
  if (xref_new) CU(cudaMemcpyAsync(d.xref, xref_new, (size_t)d.B * d.n * sizeof(double), cudaMemcpyHostToDevice, s->stream));
  s->initialized = false;
  return 0;
}

int cddp_b200_get_first_controls_async(cddp_b200_solver *s, double *u0, double *final_objective, int *status) {
  if (!s || !u0) return CDDP_B200_ERR_INVALID_ARGUMENT;
  DeviceGuard g(s->device);
  const DeviceState &d = s->d;
  int r;
  if ((r = s->ensure_scratch((size_t)d.B * d.m * sizeof(double)))) return r;
  CU(launch_first_controls(d, s->scratch, s->stream));
  s->timing.other_launches++;
  if ((r = download(s, u0, s->scratch, (size_t)d.B * d.m * sizeof(double)))) return r;
  if ((r = download(s, final_objective, d.cost, (size_t)d.B * sizeof(double)))) return r;
  if ((r = download(s, status, d.status, (size_t)d.B * sizeof(int)))) return r;
  return 0;
}

int cddp_b200_set_poll_interval(cddp_b200_solver *s, int interval) {
  if (!s || interval < -1) return CDDP_B200_ERR_INVALID_ARGUMENT;
  s->poll_interval = interval;
  return 0;
}

int cddp_b200_set_line_search_window(cddp_b200_solver *s, int enable) {
  if (!s) return CDDP_B200_ERR_INVALID_ARGUMENT;
  s->ls_window = enable ? 1 : 0;
  s->c.ls_window = s->ls_window;
  return 0;
}

int cddp_b200_set_fused_linearization(cddp_b200_solver *s, int enable) {
  if (!s) return CDDP_B200_ERR_INVALID_ARGUMENT;
  s->fuse_lin = enable ? 1 : 0;
  s->c.fuse_lin = s->fuse_lin;
  return 0;
}

int cddp_b200_set_first_alpha_speculation(cddp_b200_solver *s, int mode) {
  if (!s || mode < -1 || mode > 1) return CDDP_B200_ERR_INVALID_ARGUMENT;
  s->speculate = mode;
  s->c.speculate = mode;
  return 0;
}

int cddp_b200_enable_history(cddp_b200_solver *s, int enable) {
  if (!s) return CDDP_B200_ERR_INVALID_ARGUMENT;
  DeviceGuard g(s->device);
  if (!enable) {
    s->d.history = nullptr;  // the buffer stays allocated (history_buf) for a later re-enable
    return 0;
  }
  const int cap = s->c.opt.max_iterations + 1;
  if (!s->history_buf || cap > s->d.history_cap) {
    double *h = nullptr;
    int *l = nullptr;
    int r;
    if ((r = s->alloc(&h, (size_t)s->d.B * cap * (s->kind == 1 ? IP_HISTORY_COLS : 4)))) return r;
    if ((r = s->alloc(&l, (size_t)s->d.B))) return r;
    s->history_buf = h;
    s->d.history_len = l;
    s->d.history_cap = cap;
    CU(cudaMemset(l, 0, (size_t)s->d.B * sizeof(int)));
  }
  s->d.history = s->history_buf;
  return 0;
}

int cddp_b200_enable_trace(cddp_b200_solver *s, int enable) {
  if (!s) return CDDP_B200_ERR_INVALID_ARGUMENT;
  DeviceGuard g(s->device);
  if (!enable) {
    s->d.trace = nullptr;
    return 0;
  }
  const int cap = s->c.opt.max_iterations > 0 ? s->c.opt.max_iterations : 1;
  if (!s->trace_buf || cap > s->d.trace_cap) {
    int *t = nullptr;
    int r;
    if ((r = s->alloc(&t, (size_t)s->d.B * cap))) return r;
    s->trace_buf = t;
    s->d.trace_cap = cap;
  }
  CU(cudaMemsetAsync(s->trace_buf, 0, (size_t)s->d.B * s->d.trace_cap * sizeof(int), s->stream));
  s->d.trace = s->trace_buf;
  return 0;
}

int cddp_b200_get_trace(cddp_b200_solver *s, int *trace, int *cap) {
  if (!s) return CDDP_B200_ERR_INVALID_ARGUMENT;
  if (!s->d.trace) return CDDP_B200_ERR_STATE;
  DeviceGuard g(s->device);
  if (cap) *cap = s->d.trace_cap;
  int r;
  if ((r = download(s, trace, s->d.trace, (size_t)s->d.B * s->d.trace_cap * sizeof(int)))) return r;
  CU(cudaStreamSynchronize(s->stream));
  return 0;
}

int cddp_b200_get_history(cddp_b200_solver *s, double *history, int *lens) {
  if (!s) return CDDP_B200_ERR_INVALID_ARGUMENT;
  if (!s->d.history || s->kind == 1) return CDDP_B200_ERR_STATE;
  DeviceGuard g(s->device);
  int r;
  if ((r = download(s, history, s->d.history, (size_t)s->d.B * s->d.history_cap * 4 * sizeof(double)))) return r;
  if ((r = download(s, lens, s->d.history_len, (size_t)s->d.B * sizeof(int)))) return r;
  CU(cudaStreamSynchronize(s->stream));
  return 0;
}

int cddp_b200_get_feedforward(cddp_b200_solver *s, double *k) {
  if (!s || !k) return CDDP_B200_ERR_INVALID_ARGUMENT;
  DeviceGuard g(s->device);
  int r = download(s, k, s->d.kff, (size_t)s->d.B * s->d.N * s->d.m * sizeof(double));
  if (r) return r;
  CU(cudaStreamSynchronize(s->stream));
  return 0;
}

int cddp_b200_set_gains(cddp_b200_solver *s, const double *K, const double *k) {
  if (!s) return CDDP_B200_ERR_INVALID_ARGUMENT;
  DeviceGuard g(s->device);
  const DeviceState &d = s->d;
  if (K) CU(cudaMemcpyAsync(d.K, K, (size_t)d.B * d.N * d.m * d.n * sizeof(double), cudaMemcpyHostToDevice, s->stream));
  if (k) CU(cudaMemcpyAsync(d.kff, k, (size_t)d.B * d.N * d.m * sizeof(double), cudaMemcpyHostToDevice, s->stream));
  CU(cudaStreamSynchronize(s->stream));
  return 0;
}

int cddp_b200_set_regularization(cddp_b200_solver *s, const double *reg) {
  if (!s || !reg) return CDDP_B200_ERR_INVALID_ARGUMENT;
  DeviceGuard g(s->device);
  CU(cudaMemcpyAsync(s->d.reg, reg, (size_t)s->d.B * sizeof(double), cudaMemcpyHostToDevice, s->stream));
  CU(cudaStreamSynchronize(s->stream));
  return 0;
}

int cddp_b200_set_cost(cddp_b200_solver *s, const double *cost) {
  if (!s || !cost) return CDDP_B200_ERR_INVALID_ARGUMENT;
  DeviceGuard g(s->device);
  CU(cudaMemcpyAsync(s->d.cost, cost, (size_t)s->d.B * sizeof(double), cudaMemcpyHostToDevice, s->stream));
  CU(cudaStreamSynchronize(s->stream));
  return 0;
}

int cddp_b200_get_linearization(cddp_b200_solver *s, double *A, double *Bm) {
  if (!s || !A || !Bm) return CDDP_B200_ERR_INVALID_ARGUMENT;
  DeviceGuard g(s->device);
  const DeviceState &d = s->d;
  const size_t na = (size_t)d.B * d.N * d.n * d.n, nb = (size_t)d.B * d.N * d.n * d.m;
  int r;
  if ((r = s->ensure_scratch((na + nb) * sizeof(double)))) return r;
  CU(launch_unpack_linearization(s->c, d, s->scratch, s->scratch + na, s->stream));
  s->timing.other_launches++;
  if ((r = download(s, A, s->scratch, na * sizeof(double)))) return r;
  if ((r = download(s, Bm, s->scratch + na, nb * sizeof(double)))) return r;
  CU(cudaStreamSynchronize(s->stream));
  return 0;
}

int cddp_b200_set_linearization(cddp_b200_solver *s, const double *A, const double *Bm) {
  if (!s || !A || !Bm) return CDDP_B200_ERR_INVALID_ARGUMENT;
  DeviceGuard g(s->device);
  const DeviceState &d = s->d;
  const size_t na = (size_t)d.B * d.N * d.n * d.n, nb = (size_t)d.B * d.N * d.n * d.m;
  int r;
  if (d.layout != RECORDS_DENSE) {
    // caller-supplied Jacobians are dense by definition: carry the cost terms over to dense records
    if ((r = cddp_b200_set_record_layout(s, CDDP_B200_RECORDS_DENSE))) return r;
  }
  if ((r = s->ensure_scratch((na + nb) * sizeof(double)))) return r;
  CU(cudaMemcpyAsync(s->scratch, A, na * sizeof(double), cudaMemcpyHostToDevice, s->stream));
  CU(cudaMemcpyAsync(s->scratch + na, Bm, nb * sizeof(double), cudaMemcpyHostToDevice, s->stream));
  CU(launch_pack_linearization(s->c, d, s->scratch, s->scratch + na, s->stream));
  s->timing.other_launches++;
  CU(cudaStreamSynchronize(s->stream));
  return 0;
}

int cddp_b200_get_sweep(cddp_b200_solver *s, double *dV, int *ok, double *inf_du, double *Vx0, double *Vxx0) {
  if (!s) return CDDP_B200_ERR_INVALID_ARGUMENT;
  DeviceGuard g(s->device);
  const DeviceState &d = s->d;
  int r;
  if ((r = download(s, dV, d.dV, (size_t)d.B * 2 * sizeof(double)))) return r;
  if ((r = download(s, ok, d.bw_ok, (size_t)d.B * sizeof(int)))) return r;
  if ((r = download(s, inf_du, d.inf_du, (size_t)d.B * sizeof(double)))) return r;
  if ((r = download(s, Vx0, d.Vx0, (size_t)d.B * d.n * sizeof(double)))) return r;
  if ((r = download(s, Vxx0, d.Vxx0, (size_t)d.B * d.n * d.n * sizeof(double)))) return r;
  CU(cudaStreamSynchronize(s->stream));
  return 0;
}

int cddp_b200_get_forward(cddp_b200_solver *s, double *costs, int *accepted, double *Xnew, double *Unew) {
  if (!s) return CDDP_B200_ERR_INVALID_ARGUMENT;
  DeviceGuard g(s->device);
  const DeviceState &d = s->d;
  const size_t B = d.B, n = d.n, m = d.m, N = d.N;
  int r;
  if (costs) {
    std::vector<double> tmp(B * CDDP_B200_MAX_ALPHAS);
    if ((r = download(s, tmp.data(), d.ls_cost, tmp.size() * sizeof(double)))) return r;
    CU(cudaStreamSynchronize(s->stream));
    for (size_t b = 0; b < B; ++b)
      for (int a = 0; a < s->c.num_alphas; ++a) costs[b * s->c.num_alphas + a] = tmp[b * CDDP_B200_MAX_ALPHAS + a];
  }
  if ((r = download(s, accepted, d.accepted, B * sizeof(int)))) return r;
  if (Xnew || Unew) {
    const size_t nx = B * (N + 1) * n, nu = B * N * m;
    if ((r = s->ensure_scratch((nx + nu) * sizeof(double)))) return r;
    CU(launch_gather_current(s->c, d, Xnew ? s->scratch : nullptr, Unew ? s->scratch + nx : nullptr, 1, s->stream));
    s->timing.other_launches++;
    if ((r = download(s, Xnew, s->scratch, nx * sizeof(double)))) return r;
    if ((r = download(s, Unew, s->scratch + nx, nu * sizeof(double)))) return r;
  }
  CU(cudaStreamSynchronize(s->stream));
  return 0;
}

/* ------------------------------------------------------------------------------------------------ IPDDP */
void cddp_b200_ipddp_default_options(cddp_b200_ipddp_options *io) { /* options.hpp:75-104,148-186 */
  if (!io) return;
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
  io->barrier_strategy = CDDP_B200_BARRIER_ADAPTIVE;
  io->jacobian_regularization_value = 1e-8;
  io->jacobian_regularization_exponent = 0.25;
  io->terminal_equality = 0;
}

int cddp_b200_ipddp_create(const cddp_b200_problem *p, const cddp_b200_options *o, const cddp_b200_ipddp_options *io,
                           const cddp_b200_constraint *cs, int nc, int batch, int device, cddp_b200_solver **out) {
  return cddp_b200_ipddp_create_ex(p, o, io, cs, nc, nullptr, batch, device, out);
}

int cddp_b200_ipddp_create_ex(const cddp_b200_problem *p, const cddp_b200_options *o, const cddp_b200_ipddp_options *io,
                              const cddp_b200_constraint *cs, int nc, const char *model_source, int batch, int device,
                              cddp_b200_solver **out) {
  if (!out) return CDDP_B200_ERR_INVALID_ARGUMENT;
  *out = nullptr;
  if (!p || !o || !io || nc < 0 || (nc > 0 && !cs)) return CDDP_B200_ERR_INVALID_ARGUMENT;
  if (p->model == CDDP_B200_MODEL_LTI) return CDDP_B200_ERR_UNSUPPORTED_MODEL;
  if (io->max_filter_size < 1 || io->max_filter_size >= IP_FILTER_CAP) return CDDP_B200_ERR_INVALID_ARGUMENT;
  if (count_alphas(*o) > 16) return CDDP_B200_ERR_INVALID_ARGUMENT;  // one lane per alpha, 16 lanes per trajectory
  const int n = p->n, m = p->m;
  // flatten the constraint set into rows
  int D = 0;
  for (int i = 0; i < nc; ++i) {
    switch (cs[i].type) {
      case CDDP_B200_CON_CONTROL_BOX: D += 2 * m; break;
      case CDDP_B200_CON_STATE_BOX: D += 2 * n; break;
      case CDDP_B200_CON_BALL:
        if (cs[i].rows < 1 || cs[i].rows > n) return CDDP_B200_ERR_INVALID_ARGUMENT;
        D += 1;
        break;
      case CDDP_B200_CON_LINEAR:
        if (cs[i].rows < 1) return CDDP_B200_ERR_INVALID_ARGUMENT;
        D += cs[i].rows;
        break;
      default: return CDDP_B200_ERR_INVALID_ARGUMENT;
    }
    if (!cs[i].p0 || !cs[i].p1) return CDDP_B200_ERR_INVALID_ARGUMENT;
  }
  if (D > IP_MAX_DUAL) return CDDP_B200_ERR_INVALID_ARGUMENT;
  cddp_b200_problem pc = *p;
  pc.has_control_box = 0;  // IPDDP never clamps (ipddp_solver.cpp:1650-1651); a ControlConstraint is a constraint row pair
  pc.lb = pc.ub = nullptr;
  cddp_b200_solver *s = nullptr;
  int r = cddp_b200_create_ex(&pc, o, model_source, batch, device, &s);
  if (r) return r;
  DeviceGuard g(device);
  s->kind = 1;
  const int Dn = D > 0 ? D : 1;
  std::vector<int> rt(Dn, 0), bd(Dn, 0);
  std::vector<double> Gx((size_t)Dn * n, 0.0), Gu((size_t)Dn * m, 0.0), off(Dn, 0.0), sc(Dn, 1.0);
  int row = 0;
  for (int i = 0; i < nc; ++i) {
    const cddp_b200_constraint &C = cs[i];
    if (C.type == CDDP_B200_CON_CONTROL_BOX || C.type == CDDP_B200_CON_STATE_BOX) {  // constraint.hpp:144-217
      const bool ctrl = C.type == CDDP_B200_CON_CONTROL_BOX;
      const int k = ctrl ? m : n;
      for (int j = 0; j < k; ++j) {
        for (int half = 0; half < 2; ++half) {
          const int rr = row + half * k + j;
          rt[rr] = ctrl ? IP_ROW_CONTROL : IP_ROW_STATE;
          sc[rr] = C.scale;
          const double sgn = half ? 1.0 : -1.0;
          if (ctrl) Gu[(size_t)rr * m + j] = sgn * C.scale;
          else Gx[(size_t)rr * n + j] = sgn * C.scale;
          off[rr] = half ? C.p1[j] * C.scale : -C.p0[j] * C.scale;  // ip_upper_bound_ (:156-159)
        }
      }
      row += 2 * k;
    } else if (C.type == CDDP_B200_CON_BALL) {  // constraint.hpp:320-373
      rt[row] = IP_ROW_BALL;
      bd[row] = C.rows;
      sc[row] = C.scale;
      off[row] = C.p1[0];
      for (int j = 0; j < C.rows; ++j) Gx[(size_t)row * n + j] = C.p0[j];
      row += 1;
    } else {  // LinearConstraint (:253-284)
      for (int q = 0; q < C.rows; ++q) {
        rt[row + q] = IP_ROW_STATE;
        sc[row + q] = C.scale;
        off[row + q] = C.p1[q];
        for (int j = 0; j < n; ++j) Gx[(size_t)(row + q) * n + j] = C.p0[(size_t)q * n + j];
      }
      row += C.rows;
    }
  }
  int *drt = nullptr, *dbd = nullptr;
  double *dGx = nullptr, *dGu = nullptr, *doff = nullptr, *dsc = nullptr;
  IpDevice &ip = s->ip;
  const size_t B = batch, N = p->horizon, Dd = Dn;
#define AL(ptr, cnt)                     \
  do {                                   \
    if ((r = s->alloc(&(ptr), (cnt)))) { \
      cddp_b200_destroy(s);              \
      return r;                          \
    }                                    \
  } while (0)
  AL(drt, Dd); AL(dbd, Dd); AL(dGx, Dd * n); AL(dGu, Dd * m); AL(doff, Dd); AL(dsc, Dd);
  for (int q = 0; q < 2; ++q) { AL(ip.Y[q], B * N * Dd); AL(ip.S[q], B * N * Dd); AL(ip.G[q], B * N * Dd); }
  AL(ip.ky, B * N * Dd); AL(ip.ks, B * N * Dd); AL(ip.Ky, B * N * Dd * n); AL(ip.Ks, B * N * Dd * n);
  AL(ip.mu, B); AL(ip.merit, B); AL(ip.logsum, B); AL(ip.filter_theta, B); AL(ip.inf_pr, B); AL(ip.inf_comp, B);
  AL(ip.step_norm, B); AL(ip.alpha_du, B); AL(ip.apm, B); AL(ip.adm, B);
  AL(ip.filter, B * IP_FILTER_CAP * 2); AL(ip.filter_size, B); AL(ip.ls_stats, B * CDDP_B200_MAX_ALPHAS * 4);
  AL(ip.lamT, B * n); AL(ip.dlamT, B * n); AL(ip.lamh, B);
  if (io->terminal_equality) {  // scratch of the one-sweep / (n+1)-column terminal-equality solve (ipddp_teq.cu)
    AL(ip.kvar, B * (n + 1) * N * m); AL(ip.pvar, B * (n + 1) * (N + 1) * n); AL(ip.rvar, B * N * m);
    if (teq_small_case(n, m, (int)Dd)) { AL(ip.stage, B * N * teq_stage_stride(n, m)); AL(ip.teq_ran, B); }
  }
#undef AL
  cudaError_t e = cudaMemcpy(drt, rt.data(), Dd * sizeof(int), cudaMemcpyHostToDevice);
  if (e == cudaSuccess) e = cudaMemcpy(dbd, bd.data(), Dd * sizeof(int), cudaMemcpyHostToDevice);
  if (e == cudaSuccess) e = cudaMemcpy(dGx, Gx.data(), Gx.size() * sizeof(double), cudaMemcpyHostToDevice);
  if (e == cudaSuccess) e = cudaMemcpy(dGu, Gu.data(), Gu.size() * sizeof(double), cudaMemcpyHostToDevice);
  if (e == cudaSuccess) e = cudaMemcpy(doff, off.data(), Dd * sizeof(double), cudaMemcpyHostToDevice);
  if (e == cudaSuccess) e = cudaMemcpy(dsc, sc.data(), Dd * sizeof(double), cudaMemcpyHostToDevice);
  if (e == cudaSuccess) e = cudaMemset(ip.filter_size, 0, B * sizeof(int));
  if (e == cudaSuccess) e = cudaMemset(ip.lamT, 0, B * n * sizeof(double));
  if (e == cudaSuccess) e = cudaMemset(ip.dlamT, 0, B * n * sizeof(double));
  if (e == cudaSuccess) e = cudaMemset(ip.lamh, 0, B * sizeof(double));
  if (e == cudaSuccess) e = cudaMemset(ip.ls_stats, 0, B * CDDP_B200_MAX_ALPHAS * 4 * sizeof(double));
  if (e == cudaSuccess) e = cudaMemset(ip.ky, 0, B * N * Dd * sizeof(double));
  if (e == cudaSuccess) e = cudaMemset(ip.ks, 0, B * N * Dd * sizeof(double));
  if (e == cudaSuccess) e = cudaMemset(ip.Ky, 0, B * N * Dd * n * sizeof(double));
  if (e == cudaSuccess) e = cudaMemset(ip.Ks, 0, B * N * Dd * n * sizeof(double));
  if (e != cudaSuccess) { cddp_b200_destroy(s); return cuda_fail(e, "ipddp setup"); }
  s->ic.d = D; s->ic.nc = nc; s->ic.teq = io->terminal_equality ? 1 : 0;
  s->ic.row_type = drt; s->ic.row_bdim = dbd; s->ic.Gx = dGx; s->ic.Gu = dGu; s->ic.off = doff; s->ic.scale = dsc;
  s->ic.io = *io;
  if ((r = apply_layout(s, RECORDS_DENSE))) { cddp_b200_destroy(s); return r; }
  *out = s;
  return 0;
}

int cddp_b200_ipddp_dual_dim(cddp_b200_solver *s, int *d) {
  if (!s || !d) return CDDP_B200_ERR_INVALID_ARGUMENT;
  *d = s->kind == 1 ? s->ic.d : 0;
  return 0;
}

namespace {
// a [B][N][d] double-buffered array of the CURRENT nominal (cur[b] selects the buffer): gathered on the device into the
// scratch buffer, then one device-to-host copy
int download_current(cddp_b200_solver *s, double *dst, double *const buf[2], size_t per_instance) {
  if (!dst) return 0;
  const size_t bytes = (size_t)s->d.B * per_instance * sizeof(double);
  int r = s->ensure_scratch(bytes);
  if (r) return r;
  CU(launch_gather_by_cur(s->d, buf[0], buf[1], per_instance, s->scratch, s->stream));
  s->timing.other_launches++;
  CU(cudaMemcpyAsync(dst, s->scratch, bytes, cudaMemcpyDeviceToHost, s->stream));
  CU(cudaStreamSynchronize(s->stream));  // the scratch buffer is reused by the next call
  return 0;
}
}  // namespace

int cddp_b200_ipddp_get_solution(cddp_b200_solver *s, double *Y, double *S, double *G, double *scalars) {
  if (!s || s->kind != 1) return CDDP_B200_ERR_INVALID_ARGUMENT;
  DeviceGuard g(s->device);
  const size_t B = s->d.B, per = (size_t)s->d.N * s->ic.d;
  int r;
  if (per) {
    if ((r = download_current(s, Y, s->ip.Y, per))) return r;
    if ((r = download_current(s, S, s->ip.S, per))) return r;
    if ((r = download_current(s, G, s->ip.G, per))) return r;
  }
  if (scalars) {
    std::vector<double> tmp(8 * B);
    double *src[8] = {s->ip.mu, s->ip.merit, s->ip.inf_pr, s->ip.inf_comp, s->ip.step_norm, s->ip.alpha_du, s->ip.apm, s->ip.adm};
    for (int k = 0; k < 8; ++k)
      if ((r = download(s, tmp.data() + k * B, src[k], B * sizeof(double)))) return r;
    CU(cudaStreamSynchronize(s->stream));
    for (size_t b = 0; b < B; ++b)
      for (int k = 0; k < 8; ++k) scalars[b * 8 + k] = tmp[k * B + b];
  }
  CU(cudaStreamSynchronize(s->stream));
  return 0;
}

int cddp_b200_ipddp_get_iteration_state(cddp_b200_solver *s, double *lamT, double *filter, int *filter_size, double *scalars) {
  if (!s || s->kind != 1) return CDDP_B200_ERR_INVALID_ARGUMENT;
  DeviceGuard g(s->device);
  const size_t B = s->d.B, n = s->d.n;
  int r;
  if (lamT) {
    if (s->ic.teq) {
      if ((r = download(s, lamT, s->ip.lamT, B * n * sizeof(double)))) return r;
    } else {
      std::memset(lamT, 0, B * n * sizeof(double));
    }
  }
  if ((r = download(s, filter, s->ip.filter, B * IP_FILTER_CAP * 2 * sizeof(double)))) return r;
  if ((r = download(s, filter_size, s->ip.filter_size, B * sizeof(int)))) return r;
  if (scalars) {
    std::vector<double> tmp(3 * B);
    double *src[3] = {s->ip.filter_theta, s->ip.logsum, s->ip.lamh};
    for (int k = 0; k < 3; ++k) {
      if (!src[k]) {
        std::fill(tmp.begin() + k * B, tmp.begin() + (k + 1) * B, 0.0);
        continue;
      }
      if ((r = download(s, tmp.data() + k * B, src[k], B * sizeof(double)))) return r;
    }
    CU(cudaStreamSynchronize(s->stream));
    for (size_t b = 0; b < B; ++b)
      for (int k = 0; k < 3; ++k) scalars[b * 3 + k] = tmp[k * B + b];
  }
  CU(cudaStreamSynchronize(s->stream));
  return 0;
}

int cddp_b200_ipddp_get_gains(cddp_b200_solver *s, double *ky, double *Ky, double *ks, double *Ks) {
  if (!s || s->kind != 1) return CDDP_B200_ERR_INVALID_ARGUMENT;
  DeviceGuard g(s->device);
  const size_t cnt = (size_t)s->d.B * s->d.N * s->ic.d;
  int r;
  if (cnt) {
    if ((r = download(s, ky, s->ip.ky, cnt * sizeof(double)))) return r;
    if ((r = download(s, ks, s->ip.ks, cnt * sizeof(double)))) return r;
    if ((r = download(s, Ky, s->ip.Ky, cnt * s->d.n * sizeof(double)))) return r;
    if ((r = download(s, Ks, s->ip.Ks, cnt * s->d.n * sizeof(double)))) return r;
  }
  CU(cudaStreamSynchronize(s->stream));
  return 0;
}

int cddp_b200_ipddp_get_line_search(cddp_b200_solver *s, double *table) {
  if (!s || s->kind != 1 || !table) return CDDP_B200_ERR_INVALID_ARGUMENT;
  DeviceGuard g(s->device);
  const size_t B = s->d.B;
  std::vector<double> tmp(B * CDDP_B200_MAX_ALPHAS * 4);
  int r;
  if ((r = download(s, tmp.data(), s->ip.ls_stats, tmp.size() * sizeof(double)))) return r;
  CU(cudaStreamSynchronize(s->stream));
  const int na = s->c.num_alphas;
  for (size_t b = 0; b < B; ++b)
    for (int a = 0; a < na; ++a)
      for (int k = 0; k < 4; ++k) table[(b * na + a) * 4 + k] = tmp[(b * CDDP_B200_MAX_ALPHAS + a) * 4 + k];
  return 0;
}

int cddp_b200_ipddp_get_history(cddp_b200_solver *s, double *history, int *lens) {
  if (!s || s->kind != 1) return CDDP_B200_ERR_INVALID_ARGUMENT;
  if (!s->d.history) return CDDP_B200_ERR_STATE;
  DeviceGuard g(s->device);
  int r;
  if ((r = download(s, history, s->d.history, (size_t)s->d.B * s->d.history_cap * IP_HISTORY_COLS * sizeof(double)))) return r;
  if ((r = download(s, lens, s->d.history_len, (size_t)s->d.B * sizeof(int)))) return r;
  CU(cudaStreamSynchronize(s->stream));
  return 0;
}

int cddp_b200_reset_timing(cddp_b200_solver *s) {
  if (!s) return CDDP_B200_ERR_INVALID_ARGUMENT;
  s->timing = cddp_b200_timing{};
  return 0;
}

int cddp_b200_get_timing(cddp_b200_solver *s, cddp_b200_timing *t) {
  if (!s || !t) return CDDP_B200_ERR_INVALID_ARGUMENT;
  *t = s->timing;
  return 0;
}

int cddp_b200_enable_timing(cddp_b200_solver *s, int enable) {
  if (!s) return CDDP_B200_ERR_INVALID_ARGUMENT;
  s->timing_enabled = enable != 0;
  return 0;
}

int cddp_b200_backward_algorithmic_bytes(cddp_b200_solver *s, double *bytes) {
  if (!s || !bytes) return CDDP_B200_ERR_INVALID_ARGUMENT;
  const double n = s->d.n, m = s->d.m;
  *bytes = 8.0 * (n * n + 2 * n * m + n + 3 * m) * (double)s->d.N * (double)s->d.B;
  return 0;
}

int cddp_b200_solve_host(const cddp_b200_problem *problem, const cddp_b200_options *opts, int batch, int device,
                         const double *x0, const double *xref, const double *ref_traj, double *X, double *U, double *K,
                         double *final_objective, int *iterations_completed, int *status, double *final_step_length,
                         double *final_regularization, double *inf_du) {
  cddp_b200_solver *s = nullptr;
  int r = cddp_b200_create(problem, opts, batch, device, &s);
  if (r) return r;
  r = cddp_b200_set_instances(s, x0, xref, ref_traj, X, U);
  if (!r) r = cddp_b200_solve(s);
  if (!r)
    r = cddp_b200_get_solution(s, X, U, K, final_objective, iterations_completed, status, final_step_length,
                               final_regularization, inf_du);
  cddp_b200_destroy(s);
  return r;
}

}  // extern "C"
