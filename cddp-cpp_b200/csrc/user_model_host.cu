// Host side of the user-model plugin: compiles the model-dependent kernels (kernels_linearize.cuh, kernels_forward.cuh,
// kernels_ipddp.cuh — the very same kernel text the built-in models are compiled from) together with a user-supplied
// CUDA source for sm_100a with NVRTC, loads the cubin with the driver API and launches it in place of the built-in
// instantiations.  This is the device counterpart of subclassing cddp::DynamicalSystem
// (include/cddp-cpp/cddp_core/dynamical_system.hpp:33-152): see user_model.cuh for the source contract.
//
// libnvrtc and libcuda are dlopen'ed on first use, so that libcddp_b200.so itself loads (and every entry point that
// does not need them works) on machines without a driver, and compile-only checks run without a GPU.
#include <cuda.h>
#include <dlfcn.h>
#include <nvrtc.h>

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <mutex>
#include <string>
#include <vector>

#include "engine.h"
#include "user_model_host.h"

namespace cddp_b200 {

namespace {

#include "build/embedded_sources.inc"  // struct EmbeddedSource { const char *name, *text; } kEmbedded[]; kNumEmbedded

struct Nvrtc {
  void *h = nullptr;
  decltype(&nvrtcCreateProgram) CreateProgram = nullptr;
  decltype(&nvrtcDestroyProgram) DestroyProgram = nullptr;
  decltype(&nvrtcAddNameExpression) AddNameExpression = nullptr;
  decltype(&nvrtcCompileProgram) CompileProgram = nullptr;
  decltype(&nvrtcGetProgramLogSize) GetProgramLogSize = nullptr;
  decltype(&nvrtcGetProgramLog) GetProgramLog = nullptr;
  decltype(&nvrtcGetCUBINSize) GetCUBINSize = nullptr;
  decltype(&nvrtcGetCUBIN) GetCUBIN = nullptr;
  decltype(&nvrtcGetLoweredName) GetLoweredName = nullptr;
  decltype(&nvrtcGetErrorString) GetErrorString = nullptr;
};
struct Driver {
  void *h = nullptr;
  decltype(&cuModuleLoadData) ModuleLoadData = nullptr;
  decltype(&cuModuleUnload) ModuleUnload = nullptr;
  decltype(&cuModuleGetFunction) ModuleGetFunction = nullptr;
  decltype(&cuLaunchKernel) LaunchKernel = nullptr;
  decltype(&cuFuncSetAttribute) FuncSetAttribute = nullptr;
  decltype(&cuGetErrorString) GetErrorString = nullptr;
};

template <class F>
bool sym(void *h, const char *name, F &out) {
  out = reinterpret_cast<F>(dlsym(h, name));
  return out != nullptr;
}

Nvrtc *nvrtc(std::string &err) {
  static Nvrtc lib;
  static std::once_flag once;
  static std::string load_err;
  std::call_once(once, [] {
    const char *names[] = {"libnvrtc.so.12", "/usr/local/cuda/lib64/libnvrtc.so.12", "libnvrtc.so", "/usr/local/cuda/lib64/libnvrtc.so"};
    for (const char *n : names)
      if ((lib.h = dlopen(n, RTLD_NOW | RTLD_LOCAL))) break;
    if (!lib.h) {
      load_err = "cannot load libnvrtc (needed to compile a user-supplied dynamics model)";
      return;
    }
    bool ok = sym(lib.h, "nvrtcCreateProgram", lib.CreateProgram) & sym(lib.h, "nvrtcDestroyProgram", lib.DestroyProgram) &
              sym(lib.h, "nvrtcAddNameExpression", lib.AddNameExpression) & sym(lib.h, "nvrtcCompileProgram", lib.CompileProgram) &
              sym(lib.h, "nvrtcGetProgramLogSize", lib.GetProgramLogSize) & sym(lib.h, "nvrtcGetProgramLog", lib.GetProgramLog) &
              sym(lib.h, "nvrtcGetCUBINSize", lib.GetCUBINSize) & sym(lib.h, "nvrtcGetCUBIN", lib.GetCUBIN) &
              sym(lib.h, "nvrtcGetLoweredName", lib.GetLoweredName) & sym(lib.h, "nvrtcGetErrorString", lib.GetErrorString);
    if (!ok) {
      load_err = "libnvrtc lacks a required symbol";
      lib.h = nullptr;
    }
  });
  if (!lib.h) {
    err = load_err;
    return nullptr;
  }
  return &lib;
}

Driver *driver(std::string &err) {
  static Driver lib;
  static std::once_flag once;
  static std::string load_err;
  std::call_once(once, [] {
    const char *names[] = {"libcuda.so.1", "libcuda.so"};
    for (const char *n : names)
      if ((lib.h = dlopen(n, RTLD_NOW | RTLD_LOCAL))) break;
    if (!lib.h) {
      load_err = "cannot load libcuda (no NVIDIA driver): a user-supplied model cannot be loaded; there is no CPU fallback";
      return;
    }
    bool ok = sym(lib.h, "cuModuleLoadData", lib.ModuleLoadData) & sym(lib.h, "cuModuleUnload", lib.ModuleUnload) &
              sym(lib.h, "cuModuleGetFunction", lib.ModuleGetFunction) & sym(lib.h, "cuLaunchKernel", lib.LaunchKernel) &
              sym(lib.h, "cuFuncSetAttribute", lib.FuncSetAttribute) & sym(lib.h, "cuGetErrorString", lib.GetErrorString);
    if (!ok) {
      load_err = "libcuda lacks a required symbol";
      lib.h = nullptr;
    }
  });
  if (!lib.h) {
    err = load_err;
    return nullptr;
  }
  return &lib;
}

enum { K_LIN = 0, K_FWD16, K_FWD32, K_FWD1, K_IPINIT, K_IPFWD, K_COUNT };

std::string name_expr(int k, bool diag) {
  const char *d = diag ? "true" : "false";
  switch (k) {
    case K_LIN: return "cddp_b200::kern::linearize_kernel<CDDP_B200_MODEL_USER, cddp_b200::DensePattern>";
    case K_FWD16: return std::string("cddp_b200::kern::forward_kernel<CDDP_B200_MODEL_USER, 16, ") + d + ">";
    case K_FWD32: return std::string("cddp_b200::kern::forward_kernel<CDDP_B200_MODEL_USER, 32, ") + d + ">";
    case K_FWD1: return std::string("cddp_b200::kern::forward_first_kernel<CDDP_B200_MODEL_USER, ") + d + ">";
    case K_IPINIT: return "cddp_b200::kern::ip_initialize_kernel<CDDP_B200_MODEL_USER>";
    default: return "cddp_b200::kern::ip_forward_kernel<CDDP_B200_MODEL_USER, 0>";
  }
}

}  // namespace

struct UserKernels {
  CUmodule mod = nullptr;
  CUfunction fn[K_COUNT] = {};
  int n = 0, m = 0;
  bool diag = false;
};

// Compiles the translation unit; on success `cubin` holds the sm_100a image and `lowered` the mangled kernel names.
static int compile_user_model_uncached(const char *source, int n, int m, bool diag, std::vector<char> &cubin,
                                       std::string lowered[K_COUNT], std::string &log);

// Process-wide cache of compiled images keyed by (n, m, cost form, source text): a solver handle per MPC loop / per stream /
// per solver type of the same model then pays the NVRTC compile (4 .. 19 s) once.
namespace {
struct CompiledImage {
  std::vector<char> cubin;
  std::string lowered[K_COUNT];
};
std::mutex g_cache_mutex;
std::map<std::string, CompiledImage> g_image_cache;
}  // namespace

static int compile_user_model(const char *source, int n, int m, bool diag, std::vector<char> &cubin, std::string lowered[K_COUNT],
                              std::string &log) {
  const std::string key = std::to_string(n) + "/" + std::to_string(m) + "/" + (diag ? "d" : "f") + "/" + source;
  {
    std::lock_guard<std::mutex> lk(g_cache_mutex);
    auto it = g_image_cache.find(key);
    if (it != g_image_cache.end()) {
      cubin = it->second.cubin;
      for (int k = 0; k < K_COUNT; ++k) lowered[k] = it->second.lowered[k];
      return 0;
    }
  }
  const int r = compile_user_model_uncached(source, n, m, diag, cubin, lowered, log);
  if (r == 0) {
    std::lock_guard<std::mutex> lk(g_cache_mutex);
    CompiledImage &img = g_image_cache[key];
    img.cubin = cubin;
    for (int k = 0; k < K_COUNT; ++k) img.lowered[k] = lowered[k];
  }
  return r;
}

static int compile_user_model_uncached(const char *source, int n, int m, bool diag, std::vector<char> &cubin,
                                       std::string lowered[K_COUNT], std::string &log) {
  std::string err;
  Nvrtc *rt = nvrtc(err);
  if (!rt) {
    log = err;
    return CDDP_B200_ERR_CUDA;
  }
  std::string tu;
  tu += "#define CDDP_USER_NS " + std::to_string(n) + "\n#define CDDP_USER_NC " + std::to_string(m) + "\n";
  tu += "#include \"engine.h\"\n#include \"user_model.cuh\"\n";
  tu += "#line 1 \"user_model_source.cu\"\n";
  tu += source;
  tu += "\n#include \"kernels_linearize.cuh\"\n#include \"kernels_forward.cuh\"\n#include \"kernels_ipddp.cuh\"\n";
  std::vector<const char *> hdr_text, hdr_name;
  for (int i = 0; i < kNumEmbedded; ++i) {
    hdr_name.push_back(kEmbedded[i].name);
    hdr_text.push_back(kEmbedded[i].text);
  }
  nvrtcProgram prog = nullptr;
  nvrtcResult r = rt->CreateProgram(&prog, tu.c_str(), "cddp_b200_user_model.cu", (int)hdr_text.size(), hdr_text.data(), hdr_name.data());
  if (r != NVRTC_SUCCESS) {
    log = std::string("nvrtcCreateProgram: ") + rt->GetErrorString(r);
    return CDDP_B200_ERR_CUDA;
  }
  std::string names[K_COUNT];
  for (int k = 0; k < K_COUNT; ++k) {
    names[k] = name_expr(k, diag);
    rt->AddNameExpression(prog, names[k].c_str());
  }
  const char *opts[] = {"--gpu-architecture=sm_100a", "-std=c++17", "-default-device", "-lineinfo"};
  r = rt->CompileProgram(prog, (int)(sizeof(opts) / sizeof(opts[0])), opts);
  size_t log_size = 0;
  rt->GetProgramLogSize(prog, &log_size);
  if (log_size > 1) {
    std::vector<char> buf(log_size);
    rt->GetProgramLog(prog, buf.data());
    log.assign(buf.data());
  }
  if (r != NVRTC_SUCCESS) {
    if (log.empty()) log = rt->GetErrorString(r);
    rt->DestroyProgram(&prog);
    return CDDP_B200_ERR_USER_MODEL;
  }
  size_t sz = 0;
  rt->GetCUBINSize(prog, &sz);
  cubin.resize(sz);
  rt->GetCUBIN(prog, cubin.data());
  if (const char *dump = getenv("CDDP_B200_DUMP_CUBIN")) {  // developer aid: cuobjdump -res-usage / -sass on the JIT result
    if (FILE *f = fopen(dump, "wb")) {
      fwrite(cubin.data(), 1, cubin.size(), f);
      fclose(f);
    }
  }
  for (int k = 0; k < K_COUNT; ++k) {
    const char *low = nullptr;
    if (rt->GetLoweredName(prog, names[k].c_str(), &low) != NVRTC_SUCCESS || !low) {
      log += "\ncannot resolve kernel " + names[k];
      rt->DestroyProgram(&prog);
      return CDDP_B200_ERR_USER_MODEL;
    }
    lowered[k] = low;
  }
  rt->DestroyProgram(&prog);
  return 0;
}

int user_model_compile_only(const char *source, int n, int m, std::string &log, size_t *cubin_bytes) {
  std::vector<char> cubin;
  std::string lowered[K_COUNT];
  int r = compile_user_model(source, n, m, true, cubin, lowered, log);
  if (cubin_bytes) *cubin_bytes = cubin.size();
  return r;
}

int user_model_build(const char *source, int n, int m, bool diag, UserKernels **out, std::string &log) {
  *out = nullptr;
  std::vector<char> cubin;
  std::string lowered[K_COUNT];
  int r = compile_user_model(source, n, m, diag, cubin, lowered, log);
  if (r) return r;
  std::string err;
  Driver *dr = driver(err);
  if (!dr) {
    log = err;
    return CDDP_B200_ERR_CUDA;
  }
  cudaFree(nullptr);  // make sure the runtime's primary context exists and is current
  UserKernels *uk = new UserKernels();
  uk->n = n; uk->m = m; uk->diag = diag;
  CUresult cr = dr->ModuleLoadData(&uk->mod, cubin.data());
  if (cr != CUDA_SUCCESS) {
    const char *s = nullptr;
    dr->GetErrorString(cr, &s);
    log = std::string("cuModuleLoadData: ") + (s ? s : "?");
    delete uk;
    return CDDP_B200_ERR_CUDA;
  }
  for (int k = 0; k < K_COUNT; ++k) {
    cr = dr->ModuleGetFunction(&uk->fn[k], uk->mod, lowered[k].c_str());
    if (cr != CUDA_SUCCESS) {
      log = "cuModuleGetFunction failed for " + lowered[k];
      dr->ModuleUnload(uk->mod);
      delete uk;
      return CDDP_B200_ERR_CUDA;
    }
    dr->FuncSetAttribute(uk->fn[k], CU_FUNC_ATTRIBUTE_MAX_DYNAMIC_SHARED_SIZE_BYTES, 200 * 1024);
  }
  *out = uk;
  return 0;
}

void user_model_destroy(UserKernels *uk) {
  if (!uk) return;
  std::string err;
  Driver *dr = driver(err);
  if (dr && uk->mod) dr->ModuleUnload(uk->mod);
  delete uk;
}

static cudaError_t launch(UserKernels *uk, int k, unsigned blocks, unsigned threads, size_t shm, cudaStream_t st, void **params) {
  std::string err;
  Driver *dr = driver(err);
  if (!dr || !uk) return cudaErrorInvalidValue;
  const CUresult cr = dr->LaunchKernel(uk->fn[k], blocks, 1, 1, threads, 1, 1, (unsigned)shm, (CUstream)st, params, nullptr);
  return cr == CUDA_SUCCESS ? cudaSuccess : cudaErrorLaunchFailure;
}

cudaError_t launch_user_linearize(const Constants &c, const DeviceState &d, bool force, cudaStream_t st) {
  UserKernels *uk = static_cast<UserKernels *>(d.user);
  const int n = d.n, m = d.m;
  const int stride = (n * n + n * m + n + 2 * m + 1) & ~1;  // RecordLayout<n, m, DensePattern>::stride
  const int PSH = lin_part_width(stride) | 1;
  const size_t per_warp = sizeof(double) * 32 * PSH;
  int wpc = 4;
  const long long warps = (long long)d.n_slots * ((d.N + 1 + 31) / 32);
  const unsigned blocks = (unsigned)((warps + wpc - 1) / wpc);
  int f = force ? 1 : 0;
  void *params[] = {(void *)&c, (void *)&d, &f, &wpc};
  return launch(uk, K_LIN, blocks, 128, per_warp * wpc, st, params);
}

cudaError_t launch_user_forward(const Constants &c, const DeviceState &d, int mode, cudaStream_t st) {
  UserKernels *uk = static_cast<UserKernels *>(d.user);
  if (!uk || (c.cost_diag != 0) != uk->diag) return cudaErrorInvalidValue;
  const int wpc = 2;  // kern::kWarpsPerCta
  void *params[] = {(void *)&c, (void *)&d, &mode};
  // The speculative alphas_[0] pass (kernels_forward.cuh) saves THROUGHPUT: one rollout instead of 16 for every instance
  // it settles.  Its price is LATENCY: the full-width launch that follows starts after it and costs a whole rollout
  // however few instances are left.  It pays only while the full line search is throughput-bound, i.e. while the 16-wide
  // rollout of the work list oversubscribes the SMs (measured with the 7-DOF manipulator: 9.1 -> 3.4 ms at 1024 instances
  // without the pass, 12.6 -> 11.4 ms at 8192 with it); below that the full launch runs alone.
  const bool speculate = c.speculate >= 0 ? c.speculate != 0
                                          : d.n_slots >= 8192;  // 16 lanes x 8192 = 4096 warps = 27 per SM: past the point where one more rollout of latency is cheaper
  if (mode == FW_ITERATE && !c.opt.enable_parallel && speculate) {
    cudaError_t e = launch(uk, K_FWD1, (unsigned)((d.n_slots + 63) / 64), 64, 0, st, params);
    if (e != cudaSuccess) return e;
  } else if (mode == FW_ITERATE) {
    // options can change on a live handle (cddp_b200_set_options): flags left by an earlier speculative pass must not
    // mask instances from the full line search
    cudaError_t e = cudaMemsetAsync(d.fw_done, 0, (size_t)d.B * sizeof(int), st);
    if (e != cudaSuccess) return e;
  }
  if (c.num_alphas <= 16) return launch(uk, K_FWD16, (unsigned)((d.n_slots + wpc * 2 - 1) / (wpc * 2)), wpc * 32, 0, st, params);
  return launch(uk, K_FWD32, (unsigned)((d.n_slots + wpc - 1) / wpc), wpc * 32, 0, st, params);
}

cudaError_t launch_user_ip_initialize(const Constants &c, const DeviceState &d, const IpConstants &ic, const IpDevice &ip,
                                      cudaStream_t st) {
  UserKernels *uk = static_cast<UserKernels *>(d.user);
  void *params[] = {(void *)&c, (void *)&d, (void *)&ic, (void *)&ip};
  return launch(uk, K_IPINIT, (unsigned)((d.B + 63) / 64), 64, 0, st, params);
}

cudaError_t launch_user_ip_forward(const Constants &c, const DeviceState &d, const IpConstants &ic, const IpDevice &ip, int mode,
                                   cudaStream_t st) {
  UserKernels *uk = static_cast<UserKernels *>(d.user);
  const int per_cta = 64 / 16;  // kern::kFwThreads / LG
  const int step = ip_fw_step_doubles(d.n, d.m, ic.d), table = con_table_doubles(d.n, d.m, ic.d);
  const size_t shm = sizeof(double) * ((size_t)ip_fw_cost_doubles(d.n, d.m) + table + (size_t)per_cta * 2 * step);  // kern::ip_fw_smem_doubles
  if (shm > 200 * 1024) return cudaErrorInvalidValue;
  void *params[] = {(void *)&c, (void *)&d, (void *)&ic, (void *)&ip, &mode};
  return launch(uk, K_IPFWD, (unsigned)((d.n_slots + per_cta - 1) / per_cta), 64, shm, st, params);
}

}  // namespace cddp_b200
