// Host interface of the user-model plugin (user_model_host.cu).
#pragma once
#include <string>

#include "engine.h"

namespace cddp_b200 {

struct UserKernels;  // NVRTC-compiled module + kernel handles of one user model

// compile only (no driver needed): returns 0 / CDDP_B200_ERR_USER_MODEL (log = compiler output) / CDDP_B200_ERR_CUDA
int user_model_compile_only(const char *source, int n, int m, std::string &log, size_t *cubin_bytes);
// compile for (n, m, diagonal-cost variant) and load into the current device's primary context
int user_model_build(const char *source, int n, int m, bool diag, UserKernels **out, std::string &log);
void user_model_destroy(UserKernels *uk);

cudaError_t launch_user_linearize(const Constants &c, const DeviceState &d, bool force, cudaStream_t st);
cudaError_t launch_user_forward(const Constants &c, const DeviceState &d, int mode, cudaStream_t st);
cudaError_t launch_user_ip_initialize(const Constants &c, const DeviceState &d, const IpConstants &ic, const IpDevice &ip, cudaStream_t st);
cudaError_t launch_user_ip_forward(const Constants &c, const DeviceState &d, const IpConstants &ic, const IpDevice &ip, int mode,
                                   cudaStream_t st);

}  // namespace cddp_b200
