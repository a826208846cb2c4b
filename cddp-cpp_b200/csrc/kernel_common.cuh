// Small device helpers shared by the model-dependent kernels (kernels_*.cuh).  NVRTC-clean: no standard headers.
#pragma once
#include "engine.h"

namespace cddp_b200 {
namespace kern {

__device__ __forceinline__ double pos_inf() { return __longlong_as_double(0x7ff0000000000000LL); }

__device__ __forceinline__ const double *ref_ptr(const DeviceState &d, int b, int t) {
  // objective.cpp:84-88: per-index reference if a reference trajectory was given
  return d.ref_traj ? d.ref_traj + ((size_t)b * (d.N + 1) + t) * d.n : d.xref + (size_t)b * d.n;
}

}  // namespace kern
}  // namespace cddp_b200
