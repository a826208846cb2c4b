// Compile-time loop without <type_traits> (this header is also compiled by NVRTC, which ships no C++ standard library).
#pragma once
namespace cddp_b200 {
template <int V>
struct IntC {
  static constexpr int value = V;
};
template <int B_, int E_, class F>
__device__ __forceinline__ void static_for(F &&f) {
  if constexpr (B_ < E_) {
    f(IntC<B_>{});
    static_for<B_ + 1, E_>(f);
  }
}
}  // namespace cddp_b200
