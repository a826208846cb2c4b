// linearize_kernel: for every (instance, t) build the backward-sweep record
//   A = I + dt*Fx, B = dt*Fu                       (clddp_solver.cpp:113-118)
//   lx = 2 Q_ (x - ref_t), lu = 2 R_ u             (objective.cpp:100-120; Q_=Q*dt, R_=R*dt)
// and the terminal gradient V_x(N) = 2 Qf (x_N - ref) (objective.cpp:122-126).
// Header-only so that the SAME kernel is compiled ahead of time for the built-in models (linearize.cu) and at run time
// by NVRTC for a user-supplied model (user_model_host.cu).
#pragma once
#include "kernel_common.cuh"
#include "static_for.cuh"

namespace cddp_b200 {
namespace kern {

// One lane per (instance, t); one warp per chunk of 32 consecutive timesteps of ONE instance.  PAT = DensePattern
// (RECORDS_DENSE) or ModelPattern<MODEL> (RECORDS_STRUCTURED): only entries inside the pattern are stored; the
// analytic Jacobians are identically zero outside it.  The 32 records of a chunk are contiguous in HBM, so each
// lane puts its record into a (bank-conflict-free, odd-stride) shared-memory row and the warp then streams the
// chunk out with coalesced stores.  v1 had every thread write its own 832..1936-byte record directly (32 scattered
// 8-byte stores per instruction): 0.34 ms for 341 MB = 1 TB/s.
// The record goes through shared memory in PARTS slices of at most 56 doubles: staging whole records (26.9 KB per warp
// for the quadrotor, 82.7 KB for a dense n = 14 model) left 6 resp. 2 warps per SM, and the kernel is bound by the
// latency of its loads and of the FP64 chain that forms the Jacobians, not by the stores.
// (lin_parts / lin_part_width live in engine.h: the host launchers size the shared memory with them)

template <int MODEL, class PAT>
__global__ void __launch_bounds__(128, 3) linearize_kernel(Constants c, DeviceState d, int force, int warps_per_cta) {
  constexpr int NS = Model<MODEL>::NS, NC = Model<MODEL>::NC;
  using L = RecordLayout<NS, NC, PAT>;
  constexpr int RS = L::stride, PARTS = lin_parts(RS), H = lin_part_width(RS), PSH = H | 1;  // odd row stride
  extern __shared__ double lin_smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp >= warps_per_cta) return;
  const int N = d.N;
  const int chunks = (N + 1 + 31) / 32;
  const long long wid = (long long)blockIdx.x * warps_per_cta + warp;
  if (wid >= (long long)d.n_slots * chunks) return;
  const int b = slot_instance(d, (int)(wid / chunks)), t0 = (int)(wid % chunks) * 32, t = t0 + lane;
  // (the three per-instance words are requested together: with cur read after the early return the warp paid one more
  // dependent memory round trip before its first operand load — half of the kernel's stall samples are these chains)
  const int status = d.status[b], valid = d.lin_valid[b], cur = d.cur[b];
  if (!force && (status != CDDP_B200_STATUS_RUNNING || valid)) return;  // warp-uniform
  double *row = lin_smem + ((size_t)warp * 32 + lane) * PSH;
  const bool rec = t < N;
  double Fx[NS * NS], Fu[NS * NC], lxv[NS], luv[NC], u[NC];
  if (t <= N) {
    const double *xp = d.X[cur] + ((size_t)b * (N + 1) + t) * NS;
    double x[NS];
#pragma unroll
    for (int i = 0; i < NS; ++i) x[i] = xp[i];
    if (t == N) {
      const double *ref = d.xref + (size_t)b * NS;  // terminal cost always uses reference_state_ (objective.cpp:96)
      double e[NS];
#pragma unroll
      for (int i = 0; i < NS; ++i) e[i] = x[i] - ref[i];
      for (int i = 0; i < NS; ++i) {
        double s = 0.0;
#pragma unroll
        for (int j = 0; j < NS; ++j) s += c.Qf2[i * NS + j] * e[j];
        d.vterm[(size_t)b * NS + i] = s;
      }
    } else {
      const double *up = d.U[cur] + ((size_t)b * N + t) * NC;
#pragma unroll
      for (int i = 0; i < NC; ++i) u[i] = up[i];
      Model<MODEL>::jac(c.mp, x, u, Fx, Fu);
      const double *ref = ref_ptr(d, b, t);
      double e[NS];
#pragma unroll
      for (int i = 0; i < NS; ++i) e[i] = x[i] - ref[i];
      if (c.cost_diag) {  // Q, R diagonal (every shipped workload): the products with the exact-zero entries are skipped
#pragma unroll
        for (int i = 0; i < NS; ++i) lxv[i] = 0.0 + c.Qdt2[i * NS + i] * e[i];
#pragma unroll
        for (int i = 0; i < NC; ++i) luv[i] = 0.0 + c.Rdt2[i * NC + i] * u[i];
      } else {
#pragma unroll
        for (int i = 0; i < NS; ++i) {
          double s = 0.0;
#pragma unroll
          for (int j = 0; j < NS; ++j) s += c.Qdt2[i * NS + j] * e[j];
          lxv[i] = s;
        }
#pragma unroll
        for (int i = 0; i < NC; ++i) {
          double s = 0.0;
#pragma unroll
          for (int j = 0; j < NC; ++j) s += c.Rdt2[i * NC + j] * u[j];
          luv[i] = s;
        }
      }
    }
  }
  const int nrec = min(32, N - t0);  // records in this chunk (t < N)
  double *dst = d.rec + ((size_t)b * N + (nrec > 0 ? t0 : 0)) * RS;
  const double *src = lin_smem + (size_t)warp * 32 * PSH;
  static_for<0, PARTS>([&](auto pc) {
    constexpr int part = decltype(pc)::value, lo = part * H;
    if (rec) {  // this lane's entries lo .. lo + H of its record
      static_for<0, NS>([&](auto lc) {
        constexpr int l = decltype(lc)::value;
        static_for<0, NS>([&](auto jc) {
          constexpr int j = decltype(jc)::value;
          if constexpr (PAT::a(l, j)) {
            if constexpr (L::idxA(l, j) / H == part) row[L::idxA(l, j) - lo] = c.dt * Fx[l * NS + j] + (l == j ? 1.0 : 0.0);
          }
        });
        if constexpr (PAT::brow(l)) {
          static_for<0, NC>([&](auto ac) {
            constexpr int a = decltype(ac)::value;
            if constexpr (L::idxB(l, a) / H == part) row[L::idxB(l, a) - lo] = c.dt * Fu[l * NC + a];
          });
        }
        if constexpr ((L::offLx + l) / H == part) row[L::offLx + l - lo] = lxv[l];
      });
      static_for<0, NC>([&](auto ac) {
        constexpr int a = decltype(ac)::value;
        if constexpr ((L::offLu + a) / H == part) row[L::offLu + a - lo] = luv[a];
        if constexpr ((L::offU + a) / H == part) row[L::offU + a - lo] = u[a];
      });
      if constexpr (L::count < RS && L::count / H == part) row[L::count - lo] = 0.0;  // pad
    }
    __syncwarp();
    // lane -> fixed columns k = lane, lane + 32 of the slice, loop over the records: 256-byte runs, no index division
#pragma unroll
    for (int k0 = 0; k0 < H; k0 += 32) {
      const int k = k0 + lane;
      if (k < H && lo + k < RS) {
        for (int r = 0; r < nrec; ++r) dst[(size_t)r * RS + lo + k] = src[r * PSH + k];
      }
    }
    __syncwarp();
  });
}

}  // namespace kern
}  // namespace cddp_b200
