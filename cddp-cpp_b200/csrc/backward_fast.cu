// Batched CLDDP backward Riccati sweep, v2: sub-warp-per-trajectory, row-per-lane, warp-specialised.
//
// Reference behaviour followed: CLDDPSolver::backwardPass, src/cddp_core/clddp_solver.cpp:79-204, and the
// regularisation-retry loop of CDDPSolverBase::solve, src/cddp_core/cddp_solver_base.cpp:93-111 (same
// statements as backward.cu, which stays as the runtime-dimension fallback).
//
// Work decomposition (DESIGN.md "Backward sweep kernel"):
//   * One trajectory is owned by a GROUP of G = 4/8/16 lanes (smallest power of two > n).  Lane r < n holds
//     ROW r of V_xx in registers; lane r == n holds V_x^T as an extra row, so that  [V_xx; V_x^T] * [A|B]
//     yields P_A = V_xx A, P_B = V_xx B and the vector parts A^T V_x, B^T V_x in ONE pass of identical
//     instructions.  A warp therefore sweeps 32/G trajectories at once and 13 (+1) of every 16 lanes of
//     each DFMA do useful work (the v1 kernel left 19 of 32 lanes idle in its rank-13 products).
//   * Q_xx = l_xx + P_A^T A and Q_xu = P_A^T B use the symmetry of V_xx: lane r needs COLUMN r of P_A
//     (one shared-memory transpose) and the record entries by broadcast.  With a compile-time sparsity
//     pattern of A = I + dt*Fx (records.cuh) the unrolled loops skip every structural zero.
//   * The control-space subproblem (PD test, BoxQP, inverse of the free block) is scalar work per
//     trajectory.  Doing it redundantly in every lane (v1) costs more FP64 issue slots than the matrix
//     work, so a dedicated QP WARP solves the subproblems of all T trajectories of the CTA, one trajectory
//     per LANE, between two CTA barriers; the other CTAs resident on the SM cover its latency.
//   * Records are staged global->shared by TMA bulk copies (cp.async.bulk + mbarrier), double-buffered
//     one timestep ahead, issued by one lane per trajectory.
#include <cstdint>
#include <cstdlib>
#include <string>
#include <type_traits>

#include "boxqp_quad.cuh"
#include "boxqp_small.cuh"
#include "engine.h"
#include "kernel_common.cuh"
#include "models.cuh"
#include "records.cuh"
#include "static_for.cuh"

namespace cddp_b200 {

namespace {

template <int B_, int E_, class F>
__device__ __forceinline__ void static_for(F &&f) {
  if constexpr (B_ < E_) {
    f(std::integral_constant<int, B_>{});
    static_for<B_ + 1, E_>(f);
  }
}

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
// TMA bulk copy global -> shared, completion signalled on the mbarrier (SASS: UBLKCP)
__device__ __forceinline__ void bulk_g2s(void *dst, const void *src, uint32_t bytes, uint64_t *bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
               "l"(src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
  const uint32_t addr = smem_u32(bar);
  uint32_t done;
  do {
    asm volatile(
        "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(addr), "r"(parity)
        : "memory");
  } while (!done);
}

// Lane-group geometry: a trajectory's n+1 rows (V_xx rows + the V_x^T row) are spread over G lanes with R rows per lane
// (row = lane_in_group + G*slot).  n+1 <= 4: G=4,R=1; n+1 <= 8: G=8,R=1; else G=8,R=2.  Two rows per lane make every
// broadcast operand feed two DFMAs and halve the number of warps needed per trajectory, so that a CTA of 7 matrix
// warps (28 trajectories) + 1 QP warp fills an SM with 256 threads at up to 255 registers each.
constexpr int group_size(int ns) { return ns + 1 <= 4 ? 4 : 8; }
constexpr int rows_per_lane(int ns) { return ns + 1 <= 8 ? 1 : 2; }

enum { CTRL_OK = 1, CTRL_RESTART = 2, CTRL_FAIL = 3 };

// QPQ = false: W matrix warps + ONE QP warp, one trajectory per QP lane, CTA-wide barriers (any n, m).
// QPQ = true (m = 4, 4 trajectories per matrix warp): W = 8 matrix warps + FOUR QP warps, four lanes per trajectory
// (boxqp_quad.cuh).  The CTA is four independent SETS of two matrix warps and one QP warp (8 trajectories) that meet at
// their own named barriers, so a slow subproblem holds up 8 trajectories instead of 28, and the warps re-divide the
// register file with setmaxnreg (sm_90+): 12 warps are launched at 168 registers, the matrix warps grow to kMatrixRegs
// and the QP warps shrink to kQpRegs (8*32*216 + 4*32*72 = 384*168).
template <int NS, int NC, class PAT, int W, bool QPQ = false>
struct SweepCfg {
  using L = RecordLayout<NS, NC, PAT>;
  static constexpr int G = group_size(NS);
  static constexpr int R = rows_per_lane(NS);
  static constexpr int TPW = 32 / G;  // trajectories per warp
  static constexpr int T = W * TPW;   // trajectories per CTA
  static constexpr int QW = QPQ ? W / 2 : 1;  // QP warps
  static constexpr int kMatrixRegs = 216, kQpRegs = 72;
  static constexpr int RS = L::stride;
  // per-trajectory shared-memory block (offsets in doubles).  Every field starts on an even offset (16 bytes): nvcc
  // merges stores/loads of neighbouring doubles into 128-bit accesses and has been seen (compute-sanitizer, pendulum
  // instantiation) to do so across two 1-element fields at an ODD offset, which faults; with even field starts every
  // merged pair is 16-byte aligned.
  static constexpr int ev(int x) { return (x + 1) & ~1; }
  static constexpr int oRec = 0;                          // [2][RS] double-buffered record
  static constexpr int oPA = oRec + 2 * RS;               // [NS+1][NS]  P_A rows (+ row NS = V_x^T A); reused for the V' transpose
  static constexpr int oPB = oPA + ev((NS + 1) * NS);     // [NS+1][NC]
  static constexpr int oQux = oPB + ev((NS + 1) * NC);    // [NC][NS] Q_ux
  static constexpr int oVx = oQux + ev(NC * NS);          // [NS] scratch for the new V_x
  static constexpr int oQuu = oVx + ev(NS);               // [NC][NC] unregularised Q_uu      (matrix lanes -> QP lane)
  static constexpr int oQu = oQuu + ev(NC * NC);          // [NC]
  static constexpr int oU = oQu + ev(NC);                 // [NC] nominal control u_t
  static constexpr int oKprev = oU + ev(NC);              // [NC] BoxQP warm start k_u_[t]
  static constexpr int oKk = oKprev + ev(NC);             // [NC] k                           (QP lane -> matrix lanes)
  static constexpr int oHinv = oKk + ev(NC);              // [NC][NC] Ht = inverse of the free block of Q_uu_reg, clamped rows/cols 0
  static constexpr int oW = oHinv + ev(NC * NC);          // [NC] w = Q_uu k + Q_u
  static constexpr int oNrm = oW + ev(NC);                // sum_t ||V_x||_1 (matrix lane NS -> QP lane, at the end of the sweep)
  static constexpr int oAcc = oNrm + 2;                   // QPQ: dV0, dV1, Qu_err accumulated here (the QP lanes run at 72 registers)
  static constexpr int oCtrl = oAcc + 4;                  // int state
  static constexpr int oBar = oCtrl + 2;                  // 2 x uint64 mbarrier
  static constexpr int oKst = oBar + 2;                   // [2][NC] BoxQP warm start k_u_[t], staged with the record (TMA)
  static constexpr bool kStaged = (NC * 8) % 16 == 0;     // cp.async.bulk moves multiples of 16 bytes
  static constexpr int raw = oKst + (kStaged ? 2 * NC : 0);
  // Stride between the blocks of neighbouring trajectories, in doubles; even, so that the records stay 16-byte aligned
  // for the bulk copies.  Which residue mod 16 (= mod the 32 four-byte banks) is best was MEASURED, not derived — the
  // matrix warps' row / transposed accesses (8 consecutive doubles per trajectory, 4 trajectories per warp), their
  // broadcast reads and the QP warp's accesses (lane q -> trajectory q: 28 addresses one stride apart) pull in different
  // directions.  Headline launch, ncu shared wavefronts / bank conflicts per launch and sweep time (profiles/
  // r02_sweep_stride.txt): residue 0: 192.7 M / 130.1 M, 0.785 ms; 2 and 14: 88.0 M / 25.6 M, 0.562 ms; 4 and 12: 90.3 M /
  // 27.9 M, 0.563 ms; 8: 98.4 M / 36.0 M, 0.564 ms; **6 and 10: 82.3 M / 20.0 M, 0.552 ms**.  Four trajectories per warp use
  // residue 6; eight (G = 4) keep "== 2 (mod 4)".
  static constexpr int ST = (G == 8) ? raw + ((6 - raw % 16) + 16) % 16 : raw + ((2 - raw % 4) + 4) % 4;
  static constexpr int constDoubles = ((NS * NS + NC * NC) + 1) & ~1;
  static constexpr size_t smemBytes = sizeof(double) * (size_t)(constDoubles + T * ST);
  static constexpr int threads = (W + QW) * 32;
};

// set-level barriers of the QPQ layout: ids 1 + 2*set (reduces "any trajectory of the set still running") and 2 + 2*set
__device__ __forceinline__ bool set_barrier_or(int id, int count, bool p) {
  unsigned r;
  asm volatile(
      "{\n\t.reg .pred pi, po;\n\tsetp.ne.u32 pi, %1, 0;\n\tbarrier.cta.red.or.pred po, %2, %3, pi;\n\tselp.u32 %0, 1, 0, po;\n\t}"
      : "=r"(r)
      : "r"(p ? 1u : 0u), "r"(id), "r"(count)
      : "memory");
  return r != 0u;
}
__device__ __forceinline__ void set_barrier(int id, int count) {
  asm volatile("barrier.cta.sync %0, %1;" ::"r"(id), "r"(count) : "memory");
}

// End of a trajectory's sweep (one lane per trajectory): inf_du scaling (clddp_solver.cpp:194-201), early convergence
// (:206-213), history (cddp_solver_base.cpp:116-118), state written back for the line search.
template <int NS>
__device__ __forceinline__ void sweep_epilogue(const Constants &c, const DeviceState &d, int mode, int b, bool ok, double norm_Vx,
                                               int N, double reg, double dV0, double dV1, double Qu_err, int status,
                                               int failures) {
  double inf_du = 0.0;
  if (ok) {  // norm_Vx: accumulated by the lane that holds V_x
    double sf = c.opt.termination_scaling_max_factor;  // (:197-201)
    sf = fmax(sf, norm_Vx / (double)(N * NS)) / sf;
    inf_du = Qu_err / sf;
    d.dV[2 * b] = dV0;
    d.dV[2 * b + 1] = dV1;
    d.inf_du[b] = inf_du;
  }
  d.bw_ok[b] = ok ? 1 : 0;
  d.lin_valid[b] = 1;
  if (mode == BW_ITERATE) {
    d.reg[b] = reg;
    if (ok && inf_du < c.opt.tolerance) {  // checkEarlyConvergence, clddp_solver.cpp:206-213
      status = CDDP_B200_STATUS_OPTIMAL;
      if (d.history) {  // recordIterationHistory, cddp_solver_base.cpp:116-118
        const int hl = d.history_len[b];
        if (hl < d.history_cap) {
          double *h = d.history + ((size_t)b * d.history_cap + hl) * 4;
          h[0] = d.cost[b];
          h[1] = d.alpha[b];
          h[2] = inf_du;
          h[3] = reg;
          d.history_len[b] = hl + 1;
        }
      }
    }
    if (status != CDDP_B200_STATUS_RUNNING) d.status[b] = status;
    trace_backward(d, b, d.iter[b], failures, status == CDDP_B200_STATUS_OPTIMAL ? 0xff : status == CDDP_B200_STATUS_REG_LIMIT ? 0xfe : 0);
  }
}

#include "sweep_inline.cuh"

// FMODEL >= 0 (one-QP-warp layout, a built-in model's structured records): FUSED LINEARISATION (SURVEY 8f N2,
// clddp_solver.cpp:113-118 computes A, B inside the sweep's loop).  The QP warp is idle from barrier 2 to barrier 1 — while
// the matrix warps run phases C and A1, 40 % of the step — and holds no matrix state, so its lane q forms ONE record of
// trajectory q per step there (closed-form Jacobians of timestep lt = N-4, N-5, ... from x, u; the same statements as
// linearize_kernel), four steps ahead of the sweep, and writes it to the trajectory's record buffer; the matrix warps
// stage it through the same bulk copies as before.  The separate linearize launch (0.115 ms of the 1.06 ms iteration)
// is then skipped by the iteration loop; white-box calls (BW_SINGLE) use the records they are given.
template <int NS, int NC, class PAT, int W, int MINB, bool QPQ = false, int FMODEL = -1>
__global__ void __launch_bounds__(SweepCfg<NS, NC, PAT, W, QPQ>::threads, MINB) sweep_kernel(Constants c, DeviceState d, int mode) {
  using Cfg = SweepCfg<NS, NC, PAT, W, QPQ>;
  static_assert(!QPQ || (NC == 4 && Cfg::TPW == 4 && W % 2 == 0), "quad QP: m = 4, sets of two matrix warps");
  using L = typename Cfg::L;
  constexpr int G = Cfg::G, R = Cfg::R, TPW = Cfg::TPW, T = Cfg::T, RS = Cfg::RS, ST = Cfg::ST;
  extern __shared__ __align__(16) double smem[];
  double *sQ = smem;            // l_xx = 2 Q dt
  double *sR = sQ + NS * NS;    // l_uu = 2 R dt
  double *traj0 = smem + Cfg::constDoubles;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int N = d.N;
  for (int i = threadIdx.x; i < NS * NS; i += blockDim.x) sQ[i] = c.Qdt2[i];
  for (int i = threadIdx.x; i < NC * NC; i += blockDim.x) sR[i] = c.Rdt2[i];

  // barriers: CTA-wide (id 0) or, with QPQ, the set's own pair over its 96 threads
  const int set = QPQ ? (warp < W ? warp >> 1 : warp - W) : 0;
  auto barrier_any = [&](bool p) -> bool {
    if constexpr (QPQ) return set_barrier_or(1 + 2 * set, 96, p);
    else return __syncthreads_or(p ? 1 : 0) != 0;
  };
  auto barrier_all = [&]() {
    if constexpr (QPQ) set_barrier(2 + 2 * set, 96);
    else __syncthreads();
  };
  if (warp < W) {
    // =========================================================== matrix warps
    if constexpr (QPQ) asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(Cfg::kMatrixRegs));
    const int hw = lane / G, r = lane % G;  // hw: which of the warp's TPW trajectories, r: lane within the group
    const int q = warp * TPW + hw;
    const int b = slot_instance(d, blockIdx.x * T + q);
    double *S = traj0 + q * ST;
    const volatile double *Sv = S;  // for broadcast reads (see phase A1)
    // rows owned by this lane: row[s] = r + G*s.  row < NS: a row of V_xx; row == NS: V_x^T; row > NS: idle slot
    // (computes a duplicate of row NS-1 / of V_x, never stored)
    int row[R], rr[R];
    bool isrow[R], isvx[R], has[R];
#pragma unroll
    for (int sl = 0; sl < R; ++sl) {
      row[sl] = r + G * sl;
      rr[sl] = row[sl] < NS ? row[sl] : NS - 1;
      isrow[sl] = row[sl] < NS;
      isvx[sl] = row[sl] == NS;
      has[sl] = row[sl] <= NS;
    }
    const bool alive = b < d.B && !(mode == BW_ITERATE && d.status[b] != CDDP_B200_STATUS_RUNNING);
    const int bb = alive ? b : 0;
    uint64_t *bar = reinterpret_cast<uint64_t *>(S + Cfg::oBar);
    if (r == 0) {
      mbar_init(&bar[0], 1);
      mbar_init(&bar[1], 1);
      asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    const double *grec = d.rec + (size_t)bb * N * RS;
    double *gK = d.K + (size_t)bb * N * NC * NS;
    double *gk = d.kff + (size_t)bb * N * NC;
    const double *vterm = d.vterm + (size_t)bb * NS;

    double V[R][NS];  // per slot: a row of V_xx, or V_x^T
    uint32_t par0 = 0u, par1 = 0u;        // mbarrier phase parity of each record buffer
    bool pend0 = false, pend1 = false;    // a bulk copy into the buffer is in flight
    int t = N - 1, buf = 0;
    bool run = alive;

    auto issue = [&](int tt, int x) {
      if (r == 0) {
        // the buffer was last READ through the generic proxy (phases A1 / A2 of the step that used it, ordered before this
        // point by the __syncwarp of the caller); the bulk copy WRITES it through the async proxy
        // (fused linearisation: the record itself was WRITTEN to global memory through the generic proxy by the QP warp, a
        // CTA barrier ago, and is READ by the bulk copy through the async proxy)
        if constexpr (FMODEL >= 0) asm volatile("fence.proxy.async;" ::: "memory");
        else asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        mbar_expect_tx(&bar[x], RS * 8 + (Cfg::kStaged ? NC * 8 : 0));
        bulk_g2s(S + Cfg::oRec + x * RS, grec + (size_t)tt * RS, RS * 8, &bar[x]);
        // the BoxQP warm start of step tt rides on the same mbarrier: as a plain load issued in phase A2 it shared a
        // scoreboard with the first record reads and stalled them for an HBM round trip (ncu: 1.5 k samples on one DFMA)
        if constexpr (Cfg::kStaged) bulk_g2s(S + Cfg::oKst + x * NC, gk + (size_t)tt * NC, NC * 8, &bar[x]);
      }
      if (x) pend1 = true; else pend0 = true;
    };
    auto wait_buf = [&](int x) {
      mbar_wait(&bar[x], x ? par1 : par0);
      if (x) { par1 ^= 1u; pend1 = false; } else { par0 ^= 1u; pend0 = false; }
    };
    double nrm = 0.0;  // slot holding V_x: sum over the sweep of ||V_x||_1 (clddp_solver.cpp:107,:194)
    double kprev = 0.0;  // lanes r < NC: BoxQP warm start of the step about to be processed
    auto init_sweep = [&]() {  // V_xx = 2 Qf, V_x = 2 Qf (x_N - ref)  (clddp_solver.cpp:89-92)
#pragma unroll
      for (int sl = 0; sl < R; ++sl)
#pragma unroll
        for (int j = 0; j < NS; ++j) V[sl][j] = isrow[sl] ? c.Qf2[rr[sl] * NS + j] : vterm[j];
      t = N - 1;
      buf = 0;
      nrm = 0.0;
      if constexpr (!Cfg::kStaged) kprev = (r < NC) ? gk[(size_t)(N - 1) * NC + r] : 0.0;
    };

    __syncthreads();  // mbarriers initialised, sQ/sR loaded
    auto start_staging = [&]() {  // records N-1 (waited here: phase A1 reads it at once) and N-2 (one step ahead)
      issue(N - 1, 0);
      if (N > 1) issue(N - 2, 1);
      wait_buf(0);
    };
    if (run) {
      init_sweep();
      start_staging();
    }
    double qd[R];
#pragma unroll
    for (int sl = 0; sl < R; ++sl) qd[sl] = sQ[rr[sl] * NS + rr[sl]];

    while (true) {
      const bool wrun = __any_sync(0xffffffffu, run);
      double Qxx[R][NS], Qxu[R][NC], Qx[R];
      if (wrun) {
        // ------------------------------------------------------------ phase A1 (critical path): what the QP needs
        // (the record of this step was waited for in the previous step's shadow phase)
        // BROADCAST operand reads go through a volatile pointer on purpose: nvcc would merge neighbouring doubles into
        // LDS.128, and a broadcast LDS.128 costs 2 shared-memory cycles per warp against <=0.5 for LDS.64 (measured,
        // tools/microbench/lds_broadcast.cu) — 2x more per byte, and shared-memory issue is what bounds this kernel.
        const volatile double *rc = S + Cfg::oRec + buf * RS;
        {
          // P_B(row) = V(row) * B ; V_x slot: (B^T V_x)^T
          double PBr[R][NC];
#pragma unroll
          for (int sl = 0; sl < R; ++sl)
#pragma unroll
            for (int a = 0; a < NC; ++a) PBr[sl][a] = 0.0;
          static_for<0, NS>([&](auto lc) {
            constexpr int l = decltype(lc)::value;
            if constexpr (PAT::brow(l)) {
              static_for<0, NC>([&](auto ac) {
                constexpr int a = decltype(ac)::value;
                const double bv = rc[L::idxB(l, a)];
#pragma unroll
                for (int sl = 0; sl < R; ++sl) PBr[sl][a] = fma(V[sl][l], bv, PBr[sl][a]);
              });
            }
          });
#pragma unroll
          for (int sl = 0; sl < R; ++sl)
            if (has[sl]) {
#pragma unroll
              for (int a = 0; a < NC; ++a) S[Cfg::oPB + row[sl] * NC + a] = PBr[sl][a];
            }
        }
        __syncwarp();
        // Q_uu = l_uu + B^T P_B, one entry per lane; Q_u = l_u + B^T V_x; hand-off to the QP lane   (:125,:128)
#pragma unroll
        for (int ke = 0; ke < (NC * NC + G - 1) / G; ++ke) {
          const int e = r + ke * G;
          if (e < NC * NC) {
            const int a = e / NC, bcol = e - a * NC;
            double acc = sR[e];
            static_for<0, NS>([&](auto lc) {
              constexpr int l = decltype(lc)::value;
              if constexpr (PAT::brow(l)) acc = fma(rc[L::idxB(l, 0) + a], Sv[Cfg::oPB + l * NC + bcol], acc);
            });
            if (run) S[Cfg::oQuu + e] = acc;
          }
        }
        if (run && r < NC) {
          S[Cfg::oQu + r] = rc[L::offLu + r] + S[Cfg::oPB + NS * NC + r];
          S[Cfg::oU + r] = rc[L::offU + r];
          S[Cfg::oKprev + r] = Cfg::kStaged ? S[Cfg::oKst + buf * NC + r] : kprev;
        }
      }
      if (!barrier_any(run)) break;  // barrier 1: Q_uu/Q_u visible to the QP warp
      if (wrun) {
        // ------------------------------------------------------------ phase A2 (in the shadow of the QP warp)
        // BoxQP warm start of the NEXT step, k_u_[t-1] (clddp_solver.cpp:149): an HBM-latency load, issued here so
        // that it is off the critical path (phase A1 used to stall on it)
        if constexpr (!Cfg::kStaged)
          if (run && r < NC && t > 0) kprev = gk[(size_t)(t - 1) * NC + r];
        // TMA staging, also in the shadow: record t-1 (other buffer) was issued one step ago -> wait for it now.
        // Record t-2 is issued into THIS step's buffer once phase A2 has finished reading it (end of this block).
        if (run && t > 0) wait_buf(buf ^ 1);
        const volatile double *rc = S + Cfg::oRec + buf * RS;
        {
          // P_A(row) = V(row) * A ; V_x slot: (A^T V_x)^T                                  (:124,:126-127)
          double PAr[R][NS];
#pragma unroll
          for (int sl = 0; sl < R; ++sl)
#pragma unroll
            for (int j = 0; j < NS; ++j) PAr[sl][j] = 0.0;
          static_for<0, NS>([&](auto lc) {
            constexpr int l = decltype(lc)::value;
            static_for<0, NS>([&](auto jc) {
              constexpr int j = decltype(jc)::value;
              if constexpr (PAT::a(l, j)) {
                const double av = rc[L::idxA(l, j)];
#pragma unroll
                for (int sl = 0; sl < R; ++sl) PAr[sl][j] = fma(V[sl][l], av, PAr[sl][j]);
              }
            });
          });
#pragma unroll
          for (int sl = 0; sl < R; ++sl) {
            if (has[sl]) {
#pragma unroll
              for (int j = 0; j < NS; ++j) S[Cfg::oPA + row[sl] * NS + j] = PAr[sl][j];
            }
            double l1 = 0.0;  // ||V_x||_1 of the value function entering this step
#pragma unroll
            for (int j = 0; j < NS; ++j) l1 += fabs(V[sl][j]);
            nrm += isvx[sl] ? l1 : 0.0;
          }
        }
        __syncwarp();
        // Q_xx(row) = l_xx + (P_A^T A)(row);  Q_xu(row) = (P_A^T B)(row)   (V_xx symmetric)
        double col[R][NS];
#pragma unroll
        for (int sl = 0; sl < R; ++sl) {
#pragma unroll
          for (int l = 0; l < NS; ++l) col[sl][l] = S[Cfg::oPA + l * NS + rr[sl]];
          if (c.q_diag) {
#pragma unroll
            for (int j = 0; j < NS; ++j) Qxx[sl][j] = (j == rr[sl]) ? qd[sl] : 0.0;
          } else {
#pragma unroll
            for (int j = 0; j < NS; ++j) Qxx[sl][j] = sQ[rr[sl] * NS + j];
          }
#pragma unroll
          for (int a = 0; a < NC; ++a) Qxu[sl][a] = 0.0;
        }
        static_for<0, NS>([&](auto lc) {
          constexpr int l = decltype(lc)::value;
          static_for<0, NS>([&](auto jc) {
            constexpr int j = decltype(jc)::value;
            if constexpr (PAT::a(l, j)) {
              const double av = rc[L::idxA(l, j)];
#pragma unroll
              for (int sl = 0; sl < R; ++sl) Qxx[sl][j] = fma(col[sl][l], av, Qxx[sl][j]);
            }
          });
          if constexpr (PAT::brow(l)) {
            static_for<0, NC>([&](auto ac) {
              constexpr int a = decltype(ac)::value;
              const double bv = rc[L::idxB(l, a)];
#pragma unroll
              for (int sl = 0; sl < R; ++sl) Qxu[sl][a] = fma(col[sl][l], bv, Qxu[sl][a]);
            });
          }
        });
#pragma unroll
        for (int sl = 0; sl < R; ++sl) {
          Qx[sl] = rc[L::offLx + rr[sl]] + S[Cfg::oPA + NS * NS + rr[sl]];  // Q_x = l_x + A^T V_x
          if (isrow[sl]) {
#pragma unroll
            for (int a = 0; a < NC; ++a) S[Cfg::oQux + a * NS + row[sl]] = Qxu[sl][a];
          }
        }
        __syncwarp();  // every lane of the group is done reading this step's record
        if (run && t > 1) issue(t - 2, buf);
      }
      barrier_all();  // barrier 2: k, Ht, w, state visible to the matrix warps (and Q_ux to the other lanes)
      if (wrun) {
        // ------------------------------------------------------------ phase C (critical path): value update
        const int st = run ? *reinterpret_cast<volatile int *>(S + Cfg::oCtrl) : 0;
        const bool okh = run && st == CTRL_OK;
        // With K = -Ht Q_ux (Ht symmetric, zero in clamped rows/columns):
        //   V_xx' = Q_xx + K^T Q_uu K + Q_ux^T K + K^T Q_ux = Q_xx + Q_ux^T T,   T(:, i) = 2 K(:, i) - Ht Q_uu K(:, i)
        //   V_x'  = Q_x + K^T (Q_uu k + Q_u) + Q_ux^T k      = Q_x + Q_ux^T (k - Ht w),  w = Q_uu k + Q_u      (:188-191)
        // Each slot forms its own column of K and T from three m x m mat-vecs (no shared-memory round trip on the
        // critical path); the only broadcast stream left is Q_ux (m x n), staged during the QP's shadow.
        double Kc[R][NC], z[NC], T[R][NC];
        {
          double Ht[NC * NC], Qu_[NC * NC];
#pragma unroll
          for (int e = 0; e < NC * NC; ++e) {
            Ht[e] = Sv[Cfg::oHinv + e];
            Qu_[e] = Sv[Cfg::oQuu + e];
          }
#pragma unroll
          for (int a = 0; a < NC; ++a) {  // z = k - Ht w
            double sz = 0.0;
#pragma unroll
            for (int bcol = 0; bcol < NC; ++bcol) sz = fma(Ht[a * NC + bcol], Sv[Cfg::oW + bcol], sz);
            z[a] = Sv[Cfg::oKk + a] - sz;
          }
#pragma unroll
          for (int sl = 0; sl < R; ++sl) {
            double QK[NC];
#pragma unroll
            for (int a = 0; a < NC; ++a) {  // K(:, i) = -Ht Q_ux(:, i)   (:142-178)
              double sK = 0.0;
#pragma unroll
              for (int bcol = 0; bcol < NC; ++bcol) sK = fma(Ht[a * NC + bcol], Qxu[sl][bcol], sK);
              Kc[sl][a] = -sK;
            }
#pragma unroll
            for (int a = 0; a < NC; ++a) {  // Q_uu K(:, i)   (unregularised Q_uu)
              double sq = 0.0;
#pragma unroll
              for (int bcol = 0; bcol < NC; ++bcol) sq = fma(Qu_[a * NC + bcol], Kc[sl][bcol], sq);
              QK[a] = sq;
            }
#pragma unroll
            for (int a = 0; a < NC; ++a) {
              double sT = 2.0 * Kc[sl][a];
#pragma unroll
              for (int bcol = 0; bcol < NC; ++bcol) sT = fma(-Ht[a * NC + bcol], QK[bcol], sT);
              T[sl][a] = sT;
            }
          }
        }
        double Vn[R][NS], vxn[R];
#pragma unroll
        for (int sl = 0; sl < R; ++sl) {
          vxn[sl] = Qx[sl];
#pragma unroll
          for (int j = 0; j < NS; ++j) Vn[sl][j] = Qxx[sl][j];
#pragma unroll
          for (int a = 0; a < NC; ++a) vxn[sl] = fma(Qxu[sl][a], z[a], vxn[sl]);
        }
#pragma unroll
        for (int a = 0; a < NC; ++a) {
#pragma unroll
          for (int j = 0; j < NS; ++j) {
            const double qv = Sv[Cfg::oQux + a * NS + j];
#pragma unroll
            for (int sl = 0; sl < R; ++sl) Vn[sl][j] = fma(T[sl][a], qv, Vn[sl][j]);
          }
        }
#pragma unroll
        for (int sl = 0; sl < R; ++sl)
          if (okh && isrow[sl]) {
#pragma unroll
            for (int j = 0; j < NS; ++j) S[Cfg::oPA + row[sl] * NS + j] = Vn[sl][j];
            S[Cfg::oVx + row[sl]] = vxn[sl];
          }
        __syncwarp();
        // V is assigned in EVERY lane — also those whose step failed: they restart from init_sweep or stop and never read
        // it again — so that the old value function is dead once phase A2 has formed P_A.  Assigned under `if (okh)` it
        // stayed live through the second half of A2 and all of C (2*R*NS registers; the kernel spilled at 255).
#pragma unroll
        for (int sl = 0; sl < R; ++sl)
#pragma unroll
          for (int j = 0; j < NS; ++j)  // V_xx = (V' + V'^T)/2 (:192); the V_x slot takes the new V_x^T
            V[sl][j] = isrow[sl] ? 0.5 * (Vn[sl][j] + S[Cfg::oPA + j * NS + rr[sl]]) : Sv[Cfg::oVx + j];
        if (okh) {
          // gains to HBM last: a store keeps its address/data registers busy until the LSU has read them, which
          // stalled the value update when the stores sat in front of it
#pragma unroll
          for (int sl = 0; sl < R; ++sl)
            if (isrow[sl]) {
#pragma unroll
              for (int a = 0; a < NC; ++a) gK[((size_t)t * NC + a) * NS + row[sl]] = Kc[sl][a];  // K_u_[t] (:182)
            }
          if (r < NC) gk[(size_t)t * NC + r] = S[Cfg::oKk + r];  // k_u_[t] (:181)
          --t;
          buf ^= 1;
          if (t < 0) {  // sweep finished: white-box value function at t = 0, and the V_x(0) term of the norm
            run = false;
#pragma unroll
            for (int sl = 0; sl < R; ++sl) {
              double l1 = 0.0;
#pragma unroll
              for (int j = 0; j < NS; ++j) {
                l1 += fabs(V[sl][j]);
                if (isrow[sl]) d.Vxx0[((size_t)b * NS + row[sl]) * NS + j] = V[sl][j];
                if (isvx[sl]) d.Vx0[(size_t)b * NS + j] = V[sl][j];
              }
              if (isvx[sl]) S[Cfg::oNrm] = nrm + l1;
            }
          }
        } else if (run) {
          if (pend0) wait_buf(0);  // drain the speculative prefetch
          if (pend1) wait_buf(1);
          if (st == CTRL_RESTART) {     // backward failure: sweep again at the increased regularisation
            init_sweep();
            start_staging();
          } else {
            run = false;
          }
        }
        __syncwarp();  // all reads of the transpose buffer done before the next step overwrites it
      }
    }
  } else if constexpr (QPQ) {
    // =========================================================== QP warps, four lanes per trajectory (boxqp_quad.cuh)
    asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(Cfg::kQpRegs));
    const int i = lane & 3, qb = lane & 28;
    const int q = set * 8 + (lane >> 2);
    const int b = slot_instance(d, blockIdx.x * T + q);
    const bool alive = b < d.B && !(mode == BW_ITERATE && d.status[b] != CDDP_B200_STATUS_RUNNING);
    double *Sw = traj0 + q * ST;
    const volatile double *S = Sw;
    __syncthreads();
    double reg = alive ? d.reg[b] : 0.0;
    if (alive && mode == BW_ITERATE && i == 0) d.iter[b] += 1;  // ++iter, cddp_solver_base.cpp:75
    if (i < 3) Sw[Cfg::oAcc + i] = 0.0;  // dV0, dV1, Qu_err
    int qt = N - 1, status = CDDP_B200_STATUS_RUNNING, failures = 0;
    bool run = alive, ok = false;
    while (true) {
      if (!barrier_any(run)) break;
      {
        // Q_uu_reg (:130-131): this lane's row in original order, from the upper-triangle entries
        double Ho[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) Ho[j] = S[Cfg::oQuu + (i < j ? i * 4 + j : j * 4 + i)] + (j == i ? reg : 0.0);
        const double g = S[Cfg::oQu + i];
        double kk = 0.0, Hk = 0.0, inv[4];
        unsigned fm = 15u;
        bool good;
        if (c.has_box) {  // (:147-159)
          const double un = S[Cfg::oU + i];
          kk = S[Cfg::oKprev + i];
          // (lane-dependent index into a kernel-parameter array would make the compiler copy the array to local memory)
          const double lbi = i == 0 ? c.lb[0] : (i == 1 ? c.lb[1] : (i == 2 ? c.lb[2] : c.lb[3]));
          const double ubi = i == 0 ? c.ub[0] : (i == 1 ? c.ub[1] : (i == 2 ? c.ub[2] : c.ub[3]));
          const int qs = QuadQP::solve_box(c.opt, run, S + Cfg::oQuu, reg, Ho, g, lbi - un, ubi - un, kk, Hk, fm, inv, i, qb);
          good = !(qs == QP_HESSIAN_NOT_PD || qs == QP_NO_DESCENT);
        } else {  // k = -H^-1 Q_u (:142-144)
          good = QuadQP::inverse_row(S + Cfg::oQuu, reg, 15u, i, qb, inv);
          double v4[4];
          QuadQP::gather_rot(g, qb, i, v4);
          double sacc = 0.0;
#pragma unroll
          for (int bc = 0; bc < 4; ++bc) sacc = fma(inv[bc], v4[bc], sacc);
          kk = -sacc;
          QuadQP::gather(kk, qb, v4);
          sacc = 0.0;
#pragma unroll
          for (int j = 0; j < 4; ++j) sacc = fma(Ho[j], v4[j], sacc);
          Hk = sacc;
        }
        // dV += (Q_u.k, 0.5 k^T Q_uu k), unregularised Q_uu (:184-186); all lanes take part in the quad sums
        const double sq = fma(-reg, kk, Hk);  // (Q_uu k)_i = (Q_uu_reg k)_i - reg k_i
        const double d0 = QuadQP::quad_sum(g * kk), d1 = QuadQP::quad_sum(kk * sq), linf = QuadQP::quad_max(fabs(g));
        if (run) {
          if (good) {
            const bool fi = (fm >> i) & 1u;
#pragma unroll
            for (int bc = 0; bc < 4; ++bc) {  // inverse of the free block, clamped rows / columns zero (:161-178)
              const int j = (i + bc) & 3;
              Sw[Cfg::oHinv + i * 4 + j] = (fi && ((fm >> j) & 1u)) ? inv[bc] : 0.0;
            }
            Sw[Cfg::oW + i] = sq + g;
            Sw[Cfg::oKk + i] = kk;
            // lane 0: dV0 += d0, lane 1: dV1 += 0.5 d1, lane 2: Qu_err = max(Qu_err, linf) (:184-186, :195)
            if (i < 3) {
              const double acc = S[Cfg::oAcc + i];
              Sw[Cfg::oAcc + i] = i == 0 ? acc + d0 : (i == 1 ? acc + 0.5 * d1 : max_ref(acc, linf));
            }
            if (i == 0) *reinterpret_cast<volatile int *>(Sw + Cfg::oCtrl) = CTRL_OK;
            if (--qt < 0) {
              run = false;
              ok = true;
            }
          } else if (mode == BW_SINGLE) {
            if (i == 0) *reinterpret_cast<volatile int *>(Sw + Cfg::oCtrl) = CTRL_FAIL;
            run = false;
          } else {
            // increaseRegularization + limit test (cddp_solver_base.cpp:95-109, cddp_core.cpp:308-326)
            reg = fmin(reg * c.opt.reg_update_factor, c.opt.reg_max_value);
            ++failures;
            if (reg >= c.opt.reg_max_value) {
              status = CDDP_B200_STATUS_REG_LIMIT;
              if (i == 0) *reinterpret_cast<volatile int *>(Sw + Cfg::oCtrl) = CTRL_FAIL;
              run = false;
            } else {
              if (i == 0) *reinterpret_cast<volatile int *>(Sw + Cfg::oCtrl) = CTRL_RESTART;
              if (i < 3) Sw[Cfg::oAcc + i] = 0.0;
              qt = N - 1;
            }
          }
        }
      }
      barrier_all();
    }
    __syncwarp();
    if (alive && i == 0)
      sweep_epilogue<NS>(c, d, mode, b, ok, S[Cfg::oNrm], N, reg, S[Cfg::oAcc], S[Cfg::oAcc + 1], S[Cfg::oAcc + 2], status, failures);
  } else {
    // =========================================================== QP warp: lane q <-> trajectory q
    constexpr int MC = NC;
    const int q = lane;
    const int b = slot_instance(d, blockIdx.x * T + (q < T ? q : 0));
    const bool alive = q < T && b < d.B && !(mode == BW_ITERATE && d.status[b] != CDDP_B200_STATUS_RUNNING);
    const double *S = traj0 + (q < T ? q : T - 1) * ST;
    double *Sw = traj0 + (q < T ? q : T - 1) * ST;
    // ---- fused linearisation: this lane's trajectory, one record per step, three steps ahead of the sweep
    const bool need_lin = FMODEL >= 0 && mode == BW_ITERATE && alive && !d.lin_valid[b];
    int lt = N - 1;  // next timestep whose record is still to be formed
    auto produce = [&](int tt) {
      if constexpr (FMODEL >= 0) {
        const int cur = d.cur[b];
        const double *xp = d.X[cur] + ((size_t)b * (N + 1) + tt) * NS;
        const double *up = d.U[cur] + ((size_t)b * N + tt) * NC;
        double x[NS], u[NC], Fx[NS * NS], Fu[NS * NC], lxv[NS], luv[NC];
#pragma unroll
        for (int i2 = 0; i2 < NS; ++i2) x[i2] = xp[i2];
#pragma unroll
        for (int i2 = 0; i2 < NC; ++i2) u[i2] = up[i2];
        Model<FMODEL>::jac(c.mp, x, u, Fx, Fu);
        const double *ref = kern::ref_ptr(d, b, tt);
        double e[NS];
#pragma unroll
        for (int i2 = 0; i2 < NS; ++i2) e[i2] = x[i2] - ref[i2];
        if (c.cost_diag) {  // the statements of linearize_kernel (kernels_linearize.cuh), in its order
#pragma unroll
          for (int i2 = 0; i2 < NS; ++i2) lxv[i2] = 0.0 + c.Qdt2[i2 * NS + i2] * e[i2];
#pragma unroll
          for (int i2 = 0; i2 < NC; ++i2) luv[i2] = 0.0 + c.Rdt2[i2 * NC + i2] * u[i2];
        } else {
#pragma unroll
          for (int i2 = 0; i2 < NS; ++i2) {
            double acc = 0.0;
#pragma unroll
            for (int j2 = 0; j2 < NS; ++j2) acc += c.Qdt2[i2 * NS + j2] * e[j2];
            lxv[i2] = acc;
          }
#pragma unroll
          for (int i2 = 0; i2 < NC; ++i2) {
            double acc = 0.0;
#pragma unroll
            for (int j2 = 0; j2 < NC; ++j2) acc += c.Rdt2[i2 * NC + j2] * u[j2];
            luv[i2] = acc;
          }
        }
        double rv[RS];  // the packed record, then 16-byte stores (a record is a 16-byte multiple, 16-byte aligned)
#pragma unroll
        for (int i2 = 0; i2 < RS; ++i2) rv[i2] = 0.0;
        static_for<0, NS>([&](auto lc) {
          constexpr int l = decltype(lc)::value;
          static_for<0, NS>([&](auto jc) {
            constexpr int j = decltype(jc)::value;
            if constexpr (PAT::a(l, j)) rv[L::idxA(l, j)] = c.dt * Fx[l * NS + j] + (l == j ? 1.0 : 0.0);
          });
          if constexpr (PAT::brow(l)) {
            static_for<0, NC>([&](auto ac) {
              constexpr int a = decltype(ac)::value;
              rv[L::idxB(l, a)] = c.dt * Fu[l * NC + a];
            });
          }
          rv[L::offLx + l] = lxv[l];
        });
        static_for<0, NC>([&](auto ac) {
          constexpr int a = decltype(ac)::value;
          rv[L::offLu + a] = luv[a];
          rv[L::offU + a] = u[a];
        });
        double2 *dst = reinterpret_cast<double2 *>(d.rec + ((size_t)b * N + tt) * RS);
#pragma unroll
        for (int i2 = 0; i2 < RS / 2; ++i2) dst[i2] = make_double2(rv[2 * i2], rv[2 * i2 + 1]);
        if (tt > 0) {  // the next window's x, u (the timestep below): into L1 now, so that its loads do not start the window with an HBM round trip
          asm volatile("prefetch.global.L1 [%0];" ::"l"(xp - NS));
          asm volatile("prefetch.global.L1 [%0];" ::"l"(xp - 1));
          asm volatile("prefetch.global.L1 [%0];" ::"l"(up - NC));
          asm volatile("prefetch.global.L1 [%0];" ::"l"(up - 1));
        }
      }
    };
    if (need_lin) {
      if constexpr (FMODEL >= 0) {
        // terminal gradient V_x(N) = 2 Qf (x_N - ref) (objective.cpp:122-126), as linearize_kernel forms it
        const double *xp = d.X[d.cur[b]] + ((size_t)b * (N + 1) + N) * NS;
        const double *ref = d.xref + (size_t)b * NS;
        double e[NS];
#pragma unroll
        for (int i2 = 0; i2 < NS; ++i2) e[i2] = xp[i2] - ref[i2];
        for (int i2 = 0; i2 < NS; ++i2) {
          double acc = 0.0;
#pragma unroll
          for (int j2 = 0; j2 < NS; ++j2) acc += c.Qf2[i2 * NS + j2] * e[j2];
          d.vterm[(size_t)b * NS + i2] = acc;
        }
      }
      // records N-1, N-2 are staged right after the barrier below, N-3 during the first step: a record is always formed
      // one whole step before the bulk copy that reads it is issued, so that the fence which pushes it to the L2
      // (__threadfence at the top of the NEXT window) finds its stores long since acknowledged
      for (int k2 = 0; k2 < 3 && lt >= 0; ++k2) produce(lt--);
      __threadfence();
    }
    __syncthreads();
    double reg = alive ? d.reg[b] : 0.0;
    if (alive && mode == BW_ITERATE) d.iter[b] += 1;  // ++iter, cddp_solver_base.cpp:75
    double dV0 = 0.0, dV1 = 0.0, Qu_err = 0.0;
    int qt = N - 1, status = CDDP_B200_STATUS_RUNNING, failures = 0;
    bool run = alive, ok = false;
    constexpr unsigned all = (1u << NC) - 1u;
    while (true) {
      if (need_lin) {  // the idle window: barrier 2 of the previous step ... barrier 1 of this one
        __threadfence();  // the record of the PREVIOUS window is at the L2 before the barrier that lets a matrix lane copy it
        if (lt >= 0) produce(lt--);
      }
      if (!barrier_any(run)) break;
      if (run) {
        double H[MC * MC], g[MC], kk[MC], Hk[MC];
#pragma unroll
        for (int a = 0; a < MC; ++a) {
          g[a] = S[Cfg::oQu + a];
#pragma unroll
          for (int bcol = a; bcol < MC; ++bcol) {  // symmetric storage from the upper triangle
            const double h = S[Cfg::oQuu + a * MC + bcol] + (a == bcol ? reg : 0.0);  // Q_uu_reg (:130-131)
            H[a * MC + bcol] = h;
            H[bcol * MC + a] = h;
          }
        }
        unsigned free_mask = all;
        bool good;
        if constexpr (MC <= 4) {
          // closed-form inverse + Sylvester PD test (boxqp_small.cuh)
          double Hinv[MC * MC];
          if (c.has_box) {  // (:147-159)
            good = SymPD<MC>::run(H);  // PD test (:133-140)
            if (good) {
              double lo[MC], hi[MC];
#pragma unroll
              for (int a = 0; a < MC; ++a) {
                const double un = S[Cfg::oU + a];
                lo[a] = c.lb[a] - un;
                hi[a] = c.ub[a] - un;
                kk[a] = S[Cfg::oKprev + a];
              }
              const int qs = SmallQP<MC>::solve(c.opt, H, g, lo, hi, kk, free_mask, Hinv, Hk);
              good = !(qs == QP_HESSIAN_NOT_PD || qs == QP_NO_DESCENT);
              if (good && c.opt.qp_max_iterations <= 0) SmallQP<MC>::masked_inverse_raw(H, free_mask, Hinv);
              SmallQP<MC>::zero_clamped(free_mask, Hinv);  // ALL_CLAMPED: free_mask == 0 -> K = 0 (:163)
            }
          } else {  // k = -H^-1 Q_u (:142-144)
            good = SymInverse<MC>::run(H, Hinv);
#pragma unroll
            for (int a = 0; a < MC; ++a) {
              double sacc = 0.0;
#pragma unroll
              for (int bcol = 0; bcol < MC; ++bcol) sacc = fma(Hinv[a * MC + bcol], g[bcol], sacc);
              kk[a] = -sacc;
            }
#pragma unroll
            for (int a = 0; a < MC; ++a) {
              double sacc = 0.0;
#pragma unroll
              for (int bcol = 0; bcol < MC; ++bcol) sacc = fma(H[a * MC + bcol], kk[bcol], sacc);
              Hk[a] = sacc;
            }
          }
          if (good) {
#pragma unroll
            for (int e = 0; e < MC * MC; ++e) Sw[Cfg::oHinv + e] = Hinv[e];
          }
        } else {
          double Lf[MC * MC];
          good = SmallMat<MC>::masked_cholesky(NC, H, all, Lf);  // PD test (:133-140)
          if (good) {
            if (c.has_box) {
              double lo[MC], hi[MC];
#pragma unroll
              for (int a = 0; a < MC; ++a) {
                const double un = S[Cfg::oU + a];
                lo[a] = c.lb[a] - un;
                hi[a] = c.ub[a] - un;
                kk[a] = S[Cfg::oKprev + a];
              }
              const int qs = SmallMat<MC>::boxqp(c.opt, NC, H, g, lo, hi, kk, free_mask, Lf, true);
              good = !(qs == QP_HESSIAN_NOT_PD || qs == QP_NO_DESCENT);
              if (good && c.opt.qp_max_iterations <= 0) SmallMat<MC>::masked_cholesky(NC, H, free_mask, Lf);
            } else {
#pragma unroll
              for (int a = 0; a < MC; ++a) kk[a] = 0.0;
            }
          }
          if (good) {
            // inverse of the free block (clamped rows/columns zero): K = -Hinv Q_ux  (:142-145, :161-178)
#pragma unroll
            for (int col = 0; col < MC; ++col) {
              double e[MC];
              const bool fc = (free_mask >> col) & 1u;
              SmallMat<MC>::chol_inverse_column(NC, Lf, col, e);
#pragma unroll
              for (int a = 0; a < MC; ++a) {
                const double hv = (fc && ((free_mask >> a) & 1u)) ? e[a] : 0.0;
                Sw[Cfg::oHinv + a * MC + col] = hv;
                if (!c.has_box) kk[a] = fma(-hv, g[col], kk[a]);  // k = -H^-1 Q_u (:144)
              }
            }
#pragma unroll
            for (int a = 0; a < MC; ++a) {
              double sacc = 0.0;
#pragma unroll
              for (int bcol = 0; bcol < MC; ++bcol) sacc = fma(H[a * MC + bcol], kk[bcol], sacc);
              Hk[a] = sacc;
            }
          }
        }
        if (good) {
          double d0 = 0.0, d1 = 0.0, linf = 0.0;
#pragma unroll
          for (int a = 0; a < MC; ++a) {  // dV += (Q_u.k, 0.5 k^T Q_uu k), unregularised Q_uu (:184-186)
            const double sq = fma(-reg, kk[a], Hk[a]);  // (Q_uu k)_a = (Q_uu_reg k)_a - reg k_a
            d0 = fma(g[a], kk[a], d0);
            d1 = fma(kk[a], sq, d1);
            linf = max_ref(linf, fabs(g[a]));
            Sw[Cfg::oW + a] = sq + g[a];
            Sw[Cfg::oKk + a] = kk[a];
          }
          dV0 += d0;
          dV1 += 0.5 * d1;
          Qu_err = max_ref(Qu_err, linf);  // (:195)
          *reinterpret_cast<volatile int *>(Sw + Cfg::oCtrl) = CTRL_OK;
          if (--qt < 0) {
            run = false;
            ok = true;
          }
        } else if (mode == BW_SINGLE) {
          *reinterpret_cast<volatile int *>(Sw + Cfg::oCtrl) = CTRL_FAIL;
          run = false;
        } else {
          // increaseRegularization + limit test (cddp_solver_base.cpp:95-109, cddp_core.cpp:308-326)
          reg = fmin(reg * c.opt.reg_update_factor, c.opt.reg_max_value);
          ++failures;
          if (reg >= c.opt.reg_max_value) {
            status = CDDP_B200_STATUS_REG_LIMIT;
            *reinterpret_cast<volatile int *>(Sw + Cfg::oCtrl) = CTRL_FAIL;
            run = false;
          } else {
            *reinterpret_cast<volatile int *>(Sw + Cfg::oCtrl) = CTRL_RESTART;
            dV0 = dV1 = Qu_err = 0.0;
            qt = N - 1;
          }
        }
      }
      barrier_all();
    }
    if (alive) sweep_epilogue<NS>(c, d, mode, b, ok, S[Cfg::oNrm], N, reg, dV0, dV1, Qu_err, status, failures);
  }
}

template <int NS, int NC, class PAT, int W, int MINB, bool QPQ = false, int FMODEL = -1>
cudaError_t launch_sweep(const Constants &c, const DeviceState &d, int mode, cudaStream_t st) {
  using Cfg = SweepCfg<NS, NC, PAT, W, QPQ>;
  static_assert(QPQ || Cfg::T <= 32, "one QP-warp lane per trajectory");
  static_assert(NS + 1 <= Cfg::G * Cfg::R, "V_x rides as an extra row");
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(sweep_kernel<NS, NC, PAT, W, MINB, QPQ, FMODEL>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         (int)Cfg::smemBytes);
    if (e != cudaSuccess) return e;
    configured = true;
  }
  const int blocks = (d.n_slots + Cfg::T - 1) / Cfg::T;
  sweep_kernel<NS, NC, PAT, W, MINB, QPQ, FMODEL><<<blocks, Cfg::threads, Cfg::smemBytes, st>>>(c, d, mode);
  return cudaGetLastError();
}

}  // namespace

// Developer switch (A-B timing only): CDDP_B200_SWEEP_VARIANT=quad selects the 12-warp layout for m = 4 (SweepCfg::QPQ:
// four QP warps, four lanes per trajectory, set-level barriers, setmaxnreg).  It passes the same parity tests but is
// SLOWER on B200 (0.815 ms against 0.571 ms for the headline batch, profiles/r02_sweep_quad_qp.txt): the quad-parallel
// BoxQP is a chain of shuffle round trips (1 575 warp-instructions per step at 7.3 cycles each against 1 900 at 3.5 for
// the one-lane-per-trajectory QP warp), and eight matrix warps per SM run 9 % slower than seven.  Default: one QP warp.
static int sweep_variant() {  // read at every launch (not cached): the parity test of the variant switches it per solve
  const char *e = std::getenv("CDDP_B200_SWEEP_VARIANT");
  return (e && std::string(e) == "quad") ? 0 : 1;
}

// Fused linearisation is OPT-IN (cddp_b200_set_fused_linearization, or CDDP_B200_FUSED_LINEARIZE=1 for every handle): it passes the same parity tests — the lock-step run of the
// headline batch gives the same worst-case errors to the last digit — but it is SLOWER on B200.  Forming one quadrotor
// record is ~750 dependent instructions for a single lane (~8 k cycles in the QP warp, which shares its scheduler's FP64
// pipe with a matrix warp in its critical phase C); the idle window is 4.7 k.  Measured, headline batch: sweep 0.551 ->
// 0.701 ms with the linearize launch (0.115 ms) gone: iteration 1.061 -> 1.114 ms (first version, fence and loads on
// the window's critical path: 1.181 ms).
static bool sweep_fused_enabled() {
  static const bool v = [] {
    const char *e = std::getenv("CDDP_B200_FUSED_LINEARIZE");
    return e && std::string(e) == "1";
  }();
  return v;
}

// One-control models: the inline kernel (sweep_inline.cuh) where the subproblem is a division (no control box:
// cartpole 1024 x 100 steps 0.116 -> 0.073 ms, 16384: 0.303 -> 0.226 ms), the warp-specialised kernel where it is the
// BoxQP iteration, which the inline kernel would run redundantly on every lane and divergently across the trajectories of
// a warp (pendulum 1 x 500 steps: 0.516 vs 0.588 ms; 4096: 0.572 vs 0.773 ms).  CDDP_B200_INLINE_SWEEP=0 / =1 forces
// one of them (A-B timing and the both-kernels parity test; read per launch).
static bool inline_sweep_enabled(const Constants &c) {
  const char *e = std::getenv("CDDP_B200_INLINE_SWEEP");
  if (e && (e[0] == '0' || e[0] == '1') && e[1] == 0) return e[0] == '1';
  return !c.has_box;
}

// true if launch_backward(BW_ITERATE) forms the linearisation records itself (the iteration loop then skips linearize)
bool backward_fuses_linearization(const Constants &c, const DeviceState &d) {
  return d.layout == RECORDS_STRUCTURED && c.model == CDDP_B200_MODEL_QUADROTOR && sweep_variant() != 0 &&
         (c.fuse_lin || sweep_fused_enabled());
}

cudaError_t launch_backward_fast(const Constants &c, const DeviceState &d, int mode, cudaStream_t st, bool *handled) {
  *handled = true;
  const int n = d.n, m = d.m;
  if (d.layout == RECORDS_STRUCTURED) {
    if (c.model == CDDP_B200_MODEL_QUADROTOR) {
      if (sweep_variant() == 0) return launch_sweep<13, 4, ModelPattern<CDDP_B200_MODEL_QUADROTOR>, 8, 1, true>(c, d, mode, st);
      if (c.fuse_lin || sweep_fused_enabled())
        return launch_sweep<13, 4, ModelPattern<CDDP_B200_MODEL_QUADROTOR>, 7, 1, false, CDDP_B200_MODEL_QUADROTOR>(c, d, mode, st);
      return launch_sweep<13, 4, ModelPattern<CDDP_B200_MODEL_QUADROTOR>, 7, 1>(c, d, mode, st);
    }
    if (c.model == CDDP_B200_MODEL_CARTPOLE)  // one control: inline subproblem, no QP warp, no CTA barrier (sweep_inline.cuh)
      return inline_sweep_enabled(c) ? launch_sweep_inline<4, ModelPattern<CDDP_B200_MODEL_CARTPOLE>, 2>(c, d, mode, st)
                                    : launch_sweep<4, 1, ModelPattern<CDDP_B200_MODEL_CARTPOLE>, 7, 2>(c, d, mode, st);
    if (c.model == CDDP_B200_MODEL_UNICYCLE) return launch_sweep<3, 2, ModelPattern<CDDP_B200_MODEL_UNICYCLE>, 4, 4>(c, d, mode, st);
  } else {
    if (n == 2 && m == 1)
      return inline_sweep_enabled(c) ? launch_sweep_inline<2, DensePattern, 2>(c, d, mode, st) : launch_sweep<2, 1, DensePattern, 4, 4>(c, d, mode, st);
    if (n == 3 && m == 2) return launch_sweep<3, 2, DensePattern, 4, 4>(c, d, mode, st);
    if (n == 4 && m == 1)
      return inline_sweep_enabled(c) ? launch_sweep_inline<4, DensePattern, 2>(c, d, mode, st) : launch_sweep<4, 1, DensePattern, 7, 2>(c, d, mode, st);
    if (n == 4 && m == 2) return launch_sweep<4, 2, DensePattern, 7, 2>(c, d, mode, st);
    if (n == 6 && m == 3) return launch_sweep<6, 3, DensePattern, 7, 2>(c, d, mode, st);
    if (n == 13 && m == 4) {
      if (sweep_variant() == 0) return launch_sweep<13, 4, DensePattern, 8, 1, true>(c, d, mode, st);
      return launch_sweep<13, 4, DensePattern, 7, 1>(c, d, mode, st);
    }
    if (n == 14 && m == 7) return launch_sweep<14, 7, DensePattern, 5, 1>(c, d, mode, st);
  }
  *handled = false;
  return cudaSuccess;
}

}  // namespace cddp_b200
