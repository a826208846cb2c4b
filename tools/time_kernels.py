"""Developer script: CUDA-event timings of the three kernels for a workload variant.
usage: python tools/time_kernels.py [config] [batch] [--nobox] [--dense] [--iters K] [--cold]
--cold: time the FIRST K iterations of the solve (default: iterations 4..K+3, after 3 untimed ones)"""
import importlib
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
cddp = importlib.import_module("cddp-cpp_b200")
problems = importlib.import_module("cddp-cpp_b200.problems")


def main():
    args = [a for a in sys.argv[1:] if not a.startswith("--")]
    name = args[0] if args else "quadrotor"
    B = int(args[1]) if len(args) > 1 else 4096
    K = int(sys.argv[sys.argv.index("--iters") + 1]) if "--iters" in sys.argv else 10
    cfg = problems.make_config(name, batch=B)
    spec = dict(cfg["spec"])
    if "--nobox" in sys.argv:
        spec["lb"] = spec["ub"] = None
    opts = dict(cfg["options"], tolerance=0.0, acceptable_tolerance=0.0, max_iterations=K + 3)
    if cfg.get("solver") == "ipddp":
        s = cddp.BatchedIPDDP(spec, cddp.default_options(**opts), cddp.default_ipddp_options(**cfg.get("ipddp_options", {})),
                              cfg["constraints"], B)
        s.set_instances(cfg["x0"], cfg["xref"], None, cfg["U0"], cfg["ref_traj"])
        s.initialize()
        if "--cold" not in sys.argv:
            s.iterate(3)
        s.enable_timing(True)
        s.reset_timing()
        s.iterate(K)
        t = s.get_timing()
        sc = s.get_scalars()
        it = t.linearize_ms / t.linearize_launches + t.backward_ms / t.backward_launches + t.forward_ms / t.forward_launches
        print(f"{name} IPDDP B={B} d={s.d}: linearize {t.linearize_ms / t.linearize_launches:.3f} ms  backward "
              f"{t.backward_ms / t.backward_launches:.3f} ms  forward {t.forward_ms / t.forward_launches:.3f} ms  -> "
              f"{B / it * 1e3:.3e} instance-iterations/s; running {int((sc['status'] == 0).sum())}/{B} mean cost {np.mean(sc['cost']):.4f}")
        s.close()
        return
    s = cddp.BatchedCLDDP(spec, cddp.default_options(**opts), B)
    if "--dense" in sys.argv:
        s.set_record_layout("dense")
    if "--fused" in sys.argv:
        s.set_fused_linearization(int(sys.argv[sys.argv.index("--fused") + 1]))
    s.set_instances(cfg["x0"], cfg["xref"], cfg["X0"], cfg["U0"], cfg["ref_traj"])
    s.initialize()
    if "--cold" not in sys.argv:
        s.iterate(3)
    s.enable_timing(True)
    s.reset_timing()
    s.iterate(K)
    t = s.get_timing()
    sc = s.get_scalars()
    lay, nb = s.get_record_layout()
    alg = s.backward_algorithmic_bytes()
    bw = t.backward_ms / t.backward_launches
    if t.linearize_launches == 0:
        t.linearize_launches = 1
    print(f"{name} B={B} box={'lb' in spec and spec['lb'] is not None} layout={lay} ({nb} B/record): "
          f"linearize {t.linearize_ms / t.linearize_launches:.3f} ms  backward {bw:.3f} ms ({alg / bw / 1e6 / 6539.2 * 100:.1f}% of HBM peak, algorithmic)  "
          f"forward {t.forward_ms / t.forward_launches:.3f} ms  mean cost {np.mean(sc['cost']):.4f} finite={np.isfinite(sc['cost']).all()}")
    s.close()


if __name__ == "__main__":
    main()
