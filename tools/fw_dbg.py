"""Developer script: per-iteration forward time and accepted step sizes (user-model speculative line search)."""
import importlib, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
cddp = importlib.import_module("cddp-cpp_b200")
problems = importlib.import_module("cddp-cpp_b200.problems")
name, B = sys.argv[1], int(sys.argv[2])
cfg = problems.make_config(name, batch=B)
opts = dict(cfg["options"], max_iterations=30) if "--tol" in sys.argv else dict(cfg["options"], tolerance=0.0, acceptable_tolerance=0.0, max_iterations=20)
s = cddp.BatchedCLDDP(dict(cfg["spec"]), cddp.default_options(**opts), B)
s.set_instances(cfg["x0"], cfg["xref"], cfg["X0"], cfg["U0"], cfg["ref_traj"])
s.initialize()
s.enable_timing(True)
for it in range(int(sys.argv[3]) if len(sys.argv) > 3 and sys.argv[3].isdigit() else 10):
    s.reset_timing()
    s.iterate(1)
    t = s.get_timing()
    sc = s.get_scalars()
    al, cnt = np.unique(sc["alpha"], return_counts=True)
    print(f"it {it}: lin {t.linearize_ms:.3f} bw {t.backward_ms:.3f} fw {t.forward_ms:.3f} ms ({t.forward_launches} launches) running {int((sc['status'] == 0).sum())} "
          f"alpha {dict(zip(np.round(al, 4).tolist(), cnt.tolist()))} mean cost {np.mean(sc['cost']):.6f}")
s.close()
