"""Developer script: what bounds the backward sweep?  Times the sweep kernel of the headline workload with the
control-space subproblem progressively removed (options only, no code change):
  box            the product path
  qp_iters=K     BoxQP capped at K projected-Newton iterations (wrong results; timing only)
  nobox          no ControlConstraint -> inverse branch (clddp_solver.cpp:142-145)
and at several batch sizes (a latency-bound kernel does not get faster with a smaller batch)."""
import importlib
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
cddp = importlib.import_module("cddp-cpp_b200")
problems = importlib.import_module("cddp-cpp_b200.problems")


def run(name, B, K=10, nobox=False, **extra):
    cfg = problems.make_config(name, batch=B)
    spec = dict(cfg["spec"])
    if nobox:
        spec["lb"] = spec["ub"] = None
    opts = dict(cfg["options"], tolerance=0.0, acceptable_tolerance=0.0, max_iterations=K + 3)
    opts.update(extra)
    s = cddp.BatchedCLDDP(spec, cddp.default_options(**opts), B)
    s.set_instances(cfg["x0"], cfg["xref"], cfg["X0"], cfg["U0"], cfg["ref_traj"])
    s.initialize()
    s.iterate(3)
    s.enable_timing(True)
    s.reset_timing()
    s.iterate(K)
    t = s.get_timing()
    sc = s.get_scalars()
    s.close()
    return (t.linearize_ms / t.linearize_launches, t.backward_ms / t.backward_launches, t.forward_ms / t.forward_launches,
            float(np.mean(sc["cost"])))


if __name__ == "__main__":
    name = sys.argv[1] if len(sys.argv) > 1 else "quadrotor"
    for B in (4096, 2048, 1024, 8192):
        print(name, "B", B, "box          lin %.3f bw %.3f fw %.3f cost %.3f" % run(name, B), flush=True)
    for k in (1, 2, 3, 4, 6):
        print(name, "B 4096 qp_iters=%d   lin %.3f bw %.3f fw %.3f cost %.3f" % ((k,) + run(name, 4096, qp_max_iterations=k)), flush=True)
    print(name, "B 4096 nobox        lin %.3f bw %.3f fw %.3f cost %.3f" % run(name, 4096, nobox=True), flush=True)
