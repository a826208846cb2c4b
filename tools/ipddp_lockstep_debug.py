"""Developer script: at a given iteration of the config-4 solve, compare the CUDA path's per-alpha line-search table
(success, cost, merit, theta) with the oracle's candidates evaluated from the same state.
usage: python tools/ipddp_lockstep_debug.py ITER INSTANCE [INSTANCE ...]"""
import importlib
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import oracle_binding as ob  # noqa: E402
from test_gpu_every_instance import ip_snapshot  # noqa: E402

cddp = importlib.import_module("cddp-cpp_b200")
problems = importlib.import_module("cddp-cpp_b200.problems")
it, inst = int(sys.argv[1]), [int(a) for a in sys.argv[2:]]
B = 2048
cfg = problems.make_config("unicycle_obstacle_teq", batch=B)
opts = dict(cfg["options"])
s = cddp.BatchedIPDDP(cfg["spec"], cddp.default_options(**opts), cddp.default_ipddp_options(**cfg["ipddp_options"]), cfg["constraints"], B)
s.set_instances(cfg["x0"], cfg["xref"], None, cfg["U0"], None)
s.enable_trace(True)
s.initialize()
s.iterate(it)
pre = ip_snapshot(s)
s.iterate(1)
post = ip_snapshot(s)
code = s.get_trace()[:, it]
ls = s.get_line_search()
P, oo = ob.OracleProblem(cfg["spec"]), ob.make_options(**opts)
oi, cs = ob.make_ipddp_options(**cfg["ipddp_options"]), ob.ConstraintSet(cfg["constraints"])
run = np.array(inst)
state = {k: pre[k][run] for k in ("X", "U", "Y", "S", "G", "lamT", "filter", "filter_size", "mu", "cost", "merit", "filter_theta", "inf_pr",
                                  "inf_comp", "reg", "alpha_du", "step_norm", "inf_du")}
state["alpha_pr"] = pre["alpha"][run]
state["iter"] = np.full(run.size, float(it + 1))
o = ob.ipddp_iterate_batch(P, oo, oi, cs, cfg["x0"][run], cfg["xref"][run], state)  # own decisions: every candidate up to its first success
with ob.variant():  # the same state through the oracle built with floating-point contraction on
    P2, oo2 = ob.OracleProblem(cfg["spec"]), ob.make_options(**opts)
    o2 = ob.ipddp_iterate_batch(P2, oo2, ob.make_ipddp_options(**cfg["ipddp_options"]), ob.ConstraintSet(cfg["constraints"]),
                                cfg["x0"][run], cfg["xref"][run], state)
np.set_printoptions(linewidth=200, precision=12)
for j, b in enumerate(inst):
    print(f"instance {b} iteration {it}: recorded code {code[b]:#x}, oracle own code {o['code'][j]:#x}; alpha_pr_max gpu {pre['alpha_pr_max'][b]:.15g}")
    print(f"  step cap alpha_pr of candidate 0: oracle strict {o['trials'][j, 0, 5]:.15g}, oracle fp-contract {o2['trials'][j, 0, 5]:.15g}, "
          f"CUDA {post['alpha_pr_max'][b]:.15g}; oracle fp-contract own code {o2['code'][j]:#x}; min slack {pre['S'][b].min():.3e} min dual {pre['Y'][b].min():.3e}")
    print("  alpha | GPU success cost merit theta | oracle success cost merit theta margin alpha_pr")
    for a in range(ls.shape[1]):
        print(f"  {a:2d} | {ls[b, a]} | {o['trials'][j, a]}")
