import importlib, sys
sys.path.insert(0,'/root/repo')
cddp = importlib.import_module("cddp-cpp_b200"); problems = importlib.import_module("cddp-cpp_b200.problems")
cfg = problems.make_config("pendulum", batch=5, horizon=60)
s = cddp.BatchedCLDDP(cfg["spec"], cddp.default_options(**cfg["options"]), 5)
s.set_instances(cfg["x0"], cfg["xref"], cfg["X0"], cfg["U0"])
s.initialize(); s.linearize(); s.backward_pass(); print(s.get_sweep()["dV"])
