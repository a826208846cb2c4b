import importlib, sys, numpy as np
sys.path.insert(0, "/root/repo")
cddp = importlib.import_module("cddp-cpp_b200"); problems = importlib.import_module("cddp-cpp_b200.problems")
B = 4096
cfg = problems.make_config("quadrotor", batch=B)
opts = dict(cfg["options"], tolerance=0.0, acceptable_tolerance=0.0, max_iterations=40)
s = cddp.BatchedCLDDP(cfg["spec"], cddp.default_options(**opts), B)
s.set_instances(cfg["x0"], cfg["xref"], cfg["X0"], cfg["U0"]); s.initialize()
tot = np.zeros(17, dtype=np.int64)
for it in range(30):
    s.iterate(1)
    acc = s.get_forward()["accepted"]
    h = np.bincount(acc + 1, minlength=17)
    tot += h
    if it % 5 == 0: print(it, h.tolist())
print("total (index -1..15):", tot.tolist(), "frac within first 4:", tot[1:5].sum() / tot.sum(), "first 8:", tot[1:9].sum()/tot.sum())
