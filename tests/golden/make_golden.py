"""Generates tests/golden/*.npz.

The reference cannot be built or imported in this image (Eigen 3.4 / autodiff are network deps), so
these are NOT outputs of the reference.  They are:
  * boxqp_fixtures.npz — the two BoxQP INPUT fixtures of the reference's own test
    (tests/cddp_core/test_boxqp.cpp:59-64,89-90 and :125-220; that test asserts nothing, it only prints),
    parsed from /root/reference at generation time, plus the solution computed by the independent numpy
    restatement (oracle/np_oracle.py) and verified here against the KKT conditions of the box QP.
  * clddp_<config>.npz — full CLDDP solves of small instances of every workload by the independent
    numpy restatement (complex-step Jacobians, numpy.linalg eig/inv/solve).  The C++ oracle and the CUDA
    path are both tested against them.
Run:  python tests/golden/make_golden.py   (needs /root/reference only for the BoxQP inputs)
"""
import importlib
import os
import re
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import np_oracle as npo  # noqa: E402

problems = importlib.import_module("cddp-cpp_b200.problems")

GOLDEN_CASES = [  # name, batch, horizon, max_iterations
    ("pendulum", 2, 80, 25), ("cartpole", 2, 50, 25), ("unicycle", 3, 50, 20), ("quadrotor", 2, 40, 20),
    ("quadrotor_fig8", 2, 40, 20), ("lti", 3, 30, 10),
]


def parse_boxqp_fixtures(path):
    src = open(path).read()
    num = r"[-+]?\d+\.?\d*(?:[eE][-+]?\d+)?"
    # 5x5: "Q << ... ;" and "q << ... ;" inside ComparisonTest
    t1 = src[src.index("TEST(QPSolver, ComparisonTest)"):src.index("TEST(BoxQPSolver, LargeDimensionTest)")]
    Q5 = np.array([float(v) for v in re.findall(num, re.search(r"\bQ <<(.*?);", t1, re.S).group(1))]).reshape(5, 5)
    q5 = np.array([float(v) for v in re.findall(num, re.search(r"\bq <<(.*?);", t1, re.S).group(1))])
    t2 = src[src.index("TEST(BoxQPSolver, LargeDimensionTest)"):]
    strip = lambda s: re.sub(r"//.*", "", s)  # noqa: E731
    Q15 = np.array([float(v) for v in re.findall(num, strip(re.search(r"\bQ <<(.*?);", t2, re.S).group(1)))]).reshape(15, 15)
    q15 = np.array([float(v) for v in re.findall(num, strip(re.search(r"\bq <<(.*?);", t2, re.S).group(1)))])
    assert q15.shape == (15,)
    return Q5, q5, Q15, q15


def kkt_ok(H, g, lo, hi, x, tol=1e-6):
    grad = H @ x + g
    for i in range(len(x)):
        if x[i] <= lo[i] + 1e-12:
            assert grad[i] >= -tol, (i, grad[i])
        elif x[i] >= hi[i] - 1e-12:
            assert grad[i] <= tol, (i, grad[i])
        else:
            assert abs(grad[i]) <= tol, (i, grad[i])
    assert (x >= lo - 1e-12).all() and (x <= hi + 1e-12).all()


def main():
    ref = "/root/reference/tests/cddp_core/test_boxqp.cpp"
    Q5, q5, Q15, q15 = parse_boxqp_fixtures(ref)
    o = npo.options()
    out = {}
    for tag, H, g, lo, hi in (("5", Q5, q5, 0.0, 2.0), ("15", Q15, q15, -2.0, 2.0)):
        n = len(g)
        lo, hi = np.full(n, lo), np.full(n, hi)
        r = npo.boxqp(o, H, g, lo, hi, None)
        kkt_ok(H, g, lo, hi, r["x"])
        out.update({f"H{tag}": H, f"g{tag}": g, f"lo{tag}": lo, f"hi{tag}": hi, f"x{tag}": r["x"],
                    f"free{tag}": r["free"].astype(np.int32), f"status{tag}": r["status"], f"value{tag}": r["value"],
                    f"iters{tag}": r["iterations"]})
        print("boxqp", tag, "status", r["status"], "iters", r["iterations"], "value", r["value"])
    np.savez(os.path.join(HERE, "boxqp_fixtures.npz"), **out)

    for name, B, N, iters in GOLDEN_CASES:
        cfg = problems.make_config(name, batch=B, horizon=N)
        opts = dict(cfg["options"])
        opts["max_iterations"] = min(opts["max_iterations"], iters)
        P, o = npo.Problem(cfg["spec"]), npo.options(**opts)
        res = []
        for b in range(B):
            rt = None if cfg["ref_traj"] is None else cfg["ref_traj"][b]
            res.append(npo.solve(P, o, cfg["x0"][b], cfg["xref"][b], cfg["X0"][b], cfg["U0"][b], rt))
        # first backward sweep + all-alpha forward costs of instance 0 (step-level golden)
        rt0 = None if cfg["ref_traj"] is None else cfg["ref_traj"][0]
        X0 = cfg["X0"][0].copy()
        X0[0] = cfg["x0"][0]
        bw = npo.backward_pass(P, o, X0, cfg["U0"][0], cfg["xref"][0], o["reg_initial_value"], np.zeros((N, P.m)), rt0)
        c0 = P.trajectory_cost(X0, cfg["U0"][0], cfg["xref"][0], rt0)
        fw = [npo.forward_pass(P, o, cfg["x0"][0], X0, cfg["U0"][0], cfg["xref"][0], bw["K"], bw["k"], bw["dV"], c0, a, rt0)
              for a in npo.build_alphas(o)]
        np.savez(os.path.join(HERE, f"clddp_{name}.npz"), batch=B, horizon=N, max_iterations=opts["max_iterations"],
                 X=np.stack([r["X"] for r in res]), U=np.stack([r["U"] for r in res]), K=np.stack([r["K"] for r in res]),
                 k=np.stack([r["k"] for r in res]), cost=np.array([r["cost"] for r in res]),
                 iterations=np.array([r["iterations"] for r in res]), status=np.array([r["status"] for r in res]),
                 alpha=np.array([r["alpha"] for r in res]), reg=np.array([r["reg"] for r in res]),
                 inf_du=np.array([r["inf_du"] for r in res]),
                 bw_K=bw["K"], bw_k=bw["k"], bw_dV=bw["dV"], bw_inf_du=bw["inf_du"], bw_Vx0=bw["Vx0"], bw_Vxx0=bw["Vxx0"],
                 init_cost=c0, fw_costs=np.array([f["cost"] for f in fw]), fw_success=np.array([f["success"] for f in fw]))
        print(name, "iters", [r["iterations"] for r in res], "status", [r["status"] for r in res], "cost", [r["cost"] for r in res])


if __name__ == "__main__":
    main()
