"""Developer script: step-by-step comparison of the CUDA path with the CPU oracle on small batches.
(The pytest -m gpu suite is the judged version of these checks; this prints more detail.)"""
import importlib
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import oracle_binding as ob  # noqa: E402

cddp = importlib.import_module("cddp-cpp_b200")
problems = importlib.import_module("cddp-cpp_b200.problems")


def rel(a, b):
    a, b = np.asarray(a), np.asarray(b)
    return float(np.max(np.abs(a - b)) / (np.max(np.abs(b)) + 1e-300))


def check(name, B, solve_iters=None):
    cfg = problems.make_config(name, batch=B)
    spec = cfg["spec"]
    P = ob.OracleProblem(spec)
    oo = ob.make_options(**cfg["options"])
    go = cddp.default_options(**cfg["options"])
    s = cddp.BatchedCLDDP(spec, go, B)
    s.set_instances(cfg["x0"], cfg["xref"], cfg["X0"], cfg["U0"], cfg["ref_traj"])
    s.initialize()
    sc = s.get_scalars()
    rt = cfg["ref_traj"]
    c0 = np.array([ob.trajectory_cost(P, cfg["X0"][b], cfg["U0"][b], cfg["xref"][b], None if rt is None else rt[b]) for b in range(B)])
    print(f"[{name}] init cost rel err {rel(sc['cost'], c0):.2e}")
    s.linearize()
    A, Bm = s.get_linearization()
    Ao = np.stack([ob.linearize(P, cfg["X0"][b], cfg["U0"][b])[0] for b in range(B)])
    Bo = np.stack([ob.linearize(P, cfg["X0"][b], cfg["U0"][b])[1] for b in range(B)])
    print(f"[{name}] linearize A rel {rel(A, Ao):.2e} B rel {rel(Bm, Bo):.2e}")
    s.backward_pass()
    sw = s.get_sweep()
    sol = s.get_solution()
    kff = s.get_feedforward()
    reg0 = cfg["options"].get("reg_initial_value", 1e-6)
    for b in range(min(B, 3)):
        r = ob.backward_pass(P, oo, cfg["X0"][b], cfg["U0"][b], cfg["xref"][b], reg0, ref_traj=None if rt is None else rt[b], debug=True)
        print(f"[{name}] b={b} bw ok gpu={sw['ok'][b]} cpu={r['ok']} K rel {rel(sol['K'][b], r['K']):.2e} k rel {rel(kff[b], r['k']):.2e} "
              f"dV {sw['dV'][b]} vs {r['dV']} inf_du {sw['inf_du'][b]:.6e} vs {r['inf_du']:.6e} Vxx0 rel {rel(sw['Vxx0'][b], r['Vxx'][0]):.2e}")
        if b == 0:
            s.forward_pass()
            fw = s.get_forward()
            alphas = ob.build_alphas(oo)
            costs = []
            first = -1
            for ai, a in enumerate(alphas):
                f = ob.forward_pass(P, oo, cfg["x0"][b], cfg["X0"][b], cfg["U0"][b], cfg["xref"][b], r["K"], r["k"], r["dV"], c0[b], a,
                                    ref_traj=None if rt is None else rt[b])
                costs.append(f["cost"])
                if f["success"] and first < 0:
                    first = ai
                    Xacc = f["X"]
            print(f"[{name}] fw costs rel {rel(fw['costs'][b], costs):.2e} accepted gpu={fw['accepted'][b]} cpu={first}"
                  + (f" X rel {rel(fw['X'][b], Xacc):.2e}" if first >= 0 else ""))
    # full solve
    t0 = time.time()
    s.set_instances(cfg["x0"], cfg["xref"], cfg["X0"], cfg["U0"], cfg["ref_traj"])
    s.solve()
    g = s.get_solution()
    tg = time.time() - t0
    t0 = time.time()
    o = ob.solve_batch(P, oo, cfg["x0"], cfg["xref"], cfg["X0"], cfg["U0"], cfg["ref_traj"], nthreads=ob.hardware_threads())
    tc = time.time() - t0
    relc = np.abs(g["cost"] - o["cost"]) / np.abs(o["cost"])
    print(f"[{name}] solve: gpu {tg:.3f}s cpu {tc:.3f}s; iters equal {np.mean(g['iterations'] == o['iterations']):.3f} status equal "
          f"{np.mean(g['status'] == o['status']):.3f} cost rel max {relc.max():.2e} median {np.median(relc):.2e} "
          f"frac<1e-6 {np.mean(relc < 1e-6):.3f}")
    print(f"[{name}]   gpu iters {g['iterations'][:6]} cpu iters {o['iterations'][:6]} gpu status {g['status'][:6]} cpu {o['status'][:6]}")
    print(f"[{name}]   X rel {rel(g['X'], o['X']):.2e} U rel {rel(g['U'], o['U']):.2e} K rel {rel(g['K'], o['K']):.2e}")
    s.close()


if __name__ == "__main__":
    names = sys.argv[1:] or ["lti", "pendulum", "unicycle", "cartpole", "quadrotor", "quadrotor_fig8"]
    sizes = {"lti": 4, "pendulum": 3, "unicycle": 5, "cartpole": 16, "quadrotor": 16, "quadrotor_fig8": 4}
    for n in names:
        check(n, sizes.get(n, 4))
