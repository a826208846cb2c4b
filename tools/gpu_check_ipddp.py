"""Developer script: step-by-step comparison of the CUDA IPDDP path with the CPU oracle on small batches
(the pytest -m gpu suite holds the judged version)."""
import importlib
import math
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import oracle_binding as ob  # noqa: E402

cddp = importlib.import_module("cddp-cpp_b200")
problems = importlib.import_module("cddp-cpp_b200.problems")


def rel(a, b):
    a, b = np.asarray(a), np.asarray(b)
    if a.size == 0:
        return 0.0
    return float(np.max(np.abs(a - b)) / (np.max(np.abs(b)) + 1e-300))


def check(name, B, iters=(0, 1, 3)):
    cfg = problems.make_config(name, batch=B)
    spec, cons = cfg["spec"], cfg["constraints"]
    P = ob.OracleProblem(dict(spec, lb=None, ub=None))
    oo = ob.make_options(**cfg["options"])
    oi = ob.make_ipddp_options(**cfg.get("ipddp_options", {}))
    ocs = ob.ConstraintSet(cons)
    go = cddp.default_options(**cfg["options"])
    gi = cddp.default_ipddp_options(**cfg.get("ipddp_options", {}))
    s = cddp.BatchedIPDDP(spec, go, gi, cons, B)
    for it in iters:
        s.set_instances(cfg["x0"], cfg["xref"], None, cfg["U0"], cfg["ref_traj"])
        s.initialize()
        if it:
            s.iterate(it)
        s.linearize()
        s.backward_pass()
        s.forward_pass()
        sol, ips, gains, sw, kff, ls = s.get_solution(), s.get_ipddp_solution(), s.get_ipddp_gains(), s.get_sweep(), s.get_feedforward(), s.get_line_search()
        for b in range(min(B, 2)):
            r = ob.ipddp_probe(P, oo, oi, ocs, cfg["x0"][b], cfg["xref"][b], cfg["U0"][b], it)
            print(f"[{name} it={it} b={b}] X {rel(sol['X'][b], r['X']):.1e} U {rel(sol['U'][b], r['U']):.1e} Y {rel(ips['Y'][b], r['Y']):.1e} "
                  f"S {rel(ips['S'][b], r['S']):.1e} G {rel(ips['G'][b], r['G']):.1e} | ku {rel(kff[b], r['ku']):.1e} Ku {rel(sol['K'][b], r['Ku']):.1e} "
                  f"ky {rel(gains['ky'][b], r['ky']):.1e} Ky {rel(gains['Ky'][b], r['Ky']):.1e} ks {rel(gains['ks'][b], r['ks']):.1e} Ks {rel(gains['Ks'][b], r['Ks']):.1e}")
            print(f"    mu {ips['mu'][b]:.6e}/{r['mu']:.6e} cost {sol['cost'][b]:.10e}/{r['cost']:.10e} merit {ips['merit'][b]:.10e}/{r['merit']:.10e} "
                  f"inf_pr {ips['inf_pr'][b]:.3e}/{r['inf_pr']:.3e} inf_du {sol['inf_du'][b]:.3e}/{r['inf_du']:.3e} inf_comp {ips['inf_comp'][b]:.3e}/{r['inf_comp']:.3e} "
                  f"step {ips['step_norm'][b]:.3e}/{r['step_norm']:.3e} reg {sol['reg'][b]:.1e}/{r['reg']:.1e} dV {sw['dV'][b]} / {r['dV0']:.8e},{r['dV1']:.8e} "
                  f"apm {ips['alpha_pr_max'][b]:.6e}/{r['alpha_pr_max']:.6e} adm {ips['alpha_du_max'][b]:.6e}/{r['alpha_du_max']:.6e} ok {sw['ok'][b]}/{r['bw_ok']}")
            print(f"    line search: accept gpu {ls[b][:, 0].astype(int)} cpu {r['trial'][:, 0].astype(int)} cost rel {rel(ls[b][:, 1], r['trial'][:, 1]):.1e} "
                  f"merit rel {rel(ls[b][:, 2], r['trial'][:, 2]):.1e} theta rel {rel(ls[b][:, 3], r['trial'][:, 3]):.1e}")
    # whole solve
    s.enable_history(True)
    s.set_instances(cfg["x0"], cfg["xref"], None, cfg["U0"], cfg["ref_traj"])
    s.solve()
    g = s.get_solution()
    gi_ = s.get_ipddp_solution(False)
    h, hl = s.get_history()
    o = ob.ipddp_solve_batch(P, oo, oi, ocs, cfg["x0"], cfg["xref"], cfg["U0"], cfg["ref_traj"], nthreads=4)
    rc = np.abs(g["cost"] - o["cost"]) / np.abs(o["cost"])
    print(f"[{name}] solve: iters gpu {g['iterations'][:8]} cpu {o['iterations'][:8]} status gpu {g['status'][:8]} cpu {o['status'][:8]}")
    print(f"[{name}] solve: cost rel err max {rc.max():.2e}  same iters {np.mean(g['iterations'] == o['iterations']):.2f}  mu gpu {gi_['mu'][:4]} cpu {o['mu'][:4]}")
    r0 = ob.ipddp_solve(P, oo, oi, ocs, cfg["x0"][0], cfg["xref"][0], cfg["U0"][0], history=True)
    L = min(hl[0], len(r0["history"]))
    dev = np.abs(h[0, :L] - r0["history"][:L]) / (np.abs(r0["history"][:L]) + 1e-12)
    print(f"[{name}] history b=0: len gpu {hl[0]} cpu {len(r0['history'])}; first row with rel dev > 1e-6: "
          f"{next((i for i in range(L) if dev[i].max() > 1e-6), None)}; max dev {dev.max():.2e}")
    s.close()


if __name__ == "__main__":
    names = sys.argv[1:] or ["unicycle_obstacle"]
    for nm in names:
        check(nm, 6)
