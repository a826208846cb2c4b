import importlib, os, sys
import numpy as np
ROOT = "/root/repo"
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import oracle_binding as ob
cddp = importlib.import_module("cddp-cpp_b200"); problems = importlib.import_module("cddp-cpp_b200.problems")
cols = ["cost","merit","a_pr","a_du","inf_du","inf_pr","inf_comp","reg","mu"]
for name in sys.argv[1:]:
    cfg = problems.make_config(name, batch=6)
    spec, cons = cfg["spec"], cfg["constraints"]
    P = ob.OracleProblem(spec); oo = ob.make_options(**cfg["options"]); oi = ob.make_ipddp_options(); ocs = ob.ConstraintSet(cons)
    s = cddp.BatchedIPDDP(spec, cddp.default_options(**cfg["options"]), cddp.default_ipddp_options(), cons, 6)
    s.enable_history(True)
    s.set_instances(cfg["x0"], cfg["xref"], None, cfg["U0"], cfg["ref_traj"])
    s.solve()
    h, hl = s.get_history()
    for b in range(6):
        r0 = ob.ipddp_solve(P, oo, oi, ocs, cfg["x0"][b], cfg["xref"][b], cfg["U0"][b], history=True)
        with ob.variant():
            r1 = ob.ipddp_solve(P, oo, oi, ocs, cfg["x0"][b], cfg["xref"][b], cfg["U0"][b], history=True)
        L = min(hl[b], len(r0["history"]))
        dev = np.abs(h[b, :L] - r0["history"][:L]) / (np.abs(r0["history"][:L]) + 1e-9)
        bad = [i for i in range(L) if dev[i].max() > 1e-6]
        L1 = min(len(r1["history"]), len(r0["history"]))
        dev1 = np.abs(r1["history"][:L1] - r0["history"][:L1]) / (np.abs(r0["history"][:L1]) + 1e-9)
        bad1 = [i for i in range(L1) if dev1[i].max() > 1e-6]
        print(f"[{name} b={b}] len gpu {hl[b]} cpu {len(r0['history'])} fma {len(r1['history'])}; gpu-vs-cpu first bad row {bad[:1]}; fma-vs-strict first bad row {bad1[:1]}")
        if bad:
            i = bad[0]
            print("   cols:", {c: f"{h[b,i,k]:.6e}/{r0['history'][i,k]:.6e}" for k, c in enumerate(cols) if dev[i,k] > 1e-6})
