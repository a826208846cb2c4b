"""The oracle's decision-trace / decision-replay / one-iteration entry points (test instrumentation used by
tests/test_gpu_every_instance.py) checked against the oracle's own free-running solve on CPU: replaying a solve's own
trace must reproduce it bit for bit with zero disagreements, and chaining oracle_iterate_batch must equal the solve."""
import numpy as np
import pytest


@pytest.mark.parametrize("name,B,N", [("quadrotor", 6, 40), ("cartpole", 8, None), ("pendulum", 2, 120), ("unicycle", 4, None)])
def test_replay_of_own_trace_is_identity(ob, problems, name, B, N):
    cfg = problems.make_config(name, batch=B, horizon=N)
    opts = dict(cfg["options"], max_iterations=min(cfg["options"]["max_iterations"], 30))
    P, oo = ob.OracleProblem(cfg["spec"]), ob.make_options(**opts)
    base = ob.solve_batch(P, oo, cfg["x0"], cfg["xref"], cfg["X0"], cfg["U0"], cfg["ref_traj"], nthreads=2)
    t = ob.solve_batch_traced(P, oo, cfg["x0"], cfg["xref"], cfg["X0"], cfg["U0"], cfg["ref_traj"], nthreads=2)
    for key in ("cost", "iterations", "status", "X", "U", "K"):
        np.testing.assert_array_equal(base[key], t[key])
    r = ob.solve_batch_traced(P, oo, cfg["x0"], cfg["xref"], cfg["X0"], cfg["U0"], cfg["ref_traj"], nthreads=2,
                              replay=dict(trace=t["trace"], iterations=t["iterations"], status=t["status"]))
    rep = r["replay"]
    assert (rep["n_disagree"] == 0).all() and (rep["n_backward_disagree"] == 0).all() and (rep["infeasible"] == 0).all()
    for key in ("cost", "iterations", "status", "X", "U", "K", "reg", "alpha"):
        np.testing.assert_array_equal(t[key], r[key])
    np.testing.assert_array_equal(t["history"], r["history"])


def test_iterate_batch_chain_equals_solve(ob, problems):
    B = 5
    cfg = problems.make_config("quadrotor", batch=B, horizon=30)
    opts = dict(cfg["options"], max_iterations=9)
    P, oo = ob.OracleProblem(cfg["spec"]), ob.make_options(**opts)
    t = ob.solve_batch_traced(P, oo, cfg["x0"], cfg["xref"], cfg["X0"], cfg["U0"], nthreads=2)
    X, U = cfg["X0"].copy(), cfg["U0"].copy()
    X[:, 0] = cfg["x0"]
    k = np.zeros((B, 30, 4))
    reg = np.full(B, opts["reg_initial_value"])
    cost = np.array([ob.trajectory_cost(P, X[b], U[b], cfg["xref"][b]) for b in range(B)])
    alpha, inf_du = np.ones(B), np.full(B, np.inf)
    status = np.zeros(B, dtype=np.int32)
    for it in range(9):
        run = np.flatnonzero(status == 0)
        if run.size == 0:
            break
        o = ob.iterate_batch(P, oo, cfg["x0"][run], cfg["xref"][run], X[run], U[run], k[run], reg[run], cost[run], alpha[run],
                             inf_du[run])
        np.testing.assert_array_equal(o["code"], t["trace"][run, it])
        X[run], U[run], k[run], reg[run], cost[run], alpha[run], inf_du[run] = (o[q] for q in ("X", "U", "k", "reg", "cost", "alpha", "inf_du"))
        status[run] = o["status"]
        # the same iteration again, FOLLOWING the recorded decision, from the pre-iteration state gives the same result
    np.testing.assert_array_equal(cost, t["cost"])
    np.testing.assert_array_equal(X, t["X"])
    fin = status != 0
    np.testing.assert_array_equal(status[fin], t["status"][fin])


def test_replay_reports_a_forced_wrong_decision(ob, problems):
    """Following a decision sequence that is NOT the oracle's own (second alpha forced where the first passes) is
    reported as a disagreement with a margin far from roundoff."""
    cfg = problems.make_config("unicycle", batch=2, horizon=40)
    opts = dict(cfg["options"], max_iterations=4)
    P, oo = ob.OracleProblem(cfg["spec"]), ob.make_options(**opts)
    t = ob.solve_batch_traced(P, oo, cfg["x0"], cfg["xref"], cfg["X0"], cfg["U0"])
    tr = t["trace"].copy()
    assert (tr[:, 0] & 0xFF == 1).all()
    tr[:, 0] = (tr[:, 0] & ~0xFF) | 2
    r = ob.solve_batch_traced(P, oo, cfg["x0"], cfg["xref"], cfg["X0"], cfg["U0"],
                              replay=dict(trace=tr, iterations=np.ones(2, dtype=np.int32), status=np.zeros(2, dtype=np.int32)))
    assert (r["replay"]["n_disagree"] >= 1).all() and (r["replay"]["max_margin"] > 1e-6).all()


@pytest.mark.parametrize("name,B", [("unicycle_obstacle_teq", 4), ("unicycle_obstacle", 3), ("pendulum_ipddp", 3)])
def test_ipddp_iterate_batch_chain_equals_solve(ob, problems, name, B):
    """oracle_ipddp_iterate_batch chained from the cold-start state reproduces oracle_ipddp_solve_batch bit for bit
    (iterations, status, cost, trajectory); following its own decisions reports no disagreement."""
    cfg = problems.make_config(name, batch=B, horizon=60)
    opts = dict(cfg["options"], max_iterations=25)
    P, oo = ob.OracleProblem(cfg["spec"]), ob.make_options(**opts)
    oi, cs = ob.make_ipddp_options(**cfg.get("ipddp_options", {})), ob.ConstraintSet(cfg["constraints"])
    ref = ob.ipddp_solve_batch(P, oo, oi, cs, cfg["x0"], cfg["xref"], cfg["U0"], nthreads=2)
    # cold-start state = what IPDDPSolver::initialize leaves (probe with 0 iterations gives X, Y, S, G and the scalars)
    d = cs.dual_dim(P.n, P.m)
    st = dict(X=[], U=[], Y=[], S=[], G=[], lamT=np.zeros((B, P.n)), filter=np.zeros((B, 8, 2)), filter_size=np.zeros(B, dtype=np.int32))
    sc = {k: np.zeros(B) for k in ob.IP_STATE_SCALARS}
    for b in range(B):
        r = ob.ipddp_probe(P, oo, oi, cs, cfg["x0"][b], cfg["xref"][b], cfg["U0"][b], 0)
        for k in ("X", "U", "Y", "S", "G"):
            st[k].append(r[k])
        sc["mu"][b], sc["cost"][b], sc["merit"][b], sc["filter_theta"][b] = r["mu"], r["cost"], r["merit"], r["filter_theta"]
        sc["reg"][b] = opts.get("reg_initial_value", 1e-6)
        sc["alpha_pr"][b], sc["alpha_du"][b] = 1.0, 1.0
        if cfg.get("ipddp_options", {}).get("terminal_equality"):
            st["filter"][b, 0] = (r["merit"], r["filter_theta"])  # resetBarrierFilter seeds the filter (ipddp_solver.cpp:2513-2516)
            st["filter_size"][b] = 1
    for k in ("X", "U", "Y", "S", "G"):
        st[k] = np.stack(st[k])
    st.update(sc)
    status = np.zeros(B, dtype=np.int32)
    iters = np.zeros(B, dtype=np.int32)
    for it in range(1, 26):
        run = np.flatnonzero(status == 0)
        if run.size == 0:
            break
        sub = {k: (v[run] if isinstance(v, np.ndarray) else v) for k, v in st.items()}
        sub["iter"] = np.full(run.size, float(it))
        o = ob.ipddp_iterate_batch(P, oo, oi, cs, cfg["x0"][run], cfg["xref"][run], sub)
        assert (o["n_disagree"] == 0).all()
        for k in st:
            if k != "iter":
                st[k][run] = o[k]
        status[run] = o["status"]
        iters[run] = it
    done = status != 0
    np.testing.assert_array_equal(iters[done], ref["iterations"][done])
    np.testing.assert_array_equal(status[done], ref["status"][done])
    np.testing.assert_array_equal(st["cost"], ref["cost"])
    np.testing.assert_array_equal(st["X"], ref["X"])
