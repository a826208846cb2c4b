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
