"""CPU tests of the IPDDP oracle (oracle/cddp_oracle.cpp, section "IPDDP"): the constraint known answers the reference's
own tests hold (tests/cddp_core/test_constraint.cpp), the reference's IPDDP solve fixtures (their assertions are
convergence properties, tests/cddp_core/test_ipddp_solver.cpp:349-470, :552-620), KKT / feasibility properties of the
returned solutions, and agreement with the independent numpy restatement (oracle/np_ipddp.py).
PARITY STATUS: unpinned w.r.t. the reference binary (it cannot be built here) — see oracle/cddp_oracle.h."""
import math

import numpy as np
import pytest

from conftest import rel_err

IPDDP_CONFIGS = ["unicycle_obstacle", "unicycle_ipddp_free", "pendulum_ipddp", "pendulum_ipddp_scaled", "cartpole_ipddp", "quadrotor_ipddp",
                 "unicycle_obstacle_teq", "unicycle_teq", "cartpole_teq"]


def _prob(ob, n, m, model="unicycle"):
    spec = dict(model=model, n=n, m=m, horizon=1, dt=0.1, integrator="euler", params=[1.0, 1.0, 0.0], Q=np.zeros((n, n)),
                R=np.eye(m), Qf=np.eye(n), lb=None, ub=None)
    return ob.OracleProblem(spec)


def test_control_and_state_box_kats(ob):
    """ControlConstraintTest/StateConstraintTest.Evaluate and .Jacobians (test_constraint.cpp:22-46, :71-95, :117-170):
    evaluate = [-v; v], upper bound = [-lb; ub], Jacobian = [-I; I] in the constrained variable, zero in the other."""
    P = _prob(ob, 2, 1, "pendulum")
    P2 = ob.OracleProblem(dict(model="unicycle", n=3, m=2, horizon=1, dt=0.1, integrator="euler", params=[], Q=np.zeros((3, 3)),
                               R=np.eye(2), Qf=np.eye(3), lb=None, ub=None))
    lb, ub = np.array([-1.0, -2.0]), np.array([1.0, 2.0])
    cs = ob.ConstraintSet([dict(type="control_box", lb=lb, ub=ub)])
    for u in ([0.5, 1.0], [1.5, -2.5]):
        g, Gx, Gu = ob.eval_constraints(P2, cs, [0.5, 1.0, 0.0], u)
        u = np.asarray(u)
        assert np.allclose(g + np.concatenate([-lb, ub]), np.concatenate([-u, u]), atol=0, rtol=1e-15)
        assert not Gx.any()
        assert np.array_equal(Gu, np.vstack([-np.eye(2), np.eye(2)]))
    cs = ob.ConstraintSet([dict(type="state_box", lb=lb, ub=ub)])
    g, Gx, Gu = ob.eval_constraints(P, cs, [0.5, 1.0], [0.3])
    assert np.allclose(g + np.concatenate([-lb, ub]), [-0.5, -1.0, 0.5, 1.0], rtol=1e-15)
    assert np.array_equal(Gx, np.vstack([-np.eye(2), np.eye(2)])) and not Gu.any()


def test_ball_and_linear_kats(ob):
    """CircleConstraintTest.Evaluate/.Gradients (test_constraint.cpp:173-211): r=2 at the origin, x=(1,1) -> -2 and
    gradient (-2,-2); x=(2.5,1.5) -> -8.5 and (-5,-3).  LinearConstraintTest.Evaluate (:213-234): A=[1 1] -> A x."""
    P = _prob(ob, 2, 1, "pendulum")
    cs = ob.ConstraintSet([dict(type="ball", center=[0.0, 0.0], radius=2.0)])
    for x, val, grad in (([1.0, 1.0], -2.0, [-2.0, -2.0]), ([2.5, 1.5], -8.5, [-5.0, -3.0])):
        g, Gx, Gu = ob.eval_constraints(P, cs, x, [0.0])
        assert abs((g[0] + (-(2.0 ** 2))) - val) < 1e-12  # evaluate() = g + upper, upper = -r^2
        assert np.allclose(Gx[0], grad, rtol=1e-15) and not Gu.any()
    cs = ob.ConstraintSet([dict(type="linear", A=[[1.0, 1.0]], b=[1.0])])
    for x, val in (([0.5, 0.5], 1.0), ([0.5, -0.5], 0.0)):
        g, Gx, Gu = ob.eval_constraints(P, cs, x, [0.0])
        assert abs((g[0] + 1.0) - val) < 1e-15 and np.array_equal(Gx, [[1.0, 1.0]])


def test_constraint_order_follows_the_references_map(ob):
    """The reference keeps path constraints in a std::map keyed by name (cddp_core.hpp:420-423): Ball < Control <
    Linear < State whatever the insertion order."""
    cs = ob.ConstraintSet([dict(type="state_box", lb=[0, 0], ub=[1, 1]), dict(type="control_box", lb=[0], ub=[1]),
                           dict(type="ball", center=[0.0], radius=1.0)])
    assert [c["type"] for c in cs.constraints] == ["ball", "control_box", "state_box"]


def test_reference_ipddp_fixtures_converge(ob):
    """IPDDPTest.SolvePendulum (test_ipddp_solver.cpp:349-470) and IPDDPTest.SolveUnicycle (:552-620, with the default
    sequential line search): the reference asserts status in {Optimal, Acceptable}."""
    N = 500
    spec = dict(model="pendulum", n=2, m=1, horizon=N, dt=0.05, integrator="euler", params=[1.0, 1.0, 0.0], Q=np.zeros((2, 2)),
                R=0.1 * np.eye(1), Qf=100 * np.eye(2), lb=None, ub=None)
    o = ob.make_options(max_iterations=100, tolerance=1e-3, acceptable_tolerance=1e-4, reg_initial_value=1e-6)
    cs = ob.ConstraintSet([dict(type="control_box", lb=[-10.0], ub=[10.0])])
    r = ob.ipddp_solve(ob.OracleProblem(spec), o, ob.make_ipddp_options(), cs, [math.pi, 0.0], [0.0, 0.0], np.zeros((N, 1)))
    assert r["status"] in (1, 2) and np.abs(r["U"]).max() <= 10.0
    N = 100
    spec = dict(model="unicycle", n=3, m=2, horizon=N, dt=0.03, integrator="euler", params=[], Q=np.zeros((3, 3)),
                R=0.5 * np.eye(2), Qf=0.5 * np.diag([50.0, 50.0, 10.0]), lb=None, ub=None)
    o = ob.make_options(max_iterations=20, tolerance=1e-2)
    cs = ob.ConstraintSet([dict(type="control_box", lb=[-1.0, -math.pi], ub=[1.0, math.pi])])
    r = ob.ipddp_solve(ob.OracleProblem(spec), o, ob.make_ipddp_options(), cs, [0, 0, math.pi / 4], [2, 2, math.pi / 2],
                       np.zeros((N, 2)))
    assert r["status"] in (1, 2)
    assert np.all(np.abs(r["U"]) <= np.array([1.0, math.pi]) + 1e-12)


@pytest.mark.parametrize("name", IPDDP_CONFIGS)
def test_solution_properties(ob, problems, name):
    """Interior-point invariants of every returned solution: slacks and duals strictly positive, primal residual
    g + s = inf_pr small when converged, complementarity |y s - mu| <= inf_comp, the state trajectory is a rollout of the
    controls, the reported cost is the cost of the returned trajectory, and the barrier parameter never grows."""
    B = 3
    cfg = problems.make_config(name, batch=B)
    P, oo, oi = ob.OracleProblem(cfg["spec"]), ob.make_options(**cfg["options"]), ob.make_ipddp_options(**cfg.get("ipddp_options", {}))
    cs = ob.ConstraintSet(cfg["constraints"])
    for b in range(B):
        r = ob.ipddp_solve(P, oo, oi, cs, cfg["x0"][b], cfg["xref"][b], cfg["U0"][b], history=True)
        if cfg.get("ipddp_options", {}).get("terminal_equality") and r["status"] in (1, 2):
            assert np.abs(r["X"][-1] - cfg["xref"][b]).max() < 1e-3  # TerminalEqualityConstraint: x_N = goal
        X, U = r["X"], r["U"]
        for t in range(0, P.N, 7):
            assert rel_err(ob.discrete_dynamics(P, X[t], U[t], t * cfg["spec"]["dt"]), X[t + 1]) < 1e-12
        assert abs(ob.trajectory_cost(P, X, U, cfg["xref"][b]) - r["cost"]) <= 1e-12 * abs(r["cost"])
        h = r["history"]
        assert np.all(np.diff(h[:, 8]) <= 0.0), "barrier parameter must be non-increasing"
        if cs.nc:
            assert (r["S"] > 0).all() and (r["Y"] > 0).all()
            G = np.array([ob.eval_constraints(P, cs, X[t], U[t])[0] for t in range(P.N)])
            assert np.abs(G + r["S"]).max() <= r["inf_pr"] * (1 + 1e-9) + 1e-15
            assert np.abs(r["Y"] * r["S"] - r["mu"]).max() <= r["inf_comp"] * (1 + 1e-9) + 1e-15
            if r["status"] in (1, 2):
                assert G.max() < 1e-3, "converged solutions are feasible"
        assert 0.0 <= r["decision_margin"] <= 1.0


@pytest.mark.parametrize("name", IPDDP_CONFIGS)
def test_oracle_vs_numpy_restatement(ob, problems, name):
    """The C++ oracle and the independently written numpy restatement (oracle/np_ipddp.py: complex-step Jacobians,
    numpy.linalg.solve) agree step by step (backward-pass gains, step caps, line-search table) and over whole solves
    on instances whose line-search decisions are not roundoff-decided."""
    np_ipddp = pytest.importorskip("np_ipddp")
    B = 2
    cfg = problems.make_config(name, batch=B, horizon=40 if not name.startswith("unicycle_obstacle") else 60)
    ipo = cfg.get("ipddp_options", {})
    P, oo, oi = ob.OracleProblem(cfg["spec"]), ob.make_options(**cfg["options"]), ob.make_ipddp_options(**ipo)
    cs = ob.ConstraintSet(cfg["constraints"])
    for b in range(B):
        for iters in (0, 2):
            r = ob.ipddp_probe(P, oo, oi, cs, cfg["x0"][b], cfg["xref"][b], cfg["U0"][b], iters)
            q = np_ipddp.probe(cfg["spec"], cfg["options"], ipo, cs.constraints, cfg["x0"][b], cfg["xref"][b], cfg["U0"][b], iters)
            for key in ("X", "U", "Y", "S", "G", "ku", "Ku", "ky", "Ky", "ks", "Ks"):
                assert rel_err(r[key], q[key]) < 1e-7, (name, b, iters, key)
            for key in ("mu", "cost", "merit", "inf_du", "step_norm", "dV0", "dV1", "alpha_pr_max", "alpha_du_max"):
                assert abs(r[key] - q[key]) <= 1e-7 * max(abs(q[key]), 1e-12), (name, b, iters, key, r[key], q[key])
        r = ob.ipddp_solve(P, oo, oi, cs, cfg["x0"][b], cfg["xref"][b], cfg["U0"][b])
        if r["decision_margin"] > 1e-6:
            q = np_ipddp.solve(cfg["spec"], cfg["options"], ipo, cs.constraints, cfg["x0"][b], cfg["xref"][b], cfg["U0"][b])
            assert q["iterations"] == r["iterations"] and q["status"] == r["status"]
            assert abs(q["cost"] - r["cost"]) <= 1e-6 * abs(r["cost"])
