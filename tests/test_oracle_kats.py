"""Pins the CPU oracle against every known-answer the reference's own tests hold for the hot path
(SURVEY.md §8c).  The reference asserts no gains/trajectories/costs on this path, so these are the only
reference-held numbers there are; everything else is cross-restatement agreement (test_oracle_vs_numpy.py)."""
import math

import numpy as np
import pytest

from conftest import golden, rel_err


def _spec(model, n, m, dt=0.1, integrator="euler", params=(), **kw):
    s = dict(model=model, n=n, m=m, horizon=5, dt=dt, integrator=integrator, params=list(params),
             Q=np.zeros((n, n)), R=np.eye(m), Qf=np.zeros((n, n)), lb=None, ub=None)
    s.update(kw)
    return s


def test_quadratic_objective_identities(ob):
    """tests/cddp_core/test_objective.cpp:39-128: cost = sum (e'Qe + u'Ru)*dt + e_N'Qf e_N; gradients 2Q e dt, 2R u dt."""
    n, m, dt = 3, 2, 0.1
    Q, R, Qf = np.eye(n), 0.1 * np.eye(m), 2.0 * np.eye(n)
    goal = np.array([1.1, 0.6, 0.3])
    spec = dict(model="lti", n=n, m=m, horizon=5, dt=dt, integrator="euler", params=[], lti_A=np.eye(n),
                lti_B=np.zeros((n, m)), Q=Q, R=R, Qf=Qf, lb=None, ub=None)
    P = ob.OracleProblem(spec)
    X = np.array([[1.0, 0.5, 0.2], [1.1, 0.6, 0.3], [1.2, 0.7, 0.4], [1.3, 0.8, 0.5], [1.4, 0.9, 0.6], [1.5, 1.0, 0.7]])
    U = np.tile(np.array([0.8, 0.5]), (5, 1))
    expected = sum(((X[i] - goal) @ Q @ (X[i] - goal)) * dt + (U[i] @ R @ U[i]) * dt for i in range(5))
    expected += (X[5] - goal) @ Qf @ (X[5] - goal)
    assert abs(ob.trajectory_cost(P, X, U, goal) - expected) < 1e-6
    assert abs(ob.terminal_cost(P, X[0], goal) - (X[0] - goal) @ Qf @ (X[0] - goal)) < 1e-6
    assert abs(ob.running_cost(P, X[0], U[0], goal) - (((X[0] - goal) @ Q @ (X[0] - goal)) + U[0] @ R @ U[0]) * dt) < 1e-12
    # gradients/Hessians enter the sweep: with V = 0 terminal (Qf=0), A=I, B=0 => Q_x = l_x = 2 Q e dt at t=N-1
    spec0 = dict(spec, Qf=np.zeros((n, n)))
    P0 = ob.OracleProblem(spec0)
    r = ob.backward_pass(P0, ob.make_options(), X, U, goal, 1e-6, debug=True)
    assert r["ok"]
    np.testing.assert_allclose(r["Vx"][4], 2.0 * Q @ (X[4] - goal) * dt, atol=1e-12)
    np.testing.assert_allclose(r["Vxx"][4], 2.0 * Q * dt, atol=1e-12)


def test_control_constraint_clamp(ob):
    """tests/cddp_core/test_constraint.cpp:48-69: clamp([1.5,-2.5]) with bounds +-[1,2] = [1,-2], seen through the rollout."""
    n, m = 2, 2
    spec = dict(model="lti", n=n, m=m, horizon=1, dt=1.0, integrator="euler", params=[], lti_A=np.eye(n), lti_B=np.eye(n),
                Q=np.zeros((n, n)), R=np.eye(m), Qf=np.zeros((n, n)), lb=[-1.0, -2.0], ub=[1.0, 2.0])
    P = ob.OracleProblem(spec)
    o = ob.make_options()
    x0 = np.zeros(2)
    X, U = np.zeros((2, 2)), np.array([[1.5, -2.5]])
    f = ob.forward_pass(P, o, x0, X, U, np.zeros(2), np.zeros((1, 2, 2)), np.zeros((1, 2)), np.zeros(2), 0.0, 1.0)
    np.testing.assert_allclose(f["U"][0], [1.0, -2.0])
    U2 = np.array([[0.5, 1.0]])
    f = ob.forward_pass(P, o, x0, X, U2, np.zeros(2), np.zeros((1, 2, 2)), np.zeros((1, 2)), np.zeros(2), 0.0, 1.0)
    np.testing.assert_allclose(f["U"][0], [0.5, 1.0])


def _fd_jac(f, z, h=2e-5):
    """central differences, h = 2e-5 (include/cddp-cpp/cddp_core/helper.hpp:96-147)"""
    f0 = f(z)
    J = np.zeros((len(f0), len(z)))
    for j in range(len(z)):
        zp, zm = z.copy(), z.copy()
        zp[j] += h
        zm[j] -= h
        J[:, j] = (f(zp) - f(zm)) / (2 * h)
    return J


def test_pendulum_jacobian_closed_form(ob):
    """tests/cddp_core/test_finite_difference.cpp:27-69 (A,B vs central FD, 1e-6) and pendulum.cpp:45-66."""
    P = ob.OracleProblem(_spec("pendulum", 2, 1, dt=0.05, params=[1.0, 1.0, 0.0]))
    x, u = np.array([0.1, 0.0]), np.array([0.0])
    Fx, Fu = ob.jacobians(P, x, u)
    np.testing.assert_allclose(Fx, [[0.0, 1.0], [9.81 * math.cos(0.1), 0.0]], atol=1e-14)
    np.testing.assert_allclose(Fu, [[0.0], [1.0]], atol=1e-14)
    A = _fd_jac(lambda z: ob.continuous_dynamics(P, z, u), x)
    Bm = _fd_jac(lambda z: ob.continuous_dynamics(P, x, z), u)
    assert np.allclose(Fx, A, rtol=1e-6, atol=1e-8) and np.allclose(Fu, Bm, rtol=1e-6, atol=1e-8)


QUAD = [1.0, 0.01, 0, 0, 0, 0.01, 0, 0, 0, 0.02, 0.2]


def test_quadrotor_hover_and_jacobians(ob):
    """tests/dynamics_model/test_quadrotor.cpp:166-212 (hover => f = 0 to 1e-10) and :223-289, :291-360
    (Jacobians vs central FD <= 1e-4 Frobenius at hover and at the tilted state)."""
    P = ob.OracleProblem(_spec("quadrotor", 13, 4, dt=0.01, params=QUAD))
    x = np.zeros(13)
    x[2], x[3] = 1.0, 1.0
    hover = 9.81 / 4.0
    u = np.full(4, hover)
    xd = ob.continuous_dynamics(P, x, u)
    assert np.abs(xd).max() < 1e-10
    u2 = u.copy()
    u2[0] += 0.1
    u2[2] -= 0.1
    assert abs(ob.continuous_dynamics(P, x, u2)[10]) > 0.0
    x2 = x.copy()
    x2[4:7] = 0.1
    x2[7:10] = 0.2
    x2[10:13] = 0.1
    x2[3:7] /= np.linalg.norm(x2[3:7])
    for xs in (x, x2):
        Fx, Fu = ob.jacobians(P, xs, u)
        A = _fd_jac(lambda z: ob.continuous_dynamics(P, z, u), xs)
        Bm = _fd_jac(lambda z: ob.continuous_dynamics(P, xs, z), u)
        assert np.linalg.norm(Fx - A) < 1e-4
        assert np.linalg.norm(Fu - Bm) < 1e-4


def test_all_model_jacobians_vs_complex_step(ob, npo):
    """Exact-derivative check (the reference's autodiff is exact): <= 1e-10 against complex-step."""
    rng = np.random.default_rng(7)
    cases = [("pendulum", 2, 1, [0.7, 1.3, 0.05]), ("cartpole", 4, 1, [1.0, 0.2, 0.5, 9.81, 0.3]), ("unicycle", 3, 2, []),
             ("quadrotor", 13, 4, [1.2, 0.01, 0.001, 0, 0.001, 0.012, 0, 0, 0, 0.02, 0.25])]
    for model, n, m, params in cases:
        spec = _spec(model, n, m, params=params)
        P, Pn = ob.OracleProblem(spec), npo.Problem(spec)
        for _ in range(5):
            x, u = rng.standard_normal(n), rng.standard_normal(m)
            Fx, Fu = ob.jacobians(P, x, u)
            Gx, Gu = Pn.jacobians(x, u)
            assert np.abs(Fx - Gx).max() < 1e-10 * max(1.0, np.abs(Gx).max()), model
            assert np.abs(Fu - Gu).max() < 1e-10 * max(1.0, np.abs(Gu).max()), model


def test_cartpole_damping_only_in_jacobian(ob):
    """cartpole.cpp:60 vs :90 — the rollout ignores damping, the (autodiff) Jacobian includes it."""
    a = ob.OracleProblem(_spec("cartpole", 4, 1, params=[1.0, 0.2, 0.5, 9.81, 0.0]))
    b = ob.OracleProblem(_spec("cartpole", 4, 1, params=[1.0, 0.2, 0.5, 9.81, 0.7]))
    x, u = np.array([0.1, 0.4, -0.2, 0.9]), np.array([0.3])
    np.testing.assert_array_equal(ob.continuous_dynamics(a, x, u), ob.continuous_dynamics(b, x, u))
    assert abs(ob.jacobians(a, x, u)[0][3, 3] - ob.jacobians(b, x, u)[0][3, 3]) > 1e-3


def test_integrators_order(ob):
    """dynamical_system.cpp:28-65: euler/heun/rk3/rk4 on the pendulum converge at orders 1/2/3/4."""
    errs = {}
    for integ in ("euler", "heun", "rk3", "rk4"):
        e = []
        for dt in (0.02, 0.01):
            P = ob.OracleProblem(_spec("pendulum", 2, 1, dt=dt, integrator=integ, params=[1.0, 1.0, 0.1]))
            Pf = ob.OracleProblem(_spec("pendulum", 2, 1, dt=dt / 64, integrator="rk4", params=[1.0, 1.0, 0.1]))
            x, u = np.array([0.5, 0.2]), np.array([0.3])
            xr = x.copy()
            for _ in range(64):
                xr = ob.discrete_dynamics(Pf, xr, u)
            e.append(np.abs(ob.discrete_dynamics(P, x, u) - xr).max())
        errs[integ] = math.log2(e[0] / e[1])
    assert 1.7 < errs["euler"] < 2.3 and 2.7 < errs["heun"] < 3.3 and 3.6 < errs["rk3"] < 4.4 and 4.5 < errs["rk4"] < 5.5


def test_alpha_schedule(ob, npo):
    """cddp_context_utils.cpp:37-57 with options.hpp:43-49 defaults: 1, 1/2, ..., 2^-10 (11 entries)."""
    a = ob.build_alphas(ob.make_options())
    np.testing.assert_array_equal(a, 0.5 ** np.arange(11))
    o = ob.make_options(ls_max_iterations=40, ls_min_step_size=1e-3)
    a = ob.build_alphas(o)
    assert a[-1] == 1e-3 and a[-2] >= 1e-3 and len(a) == 11
    np.testing.assert_array_equal(a, npo.build_alphas(npo.options(ls_max_iterations=40, ls_min_step_size=1e-3)))


def test_boxqp_reference_fixtures(ob):
    """The reference's BoxQP test inputs (test_boxqp.cpp:59-64,89-90; :125-220 — that test only prints):
    the oracle's answer must satisfy the box-QP KKT conditions and match the independent numpy solution."""
    g = golden("boxqp_fixtures.npz")
    o = ob.make_options()
    for tag in ("5", "15"):
        H, q, lo, hi = g["H" + tag], g["g" + tag], g["lo" + tag], g["hi" + tag]
        r = ob.boxqp(o, H, q, lo, hi, None)
        assert r["status"] == int(g["status" + tag]) == 4  # SUCCESS
        np.testing.assert_allclose(r["x"], g["x" + tag], rtol=1e-10, atol=1e-12)
        np.testing.assert_array_equal(r["free"], g["free" + tag])
        assert abs(r["value"] - float(g["value" + tag])) < 1e-12
        grad = H @ r["x"] + q
        for i in range(len(q)):
            if r["x"][i] == lo[i]:
                assert grad[i] > -1e-6
            elif r["x"][i] == hi[i]:
                assert grad[i] < 1e-6
            else:
                assert abs(grad[i]) < 1e-6
    # 5x5 closed form: unconstrained minimiser clipped coordinates -> value -16/3
    assert abs(float(g["value5"]) + 16.0 / 3.0) < 1e-9


def test_boxqp_edge_cases(ob):
    o = ob.make_options()
    H = np.array([[2.0, 0.0], [0.0, 2.0]])
    # all clamped: gradient pushes both coordinates out of the box
    r = ob.boxqp(o, H, np.array([10.0, 10.0]), np.array([0.0, 0.0]), np.array([1.0, 1.0]), np.zeros(2))
    assert r["status"] == 5 and (r["free"] == 0).all() and (r["x"] == 0).all()
    # interior optimum: one Newton step
    r = ob.boxqp(o, H, np.array([-1.0, -1.0]), np.array([-5.0, -5.0]), np.array([5.0, 5.0]), np.zeros(2))
    assert r["status"] == 4 and np.allclose(r["x"], [0.5, 0.5])
    # indefinite Hessian on the free block
    r = ob.boxqp(o, np.array([[1.0, 3.0], [3.0, 1.0]]), np.array([-1.0, 1.0]), -np.ones(2), np.ones(2), np.zeros(2))
    assert r["status"] in (-1, 0)  # HESSIAN_NOT_PD (Cholesky) / NO_DESCENT


def test_lti_lqr_closed_form(ob, problems):
    """LTISystem returns (A_d - I)/dt, B_d/dt (lti_system.cpp:78-92) so the sweep reconstructs A_d, B_d and the
    gains equal the textbook finite-horizon discrete LQR with stage cost x'(2Q dt)x/2... i.e. Riccati on
    (Q_ = Q dt, R_ = R dt, Qf); converges in one accepted full step."""
    cfg = problems.make_config("lti", batch=2, horizon=30)
    spec = cfg["spec"]
    Ad, Bd, dt = spec["lti_A"], spec["lti_B"], spec["dt"]
    Qs, Rs, Qf = spec["Q"] * dt, spec["R"] * dt, spec["Qf"]
    N = spec["horizon"]
    S = Qf.copy()
    Ks = [None] * N
    for t in range(N - 1, -1, -1):
        Kt = -np.linalg.solve(Rs + Bd.T @ S @ Bd, Bd.T @ S @ Ad)
        S = Qs + Ad.T @ S @ Ad + Ad.T @ S @ Bd @ Kt
        S = 0.5 * (S + S.T)
        Ks[t] = Kt
    P = ob.OracleProblem(spec)
    o = ob.make_options(**dict(cfg["options"], reg_initial_value=0.0, reg_min_value=0.0))
    r = ob.solve(P, o, cfg["x0"][0], cfg["xref"][0], cfg["X0"][0], cfg["U0"][0])
    assert r["status"] == 1 and r["iterations"] <= 3 and r["alpha"] == 1.0
    assert rel_err(r["K"], np.stack(Ks)) < 1e-10
    # optimal cost = x0' S x0
    assert abs(r["cost"] - cfg["x0"][0] @ S @ cfg["x0"][0]) < 1e-9 * abs(r["cost"])


def test_solver_outcomes_are_statuses(ob, problems):
    """Outcomes are status strings, never errors (cddp_solver_base.cpp:69,82,162; clddp_solver.cpp:209,270,274)."""
    assert ob.load().oracle_status_string(1) == b"OptimalSolutionFound"
    assert ob.load().oracle_status_string(2) == b"AcceptableSolutionFound"
    assert ob.load().oracle_status_string(3) == b"MaxIterationsReached"
    assert ob.load().oracle_status_string(4) == b"RegularizationLimitReached_NotConverged"
    cfg = problems.make_config("pendulum", batch=1, horizon=50)
    P = ob.OracleProblem(cfg["spec"])
    r = ob.solve(P, ob.make_options(max_iterations=1), cfg["x0"][0], cfg["xref"][0], cfg["X0"][0], cfg["U0"][0])
    assert r["status"] == 3 and r["iterations"] == 1
    # tiny reg_max => the first rejected line search hits the limit
    o = ob.make_options(max_iterations=20, reg_initial_value=1.0, reg_max_value=5.0, armijo_constant=1e9)
    r = ob.solve(P, o, cfg["x0"][0], cfg["xref"][0], cfg["X0"][0], cfg["U0"][0])
    assert r["status"] == 4 and r["iterations"] == 1


@pytest.mark.parametrize("name", ["pendulum", "unicycle", "quadrotor_fig8"])
def test_reference_solver_test_properties(ob, problems, name):
    """What the reference's own CLDDP tests assert (test_clddp_solver.cpp:149-151,297,752-763): status in
    {Optimal, Acceptable}, iterations > 0, final cost < initial cost."""
    cfg = problems.make_config(name, batch=1, horizon=100 if name != "pendulum" else 500)
    P = ob.OracleProblem(cfg["spec"])
    o = ob.make_options(**cfg["options"])
    rt = None if cfg["ref_traj"] is None else cfg["ref_traj"][0]
    c0 = ob.trajectory_cost(P, cfg["X0"][0], cfg["U0"][0], cfg["xref"][0], rt)
    r = ob.solve(P, o, cfg["x0"][0], cfg["xref"][0], cfg["X0"][0], cfg["U0"][0], rt)
    assert r["iterations"] > 0 and r["cost"] < c0
    if name != "quadrotor_fig8":
        assert r["status"] in (1, 2)
    if name == "quadrotor_fig8":
        q = r["X"][:, 3:7]
        assert np.abs(np.linalg.norm(q, axis=1) - 1.0).max() < 0.1
