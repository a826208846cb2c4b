"""Workload definitions: BASELINE.json's configs made concrete plus the extra
reference fixtures the parity tests use.  Pure numpy, deterministic (seeded); consumed by tests/,
bench.py and __graft_entry__.smoke().  Nothing here touches the GPU or the oracle.

Every config cites the reference file it is taken from (relative to astomodynamics/cddp-cpp @ f71fa80)
and lists its deviations from that file.
"""
from __future__ import annotations

import math

import numpy as np

SEED_BASE = 20260101  # default_rng(seed = 20260101 + config_id)


def _diag(vals):
    return np.diag(np.asarray(vals, dtype=np.float64))


# ---------------------------------------------------------------------------------------------
# Minimal numpy dynamics, used ONLY to build initial trajectories (e.g. the hover rollout of the
# quadrotor example, examples/cddp_quadrotor_point.cpp:87-95).  Not a solver component.
# ---------------------------------------------------------------------------------------------
def _quadrotor_f(params, x, u):
    mass, I, L = params[0], np.asarray(params[1:10]).reshape(3, 3), params[10]
    q = x[3:7].copy()
    nq = math.sqrt(float(q @ q))
    q = q / nq if nq > 1e-6 else np.array([1.0, 0.0, 0.0, 0.0])
    qw, qx, qy, qz = q
    w = x[10:13]
    xd = np.zeros(13)
    xd[0:3] = x[7:10]
    xd[3] = -0.5 * (qx * w[0] + qy * w[1] + qz * w[2])
    xd[4] = 0.5 * (qw * w[0] + qy * w[2] - qz * w[1])
    xd[5] = 0.5 * (qw * w[1] - qx * w[2] + qz * w[0])
    xd[6] = 0.5 * (qw * w[2] + qx * w[1] - qy * w[0])
    thrust = u[0] + u[1] + u[2] + u[3]
    tau = np.array([L * (u[0] - u[2]), L * (u[1] - u[3]), 0.1 * (u[0] - u[1] + u[2] - u[3])])
    r3 = np.array([2 * (qx * qz + qy * qw), 2 * (qy * qz - qx * qw), 1 - 2 * (qx * qx + qy * qy)])
    xd[7:10] = (1.0 / mass) * (r3 * thrust) - np.array([0.0, 0.0, 9.81])
    xd[10:13] = np.linalg.inv(I) @ (tau - np.cross(w, I @ w))
    return xd


def _rk4(f, x, u, dt):
    k1 = f(x, u)
    k2 = f(x + 0.5 * dt * k1, u)
    k3 = f(x + 0.5 * dt * k2, u)
    k4 = f(x + dt * k3, u)
    return x + (dt / 6.0) * (k1 + 2 * k2 + 2 * k3 + k4)


def quadrotor_rollout(params, dt, x0, U):
    X = np.zeros((U.shape[0] + 1, 13))
    X[0] = x0
    f = lambda x, u: _quadrotor_f(params, x, u)  # noqa: E731
    for t in range(U.shape[0]):
        X[t + 1] = _rk4(f, X[t], U[t], dt)
    return X


# ---------------------------------------------------------------------------------------------
# option dictionaries (keys are cddp_b200_options field names; unspecified = reference default)
# ---------------------------------------------------------------------------------------------
def make_config(name: str, batch: int | None = None, horizon: int | None = None, seed_offset: int = 0) -> dict:
    """Returns {"spec", "options", "x0", "xref", "X0", "U0", "ref_traj", "name", "config_id", "notes"}."""
    builders = {
        "pendulum": _cfg_pendulum, "cartpole": _cfg_cartpole, "quadrotor": _cfg_quadrotor,
        "unicycle": _cfg_unicycle, "lti": _cfg_lti, "quadrotor_fig8": _cfg_quadrotor_fig8,
        "unicycle_obstacle": _cfg_unicycle_obstacle, "cartpole_ipddp": _cfg_cartpole_ipddp,
        "pendulum_ipddp": _cfg_pendulum_ipddp, "pendulum_ipddp_scaled": _cfg_pendulum_ipddp_scaled, "unicycle_ipddp_free": _cfg_unicycle_ipddp_free,
        "quadrotor_ipddp": _cfg_quadrotor_ipddp, "bicycle_user": _cfg_bicycle_user, "bicycle_user_ipddp": _cfg_bicycle_user_ipddp,
        "chain7_user": _cfg_chain7_user, "unicycle_obstacle_teq": _cfg_unicycle_obstacle_teq,
        "unicycle_teq": _cfg_unicycle_teq, "cartpole_teq": _cfg_cartpole_teq, "chain7_user_ipddp": _cfg_chain7_user_ipddp,
        "manip7_user": _cfg_manip7_user, "manip7_user_ipddp": _cfg_manip7_user_ipddp,
    }
    if name not in builders:
        raise KeyError(f"unknown config {name!r}; have {sorted(builders)}")
    return builders[name](batch, horizon, seed_offset)


def _cfg_pendulum(batch, horizon, seed_offset):
    """BASELINE config #1.  tests/cddp_core/test_clddp_solver.cpp:28-118 (CLDDPTest.SolvePendulum).
    Deviations: none for batch 1 (instance 0 is the unperturbed fixture); further instances perturb x0."""
    B = batch or 1
    N = horizon or 500
    dt = 0.05
    rng = np.random.default_rng(SEED_BASE + 1 + seed_offset)
    spec = dict(model="pendulum", n=2, m=1, horizon=N, dt=dt, integrator="euler", params=[1.0, 1.0, 0.0],
                Q=np.zeros((2, 2)), R=0.1 * np.eye(1), Qf=100.0 * np.eye(2), lb=[-10.0], ub=[10.0])
    options = dict(max_iterations=100, tolerance=1e-3, acceptable_tolerance=1e-4, reg_initial_value=1e-6)
    x0 = np.tile(np.array([math.pi, 0.0]), (B, 1))
    if B > 1:
        x0[1:] += 0.05 * rng.standard_normal((B - 1, 2))
    xref = np.zeros((B, 2))
    X0 = np.repeat(x0[:, None, :], N + 1, axis=1)
    U0 = np.zeros((B, N, 1))
    return dict(name="pendulum", config_id=1, spec=spec, options=options, x0=x0, xref=xref, X0=X0, U0=U0, ref_traj=None,
                notes="pendulum swing-up n=2 m=1 N=500, CLDDP + box +-10 (test_clddp_solver.cpp:28-118)")


def _cfg_cartpole(batch, horizon, seed_offset):
    """BASELINE config #2.  examples/cddp_cartpole.cpp:27-68.
    Deviations: the example adds a +-5 box and runs IPDDP; BASELINE asks for the
    unconstrained iLQR (CLDDP inverse branch, clddp_solver.cpp:142-145).  x0 perturbed, sigma=0.05."""
    B = batch or 1024
    N = horizon or 100
    dt = 0.05
    rng = np.random.default_rng(SEED_BASE + 2 + seed_offset)
    spec = dict(model="cartpole", n=4, m=1, horizon=N, dt=dt, integrator="rk4", params=[1.0, 0.2, 0.5, 9.81, 0.0],
                Q=np.zeros((4, 4)), R=0.1 * np.eye(1), Qf=100.0 * np.eye(4), lb=None, ub=None)
    options = dict(max_iterations=80, tolerance=1e-6, acceptable_tolerance=1e-5, reg_initial_value=1e-5)
    x0 = np.zeros((B, 4))
    x0[1:] += 0.05 * rng.standard_normal((B - 1, 4))
    xref = np.tile(np.array([0.0, math.pi, 0.0, 0.0]), (B, 1))
    X0 = np.repeat(x0[:, None, :], N + 1, axis=1)
    U0 = np.zeros((B, N, 1))
    return dict(name="cartpole", config_id=2, spec=spec, options=options, x0=x0, xref=xref, X0=X0, U0=U0, ref_traj=None,
                notes="cartpole swing-up n=4 m=1 N=100, unconstrained iLQR (examples/cddp_cartpole.cpp:27-68 minus the box)")


QUADROTOR_PARAMS = [1.0, 0.01, 0, 0, 0, 0.01, 0, 0, 0, 0.02, 0.2]  # mass, inertia(9), arm_length


def _cfg_quadrotor(batch, horizon, seed_offset):
    """BASELINE config #3 (the headline).  examples/cddp_quadrotor_point.cpp:23-98.
    Deviations: the example is N=120 and runs IPDDP; BASELINE asks for N=100 and the
    control-box CLDDP (boxQP) path.  Start position perturbed sigma=0.1, goal position sigma_g=0.5."""
    B = batch or 4096
    N = horizon or 100
    dt = 0.02
    rng = np.random.default_rng(SEED_BASE + 3 + seed_offset)
    Q = np.zeros((13, 13))
    Q[4, 4] = Q[5, 5] = Q[6, 6] = 0.1
    R = 0.1 * np.eye(4)
    Qf = _diag([500, 500, 500, 1, 1, 1, 1, 10, 10, 10, 0, 0, 0])
    spec = dict(model="quadrotor", n=13, m=4, horizon=N, dt=dt, integrator="rk4", params=QUADROTOR_PARAMS,
                Q=Q, R=R, Qf=Qf, lb=[0.0] * 4, ub=[5.0] * 4)
    options = dict(max_iterations=120, ls_max_iterations=15, reg_initial_value=1e-4)
    x0 = np.zeros((B, 13))
    x0[:, 3] = 1.0
    xref = np.zeros((B, 13))
    xref[:, 0], xref[:, 2], xref[:, 3] = 3.0, 2.0, 1.0
    if B > 1:
        x0[1:, 0:3] += 0.1 * rng.standard_normal((B - 1, 3))
        xref[1:, 0:3] += 0.5 * rng.standard_normal((B - 1, 3))
    hover = 1.0 * 9.81 / 4.0
    U0 = np.full((B, N, 4), hover)
    # hover rollout (cddp_quadrotor_point.cpp:87-95).  From rest at identity attitude it is stationary
    # up to roundoff, so roll out instance 0 once and translate it for the others.
    Xr = quadrotor_rollout(QUADROTOR_PARAMS, dt, x0[0], U0[0])
    X0 = np.repeat(Xr[None, :, :], B, axis=0)
    X0[:, :, 0:3] += (x0[:, None, 0:3] - x0[0, None, 0:3])
    return dict(name="quadrotor", config_id=3, spec=spec, options=options, x0=x0, xref=xref, X0=X0, U0=U0, ref_traj=None,
                notes="quadrotor point-to-point n=13 m=4 N=100, control box 0..5 (boxQP), CLDDP "
                      "(examples/cddp_quadrotor_point.cpp:23-98 with N=100, CLDDP instead of IPDDP)")


def _cfg_unicycle(batch, horizon, seed_offset):
    """tests/cddp_core/test_clddp_solver.cpp:231-290 (CLDDPTest.SolveUnicycle).
    Note the reference quirk: setInitialTrajectory(X=0) overwrites initial_state_ with X[0] = 0
    (cddp_core.cpp:139-141), so the fixture really starts from the origin.  Deviation: the test sets
    enable_parallel=true (min-cost selection); here the default sequential rule is used unless the
    caller sets options['enable_parallel']=1."""
    B = batch or 1
    N = horizon or 100
    dt = 0.03
    rng = np.random.default_rng(SEED_BASE + 6 + seed_offset)
    spec = dict(model="unicycle", n=3, m=2, horizon=N, dt=dt, integrator="euler", params=[],
                Q=np.zeros((3, 3)), R=0.5 * np.eye(2), Qf=0.5 * _diag([50.0, 50.0, 10.0]),
                lb=[-1.0, -math.pi], ub=[1.0, math.pi])
    options = dict(max_iterations=20, tolerance=1e-2)
    x0 = np.zeros((B, 3))
    xref = np.tile(np.array([2.0, 2.0, math.pi / 2.0]), (B, 1))
    if B > 1:
        xref[1:] += 0.2 * rng.standard_normal((B - 1, 3))
    X0 = np.zeros((B, N + 1, 3))
    U0 = np.zeros((B, N, 2))
    return dict(name="unicycle", config_id=6, spec=spec, options=options, x0=x0, xref=xref, X0=X0, U0=U0, ref_traj=None,
                notes="unicycle n=3 m=2 N=100 box CLDDP (test_clddp_solver.cpp:231-290)")


def _cfg_lti(batch, horizon, seed_offset):
    """LQ problem on LTISystem (src/dynamics_model/lti_system.cpp:71-92): the one model whose
    linearisation equals its rollout, so CLDDP converges in one accepted step and the gains are the
    textbook finite-horizon LQR gains (closed-form known answer).  Not a reference fixture."""
    B = batch or 4
    N = horizon or 50
    dt = 0.1
    rng = np.random.default_rng(SEED_BASE + 7 + seed_offset)
    n, m = 4, 2
    S = rng.standard_normal((n, n))
    S = 0.5 * (S - S.T)
    # A_d = expm(dt*S) via a short Taylor series (S skew-symmetric => well conditioned)
    Ad = np.eye(n)
    term = np.eye(n)
    for i in range(1, 20):
        term = term @ (dt * S) / i
        Ad = Ad + term
    Bd = dt * rng.standard_normal((n, m))
    spec = dict(model="lti", n=n, m=m, horizon=N, dt=dt, integrator="euler", params=[], lti_A=Ad, lti_B=Bd,
                Q=np.eye(n), R=0.5 * np.eye(m), Qf=10.0 * np.eye(n), lb=None, ub=None)
    options = dict(max_iterations=10, tolerance=1e-8, acceptable_tolerance=1e-12, reg_initial_value=1e-10,
                   reg_min_value=1e-12)
    x0 = rng.standard_normal((B, n))
    xref = np.zeros((B, n))
    X0 = np.repeat(x0[:, None, :], N + 1, axis=1)
    U0 = np.zeros((B, N, m))
    return dict(name="lti", config_id=7, spec=spec, options=options, x0=x0, xref=xref, X0=X0, U0=U0, ref_traj=None,
                notes="LTI LQ known-answer problem n=4 m=2")


def _cfg_quadrotor_fig8(batch, horizon, seed_offset):
    """Figure-8 tracking with a per-timestep reference trajectory, patterned on
    tests/cddp_core/test_clddp_solver.cpp:570-710 (CLDDPTest.SolveQuadrotor: box 0..4, reference_states).
    Shortened horizon by default so the oracle finishes in seconds."""
    B = batch or 2
    N = horizon or 100
    dt = 0.02
    rng = np.random.default_rng(SEED_BASE + 8 + seed_offset)
    Q = np.zeros((13, 13))
    Q[0, 0] = Q[1, 1] = Q[2, 2] = 1.0
    R = 0.01 * np.eye(4)
    Qf = _diag([10, 10, 10, 0, 0, 0, 0, 1, 1, 1, 0, 0, 0])
    spec = dict(model="quadrotor", n=13, m=4, horizon=N, dt=dt, integrator="rk4", params=QUADROTOR_PARAMS,
                Q=Q, R=R, Qf=Qf, lb=[0.0] * 4, ub=[4.0] * 4)
    options = dict(max_iterations=30, tolerance=1e-4, acceptable_tolerance=1e-6, reg_initial_value=1e-4)
    tt = np.arange(N + 1) * dt
    T = N * dt
    ang = 2 * math.pi * tt / T
    ref = np.zeros((B, N + 1, 13))
    for b in range(B):
        amp = 1.0 + (0.2 * rng.standard_normal() if b else 0.0)
        ref[b, :, 0] = amp * np.sin(ang)
        ref[b, :, 1] = amp * np.sin(ang) * np.cos(ang)
        ref[b, :, 2] = 1.0
        ref[b, :, 3] = 1.0
    x0 = np.zeros((B, 13))
    x0[:, 2], x0[:, 3] = 1.0, 1.0
    xref = ref[:, -1, :].copy()
    hover = 9.81 / 4.0
    U0 = np.full((B, N, 4), hover)
    X0 = np.repeat(x0[:, None, :], N + 1, axis=1)
    return dict(name="quadrotor_fig8", config_id=8, spec=spec, options=options, x0=x0, xref=xref, X0=X0, U0=U0,
                ref_traj=ref, notes="quadrotor figure-8 tracking with reference_states (test_clddp_solver.cpp:570-710 pattern)")


# ---------------------------------------------------------------------------------------------
# IPDDP workloads ("solver": "ipddp"; "constraints" = the path-constraint set, "ipddp_options" = overrides of
# CDDPOptions::ipddp).  X0 is None: IPDDP re-rolls the state trajectory out from U0 (ipddp_solver.cpp:876-882).
# ---------------------------------------------------------------------------------------------
def _cfg_unicycle_obstacle(batch, horizon, seed_offset):
    """BASELINE config #4 (path-inequality part): unicycle obstacle avoidance with IPDDP, patterned on
    examples/python_portfolio_lib.py:374-473 (ball r=0.4 at (1,1), control box +-(1.1, pi), R=0.05 I,
    Qf=diag(100,100,50), dt=0.03, tol 1e-4) with N=200 as BASELINE.json words it.  Deviations: the
    terminal-equality constraint of config #4 is not included (terminal constraints are out of scope of this
    round); per-instance goals are perturbed."""
    B = batch or 1
    N = horizon or 200
    dt = 0.03
    rng = np.random.default_rng(SEED_BASE + 4 + seed_offset)
    spec = dict(model="unicycle", n=3, m=2, horizon=N, dt=dt, integrator="euler", params=[],
                Q=np.zeros((3, 3)), R=0.05 * np.eye(2), Qf=_diag([100.0, 100.0, 50.0]), lb=None, ub=None)
    options = dict(max_iterations=150, tolerance=1e-4, acceptable_tolerance=1e-6, reg_initial_value=1e-6)
    constraints = [dict(type="control_box", lb=[-1.1, -math.pi], ub=[1.1, math.pi]),
                   dict(type="ball", center=[1.0, 1.0], radius=0.4)]
    x0 = np.zeros((B, 3))
    xref = np.tile(np.array([2.0, 2.0, math.pi / 2.0]), (B, 1))
    if B > 1:
        xref[1:, 0:2] += 0.15 * rng.standard_normal((B - 1, 2))
        xref[1:, 2] += 0.1 * rng.standard_normal(B - 1)
    U0 = np.zeros((B, N, 2))
    return dict(name="unicycle_obstacle", config_id=4, solver="ipddp", spec=spec, options=options, constraints=constraints,
                ipddp_options={}, x0=x0, xref=xref, X0=None, U0=U0, ref_traj=None,
                notes="unicycle obstacle avoidance n=3 m=2 N=200, IPDDP, control box + ball (python_portfolio_lib.py:374-473)")


def _cfg_unicycle_obstacle_teq(batch, horizon, seed_offset):
    """BASELINE config #4 in full: the obstacle-avoidance problem above PLUS TerminalEqualityConstraint(goal)
    (addTerminalConstraint pattern of tests/cddp_core/test_ipddp_solver.cpp:1417-1419), i.e. path-inequality +
    terminal-equality, IPDDP, n=3 m=2 N=200."""
    cfg = _cfg_unicycle_obstacle(batch, horizon, seed_offset)
    cfg.update(name="unicycle_obstacle_teq", config_id=4, ipddp_options=dict(terminal_equality=1),
               notes="unicycle obstacle avoidance n=3 m=2 N=200, IPDDP, control box + ball + terminal equality (BASELINE config #4)")
    return cfg


def _cfg_unicycle_teq(batch, horizon, seed_offset):
    """Terminal equality only (no path constraints): IPDDPTest.SolveWithTerminalEqualityOnly pattern
    (test_ipddp_solver.cpp:1580-1637) on the unicycle."""
    cfg = _cfg_unicycle_obstacle(batch, horizon or 100, seed_offset)
    cfg.update(name="unicycle_teq", config_id=15, constraints=[], ipddp_options=dict(terminal_equality=1),
               notes="unicycle, IPDDP with a terminal equality constraint only")
    return cfg


def _cfg_cartpole_teq(batch, horizon, seed_offset):
    """Cartpole swing-up (rk4, n=4, m=1) with control box + terminal equality."""
    cfg = _cfg_cartpole_ipddp(batch, horizon, seed_offset)
    cfg.update(name="cartpole_teq", config_id=16, constraints=[dict(type="control_box", lb=[-5.0], ub=[5.0])],
               ipddp_options=dict(terminal_equality=1), notes="cartpole swing-up, IPDDP, control box + terminal equality")
    return cfg


def _cfg_unicycle_ipddp_free(batch, horizon, seed_offset):
    """IPDDP with an empty constraint set (unconstrained branch, ipddp_solver.cpp:1055-1118)."""
    cfg = _cfg_unicycle_obstacle(batch, horizon or 100, seed_offset)
    cfg.update(name="unicycle_ipddp_free", config_id=14, constraints=[], notes="unicycle, IPDDP without constraints")
    return cfg


def _cfg_pendulum_ipddp_scaled(batch, horizon, seed_offset):
    """pendulum_ipddp with ControlConstraint(lb, ub, scale_factor = 0.25): the scale multiplies evaluate(), the IP upper
    bound and the Jacobians (constraint.hpp:147-218), so slack / dual initialisation, theta and the barrier schedule
    differ from the unscaled problem."""
    cfg = _cfg_pendulum_ipddp(batch, horizon, seed_offset)
    cfg["constraints"] = [dict(type="control_box", lb=[-20.0], ub=[20.0], scale=0.25)]
    cfg["name"] = "pendulum_ipddp_scaled"
    return cfg


def _cfg_pendulum_ipddp(batch, horizon, seed_offset):
    """examples/cddp_pendulum.cpp:27-67: Pendulum(dt=0.02, l=0.5, m=1, b=0.01) N=100, control box +-20, IPDDP."""
    B = batch or 1
    N = horizon or 100
    dt = 0.02
    rng = np.random.default_rng(SEED_BASE + 11 + seed_offset)
    spec = dict(model="pendulum", n=2, m=1, horizon=N, dt=dt, integrator="euler", params=[0.5, 1.0, 0.01],
                Q=np.zeros((2, 2)), R=0.1 * np.eye(1), Qf=100.0 * np.eye(2), lb=None, ub=None)
    options = dict(max_iterations=100, tolerance=1e-5, acceptable_tolerance=1e-6, reg_initial_value=1e-6)
    constraints = [dict(type="control_box", lb=[-20.0], ub=[20.0])]
    x0 = np.tile(np.array([math.pi, 0.0]), (B, 1))
    if B > 1:
        x0[1:] += 0.05 * rng.standard_normal((B - 1, 2))
    xref = np.zeros((B, 2))
    U0 = np.zeros((B, N, 1))
    return dict(name="pendulum_ipddp", config_id=11, solver="ipddp", spec=spec, options=options, constraints=constraints,
                ipddp_options={}, x0=x0, xref=xref, X0=None, U0=U0, ref_traj=None,
                notes="pendulum swing-up, IPDDP + control box +-20 (examples/cddp_pendulum.cpp:27-67)")


def _cfg_cartpole_ipddp(batch, horizon, seed_offset):
    """examples/cddp_cartpole.cpp:27-68: CartPole rk4, control box +-5, IPDDP; plus a state box on the cart position
    (StateConstraint) and a linear constraint so that every constraint kind is exercised."""
    B = batch or 1
    N = horizon or 100
    dt = 0.05
    rng = np.random.default_rng(SEED_BASE + 12 + seed_offset)
    spec = dict(model="cartpole", n=4, m=1, horizon=N, dt=dt, integrator="rk4", params=[1.0, 0.2, 0.5, 9.81, 0.0],
                Q=np.zeros((4, 4)), R=0.1 * np.eye(1), Qf=100.0 * np.eye(4), lb=None, ub=None)
    options = dict(max_iterations=120, tolerance=1e-5, acceptable_tolerance=1e-6, reg_initial_value=1e-5)
    big = 1e3
    constraints = [dict(type="control_box", lb=[-5.0], ub=[5.0]),
                   dict(type="state_box", lb=[-2.0, -big, -big, -big], ub=[2.0, big, big, big]),
                   dict(type="linear", A=[[0.0, 0.0, 1.0, 0.0], [0.0, 0.0, -1.0, 0.0]], b=[6.0, 6.0])]
    x0 = np.zeros((B, 4))
    if B > 1:
        x0[1:] += 0.05 * rng.standard_normal((B - 1, 4))
    xref = np.tile(np.array([0.0, math.pi, 0.0, 0.0]), (B, 1))
    U0 = np.zeros((B, N, 1))
    return dict(name="cartpole_ipddp", config_id=12, solver="ipddp", spec=spec, options=options, constraints=constraints,
                ipddp_options={}, x0=x0, xref=xref, X0=None, U0=U0, ref_traj=None,
                notes="cartpole swing-up, IPDDP, control box +-5 + state box + linear (examples/cddp_cartpole.cpp:27-68)")


def _cfg_quadrotor_ipddp(batch, horizon, seed_offset):
    """examples/cddp_quadrotor_point.cpp:23-98 as written (IPDDP, control box 0..5) with a shorter default horizon."""
    cfg = _cfg_quadrotor(batch or 2, horizon or 60, seed_offset)
    spec = dict(cfg["spec"], lb=None, ub=None)
    options = dict(max_iterations=80, tolerance=1e-4, acceptable_tolerance=1e-6, reg_initial_value=1e-4, ls_max_iterations=11)
    cfg.update(name="quadrotor_ipddp", config_id=13, solver="ipddp", spec=spec, options=options, X0=None, ipddp_options={},
               constraints=[dict(type="control_box", lb=[0.0] * 4, ub=[5.0] * 4)],
               notes="quadrotor point-to-point, IPDDP + control box 0..5 (examples/cddp_quadrotor_point.cpp:23-98)")
    return cfg


# ---------------------------------------------------------------------------------------------
# User-model plugin workloads: spec["model"] = "user" with spec["model_source"] = CUDA source of the dynamics
# (include/cddp_b200.h, cddp_b200_create_ex); spec["twin_model"] names the CPU checker's native implementation of the same model.
# ---------------------------------------------------------------------------------------------
BICYCLE_SOURCE = """
// Bicycle (src/dynamics_model/bicycle.cpp:29-47): state (x, y, theta, v), control (a, delta); p[0] = wheelbase
template <class T>
__device__ void cddp_user_dynamics(const T *x, const T *u, const double *p, T *xdot) {
  xdot[0] = x[3] * cos(x[2]);
  xdot[1] = x[3] * sin(x[2]);
  xdot[2] = (x[3] / p[0]) * tan(u[1]);
  xdot[3] = u[0];
}
"""

CHAIN7_SOURCE = """
// 7-joint chain (NOT a reference model; stands in for BASELINE config #5's 7-DOF manipulator): gravity, viscous friction,
// nearest-neighbour elastic coupling.  state (q[7], qdot[7]), control tau[7]; p = g, c, k, I_1..I_7
template <class T>
__device__ void cddp_user_dynamics(const T *x, const T *u, const double *p, T *xdot) {
  const double g = p[0], c = p[1], k = p[2];
  for (int i = 0; i < 7; ++i) {
    xdot[i] = x[7 + i];
    T acc = u[i] - g * sin(x[i]) - c * x[7 + i];
    if (i > 0) acc = acc - k * sin(x[i] - x[i - 1]);
    if (i < 6) acc = acc - k * sin(x[i] - x[i + 1]);
    xdot[7 + i] = acc / p[3 + i];
  }
}
"""


MANIP7_SOURCE = """
// 7-DOF serial manipulator, the n-joint generalisation of the reference's 3-DOF Manipulator (src/dynamics_model/
// manipulator.cpp:29-51 dynamics, :174-208 mass matrix and gravity vector: point masses at the link ends, base joint about
// the vertical, M_ij = (sum of the masses outboard of max(i,j)) l_i l_j cos(q_{i+1} + ... + q_j)), with the Coriolis /
// centrifugal vector that belongs to that M(q) (Christoffel symbols) and viscous joint friction:
//     M(q) qdd + h(q, qd) + G(q) + b qd = tau,   state (q[7], qd[7]), control tau[7]
// p = g, b, m_0..m_6, l_0..l_6.  sigma_j = q_1 + ... + q_j (sigma_0 = 0), w_j = d/dt sigma_j:
//     M_ij   = mu_max(i,j) l_i l_j cos(sigma_j - sigma_i),  mu_j = m_j + ... + m_6
//     h_k    = sum_j Mdot_kj qd_j - 1/2 d/dq_k (qd^T M qd),  Mdot_ij = -S_ij (w_j - w_i),  S_ij = mu_j l_i l_j sin(sigma_j - sigma_i) (i < j)
//              d/dq_k (qd^T M qd) = -2 sum_{i < k <= j} S_ij qd_i qd_j
//     G_0 = 0,  G_k = -sum_{j >= k} mu_j g l_j cos(sigma_j)                     (manipulator.cpp:200-206)
// qdd from an LDL^T factorisation of the 7 x 7 mass matrix (positive definite for positive masses and lengths).
template <class T>
__device__ void cddp_user_dynamics(const T *x, const T *u, const double *p, T *xdot) {
  // every loop has a constant trip count and is unrolled, so that c, s, w, r and the 28 entries of the upper triangle of M
  // are addressed statically and live in registers (as run-time-indexed local arrays one RK4 step cost 60 us)
  const double g = p[0], bv = p[1];
  const double *mass = p + 2, *len = p + 9;
  double mu[7];
  mu[6] = mass[6];
#pragma unroll
  for (int j = 5; j >= 0; --j) mu[j] = mu[j + 1] + mass[j];
  T c[7], s[7], w[7];
  {
    T sg = x[0] * 0.0, sv = x[0] * 0.0;
#pragma unroll
    for (int j = 0; j < 7; ++j) {
      if (j > 0) {
        sg = sg + x[j];
        sv = sv + x[7 + j];
      }
      c[j] = cos(sg);
      s[j] = sin(sg);
      w[j] = sv;
    }
  }
  T M[7][7], r[7];
#pragma unroll
  for (int k = 0; k < 7; ++k) r[k] = u[k] - bv * x[7 + k];
#pragma unroll
  for (int i = 0; i < 7; ++i) {
    M[i][i] = x[0] * 0.0 + mu[i] * len[i] * len[i];
#pragma unroll
    for (int j = i + 1; j < 7; ++j) {
      const double a = mu[j] * len[i] * len[j];
      M[i][j] = a * (c[i] * c[j] + s[i] * s[j]);
      const T S = a * (s[j] * c[i] - c[j] * s[i]);
      const T md = S * (w[j] - w[i]);  // = -Mdot_ij
      r[i] = r[i] + md * x[7 + j];
      r[j] = r[j] + md * x[7 + i];
      const T P = S * (x[7 + i] * x[7 + j]);
#pragma unroll
      for (int k = i + 1; k <= j; ++k) r[k] = r[k] - P;
    }
  }
  {
    T acc = x[0] * 0.0;  // -G_k = sum_{j >= k} mu_j g l_j cos(sigma_j), k >= 1
#pragma unroll
    for (int k = 6; k >= 1; --k) {
      acc = acc + (mu[k] * g * len[k]) * c[k];
      r[k] = r[k] + acc;
    }
  }
  // LDL^T in place (unit lower factor in the upper-triangle slots M[j][i], i > j; pivots on the diagonal)
  T rd[7];  // reciprocal pivots: one division per column instead of one per entry (a division is a ~20-instruction chain)
#pragma unroll
  for (int j = 0; j < 7; ++j) {
    T d = M[j][j];
#pragma unroll
    for (int k = 0; k < j; ++k) d = d - M[k][j] * M[k][j] * M[k][k];
    M[j][j] = d;
    rd[j] = 1.0 / d;
#pragma unroll
    for (int i = j + 1; i < 7; ++i) {
      T v = M[j][i];
#pragma unroll
      for (int k = 0; k < j; ++k) v = v - M[k][i] * M[k][j] * M[k][k];
      M[j][i] = v * rd[j];
    }
  }
#pragma unroll
  for (int i = 0; i < 7; ++i)
#pragma unroll
    for (int k = 0; k < i; ++k) r[i] = r[i] - M[k][i] * r[k];
#pragma unroll
  for (int i = 0; i < 7; ++i) r[i] = r[i] * rd[i];
#pragma unroll
  for (int i = 6; i >= 0; --i)
#pragma unroll
    for (int k = i + 1; k < 7; ++k) r[i] = r[i] - M[i][k] * r[k];
#pragma unroll
  for (int i = 0; i < 7; ++i) {
    xdot[i] = x[7 + i];
    xdot[7 + i] = r[i];
  }
}
"""


def _cfg_bicycle_user(batch, horizon, seed_offset):
    """The reference's Bicycle model (src/dynamics_model/bicycle.cpp) supplied through the user-model plugin; CLDDP with a
    control box (acceleration, steering), parking-style point-to-point problem."""
    B = batch or 4
    N = horizon or 100
    dt = 0.05
    rng = np.random.default_rng(SEED_BASE + 21 + seed_offset)
    spec = dict(model="user", twin_model="bicycle", model_source=BICYCLE_SOURCE, n=4, m=2, horizon=N, dt=dt, integrator="rk4",
                params=[2.0], Q=_diag([0.0, 0.0, 0.0, 0.01]), R=_diag([0.1, 0.5]), Qf=_diag([100.0, 100.0, 50.0, 10.0]),
                lb=[-2.0, -0.6], ub=[2.0, 0.6])
    options = dict(max_iterations=60, tolerance=1e-5, acceptable_tolerance=1e-7, reg_initial_value=1e-5)
    x0 = np.zeros((B, 4))
    xref = np.tile(np.array([4.0, 3.0, math.pi / 2.0, 0.0]), (B, 1))
    if B > 1:
        xref[1:, 0:2] += 0.3 * rng.standard_normal((B - 1, 2))
    X0 = np.repeat(x0[:, None, :], N + 1, axis=1)
    U0 = np.zeros((B, N, 2))
    U0[:, :, 0] = 0.2
    return dict(name="bicycle_user", config_id=21, spec=spec, options=options, x0=x0, xref=xref, X0=X0, U0=U0, ref_traj=None,
                notes="bicycle (reference model bicycle.cpp) through the user-model plugin, n=4 m=2 N=100, CLDDP + control box")


def _cfg_bicycle_user_ipddp(batch, horizon, seed_offset):
    """Same plugin model under IPDDP with a control box and a circular obstacle on the path."""
    cfg = _cfg_bicycle_user(batch, horizon, seed_offset)
    spec = dict(cfg["spec"], lb=None, ub=None)
    cfg.update(name="bicycle_user_ipddp", config_id=22, solver="ipddp", spec=spec, X0=None, ipddp_options={},
               options=dict(max_iterations=60, tolerance=1e-4, acceptable_tolerance=1e-6, reg_initial_value=1e-4),
               constraints=[dict(type="control_box", lb=[-2.0, -0.6], ub=[2.0, 0.6]), dict(type="ball", center=[3.0, 1.2], radius=0.5)],
               notes="bicycle through the user-model plugin, IPDDP, control box + ball obstacle")
    return cfg


def _cfg_chain7_user(batch, horizon, seed_offset):
    """BASELINE config #5 stand-in: n=14, m=7, N=150, control box, CLDDP (the reference's manipulator example is CLDDP +
    box, examples/cddp_manipulator.cpp:23-70; cost pattern Q = diag(1 x7, 0.1 x7), R = 0.1 I, Qf = 100 Q from the same
    file).  The 7-DOF model itself does not exist in the reference: it is a plugin model of this build."""
    B = batch or 4
    N = horizon or 150
    dt = 0.01
    rng = np.random.default_rng(SEED_BASE + 5 + seed_offset)
    inertia = [1.0, 0.9, 0.8, 0.7, 0.6, 0.5, 0.4]
    qw = [1.0] * 7 + [0.1] * 7
    spec = dict(model="user", twin_model="chain7", model_source=CHAIN7_SOURCE, n=14, m=7, horizon=N, dt=dt, integrator="rk4",
                params=[9.81, 0.5, 4.0] + inertia, Q=_diag(qw), R=0.1 * np.eye(7), Qf=100.0 * _diag(qw),
                lb=[-50.0] * 7, ub=[50.0] * 7)
    options = dict(max_iterations=60, tolerance=1e-4, acceptable_tolerance=1e-6, reg_initial_value=1e-5)
    x0 = np.zeros((B, 14))
    xref = np.zeros((B, 14))
    xref[:, :7] = np.array([0.8, -0.6, 0.5, -0.4, 0.3, -0.2, 0.1])
    if B > 1:
        xref[1:, :7] += 0.1 * rng.standard_normal((B - 1, 7))
    X0 = np.repeat(x0[:, None, :], N + 1, axis=1)
    U0 = np.zeros((B, N, 7))
    return dict(name="chain7_user", config_id=5, spec=spec, options=options, x0=x0, xref=xref, X0=X0, U0=U0, ref_traj=None,
                notes="7-joint chain (plugin model; stands in for BASELINE config #5) n=14 m=7 N=150, CLDDP + control box +-50")


MANIP7_PARAMS = [9.81, 0.05] + [1.0, 1.0, 1.0, 0.8, 0.6, 0.5, 0.4] + [0.3, 0.3, 0.25, 0.25, 0.2, 0.15, 0.1]  # g, b, masses, lengths


def _cfg_manip7_user(batch, horizon, seed_offset):
    """BASELINE config #5: 7-DOF manipulator n=14 m=7 N=150, control box, CLDDP.  Problem data patterned on the reference's
    3-DOF example (examples/cddp_manipulator.cpp:23-70): dt = 0.01, torque box +-50, Q = diag(1 x7, 0.1 x7), R = 0.1 I,
    Qf = 100 Q, max_iterations 80, line_search.max_iterations 20 -> capped at 16 candidates here; the nominal is the arm at
    rest (see below) instead of the example's straight line.  The model is the 7-joint
    generalisation of the reference's Manipulator (MANIP7_SOURCE), supplied through the user-model plugin; the reference
    itself has only the 3-DOF model.  Per-instance goals are perturbed."""
    B = batch or 4
    N = horizon or 150
    dt = 0.01
    rng = np.random.default_rng(SEED_BASE + 5 + seed_offset)
    qw = [1.0] * 7 + [0.1] * 7
    spec = dict(model="user", twin_model="manip7", model_source=MANIP7_SOURCE, n=14, m=7, horizon=N, dt=dt, integrator="rk4",
                params=MANIP7_PARAMS, Q=_diag(qw), R=0.1 * np.eye(7), Qf=100.0 * _diag(qw), lb=[-50.0] * 7, ub=[50.0] * 7)
    options = dict(max_iterations=80, ls_max_iterations=16)
    x0 = np.zeros((B, 14))
    x0[:, 1] = -math.pi / 2.0  # arm hanging: sigma_j = -pi/2 for every j >= 1
    xref = np.zeros((B, 14))
    xref[:, :7] = np.array([math.pi / 2, -math.pi / 6, -math.pi / 3, math.pi / 4, -math.pi / 4, math.pi / 6, 0.0])
    if B > 1:
        xref[1:, :7] += 0.1 * rng.standard_normal((B - 1, 7))
    # the hanging arm at rest is an equilibrium of the model, so X0 = x0 repeated with U0 = 0 is a dynamically consistent
    # nominal.  (The example's straight-line X0 with U0 = 0 is not: CLDDP takes the cost of the GIVEN trajectory as its
    # starting point, clddp_solver.cpp:62-66, no rollout can match the cost of a path that reaches the goal for free, every
    # line search fails and the solve ends at the regularisation limit with the nominal unchanged.)
    X0 = np.repeat(x0[:, None, :], N + 1, axis=0 + 1)
    U0 = np.zeros((B, N, 7))
    return dict(name="manip7_user", config_id=5, spec=spec, options=options, x0=x0, xref=xref, X0=X0, U0=U0, ref_traj=None,
                notes="7-DOF manipulator (plugin model: M(q) qdd + h(q,qd) + G(q) + b qd = tau, the 7-joint generalisation of "
                      "src/dynamics_model/manipulator.cpp) n=14 m=7 N=150, CLDDP + torque box +-50")


def _cfg_manip7_user_ipddp(batch, horizon, seed_offset):
    """BASELINE config #5 as worded ("7-DOF manipulator n=14 m=7 N=150, mixed constraints"): the same model under IPDDP
    with MIXED path constraints — torque box (ControlConstraint, 14 rows) + joint-angle / joint-rate box (StateConstraint,
    28 rows): d = 42."""
    cfg = _cfg_manip7_user(batch, horizon, seed_offset)
    spec = dict(cfg["spec"], lb=None, ub=None)
    cfg.update(name="manip7_user_ipddp", config_id=5, solver="ipddp", spec=spec, X0=None, ipddp_options={},
               options=dict(max_iterations=80, tolerance=1e-4, acceptable_tolerance=1e-6, reg_initial_value=1e-5),
               constraints=[dict(type="control_box", lb=[-50.0] * 7, ub=[50.0] * 7),
                            dict(type="state_box", lb=[-3.3] * 7 + [-8.0] * 7, ub=[3.3] * 7 + [8.0] * 7)],
               notes="7-DOF manipulator (plugin model) n=14 m=7 N=150, IPDDP, torque box + joint/rate box (d=42)")
    return cfg


def _cfg_chain7_user_ipddp(batch, horizon, seed_offset):
    """BASELINE config #5 as worded ("7-DOF manipulator n=14 m=7 N=150, mixed constraints"): the plugin chain model under IPDDP
    with MIXED path constraints — torque box (ControlConstraint, 14 rows) + joint-angle / joint-rate box (StateConstraint,
    28 rows): d = 42."""
    cfg = _cfg_chain7_user(batch, horizon, seed_offset)
    spec = dict(cfg["spec"], lb=None, ub=None)
    cfg.update(name="chain7_user_ipddp", config_id=5, solver="ipddp", spec=spec, X0=None, ipddp_options={},
               options=dict(max_iterations=60, tolerance=1e-4, acceptable_tolerance=1e-6, reg_initial_value=1e-5),
               constraints=[dict(type="control_box", lb=[-50.0] * 7, ub=[50.0] * 7),
                            dict(type="state_box", lb=[-1.0] * 7 + [-3.0] * 7, ub=[1.0] * 7 + [3.0] * 7)],
               notes="7-joint chain (plugin model; BASELINE config #5 stand-in) n=14 m=7 N=150, IPDDP, torque box + joint/rate box (d=42)")
    return cfg


def shard(cfg: dict, rank: int, world: int) -> dict:
    """Contiguous batch shard for one rank: B_g = ceil(B / G)."""
    B = cfg["x0"].shape[0]
    per = (B + world - 1) // world
    lo, hi = min(rank * per, B), min((rank + 1) * per, B)
    out = dict(cfg)
    for key in ("x0", "xref", "X0", "U0", "ref_traj"):
        out[key] = None if cfg[key] is None else cfg[key][lo:hi]
    out["shard"] = (lo, hi)
    return out
