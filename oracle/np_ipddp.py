"""oracle/np_ipddp.py — SECOND, independent CPU restatement of the reference's IPDDP path (numpy).

TEST INFRASTRUCTURE ONLY (see oracle/cddp_oracle.h).  It exists to pin the C++ IPDDP oracle
(oracle/cddp_oracle.cpp, section "IPDDP"): written separately, in matrix form, straight from the Eigen statements of
src/cddp_core/ipddp_solver.cpp @ f71fa80, with different numerical routes:

  * dynamics Jacobians: complex-step differentiation (np_oracle.Problem.jacobians) instead of dual numbers;
  * Q_uu_reg systems: numpy.linalg.solve (LU) instead of the pivoted LDLT restatement — like Eigen::LDLT, LU does not
    reject indefinite matrices (ipddp_solver.cpp:1431-1435 only fails on info() != Success);
  * barrier merit / theta: summed constraint-major over dictionaries of trajectories, as the reference's std::map code
    does (:2778-2880); the C++ oracle does the same, the CUDA path sums time-major.

Scope: cold start, use_ilqr = true, path inequality constraints (control/state box, ball, linear), optional
TerminalEqualityConstraint on the reference state (ipddp_solver.cpp:1120-1353, :484-639; singular values by numpy.linalg.svd,
the reduced systems by numpy.linalg.solve).
PARITY STATUS: "parity unpinned" — the reference binary cannot be built in this image.
"""
from __future__ import annotations

import math

import numpy as np

import np_oracle as npo

K_SLACK_OFFSET, EPS_SLACK, MAX_RATIO = 1e-4, 1e-10, 1e6  # ipddp_solver.cpp:35-38

IP_DEFAULTS = dict(  # options.hpp:75-104, :148-186
    dual_var_init_scale=1e-1, slack_var_init_scale=1e-2, barrier_tol_mult=0.1, barrier_update_dual_weight=0.01,
    mu_kappa_epsilon=10.0, theta_0_floor=1.0, mu_initial=1.0, mu_min_value=1e-10, mu_update_factor=0.5, mu_update_power=1.2,
    min_fraction_to_boundary=0.99, merit_acceptance_threshold=1e-6, violation_acceptance_threshold=1e-6,
    max_violation_threshold=1e4, min_violation_for_armijo_check=1e-7, theta_norm_l2=0, max_filter_size=5, barrier_strategy=0,
    jacobian_regularization_value=1e-8, jacobian_regularization_exponent=0.25, terminal_equality=0,
)
NAMES = {"control_box": "ControlConstraint", "state_box": "StateConstraint", "ball": "BallConstraint", "linear": "LinearConstraint"}


class Con:
    """One path constraint: evaluate(x,u) - upper, dg/dx, dg/du (constraint.hpp:144-440)."""

    def __init__(self, c, n, m):
        self.t, self.n, self.m = c["type"], n, m
        self.scale = float(c.get("scale", 1.0))
        if self.t in ("control_box", "state_box"):
            self.lb, self.ub = np.asarray(c["lb"], float), np.asarray(c["ub"], float)
            self.dim = 2 * len(self.lb)
            self.upper = np.concatenate([-self.lb * self.scale, self.ub * self.scale])
        elif self.t == "ball":
            self.center, self.radius = np.asarray(c["center"], float), float(c["radius"])
            self.dim = 1
            self.upper = -np.array([self.radius * self.radius]) * self.scale
        else:
            self.A, self.b = np.atleast_2d(np.asarray(c["A"], float)), np.asarray(c["b"], float)
            self.dim = len(self.b)
            self.upper = self.b

    def g(self, x, u):
        if self.t == "control_box":
            return np.concatenate([-u, u]) * self.scale - self.upper
        if self.t == "state_box":
            return np.concatenate([-x, x]) * self.scale - self.upper
        if self.t == "ball":
            diff = x[: len(self.center)] - self.center
            return -np.array([self.scale * float(diff @ diff)]) - self.upper
        return self.A @ x - self.upper

    def jac(self, x, u):
        Gx, Gu = np.zeros((self.dim, self.n)), np.zeros((self.dim, self.m))
        if self.t == "control_box":
            Gu[: self.m], Gu[self.m:] = -np.eye(self.m) * self.scale, np.eye(self.m) * self.scale
        elif self.t == "state_box":
            Gx[: self.n], Gx[self.n:] = -np.eye(self.n) * self.scale, np.eye(self.n) * self.scale
        elif self.t == "ball":
            k = len(self.center)
            Gx[0, :k] = -2.0 * self.scale * (x[:k] - self.center)
        else:
            Gx[:] = self.A
        return Gx, Gu


class State:
    pass


def _clip(v, lo, hi):
    return np.minimum(np.maximum(v, lo), hi)


def _setup(spec, opts, ipopts, constraints, x0, xref, U0):
    s = State()
    s.P = npo.Problem(dict(spec, lb=None, ub=None))
    s.o = npo.options(**opts)
    s.io = dict(IP_DEFAULTS, **(ipopts or {}))
    cs = sorted(constraints or [], key=lambda c: c.get("name", NAMES[c["type"]]))  # std::map order
    s.cons = [Con(c, s.P.n, s.P.m) for c in cs]
    s.x0, s.xref = np.asarray(x0, float), np.asarray(xref, float)
    P, N = s.P, s.P.N
    # cold start (:818-913)
    s.U = np.zeros((N, P.m)) if U0 is None else np.array(U0, float)
    s.X = np.zeros((N + 1, P.n))
    s.X[0] = s.x0
    for t in range(N):
        s.X[t + 1] = P.step(s.X[t], s.U[t])
    s.teq = bool(s.io["terminal_equality"])
    s.lamT, s.dlamT = np.zeros(P.n), np.zeros(P.n)
    s.mu = max(s.o["tolerance"] / 10.0, s.io["mu_min_value"]) if (not s.cons and not s.teq) else s.io["mu_initial"]
    s.reg, s.step_norm, s.alpha_pr, s.alpha_du = s.o["reg_initial_value"], 0.0, 1.0, 1.0
    s.G = [np.array([c.g(s.X[t], s.U[t]) for t in range(N)]) for c in s.cons]
    s.S = [np.maximum(s.io["slack_var_init_scale"], -g + K_SLACK_OFFSET) for g in s.G]  # :2447-2466
    s.Y = [(s.mu * s.io["dual_var_init_scale"]) / np.maximum(sl, EPS_SLACK) for sl in s.S]
    s.cost = P.trajectory_cost(s.X, s.U, s.xref)
    s.filter = []
    _reset_filter(s)
    s.inf_du, s.dV = 0.0, np.zeros(2)
    return s


def _theta(s, G, S, hT=None):  # computeTheta :2778-2848
    total = mx = 0.0
    for g, sl in zip(G, S):
        for t in range(len(g)):
            r = g[t] + sl[t]
            total += float(r @ r) if s.io["theta_norm_l2"] else float(np.abs(r).sum())
            mx = max(mx, float(np.abs(r).max()))
    if hT is not None:
        total += float(hT @ hT) if s.io["theta_norm_l2"] else float(np.abs(hT).sum())
        mx = max(mx, float(np.abs(hT).max()))
    th = math.sqrt(total) if s.io["theta_norm_l2"] else total
    return max(th, mx)


def _merit(s, S, cost, lamT=None, hT=None):  # computeBarrierMerit :2850-2880
    merit = cost
    for sl in S:
        for t in range(len(sl)):
            merit -= s.mu * float(np.log(np.maximum(sl[t], EPS_SLACK)).sum())
    if lamT is not None and hT is not None:
        merit += float(lamT @ hT)
    return merit


def _primal_comp(s, G, S, Y, mu, hT=None):  # :2882-2937
    ip = ic = 0.0
    for g, sl, y in zip(G, S, Y):
        ip = max(ip, float(np.abs(g + sl).max()))
        ic = max(ic, float(np.abs(y * sl - mu).max()))
    if hT is not None:
        ip = max(ip, float(np.abs(hT).max()))
    return ip, ic


def _hT(s, X):
    return (X[s.P.N] - s.xref) if s.teq else None


def _reset_filter(s):  # resetBarrierFilter :2484-2517
    h = _hT(s, s.X)
    s.inf_pr, s.inf_comp = _primal_comp(s, s.G, s.S, s.Y, s.mu, h)
    s.merit = s.phi = _merit(s, s.S, s.cost, s.lamT if s.teq else None, h)
    s.filter_theta = max(_theta(s, s.G, s.S, h), 1e-8)
    s.filter = []
    if s.teq:
        _accept_filter(s.filter, s.phi, s.filter_theta)


def _backward(s):
    """IPDDPSolver::backwardPass, branches 1 (:1055-1118) and 3 (:1355-1569)."""
    P, N, n, m = s.P, s.P.N, s.P.n, s.P.m
    A, B = [], []
    for t in range(N):
        Fx, Fu = P.jacobians(s.X[t], s.U[t])
        A.append(np.eye(n) + P.dt * Fx)
        B.append(P.dt * Fu)
    s.A, s.B = A, B
    Vx = 2.0 * P.Qf @ (s.X[N] - s.xref)
    Vxx = 2.0 * P.Qf
    Vxx = 0.5 * (Vxx + Vxx.T)
    d = sum(c.dim for c in s.cons)
    if s.teq:
        return _backward_teq(s, Vx, Vxx, d)
    s.ku, s.Ku = np.zeros((N, m)), np.zeros((N, m, n))
    s.ky, s.Ky, s.ks, s.Ks = np.zeros((N, d)), np.zeros((N, d, n)), np.zeros((N, d)), np.zeros((N, d, n))
    dV = np.zeros(2)
    inf_du = inf_pr = inf_comp = step_norm = 0.0
    for t in range(N - 1, -1, -1):
        x, u = s.X[t], s.U[t]
        ref = P.ref_at(s.xref, None, t)
        l_x, l_u = 2.0 * P.Qs @ (x - ref), 2.0 * P.Rs @ u
        l_xx, l_uu = 2.0 * P.Qs, 2.0 * P.Rs
        Q_x = l_x + A[t].T @ Vx
        Q_u = l_u + B[t].T @ Vx
        Q_xx = l_xx + A[t].T @ Vxx @ A[t]
        Q_ux = B[t].T @ Vxx @ A[t]
        Q_uu = l_uu + B[t].T @ Vxx @ B[t]
        if not s.cons:
            Q_uu = 0.5 * (Q_uu + Q_uu.T)
            Q_uu = Q_uu + s.reg * np.eye(m)
            k_u = -np.linalg.solve(Q_uu, Q_u)
            K_u = -np.linalg.solve(Q_uu, Q_ux)
        else:
            y = np.concatenate([Y[t] for Y in s.Y])
            sl = np.concatenate([S[t] for S in s.S])
            g = np.concatenate([G[t] for G in s.G])
            jac = [c.jac(x, u) for c in s.cons]
            Q_yx, Q_yu = np.vstack([j[0] for j in jac]), np.vstack([j[1] for j in jac])
            Q_x = l_x + Q_yx.T @ y + A[t].T @ Vx
            Q_u = l_u + Q_yu.T @ y + B[t].T @ Vx
            s_safe = np.maximum(sl, max(s.mu * 1e-3, EPS_SLACK))
            YSinv = np.diag(_clip(y / s_safe, 0.0, MAX_RATIO))
            primal = g + sl
            comp = y * sl - s.mu
            rhat = y * primal - comp
            Q_uu_reg = 0.5 * (Q_uu + Q_uu.T) + Q_yu.T @ YSinv @ Q_yu + s.reg * np.eye(m)
            S_inv_rhat = _clip(rhat / s_safe, -MAX_RATIO, MAX_RATIO)
            big = np.column_stack([Q_u + Q_yu.T @ S_inv_rhat, Q_ux + Q_yu.T @ YSinv @ Q_yx])
            kK = -np.linalg.solve(Q_uu_reg, big)
            k_u, K_u = kK[:, 0], kK[:, 1:]
            temp = Q_yu @ k_u
            s.ky[t] = _clip((rhat + y * temp) / s_safe, -MAX_RATIO, MAX_RATIO)
            s.Ky[t] = _clip(YSinv @ (Q_yx + Q_yu @ K_u), -MAX_RATIO, MAX_RATIO)
            s.ks[t] = -primal - temp
            s.Ks[t] = -Q_yx - Q_yu @ K_u
            Q_u = Q_u + Q_yu.T @ S_inv_rhat
            Q_x = Q_x + Q_yx.T @ S_inv_rhat
            Q_xx = Q_xx + Q_yx.T @ YSinv @ Q_yx
            Q_ux = Q_ux + Q_yu.T @ YSinv @ Q_yx
            Q_uu = Q_uu + Q_yu.T @ YSinv @ Q_yu
            inf_pr = max(inf_pr, float(np.abs(primal).max()))
            inf_comp = max(inf_comp, float(np.abs(comp).max()))
        s.ku[t], s.Ku[t] = k_u, K_u
        dV += np.array([k_u @ Q_u, 0.5 * k_u @ (Q_uu @ k_u)])
        Vx = Q_x + K_u.T @ Q_u + Q_ux.T @ k_u + K_u.T @ Q_uu @ k_u
        Vxx = Q_xx + K_u.T @ Q_ux + Q_ux.T @ K_u + K_u.T @ Q_uu @ K_u
        Vxx = 0.5 * (Vxx + Vxx.T)
        inf_du = max(inf_du, float(np.abs(Q_u).max()))
        step_norm = max(step_norm, float(np.abs(k_u).max()))
    s.dV = dV
    if s.cons:  # rolloutLinearPolicy (:368-392) + dS, dY (:1516-1538)
        dx = np.zeros(n)
        s.dS, s.dY = np.zeros((N, d)), np.zeros((N, d))
        for t in range(N):
            s.dS[t] = s.ks[t] + s.Ks[t] @ dx
            s.dY[t] = _clip(s.ky[t] + s.Ky[t] @ dx, -MAX_RATIO, MAX_RATIO)
            du = s.ku[t] + s.Ku[t] @ dx
            dx = A[t] @ dx + B[t] @ du
        s.inf_pr, s.inf_comp = inf_pr, inf_comp
    else:
        s.inf_pr = s.inf_comp = 0.0
    s.inf_du, s.step_norm = inf_du, step_norm
    return True


def _seq_lqr(Q, q, R, r, M, A, B):
    """solveSequentialLQR (:411-482), d = 0."""
    N = len(R)
    n, m = Q[0].shape[0], R[0].shape[0]
    K, k = np.zeros((N, m, n)), np.zeros((N, m))
    P, p = [None] * (N + 1), [None] * (N + 1)
    P[N], p[N] = 0.5 * (Q[N] + Q[N].T), q[N]
    for t in range(N - 1, -1, -1):
        BtP = B[t].T @ P[t + 1]
        Q_uu = 0.5 * (R[t] + BtP @ B[t] + R[t].T + B[t].T @ P[t + 1].T @ B[t])
        Q_ux = BtP @ A[t] + M[t].T
        Q_x = q[t] + A[t].T @ p[t + 1]
        Q_u = r[t] + B[t].T @ p[t + 1]
        K[t] = -np.linalg.solve(Q_uu, Q_ux)
        k[t] = -np.linalg.solve(Q_uu, Q_u)
        Pt = Q[t] + A[t].T @ P[t + 1] @ A[t] + Q_ux.T @ K[t] + K[t].T @ Q_ux + K[t].T @ Q_uu @ K[t]
        P[t] = 0.5 * (Pt + Pt.T)
        p[t] = Q_x + Q_ux.T @ k[t] + K[t].T @ Q_u + K[t].T @ Q_uu @ k[t]
    return K, k, P, np.array(p)


def _backward_teq(s, Vx, Vxx, d):
    """Terminal-equality branch (:1120-1353) + solveTerminalEqualityLQR (:484-639); H_T = I, b_T = -h_T."""
    P, N, n, m = s.P, s.P.N, s.P.n, s.P.m
    A, B = s.A, s.B
    hT = s.X[N] - s.xref
    inf_pr, inf_comp = float(np.abs(hT).max()), 0.0
    Q, q, R, r, M = [None] * (N + 1), [None] * (N + 1), [None] * N, [None] * N, [None] * N
    Q[N], q[N] = Vxx, Vx
    models = []
    for t in range(N):
        x, u = s.X[t], s.U[t]
        Q[t] = 0.5 * (2.0 * P.Qs + (2.0 * P.Qs).T)
        q[t] = 2.0 * P.Qs @ (x - s.xref)
        R[t] = 0.5 * (2.0 * P.Rs + (2.0 * P.Rs).T)
        r[t] = 2.0 * P.Rs @ u
        M[t] = np.zeros((n, m))
        if s.cons:
            y = np.concatenate([Y[t] for Y in s.Y])
            sl = np.concatenate([S[t] for S in s.S])
            g = np.concatenate([G[t] for G in s.G])
            jac = [c.jac(x, u) for c in s.cons]
            Q_yx, Q_yu = np.vstack([j[0] for j in jac]), np.vstack([j[1] for j in jac])
            s_safe = np.maximum(sl, max(s.mu * 1e-3, EPS_SLACK))
            YSinv = np.diag(_clip(y / s_safe, 0.0, MAX_RATIO))
            primal, comp = g + sl, y * sl - s.mu
            rhat = y * primal - comp
            Sir = _clip(rhat / s_safe, -MAX_RATIO, MAX_RATIO)
            q[t] = q[t] + Q_yx.T @ (y + Sir)
            r[t] = r[t] + Q_yu.T @ (y + Sir)
            Q[t] = Q[t] + Q_yx.T @ YSinv @ Q_yx
            M[t] = M[t] + (Q_yu.T @ YSinv @ Q_yx).T
            R[t] = R[t] + Q_yu.T @ YSinv @ Q_yu
            Q[t], R[t] = 0.5 * (Q[t] + Q[t].T), 0.5 * (R[t] + R[t].T)
            models.append((y, s_safe, Q_yx, Q_yu, YSinv, primal, rhat))
            inf_pr = max(inf_pr, float(np.abs(primal).max()))
            inf_comp = max(inf_comp, float(np.abs(comp).max()))
        R[t] = R[t] + s.reg * np.eye(m)
    qb = [v.copy() for v in q]
    qb[N] = qb[N] + s.lamT
    var, xT = [], []
    for i in range(n + 1):
        qv = [v.copy() for v in qb]
        if i > 0:
            qv[N][i - 1] += 1.0
        K, k, Pm, pm = _seq_lqr(Q, qv, R, r, M, A, B)
        dx = np.zeros(n)
        for t in range(N):
            dx = A[t] @ dx + B[t] @ (k[t] + K[t] @ dx)
        var.append((K, k, pm))
        xT.append(dx)
    S_mat = np.column_stack([xT[i + 1] - xT[0] for i in range(n)])
    rhs = -hT - xT[0]
    AtA, Atb = S_mat.T @ S_mat, S_mat.T @ rhs
    trace_term = AtA.trace() / max(n, 1) if AtA.trace() > 1.0 else 1.0
    base_floor = max(1e-10, s.io["jacobian_regularization_value"] * max(s.mu, 0.0) ** s.io["jacobian_regularization_exponent"])
    reg = max(base_floor, 1e-6 * trace_term)
    sv = np.linalg.svd(S_mat, compute_uv=False)
    reg_base = max(reg, max(1e-8 * sv.max() - sv.min(), 0.0))
    cap = 100.0 * (1.0 + np.linalg.norm(rhs))
    best, best_res, found = np.zeros(n), math.inf, False
    for scale in (1.0, 10.0, 100.0, 1e3, 1e4):
        lam = np.linalg.solve(AtA + max(reg_base * scale, 1e-12) * np.eye(n), Atb)
        if not np.isfinite(lam).all():
            continue
        ln = np.linalg.norm(lam)
        if ln > cap:
            lam = lam * (cap / max(ln, 1e-12))
        res = np.linalg.norm(S_mat @ lam - rhs)
        if not found or res < best_res:
            best, best_res, found = lam, res, True
    s.Ku, s.ku = var[0][0], var[0][1].copy()
    pout = var[0][2].copy()
    for i in range(n):
        s.ku += best[i] * (var[i + 1][1] - var[0][1])
        pout += best[i] * (var[i + 1][2] - var[0][2])
    s.dlamT = best
    s.inf_du = max(float(np.abs(r[t] + B[t].T @ pout[t + 1]).max()) for t in range(N))
    s.step_norm = float(np.abs(s.ku).max())
    s.ky, s.Ky, s.ks, s.Ks = np.zeros((N, d)), np.zeros((N, d, n)), np.zeros((N, d)), np.zeros((N, d, n))
    s.dS, s.dY = np.zeros((N, d)), np.zeros((N, d))
    dx = np.zeros(n)
    for t in range(N):
        if s.cons:
            y, s_safe, Q_yx, Q_yu, YSinv, primal, rhat = models[t]
            temp = Q_yu @ s.ku[t]
            s.ky[t] = _clip((rhat + y * temp) / s_safe, -MAX_RATIO, MAX_RATIO)
            s.Ky[t] = _clip(YSinv @ (Q_yx + Q_yu @ s.Ku[t]), -MAX_RATIO, MAX_RATIO)
            s.ks[t] = -primal - temp
            s.Ks[t] = -Q_yx - Q_yu @ s.Ku[t]
            s.dS[t] = s.ks[t] + s.Ks[t] @ dx
            s.dY[t] = _clip(s.ky[t] + s.Ky[t] @ dx, -MAX_RATIO, MAX_RATIO)
        dx = A[t] @ dx + B[t] @ (s.ku[t] + s.Ku[t] @ dx)
    s.inf_pr, s.inf_comp = inf_pr, inf_comp
    s.dV = np.zeros(2)
    return True


def _max_steps(s):  # computeMaxStepSizes :2939-2988
    if not s.cons:
        return 1.0, 1.0
    tau = max(s.io["min_fraction_to_boundary"], 1.0 - s.mu)
    S, Y = np.hstack(s.S), np.hstack(s.Y)
    apm = adm = 1.0
    neg = s.dS < 0.0
    if neg.any():
        apm = min(apm, float((-tau * S[neg] / s.dS[neg]).min()))
    neg = s.dY < 0.0
    if neg.any():
        adm = min(adm, float((-tau * Y[neg] / s.dY[neg]).min()))
    return min(max(apm, 0.0), 1.0), min(max(adm, 0.0), 1.0)


def _forward(s, alpha):
    """IPDDPSolver::forwardPass (:1571-1876).  Returns None on rejection, else a dict."""
    P, N, n = s.P, s.P.N, s.P.n
    apm, adm = _max_steps(s)
    tau = 1.0 if not s.cons else max(s.io["min_fraction_to_boundary"], 1.0 - s.mu)
    a_pr, a_du = min(alpha, apm), min(alpha, adm)
    X, U = np.zeros((N + 1, n)), np.zeros((N, P.m))
    X[0] = s.x0
    Sn = [sl.copy() for sl in s.S]
    Yn = [y.copy() for y in s.Y]
    for t in range(N):
        dx = X[t] - s.X[t]
        off = 0
        for i, c in enumerate(s.cons):
            sl_ = slice(off, off + c.dim)
            s_new = s.S[i][t] + a_pr * s.ks[t, sl_] + s.Ks[t, sl_] @ dx
            y_new = s.Y[i][t] + a_du * s.ky[t, sl_] + s.Ky[t, sl_] @ dx
            if (s_new < (1.0 - tau) * s.S[i][t]).any() or (y_new < (1.0 - tau) * s.Y[i][t]).any():
                return None
            if not (np.isfinite(s_new).all() and np.isfinite(y_new).all()):
                return None
            Sn[i][t], Yn[i][t] = s_new, y_new
            off += c.dim
        U[t] = s.U[t] + a_pr * s.ku[t] + s.Ku[t] @ dx
        X[t + 1] = P.step(X[t], U[t])
        if not (np.isfinite(X[t + 1]).all() and np.isfinite(U[t]).all()):
            return None
    cost = P.trajectory_cost(X, U, s.xref)
    Gn = [np.array([c.g(X[t], U[t]) for t in range(N)]) for c in s.cons]
    lam_new = s.lamT + a_pr * s.dlamT if s.teq else s.lamT
    h_new = _hT(s, X)
    phi = _merit(s, Sn, cost, lam_new if s.teq else None, h_new)
    theta = _theta(s, Gn, Sn, h_new)
    ip, ic = _primal_comp(s, Gn, Sn, Yn, s.mu, h_new)
    if not all(map(math.isfinite, (phi, theta, ip, ic))):
        return None
    if not s.cons and not s.teq:  # :1785-1794
        dJ = s.cost - cost
        expected = -a_pr * (s.dV[0] + 0.5 * a_pr * s.dV[1])
        ratio = dJ / expected if expected > 0.0 else math.copysign(1.0, dJ)
        ok = ratio > 1e-6
    else:  # :1796-1839
        io = s.io
        expected_improvement = a_pr * s.dV[0]
        cv_old = s.filter[-1][1] if s.filter else 0.0
        high_ref = cv_old if s.filter else s.filter_theta
        if theta > io["max_violation_threshold"]:
            ok = theta < (1 - io["violation_acceptance_threshold"]) * high_ref
        elif max(theta, cv_old) < io["min_violation_for_armijo_check"] and expected_improvement < 0:
            ok = phi < s.merit + s.o["armijo_constant"] * expected_improvement
        else:
            ok = (phi < s.merit - io["merit_acceptance_threshold"] * theta or
                  theta < (1 - io["violation_acceptance_threshold"]) * cv_old)
    if not ok:
        return None
    return dict(X=X, U=U, S=Sn, Y=Yn, G=Gn, cost=cost, merit=phi, theta=theta, inf_pr=ip, inf_comp=ic, a_pr=a_pr, a_du=a_du,
                lamT=lam_new)


def _accept_filter(f, merit, theta):  # interior_point_utils.cpp:81-97
    if any(m <= merit and t <= theta for m, t in f):
        return
    f[:] = [(m, t) for m, t in f if not (merit <= m and theta <= t)]
    f.append((merit, theta))


def _prune(f):  # interior_point_utils.cpp:116-141
    if not f:
        return
    bv = min(f, key=lambda p: p[1])
    bm = min(f, key=lambda p: p[0])
    f[:] = [bv]
    if abs(bm[1] - bv[1]) > 1e-12 or abs(bm[0] - bv[0]) > 1e-12:
        f.append(bm)


def _apply(s, r):
    """applyForwardPassResult (:1878-1951) + updateBarrierParameters(true) (:2548-2660)."""
    s.X, s.U, s.cost, s.merit = r["X"], r["U"], r["cost"], r["merit"]
    s.alpha_pr, s.alpha_du = r["a_pr"], r["a_du"]
    s.S, s.Y, s.G, s.lamT = r["S"], r["Y"], r["G"], r["lamT"]
    s.inf_pr, s.inf_comp, s.phi, s.filter_theta = r["inf_pr"], r["inf_comp"], r["merit"], r["theta"]
    io, mu_old = s.io, s.mu
    if s.cons:
        if io["barrier_strategy"] == 0:
            kkt = max(s.inf_pr, s.inf_du, s.inf_comp)
            if kkt <= max(io["mu_update_factor"] * s.mu, 2.0 * s.mu):
                factor = io["mu_update_factor"]
                if s.mu > 1e-20:
                    ratio = kkt / max(s.mu, 1e-20)
                    if ratio < 0.01:
                        factor = 0.1 * io["mu_update_factor"]
                    elif ratio < 0.1:
                        factor = 0.3 * io["mu_update_factor"]
                    elif ratio < 0.5:
                        factor = 0.6 * io["mu_update_factor"]
                s.mu = max(min(factor * s.mu, s.mu ** io["mu_update_power"]), max(io["mu_min_value"], s.o["tolerance"] / 100.0))
        else:
            kkt = max(s.inf_pr, s.inf_du * io["barrier_update_dual_weight"], s.inf_comp)
            if kkt <= io["mu_kappa_epsilon"] * s.mu:
                s.mu = max(io["mu_min_value"], min(io["mu_update_factor"] * s.mu, s.mu ** io["mu_update_power"]))
    h = _hT(s, s.X)
    filter_theta = max(_theta(s, s.G, s.S, h), 1e-8)
    if s.mu < mu_old and s.mu > 0.0:
        s.filter = []
        if s.teq:
            _accept_filter(s.filter, s.phi, filter_theta)
    else:
        _accept_filter(s.filter, s.phi, filter_theta)
        if len(s.filter) > io["max_filter_size"]:
            _prune(s.filter)
    s.inf_pr, s.inf_comp = _primal_comp(s, s.G, s.S, s.Y, s.mu, h)
    s.merit = s.phi = _merit(s, s.S, s.cost, s.lamT if s.teq else None, h)
    s.filter_theta = filter_theta


def _flat(s):
    cat = lambda L: np.hstack(L) if L else np.zeros((s.P.N, 0))  # noqa: E731
    apm, adm = _max_steps(s)
    return dict(X=s.X, U=s.U, Y=cat(s.Y), S=cat(s.S), G=cat(s.G), ku=s.ku, Ku=s.Ku, ky=s.ky, Ky=s.Ky, ks=s.ks, Ks=s.Ks,
                mu=s.mu, cost=s.cost, merit=s.merit, inf_pr=s.inf_pr, inf_du=s.inf_du, inf_comp=s.inf_comp,
                step_norm=s.step_norm, reg=s.reg, dV0=s.dV[0], dV1=s.dV[1], alpha_pr_max=apm, alpha_du_max=adm,
                lamT=s.lamT, dlamT=s.dlamT)


def probe(spec, opts, ipopts, constraints, x0, xref, U0, iters):
    """initialize + `iters` iterations (no convergence tests) + one backward pass (mirrors oracle_ipddp_probe)."""
    s = _setup(spec, opts, ipopts, constraints, x0, xref, U0)
    alphas = npo.build_alphas(s.o)
    for it in range(iters + 1):
        _backward(s)
        if it == iters:
            break
        r = next((r for r in (_forward(s, a) for a in alphas) if r is not None), None)
        if r is not None:
            _apply(s, r)
            s.reg = max(s.reg / s.o["reg_update_factor"], s.o["reg_min_value"])
        else:
            s.reg = min(s.reg * s.o["reg_update_factor"], s.o["reg_max_value"])
            if s.cons and s.teq:
                s.reg = min(s.reg * s.o["reg_update_factor"], s.o["reg_max_value"])
    return _flat(s)


def solve(spec, opts, ipopts, constraints, x0, xref, U0):
    """CDDP::solve("IPDDP"): CDDPSolverBase::solve (cddp_solver_base.cpp:29-186) with the IPDDP hooks."""
    s = _setup(spec, opts, ipopts, constraints, x0, xref, U0)
    o, io = s.o, s.io
    alphas = npo.build_alphas(o)
    no_barrier = not s.cons
    it, status = 0, npo.MAX_ITERATIONS
    while it < o["max_iterations"]:
        it += 1
        _backward(s)  # LU never reports failure; Eigen::LDLT only does on an exact zero pivot
        if no_barrier:  # checkEarlyConvergence :925-958
            early = s.inf_pr < o["tolerance"] and s.inf_du < o["tolerance"]
        else:
            tol = max(o["tolerance"], io["barrier_tol_mult"] * s.mu)
            early = (s.inf_pr < tol and s.inf_du < tol and s.inf_comp < tol and
                     abs(s.alpha_pr) * s.step_norm < o["tolerance"] * 10.0)
        if early:
            status = npo.OPTIMAL
            break
        r = None
        if not o["enable_parallel"]:
            for a in alphas:  # first success wins, cddp_solver_base.cpp:255-263
                r = _forward(s, a)
                if r is not None:
                    break
        else:  # lowest merit among the successes, scanning in alpha order (:264-285)
            for a in alphas:
                c = _forward(s, a)
                if c is not None and (r is None or c["merit"] < r["merit"]):
                    r = c
        if r is not None:
            dJ = s.cost - r["cost"]
            _apply(s, r)
            s.reg = max(s.reg / o["reg_update_factor"], o["reg_min_value"])
            if no_barrier:  # checkConvergence :1953-2025
                if s.inf_pr < o["tolerance"] and s.inf_du < o["tolerance"]:
                    status = npo.OPTIMAL
                    break
                if o["acceptable_tolerance"] > 0.0:
                    sq = math.sqrt(o["acceptable_tolerance"])
                    acc = s.inf_pr < sq and s.inf_du < sq and it > 50
                    if dJ > 0.0:
                        acc = acc or (dJ < o["acceptable_tolerance"] and it > 50 and s.inf_pr < sq and s.inf_du < sq)
                    if acc:
                        status = npo.ACCEPTABLE
                        break
            else:
                tol = max(o["tolerance"], io["barrier_tol_mult"] * s.mu)
                if s.inf_pr < tol and s.inf_du < tol and s.inf_comp < tol and s.step_norm < o["tolerance"] * 10.0:
                    status = npo.OPTIMAL
                    break
                if o["acceptable_tolerance"] > 0.0:
                    at = math.sqrt(o["acceptable_tolerance"])
                    bat = max(io["mu_min_value"] * 100.0, o["tolerance"] / 10.0)
                    kkt = s.inf_pr < at and s.inf_du < at and s.inf_comp < at
                    done = s.mu <= bat
                    acc = kkt and done and it > 10 and abs(dJ) < o["acceptable_tolerance"]
                    acc = acc or (kkt and done and it >= 1 and s.step_norm < o["tolerance"] * 10.0 and s.inf_pr < 1e-4)
                    if acc:
                        status = npo.ACCEPTABLE
                        break
        else:  # handleForwardPassFailure :2037-2082
            s.reg = min(s.reg * o["reg_update_factor"], o["reg_max_value"])
            if not no_barrier and s.teq:
                s.reg = min(s.reg * o["reg_update_factor"], o["reg_max_value"])
            if s.reg >= o["reg_max_value"]:
                base = math.sqrt(max(o["acceptable_tolerance"], o["tolerance"]))
                at = base if no_barrier else max(base, io["barrier_tol_mult"] * s.mu)
                acc = (o["acceptable_tolerance"] > 0.0 and s.inf_pr < at and s.inf_du < at and
                       (no_barrier or s.inf_comp < at))
                status = npo.ACCEPTABLE if acc else npo.REG_LIMIT
                break
    out = _flat(s)
    out.update(iterations=it, status=status)
    return out
