"""oracle/np_oracle.py — SECOND, independent CPU restatement of the reference CLDDP hot path (numpy).

TEST INFRASTRUCTURE ONLY (see oracle/cddp_oracle.h).  It exists to pin the C++ oracle
(oracle/cddp_oracle.cpp): the two were written separately from the cited reference lines and must
agree.  Where the C++ oracle hand-derives things, this file deliberately takes a different route:

  * Jacobians: complex-step differentiation of the continuous dynamics (exact to roundoff), instead
    of dual numbers / closed forms.  The reference uses autodiff for cartpole and quadrotor
    (cartpole.cpp:95-103, quadrotor.cpp:116-140) and closed forms for pendulum/unicycle.
  * PD test: numpy.linalg.eigvals(Q_uu_reg).real.min() <= 0 — the same statement as the reference's
    Eigen::EigenSolver test (clddp_solver.cpp:133-134).
  * Unconstrained gains: numpy.linalg.inv then multiply (clddp_solver.cpp:142-145).
  * BoxQP free-block solves: numpy.linalg.solve on the gathered free block (boxqp.cpp:89-111,147).

PARITY STATUS: "parity unpinned" — neither restatement can be compared with the real reference
binary (Eigen/autodiff unavailable in this image).  Citations: astomodynamics/cddp-cpp @ f71fa80.
"""
from __future__ import annotations

import math

import numpy as np

# ------------------------------------------------------------------------------------------------
# options (options.hpp:41-66, :93-105, :208-251; boxqp.hpp:30-41)
# ------------------------------------------------------------------------------------------------
DEFAULTS = dict(
    tolerance=1e-5, acceptable_tolerance=1e-6, max_iterations=1, enable_parallel=0, max_cpu_time=0.0,
    termination_scaling_max_factor=100.0, ls_max_iterations=11, ls_initial_step_size=1.0, ls_min_step_size=1e-8,
    ls_step_reduction_factor=0.5, reg_initial_value=1e-6, reg_update_factor=10.0, reg_max_value=1e7,
    reg_min_value=1e-10, qp_max_iterations=100, qp_min_gradient_norm=1e-8, qp_min_relative_improvement=1e-8,
    qp_step_decrease_factor=0.6, qp_min_step_size=1e-22, qp_armijo_constant=0.1, armijo_constant=1e-4,
)

RUNNING, OPTIMAL, ACCEPTABLE, MAX_ITERATIONS, REG_LIMIT, MAX_CPU_TIME = range(6)
QP_HESSIAN_NOT_PD, QP_NO_DESCENT, QP_MAX_ITER, QP_MAX_LS, QP_NO_BOUNDS, QP_SUCCESS, QP_ALL_CLAMPED = -1, 0, 1, 2, 3, 4, 5


def options(**kw):
    o = dict(DEFAULTS)
    for k in kw:
        if k not in o:
            raise KeyError(k)
    o.update(kw)
    return o


def build_alphas(o):
    """detail::buildLineSearchAlphas, cddp_context_utils.cpp:37-57."""
    out, a = [], o["ls_initial_step_size"]
    for i in range(o["ls_max_iterations"]):
        out.append(a)
        a *= o["ls_step_reduction_factor"]
        if a < o["ls_min_step_size"] and i < o["ls_max_iterations"] - 1:
            out.append(o["ls_min_step_size"])
            break
    if not out:
        out.append(o["ls_initial_step_size"])
    return np.array(out)


# ------------------------------------------------------------------------------------------------
# models: continuous dynamics written for real OR complex arguments
# ------------------------------------------------------------------------------------------------
def f_pendulum(p, x, u, jac=False):  # pendulum.cpp:29-42
    length, mass, damping = p[0], p[1], p[2]
    g = 9.81
    inertia = mass * length * length
    return np.array([x[1], (u[0] - damping * x[1] + mass * g * length * np.sin(x[0])) / inertia])


def f_cartpole(p, x, u, jac=False):  # cartpole.cpp:38-63 (double) / :65-93 (autodiff: adds damping, :90)
    mc, mp, l, g, d = p[0], p[1], p[2], p[3], p[4]
    th, w, F = x[1], x[3], u[0]
    s, c = np.sin(th), np.cos(th)
    den = mc + mp * s * s
    xdd = (F + mp * s * (l * w * w + g * c)) / den
    num = -F * c - mp * l * w * w * c * s - (mc + mp) * g * s
    if jac:
        num = num - d * w
    thdd = num / (l * den)
    return np.array([x[2], w, xdd, thdd])


def f_unicycle(p, x, u, jac=False):  # unicycle.cpp:28-38
    return np.array([u[0] * np.cos(x[2]), u[0] * np.sin(x[2]), u[1] + 0 * x[0]])


def f_quadrotor(p, x, u, jac=False):  # quadrotor.cpp:33-96
    mass, L = p[0], p[10]
    I = np.asarray(p[1:10], dtype=float).reshape(3, 3)
    q = x[3:7]
    nq = np.sqrt(q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3])
    if abs(nq) > 1e-6:
        qw, qx, qy, qz = q / nq
    else:
        qw, qx, qy, qz = 1.0, 0.0, 0.0, 0.0
    w = x[10:13]
    thrust = u[0] + u[1] + u[2] + u[3]
    tau = np.array([L * (u[0] - u[2]), L * (u[1] - u[3]), 0.1 * (u[0] - u[1] + u[2] - u[3])])
    out = np.zeros(13, dtype=np.result_type(x.dtype, u.dtype))
    out[0:3] = x[7:10]
    out[3] = -0.5 * (qx * w[0] + qy * w[1] + qz * w[2])
    out[4] = 0.5 * (qw * w[0] + qy * w[2] - qz * w[1])
    out[5] = 0.5 * (qw * w[1] - qx * w[2] + qz * w[0])
    out[6] = 0.5 * (qw * w[2] + qx * w[1] - qy * w[0])
    # thrust along the body z axis = third column of R(q)
    r3 = np.array([2 * (qx * qz + qy * qw), 2 * (qy * qz - qx * qw), 1 - 2 * (qx * qx + qy * qy)])
    out[7:10] = r3 * thrust / mass
    out[9] = out[9] - 9.81
    Iw = I @ w
    cross = np.array([w[1] * Iw[2] - w[2] * Iw[1], w[2] * Iw[0] - w[0] * Iw[2], w[0] * Iw[1] - w[1] * Iw[0]])
    out[10:13] = np.linalg.inv(I) @ (tau - cross)
    return out


def f_bicycle(p, x, u, jac=False):  # bicycle.cpp:29-47; params: wheelbase
    return np.array([x[3] * np.cos(x[2]), x[3] * np.sin(x[2]), (x[3] / p[0]) * np.tan(u[1]), u[0] + 0 * x[0]])


def f_chain7(p, x, u, jac=False):  # test plugin model (not in the reference), see oracle/cddp_oracle.cpp chain7_f
    g, c, k, inertia = p[0], p[1], p[2], np.asarray(p[3:10])
    q, qd = x[:7], x[7:]
    acc = u - g * np.sin(q) - c * qd
    acc = acc - k * np.concatenate([[0.0], np.sin(q[1:] - q[:-1])]) - k * np.concatenate([np.sin(q[:-1] - q[1:]), [0.0]])
    return np.concatenate([qd, acc / inertia])


def f_manip7(p, x, u, jac=False):
    """7-DOF manipulator (BASELINE config #5), matrix form: M(q) qdd + h(q,qd) + G(q) + b qd = tau with
    M_ij = mu_max(i,j) l_i l_j cos(sigma_i - sigma_j), h from the Christoffel symbols of M, see oracle/cddp_oracle.cpp manip7_f.
    Third, independent statement: h is formed from dM/dq_k tensors."""
    g, bv = p[0], p[1]
    mass, ln = np.asarray(p[2:9], float), np.asarray(p[9:16], float)
    q, qd = x[:7], x[7:]
    mu = np.cumsum(mass[::-1])[::-1]
    T = np.tril(np.ones((7, 7)))
    T[:, 0] = 0.0  # sigma_j = q_1 + ... + q_j
    sig = T @ q
    A = np.array([[mu[max(i, j)] * ln[i] * ln[j] for j in range(7)] for i in range(7)])
    D = sig[:, None] - sig[None, :]
    M = A * np.cos(D)
    # dM/dq_k = -A sin(D) * (dsig_i/dq_k - dsig_j/dq_k)
    dM = [-(A * np.sin(D)) * (T[:, k][:, None] - T[:, k][None, :]) for k in range(7)]
    Mdot = sum(dM[k] * qd[k] for k in range(7))
    h = Mdot @ qd - 0.5 * np.array([qd @ dM[k] @ qd for k in range(7)])
    Gv = np.array([0.0 if k == 0 else -np.sum(mu[k:] * g * ln[k:] * np.cos(sig[k:])) for k in range(7)], dtype=M.dtype)
    qdd = np.linalg.solve(M, u - h - Gv - bv * qd)
    return np.concatenate([qd, qdd])


MODELS = {"pendulum": f_pendulum, "cartpole": f_cartpole, "unicycle": f_unicycle, "quadrotor": f_quadrotor,
          "bicycle": f_bicycle, "chain7": f_chain7, "manip7": f_manip7}


class Problem:
    def __init__(self, spec):
        self.spec = spec
        self.model = spec.get("twin_model", spec["model"])
        self.n, self.m, self.N = int(spec["n"]), int(spec["m"]), int(spec["horizon"])
        self.dt = float(spec["dt"])
        self.integrator = spec.get("integrator", "rk4")
        self.params = list(spec.get("params", [])) + [0.0] * 16
        # QuadraticObjective ctor: Q_ = Q*dt, R_ = R*dt (objective.cpp:38-39)
        self.Qs = np.asarray(spec["Q"], float) * self.dt
        self.Rs = np.asarray(spec["R"], float) * self.dt
        self.Qf = np.asarray(spec["Qf"], float)
        self.lb = None if spec.get("lb") is None else np.asarray(spec["lb"], float)
        self.ub = None if spec.get("ub") is None else np.asarray(spec["ub"], float)
        self.Ad = None if spec.get("lti_A") is None else np.asarray(spec["lti_A"], float)
        self.Bd = None if spec.get("lti_B") is None else np.asarray(spec["lti_B"], float)

    # --- dynamics ---
    def f(self, x, u, jac=False):
        return MODELS[self.model](self.params, x, u, jac)

    def step(self, x, u):
        """getDiscreteDynamics -> euler/heun/rk3/rk4 (dynamical_system.cpp:28-83); LTI: lti_system.cpp:71-76."""
        if self.model == "lti":
            return self.Ad @ x + self.Bd @ u
        dt, f = self.dt, self.f
        it = self.integrator
        if it == "euler":
            return x + dt * f(x, u)
        if it == "heun":
            k1 = f(x, u)
            k2 = f(x + dt * k1, u)
            return x + 0.5 * dt * (k1 + k2)
        if it == "rk3":
            k1 = f(x, u)
            k2 = f(x + 0.5 * dt * k1, u)
            k3 = f(x - dt * k1 + 2 * dt * k2, u)
            return x + (dt / 6.0) * (k1 + 4 * k2 + k3)
        k1 = f(x, u)
        k2 = f(x + 0.5 * dt * k1, u)
        k3 = f(x + 0.5 * dt * k2, u)
        k4 = f(x + dt * k3, u)
        return x + (dt / 6.0) * (k1 + 2 * k2 + 2 * k3 + k4)

    def jacobians(self, x, u):
        """CONTINUOUS-time Jacobians (Fx, Fu) by complex step; LTI per lti_system.cpp:78-92."""
        n, m = self.n, self.m
        if self.model == "lti":
            return (self.Ad - np.eye(n)) / self.dt, self.Bd / self.dt
        h = 1e-30
        Fx, Fu = np.zeros((n, n)), np.zeros((n, m))
        xc, uc = x.astype(complex), u.astype(complex)
        for j in range(n):
            xp = xc.copy()
            xp[j] += 1j * h
            Fx[:, j] = np.imag(self.f(xp, uc, jac=True)) / h
        for j in range(m):
            up = uc.copy()
            up[j] += 1j * h
            Fu[:, j] = np.imag(self.f(xc, up, jac=True)) / h
        return Fx, Fu

    # --- objective (objective.cpp:80-154) ---
    def ref_at(self, xref, ref_traj, t):
        return xref if ref_traj is None else ref_traj[t]

    def running_cost(self, x, u, ref):
        e = x - ref
        return float((e @ self.Qs) @ e + (u @ self.Rs) @ u)

    def terminal_cost(self, x, xref):
        e = x - xref
        return float((e @ self.Qf) @ e)

    def trajectory_cost(self, X, U, xref, ref_traj=None):
        J = 0.0
        for t in range(self.N):
            J += self.running_cost(X[t], U[t], self.ref_at(xref, ref_traj, t))
        return J + self.terminal_cost(X[self.N], xref)


# ------------------------------------------------------------------------------------------------
# BoxQP (boxqp.cpp:25-250)
# ------------------------------------------------------------------------------------------------
def boxqp(o, H, g, lower, upper, x0=None):
    n = len(g)
    if x0 is not None and len(x0) == n:
        x = np.minimum(np.maximum(x0, lower), upper)
    else:
        x = np.zeros(n)
        for i in range(n):
            fl, fu = math.isfinite(lower[i]), math.isfinite(upper[i])
            x[i] = 0.5 * (lower[i] + upper[i]) if (fl and fu) else lower[i] if fl else upper[i] if fu else 0.0
    val = lambda z: 0.5 * z @ (H @ z) + g @ z  # noqa: E731
    status = QP_MAX_ITER
    clamped = np.zeros(n, dtype=bool)
    free = np.ones(n, dtype=bool)
    value, old = val(x), math.inf
    Hf_idx = None  # indices of the block the current factor belongs to
    iters = facts = 0
    for it in range(o["qp_max_iterations"]):
        iters = it + 1
        if it > 0 and abs(old - value) < o["qp_min_relative_improvement"] * abs(old):
            status = QP_SUCCESS
            break
        old = value
        grad = g + H @ x
        old_clamped = clamped
        clamped = ((x == lower) & (grad > 0)) | ((x == upper) & (grad < 0))
        free = ~clamped
        if clamped.all():
            status = QP_ALL_CLAMPED
            break
        if it == 0 or (old_clamped != clamped).any():
            Hf_idx = np.flatnonzero(free)
            Hff = H[np.ix_(Hf_idx, Hf_idx)]
            if not np.all(np.linalg.eigvalsh(0.5 * (Hff + Hff.T)) > 0):
                status = QP_HESSIAN_NOT_PD
                break
            facts += 1
        gn = math.sqrt(float(np.sum(grad[free] ** 2)))
        if gn < o["qp_min_gradient_norm"]:
            status = QP_SUCCESS
            break
        gc = g + H[:, clamped] @ x[clamped]
        search = np.zeros(n)
        fi = np.flatnonzero(free)
        search[fi] = -np.linalg.solve(H[np.ix_(fi, fi)], gc[fi]) - x[fi]
        sdotg = float(search @ grad)
        if sdotg >= 0:
            status = QP_NO_DESCENT
            break
        step, accepted = 1.0, None
        while step > o["qp_min_step_size"]:
            xn = np.minimum(np.maximum(x + step * search, lower), upper)
            if val(xn) - value <= o["qp_armijo_constant"] * step * sdotg:
                accepted = xn
                break
            step *= o["qp_step_decrease_factor"]
        if accepted is None:
            status = QP_MAX_LS
            break
        x = accepted
        value = val(x)
    return dict(status=status, x=x, free=free, factor_idx=Hf_idx, iterations=iters, factorizations=facts, value=value)


# ------------------------------------------------------------------------------------------------
# CLDDP backward / forward / solve
# ------------------------------------------------------------------------------------------------
def backward_pass(P, o, X, U, xref, reg, k_prev, ref_traj=None, AB=None):
    """clddp_solver.cpp:79-204.  Returns dict(ok, K, k, dV, inf_du, Vx0, Vxx0)."""
    n, m, N = P.n, P.m, P.N
    Vx = 2.0 * P.Qf @ (X[N] - xref)
    Vxx = 2.0 * P.Qf
    K_all, k_all = np.zeros((N, m, n)), np.array(k_prev, dtype=float).copy()
    dV = np.zeros(2)
    norm_Vx, Qu_err = float(np.abs(Vx).sum()), 0.0
    for t in range(N - 1, -1, -1):
        x, u = X[t], U[t]
        if AB is None:
            Fx, Fu = P.jacobians(x, u)
            A = P.dt * Fx
            A[np.diag_indices(n)] += 1.0
            B = P.dt * Fu
        else:
            A, B = AB[0][t], AB[1][t]
        ref = P.ref_at(xref, ref_traj, t)
        lx, lu = 2.0 * P.Qs @ (x - ref), 2.0 * P.Rs @ u
        lxx, luu = 2.0 * P.Qs, 2.0 * P.Rs
        Qx = lx + A.T @ Vx
        Qu = lu + B.T @ Vx
        Qxx = lxx + A.T @ Vxx @ A
        Qux = B.T @ Vxx @ A
        Quu = luu + B.T @ Vxx @ B
        Quu_reg = Quu + reg * np.eye(m)
        if np.linalg.eigvals(Quu_reg).real.min() <= 0:
            return dict(ok=False, fail_t=t)
        if P.lb is None:
            Hinv = np.linalg.inv(Quu_reg)
            k = -Hinv @ Qu
            K = -Hinv @ Qux
        else:
            r = boxqp(o, Quu_reg, Qu, P.lb - u, P.ub - u, k_all[t])
            if r["status"] in (QP_HESSIAN_NOT_PD, QP_NO_DESCENT):
                return dict(ok=False, fail_t=t)
            k = r["x"]
            K = np.zeros((m, n))
            fr = np.flatnonzero(r["free"])
            if fr.size:
                # result.Hfree is the LAST factor computed; with the defaults it always belongs to the final
                # free set except on the relative-improvement exit (stale by one iteration) — same block here
                fi = r["factor_idx"]
                if fi is not None and fi.size == fr.size and (fi == fr).all():
                    K[fr] = -np.linalg.solve(Quu_reg[np.ix_(fr, fr)], Qux[fr])
                else:  # stale factor of a different block size: the reference would solve with mismatched
                    # dimensions (Eigen assertion / UB); neither oracle defines it — flag it
                    return dict(ok=False, fail_t=t, stale_factor=True)
        k_all[t], K_all[t] = k, K
        dV += np.array([Qu @ k, 0.5 * k @ (Quu @ k)])
        Vx = Qx + K.T @ Quu @ k + Qux.T @ k + K.T @ Qu
        Vxx = Qxx + K.T @ Quu @ K + Qux.T @ K + K.T @ Qux
        Vxx = 0.5 * (Vxx + Vxx.T)
        norm_Vx += float(np.abs(Vx).sum())
        Qu_err = max(Qu_err, float(np.abs(Qu).max()))
    sf = o["termination_scaling_max_factor"]
    sf = max(sf, norm_Vx / (N * n)) / sf
    return dict(ok=True, K=K_all, k=k_all, dV=dV, inf_du=Qu_err / sf, Vx0=Vx, Vxx0=Vxx)


def forward_pass(P, o, x0, X, U, xref, K, k, dV, cost, alpha, ref_traj=None):
    """clddp_solver.cpp:215-262."""
    N = P.N
    Xn, Un = X.copy(), U.copy()
    Xn[0] = x0
    J = 0.0
    for t in range(N):
        x = Xn[t]
        u = Un[t] + alpha * k[t] + K[t] @ (x - X[t])
        if P.lb is not None:
            u = np.minimum(np.maximum(u, P.lb), P.ub)
        Un[t] = u
        J += P.running_cost(x, u, P.ref_at(xref, ref_traj, t))
        Xn[t + 1] = P.step(x, u)
    J += P.terminal_cost(Xn[N], xref)
    dJ = cost - J
    expected = -alpha * (dV[0] + 0.5 * alpha * dV[1])
    ratio = dJ / expected if expected > 0.0 else math.copysign(1.0, dJ)
    return dict(success=bool(ratio > o["armijo_constant"]), X=Xn, U=Un, cost=J)


def solve(P, o, x0, xref, X0, U0, ref_traj=None):
    """CDDP::solve("CLDDP"): cddp_core.cpp:235-306 + clddp_solver.cpp:28-75 + cddp_solver_base.cpp:29-186."""
    X, U = np.array(X0, dtype=float).copy(), np.array(U0, dtype=float).copy()
    X[0] = x0  # initializeProblemIfNecessary, cddp_core.cpp:294
    N, m, n = P.N, P.m, P.n
    K, k = np.zeros((N, m, n)), np.zeros((N, m))
    cost = P.trajectory_cost(X, U, xref, ref_traj)  # no re-rollout (clddp_solver.cpp:72-74)
    reg = o["reg_initial_value"]
    alpha_pr = o["ls_initial_step_size"]
    alphas = build_alphas(o)
    inf_du = math.inf
    status, it = MAX_ITERATIONS, 0
    hist = [(cost, alpha_pr, inf_du, reg)]
    while it < o["max_iterations"]:
        it += 1
        bw = None
        while True:
            bw = backward_pass(P, o, X, U, xref, reg, k, ref_traj)
            if bw["ok"]:
                break
            reg = min(reg * o["reg_update_factor"], o["reg_max_value"])  # cddp_core.cpp:308-314
            if reg >= o["reg_max_value"]:
                status = REG_LIMIT
                break
        if not bw["ok"]:
            break
        K, k, dV, inf_du = bw["K"], bw["k"], bw["dV"], bw["inf_du"]
        if inf_du < o["tolerance"]:  # checkEarlyConvergence
            status = OPTIMAL
            hist.append((cost, alpha_pr, inf_du, reg))
            break
        best = None
        for a in alphas:
            r = forward_pass(P, o, x0, X, U, xref, K, k, dV, cost, a, ref_traj)
            if not r["success"]:
                continue
            if not o["enable_parallel"]:
                best = (a, r)
                break
            if best is None or r["cost"] < best[1]["cost"]:
                best = (a, r)
        if best is not None:
            a, r = best
            dJ = cost - r["cost"]
            X, U, cost, alpha_pr = r["X"], r["U"], r["cost"], a
            hist.append((cost, alpha_pr, inf_du, reg))
            reg = max(reg / o["reg_update_factor"], o["reg_min_value"])
            if inf_du < o["tolerance"]:
                status = OPTIMAL
                break
            if 0.0 < dJ < o["acceptable_tolerance"]:
                status = ACCEPTABLE
                break
        else:
            reg = min(reg * o["reg_update_factor"], o["reg_max_value"])
            if reg >= o["reg_max_value"]:
                status = REG_LIMIT
                break
    return dict(X=X, U=U, K=K, k=k, cost=cost, alpha=alpha_pr, reg=reg, inf_du=inf_du, iterations=it, status=status,
                history=np.array(hist))
