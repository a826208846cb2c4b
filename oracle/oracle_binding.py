"""ctypes binding of the CPU oracle (oracle/liboracle.so).

TEST INFRASTRUCTURE ONLY: importable from tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs.  The product package never imports this module.
The problem/option dictionaries are the ones cddp-cpp_b200/problems.py produces.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "liboracle.so")

MODEL_IDS = {"pendulum": 0, "cartpole": 1, "unicycle": 2, "quadrotor": 3, "lti": 4, "bicycle": 5, "chain7": 6, "manip7": 7}
INTEGRATORS = {"euler": 0, "heun": 1, "rk3": 2, "rk4": 3}
STATUS_STRINGS = {0: "Running", 1: "OptimalSolutionFound", 2: "AcceptableSolutionFound", 3: "MaxIterationsReached",
                  4: "RegularizationLimitReached_NotConverged", 5: "MaxCpuTimeReached"}
QP_STATUS = {-1: "HESSIAN_NOT_PD", 0: "NO_DESCENT", 1: "MAX_ITER_EXCEEDED", 2: "MAX_LS_EXCEEDED", 3: "NO_BOUNDS",
             4: "SUCCESS", 5: "ALL_CLAMPED"}


class Options(C.Structure):
    _fields_ = [
        ("tolerance", C.c_double), ("acceptable_tolerance", C.c_double), ("max_iterations", C.c_int),
        ("enable_parallel", C.c_int), ("max_cpu_time", C.c_double), ("termination_scaling_max_factor", C.c_double),
        ("ls_max_iterations", C.c_int), ("reserved1", C.c_int), ("ls_initial_step_size", C.c_double),
        ("ls_min_step_size", C.c_double), ("ls_step_reduction_factor", C.c_double), ("reg_initial_value", C.c_double),
        ("reg_update_factor", C.c_double), ("reg_max_value", C.c_double), ("reg_min_value", C.c_double),
        ("qp_max_iterations", C.c_int), ("reserved2", C.c_int), ("qp_min_gradient_norm", C.c_double),
        ("qp_min_relative_improvement", C.c_double), ("qp_step_decrease_factor", C.c_double),
        ("qp_min_step_size", C.c_double), ("qp_armijo_constant", C.c_double), ("armijo_constant", C.c_double),
    ]


class Problem(C.Structure):
    _fields_ = [
        ("model", C.c_int), ("n", C.c_int), ("m", C.c_int), ("horizon", C.c_int), ("dt", C.c_double),
        ("integrator", C.c_int), ("has_control_box", C.c_int), ("model_params", C.c_double * 16),
        ("lti_A", C.POINTER(C.c_double)), ("lti_B", C.POINTER(C.c_double)), ("Q", C.POINTER(C.c_double)),
        ("R", C.POINTER(C.c_double)), ("Qf", C.POINTER(C.c_double)), ("lb", C.POINTER(C.c_double)),
        ("ub", C.POINTER(C.c_double)),
    ]


class Result(C.Structure):
    _fields_ = [("final_objective", C.c_double), ("final_step_length", C.c_double),
                ("final_regularization", C.c_double), ("inf_du", C.c_double), ("iterations", C.c_int),
                ("status", C.c_int), ("history_len", C.c_int), ("reserved", C.c_int)]


_lib = None


def build():
    subprocess.run(["make", "-C", _HERE, "-s"], check=True)


def load():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        build()
    lib = C.CDLL(LIB_PATH)
    lib.oracle_running_cost.restype = C.c_double
    lib.oracle_terminal_cost.restype = C.c_double
    lib.oracle_trajectory_cost.restype = C.c_double
    lib.oracle_status_string.restype = C.c_char_p
    _lib = lib
    return lib


class variant:
    """Context manager: route calls to liboracle_fma.so (same algorithm, fp contraction on)."""

    def __enter__(self):
        global _lib
        self.prev = load()
        path = os.path.join(_HERE, "liboracle_fma.so")
        if not os.path.exists(path):
            build()
        lib = C.CDLL(path)
        lib.oracle_running_cost.restype = C.c_double
        lib.oracle_terminal_cost.restype = C.c_double
        lib.oracle_trajectory_cost.restype = C.c_double
        lib.oracle_status_string.restype = C.c_char_p
        _lib = lib
        return self

    def __exit__(self, *exc):
        global _lib
        _lib = self.prev


def _f64(a):
    return np.ascontiguousarray(np.asarray(a, dtype=np.float64))


def _p(a):
    return None if a is None else C.c_void_p(a.ctypes.data)


def make_options(**overrides) -> Options:
    o = Options()
    load().oracle_default_options(C.byref(o))
    for k, v in overrides.items():
        if not hasattr(o, k):
            raise AttributeError(k)
        setattr(o, k, v)
    return o


class OracleProblem:
    def __init__(self, spec: dict):
        self.spec = spec
        n, m = int(spec["n"]), int(spec["m"])
        self.n, self.m, self.N = n, m, int(spec["horizon"])
        self._keep = {}
        p = Problem()
        model = spec.get("twin_model", spec["model"])  # user-plugin workloads name the oracle's native twin of the model
        p.model = MODEL_IDS[model] if isinstance(model, str) else int(model)
        p.n, p.m, p.horizon, p.dt = n, m, self.N, float(spec["dt"])
        integ = spec.get("integrator", "rk4")
        p.integrator = INTEGRATORS[integ] if isinstance(integ, str) else int(integ)
        p.has_control_box = 1 if spec.get("lb") is not None else 0
        params = list(spec.get("params", []))
        for i in range(16):
            p.model_params[i] = float(params[i]) if i < len(params) else 0.0
        dp = C.POINTER(C.c_double)
        for key in ("lti_A", "lti_B", "Q", "R", "Qf", "lb", "ub"):
            v = spec.get(key)
            if v is None:
                setattr(p, key, None)
            else:
                arr = _f64(v)
                self._keep[key] = arr
                setattr(p, key, arr.ctypes.data_as(dp))
        self.struct = p

    @property
    def ref(self):
        return C.byref(self.struct)


def build_alphas(opts: Options) -> np.ndarray:
    buf = np.zeros(64)
    cnt = load().oracle_build_alphas(C.byref(opts), _p(buf))
    return buf[:cnt].copy()


def continuous_dynamics(P: OracleProblem, x, u, t=0.0):
    x, u = _f64(x), _f64(u)
    out = np.zeros(P.n)
    load().oracle_continuous_dynamics(P.ref, _p(x), _p(u), C.c_double(t), _p(out))
    return out


def discrete_dynamics(P: OracleProblem, x, u, t=0.0):
    x, u = _f64(x), _f64(u)
    out = np.zeros(P.n)
    load().oracle_discrete_dynamics(P.ref, _p(x), _p(u), C.c_double(t), _p(out))
    return out


def jacobians(P: OracleProblem, x, u, t=0.0):
    x, u = _f64(x), _f64(u)
    Fx, Fu = np.zeros((P.n, P.n)), np.zeros((P.n, P.m))
    load().oracle_jacobians(P.ref, _p(x), _p(u), C.c_double(t), _p(Fx), _p(Fu))
    return Fx, Fu


def running_cost(P, x, u, ref):
    x, u, ref = _f64(x), _f64(u), _f64(ref)
    return load().oracle_running_cost(P.ref, _p(x), _p(u), _p(ref))


def terminal_cost(P, x, ref):
    x, ref = _f64(x), _f64(ref)
    return load().oracle_terminal_cost(P.ref, _p(x), _p(ref))


def trajectory_cost(P, X, U, xref, ref_traj=None):
    X, U, xref = _f64(X), _f64(U), _f64(xref)
    rt = _f64(ref_traj) if ref_traj is not None else None
    return load().oracle_trajectory_cost(P.ref, _p(X), _p(U), _p(xref), _p(rt))


def boxqp(opts: Options, H, g, lower, upper, x0=None, rhs=None):
    H, g, lower, upper = _f64(H), _f64(g), _f64(lower), _f64(upper)
    n = g.shape[0]
    x0a = _f64(x0) if x0 is not None else None
    x = np.zeros(n)
    free = np.zeros(n, dtype=np.int32)
    it, fac = C.c_int(0), C.c_int(0)
    fv, gn = C.c_double(0), C.c_double(0)
    ncols = 0
    rhs_a = out = None
    if rhs is not None:
        rhs_a = _f64(rhs)
        ncols = rhs_a.shape[1]
        out = np.zeros((n, ncols))
    st = load().oracle_boxqp(C.byref(opts), n, _p(H), _p(g), _p(lower), _p(upper), _p(x0a), _p(x), _p(free),
                             C.byref(it), C.byref(fac), C.byref(fv), C.byref(gn), ncols, _p(rhs_a), _p(out))
    return dict(status=st, x=x, free=free, iterations=it.value, factorizations=fac.value, value=fv.value,
                grad_norm=gn.value, solve=out)


def linearize(P, X, U):
    X, U = _f64(X), _f64(U)
    A, B = np.zeros((P.N, P.n, P.n)), np.zeros((P.N, P.n, P.m))
    load().oracle_linearize(P.ref, _p(X), _p(U), _p(A), _p(B))
    return A, B


def backward_pass(P, opts, X, U, xref, reg, k_prev=None, ref_traj=None, A=None, B=None, debug=False):
    X, U, xref = _f64(X), _f64(U), _f64(xref)
    rt = _f64(ref_traj) if ref_traj is not None else None
    K = np.zeros((P.N, P.m, P.n))
    k = np.zeros((P.N, P.m)) if k_prev is None else _f64(k_prev).copy()
    dV = np.zeros(2)
    inf_du = C.c_double(0)
    fail_t = C.c_int(-1)
    Vx = np.zeros((P.N + 1, P.n)) if debug else None
    Vxx = np.zeros((P.N + 1, P.n, P.n)) if debug else None
    if A is not None:
        A, B = _f64(A), _f64(B)
        ok = load().oracle_backward_pass_AB(P.ref, C.byref(opts), _p(A), _p(B), _p(X), _p(U), _p(xref), _p(rt),
                                            C.c_double(reg), _p(K), _p(k), _p(dV), C.byref(inf_du), _p(Vx), _p(Vxx),
                                            C.byref(fail_t))
    else:
        ok = load().oracle_backward_pass(P.ref, C.byref(opts), _p(X), _p(U), _p(xref), _p(rt), C.c_double(reg), _p(K),
                                         _p(k), _p(dV), C.byref(inf_du), _p(Vx), _p(Vxx), C.byref(fail_t))
    return dict(ok=bool(ok), K=K, k=k, dV=dV, inf_du=inf_du.value, fail_t=fail_t.value, Vx=Vx, Vxx=Vxx)


def forward_pass(P, opts, x0, X, U, xref, K, k, dV, cost, alpha, ref_traj=None):
    x0, X, U, xref, K, k, dV = map(_f64, (x0, X, U, xref, K, k, dV))
    rt = _f64(ref_traj) if ref_traj is not None else None
    Xn, Un = np.zeros_like(X), np.zeros_like(U)
    Jn = C.c_double(0)
    ok = load().oracle_forward_pass(P.ref, C.byref(opts), _p(x0), _p(X), _p(U), _p(xref), _p(rt), _p(K), _p(k), _p(dV),
                                    C.c_double(cost), C.c_double(alpha), _p(Xn), _p(Un), C.byref(Jn))
    return dict(success=bool(ok), X=Xn, U=Un, cost=Jn.value)


def solve(P, opts, x0, xref, X0, U0, ref_traj=None, history=False):
    x0, xref = _f64(x0), _f64(xref)
    X, U = _f64(X0).copy(), _f64(U0).copy()
    rt = _f64(ref_traj) if ref_traj is not None else None
    K, k = np.zeros((P.N, P.m, P.n)), np.zeros((P.N, P.m))
    res = Result()
    hist = np.zeros((opts.max_iterations + 1, 4)) if history else None
    load().oracle_solve(P.ref, C.byref(opts), _p(x0), _p(xref), _p(rt), _p(X), _p(U), _p(K), _p(k), C.byref(res), _p(hist))
    out = dict(X=X, U=U, K=K, k=k, cost=res.final_objective, alpha=res.final_step_length,
               reg=res.final_regularization, inf_du=res.inf_du, iterations=res.iterations, status=res.status)
    if history:
        out["history"] = hist[: res.history_len].copy()
    return out


def solve_batch(P, opts, x0, xref, X0, U0, ref_traj=None, nthreads=1):
    x0, xref = _f64(x0), _f64(xref)
    B = x0.shape[0]
    X, U = _f64(X0).copy(), _f64(U0).copy()
    rt = _f64(ref_traj) if ref_traj is not None else None
    K, k = np.zeros((B, P.N, P.m, P.n)), np.zeros((B, P.N, P.m))
    res = (Result * B)()
    load().oracle_solve_batch(P.ref, C.byref(opts), B, int(nthreads), _p(x0), _p(xref), _p(rt), _p(X), _p(U), _p(K),
                              _p(k), res)
    return dict(X=X, U=U, K=K, k=k, cost=np.array([r.final_objective for r in res]),
                alpha=np.array([r.final_step_length for r in res]), reg=np.array([r.final_regularization for r in res]),
                inf_du=np.array([r.inf_du for r in res]), iterations=np.array([r.iterations for r in res], dtype=np.int32),
                status=np.array([r.status for r in res], dtype=np.int32))


class ReplayReport(C.Structure):
    _fields_ = [("n_disagree", C.c_int), ("n_backward_disagree", C.c_int), ("infeasible", C.c_int), ("reserved", C.c_int),
                ("max_margin", C.c_double)]


TRACE_LS_FAILED, TRACE_EARLY_EXIT, TRACE_BW_LIMIT = 0, 0xFF, 0xFE


def solve_batch_traced(P, opts, x0, xref, X0, U0, ref_traj=None, nthreads=1, replay=None, history=True):
    """oracle_solve_batch_traced.  replay=None: normal solve, the decision trace comes back as out["trace"]
    ([B][max_iterations] int32).  replay=dict(trace=..., iterations=..., status=...): the oracle follows that decision
    sequence (e.g. the CUDA path's); out["replay"] holds the per-instance disagreement report."""
    x0, xref = _f64(x0), _f64(xref)
    B = x0.shape[0]
    X, U = _f64(X0).copy(), _f64(U0).copy()
    rt = _f64(ref_traj) if ref_traj is not None else None
    K, k = np.zeros((B, P.N, P.m, P.n)), np.zeros((B, P.N, P.m))
    res = (Result * B)()
    mi = max(int(opts.max_iterations), 1)
    hist = np.zeros((B, opts.max_iterations + 1, 4)) if history else None
    trace = np.zeros((B, mi), dtype=np.int32)
    rep = (ReplayReport * B)()
    if replay is None:
        load().oracle_solve_batch_traced(P.ref, C.byref(opts), B, int(nthreads), _p(x0), _p(xref), _p(rt), _p(X), _p(U), _p(K),
                                         _p(k), res, _p(hist), _p(trace), None, None, None, None)
    else:
        rtrace = np.ascontiguousarray(replay["trace"], dtype=np.int32)
        assert rtrace.shape == (B, mi), (rtrace.shape, (B, mi))
        rit = np.ascontiguousarray(replay["iterations"], dtype=np.int32)
        rst = np.ascontiguousarray(replay["status"], dtype=np.int32)
        load().oracle_solve_batch_traced(P.ref, C.byref(opts), B, int(nthreads), _p(x0), _p(xref), _p(rt), _p(X), _p(U), _p(K),
                                         _p(k), res, _p(hist), None, _p(rtrace), _p(rit), _p(rst), rep)
        trace = rtrace
    out = dict(X=X, U=U, K=K, k=k, cost=np.array([r.final_objective for r in res]),
               alpha=np.array([r.final_step_length for r in res]), reg=np.array([r.final_regularization for r in res]),
               inf_du=np.array([r.inf_du for r in res]), iterations=np.array([r.iterations for r in res], dtype=np.int32),
               status=np.array([r.status for r in res], dtype=np.int32), trace=trace,
               history_len=np.array([r.history_len for r in res], dtype=np.int32))
    if history:
        out["history"] = hist
    if replay is not None:
        out["replay"] = dict(n_disagree=np.array([r.n_disagree for r in rep]),
                             n_backward_disagree=np.array([r.n_backward_disagree for r in rep]),
                             infeasible=np.array([r.infeasible for r in rep]),
                             max_margin=np.array([r.max_margin for r in rep]))
    return out


def iterate_batch(P, opts, x0, xref, X, U, k, reg, cost, alpha, inf_du, ref_traj=None, nthreads=1, follow=None,
                  follow_status=None):
    """oracle_iterate_batch: one main-loop entry per instance from the given solver state (arrays are copied)."""
    x0, xref = _f64(x0), _f64(xref)
    B = x0.shape[0]
    X, U, k = _f64(X).copy(), _f64(U).copy(), _f64(k).copy()
    reg, cost, alpha, inf_du = (_f64(a).copy() for a in (reg, cost, alpha, inf_du))
    rt = _f64(ref_traj) if ref_traj is not None else None
    K = np.zeros((B, P.N, P.m, P.n))
    dV = np.zeros((B, 2))
    code, status = np.zeros(B, dtype=np.int32), np.zeros(B, dtype=np.int32)
    rep = (ReplayReport * B)()
    fo = None if follow is None else np.ascontiguousarray(follow, dtype=np.int32)
    fs = None if follow_status is None else np.ascontiguousarray(follow_status, dtype=np.int32)
    load().oracle_iterate_batch(P.ref, C.byref(opts), B, int(nthreads), _p(x0), _p(xref), _p(rt), _p(X), _p(U), _p(K), _p(k),
                                _p(reg), _p(cost), _p(alpha), _p(inf_du), _p(dV), _p(fo), _p(fs), _p(code), _p(status), rep)
    return dict(X=X, U=U, K=K, k=k, reg=reg, cost=cost, alpha=alpha, inf_du=inf_du, dV=dV, code=code, status=status,
                n_disagree=np.array([r.n_disagree for r in rep]),
                n_backward_disagree=np.array([r.n_backward_disagree for r in rep]),
                infeasible=np.array([r.infeasible for r in rep]), max_margin=np.array([r.max_margin for r in rep]))


def hardware_threads() -> int:
    return int(load().oracle_hardware_threads())


# ------------------------------------------------------------------------------------------------
# IPDDP (src/cddp_core/ipddp_solver.cpp): cold start, path inequality constraints
# ------------------------------------------------------------------------------------------------
CONSTRAINT_TYPES = {"control_box": 0, "state_box": 1, "ball": 2, "linear": 3}
DEFAULT_CONSTRAINT_NAMES = {"control_box": "ControlConstraint", "state_box": "StateConstraint", "ball": "BallConstraint",
                            "linear": "LinearConstraint"}
IPDDP_HISTORY_COLS = 9


class Constraint(C.Structure):
    _fields_ = [("type", C.c_int), ("rows", C.c_int), ("scale", C.c_double), ("p0", C.POINTER(C.c_double)),
                ("p1", C.POINTER(C.c_double))]


class IpddpOptions(C.Structure):
    _fields_ = [(k, C.c_double) for k in (
        "dual_var_init_scale", "slack_var_init_scale", "barrier_tol_mult", "barrier_update_dual_weight",
        "mu_kappa_epsilon", "theta_0_floor", "mu_initial", "mu_min_value", "mu_update_factor", "mu_update_power",
        "min_fraction_to_boundary", "merit_acceptance_threshold", "violation_acceptance_threshold",
        "max_violation_threshold", "min_violation_for_armijo_check", "jacobian_regularization_value",
        "jacobian_regularization_exponent")] + [
        ("theta_norm_l2", C.c_int), ("max_filter_size", C.c_int), ("barrier_strategy", C.c_int), ("terminal_equality", C.c_int)]


class IpddpResult(C.Structure):
    _fields_ = [(k, C.c_double) for k in ("final_objective", "final_step_length", "final_regularization", "inf_du",
                                          "inf_pr", "inf_comp", "mu", "merit", "decision_margin")] + [
        ("iterations", C.c_int), ("status", C.c_int), ("history_len", C.c_int), ("dual_dim", C.c_int)]


def sort_constraints(constraints):
    """The reference keeps path constraints in a std::map keyed by name and iterates it in key order
    (cddp_core.hpp:420-423, ipddp_solver.cpp:1375-1390): sort the same way."""
    return sorted(constraints, key=lambda c: c.get("name", DEFAULT_CONSTRAINT_NAMES[c["type"]]))


class ConstraintSet:
    """ctypes array of oracle_constraint / cddp_b200_constraint (identical layouts) built from a list of dicts:
    {'type': 'control_box'|'state_box', 'lb': [...], 'ub': [...], 'scale': 1.0}
    {'type': 'ball', 'center': [...], 'radius': r, 'scale': 1.0}   {'type': 'linear', 'A': [[...]], 'b': [...]}"""

    def __init__(self, constraints, struct=Constraint):
        cs = sort_constraints(list(constraints or []))
        self.constraints = cs
        self.nc = len(cs)
        self._keep = []
        arr = (struct * max(self.nc, 1))()
        dp = C.POINTER(C.c_double)
        for i, c in enumerate(cs):
            t = c["type"]
            arr[i].type = CONSTRAINT_TYPES[t]
            arr[i].scale = float(c.get("scale", 1.0))
            if t in ("control_box", "state_box"):
                a0, a1 = _f64(c["lb"]), _f64(c["ub"])
                arr[i].rows = a0.shape[0]
            elif t == "ball":
                a0, a1 = _f64(c["center"]), _f64([c["radius"]])
                arr[i].rows = a0.shape[0]
            else:
                a0, a1 = _f64(c["A"]), _f64(c["b"])
                arr[i].rows = a1.shape[0]
            self._keep += [a0, a1]
            arr[i].p0 = a0.ctypes.data_as(dp)
            arr[i].p1 = a1.ctypes.data_as(dp)
        self.array = arr

    def dual_dim(self, n, m):
        d = 0
        for c in self.constraints:
            d += {"control_box": 2 * m, "state_box": 2 * n, "ball": 1}.get(c["type"], len(np.atleast_1d(c.get("b", []))))
        return d


def make_ipddp_options(**overrides) -> IpddpOptions:
    io = IpddpOptions()
    load().oracle_ipddp_default_options(C.byref(io))
    for k, v in overrides.items():
        if not hasattr(io, k):
            raise AttributeError(k)
        setattr(io, k, v)
    return io


def eval_constraints(P, cset: ConstraintSet, x, u):
    x, u = _f64(x), _f64(u)
    d = cset.dual_dim(P.n, P.m)
    g, Gx, Gu = np.zeros(d), np.zeros((d, P.n)), np.zeros((d, P.m))
    load().oracle_eval_constraints(P.ref, cset.array, cset.nc, _p(x), _p(u), _p(g), _p(Gx), _p(Gu))
    return g, Gx, Gu


def ipddp_solve(P, opts, iopts, cset: ConstraintSet, x0, xref, U0, ref_traj=None, history=False):
    x0, xref = _f64(x0), _f64(xref)
    d = cset.dual_dim(P.n, P.m)
    X, U = np.zeros((P.N + 1, P.n)), _f64(U0).copy()
    rt = _f64(ref_traj) if ref_traj is not None else None
    K = np.zeros((P.N, P.m, P.n))
    Y, S = np.zeros((P.N, max(d, 1))), np.zeros((P.N, max(d, 1)))
    res = IpddpResult()
    hist = np.zeros((opts.max_iterations + 1, IPDDP_HISTORY_COLS)) if history else None
    load().oracle_ipddp_solve(P.ref, C.byref(opts), C.byref(iopts), cset.array, cset.nc, _p(x0), _p(xref), _p(rt), _p(X),
                              _p(U), _p(K), _p(Y), _p(S), C.byref(res), _p(hist))
    out = dict(X=X, U=U, K=K, Y=Y[:, :d], S=S[:, :d], cost=res.final_objective, alpha=res.final_step_length,
               reg=res.final_regularization, inf_du=res.inf_du, inf_pr=res.inf_pr, inf_comp=res.inf_comp, mu=res.mu,
               merit=res.merit, decision_margin=res.decision_margin, iterations=res.iterations, status=res.status)
    if history:
        out["history"] = hist[: res.history_len].copy()
    return out


def ipddp_solve_batch(P, opts, iopts, cset: ConstraintSet, x0, xref, U0, ref_traj=None, nthreads=1):
    x0, xref = _f64(x0), _f64(xref)
    B = x0.shape[0]
    d = cset.dual_dim(P.n, P.m)
    X, U = np.zeros((B, P.N + 1, P.n)), _f64(U0).copy()
    rt = _f64(ref_traj) if ref_traj is not None else None
    K = np.zeros((B, P.N, P.m, P.n))
    Y, S = np.zeros((B, P.N, d)), np.zeros((B, P.N, d))
    res = (IpddpResult * B)()
    load().oracle_ipddp_solve_batch(P.ref, C.byref(opts), C.byref(iopts), cset.array, cset.nc, B, int(nthreads), _p(x0),
                                    _p(xref), _p(rt), _p(X), _p(U), _p(K), _p(Y) if d else None, _p(S) if d else None, res)
    f = lambda k, dt=np.float64: np.array([getattr(r, k) for r in res], dtype=dt)  # noqa: E731
    return dict(X=X, U=U, K=K, Y=Y, S=S, cost=f("final_objective"), alpha=f("final_step_length"),
                reg=f("final_regularization"), inf_du=f("inf_du"), inf_pr=f("inf_pr"), inf_comp=f("inf_comp"), mu=f("mu"),
                merit=f("merit"), decision_margin=f("decision_margin"), iterations=f("iterations", np.int32),
                status=f("status", np.int32))


MAX_ALPHAS = 64
IP_STATE_SCALARS = ("mu", "cost", "merit", "filter_theta", "inf_pr", "inf_comp", "reg", "alpha_pr", "alpha_du", "step_norm",
                    "inf_du", "iter")


def ipddp_iterate_batch(P, opts, iopts, cset: ConstraintSet, x0, xref, state, ref_traj=None, nthreads=1, follow=None,
                        follow_status=None):
    """oracle_ipddp_iterate_batch: ONE IPDDP main-loop entry per instance from the solver state `state` =
    dict(X, U, Y, S, G, lamT, filter [B][8][2], filter_size [B], + the IP_STATE_SCALARS as [B] arrays; "iter" = the 1-based
    index of the iteration about to run).  Arrays are copied; the state after the iteration comes back in the same form
    plus code / status / the disagreement report."""
    x0, xref = _f64(x0), _f64(xref)
    B = x0.shape[0]
    d = cset.dual_dim(P.n, P.m)
    X, U = _f64(state["X"]).copy(), _f64(state["U"]).copy()
    Y, S, G = (_f64(state[k]).copy().reshape(B, P.N, d) if d else np.zeros((B, P.N, 1)) for k in ("Y", "S", "G"))
    lamT = _f64(state["lamT"]).copy() if state.get("lamT") is not None else np.zeros((B, P.n))
    filt = _f64(state["filter"]).copy()
    fsz = np.ascontiguousarray(state["filter_size"], dtype=np.int32).copy()
    sc = np.ascontiguousarray(np.stack([np.asarray(state[k], dtype=np.float64) for k in IP_STATE_SCALARS], axis=1))
    rt = _f64(ref_traj) if ref_traj is not None else None
    code, status = np.zeros(B, dtype=np.int32), np.zeros(B, dtype=np.int32)
    rep = (ReplayReport * B)()
    fo = None if follow is None else np.ascontiguousarray(follow, dtype=np.int32)
    fs = None if follow_status is None else np.ascontiguousarray(follow_status, dtype=np.int32)
    table = np.full((B, MAX_ALPHAS, 6), np.nan)
    load().oracle_ipddp_iterate_batch(P.ref, C.byref(opts), C.byref(iopts), cset.array, cset.nc, B, int(nthreads), _p(x0), _p(xref),
                                      _p(rt), _p(X), _p(U), _p(Y), _p(S), _p(G), _p(lamT), _p(filt), _p(fsz), _p(sc), _p(fo), _p(fs),
                                      _p(code), _p(status), rep, _p(table))
    out = dict(X=X, U=U, Y=Y[:, :, :d], S=S[:, :, :d], G=G[:, :, :d], lamT=lamT, filter=filt, filter_size=fsz, code=code, status=status,
               n_disagree=np.array([r.n_disagree for r in rep]), n_backward_disagree=np.array([r.n_backward_disagree for r in rep]),
               infeasible=np.array([r.infeasible for r in rep]), max_margin=np.array([r.max_margin for r in rep]),
               kind=np.array([r.reserved for r in rep]), trials=table)
    for i, k in enumerate(IP_STATE_SCALARS):
        out[k] = sc[:, i].copy()
    return out


def ipddp_probe(P, opts, iopts, cset: ConstraintSet, x0, xref, U0, iters, ref_traj=None):
    """initialize + `iters` full iterations + ONE backward pass; returns every intermediate (white box)."""
    x0, xref, U0 = _f64(x0), _f64(xref), _f64(U0)
    rt = _f64(ref_traj) if ref_traj is not None else None
    N, n, m, d = P.N, P.n, P.m, cset.dual_dim(P.n, P.m)
    dd = max(d, 1)
    z = lambda *s: np.zeros(s)  # noqa: E731
    out = dict(X=z(N + 1, n), U=z(N, m), Y=z(N, dd), S=z(N, dd), G=z(N, dd), ku=z(N, m), Ku=z(N, m, n), ky=z(N, dd),
               Ky=z(N, dd, n), ks=z(N, dd), Ks=z(N, dd, n), dS=z(N, dd), dY=z(N, dd))
    scal = np.zeros(16)
    na = len(build_alphas(opts))
    trial = np.zeros((na, 4))
    load().oracle_ipddp_probe(P.ref, C.byref(opts), C.byref(iopts), cset.array, cset.nc, _p(x0), _p(xref), _p(rt), _p(U0),
                              int(iters), *[_p(out[k]) for k in ("X", "U", "Y", "S", "G", "ku", "Ku", "ky", "Ky", "ks", "Ks",
                                                                 "dS", "dY")], _p(scal), _p(trial))
    names = ("mu", "cost", "merit", "inf_pr", "inf_du", "inf_comp", "step_norm", "reg", "dV0", "dV1", "alpha_pr_max",
             "alpha_du_max", "filter_theta", "theta", "filter_size", "bw_ok")
    out.update({k: scal[i] for i, k in enumerate(names)})
    out["trial"] = trial
    for k in ("Y", "S", "G", "ky", "Ky", "ks", "Ks", "dS", "dY"):
        out[k] = out[k][:, :d]
    return out
