"""cddp-cpp_b200 — B200-native batched CLDDP/iLQR engine (Python harness side).

The product is the C-ABI shared library ``libcddp_b200.so`` (``include/cddp_b200.h``) built from the
hand-written sm_100a CUDA in ``csrc/`` plus the C++ host mirror of the reference API in ``host/``.
This module is only the ctypes binding the tests and ``bench.py`` use to reach that ABI; it holds no
solver arithmetic and has NO fallback: if the shared library is missing, importing the binding
raises.

The directory name contains a hyphen, so import it with::

    import importlib; cddp = importlib.import_module("cddp-cpp_b200")
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libcddp_b200.so")

MAX_N, MAX_M, MAX_ALPHAS = 16, 8, 32

MODEL_PENDULUM, MODEL_CARTPOLE, MODEL_UNICYCLE, MODEL_QUADROTOR, MODEL_LTI = range(5)
MODEL_USER = 5
MODEL_IDS = {"pendulum": 0, "cartpole": 1, "unicycle": 2, "quadrotor": 3, "lti": 4, "user": 5}
EULER, HEUN, RK3, RK4 = range(4)
INTEGRATORS = {"euler": 0, "heun": 1, "rk3": 2, "rk4": 3}

STATUS_RUNNING, STATUS_OPTIMAL, STATUS_ACCEPTABLE, STATUS_MAX_ITERATIONS, STATUS_REG_LIMIT, STATUS_MAX_CPU_TIME = range(6)


class Options(C.Structure):
    """cddp_b200_options — the CLDDP-relevant subset of cddp::CDDPOptions (options.hpp:208-251)."""

    _fields_ = [
        ("tolerance", C.c_double),
        ("acceptable_tolerance", C.c_double),
        ("max_iterations", C.c_int),
        ("enable_parallel", C.c_int),
        ("max_cpu_time", C.c_double),
        ("termination_scaling_max_factor", C.c_double),
        ("ls_max_iterations", C.c_int),
        ("reserved1", C.c_int),
        ("ls_initial_step_size", C.c_double),
        ("ls_min_step_size", C.c_double),
        ("ls_step_reduction_factor", C.c_double),
        ("reg_initial_value", C.c_double),
        ("reg_update_factor", C.c_double),
        ("reg_max_value", C.c_double),
        ("reg_min_value", C.c_double),
        ("qp_max_iterations", C.c_int),
        ("reserved2", C.c_int),
        ("qp_min_gradient_norm", C.c_double),
        ("qp_min_relative_improvement", C.c_double),
        ("qp_step_decrease_factor", C.c_double),
        ("qp_min_step_size", C.c_double),
        ("qp_armijo_constant", C.c_double),
        ("armijo_constant", C.c_double),
    ]


class Problem(C.Structure):
    """cddp_b200_problem — batch-shared problem definition."""

    _fields_ = [
        ("model", C.c_int),
        ("n", C.c_int),
        ("m", C.c_int),
        ("horizon", C.c_int),
        ("dt", C.c_double),
        ("integrator", C.c_int),
        ("has_control_box", C.c_int),
        ("model_params", C.c_double * 16),
        ("lti_A", C.POINTER(C.c_double)),
        ("lti_B", C.POINTER(C.c_double)),
        ("Q", C.POINTER(C.c_double)),
        ("R", C.POINTER(C.c_double)),
        ("Qf", C.POINTER(C.c_double)),
        ("lb", C.POINTER(C.c_double)),
        ("ub", C.POINTER(C.c_double)),
    ]


class Constraint(C.Structure):
    """cddp_b200_constraint — one entry of the IPDDP path-constraint set."""

    _fields_ = [("type", C.c_int), ("rows", C.c_int), ("scale", C.c_double), ("p0", C.POINTER(C.c_double)),
                ("p1", C.POINTER(C.c_double))]


class IpddpOptions(C.Structure):
    """cddp_b200_ipddp_options — CDDPOptions::ipddp + ::filter (options.hpp:75-104, :148-186)."""

    _fields_ = [(k, C.c_double) for k in (
        "dual_var_init_scale", "slack_var_init_scale", "barrier_tol_mult", "barrier_update_dual_weight",
        "mu_kappa_epsilon", "theta_0_floor", "mu_initial", "mu_min_value", "mu_update_factor", "mu_update_power",
        "min_fraction_to_boundary", "merit_acceptance_threshold", "violation_acceptance_threshold",
        "max_violation_threshold", "min_violation_for_armijo_check", "jacobian_regularization_value",
        "jacobian_regularization_exponent")] + [
        ("theta_norm_l2", C.c_int), ("max_filter_size", C.c_int), ("barrier_strategy", C.c_int), ("terminal_equality", C.c_int)]


CONSTRAINT_TYPES = {"control_box": 0, "state_box": 1, "ball": 2, "linear": 3}
DEFAULT_CONSTRAINT_NAMES = {"control_box": "ControlConstraint", "state_box": "StateConstraint", "ball": "BallConstraint",
                            "linear": "LinearConstraint"}
IPDDP_HISTORY_COLS = 9


class Timing(C.Structure):
    _fields_ = [
        ("linearize_ms", C.c_double),
        ("backward_ms", C.c_double),
        ("forward_ms", C.c_double),
        ("linearize_launches", C.c_longlong),
        ("backward_launches", C.c_longlong),
        ("forward_launches", C.c_longlong),
        ("other_launches", C.c_longlong),
    ]


# every symbol include/cddp_b200.h declares (tests/test_abi.py checks the library exports all of them)
ABI_SYMBOLS = [
    "cddp_b200_abi_version", "cddp_b200_error_string", "cddp_b200_last_cuda_error", "cddp_b200_status_string",
    "cddp_b200_device_count", "cddp_b200_default_options", "cddp_b200_build_alphas", "cddp_b200_create",
    "cddp_b200_destroy", "cddp_b200_set_stream", "cddp_b200_set_options", "cddp_b200_set_instances",
    "cddp_b200_set_instances_device", "cddp_b200_initialize", "cddp_b200_linearize", "cddp_b200_backward_pass",
    "cddp_b200_forward_pass", "cddp_b200_iterate", "cddp_b200_solve", "cddp_b200_num_running",
    "cddp_b200_synchronize", "cddp_b200_get_solution", "cddp_b200_enable_history", "cddp_b200_get_history",
    "cddp_b200_get_feedforward", "cddp_b200_set_gains", "cddp_b200_set_regularization", "cddp_b200_set_cost",
    "cddp_b200_get_linearization", "cddp_b200_set_linearization", "cddp_b200_get_sweep", "cddp_b200_get_forward",
    "cddp_b200_reset_timing", "cddp_b200_get_timing", "cddp_b200_enable_timing",
    "cddp_b200_backward_algorithmic_bytes", "cddp_b200_solve_host", "cddp_b200_set_record_layout",
    "cddp_b200_get_record_layout", "cddp_b200_set_poll_interval", "cddp_b200_set_line_search_window", "cddp_b200_set_first_alpha_speculation", "cddp_b200_set_fused_linearization", "cddp_b200_get_solution_async",
    "cddp_b200_mpc_advance", "cddp_b200_get_first_controls_async",
    "cddp_b200_ipddp_default_options", "cddp_b200_ipddp_create", "cddp_b200_ipddp_dual_dim",
    "cddp_b200_ipddp_get_solution", "cddp_b200_ipddp_get_iteration_state", "cddp_b200_ipddp_get_gains", "cddp_b200_ipddp_get_line_search",
    "cddp_b200_ipddp_get_history", "cddp_b200_create_ex", "cddp_b200_ipddp_create_ex", "cddp_b200_compile_user_model",
    "cddp_b200_last_compile_log", "cddp_b200_enable_trace", "cddp_b200_get_trace",
]


class CddpB200Error(RuntimeError):
    def __init__(self, code: int, what: str):
        super().__init__(what)
        self.code = code


_lib = None


def load_library() -> C.CDLL:
    """dlopen libcddp_b200.so.  Raises (never falls back) if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise CddpB200Error(-1, f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                                "(there is no CPU fallback)")
    lib = C.CDLL(LIB_PATH)
    dp = C.POINTER(C.c_double)
    ip = C.POINTER(C.c_int)
    vp = C.c_void_p
    lib.cddp_b200_abi_version.restype = C.c_int
    lib.cddp_b200_error_string.restype = C.c_char_p
    lib.cddp_b200_error_string.argtypes = [C.c_int]
    lib.cddp_b200_last_cuda_error.restype = C.c_char_p
    lib.cddp_b200_status_string.restype = C.c_char_p
    lib.cddp_b200_status_string.argtypes = [C.c_int]
    lib.cddp_b200_device_count.argtypes = [ip]
    lib.cddp_b200_default_options.restype = None
    lib.cddp_b200_default_options.argtypes = [C.POINTER(Options)]
    lib.cddp_b200_build_alphas.argtypes = [C.POINTER(Options), dp, C.c_int, ip]
    lib.cddp_b200_create.argtypes = [C.POINTER(Problem), C.POINTER(Options), C.c_int, C.c_int, C.POINTER(vp)]
    lib.cddp_b200_destroy.argtypes = [vp]
    lib.cddp_b200_set_stream.argtypes = [vp, vp]
    lib.cddp_b200_set_options.argtypes = [vp, C.POINTER(Options)]
    lib.cddp_b200_set_record_layout.argtypes = [vp, C.c_int]
    lib.cddp_b200_get_record_layout.argtypes = [vp, ip, ip]
    lib.cddp_b200_set_instances.argtypes = [vp, vp, vp, vp, vp, vp]
    lib.cddp_b200_set_instances_device.argtypes = [vp, vp, vp, vp, vp, vp]
    for name in ("initialize", "linearize", "backward_pass", "forward_pass", "solve", "synchronize", "reset_timing"):
        getattr(lib, "cddp_b200_" + name).argtypes = [vp]
    lib.cddp_b200_iterate.argtypes = [vp, C.c_int]
    lib.cddp_b200_num_running.argtypes = [vp, ip]
    lib.cddp_b200_get_solution.argtypes = [vp] + [vp] * 9
    lib.cddp_b200_get_solution_async.argtypes = [vp] + [vp] * 9
    lib.cddp_b200_set_poll_interval.argtypes = [vp, C.c_int]
    lib.cddp_b200_set_line_search_window.argtypes = [vp, C.c_int]
    lib.cddp_b200_set_first_alpha_speculation.argtypes = [vp, C.c_int]
    lib.cddp_b200_set_fused_linearization.argtypes = [vp, C.c_int]
    lib.cddp_b200_mpc_advance.argtypes = [vp, C.c_int, vp, vp]
    lib.cddp_b200_get_first_controls_async.argtypes = [vp, vp, vp, vp]
    lib.cddp_b200_enable_history.argtypes = [vp, C.c_int]
    lib.cddp_b200_get_history.argtypes = [vp, vp, vp]
    lib.cddp_b200_get_feedforward.argtypes = [vp, vp]
    lib.cddp_b200_set_gains.argtypes = [vp, vp, vp]
    lib.cddp_b200_set_regularization.argtypes = [vp, vp]
    lib.cddp_b200_set_cost.argtypes = [vp, vp]
    lib.cddp_b200_get_linearization.argtypes = [vp, vp, vp]
    lib.cddp_b200_set_linearization.argtypes = [vp, vp, vp]
    lib.cddp_b200_get_sweep.argtypes = [vp, vp, vp, vp, vp, vp]
    lib.cddp_b200_get_forward.argtypes = [vp, vp, vp, vp, vp]
    lib.cddp_b200_get_timing.argtypes = [vp, C.POINTER(Timing)]
    lib.cddp_b200_enable_timing.argtypes = [vp, C.c_int]
    lib.cddp_b200_backward_algorithmic_bytes.argtypes = [vp, dp]
    lib.cddp_b200_solve_host.argtypes = [C.POINTER(Problem), C.POINTER(Options), C.c_int, C.c_int] + [vp] * 12
    lib.cddp_b200_ipddp_default_options.restype = None
    lib.cddp_b200_ipddp_default_options.argtypes = [C.POINTER(IpddpOptions)]
    lib.cddp_b200_ipddp_create.argtypes = [C.POINTER(Problem), C.POINTER(Options), C.POINTER(IpddpOptions), vp, C.c_int,
                                           C.c_int, C.c_int, C.POINTER(vp)]
    lib.cddp_b200_create_ex.argtypes = [C.POINTER(Problem), C.POINTER(Options), C.c_char_p, C.c_int, C.c_int, C.POINTER(vp)]
    lib.cddp_b200_ipddp_create_ex.argtypes = [C.POINTER(Problem), C.POINTER(Options), C.POINTER(IpddpOptions), vp, C.c_int,
                                              C.c_char_p, C.c_int, C.c_int, C.POINTER(vp)]
    lib.cddp_b200_compile_user_model.argtypes = [C.c_char_p, C.c_int, C.c_int, C.POINTER(C.c_size_t)]
    lib.cddp_b200_last_compile_log.restype = C.c_char_p
    lib.cddp_b200_ipddp_dual_dim.argtypes = [vp, ip]
    lib.cddp_b200_ipddp_get_solution.argtypes = [vp, vp, vp, vp, vp]
    lib.cddp_b200_ipddp_get_gains.argtypes = [vp, vp, vp, vp, vp]
    lib.cddp_b200_ipddp_get_iteration_state.argtypes = [vp, vp, vp, vp, vp]
    lib.cddp_b200_ipddp_get_line_search.argtypes = [vp, vp]
    lib.cddp_b200_ipddp_get_history.argtypes = [vp, vp, vp]
    for name in ABI_SYMBOLS:
        fn = getattr(lib, name)
        if name not in ("cddp_b200_error_string", "cddp_b200_last_cuda_error", "cddp_b200_status_string",
                        "cddp_b200_default_options", "cddp_b200_ipddp_default_options", "cddp_b200_last_compile_log"):
            fn.restype = C.c_int
    _lib = lib
    return lib


def _check(rc: int) -> None:
    if rc != 0:
        lib = load_library()
        msg = lib.cddp_b200_error_string(rc).decode()
        if rc in (3, 4):
            msg += " — " + lib.cddp_b200_last_cuda_error().decode()
        if rc == 6:
            msg += "\n" + lib.cddp_b200_last_compile_log().decode()
        raise CddpB200Error(rc, msg)


def default_options(**overrides) -> Options:
    """CDDPOptions() defaults (options.hpp) with keyword overrides."""
    o = Options()
    load_library().cddp_b200_default_options(C.byref(o))
    for k, v in overrides.items():
        if not hasattr(o, k):
            raise AttributeError(f"unknown option {k}")
        setattr(o, k, v)
    return o


def build_alphas(opts: Options) -> np.ndarray:
    buf = np.zeros(256)
    cnt = C.c_int(0)
    _check(load_library().cddp_b200_build_alphas(C.byref(opts), buf.ctypes.data_as(C.POINTER(C.c_double)), 256, C.byref(cnt)))
    return buf[: cnt.value].copy()


def _f64(a, shape=None) -> Optional[np.ndarray]:
    if a is None:
        return None
    arr = np.ascontiguousarray(np.asarray(a, dtype=np.float64))
    if shape is not None and tuple(arr.shape) != tuple(shape):
        raise ValueError(f"expected shape {tuple(shape)}, got {tuple(arr.shape)}")
    return arr


def _ptr(a) -> Optional[int]:
    return None if a is None else a.ctypes.data


class ProblemSpec:
    """Keeps the numpy arrays a cddp_b200_problem points at alive; builds the ctypes struct.

    ``spec`` is a dict: model, n, m, horizon, dt, integrator, params, Q, R, Qf, lb, ub, lti_A, lti_B.
    """

    def __init__(self, spec: dict, struct_cls=Problem):
        self.spec = spec
        n, m = int(spec["n"]), int(spec["m"])
        self.n, self.m, self.N = n, m, int(spec["horizon"])
        self.Q = _f64(spec["Q"], (n, n))
        self.R = _f64(spec["R"], (m, m))
        self.Qf = _f64(spec["Qf"], (n, n))
        self.lb = _f64(spec.get("lb"))
        self.ub = _f64(spec.get("ub"))
        self.lti_A = _f64(spec.get("lti_A"))
        self.lti_B = _f64(spec.get("lti_B"))
        p = struct_cls()
        model = spec["model"]
        p.model = MODEL_IDS[model] if isinstance(model, str) else int(model)
        p.n, p.m, p.horizon, p.dt = n, m, self.N, float(spec["dt"])
        integ = spec.get("integrator", "rk4")
        p.integrator = INTEGRATORS[integ] if isinstance(integ, str) else int(integ)
        p.has_control_box = 1 if self.lb is not None else 0
        params = list(spec.get("params", []))
        for i in range(16):
            p.model_params[i] = float(params[i]) if i < len(params) else 0.0
        dp = C.POINTER(C.c_double)
        cast = lambda a: a.ctypes.data_as(dp) if a is not None else None  # noqa: E731
        p.lti_A, p.lti_B = cast(self.lti_A), cast(self.lti_B)
        p.Q, p.R, p.Qf = cast(self.Q), cast(self.R), cast(self.Qf)
        p.lb, p.ub = cast(self.lb), cast(self.ub)
        self.struct = p


class BatchedCLDDP:
    """Owner of one cddp_b200_solver handle (one GPU, one batch of instances)."""

    def __init__(self, spec: dict, opts: Options, batch: int, device: int = 0):
        self.lib = load_library()
        self.pspec = ProblemSpec(spec)
        self.opts = opts
        self.B, self.n, self.m, self.N = int(batch), self.pspec.n, self.pspec.m, self.pspec.N
        self.handle = C.c_void_p()
        src = spec.get("model_source")  # CUDA source of a user-defined dynamics model (model == "user")
        _check(self.lib.cddp_b200_create_ex(C.byref(self.pspec.struct), C.byref(opts), src.encode() if src else None, self.B,
                                            device, C.byref(self.handle)))
        self.num_alphas = len(build_alphas(opts))
        self._keep = []

    def close(self):
        if getattr(self, "handle", None) is not None and self.handle:
            self.lib.cddp_b200_destroy(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    # ---- configuration ----
    def set_stream(self, stream_ptr: int):
        _check(self.lib.cddp_b200_set_stream(self.handle, C.c_void_p(stream_ptr)))

    def set_options(self, opts: Options):
        _check(self.lib.cddp_b200_set_options(self.handle, C.byref(opts)))
        self.opts = opts
        self.num_alphas = len(build_alphas(opts))

    def set_record_layout(self, layout):
        """'dense' (stacked n*n+n*m Jacobians) or 'structured' (the model's structural non-zeros only)."""
        code = {"dense": 0, "structured": 1}[layout] if isinstance(layout, str) else int(layout)
        _check(self.lib.cddp_b200_set_record_layout(self.handle, code))

    def get_record_layout(self):
        lay, nb = C.c_int(0), C.c_int(0)
        _check(self.lib.cddp_b200_get_record_layout(self.handle, C.byref(lay), C.byref(nb)))
        return ("dense", "structured")[lay.value], nb.value

    def mpc_advance(self, steps: int, x0_new=None, xref_new=None):
        a = _f64(x0_new, (self.B, self.n)) if x0_new is not None else None
        b = _f64(xref_new, (self.B, self.n)) if xref_new is not None else None
        self._keep_mpc = (a, b)
        _check(self.lib.cddp_b200_mpc_advance(self.handle, int(steps), _ptr(a), _ptr(b)))
        _check(self.lib.cddp_b200_synchronize(self.handle))

    def get_first_controls(self):
        u0, cost, st = np.empty((self.B, self.m)), np.empty(self.B), np.empty(self.B, dtype=np.int32)
        _check(self.lib.cddp_b200_get_first_controls_async(self.handle, _ptr(u0), _ptr(cost), _ptr(st)))
        _check(self.lib.cddp_b200_synchronize(self.handle))
        return u0, cost, st

    def set_poll_interval(self, interval: int):
        _check(self.lib.cddp_b200_set_poll_interval(self.handle, int(interval)))

    def set_line_search_window(self, enable: bool):
        _check(self.lib.cddp_b200_set_line_search_window(self.handle, int(enable)))

    def set_fused_linearization(self, enable: bool):
        _check(self.lib.cddp_b200_set_fused_linearization(self.handle, int(enable)))

    def set_first_alpha_speculation(self, mode: int):
        _check(self.lib.cddp_b200_set_first_alpha_speculation(self.handle, int(mode)))

    def set_instances(self, x0, xref, X0=None, U0=None, ref_traj=None):
        B, n, m, N = self.B, self.n, self.m, self.N
        a = [_f64(x0, (B, n)), _f64(xref, (B, n)), _f64(ref_traj, (B, N + 1, n)) if ref_traj is not None else None,
             _f64(X0, (B, N + 1, n)) if X0 is not None else None, _f64(U0, (B, N, m)) if U0 is not None else None]
        self._keep = a  # async H2D copies read these
        _check(self.lib.cddp_b200_set_instances(self.handle, *[_ptr(v) for v in a]))
        _check(self.lib.cddp_b200_synchronize(self.handle))

    def set_instances_device(self, x0_ptr, xref_ptr, X0_ptr, U0_ptr, ref_traj_ptr=None):
        _check(self.lib.cddp_b200_set_instances_device(self.handle, x0_ptr, xref_ptr, ref_traj_ptr, X0_ptr, U0_ptr))

    # ---- steps ----
    def initialize(self):
        _check(self.lib.cddp_b200_initialize(self.handle))

    def linearize(self):
        _check(self.lib.cddp_b200_linearize(self.handle))

    def backward_pass(self):
        _check(self.lib.cddp_b200_backward_pass(self.handle))

    def forward_pass(self):
        _check(self.lib.cddp_b200_forward_pass(self.handle))

    def iterate(self, iterations: int):
        _check(self.lib.cddp_b200_iterate(self.handle, int(iterations)))

    def solve(self):
        _check(self.lib.cddp_b200_solve(self.handle))

    def synchronize(self):
        _check(self.lib.cddp_b200_synchronize(self.handle))

    def num_running(self) -> int:
        r = C.c_int(0)
        _check(self.lib.cddp_b200_num_running(self.handle, C.byref(r)))
        return r.value

    # ---- results ----
    def get_solution(self, want_K: bool = True) -> dict:
        B, n, m, N = self.B, self.n, self.m, self.N
        out = {
            "X": np.empty((B, N + 1, n)), "U": np.empty((B, N, m)),
            "K": np.empty((B, N, m, n)) if want_K else None,
            "cost": np.empty(B), "iterations": np.empty(B, dtype=np.int32), "status": np.empty(B, dtype=np.int32),
            "alpha": np.empty(B), "reg": np.empty(B), "inf_du": np.empty(B),
        }
        _check(self.lib.cddp_b200_get_solution(self.handle, _ptr(out["X"]), _ptr(out["U"]), _ptr(out["K"]),
                                               _ptr(out["cost"]), _ptr(out["iterations"]), _ptr(out["status"]),
                                               _ptr(out["alpha"]), _ptr(out["reg"]), _ptr(out["inf_du"])))
        return out

    def get_scalars(self) -> dict:
        B = self.B
        out = {"cost": np.empty(B), "iterations": np.empty(B, dtype=np.int32), "status": np.empty(B, dtype=np.int32),
               "alpha": np.empty(B), "reg": np.empty(B), "inf_du": np.empty(B)}
        _check(self.lib.cddp_b200_get_solution(self.handle, None, None, None, _ptr(out["cost"]), _ptr(out["iterations"]),
                                               _ptr(out["status"]), _ptr(out["alpha"]), _ptr(out["reg"]),
                                               _ptr(out["inf_du"])))
        return out

    def enable_history(self, enable: bool = True):
        _check(self.lib.cddp_b200_enable_history(self.handle, 1 if enable else 0))

    def get_history(self):
        cap = self.opts.max_iterations + 1
        h = np.empty((self.B, cap, 4))
        lens = np.empty(self.B, dtype=np.int32)
        _check(self.lib.cddp_b200_get_history(self.handle, _ptr(h), _ptr(lens)))
        return h, lens

    # ---- white box ----
    def enable_trace(self, enable: bool = True):
        _check(self.lib.cddp_b200_enable_trace(self.handle, int(enable)))

    def get_trace(self) -> np.ndarray:
        """[B][max_iterations] decision codes (cddp_b200_get_trace)."""
        cap = C.c_int(0)
        tr = np.zeros((self.B, max(int(self.opts.max_iterations), 1)), dtype=np.int32)
        _check(self.lib.cddp_b200_get_trace(self.handle, C.c_void_p(tr.ctypes.data), C.byref(cap)))
        assert cap.value == tr.shape[1], (cap.value, tr.shape)
        return tr

    def get_feedforward(self) -> np.ndarray:
        k = np.empty((self.B, self.N, self.m))
        _check(self.lib.cddp_b200_get_feedforward(self.handle, _ptr(k)))
        return k

    def set_gains(self, K=None, k=None):
        K = _f64(K, (self.B, self.N, self.m, self.n)) if K is not None else None
        k = _f64(k, (self.B, self.N, self.m)) if k is not None else None
        _check(self.lib.cddp_b200_set_gains(self.handle, _ptr(K), _ptr(k)))

    def set_regularization(self, reg):
        reg = _f64(np.broadcast_to(np.asarray(reg, dtype=np.float64), (self.B,)).copy(), (self.B,))
        _check(self.lib.cddp_b200_set_regularization(self.handle, _ptr(reg)))

    def set_cost(self, cost):
        cost = _f64(cost, (self.B,))
        _check(self.lib.cddp_b200_set_cost(self.handle, _ptr(cost)))

    def get_linearization(self):
        A = np.empty((self.B, self.N, self.n, self.n))
        Bm = np.empty((self.B, self.N, self.n, self.m))
        _check(self.lib.cddp_b200_get_linearization(self.handle, _ptr(A), _ptr(Bm)))
        return A, Bm

    def set_linearization(self, A, Bm):
        A = _f64(A, (self.B, self.N, self.n, self.n))
        Bm = _f64(Bm, (self.B, self.N, self.n, self.m))
        _check(self.lib.cddp_b200_set_linearization(self.handle, _ptr(A), _ptr(Bm)))

    def get_sweep(self) -> dict:
        B, n = self.B, self.n
        out = {"dV": np.empty((B, 2)), "ok": np.empty(B, dtype=np.int32), "inf_du": np.empty(B),
               "Vx0": np.empty((B, n)), "Vxx0": np.empty((B, n, n))}
        _check(self.lib.cddp_b200_get_sweep(self.handle, _ptr(out["dV"]), _ptr(out["ok"]), _ptr(out["inf_du"]),
                                            _ptr(out["Vx0"]), _ptr(out["Vxx0"])))
        return out

    def get_forward(self) -> dict:
        B, n, m, N = self.B, self.n, self.m, self.N
        out = {"costs": np.empty((B, self.num_alphas)), "accepted": np.empty(B, dtype=np.int32),
               "X": np.empty((B, N + 1, n)), "U": np.empty((B, N, m))}
        _check(self.lib.cddp_b200_get_forward(self.handle, _ptr(out["costs"]), _ptr(out["accepted"]), _ptr(out["X"]),
                                              _ptr(out["U"])))
        return out

    # ---- measurement ----
    def enable_timing(self, enable: bool = True):
        _check(self.lib.cddp_b200_enable_timing(self.handle, 1 if enable else 0))

    def reset_timing(self):
        _check(self.lib.cddp_b200_reset_timing(self.handle))

    def get_timing(self) -> Timing:
        t = Timing()
        _check(self.lib.cddp_b200_get_timing(self.handle, C.byref(t)))
        return t

    def backward_algorithmic_bytes(self) -> float:
        v = C.c_double(0.0)
        _check(self.lib.cddp_b200_backward_algorithmic_bytes(self.handle, C.byref(v)))
        return v.value


def default_ipddp_options(**overrides) -> IpddpOptions:
    io = IpddpOptions()
    load_library().cddp_b200_ipddp_default_options(C.byref(io))
    for k, v in overrides.items():
        if not hasattr(io, k):
            raise AttributeError(f"unknown IPDDP option {k}")
        setattr(io, k, v)
    return io


class ConstraintSet:
    """Array of cddp_b200_constraint from a list of dicts, sorted by constraint name like the reference's std::map
    (cddp_core.hpp:420-423):  {'type': 'control_box'|'state_box', 'lb', 'ub', 'scale'} | {'type': 'ball', 'center',
    'radius', 'scale'} | {'type': 'linear', 'A', 'b'}; optional 'name' overrides the class's default name."""

    def __init__(self, constraints):
        cs = sorted(list(constraints or []), key=lambda c: c.get("name", DEFAULT_CONSTRAINT_NAMES[c["type"]]))
        self.constraints, self.nc, self._keep = cs, len(cs), []
        arr = (Constraint * max(self.nc, 1))()
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
            arr[i].p0, arr[i].p1 = a0.ctypes.data_as(dp), a1.ctypes.data_as(dp)
        self.array = arr


class BatchedIPDDP(BatchedCLDDP):
    """IPDDP handle (cddp_b200_ipddp_create): same driving entry points as BatchedCLDDP plus the interior-point results."""

    def __init__(self, spec: dict, opts: Options, ipddp_opts: IpddpOptions, constraints, batch: int, device: int = 0):
        self.lib = load_library()
        self.pspec = ProblemSpec(dict(spec, lb=None, ub=None))
        self.opts, self.ipddp_opts = opts, ipddp_opts
        self.cset = constraints if isinstance(constraints, ConstraintSet) else ConstraintSet(constraints)
        self.B, self.n, self.m, self.N = int(batch), self.pspec.n, self.pspec.m, self.pspec.N
        self.handle = C.c_void_p()
        src = spec.get("model_source")
        _check(self.lib.cddp_b200_ipddp_create_ex(C.byref(self.pspec.struct), C.byref(opts), C.byref(ipddp_opts), self.cset.array,
                                                  self.cset.nc, src.encode() if src else None, self.B, device,
                                                  C.byref(self.handle)))
        self.num_alphas = len(build_alphas(opts))
        d = C.c_int(0)
        _check(self.lib.cddp_b200_ipddp_dual_dim(self.handle, C.byref(d)))
        self.d = d.value
        self._keep = []

    def get_ipddp_solution(self, trajectories: bool = True) -> dict:
        B, N, d = self.B, self.N, self.d
        out = {"Y": np.zeros((B, N, d)), "S": np.zeros((B, N, d)), "G": np.zeros((B, N, d))} if trajectories else {}
        sc = np.empty((B, 8))
        _check(self.lib.cddp_b200_ipddp_get_solution(self.handle, _ptr(out.get("Y")) if d else None,
                                                     _ptr(out.get("S")) if d else None, _ptr(out.get("G")) if d else None,
                                                     _ptr(sc)))
        for i, k in enumerate(("mu", "merit", "inf_pr", "inf_comp", "step_norm", "alpha_du", "alpha_pr_max", "alpha_du_max")):
            out[k] = sc[:, i].copy()
        return out

    def get_iteration_state(self) -> dict:
        """cddp_b200_ipddp_get_iteration_state: Lambda_T, filter points, filter size, {filter_theta, logsum, lamh}."""
        B = self.B
        out = {"lamT": np.zeros((B, self.n)), "filter": np.zeros((B, 8, 2)), "filter_size": np.zeros(B, dtype=np.int32)}
        sc = np.zeros((B, 3))
        _check(self.lib.cddp_b200_ipddp_get_iteration_state(self.handle, _ptr(out["lamT"]), _ptr(out["filter"]),
                                                            C.c_void_p(out["filter_size"].ctypes.data), _ptr(sc)))
        out["filter_theta"], out["logsum"], out["lamh"] = sc[:, 0].copy(), sc[:, 1].copy(), sc[:, 2].copy()
        return out

    def get_ipddp_gains(self) -> dict:
        B, N, d, n = self.B, self.N, self.d, self.n
        out = {"ky": np.zeros((B, N, d)), "Ky": np.zeros((B, N, d, n)), "ks": np.zeros((B, N, d)), "Ks": np.zeros((B, N, d, n))}
        _check(self.lib.cddp_b200_ipddp_get_gains(self.handle, _ptr(out["ky"]), _ptr(out["Ky"]), _ptr(out["ks"]), _ptr(out["Ks"])))
        return out

    def get_line_search(self) -> np.ndarray:
        t = np.empty((self.B, self.num_alphas, 4))
        _check(self.lib.cddp_b200_ipddp_get_line_search(self.handle, _ptr(t)))
        return t

    def get_history(self):
        cap = self.opts.max_iterations + 1
        h = np.empty((self.B, cap, IPDDP_HISTORY_COLS))
        lens = np.empty(self.B, dtype=np.int32)
        _check(self.lib.cddp_b200_ipddp_get_history(self.handle, _ptr(h), _ptr(lens)))
        return h, lens


def compile_user_model(source: str, n: int, m: int) -> int:
    """cddp_b200_compile_user_model: NVRTC-compiles a user dynamics source for (n, m); needs no GPU.  Returns the cubin
    size; raises CddpB200Error(code 6) with the compiler log on failure."""
    sz = C.c_size_t(0)
    _check(load_library().cddp_b200_compile_user_model(source.encode(), int(n), int(m), C.byref(sz)))
    return sz.value


def solve_host(spec: dict, opts: Options, x0, xref, X0, U0, ref_traj=None, device: int = 0) -> dict:
    """cddp_b200_solve_host: the one-shot batched call (host buffers in/out, copies inside)."""
    lib = load_library()
    ps = ProblemSpec(spec)
    x0 = _f64(x0)
    B, n, m, N = x0.shape[0], ps.n, ps.m, ps.N
    xref = _f64(xref, (B, n))
    X = _f64(X0, (B, N + 1, n)).copy()
    U = _f64(U0, (B, N, m)).copy()
    rt = _f64(ref_traj, (B, N + 1, n)) if ref_traj is not None else None
    out = {"X": X, "U": U, "K": np.empty((B, N, m, n)), "cost": np.empty(B), "iterations": np.empty(B, dtype=np.int32),
           "status": np.empty(B, dtype=np.int32), "alpha": np.empty(B), "reg": np.empty(B), "inf_du": np.empty(B)}
    _check(lib.cddp_b200_solve_host(C.byref(ps.struct), C.byref(opts), B, device, _ptr(x0), _ptr(xref), _ptr(rt),
                                    _ptr(X), _ptr(U), _ptr(out["K"]), _ptr(out["cost"]), _ptr(out["iterations"]),
                                    _ptr(out["status"]), _ptr(out["alpha"]), _ptr(out["reg"]), _ptr(out["inf_du"])))
    return out


def status_string(code: int) -> str:
    return load_library().cddp_b200_status_string(int(code)).decode()
