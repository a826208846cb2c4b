"""User-model plugin (cddp_b200_create_ex / cddp_b200_compile_user_model): the device counterpart of subclassing
cddp::DynamicalSystem (include/cddp-cpp/cddp_core/dynamical_system.hpp:33-152).  CPU tests: NVRTC compiles the engine's own
kernels around a user source without a GPU, compile errors are reported with the compiler log, the oracle's native twins
of the plugin models agree with the numpy restatement.  GPU tests: parity of the plugin path with the oracle."""
import numpy as np
import pytest

from conftest import rel_err


def test_user_source_compiles_without_a_gpu(cddp, problems):
    for name in ("bicycle_user", "chain7_user", "manip7_user"):
        spec = problems.make_config(name, batch=1)["spec"]
        assert cddp.compile_user_model(spec["model_source"], spec["n"], spec["m"]) > 10000


def test_user_source_with_analytic_jacobian_compiles(cddp):
    src = """
    #define CDDP_USER_HAS_JACOBIAN
    template <class T> __device__ void cddp_user_dynamics(const T *x, const T *u, const double *p, T *xd) {
      xd[0] = x[1]; xd[1] = u[0] - p[0] * sin(x[0]);
    }
    __device__ void cddp_user_jacobian(const double *x, const double *u, const double *p, double *Fx, double *Fu) {
      Fx[0] = 0.0; Fx[1] = 1.0; Fx[2] = -p[0] * cos(x[0]); Fx[3] = 0.0; Fu[0] = 0.0; Fu[1] = 1.0;
    }"""
    assert cddp.compile_user_model(src, 2, 1) > 10000


def test_compile_error_is_reported_with_the_log(cddp):
    bad = "template <class T> __device__ void cddp_user_dynamics(const T *x, const T *u, const double *p, T *xd) { xd[0] = nope; }"
    with pytest.raises(cddp.CddpB200Error) as ei:
        cddp.compile_user_model(bad, 2, 1)
    assert ei.value.code == 6 and "nope" in str(ei.value) and "user_model_source.cu" in str(ei.value)
    with pytest.raises(cddp.CddpB200Error):  # missing definition: unresolved at link of the cubin / instantiation
        cddp.compile_user_model("// no dynamics here", 2, 1)


def test_model_user_without_source_is_unsupported_not_a_fallback(cddp, problems):
    """A host-only DynamicalSystem has no device dynamics: CDDP_B200_ERR_UNSUPPORTED_MODEL (code 2), never a CPU fallback."""
    spec = dict(problems.make_config("bicycle_user", batch=1)["spec"])
    spec.pop("model_source")
    with pytest.raises(cddp.CddpB200Error) as ei:
        cddp.BatchedCLDDP(spec, cddp.default_options(), 1)
    assert ei.value.code == 2


@pytest.mark.parametrize("name", ["bicycle_user", "chain7_user", "manip7_user"])
def test_oracle_twin_matches_numpy(ob, npo, problems, name):
    cfg = problems.make_config(name, batch=2, horizon=30)
    P, Pn = ob.OracleProblem(cfg["spec"]), npo.Problem(cfg["spec"])
    rng = np.random.default_rng(3)
    for _ in range(5):
        x, u = rng.standard_normal(P.n), 0.3 * rng.standard_normal(P.m)
        tol = 1e-12 if name == "manip7_user" else 1e-14  # 7 x 7 solve: Gaussian elimination vs LAPACK
        assert rel_err(ob.continuous_dynamics(P, x, u), Pn.f(x, u)) < tol
        assert rel_err(ob.discrete_dynamics(P, x, u), Pn.step(x, u)) < tol
        Fx, Fu = ob.jacobians(P, x, u)
        Fx2, Fu2 = Pn.jacobians(x, u)
        assert rel_err(Fx, Fx2) < 1e-11 and rel_err(Fu, Fu2) < 1e-11
    oo, on = ob.make_options(**cfg["options"]), npo.options(**cfg["options"])
    r = ob.solve(P, oo, cfg["x0"][0], cfg["xref"][0], cfg["X0"][0], cfg["U0"][0])
    q = npo.solve(Pn, on, cfg["x0"][0], cfg["xref"][0], cfg["X0"][0], cfg["U0"][0])
    assert r["iterations"] == q["iterations"] and abs(r["cost"] - q["cost"]) <= 1e-8 * abs(q["cost"])


# ------------------------------------------------------------------------------------------------ GPU
def test_manip7_is_a_lagrangian_system_and_the_plugin_text_is_the_twin(ob, npo, problems):
    """BASELINE config #5's model.  (1) With tau = 0 and no friction the total energy 1/2 qd^T M(q) qd + V(q) is conserved
    along the rollout: the Coriolis / centrifugal vector is the one that belongs to M(q) and G = dV/dq.  (2) The plugin
    source compiled on the HOST (T = double) agrees with the oracle's native twin to roundoff: the three statements of the
    model (CUDA text, C++ twin, numpy tensor form) are the same function.  (3) The state Jacobian of the accelerations is
    dense (every joint couples with every other one through M(q)^-1), unlike the chain model it replaces."""
    import ctypes
    import subprocess
    import tempfile
    cfg = problems.make_config("manip7_user", batch=1)
    spec = cfg["spec"]
    par = list(spec["params"])
    par[1] = 0.0
    Pn = npo.Problem(dict(spec, params=par, dt=1e-3))
    g, mass, ln = par[0], np.array(par[2:9]), np.array(par[9:16])
    mu = np.cumsum(mass[::-1])[::-1]

    def energy(x):
        q, qd = x[:7], x[7:]
        T = np.tril(np.ones((7, 7)))
        T[:, 0] = 0
        sig = T @ q
        A = np.array([[mu[max(i, j)] * ln[i] * ln[j] for j in range(7)] for i in range(7)])
        return 0.5 * qd @ (A * np.cos(sig[:, None] - sig[None, :])) @ qd - np.sum(mu[1:] * g * ln[1:] * np.sin(sig[1:]))

    rng = np.random.default_rng(1)
    x = rng.standard_normal(14)
    e0 = energy(x)
    for _ in range(300):
        x = Pn.step(x, np.zeros(7))
    assert abs(energy(x) - e0) < 1e-7 * abs(e0)
    src = ("#include <cmath>\n#define __device__\nusing std::sin; using std::cos;\n" + problems.MANIP7_SOURCE +
           '\nextern "C" void f(const double*x,const double*u,const double*p,double*xd){cddp_user_dynamics<double>(x,u,p,xd);}\n')
    d = tempfile.mkdtemp()
    open(d + "/m.cpp", "w").write(src)
    subprocess.check_call(["g++", "-O2", "-shared", "-fPIC", "-ffp-contract=off", "-o", d + "/m.so", d + "/m.cpp"])
    lib = ctypes.CDLL(d + "/m.so")
    P = ob.OracleProblem(spec)
    pp = np.array(spec["params"], float)
    for _ in range(10):
        x, u, out = 2 * rng.standard_normal(14), 20 * rng.standard_normal(7), np.zeros(14)
        lib.f(*(a.ctypes.data_as(ctypes.c_void_p) for a in (x, u, pp, out)))
        assert rel_err(out, ob.continuous_dynamics(P, x, u)) < 1e-13
    Fx, _ = ob.jacobians(P, rng.standard_normal(14), rng.standard_normal(7))
    assert (np.abs(Fx[7:, 1:]) > 1e-9).mean() > 0.95


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["bicycle_user", "chain7_user", "manip7_user"])
def test_plugin_clddp_parity(cddp, ob, problems, name):
    """One iteration step by step (linearisation by dual numbers in-kernel vs the oracle, sweep, line search) and whole
    solves, exactly as for the built-in models (tests/test_gpu_parity.py)."""
    B = 5
    cfg = problems.make_config(name, batch=B, horizon=60)
    opts = cfg["options"]
    s = cddp.BatchedCLDDP(cfg["spec"], cddp.default_options(**opts), B)
    s.set_instances(cfg["x0"], cfg["xref"], cfg["X0"], cfg["U0"])
    P, oo = ob.OracleProblem(cfg["spec"]), ob.make_options(**opts)
    s.initialize()
    s.linearize()
    A, Bm = s.get_linearization()
    X0 = cfg["X0"].copy()
    X0[:, 0] = cfg["x0"]
    s.backward_pass()
    sw, K, k = s.get_sweep(), s.get_solution()["K"], s.get_feedforward()
    s.forward_pass()
    fw = s.get_forward()
    alphas = ob.build_alphas(oo)
    for b in range(B):
        Ao, Bo = ob.linearize(P, X0[b], cfg["U0"][b])
        assert rel_err(A[b], Ao) < 1e-12 and rel_err(Bm[b], Bo) < 1e-12
        r = ob.backward_pass(P, oo, X0[b], cfg["U0"][b], cfg["xref"][b], opts.get("reg_initial_value", 1e-6))
        assert r["ok"] and sw["ok"][b] == 1
        assert rel_err(K[b], r["K"]) < 1e-9 and rel_err(k[b], r["k"]) < 1e-9 and rel_err(sw["dV"][b], r["dV"]) < 1e-9
        c0 = ob.trajectory_cost(P, X0[b], cfg["U0"][b], cfg["xref"][b])
        first = -1
        for ai, a in enumerate(alphas):
            f = ob.forward_pass(P, oo, cfg["x0"][b], X0[b], cfg["U0"][b], cfg["xref"][b], r["K"], r["k"], r["dV"], c0, a)
            if np.isfinite(f["cost"]):
                assert abs(fw["costs"][b, ai] - f["cost"]) < 1e-8 * abs(f["cost"])
            if f["success"] and first < 0:
                first = ai
        assert fw["accepted"][b] == first
    s.set_instances(cfg["x0"], cfg["xref"], cfg["X0"], cfg["U0"])
    s.solve()
    g = s.get_solution()
    o = ob.solve_batch(P, oo, cfg["x0"], cfg["xref"], cfg["X0"], cfg["U0"], nthreads=4)
    with ob.variant():
        o2 = ob.solve_batch(P, oo, cfg["x0"], cfg["xref"], cfg["X0"], cfg["U0"], nthreads=4)
    stable = (np.abs(o["cost"] - o2["cost"]) <= 1e-7 * np.abs(o["cost"])) & (o["iterations"] == o2["iterations"])
    assert stable.sum() >= 3
    assert (g["iterations"][stable] == o["iterations"][stable]).all() and (g["status"][stable] == o["status"][stable]).all()
    assert (np.abs(g["cost"][stable] - o["cost"][stable]) <= 1e-6 * np.abs(o["cost"][stable])).all()
    assert (g["U"] >= np.asarray(cfg["spec"]["lb"]) - 1e-12).all() and (g["U"] <= np.asarray(cfg["spec"]["ub"]) + 1e-12).all()
    s.close()


@pytest.mark.gpu
def test_speculative_first_alpha_and_option_switch(cddp, ob, problems):
    """User models try alphas_[0] on one lane per instance before the 16-wide line search (sequential rule only).  (1) Per
    iteration, the accepted step and the cost must be those of the oracle's sequential line search.  (2) Switching a live
    handle to enable_parallel (full line search alone) must not be masked by flags the speculative pass left behind."""
    B = 6
    cfg = problems.make_config("bicycle_user", batch=B, horizon=60)
    opts = dict(cfg["options"], max_iterations=30)
    P, oo = ob.OracleProblem(cfg["spec"]), ob.make_options(**opts)
    s = cddp.BatchedCLDDP(cfg["spec"], cddp.default_options(**opts), B)
    s.set_first_alpha_speculation(1)  # a batch this small would not speculate on its own (latency-bound)
    s.set_instances(cfg["x0"], cfg["xref"], cfg["X0"], cfg["U0"])
    s.enable_history(True)
    s.initialize()
    s.iterate(4)
    h, lens = s.get_history()
    for b in range(B):
        o = ob.solve(P, ob.make_options(**dict(opts, max_iterations=4)), cfg["x0"][b], cfg["xref"][b], cfg["X0"][b], cfg["U0"][b],
                     history=True)
        n = min(lens[b], o["history"].shape[0])
        assert n >= 2
        assert rel_err(h[b, :n, :2], o["history"][:n, :2]) < 1e-7  # accepted cost and alpha per iteration
    mid = s.get_solution()
    # (2) continue with enable_parallel on the SAME handle: every instance (all still far from converged) must keep moving
    assert (mid["status"] == 0).all()
    s.set_options(cddp.default_options(**dict(opts, enable_parallel=1)))
    s.iterate(3)
    a = s.get_solution()
    s.close()
    for b in range(B):
        assert not np.array_equal(a["X"][b], mid["X"][b]), b
        assert a["cost"][b] < mid["cost"][b]


@pytest.mark.gpu
def test_plugin_ipddp_parity(cddp, ob, problems):
    B = 4
    cfg = problems.make_config("bicycle_user_ipddp", batch=B, horizon=60)
    s = cddp.BatchedIPDDP(cfg["spec"], cddp.default_options(**cfg["options"]), cddp.default_ipddp_options(), cfg["constraints"], B)
    P, oo, oi, cs = ob.OracleProblem(cfg["spec"]), ob.make_options(**cfg["options"]), ob.make_ipddp_options(), ob.ConstraintSet(cfg["constraints"])
    for iters in (0, 2):
        s.set_instances(cfg["x0"], cfg["xref"], None, cfg["U0"])
        s.initialize()
        if iters:
            s.iterate(iters)
        s.linearize()
        s.backward_pass()
        sol, ips, gains, kff = s.get_solution(), s.get_ipddp_solution(), s.get_ipddp_gains(), s.get_feedforward()
        for b in range(B):
            if iters and ob.ipddp_solve(P, ob.make_options(**dict(cfg["options"], max_iterations=iters)), oi, cs, cfg["x0"][b],
                                        cfg["xref"][b], cfg["U0"][b])["decision_margin"] < 1e-9:
                continue
            r = ob.ipddp_probe(P, oo, oi, cs, cfg["x0"][b], cfg["xref"][b], cfg["U0"][b], iters)
            for key, val in (("X", sol["X"][b]), ("U", sol["U"][b]), ("S", ips["S"][b]), ("Y", ips["Y"][b]), ("ku", kff[b]),
                             ("Ku", sol["K"][b]), ("ky", gains["ky"][b]), ("Ks", gains["Ks"][b])):
                assert rel_err(val, r[key]) < 1e-9, (iters, b, key)
            assert abs(ips["alpha_pr_max"][b] - r["alpha_pr_max"]) <= 1e-9 * r["alpha_pr_max"]
    s.set_instances(cfg["x0"], cfg["xref"], None, cfg["U0"])
    s.solve()
    g = s.get_solution()
    assert np.isfinite(g["cost"]).all()
    dist = np.hypot(g["X"][:, :, 0] - 3.0, g["X"][:, :, 1] - 1.2)
    assert (dist.min(axis=1) > 0.5 - 5e-2).all()
    s.close()


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["chain7_user_ipddp", "manip7_user_ipddp"])
def test_config5_mixed_constraints_ipddp(cddp, ob, problems, name):
    """BASELINE config #5 as worded (n=14, m=7, N=150, mixed constraints): the 7-DOF manipulator plugin model (and the
    earlier 7-joint chain) under IPDDP with a torque box and a joint / rate box (d = 42).  First backward pass against the
    oracle; whole solves are held to validity properties (the state-box rows at t = 0 make every line search
    roundoff-decided, see tests/test_gpu_ipddp.py)."""
    B = 4
    cfg = problems.make_config(name, batch=B, horizon=150)
    s = cddp.BatchedIPDDP(cfg["spec"], cddp.default_options(**cfg["options"]), cddp.default_ipddp_options(), cfg["constraints"], B)
    assert s.d == 42
    P, oo, oi, cs = ob.OracleProblem(cfg["spec"]), ob.make_options(**cfg["options"]), ob.make_ipddp_options(), ob.ConstraintSet(cfg["constraints"])
    s.set_instances(cfg["x0"], cfg["xref"], None, cfg["U0"])
    s.initialize()
    s.linearize()
    s.backward_pass()
    sol, ips, gains, kff = s.get_solution(), s.get_ipddp_solution(), s.get_ipddp_gains(), s.get_feedforward()
    for b in range(B):
        r = ob.ipddp_probe(P, oo, oi, cs, cfg["x0"][b], cfg["xref"][b], cfg["U0"][b], 0)
        for key, val in (("X", sol["X"][b]), ("S", ips["S"][b]), ("Y", ips["Y"][b]), ("G", ips["G"][b]), ("ku", kff[b]), ("Ku", sol["K"][b]),
                         ("ky", gains["ky"][b]), ("Ky", gains["Ky"][b]), ("ks", gains["ks"][b]), ("Ks", gains["Ks"][b])):
            assert rel_err(val, r[key]) < 1e-9, (b, key)
        assert abs(ips["alpha_pr_max"][b] - r["alpha_pr_max"]) <= 1e-9 * r["alpha_pr_max"]
        assert abs(sol["inf_du"][b] - r["inf_du"]) <= 1e-9 * r["inf_du"]
    s.set_instances(cfg["x0"], cfg["xref"], None, cfg["U0"])
    s.solve()
    g, gi = s.get_solution(want_K=False), s.get_ipddp_solution()
    assert np.isfinite(g["cost"]).all() and (gi["S"] > 0).all() and (gi["Y"] > 0).all()
    conv = np.isin(g["status"], (1, 2))
    assert conv.any()
    qmax = 1.0 if name.startswith("chain7") else 3.3
    assert (np.abs(g["U"][conv]) <= 50.0 + 1e-6).all() and (np.abs(g["X"][conv][:, :-1, :7]) <= qmax + 1e-4).all()
    for b in range(B):
        assert abs(ob.trajectory_cost(P, g["X"][b], g["U"][b], cfg["xref"][b]) - g["cost"][b]) <= 1e-10 * abs(g["cost"][b])
    s.close()
