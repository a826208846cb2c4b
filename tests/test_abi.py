"""The C-ABI shared library loads and exports every symbol include/cddp_b200.h declares; option defaults equal
the reference's member initialisers; argument validation and the no-CPU-fallback rule.  No compute calls."""
import ctypes as C
import os
import re
import subprocess

import numpy as np
import pytest

from conftest import ROOT, has_cuda


def header_symbols():
    src = open(os.path.join(ROOT, "include", "cddp_b200.h")).read()
    return sorted(set(re.findall(r"CDDP_B200_API\s+[\w\s\*]+?\b(cddp_b200_\w+)\s*\(", src)))


def test_header_symbols_exported(cddp):
    names = header_symbols()
    assert len(names) >= 35
    lib = cddp.load_library()
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/cddp_b200.h but not exported"
    assert sorted(cddp.ABI_SYMBOLS) == names, "python binding's symbol list drifted from the header"
    out = subprocess.run(["nm", "-D", "--defined-only", cddp.LIB_PATH], capture_output=True, text=True).stdout
    exported = set(re.findall(r"\bT (cddp_b200_\w+)", out))
    assert exported == set(names), "the library must export exactly the declared C ABI (hidden visibility otherwise)"


def test_library_is_sm100a_cuda(cddp):
    out = subprocess.run(["cuobjdump", "-lelf", cddp.LIB_PATH], capture_output=True, text=True).stdout
    assert "sm_100a" in out, "libcddp_b200.so must carry sm_100a SASS"


def test_abi_version_and_strings(cddp):
    lib = cddp.load_library()
    assert lib.cddp_b200_abi_version() == 1
    # reference status strings: cddp_solver_base.cpp:69,82,162; clddp_solver.cpp:209,270,274
    expect = {1: "OptimalSolutionFound", 2: "AcceptableSolutionFound", 3: "MaxIterationsReached",
              4: "RegularizationLimitReached_NotConverged", 5: "MaxCpuTimeReached"}
    for k, v in expect.items():
        assert cddp.status_string(k) == v
    assert lib.cddp_b200_error_string(0) == b"ok"
    assert b"no CPU fallback" in lib.cddp_b200_error_string(2)


def test_default_options_equal_reference(cddp):
    """options.hpp:41-50,58-66,103-104,211-215,245; boxqp.hpp:30-41."""
    o = cddp.default_options()
    ref = dict(tolerance=1e-5, acceptable_tolerance=1e-6, max_iterations=1, enable_parallel=0, max_cpu_time=0.0,
               termination_scaling_max_factor=100.0, ls_max_iterations=11, ls_initial_step_size=1.0, ls_min_step_size=1e-8,
               ls_step_reduction_factor=0.5, reg_initial_value=1e-6, reg_update_factor=10.0, reg_max_value=1e7,
               reg_min_value=1e-10, qp_max_iterations=100, qp_min_gradient_norm=1e-8, qp_min_relative_improvement=1e-8,
               qp_step_decrease_factor=0.6, qp_min_step_size=1e-22, qp_armijo_constant=0.1, armijo_constant=1e-4)
    for k, v in ref.items():
        assert getattr(o, k) == v, k


def test_options_layout_matches_oracle(cddp, ob):
    """One ctypes layout feeds both the product and the oracle (oracle/cddp_oracle.h)."""
    assert C.sizeof(cddp.Options) == C.sizeof(ob.Options)
    assert [f[0] for f in cddp.Options._fields_] == [f[0] for f in ob.Options._fields_]
    assert C.sizeof(cddp.Problem) == C.sizeof(ob.Problem)
    a, b = cddp.default_options(), ob.make_options()
    assert bytes(a) == bytes(b)


def test_alpha_schedule(cddp, ob):
    """detail::buildLineSearchAlphas (cddp_context_utils.cpp:37-57)."""
    np.testing.assert_array_equal(cddp.build_alphas(cddp.default_options()), 0.5 ** np.arange(11))
    for kw in (dict(ls_max_iterations=15), dict(ls_max_iterations=40, ls_min_step_size=1e-3), dict(ls_step_reduction_factor=0.1, ls_max_iterations=12)):
        np.testing.assert_array_equal(cddp.build_alphas(cddp.default_options(**kw)), ob.build_alphas(ob.make_options(**kw)))


def test_argument_validation(cddp, problems):
    lib = cddp.load_library()
    cfg = problems.make_config("pendulum", batch=1, horizon=10)
    h = C.c_void_p()
    o = cddp.default_options()
    bad = dict(cfg["spec"], n=3)  # pendulum must be n=2
    assert lib.cddp_b200_create(C.byref(cddp.ProblemSpec(bad.copy() | {"Q": np.zeros((3, 3)), "Qf": np.zeros((3, 3))}).struct),
                                C.byref(o), 1, 0, C.byref(h)) == 1
    ps = cddp.ProblemSpec(cfg["spec"])
    assert lib.cddp_b200_create(C.byref(ps.struct), C.byref(o), 0, 0, C.byref(h)) == 1  # batch < 1
    ps.struct.model = 99
    assert lib.cddp_b200_create(C.byref(ps.struct), C.byref(o), 1, 0, C.byref(h)) == 2  # unsupported model
    too_many = cddp.default_options(ls_max_iterations=40, ls_step_reduction_factor=0.9)  # 40 alphas > 32 lanes
    ps = cddp.ProblemSpec(cfg["spec"])
    assert lib.cddp_b200_create(C.byref(ps.struct), C.byref(too_many), 1, 0, C.byref(h)) == 1
    assert lib.cddp_b200_create(None, C.byref(o), 1, 0, C.byref(h)) == 1
    assert lib.cddp_b200_solve(None) == 1 and lib.cddp_b200_destroy(None) == 0


@pytest.mark.skipif(has_cuda(), reason="checks the behaviour on a machine WITHOUT a GPU")
def test_no_cpu_fallback_without_gpu(cddp, problems):
    """north_star: no CPU fallback — without a device, create fails loudly with CDDP_B200_ERR_CUDA."""
    cfg = problems.make_config("pendulum", batch=1, horizon=10)
    with pytest.raises(cddp.CddpB200Error) as e:
        cddp.BatchedCLDDP(cfg["spec"], cddp.default_options(), 1)
    assert e.value.code == 3
    with pytest.raises(cddp.CddpB200Error):
        cddp.solve_host(cfg["spec"], cddp.default_options(), cfg["x0"], cfg["xref"], cfg["X0"], cfg["U0"])


def test_product_never_touches_the_oracle():
    """The oracle is test infrastructure: nothing under cddp-cpp_b200/ or include/ may reference it.  ANY mention of
    the word is flagged; the only exemptions are host/tests/ (the C++ GPU test uses it as its checker) and the one
    Makefile rule that builds that test binary — every other Makefile line is checked."""
    bad = []
    for base in ("cddp-cpp_b200", "include"):
        for dp, _, fs in os.walk(os.path.join(ROOT, base)):
            if "build" in dp.split(os.sep) or "__pycache__" in dp:
                continue
            for f in fs:
                if not f.endswith((".py", ".cu", ".cuh", ".h", ".hpp", ".cpp", "Makefile")):
                    continue
                path = os.path.join(dp, f)
                if os.sep + "tests" + os.sep in path:
                    continue  # host/tests may link the oracle only as a checker
                txt = open(path, errors="ignore").read()
                # prose may name the oracle (comments, docstrings); code, include paths and link lines may not
                if f.endswith(".py"):
                    txt = re.sub(r'(\'\'\'|""")[\s\S]*?\1', "", txt)
                    txt = re.sub(r"#.*", "", txt)
                elif f != "Makefile":
                    txt = re.sub(r"/\*[\s\S]*?\*/", "", txt)
                    txt = re.sub(r"//.*", "", txt)
                lines = txt.split("\n")
                rule = ""
                for ln in lines:
                    if f == "Makefile" and ln and not ln[0].isspace() and ":" in ln:
                        rule = ln.split(":")[0].strip()
                    if re.search(r"oracle", ln, re.IGNORECASE):
                        in_test_rule = f == "Makefile" and (rule.startswith("tests/") or ln.lstrip().startswith("#"))
                        if not in_test_rule:
                            bad.append(f"{path}: {ln.strip()}")
    assert not bad, bad
