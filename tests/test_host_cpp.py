"""The C++ host mirror of the reference API (cddp-cpp_b200/host/): facade / registry / plugin contract on CPU with a
mock solver (the reference's own test_cddp_core.cpp cases), and — on the GPU — cddp::CDDP::solve("CLDDP") through
the registry to the B200 solver plus the batched facade, checked against the oracle inside the C++ test binary."""
import os
import subprocess

import pytest

from conftest import ROOT

HOST = os.path.join(ROOT, "cddp-cpp_b200", "host")


def _run(binary):
    path = os.path.join(HOST, "tests", binary)
    if not os.path.exists(path):
        subprocess.run(["make", "-C", os.path.join(ROOT, "oracle"), "-s"], check=True)
        subprocess.run(["make", "-C", HOST, "-s"], check=True)
    r = subprocess.run([path], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert " 0 failed" in r.stdout
    return r.stdout


def test_facade_and_registry_contract():
    out = _run("test_cddp_core")
    for case in ("SolverPrecedence", "UnknownSolverErrorHandling", "HostOnlyDynamicsIsRejectedNotFallenBack"):
        assert case in out


def test_host_library_links_only_the_c_abi():
    """The host mirror reaches the device ONLY through the C ABI symbols of include/cddp_b200.h."""
    out = subprocess.run(["nm", "-D", "--undefined-only", os.path.join(HOST, "libcddp_b200_host.so")], capture_output=True, text=True).stdout
    used = {l.split()[-1] for l in out.splitlines() if "cddp_b200_" in l}
    assert used and all(u.startswith("cddp_b200_") for u in used)
    assert "oracle" not in out


@pytest.mark.gpu
def test_plugin_route_and_batched_facade_on_gpu():
    out = _run("test_b200_solver")
    assert "SolvePendulum" in out and "SolveQuadrotorBatch" in out and "SolveUnicycleObstacleIPDDP" in out and "SolveUserDefinedBicycle" in out
