"""The C++ host-side mirror of the reference's controller surface (include/inria_wbc_b200/batch_pos_tracker.hpp) over the
C ABI: a host-only API test and the GPU re-host of the reference's tests/test_determinism.cpp."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "inria_wbc_b200")


def _build(name):
    src = os.path.join(ROOT, "tests", "cpp", name + ".cpp")
    exe = os.path.join(ROOT, "tests", "cpp", name + ".bin")
    deps = [src, os.path.join(ROOT, "include", "inria_wbc_b200", "batch_pos_tracker.hpp"), os.path.join(PKG, "libiwbc_b200.so")]
    if not os.path.exists(exe) or any(os.path.getmtime(exe) < os.path.getmtime(d) for d in deps):
        subprocess.check_call(["g++", "-O1", "-std=c++17", "-I", os.path.join(ROOT, "include"), src, "-o", exe, "-L", PKG, "-liwbc_b200",
                               "-Wl,-rpath," + PKG, "-Wl,-rpath,/usr/local/cuda/lib64", "-pthread"])
    return exe


def _run(exe, *args):
    return subprocess.run([exe, os.path.join(PKG, "etc"), os.path.join(PKG, "robots"), *args], capture_output=True, text=True, timeout=900)


def test_host_api_mirrors_reference_error_behaviour():
    r = _run(_build("test_host_api"))
    assert r.returncode == 0, r.stdout + r.stderr
    assert "OK" in r.stdout


@pytest.mark.gpu
def test_batch_determinism_rehost():
    """tests/test_determinism.cpp at the reference's scale: {arm, squat} x {talos.urdf, talos_fast.urdf}, 5 fresh controllers each,
    50 000 ticks, pairwise 1e-8 on the full and the mimic-filtered command vectors."""
    r = _run(_build("test_batch_determinism"), "50000", "5")
    print(r.stdout)
    assert r.returncode == 0, r.stdout + r.stderr
    assert r.stdout.strip().endswith("OK") and r.stdout.count("-> ok") == 4


@pytest.mark.gpu
def test_batch_behavior_rehost():
    r = _run(_build("test_batch_behavior"))
    assert r.returncode == 0, r.stdout + r.stderr
    assert r.stdout.startswith("OK")
