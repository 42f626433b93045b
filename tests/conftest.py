import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


# the reference's controller-only tests switch these off (tests/test_determinism.cpp:67-69); activated, they demand sensor data
OPEN_LOOP = {"stabilizer": {"activated": False}, "collision_detection": {"activated": False}, "closed_loop": False}


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session", autouse=True)
def _built():
    """The product library and the oracle must exist (python __graft_entry__.py builds both)."""
    import subprocess
    so = os.path.join(ROOT, "inria_wbc_b200", "libiwbc_b200.so")
    if not os.path.exists(so):
        subprocess.check_call([sys.executable, os.path.join(ROOT, "__graft_entry__.py")])
    yield


@pytest.fixture(scope="session")
def franka():
    import inria_wbc_b200 as iw
    return iw.Problem("franka/pos_tracker.yaml")


@pytest.fixture(scope="session")
def talos():
    import inria_wbc_b200 as iw
    # stabilizer off, as the reference's open-loop tests run the controller (tests/test_determinism.cpp:67): with it activated the
    # controller needs F/T + IMU sensor data for every update (tests/test_stabilizer.py)
    return iw.Problem("talos/talos_pos_tracker.yaml", overrides=OPEN_LOOP)


@pytest.fixture(scope="session")
def talos_fast():
    """Talos with the `talos_fast.urdf` model of the reference's `-f` runs (grippers reduced to their driven joint)."""
    import inria_wbc_b200 as iw
    return iw.Problem("talos/talos_pos_tracker.yaml", overrides=dict(OPEN_LOOP, urdf="talos/talos_fast.urdf"))


@pytest.fixture(scope="session")
def tiago():
    import inria_wbc_b200 as iw
    return iw.Problem("tiago/pos_tracker.yaml", overrides={"urdf": "tiago/tiago_steel_no_wheel.urdf"})


@pytest.fixture(scope="session")
def icub():
    import inria_wbc_b200 as iw
    return iw.Problem("icub/humanoid_pos_tracker.yaml", overrides=OPEN_LOOP)


def random_states(P, B, seed, qn=0.02, vn=0.05, base=False):
    import numpy as np
    rng = np.random.default_rng(seed)
    q = np.tile(P.q0(), (B, 1))
    v = np.zeros((B, P.nv))
    na = P.na
    q[:, P.nq - na:] += rng.normal(0, qn, (B, na))
    v[:, P.nv - na:] += rng.normal(0, vn, (B, na))
    if base and P.nv > na:
        v[:, :6] = rng.normal(0, vn, (B, 6))
        q[:, :3] += rng.normal(0, 0.05, (B, 3))
        dq = rng.normal(0, 0.1, (B, 4))
        qq = q[:, 3:7] + dq
        q[:, 3:7] = qq / np.linalg.norm(qq, axis=1, keepdims=True)
    return q, v
