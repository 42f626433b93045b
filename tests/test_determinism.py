"""Re-hosted tests/test_determinism.cpp (run-to-run reproducibility of open-loop roll-outs, 1e-8 on q),
on the CPU oracle; tests/test_gpu_parity.py repeats it on the GPU path (GPU<->GPU bitwise, GPU<->oracle 1e-8)."""
import os

import numpy as np
import pytest

import inria_wbc_b200 as iw
import oracle_lib as ol
from inria_wbc_b200 import behaviors, ETC_PATH

TICKS = 1500  # the reference runs 50 000 ticks x 5 controllers x 4 sub-tests; sized for a CPU suite of a few minutes


def _fresh():
    # stabilizer / collision detection / closed loop forced off as tests/test_determinism.cpp:66-70 does
    return iw.Problem("talos/talos_pos_tracker.yaml", overrides={"closed_loop": False, "check_model_collisions": False,
                                                                  "stabilizer": {"activated": False},
                                                                  "collision_detection": {"activated": False}})


@pytest.mark.parametrize("behavior", ["arm", "squat"])
def test_rollouts_are_reproducible(behavior):
    runs = []
    for k in range(3):
        P = _fresh()  # a fresh controller per run, like the reference's sub-test
        cfg = os.path.join(ETC_PATH, "talos", behavior + ".yaml")
        b = behaviors.Cartesian(P, cfg) if behavior == "arm" else behaviors.MoveCom(P, cfg)
        refs = b.table(TICKS)
        r = ol.rollout(P, P.q0(), np.zeros(P.nv), refs, TICKS)
        assert r["rc"] == 0, "QP failed at tick %d" % r["rc"]
        runs.append(r["q_cmd"])
    for k in range(1, 3):
        assert np.abs(runs[0] - runs[k]).max() < 1e-8
        assert np.array_equal(runs[0], runs[k])  # in fact bitwise
    # the behavior really moves the robot
    assert np.abs(runs[0][-1] - runs[0][0]).max() > 1e-3
    # mimic-filtered command vector (Controller::q(true)): drop the mimic dofs, same check
    dofs = P.names("dofs")
    keep = [i for i in range(6)] + [6 + i for i, d in enumerate(dofs) if d not in set(P.names("mimics"))]
    assert len(keep) == P.nv - 12
    assert np.array_equal(runs[0][:, keep], runs[1][:, keep])


def test_squat_tracks_the_com_reference():
    P = _fresh()
    b = behaviors.MoveCom(P, os.path.join(ETC_PATH, "talos", "squat.yaml"))
    refs = b.table(TICKS)
    r = ol.rollout(P, P.q0(), np.zeros(P.nv), refs, TICKS)
    assert r["rc"] == 0
    off, _ = P.ref_slot("com")
    d = ol.cycle_debug(P, r["q"], r["v"], refs[-1])
    # the CoM task (weight 1000, kp 30) pulls the model CoM down with the reference; the posture / feet tasks resist,
    # so it lags by ~2 cm on this model
    assert np.abs(d["com"][:3] - refs[-1, off:off + 3]).max() < 3e-2
    assert d["com"][2] < P.default_refs()[off + 2] - 0.1
    assert refs[-1, off + 2] < refs[0, off + 2] - 0.05
