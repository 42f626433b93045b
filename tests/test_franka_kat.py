"""The reference's only known-answer value for this path: Franka pos-tracker +
cartesian_line, mean TSID-model ee position error over 10 s @ 1 kHz, tests/ref_test_franka.yaml."""
import json
import os

import numpy as np

import oracle_lib as ol
from inria_wbc_b200 import behaviors, ETC_PATH

GOLD = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "ref_test_franka.json")))


def _mean_error(P, refs):
    ticks = refs.shape[0]
    off, _ = P.ref_slot("ee")
    r = ol.rollout(P, P.q0(), np.zeros(P.nv), refs, ticks)
    assert r["rc"] == 0 and (r["status"] == 0).all()
    # the reference reads the model pose held by tsid's data, i.e. at the state the tick started from
    # (controller->model_joint_pos, tests/test_franka.cpp:214-218)
    qs = np.vstack([P.q0()[None], r["q_cmd"]])
    ee = np.array([ol.frame(P, qs[t], np.zeros(P.nv), 0)["oMf"][9:12] for t in range(ticks + 1)])
    before = np.linalg.norm(ee[:-1] - refs[:, off:off + 3], axis=1).mean()
    after = np.linalg.norm(ee[1:] - refs[:, off:off + 3], axis=1).mean()
    return before, after


def test_franka_known_answer(franka):
    ticks = int(GOLD["duration_s"] / GOLD["dt"])
    gold = GOLD["cases"]["servo"]
    # (1) the reference's own acceptance criterion with the behavior as it is in the tree today
    refs = behaviors.Cartesian(franka, os.path.join(ETC_PATH, "franka", "cartesian_line.yaml")).table(ticks)
    before, _ = _mean_error(franka, refs)
    assert before <= gold["ee_error_tsid"] + GOLD["tolerance"]
    # (2) the recorded value itself: the golden file (2021-07-13) predates the velocity / acceleration feed-forward of
    # generic::cartesian (src/behaviors/generic/cartesian.cpp:76-77); with position-only references the oracle
    # reproduces both recorded numbers (TSID model, and the servo-tracked DART model = one tick later) to 3e-4 relative
    refs = behaviors.Cartesian(franka, os.path.join(ETC_PATH, "franka", "cartesian_line.yaml"), feed_forward=False).table(ticks)
    before, after = _mean_error(franka, refs)
    assert abs(before - gold["ee_error_tsid"]) < 3e-4 * gold["ee_error_tsid"], before
    assert abs(after - gold["ee_error_dart"]) < 3e-4 * gold["ee_error_dart"], after
    assert abs((before - after) - (gold["ee_error_tsid"] - gold["ee_error_dart"])) < 1e-6
    for case in ("velocity", "spd"):
        assert GOLD["cases"][case]["ee_error_tsid"] == gold["ee_error_tsid"]  # open loop: independent of the actuator model
