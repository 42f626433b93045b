"""the reference's own, unmodified controller YAML files (read where they lie under /root/reference, only in the
build container -- the GPU box has no /root/reference) load through the host compiler and give the same compiled problem as the
normalised copies packaged under inria_wbc_b200/etc/ (tools/import_reference_configs.py).  `base_path` is overridden the way the
reference's tests do it (tests/test_determinism.cpp:70): the files hard-code /home/pal/inria_wbc/etc/<robot>."""
import os

import numpy as np
import pytest

import inria_wbc_b200 as iw

REF = "/root/reference/etc"
pytestmark = pytest.mark.skipif(not os.path.isdir(REF), reason="the reference tree is only mounted in the build container")

CASES = [("talos/talos_pos_tracker.yaml", {}), ("icub/humanoid_pos_tracker.yaml", {}), ("franka/pos_tracker.yaml", {}),
         ("tiago/pos_tracker.yaml", {"urdf": "tiago/tiago_steel_no_wheel.urdf"})]  # tests/test_all_robots.cpp:329 uses the no-wheel model


@pytest.mark.parametrize("yaml_rel,ov", CASES)
def test_unmodified_reference_controller_configs_load(yaml_rel, ov):
    robot = yaml_rel.split("/")[0]
    P = iw.Problem(os.path.join(REF, yaml_rel), overrides=dict(ov, base_path=os.path.join(REF, robot)))
    Q = iw.Problem(yaml_rel, overrides=ov)
    assert P.qp_shape() == Q.qp_shape() and (P.nq, P.nv, P.na, P.ncontacts, P.nref) == (Q.nq, Q.nv, Q.na, Q.ncontacts, Q.nref)
    assert np.array_equal(P.default_refs(), Q.default_refs()) and np.array_equal(P.q0(), Q.q0())
    assert P.names("tasks") == Q.names("tasks") and P.names("contacts") == Q.names("contacts")
    sp, sq = P.struct(), Q.struct()
    for k in range(sp.ntasks):
        assert sp.tasks[k].weight == sq.tasks[k].weight and sp.tasks[k].type == sq.tasks[k].type


def test_unmodified_reference_behaviour_configs_load():
    P = iw.Problem("talos/talos_pos_tracker.yaml")
    ref = iw.Behavior(P, os.path.join(REF, "talos", "walk.yaml"))
    own = iw.Behavior(iw.Problem("talos/talos_pos_tracker.yaml"), "talos/walk.yaml")
    assert (ref.total_ticks, ref.nsegments, ref.masks) == (own.total_ticks, own.nsegments, own.masks) == (64000, 64, own.masks)
    for k in (0, 1, 2, 33, 63):
        a, b = ref.segment(k), own.segment(k)
        assert a[:4] == b[:4] and np.array_equal(a[4], b[4])
    F = iw.Problem("franka/pos_tracker.yaml")
    assert iw.Behavior(F, os.path.join(REF, "franka", "cartesian_line.yaml")).total_ticks == 4000
    assert iw.Behavior(P, os.path.join(REF, "talos", "move_com.yaml")).nsegments == 4
