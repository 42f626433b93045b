"""Post-solve model collision check (src/safety/collision_check.cpp): the compiled sphere table, and the batched
device check against the CPU restatement."""
import os

import numpy as np
import pytest

import inria_wbc_b200 as iw
import oracle_lib as ol
from conftest import random_states


def test_collision_tables(talos, tiago, franka):
    cm = iw.CollisionModel(talos, "talos/collisions/talos_collisions_margin.yaml")
    assert cm.nmembers == 5 and cm.skipped_links == 0 and cm.nspheres > 100  # arm_left, arm_right, leg_left, leg_right, torso
    pr, jm = cm.spheres()
    assert (pr[:, 3] > 0).all() and (np.diff(jm[:, 1]) >= 0).all()  # grouped by member, members sorted by name
    assert jm[:, 0].min() >= 0  # every link of a floating-base robot hangs below root_joint
    t = iw.CollisionModel(tiago, "tiago/collisions.yaml")
    assert t.nmembers == 2 and t.nspheres == 16
    assert (t.spheres()[1][:, 0] == -1).sum() == 1  # base_link of the fixed-base Tiago rides on the universe
    with pytest.raises(iw.IwbcError, match="no sphere"):
        iw.CollisionModel(franka, "tiago/collisions.yaml")
    # the reference configuration is collision free, and the restatement finds folded arms
    assert ol.collision(talos, talos.q0(), pr, jm, cm.nmembers)[0] == 0
    q = talos.q0().copy()
    names = talos.names("dofs")
    q[7 + names.index("arm_left_2_joint")] = -0.5   # left arm swung into the torso / leg
    q[7 + names.index("arm_left_4_joint")] = -2.0
    assert ol.collision(talos, q, pr, jm, cm.nmembers)[0] == 1


@pytest.mark.gpu
def test_collision_check_on_device_matches_the_restatement(talos, tiago):
    import torch
    for P, yaml, qn in ((talos, "talos/collisions/talos_collisions_margin.yaml", 0.35), (tiago, "tiago/collisions.yaml", 0.8)):
        cm = iw.CollisionModel(P, yaml)
        pr, jm = cm.spheres()
        B = 512
        q, _ = random_states(P, B, 77, qn=qn, base=P.nv > P.na)
        q[0] = P.q0()
        bt = iw.Batch(P, 0, B)
        dq = torch.from_numpy(q).cuda()
        hit = torch.full((B,), -1, dtype=torch.int32, device="cuda")
        pair = torch.full((B, 4), -9, dtype=torch.int32, device="cuda")
        bt.check_collisions_device(cm, B, dq.data_ptr(), hit.data_ptr(), pair.data_ptr())
        torch.cuda.synchronize()
        hit, pair = hit.cpu().numpy(), pair.cpu().numpy()
        want = np.array([ol.collision(P, q[i], pr, jm, cm.nmembers) for i in range(B)])
        assert np.array_equal(hit, want[:, 0])
        assert np.array_equal(pair, want[:, 1:])
        assert 0.02 < hit.mean() < 0.98, hit.mean()  # the sample has both outcomes
        if P is talos:
            assert hit[0] == 0  # the reference configuration of the (stand-in) Talos model is free
        assert bt.launch_count == 1
