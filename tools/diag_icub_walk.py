#!/usr/bin/env python3
"""Growth of the GPU-vs-oracle difference along the iCub walking roll-out of tests/test_behaviors.py (the stand-in model reaches the
edge of feasibility around tick 930: how fast does the difference grow before it?)."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import inria_wbc_b200 as iw  # noqa: E402
import oracle_lib as ol  # noqa: E402
import behaviors_oracle as bo  # noqa: E402

OFF = {"stabilizer": {"activated": False}, "collision_detection": {"activated": False}}
P = iw.Problem("icub/humanoid_pos_tracker.yaml", overrides=OFF)
beh = iw.Behavior(P, "icub/walk.yaml", overrides={"num_of_cycles": 1})
ora = bo.Walk(P, os.path.join(iw.ETC_PATH, "icub", "walk.yaml"), num_of_cycles=1)
q, v = P.q0().copy(), np.zeros(P.nv)
dq = torch.from_numpy(P.q0()[None].copy()).cuda()
dv_ = torch.zeros(1, P.nv, dtype=torch.float64, device="cuda")
d_refs = torch.empty(1, P.nref, dtype=torch.float64, device="cuda")
st = torch.full((1,), -7, dtype=torch.int32, device="cuda")
bt = iw.Batch(P, 0, 4)
step = 50
for t0 in range(0, 950, step):
    for t in range(t0, t0 + step):
        r, m = ora.tick(t)
        o = ol.cycle_batch(P, q[None], v[None], r, contact_mask=m)
        if o["status"][0] != 0:
            print("oracle not optimal at", t)
            sys.exit(0)
        q, v = o["q_next"][0], o["v_next"][0]
    bt.behavior_rollout_device(beh, 1, dq.data_ptr(), dv_.data_ptr(), d_refs.data_ptr(), None, t0, step, init_refs=(t0 == 0), out_ptrs={"status": st.data_ptr()})
    torch.cuda.synchronize()
    print("tick %4d  status %d  max|dq| %.3g  max|dv| %.3g  iters(oracle last) %d" % (t0 + step, int(st[0]), np.abs(dq.cpu().numpy()[0] - q).max(),
                                                                                 np.abs(dv_.cpu().numpy()[0] - v).max(), o["iters"][0]))


# tick by tick from a fresh roll-out: the GPU's first non-optimal tick, and what the oracle says on the GPU's own state there
dq = torch.from_numpy(P.q0()[None].copy()).cuda()
dv_ = torch.zeros(1, P.nv, dtype=torch.float64, device="cuda")
it = torch.zeros(1, dtype=torch.int32, device="cuda")
bt2 = iw.Batch(P, 0, 4)
bt2.behavior_rollout_device(beh, 1, dq.data_ptr(), dv_.data_ptr(), d_refs.data_ptr(), None, 0, 600, out_ptrs={"status": st.data_ptr()})
torch.cuda.synchronize()
for t in range(600, 1100):
    pq, pv = dq.cpu().numpy().copy(), dv_.cpu().numpy().copy()
    bt2.behavior_rollout_device(beh, 1, dq.data_ptr(), dv_.data_ptr(), d_refs.data_ptr(), None, t, 1, init_refs=False,
                                out_ptrs={"status": st.data_ptr(), "iters": it.data_ptr()})
    torch.cuda.synchronize()
    r, m = ora.tick(t)
    o = ol.cycle_batch(P, pq, pv, d_refs.cpu().numpy(), contact_mask=m)
    if int(st[0]) != 0 or o["status"][0] != 0 or int(it[0]) != o["iters"][0]:
        print("tick", t, "mask", m, "GPU status/iters", int(st[0]), int(it[0]), " oracle on the GPU's state:", o["status"][0], o["iters"][0],
              " |dq_next|", np.abs(o["q_next"][0] - dq.cpu().numpy()[0]).max())
    if int(st[0]) != 0:
        break
