"""send_cmd latch, q_solver and qp_step_back (second half; src/controllers/controller.cpp:238-242, 263-282, 309-312,
445-450) on the device against the line-by-line restatement in oracle/safety_oracle.py: Talos with the left-hand reference pulled
into the torso; some robots collide (and latch) after ~500 ticks, others never do."""
import os
import sys

import numpy as np
import pytest

import inria_wbc_b200 as iw
import oracle_lib as ol

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import safety_oracle as so  # noqa: E402

OFFSETS = [[0.0, -0.35, 0.0], [-0.15, -0.3, 0.1], [0.0, -0.2, -0.1], [0.0, 0.0, 0.1], [-0.05, -0.4, 0.05]]


def test_restated_latch_logic(talos):
    """CPU: the restatement itself -- the latch needs a collision AND a non-zero command; after it q_ stays, q_solver_ goes on."""
    cm = iw.CollisionModel(talos, "talos/collisions/talos_collisions_margin.yaml")
    pr, jm = cm.spheres()
    refs = talos.default_refs()
    lh, _ = talos.ref_slot("lh")
    refs[lh:lh + 3] += OFFSETS[4]
    c = so.SafeController(ol, talos, talos.q0(), np.zeros(talos.nv), (pr, jm, cm.nmembers))
    for t in range(560):
        c.solve(refs)
        if not c.send_cmd:
            break
    assert not c.send_cmd and c.is_model_colliding and 100 < t < 560
    q_latched = c.q_.copy()
    for _ in range(5):
        c.solve(refs)
    assert np.array_equal(c.q_, q_latched) and np.abs(c.q_solver - q_latched).max() > 1e-5
    qp, vp = c.q_prev.copy(), c.v_prev.copy()
    c.qp_step_back()
    assert np.array_equal(c.q_tsid, qp) and np.array_equal(c.v_tsid, vp)


@pytest.mark.gpu
def test_send_cmd_latch_and_step_back_on_device(talos):
    import torch
    P = talos
    cm = iw.CollisionModel(P, "talos/collisions/talos_collisions_margin.yaml")
    pr, jm = cm.spheres()
    B, T = len(OFFSETS), 560
    refs = np.tile(P.default_refs(), (B, 1))
    lh, _ = P.ref_slot("lh")
    refs[:, lh:lh + 3] += np.array(OFFSETS)
    q0, v0 = np.tile(P.q0(), (B, 1)), np.zeros((B, P.nv))
    dev = "cuda"
    dq, dv_, dr = (torch.from_numpy(x.copy()).to(dev) for x in (q0, v0, refs))
    f64 = dict(dtype=torch.float64, device=dev)
    S = {"send_cmd": torch.ones(B, dtype=torch.int32, device=dev), "colliding": torch.zeros(B, dtype=torch.int32, device=dev),
         "q_prev": torch.zeros(B, P.nq, **f64), "v_prev": torch.zeros(B, P.nv, **f64), "q_solver": torch.zeros(B, P.nv, **f64),
         "q_cmd": torch.zeros(B, P.nv, **f64), "tau_cmd": torch.zeros(B, P.na, **f64), "dq_cmd": torch.zeros(B, P.nv, **f64),
         "ddq_cmd": torch.zeros(B, P.nv, **f64)}
    ptrs = {k: t.data_ptr() for k, t in S.items()}
    st = torch.full((B,), -7, dtype=torch.int32, device=dev)
    bt = iw.Batch(P, 0, B)
    bt.safe_rollout_device(cm, B, dq.data_ptr(), dv_.data_ptr(), dr.data_ptr(), P.nref, P.default_contact_mask, T, ptrs, st.data_ptr())
    torch.cuda.synchronize()
    assert (st.cpu().numpy() == 0).all()
    ctrl = []
    for i in range(B):
        c = so.SafeController(ol, P, q0[i], v0[i], (pr, jm, cm.nmembers))
        for t in range(T):
            c.solve(refs[i])
        ctrl.append(c)
    got = {k: t.cpu().numpy() for k, t in S.items()}
    assert np.array_equal(got["send_cmd"], [int(c.send_cmd) for c in ctrl]) and 0 < got["send_cmd"].sum() < B  # both outcomes present
    assert np.array_equal(got["colliding"], [int(c.is_model_colliding) for c in ctrl])
    for i, c in enumerate(ctrl):
        for name, want in (("q_cmd", c.q_), ("q_solver", c.q_solver), ("dq_cmd", c.dq_), ("ddq_cmd", c.ddq_), ("q_prev", c.q_prev), ("v_prev", c.v_prev)):
            assert np.abs(got[name][i] - want).max() < 1e-8 * max(1.0, np.abs(want).max()), (i, name)
        assert np.abs(got["tau_cmd"][i] - c.tau_).max() < 1e-7 * np.abs(c.tau_).max(), i
        assert np.abs(dq[i].cpu().numpy() - c.q_tsid).max() < 1e-8, i
        if not c.send_cmd:  # the command stopped where the collision was detected, the solver's own state went on
            assert np.abs(got["q_cmd"][i] - got["q_solver"][i]).max() > 1e-4
    # qp_step_back() on the latched robots only, then set_send_cmd(true) and one more tick
    sel = (1 - S["send_cmd"]).contiguous()
    before = dq.cpu().numpy().copy()
    bt.qp_step_back_device(B, dq.data_ptr(), dv_.data_ptr(), ptrs, sel.data_ptr())
    torch.cuda.synchronize()
    after = dq.cpu().numpy()
    for i, c in enumerate(ctrl):
        if c.send_cmd:
            assert np.array_equal(after[i], before[i])
        else:
            c.qp_step_back()
            assert np.array_equal(after[i], got["q_prev"][i]) and np.abs(after[i] - c.q_tsid).max() < 1e-8
            c.send_cmd = True
    S["send_cmd"].fill_(1)
    bt.safe_rollout_device(cm, B, dq.data_ptr(), dv_.data_ptr(), dr.data_ptr(), P.nref, P.default_contact_mask, 1, ptrs, st.data_ptr())
    torch.cuda.synchronize()
    for i, c in enumerate(ctrl):
        c.solve(refs[i])
        assert np.abs(S["q_cmd"][i].cpu().numpy() - c.q_).max() < 1e-8 and int(S["send_cmd"][i]) == int(c.send_cmd), i
