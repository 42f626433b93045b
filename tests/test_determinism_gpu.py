"""tests/test_determinism.cpp of the reference (/root/reference/tests/test_determinism.cpp:24-138) at its own scale, GPU against
the CPU oracle: {arm.yaml, squat.yaml} x {talos.urdf, talos_fast.urdf}, 50 000 open-loop ticks with the behaviour evaluated on the
device every tick; the per-tick command vector q(false) of the CUDA path against the oracle driven by the host table of the same
behaviour, 1e-8 absolute (the reference's own tolerance on q, :51).  The GPU<->GPU half (5 fresh controllers, pairwise) is
tests/cpp/test_batch_determinism.cpp (tests/test_cpp_shim.py)."""
import os
from concurrent.futures import ThreadPoolExecutor

import numpy as np
import pytest

import inria_wbc_b200 as iw
import oracle_lib as ol
from inria_wbc_b200 import behaviors, ETC_PATH

pytestmark = pytest.mark.gpu
TICKS = 50000
CASES = [(b, u) for b in ("arm", "squat") for u in ("talos/talos.urdf", "talos/talos_fast.urdf")]


def _fresh(urdf):
    # stabilizer / collision detection / closed loop forced off as tests/test_determinism.cpp:66-70 does
    return iw.Problem("talos/talos_pos_tracker.yaml", overrides={"closed_loop": False, "stabilizer": {"activated": False},
                                                                  "collision_detection": {"activated": False}, "urdf": urdf})


def _oracle_run(case):
    """The CPU restatement over the same ticks; stops at the first QP that is not solved to optimality (the reference throws there)."""
    beh, urdf = case
    P = _fresh(urdf)
    cfg = os.path.join(ETC_PATH, "talos", beh + ".yaml")
    host = behaviors.Cartesian(P, cfg) if beh == "arm" else behaviors.MoveCom(P, cfg)
    period = iw.Behavior(P, "talos/%s.yaml" % beh).total_ticks
    tab = host.table(period)  # loop: true -> the table repeats
    q, v = P.q0(), np.zeros(P.nv)
    cmds, failed_at = [], None
    for k in range((TICKS + period - 1) // period):
        n = min(period, TICKS - k * period)
        r = ol.rollout(P, q, v, tab[:n], n)
        if r["rc"] != 0:  # rc = 1-based tick of the first non-optimal QP inside this chunk
            failed_at = k * period + r["rc"] - 1
            cmds.append(r["q_cmd"][:r["rc"] - 1])
            break
        q, v = r["q"], r["v"]
        cmds.append(r["q_cmd"])
    return np.concatenate(cmds), failed_at


def test_gpu_rollouts_match_the_oracle_over_50000_ticks():
    with ThreadPoolExecutor(len(CASES)) as ex:
        futures = [ex.submit(_oracle_run, c) for c in CASES]  # (ctypes releases the GIL: the four CPU roll-outs run beside the GPU)
        gpu = []
        for beh, urdf in CASES:
            P = _fresh(urdf)
            b = iw.Behavior(P, "talos/%s.yaml" % beh)
            bt = iw.Batch(P, 0, 1)
            q, v = P.q0()[None].copy(), np.zeros((1, P.nv))
            trace = np.zeros((TICKS, 1, P.nv))
            out = bt.behavior_run(b, q, v, TICKS, trace=trace, fields=("status",))
            gpu.append((trace[:, 0], int(out.status[0])))
        for (beh, urdf), f, (trace, status) in zip(CASES, futures, gpu):
            cmds, failed_at = f.result()
            n = min(len(cmds), TICKS)
            err_t = np.abs(trace[:n] - cmds[:n]).max(axis=1)
            if failed_at is None and status == 0:
                assert n == TICKS and err_t.max() < 1e-8, (beh, urdf, err_t.max())
                note = ""
            else:
                # The stand-in Talos drifts, over ~30 000 ticks of squats, until torso_2_joint sits EXACTLY on its lower limit; TSID's
                # position bound then returns 1e6 (its `q == qmin` branch) and the QP is infeasible: the reference throws there, the
                # oracle roll-out stops, the batched path reports the status and keeps the state.  Whether the limit is hit exactly
                # in one squat cycle or the next is decided by the last bit of the state, so the two sides may disagree on WHEN
                # (measured: 30 122 vs 34 024, and after a change of summation order in the kernel the other way round).  Required:
                # agreement to 1e-8 on every tick before the first of the two events, and at that tick the oracle, evaluated on the
                # GPU's own state, reports the same status after the same iterations as the GPU (non-optimal when the event is the
                # GPU's; optimal when it is the oracle's, whose own state differs in the last bit).
                assert status != 0 or failed_at is not None
                bad = np.where(err_t > 1e-8)[0]
                td = int(bad[0]) if len(bad) else n
                assert td > 20000 and err_t[:td].max() < 1e-8, (beh, urdf, td)
                if td < TICKS:
                    P = _fresh(urdf)
                    b2, bt2 = iw.Behavior(P, "talos/%s.yaml" % beh), iw.Batch(P, 0, 1)
                    q, v = P.q0()[None].copy(), np.zeros((1, P.nv))
                    bt2.behavior_run(b2, q, v, td, fields=("status",))
                    qb, vb = q.copy(), v.copy()
                    o = bt2.behavior_run(b2, q, v, 1, tick0=td, fields=("status", "iters"))
                    cfg = os.path.join(ETC_PATH, "talos", beh + ".yaml")
                    host = behaviors.Cartesian(P, cfg) if beh == "arm" else behaviors.MoveCom(P, cfg)
                    ref = host.table(b2.total_ticks)[td % b2.total_ticks]
                    oc = ol.cycle_batch(P, qb, vb, ref)
                    assert o.status[0] == oc["status"][0] and o.iters[0] == oc["iters"][0], (beh, urdf, td)
                    assert o.status[0] != 0 or failed_at == td, (beh, urdf, td, failed_at)
                note = " (first non-optimal QP: GPU tick %d, oracle tick %s -- the joint limit is met exactly, see the test)" % (td, failed_at)
                n = td
            err = err_t[:n].max()
            print("determinism %s %s: %d ticks GPU vs oracle, max |dq| %.3g%s" % (beh, urdf, n, err, note))
            assert np.abs(trace[n - 1] - trace[0]).max() > 1e-3  # the behaviour really moves the robot
