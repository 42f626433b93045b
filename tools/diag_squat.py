#!/usr/bin/env python3
"""Diagnostic (GPU box): where does the device roll-out of etc/talos/squat.yaml first leave the oracle's, and what do both say
about the QP at that tick?"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import inria_wbc_b200 as iw  # noqa: E402
import oracle_lib as ol  # noqa: E402
from inria_wbc_b200 import behaviors, ETC_PATH  # noqa: E402

urdf = sys.argv[1] if len(sys.argv) > 1 else "talos/talos.urdf"
T = 36000
ov = {"closed_loop": False, "stabilizer": {"activated": False}, "collision_detection": {"activated": False}, "urdf": urdf}
P = iw.Problem("talos/talos_pos_tracker.yaml", overrides=ov)
beh = iw.Behavior(P, "talos/squat.yaml")
host = behaviors.MoveCom(P, os.path.join(ETC_PATH, "talos", "squat.yaml"))
period = beh.total_ticks
tab = host.table(period)
bt = iw.Batch(P, 0, 1)
q, v = P.q0()[None].copy(), np.zeros((1, P.nv))
trace = np.zeros((T, 1, P.nv))
out = bt.behavior_run(beh, q, v, T, trace=trace, fields=("status",))
print("gpu final status", out.status)
qo, vo = P.q0(), np.zeros(P.nv)
cm = []
for k in range(T // period):
    r = ol.rollout(P, qo, vo, tab, period)
    cm.append(r["q_cmd"] if r["rc"] == 0 else r["q_cmd"][:r["rc"] - 1])
    if r["rc"]:
        print("oracle fails at tick", k * period + r["rc"] - 1)
        break
    qo, vo = r["q"], r["v"]
cm = np.concatenate(cm)
n = min(len(cm), T)
err = np.abs(trace[:n, 0] - cm[:n]).max(axis=1)
bad = np.where(err > 1e-8)[0]
print("compared", n, "first tick with |dq| > 1e-8:", bad[:5], "err there", err[bad[:5]], "max err before", err[:bad[0]].max() if len(bad) else err.max())
if len(bad):
    td = int(bad[0])
    # replay on the GPU to td - 3, then tick by tick with status / iters
    bt2 = iw.Batch(P, 0, 1)
    beh2 = iw.Behavior(P, "talos/squat.yaml")
    q, v = P.q0()[None].copy(), np.zeros((1, P.nv))
    bt2.behavior_run(beh2, q, v, td - 3, fields=("status",))
    for t in range(td - 3, td + 3):
        qb, vb = q.copy(), v.copy()
        o = bt2.behavior_run(beh2, q, v, 1, tick0=t, fields=("status", "iters"))
        ref = tab[t % period]
        oc = ol.cycle_batch(P, qb, vb, ref)
        dof = P.names("dofs")
        S = P.struct()
        lo, hi = np.array(S.q_lo[:P.na]), np.array(S.q_hi[:P.na])
        near = [(dof[i], float(qb[0, 7 + i] - lo[i]), float(hi[i] - qb[0, 7 + i])) for i in range(P.na) if min(qb[0, 7 + i] - lo[i], hi[i] - qb[0, 7 + i]) < 1e-3]
        print("tick", t, "gpu status", o.status, "iters", o.iters, "| oracle on the GPU's state: status", oc["status"], "iters", oc["iters"], "near limits", near)
