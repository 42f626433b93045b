#!/usr/bin/env python3
"""e2e timing of Batch.update (pinned host buffers) for a few pipeline chunk plans (IWBC_PIPE): tuning aid."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
import bench, inria_wbc_b200 as iw
wl = sys.argv[1] if len(sys.argv) > 1 else "talos_ds_16384"
yaml, ov, B, recipe = bench.WORKLOADS[wl]
P = iw.Problem(yaml, overrides=ov)
q, v, refs = bench.make_inputs(P, B, recipe, 0xB200 + 3)
bt = iw.Batch(P, 0, B)
hq, hv, hr = (torch.from_numpy(x).pin_memory() for x in (q, v, refs))
res = iw.Results(P, B, fields=("tau", "dv", "f", "q_next", "v_next", "status"))
for fld in ("tau", "dv", "f", "q_next", "v_next", "status"):
    setattr(res, fld, torch.from_numpy(getattr(res, fld)).pin_memory().numpy())
for plan in ("4,10", "2,5", "1,3", "2,9", "9,9", "18,18", "40,40", "1,1"):
    os.environ["IWBC_PIPE"] = plan
    for _ in range(3):
        bt.update(hq.numpy(), hv.numpy(), hr.numpy(), out=res)
    t0 = time.perf_counter()
    for _ in range(20):
        bt.update(hq.numpy(), hv.numpy(), hr.numpy(), out=res)
    dt = (time.perf_counter() - t0) / 20
    print(plan, round(dt * 1e3, 3), "ms", round(B / dt / 1e6, 3), "M/s")
