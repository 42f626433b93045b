#!/usr/bin/env python3
"""p50 / p95 latency of ONE batched control update as a function of the batch size (BASELINE.json's "p50 cycle latency per batch"):
host buffers in -> host results out through Batch.update (iwbc_batch_update), and the device-resident tick (rollout_device, one
tick per call + synchronise).  200 timed calls after 20 warm-ups per point.
    python tools/latency_sweep.py [workload]"""
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
import inria_wbc_b200 as iw  # noqa: E402

wl = sys.argv[1] if len(sys.argv) > 1 else "talos_ds_16384"
yaml, ov, _, recipe = bench.WORKLOADS[wl]
P = iw.Problem(yaml, overrides=ov)
out = {"workload": wl, "points": []}
for B in (1, 16, 256, 1024, 4096, 16384):
    q, v, refs = bench.make_inputs(P, B, recipe, 0xB200 + 3)
    bt = iw.Batch(P, 0, B)
    hq, hv, hr = (torch.from_numpy(x).pin_memory() for x in (q, v, refs))
    res = iw.Results(P, B, fields=("tau", "dv", "f", "q_next", "v_next", "status"))
    for fld in ("tau", "dv", "f", "q_next", "v_next", "status"):
        setattr(res, fld, torch.from_numpy(getattr(res, fld)).pin_memory().numpy())
    lat = []
    for k in range(220):
        t0 = time.perf_counter()
        bt.update(hq.numpy(), hv.numpy(), hr.numpy(), out=res)
        lat.append(time.perf_counter() - t0)
    lat = np.array(lat[20:]) * 1e3
    dq, dv_, dr = (torch.from_numpy(x).cuda() for x in (q.copy(), v.copy(), refs))
    st = torch.zeros(B, dtype=torch.int32, device="cuda")
    ts = torch.cuda.Stream()
    dl = []
    for k in range(220):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        bt.rollout_device(B, dq.data_ptr(), dv_.data_ptr(), dr.data_ptr(), P.nref, P.default_contact_mask, 1, {"status": st.data_ptr()}, ts.cuda_stream)
        ts.synchronize()
        dl.append(time.perf_counter() - t0)
    dl = np.array(dl[20:]) * 1e3
    pt = {"batch": B, "host_call_p50_ms": float(np.percentile(lat, 50)), "host_call_p95_ms": float(np.percentile(lat, 95)),
          "device_tick_p50_ms": float(np.percentile(dl, 50)), "device_tick_p95_ms": float(np.percentile(dl, 95))}
    out["points"].append(pt)
    print(pt, flush=True)
print(json.dumps(out))
