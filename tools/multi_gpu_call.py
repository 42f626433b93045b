#!/usr/bin/env python3
"""One host call over G GPUs (iwbc_multi_update) against G separate single-GPU calls: Talos double support, 16 384 states per GPU,
pinned host buffers.  python tools/multi_gpu_call.py [G]"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import bench  # noqa: E402
import inria_wbc_b200 as iw  # noqa: E402

G = int(sys.argv[1]) if len(sys.argv) > 1 else torch.cuda.device_count()
per = 16384
yaml, ov, _, recipe = bench.WORKLOADS["talos_ds_16384"]
P = iw.Problem(yaml, overrides=ov)
B = per * G
q, v, refs = bench.make_inputs(P, B, recipe, 0xB200 + 3)
pin = lambda a: torch.from_numpy(a).pin_memory().numpy()  # noqa: E731
q, v, refs = pin(q), pin(v), pin(refs)
res = iw.Results(P, B, fields=("tau", "dv", "f", "q_next", "v_next", "status"))
for f in ("tau", "dv", "f", "q_next", "v_next", "status"):
    setattr(res, f, pin(getattr(res, f)))
mb = iw.MultiBatch(P, list(range(G)), B)
for _ in range(3):
    mb.update(q, v, refs, out=res)
lat = []
for _ in range(20):
    t0 = time.perf_counter()
    mb.update(q, v, refs, out=res)
    lat.append(time.perf_counter() - t0)
multi = B / np.median(lat)
one = iw.Batch(P, 0, per)
r1 = iw.Results(P, per, fields=("tau", "dv", "f", "q_next", "v_next", "status"))
for f in ("tau", "dv", "f", "q_next", "v_next", "status"):
    setattr(r1, f, pin(getattr(r1, f)))
for _ in range(3):
    one.update(q[:per], v[:per], refs[:per], out=r1)
l1 = []
for _ in range(20):
    t0 = time.perf_counter()
    one.update(q[:per], v[:per], refs[:per], out=r1)
    l1.append(time.perf_counter() - t0)
single = per / np.median(l1)
same = all(np.array_equal(getattr(res, f)[:per], getattr(r1, f)) for f in ("tau", "dv", "f", "status"))
print(json.dumps({"gpus": G, "states": B, "one_call_updates_per_s": multi, "p50_ms": 1e3 * float(np.median(lat)), "single_gpu_updates_per_s": single,
                  "efficiency_vs_G_x_single": multi / (G * single), "first_shard_bit_identical_to_single_gpu_call": bool(same),
                  "optimal_fraction": float((res.status == 0).mean())}))
