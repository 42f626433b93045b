#!/usr/bin/env python3
"""Steady-state phase profile of the fused cycle kernel (mean SM cycles per instance and phase) on a bench workload.
   python tools/phase_profile.py [workload] [batch]"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
import inria_wbc_b200 as iw  # noqa: E402

wl = sys.argv[1] if len(sys.argv) > 1 else bench.DEFAULT_WORKLOAD
yaml, ov, B, recipe = bench.WORKLOADS[wl]
if len(sys.argv) > 2:
    B = int(sys.argv[2])
P = iw.Problem(yaml, overrides=ov)
q, v, refs = bench.make_inputs(P, B, recipe, 0xB200 + 3)
bt = iw.Batch(P, 0, B)
bt.update(q, v, refs)
bt.phase_profile(True)
bt.enable_timing(True)
for _ in range(3):
    r = bt.update(q, v, refs)
prof = bt.phase_profile(False)
ms = bt.kernel_ms(True)
out = {"workload": wl, "batch": B, "launch": bt.launch_info(), "fast": bt.fast_info(), "kernel_ms": ms, "mean_cycles_per_instance": prof,
       "total_cycles": {k: sum(x for kk, x in d.items() if kk not in ("instances", "directions", "adds") and isinstance(x, float))
                        for k, d in (prof or {}).items()},
       "mean_iters": float(r.iters.mean())}
print(json.dumps(out))
