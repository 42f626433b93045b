#!/usr/bin/env python3
"""Measured worst-case parity per configuration (GPU through the C ABI vs the CPU oracle): max relative error of tau / dv / f,
and the fraction of instances whose final active set is identical / an equivalent description of the same vertex (differs only
in rows with s_i = 0 at the common solution) / a tie (eiquadprog's strict-< argmin decided by rounding inside its own termination
slack).  Writes one JSON object per configuration; `python tools/parity_report.py > profiles/r02_parity_report.jsonl` on a B200."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import bench  # noqa: E402
import inria_wbc_b200 as iw  # noqa: E402
import oracle_lib as ol  # noqa: E402
from conftest import random_states  # noqa: E402
from test_cpu_sim import active_sets_equivalent  # noqa: E402

OFF = {"stabilizer": {"activated": False}, "collision_detection": {"activated": False}}


def rel(a, b):
    if b.size == 0:
        return 0.0
    return float((np.abs(a - b).max(axis=1) / np.maximum(np.abs(b).max(axis=1), 1e-9)).max())


def report(name, P, q, v, refs=None, mask=None):
    B = len(q)
    bt = iw.Batch(P, 0, B)
    r = bt.update(q, v, refs, contact_mask=mask)
    o = ol.cycle_batch(P, q, v, refs, contact_mask=mask, nthreads=os.cpu_count())
    ok = (o["status"] == 0) & (r.status == 0)
    same = ok & np.all(r.active == o["active"], axis=1)
    diff = np.where(ok & ~same)[0]
    eq = []
    for i in diff:
        rr = None if refs is None else (refs if refs.ndim == 1 else refs[i])
        m = None if mask is None else int(np.broadcast_to(mask, (B,))[i])
        eq.append(active_sets_equivalent(P, q[i], v[i], rr, o["active"][i], r.active[i], m))
    eq = np.array(eq, dtype=bool)
    equiv, tie = diff[eq] if len(diff) else diff, diff[~eq] if len(diff) else diff
    out = {"config": name, "B": B, "qp": P.qp_shape(), "status_equal": bool(np.array_equal(r.status, o["status"])),
           "optimal_frac": float(ok.mean()), "identical_active_set_frac": float(same.sum() / max(1, ok.sum())),
           "equivalent_vertex_frac": float(len(equiv) / max(1, ok.sum())), "tie_frac": float(len(tie) / max(1, ok.sum())),
           "iters_mean": float(o["iters"][ok].mean()), "iters_max": int(o["iters"][ok].max()),
           "iters_equal_frac_same_set": float((r.iters[same] == o["iters"][same]).mean()) if same.any() else None}
    for grp, idx in (("same_set", np.where(same)[0]), ("equivalent", equiv), ("tie", tie)):
        out["max_rel_" + grp] = {k: rel(getattr(r, k)[idx], o[k][idx]) for k in ("tau", "dv", "f", "q_next", "v_next")} if len(idx) else None
    if len(tie):
        out["tie_max_abs_f"] = float(np.abs(r.f[tie] - o["f"][tie]).max())
    print(json.dumps(out), flush=True)
    return out


def main():
    franka = iw.Problem("franka/pos_tracker.yaml")
    tiago = iw.Problem("tiago/pos_tracker.yaml", overrides={"urdf": "tiago/tiago_steel_no_wheel.urdf"})
    icub = iw.Problem("icub/humanoid_pos_tracker.yaml", overrides=OFF)
    talos = iw.Problem("talos/talos_pos_tracker.yaml", overrides=OFF)
    fast = iw.Problem("talos/talos_pos_tracker.yaml", overrides=dict(OFF, urdf="talos/talos_fast.urdf"))
    q, v, refs = bench.make_inputs(franka, 4096, "franka", 0xB200)
    report("franka_config2", franka, q, v, refs)
    q, v, refs = bench.make_inputs(tiago, 4096, "arm", 0xB200 + 5)
    report("tiago_config4_recipe", tiago, q, v, refs)
    q, v = random_states(tiago, 2048, 2, qn=0.2, vn=0.3)
    report("tiago_active_bounds", tiago, q, v)
    q, v, refs = bench.make_inputs(icub, 2048, "talos", 0xB200 + 4)
    report("icub_config4_recipe", icub, q, v, refs)
    q, v = random_states(icub, 512, 9, qn=0.1, vn=0.6)
    report("icub_hard", icub, q, v)
    q, v, refs = bench.make_inputs(talos, 4096, "talos", 0xB200 + 3)
    report("talos_config3", talos, q, v, refs)
    q, v, refs = bench.make_inputs(fast, 2048, "talos", 0xB200 + 3)
    report("talos_fast_config3", fast, q, v, refs)
    mask = np.repeat(np.array([3, 1, 2], dtype=np.uint32), 512)
    report("talos_fast_ds_ss_masks", fast, q[:1536], v[:1536], refs[:1536], mask)
    q, v = random_states(talos, 1536, 77, base=True)
    report("talos_moving_base_ds", talos, q, v)
    report("talos_moving_base_ss", talos, q, v, None, np.repeat(np.array([1, 2], dtype=np.uint32), 768))
    q, v = random_states(talos, 768, 5, qn=0.08, vn=0.4)
    report("talos_hard_ds_ss", talos, q, v, None, np.repeat(np.array([3, 1, 2], dtype=np.uint32), 256))


if __name__ == "__main__":
    main()
