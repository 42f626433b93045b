#!/usr/bin/env python3
"""Quick GPU-vs-oracle check used during development (the real tests are tests/test_gpu_*.py)."""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import inria_wbc_b200 as iw  # noqa: E402
import oracle_lib as ol  # noqa: E402


def rel(a, b):
    return np.abs(a - b).max() / max(1e-30, np.abs(b).max())


def states(P, B, seed, qn=0.02, vn=0.05):
    rng = np.random.default_rng(seed)
    q = np.tile(P.q0(), (B, 1))
    v = np.zeros((B, P.nv))
    na = P.na
    q[:, P.nq - na:] += rng.normal(0, qn, (B, na))
    v[:, P.nv - na:] += rng.normal(0, vn, (B, na))
    return q, v


def check(name, yaml, B, overrides=None, **kw):
    P = iw.Problem(yaml, overrides=overrides)
    q, v = states(P, B, 1234, **kw)
    bt = iw.Batch(P, 0, B)
    d = bt.debug_instance(q[0], v[0])
    o = ol.cycle_debug(P, q[0], v[0])
    print(name, "launch", bt.launch_info(), "cycles", bt.debug_cycles())
    print(name, "M", rel(d["M"], o["M"]), "nle", rel(d["nle"], o["nle"]), "H", rel(d["H"], o["H"]), "g", rel(d["g"], o["g"]),
          "x", rel(d["x"], o["x"]))
    t0 = time.time()
    r = bt.update(q, v)
    t1 = time.time()
    ref = ol.cycle_batch(P, q, v, nthreads=os.cpu_count())
    t2 = time.time()
    print(name, "B", B, "gpu s", t1 - t0, "cpu s", t2 - t1, "status gpu", np.bincount(r.status + 1), "cpu", np.bincount(ref["status"] + 1),
          "iters", r.iters.mean(), ref["iters"].mean())
    ok = (ref["status"] == 0) & (r.status == 0)
    print("   status equal:", np.array_equal(r.status, ref["status"]))
    for k in ("tau", "dv", "f", "q_next", "v_next", "q_cmd", "momentum"):
        a, b = getattr(r, k)[ok], ref[k][ok]
        if b.size:
            err = np.abs(a - b).max(axis=1) / np.maximum(1e-12, np.abs(b).max(axis=1))
            print("   %-8s max rel %.3e  (abs %.3e)" % (k, err.max(), np.abs(a - b).max()))
    print("   active equal (optimal):", np.array_equal(r.active[ok], ref["active"][ok]), "iters equal:", np.array_equal(r.iters[ok], ref["iters"][ok]),
          "max iters", r.iters[ok].max())
    return P, bt


if __name__ == "__main__":
    check("franka", "franka/pos_tracker.yaml", 512, qn=0.3, vn=0.3)
    check("tiago", "tiago/pos_tracker.yaml", 512, overrides={"urdf": "tiago/tiago_steel_no_wheel.urdf"}, qn=0.2, vn=0.3)
    check("talos", "talos/talos_pos_tracker.yaml", 512)
    check("talos-hard", "talos/talos_pos_tracker.yaml", 512, qn=0.15, vn=0.8)
