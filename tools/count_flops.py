#!/usr/bin/env python3
"""Counts the floating-point operations of one control update by running the INSTRUMENTED CPU oracle (oracle/count_scalar.hpp,
`make -C oracle count`) on the bench recipes, per cycle phase ("F_dyn, F_asm to be counted by instrumenting the
oracle").  Writes profiles/flops_per_update.json, which bench.py reads for the roofline numerator (F_dyn and F_asm measured;
F_qp stays the reference-algorithm formula of SURVEY.md 8(d) so that a smarter kernel cannot inflate its own fraction).
    python tools/count_flops.py            # CPU only; needs the product library for the host-side problem compiler"""
import ctypes as C
import json
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import bench  # noqa: E402
import inria_wbc_b200 as iw  # noqa: E402

subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle"), "-s", "count"])
L = C.CDLL(os.path.join(ROOT, "oracle", "liboracle_count.so"))
PH = ("other", "dynamics", "assembly", "qp", "decode")
OPS = ("add", "mul", "div", "func")
N = 256
out = {"_how": "instrumented oracle (every double counts its +,-,*,/ and sqrt/trig/exp; add and mul = 1 FLOP each, FMA = 2), mean over %d "
               "instances of the bench recipe; 'qp' = H, g assembly + eiquadprog as the restatement executes it (it skips structural zeros of the "
               "cost rows); the roofline numerator uses dynamics + assembly from here and the reference-algorithm F_qp formula" % N}
for wl, (yaml, ov, B, recipe) in bench.WORKLOADS.items():
    if recipe == "walk":
        continue  # needs the GPU to walk the seed instance; the stand-in recipe covers the same shapes
    P = iw.Problem(yaml, overrides=ov)
    q, v, refs = bench.make_inputs(P, N, recipe, 0xB200 + 3)
    masks = None
    if recipe == "walk_standin":
        P.set_task_weight("momentum", 0.0)
        masks = bench.make_walk_masks(P, N, refs, 0xB200 + 5)
    res = {k: np.zeros(s) for k, s in (("tau", (N, P.na)), ("dv", (N, P.nv)), ("f", (N, 12 * max(P.ncontacts, 1))), ("qn", (N, P.nq)), ("vn", (N, P.nv)),
                                        ("qc", (N, P.nv)), ("mom", (N, 3)))}
    status, iters = np.zeros(N, np.int32), np.zeros(N, np.int32)
    active = np.zeros((N, 12), np.uint32)
    cnt = (C.c_long * 20)()
    L.oracle_flop_counts(cnt, 1)
    p = lambda a: None if a is None else C.c_void_p(a.ctypes.data)  # noqa: E731
    rc = L.oracle_cycle_batch(C.c_void_p(P.data_ptr), C.c_int(N), p(q), p(v), p(refs), C.c_long(P.nref), p(masks), p(res["tau"]), p(res["dv"]),
                              p(res["f"]), p(res["qn"]), p(res["vn"]), p(status), p(iters), p(active), p(res["qc"]), p(res["mom"]), C.c_int(1))
    assert rc == 0 and L.oracle_flop_counts(cnt, 1) == 0
    # the instrumented build must walk the same path as the plain oracle
    import oracle_lib as ol
    ref = ol.cycle_batch(P, q, v, refs, contact_mask=masks)
    assert np.array_equal(ref["status"], status) and np.array_equal(ref["iters"], iters) and np.array_equal(ref["tau"], res["tau"]), wl
    c = np.array(list(cnt), dtype=np.float64).reshape(5, 4) / N
    d = {ph: {"flops": float(c[i].sum()), **{o: float(c[i, k]) for k, o in enumerate(OPS)}} for i, ph in enumerate(PH)}
    d["mean_iterations"] = float(iters.mean())
    d["f_dyn"] = d["dynamics"]["flops"]
    d["f_asm"] = d["assembly"]["flops"]
    d["f_qp_executed_by_oracle"] = d["qp"]["flops"]
    out[wl] = d
    print(wl, "F_dyn %.0f  F_asm %.0f  F_qp(oracle) %.0f  decode %.0f  iters %.2f" % (d["f_dyn"], d["f_asm"], d["qp"]["flops"], d["decode"]["flops"], d["mean_iterations"]))
out["talos_walk_131072"] = dict(out["talos_walk_standin_131072"], note="counted on the CPU-only stand-in recipe of the same shapes")
json.dump(out, open(os.path.join(ROOT, "profiles", "flops_per_update.json"), "w"), indent=1)
