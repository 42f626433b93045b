"""The instrumented oracle (oracle/count_scalar.hpp): same path and same numbers as the plain oracle, and the
committed per-update FLOP counts that bench.py's roofline numerator reads are reproducible."""
import ctypes as C
import json
import os
import subprocess

import numpy as np

import oracle_lib as ol

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _count(P, q, v, refs):
    subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle"), "-s", "count"])
    L = C.CDLL(os.path.join(ROOT, "oracle", "liboracle_count.so"))
    N = q.shape[0]
    z = lambda *s: np.zeros(s)  # noqa: E731
    tau, dv, f, qn, vn, qc, mom = z(N, P.na), z(N, P.nv), z(N, 12 * max(P.ncontacts, 1)), z(N, P.nq), z(N, P.nv), z(N, P.nv), z(N, 3)
    status, iters, active = np.zeros(N, np.int32), np.zeros(N, np.int32), np.zeros((N, 12), np.uint32)
    cnt = (C.c_long * 20)()
    L.oracle_flop_counts(cnt, 1)
    p = lambda a: C.c_void_p(a.ctypes.data)  # noqa: E731
    rc = L.oracle_cycle_batch(C.c_void_p(P.data_ptr), C.c_int(N), p(q), p(v), p(refs), C.c_long(P.nref), None, p(tau), p(dv), p(f), p(qn), p(vn),
                              p(status), p(iters), p(active), p(qc), p(mom), C.c_int(1))
    assert rc == 0 and L.oracle_flop_counts(cnt, 0) == 0
    return np.array(list(cnt), dtype=np.float64).reshape(5, 4) / N, tau, status, iters


def test_instrumented_oracle_walks_the_same_path(talos, franka):
    import bench
    assert ol.lib().oracle_flop_counts(None, 0) == -1  # the plain build does not count
    committed = json.load(open(os.path.join(ROOT, "profiles", "flops_per_update.json")))
    for P, recipe, wl in ((franka, "franka", "franka_65536"), (talos, "talos", "talos_ds_16384")):
        q, v, refs = bench.make_inputs(P, 32, recipe, 0xB200 + 3)
        c, tau, status, iters = _count(P, q, v, refs)
        ref = ol.cycle_batch(P, q, v, refs)
        assert np.array_equal(ref["tau"], tau) and np.array_equal(ref["status"], status) and np.array_equal(ref["iters"], iters)
        f_dyn, f_asm = c[1].sum(), c[2].sum()
        assert f_dyn > 0 and f_asm > 0 and c[3].sum() > 0
        # dynamics and assembly do not depend on the state (fixed loop bounds): the committed figures are reproduced
        assert abs(f_dyn - committed[wl]["f_dyn"]) <= 0.01 * f_dyn and abs(f_asm - committed[wl]["f_asm"]) <= 0.02 * f_asm
