"""The *device source* of the fused cycle, compiled for the host with a one-thread group
(tests/cpu_sim.cpp, IWBC_CPU_SIM), against the oracle.  Pins the control flow and arithmetic of the
CUDA path in the CPU suite; the -m gpu tests repeat these comparisons on the real kernel."""
import numpy as np
import pytest

import oracle_lib as ol
from conftest import random_states
from test_oracle_qp import random_qp

TOL = 1e-8  # north_star: tau, dv, f within 1e-8 relative, same final active set


def active_sets_equivalent(P, q, v, ref_row, a_bits, b_bits, mask=None):
    """Equal, or differing only in *weakly active* rows (s_i = 0 at the common solution).  Contact6d's friction pyramid is
    degenerate whenever a corner force vanishes: its 4 rows meet at f_i = 0 in a 3-D space, so several active sets describe
    the same vertex and rounding picks among them (parity section)."""
    if np.array_equal(a_bits, b_bits):
        return True
    d = ol.cycle_debug(P, q, v, ref_row, contact_mask=mask)
    s = d["CI"] @ d["x"] + d["ci0"]
    scale = max(1.0, np.abs(d["ci0"][np.abs(d["ci0"]) < 1e9]).max())
    diff = [k for k in range(len(s)) if ((a_bits[k >> 5] ^ b_bits[k >> 5]) >> (k & 31)) & 1]
    return all(abs(s[k]) < 1e-7 * scale for k in diff)


def _compare(P, q, v, refs=None, mask=None, tol=TOL):
    ref = ol.cycle_batch(P, q, v, refs, contact_mask=mask, nthreads=4)
    sim = ol.sim_cycle_batch(P, q, v, refs, contact_mask=mask)
    assert np.array_equal(ref["status"], sim["status"])
    ok = ref["status"] == 0
    assert ok.mean() > 0.9
    for i in np.where(ok)[0]:
        rr = None if refs is None else (refs if refs.ndim == 1 else refs[i])
        assert active_sets_equivalent(P, q[i], v[i], rr, ref["active"][i], sim["active"][i], mask), i
    for k in ("tau", "dv", "f", "q_next", "v_next", "q_cmd", "momentum"):
        a, b = sim[k][ok], ref[k][ok]
        if b.size == 0:
            continue
        err = np.abs(a - b).max(axis=1) / np.maximum(np.abs(b).max(axis=1), 1e-9)
        assert err.max() < tol, (k, err.max())
    return ref


def test_franka(franka):
    q, v = random_states(franka, 64, 1, qn=0.3, vn=0.3)
    _compare(franka, q, v)


def test_tiago_with_active_bounds(tiago):
    q, v = random_states(tiago, 64, 2, qn=0.2, vn=0.3)
    ref = _compare(tiago, q, v)
    assert ref["iters"].max() > 3  # the inequality loop really runs


def test_talos_double_support(talos):
    q, v = random_states(talos, 32, 3)
    refs = np.tile(talos.default_refs(), (32, 1))
    off, _ = talos.ref_slot("com")
    refs[:, off + 2] -= np.random.default_rng(0).uniform(0, 0.2, 32)
    _compare(talos, q, v, refs)


def test_talos_moving_base_and_single_support(talos):
    q, v = random_states(talos, 16, 4, base=True)
    _compare(talos, q, v)
    _compare(talos, q, v, mask=0b01)
    _compare(talos, q, v, mask=0b10)


def test_talos_active_inequalities(talos):
    # velocities large enough to activate joint / torque bounds but still feasible
    q, v = random_states(talos, 48, 5, qn=0.05, vn=0.3)
    ref = ol.cycle_batch(talos, q, v, nthreads=4)
    sim = ol.sim_cycle_batch(talos, q, v)
    assert np.array_equal(ref["status"], sim["status"])
    ok = (ref["status"] == 0) & (ref["iters"] > 1)
    assert ok.sum() >= 5
    for i in np.where(ok)[0]:
        assert active_sets_equivalent(talos, q[i], v[i], None, ref["active"][i], sim["active"][i]), i
    err = np.abs(sim["dv"][ok] - ref["dv"][ok]).max(axis=1) / np.abs(ref["dv"][ok]).max(axis=1)
    assert err.max() < 1e-6  # conditioning of H (force regularisation ~1e-9) amplifies rounding on hard states: DESIGN.md


def _compare_fast(P, q, v, refs=None, mask=None, min_fast=0.5):
    """fast_kernel.cuh chained with the active-set kernel (as batch.cu launches them) against the oracle."""
    ref = ol.cycle_batch(P, q, v, refs, contact_mask=mask, nthreads=4)
    sim = ol.sim_fast_batch(P, q, v, refs, contact_mask=mask)
    assert np.array_equal(ref["status"], sim["status"])
    fast = sim["fast"] == 1
    assert fast.mean() >= min_fast
    # the fast path finishes exactly the instances eiquadprog ends after one iteration, with an empty active set
    assert np.array_equal(fast, (ref["iters"] == 1) & (ref["status"] == 0))
    assert not sim["active"][fast].any() and not ref["active"][fast].any()
    assert np.array_equal(ref["iters"][fast], sim["iters"][fast])
    ok = ref["status"] == 0
    for k in ("tau", "dv", "f", "q_next", "v_next", "q_cmd", "momentum"):
        a, b = sim[k][ok], ref[k][ok]
        err = np.abs(a - b).max(axis=1) / np.maximum(np.abs(b).max(axis=1), 1e-9)
        if fast.any():
            assert err[fast[ok]].max() < TOL, (k, err.max())
    return ref, sim


def test_fast_path_talos_double_support(talos):
    import bench
    q, v, refs = bench.make_inputs(talos, 96, "talos", 0xB200 + 3)
    _compare_fast(talos, q, v, refs, min_fast=0.8)


def test_fast_path_moving_base_and_single_support(talos):
    q, v = random_states(talos, 24, 4, base=True)
    _compare_fast(talos, q, v, min_fast=0.0)
    _compare_fast(talos, q, v, mask=0b01, min_fast=0.0)
    _compare_fast(talos, q, v, mask=0b10, min_fast=0.0)


def test_icub_double_support(icub):
    """iCub (config 4): n = 62, me = 18, mi = 132, 83 cost rows, contact frames with z pointing to the ground."""
    assert icub.qp_shape() == (62, 18, 132, 83)
    q, v = random_states(icub, 32, 6)
    _compare(icub, q, v)
    _compare_fast(icub, q, v, min_fast=0.5)
    q, v = random_states(icub, 24, 7, qn=0.1, vn=0.6)  # joint bounds become active
    ref = _compare(icub, q, v)
    assert ref["iters"].max() > 1
    _compare_fast(icub, q, v, min_fast=0.0)


@pytest.mark.parametrize("n,me,mi", [(9, 0, 4), (20, 5, 30), (74, 18, 244)])
def test_dense_qp(n, me, mi):
    rng = np.random.default_rng(n)
    for _ in range(3):
        H, g, CE, ce0, CI, ci0 = random_qp(rng, n, me, mi)
        a = ol.qp_solve(H, g, CE if me else None, ce0 if me else None, CI, ci0)
        b = ol.sim_qp_solve(H, g, CE if me else None, ce0 if me else None, CI, ci0)
        assert a["status"] == b["status"] == 0
        assert a["iters"] == b["iters"]
        bits = np.zeros(ol.ORACLE_ACTIVE_WORDS, np.uint32)
        for r in a["active"]:
            bits[r >> 5] |= np.uint32(1 << (r & 31))
        assert np.array_equal(bits, b["active"])
        np.testing.assert_allclose(b["x"], a["x"], rtol=0, atol=1e-9 * np.abs(a["x"]).max())


def test_shared_memory_footprints_keep_their_occupancy(talos, talos_fast, icub):
    """The fast pass runs 4 CTAs per SM and the active-set continuation 2: their dynamic shared memory must stay under
    (233 472 B per SM) / k - 1 kB.  Talos double support sits ~100 B under the continuation's limit; a layout change that crosses it
    halves that kernel's occupancy (measured: -20 % on the walking workload)."""
    import ctypes as C
    L = ol.sim_lib()
    for P in (talos, talos_fast, icub):
        for mask in (3, 1, 2):
            out = np.zeros(3, np.int64)
            assert L.sim_smem_bytes(C.c_void_p(P.data_ptr), C.c_uint(mask), C.c_void_p(out.ctypes.data)) == 0
            assert out[0] <= 233472 // 4 - 1024, (P.nv, mask, out)
            assert out[1] <= 233472 // 2 - 1024, (P.nv, mask, out)
