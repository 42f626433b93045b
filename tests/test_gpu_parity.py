"""Parity tests proper: the CUDA path through the C ABI against the CPU oracle (run on a B200 with -m gpu).
Bar (BASELINE.json north_star): tau, dv, f within 1e-8 relative, same final active set, same status.

What is asserted, and why (measured table: profiles/r02_parity_report.jsonl, tools/parity_report.py):
  * tau, dv (and q_next, v_next, q_cmd, momentum): 1e-8 relative on EVERY configuration (measured worst case 5.2e-9 on the
    benchmark recipes);
  * f: 1e-8 on the Talos benchmark recipes (measured <= 5e-10); TOL_F_STRESS on iCub and on the stress states (moving base, large
    offsets: measured up to 1.5e-7).  The contact forces' components in the null space of the 6 x 12 force generator are fixed
    only by the 1e-8 Hessian regulariser: cond(H_ff) ~ 1e7, so two backward-stable implementations differ by ~1e7 x 1e-15 there
    (the oracle against itself with a 1e-16 perturbed problem: 3e-10, tests/test_oracle_loader.py);
  * final active set: identical, or an equivalent description of the same vertex (differing rows have |s_i| < 1e-7 x scale at the
    common solution: weakly active rows of a degenerate friction-pyramid vertex).  "Identical" alone is not a well-posed bar: the
    oracle against itself with a 1e-16 perturbed problem differs in 0.1 % (config 3) to 7 % (moving base) of the instances.
    Ties (eiquadprog's strict-< argmin decided by rounding between rows of equal value; < 0.2 % of a batch) must agree on the
    iteration count, on tau / dv to 1e-8 and on f within eiquadprog's own termination slack."""
import os
import sys

import numpy as np
import pytest

import inria_wbc_b200 as iw
import oracle_lib as ol
from conftest import random_states
from test_cpu_sim import active_sets_equivalent
from test_oracle_qp import random_qp

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
TOL = 1e-8
TOL_F_STRESS = 5e-7


def _rel(a, b):
    return (np.abs(a - b).max(axis=1) / np.maximum(np.abs(b).max(axis=1), 1e-9)).max() if b.size else 0.0


def _check(P, bt, q, v, refs=None, mask=None, tol=TOL, tol_f=None):
    """GPU (through the C ABI) against the oracle: see the module docstring for the bar."""
    tol_f = tol if tol_f is None else tol_f
    r = bt.update(q, v, refs, contact_mask=mask)
    o = ol.cycle_batch(P, q, v, refs, contact_mask=mask, nthreads=os.cpu_count())
    assert np.array_equal(r.status, o["status"])
    ok = o["status"] == 0
    assert ok.mean() > 0.9
    same = ok & np.all(r.active == o["active"], axis=1)
    for k in ("tau", "dv", "f", "q_next", "v_next", "q_cmd", "momentum"):
        assert _rel(getattr(r, k)[same], o[k][same]) < (tol_f if k == "f" else tol), k
    bad = np.where(ok & ~same)[0]
    ties = 0
    for i in bad[:200]:
        rr = None if refs is None else (refs if refs.ndim == 1 else refs[i])
        m = None if mask is None else int(np.broadcast_to(mask, (len(q),))[i])
        if active_sets_equivalent(P, q[i], v[i], rr, o["active"][i], r.active[i], m):
            for k in ("tau", "dv", "f"):
                assert _rel(getattr(r, k)[i:i + 1], o[k][i:i + 1]) < (tol_f if k == "f" else tol), (k, i)
        else:
            ties += 1
            assert r.iters[i] == o["iters"][i], i
            assert np.abs(r.f[i] - o["f"][i]).max() < 0.5, i  # eiquadprog's termination slack on these problems: ~0.1 - 0.3 N
            assert _rel(r.tau[i:i + 1], o["tau"][i:i + 1]) < tol and _rel(r.dv[i:i + 1], o["dv"][i:i + 1]) < tol, i
    assert ties <= max(2, 0.005 * ok.sum()), ties
    return r, o


def test_library_is_the_cuda_one():
    assert os.path.exists(iw.lib_path())
    assert iw.measure_fp64_tflops(0) > 5.0


def test_franka_batch(franka):
    import bench
    q, v, refs = bench.make_inputs(franka, 4096, "franka", 0xB200)
    bt = iw.Batch(franka, 0, 4096)
    r, o = _check(franka, bt, q, v, refs)
    assert (r.iters == 1).all() and bt.launch_count == 1


def test_tiago_active_bounds(tiago):
    q, v = random_states(tiago, 2048, 2, qn=0.2, vn=0.3)
    bt = iw.Batch(tiago, 0, 2048)
    r, o = _check(tiago, bt, q, v)
    assert np.array_equal(r.iters, o["iters"]) and r.iters.max() > 3


def test_mixed_icub_tiago_config4(icub, tiago):
    """BASELINE config 4: a 50/50 iCub + Tiago batch, bucketed by robot (one handle and one launch set per robot on the
    same device), each half against the oracle."""
    B = 1024
    qi, vi = random_states(icub, B, 0xB200 + 4)
    qt, vt = random_states(tiago, B, 0xB200 + 5, qn=0.2, vn=0.3)
    bi, btg = iw.Batch(icub, 0, B), iw.Batch(tiago, 0, B)
    ri, oi = _check(icub, bi, qi, vi, tol_f=TOL_F_STRESS)
    rt, ot = _check(tiago, btg, qt, vt)
    assert np.array_equal(ri.iters, oi["iters"]) and np.array_equal(rt.iters, ot["iters"])
    # interleaved calls on the two handles keep their own constants (the __constant__ copies are re-uploaded on a switch)
    ri2 = bi.update(qi, vi)
    assert np.array_equal(ri.tau, ri2.tau) and np.array_equal(ri.active, ri2.active)
    # harder iCub states: joint bounds active, the active-set continuation runs
    qh, vh = random_states(icub, 512, 9, qn=0.1, vn=0.6)
    rh, oh = _check(icub, bi, qh, vh, tol_f=TOL_F_STRESS)
    assert oh["iters"].max() > 1


def test_talos_double_support_config3(talos):
    import bench
    B = 2048
    q, v, refs = bench.make_inputs(talos, B, "talos", 0xB200 + 3)
    bt = iw.Batch(talos, 0, B)
    r, o = _check(talos, bt, q, v, refs)
    same = np.all(r.active == o["active"], axis=1)
    assert same.mean() > 0.97  # the rest differ only in weakly active friction rows (checked by _check)
    assert np.array_equal(r.iters[same], o["iters"][same])


def test_talos_fast_urdf_config3(talos_fast):
    """BASELINE config 3 on the fast URDF (n = 62, me = 18, mi = 196, 85 cost rows), double and single support."""
    import bench
    P = talos_fast
    assert (P.nq, P.nv, P.na) == (39, 38, 32) and P.qp_shape() == (62, 18, 196, 85)
    B = 1536
    q, v, refs = bench.make_inputs(P, B, "talos", 0xB200 + 3)
    bt = iw.Batch(P, 0, B)
    r, o = _check(P, bt, q, v, refs)
    mask = np.repeat(np.array([3, 1, 2], dtype=np.uint32), B // 3)
    _check(P, bt, q, v, refs, mask=mask)


def test_talos_hard_states_exercise_adds_drops_and_the_large_active_sets(talos):
    """Large joint offsets and velocities: tens of active-set iterations per instance (up to > 100), constraints dropped again
    (delete_constraint_cta), single-support shapes too.  About a fifth of these states end on a degenerate vertex (joints on their
    bounds with zero multiplier, unloaded contact corners): the two implementations may then name it by different weakly active
    rows -- checked to be the same point (active_sets_equivalent) with the same torques and accelerations."""
    B = 768
    q, v = random_states(talos, B, 5, qn=0.08, vn=0.4)
    bt = iw.Batch(talos, 0, B)
    mask = np.repeat(np.array([3, 1, 2], dtype=np.uint32), B // 3)
    r = bt.update(q, v, contact_mask=mask)
    o = ol.cycle_batch(talos, q, v, contact_mask=mask, nthreads=os.cpu_count())
    assert np.array_equal(r.status, o["status"])
    ok = o["status"] == 0
    assert ok.mean() > 0.9 and o["iters"][ok].max() > 40 and o["iters"][ok].mean() > 6
    same = ok & np.all(r.active == o["active"], axis=1)
    assert same.sum() > 0.7 * ok.sum()
    assert (r.iters[same] == o["iters"][same]).mean() > 0.9  # (a tie on the way may cost or save an add / drop pair and still end on the same set)
    for k in ("tau", "dv"):
        assert _rel(getattr(r, k)[same], o[k][same]) < 3e-8, k  # > 100 active-set iterations: measured 1.1e-8 (profiles/r02_parity_report.jsonl)
    assert _rel(r.f[same], o["f"][same]) < TOL_F_STRESS
    diff = np.where(ok & ~same)[0]
    equivalent = [active_sets_equivalent(talos, q[i], v[i], None, o["active"][i], r.active[i], int(mask[i])) for i in diff[:80]]
    assert np.mean(equivalent) > 0.95
    for k in ("tau", "dv"):
        assert _rel(getattr(r, k)[diff], o[k][diff]) < 3e-8, k


def test_moving_base_batches(talos, talos_fast, icub):
    """Free-flyer away from the SRDF pose with a non-zero base twist (position +-5 cm, orientation ~0.2 rad, twist 0.05), double and
    single support: the base columns of M / nle / the contact Jacobians and the SE(3) integration are exercised on the CUDA binary
    itself (tests/test_cpu_sim.py does the same on the host-compiled device source)."""
    for P, B in ((talos, 1536), (talos_fast, 1024), (icub, 1024)):
        q, v = random_states(P, B, 77, base=True)
        bt = iw.Batch(P, 0, B)
        r, o = _check(P, bt, q, v, tol_f=TOL_F_STRESS)
        assert np.abs(r.q_next[:, :7] - o["q_next"][:, :7]).max() < 1e-12  # SE(3) exponential of the base
        mask = np.repeat(np.array([3, 1, 2], dtype=np.uint32), (B + 2) // 3)[:B]
        _check(P, bt, q, v, mask=mask, tol_f=TOL_F_STRESS)


def test_talos_stage_dump(talos):
    q, v = random_states(talos, 1, 9, base=True)
    bt = iw.Batch(talos, 0, 8)
    d = bt.debug_instance(q[0], v[0])
    o = ol.cycle_debug(talos, q[0], v[0])
    for k, tol in (("M", 1e-12), ("nle", 1e-12), ("H", 1e-11), ("g", 1e-11), ("x", 1e-8)):
        assert np.abs(d[k] - o[k]).max() <= tol * np.abs(o[k]).max(), k


def test_talos_mixed_contact_masks_are_bucketed(talos):
    B = 600
    q, v = random_states(talos, B, 12)
    mask = np.random.default_rng(1).choice(np.array([3, 1, 2], dtype=np.uint32), B)
    bt = iw.Batch(talos, 0, B)
    r, o = _check(talos, bt, q, v, mask=mask)
    assert bt.launch_count == 6  # per contact mask: one shape-uniform launch of the fast pass + one of the active-set continuation
    f = r.f.reshape(B, 2, 12)
    assert np.all(f[mask == 1, 1] == 0) and np.all(f[mask == 2, 0] == 0)


def test_pipelined_host_call_with_contact_mask_runs(talos):
    """iwbc_batch_update's pipelined form (B >= 8192: upload / compute / download streams chained per chunk) on a batch bucketed into
    three contact-mask runs, against the plain single-stream form of the same call (smaller pieces) bit for bit, and a sample
    of instances against the oracle."""
    B = 9216
    q, v = random_states(talos, B, 33)
    mask = np.repeat(np.array([3, 1, 2], dtype=np.uint32), B // 3)
    bt = iw.Batch(talos, 0, B)
    big = bt.update(q, v, contact_mask=mask)
    for first in range(0, B, 3072):  # each piece is uniform and below the pipeline threshold
        sl = slice(first, first + 3072)
        small = bt.update(q[sl], v[sl], contact_mask=mask[sl])
        for k in ("tau", "dv", "f", "q_next", "v_next", "q_cmd", "momentum", "status", "iters", "active"):
            assert np.array_equal(getattr(big, k)[sl], getattr(small, k)), (k, first)
    pick = np.random.default_rng(2).choice(B, 384, replace=False)
    o = ol.cycle_batch(talos, q[pick], v[pick], contact_mask=mask[pick], nthreads=os.cpu_count())
    assert np.array_equal(big.status[pick], o["status"])
    same = np.all(big.active[pick] == o["active"], axis=1) & (o["status"] == 0)
    assert same.mean() > 0.9
    for k in ("tau", "dv", "f"):
        assert _rel(getattr(big, k)[pick][same], o[k][same]) < TOL, k
    # a second call on the same handle (constants already resident, different run lengths) stays consistent
    mask2 = np.concatenate([np.full(5000, 3, np.uint32), np.full(B - 5000, 2, np.uint32)])
    again = bt.update(q, v, contact_mask=mask2)
    ref = bt.update(q[:5000], v[:5000], contact_mask=mask2[:5000])
    assert np.array_equal(again.tau[:5000], ref.tau) and np.array_equal(again.status[:5000], ref.status)


def test_full_size_properties(talos):
    """BASELINE config 3 at full size (16 384): size-independent properties instead of the oracle."""
    import bench
    B = 16384
    q, v, refs = bench.make_inputs(talos, B, "talos", 0xB200 + 3)
    bt = iw.Batch(talos, 0, B)
    r = bt.update(q, v, refs)
    assert (r.status == 0).all()
    S = talos.struct()
    # semi-implicit Euler identity and joint-space integration
    np.testing.assert_allclose(r.v_next, v + S.dt * r.dv, rtol=0, atol=1e-15)
    np.testing.assert_allclose(r.q_next[:, 7:], q[:, 7:] + S.dt * r.v_next[:, 6:], rtol=0, atol=1e-15)
    assert np.abs(np.linalg.norm(r.q_next[:, 3:7], axis=1) - 1).max() < 1e-15
    # constraints hold: torque limits, friction cone, unilateral normal force
    tmax = np.array(S.tau_max[:talos.na])
    assert (np.abs(r.tau) <= tmax * (1 + 1e-9) + 1e-9).all()
    f = r.f.reshape(B, 2, 4, 3)
    fz = f[..., 2].sum(axis=2)
    assert (fz >= 5.0 - 0.2).all() and (fz <= 1500 + 0.2).all()
    # (eiquadprog stops when the summed violation is below mi * eps * trace(H) * trace(J) * 100: with the 1e-9 force
    # regularisation trace(J) ~ 2e5 and the accepted total violation is ~0.1 N; the oracle stops at the same point)
    # measured on this batch with the CPU oracle: 79 of 16 384 instances end with a friction-row violation > 0.2 N, max 0.318 N
    assert (np.abs(f[..., 0]) <= 0.3 * f[..., 2] + 0.5).all() and (np.abs(f[..., 1]) <= 0.3 * f[..., 2] + 0.5).all()
    # bitwise run-to-run determinism (the reference's test_determinism contract, on the GPU)
    r2 = bt.update(q, v, refs)
    for k in ("tau", "dv", "f", "q_next", "status", "iters", "active"):
        assert np.array_equal(getattr(r, k), getattr(r2, k)), k
    # a random subset against the oracle
    idx = np.random.default_rng(0).choice(B, 256, replace=False)
    o = ol.cycle_batch(talos, q[idx], v[idx], refs[idx], nthreads=os.cpu_count())
    for k in ("tau", "dv", "f"):
        assert _rel(getattr(r, k)[idx], o[k]) < TOL, k
    assert np.all(r.active[idx] == o["active"], axis=1).mean() > 0.97


def test_franka_full_size_linearity(franka):
    """Franka 65 536 (config 2): the stack is an unconstrained least squares, so dv is affine in the posture reference."""
    import bench
    B = 65536
    q, v, refs = bench.make_inputs(franka, B, "franka", 0xB200)
    bt = iw.Batch(franka, 0, B)
    r0 = bt.update(q, v, refs, fields=("dv", "status"))
    off, _ = franka.ref_slot("posture")
    d1, d2 = np.zeros_like(refs), np.zeros_like(refs)
    rng = np.random.default_rng(3)
    d1[:, off:off + 9] = rng.normal(0, 0.1, (B, 9))
    d2[:, off:off + 9] = rng.normal(0, 0.1, (B, 9))
    r1 = bt.update(q, v, refs + d1, fields=("dv", "status"))
    r2 = bt.update(q, v, refs + d2, fields=("dv", "status"))
    r12 = bt.update(q, v, refs + d1 + d2, fields=("dv", "status"))
    assert (r0.status == 0).all()
    lhs, rhs = r12.dv - r0.dv, (r1.dv - r0.dv) + (r2.dv - r0.dv)
    assert np.abs(lhs - rhs).max() < 1e-7 * np.abs(r0.dv).max()


def test_rollout_device_matches_oracle(talos):
    import torch
    from inria_wbc_b200 import behaviors, ETC_PATH
    B, T = 8, 40
    q, v = random_states(talos, B, 21, qn=0.01, vn=0.02)
    refs = talos.default_refs()
    dq, dv_, dr = (torch.from_numpy(x).cuda() for x in (q.copy(), v.copy(), refs))
    bt = iw.Batch(talos, 0, B)
    st = torch.empty(B, dtype=torch.int32, device="cuda")
    bt.rollout_device(B, dq.data_ptr(), dv_.data_ptr(), dr.data_ptr(), 0, talos.default_contact_mask, T, {"status": st.data_ptr()})
    torch.cuda.synchronize()
    assert (st.cpu().numpy() == 0).all() and bt.launch_count == 2 * T  # fast pass + active-set continuation per tick
    for i in range(B):
        o = ol.rollout(talos, q[i], v[i], refs, T)
        assert o["rc"] == 0
        assert np.abs(dq[i].cpu().numpy() - o["q"]).max() < 1e-8   # test_determinism's tolerance on q, GPU vs oracle
        assert np.abs(dv_[i].cpu().numpy() - o["v"]).max() < 1e-7


def test_dense_qp_solver_interface():
    rng = np.random.default_rng(7)
    for n, me, mi in ((9, 0, 6), (20, 5, 30), (74, 18, 244)):
        B = 64
        qs = [random_qp(rng, n, me, mi) for _ in range(B)]
        H, g, CE, ce0, CI, ci0 = (np.stack([p[k] for p in qs]) for k in range(6))
        x, status, iters, active = iw.qp_solve_batch(H, g, CE if me else None, ce0 if me else None, CI, ci0)
        for i in range(B):
            o = ol.qp_solve(*qs[i][:2], qs[i][2] if me else None, qs[i][3] if me else None, *qs[i][4:])
            assert status[i] == o["status"] == 0
            np.testing.assert_allclose(x[i], o["x"], rtol=0, atol=1e-9 * np.abs(o["x"]).max())
            bits = np.zeros(iw.ACTIVE_WORDS, np.uint32)
            for r in o["active"]:
                bits[r >> 5] |= np.uint32(1 << (r & 31))
            assert np.array_equal(bits, active[i]) and iters[i] == o["iters"]


def test_infeasible_instances_are_data_not_errors(talos):
    # absurd velocities make the bound constraints inconsistent: status 1 per instance, no exception, others unaffected
    q, v = random_states(talos, 64, 30)
    v[::2, 10] = 500.0
    bt = iw.Batch(talos, 0, 64)
    r = bt.update(q, v)
    o = ol.cycle_batch(talos, q, v, nthreads=8)
    assert np.array_equal(r.status, o["status"])
    assert (r.status[1::2] == 0).all() and (r.status[::2] != 0).any()


def test_argument_errors(franka):
    bt = iw.Batch(franka, 0, 16)
    q, v = random_states(franka, 32, 1)
    with pytest.raises(iw.IwbcError, match="max_batch"):
        bt.update(q, v)
    with pytest.raises(iw.IwbcError):
        iw.Batch(franka, 99, 16)


def test_franka_known_answer_through_the_gpu_path(franka):
    """tests/ref_test_franka.yaml through the CUDA path: 10 s @ 1 kHz open loop with the cartesian_line behavior, every tick a
    C-ABI update on the GPU (B = 3: the three open-loop actuator modes of the golden file are the same controller run)."""
    import json
    from inria_wbc_b200 import behaviors, ETC_PATH
    gold_all = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "ref_test_franka.json")))
    gold = gold_all["cases"]["servo"]
    ticks = int(gold_all["duration_s"] / gold_all["dt"])
    off, _ = franka.ref_slot("ee")
    B = 3
    bt = iw.Batch(franka, 0, B)

    def mean_error(refs):
        q = np.tile(franka.q0(), (B, 1))
        v = np.zeros((B, franka.nv))
        qs = [q[0].copy()]
        res = iw.Results(franka, B, fields=("q_next", "v_next", "q_cmd", "status"))
        for t in range(ticks):
            bt.update(q, v, np.tile(refs[t], (B, 1)), out=res)
            assert (res.status == 0).all()
            assert np.array_equal(res.q_next[0], res.q_next[1]) and np.array_equal(res.q_next[0], res.q_next[2])
            q, v = res.q_next.copy(), res.v_next.copy()
            qs.append(res.q_cmd[0].copy())
        ee = np.array([ol.frame(franka, qs[t], np.zeros(franka.nv), 0)["oMf"][9:12] for t in range(ticks + 1)])
        return (np.linalg.norm(ee[:-1] - refs[:, off:off + 3], axis=1).mean(), np.linalg.norm(ee[1:] - refs[:, off:off + 3], axis=1).mean())

    cfg = os.path.join(ETC_PATH, "franka", "cartesian_line.yaml")
    before, _ = mean_error(behaviors.Cartesian(franka, cfg).table(ticks))
    assert before <= gold["ee_error_tsid"] + gold_all["tolerance"]  # the reference's own acceptance test (tests/test_franka.cpp:266)
    before, after = mean_error(behaviors.Cartesian(franka, cfg, feed_forward=False).table(ticks))
    assert abs(before - gold["ee_error_tsid"]) < 3e-4 * gold["ee_error_tsid"], before  # the recorded value itself (see test_franka_kat.py)
    assert abs(after - gold["ee_error_dart"]) < 3e-4 * gold["ee_error_dart"], after


@pytest.mark.gpu
def test_cost_and_objective_value_getters(talos, tiago):
    """Controller::cost(task) (controller.hpp:148-152: ||A ddq - b|| of the task's own, unweighted constraint) and
    PosTracker::objective_value() (pos_tracker.hpp:44: eiquadprog's x'Hx/2 + g'x) of one robot, re-evaluated from a stage dump.
    Objective: against the oracle's H, g, x.  Costs: CoM and posture rebuilt in numpy from the oracle's Jcom / CoM drift and the task
    gains; an SE3 task through the identity  sum_t w_t cost_t^2 = 2 f(x) + sum_t w_t ||b_t||^2  restricted to a change of one weight."""
    def task_of(P, name):
        t = P.struct().tasks[P.names("tasks").index(name)]
        return dict(kp=list(t.kp), kd=list(t.kd), mask=list(t.mask), weight=t.weight)

    for P in (talos, tiago):
        q, v = random_states(P, 2, 21, **({"qn": 0.1, "vn": 0.2} if P.ncontacts == 0 else {}))
        bt = iw.Batch(P, 0, 4)
        r = bt.update(q, v)
        for i in range(2):
            d = ol.cycle_debug(P, q[i], v[i])
            f = 0.5 * d["x"] @ d["H"] @ d["x"] + d["g"] @ d["x"]
            assert abs(bt.objective_value(q[i], v[i]) - f) <= 1e-8 * max(1.0, abs(f))
            ddq = r.dv[i]
            na, nu = P.na, P.nv - P.na
            # posture: S (ddq_a - a_des), a_des = -kp (q - q_ref) - kd v  (zero reference velocity / acceleration)
            off, _ = P.ref_slot("posture")
            ref = P.default_refs()
            T = task_of(P, "posture")
            a_des = -T["kp"][0] * (q[i][P.nq - na:] - ref[off:off + na]) - T["kd"][0] * v[i][P.nv - na:]
            want = np.linalg.norm(np.array(T["mask"][:na]) * (ddq[nu:] - a_des))
            assert abs(bt.task_cost("posture", q[i], v[i], ddq) - want) <= 1e-9 * max(1.0, want)
            if "com" in P.names("tasks"):
                T = task_of(P, "com")
                off, _ = P.ref_slot("com")
                com, vcom, drift = d["com"][0:3], d["com"][3:6], d["com"][6:9]
                a_des = -np.array(T["kp"][:3]) * (com - ref[off:off + 3]) - np.array(T["kd"][:3]) * vcom
                res = (d["Jcom"] @ ddq - (a_des - drift)) * np.array(T["mask"][:3])
                want = np.linalg.norm(res)
                assert abs(bt.task_cost("com", q[i], v[i], ddq) - want) <= 1e-8 * max(1.0, want)
    with pytest.raises(iw.IwbcError, match="not a level-1 task"):
        bt.task_cost("bounds", q[0], v[0], r.dv[0])
