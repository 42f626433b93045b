"""Device-side behaviours: the compiled segment table and its GPU evaluation against the numpy restatement
of the reference's humanoid::walk (oracle/behaviors_oracle.py, which follows src/behaviors/humanoid/walk.cpp and
include/inria_wbc/trajs/trajectory_generator.hpp), and the closed device-side roll-out against the CPU oracle."""
import os
import sys

import numpy as np
import pytest

import inria_wbc_b200 as iw

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import behaviors_oracle as bo  # noqa: E402

OFF = {"stabilizer": {"activated": False}, "collision_detection": {"activated": False}}
WALK = os.path.join(iw.ETC_PATH, "talos", "walk.yaml")


@pytest.fixture(scope="module")
def walk_problem():
    P = iw.Problem("talos/talos_pos_tracker.yaml", overrides=OFF)
    P.set_task_weight("momentum", 0.0)  # BEHAVIOR.customize_task_weights (etc/talos/walk.yaml:8-9; src/behaviors/behavior.cpp:7-22)
    return P


def test_walk_table_matches_the_reference_state_machine(walk_problem):
    P = walk_problem
    for ncyc in (1, 3):
        beh = iw.Behavior(P, "talos/walk.yaml", overrides={"num_of_cycles": ncyc})
        ora = bo.Walk(P, WALK, num_of_cycles=ncyc)
        assert beh.total_ticks == ora.total_ticks == 1000 * (4 + 6 * ncyc)
        assert beh.nsegments == len(ora.cycle) and beh.nchannels == 5
        refs, masks = ora.table()
        assert sorted(beh.masks) == sorted(set(int(m) for m in masks))
        slots = [ora.slot[k] for k in ("com", "lf", "rf", "lh", "rh")]
        for k in range(beh.nsegments):
            start, ticks, mask, mask_last, wp = beh.segment(k)
            assert start == sum(ora.seg_len[:k]) and ticks == ora.seg_len[k]
            assert np.all(masks[start:start + ticks - 1] == mask) and masks[start + ticks - 1] == mask_last
            for c, off in enumerate(slots):
                # first sample of the segment is the start waypoint; a constant channel keeps it, a moving one follows the polynomial
                assert np.array_equal(refs[start, off:off + 3], wp[c, 0:3]), (k, c)
                if wp[c, 7] == 0:
                    assert np.array_equal(refs[start + ticks - 1, off:off + 3], wp[c, 0:3])
                else:
                    i = ticks // 3
                    want = bo.minimum_jerk_polynom(wp[c, 0:3], wp[c, 3:6], P.struct().dt * i, wp[c, 6])
                    assert np.abs(refs[start + i, off:off + 3] - want).max() < 1e-15


def test_walk_contact_schedule(walk_problem):
    """walk.cpp:203-221: a foot's contact is removed at the first tick of its LIFT_UP state and comes back at the LAST tick of
    its LIFT_DOWN state."""
    P = walk_problem
    beh = iw.Behavior(P, "talos/walk.yaml", overrides={"num_of_cycles": 1})
    lf, rf = 1 << P.contact_index("contact_lfoot"), 1 << P.contact_index("contact_rfoot")
    seq = [beh.segment(k)[2:4] for k in range(beh.nsegments)]
    ds = lf | rf
    assert seq == [(ds, ds), (rf, rf), (rf, ds), (ds, ds), (lf, lf), (lf, ds), (ds, ds), (rf, rf), (rf, ds), (ds, ds)]


def test_loading_a_behaviour_customizes_task_weights():
    """Behavior::_customize_tasks (src/behaviors/behavior.cpp:7-22): etc/talos/walk.yaml sets the momentum task's weight to 0."""
    P = iw.Problem("talos/talos_pos_tracker.yaml", overrides=OFF)
    k = P.names("tasks").index("momentum")
    assert P.struct().tasks[k].weight > 0
    iw.Behavior(P, "talos/walk.yaml")
    assert P.struct().tasks[k].weight == 0.0


def test_behavior_config_errors(walk_problem, franka):
    with pytest.raises(iw.IwbcError, match="num_of_cycles"):
        iw.Behavior(walk_problem, "talos/walk.yaml", overrides={"num_of_cycles": 0})
    # the base-class constructor customizes the task weights first (unknown task on a Franka), then Walk asserts its tasks (walk.cpp:14-20)
    with pytest.raises(iw.IwbcError, match="momentum"):
        iw.Behavior(franka, "talos/walk.yaml")
    with pytest.raises(iw.IwbcError, match="lf"):
        iw.Behavior(franka, "talos/walk.yaml", overrides={"customize_task_weights": "{}"})
    with pytest.raises(iw.IwbcError, match="momentum"):
        iw.Behavior(franka, "talos/walk.yaml")
    with pytest.raises(iw.IwbcError, match="device-side form"):
        iw.Behavior(walk_problem, "talos/walk.yaml", overrides={"name": "humanoid::no-such-behavior"})


def test_cartesian_and_move_com_tables(franka, talos):
    """generic::cartesian (src/behaviors/generic/cartesian.cpp) and humanoid::move_com (src/behaviors/humanoid/move_com.cpp) compiled
    into segment tables, against the host tables of inria_wbc_b200.behaviors (numpy restatements of the same files)."""
    from inria_wbc_b200 import behaviors
    beh = iw.Behavior(franka, "franka/cartesian_line.yaml")
    host = behaviors.Cartesian(franka, os.path.join(iw.ETC_PATH, "franka", "cartesian_line.yaml"))
    tab = host.table(4000)
    off, _ = franka.ref_slot("ee")
    assert beh.total_ticks == 4000 and beh.nsegments == 2 and beh.nchannels == 1 and beh.masks == [franka.default_contact_mask]
    for k in range(2):
        start, ticks, _, _, wp = beh.segment(k)
        assert (start, ticks) == (2000 * k, 2000) and wp[0, 7] == 1 and wp[0, 6] == 2.0
        assert np.abs(wp[0, 0:3] - tab[start, off:off + 3]).max() < 1e-15
    assert np.abs(beh.segment(0)[4][0, 3:6] - beh.segment(1)[4][0, 0:3]).max() == 0  # the loop goes back from the target
    mc = iw.Behavior(talos, "talos/move_com.yaml")
    hmc = behaviors.MoveCom(talos, os.path.join(iw.ETC_PATH, "talos", "move_com.yaml"))
    tab = hmc.table(mc.total_ticks)
    off, _ = talos.ref_slot("com")
    assert mc.nsegments == len(hmc.segs) and mc.nchannels == 1
    for k in range(mc.nsegments):
        start, ticks, _, _, wp = mc.segment(k)
        assert np.abs(wp[0, 0:3] - tab[start, off:off + 3]).max() < 1e-15


@pytest.mark.gpu
def test_walk_refs_on_device_match_the_restatement(walk_problem):
    import torch
    P = walk_problem
    beh = iw.Behavior(P, "talos/walk.yaml", overrides={"num_of_cycles": 2})
    ora = bo.Walk(P, WALK, num_of_cycles=2)
    refs_t, masks_t = ora.table()
    B = 4096
    rng = np.random.default_rng(5)
    phase = rng.integers(0, ora.total_ticks + 500, B).astype(np.int32)  # some instances are past the end: the last references stay
    edges = np.cumsum(ora.seg_len)
    phase[:len(edges)] = edges - 1  # the last tick of every segment (contact re-added there)
    phase[len(edges):2 * len(edges)] = edges % ora.total_ticks
    d_phase = torch.from_numpy(phase).cuda()
    d_refs = torch.full((B, P.nref), float("nan"), dtype=torch.float64, device="cuda")
    d_mask = torch.zeros(B, dtype=torch.int32, device="cuda")
    for tick in (0, 7):
        beh.eval_device(0, B, d_phase.data_ptr(), tick, d_refs.data_ptr(), d_mask.data_ptr(), init_refs=True)
        torch.cuda.synchronize()
        g = np.minimum(phase.astype(np.int64) + tick, ora.total_ticks - 1)
        assert np.array_equal(d_mask.cpu().numpy().astype(np.uint32), masks_t[g])
        got = d_refs.cpu().numpy()
        assert np.isfinite(got).all()
        assert np.abs(got - refs_t[g]).max() < 1e-14


@pytest.mark.gpu
def test_walk_rollout_on_device_matches_the_oracle(walk_problem):
    """Closed device-side roll-out (behaviour -> refs, masks -> device-side bucketing -> one launch set per contact mask -> state
    integrated in place) against tick-by-tick calls of the CPU oracle with the restated behaviour's references.  Instances sit
    at different phases of the walk so that every tick mixes double- and single-support shapes and two of them cross a contact
    removal / re-insertion."""
    import torch
    import oracle_lib as ol
    P = walk_problem
    beh = iw.Behavior(P, "talos/walk.yaml", overrides={"num_of_cycles": 1})
    ora = bo.Walk(P, WALK, num_of_cycles=1)
    T = 30
    phase = np.array([0, 400, 985, 1500, 2985, 3300, 3990, 4500, 9990], dtype=np.int32)
    B = len(phase)
    q = np.tile(P.q0(), (B, 1))
    v = np.zeros((B, P.nv))
    dq, dv_ = torch.from_numpy(q.copy()).cuda(), torch.from_numpy(v.copy()).cuda()
    d_phase = torch.from_numpy(phase).cuda()
    d_refs = torch.empty(B, P.nref, dtype=torch.float64, device="cuda")
    d_mask = torch.zeros(B, dtype=torch.int32, device="cuda")
    st = torch.full((B,), -7, dtype=torch.int32, device="cuda")
    tau = torch.zeros(B, P.na, dtype=torch.float64, device="cuda")
    bt = iw.Batch(P, 0, B)
    bt.behavior_rollout_device(beh, B, dq.data_ptr(), dv_.data_ptr(), d_refs.data_ptr(), d_phase.data_ptr(), 0, T,
                               out_ptrs={"status": st.data_ptr(), "tau": tau.data_ptr()}, mask_ptr=d_mask.data_ptr())
    torch.cuda.synchronize()
    got_q, got_v, got_st, got_tau = dq.cpu().numpy(), dv_.cpu().numpy(), st.cpu().numpy(), tau.cpu().numpy()
    masks_seen = set()
    for i in range(B):
        qi, vi = q[i].copy(), v[i].copy()
        for t in range(T):
            r, m = ora.tick(int(phase[i]) + t)
            masks_seen.add(m)
            o = ol.cycle_batch(P, qi[None], vi[None], r, contact_mask=m)
            qi, vi = o["q_next"][0], o["v_next"][0]
        assert got_st[i] == o["status"][0] == 0, i
        assert np.abs(got_q[i] - qi).max() < 1e-8, i    # tests/test_determinism.cpp:51,63 tolerance on q
        assert np.abs(got_v[i] - vi).max() < 1e-6, i
        assert np.abs(got_tau[i] - o["tau"][0]).max() <= 1e-7 * np.abs(o["tau"][0]).max(), i
        assert int(d_mask[i]) == ora.tick(int(phase[i]) + T - 1)[1]
    assert len(masks_seen) == 3


@pytest.mark.gpu
def test_cartesian_and_move_com_on_device(franka, talos):
    """Every tick of the looping Franka cartesian line and of Talos move_com evaluated on the device in one call (instance i sits
    at tick i) against the host tables: positions, rotations and the feed-forward velocity / acceleration."""
    import torch
    from inria_wbc_b200 import behaviors
    for P, name, Host in ((franka, "franka/cartesian_line.yaml", behaviors.Cartesian), (talos, "talos/move_com.yaml", behaviors.MoveCom)):
        beh = iw.Behavior(P, name)
        B = beh.total_ticks + 300  # past the end: the loop wraps
        tab = Host(P, os.path.join(iw.ETC_PATH, name)).table(B)
        d_phase = torch.arange(B, dtype=torch.int32, device="cuda")
        d_refs = torch.full((B, P.nref), float("nan"), dtype=torch.float64, device="cuda")
        beh.eval_device(0, B, d_phase.data_ptr(), 0, d_refs.data_ptr(), None, init_refs=True)
        torch.cuda.synchronize()
        got = d_refs.cpu().numpy()
        assert np.abs(got - tab).max() < 1e-13, (name, np.abs(got - tab).max())


@pytest.mark.gpu
def test_franka_known_answer_with_the_behaviour_on_the_device(franka):
    """tests/ref_test_franka.yaml with BOTH halves on the GPU: generic::cartesian evaluated on the device every tick
    and the control cycle, 10 s @ 1 kHz open loop; the reference's own acceptance criterion (tests/test_franka.cpp:266)."""
    import json
    import torch
    import oracle_lib as ol
    gold_all = json.load(open(os.path.join(ROOT, "tests", "golden", "ref_test_franka.json")))
    gold = gold_all["cases"]["servo"]
    ticks = int(gold_all["duration_s"] / gold_all["dt"])
    beh = iw.Behavior(franka, "franka/cartesian_line.yaml")
    bt = iw.Batch(franka, 0, 8)
    dq = torch.from_numpy(franka.q0()[None].copy()).cuda()
    dv_ = torch.zeros(1, franka.nv, dtype=torch.float64, device="cuda")
    d_refs = torch.empty(1, franka.nref, dtype=torch.float64, device="cuda")
    st = torch.zeros(1, dtype=torch.int32, device="cuda")
    hist = torch.empty(ticks + 1, franka.nq, dtype=torch.float64, device="cuda")
    refs_hist = torch.empty(ticks, franka.nref, dtype=torch.float64, device="cuda")
    bad = torch.zeros(1, dtype=torch.int32, device="cuda")
    ts = torch.cuda.Stream()  # an explicit stream: handle 0 would mean "the batch's own stream" to the C ABI
    torch.cuda.synchronize()
    with torch.cuda.stream(ts):
        for t in range(ticks):
            hist[t].copy_(dq[0])
            bt.behavior_rollout_device(beh, 1, dq.data_ptr(), dv_.data_ptr(), d_refs.data_ptr(), None, t, 1, init_refs=(t == 0),
                                       out_ptrs={"status": st.data_ptr()}, stream=ts.cuda_stream)
            refs_hist[t].copy_(d_refs[0])
            bad |= st
        hist[ticks].copy_(dq[0])
    torch.cuda.synchronize()
    assert int(bad[0]) == 0
    qs, refs = hist.cpu().numpy(), refs_hist.cpu().numpy()
    off, _ = franka.ref_slot("ee")
    # the reference reads the model pose at the state the tick started from (tests/test_franka.cpp:214-218)
    ee = np.array([ol.frame(franka, qs[t], np.zeros(franka.nv), 0)["oMf"][9:12] for t in range(ticks)])
    err = np.linalg.norm(ee - refs[:, off:off + 3], axis=1).mean()
    assert err <= gold["ee_error_tsid"] + gold_all["tolerance"], err
    # ... and the same number as the CPU oracle driven by the host table of the same behaviour (feed-forward terms included, as the
    # tree has them today; the 2021 recording predates them: tests/test_franka_kat.py)
    from inria_wbc_b200 import behaviors
    tab = behaviors.Cartesian(franka, os.path.join(iw.ETC_PATH, "franka", "cartesian_line.yaml")).table(ticks)
    r = ol.rollout(franka, franka.q0(), np.zeros(franka.nv), tab, ticks)
    dq_max = np.abs(qs[-1] - r["q"]).max()
    assert r["rc"] == 0 and dq_max < 1e-8, dq_max  # tests/test_determinism.cpp:51,63 tolerance after 10 000 ticks


@pytest.mark.gpu
def test_arm_behaviour_rollout_takes_the_thread_per_instance_kernel(franka):
    """A behaviour roll-out of a LARGE arm batch runs the thread-per-instance kernel over the behaviour's per-mask instance list, in
    several passes per launch when the batch exceeds the resident capacity (cycle_tu.cuh: launch_tpi); the same robots in a small
    batch run the warp-per-instance kernel.  Same instances, same phases: the two must agree to rounding after 40 closed ticks."""
    B, n = 98304, 96  # > 75 776 resident threads: two passes
    rng = np.random.default_rng(5)
    beh = iw.Behavior(franka, "franka/cartesian_line.yaml")
    q = np.tile(franka.q0(), (B, 1)) + 0.05 * rng.standard_normal((B, franka.nq))
    v = 0.05 * rng.standard_normal((B, franka.nv))
    phase = rng.integers(0, beh.total_ticks, B).astype(np.int32)
    pick = np.concatenate([np.arange(n // 2), B - 1 - np.arange(n // 2)])  # instances of the first and of the last pass
    qs, vs = q[pick].copy(), v[pick].copy()
    big, small = iw.Batch(franka, 0, B), iw.Batch(franka, 0, n)
    l0 = big.launch_count
    rb = big.behavior_run(beh, q, v, 40, phase=phase, fields=("tau", "status"))
    assert big.launch_count - l0 >= 40 * 2  # (behaviour kernel + cycle passes per tick)
    rs = small.behavior_run(beh, qs, vs, 40, phase=phase[pick], fields=("tau", "status"))
    assert (rb.status == 0).all() and (rs.status == 0).all()
    assert np.abs(q[pick] - qs).max() < 1e-10 and np.abs(v[pick] - vs).max() < 1e-8
    assert np.abs(rb.tau[pick] - rs.tau).max() < 1e-8 * max(1.0, np.abs(rs.tau).max())


def test_walk_on_spot_table(talos):
    """humanoid::walk-on-spot (src/behaviors/humanoid/walk_on_spot.cpp): INIT once, then six states for ever."""
    P = iw.Problem("talos/talos_pos_tracker.yaml", overrides=OFF)
    beh = iw.Behavior(P, "talos/walk_on_spot.yaml")
    ora = bo.WalkOnSpot(P, os.path.join(iw.ETC_PATH, "talos", "walk_on_spot.yaml"))
    assert beh.total_ticks == 7000 and beh.nsegments == 7 and beh.nchannels == 3
    refs, masks = ora.play(7000)
    lf, rf = 1 << P.contact_index("contact_lfoot"), 1 << P.contact_index("contact_rfoot")
    ds = lf | rf
    assert [beh.segment(k)[2:4] for k in range(7)] == [(ds, ds), (rf, rf), (rf, ds), (ds, ds), (lf, lf), (lf, ds), (ds, ds)]
    slots = [ora.slot[k] for k in ("com", "lf", "rf")]
    for k in range(7):
        start, ticks, mask, mask_last, wp = beh.segment(k)
        assert np.all(masks[start:start + ticks - 1] == mask) and masks[start + ticks - 1] == mask_last
        for c, off in enumerate(slots):
            assert np.array_equal(refs[start, off:off + 3], wp[c, 0:3])


@pytest.mark.gpu
def test_walk_on_spot_on_device_loops_past_its_init_segment(talos):
    import torch
    P = iw.Problem("talos/talos_pos_tracker.yaml", overrides=OFF)
    beh = iw.Behavior(P, "talos/walk_on_spot.yaml")
    ora = bo.WalkOnSpot(P, os.path.join(iw.ETC_PATH, "talos", "walk_on_spot.yaml"))
    B = 7000 + 2 * 6000 + 17  # INIT + two and a bit loops
    refs, masks = ora.play(B)
    d_phase = torch.arange(B, dtype=torch.int32, device="cuda")
    d_refs = torch.full((B, P.nref), float("nan"), dtype=torch.float64, device="cuda")
    d_mask = torch.zeros(B, dtype=torch.int32, device="cuda")
    beh.eval_device(0, B, d_phase.data_ptr(), 0, d_refs.data_ptr(), d_mask.data_ptr(), init_refs=True)
    torch.cuda.synchronize()
    assert np.array_equal(d_mask.cpu().numpy().astype(np.uint32), masks)
    assert np.abs(d_refs.cpu().numpy() - refs).max() < 1e-14


def test_remaining_reference_behaviour_files_compile(talos, icub):
    """squat = humanoid::move_com, arm / align = generic::cartesian, clapping: every behaviour file of etc/ except move_feet and the
    teleoperation trajectory replays has a device-side form."""
    from inria_wbc_b200 import trajs
    for P, name, ticks, nseg, nch in ((talos, "talos/squat.yaml", 4000, 2, 1), (talos, "talos/arm.yaml", 2000, 2, 2), (icub, "icub/squat.yaml", 2000, 2, 1),
                                      (icub, "icub/arm.yaml", 2000, 2, 2), (talos, "talos/clapping.yaml", None, 2, 2), (talos, "talos/align.yaml", 10000, 1, 2),
                                      (talos, "talos/move_feet.yaml", 4000, 1, 2)):
        b = iw.Behavior(P, name)
        assert (b.nsegments, b.nchannels) == (nseg, nch) and (ticks is None or b.total_ticks == ticks), name
    # move-feet drives the foot task AND the contact sharing its frame (move_feet.cpp:86-87); the device writes both slots
    mf = iw.Behavior(talos, "talos/move_feet.yaml")
    wp = mf.segment(0)[4]
    lf = talos.ref_slot("lf")[0]
    assert np.allclose(wp[0, 3:6] - wp[0, 0:3], 0.05) and np.array_equal(wp[0, 0:3], talos.default_refs()[lf:lf + 3])
    # clapping: hands towards each other along y and back (clapping.cpp:26-32)
    b = iw.Behavior(talos, "talos/clapping.yaml")
    (_, _, _, _, go), (_, _, _, _, back) = b.segment(0), b.segment(1)
    ref0 = talos.default_refs()
    lh, rh = talos.ref_slot("lh")[0], talos.ref_slot("rh")[0]
    assert np.array_equal(go[0, 0:3], ref0[lh:lh + 3]) and np.array_equal(go[1, 0:3], ref0[rh:rh + 3])
    assert go[0, 4] < go[0, 1] and go[1, 4] > go[1, 1] and np.array_equal(go[:, 3:6], back[:, 0:3]) and np.array_equal(back[:, 3:6], go[:, 0:3])


def _write_recorded_trajectory(tmp_path, P, rows=37):
    """A small recorded trajectory in the reference's format (src/trajs/loader.cpp:11-100): one CSV per task, 3 + 9 numbers per row."""
    rng = np.random.default_rng(11)
    ref0 = P.default_refs()
    d = tmp_path / "traj_rec"
    d.mkdir()
    tabs = {}
    for name, width in (("lh", 12), ("rh", 12), ("com", 3)):
        off = P.ref_slot(name)[0]
        base = ref0[off:off + width]
        tab = base[None, :] + np.cumsum(rng.normal(0, 1e-3, (rows, width)), axis=0)
        np.savetxt(d / (name + ".csv"), tab, fmt="%.17g")
        tabs[name] = np.loadtxt(d / (name + ".csv"))
    (d / "trajectory.yaml").write_text("refs:\n  lh: lh.csv\n  rh: rh.csv\n  com: com.csv\n")
    (tmp_path / "replay.yaml").write_text("BEHAVIOR:\n  name : generic::cartesian_traj\n  trajectories: traj_rec/trajectory.yaml\n  loop: true\n  scale: 0.5\n")
    return str(tmp_path / "replay.yaml"), tabs


def _replay_reference(P, tabs, scale, ticks):
    """cartesian_traj.cpp:19-45 restated: row `time_` of every table (translation / CoM scaled), then time_ += step_ with the step
    flipping at both ends."""
    refs = np.tile(P.default_refs(), (ticks, 1))
    rows = len(tabs["com"])
    time_, step = 0, 1
    for g in range(ticks):
        for name in ("lh", "rh"):
            off = P.ref_slot(name)[0]
            refs[g, off:off + 3] = tabs[name][time_, :3] * scale
            refs[g, off + 3:off + 12] = tabs[name][time_, 3:]
            refs[g, off + 12:off + 24] = 0.0
        off = P.ref_slot("com")[0]
        refs[g, off:off + 3] = tabs["com"][time_] * scale
        refs[g, off + 3:off + 9] = 0.0
        time_ += step
        if time_ >= rows - 1:
            step = -1
        elif time_ <= 0:
            step = 1
    return refs


def test_recorded_trajectory_replay_table(talos, tmp_path):
    yaml_path, tabs = _write_recorded_trajectory(tmp_path, talos)
    beh = iw.Behavior(talos, yaml_path)
    assert beh.total_ticks == 2 * 37 - 2 and beh.nchannels == 3 and beh.masks == [talos.default_contact_mask]
    ref = "/root/reference/etc/talos/traj_teleop1.yaml"
    if os.path.exists(ref):  # the reference's own recording (2425 rows), read where it lies (build container only)
        assert iw.Behavior(talos, ref).total_ticks == 2 * 2425 - 2
    (tmp_path / "traj_rec" / "com.csv").write_text("0 0 0\n1 1\n")
    with pytest.raises(iw.IwbcError, match="wrong number"):
        iw.Behavior(talos, yaml_path)


@pytest.mark.gpu
def test_recorded_trajectory_replay_on_device(talos, tmp_path):
    import torch
    yaml_path, tabs = _write_recorded_trajectory(tmp_path, talos)
    beh = iw.Behavior(talos, yaml_path)
    B = 3 * beh.total_ticks + 5
    want = _replay_reference(talos, tabs, 0.5, B)
    d_phase = torch.arange(B, dtype=torch.int32, device="cuda")
    d_refs = torch.full((B, talos.nref), float("nan"), dtype=torch.float64, device="cuda")
    beh.eval_device(0, B, d_phase.data_ptr(), 0, d_refs.data_ptr(), None, init_refs=True)
    torch.cuda.synchronize()
    assert np.array_equal(d_refs.cpu().numpy(), want)


@pytest.mark.gpu
def test_icub_walk_rollout_on_device(icub):
    """iCub walking (etc/icub/walk.yaml: 1 s CoM shifts, 0.25 s foot trajectories) as one closed device-side roll-out of 600 ticks against
    600 oracle ticks driven by the restated behaviour, then tick by tick.  From tick ~690 on the stand-in iCub model walks along the
    edge of feasibility (20+ active-set iterations per tick): the oracle's outcome flips there under a 1e-13 perturbation of the state
    (measured first non-optimal tick on the GPU: 744 .. 994 depending on the summation orders inside the kernel; the oracle on its own
    trajectory: ~930), so beyond tick 600 the comparison is state-wise: at every tick the oracle is evaluated on the GPU's own state, and
    while both solve the QP they must land on the same next state to 1e-8 (also where their active-set paths differ in length).  The
    comparison ends at the first tick one of them does not solve -- which must be a QP on the edge of feasibility (after tick 650, 20+
    iterations); measured so far, the two always agreed there too (status 1 after 49 iterations on both sides)."""
    import torch
    import oracle_lib as ol
    P = iw.Problem("icub/humanoid_pos_tracker.yaml", overrides=OFF)
    beh = iw.Behavior(P, "icub/walk.yaml", overrides={"num_of_cycles": 1})
    ora = bo.Walk(P, os.path.join(iw.ETC_PATH, "icub", "walk.yaml"), num_of_cycles=1)
    assert beh.total_ticks == ora.total_ticks == 5500
    T = 600
    q, v = P.q0().copy(), np.zeros(P.nv)
    for t in range(T):
        r, m = ora.tick(t)
        o = ol.cycle_batch(P, q[None], v[None], r, contact_mask=m)
        assert o["status"][0] == 0, t
        q, v = o["q_next"][0], o["v_next"][0]
    dq = torch.from_numpy(P.q0()[None].copy()).cuda()
    dv_ = torch.zeros(1, P.nv, dtype=torch.float64, device="cuda")
    d_refs = torch.empty(1, P.nref, dtype=torch.float64, device="cuda")
    st = torch.full((1,), -7, dtype=torch.int32, device="cuda")
    it = torch.zeros(1, dtype=torch.int32, device="cuda")
    bt = iw.Batch(P, 0, 4)
    bt.behavior_rollout_device(beh, 1, dq.data_ptr(), dv_.data_ptr(), d_refs.data_ptr(), None, 0, T, out_ptrs={"status": st.data_ptr()})
    torch.cuda.synchronize()
    assert int(st[0]) == 0
    assert np.abs(dq.cpu().numpy()[0] - q).max() < 1e-8 and np.abs(dv_.cpu().numpy()[0] - v).max() < 1e-6
    ended = None
    for t in range(T, 1500):
        pq, pv = dq.cpu().numpy().copy(), dv_.cpu().numpy().copy()
        bt.behavior_rollout_device(beh, 1, dq.data_ptr(), dv_.data_ptr(), d_refs.data_ptr(), None, t, 1, init_refs=False,
                                   out_ptrs={"status": st.data_ptr(), "iters": it.data_ptr()})
        torch.cuda.synchronize()
        o = ol.cycle_batch(P, pq, pv, d_refs.cpu().numpy(), contact_mask=ora.tick(t)[1])
        gs, gi, os_, oi = int(st[0]), int(it[0]), int(o["status"][0]), int(o["iters"][0])
        if gs == 0 and os_ == 0:
            # both solve it: the same vertex, also where the two active-set paths differ in length (measured: 25 vs 27 iterations)
            assert np.abs(o["q_next"][0] - dq.cpu().numpy()[0]).max() < 1e-8, (t, gi, oi)
            continue
        ended = (t, gs, gi, os_, oi)
        break
    # the comparison may only end at a QP on the edge of feasibility: late in the CoM shift, after a long active-set iteration
    assert ended is not None and ended[0] > 650 and max(ended[2], ended[4]) >= 20, ended
