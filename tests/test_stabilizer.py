"""Stabilizer + CoP estimator + filters as a batched pre-step.

CPU part: the restatement itself (oracle/stabilizer_oracle.py, which follows src/stabilizers/stabilizer.cpp, src/estimators/cop.cpp,
include/inria_wbc/estimators/filtering.hpp and HumanoidPosTracker::update line by line) against closed-form properties, and the
stabilizer parameter sets read from the reference's stab_*.yaml files.
GPU part: B robots driven for T ticks with synthetic F/T + IMU streams (loaded and unloaded feet, double / single support, behaviour
type switches that resize the filter windows) through iwbc_batch_update_sensors, against one restated controller per robot: the
reference rows the solve saw, and tau / dv / f of the solve."""
import os
import sys

import numpy as np
import pytest

import inria_wbc_b200 as iw
import oracle_lib as ol

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import stabilizer_oracle as so  # noqa: E402


@pytest.fixture(scope="module")
def talos_stab():
    return iw.Problem("talos/talos_pos_tracker.yaml", overrides={"collision_detection": {"activated": False}})  # stabilizer.activated: true, as shipped


@pytest.fixture(scope="module")
def icub_stab():
    return iw.Problem("icub/humanoid_pos_tracker.yaml")


def test_parameter_sets_come_from_the_reference_files(talos_stab, icub_stab):
    import yaml
    for P, robot in ((talos_stab, "talos"), (icub_stab, "icub")):
        S = P.struct()
        assert S.stabilizer_activated == 1 and S.stab_present == 1
        for bt, name in ((0, "stab_fixed_base.yaml"), (1, "stab_single_support.yaml"), (2, "stab_double_support.yaml")):
            y = yaml.safe_load(open(os.path.join(iw.ETC_PATH, robot, name)))
            c = S.stab[bt]
            assert c.filter_size == y["filter_size"] and bool(c.use_zmp) == y["use_zmp"] and bool(c.use_momentum) == y["use_momentum"]
            for key, field in (("com", "com_gains"), ("ankle", "ankle_gains"), ("ffda", "ffda_gains"), ("zmp_p", "zmp_p"), ("zmp_d", "zmp_d"),
                               ("zmp_w", "zmp_w"), ("momentum_p", "momentum_p"), ("momentum_d", "momentum_d")):
                assert list(getattr(c, field)) == [float(x) for x in y[key]], (robot, name, key)
            if c.use_zmp:  # zmp_w is one of the weight vectors a contact may carry
                assert list(S.wreg_variants[c.zmp_w_variant]) == [float(x) for x in y["zmp_w"]]
        assert list(S.wreg_variants[0]) == [1, 1, 1e-3, 2, 2, 2] and 1 <= S.nwreg <= 4
    assert talos_stab.struct().stab_talos == 1 and talos_stab.struct().stab_imu_frame >= 0 and icub_stab.struct().stab_talos == 0


def test_filter_follows_the_reference_semantics():
    f = so.MovingAverageFilter(2, 3)
    xs = np.arange(1.0, 11.0)
    for k, x in enumerate(xs):
        out = f.filter([x, -x])
        lo = max(0, k - 2)
        assert np.allclose(out, [xs[lo:k + 1].mean(), -xs[lo:k + 1].mean()]) and f.data_ready() == (k >= 2)
    f.set_window_size(5)  # grow: old columns in front, zeros behind, the count stays (filtering.hpp:63-86)
    assert f.data_ready() is False and f.cnt == 3
    assert np.allclose(f.filter([0.0, 0.0]), [(8 + 9 + 10 + 0) / 4.0, -(8 + 9 + 10) / 4.0])
    f.set_window_size(2)  # shrink: the last columns
    assert f.data_ready() and np.allclose(f.filter([4.0, 4.0]), [(0 + 4) / 2.0, 2.0])


def test_euler_angles_round_trip_and_first_angle_range():
    rng = np.random.default_rng(0)
    for _ in range(200):
        e = rng.uniform(-1.4, 1.4, 3)
        R = so.rot_xyz(e)
        got = so.euler_angles_012(R)
        assert np.abs(so.rot_xyz(got) - R).max() < 1e-12
        assert 0.0 <= got[0] <= np.pi + 1e-12  # Eigen's convention: the first angle lies in [0, pi]


def test_cop_and_zmp_distributor_closed_forms():
    # a pure vertical force F applied at (x, y) on the sole: the sensor at height h sees torque (y F, -x F, 0); the CoP is (x, y)
    foot = np.array([0.1, 0.2, 0.107])
    for x, y in ((0.03, -0.02), (-0.05, 0.01)):
        F = 400.0
        c = so.Cop.foot_cop(foot, np.array([y * F, -x * F, 0.0]), np.array([0.0, 0.0, F]))
        assert np.allclose(c, foot[:2] + [x, y])
    # distributor: forces sum to -M g, split by the projection of the ZMP on the line between the feet
    lf, rf, mass = np.array([0.0, 0.1, 0.1]), np.array([0.0, -0.1, 0.1]), 90.0
    for zy, want in ((0.1, 0.0), (-0.1, 1.0), (0.0, 0.5), (0.05, 0.25), (0.3, 0.0), (-0.3, 1.0)):
        alpha, left, right = so.zmp_distributor(mass, np.array([0.01, zy]), lf, rf, True, True)
        assert abs(alpha - want) < 1e-12 and abs(left[2] + right[2] + mass * 9.81) < 1e-9
    assert so.zmp_distributor(mass, np.zeros(2), lf, rf, False, True)[0] == 1.0 and so.zmp_distributor(mass, np.zeros(2), lf, rf, True, False)[0] == 0.0
    with pytest.raises(RuntimeError):
        so.zmp_distributor(mass, np.zeros(2), lf, rf, False, False)


def _streams(P, B, T, seed):
    """Synthetic sensor streams and contact schedules: robots 0 .. B/2 stay in double support; the others lift a foot for a while
    (contact removed, that foot's force drops below FMIN) and switch their behaviour type with it."""
    rng = np.random.default_rng(seed)
    S = P.struct()
    Mg = S.total_mass * 9.81
    cl, cr = S.stab_contact_lf, S.stab_contact_rf
    both = (1 << cl) | (1 << cr)
    sensors = np.zeros((T, B, 15))
    masks = np.full((T, B), both, dtype=np.uint32)
    btype = np.full((T, B), so.BT_DOUBLE_SUPPORT, dtype=np.int32)
    for i in range(B):
        share = rng.uniform(0.35, 0.65)
        lift = None
        if i >= B // 2:
            t0 = int(rng.integers(8, max(9, T - 12)))
            lift = (t0, t0 + int(rng.integers(4, 10)), int(rng.integers(0, 2)))
        for t in range(T):
            lz, rz = share * Mg, (1 - share) * Mg
            if lift and lift[0] <= t < lift[1]:
                masks[t, i] = (1 << cr) if lift[2] == 0 else (1 << cl)
                btype[t, i] = so.BT_SINGLE_SUPPORT
                if lift[2] == 0:
                    lz, rz = rng.uniform(0, 20), Mg
                else:
                    lz, rz = Mg, rng.uniform(0, 20)
            sensors[t, i, 0:3] = [rng.normal(0, 3), rng.normal(0, 3), lz + rng.normal(0, 4)]
            sensors[t, i, 3:6] = [rng.normal(0, 3), rng.normal(0, 3), rz + rng.normal(0, 4)]
            sensors[t, i, 6:9] = rng.normal(0, 4, 3) * (lz > 30)
            sensors[t, i, 9:12] = rng.normal(0, 4, 3) * (rz > 30)
            sensors[t, i, 12:15] = rng.normal(0, 0.02, 3)
    return sensors, masks, btype


def _dict(row):
    return dict(lf_force=row[0:3], rf_force=row[3:6], lf_torque=row[6:9], rf_torque=row[9:12], imu_vel=row[12:15])


def test_restated_controller_applies_every_law(talos_stab):
    """CPU: over a stream with loaded feet the restated controller touches exactly the references the reference's laws touch, installs
    zmp_w once the CoP filter is full, and leaves the behaviour's row alone."""
    P = talos_stab
    sensors, masks, btype = _streams(P, 2, 24, 5)
    c = so.HumanoidController(ol, P)
    c.set_behavior_type(so.BT_DOUBLE_SUPPORT)
    refs = P.default_refs()
    keep = refs.copy()
    touched = set()
    names = {P.ref_slot(n)[0]: n for n in ("com", "lf", "rf", "torso", "momentum", "contact_lfoot", "contact_rfoot")}
    for t in range(24):
        o, r = c.update(refs, _dict(sensors[t, 0]), int(masks[t, 0]))
        assert o["status"][0] == 0 and np.array_equal(refs, keep)
        for off, n in names.items():
            ln = P.ref_slot(n)[1]
            if np.any(r[off:off + ln] != refs[off:off + ln]):
                touched.add(n)
        if t < 6:
            assert c.wreg_variant == 0  # the CoP filter (window 7) is not full yet: no ZMP distributor
    assert touched == set(names.values()) and c.wreg_variant == P.struct().stab[2].zmp_w_variant != 0


@pytest.mark.gpu
@pytest.mark.parametrize("robot", ["talos", "icub"])
def test_batched_stabilizer_matches_the_restated_controllers(robot, talos_stab, icub_stab):
    P = talos_stab if robot == "talos" else icub_stab
    B, T = 48, 36
    sensors, masks, btype = _streams(P, B, T, 11)
    rng = np.random.default_rng(3)
    refs = np.tile(P.default_refs(), (B, 1))
    com, _ = P.ref_slot("com")
    refs[:, com + 2] -= rng.uniform(0, 0.05, B)  # every robot has its own behaviour references
    bt = iw.Batch(P, 0, B)
    with pytest.raises(iw.IwbcError, match="stabilizer needs the LF force"):  # humanoid_pos_tracker.cpp:180: no sensors, no update
        bt.update(np.tile(P.q0(), (B, 1)), np.zeros((B, P.nv)))
    stab = bt.stabilizer()
    ctrl = [so.HumanoidController(ol, P) for _ in range(B)]
    q, v = np.tile(P.q0(), (B, 1)), np.zeros((B, P.nv))
    from test_cpu_sim import active_sets_equivalent
    worst = dict(refs=0.0, tau=0.0, dv=0.0, f=0.0, f_other_abs=0.0)
    same_sets = total = 0
    for t in range(T):
        out, used = bt.update_sensors(stab, q, v, sensors[t], refs, contact_mask=masks[t], behavior_type=btype[t], want_refs=True)
        qn, vn = q.copy(), v.copy()
        for i in range(B):
            if ctrl[i].behavior_type != btype[t, i]:
                ctrl[i].set_behavior_type(int(btype[t, i]))
            o, r = ctrl[i].update(refs[i], _dict(sensors[t, i]), int(masks[t, i]), q=q[i], v=v[i])
            scale = np.maximum(1.0, np.abs(r))
            worst["refs"] = max(worst["refs"], float((np.abs(used[i] - r) / scale).max()))
            assert out.status[i] == o["status"][0] == 0, (t, i)
            same = np.array_equal(out.active[i], o["active"][0])
            same_sets += same
            total += 1
            for k in ("tau", "dv", "f"):
                err = np.abs(getattr(out, k)[i] - o[k][0]).max() / max(np.abs(o[k][0]).max(), 1e-9)
                if k != "f" or same:
                    worst[k] = max(worst[k], float(err))
            if not same:
                # another description of the same vertex, or an argmin tie inside eiquadprog's termination slack (tests/test_gpu_parity.py):
                # the force split between corners may then differ by a fraction of a newton, nothing else does
                worst["f_other_abs"] = max(worst["f_other_abs"], float(np.abs(out.f[i] - o["f"][0]).max()))
                assert out.iters[i] == o["iters"][0] or active_sets_equivalent(P, q[i], v[i], r, o["active"][0], out.active[i], int(masks[t, i])), (t, i)
            qn[i], vn[i] = o["q_next"][0], o["v_next"][0]
        q, v = qn, vn  # closed loop on the restated controllers' state: both sides see identical inputs at every tick
    print("stabilizer %s: %d robots x %d ticks, worst relative error refs %.2e tau %.2e dv %.2e f %.2e (identical active sets: %d / %d; the others: "
          "max |df| %.3g N)" % (robot, B, T, worst["refs"], worst["tau"], worst["dv"], worst["f"], same_sets, total, worst["f_other_abs"]))
    assert worst["refs"] < 1e-9 and worst["tau"] < 1e-8 and worst["dv"] < 1e-8 and worst["f"] < 5e-7 and worst["f_other_abs"] < 0.5
    assert same_sets > 0.9 * total
    if robot == "talos":  # (the iCub parameter sets have use_zmp: false)
        assert any(c.wreg_variant != 0 for c in ctrl)  # the ZMP distributor really ran


@pytest.mark.gpu
def test_stabilizer_state_reset_and_deactivated_controller(talos):
    """A controller loaded with stabilizer.activated = false still accepts the sensor call (the CoP is estimated, nothing is rewritten:
    humanoid_pos_tracker.cpp:161-179) and gives the plain update's result."""
    P = talos
    B = 8
    q, v = np.tile(P.q0(), (B, 1)), np.zeros((B, P.nv))
    sensors = _streams(P, B, 4, 1)[0]
    bt = iw.Batch(P, 0, B)
    stab = bt.stabilizer()
    plain = bt.update(q, v)
    out, used = bt.update_sensors(stab, q, v, sensors[0], want_refs=True)
    assert np.array_equal(used, np.tile(P.default_refs(), (B, 1)))
    for k in ("tau", "dv", "f", "status"):
        assert np.array_equal(getattr(out, k), getattr(plain, k)), k
    stab.reset()
