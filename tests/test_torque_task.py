"""The `torque` task type of the reference's factory (make_torque, /root/reference/src/controllers/tasks.cpp:227-271:
tsid::tasks::TaskActuationEquality with a joint mask and an optional per-joint `scaling`, added with addActuationTask(task, w, 1)).
No configuration file of the reference uses it and upstream TSID is not in the reference tree, so this row is pinned by identities,
not golden vectors: the oracle's Hessian / gradient contribution equals w D' S^2 D / -w D' S^2 (tau_ref - h) built here in numpy from
the oracle's own M, nle and (for the humanoid) constraint rows; the product's config compiler and the oracle's independent loader
agree on the struct; the device source (host-compiled) and the CUDA kernels agree with the oracle."""
import os

import numpy as np
import pytest

import inria_wbc_b200 as iw
import oracle_lib as ol
from conftest import random_states

OFF = {"stabilizer": {"activated": False}, "collision_detection": {"activated": False}}
TIAGO_URDF = {"urdf": "tiago/tiago_steel_no_wheel.urdf"}


def _tasks_with_torque(tmp_path, robot, extra):
    src = open(os.path.join(iw.ETC_PATH, robot, "tasks.yaml")).read()
    p = tmp_path / ("tasks_torque_%s.yaml" % robot)
    p.write_text(src + "\n" + extra + "\n")
    return str(p)


def _tiago(tmp_path, scaling=True):
    na = 12
    extra = "min_torque: {type: torque, weight: 0.02, mask: '%s'%s}" % (
        "110111111011", (", scaling: [%s]" % ", ".join("%.2f" % (0.5 + 0.1 * i) for i in range(na))) if scaling else "")
    return dict(TIAGO_URDF, tasks=_tasks_with_torque(tmp_path, "tiago", extra))


def _talos(tmp_path):
    return dict(OFF, tasks=_tasks_with_torque(tmp_path, "talos", "min_torque: {type: torque, weight: 0.0005}"))


def test_config_compiler_and_independent_loader_agree(tmp_path):
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle"))
    import problem_loader as pl
    from test_oracle_loader import _cmp, _flat
    for yaml_path, ov in (("tiago/pos_tracker.yaml", _tiago(tmp_path)), ("talos/talos_pos_tracker.yaml", _talos(tmp_path))):
        prod, mine = iw.Problem(yaml_path, overrides=ov), pl.load(yaml_path, overrides=ov)
        errs = []
        _cmp(_flat(mine.struct()), _flat(prod.struct()), "P", errs)
        assert not errs, "\n".join(errs[:20])
        assert np.abs(mine.default_refs() - prod.default_refs()).max() < 1e-13
        assert "min_torque" in prod.names("tasks")
        off, n = prod.ref_slot("min_torque")
        assert n == 2 * prod.na
    # the reference's own checks (tasks.cpp:246-256)
    bad = dict(TIAGO_URDF, tasks=_tasks_with_torque(tmp_path, "tiago", "t2: {type: torque, weight: 1.0, mask: '111'}"))
    with pytest.raises(iw.IwbcError, match="wrong size in torque mask"):
        iw.Problem("tiago/pos_tracker.yaml", overrides=bad)
    with pytest.raises(pl.ConfigError, match="wrong size in torque mask"):
        pl.load("tiago/pos_tracker.yaml", overrides=bad)


def test_oracle_contribution_is_the_weighted_torque_norm(tmp_path):
    """H_with - H_without = w D' S^2 D and g_with - g_without = -w D' S^2 (tau_ref - h), D = [M_a | -J_a^T] taken from the oracle's
    own actuation-bound rows (fixed base: D = M)."""
    # fixed base, no contacts: D = M
    P0 = iw.Problem("tiago/pos_tracker.yaml", overrides=TIAGO_URDF)
    P1 = iw.Problem("tiago/pos_tracker.yaml", overrides=_tiago(tmp_path))
    q, v = random_states(P0, 3, 11, qn=0.2, vn=0.3)
    off, _ = P1.ref_slot("min_torque")
    na = P1.na
    for i in range(3):
        r1 = P1.default_refs()
        r1[off:off + na] = 0.3 * np.arange(na) - 1.0  # a non-zero torque reference
        a, b = ol.cycle_debug(P0, q[i], v[i]), ol.cycle_debug(P1, q[i], v[i], ref=r1)
        mask = np.array([c == "1" for c in "110111111011"], float)
        s2 = (mask * r1[off + na:off + 2 * na]) ** 2
        M, h = a["M"], a["nle"]
        assert np.abs((b["H"] - a["H"]) - 0.02 * M.T @ np.diag(s2) @ M).max() < 1e-12 * np.abs(b["H"]).max()
        assert np.abs((b["g"] - a["g"]) + 0.02 * M.T @ (s2 * (r1[off:off + na] - h))).max() < 1e-12 * max(1.0, np.abs(b["g"]).max())
    # floating base with contacts: the actuation-bound rows of CI are +-[M_a | -J_a^T] (TaskActuationBounds)
    T0 = iw.Problem("talos/talos_pos_tracker.yaml", overrides=OFF)
    T1 = iw.Problem("talos/talos_pos_tracker.yaml", overrides=_talos(tmp_path))
    q, v = random_states(T0, 2, 5, base=True)
    for i in range(2):
        a, b = ol.cycle_debug(T0, q[i], v[i]), ol.cycle_debug(T1, q[i], v[i])
        n, na, nu = a["H"].shape[0], T0.na, T0.nv - T0.na
        M, h = a["M"], a["nle"]
        # find the block of CI whose dv part is M_a: rows r with CI[r, :nv] == M[nu + j]
        CI = a["CI"]
        rows = [int(np.argmin(np.abs(CI[:, :T0.nv] - M[nu + j]).sum(axis=1))) for j in range(na)]
        D = CI[rows]
        assert np.abs(D[:, :T0.nv] - M[nu:]).max() < 1e-12
        assert np.abs((b["H"] - a["H"]) - 0.0005 * D.T @ D).max() < 1e-11 * np.abs(b["H"]).max()
        assert np.abs((b["g"] - a["g"]) + 0.0005 * D.T @ (0.0 - h[nu:])).max() < 1e-11 * max(1.0, np.abs(b["g"]).max())
        assert b["status"][0] == 0


def _close(a, b, keys, tol):
    for k in keys:
        err = np.abs(a[k] - b[k]).max(axis=1) / np.maximum(np.abs(b[k]).max(axis=1), 1e-9)
        assert err.max() < tol, (k, err.max())


def test_device_source_on_the_host_matches_the_oracle(tmp_path):
    P = iw.Problem("tiago/pos_tracker.yaml", overrides=_tiago(tmp_path))
    q, v = random_states(P, 24, 3, qn=0.2, vn=0.3)
    ref, sim = ol.cycle_batch(P, q, v, nthreads=4), ol.sim_cycle_batch(P, q, v)
    assert np.array_equal(ref["status"], sim["status"]) and (ref["status"] == 0).all()
    assert np.array_equal(ref["iters"], sim["iters"])
    _close(sim, ref, ("tau", "dv", "q_next", "v_next"), 1e-8)
    # the task does something: torques differ from the stack without it
    P0 = iw.Problem("tiago/pos_tracker.yaml", overrides=TIAGO_URDF)
    base = ol.cycle_batch(P0, q, v, nthreads=4)
    assert np.abs(base["tau"] - ref["tau"]).max() > 1e-3
    T = iw.Problem("talos/talos_pos_tracker.yaml", overrides=_talos(tmp_path))
    q, v = random_states(T, 6, 4)
    ref, sim = ol.cycle_batch(T, q, v, nthreads=4), ol.sim_cycle_batch(T, q, v)
    assert np.array_equal(ref["status"], sim["status"]) and (ref["status"] == 0).all()
    _close(sim, ref, ("tau", "dv", "f"), 1e-7)


@pytest.mark.gpu
def test_cuda_path_matches_the_oracle(tmp_path):
    """Warp-per-instance kernel (small batch), thread-per-instance kernel (large batch) and the generic CTA kernel (Talos: a problem
    with a torque task never takes the fast pair, its Hessian couples dv and f)."""
    P = iw.Problem("tiago/pos_tracker.yaml", overrides=_tiago(tmp_path))
    for B in (256, 16384):
        q, v = random_states(P, B, 8, qn=0.2, vn=0.3)
        bt = iw.Batch(P, 0, B)
        r = bt.update(q, v)
        n = min(B, 512)
        o = ol.cycle_batch(P, q[:n], v[:n], nthreads=8)
        assert np.array_equal(r.status[:n], o["status"]) and (r.status == 0).all()
        assert np.array_equal(r.iters[:n], o["iters"])
        _close({k: getattr(r, k)[:n] for k in ("tau", "dv", "q_next", "v_next")}, o, ("tau", "dv", "q_next", "v_next"), 1e-8)
    T = iw.Problem("talos/talos_pos_tracker.yaml", overrides=_talos(tmp_path))
    q, v = random_states(T, 64, 4)
    bt = iw.Batch(T, 0, 64)
    r = bt.update(q, v)
    o = ol.cycle_batch(T, q, v, nthreads=8)
    assert np.array_equal(r.status, o["status"]) and (r.status == 0).all()
    _close({k: getattr(r, k) for k in ("tau", "dv", "f")}, o, ("tau", "dv", "f"), 1e-7)
