"""The `cop` task type of the reference's factory (make_cop, /root/reference/src/controllers/tasks.cpp:156-178:
tsid::tasks::TaskCopEquality with reference (0, 0, 0), added with addForceTask(task, w, 1)).  No configuration file of the reference
uses it and upstream TSID is not in the reference tree (parity unpinned): the row is held by identities -- the oracle's Hessian
contribution equals w A'A with A = [(p_world_j - cop_ref) n'] per contact point, rebuilt here in numpy from the oracle's own joint
placements and the contact geometry of the task file; gradient unchanged (zero right-hand side); config compiler and independent
loader agree on the struct; host-compiled device source and the CUDA kernel agree with the oracle."""
import os

import numpy as np
import pytest
import yaml

import inria_wbc_b200 as iw
import oracle_lib as ol
from conftest import random_states

OFF = {"stabilizer": {"activated": False}, "collision_detection": {"activated": False}}
W = 2.0e-4


def _talos(tmp_path):
    src = open(os.path.join(iw.ETC_PATH, "talos", "tasks.yaml")).read()
    p = tmp_path / "tasks_cop_talos.yaml"
    p.write_text(src + "\ncop_task: {type: cop, weight: %g}\n" % W)
    return dict(OFF, tasks=str(p))


def test_config_compiler_and_independent_loader_agree(tmp_path):
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle"))
    import problem_loader as pl
    from test_oracle_loader import _cmp, _flat
    ov = _talos(tmp_path)
    prod, mine = iw.Problem("talos/talos_pos_tracker.yaml", overrides=ov), pl.load("talos/talos_pos_tracker.yaml", overrides=ov)
    errs = []
    _cmp(_flat(mine.struct()), _flat(prod.struct()), "P", errs)
    assert not errs, "\n".join(errs[:20])
    assert np.abs(mine.default_refs() - prod.default_refs()).max() < 1e-13
    off, n = prod.ref_slot("cop_task")
    assert n == 6 and mine.ref_slot("cop_task") == (off, n)
    assert prod.default_refs()[off:off + 6].tolist() == [0, 0, 0, 0, 0, 1]
    assert prod.qp_shape() == iw.Problem("talos/talos_pos_tracker.yaml", overrides=OFF).qp_shape()[:3] + (prod.qp_shape()[3],)


def test_oracle_contribution_is_the_cop_moment(tmp_path):
    T0 = iw.Problem("talos/talos_pos_tracker.yaml", overrides=OFF)
    T1 = iw.Problem("talos/talos_pos_tracker.yaml", overrides=_talos(tmp_path))
    tasks = yaml.safe_load(open(os.path.join(iw.ETC_PATH, "talos", "tasks.yaml")))
    contacts = [(k, v) for k, v in tasks.items() if v.get("type") == "contact"]
    q, v = random_states(T0, 2, 9, base=True)
    off, _ = T1.ref_slot("cop_task")
    for i in range(2):
        ref = T1.default_refs()
        ref[off:off + 3] = [0.02, -0.01, 0.0]  # a non-zero CoP reference
        a, b = ol.cycle_debug(T0, q[i], v[i]), ol.cycle_debug(T1, q[i], v[i], ref=ref)
        n, nv = a["H"].shape[0], T0.nv
        A = np.zeros((3, n))
        for c, (name, node) in enumerate(contacts):
            fr = ol.frame(T0, q[i], v[i], T0.struct().contacts[c].frame)["oMf"]  # R row-major (9), p (3)
            R, p = np.array(fr[:9]).reshape(3, 3), np.array(fr[9:12])
            lxn, lyn, lxp, lyp, lz = (float(node[k]) for k in ("lxn", "lyn", "lxp", "lyp", "lz"))
            pts = [(-lxn, -lyn, lz), (-lxn, lyp, lz), (lxp, -lyn, lz), (lxp, lyp, lz)]  # tasks.cpp:353-356
            for j, P in enumerate(pts):
                pw = R @ np.array(P) + p
                A[:, nv + 12 * c + 3 * j + 2] = pw - ref[off:off + 3]  # normal (0, 0, 1): only the z coefficient of the point's force
        assert np.abs((b["H"] - a["H"]) - W * A.T @ A).max() < 1e-12 * np.abs(b["H"]).max()
        assert np.abs(b["g"] - a["g"]).max() < 1e-12 * max(1.0, np.abs(a["g"]).max())
        assert b["status"][0] == 0
        assert np.abs(b["x"] - a["x"]).max() > 1e-6  # the task does something to the force distribution


def _close(a, b, keys, tol):
    for k in keys:
        err = np.abs(a[k] - b[k]).max(axis=1) / np.maximum(np.abs(b[k]).max(axis=1), 1e-9)
        assert err.max() < tol, (k, err.max())


def test_device_source_on_the_host_matches_the_oracle(tmp_path):
    T = iw.Problem("talos/talos_pos_tracker.yaml", overrides=_talos(tmp_path))
    q, v = random_states(T, 6, 4)
    ref, sim = ol.cycle_batch(T, q, v, nthreads=4), ol.sim_cycle_batch(T, q, v)
    assert np.array_equal(ref["status"], sim["status"]) and (ref["status"] == 0).all()
    _close(sim, ref, ("tau", "dv", "f"), 1e-7)


@pytest.mark.gpu
def test_cuda_path_matches_the_oracle(tmp_path):
    """The generic CTA kernel: a problem with a cop task never takes the fast pair (its force block of H depends on the state)."""
    T = iw.Problem("talos/talos_pos_tracker.yaml", overrides=_talos(tmp_path))
    for mask in (None, 1):
        q, v = random_states(T, 48, 6)
        bt = iw.Batch(T, 0, 48)
        r = bt.update(q, v, contact_mask=mask)
        o = ol.cycle_batch(T, q, v, contact_mask=mask, nthreads=8)
        assert np.array_equal(r.status, o["status"])
        ok = o["status"] == 0
        assert ok.mean() > 0.5
        _close({k: getattr(r, k)[ok] for k in ("tau", "dv", "f")}, {k: o[k][ok] for k in ("tau", "dv", "f")}, ("tau", "dv", "f"), 1e-7)
