"""Handle-lifetime regression: a freed batch's address may be handed to the next batch; the __constant__ tables of the per-shape
kernel units must then be refreshed (ownership is tracked by a never-reused batch id, not by address).  iCub and the fast-URDF
Talos share the (nv, me) = (38, 18) unit, which is where a stale table shows."""
import numpy as np
import pytest

import inria_wbc_b200 as iw
import oracle_lib as ol
from conftest import random_states

pytestmark = pytest.mark.gpu


def test_constant_tables_follow_the_batch_not_its_address(icub, talos_fast, talos):
    for rep in range(3):
        for P in (icub, talos_fast, talos):
            q, v = random_states(P, 96, 100 + rep)
            bt = iw.Batch(P, 0, 96)          # created and freed in turn: allocations tend to reuse the same address
            r = bt.update(q, v)
            o = ol.cycle_batch(P, q, v, nthreads=4)
            assert np.array_equal(r.status, o["status"]), (rep, P.nq)
            same = np.all(r.active == o["active"], axis=1) & (o["status"] == 0)
            assert same.mean() > 0.9
            err = np.abs(r.tau[same] - o["tau"][same]).max(axis=1) / np.abs(o["tau"][same]).max(axis=1)
            assert err.max() < 1e-8
            del bt


def test_two_handles_driven_from_two_host_threads(icub, talos_fast, talos, tiago):
    """Handles share no device state (shape tables are kernel parameters, the problem and the plans are the handle's own
    global-memory copies): batches of different robots -- iCub and the fast-URDF Talos even share one compiled kernel pair --
    driven concurrently from host threads on ONE device, with contact masks changing between calls, give bit-identical results
    to serial use."""
    import threading
    jobs = []
    for k, P in enumerate((icub, talos_fast, talos, tiago)):
        q, v = random_states(P, 192, 40 + k, base=P.nv > P.na)
        masks = [None] if P.ncontacts == 0 else [None, 1, 2, None]
        jobs.append((P, q, v, masks))

    def run(job, out, reps):
        P, q, v, masks = job
        bt = iw.Batch(P, 0, 192)
        for _ in range(reps):
            out.append([bt.update(q, v, contact_mask=m) for m in masks])

    serial = []
    for job in jobs:
        res = []
        run(job, res, 1)
        serial.append(res[0])
    outs = [[] for _ in jobs]
    threads = [threading.Thread(target=run, args=(job, out, 6)) for job, out in zip(jobs, outs)]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    for want, got in zip(serial, outs):
        assert len(got) == 6
        for rep in got:
            for a, b in zip(want, rep):
                for k in ("tau", "dv", "f", "q_next", "v_next", "status", "iters", "active"):
                    assert np.array_equal(getattr(a, k), getattr(b, k)), k


def test_one_call_over_several_shards(talos, franka):
    """iwbc_multi_update: one host call, contiguous shards, one worker thread per shard.  On a one-GPU box the shards sit on the same
    device (ordinal repeated): results must be bit-identical to the single-handle call, in instance order, for mixed contact masks too."""
    for P, B in ((talos, 3001), (franka, 5000)):
        q, v = random_states(P, B, 17, base=P.nv > P.na)
        mask = None if P.ncontacts == 0 else np.random.default_rng(2).choice(np.array([3, 1, 2], dtype=np.uint32), B)
        one = iw.Batch(P, 0, B).update(q, v, contact_mask=mask)
        for ndev in (1, 3):
            mb = iw.MultiBatch(P, [0] * ndev, B)
            got = mb.update(q, v, contact_mask=mask)
            for k in ("tau", "dv", "f", "q_next", "v_next", "status", "iters", "active"):
                assert np.array_equal(getattr(got, k), getattr(one, k)), (k, ndev)
            assert mb.launch_count > 0
        with pytest.raises(iw.IwbcError, match="capacity"):
            iw.MultiBatch(P, [0, 0], 64).update(q, v)
