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
