"""world_size-2 gloo worker: each rank steps its own shard (the CPU oracle stands in for the kernel, this is a
host-logic test) and the ranks agree on max-time / summed counters exactly as bench.py does over NCCL."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import inria_wbc_b200 as iw  # noqa: E402
import oracle_lib as ol  # noqa: E402
from inria_wbc_b200.sharding import reduce_timing, shard_range  # noqa: E402
from conftest import random_states  # noqa: E402


def main():
    dist.init_process_group("gloo")
    rank, world = dist.get_rank(), dist.get_world_size()
    P = iw.Problem("franka/pos_tracker.yaml")
    total = 101
    q, v = random_states(P, total, 5, qn=0.3, vn=0.3)
    lo, hi = shard_range(total, rank, world)
    out = ol.cycle_batch(P, q[lo:hi], v[lo:hi])
    fake_ms = 10.0 + rank
    t, (n_done, checksum) = reduce_timing(dist, torch.device("cpu"), fake_ms, [hi - lo, float(out["dv"].sum())])
    full = ol.cycle_batch(P, q, v)
    if rank == 0:
        assert t == 10.0 + world - 1, t
        assert n_done == total, n_done
        assert abs(checksum - full["dv"].sum()) < 1e-9 * abs(full["dv"]).sum()
        # gather the shards back into instance order (host gather of torques / accelerations)
    parts = [None] * world
    dist.all_gather_object(parts, (lo, hi, out["tau"]))
    tau = np.zeros_like(full["tau"])
    covered = np.zeros(total, int)
    for a, b, t_ in parts:
        tau[a:b] = t_
        covered[a:b] += 1
    assert (covered == 1).all()
    assert np.array_equal(tau, full["tau"])
    dist.barrier()
    dist.destroy_process_group()
    if rank == 0:
        print("GLOO_OK")


if __name__ == "__main__":
    main()
