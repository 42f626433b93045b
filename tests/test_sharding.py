"""N > 1 host logic on CPU: world_size-2 gloo run of the sharding / reduction code bench.py uses over NCCL."""
import os
import subprocess
import sys

from inria_wbc_b200.sharding import shard_range

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_shard_ranges_partition_the_batch():
    for total in (0, 1, 7, 16384, 1000003):
        for world in (1, 2, 4, 8):
            r = [shard_range(total, g, world) for g in range(world)]
            assert r[0][0] == 0 and r[-1][1] == total
            assert all(r[g][1] == r[g + 1][0] for g in range(world - 1))
            sizes = [b - a for a, b in r]
            assert max(sizes) - min(sizes) <= 1


def test_two_rank_gloo_run():
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", OMP_NUM_THREADS="1")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
           "--master-port", "29431", os.path.join(ROOT, "tests", "_gloo_worker.py")]
    p = subprocess.run(cmd, env=env, capture_output=True, text=True, timeout=240)
    assert p.returncode == 0, p.stdout[-2000:] + p.stderr[-2000:]
    assert "GLOO_OK" in p.stdout
