"""Multi-GPU plumbing of the batched cycle: the batch shards by instance range, one process per GPU,
no collective on the data path.  torch.distributed is only used for the barrier and
for the max-over-ranks / sum-over-ranks of the timing and counters."""


from ._lib import shard_range  # noqa: F401  (the library's own shard arithmetic: iwbc_shard_range, the one iwbc_multi_update uses)


def reduce_timing(dist, device, step_ms_total, counts):
    """max over ranks of the elapsed time, sum over ranks of the counters (units processed etc.)."""
    import torch
    t = torch.tensor([float(step_ms_total)], dtype=torch.float64, device=device)
    c = torch.tensor([float(x) for x in counts], dtype=torch.float64, device=device)
    if dist is not None and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(c, op=dist.ReduceOp.SUM)
    return float(t[0]), [float(x) for x in c]
