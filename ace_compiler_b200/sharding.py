"""Partitioning of independent work units (encrypted images / ciphertexts) across ranks.

The reference's only parallelism is an OpenMP loop over images sharing one read-only context
(fhe-cmplr/rtlib/ant/dataset/resnet_cifar.main.inc:81); here image i goes to GPU i mod G and
no data-path collective exists.  torch.distributed only carries the barrier and the
max-over-ranks reduction of the device timings."""


def shard_units(total, rank, world):
    """indices of the units rank `rank` of `world` processes (round-robin)"""
    if not 0 <= rank < world:
        raise ValueError("rank out of range")
    return list(range(rank, total, world))


def max_over_ranks(value, dist=None, device=None):
    """max of a python float over all ranks (identity when not distributed)"""
    if dist is None or not dist.is_initialized():
        return value
    import torch
    t = torch.tensor([value], dtype=torch.float64, device=device or "cpu")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())
