"""Capture sharding across ranks and the one collective on the path.

Captures are independent: rank r of W demodulates a contiguous block of the batch with no
data-path collective; afterwards the decoded bit streams are gathered (NCCL all-gather over
NVLink on GPUs, gloo in the CPU tests)."""


def shard_range(n_captures, rank, world):
    """Contiguous block of captures owned by `rank`: sizes differ by at most one."""
    base, extra = divmod(n_captures, world)
    start = rank * base + min(rank, extra)
    return start, start + base + (1 if rank < extra else 0)


def gather_bits(dist, local_bits, n_captures, world):
    """All-gather the per-rank [n_local, L] uint8 bit tensors into [n_captures, L] (rank order).

    Ranks may own one capture fewer than others; shorter shards are padded for the collective."""
    import torch

    n_max = -(-n_captures // world)
    L = local_bits.shape[1]
    padded = torch.zeros((n_max, L), dtype=local_bits.dtype, device=local_bits.device)
    padded[: local_bits.shape[0]] = local_bits
    out = torch.empty((world * n_max, L), dtype=local_bits.dtype, device=local_bits.device)
    dist.all_gather_into_tensor(out, padded)
    parts = []
    for r in range(world):
        a, b = shard_range(n_captures, r, world)
        parts.append(out[r * n_max: r * n_max + (b - a)])
    return torch.cat(parts, dim=0)
