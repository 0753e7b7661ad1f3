"""Multi-GPU sharding of a line corpus: byte-balanced contiguous ranges of lines,
one per rank, and the final gather of per-rank match counts.

The hot path has no exchange step: every rank matches its own range with its own
copy of the program (the reference's matching is read-only and shared-nothing
as well, doc/Performance.md "Threads").  The only communication is the gather
of counts (and, if wanted, of the matched line indices) at the end, which goes
through torch.distributed - NCCL on GPUs, gloo in the CPU tests.
"""

import numpy as np


def byte_balanced_ranges(lens, world):
    """Cut lines [0, n) into `world` contiguous ranges of (nearly) equal bytes.

    lens: per-line byte counts.  Returns a list of (first_line, end_line).
    Every line belongs to exactly one range; ranges are in order; a range may
    be empty when there are fewer lines than ranks."""
    lens = np.asarray(lens, dtype=np.uint64)
    n = len(lens)
    if world <= 1 or n == 0:
        return [(0, n)] + [(n, n)] * (max(world, 1) - 1)
    csum = np.cumsum(lens, dtype=np.uint64)
    total = int(csum[-1])
    cuts = [0]
    for r in range(1, world):
        target = (total * r) // world
        # first line whose end lies beyond the target
        k = int(np.searchsorted(csum, target, side="left"))
        cuts.append(min(max(k, cuts[-1]), n))
    cuts.append(n)
    return [(cuts[i], cuts[i + 1]) for i in range(world)]


def byte_balanced_ranges_periodic(pool_lens, reps, world):
    """byte_balanced_ranges(np.tile(pool_lens, reps), world) without building
    the tiled array: bench.py's corpus is a pool of lines repeated `reps`
    times, hundreds of millions of lines at eight ranks."""
    pool_lens = np.asarray(pool_lens, dtype=np.uint64)
    m = len(pool_lens)
    n = m * reps
    if world <= 1 or n == 0:
        return [(0, n)] + [(n, n)] * (max(world, 1) - 1)
    csum = np.cumsum(pool_lens, dtype=np.uint64)
    per_pool = int(csum[-1])
    total = per_pool * reps
    cuts = [0]
    for r in range(1, world):
        target = (total * r) // world
        if target == 0 or per_pool == 0:
            k = int(np.searchsorted(csum, target, side="left")) if per_pool else 0
        else:
            # the first line whose end reaches the target lies in pool copy q
            q = (target - 1) // per_pool
            k = q * m + int(np.searchsorted(csum, target - q * per_pool, side="left"))
        cuts.append(min(max(k, cuts[-1]), n))
    cuts.append(n)
    return [(cuts[i], cuts[i + 1]) for i in range(world)]


def gather_counts(local_counts, group=None):
    """All ranks get every rank's vector of counts (e.g. matches per result id).

    local_counts: 1-D torch int64 tensor, same length on every rank, on the
    device the process group works with."""
    import torch.distributed as dist
    world = dist.get_world_size(group)
    out = [local_counts.new_empty(local_counts.shape) for _ in range(world)]
    dist.all_gather(out, local_counts, group=group)
    return out


def gather_matched_lines(local_indices, first_line, group=None):
    """Global indices of matched lines from all ranks, in line order.

    local_indices: 1-D int64 tensor of indices relative to this rank's range."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group)
    n = torch.tensor([local_indices.numel()], dtype=torch.int64, device=local_indices.device)
    sizes = [torch.zeros_like(n) for _ in range(world)]
    dist.all_gather(sizes, n, group=group)
    cap = int(max(int(s.item()) for s in sizes))
    padded = torch.full((max(cap, 1),), -1, dtype=torch.int64, device=local_indices.device)
    padded[:local_indices.numel()] = local_indices + first_line
    parts = [torch.empty_like(padded) for _ in range(world)]
    dist.all_gather(parts, padded, group=group)
    return torch.cat([p[:int(s.item())] for p, s in zip(parts, sizes)])
