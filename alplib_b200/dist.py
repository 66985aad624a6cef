"""Multi-GPU plumbing: one process per GPU (torchrun), work sharded by flux-row range or by scan-grid row, one small
all-reduce of the merged rates / histograms / rate maps (NCCL over NVLink on the GPU box, gloo in the CPU tests).

No per-sample data crosses the interconnect: every flux row, sample and grid point is independent, and
the Philox variates are keyed by the GLOBAL row index, so results do not depend on the number of ranks.
"""
import numpy as np
import torch


def shard_range(n, rank, world):
    """Contiguous block [lo, hi) of n items owned by `rank` of `world` (the first n % world ranks get one extra)."""
    base, extra = divmod(int(n), int(world))
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def shard_rows(table, rank, world):
    """(rows owned by this rank, global index of its first row).  Concatenating the ranks' outputs in rank order
    reproduces the single-process sample order (the reference loops rows in order, fluxes.py:150-151, 232-233)."""
    table = np.asarray(table)
    lo, hi = shard_range(table.shape[0], rank, world)
    return np.ascontiguousarray(table[lo:hi]), lo


def world(group=None):
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(group), dist.get_world_size(group)
    return 0, 1


def all_reduce_sum(t, group=None):
    """In-place sum over ranks (no-op in a single process)."""
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(t, group=group)
    return t


def merge_row_blocks(local, lo, n_total, group=None):
    """Assemble a [n_total, ...] tensor from per-rank contiguous row blocks (`local` = rows [lo, lo+len)) with one
    all-reduce (disjoint blocks: the sum is a gather).  Used for scan maps sharded by mass."""
    shape = (int(n_total),) + tuple(local.shape[1:])
    full = torch.zeros(shape, dtype=local.dtype, device=local.device)
    if local.shape[0]:
        full[lo:lo + local.shape[0]] = local
    return all_reduce_sum(full, group)
