"""Multi-GPU plumbing: one process per GPU (torchrun), work sharded by flux-row range or by scan-grid row, one small
merge of the rates / histograms / rate maps: a single libalpk kernel per rank over NVLink peer memory (peer.py,
csrc/alpk_peer.cu) when the ranks share a node, torch.distributed's all-reduce / all-gather otherwise (NCCL; gloo in the
CPU tests).

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


def shard_kept_rows(table, keep, rank, world):
    """Balanced sharding of a flux table whose rows are filtered before sampling (the per-row `return` of
    simulate_single, fluxes.py:136, 214-216): the KEPT rows are split into contiguous, equally sized blocks and rank r gets
    the slice of the table that spans its block.  Rows of that slice that fail the filter are dropped again by the flux
    class itself, so the rank simulates exactly its kept rows; low-energy rows under the production threshold no longer
    make the first ranks idle.  Returns (table slice, global index of its first row); the slice is empty when this rank
    owns no kept row.  Concatenating the ranks' outputs in rank order reproduces the single-process sample order."""
    table = np.asarray(table)
    idx = np.flatnonzero(np.asarray(keep, dtype=bool))
    lo, hi = shard_range(idx.shape[0], rank, world)
    if hi <= lo:
        return table[0:0], 0
    first, last = int(idx[lo]), int(idx[hi - 1])
    return np.ascontiguousarray(table[first:last + 1]), first


def shard_flux(flux, rank=None, nranks=None, group=None):
    """Make `flux` (a flux object with a row-filter rule, before simulate()) this rank's shard of a single job: its flux
    table is replaced by the rank's block of kept rows (shard_kept_rows) and `row_offset` set so that the Philox keys stay
    the global row indices.  Returns the flux.  Every flux row is independent (fluxes.py:150-151, 232-233), so the
    concatenation of the ranks' sample arrays equals the single-process arrays bit for bit."""
    if rank is None or nranks is None:
        rank, nranks = world(group)
    name = flux._TABLE_ATTR
    table = np.asarray(getattr(flux, name), dtype=np.float64)
    sub, off = shard_kept_rows(table, flux._row_keep(table), rank, nranks)
    setattr(flux, name, sub)
    flux.row_offset = int(getattr(flux, "row_offset", 0)) + off
    return flux


def world(group=None):
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(group), dist.get_world_size(group)
    return 0, 1


def _host_staged(t, group):
    """gloo moves host memory: with that backend (CPU tests, or more ranks than GPUs) CUDA tensors are staged through the
    host; NCCL takes the device tensor as it is."""
    import torch.distributed as dist
    return t.is_cuda and dist.get_backend(group) == "gloo"


def all_reduce_sum(t, group=None):
    """In-place sum over ranks (no-op in a single process)."""
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        if _host_staged(t, group):
            h = t.cpu()
            dist.all_reduce(h, group=group)
            t.copy_(h)
        else:
            dist.all_reduce(t, group=group)
    return t


def merged_rates_and_hists(rates, hists=(), group=None):
    """The path's one exchange step: total rates and binned spectra of all ranks merged with ONE all-reduce
    over the packed buffer [rate_0, rate_1, ... || hist_0 || hist_1 ...].  On one node this is a single kernel per rank
    over NVLink peer memory (peer.PeerGroup.all_reduce_sum: the segments are packed, pushed, and summed in rank order by the
    kernel itself -- every rank gets the same bits); otherwise one NCCL all-reduce of the concatenated buffer.
    rates: 1-element float64 tensors; hists: float64 tensors.  Returns (rates tensor [n_rates], list of merged hists) --
    views of the packed buffer, on the tensors' device."""
    from .peer import peer_group
    parts = [r.reshape(-1) for r in rates] + [h.reshape(-1) for h in hists]
    pg = peer_group(group, parts[0])
    if pg is not None and pg.fits_reduce(sum(int(p.numel()) for p in parts)):
        buf = pg.all_reduce_sum([p.contiguous() for p in parts])
    else:
        buf = torch.cat(parts) if len(parts) > 1 else parts[0].clone()
        all_reduce_sum(buf, group)
    n_r = len(rates)
    out_h, pos = [], n_r
    for h in hists:
        out_h.append(buf[pos:pos + h.numel()].view(h.shape))
        pos += h.numel()
    return buf[:n_r], out_h


def gather_row_blocks(local, counts, group=None):
    """All-gather of per-rank row blocks with known per-rank row counts (disjoint ownership, e.g. scan maps sharded by
    mass): returns the list of every rank's block.  One NCCL all-gather -- no zero-filled reduce."""
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return [local]
    if _host_staged(local, group):
        return [b.to(local.device) for b in gather_row_blocks(local.cpu(), counts, group)]
    tail = tuple(local.shape[1:])
    outs = [torch.empty((int(c),) + tail, dtype=local.dtype, device=local.device) for c in counts]
    if len(set(int(c) for c in counts)) == 1:
        dist.all_gather(outs, local.contiguous(), group=group)
    else:                                               # ragged blocks: pad to the largest (blocks differ by <= 1 row)
        m = max(int(c) for c in counts)
        pad = torch.zeros((m,) + tail, dtype=local.dtype, device=local.device)
        pad[:local.shape[0]] = local
        tmp = [torch.empty_like(pad) for _ in counts]
        dist.all_gather(tmp, pad, group=group)
        outs = [t[:int(c)] for t, c in zip(tmp, counts)]
    return outs


def merge_row_blocks(local, lo, n_total, group=None):
    """Assemble a [n_total, ...] tensor from per-rank contiguous row blocks (`local` = rows [lo, lo+len)) with one
    all-reduce (disjoint blocks: the sum is a gather).  Used for scan maps sharded by mass."""
    shape = (int(n_total),) + tuple(local.shape[1:])
    full = torch.zeros(shape, dtype=local.dtype, device=local.device)
    if local.shape[0]:
        full[lo:lo + local.shape[0]] = local
    return all_reduce_sum(full, group)
