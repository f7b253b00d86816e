"""Halo-sharded painting across the GPUs of one node.

The reference's only parallelism is `np.array_split(range(N), n_cpus)` batches of halos painted
by joblib threads into one shared array (paint_bucket.py:2402-2433).  Here each rank (one process
per GPU, torch.distributed / NCCL over NVLink 5) paints its block of the catalog into a full-sky
partial map in its own HBM and the partial maps are summed with NCCL all-reduces (fp64, npix values
in total; ring band by ring band, overlapped with the next band's QUADPACK phase, for tabulated
templates); the result is added to the canvas on every rank, so all ranks end with the same
map the single-process path would have produced (up to fp64 summation order).
"""
import numpy as np


def world():
    """(rank, world_size) of the default process group, (0, 1) when not distributed."""
    try:
        import torch.distributed as dist
        if dist.is_available() and dist.is_initialized():
            return dist.get_rank(), dist.get_world_size()
    except ImportError:
        pass
    return 0, 1


def shard_rows(n_halos, rank, world_size):
    """Catalog rows painted by `rank`: contiguous blocks, as np.array_split(range(N), n) gives."""
    return np.array_split(np.arange(n_halos), world_size)[rank]


def reduce_partial_map(partial, op="all"):
    """Sum the partial maps of all ranks in place (tensor on the rank's device, or CPU with gloo)."""
    import torch.distributed as dist
    if op == "all":
        dist.all_reduce(partial, op=dist.ReduceOp.SUM)
    else:
        dist.reduce(partial, dst=int(op), op=dist.ReduceOp.SUM)
    return partial


def spray_sharded(painter, canvas, R_mode, host_cols, template_kwargs, pinned=None):
    import torch
    rank, ws = world()
    if ws == 1:
        d_map = canvas.pixels_device
        stats = painter._paint_rows(canvas, None, R_mode, host_cols, template_kwargs, d_map, pinned=pinned)
        stats["world_size"] = 1
        return stats
    import torch.distributed as dist
    rows = shard_rows(canvas.catalog.size, rank, ws)
    base = canvas.pixels_device
    partial = torch.zeros_like(base)
    works = []

    def band_done(k, p_lo, p_hi):
        # the band is final on this rank: sum it over the ranks on NCCL's stream while the next band of
        # halos is still being integrated (table templates); other paths call this once for the whole map
        if p_hi > p_lo:
            works.append(dist.all_reduce(partial[p_lo:p_hi], op=dist.ReduceOp.SUM, async_op=True))

    stats = painter._paint_rows(canvas, rows, R_mode, host_cols, template_kwargs, partial, pinned=pinned,
                                band_done=band_done)
    for w in works:
        w.wait()
    base += partial
    stats["world_size"] = ws
    return stats
