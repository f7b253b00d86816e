"""Multi-GPU painting: one process per GPU (torch.distributed, NCCL over NVLink 5), ring-band ownership.

The reference's only parallelism is `np.array_split(range(N), n_cpus)` batches of halos painted by joblib
threads into one shared array (paint_bucket.py:2402-2433).  Here every rank OWNS one ring band of the map
(_bands.Band: an equal share of the sphere's area, contiguous in RING order) and paints the halos whose disc
touches it, clipped to its own rings -- no partial maps and no reduction.  The host copy
of the map is one array shared by the ranks of the node (_bands.SharedHostMap); each rank copies its band into
it over its own PCIe link.

Under world_size > 1 the Canvas / Painter calls are SPMD-collective: every rank makes the same calls on the
same catalog.  Templates that run on the host (arbitrary Python callables) keep the reference's halo batches:
rank r paints rows np.array_split(range(N), world)[r] into a full-sky partial map, the partial maps are summed
(NCCL all-reduce) and every rank adds its own band of the sum.
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
    """Catalog rows painted by `rank` on the host-template path: contiguous blocks, as
    np.array_split(range(N), n) gives (paint_bucket.py:2414)."""
    return np.array_split(np.arange(n_halos), world_size)[rank]


def reduce_partial_map(partial, op="all"):
    """Sum the partial maps of all ranks in place (tensor on the rank's device, or CPU with gloo)."""
    import torch.distributed as dist
    if op == "all":
        dist.all_reduce(partial, op=dist.ReduceOp.SUM)
    else:
        dist.reduce(partial, dst=int(op), op=dist.ReduceOp.SUM)
    return partial


def all_agree(flag):
    """True iff `flag` is true on every rank (one tiny all-reduce; the identity when not distributed)."""
    rank, ws = world()
    if ws == 1:
        return bool(flag)
    import torch
    import torch.distributed as dist
    dev = "cuda" if dist.get_backend() == "nccl" else "cpu"
    t = torch.tensor([1 if flag else 0], dtype=torch.int32, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MIN)
    return bool(int(t.item()))


def barrier():
    rank, ws = world()
    if ws > 1:
        import torch.distributed as dist
        dist.barrier()


_READBACK_WEIGHTS = {}


def readback_weights(megabytes=256, reps=3):
    """Device->host bandwidth of every rank's GPU while ALL ranks copy at once [GB/s], the same list on every rank
    (collective, measured once per process group and cached).  On a multi-GPU host the PCIe paths are not equal --
    on the 8xB200 box of the bench the four GPUs of one half reach 11.7 GB/s each and the other four 18.5 GB/s
    when all eight copy -- so bands sized by these numbers (Canvas(band_balance="readback")) finish their read-back
    together, where equal-area bands wait for the slowest link."""
    import torch
    import torch.distributed as dist
    rank, ws = world()
    if ws == 1:
        return [1.0]
    if ws in _READBACK_WEIGHTS:
        return _READBACK_WEIGHTS[ws]
    from . import _engine as eng
    n = megabytes * (1 << 20) // 8
    d = torch.zeros(n, dtype=torch.float64, device="cuda")
    h = torch.empty(n, dtype=torch.float64, pin_memory=True)
    seen = []
    for _ in range(reps + 1):
        dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        eng.copy_d2h(h, d)
        e1.record()
        e1.synchronize()
        seen.append(8.0 * n / (e0.elapsed_time(e1) * 1e-3) / 1e9)
    t = torch.zeros(ws, dtype=torch.float64, device="cuda")
    t[rank] = float(np.median(seen[1:]))       # the first copy warms the path up
    dist.all_reduce(t)
    _READBACK_WEIGHTS[ws] = [float(x) for x in t.cpu()]
    return _READBACK_WEIGHTS[ws]


def gather_bands(d_band, band, dtype):
    """Full map on every rank's device from the owners' bands: one broadcast per band over NVLink."""
    import torch
    import torch.distributed as dist
    full = torch.empty(12 * band.nside * band.nside, dtype=dtype, device=d_band.device)
    full[band.pix_lo:band.pix_hi].copy_(d_band)
    for g in range(band.world_size):
        lo, hi = band.pix_bounds[g], band.pix_bounds[g + 1]
        if hi > lo:
            dist.broadcast(full[lo:hi], src=g)
    return full
