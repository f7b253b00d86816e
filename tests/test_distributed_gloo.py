"""world_size-2 gloo tests (CPU) of the N>1 host logic.

1. The host-template path: shard the catalog in np.array_split blocks, paint each shard (here with the oracle,
   since there is no GPU), all-reduce the partial maps and compare with the serial map.
2. Ring-band ownership: every rank paints (oracle) the halos its _bands.Plan selects, keeps only the pixels of
   its own band and writes them into the node's SharedHostMap; the assembled map must equal the serial one, and
   parallel.all_agree / the collective buffer hand-over must behave."""
import os
import socket

import numpy as np


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out_dir):
    import torch
    import torch.distributed as dist
    import astropaint_b200 as ap
    from astropaint_b200 import parallel, synthetic
    from oracle import reference_loop as ref, profiles_np as oprof
    ap.paint_bucket.verbose = False
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    assert parallel.world() == (rank, world)
    nside = 32
    cat = ap.Catalog(synthetic.random_box(41, seed=7))
    rows = parallel.shard_rows(cat.size, rank, world)
    partial = np.zeros(12 * nside * nside)
    ref.spray(partial, cat.data, nside, oprof.gaussian_readme, R_times=3, halo_list=rows)
    t = torch.from_numpy(partial)
    parallel.reduce_partial_map(t, "all")
    np.save(os.path.join(out_dir, f"rank{rank}.npy"), t.numpy())
    dist.destroy_process_group()


def test_sharded_paint_and_allreduce_equals_serial(tmp_path):
    import torch.multiprocessing as mp
    import astropaint_b200 as ap
    from astropaint_b200 import synthetic
    from oracle import reference_loop as ref, profiles_np as oprof
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    nside = 32
    cat = ap.Catalog(synthetic.random_box(41, seed=7))
    serial = np.zeros(12 * nside * nside)
    ref.spray(serial, cat.data, nside, oprof.gaussian_readme, R_times=3)
    for r in range(world):
        got = np.load(tmp_path / f"rank{r}.npy")
        assert np.abs(got - serial).max() <= 1e-12 * np.abs(serial).max()


def _band_worker(rank, world, port, out_dir):
    import torch.distributed as dist
    import astropaint_b200 as ap
    from astropaint_b200 import parallel, synthetic, _bands
    from oracle import reference_loop as ref, profiles_np as oprof
    ap.paint_bucket.verbose = False
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    nside, R_times = 32, 3
    cat = ap.Catalog(synthetic.random_box(41, seed=7))
    band = _bands.Band(nside, rank, world)
    order = cat.theta_order()[0]
    plan = _bands.Plan(band, cat.disc_reach(R_times), n_chunks=2)
    rows = order[plan.halo_lo:plan.halo_hi]
    full = np.zeros(12 * nside * nside)
    ref.spray(full, cat.data, nside, oprof.gaussian_readme, R_times=R_times, halo_list=rows)
    host = _bands.SharedHostMap(12 * nside * nside, np.float64, band, register=False)     # collective
    host.array[band.pix_lo:band.pix_hi] = full[band.pix_lo:band.pix_hi]                   # the rank's own rings only
    parallel.barrier()
    np.save(os.path.join(out_dir, f"band_rank{rank}.npy"), np.array(host.array))
    # agreement helper: true only if every rank says so
    assert parallel.all_agree(True) is True
    assert parallel.all_agree(rank != 0) is False
    held = host.array[:4]                         # a view keeps the buffer "handed out"
    assert host.handed_out()
    del held
    assert not host.handed_out()
    dist.destroy_process_group()


def test_band_owned_paint_into_shared_host_map_equals_serial(tmp_path):
    import torch.multiprocessing as mp
    import astropaint_b200 as ap
    from astropaint_b200 import synthetic
    from oracle import reference_loop as ref, profiles_np as oprof
    world = 2
    mp.spawn(_band_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    nside = 32
    cat = ap.Catalog(synthetic.random_box(41, seed=7))
    serial = np.zeros(12 * nside * nside)
    ref.spray(serial, cat.data, nside, oprof.gaussian_readme, R_times=3)
    for r in range(world):
        got = np.load(tmp_path / f"band_rank{r}.npy")
        assert np.abs(got - serial).max() <= 1e-12 * np.abs(serial).max()
