"""world_size-2 gloo test (CPU) of the N>1 host logic: shard the catalog like spray_sharded does,
paint each shard (here with the oracle, since there is no GPU), all-reduce the partial maps and
compare with the serial map."""
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
