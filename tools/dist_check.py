"""N-rank NCCL check (run under torchrun): the band-owned Painter.spray of N ranks == a single-rank spray.

    python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 tools/dist_check.py

For every case each rank paints ITS ring band (parallel.py, _bands.py); `canvas.pixels` assembles the bands in
the node's shared host map and `canvas.pixels_device` gathers them over NVLink.  The comparison map is painted
by every rank alone (Canvas._band_override = the one-rank layout).  Prints one "OK" line per rank and case;
tests/test_gpu_bands.py::test_two_rank_spray_equals_single_rank runs this when two GPUs are visible."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import astropaint_b200 as ap  # noqa: E402
from astropaint_b200 import synthetic, parallel, _bands  # noqa: E402
from astropaint_b200.profiles import NFW, Battaglia16  # noqa: E402

ap.paint_bucket.verbose = False
local = int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
rank, ws = parallel.world()


def py_template(R, R_200c, v_r):
    return v_r * np.exp(-(R / R_200c) ** 2)


cat = ap.Catalog(synthetic.websky_like(20000, seed=3))
small = ap.Catalog(synthetic.websky_like(300, seed=4, m_min=2e14))
ap.Painter.PREFETCH_MIN_HALOS = 2000
_bands.MIN_CHUNK_HALOS = 1000
CASES = [("NFW.kSZ_T", NFW.kSZ_T, cat, 2048, False), ("NFW.kSZ_T readback-balanced bands", NFW.kSZ_T, cat, 2048, "readback"),
         ("Battaglia16.kSZ_T", Battaglia16.kSZ_T, cat, 2048, False),
         ("NFW.BG", NFW.BG, cat, 2048, False), ("NFW.kSZ_T inclusive", NFW.kSZ_T, cat, 1024, True),
         ("python template (host path)", py_template, small, 512, False)]
for name, tmpl, c, nside, inclusive in CASES:
    balance = "readback" if inclusive == "readback" else "area"
    inclusive = inclusive is True
    canvas = ap.Canvas(c, nside, R_times=5, inclusive=inclusive, band_balance=balance)
    p = ap.Painter(tmpl)
    p.spray(canvas)                                      # collective: this rank paints its band
    n_painted = torch.zeros(1, dtype=torch.int64, device="cuda") + p.last_spray["d_count"]
    dist.all_reduce(n_painted)
    got_dev = canvas.pixels_device.clone()               # gathered over NVLink
    got_host = canvas.pixels.copy()                      # bands assembled in the shared host map
    canvas.clean()
    p.spray(canvas)                                      # second round: streams into the reused host map
    again = canvas.pixels.copy()
    # the same spray by this rank alone
    c2 = ap.Canvas(c, nside, R_times=5, inclusive=inclusive)
    c2._band_override = _bands.Band(nside, 0, 1)
    p2 = ap.Painter(tmpl)
    p2.spray(c2)
    want = c2._dmap
    scale = float(want.abs().max())
    err_dev = float((got_dev - want).abs().max()) / scale
    want_h = want.cpu().numpy()
    err_host = float(np.abs(got_host - want_h).max()) / scale
    err_again = float(np.abs(again - want_h).max()) / scale
    n_want = int(p2.last_spray["d_count"].item())
    ok = err_dev < 1e-12 and err_host < 1e-12 and err_again < 1e-12 and int(n_painted.item()) == n_want and scale > 0
    print(f"rank {rank}/{ws} {name}: path {p.last_spray['path']}, device {err_dev:.2e} host {err_host:.2e} "
          f"streamed {err_again:.2e}, pairs {int(n_painted.item())} vs {n_want}: {'OK' if ok else 'MISMATCH'}", flush=True)
    assert ok
# stack_cutouts: each rank stacks its share of the halo list from the gathered map; the stacks are summed over NVLink
canvas = ap.Canvas(cat, 2048, R_times=5)
ap.Painter(NFW.kSZ_T).spray(canvas)
got = canvas.stack_cutouts(halo_list=np.arange(500), lon_range=[-0.5, 0.5], xpix=64, inplace=False)
c2 = ap.Canvas(cat, 2048, R_times=5)
c2._band_override = _bands.Band(2048, 0, 1)
ap.Painter(NFW.kSZ_T).spray(c2)
d_lon = torch.from_numpy(cat.data["lon"].values[:500].copy()).cuda()
d_lat = torch.from_numpy(cat.data["lat"].values[:500].copy()).cuda()
from astropaint_b200 import _engine as eng  # noqa: E402
want = eng.stack_cutouts(c2._dmap, 2048, d_lon, d_lat, [-0.5, 0.5], [-0.5, 0.5], 64, 64).cpu().numpy()
err = float(np.abs(got - want).max() / np.abs(want).max())
print(f"rank {rank}/{ws} stack_cutouts sharded over ranks: {err:.2e}: {'OK' if err < 1e-12 else 'MISMATCH'}", flush=True)
assert err < 1e-12
dist.barrier()
dist.destroy_process_group()
