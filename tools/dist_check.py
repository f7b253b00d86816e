"""2-rank NCCL check: sharded Painter.spray == single-rank spray (run under torchrun)."""
import os, sys
import numpy as np, torch, torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import astropaint_b200 as ap
from astropaint_b200 import synthetic, parallel
from astropaint_b200.profiles import NFW, Battaglia16
ap.paint_bucket.verbose = False
local = int(os.environ["LOCAL_RANK"]); torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
rank, ws = parallel.world()
cat = ap.Catalog(synthetic.websky_like(20000, seed=3))
for tmpl, nside in [(NFW.kSZ_T, 2048), (Battaglia16.kSZ_T, 2048)]:
    canvas = ap.Canvas(cat, nside, R_times=5)
    p = ap.Painter(tmpl); p.spray(canvas)           # sharded: world_size > 1
    got = canvas.pixels_device.clone()
    # serial on this rank alone, bypassing the sharding
    c2 = ap.Canvas(cat, nside, R_times=5)
    R_mode = p.template_args_list[p.R_arg_indx]
    hc = p._shake_canister(cat, {})
    p._paint_rows(c2, None, R_mode, hc, {}, c2.pixels_device)
    want = c2._dmap
    err = float((got - want).abs().max() / want.abs().max())
    print(f"rank {rank}/{ws} {tmpl.__name__}: sharded vs serial max rel err {err:.3e}, painted {int(p.last_spray['d_count'].item())}", flush=True)
    assert err < 1e-12
dist.destroy_process_group()
