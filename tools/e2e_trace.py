"""Timeline of one end-to-end spray (public API, host catalog -> host map) from CUDA events: when every chunk's
kernels, the catalog uploads and the read-back copies finish relative to the start of the spray.

    python tools/e2e_trace.py [halos]                                   # one GPU
    python -m torch.distributed.run --nproc-per-node N ... tools/e2e_trace.py [halos]
"""
import json
import os
import sys
import time

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import astropaint_b200 as ap  # noqa: E402
from astropaint_b200 import synthetic  # noqa: E402
from astropaint_b200.profiles import Battaglia16  # noqa: E402

ap.paint_bucket.verbose = False
n_halos = int(float(sys.argv[1])) if len(sys.argv) > 1 else 10_000_000
world = int(os.environ.get("WORLD_SIZE", "1"))
rank = int(os.environ.get("RANK", "0"))
local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
cat = ap.Catalog(synthetic.websky_like(n_halos, seed=1))
cat.pin_memory()
canvas = ap.Canvas(cat, 8192, R_times=5, band_balance=os.environ.get("AP_BAND_BALANCE", "readback"))
painter = ap.Painter(Battaglia16.kSZ_T)


def sync():
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()


for _ in range(3):
    canvas.clean(); painter.spray(canvas); m = canvas.pixels; del m
ap.Painter.trace = True
out = []
for _ in range(3):
    sync()
    t0 = time.perf_counter()
    canvas.clean()
    t1 = time.perf_counter()
    painter.spray(canvas)
    t2 = time.perf_counter()
    m = canvas.pixels
    t3 = time.perf_counter()
    del m
    sync()
    t4 = time.perf_counter()
    tr = painter.last_spray["trace"]
    start = tr["main"][0][1]
    lanes = {lane: [(name, round(start.elapsed_time(ev), 2)) for name, ev in evs] for lane, evs in tr.items()}
    out.append({"rank": rank, "host_ms": {"clean": 1e3 * (t1 - t0), "spray_returns": 1e3 * (t2 - t0), "pixels_returns": 1e3 * (t3 - t0),
                                          "all_ranks_done": 1e3 * (t4 - t0)},
                "n_halos": painter.last_spray["n_halos"], "device_ms_since_spray_start": lanes})
for r in range(world):
    if r == rank:
        print(json.dumps(out[-1]), flush=True)
    if world > 1:
        dist.barrier()
if world > 1:
    dist.destroy_process_group()
