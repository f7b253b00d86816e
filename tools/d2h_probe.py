"""Host-side topology and concurrent device->host bandwidth of the ranks of one node (run under torchrun):
page-locked torch memory vs the registered shared mapping SharedHostMap uses, with and without binding the
process to the NUMA node of its GPU before the pages are first touched."""
import glob
import json
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from astropaint_b200 import _bands, _engine as eng, _capi  # noqa: E402

world = int(os.environ.get("WORLD_SIZE", "1"))
rank = int(os.environ.get("RANK", "0"))
local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)


def bar():
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()


props = torch.cuda.get_device_properties(local)
bus = f"{props.pci_domain_id:04x}:{props.pci_bus_id:02x}:{props.pci_device_id:02x}.0"
try:
    gpu_node = int(open(f"/sys/bus/pci/devices/{bus}/numa_node").read())
except OSError:
    gpu_node = None
nodes = {}
for p in sorted(glob.glob("/sys/devices/system/node/node*/cpulist")):
    nodes[p.split("/")[-2]] = open(p).read().strip()
info = {"rank": rank, "pci": bus, "gpu_numa_node": gpu_node, "affinity": sorted(os.sched_getaffinity(0))[:4] + ["...", len(os.sched_getaffinity(0))],
        "nodes": nodes}

nside = 8192
npix = 12 * nside * nside
band = _bands.Band(nside, rank, world)
d = torch.zeros(band.npix, dtype=torch.float64, device=dev)


def timed_copy(view, reps=3):
    best = 0.0
    for _ in range(reps):
        bar()
        t0 = time.perf_counter()
        eng.copy_d2h(view, d)
        torch.cuda.synchronize()
        t1 = time.perf_counter()
        bar()
        best = max(best, 8 * band.npix / (t1 - t0) / 1e9)
    return round(best, 1)


pinned = torch.empty(band.npix, dtype=torch.float64, pin_memory=True)
info["GBps_torch_pinned"] = timed_copy(pinned)
del pinned
host = _bands.SharedHostMap(npix, np.float64, band)
info["GBps_shared_registered"] = timed_copy(host.band_view(0, band.npix))
host.close()
del host


def cpus_of(node):
    out = []
    for part in open(f"/sys/devices/system/node/node{node}/cpulist").read().strip().split(","):
        a, _, b = part.partition("-")
        out += list(range(int(a), int(b or a) + 1))
    return out


if gpu_node is not None and gpu_node >= 0:
    allowed = os.sched_getaffinity(0)
    want = set(cpus_of(gpu_node)) & allowed
    info["cpus_on_gpu_node_allowed"] = len(want)
    if want:
        os.sched_setaffinity(0, want)
        host = _bands.SharedHostMap(npix, np.float64, band)
        info["GBps_shared_registered_numa_bound"] = timed_copy(host.band_view(0, band.npix))
        pinned = torch.empty(band.npix, dtype=torch.float64, pin_memory=True)
        info["GBps_torch_pinned_numa_bound"] = timed_copy(pinned)
for r in range(world):
    if r == rank:
        print(json.dumps(info), flush=True)
    if world > 1:
        dist.barrier()
if rank == 0:
    os.system("nvidia-smi topo -m 2>&1 | head -20")
if world > 1:
    dist.destroy_process_group()
