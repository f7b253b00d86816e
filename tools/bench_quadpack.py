"""Times the QUADPACK node kernel (ap_los_nodes, kind b16_tau) alone on a Websky-shaped catalog and
writes the node values so that two builds of the library can be compared bit for bit:

    AP_LIB_PATH=/path/to/variant.so python tools/bench_quadpack.py --n 2000000 --out gpurun_out/q_variant.npy

prints {"ms": ..., "ms_per_1e7_halos": ..., "checksum": ...}; with --ref other.npy also the largest relative
difference against a previous run's values."""
import argparse, json, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from astropaint_b200 import _engine as eng, synthetic
import astropaint_b200 as ap

ap.paint_bucket.verbose = False
a = argparse.ArgumentParser()
a.add_argument("--n", type=int, default=2_000_000)
a.add_argument("--nside", type=int, default=8192)
a.add_argument("--reps", type=int, default=3)
a.add_argument("--out", default=None)
a.add_argument("--ref", default=None)
args = a.parse_args()

dev = torch.device("cuda", 0)
cat = ap.Catalog(synthetic.websky_like(args.n, seed=1))
halos = eng.DeviceHalos(cat.data, 5, dev)
cols = [eng.to_device(cat.data[c].to_numpy(), dev) for c in ("R_200c", "M_200c", "redshift")]
n_pix, _, rmin, rmax = eng.disc_count(halos, args.nside, want_range=True)
nodes, n_nodes = eng.table_nodes(halos, n_pix, rmin, rmax, 15, "linspace")
values = eng.los_nodes("b16_tau", halos, cols, nodes, n_nodes)  # warm-up
torch.cuda.synchronize()
ts = []
for _ in range(args.reps):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    values = eng.los_nodes("b16_tau", halos, cols, nodes, n_nodes)
    e1.record()
    torch.cuda.synchronize()
    ts.append(e0.elapsed_time(e1))
v = values.cpu().numpy()
out = {"lib": os.path.basename(os.environ.get("AP_LIB_PATH", "default")), "n": args.n, "ms": min(ts),
       "ms_per_1e7_halos": min(ts) * 1e7 / args.n, "checksum": float(v.sum())}
if args.ref and os.path.exists(args.ref):
    r = np.load(args.ref)
    d = np.abs(v - r) / np.maximum(np.abs(r), 1e-300)
    out["max_rel_diff_vs_ref"] = float(d.max())
    out["n_rel_diff_gt_1e-9"] = int((d > 1e-9).sum())
if args.out:
    np.save(args.out, v)
print(json.dumps(out))
