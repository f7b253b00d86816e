"""Catalog.build_dataframe (paint_bucket.py:308-364) for 1e7 halos: numpy/pandas on the host against the device
kernel (ap_catalog_build), raw columns in host memory both times."""
import json, os, sys, time
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import astropaint_b200 as ap
from astropaint_b200 import synthetic
ap.paint_bucket.verbose = False
n = int(sys.argv[1]) if len(sys.argv) > 1 else 10_000_000
raw = synthetic.websky_like(n, seed=1)
out = {"n": n}
for device in (True, False, True):
    df = raw.copy()
    torch.cuda.synchronize(); t0 = time.perf_counter()
    cat = ap.Catalog(df, device=device)
    torch.cuda.synchronize(); out["device_s" if device else "host_s"] = time.perf_counter() - t0
    # kernel alone
print(json.dumps(out))
