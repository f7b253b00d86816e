"""Canvas(inclusive=True) against the default exclusive query: 1e6 Websky-shaped halos, NFW.kSZ_T, nside 8192."""
import json, os, sys, time
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import astropaint_b200 as ap
from astropaint_b200 import synthetic
from astropaint_b200.profiles import NFW
ap.paint_bucket.verbose = False
cat = ap.Catalog(synthetic.websky_like(1_000_000, seed=1))
out = {}
for inclusive in (False, True):
    canvas = ap.Canvas(cat, 8192, R_times=5, inclusive=inclusive)
    p = ap.Painter(NFW.kSZ_T)
    p.spray(canvas)
    torch.cuda.synchronize()
    ts = []
    for _ in range(3):
        canvas.clean()
        torch.cuda.synchronize(); t0 = time.perf_counter()
        p.spray(canvas)
        torch.cuda.synchronize(); ts.append(time.perf_counter() - t0)
    n = int(p.last_spray["d_count"].item())
    out["inclusive" if inclusive else "exclusive"] = {"halo_pixels": n, "ms": 1e3 * min(ts), "halo_pixels_per_s": n / min(ts)}
print(json.dumps(out))
