"""Times BASELINE configs 1, 2, 4, 5 through the public API on one GPU (config 3 is bench.py)."""
import json, os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import astropaint_b200 as ap
from astropaint_b200 import synthetic
from astropaint_b200.profiles import spherical, NFW, Battaglia16
ap.paint_bucket.verbose = False

def timed_spray(cat, nside, tmpl, reps=3, **kw):
    canvas = ap.Canvas(cat, nside, R_times=5)
    p = ap.Painter(tmpl)
    p.spray(canvas, **kw)  # warm-up (uploads columns)
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        canvas.clean()
        torch.cuda.synchronize(); t0 = time.perf_counter()
        p.spray(canvas, **kw)
        torch.cuda.synchronize(); ts.append(time.perf_counter() - t0)
    n = int(p.last_spray["d_count"].item())
    return canvas, {"halo_pixels": n, "ms": 1e3 * min(ts), "halo_pixels_per_s": n / min(ts), "path": p.last_spray["path"]}

out = {}
cat = ap.Catalog(synthetic.random_box(1000, seed=0))
_, out["C1 1e3 random box, Gaussian, nside 512"] = timed_spray(cat, 512, spherical.gaussian_2D)
cat = ap.Catalog(synthetic.websky_like(1_000_000, seed=1))
_, out["C2 1e6 Websky-shaped, Battaglia16.tau_2D_interp, nside 4096"] = timed_spray(cat, 4096, Battaglia16.tau_2D_interp)
canvas, out["C4 1e6 Websky-shaped, NFW.BG, nside 8192"] = timed_spray(cat, 8192, NFW.BG)
_, out["extra 1e6 Websky-shaped, NFW.kSZ_T, nside 8192"] = timed_spray(cat, 8192, NFW.kSZ_T)
# C5: stack 1e5 cutouts 0.4 x 0.4 deg, 200 x 200 from the BG map
halos = np.arange(100000)
canvas.stack_cutouts(halos[:1000], lon_range=[-0.2, 0.2], lat_range=[-0.2, 0.2], xpix=200)
torch.cuda.synchronize(); t0 = time.perf_counter()
canvas.stack_cutouts(halos, lon_range=[-0.2, 0.2], lat_range=[-0.2, 0.2], xpix=200)
torch.cuda.synchronize(); dt = time.perf_counter() - t0
out["C5 stack 1e5 cutouts 200x200 from nside 8192"] = {"samples": 4e9, "ms": 1e3 * dt, "cutouts_per_s": 1e5 / dt,
                                                        "nominal_gather_GBps": 32e9 / dt / 1e9}
print(json.dumps(out, indent=1))
