import os, sys, time, json
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import astropaint_b200 as ap
from astropaint_b200 import synthetic, _engine as eng
from astropaint_b200.profiles import NFW
ap.paint_bucket.verbose = False
cat = ap.Catalog(synthetic.websky_like(1_000_000, seed=1))
canvas = ap.Canvas(cat, 8192, R_times=5)
ap.Painter(NFW.kSZ_T).spray(canvas)
d_map = canvas.pixels_device
n = 100000
lon = eng.to_device(cat.data.lon.values[:n], d_map.device); lat = eng.to_device(cat.data.lat.values[:n], d_map.device)
for _ in range(2): eng.stack_cutouts(d_map, 8192, lon, lat, [-0.2, 0.2], [-0.2, 0.2], 200, 200)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(3): st = eng.stack_cutouts(d_map, 8192, lon, lat, [-0.2, 0.2], [-0.2, 0.2], 200, 200)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 3
print(json.dumps({"stack_ms": ms, "cutouts_per_s": n / ms * 1e3, "samples_per_s": n * 4e4 / ms * 1e3, "nominal_GBps": n * 4e4 * 8 / ms / 1e6, "frac_hbm": n * 4e4 * 8 / ms / 1e6 / 6553}))

# oriented stacking through the public API: rotate each cutout along its transverse velocity, taper, flip by the
# sign of v_r (Birkinshaw_Gull_stacking.ipynb / parallel_stacking.ipynb); device chain vs the host loop (scipy)
from astropaint_b200 import transform
halos = np.arange(n)
angle = np.rad2deg(np.arctan2(cat.data.v_th.values, cat.data.v_ph.values))
sign = -np.sign(cat.data.v_r.values)
funcs = [transform.rotate_patch, transform.taper_patch, transform.scale_patch]
kw = lambda: [{"angle": angle}, {}, {"weight": sign}]
canvas.stack_cutouts(halos[:2000], lon_range=[-0.2, 0.2], xpix=200, apply_func=funcs, func_kwargs=kw())
torch.cuda.synchronize(); t0 = time.perf_counter()
canvas.stack_cutouts(halos, lon_range=[-0.2, 0.2], xpix=200, apply_func=funcs, func_kwargs=kw())
torch.cuda.synchronize(); dev_s = time.perf_counter() - t0
m = 200
host = [lambda p, angle: transform.rotate_patch(p, angle), lambda p: transform.taper_patch(p), lambda p, weight: p * weight]
t0 = time.perf_counter()
canvas.stack_cutouts(halos[:m], lon_range=[-0.2, 0.2], xpix=200, apply_func=host, func_kwargs=kw())
host_s = (time.perf_counter() - t0) / m * n
print(json.dumps({"oriented_stack_1e5_device_s": dev_s, "oriented_stack_1e5_host_loop_s_extrapolated": host_s,
                  "host_sample": m}))
