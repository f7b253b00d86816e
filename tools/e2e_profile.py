import os, sys, time, cProfile, pstats
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import astropaint_b200 as ap
from astropaint_b200 import synthetic
from astropaint_b200.profiles import Battaglia16
ap.paint_bucket.verbose = False
cat = ap.Catalog(synthetic.websky_like(10_000_000, seed=1)).pin_memory()
canvas = ap.Canvas(cat, 8192, R_times=5)
p = ap.Painter(Battaglia16.kSZ_T)
def step():
    canvas.clean(); p.spray(canvas); return canvas.pixels
step(); torch.cuda.synchronize()
for _ in range(2):
    t0 = time.perf_counter(); canvas.clean(); t1 = time.perf_counter(); p.spray(canvas); torch.cuda.synchronize(); t2 = time.perf_counter(); x = canvas.pixels; t3 = time.perf_counter()
    print(f"clean {1e3*(t1-t0):.1f} ms, spray(+sync) {1e3*(t2-t1):.1f} ms, pixels D2H {1e3*(t3-t2):.1f} ms ({6.442/(t3-t2):.1f} GB/s)")
pr = cProfile.Profile(); pr.enable(); step(); torch.cuda.synchronize(); pr.disable()
pstats.Stats(pr).sort_stats("cumulative").print_stats(18)
