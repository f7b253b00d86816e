"""e2e time of the public API (Canvas.clean(); Painter(kSZ_T).spray(canvas); canvas.pixels on a pinned 1e7-halo
catalog, nside 8192) as a function of Painter.PREFETCH_CHUNKS (ring bands streamed to the host during the
QUADPACK phase) and of the area ratio of consecutive bands:  python tools/e2e_chunks.py 8:1.0 10:0.8 12:0.8"""
import os, sys, time, json
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import astropaint_b200 as ap
from astropaint_b200 import synthetic
from astropaint_b200.profiles import Battaglia16
ap.paint_bucket.verbose = False
from astropaint_b200 import _engine as eng
chunks = [a for a in sys.argv[1:]] or ["8:1.0", "10:0.8"]      # PREFETCH_CHUNKS:BAND_RATIO
cat = ap.Catalog(synthetic.websky_like(10_000_000, seed=1)).pin_memory()
canvas = ap.Canvas(cat, 8192, R_times=5)
p = ap.Painter(Battaglia16.kSZ_T)


def step():
    canvas.clean()
    p.spray(canvas)
    return canvas.pixels


ref = None
for c in chunks:
    ap.Painter.PREFETCH_CHUNKS = int(c.split(":")[0])
    eng.BAND_RATIO = float(c.split(":")[1]) if ":" in c else 1.0
    step(); step()
    torch.cuda.synchronize()
    ts = []
    for _ in range(3):
        t0 = time.perf_counter()
        x = step()
        ts.append(1e3 * (time.perf_counter() - t0))
    chk = float(np.abs(x).sum())
    ref = ref or chk
    print(json.dumps({"chunks": c, "ms": [round(t, 1) for t in ts], "min_ms": round(min(ts), 1),
                      "abs_sum_rel_to_first": chk / ref}), flush=True)
