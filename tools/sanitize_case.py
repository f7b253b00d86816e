"""Small end-to-end case for compute-sanitizer (memcheck): every kernel family once."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import astropaint_b200 as ap
from astropaint_b200 import synthetic
from astropaint_b200.profiles import spherical, NFW, Battaglia16
ap.paint_bucket.verbose = False
cat = ap.Catalog(synthetic.random_box(30, seed=1))
cv = ap.Canvas(cat, 64, R_times=5)
for t in (spherical.gaussian_2D, NFW.kSZ_T, NFW.BG):
    ap.Painter(t).spray(cv)
ap.Painter(lambda R, R_200c: R / R_200c).spray(cv)
cv.stack_cutouts(np.arange(10), lon_range=[-1, 1], xpix=24)
list(cv.cutouts(np.arange(3), xpix=16))
web = ap.Catalog(synthetic.websky_like(300, seed=2, m_min=5e13))
cv = ap.Canvas(web, 1024, R_times=5)
ap.Painter(Battaglia16.kSZ_T).spray(cv)
_ = cv.pixels
ap.Painter.PREFETCH_MIN_HALOS = 0
cv.clean(); ap.Painter(Battaglia16.kSZ_T).spray(cv); _ = cv.pixels
ap.Painter(NFW.rho_2D_interp).spray(cv)
cv32 = ap.Canvas(web, 512, R_times=5, accumulate="fp32")
ap.Painter(Battaglia16.tau_2D_interp).spray(cv32); _ = cv32.pixels
print("sanitize case done")
