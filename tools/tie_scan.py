"""Empirical near-tie check: device disc query (CUDA libm) vs the host build of the same header (glibc)
on N Websky-shaped halos at nside 8192: per-halo pixel count and sum of pixel ids must agree."""
import os, sys, ctypes as C, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import astropaint_b200 as ap
from astropaint_b200 import synthetic, _engine as eng
ap.paint_bucket.verbose = False
N = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
L = C.CDLL(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "astropaint_b200", "_native", "libap_hostcheck.so"))
L.hc_query_disc_stats.argtypes = [C.c_int64, C.c_int64] + [C.c_void_p] * 6
cat = ap.Catalog(synthetic.websky_like(N, seed=21))
dev = torch.device("cuda", 0)
halos = eng.DeviceHalos(cat.data, 5, dev)
n_pix, _, _, _ = eng.disc_count(halos, 8192)
# device: sum of pixel ids per halo via the flat list in batches
cnt_d = n_pix.cpu().numpy(); sum_d = np.zeros(N, dtype=np.int64)
B = 200000
for lo in range(0, N, B):
    hi = min(N, lo + B)
    off, pix, _, _ = eng.disc_pixels(halos, 8192, n_pix, lo, hi)
    seg = torch.repeat_interleave(torch.arange(hi - lo, device=dev), n_pix[lo:hi])
    s = torch.zeros(hi - lo, dtype=torch.int64, device=dev).index_add_(0, seg, pix)
    sum_d[lo:hi] = s.cpu().numpy()
d = cat.data
x, y, z = (np.ascontiguousarray(d[k].values) for k in "xyz")
rad = np.ascontiguousarray(5 * np.deg2rad(d.R_ang_200c.values / 60))
cnt_h = np.zeros(N, dtype=np.int64); sum_h = np.zeros(N, dtype=np.int64)
t0 = time.time()
L.hc_query_disc_stats(8192, N, x.ctypes.data, y.ctypes.data, z.ctypes.data, rad.ctypes.data, cnt_h.ctypes.data, sum_h.ctypes.data)
bad = np.nonzero((cnt_h != cnt_d) | (sum_h != sum_d))[0]
print(f"halos {N}, halo-pixels {cnt_d.sum()}, ring boundaries ~{2 * 11 * N:.2e}: {len(bad)} halos differ between CUDA and glibc ({time.time()-t0:.1f} s host)")
print(bad[:10], cnt_h[bad[:10]], cnt_d[bad[:10]])
