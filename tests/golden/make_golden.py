"""Generates tests/golden/golden_small.npz from the CPU oracle (oracle/), which restates the
reference's loop; the reference itself cannot be imported here (healpy / astropy absent), so these
vectors pin the ORACLE (regression) and give the GPU tests a fixture that does not need scipy time.

    python tests/golden/make_golden.py
"""
import hashlib
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from oracle import reference_loop as ref, profiles_np as oprof  # noqa: E402
import astropaint_b200 as ap  # noqa: E402  (host-side Catalog only: derived columns)
from astropaint_b200 import synthetic  # noqa: E402

ap.paint_bucket.verbose = False


def sparse(m):
    idx = np.nonzero(m)[0]
    return idx.astype(np.int64), m[idx]


def main():
    out = {}
    # 1. the reference's test fixture: 6 axis halos, nside 64 (tests/conftest.py:10-34)
    cat = ap.Catalog("test")
    discs = ref.Discs(cat.data, 64, 5)
    pix = [discs.pixel_index(h) for h in range(6)]
    out["test6_counts"] = np.array([len(p) for p in pix])
    out["test6_sha256"] = np.frombuffer(hashlib.sha256(np.concatenate(pix).tobytes()).digest(), dtype=np.uint8)
    m = np.zeros(12 * 64 * 64)
    ref.spray(m, cat.data, 64, oprof.gaussian_readme, R_times=5)
    out["test6_gaussian_map"] = m
    # 2. BASELINE config 1 (reduced): random box, Gaussian, nside 128, R_times 5
    cat = ap.Catalog(synthetic.random_box(50, seed=0))
    m = np.zeros(12 * 128 * 128)
    out["box_pairs"] = np.array(ref.spray(m, cat.data, 128, oprof.gaussian_readme, R_times=5))
    out["box_gaussian_map"] = m
    m = np.zeros(12 * 128 * 128)
    ref.spray(m, cat.data, 128, oprof.nfw_kSZ_T, R_times=3)
    out["box_nfw_ksz_map"] = m
    m = np.zeros(12 * 128 * 128)
    ref.spray(m, cat.data, 128, oprof.nfw_BG, R_times=3)
    out["box_nfw_bg_map"] = m
    stack = ref.stack_cutouts(out["box_gaussian_map"], cat.data, 128, np.arange(10), [-2, 2], None, 32)
    out["box_stack"] = stack
    # 3. BASELINE configs 2/3 (reduced): Websky-shaped, Battaglia16 kSZ, nside 1024
    cat = ap.Catalog(synthetic.websky_like(16, seed=5, m_min=2e14))
    m = np.zeros(12 * 1024 * 1024)
    out["websky_pairs"] = np.array(ref.spray(m, cat.data, 1024, oprof.b16_kSZ_T, R_times=5))
    out["websky_ksz_idx"], out["websky_ksz_val"] = sparse(m)
    np.savez_compressed(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden_small.npz"), **out)
    print({k: (v.shape, v.dtype) for k, v in out.items()})


if __name__ == "__main__":
    main()
