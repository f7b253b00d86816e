"""CPU tests of the oracle itself.  The reference's tests pin no numbers for this path and healpy
is absent (parity unpinned), so the oracle is pinned by construction-true properties and by the
committed golden vectors (tests/golden/, generated from it: a regression pin)."""
import hashlib
import os

import numpy as np
import pytest

from oracle import healpix_np as H, reference_loop as ref, profiles_np as oprof

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "golden_small.npz")


@pytest.mark.parametrize("nside", [1, 2, 3, 4, 12, 16, 64, 100])
def test_pix_vec_ang_round_trips(nside):
    hp = H.HealpixRing(nside)
    pix = np.arange(hp.npix)
    x, y, z = hp.pix2vec(pix)
    assert np.allclose(x * x + y * y + z * z, 1.0, atol=1e-14)
    assert np.array_equal(hp.vec2pix(x, y, z), pix)
    assert np.array_equal(hp.vec2pix(3.7 * x, 3.7 * y, 3.7 * z), pix)   # un-normalised input
    th, ph = hp.pix2ang(pix)
    assert np.array_equal(hp.ang2pix(th, ph), pix)
    # ring structure: startpix/ringpix tile [0, npix) and z decreases ring by ring
    total, zs = 0, []
    for r in range(1, 4 * nside):
        sp, rp, _ = hp.ring_info(r)
        assert sp == total
        total += rp
        zs.append(hp.ring2z(r))
    assert total == hp.npix
    assert np.all(np.diff(zs) < 0)


@pytest.mark.parametrize("nside", [1, 3, 4, 12, 16, 64])
def test_query_disc_is_brute_force_disc(nside):
    """{p : angdist(p, c) <= r} over ALL pixels, 250 random discs incl. poles / whole sky / empty."""
    hp = H.HealpixRing(nside)
    rng = np.random.default_rng(nside)
    pix = np.arange(hp.npix)
    th, ph = hp.pix2ang(pix)
    for t in range(250):
        v = rng.normal(size=3) * rng.uniform(0.1, 100)
        if t % 25 == 0:
            v = np.array([0.0, 0.0, 1.0 if t % 50 else -1.0])
        r = rng.choice([rng.uniform(0, 0.2), rng.uniform(0, np.pi * 1.05), rng.uniform(0, 3 / nside)])
        q = hp.query_disc(v, r)
        assert np.all(np.diff(q) > 0)
        thc = np.arctan2(np.hypot(v[0], v[1]), v[2])
        phc = np.arctan2(v[1], v[0])
        d = H.angdist(np.array([th, ph]), (thc, phc))
        diff = np.setxor1d(q, pix[d <= r])
        assert len(diff) == 0 or np.abs(d[diff] - r).max() < 1e-9   # only exact boundary ties may differ


@pytest.mark.parametrize("nside", [2, 4, 16, 64])
def test_inclusive_query_covers_every_overlapping_pixel(nside):
    """hp.query_disc(inclusive=True) (paint_bucket.py:1228 with Canvas(inclusive=True), :782) by construction:
    (a) sorted and unique, (b) a superset of the exclusive query, (c) EVERY point inside the disc falls in a
    returned pixel (so every overlapping pixel is returned), (d) no returned pixel has its centre farther
    than radius + max_pixrad, (e) fact=1 (no trimming) is a superset of fact=4."""
    hp = H.HealpixRing(nside)
    rng = np.random.default_rng(40 + nside)
    pixsize = np.sqrt(4 * np.pi / hp.npix)
    for t in range(40):
        v = rng.normal(size=3) * rng.uniform(0.1, 100)
        if t % 10 == 0:
            v = np.array([1e-3, -2e-3, 1.0 if t % 20 else -1.0])   # pole inside the disc
        r = [rng.uniform(0, 4) * pixsize, rng.uniform(0, 1.5), 0.0, rng.uniform(1.5, 3.3)][t % 4]
        inc = hp.query_disc(v, r, inclusive=True)
        exc = hp.query_disc(v, r)
        assert np.all(np.diff(inc) > 0)
        assert np.isin(exc, inc).all()
        assert np.isin(inc, hp.query_disc(v, r, inclusive=True, fact=1)).all()
        vn = v / np.linalg.norm(v)
        m = 3000
        u = rng.normal(size=(m, 3))
        u -= (u @ vn)[:, None] * vn
        u /= np.linalg.norm(u, axis=1)[:, None]
        ang = min(r, np.pi) * np.sqrt(rng.uniform(0, 1, m))
        ang[:300] = min(r, np.pi) * (1 - 1e-9)          # the rim, where inclusion is decided
        pts = np.cos(ang)[:, None] * vn + np.sin(ang)[:, None] * u
        assert np.isin(hp.vec2pix(pts[:, 0], pts[:, 1], pts[:, 2]), inc).all()
        if len(inc):
            cx, cy, cz = hp.pix2vec(inc)
            dist = np.arccos(np.clip(cx * vn[0] + cy * vn[1] + cz * vn[2], -1, 1))
            assert (dist <= r + hp.max_pixrad() + 1e-9).all()


@pytest.mark.parametrize("nside", [1, 3, 16, 256, 8192])
def test_inclusive_fast_path_equals_plain_restatement(nside):
    """the pre-filter used to keep the inclusive oracle fast in the big parity tests changes nothing"""
    hp = H.HealpixRing(nside)
    rng = np.random.default_rng(70 + nside)
    pixsize = np.sqrt(4 * np.pi / hp.npix)
    for t in range(60 if nside <= 256 else 6):
        v = rng.normal(size=3) * rng.uniform(0.1, 100)
        if t % 10 == 0:
            v = np.array([1e-3, -2e-3, 1.0 if t % 20 else -1.0])
        if nside <= 256:
            r = [rng.uniform(0, 6) * pixsize, rng.uniform(0, 1.5), 0.0, rng.uniform(1.5, 3.3)][t % 4]
        else:
            r = rng.uniform(0, 12) * pixsize
        for fact in (4, 2):
            a = hp.query_disc(v, r, inclusive=True, fact=fact)
            b = hp.query_disc(v, r, inclusive=True, fact=fact, fast=True)
            assert np.array_equal(a, b), (nside, t, fact)


@pytest.mark.parametrize("nside", [1, 2, 5, 16, 128])
def test_ring_xyf_round_trip_and_max_pixrad(nside):
    hp = H.HealpixRing(nside)
    rng = np.random.default_rng(nside)
    for pix in np.unique(np.concatenate([rng.integers(0, hp.npix, 300), [0, hp.npix - 1, hp.ncap, hp.npix - hp.ncap - 1]])):
        ix, iy, face = hp.ring2xyf(int(pix))
        assert 0 <= ix < nside and 0 <= iy < nside and 0 <= face < 12
        assert hp.xyf2ring(ix, iy, face) == pix
    # max_pixrad bounds the centre-to-corner distance: between half a pixel side and a full one
    side = np.sqrt(4 * np.pi / hp.npix)
    assert 0.5 * side < hp.max_pixrad() < 1.1 * side


def test_cartesian_projection_orientation():
    """centre of the cutout is the halo; +y is north; +x is west (astro flip)."""
    nside = 256
    hp = H.HealpixRing(nside)
    m = np.arange(hp.npix, dtype=np.float64)
    lon, lat = 40.0, 25.0
    img = H.cartesian_projmap(hp, m, lon, lat, [-1, 1], [-1, 1], 101)
    assert img.shape == (101, 101)
    th, ph = hp.pix2ang(img[50, 50].astype(np.int64))
    assert abs(np.degrees(ph[0]) - lon) < 0.3 and abs(90 - np.degrees(th[0]) - lat) < 0.3
    th_n, _ = hp.pix2ang(img[100, 50].astype(np.int64))
    assert (90 - np.degrees(th_n[0])) > lat + 0.7
    _, ph_w = hp.pix2ang(img[50, 100].astype(np.int64))
    assert np.degrees(ph_w[0]) < lon - 0.7
    # ysize=None follows the lat/lon aspect ratio
    assert H.cartesian_projmap(hp, m, lon, lat, [-2, 2], [-1, 1], 40).shape == (20, 40)


def test_los_projection_of_tophat_is_analytic():
    """README.md:170-218: LOS_integrate(tophat_3D) == 2 sqrt(R_200c^2 - R^2) (known answer)."""
    R200 = 1.7

    def tophat(r, R_200c):
        return np.where(np.asarray(r) > R_200c, 0.0, 1.0)

    import warnings
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        R = np.linspace(0.05, 1.6, 12)
        got = oprof.los_integrate(tophat, R, R200)
    assert np.allclose(got, 2 * np.sqrt(R200 ** 2 - R ** 2), rtol=2e-2)   # quad runs at epsrel=1e-2


def test_nfw_complex_forms_equal_real_closed_forms():
    R = np.concatenate([np.linspace(0.01, 0.95, 30), np.linspace(1.05, 8, 30)]) * 0.4
    x = R / 0.4
    f_lt = 1 - 2 / np.sqrt(1 - x[x < 1] ** 2) * np.arctanh(np.sqrt((1 - x[x < 1]) / (1 + x[x < 1])))
    f_gt = 1 - 2 / np.sqrt(x[x > 1] ** 2 - 1) * np.arctan(np.sqrt((x[x > 1] - 1) / (x[x > 1] + 1)))
    want = 8 * 1e15 * 0.4 * np.concatenate([f_lt, f_gt]) / (x ** 2 - 1)
    assert np.allclose(oprof.nfw_rho_2D_bartlemann(R, 1e15, 0.4), want, rtol=1e-12)


def test_oracle_reproduces_golden_vectors():
    import astropaint_b200 as ap
    from astropaint_b200 import synthetic
    g = np.load(GOLD)
    cat = ap.Catalog("test")
    discs = ref.Discs(cat.data, 64, 5)
    pix = [discs.pixel_index(h) for h in range(6)]
    assert np.array_equal(g["test6_counts"], [len(p) for p in pix])
    sha = np.frombuffer(hashlib.sha256(np.concatenate(pix).tobytes()).digest(), dtype=np.uint8)
    assert np.array_equal(g["test6_sha256"], sha)
    # symmetry of the fixture: the four equatorial discs map onto each other under 90-degree
    # rotations about z, the two polar discs under the north-south mirror
    c = g["test6_counts"].tolist()
    assert len(set(c[:4])) == 1 and c[4] == c[5]
    m = np.zeros(12 * 64 * 64)
    ref.spray(m, cat.data, 64, oprof.gaussian_readme, R_times=5)
    assert np.allclose(m, g["test6_gaussian_map"], rtol=1e-13, atol=1e-13 * np.abs(m).max())
    cat = ap.Catalog(synthetic.random_box(50, seed=0))
    m = np.zeros(12 * 128 * 128)
    n = ref.spray(m, cat.data, 128, oprof.nfw_kSZ_T, R_times=3)
    assert np.allclose(m, g["box_nfw_ksz_map"], rtol=1e-12, atol=1e-13 * np.abs(m).max())
    assert n > 0
    cat = ap.Catalog(synthetic.websky_like(16, seed=5, m_min=2e14))
    m = np.zeros(12 * 1024 * 1024)
    import warnings
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        n = ref.spray(m, cat.data, 1024, oprof.b16_kSZ_T, R_times=5)
    assert n == int(g["websky_pairs"])
    idx = np.nonzero(m)[0]
    assert np.array_equal(idx, g["websky_ksz_idx"])
    assert np.allclose(m[idx], g["websky_ksz_val"], rtol=1e-12)


def test_catalog_columns_match_oracle_build_dataframe():
    import astropaint_b200 as ap
    from astropaint_b200 import synthetic
    df = synthetic.websky_like(500, seed=3)
    cat = ap.Catalog(df)
    want = ref.build_dataframe(df)
    assert list(cat.data.columns) == list(want.columns)
    assert len(cat.data.columns) == 24
    for c in want.columns:
        assert np.allclose(cat.data[c].values, want[c].values, rtol=1e-14, atol=0), c


@pytest.mark.parametrize("shape", [(16, 16), (21, 21), (12, 19), (5, 5)])
def test_patch_rotate_restatement_is_scipy_rotate(shape):
    """oracle/patch_ops_np.py (what csrc/patch_ops.cuh is read against) == scipy.ndimage.rotate itself, the
    function the reference's transform.rotate_patch calls (lib/transform.py:266-275): a PINNED oracle."""
    from scipy import ndimage
    from oracle import patch_ops_np as P
    rng = np.random.default_rng(shape[0])
    p = rng.normal(size=shape)
    assert np.abs(ndimage.spline_filter(p, order=3, mode="constant") - P.spline_prefilter(p)).max() < 1e-13
    for angle in [0, 13.7, 45, 90, -120.3, 180, 271.1, 359.99]:
        want = ndimage.rotate(p, angle, reshape=False)
        got = P.rotate_patch(p, angle)
        assert np.array_equal(want == 0, got == 0)           # same pixels fall outside the source
        assert np.abs(got - want).max() < 1e-13
    if shape[0] == shape[1]:     # the reference's window is square
        from astropaint_b200.lib import transform
        assert np.array_equal(P.taper_patch(p), transform.taper_patch(p.copy()))


def test_ud_grade_restatement_by_construction():
    """oracle ud_grade (healpy.ud_grade as Canvas.ud_grade calls it, paint_bucket.py:2233-2257; unpinned): ring2nest
    is a bijection, the children of a parent tile the parent (their mean direction is the parent's centre),
    degrade(upgrade(m)) == m, means are preserved, UNSEEN children are skipped / poison under pess."""
    hp8, hp4 = H.HealpixRing(8), H.HealpixRing(4)
    nest = np.array([H.ring2nest(hp8, p) for p in range(hp8.npix)])
    assert np.array_equal(np.sort(nest), np.arange(hp8.npix))
    x, y, z = hp8.pix2vec(np.arange(hp8.npix))
    x4, y4, z4 = hp4.pix2vec(np.arange(hp4.npix))
    for fine, coarse in ((x, x4), (y, y4), (z, z4)):
        assert np.abs(H.ud_grade(fine, 4) - coarse).max() < 0.02       # pixel size at nside 4 is 0.25 rad
    rng = np.random.default_rng(0)
    m = rng.normal(size=hp4.npix)
    up = H.ud_grade(m, 16)
    assert np.array_equal(np.sort(np.unique(up)), np.sort(m)) and len(up) == 12 * 256
    assert np.allclose(H.ud_grade(up, 4), m, rtol=0, atol=1e-15)
    fine = rng.normal(size=hp8.npix)
    assert abs(H.ud_grade(fine, 2).mean() - fine.mean()) < 1e-15
    holed = fine.copy()
    holed[rng.integers(0, hp8.npix, 40)] = H.UNSEEN
    d = H.ud_grade(holed, 4)
    dp = H.ud_grade(holed, 4, pess=True)
    assert (d != H.UNSEEN).all() and (dp == H.UNSEEN).sum() >= 10
    assert np.allclose(H.ud_grade(m, 8, power=-2), H.ud_grade(m, 8) * 2.0 ** -2)



# ------------------------------------------------------------------ the one output of the REAL reference
NB_GOLD = os.path.join(os.path.dirname(GOLD), "notebook_catalog_head.json")
NB_INPUTS = ("x", "y", "z", "v_x", "v_y", "v_z", "M_200c", "redshift")
NB_DERIVED = ("D_c", "lat", "R_200c", "c_200c", "R_ang_200c", "rho_s", "R_s", "v_r", "v_th", "v_ph", "v_lat", "v_lon")


def notebook_head():
    import json
    import pandas as pd
    g = json.load(open(NB_GOLD))
    inp = pd.DataFrame({k: g["values"][k] for k in NB_INPUTS})
    return g, inp


def check_against_notebook(df, g):
    """Tolerance, written out: the notebook prints inputs with 6-9 significant digits (redshift: 6), so a
    derived column can only be reproduced to the propagated input rounding (<= 1e-6 relative: rho_c(z) and
    c_200c(M, z) move by ~1e-6 per 1e-6 in z) plus half a unit of the last printed digit of the output.
    Velocity components are sums with cancellation: relative to |v|."""
    vnorm = np.sqrt(df.v_x ** 2 + df.v_y ** 2 + df.v_z ** 2).values
    for col in NB_DERIVED:
        want = np.array(g["values"][col])
        half_ulp = np.array([0.5 * 10.0 ** (-len(t.split("e")[0].split(".")[1])) * (10.0 ** int(t.split("e")[1]) if "e" in t else 1.0)
                             for t in g["printed"][col]])
        scale = vnorm if col.startswith("v_") else np.abs(want)
        err = np.abs(df[col].values - want)
        assert np.all(err <= 1.5e-6 * scale + half_ulp), (col, err / scale)


def test_oracle_catalog_matches_reference_notebook_output():
    """REAL-reference pin: `catalog.data.head()` of examples/Birkinshaw_Gull_stacking.ipynb cell 3 -- the
    reference's own Catalog.build_dataframe (paint_bucket.py:308-364; astropy Planck18_arXiv_v2, Dutton-Maccio
    style c(M, z), cart2sph velocity Jacobian) on five WebSky halos.  Pins the oracle's cosmology constants,
    R_200c / c_200c / rho_s / R_s / R_ang_200c (hence D_a) and the velocity-component conventions (v_lat = -v_th)."""
    g, inp = notebook_head()
    check_against_notebook(ref.build_dataframe(inp.copy()), g)


def test_host_catalog_matches_reference_notebook_output():
    from astropaint_b200 import Catalog
    g, inp = notebook_head()
    cat = Catalog(inp.copy(), device=False)
    check_against_notebook(cat.data, g)
    cols = list(cat.data.columns)                     # the printed frame: "[5 rows x 24 columns]", 10 + ... + 10
    assert len(cols) == 24 and cols[:10] == g["columns"][:10] and cols[-10:] == g["columns"][-10:]
    assert sorted(cols[10:14]) == ["D_a", "lon", "phi", "theta"]


# ------------------------------------------------------------------ healpy's published docstring examples
def test_oracle_reproduces_healpy_docstring_examples():
    """Known-answer pin of the RING scheme restatement: every numeric example of healpy.pixelfunc's public
    docstrings that touches the functions the path calls (tests/golden/healpy_docstring_kat.py)."""
    from .golden import healpy_docstring_kat as K
    hp16 = H.HealpixRing(16)
    a = K.ANG2PIX_16
    assert hp16.ang2pix(np.array(a["theta"], float), np.array(a["phi"], float)).tolist() == a["pix"]
    a = K.PIX2ANG_16
    th, ph = hp16.pix2ang(np.array(a["pix"]))
    assert np.abs(th - a["theta"]).max() < 5e-9 and np.abs(ph - a["phi"]).max() < 5e-9
    th, ph = hp16.pix2ang(np.array([1440]))
    assert abs(th[0] - K.PIX2ANG_16_1440[0]) < 1e-15 and ph[0] == 0.0
    a = K.PIX2ANG_PIX11
    for ns, t, p in zip(a["nside"], a["theta"], a["phi"]):
        th, ph = H.HealpixRing(ns).pix2ang(np.array([11]))
        assert abs(th[0] - t) < 5e-9 and abs(ph[0] - p) < 5e-9
    x, y, z = hp16.pix2vec(np.array([1504]))
    assert np.abs(np.array([x[0], y[0], z[0]]) - K.PIX2VEC_16_1504).max() < 2e-16
    a = K.PIX2VEC_16
    x, y, z = hp16.pix2vec(np.array(a["pix"]))
    assert max(np.abs(x - a["x"]).max(), np.abs(y - a["y"]).max(), np.abs(z - a["z"]).max()) < 5e-9
    a = K.PIX2VEC_PIX11
    for i, ns in enumerate(a["nside"]):
        x, y, z = H.HealpixRing(ns).pix2vec(np.array([11]))
        assert max(abs(x[0] - a["x"][i]), abs(y[0] - a["y"][i]), abs(z[0] - a["z"][i])) < 5e-9
    a = K.VEC2PIX_16
    assert hp16.vec2pix(np.array(a["x"], float), np.array(a["y"], float), np.array(a["z"], float)).tolist() == a["pix"]
    assert H.ring2nest(hp16, 1504) == K.RING2NEST_16_1504
    assert [H.ring2nest(H.HealpixRing(2), i) for i in range(10)] == K.RING2NEST_2
    assert [H.ring2nest(H.HealpixRing(n), 11) for n in K.RING2NEST_PIX11["nside"]] == K.RING2NEST_PIX11["nest"]
    assert abs(hp16.max_pixrad() - K.MAX_PIXRAD_16) < 5e-11
    assert abs(np.sqrt(4 * np.pi / H.nside2npix(128)) * 10800 / np.pi - K.NSIDE2RESOL_128_ARCMIN) < 1e-13
    assert abs(4 * np.pi / H.nside2npix(128) - K.NSIDE2PIXAREA_128) < 1e-19
    assert H.ud_grade(np.arange(48.), 1).tolist() == K.UD_GRADE_48_TO_1


def _disc_agreement_job(job):
    """one chunk of the two-statement comparison (run in a worker process)"""
    from oracle.query_disc_ineq import DiscByInequality, compare
    nside, seed, n, special = job
    rng = np.random.default_rng(seed)
    hp = H.HealpixRing(nside)
    ineq = DiscByInequality(nside)
    mu = rng.uniform(-1, 1, n)
    ph = rng.uniform(0, 2 * np.pi, n)
    # radii: log-uniform 0.3 .. 70 arcmin like the bench catalog
    rad = np.deg2rad(10 ** rng.uniform(np.log10(0.3), np.log10(70), n) / 60.0)
    if special:      # poles, the polar-cap / equatorial-belt boundary, phi = 0 wrap; sky-sized and sub-pixel discs
        mu[:40] = np.repeat([1.0, -1.0, 1 - 1e-9, -1 + 1e-9, 2.0 / 3.0, -2.0 / 3.0, 0.0, 1e-12], 5)
        ph[:8] = [0.0, np.pi / 2, np.pi, 1.5 * np.pi, 1e-15, 2 * np.pi - 1e-15, np.pi / 4, 0.75 * np.pi]
        rad[40:60] = rng.uniform(0.2, 3.2, 20) if nside == 512 else rng.uniform(20, 200, 20) / nside
        rad[60:80] = rng.uniform(0, 1.5 / nside, 20)
    scale = rng.uniform(10, 5000, n)            # the reference passes un-normalised Mpc vectors
    st = np.sqrt(np.clip(1 - mu * mu, 0, None))
    log = []
    for i in range(n):
        vec = (scale[i] * st[i] * np.cos(ph[i]), scale[i] * st[i] * np.sin(ph[i]), scale[i] * mu[i])
        only_a, only_b, margins = compare(hp, ineq, vec, float(rad[i]))
        for tag, arr in (("floor-form only", only_a), ("inequality only", only_b)):
            for p in arr:
                log.append((nside, seed, i, tag, int(p), margins.get(int(p))))
    return n, log


def test_two_independent_statements_of_query_disc_agree():
    """the floor-form restatement of healpix's query_disc (oracle/healpix_np.py)
    against the inequality form cos(angdist(pixel centre, c)) > cos r with integer ring / pixel enumeration
    (oracle/query_disc_ineq.py) on 1e5 random discs at nside 512 .. 8192 (Websky-shaped radii plus polar, sky-sized and
    sub-pixel discs).  Any disagreement must be a pixel whose centre lies on the disc's edge to rounding error, and
    every one is logged with its distance from the edge."""
    import multiprocessing as mp
    jobs = [(nside, 1000 * nside + k, 3125, k == 0) for nside in (512, 2048, 4096, 8192) for k in range(8)]
    with mp.get_context("fork").Pool(min(8, os.cpu_count() or 1)) as pool:
        results = pool.map(_disc_agreement_job, jobs)
    n_discs = sum(r[0] for r in results)
    log = [row for r in results for row in r[1]]
    assert n_discs == 100000
    if log:
        print(f"\n{len(log)} disagreeing pixels over {n_discs} discs:")
        for row in log[:50]:
            print("  nside %d seed %d disc %d %s pixel %d  cos(dist) - cos(r) = %s" % row)
    # a disagreement is only legitimate ON the edge: |cos(dist) - cos(r)| at rounding level
    for nside, seed, i, tag, p, margin in log:
        assert margin is not None and abs(margin) < 1e-13, (nside, seed, i, tag, p, margin)
    assert len(log) <= 5


def test_query_disc_reproduces_healpys_own_test_vectors():
    """healpy/test/test_query_disc.py asserts two pixel lists (tests/golden/healpy_docstring_kat.py::QUERY_DISC_TEST):
    both statements of the exclusive query and the inclusive query (fact 4, healpy's default) reproduce them; angdist
    reproduces its docstring example."""
    from .golden import healpy_docstring_kat as K
    from oracle.query_disc_ineq import DiscByInequality
    q = K.QUERY_DISC_TEST
    hp = H.HealpixRing(q["nside"])
    r = np.radians(q["radius_deg"])
    assert hp.query_disc(q["vec"], r).tolist() == q["exclusive"]
    assert DiscByInequality(q["nside"]).query(q["vec"], r).tolist() == q["exclusive"]
    assert hp.query_disc(q["vec"], r, inclusive=True).tolist() == q["inclusive"]
    a = K.ANGDIST_EXAMPLE
    v1 = np.array([np.sin(a["dir1"][0]) * np.cos(a["dir1"][1]), np.sin(a["dir1"][0]) * np.sin(a["dir1"][1]), np.cos(a["dir1"][0])])
    v2 = np.array([np.sin(a["dir2"][0]) * np.cos(a["dir2"][1]), np.sin(a["dir2"][0]) * np.sin(a["dir2"][1]), np.cos(a["dir2"][0])])
    d = np.arctan2(np.linalg.norm(np.cross(v1, v2)), np.dot(v1, v2))
    assert abs(d - a["value"]) < 5e-16
