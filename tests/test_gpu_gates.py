"""SURVEY 8(d) parity gates at the BASELINE configs' own sizes:

  C1  all 1e3 halos of the random box at nside 512, R_times 5: every pixel set bit-exact, map 1e-6
  C2 / C3 / C4  a 1e4-halo subsample of the config's catalog that holds the 100 largest and the 100 smallest
      discs and every pole-containing disc, at the config's real nside: pixel sets bit-exact;
      1e3 halos: maps to 1e-6 (Battaglia16.tau_2D_interp at 4096, Battaglia16.kSZ_T and NFW.BG at 8192)
  C5  cutouts (bit-exact) and their stack (1e-12) of 200 halos, 0.4 x 0.4 deg, 200 x 200, from an nside 8192 map
  near-tie audit of the disc query over the full 1e7-halo bench catalog (device flags, host libm recompute)

The oracle (oracle/) is the checker; the CUDA path is what is measured against it.
"""
import multiprocessing as mp
import os

import numpy as np
import pytest

from .conftest import map_close

pytestmark = pytest.mark.gpu

INPUT_COLS = ["x", "y", "z", "v_x", "v_y", "v_z", "M_200c", "redshift"]


def _oracle_map_job(job):
    df, nside, R_times, name, lo, hi = job
    import warnings
    from oracle import reference_loop as ref, profiles_np as oprof
    warnings.simplefilter("ignore")
    m = np.zeros(12 * nside * nside)
    n = ref.spray(m, df, nside, getattr(oprof, name), R_times=R_times, halo_list=np.arange(lo, hi))
    idx = np.nonzero(m)[0]
    return idx, m[idx], n


def _oracle_map(df, nside, R_times, name):
    """the oracle's spray of every halo of df, split over the host cores (sparse partial maps summed in order)"""
    n = len(df)
    procs = max(1, min(os.cpu_count() or 1, 16, n // 20))
    bounds = np.linspace(0, n, procs + 1).astype(int)
    jobs = [(df, nside, R_times, name, bounds[i], bounds[i + 1]) for i in range(procs)]
    with mp.get_context("fork").Pool(procs) as pool:
        parts = pool.map(_oracle_map_job, jobs)
    m = np.zeros(12 * nside * nside)
    pairs = 0
    for idx, val, k in parts:
        m[idx] += val
        pairs += k
    return m, pairs


def _subsample(cat, nside, R_times, n_total, rng):
    """positions of n_total halos: the 100 largest and 100 smallest discs, every pole-containing disc, random rest"""
    import torch
    from astropaint_b200 import _engine as eng
    dev = torch.device("cuda", 0)
    halos = eng.DeviceHalos(cat.data, R_times, dev)
    n_pix = eng.disc_count(halos, nside)[0].cpu().numpy()
    theta = cat.data["theta"].values
    radius = R_times * np.deg2rad(cat.data["R_ang_200c"].values / 60.0)
    polar = np.nonzero((theta - radius <= 0) | (theta + radius >= np.pi))[0]
    order = np.argsort(n_pix, kind="stable")
    chosen = np.unique(np.concatenate([order[:100], order[-100:], polar]))
    rest = np.setdiff1d(rng.choice(cat.size, size=min(cat.size, n_total + len(chosen)), replace=False), chosen)
    idx = np.concatenate([chosen, rest[:max(0, n_total - len(chosen))]])
    del halos
    return np.sort(idx), n_pix, len(polar)


def _check_pixel_sets(cat, nside, R_times, idx):
    import astropaint_b200 as ap
    from oracle import reference_loop as ref
    canvas = ap.Canvas(cat, nside, R_times=R_times)
    discs = ref.Discs(cat.data, nside, R_times)
    n_checked = n_pixels = 0
    for h, got in zip(idx, canvas.discs.gen_pixel_index(halo_list=idx)):
        want = discs.pixel_index(int(h))
        assert got.dtype == np.int64 and np.array_equal(got, want), f"halo {h}: pixel set differs at nside {nside}"
        n_checked += 1
        n_pixels += len(want)
    return n_checked, n_pixels


def test_C1_in_full_every_pixel_set_and_the_map():
    import astropaint_b200 as ap
    from astropaint_b200 import synthetic
    from astropaint_b200.profiles import spherical
    from oracle import reference_loop as ref, profiles_np as oprof
    nside, R_times = 512, 5
    cat = ap.Catalog(synthetic.random_box(1000, seed=0))
    n_checked, n_pixels = _check_pixel_sets(cat, nside, R_times, np.arange(cat.size))
    assert n_checked == 1000 and n_pixels > 1.0e8        # N_hp ~ 1.1e8
    canvas = ap.Canvas(cat, nside, R_times=R_times)
    p = ap.Painter(spherical.gaussian_2D)
    p.spray(canvas)
    want = np.zeros(12 * nside * nside)
    pairs = ref.spray(want, cat.data, nside, oprof.gaussian_readme, R_times=R_times)
    assert int(p.last_spray["d_count"].item()) == pairs == n_pixels
    map_close(canvas.pixels, want, 1e-6)


@pytest.fixture(scope="module")
def websky_1e6():
    import astropaint_b200 as ap
    from astropaint_b200 import synthetic
    ap.paint_bucket.verbose = False
    return ap.Catalog(synthetic.websky_like(1_000_000, seed=1))


@pytest.mark.parametrize("config,nside", [("C2", 4096), ("C4", 8192)])
def test_C2_C4_pixel_sets_on_the_gate_subsample(websky_1e6, config, nside):
    idx, n_pix, n_polar = _subsample(websky_1e6, nside, 5, 10_000, np.random.default_rng(7))
    assert len(idx) == 10_000
    n_checked, n_pixels = _check_pixel_sets(websky_1e6, nside, 5, idx)
    assert n_checked == 10_000 and n_pixels >= int(np.sort(n_pix)[-100:].sum())
    print(f"{config}: {n_checked} discs, {n_pixels} pixels, {n_polar} polar discs, largest {n_pix.max()}, smallest {n_pix.min()}")


@pytest.fixture(scope="module")
def websky_1e7():
    import astropaint_b200 as ap
    from astropaint_b200 import synthetic
    ap.paint_bucket.verbose = False
    return ap.Catalog(synthetic.websky_like(10_000_000, seed=1))      # the bench catalog (BASELINE config 3)


def test_C3_pixel_sets_on_the_gate_subsample_of_the_1e7_catalog(websky_1e7):
    cat = websky_1e7
    idx, n_pix, n_polar = _subsample(cat, 8192, 5, 10_000, np.random.default_rng(8))
    n_checked, n_pixels = _check_pixel_sets(cat, 8192, 5, idx)
    assert n_checked == 10_000
    print(f"C3: {n_checked} discs, {n_pixels} pixels, {n_polar} polar discs, largest {n_pix.max()}, smallest {n_pix.min()}")


def test_polar_discs_at_nside_8192():
    """The Websky-shaped catalogs above hold no pole-containing disc (a 1.5' disc on a uniform sphere), so the
    polar branches of the query (whole rings above irmin / below irmax, the caps' 4 i pixels per ring) are gated
    here at the bench resolution: 300 halos moved to within their own disc radius of either pole."""
    import astropaint_b200 as ap
    from astropaint_b200 import synthetic
    from astropaint_b200.profiles import NFW
    from oracle import reference_loop as ref, profiles_np as oprof
    df = synthetic.websky_like(300, seed=77, m_min=2e13)
    cat0 = ap.Catalog(df.copy())
    rng = np.random.default_rng(77)
    radius = 5 * np.deg2rad(cat0.data["R_ang_200c"].values / 60.0)
    theta = rng.uniform(0, 1.5, 300) * radius                       # two thirds of them contain the pole
    theta[:4] = [0.0, 1e-12, radius[2], 0.5 * radius[3]]
    theta[150:] = np.pi - theta[150:]
    phi = rng.uniform(0, 2 * np.pi, 300)
    D = np.sqrt(df.x ** 2 + df.y ** 2 + df.z ** 2).values
    df["x"], df["y"], df["z"] = D * np.sin(theta) * np.cos(phi), D * np.sin(theta) * np.sin(phi), D * np.cos(theta)
    cat = ap.Catalog(df)
    nside = 8192
    n_checked, n_pixels = _check_pixel_sets(cat, nside, 5, np.arange(300))
    canvas = ap.Canvas(cat, nside, R_times=5)
    p = ap.Painter(NFW.kSZ_T)
    p.spray(canvas)
    want = np.zeros(12 * nside * nside)
    pairs = ref.spray(want, cat.data, nside, oprof.nfw_kSZ_T, R_times=5)
    assert int(p.last_spray["d_count"].item()) == pairs == n_pixels
    assert want[:4].any() and want[-4:].any()                        # the polar pixels themselves are painted
    map_close(canvas.pixels, want, 1e-6)


MAP_CASES = [("C2", 4096, "Battaglia16.tau_2D_interp", "b16_tau_2D_interp"),
             ("C3", 8192, "Battaglia16.kSZ_T", "b16_kSZ_T"),
             ("C4", 8192, "NFW.BG", "nfw_BG")]


@pytest.mark.parametrize("config,nside,template,oracle_name", MAP_CASES, ids=[c[0] for c in MAP_CASES])
def test_C2_C3_C4_maps_of_1e3_halos(websky_1e6, config, nside, template, oracle_name):
    import astropaint_b200 as ap
    from astropaint_b200.profiles import NFW, Battaglia16
    rng = np.random.default_rng(11)
    rows = np.sort(rng.choice(websky_1e6.size, size=1000, replace=False))
    sub = ap.Catalog(websky_1e6.data.iloc[rows][INPUT_COLS].reset_index(drop=True))
    mod, fn = template.split(".")
    tmpl = getattr({"NFW": NFW, "Battaglia16": Battaglia16}[mod], fn)
    canvas = ap.Canvas(sub, nside, R_times=5)
    p = ap.Painter(tmpl)
    p.spray(canvas)
    assert p.last_spray["path"].startswith("device_")
    want, pairs = _oracle_map(sub.data, nside, 5, oracle_name)
    assert int(p.last_spray["d_count"].item()) == pairs
    err = map_close(canvas.pixels, want, 1e-6)
    print(f"{config}: 1000 halos, {pairs} halo-pixels, max-abs error / max |map| = {err:.2e}")


def test_C5_cutouts_and_stack_at_nside_8192(websky_1e6):
    import astropaint_b200 as ap
    from astropaint_b200.profiles import NFW
    from oracle import reference_loop as ref
    nside = 8192
    canvas = ap.Canvas(websky_1e6, nside, R_times=5)
    ap.Painter(NFW.BG).spray(canvas)
    pixels = canvas.pixels
    rng = np.random.default_rng(5)
    halo_list = np.sort(rng.choice(100_000, size=200, replace=False))
    halo_list[:2] = [int(np.argmin(websky_1e6.data["theta"].values[:100_000])),      # closest to the poles
                     int(np.argmax(websky_1e6.data["theta"].values[:100_000]))]
    rng_deg = [-0.2, 0.2]
    got = list(canvas.cutouts(halo_list=halo_list, lon_range=rng_deg, lat_range=rng_deg, xpix=200, ypix=200))
    want = list(ref.cutouts(pixels, websky_1e6.data, nside, halo_list, rng_deg, rng_deg, 200, 200))
    assert len(got) == len(want) == 200
    n_nonzero = 0
    for g, w in zip(got, want):
        assert g.shape == (200, 200) and np.array_equal(g, w)       # nearest-pixel look-ups: bit-exact
        n_nonzero += int(np.count_nonzero(w))
    assert n_nonzero > 0
    stack = canvas.stack_cutouts(halo_list=halo_list, lon_range=rng_deg, lat_range=rng_deg, xpix=200, ypix=200, inplace=False)
    want_stack = np.sum(want, axis=0)
    assert np.abs(stack - want_stack).max() <= 1e-12 * np.abs(want_stack).max()


def test_near_tie_audit_of_the_full_bench_catalog(websky_1e7):
    """device flags every halo with a span boundary within tol pixels of an integer, the host
    recomputes those discs with glibc; over the 1e7-halo bench catalog (7.9e8 halo-pixels, 2.2e8 ring boundaries)
    no pixel set may differ."""
    import torch
    from astropaint_b200 import _engine as eng
    halos = eng.DeviceHalos(websky_1e7.data, 5, torch.device("cuda", 0))
    tight = eng.disc_tie_audit(halos, 8192, tol=1e-9)
    loose = eng.disc_tie_audit(halos, 8192, tol=1e-5)       # exercises the recompute on a few thousand discs
    print(f"near-tie audit, 1e7 halos at nside 8192: tol 1e-9 -> {tight['n_flagged']} flagged, {tight['n_flips']} flips; "
          f"tol 1e-5 -> {loose['n_flagged']} flagged, {loose['n_flips']} flips; smallest margin {tight['min_margin']:.3e} pixel")
    assert tight["n_flips"] == 0 and loose["n_flips"] == 0
    assert loose["n_flagged"] > 100 and loose["n_flagged"] >= tight["n_flagged"]
    assert tight["min_margin"] >= 0.0


def test_table_host_entry_and_pixel_ang_generator():
    """C ABI: ap_spray_table_host (host buffers in and out for the tabulated templates) == Painter.spray;
    Canvas.Disc.gen_pixel_ang == hp.pix2ang of the disc's pixels (paint_bucket.py:1231-1238)."""
    import ctypes as C
    import astropaint_b200 as ap
    from astropaint_b200 import synthetic, _capi
    from astropaint_b200.profiles import Battaglia16
    from oracle import healpix_np as H
    nside = 1024
    cat = ap.Catalog(synthetic.websky_like(400, seed=31, m_min=5e13))
    canvas = ap.Canvas(cat, nside, R_times=5)
    p = ap.Painter(Battaglia16.kSZ_T)
    p.spray(canvas)
    want = canvas.pixels.copy()
    d = cat.data
    col = lambda k: np.ascontiguousarray(d[k].values, dtype=np.float64)   # noqa: E731
    radius = np.ascontiguousarray(5 * np.deg2rad(d.R_ang_200c.values / 60))
    scale = np.ascontiguousarray(-col("v_r") / Battaglia16.c * Battaglia16.T_cmb * (1 + col("redshift")))
    h_map = np.zeros(12 * nside * nside)
    cnt = C.c_uint64(0)
    L = _capi.lib()
    args = [col("x"), col("y"), col("z"), radius, col("theta"), col("phi"), col("D_a"), col("R_200c"), col("M_200c"),
            col("redshift"), scale]
    rc = L.ap_spray_table_host(0, nside, len(d), *[_capi.ptr(a) for a in args], 15, 0, _capi.ptr(h_map), C.byref(cnt))
    assert rc == 0, L.ap_last_error()
    assert cnt.value == int(p.last_spray["d_count"].item())
    map_close(h_map, want, 1e-12)
    assert L.ap_spray_table_host(9, nside, len(d), *[_capi.ptr(a) for a in args], 15, 0, _capi.ptr(h_map), C.byref(cnt)) != 0
    # gen_pixel_ang
    hp = H.HealpixRing(nside)
    for (theta, phi), pix in zip(canvas.discs.gen_pixel_ang(halo_list=[0, 5, 17]), canvas.discs.gen_pixel_index(halo_list=[0, 5, 17])):
        t, f = hp.pix2ang(pix)
        assert np.allclose(theta, t, rtol=0, atol=1e-14) and np.allclose(phi, f, rtol=0, atol=1e-14)


def test_device_query_disc_reproduces_healpys_own_test_vectors():
    """the CUDA disc query (exclusive and inclusive) on the two pixel lists healpy/test/test_query_disc.py asserts"""
    import pandas as pd
    import astropaint_b200 as ap
    from .golden import healpy_docstring_kat as K
    q = K.QUERY_DISC_TEST
    v = np.array(q["vec"]) * 100.0                       # 100 Mpc away in that direction
    cat = ap.Catalog(pd.DataFrame({"x": [v[0]], "y": [v[1]], "z": [v[2]], "v_x": [0.0], "v_y": [0.0], "v_z": [0.0],
                                   "M_200c": [1e14]}))
    # R_times such that R_times * arcmin2rad(R_ang_200c) is the test's radius of 6 degrees
    R_times = q["radius_deg"] * 60.0 / float(cat.data["R_ang_200c"].values[0])
    for inclusive, want in ((False, q["exclusive"]), (True, q["inclusive"])):
        canvas = ap.Canvas(cat, q["nside"], R_times=R_times, inclusive=inclusive)
        got = next(canvas.discs.gen_pixel_index())
        assert got.tolist() == want, (inclusive, got)
