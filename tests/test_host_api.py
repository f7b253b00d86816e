"""CPU tests of the host-side mirror of the reference API (no compute calls: no GPU here)."""
import ctypes as C
import inspect
import os
import re

import numpy as np
import pytest


def test_library_loads_and_exports_every_declared_symbol():
    from astropaint_b200 import _capi
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    header = open(os.path.join(root, "include", "astropaint_b200.h")).read()
    declared = set(re.findall(r"\b(ap_[a-z0-9_]+)\s*\(", header))
    assert len(declared) >= 14
    lib = C.CDLL(_capi.LIB_PATH)
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in include/astropaint_b200.h but not exported"
    assert declared == set(_capi.SIGNATURES), "ctypes SIGNATURES out of sync with the header"
    assert _capi.lib().ap_version() == 200


def test_no_cpu_fallback_product_path_fails_loudly():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    import astropaint_b200 as ap
    from astropaint_b200.profiles import spherical
    canvas = ap.Canvas(ap.Catalog("test"), 16)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        ap.Painter(spherical.gaussian_2D).spray(canvas)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        next(canvas.discs.gen_pixel_index())
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        canvas.stack_cutouts()


def test_product_never_imports_oracle():
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    for dirpath, _, files in os.walk(os.path.join(root, "astropaint_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, re.M), f


def test_import_surface_matches_reference():
    import astropaint_b200 as ap
    from astropaint_b200 import Catalog, Canvas, Painter, transform, utils, LOS_integrate, interpolate  # noqa: F401
    from astropaint_b200.lib.utils import LOS_integrate as L2, interpolate as I2  # noqa: F401
    from astropaint_b200.profiles import NFW, Battaglia16, spherical  # noqa: F401
    assert inspect.signature(Canvas.__init__).parameters.keys() >= {"catalog", "nside", "mode", "R_times", "inclusive"}
    sig = inspect.signature(Painter.spray)
    assert list(sig.parameters)[:7] == ["self", "canvas", "distance_units", "n_cpus", "cache", "lazy", "lazy_grid"]
    sig = inspect.signature(Canvas.stack_cutouts)
    assert list(sig.parameters)[1:] == ["halo_list", "lon_range", "lat_range", "xpix", "ypix", "inplace",
                                        "parallel", "apply_func", "func_kwargs"]


def test_catalog_fixture_and_columns(test_catalog):
    """reference tests/test_Catalog.py:7-26"""
    d = test_catalog.data
    for c in ["x", "y", "z", "v_x", "v_y", "v_z", "M_200c", "redshift", "D_c", "lat", "lon", "theta", "phi", "D_a",
              "R_200c", "c_200c", "R_ang_200c", "rho_s", "R_s", "v_r", "v_th", "v_ph", "v_lat", "v_lon"]:
        assert c in d.columns
    assert test_catalog.size == len(d) == 6
    assert np.allclose(d.D_c, 100) and np.allclose(d.redshift, 0)
    assert np.allclose(sorted(d.lat), [-90, 0, 0, 0, 0, 90])


def test_random_box_uses_reference_rng_order():
    import astropaint_b200 as ap
    from astropaint_b200 import synthetic
    np.random.seed(0)
    cat = ap.Catalog()
    cat.generate_random_box(n_tot=100)
    want = synthetic.random_box(100, seed=0)
    for c in want.columns:
        assert np.array_equal(cat.data[c].values, want[c].values)


def test_canvas_state_flags(test_canvas):
    """reference tests/test_Canvas.py:9-61 (the parts that do not need spherical harmonics)"""
    test_canvas.clean()
    assert not test_canvas._pixel_is_outdated and test_canvas._alm_is_outdated and test_canvas._Cl_is_outdated
    assert (test_canvas.pixels == 0).all() and test_canvas.pixels.shape == (12 * 64 * 64,)
    test_canvas.pixels = 1
    assert not test_canvas._pixel_is_outdated and test_canvas._alm_is_outdated and test_canvas._Cl_is_outdated
    test_canvas.alm = 1
    assert test_canvas._pixel_is_outdated and not test_canvas._alm_is_outdated and test_canvas._Cl_is_outdated
    with pytest.raises(AttributeError):
        test_canvas.Cl = 1
    assert test_canvas.lmax == 3 * 64 - 1 and test_canvas.npix == 12 * 64 * 64
    assert test_canvas.R_max == test_canvas.catalog.data.R_200c.max()


def test_painter_introspection_and_marshalling(test_catalog):
    import astropaint_b200 as ap
    from astropaint_b200.profiles import NFW, Battaglia16

    p = ap.Painter(NFW.kSZ_T)
    assert p.template_args_list == ["R", "rho_s", "R_s", "v_r"] and p.template_kwargs_list == ["T_cmb"]
    assert p.R_arg_indx == 0 and p.template_name == "kSZ_T"
    p = ap.Painter(NFW.BG)
    assert p.template_args_list[0] == "R_vec"
    p = ap.Painter(Battaglia16.tau_2D_interp)       # decorators keep the wrapped signature
    assert p.template_args_list == ["R", "R_200c", "M_200c", "redshift"]
    with pytest.raises(AssertionError):
        ap.Painter(lambda r, M_200c: r)
    with pytest.raises(AssertionError):
        ap.Painter(lambda R, R_vec: R)

    def tmpl(R, R_200c, amplitude, *, power=2.0):
        return amplitude * (R / R_200c) ** power

    p = ap.Painter(tmpl)
    cols = p._shake_canister(test_catalog, {"amplitude": 3.0, "power": [1.0, 2, 3, 4, 5, 6]})
    assert set(cols) == {"R_200c", "amplitude", "power"}
    assert np.array_equal(cols["amplitude"], np.full(6, 3.0)) and cols["power"][5] == 6
    with pytest.raises(ValueError):
        p._shake_canister(test_catalog, {"amplitude": [1.0, 2.0]})
    vals = p.calculate_template(np.linspace(0, 1, 5), test_catalog, halo_list=[0, 2], amplitude=2.0)
    assert vals.shape == (2, 5)
    canvas = ap.Canvas(test_catalog, 16)
    with pytest.raises(ValueError):
        p.spray(canvas, n_cpus=0, amplitude=1.0)
    with pytest.raises(AssertionError):
        p.spray(canvas, distance_units="km", amplitude=1.0)
    with pytest.raises(AssertionError):
        ap.Painter(NFW.BG).spray(canvas, cache=True)


def test_decorators_keep_reference_semantics():
    from astropaint_b200 import LOS_integrate, interpolate
    from oracle import profiles_np as oprof

    def g3(r, a):
        return np.exp(-(r / a) ** 2)

    @LOS_integrate
    def g2(R, a):
        return g3(R, a)

    @interpolate(n_samples=12, sampling_method="logspace")
    @LOS_integrate
    def g2i(R, a):
        return g3(R, a)

    @interpolate
    def bare(R, a):
        return g3(R, a)

    assert inspect.getfullargspec(g2i).args == ["R", "a"]
    assert g2i._ap_interp["n_samples"] == 12 and g2i._ap_interp["sampling_method"] == "logspace"
    assert bare._ap_interp["n_samples"] == 20
    R = np.linspace(0.05, 2, 40)
    assert np.allclose(g2(R, 0.8), np.sqrt(np.pi) * 0.8 * np.exp(-(R / 0.8) ** 2), rtol=2e-2)   # analytic Abel pair
    want = oprof.interpolate(lambda RR, a: oprof.los_integrate(g3, RR, a), R, 0.8, n_samples=12,
                             sampling_method="logspace")
    assert np.allclose(g2i(R, 0.8), want, rtol=1e-12)
    assert np.allclose(g2i(R[:5], 0.8), g2(R[:5], 0.8))      # len(R) < n_samples: direct evaluation
    assert np.isscalar(float(g2(0.3, 0.8)))


def test_shard_rows_is_array_split():
    from astropaint_b200 import parallel
    for n, w in [(10, 3), (7, 8), (1000, 8), (0, 2)]:
        parts = [parallel.shard_rows(n, r, w) for r in range(w)]
        assert np.array_equal(np.concatenate(parts), np.arange(n))
        want = np.array_split(np.arange(n), w)
        assert all(np.array_equal(a, b) for a, b in zip(parts, want))


def test_band_plan_ownership_and_streaming_are_safe():
    """Ring-band ownership (_bands.Plan): (a) the bands of all ranks tile the map; (b) every halo whose disc
    touches a rank's band lies inside the contiguous slice of the colatitude order that rank paints; (c) after
    chunk c of a rank's schedule no LATER halo touches a ring the chunk declared final.  Checked against the
    oracle's pixel sets, including polar, sky-sized and inclusive-sized discs."""
    from astropaint_b200 import _bands
    from oracle import healpix_np as H

    rng = np.random.default_rng(0)
    for nside, ws, n_chunks in [(16, 1, 5), (16, 3, 2), (64, 8, 3), (64, 2, 6), (128, 5, 4)]:
        hp = H.HealpixRing(nside)
        # equal-area bands, or bands sized by per-rank weights (Canvas(band_balance="readback"))
        weights = None if nside != 64 else list(rng.uniform(0.5, 2.0, ws))
        bands = [_bands.Band(nside, r, ws, weights) for r in range(ws)]
        if weights is not None:
            share = np.array([b.npix for b in bands]) / hp.npix
            assert np.allclose(share, np.array(weights) / np.sum(weights), atol=2.0 / nside)
        assert bands[0].pix_lo == 0 and bands[-1].pix_hi == hp.npix
        assert all(bands[r].pix_hi == bands[r + 1].pix_lo for r in range(ws - 1))
        assert all(b.pix_lo == _bands.ring_start(nside, b.ring_lo) for b in bands)
        n = 300
        theta = np.arccos(rng.uniform(-1, 1, n))
        theta[:12] = [0.0, np.pi, 1e-9, np.pi - 1e-9] * 3
        phi = rng.uniform(0, 2 * np.pi, n)
        radius = rng.choice([1.0, 0.1, 3.0], n) * rng.uniform(0, 4.0 / nside, n) + (rng.uniform(0, 1, n) < 0.02) * 0.7
        order = np.argsort(theta, kind="stable")
        ts, rs = theta[order], radius[order]
        pix_sets = []
        for i in range(n):
            v = (np.sin(ts[i]) * np.cos(phi[order[i]]), np.sin(ts[i]) * np.sin(phi[order[i]]), np.cos(ts[i]))
            pix_sets.append(hp.query_disc(v, rs[i]))
        ring_starts = np.array([_bands.ring_start(nside, r) for r in range(1, 4 * nside + 1)])
        for b in bands:
            reach = (np.minimum.accumulate((ts - rs)[::-1])[::-1], np.maximum.accumulate(ts + rs))
            plan = _bands.Plan(b, reach, n_chunks=n_chunks)
            assert plan.chunks[0][0] == plan.halo_lo and plan.chunks[-1][1] == plan.halo_hi
            assert plan.chunks[-1][2] == b.ring_hi
            for i, q in enumerate(pix_sets):
                touches = len(q) and q.max() >= b.pix_lo and q.min() < b.pix_hi
                if touches:
                    assert plan.halo_lo <= i < plan.halo_hi, (nside, ws, b.rank, i)
            for (a, e, done) in plan.chunks:
                done_pix = _bands.ring_start(nside, done)
                for i in range(e, plan.halo_hi):
                    q = pix_sets[i]
                    q = q[(q >= b.pix_lo) & (q < b.pix_hi)]
                    if len(q):
                        assert q.min() >= done_pix, (nside, ws, b.rank, e, i, q.min(), done_pix)
        # ring helpers agree with the oracle's ring scheme
        for r in (1, nside - 1, nside, 2 * nside, 3 * nside, 3 * nside + 1, 4 * nside - 1):
            assert _bands.ring_start(nside, r) == hp.ring_info(r)[0]
            assert abs(_bands.ring_z(nside, r) - hp.ring2z(r)) < 1e-15


def test_fits_map_round_trip_and_layout(tmp_path):
    """Canvas.save_map_to_file / load_map_from_file (paint_bucket.py:1478-1558): healpy.write_map's default layout
    (float32 column TEMPERATURE, 1024 pixels per row, RING keywords) read back by healpy.read_map's restatement."""
    from astropaint_b200.lib import fits_map
    rng = np.random.default_rng(0)
    for nside in (4, 16, 64):
        m = rng.normal(size=12 * nside * nside)
        fn = str(tmp_path / f"map_{nside}.fits")
        fits_map.write_map(fn, m, overwrite=True)
        raw = open(fn, "rb").read()
        assert len(raw) % 2880 == 0 and raw[:30].decode().startswith("SIMPLE  =                    T")
        hdr = raw[2880:2880 * 2].decode()
        per_row = 1024 if 12 * nside * nside > 1024 else 1
        for card in ("XTENSION= 'BINTABLE'", "PIXTYPE = 'HEALPIX '", "ORDERING= 'RING    '", f"TFORM1  = '{per_row}E",
                     "TTYPE1  = 'TEMPERATURE'", "INDXSCHM= 'IMPLICIT'"):
            assert card in hdr, card
        assert f"{nside:>20}" in hdr[hdr.index("NSIDE   ="):hdr.index("NSIDE   =") + 80]
        back = fits_map.read_map(fn)
        assert back.dtype == np.float64 and np.array_equal(back, m.astype(np.float32).astype(np.float64))
        with pytest.raises(OSError):
            fits_map.write_map(fn, m, overwrite=False)
        fits_map.write_map(fn, m, overwrite=True, dtype=np.float64)
        assert np.array_equal(fits_map.read_map(fn), m)
    with pytest.raises(ValueError):
        fits_map.write_map(str(tmp_path / "bad.fits"), np.zeros(13))


def test_catalog_order_and_staging_follow_in_place_edits():
    """The colatitude order, the disc reach and the page-locked staging are caches of catalog columns: an in-place
    edit of the geometry (or of a staged column) must invalidate them, not be painted from stale copies."""
    import astropaint_b200 as ap
    from astropaint_b200 import synthetic
    cat = ap.Catalog(synthetic.websky_like(5000, seed=2))
    order0 = cat.theta_order()[0].copy()
    reach0 = cat.disc_reach(5)[0].copy()
    assert cat.theta_order()[0] is cat.theta_order()[0]                     # cached
    cat.data["z"] = -cat.data["z"].values                                   # replaced column: mirror the sky
    order1 = cat.theta_order()[0]
    assert not np.array_equal(order0, order1)
    assert not np.array_equal(reach0, cat.disc_reach(5)[0])
    z = cat.data["z"].values
    if z.flags.writeable:
        z *= 0.5                                                            # in-place bulk edit of the same array
        cat.theta_order()
        assert cat._order_fp == cat._geometry_fingerprint()
    fp = ap.Catalog._fingerprint(np.arange(1000.0))
    assert fp != ap.Catalog._fingerprint(np.arange(1000.0) * 2)


def test_redshift_table_inversion_equals_root_finding():
    """Catalog(calculate_redshifts=True): the table inversion of D_c(z) used for catalogs agrees with the per-halo root
    finder (the reference's z_at_value, lib/transform.py:116-119) far inside any tolerance of the path."""
    from astropaint_b200.lib.cosmology import cosmo
    rng = np.random.default_rng(1)
    D = np.concatenate([rng.uniform(1.0, 9000.0, 40), [0.5, 150.0, 7734.0]])
    table = cosmo.z_at_comoving_distance(D)                       # > 32 values: table path
    exact = np.array([cosmo.z_at_comoving_distance(float(d)) for d in D[:12]])
    assert np.abs(table[:12] - exact).max() < 1e-9
    assert np.all(np.diff(cosmo.z_at_comoving_distance(np.linspace(1, 9000, 100))) > 0)
    back = np.array([cosmo.comoving_distance(float(zz)) for zz in table[:8]])
    assert np.abs(back - D[:8]).max() < 1e-6 * D[:8].max()
