"""GPU parity tests: the CUDA path, called through the reference-shaped API and the C ABI, against
the CPU oracle (oracle/) on the same seeded inputs.  Bar: pixel sets bit-exact, map values within
1e-6 relative (fp64), as BASELINE.json's north_star states."""
import numpy as np
import pytest

from .conftest import map_close, apply_func_to_cutout

pytestmark = pytest.mark.gpu


def _oracle():
    from oracle import reference_loop as ref, profiles_np as oprof, healpix_np as ohp
    return ref, oprof, ohp


def _catalogs():
    import astropaint_b200 as ap
    from astropaint_b200 import synthetic
    return {
        "test6": ap.Catalog("test"),
        "box": ap.Catalog(synthetic.random_box(60, seed=0)),
        "websky": ap.Catalog(synthetic.websky_like(300, seed=1)),
    }


@pytest.fixture(scope="module")
def cats():
    return _catalogs()


# ------------------------------------------------------------------ pixel sets and geometry
@pytest.mark.parametrize("name,nside,R_times", [("test6", 64, 1), ("test6", 64, 20), ("test6", 16, 80),
                                                ("box", 128, 5), ("box", 512, 5), ("box", 32, 1),
                                                ("websky", 2048, 5), ("websky", 8192, 5), ("websky", 4096, 1)])
def test_pixel_sets_bit_exact(cats, name, nside, R_times):
    import astropaint_b200 as ap
    ref, _, _ = _oracle()
    cat = cats[name]
    canvas = ap.Canvas(cat, nside, R_times=R_times)
    discs = ref.Discs(cat.data, nside, R_times)
    n = 0
    for halo, pix in enumerate(canvas.discs.gen_pixel_index()):
        want = discs.pixel_index(halo)
        assert pix.dtype == np.int64
        assert np.array_equal(pix, want), f"halo {halo}: {len(pix)} vs {len(want)} pixels"
        n += len(pix)
    assert n > 0


@pytest.mark.parametrize("name,nside,R_times", [("test6", 64, 5), ("box", 128, 5), ("websky", 4096, 5)])
def test_R_and_R_vec(cats, name, nside, R_times):
    import astropaint_b200 as ap
    ref, _, _ = _oracle()
    cat = cats[name]
    canvas = ap.Canvas(cat, nside, R_times=R_times)
    discs = ref.Discs(cat.data, nside, R_times)
    for halo, (R, Rv) in enumerate(zip(canvas.discs.gen_cent2pix_mpc(), canvas.discs.gen_cent2pix_mpc_vec())):
        want = discs.cent2pix_mpc(halo)
        wantv = discs.cent2pix_mpc_vec(halo)
        if len(want) == 0:
            assert len(R) == 0
            continue
        D_a = cat.data.D_a[halo]
        assert np.abs(R - want).max() <= 1e-9 * max(want.max(), 1e-300)
        assert np.abs(Rv - wantv).max() <= 1e-12 * D_a


# ------------------------------------------------------------------ spray: device profiles
def _spray_both(cat, nside, R_times, template, otemplate, **kw):
    import astropaint_b200 as ap
    ref, _, _ = _oracle()
    canvas = ap.Canvas(cat, nside, R_times=R_times)
    painter = ap.Painter(template)
    painter.spray(canvas, **kw)
    want = np.zeros(12 * nside * nside)
    n_pairs = ref.spray(want, cat.data, nside, otemplate, R_times=R_times, **kw)
    got = canvas.pixels
    assert int(painter.last_spray["d_count"].item()) == n_pairs
    return got, want, painter


def test_constant_density_is_coverage_count(cats):
    from astropaint_b200.profiles import spherical
    _, oprof, _ = _oracle()
    got, want, p = _spray_both(cats["box"], 128, 5, spherical.constant_density_2D, oprof.constant_density_2D,
                               constant=1.0)
    assert p.last_spray["path"] == "device_profile"
    assert np.array_equal(got, want)  # sums of 1.0 are exact: pins pixel sets and multiplicity
    assert got.max() > 1


PROFILE_CASES = [
    ("linear", "box", 128, 5, dict(intercept=0.5, slope=2.0)),
    ("gaussian", "box", 512, 5, {}),
    ("gaussian", "test6", 64, 5, {}),
    ("solid_sphere", "box", 128, 1, {}),
    ("solid_sphere", "box", 128, 2, {}),      # NaN outside R_200c, exactly like the reference
    ("nfw_rho", "box", 256, 3, {}),
    ("nfw_tau", "websky", 4096, 5, {}),
    ("nfw_ksz", "websky", 4096, 5, {}),
    ("nfw_ksz", "websky", 4096, 5, dict(T_cmb=3.0)),
    ("nfw_deflection", "box", 128, 5, {}),
    ("nfw_bg", "websky", 4096, 5, {}),
    ("nfw_bg", "box", 128, 5, {}),
]


def _pair(kind):
    from astropaint_b200.profiles import spherical, NFW
    _, o, _ = _oracle()
    return {
        "linear": (spherical.linear_density_2D, o.linear_density_2D),
        "gaussian": (spherical.gaussian_2D, o.gaussian_readme),
        "solid_sphere": (spherical.solid_sphere_2D, o.solid_sphere_2D),
        "nfw_rho": (NFW.rho_2D_bartlemann, o.nfw_rho_2D_bartlemann),
        "nfw_tau": (NFW.tau_2D, o.nfw_tau_2D),
        "nfw_ksz": (NFW.kSZ_T, o.nfw_kSZ_T),
        "nfw_deflection": (NFW.deflection_angle, o.nfw_deflection_angle),
        "nfw_bg": (NFW.BG, o.nfw_BG),
    }[kind]


@pytest.mark.parametrize("kind,name,nside,R_times,kw", PROFILE_CASES)
def test_device_profiles_match_oracle(cats, kind, name, nside, R_times, kw):
    t, o = _pair(kind)
    got, want, p = _spray_both(cats[name], nside, R_times, t, o, **kw)
    assert p.last_spray["path"] == "device_profile"
    map_close(got, want, 1e-6)


def test_spray_is_additive(cats):
    import astropaint_b200 as ap
    from astropaint_b200.profiles import spherical
    canvas = ap.Canvas(cats["box"], 64, R_times=3)
    painter = ap.Painter(spherical.gaussian_2D)
    painter.spray(canvas)
    once = canvas.pixels.copy()
    painter.spray(canvas)          # host array handed out -> re-uploaded, then painted on top
    map_close(canvas.pixels, 2 * once, 1e-12)
    canvas.clean()
    assert (canvas.pixels == 0).all()


# ------------------------------------------------------------------ Battaglia16 (device QUADPACK + spline)
def test_device_quadpack_matches_scipy():
    """The device qagie against scipy.integrate.quad itself: value, error estimate AND number of
    integrand evaluations (i.e. the same adaptive decisions) on 3000 (halo, R) pairs."""
    import torch
    from scipy import integrate
    from astropaint_b200 import _capi
    from astropaint_b200._capi import ptr
    ref, o, _ = _oracle()
    rng = np.random.default_rng(7)
    n = 3000
    M = 10 ** rng.uniform(13, 15.5, n)
    z = rng.uniform(0, 4.6, n)
    R200 = np.asarray(ref.M_200c_to_R_200c(M, z))
    R = R200 * rng.uniform(0, 5, n) * rng.choice([1.0, 1.0, 0.01, 1e-4], n)
    R = np.maximum(R, 1e-9)
    dev = torch.device("cuda", 0)
    t = [torch.from_numpy(a).to(dev) for a in (R, R200, M, z)]
    res = torch.empty(n, dtype=torch.float64, device=dev)
    err = torch.empty_like(res)
    nev = torch.empty(n, dtype=torch.int32, device=dev)
    _capi.check(_capi.lib().ap_b16_tau_eval(n, *[ptr(a) for a in t], ptr(res), ptr(err), ptr(nev), None))
    res, err, nev = res.cpu().numpy(), err.cpu().numpy(), nev.cpu().numpy()
    import warnings
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        for i in range(n):
            f = lambda r: o.b16_tau_3D(r, R200[i], M[i], z[i]) * 2. * r / np.sqrt(r ** 2 - R[i] ** 2)  # noqa: E731
            v, e, info = integrate.quad(f, R[i], np.inf, epsabs=0., epsrel=1.e-2, full_output=1)[:3]
            assert nev[i] == info["neval"], (i, R[i] / R200[i], nev[i], info["neval"])
            # same decisions (neval) but not the same bits: the integrand is formed as exp/log instead of pow
            # and the epsilon algorithm (dqelg: 1/(1/d1 + 1/d2 - 1/d3) of near-equal differences) amplifies
            # last-bit differences to ~1e-9 on a few quadratures (3e7 device quadratures, two builds whose
            # integrands differ by one rounding: max relative difference 1.4e-8) -- bar: 1e-7, map bar 1e-6
            assert abs(res[i] - v) <= 1e-7 * abs(v)
            assert abs(err[i] - e) <= 1e-3 * abs(e) + 1e-300   # error estimates are differences of near-equal sums


def test_b16_tau_2D_direct(cats):
    import astropaint_b200 as ap
    from astropaint_b200.profiles import Battaglia16
    _, o, _ = _oracle()
    cat = ap.Catalog(cats["websky"].data[["x", "y", "z", "v_x", "v_y", "v_z", "M_200c", "redshift"]].iloc[:40])
    got, want, p = _spray_both(cat, 2048, 5, Battaglia16.tau_2D, o.b16_tau_2D)
    assert p.last_spray["path"] == "device_profile"
    map_close(got, want, 1e-6)


@pytest.mark.parametrize("nside,n,m_min", [(4096, 120, 1e13), (1024, 60, 1e14), (8192, 40, 3e14)])
def test_b16_tau_2D_interp_and_kSZ(nside, n, m_min):
    """BASELINE configs 2/3 on sub-samples: discs both above and below the 15-sample bypass."""
    import astropaint_b200 as ap
    from astropaint_b200 import synthetic
    from astropaint_b200.profiles import Battaglia16
    _, o, _ = _oracle()
    cat = ap.Catalog(synthetic.websky_like(n, seed=11, m_min=m_min))
    got, want, p = _spray_both(cat, nside, 5, Battaglia16.tau_2D_interp, o.b16_tau_2D_interp)
    assert p.last_spray["path"] == "device_table"
    map_close(got, want, 1e-6)
    got, want, p = _spray_both(cat, nside, 5, Battaglia16.kSZ_T, o.b16_kSZ_T)
    assert p.last_spray["path"] == "device_table"
    map_close(got, want, 1e-6)


# ------------------------------------------------------------------ Python templates
def test_python_template_flat_path(cats):
    """The README's user-defined template (README.md:99-101) as plain Python."""
    def a_nonsense_template(R, R_200c, x, y, z):
        return np.exp(-(R / R_200c / 3) ** 2) * (x + y + z)

    _, o, _ = _oracle()
    got, want, p = _spray_both(cats["box"], 128, 5, a_nonsense_template, o.gaussian_readme)
    assert p.last_spray["path"] == "flat"
    # summation order differs (atomics vs sequential np.add.at): pixels where halos of opposite
    # sign cancel lose relative precision, so the bar is the map tolerance, not 1e-9
    map_close(got, want, 1e-6)


def test_python_template_scalar_and_kwargs(cats):
    def shifted(R, R_200c, *, offset=0.0, gain=1.0):
        return gain * (R / R_200c) + offset

    gains = np.linspace(1, 2, cats["box"].size)
    got, want, p = _spray_both(cats["box"], 64, 2, shifted, shifted, offset=0.25, gain=gains)
    map_close(got, want, 1e-6)


def test_python_R_vec_template(cats):
    def dipole(R_vec, v_x, v_y, v_z):
        return R_vec[:, 0] * v_x + R_vec[:, 1] * v_y + R_vec[:, 2] * v_z

    got, want, p = _spray_both(cats["box"], 64, 2, dipole, dipole)
    map_close(got, want, 1e-6)


def test_python_interpolate_template(cats):
    """README.md:170-206: @interpolate(n_samples=20) @LOS_integrate top-hat; the analytic answer
    2 sqrt(R_200c^2 - R^2) is the known-answer check of the projection operator."""
    import astropaint_b200 as ap
    from astropaint_b200 import LOS_integrate, interpolate
    from scipy import integrate
    from scipy.interpolate import InterpolatedUnivariateSpline

    def tophat_3D(r, R_200c):
        return np.where(np.asarray(r) > R_200c, 0.0, 1.0)

    @interpolate(n_samples=20)
    @LOS_integrate
    def tophat_2D_interp(R, R_200c):
        return tophat_3D(R, R_200c)

    def o_tophat(R, R_200c):  # the reference decorators, spelled out
        def los(Rs):
            scalar = not hasattr(Rs, "__iter__")
            out = [integrate.quad(lambda r: tophat_3D(r, R_200c) * 2. * r / np.sqrt(r ** 2 - R_i ** 2),
                                  R_i, np.inf, epsabs=0., epsrel=1.e-2)[0] for R_i in ([Rs] if scalar else Rs)]
            return np.asarray(out[0] if scalar else out)
        if len(R) < 20:
            return los(R)
        xs = np.linspace(R.min(), R.max(), 20)
        return InterpolatedUnivariateSpline(xs, np.array([los(x) for x in xs]), k=3)(R)

    cat = ap.Catalog(cats["box"].data[["x", "y", "z", "v_x", "v_y", "v_z", "M_200c"]].iloc[:12])
    import warnings
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        got, want, p = _spray_both(cat, 64, 1, tophat_2D_interp, o_tophat)
    assert p.last_spray["path"] == "python_table"
    map_close(got, want, 1e-6)


# ------------------------------------------------------------------ cutouts / stack
def test_cutouts_and_stack_match_oracle(cats):
    import astropaint_b200 as ap
    from astropaint_b200.profiles import spherical
    ref, _, _ = _oracle()
    cat = cats["box"]
    nside = 256
    canvas = ap.Canvas(cat, nside, R_times=5)
    ap.Painter(spherical.gaussian_2D).spray(canvas)
    pixels = canvas.pixels.copy()
    halos = np.arange(0, 40, 3)
    for lon_range, lat_range, xpix, ypix in [([-1, 1], None, 200, None), ([-3, 2], [-1, 2.5], 50, None),
                                             ([-0.2, 0.2], [-0.2, 0.2], 17, 17)]:
        got = list(canvas.cutouts(halos, lon_range=lon_range, lat_range=lat_range, xpix=xpix, ypix=ypix))
        want = list(ref.cutouts(pixels, cat.data, nside, halos, lon_range, lat_range, xpix, ypix))
        assert len(got) == len(want) == len(halos)
        for g, w in zip(got, want):
            assert g.shape == w.shape
            assert np.array_equal(g, w)   # nearest-pixel lookups: identical pixels -> identical values
    stack = canvas.stack_cutouts(halos, lon_range=[-1, 1], xpix=64, inplace=False)
    want = ref.stack_cutouts(pixels, cat.data, nside, halos, [-1, 1], None, 64)
    map_close(stack, want, 1e-12)
    canvas.stack_cutouts("all", lon_range=[-0.5, 0.5], xpix=32)
    want = ref.stack_cutouts(pixels, cat.data, nside, None, [-0.5, 0.5], None, 32)
    map_close(canvas.stack, want, 1e-12)


@pytest.mark.parametrize("mult_by, add_to", [(1, 1), (1, [1]), ([1], [1]), (1, [1, 2, 3, 4, 5, 6])])
def test_cutouts_apply_func_kwarg_types(test_canvas, mult_by, add_to):
    """Reference tests/test_Canvas.py:64-83, plus value checks."""
    ref, _, _ = _oracle()
    halo_list = test_canvas.catalog.data.index
    cuts = list(test_canvas.cutouts(halo_list, apply_func=apply_func_to_cutout,
                                    func_kwargs={"mult_by": mult_by, "add_to": add_to}))
    assert len(cuts) == 6
    a = np.atleast_1d(add_to)
    for h, c in enumerate(cuts):
        assert c.shape == (200, 200)
        assert np.all(c == (a[h] if len(a) > 1 else a[0]))   # canvas is empty


def test_cutouts_apply_multiple_funcs(test_canvas):
    halo_list = test_canvas.catalog.data.index
    cuts = list(test_canvas.cutouts(halo_list, apply_func=[apply_func_to_cutout] * 2,
                                    func_kwargs=[{"mult_by": 1, "add_to": 1}] * 2))
    assert all(np.all(c == 2) for c in cuts)


def test_stack_with_apply_func(cats):
    import astropaint_b200 as ap
    from astropaint_b200.profiles import spherical
    ref, _, _ = _oracle()
    cat = cats["box"]
    canvas = ap.Canvas(cat, 128, R_times=5)
    ap.Painter(spherical.gaussian_2D).spray(canvas)
    pixels = canvas.pixels.copy()
    sign = np.sign(cat.data.v_r.values)
    halos = np.arange(20)
    flip = lambda patch, s: patch * s  # noqa: E731
    got = canvas.stack_cutouts(halos, lon_range=[-2, 2], xpix=40, inplace=False, apply_func=flip,
                               func_kwargs={"s": sign})
    want = ref.stack_cutouts(pixels, cat.data, 128, halos, [-2, 2], None, 40, apply_func=flip,
                             func_kwargs={"s": sign})
    map_close(got, want, 1e-12)


# ------------------------------------------------------------------ API errors
def test_spray_argument_errors(cats):
    import astropaint_b200 as ap
    from astropaint_b200.profiles import spherical, NFW
    canvas = ap.Canvas(cats["test6"], 16)
    p = ap.Painter(spherical.gaussian_2D)
    with pytest.raises(ValueError):
        p.spray(canvas, n_cpus=0)
    with pytest.raises(AssertionError):
        p.spray(canvas, distance_units="furlongs")
    with pytest.raises(AssertionError):
        ap.Painter(NFW.BG).spray(canvas, cache=True)
    with pytest.raises(TypeError):
        ap.Painter(spherical.constant_density_2D).spray(canvas)     # `constant` not provided
    with pytest.raises(AssertionError):
        ap.Painter(lambda r: r)                                     # neither R nor R_vec
    p.spray(canvas, n_cpus=-1)


# ------------------------------------------------------------------ raw C ABI
def test_c_abi_error_codes_and_host_entry(cats):
    import ctypes as C
    import torch
    from astropaint_b200 import _capi, _engine as eng
    ref, oprof, _ = _oracle()
    L = _capi.lib()
    cat = cats["box"]
    dev = torch.device("cuda", 0)
    halos = eng.DeviceHalos(cat.data, 5, dev)
    s = halos.struct()
    n_pix = torch.empty(halos.n, dtype=torch.int64, device=dev)
    assert L.ap_disc_count(0, C.byref(s), 0, _capi.ptr(n_pix), None, None, None, None) != 0
    assert b"nside" in L.ap_last_error()
    n_inc = torch.empty_like(n_pix)                                        # inclusive != 0: healpy's inclusive query
    assert L.ap_disc_count(64, C.byref(s), 1, _capi.ptr(n_inc), None, None, None, None) == 0
    assert L.ap_disc_count(64, C.byref(s), 0, _capi.ptr(n_pix), None, None, None, None) == 0
    assert bool((n_inc >= n_pix).all()) and bool((n_inc > n_pix).any())
    assert L.ap_disc_count(64, C.byref(s), 0, None, None, None, None, None) != 0
    d_map = torch.zeros(12 * 64 * 64, dtype=torch.float64, device=dev)
    assert L.ap_paint_profile(3, 64, C.byref(s), None, None, 2, _capi.ptr(d_map), None, None) != 0
    assert L.ap_paint_profile(77, 64, C.byref(s), None, None, 0, _capi.ptr(d_map), None, None) != 0
    empty = _capi.ApHalos()
    assert L.ap_disc_count(64, C.byref(empty), 0, _capi.ptr(n_pix), None, None, None, None) == 0
    # host-buffer entry point: Gaussian template, all HOST pointers
    nside = 64
    d = cat.data
    cols = [np.ascontiguousarray(d[k].values, dtype=np.float64) for k in ("x", "y", "z")]
    radius = np.ascontiguousarray(5 * np.deg2rad(d.R_ang_200c.values / 60))
    geo = [np.ascontiguousarray(d[k].values, dtype=np.float64) for k in ("theta", "phi", "D_a")]
    params = [np.ascontiguousarray(d[k].values, dtype=np.float64) for k in ("R_200c", "x", "y", "z")]
    arr = (C.c_void_p * 4)(*[p.ctypes.data for p in params])
    strides = (C.c_int32 * 4)(1, 1, 1, 1)
    h_map = np.zeros(12 * nside * nside)
    cnt = C.c_uint64(0)
    rc = L.ap_spray_host(3, nside, len(d), *[_capi.ptr(c) for c in cols], _capi.ptr(radius),
                         *[_capi.ptr(g) for g in geo], arr, strides, 4, _capi.ptr(h_map), C.byref(cnt))
    assert rc == 0, L.ap_last_error()
    want = np.zeros_like(h_map)
    n_pairs = ref.spray(want, d, nside, oprof.gaussian_readme, R_times=5)
    assert cnt.value == n_pairs
    map_close(h_map, want, 1e-9)


# ------------------------------------------------------------------ full-size properties (nside 8192)
def test_full_size_properties():
    """Size-independent properties at BASELINE's nside 8192 on a 2e5-halo Websky-shaped catalog:
    coverage count is integral and sums to the per-halo disc counts; linearity of the painter."""
    import torch
    import astropaint_b200 as ap
    from astropaint_b200 import synthetic, _engine as eng
    from astropaint_b200.profiles import spherical, NFW
    nside = 8192
    cat = ap.Catalog(synthetic.websky_like(200000, seed=2))
    canvas = ap.Canvas(cat, nside, R_times=5)
    p = ap.Painter(spherical.constant_density_2D)
    p.spray(canvas, constant=1.0)
    dmap = canvas.pixels_device
    total = int(p.last_spray["d_count"].item())
    halos = eng.DeviceHalos(cat.data, 5, dmap.device)
    n_pix, _, _, _ = eng.disc_count(halos, nside)
    assert total == int(n_pix.sum().item())
    assert float(dmap.sum().item()) == float(total)
    assert bool((dmap == dmap.round()).all())
    # linearity: spray(kSZ with T) + spray(kSZ with 2T) == spray(kSZ with 3T)
    a = ap.Canvas(cat, nside, R_times=5)
    ap.Painter(NFW.kSZ_T).spray(a, T_cmb=1.0)
    ap.Painter(NFW.kSZ_T).spray(a, T_cmb=2.0)
    b = ap.Canvas(cat, nside, R_times=5)
    ap.Painter(NFW.kSZ_T).spray(b, T_cmb=3.0)
    da, db = a.pixels_device, b.pixels_device
    scale = float(db.abs().max().item())
    assert float((da - db).abs().max().item()) <= 1e-9 * scale
    del da, db, dmap
    torch.cuda.empty_cache()


# ------------------------------------------------------------------ committed golden fixtures
def test_cuda_path_matches_golden_fixtures():
    """tests/golden/golden_small.npz (made by tests/golden/make_golden.py from the oracle)."""
    import hashlib
    import os
    import astropaint_b200 as ap
    from astropaint_b200 import synthetic
    from astropaint_b200.profiles import spherical, NFW, Battaglia16
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "golden_small.npz"))
    cat = ap.Catalog("test")
    canvas = ap.Canvas(cat, 64, R_times=5)
    pix = list(canvas.discs.gen_pixel_index())
    assert np.array_equal(g["test6_counts"], [len(p) for p in pix])
    sha = np.frombuffer(hashlib.sha256(np.concatenate(pix).tobytes()).digest(), dtype=np.uint8)
    assert np.array_equal(g["test6_sha256"], sha)
    ap.Painter(spherical.gaussian_2D).spray(canvas)
    map_close(canvas.pixels, g["test6_gaussian_map"], 1e-6)
    cat = ap.Catalog(synthetic.random_box(50, seed=0))
    canvas = ap.Canvas(cat, 128, R_times=5)
    p = ap.Painter(spherical.gaussian_2D)
    p.spray(canvas)
    assert int(p.last_spray["d_count"].item()) == int(g["box_pairs"])
    map_close(canvas.pixels, g["box_gaussian_map"], 1e-6)
    stack = canvas.stack_cutouts(np.arange(10), lon_range=[-2, 2], xpix=32, inplace=False)
    map_close(stack, g["box_stack"], 1e-6)
    for tmpl, key in [(NFW.kSZ_T, "box_nfw_ksz_map"), (NFW.BG, "box_nfw_bg_map")]:
        canvas = ap.Canvas(cat, 128, R_times=3)
        ap.Painter(tmpl).spray(canvas)
        map_close(canvas.pixels, g[key], 1e-6)
    cat = ap.Catalog(synthetic.websky_like(16, seed=5, m_min=2e14))
    canvas = ap.Canvas(cat, 1024, R_times=5)
    p = ap.Painter(Battaglia16.kSZ_T)
    p.spray(canvas)
    assert int(p.last_spray["d_count"].item()) == int(g["websky_pairs"])
    want = np.zeros(12 * 1024 * 1024)
    want[g["websky_ksz_idx"]] = g["websky_ksz_val"]
    map_close(canvas.pixels, want, 1e-6)


def test_smoke_entry():
    import __graft_entry__ as ge
    ge.smoke()


def test_uniform_spline_kernel_equals_generic_and_scipy():
    import torch
    from scipy.interpolate import InterpolatedUnivariateSpline
    from astropaint_b200 import _engine as eng
    dev = torch.device("cuda", 0)
    rng = np.random.default_rng(3)
    for ns in (15, 20, 9):
        n = 1000
        a = rng.uniform(0.01, 1.0, n)
        b = a + rng.uniform(0.1, 5.0, n)
        nodes = np.stack([np.linspace(a[i], b[i], ns) for i in range(n)])
        vals = np.exp(-nodes) * (1 + 0.05 * rng.normal(size=nodes.shape))
        nn = np.full(n, ns, dtype=np.int32)
        nn[::7] = 0
        t_nodes, t_vals = torch.from_numpy(nodes).to(dev), torch.from_numpy(vals).to(dev)
        t_nn = torch.from_numpy(nn).to(dev)
        c_gen = eng.spline_build(t_nodes, t_vals, t_nn, uniform=False).cpu().numpy()
        c_uni = eng.spline_build(t_nodes, t_vals, t_nn, uniform=True).cpu().numpy()
        scale = np.abs(c_gen).max(axis=(1, 2), keepdims=True) + 1e-300
        assert np.abs(c_uni - c_gen).max() <= 1e-9 * scale.max()
        assert np.all(c_uni[::7] == 0)
        for i in (1, 2, 500):
            X = np.linspace(a[i], b[i], 50)
            k = np.clip(np.searchsorted(nodes[i], X, side="right") - 1, 0, ns - 2)
            t = X - nodes[i][k]
            c = c_uni[i][k]
            got = c[:, 0] + t * (c[:, 1] + t * (c[:, 2] + t * c[:, 3]))
            want = InterpolatedUnivariateSpline(nodes[i], vals[i], k=3)(X)
            assert np.abs(got - want).max() <= 1e-10 * np.abs(want).max()


@pytest.mark.parametrize("kind,name,nside", [("gaussian", "box", 128), ("nfw_ksz", "websky", 4096), ("nfw_bg", "box", 128)])
def test_fp32_accumulate_mode(cats, kind, name, nside):
    """BASELINE north_star: map values within 1e-4 in fp32-accumulate mode (max-abs, relative to the map's
    maximum: single-precision sums of opposite-sign contributions cannot hold a per-pixel relative bar)."""
    import astropaint_b200 as ap
    ref, _, _ = _oracle()
    t, o = _pair(kind)
    cat = cats[name]
    canvas = ap.Canvas(cat, nside, R_times=5, accumulate="fp32")
    ap.Painter(t).spray(canvas)
    got = canvas.pixels
    assert got.dtype == np.float32
    want = np.zeros(12 * nside * nside)
    ref.spray(want, cat.data, nside, o, R_times=5)
    assert np.abs(got - want).max() <= 1e-4 * np.abs(want).max()
    stack = canvas.stack_cutouts(np.arange(5), lon_range=[-1, 1], xpix=16, inplace=False)
    assert stack.shape == (16, 16) and np.isfinite(stack).all()


def test_fp32_accumulate_battaglia_and_python_paths():
    import astropaint_b200 as ap
    from astropaint_b200 import synthetic
    from astropaint_b200.profiles import Battaglia16
    ref, o, _ = _oracle()
    cat = ap.Catalog(synthetic.websky_like(80, seed=11, m_min=1e14))
    canvas = ap.Canvas(cat, 1024, R_times=5, accumulate="fp32")
    ap.Painter(Battaglia16.kSZ_T).spray(canvas)
    want = np.zeros(12 * 1024 * 1024)
    ref.spray(want, cat.data, 1024, o.b16_kSZ_T, R_times=5)
    assert np.abs(canvas.pixels - want).max() <= 1e-4 * np.abs(want).max()

    def tmpl(R, R_200c):
        return 1.0 / (1.0 + (R / R_200c) ** 2)

    canvas = ap.Canvas(cat, 1024, R_times=5, accumulate="fp32")
    ap.Painter(tmpl).spray(canvas)
    want = np.zeros(12 * 1024 * 1024)
    ref.spray(want, cat.data, 1024, tmpl, R_times=5)
    assert np.abs(canvas.pixels - want).max() <= 1e-4 * np.abs(want).max()


def test_device_radius_is_numpy_radius(cats):
    """the disc radius is formed on the device; it must round exactly like the reference's numpy
    expression R_times * np.deg2rad(R_ang_200c / 60) (paint_bucket.py:1226-1227)"""
    import torch
    from astropaint_b200 import _engine as eng
    from astropaint_b200 import synthetic
    import astropaint_b200 as ap
    big = ap.Catalog(synthetic.websky_like(200000, seed=9))
    for cat in list(cats.values()) + [big]:
        for R_times in (1, 5, 2.5):
            halos = eng.DeviceHalos(cat.data, R_times, torch.device("cuda", 0))
            want = R_times * np.deg2rad(cat.data.R_ang_200c.values / 60)
            assert np.array_equal(halos.cols["__radius__"].cpu().numpy(), want)


def test_disc_query_cuda_vs_glibc_at_scale():
    """5e5 Websky-shaped discs at nside 8192 (4e7 halo-pixels, 1.1e7 ring boundaries): per-halo pixel
    count and sum of pixel ids from the device (CUDA libm) equal those of the same header compiled for
    the host (glibc) — the near-tie scan of DESIGN.md §5."""
    import ctypes as C
    import os
    import torch
    import astropaint_b200 as ap
    from astropaint_b200 import synthetic, _engine as eng
    N, nside = 500000, 8192
    L = C.CDLL(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "astropaint_b200", "_native",
                            "libap_hostcheck.so"))
    L.hc_query_disc_stats.argtypes = [C.c_int64, C.c_int64] + [C.c_void_p] * 6
    cat = ap.Catalog(synthetic.websky_like(N, seed=21))
    dev = torch.device("cuda", 0)
    halos = eng.DeviceHalos(cat.data, 5, dev)
    n_pix, _, _, _ = eng.disc_count(halos, nside)
    off, pix, _, _ = eng.disc_pixels(halos, nside, n_pix)
    seg = torch.repeat_interleave(torch.arange(N, device=dev), n_pix)
    sum_d = torch.zeros(N, dtype=torch.int64, device=dev).index_add_(0, seg, pix).cpu().numpy()
    cnt_d = n_pix.cpu().numpy()
    d = cat.data
    x, y, z = (np.ascontiguousarray(d[k].values) for k in "xyz")
    rad = np.ascontiguousarray(5 * np.deg2rad(d.R_ang_200c.values / 60))
    cnt_h, sum_h = np.zeros(N, dtype=np.int64), np.zeros(N, dtype=np.int64)
    L.hc_query_disc_stats(nside, N, x.ctypes.data, y.ctypes.data, z.ctypes.data, rad.ctypes.data, cnt_h.ctypes.data,
                          sum_h.ctypes.data)
    assert np.array_equal(cnt_d, cnt_h) and np.array_equal(sum_d, sum_h)
    assert cnt_d.sum() > 3e7


@pytest.mark.parametrize("kind,nside,n,m_min", [("nfw_rho_2D", 2048, 30, 1e14), ("nfw_rho_2D_interp", 2048, 80, 1e14),
                                               ("nfw_rho_2D_interp", 8192, 40, 3e14),
                                               ("b16_rho_gas_2D", 2048, 30, 1e14), ("b16_ne_2D", 2048, 80, 1e14),
                                               ("b16_ne_2D", 8192, 40, 3e14)])
def test_los_projected_templates_on_device(kind, nside, n, m_min):
    """the remaining @LOS_integrate / @interpolate built-ins (NFW.rho_2D, rho_2D_interp; Battaglia16
    rho_gas_2D, ne_2D): device QUADPACK for every 3-D profile, log-spaced tables included"""
    import warnings
    import astropaint_b200 as ap
    from astropaint_b200 import synthetic
    from astropaint_b200.profiles import NFW, Battaglia16
    _, o, _ = _oracle()
    t, ot, path = {"nfw_rho_2D": (NFW.rho_2D, o.nfw_rho_2D, "device_profile"),
                   "nfw_rho_2D_interp": (NFW.rho_2D_interp, o.nfw_rho_2D_interp, "device_table"),
                   "b16_rho_gas_2D": (Battaglia16.rho_gas_2D, o.b16_rho_gas_2D, "device_profile"),
                   "b16_ne_2D": (Battaglia16.ne_2D, o.b16_ne_2D, "device_table")}[kind]
    cat = ap.Catalog(synthetic.websky_like(n, seed=13, m_min=m_min))
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        got, want, p = _spray_both(cat, nside, 5, t, ot)
    assert p.last_spray["path"] == path
    assert np.abs(want).max() > 0
    map_close(got, want, 1e-6)


def test_prefetch_pipeline_matches_plain_spray(monkeypatch):
    """colatitude-ordered chunks with device->host streaming of the finished rings == the plain path"""
    import astropaint_b200 as ap
    from astropaint_b200 import synthetic, _bands
    from astropaint_b200.profiles import Battaglia16
    cat = ap.Catalog(synthetic.websky_like(4000, seed=17, m_min=5e13))
    canvas = ap.Canvas(cat, 2048, R_times=5)
    p = ap.Painter(Battaglia16.kSZ_T)
    p.spray(canvas)
    assert p.last_spray["path"] == "device_table"
    plain = canvas.pixels.copy()              # first read allocates the pinned host buffer
    n_plain = int(p.last_spray["d_count"].item())
    monkeypatch.setattr(ap.Painter, "PREFETCH_MIN_HALOS", 0)
    monkeypatch.setattr(_bands, "MIN_CHUNK_HALOS", 500)
    canvas.clean()
    p.spray(canvas)
    assert p.last_spray["path"] == "device_table+stream" and p.last_spray["chunks"] == 8
    assert int(p.last_spray["d_count"].item()) == n_plain
    got = canvas.pixels
    map_close(got, plain, 1e-12)
    # additive second spray on the (host-authoritative) canvas, again through the prefetch path
    p.spray(canvas)
    map_close(canvas.pixels, 2 * plain, 1e-12)
    # a spray that is never read back must not leave a dangling copy
    canvas.clean()
    p.spray(canvas)
    canvas.clean()
    assert (canvas.pixels == 0).all()


@pytest.mark.parametrize("nside", [4, 64, 1024])
def test_disc_edge_cases_on_device(nside):
    """the degenerate / adversarial discs of tests/test_hostcheck.py through the device kernels (K1 count,
    K1b flat list, and the fused painter via a coverage-count spray)"""
    import pandas as pd
    import torch
    from astropaint_b200 import _engine as eng
    _, _, ohp = _oracle()
    hp = ohp.HealpixRing(nside)
    centres = [(0, 0, 1), (0, 0, -1), (1, 0, 0), (-1, 0, 0), (0, 1, 0), (0, -1, 0), (1e-9, 0, 1), (0, -1e-9, -1),
               (3, 4, 0), (1, 1, 1), (-2, 5, -0.001)]
    pix = np.unique(np.linspace(0, hp.npix - 1, 15).astype(np.int64))
    x, y, z = hp.pix2vec(pix)
    centres += list(zip(x, y, z))
    pixsize = np.sqrt(4 * np.pi / hp.npix)
    # radius pi/2 around an axis puts thousands of pixel centres EXACTLY on the disc boundary; how such an
    # exact tie falls is decided by the last bit of libm's atan2 (CUDA vs glibc here, and any two libm
    # builds under the reference) -- DESIGN.md section 5.  It is exercised on the host build only.
    # (np.pi - 1e-12 is likewise host-only: around a pixel centre it mirrors the centre's ring onto a ring
    # centre, ring_above() sits on an exact tie and one ulp of cos() moves a whole ring in or out)
    radii = [0.0, 1e-12, 0.5 * pixsize, pixsize, 2.5 * pixsize, 0.1, 1.0, 1.5, 3.0, 3.2, 4.0]
    rows = [(c, r) for c in centres for r in radii]
    c = np.array([row[0] for row in rows], dtype=np.float64)
    r_ang = np.rad2deg(np.array([row[1] for row in rows])) * 60
    r_eff = np.deg2rad(r_ang / 60)            # what the device forms (ap_disc_radius, R_times = 1)
    theta = np.arctan2(np.hypot(c[:, 0], c[:, 1]), c[:, 2])
    phi = np.arctan2(c[:, 1], c[:, 0]) % (2 * np.pi)
    df = pd.DataFrame({"x": c[:, 0], "y": c[:, 1], "z": c[:, 2], "theta": theta, "phi": phi,
                       "D_a": np.ones(len(rows)), "R_ang_200c": r_ang})
    dev = torch.device("cuda", 0)
    halos = eng.DeviceHalos(df, 1, dev)
    n_pix, _, rmin, rmax = eng.disc_count(halos, nside, want_range=True)
    off, pixd, R, _ = eng.disc_pixels(halos, nside, n_pix, want_R=True)
    off, pixd, R = off.cpu().numpy(), pixd.cpu().numpy(), R.cpu().numpy()
    rmin, rmax = rmin.cpu().numpy(), rmax.cpu().numpy()
    cover = np.zeros(hp.npix)
    for i in range(len(rows)):
        want = hp.query_disc(c[i], r_eff[i])
        got = pixd[off[i]:off[i + 1]]
        if 0.0 < r_eff[i] < 1e-9:
            # cos(radius) == 1.0: whether the centre's own pixel is returned is decided by the last bit of
            # cos(atan2(.)) (ysq = 1 - z^2 - x^2 ~ +-1e-16), i.e. by the libm -- CUDA and glibc may differ,
            # as any two libm builds under healpy may.  Either answer must be {} or {centre pixel}.
            own = hp.vec2pix(*c[i])
            assert len(got) <= 1 and len(want) <= 1 and all(g == own for g in got), (nside, rows[i], got, want)
            np.add.at(cover, got, 1.0)
            continue
        assert np.array_equal(got, want), (nside, rows[i], len(got), len(want))
        np.add.at(cover, want, 1.0)
        if len(want):
            Ri = R[off[i]:off[i + 1]]
            assert abs(rmin[i] - Ri.min()) <= 1e-12 and abs(rmax[i] - Ri.max()) <= 1e-12
    d_map = torch.zeros(hp.npix, dtype=torch.float64, device=dev)
    one = torch.ones(1, dtype=torch.float64, device=dev)
    eng.paint_profile(halos, nside, 0, [one], d_map)          # AP_CONSTANT: coverage count
    assert np.array_equal(d_map.cpu().numpy(), cover)


# ------------------------------------------------------------------ Canvas(inclusive=True)
@pytest.mark.parametrize("name,nside,R_times", [("test6", 64, 1), ("test6", 16, 40), ("box", 128, 5), ("box", 32, 1),
                                                ("websky", 2048, 5), ("websky", 8192, 5), ("websky", 4096, 1)])
def test_inclusive_pixel_sets_bit_exact(cats, name, nside, R_times):
    """Canvas(inclusive=True): hp.query_disc(..., inclusive=True) (paint_bucket.py:782, :1228) per halo --
    device pixel sets == the oracle's restatement of healpix query_disc_inclusive (fact 4), R alongside."""
    import astropaint_b200 as ap
    ref, _, _ = _oracle()
    cat = cats[name]
    canvas = ap.Canvas(cat, nside, R_times=R_times, inclusive=True)
    discs = ref.Discs(cat.data, nside, R_times, inclusive=True)
    excl = ref.Discs(cat.data, nside, R_times)
    n = n_excl = 0
    for halo, (pix, R) in enumerate(zip(canvas.discs.gen_pixel_index(), canvas.discs.gen_cent2pix_mpc())):
        want = discs.pixel_index(halo)
        assert np.array_equal(pix, want), f"halo {halo}: {len(pix)} vs {len(want)} pixels"
        wantR = discs.cent2pix_mpc(halo)
        assert np.abs(R - wantR).max() <= 1e-9 * max(wantR.max(), 1e-300)
        n += len(pix)
        n_excl += len(excl.pixel_index(halo))
    assert n > n_excl > 0


@pytest.mark.parametrize("template", ["gaussian", "kSZ", "BG", "python"])
def test_inclusive_spray_matches_oracle(cats, template):
    """the fused painter (built-in, tabulated + bypass, R_vec) and the flat path under Canvas(inclusive=True)"""
    import astropaint_b200 as ap
    from astropaint_b200.profiles import spherical, Battaglia16, NFW
    ref, oprof, _ = _oracle()
    if template == "gaussian":
        cat, nside, tmpl, otmpl = cats["box"], 64, spherical.gaussian_2D, oprof.gaussian_readme
    elif template == "kSZ":
        cat, nside, tmpl, otmpl = cats["websky"], 2048, Battaglia16.kSZ_T, oprof.b16_kSZ_T
    elif template == "BG":
        cat, nside, tmpl, otmpl = cats["websky"], 2048, NFW.BG, oprof.nfw_BG
    else:
        def tmpl(R, R_200c, v_r):
            return np.cos(R / R_200c) * v_r
        cat, nside, otmpl = cats["box"], 64, tmpl
    canvas = ap.Canvas(cat, nside, R_times=5, inclusive=True)
    ap.Painter(tmpl).spray(canvas)
    want = np.zeros(12 * nside * nside)
    ref.spray(want, cat.data, nside, otmpl, R_times=5, inclusive=True)
    map_close(canvas.pixels, want)
    plain = np.zeros_like(want)
    ref.spray(plain, cat.data, nside, otmpl, R_times=5)
    assert np.count_nonzero(want) > np.count_nonzero(plain)     # the inclusive map really is wider
    # the mode is scoped to the spray: an exclusive canvas afterwards is unaffected
    canvas2 = ap.Canvas(cat, nside, R_times=5)
    ap.Painter(tmpl).spray(canvas2)
    map_close(canvas2.pixels, plain)


# ------------------------------------------------------------------ Catalog.build_dataframe on the device
@pytest.mark.parametrize("with_redshift", [True, False])
def test_device_catalog_build_matches_host(with_redshift):
    """ap_catalog_build (paint_bucket.py:308-364) against the oracle's numpy restatement and against the
    package's own host build: 17 derived columns, 1e-13 relative (velocity components relative to |v|)."""
    import astropaint_b200 as ap
    from astropaint_b200 import synthetic
    ref, _, _ = _oracle()
    raw = synthetic.websky_like(50_000, seed=11)
    if not with_redshift:
        raw = raw.drop(columns=["redshift"])
    dev_cat = ap.Catalog(raw.copy(), default_redshift=0.3, device=True)
    host_cat = ap.Catalog(raw.copy(), default_redshift=0.3, device=False)
    want = ref.build_dataframe(raw.copy(), default_redshift=0.3)
    assert list(dev_cat.data.columns) == list(host_cat.data.columns)
    vnorm = np.sqrt(raw.v_x ** 2 + raw.v_y ** 2 + raw.v_z ** 2).values
    for col in ["D_c", "redshift", "lat", "lon", "theta", "phi", "D_a", "R_200c", "c_200c", "R_ang_200c", "rho_s", "R_s",
                "v_r", "v_th", "v_ph", "v_lat", "v_lon"]:
        g = dev_cat.data[col].values
        for w in (want[col].values, host_cat.data[col].values):
            scale = vnorm if col.startswith("v_") else np.maximum(np.abs(w), 1e-300)
            if col in ("lat", "lon", "theta", "phi"):
                scale = np.maximum(np.abs(w), 1.0)      # angles: absolute 1e-13 rad / deg near zero
            assert (np.abs(g - w) / scale).max() <= 1e-13, col
    # a canvas over the device-built catalog paints the same pixels as over the host-built one
    from astropaint_b200.profiles import NFW
    a, b = ap.Canvas(dev_cat, 1024, R_times=3), ap.Canvas(host_cat, 1024, R_times=3)
    ap.Painter(NFW.kSZ_T).spray(a)
    ap.Painter(NFW.kSZ_T).spray(b)
    map_close(a.pixels, b.pixels)      # columns differ in the last bits (CUDA pow/atan2 vs numpy's): map bar 1e-6


# ------------------------------------------------------------------ apply_func library on the device (K6)
@pytest.mark.parametrize("shape", [(200, 200), (33, 33), (24, 41), (7, 7)])
def test_patch_ops_match_scipy(shape):
    """csrc/patch_ops.cuh against scipy itself (the reference's transform.rotate_patch IS
    scipy.ndimage.rotate(reshape=False), lib/transform.py:266-275): prefilter, rotation with per-cutout
    angles, taper, weights, and the fused accumulate.  Tolerance 1e-12 of the patch scale."""
    import torch
    from scipy import ndimage
    from astropaint_b200 import _engine as eng, _capi
    from astropaint_b200._capi import ptr
    rng = np.random.default_rng(shape[0] + shape[1])
    n = 9
    patches = rng.normal(size=(n,) + shape)
    angles = np.array([0.0, 13.7, 45.0, 90.0, -120.3, 180.0, 271.1, 359.99, 33.3])
    dev = torch.device("cuda", 0)
    d = torch.from_numpy(patches).to(dev)
    coef = d.clone()
    _capi.check(_capi.lib().ap_patch_spline_prefilter(n, shape[0], shape[1], ptr(coef), None))
    want_coef = np.stack([ndimage.spline_filter(p, order=3, mode="constant") for p in patches])
    assert np.abs(coef.cpu().numpy() - want_coef).max() < 1e-12
    got = eng.patch_rotate(d.clone(), angles).cpu().numpy()
    want = np.stack([ndimage.rotate(p, a, reshape=False) for p, a in zip(patches, angles)])
    assert np.abs(got - want).max() < 1e-12
    assert (np.count_nonzero((got == 0) != (want == 0))) <= 2 * n      # border ties of in-bounds tests only
    w = rng.normal(size=n)
    if shape[0] == shape[1]:
        han = np.hanning(shape[0])
        stack = torch.zeros(shape, dtype=torch.float64, device=dev)
        eng.patch_rotate(d.clone(), angles, win=(han, han), weight=torch.from_numpy(w).to(dev), stack=stack)
        want_stack = sum(wi * r * np.outer(han, han) for wi, r in zip(w, want))
        assert np.abs(stack.cpu().numpy() - want_stack).max() < 1e-11
    scaled = eng.patch_scale(d.clone(), weight=torch.from_numpy(w).to(dev)).cpu().numpy()
    assert np.array_equal(scaled, patches * w[:, None, None])


def test_stack_with_library_apply_funcs_stays_on_device(cats):
    """Canvas.cutouts / stack_cutouts with apply_func = [rotate_patch, taper_patch, scale_patch]
    (oriented stacking, Birkinshaw_Gull_stacking.ipynb) == the reference loop calling scipy per halo."""
    import astropaint_b200 as ap
    from astropaint_b200 import transform
    from astropaint_b200.profiles import NFW
    ref, _, _ = _oracle()
    cat = cats["box"]
    nside = 256
    canvas = ap.Canvas(cat, nside, R_times=5)
    ap.Painter(NFW.kSZ_T).spray(canvas)
    pixels = canvas.pixels.copy()
    n = len(cat.data)
    angle = np.rad2deg(np.arctan2(cat.data.v_th.values, cat.data.v_ph.values))
    sign = -np.sign(cat.data.v_r.values)
    halos = np.arange(0, 30, 2)
    funcs = [transform.rotate_patch, transform.taper_patch, transform.scale_patch]
    kw = lambda: [{"angle": angle.copy()}, {}, {"weight": sign.copy()}]  # noqa: E731
    want_cuts = list(ref.cutouts(pixels, cat.data, nside, halos, [-1, 1], None, 48, None, funcs, kw()))
    got_cuts = list(canvas.cutouts(halos, lon_range=[-1, 1], xpix=48, apply_func=funcs, func_kwargs=kw()))
    assert len(got_cuts) == len(want_cuts)
    scale = max(np.abs(w).max() for w in want_cuts)
    for g, w in zip(got_cuts, want_cuts):
        assert np.abs(g - w).max() <= 1e-12 * scale
    got = canvas.stack_cutouts(halos, lon_range=[-1, 1], xpix=48, inplace=False, apply_func=funcs, func_kwargs=kw())
    want = ref.stack_cutouts(pixels, cat.data, nside, halos, [-1, 1], None, 48, None, funcs, kw())
    assert np.abs(got - want).max() <= 1e-11 * scale
    # single callable + shared scalar kwarg
    got = canvas.stack_cutouts(halos, lon_range=[-1, 1], xpix=48, inplace=False, apply_func=transform.rotate_patch,
                               func_kwargs={"angle": 30.0})
    want = ref.stack_cutouts(pixels, cat.data, nside, halos, [-1, 1], None, 48, None, transform.rotate_patch,
                             {"angle": 30.0})
    assert np.abs(got - want).max() <= 1e-11 * scale
    assert canvas._patch_plan([lambda p: p], [{}], 48, 48) is None      # unknown callables keep the host loop


def test_quadrature_fast_path_equals_general_routine_on_device():
    """ap_los_nodes with the rightmost-bisection fast path == the general dqagie kernel, bit for bit, on every
    (halo, node) of a Websky-shaped catalog and for all four LOS integrands."""
    import torch
    import astropaint_b200 as ap
    from astropaint_b200 import synthetic, _engine as eng, _capi
    cat = ap.Catalog(synthetic.websky_like(30_000, seed=4))
    dev = torch.device("cuda", 0)
    halos = eng.DeviceHalos(cat.data, 5, dev)
    n_pix, _, rmin, rmax = eng.disc_count(halos, 8192, want_range=True)
    L = _capi.lib()
    for kind, names, method in [("b16_tau", ("R_200c", "M_200c", "redshift"), "linspace"),
                                ("b16_ne", ("R_200c", "M_200c", "redshift"), "linspace"),
                                ("b16_rho_gas", ("R_200c", "M_200c", "redshift"), "linspace"),
                                ("nfw_rho", ("rho_s", "R_s"), "logspace")]:
        cols = [eng.to_device(cat.data[c].to_numpy(), dev) for c in names]
        nodes, n_nodes = eng.table_nodes(halos, n_pix, rmin, rmax, 15, method)
        try:
            assert L.ap_quad_fast_path(1) in (0, 1)
            fast = eng.los_nodes(kind, halos, cols, nodes, n_nodes).cpu().numpy()
            L.ap_quad_fast_path(0)
            general = eng.los_nodes(kind, halos, cols, nodes, n_nodes).cpu().numpy()
        finally:
            L.ap_quad_fast_path(1)
        assert np.isfinite(general).all() and np.abs(general).max() > 0
        assert np.array_equal(fast, general), kind


# ------------------------------------------------------------------ map sinks
def test_ud_grade_and_coverage_match_oracle(cats):
    import astropaint_b200 as ap
    from astropaint_b200.profiles import spherical
    ref, _, ohp = _oracle()
    cat = cats["box"]
    canvas = ap.Canvas(cat, 32, R_times=3)
    ap.Painter(spherical.gaussian_2D).spray(canvas)
    pixels = canvas.pixels.copy()
    for nside_out, kw in [(8, {}), (16, {"power": -2}), (64, {}), (128, {"power": 1}), (32, {}), (4, {"pess": True})]:
        got = canvas.ud_grade(nside_out, inplace=False, **kw)
        want = ohp.ud_grade(pixels, nside_out, **kw)
        map_close(got, want, 1e-12)
    holed = pixels.copy()
    holed[::7] = ohp.UNSEEN
    canvas.pixels = holed
    for pess in (False, True):
        got = canvas.ud_grade(8, inplace=False, pess=pess)
        want = ohp.ud_grade(holed, 8, pess=pess)
        assert np.array_equal(got == ohp.UNSEEN, want == ohp.UNSEEN)
        ok = want != ohp.UNSEEN
        assert np.abs(got[ok] - want[ok]).max() <= 1e-12 * np.abs(want[ok]).max()
    canvas.pixels = pixels
    canvas.ud_grade(16)                                   # in place: the canvas itself changes resolution
    assert canvas.nside == 16 and canvas.npix == 12 * 256 and canvas.lmax == 47 and len(canvas.pixels) == 12 * 256
    map_close(canvas.pixels, ohp.ud_grade(pixels, 16), 1e-12)
    # show_discs' coverage map (paint_bucket.py:1433-1458) for both disc-query modes
    for inclusive in (False, True):
        c2 = ap.Canvas(cat, 32, R_times=2, inclusive=inclusive)
        discs = ref.Discs(cat.data, 32, 2, inclusive=inclusive)
        want = np.zeros(c2.npix)
        for h in range(len(cat.data)):
            want[discs.pixel_index(h)] = 1
        assert np.array_equal(c2.discs_coverage(), want)



def test_device_catalog_build_matches_reference_notebook_output():
    """ap_catalog_build against the REAL reference's printed output (examples/Birkinshaw_Gull_stacking.ipynb
    cell 3, tests/golden/notebook_catalog_head.json; tolerance = printed precision, see test_oracle.py)."""
    import astropaint_b200 as ap
    from .test_oracle import notebook_head, check_against_notebook
    g, inp = notebook_head()
    cat = ap.Catalog(inp.copy(), device=True)
    check_against_notebook(cat.data, g)


def test_device_ud_grade_reproduces_healpy_docstring_example(cats):
    """hp.ud_grade(np.arange(48.), 1) from healpy's docstring (tests/golden/healpy_docstring_kat.py): RING
    numbering, RING<->NEST and the averaging of ap_ud_grade against the published known answer."""
    import astropaint_b200 as ap
    from .golden import healpy_docstring_kat as K
    canvas = ap.Canvas(cats["box"], 2)
    canvas.pixels = np.arange(48.)
    assert canvas.ud_grade(1, inplace=False).tolist() == K.UD_GRADE_48_TO_1


@pytest.mark.parametrize("nside", [2, 16, 1024, 8192])
def test_device_ang2pix_pix2ang(cats, nside):
    """ap_ang2pix / ap_pix2ang (hp.ang2pix, hp.pix2ang) against the oracle, healpy's docstring examples, and through
    the public generators gen_center_ipix / gen_center_ang(snap2pixel=True) / find_centers_indx."""
    import astropaint_b200 as ap
    from astropaint_b200 import _engine as eng
    from .golden import healpy_docstring_kat as K
    _, _, ohp = _oracle()
    hp = ohp.HealpixRing(nside)
    rng = np.random.default_rng(nside)
    n = 200_000
    theta = np.arccos(rng.uniform(-1, 1, n))
    theta[:500] = rng.uniform(0, 0.02, 500)
    theta[500:1000] = np.pi - rng.uniform(0, 0.02, 500)
    theta[1000:1002] = [0.0, np.pi]
    phi = rng.uniform(-4 * np.pi, 4 * np.pi, n)
    pix = eng.ang2pix(nside, theta, phi)
    assert np.array_equal(pix, hp.ang2pix(theta, phi))
    th, ph = eng.pix2ang(nside, pix)
    wt, wp = hp.pix2ang(pix)
    assert np.abs(th - wt).max() <= 4.5e-16 and np.abs(ph - wp).max() <= 9e-16
    assert np.array_equal(eng.ang2pix(nside, th, ph), pix)          # round trip through the pixel centre
    with pytest.raises(ValueError):
        eng.ang2pix(nside, np.array([-0.1]), np.array([0.0]))
    with pytest.raises(ValueError):
        eng.pix2ang(nside, np.array([12 * nside * nside]))
    if nside == 16:
        a = K.ANG2PIX_16
        assert eng.ang2pix(16, a["theta"], a["phi"]).tolist() == a["pix"]
        a = K.PIX2ANG_16
        t, f = eng.pix2ang(16, a["pix"])
        assert np.abs(t - a["theta"]).max() < 5e-9 and np.abs(f - a["phi"]).max() < 5e-9
    cat = cats["box"]
    canvas = ap.Canvas(cat, nside)
    d = cat.data
    want = hp.ang2pix(d.theta.values, d.phi.values)
    assert list(canvas.discs.gen_center_ipix()) == want.tolist()
    assert list(canvas.discs.gen_center_ipix([3, 1])) == want[[3, 1]].tolist()
    snapped = np.array(list(canvas.discs.gen_center_ang(snap2pixel=True)))
    wt, wp = hp.pix2ang(want)
    assert np.abs(snapped[:, 0] - wt).max() <= 4.5e-16 and np.abs(snapped[:, 1] - wp).max() <= 9e-16
    canvas.discs.find_centers_indx()
    assert np.array_equal(canvas.discs.center_index, want)
