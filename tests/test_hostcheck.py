"""CPU tests of the DEVICE arithmetic: csrc/{healpix,geom,qagi,profiles,spline}.cuh compiled for
the host (libap_hostcheck.so, test-only) against the oracle and scipy.  Integer results (pixel
sets, vec2pix) must be bit-exact; QUADPACK must reproduce scipy's decisions (same neval)."""
import ctypes as C
import os

import numpy as np
import pytest

from oracle import healpix_np as H, profiles_np as oprof, reference_loop as ref

d, i64 = C.c_double, C.c_int64


@pytest.fixture(scope="module")
def hc():
    here = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    L = C.CDLL(os.path.join(here, "astropaint_b200", "_native", "libap_hostcheck.so"))
    L.hc_query_disc.restype = i64
    L.hc_query_disc.argtypes = [i64, d, d, d, d, C.c_void_p, i64]
    L.hc_query_disc_inclusive.restype = i64
    L.hc_query_disc_inclusive.argtypes = [i64, d, d, d, d, C.c_int, C.c_void_p, i64]
    L.hc_max_pixrad.restype = d
    L.hc_max_pixrad.argtypes = [i64]
    L.hc_disc_R.restype = i64
    L.hc_disc_R.argtypes = [i64, d, d, d, d, d, d, d, C.c_void_p, C.c_void_p, i64, C.c_void_p, C.c_void_p]
    L.hc_vec2pix.argtypes = [i64, i64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    L.hc_ang2pix.argtypes = [i64, i64, C.c_void_p, C.c_void_p, C.c_void_p]
    L.hc_pix2ang.argtypes = [i64, i64, C.c_void_p, C.c_void_p, C.c_void_p]
    L.hc_qagi_b16.argtypes = [d, d, d, d, C.POINTER(d), C.POINTER(d), C.POINTER(C.c_int), C.POINTER(C.c_int)]
    L.hc_critical_density.restype = d
    L.hc_critical_density.argtypes = [d]
    L.hc_b16_tau_3D.restype = d
    L.hc_b16_tau_3D.argtypes = [d, d, d, d]
    L.hc_los_fast_path.restype = C.c_int
    L.hc_los_fast_path.argtypes = [C.c_int, d, d, d, d, C.POINTER(d), C.POINTER(d), C.POINTER(C.c_int), C.POINTER(C.c_int)]
    L.hc_los.restype = C.c_int
    L.hc_los.argtypes = [C.c_int, d, d, d, d, C.POINTER(d), C.POINTER(C.c_int)]
    L.hc_profile_R.argtypes = [C.c_int, C.c_void_p, i64, C.c_void_p, C.c_void_p]
    L.hc_profile_vec.argtypes = [C.c_int, C.c_void_p, i64, C.c_void_p, C.c_void_p]
    L.hc_spline.argtypes = [C.c_int, C.c_void_p, C.c_void_p, i64, C.c_void_p, C.c_void_p]
    return L


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


@pytest.mark.parametrize("nside", [1, 2, 3, 8, 12, 64, 100, 512, 8192])
def test_disc_pixel_sets_bit_exact(hc, nside):
    hp = H.HealpixRing(nside)
    rng = np.random.default_rng(100 + nside)
    axes = [np.array(v, dtype=float) for v in [(0, 0, 1), (0, 0, -1), (1, 0, 0), (-1, 0, 0), (0, 1, 0), (0, -1, 0)]]
    n_cases = 300 if nside <= 512 else 120
    buf = np.empty(min(hp.npix, 1 << 24), dtype=np.int64)
    for t in range(n_cases):
        v = axes[t] * 100 if t < 6 else rng.normal(size=3) * rng.uniform(0.1, 100)
        if nside <= 512:
            r = rng.choice([rng.uniform(0, 0.2), rng.uniform(0, np.pi * 1.05), rng.uniform(0, 3 / nside)])
        else:
            r = rng.choice([rng.uniform(0, 5e-3), rng.uniform(0, 3 / nside)])
        want = hp.query_disc(v, r)
        n = hc.hc_query_disc(nside, v[0], v[1], v[2], r, _p(buf), len(buf))
        assert n == len(want)
        assert np.array_equal(buf[:n], want)


@pytest.mark.parametrize("nside", [1, 2, 3, 8, 64, 1024, 8192])
def test_inclusive_disc_pixel_sets_bit_exact(hc, nside):
    """query_disc(inclusive=True, fact) of the device header (host build) == the oracle, fact 4 (healpy's
    default, what Canvas(inclusive=True) gets), 1 (no trimming) and 2."""
    hp = H.HealpixRing(nside)
    assert hc.hc_max_pixrad(nside) == hp.max_pixrad()
    rng = np.random.default_rng(300 + nside)
    pixsize = np.sqrt(4 * np.pi / hp.npix)
    n_cases = 40
    buf = np.empty(min(hp.npix, 1 << 22), dtype=np.int64)
    for t in range(n_cases):
        v = rng.normal(size=3) * rng.uniform(0.1, 100)
        if t % 8 == 0:
            v = np.array([0.0, 0.0, 1.0 if t % 16 else -1.0])
        if t % 8 == 1:
            v = np.array([1e-9, 1e-9, 1.0])
        if nside <= 64:
            r = [0.0, pixsize * rng.uniform(0, 4), rng.uniform(0, 3.2), pixsize * 0.01][t % 4]
        else:
            r = pixsize * rng.uniform(0, 20) if t % 4 else pixsize * rng.uniform(30, 80)
        for fact in (4, 1, 2):
            # plain restatement at small nside; above, the oracle's validated pre-filter (tests/test_oracle.py)
            want = hp.query_disc(v, r, inclusive=True, fact=fact, fast=nside > 64)
            n = hc.hc_query_disc_inclusive(nside, v[0], v[1], v[2], r, fact, _p(buf), len(buf))
            assert n == len(want), (t, v, r, fact)
            assert np.array_equal(buf[:n], want)


def test_R_and_R_vec_match_healpy_formulas(hc):
    rng = np.random.default_rng(5)
    for nside in (8, 64, 512):
        hp = H.HealpixRing(nside)
        for _ in range(100):
            v = rng.normal(size=3) * rng.uniform(1, 100)
            r = rng.choice([rng.uniform(0, 0.2), rng.uniform(0, np.pi), rng.uniform(0, 5 / nside)])
            th = np.arctan2(np.hypot(v[0], v[1]), v[2])
            ph = np.arctan2(v[1], v[0]) % (2 * np.pi)
            Da = np.linalg.norm(v)
            q = hp.query_disc(v, r)
            if len(q) == 0:
                continue
            R = np.empty(len(q))
            Rv = np.empty((len(q), 3))
            rmin, rmax = d(), d()
            hc.hc_disc_R(nside, v[0], v[1], v[2], r, th, ph, Da, _p(R), _p(Rv), len(q), C.byref(rmin), C.byref(rmax))
            Rref = Da * H.angdist(np.array(hp.pix2ang(q)), (th, ph))
            Rvref = Da * (np.asarray(hp.pix2vec(q)).T - H.ang2vec(th, ph))
            ok = Rref > 0
            assert np.abs(R[ok] - Rref[ok]).max() <= 1e-10 * Rref.max()
            assert np.abs(Rv - Rvref).max() <= 1e-13 * Da
            assert abs(rmin.value - Rref.min()) <= 1e-10 * Rref.max()
            assert abs(rmax.value - Rref.max()) <= 1e-10 * Rref.max()


@pytest.mark.parametrize("nside", [1, 3, 4, 12, 64, 100, 8192])
def test_vec2pix_bit_exact(hc, nside):
    hp = H.HealpixRing(nside)
    rng = np.random.default_rng(nside)
    v = rng.normal(size=(3, 20000))
    v[:, :100] = np.array([[0, 0, 1.0]]).T + 1e-3 * rng.normal(size=(3, 100))     # polar caps
    v[:, 100:200] = np.array([[0, 0, -1.0]]).T + 1e-3 * rng.normal(size=(3, 100))
    x, y, z = (np.ascontiguousarray(a) for a in v)
    out = np.empty(x.size, dtype=np.int64)
    hc.hc_vec2pix(nside, x.size, _p(x), _p(y), _p(z), _p(out))
    assert np.array_equal(out, hp.vec2pix(x, y, z))


def test_quadpack_restatement_equals_scipy(hc):
    from scipy import integrate
    import warnings
    rng = np.random.default_rng(2)
    n_checked = 0
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        for _ in range(120):
            M = 10 ** rng.uniform(13, 15.5)
            z = rng.uniform(0, 4.6)
            R200 = float(ref.M_200c_to_R_200c(M, z))
            for R in np.linspace(rng.uniform(0, 0.1) * R200 * rng.choice([1, 0.01]), 5 * R200 * rng.uniform(0.2, 1), 8):
                if R <= 0:
                    continue
                f = lambda r: oprof.b16_tau_3D(r, R200, M, z) * 2. * r / np.sqrt(r ** 2 - R ** 2)  # noqa: E731
                v, e, info = integrate.quad(f, R, np.inf, epsabs=0., epsrel=1e-2, full_output=1)[:3]
                res, ae, ne, ier = d(), d(), C.c_int(), C.c_int()
                hc.hc_qagi_b16(R, R200, M, z, C.byref(res), C.byref(ae), C.byref(ne), C.byref(ier))
                assert ne.value == info["neval"]
                assert abs(res.value - v) <= 1e-10 * abs(v)
                n_checked += 1
    assert n_checked > 800


def test_cosmology_and_tau3d(hc):
    for z in (0.0, 0.3, 1.0, 4.6):
        assert abs(hc.hc_critical_density(z) / oprof.critical_density(z) - 1) < 1e-14
        assert abs(hc.hc_b16_tau_3D(0.3, 1.1, 3e14, z) / oprof.b16_tau_3D(0.3, 1.1, 3e14, z) - 1) < 1e-13


def test_spline_is_scipy_not_a_knot(hc):
    from scipy.interpolate import InterpolatedUnivariateSpline
    rng = np.random.default_rng(9)
    for t in range(100):
        n = int(rng.integers(4, 30))
        x = np.sort(rng.uniform(0, 5, n))
        if t % 2:
            x = np.linspace(x[0], x[-1], n)
        if t % 5 == 0:
            x = np.logspace(-3, 0.5, n)
        y = np.exp(-x) * (1 + 0.1 * rng.normal(size=n))
        X = rng.uniform(x[0], x[-1], 200)
        out = np.empty(200)
        assert hc.hc_spline(n, _p(x), _p(y), 200, _p(X), _p(out)) == 0
        want = InterpolatedUnivariateSpline(x, y, k=3)(X)
        assert np.abs(out - want).max() <= 1e-9 * np.abs(y).max()


PROFILES = [
    (0, [2.5], lambda R, p: np.full_like(R, p[0])),
    (1, [0.5, 2.0], lambda R, p: oprof.linear_density_2D(R, *p)),
    (2, [1e15, 2.0], lambda R, p: oprof.solid_sphere_2D(R, *p)),
    (3, [1.3, 4.0, -2.0, 7.0], lambda R, p: oprof.gaussian_readme(R, *p)),
    (4, [1e15, 0.4], lambda R, p: oprof.nfw_rho_2D_bartlemann(R, *p)),
    (5, [1e15, 0.4], lambda R, p: oprof.nfw_tau_2D(R, *p)),
    (6, [1e15, 0.4, -320.0, 2.7251], lambda R, p: oprof.nfw_kSZ_T(R, p[0], p[1], p[2], T_cmb=p[3])),
    (7, [4.5, 1.6, 8e14], lambda R, p: oprof.nfw_deflection_angle(R, *p)),
]


@pytest.mark.parametrize("pid,params,fn", PROFILES)
def test_device_profiles_equal_reference_formulas(hc, pid, params, fn):
    R = np.concatenate([np.linspace(0.01, 1.9, 40), np.linspace(2.1, 9, 20)])
    p = np.array(params, dtype=np.float64)
    out = np.empty_like(R)
    assert hc.hc_profile_R(pid, _p(p), R.size, _p(R), _p(out)) == 0
    with np.errstate(invalid="ignore"):
        want = np.asarray(fn(R, params), dtype=np.float64)
    assert np.array_equal(np.isnan(out), np.isnan(want))
    ok = ~np.isnan(want)
    assert np.allclose(out[ok], want[ok], rtol=1e-12, atol=0)


def test_device_BG_equals_reference_formula(hc):
    rng = np.random.default_rng(4)
    Rv = rng.normal(size=(50, 3)) * 2
    p = np.array([4.5, 1.6, 8e14, 0.7, 2.1, 150.0, -90.0, 2.7251])
    out = np.empty(50)
    assert hc.hc_profile_vec(8, _p(p), 50, _p(Rv), _p(out)) == 0
    want = oprof.nfw_BG(Rv, *p[:7], T_cmb=p[7])
    assert np.allclose(out, want, rtol=1e-12)


@pytest.mark.parametrize("nside", [64, 8192])
def test_vec2pix_near_azimuth_equals_vec2pix(hc, nside):
    """the cutout kernel's small-angle azimuth path (phi = phi_c + atan series) against the oracle"""
    hc.hc_vec2pix_near.argtypes = [i64, i64, C.c_void_p, C.c_void_p, C.c_void_p, d, C.c_void_p]
    hp = H.HealpixRing(nside)
    rng = np.random.default_rng(nside + 1)
    bad = 0
    for t in range(60):
        lon = rng.uniform(0, 2 * np.pi)
        lat = np.arcsin(rng.uniform(-1, 1)) if t % 6 else rng.choice([1.55, -1.56, 1.2, -0.73])
        n = 4000
        dl = rng.uniform(-0.01, 0.01, n) / max(np.cos(lat), 1e-3)
        db = rng.uniform(-0.01, 0.01, n)
        if t % 10 == 0:
            dl = rng.uniform(-np.pi, np.pi, n)          # far in azimuth: general atan2 branch
        th, ph = np.pi / 2 - np.clip(lat + db, -np.pi / 2, np.pi / 2), lon + dl
        x, y, z = np.sin(th) * np.cos(ph), np.sin(th) * np.sin(ph), np.cos(th)
        x, y, z = (np.ascontiguousarray(a) for a in (x, y, z))
        out = np.empty(n, dtype=np.int64)
        hc.hc_vec2pix_near(nside, n, _p(x), _p(y), _p(z), lon, _p(out))
        bad += int((out != hp.vec2pix(x, y, z)).sum())
    assert bad == 0


@pytest.mark.parametrize("kind", [0, 1, 2, 3])
def test_los_kinds_equal_scipy(hc, kind):
    """every 3-D profile the device projects: value and neval against scipy.integrate.quad"""
    from scipy import integrate
    import warnings
    hc.hc_los.argtypes = [C.c_int, d, d, d, d, C.POINTER(d), C.POINTER(C.c_int)]
    f3 = [oprof.b16_tau_3D, oprof.b16_ne_3D, oprof.b16_rho_gas_3D, oprof.nfw_rho_3D][kind]
    rng = np.random.default_rng(kind)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        for _ in range(40):
            M = 10 ** rng.uniform(13, 15.5)
            z = rng.uniform(0, 4)
            R200 = float(ref.M_200c_to_R_200c(M, z))
            if kind == 3:
                c200 = float(ref.M_200c_to_c_200c(M, z))
                args = (float(ref.M_200c_to_rho_s(M, z, R200, c200)), R200 / c200)
            else:
                args = (R200, M, z)
            for R in R200 * rng.uniform(0.001, 5, 6):
                f = lambda r: f3(r, *args) * 2. * r / np.sqrt(r ** 2 - R ** 2)  # noqa: E731
                v, e, info = integrate.quad(f, R, np.inf, epsabs=0., epsrel=1e-2, full_output=1)[:3]
                res, ne = d(), C.c_int()
                a3 = list(args) + [0.0] * (3 - len(args))
                assert hc.hc_los(kind, R, *a3, C.byref(res), C.byref(ne)) == 0
                assert ne.value == info["neval"]
                assert abs(res.value - v) <= 1e-10 * abs(v)


@pytest.mark.parametrize("nside", [1, 4, 64, 1024])
def test_disc_edge_cases_bit_exact(hc, nside):
    """degenerate and adversarial discs: zero / tiny / >= pi radius, centres on the poles, on phi = 0 and
    pi, on the equator, exactly on pixel centres and on ring boundaries"""
    hp = H.HealpixRing(nside)
    buf = np.empty(hp.npix, dtype=np.int64)
    centres = [(0, 0, 1), (0, 0, -1), (1, 0, 0), (-1, 0, 0), (0, 1, 0), (0, -1, 0), (1, 1e-300, 0), (-1, -1e-300, 0),
               (1e-9, 0, 1), (0, -1e-9, -1), (3, 4, 0), (1, 1, 1), (-2, 5, -0.001)]
    pix = np.unique(np.linspace(0, hp.npix - 1, 25).astype(np.int64))
    x, y, z = hp.pix2vec(pix)                                    # exact pixel centres
    centres += list(zip(x, y, z))
    for r in range(1, 4 * nside, max(1, nside // 3)):            # points exactly on ring latitudes
        zz = hp.ring2z(r)
        s = np.sqrt(max(0.0, 1 - zz * zz))
        centres.append((s * np.cos(0.3), s * np.sin(0.3), zz))
    pixsize = np.sqrt(4 * np.pi / hp.npix)
    radii = [0.0, 1e-300, 1e-12, 0.5 * pixsize, pixsize, 2.5 * pixsize, 0.1, 1.0, np.pi / 2, np.pi - 1e-12, np.pi,
             np.nextafter(np.pi, 4), 4.0, 100.0]
    n_cases = 0
    for c in centres:
        for r in radii:
            want = hp.query_disc(c, r)
            n = hc.hc_query_disc(nside, float(c[0]), float(c[1]), float(c[2]), float(r), _p(buf), len(buf))
            assert n == len(want), (nside, c, r, n, len(want))
            assert np.array_equal(buf[:n], want), (nside, c, r)
            n_cases += 1
    assert n_cases > 300


def test_rightmost_fast_path_is_bit_identical_to_dqagie(hc):
    """qagi_upper_rm (the rightmost-bisection specialisation of dqagie, csrc/qagi.cuh) against scipy.integrate.quad
    -- the reference's LOS_integrate (lib/utils.py:448) -- on Websky-shaped halos x 15 linspace nodes plus
    degenerate R: wherever it accepts a quadrature, neval equals scipy's and the value agrees to 1e-7 (the bar
    of the general restatement); it must accept nearly all of them; and the public entry (fast path + general
    fallback) must give neval == scipy everywhere."""
    from scipy import integrate
    import astropaint_b200 as ap
    from astropaint_b200 import synthetic
    ap.paint_bucket.verbose = False
    D = ap.Catalog(synthetic.websky_like(400, seed=2), device=False).data
    rng = np.random.default_rng(1)
    pix = np.sqrt(4 * np.pi / (12 * 8192 ** 2))
    n_acc = n_tot = 0
    import warnings
    for i in range(120):
        rad = 5 * np.deg2rad(D.R_ang_200c[i] / 60)
        Rs = list(np.linspace(D.D_a[i] * pix * 0.3 * rng.random(), D.D_a[i] * rad, 15))
        if i % 10 == 0:
            Rs += [1e-9, 1e-5 * D.R_200c[i], 50 * D.R_200c[i]]
        for R in Rs[::3] if i % 4 else Rs:
            args = (float(D.R_200c[i]), float(D.M_200c[i]), float(D.redshift[i]))
            r1, e1, n1, i1 = d(), d(), C.c_int(), C.c_int()
            ok = hc.hc_los_fast_path(0, R, *args, C.byref(r1), C.byref(e1), C.byref(n1), C.byref(i1))
            r3, n3 = d(), C.c_int()
            hc.hc_los(0, R, *args, C.byref(r3), C.byref(n3))
            f = lambda r: oprof.b16_tau_3D(r, *args) * 2. * r / np.sqrt(r ** 2 - R ** 2)  # noqa: E731
            with warnings.catch_warnings():
                warnings.simplefilter("ignore")
                v, e, info = integrate.quad(f, R, np.inf, epsabs=0., epsrel=1.e-2, full_output=1)[:3]
            assert n3.value == info["neval"] and abs(r3.value - v) <= 1e-7 * abs(v)
            n_tot += 1
            if ok:
                n_acc += 1
                assert n1.value == info["neval"], (i, R, n1.value, info["neval"])
                assert r1.value == r3.value            # the public entry returned the fast path's bits
                assert abs(e1.value - e) <= 1e-3 * abs(e) + 1e-300
    assert n_acc >= 0.97 * n_tot, (n_acc, n_tot)


def test_call_free_pixel_geometry(hc):
    """geom.cuh: sin^2(d/2) (series / folded Taylor) and 2 asin(sqrt(hav)) (series / fdlibm rational) against
    numpy over their whole domains: a few ulp, far inside the 1e-6 map bar"""
    hc.hc_geom_funcs.argtypes = [i64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    rng = np.random.default_rng(5)
    dd = np.concatenate([rng.uniform(-3 * np.pi, 3 * np.pi, 200000), rng.uniform(-0.3, 0.3, 100000),
                         10.0 ** rng.uniform(-12, 0, 50000), [0.0, np.pi, -np.pi, np.pi / 2, 0.25, -0.25, 2 * np.pi]])
    hav = np.concatenate([rng.uniform(0, 1, 200000), 10.0 ** rng.uniform(-30, 0, 100000),
                          1.0 - 10.0 ** rng.uniform(-16, 0, 50000), [0.0, 1.0, 0.25, 0.5, 1e-4, 0.9999]])
    n = min(len(dd), len(hav))
    dd, hav = np.ascontiguousarray(dd[:n]), np.ascontiguousarray(hav[:n])
    s2, sep = np.empty(n), np.empty(n)
    hc.hc_geom_funcs(n, _p(dd), _p(hav), _p(s2), _p(sep))
    want_s2 = np.sin(0.5 * dd) ** 2
    assert np.abs(s2 - want_s2).max() <= 1e-15                      # absolute (values in [0, 1])
    small = np.abs(dd) < 0.25
    assert (np.abs(s2[small] - want_s2[small]) <= 1e-15 * want_s2[small] + 1e-300).all()
    want_sep = 2.0 * np.arctan2(np.sqrt(hav), np.sqrt(1.0 - hav))   # = 2 asin(sqrt(hav)), accurate at both ends
    ok = hav > 0
    assert (np.abs(sep[ok] - want_sep[ok]) <= 2e-15 * want_sep[ok] + 1e-15).all()   # near pi: pi - small, absolute
    assert sep[hav == 0.0].max() == 0.0



def test_device_header_reproduces_healpy_docstring_examples(hc):
    """healpix.cuh (host build) on healpy's published docstring examples (tests/golden/healpy_docstring_kat.py)."""
    from .golden import healpy_docstring_kat as K
    a = K.VEC2PIX_16
    x, y, z = (np.array(a[k], dtype=np.float64) for k in "xyz")
    out = np.empty(2, dtype=np.int64)
    hc.hc_vec2pix(16, 2, _p(x), _p(y), _p(z), _p(out))
    assert out.tolist() == a["pix"]
    a = K.ANG2PIX_16                                   # ang2pix == vec2pix(dir2vec(theta, phi))
    th, ph = np.array(a["theta"], float), np.array(a["phi"], float)
    x, y, z = np.sin(th) * np.cos(ph), np.sin(th) * np.sin(ph), np.cos(th)
    out = np.empty(5, dtype=np.int64)
    hc.hc_vec2pix(16, 5, _p(x), _p(y), _p(z), _p(out))
    assert out.tolist() == a["pix"]
    assert abs(hc.hc_max_pixrad(16) - K.MAX_PIXRAD_16) < 5e-11
    out = np.empty(5, dtype=np.int64)
    hc.hc_ang2pix(16, 5, _p(th), _p(ph), _p(out))
    assert out.tolist() == a["pix"]
    a = K.PIX2ANG_16
    pix, t, f = np.array(a["pix"], dtype=np.int64), np.empty(5), np.empty(5)
    hc.hc_pix2ang(16, 5, _p(pix), _p(t), _p(f))
    assert np.abs(t - a["theta"]).max() < 5e-9 and np.abs(f - a["phi"]).max() < 5e-9
    assert abs(t[0] - K.PIX2ANG_16_1440[0]) < 1e-15
    a = K.PIX2ANG_PIX11
    for ns, wt, wf in zip(a["nside"], a["theta"], a["phi"]):
        pix = np.array([11], dtype=np.int64)
        hc.hc_pix2ang(ns, 1, _p(pix), _p(t), _p(f))
        assert abs(t[0] - wt) < 5e-9 and abs(f[0] - wf) < 5e-9
    # pix2vec(16, 1504): a disc of one tenth of a pixel around that direction holds exactly pixel 1504
    buf = np.empty(16, dtype=np.int64)
    n = hc.hc_query_disc(16, *K.PIX2VEC_16_1504, 0.005, _p(buf), 16)
    assert buf[:n].tolist() == [1504]


@pytest.mark.parametrize("nside", [1, 2, 3, 16, 100, 1024, 8192])
def test_ang2pix_pix2ang_bit_exact(hc, nside):
    """healpix.cuh ang2pix_ring / pix2ang_ring (hp.ang2pix, hp.pix2ang: paint_bucket.py:996, 1183, 1199, 1238)
    against the oracle: pixel numbers equal, angles equal to 1 ulp (atan2 / acos of the same z), round trip."""
    hp = H.HealpixRing(nside)
    rng = np.random.default_rng(nside)
    n = 20000
    theta = np.arccos(rng.uniform(-1, 1, n))
    theta[:200] = rng.uniform(0, 0.02, 200)                    # inside healpix's sin(theta) branch (< 0.01) and just outside
    theta[200:400] = np.pi - rng.uniform(0, 0.02, 200)
    theta[400:404] = [0.0, np.pi, np.pi / 2, np.arccos(2 / 3)]
    phi = rng.uniform(-4 * np.pi, 4 * np.pi, n)                 # hp.ang2pix wraps any phi
    phi[400:404] = [0.0, 0.0, 2 * np.pi, np.pi / 4]
    pix = np.empty(n, dtype=np.int64)
    hc.hc_ang2pix(nside, n, _p(theta), _p(phi), _p(pix))
    assert np.array_equal(pix, hp.ang2pix(theta, phi))
    if hp.npix <= 40000:
        allp = np.arange(hp.npix, dtype=np.int64)
    else:
        allp = np.unique(np.concatenate([pix, np.arange(4000), hp.npix - 1 - np.arange(4000)]))
    th, ph = np.empty(allp.size), np.empty(allp.size)
    hc.hc_pix2ang(nside, allp.size, _p(allp), _p(th), _p(ph))
    wt, wp = hp.pix2ang(allp)
    assert np.abs(th - wt).max() <= 4.5e-16 and np.array_equal(ph, wp)
    back = np.empty(allp.size, dtype=np.int64)
    hc.hc_ang2pix(nside, allp.size, _p(th), _p(ph), _p(back))
    assert np.array_equal(back, allp)



def test_device_header_reproduces_healpys_query_disc_test_vectors(hc):
    """healpix.cuh (host build of the device code) on the two pixel lists healpy's own test-suite asserts"""
    from .golden import healpy_docstring_kat as K
    q = K.QUERY_DISC_TEST
    r = float(np.radians(q["radius_deg"]))
    out = np.zeros(64, dtype=np.int64)
    n = hc.hc_query_disc(q["nside"], *q["vec"], r, out.ctypes.data, 64)
    assert out[:n].tolist() == q["exclusive"]
    n = hc.hc_query_disc_inclusive(q["nside"], *q["vec"], r, 4, out.ctypes.data, 64)
    assert out[:n].tolist() == q["inclusive"]
