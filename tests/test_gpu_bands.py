"""Ring-band ownership of the map (astropaint_b200/_bands.py, ap_band_clip): every rank of an N-rank layout
paints only its own rings, and the bands of all ranks put side by side ARE the map a single rank paints.
The N ranks are emulated on one GPU (Canvas._band_override), so the driver's single-GPU run covers the
multi-GPU painting logic; the real NCCL run is test_two_rank_spray_equals_single_rank (needs 2 GPUs)."""
import os
import subprocess
import sys

import numpy as np
import pytest

from .conftest import map_close

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _banded_map(cat, nside, R_times, template, ws, inclusive=False, **kw):
    """(map assembled from the ws emulated ranks' bands, total painted pairs, max halos any rank painted)"""
    import astropaint_b200 as ap
    from astropaint_b200 import _bands
    parts, pairs, worst = [], 0, 0
    for r in range(ws):
        canvas = ap.Canvas(cat, nside, R_times=R_times, inclusive=inclusive)
        canvas._band_override = _bands.Band(nside, r, ws)
        p = ap.Painter(template)
        p.spray(canvas, **kw)
        assert p.last_spray["path"].startswith("device_")
        assert canvas._dmap.numel() == canvas._band_override.npix
        parts.append(canvas._dmap.cpu().numpy())
        pairs += int(p.last_spray["d_count"].item())
        worst = max(worst, p.last_spray["n_halos"])
    return np.concatenate(parts), pairs, worst


def _whole_map(cat, nside, R_times, template, inclusive=False, **kw):
    import astropaint_b200 as ap
    canvas = ap.Canvas(cat, nside, R_times=R_times, inclusive=inclusive)
    p = ap.Painter(template)
    p.spray(canvas, **kw)
    return canvas.pixels.copy(), int(p.last_spray["d_count"].item())


CASES = [
    # name, catalog, nside, R_times, template, ranks
    ("nfw_ksz", "websky", 2048, 5, "NFW.kSZ_T", 3),
    ("nfw_bg_vec", "websky", 2048, 5, "NFW.BG", 8),
    ("b16_ksz_table", "websky", 2048, 5, "Battaglia16.kSZ_T", 8),
    ("b16_tau_bypass", "websky", 512, 5, "Battaglia16.tau_2D_interp", 5),     # most discs < 15 pixels: bypass list
    ("gaussian_sky_discs", "box", 128, 5, "spherical.gaussian_2D", 8),        # polar caps, sky-sized discs
    ("nfw_rho_logspace", "websky_big", 1024, 5, "NFW.rho_2D_interp", 2),
    ("nside_not_a_power_of_two", "box", 96, 2, "NFW.kSZ_T", 5),
]


def _template(name):
    from astropaint_b200.profiles import NFW, Battaglia16, spherical
    mod, fn = name.split(".")
    return getattr({"NFW": NFW, "Battaglia16": Battaglia16, "spherical": spherical}[mod], fn)


@pytest.fixture(scope="module")
def band_cats():
    import astropaint_b200 as ap
    from astropaint_b200 import synthetic
    ap.paint_bucket.verbose = False
    return {"websky": ap.Catalog(synthetic.websky_like(6000, seed=11)),
            "websky_big": ap.Catalog(synthetic.websky_like(300, seed=12, m_min=3e14)),
            "box": ap.Catalog(synthetic.random_box(200, seed=4))}


@pytest.mark.parametrize("name,cat,nside,R_times,template,ws", CASES, ids=[c[0] for c in CASES])
def test_emulated_bands_tile_the_single_rank_map(band_cats, name, cat, nside, R_times, template, ws):
    tmpl = _template(template)
    want, n_want = _whole_map(band_cats[cat], nside, R_times, tmpl)
    got, n_got, worst = _banded_map(band_cats[cat], nside, R_times, tmpl, ws)
    assert n_got == n_want              # every (halo, pixel) pair painted by exactly one rank
    assert np.abs(want).max() > 0
    # same values, same pixels; only the order of the fp64 atomic adds differs (sky-sized discs of mixed sign
    # cancel in a pixel, which is where the few 1e-12 of per-pixel relative difference come from)
    map_close(got, want, 1e-10)
    if cat == "websky" and ws >= 3:     # ownership really shards the work (discs are small against a band)
        assert worst < 0.75 * band_cats[cat].size


def test_emulated_bands_inclusive_query(band_cats):
    tmpl = _template("NFW.kSZ_T")
    want, n_want = _whole_map(band_cats["websky"], 1024, 5, tmpl, inclusive=True)
    got, n_got, _ = _banded_map(band_cats["websky"], 1024, 5, tmpl, 4, inclusive=True)
    assert n_got == n_want
    map_close(got, want, 1e-12)


def test_constant_profile_counts_are_exact_per_band(band_cats):
    """constant = 1: each band holds integer disc-coverage counts and the bands' sums add to the pair count."""
    tmpl = _template("spherical.constant_density_2D")
    want, n_want = _whole_map(band_cats["box"], 64, 3, tmpl, constant=1.0)
    got, n_got, _ = _banded_map(band_cats["box"], 64, 3, tmpl, 7, constant=1.0)
    assert n_got == n_want == int(want.sum())
    assert np.array_equal(got, want)


def test_pixels_handed_out_survive_clean_and_respray():
    """`tsz = canvas.pixels; canvas.clean(); spray kSZ; total = tsz + canvas.pixels` -- the array handed
    out before clean() must never be written again (the reference allocates a fresh array in clean(),
    paint_bucket.py:1316); a buffer nobody references any more IS reused (no 6 GB allocation per step)."""
    import astropaint_b200 as ap
    from astropaint_b200 import synthetic
    from astropaint_b200.profiles import NFW
    cat = ap.Catalog(synthetic.websky_like(3000, seed=21))
    canvas = ap.Canvas(cat, 1024, R_times=5)
    ap.Painter(NFW.tau_2D).spray(canvas)
    m1 = canvas.pixels
    keep = m1.copy()
    canvas.clean()
    ap.Painter(NFW.kSZ_T).spray(canvas)
    m2 = canvas.pixels
    assert m2 is not m1 and not np.shares_memory(m1, m2)
    assert np.array_equal(m1, keep)
    assert not np.array_equal(m1, m2)
    # without clean() the reference keeps painting into the same array: aliasing is the contract there
    buf_id = id(canvas._host)
    del m1, m2
    canvas.clean()
    ap.Painter(NFW.kSZ_T).spray(canvas)
    _ = canvas.pixels
    assert id(canvas._host) == buf_id          # nobody held the old array: the pinned buffer was reused


def test_device_twin_is_skipped_for_kwargs_it_does_not_have(band_cats):
    """NFW.deflection_angle(suppress=False / suppression_R=4) are features of the Python template only;
    the spray must not silently paint the default (suppressed, 8 R_200c) device profile."""
    import astropaint_b200 as ap
    from astropaint_b200.profiles import NFW
    from oracle import reference_loop as ref, profiles_np as oprof
    cat = band_cats["websky_big"]
    nside = 512
    for kw in (dict(suppress=False), dict(suppression_R=4)):
        canvas = ap.Canvas(cat, nside, R_times=5)
        p = ap.Painter(NFW.deflection_angle)
        p.spray(canvas, **kw)
        assert p.last_spray["path"] == "flat"
        want = np.zeros(12 * nside * nside)
        ref.spray(want, cat.data, nside, lambda R, c_200c, R_200c, M_200c: oprof.nfw_deflection_angle(
            R, c_200c, R_200c, M_200c, **kw), R_times=5)
        map_close(canvas.pixels, want, 1e-6)
    canvas = ap.Canvas(cat, nside, R_times=5)
    p = ap.Painter(NFW.deflection_angle)
    p.spray(canvas)
    assert p.last_spray["path"] == "device_profile"
    with pytest.raises(TypeError):
        ap.Painter(NFW.deflection_angle).spray(ap.Canvas(cat, nside, R_times=5), no_such_kwarg=1.0)


def test_two_rank_spray_equals_single_rank():
    """Driver-visible multi-GPU correctness: torchrun --nproc-per-node 2 tools/dist_check.py paints NFW.kSZ_T,
    Battaglia16.kSZ_T, NFW.BG, an inclusive canvas and a host (Python) template over two ranks with NCCL and
    asserts band-owned == single-rank <= 1e-12 on the shared host map and on the gathered device map."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs (run with gpurun --gpus 2)")
    import socket
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr",
           "127.0.0.1", "--master-port", str(port), os.path.join(ROOT, "tools", "dist_check.py")]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=900)
    sys.stdout.write(out.stdout[-4000:])
    sys.stderr.write(out.stderr[-4000:])
    assert out.returncode == 0
    assert out.stdout.count("OK") >= 2 * 7


def test_deterministic_accumulation_is_bit_reproducible(band_cats):
    """Canvas(deterministic=True): sort-then-segmented-reduce instead of atomics (SURVEY 7.2 K3 alternative).  Two runs
    give bit-identical maps (the atomic path only agrees to rounding), the map equals the atomic one to 1e-12, the
    emulated ranks' bands tile it bit for bit, and a host-evaluated template refuses the mode."""
    import astropaint_b200 as ap
    from astropaint_b200 import _bands
    from astropaint_b200.profiles import NFW, Battaglia16
    for cat_name, nside, tmpl in [("box", 128, NFW.kSZ_T), ("websky", 2048, Battaglia16.kSZ_T), ("websky", 512, Battaglia16.tau_2D_interp)]:
        cat = band_cats[cat_name]
        maps = []
        for _ in range(2):
            canvas = ap.Canvas(cat, nside, R_times=5, deterministic=True)
            p = ap.Painter(tmpl)
            p.spray(canvas)
            assert p.last_spray["path"].endswith("+deterministic")
            maps.append(canvas.pixels.copy())
        assert np.array_equal(maps[0], maps[1])
        atomic, n_atomic = _whole_map(cat, nside, 5, tmpl)
        assert int(p.last_spray["d_count"].item()) == n_atomic
        map_close(maps[0], atomic, 1e-10)
        parts = []
        for r in range(3):
            canvas = ap.Canvas(cat, nside, R_times=5, deterministic=True)
            canvas._band_override = _bands.Band(nside, r, 3)
            ap.Painter(tmpl).spray(canvas)
            parts.append(canvas._dmap.cpu().numpy())
        assert np.array_equal(np.concatenate(parts), maps[0])      # same pixels, same values, same order of addition
    with pytest.raises(NotImplementedError):
        ap.Painter(lambda R, R_200c: np.exp(-R / R_200c)).spray(ap.Canvas(band_cats["websky_big"], 256, deterministic=True))


def test_map_file_round_trip_through_the_canvas(tmp_path, band_cats):
    """Canvas.save_map_to_file / load_map_from_file (paint_bucket.py:1478-1558): the painted map written in healpy's
    FITS layout (float32, as hp.write_map's default) and read back into a second canvas, which keeps painting on it."""
    import astropaint_b200 as ap
    from astropaint_b200.profiles import NFW
    cat = band_cats["websky"]
    canvas = ap.Canvas(cat, 256, R_times=5)
    ap.Painter(NFW.kSZ_T).spray(canvas)
    painted = canvas.pixels.copy()
    fn = str(tmp_path / "ksz.fits")
    canvas.save_map_to_file(filename=fn)
    default_name = canvas._map_filename(None, "run1", "v2")
    assert default_name == "run1_kSZ_T_NSIDE=256_v2.fits"
    other = ap.Canvas(cat, 256, R_times=5)
    other.load_map_from_file(filename=fn)
    assert np.array_equal(other.pixels, painted.astype(np.float32).astype(np.float64))
    assert other._alm_is_outdated and other._state == "host"
    ap.Painter(NFW.kSZ_T).spray(other)                      # additive on the loaded map
    map_close(other.pixels, painted.astype(np.float32).astype(np.float64) + painted, 1e-7)
    assert np.array_equal(canvas.load_map_from_file(filename=fn, inplace=False), painted.astype(np.float32).astype(np.float64))
