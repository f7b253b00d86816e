"""TEST INFRASTRUCTURE — CPU oracle, never imported by the product path.

Line-by-line CPU restatement of the reference's hot path.  Pinning status: build_dataframe IS pinned
against the real reference (the `catalog.data.head()` output its author committed in
examples/Birkinshaw_Gull_stacking.ipynb cell 3 -> tests/golden/notebook_catalog_head.json, 12 derived columns
to printed precision); everything through healpy is parity unpinned (healpy/astropy are absent, see
oracle/healpix_np.py):

  Catalog.build_dataframe      astropaint/paint_bucket.py:308-364 (+ lib/transform.py:49-133,153-179)
  Canvas.Disc generators       astropaint/paint_bucket.py:1175-1290
  Painter.spray                astropaint/paint_bucket.py:2296-2495 (serial branch 2378-2389)
  Painter._shake_canister      astropaint/paint_bucket.py:2598-2676
  Canvas.cutouts               astropaint/paint_bucket.py:1601-1743
  Canvas.stack_cutouts         astropaint/paint_bucket.py:1745-1941 (serial branch 1932-1935)

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this module.
"""
import inspect

import numpy as np
import pandas as pd

from . import healpix_np as hpx
from . import profiles_np as prof


# ---------------------------------------------------------------------------------
# lib/transform.py
# ---------------------------------------------------------------------------------
def M_200c_to_R_200c(M_200c, redshift):
    """transform.py:49-56"""
    crit_dens = prof.crit_dens_0 * (1 + redshift) ** 3
    return np.power((3 * M_200c) / (800 * np.pi * crit_dens), 1 / 3)


def M_200c_to_c_200c(M_200c, redshift):
    """transform.py:58-78 (Coe 2010)"""
    A, d, m = 5.71, -0.047, -0.097
    M_0 = 2.E12 / prof.h
    return A * (1 + redshift) ** d * (M_200c / M_0) ** m


def M_200c_to_rho_s(M_200c, redshift, R_200c, c_200c):
    """transform.py:81-94"""
    R_s = R_200c / c_200c
    A_c = np.log(1 + c_200c) - c_200c / (1 + c_200c)
    return M_200c / (16 * np.pi * (R_s ** 3) * A_c)


def rad2arcmin(angle):
    """transform.py:143-145"""
    return np.rad2deg(angle) * 60


def arcmin2rad(angle):
    """transform.py:148-150"""
    return np.deg2rad(angle / 60)


def get_cart2sph_jacobian(th, ph):
    """transform.py:153-179"""
    row1 = np.stack((np.sin(th) * np.cos(ph), np.cos(th) * np.cos(ph), -np.sin(ph)))
    row2 = np.stack((np.sin(th) * np.sin(ph), np.cos(th) * np.sin(ph), np.cos(ph)))
    row3 = np.stack((np.cos(th), -np.sin(th), 0.0 * np.cos(th)))
    return np.stack((row1, row2, row3))


def build_dataframe(data, default_redshift=0):
    """paint_bucket.py:308-364 (calculate_redshifts=False branch)."""
    df = pd.DataFrame(data).reset_index(drop=True).copy()
    x, y, z = (df[k].values.astype(np.float64) for k in ("x", "y", "z"))
    # astropy cartesian_to_spherical: r, lat in [-pi/2, pi/2], lon wrapped to [0, 2pi)
    s = np.hypot(x, y)
    df["D_c"] = np.hypot(s, z)
    lat = np.arctan2(z, s)
    lon = np.arctan2(y, x)
    lon = np.where(lon < 0, lon + 2 * np.pi, lon)
    df["lat"], df["lon"] = lat, lon
    if "redshift" not in df.columns:
        df["redshift"] = pd.Series([default_redshift] * len(df))
    df["theta"] = np.pi / 2 - df["lat"]
    df["phi"] = df["lon"]
    df["lon"], df["lat"] = np.rad2deg((df["lon"], df["lat"]))
    df["D_a"] = df["D_c"] / (1 + df["redshift"])
    df["R_200c"] = M_200c_to_R_200c(df["M_200c"], df["redshift"])
    df["c_200c"] = M_200c_to_c_200c(df["M_200c"], df["redshift"])
    df["R_ang_200c"] = rad2arcmin(np.true_divide(df["R_200c"], df["D_a"]))
    df["rho_s"] = M_200c_to_rho_s(df["M_200c"], df["redshift"], df["R_200c"], df["c_200c"])
    df["R_s"] = np.true_divide(df["R_200c"], df["c_200c"])
    J = get_cart2sph_jacobian(df["theta"].values, df["phi"].values)
    v_cart = np.array([df["v_x"], df["v_y"], df["v_z"]])
    df["v_r"], df["v_th"], df["v_ph"] = np.einsum("ij...,i...->j...", J, v_cart)
    df["v_lat"] = -df["v_th"]
    df["v_lon"] = df["v_ph"]
    return df


# ---------------------------------------------------------------------------------
# Canvas.Disc generators (per halo)
# ---------------------------------------------------------------------------------
class Discs:
    def __init__(self, df, nside, R_times=1, inclusive=False, fast_inclusive=True):
        # fast_inclusive: the validated pre-filter of HealpixRing.query_disc_inclusive_ranges (test speed only)
        self.fast_inclusive = fast_inclusive
        self.df = df
        self.nside = nside
        self.R_times = R_times
        self.inclusive = inclusive
        self.hp = hpx.HealpixRing(nside)
        self._x = df["x"].values
        self._y = df["y"].values
        self._z = df["z"].values
        self._rang = df["R_ang_200c"].values
        self._theta = df["theta"].values
        self._phi = df["phi"].values
        self._D_a = df["D_a"].values

    def pixel_index(self, halo):
        """gen_pixel_index, paint_bucket.py:1215-1229"""
        return self.hp.query_disc((self._x[halo], self._y[halo], self._z[halo]),
                                  self.R_times * arcmin2rad(self._rang[halo]),
                                  inclusive=self.inclusive, fast=self.fast_inclusive)

    def cent2pix_rad(self, halo, pixel_index=None):
        """gen_cent2pix_rad, :1256-1264"""
        if pixel_index is None:
            pixel_index = self.pixel_index(halo)
        pixel_ang = self.hp.pix2ang(pixel_index)
        return hpx.angdist(np.array(pixel_ang), (self._theta[halo], self._phi[halo]))

    def cent2pix_mpc(self, halo, pixel_index=None):
        """gen_cent2pix_mpc, :1266-1273"""
        return self._D_a[halo] * self.cent2pix_rad(halo, pixel_index)

    def cent2pix_mpc_vec(self, halo, pixel_index=None):
        """gen_cent2pix_mpc_vec, :1283-1290"""
        if pixel_index is None:
            pixel_index = self.pixel_index(halo)
        pix_vec = np.asarray(self.hp.pix2vec(pixel_index)).T
        cent_vec = hpx.ang2vec(self._theta[halo], self._phi[halo])
        return self._D_a[halo] * (pix_vec - cent_vec)


# ---------------------------------------------------------------------------------
# Painter
# ---------------------------------------------------------------------------------
def analyze_template(template):
    """paint_bucket.py:2542-2596"""
    spec = inspect.getfullargspec(template)
    args = [a for a in spec.args if a not in ("self", "cls")]
    assert sum(a in args for a in ("R", "R_vec")) == 1
    R_idx = args.index("R") if "R" in args else args.index("R_vec")
    return args, list(spec.kwonlyargs), R_idx


def shake_canister(df, args, template_kwargs):
    """paint_bucket.py:2598-2676: columns named like args[1:] + user kwargs (scalars tiled)."""
    cols = [a for a in args[1:] if a in df.columns]
    spray = {c: df[c].values for c in cols}
    n = len(df)
    for k, v in template_kwargs.items():
        if not hasattr(v, "__len__"):
            spray[k] = np.full(n, v)
        else:
            v = np.asarray(v)
            spray[k] = np.tile(v, n) if len(v) == 1 else v
    return spray


def spray(pixels, df, nside, template, R_times=1, inclusive=False, halo_list=None,
          count=None, **template_kwargs):
    """Painter.spray serial branch (paint_bucket.py:2378-2389): in-place np.add.at on `pixels`.

    Returns the number of (halo, pixel) pairs painted.  `count` (optional int64[n]) receives
    the per-halo pixel counts.
    """
    args, _, R_idx = analyze_template(template)
    R_mode = args[R_idx]
    spray_cols = shake_canister(df, args, template_kwargs)
    discs = Discs(df, nside, R_times, inclusive)
    n_pairs = 0
    halos = range(len(df)) if halo_list is None else halo_list
    for halo in halos:
        pixel_index = discs.pixel_index(halo)
        if R_mode == "R":
            R = discs.cent2pix_mpc(halo, pixel_index)
        else:
            R = discs.cent2pix_mpc_vec(halo, pixel_index)
        spray_dict = {R_mode: R, **{k: v[halo] for k, v in spray_cols.items()}}
        val = template(**spray_dict)
        if np.iscomplexobj(val):
            val = np.real(val)
        np.add.at(pixels, pixel_index, val)
        n_pairs += len(pixel_index)
        if count is not None:
            count[halo] = len(pixel_index)
    return n_pairs


# ---------------------------------------------------------------------------------
# Canvas.cutouts / stack_cutouts
# ---------------------------------------------------------------------------------
def cutouts(pixels, df, nside, halo_list=None, lon_range=(-1, 1), lat_range=None,
            xpix=200, ypix=None, apply_func=None, func_kwargs=None):
    """paint_bucket.py:1601-1743"""
    if lat_range is None:
        lat_range = lon_range
    if halo_list is None:
        halo_list = range(len(df))
    if apply_func is not None and not isinstance(apply_func, list):
        apply_func = [apply_func]
        func_kwargs = [func_kwargs or {}]
    if func_kwargs:
        for f_kw in func_kwargs:
            for key, value in list(f_kw.items()):
                if not hasattr(value, "__len__"):
                    f_kw[key] = [value]
    hp = hpx.HealpixRing(nside)
    lon = df["lon"].values
    lat = df["lat"].values
    lonra = list(lon_range)
    latra = list(lat_range)
    for halo in halo_list:
        cut = hpx.cartesian_projmap(hp, pixels, lon[halo], lat[halo], lonra, latra, xpix, ypix)
        if apply_func is not None:
            for func, f_kw in zip(apply_func, func_kwargs):
                fd = {}
                for key, value in f_kw.items():
                    fd[key] = value[halo] if len(value) > 1 else value[0]
                cut = func(cut, **fd)
        yield cut


def stack_cutouts(pixels, df, nside, halo_list=None, lon_range=(-1, 1), lat_range=None,
                  xpix=200, ypix=None, apply_func=None, func_kwargs=None):
    """paint_bucket.py:1745-1941 serial branch"""
    if ypix is None:
        ypix = xpix
    stack = np.zeros((xpix, ypix))
    for cut in cutouts(pixels, df, nside, halo_list, lon_range, lat_range, xpix, ypix,
                       apply_func, func_kwargs):
        stack += cut
    return stack
