"""Synthetic halo catalogs for the parity tests and bench.py (the reference's WebSky / Sehgal /
MICE csv files are git-LFS blobs that are not available; SURVEY.md section 8d fixes the recipes)."""
import numpy as np
import pandas as pd

from .lib.cosmology import cosmo


def random_box(n_tot=1000, seed=0, box_size=50, v_max=100, mass_min=1E14, mass_max=1E15):
    """BASELINE config 1: the reference's Catalog.generate_random_box defaults with np.random.seed(seed)
    (legacy global RNG, same call order as paint_bucket.py:198-217)."""
    np.random.seed(seed)
    x, y, z = np.random.uniform(low=-box_size / 2, high=box_size / 2, size=(3, n_tot))
    v_x, v_y, v_z = np.random.uniform(low=-v_max, high=v_max, size=(3, n_tot))
    M = np.exp(np.random.uniform(low=np.log(mass_min), high=np.log(mass_max), size=n_tot))
    return pd.DataFrame({"x": x, "y": y, "z": z, "v_x": v_x, "v_y": v_y, "v_z": v_z, "M_200c": M})


_Z_TABLE = None


def _redshift_of_distance(D):
    global _Z_TABLE
    if _Z_TABLE is None:
        z = np.concatenate([[0.0], np.geomspace(1e-4, 6.0, 600)])
        _Z_TABLE = (cosmo.comoving_distance(z), z)
    from scipy.interpolate import CubicSpline
    return CubicSpline(_Z_TABLE[0], _Z_TABLE[1])(D)


def websky_like(n, seed=1, d_min=150.0, z_max=4.6, m_min=1e13, m_max=3.5e15, sigma_v=250.0):
    """BASELINE configs 2-5: "Websky-shaped" full-sky light-cone catalog.

    Directions uniform on the sphere; comoving distance pdf ~ D^2 on [d_min, D_c(z_max)]; redshift
    from D_c (flat LCDM, Planck18); mass pdf ~ M^-2.9 on [m_min, m_max]; Gaussian velocities.
    """
    rng = np.random.default_rng(seed)
    d_max = float(cosmo.comoving_distance(z_max))
    mu = rng.uniform(-1.0, 1.0, n)
    ph = rng.uniform(0.0, 2 * np.pi, n)
    D = (d_min ** 3 + rng.uniform(0.0, 1.0, n) * (d_max ** 3 - d_min ** 3)) ** (1.0 / 3.0)
    st = np.sqrt(1.0 - mu * mu)
    a = -1.9
    u = rng.uniform(0.0, 1.0, n)
    M = (m_min ** a + u * (m_max ** a - m_min ** a)) ** (1.0 / a)
    v = rng.normal(0.0, sigma_v, size=(3, n))
    return pd.DataFrame({"x": D * st * np.cos(ph), "y": D * st * np.sin(ph), "z": D * mu,
                         "v_x": v[0], "v_y": v[1], "v_z": v[2], "M_200c": M,
                         "redshift": _redshift_of_distance(D)})
