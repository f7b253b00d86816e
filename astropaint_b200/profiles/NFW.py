"""[Projected] NFW halo profiles (reference: astropaint/profiles/NFW.py).

The reference evaluates the projected NFW forms through complex arithmetic and keeps the real
part (NFW.py:75-80, :138-150); the same real-valued branches are written out here and in the
device twins (csrc/profiles.cuh :: nfw_g).
"""
import numpy as np

from ..lib.cosmology import cosmo, sigma_T, m_p
from ..lib.utils import interpolate, LOS_integrate
from ..lib import transform
from .._registry import (device_profile, device_table, AP_NFW_RHO2D, AP_NFW_TAU2D, AP_NFW_KSZ,
                         AP_NFW_DEFLECTION, AP_NFW_BG, AP_NFW_RHO2D_LOS)

f_b = cosmo.Ob0 / cosmo.Om0
c = 299792.  # km/s
h = cosmo.h
T_cmb = 2.7251
Gcm2 = 4.785E-20  # G/c^2 (Mpc/M_sun)


def rho_3D(r, rho_s, r_s):
    """3-D NFW density: 4 rho_s r_s^3 / (r (r + r_s)^2)."""
    return 4 * rho_s * r_s ** 3 / r / (r + r_s) ** 2


def _g(x):
    """Re[2 / sqrt(1 - x^2) * arctanh(sqrt((1 - x) / (1 + x)))] for real x > 0."""
    x = np.asarray(x, dtype=np.float64)
    out = np.empty_like(x)
    lo = x < 1
    with np.errstate(divide="ignore", invalid="ignore"):
        xl = x[lo]
        out[lo] = 2 / np.sqrt((1 - xl) * (1 + xl)) * np.arctanh(np.sqrt((1 - xl) / (1 + xl)))
        xh = x[~lo]
        out[~lo] = 2 / np.sqrt((xh - 1) * (xh + 1)) * np.arctan(np.sqrt((xh - 1) / (1 + xh)))
    return out


@device_profile(AP_NFW_RHO2D, ["rho_s", "R_s"])
def rho_2D_bartlemann(R, rho_s, R_s):
    """Projected NFW surface mass density, Bartelmann 1996 eq. 7 [M_sun / Mpc^2]."""
    x = np.asarray(R / R_s, dtype=np.float64)
    with np.errstate(divide="ignore", invalid="ignore"):
        f = (1 - _g(x)) / ((x - 1) * (x + 1))
    return 8 * rho_s * R_s * f


@device_profile(AP_NFW_RHO2D_LOS, ["rho_s", "R_s"])
@LOS_integrate
def rho_2D(R, rho_s, R_s):
    """3-D NFW profile integrated along the line of sight."""
    return rho_3D(R, rho_s, R_s)


@device_table("nfw_rho", ["rho_s", "R_s"], n_samples=20, sampling_method="logspace")
@interpolate(n_samples=20, sampling_method="logspace")
@LOS_integrate
def rho_2D_interp(R, rho_s, R_s):
    """rho_2D on 20 log-spaced lines of sight, splined in between."""
    return rho_3D(R, rho_s, R_s)


@device_profile(AP_NFW_DEFLECTION, ["c_200c", "R_200c", "M_200c"])
def deflection_angle(R, c_200c, R_200c, M_200c, *, suppress=True, suppression_R=8):
    """Lensing deflection angle of an NFW halo (Baxter et al. 2015 eq. 6), suppressed beyond
    suppression_R x R_200c."""
    A = M_200c * c_200c ** 2 / (np.log(1 + c_200c) - c_200c / (1 + c_200c)) / 4. / np.pi
    C = 16 * np.pi * Gcm2 * A / c_200c / R_200c
    x = np.asarray(R / (R_200c / c_200c), dtype=np.float64)
    with np.errstate(divide="ignore", invalid="ignore"):
        alpha = C * (np.log(x / 2) + _g(x)) / x
    if suppress:
        alpha = alpha * np.exp(-(R / (suppression_R * R_200c)) ** 3)
    return alpha


@device_profile(AP_NFW_TAU2D, ["rho_s", "R_s"])
def tau_2D(R, rho_s, R_s):
    """Projected NFW optical depth."""
    X_H = 0.76
    x_e = (X_H + 1) / 2 * X_H
    f_s = 0.02
    mu = 4 / (2 * X_H + 1 + X_H * x_e)
    return sigma_T * x_e * X_H * (1 - f_s) * f_b * rho_2D_bartlemann(R, rho_s, R_s) / mu / m_p


@device_profile(AP_NFW_KSZ, ["rho_s", "R_s", "v_r"], {"T_cmb": T_cmb})
def kSZ_T(R, rho_s, R_s, v_r, *, T_cmb=T_cmb):
    """Kinetic Sunyaev-Zeldovich temperature shift of an NFW gas halo."""
    return -tau_2D(R, rho_s, R_s) * v_r / c * T_cmb


@device_profile(AP_NFW_BG, ["c_200c", "R_200c", "M_200c", "theta", "phi", "v_th", "v_ph"],
                {"T_cmb": T_cmb})
def BG(R_vec, c_200c, R_200c, M_200c, theta, phi, v_th, v_ph, *, T_cmb=T_cmb):
    """Birkinshaw-Gull (moving lens) dipole."""
    R = np.linalg.norm(R_vec, axis=-1)
    R_hat = np.true_divide(R_vec, np.expand_dims(R, axis=-1))
    alpha = deflection_angle(R, c_200c, R_200c, M_200c)
    v_vec = transform.convert_velocity_sph2cart(theta, phi, 0, v_th, v_ph)
    return -alpha * np.dot(R_hat, v_vec) / c * T_cmb
