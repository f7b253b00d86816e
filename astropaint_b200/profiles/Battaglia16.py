"""[Projected] gas profiles of Battaglia 2016 (reference: astropaint/profiles/Battaglia16.py).

tau_2D_interp and kSZ_T — the templates of BASELINE configs 2 and 3 — run fully on device:
QUADPACK qagie per node (csrc/qagi.cuh), not-a-knot spline (csrc/spline.cuh), fused painter.
"""
import numpy as np

from ..lib.cosmology import cosmo, sigma_T, m_p  # noqa: F401
from ..lib.utils import interpolate, LOS_integrate
from .._registry import device_profile, device_table, AP_B16_TAU2D, AP_B16_RHOGAS2D

f_b = cosmo.Ob0 / cosmo.Om0
c = 299792.  # km/s
h = cosmo.h
T_cmb = 2.7251
Gcm2 = 4.785E-20

# fit parameters (Battaglia 2016, AGN feedback, 200c)
xc = 0.5
gamma = -0.2


def _m14(M_200c):
    return (M_200c / h) / 1.e14


def _rho0(M_200c, redshift):
    return 4.e3 * _m14(M_200c) ** 0.29 * (1. + redshift) ** (-0.66)


def _alpha(M_200c, redshift):
    return 0.88 * _m14(M_200c) ** (-0.03) * (1. + redshift) ** 0.19


def _beta(M_200c, redshift):
    return 3.83 * _m14(M_200c) ** 0.04 * (1. + redshift) ** (-0.025)


def rho_gas_3D(r, R_200c, M_200c, redshift):
    """3-D gas density: generalised-NFW fit x rho_crit(z) x f_b."""
    y = (r / R_200c) / xc
    a = _alpha(M_200c, redshift)
    fit = (1. + y ** a) ** (-(_beta(M_200c, redshift) + gamma) / a)
    fit = fit * y ** gamma * _rho0(M_200c, redshift)
    return fit * cosmo.critical_density(redshift) * (cosmo.Ob0 / cosmo.Om0)


def ne_3D(r, R_200c, M_200c, redshift):
    """3-D electron number density (fully ionised, primordial He)."""
    me, mH = 9.10938291e-31, 1.67262178e-27
    mHe = 4. * mH
    xH = 0.76
    nH_ne = 2. * xH / (xH + 1.)
    nHe_ne = (1. - xH) / (2. * (1. + xH))
    factor = (me + nH_ne * mH + nHe_ne * mHe) / 1.989e30 * h
    return rho_gas_3D(r, R_200c, M_200c, redshift) / factor


def tau_3D(r, R_200c, M_200c, redshift):
    """3-D Thomson optical depth per unit length."""
    return sigma_T * ne_3D(r, R_200c, M_200c, redshift)


@interpolate(min_frac=0.1, sampling_method="logspace")
@LOS_integrate
def rho_gas_2D_interp(R, R_200c, M_200c, redshift):
    return rho_gas_3D(R, R_200c, M_200c, redshift)


@device_profile(AP_B16_RHOGAS2D, ["R_200c", "M_200c", "redshift"])
@LOS_integrate
def rho_gas_2D(R, R_200c, M_200c, redshift):
    return rho_gas_3D(R, R_200c, M_200c, redshift)


@device_table("b16_ne", ["R_200c", "M_200c", "redshift"], n_samples=20)
@interpolate
@LOS_integrate
def ne_2D(R, R_200c, M_200c, redshift):
    return ne_3D(R, R_200c, M_200c, redshift)


@device_profile(AP_B16_TAU2D, ["R_200c", "M_200c", "redshift"])
@LOS_integrate
def tau_2D(R, R_200c, M_200c, redshift):
    """Projected optical depth: one QUADPACK integral per line of sight."""
    return tau_3D(R, R_200c, M_200c, redshift)


@device_table("b16_tau", ["R_200c", "M_200c", "redshift"], n_samples=15)
@interpolate(n_samples=15)
@LOS_integrate
def tau_2D_interp(R, R_200c, M_200c, redshift):
    """tau_2D on 15 lines of sight spanning the halo's own R range, splined in between."""
    return tau_3D(R, R_200c, M_200c, redshift)


def _ksz_scale(cols, kw):
    return -cols["v_r"] / c * kw["T_cmb"] * (1 + cols["redshift"])


@device_table("b16_tau", ["R_200c", "M_200c", "redshift"], n_samples=15,
              scale={"fn": _ksz_scale, "args": ["v_r", "redshift"], "kwonly": {"T_cmb": T_cmb}})
def kSZ_T(R, R_200c, M_200c, v_r, redshift, *, T_cmb=T_cmb):
    """Kinetic Sunyaev-Zeldovich temperature shift of the Battaglia gas profile."""
    tau = tau_2D_interp(R, R_200c, M_200c, redshift)
    return -tau * v_r / c * T_cmb * (1 + redshift)
