"""TEST INFRASTRUCTURE — CPU oracle, never imported by the product path.

numpy/scipy restatement of the reference's templates and decorators, each function citing
the reference lines it follows.  astropy is absent: the Planck18_arXiv_v2 cosmology and the CODATA
constants are restated from their published values.  H0 / Om0 / critical density / G enter R_200c, c_200c,
rho_s, R_s and R_ang_200c, which ARE pinned to ~1e-6 against the real reference's committed notebook output
(tests/golden/notebook_catalog_head.json); sigma_T, m_p, T_cmb and the Battaglia16 fit stay unpinned.

  astropaint/lib/utils.py:406-425   sample_array
  astropaint/lib/utils.py:428-454   LOS_integrate
  astropaint/lib/utils.py:457-521   interpolate
  astropaint/lib/transform.py:182-225  get_sph2cart_jacobian / convert_velocity_sph2cart
  astropaint/profiles/spherical.py, NFW.py, Battaglia16.py
"""
import math

import numpy as np
from scipy import integrate
from scipy.interpolate import InterpolatedUnivariateSpline

# ---------------------------------------------------------------------------------
# constants (astropy.constants CODATA2018 / IAU2015; astropy.cosmology Planck18_arXiv_v2)
# ---------------------------------------------------------------------------------
_G = 6.6743e-11            # m^3 kg^-1 s^-2
_C_SI = 299792458.0        # m/s
_MPC = 3.085677581491367e22  # m
_GM_SUN = 1.3271244e20     # m^3 s^-2
_MSUN = _GM_SUN / _G       # kg
_SIGMA_SB = 5.6703744191844314e-08
_KB_EV = 8.617333262145179e-05  # eV/K

H0 = 67.66
Om0 = 0.30966
Ob0 = 0.04897
Tcmb0 = 2.7255
Neff = 3.046
m_nu_eV = (0.0, 0.0, 0.06)
h = H0 / 100.0
f_b = Ob0 / Om0

_H0_SI = H0 * 1.0e3 / _MPC
_RHOC0_SI = 3.0 * _H0_SI ** 2 / (8.0 * math.pi * _G)          # kg / m^3
crit_dens_0 = _RHOC0_SI * _MPC ** 3 / _MSUN                   # M_sun / Mpc^3
_Ogamma0 = 4.0 * _SIGMA_SB / _C_SI ** 3 * Tcmb0 ** 4 / _RHOC0_SI
_Tnu0 = 0.7137658555036082 * Tcmb0
_nu_y = np.array([m / (_KB_EV * _Tnu0) for m in m_nu_eV if m > 0])
_n_massless = sum(1 for m in m_nu_eV if m == 0)


def _nu_relative_density(z):
    """astropy FLRW.nu_relative_density (Komatsu et al. 2011 eq. 26 fit)."""
    prefac = 0.22710731766
    p, invp, k = 1.83, 0.54644808743, 0.3173
    curr = _nu_y / (1.0 + z)
    rel_mass = ((1.0 + (k * curr) ** p) ** invp).sum() + _n_massless
    return prefac * (Neff / 3.0) * rel_mass


_Onu0 = _Ogamma0 * _nu_relative_density(0.0)
_Ode0 = 1.0 - Om0 - _Ogamma0 - _Onu0


def critical_density(z):
    """cosmo.critical_density(z).to(Msun/Mpc^3).value (Battaglia16.py:89)."""
    zp1 = 1.0 + z
    Or = _Ogamma0 * (1.0 + _nu_relative_density(z))
    e2 = zp1 ** 3 * (Or * zp1 + Om0) + _Ode0
    return crit_dens_0 * e2


sigma_T = 6.6524587321e-29 / _MPC ** 2      # Mpc^2   (NFW.py:25, Battaglia16.py:32)
m_p = 1.67262192369e-27 / _MSUN             # M_sun   (NFW.py:26)
c = 299792.                                  # km/s    (NFW.py:28)
T_cmb = 2.7251                               # NFW.py:30
Gcm2 = 4.785E-20                             # NFW.py:31


# ---------------------------------------------------------------------------------
# decorators (plain-function restatements; the product keeps the decorator form)
# ---------------------------------------------------------------------------------
def sample_array(array, n_samples, method="linspace", eps=0.001):
    """utils.py:406-425"""
    array_min = array.min()
    array_max = array.max()
    if method == "linspace":
        return np.linspace(array_min, array_max, n_samples)
    if method == "logspace":
        if array_min < eps:
            array_min = eps
        return np.logspace(np.log10(array_min), np.log10(array_max), n_samples)
    raise ValueError(method)


def los_integrate(profile_3D, R, *args):
    """utils.py:428-454: one QUADPACK qagie per R, epsabs=0, epsrel=1e-2."""
    scalar = not hasattr(R, "__iter__")
    Rs = [R] if scalar else R
    out = []
    for R_i in Rs:
        f = lambda r: profile_3D(r, *args) * 2. * r / np.sqrt(r ** 2 - R_i ** 2)
        out.append(integrate.quad(f, R_i, np.inf, epsabs=0., epsrel=1.e-2)[0])
    if scalar:
        return np.asarray(out[0])
    return np.asarray(out)


def interpolate(profile, R, *args, n_samples=20, sampling_method="linspace", k=3):
    """utils.py:457-521 (min_frac=None branch)."""
    if len(R) < n_samples:
        return profile(R, *args)
    R_samples = sample_array(R, n_samples, method=sampling_method)
    sample_values = np.array([profile(R_s, *args) for R_s in R_samples])
    return InterpolatedUnivariateSpline(R_samples, sample_values, k=k)(R)


# ---------------------------------------------------------------------------------
# README.md:99-101 and profiles/spherical.py
# ---------------------------------------------------------------------------------
def gaussian_readme(R, R_200c, x, y, z):
    return np.exp(-(R / R_200c / 3) ** 2) * (x + y + z)


def solid_sphere_2D(R, M_200c, R_200c):
    """spherical.py:25-42 (NaN outside R_200c, as the reference)."""
    with np.errstate(invalid="ignore"):
        return M_200c / 2 / np.pi * np.sqrt(R_200c ** 2 - R ** 2) / R_200c ** 3


def constant_density_2D(R, constant):
    """spherical.py:44-59"""
    return constant


def linear_density_2D(R, intercept, slope):
    """spherical.py:62-79"""
    return intercept + R * slope


# ---------------------------------------------------------------------------------
# profiles/NFW.py
# ---------------------------------------------------------------------------------
def nfw_rho_2D_bartlemann(R, rho_s, R_s):
    """NFW.py:65-80 (np.complex -> complex128)."""
    x = np.asarray(R / R_s, dtype=np.complex128)
    f = 1 - 2 / np.sqrt(1 - x ** 2) * np.arctanh(np.sqrt((1 - x) / (1 + x)))
    f = f.real
    f = np.true_divide(f, x ** 2 - 1)
    return 8 * rho_s * R_s * f


def nfw_tau_2D(R, rho_s, R_s):
    """NFW.py:153-169"""
    X_H = 0.76
    x_e = (X_H + 1) / 2 * X_H
    f_s = 0.02
    mu = 4 / (2 * X_H + 1 + X_H * x_e)
    Sigma = nfw_rho_2D_bartlemann(R, rho_s, R_s)
    return sigma_T * x_e * X_H * (1 - f_s) * f_b * Sigma / mu / m_p


def nfw_kSZ_T(R, rho_s, R_s, v_r, *, T_cmb=T_cmb):
    """NFW.py:172-178"""
    return -nfw_tau_2D(R, rho_s, R_s) * v_r / c * T_cmb


def nfw_deflection_angle(R, c_200c, R_200c, M_200c, *, suppress=True, suppression_R=8):
    """NFW.py:111-150"""
    A = M_200c * c_200c ** 2 / (np.log(1 + c_200c) - c_200c / (1 + c_200c)) / 4. / np.pi
    C = 16 * np.pi * Gcm2 * A / c_200c / R_200c
    R_s = R_200c / c_200c
    x = (R / R_s).astype(np.complex128)
    f = np.true_divide(1, x) * (np.log(x / 2) + 2 / np.sqrt(1 - x ** 2) *
                                np.arctanh(np.sqrt(np.true_divide(1 - x, 1 + x))))
    alpha = C * f
    if suppress:
        alpha *= np.exp(-(R / (suppression_R * R_200c)) ** 3)
    return alpha.real


def sph2cart_velocity(th, ph, v_r, v_th, v_ph):
    """transform.py:182-225"""
    J = np.array([[np.sin(th) * np.cos(ph), np.sin(th) * np.sin(ph), np.cos(th)],
                  [np.cos(th) * np.cos(ph), np.cos(th) * np.sin(ph), -np.sin(th)],
                  [-np.sin(ph), np.cos(ph), 0.0 * np.cos(th)]])
    v = np.einsum('ij...,i...->j...', J, np.array([v_r, v_th, v_ph]))
    return v[0], v[1], v[2]


def nfw_BG(R_vec, c_200c, R_200c, M_200c, theta, phi, v_th, v_ph, *, T_cmb=T_cmb):
    """NFW.py:181-193"""
    R = np.linalg.norm(R_vec, axis=-1)
    R_hat = np.true_divide(R_vec, np.expand_dims(R, axis=-1))
    alpha = nfw_deflection_angle(R, c_200c, R_200c, M_200c)
    v_vec = sph2cart_velocity(theta, phi, 0, v_th, v_ph)
    return -alpha * np.dot(R_hat, v_vec) / c * T_cmb


# ---------------------------------------------------------------------------------
# profiles/Battaglia16.py
# ---------------------------------------------------------------------------------
_xc = 0.5
_gamma = -0.2


def b16_rho_gas_3D(r, R_200c, M_200c, redshift):
    """Battaglia16.py:49-94"""
    m = (M_200c / h) / 1.e14
    rho0 = 4.e3 * m ** 0.29 * (1. + redshift) ** (-0.66)
    alpha = 0.88 * m ** (-0.03) * (1. + redshift) ** 0.19
    beta = 3.83 * m ** 0.04 * (1. + redshift) ** (-0.025)
    x = r / R_200c
    fit = 1. + (x / _xc) ** alpha
    fit **= -(beta + _gamma) / alpha
    fit *= (x / _xc) ** _gamma
    fit *= rho0
    rho_3d = fit * critical_density(redshift)
    rho_3d *= Ob0 / Om0
    return rho_3d


def b16_ne_3D(r, R_200c, M_200c, redshift):
    """Battaglia16.py:97-129"""
    result = b16_rho_gas_3D(r, R_200c, M_200c, redshift)
    me = 9.10938291e-31
    mH = 1.67262178e-27
    mHe = 4. * mH
    xH = 0.76
    nH_ne = 2. * xH / (xH + 1.)
    nHe_ne = (1. - xH) / (2. * (1. + xH))
    factor = (me + nH_ne * mH + nHe_ne * mHe)
    msun = 1.989e30
    factor /= msun
    factor *= h
    result /= factor
    return result


def b16_tau_3D(r, R_200c, M_200c, redshift):
    """Battaglia16.py:132-155"""
    return sigma_T * b16_ne_3D(r, R_200c, M_200c, redshift)


def b16_tau_2D(R, R_200c, M_200c, redshift):
    """Battaglia16.py:221-239"""
    return los_integrate(b16_tau_3D, R, R_200c, M_200c, redshift)


def b16_tau_2D_interp(R, R_200c, M_200c, redshift):
    """Battaglia16.py:242-262 = interpolate(n_samples=15)(LOS_integrate(tau_3D))"""
    return interpolate(b16_tau_2D, R, R_200c, M_200c, redshift, n_samples=15)


def b16_kSZ_T(R, R_200c, M_200c, v_r, redshift, *, T_cmb=T_cmb):
    """Battaglia16.py:265-271"""
    tau = b16_tau_2D_interp(R, R_200c, M_200c, redshift)
    return -tau * v_r / c * T_cmb * (1 + redshift)


def nfw_rho_3D(r, rho_s, r_s):
    """NFW.py:38-58"""
    return 4 * rho_s * r_s ** 3 / r / (r + r_s) ** 2


def nfw_rho_2D(R, rho_s, R_s):
    """NFW.py:83-93 = LOS_integrate(rho_3D)"""
    return los_integrate(nfw_rho_3D, R, rho_s, R_s)


def nfw_rho_2D_interp(R, rho_s, R_s):
    """NFW.py:96-108 = interpolate(n_samples=20, sampling_method="logspace")(LOS_integrate(rho_3D))"""
    return interpolate(nfw_rho_2D, R, rho_s, R_s, n_samples=20, sampling_method="logspace")


def b16_rho_gas_2D(R, R_200c, M_200c, redshift):
    """Battaglia16.py:181-198"""
    return los_integrate(b16_rho_gas_3D, R, R_200c, M_200c, redshift)


def b16_ne_2D(R, R_200c, M_200c, redshift):
    """Battaglia16.py:201-218 = interpolate(LOS_integrate(ne_3D)) with the decorator defaults (20, linspace)"""
    return interpolate(lambda RR, *a: los_integrate(b16_ne_3D, RR, *a), R, R_200c, M_200c, redshift, n_samples=20)
