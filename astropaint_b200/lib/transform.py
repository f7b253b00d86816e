"""Halo-property and coordinate transforms used by the painting path.

Mirrors the names of the reference's ``astropaint/lib/transform.py`` that the hot path and its
callers use (M_200c_to_* :49-105, D_c_to_D_a :110, radius_to_angsize :121, rad2arcmin /
arcmin2rad :143-150, Jacobians :153-195, velocity conversions :198-239, fwhm2sigma :242,
rotate_patch / taper_patch :266-283).  Host-side numpy, evaluated once per catalog.
"""
import numpy as np

from .cosmology import cosmo, sigma_T, m_p

T_0 = 2.725E6   # uK
c = 299792.     # km/s
c2 = c ** 2
Gcm2 = 4.785E-20  # G / c^2 [Mpc / M_sun]
h = cosmo.h
crit_dens_0 = cosmo.critical_density0
f_b = cosmo.Ob0 / cosmo.Om0


def M_200c_to_R_200c(M_200c, redshift):
    rho = crit_dens_0 * (1 + redshift) ** 3
    return np.power((3 * M_200c) / (800 * np.pi * rho), 1 / 3)


def M_200c_to_c_200c(M_200c, redshift):
    # Coe 2010 (1005.0411) fit, as the reference uses
    A, d, m = 5.71, -0.047, -0.097
    M_0 = 2.E12 / h
    return A * (1 + redshift) ** d * (M_200c / M_0) ** m


def M_200c_to_rho_s(M_200c, redshift, R_200c=None, c_200c=None):
    if R_200c is None:
        R_200c = M_200c_to_R_200c(M_200c, redshift)
    if c_200c is None:
        c_200c = M_200c_to_c_200c(M_200c, redshift)
    R_s = R_200c / c_200c
    A_c = np.log(1 + c_200c) - c_200c / (1 + c_200c)
    return M_200c / (16 * np.pi * (R_s ** 3) * A_c)


def M_to_tau(M):
    X_H = 0.76
    x_e = (X_H + 1) / 2 * X_H
    f_s = 0.02
    mu = 4 / (2 * X_H + 1 + X_H * x_e)
    return sigma_T * x_e * X_H * (1 - f_s) * f_b * M / mu / m_p


def D_c_to_D_a(D_c, redshift):
    return D_c / (1 + redshift)


def D_c_to_redshift(D_c, units=None):
    return cosmo.z_at_comoving_distance(D_c)


def rad2arcmin(angle):
    return np.rad2deg(angle) * 60


def arcmin2rad(angle):
    return np.deg2rad(angle / 60)


def radius_to_angsize(radius, D_a, arcmin=True):
    ang_size = np.true_divide(radius, D_a)
    if arcmin:
        ang_size = rad2arcmin(ang_size)
    return ang_size


def get_cart2sph_jacobian(th, ph):
    s_t, c_t, s_p, c_p = np.sin(th), np.cos(th), np.sin(ph), np.cos(ph)
    return np.stack((np.stack((s_t * c_p, c_t * c_p, -s_p)),
                     np.stack((s_t * s_p, c_t * s_p, c_p)),
                     np.stack((c_t, -s_t, 0.0 * c_t))))


def get_sph2cart_jacobian(th, ph):
    s_t, c_t, s_p, c_p = np.sin(th), np.cos(th), np.sin(ph), np.cos(ph)
    return np.stack((np.stack((s_t * c_p, s_t * s_p, c_t)),
                     np.stack((c_t * c_p, c_t * s_p, -s_t)),
                     np.stack((-s_p, c_p, 0.0 * c_t))))


def convert_velocity_sph2cart(th, ph, v_r, v_th, v_ph):
    J = get_sph2cart_jacobian(th, ph)
    v = np.einsum('ij...,i...->j...', J, np.array([v_r, v_th, v_ph]))
    return v[0], v[1], v[2]


def convert_velocity_cart2sph(th, ph, v_x, v_y, v_z):
    J = get_cart2sph_jacobian(th, ph)
    v_r, v_th, v_ph = np.einsum('ij...,i...->j...', J, np.array([v_x, v_y, v_z]))
    return v_r, v_th, v_ph


def fwhm2sigma(fwhm, arcmin=True):
    if arcmin:
        fwhm = arcmin2rad(fwhm)
    return fwhm ** 2 / 8 / np.log(2)


def rotate_patch(patch, angle):
    """Rotate a cutout by `angle` degrees (scipy.ndimage.rotate, cubic, same shape)."""
    from scipy.ndimage import rotate
    return rotate(patch, angle, reshape=False)


def taper_patch(patch):
    """Multiply a cutout by a 2-D Hanning window (in place, like the reference)."""
    han = np.hanning(patch.shape[0])
    patch *= np.outer(han, han)
    return patch


def scale_patch(patch, weight=1.0):
    """Multiply a cutout by a (per-halo) number -- the sign flip / weighting of oriented stacking
    (parallel_stacking.ipynb).  Not in the reference's transform.py; provided so that the common
    `lambda patch, w: patch * w` has a device twin."""
    return patch * weight


# Canvas.cutouts / stack_cutouts run chains made ONLY of these on the device (csrc/patch_ops.cuh); any other
# callable in `apply_func` keeps the reference's per-halo host loop.  value = (op, name of its keyword)
rotate_patch._ap_patch_op = ("rotate", "angle")
taper_patch._ap_patch_op = ("taper", None)
scale_patch._ap_patch_op = ("scale", "weight")
