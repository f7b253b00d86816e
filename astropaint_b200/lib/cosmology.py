"""Flat LambdaCDM with the Planck18_arXiv_v2 parameters, without astropy.

The reference takes its cosmology from ``astropy.cosmology.Planck18_arXiv_v2``
(astropaint/lib/transform.py:24, profiles/Battaglia16.py:16).  astropy is not available in this
image, so the handful of quantities the hot path needs are restated here from the published
parameter values; `csrc/profiles.cuh` holds the same numbers for the device functions.
"""
import math

import numpy as np

G = 6.6743e-11                    # m^3 kg^-1 s^-2   (CODATA 2018)
c_SI = 299792458.0                # m / s
Mpc = 3.085677581491367e22        # m
GM_sun = 1.3271244e20             # m^3 s^-2         (IAU 2015 nominal)
M_sun = GM_sun / G                # kg
sigma_SB = 5.6703744191844314e-08
k_B_eV = 8.617333262145179e-05    # eV / K
sigma_T_SI = 6.6524587321e-29     # m^2
m_p_SI = 1.67262192369e-27        # kg


class FlatLambdaCDM:
    def __init__(self, H0, Om0, Ob0, Tcmb0, Neff, m_nu):
        self.H0, self.Om0, self.Ob0, self.Tcmb0, self.Neff = H0, Om0, Ob0, Tcmb0, Neff
        self.h = H0 / 100.0
        self._H0_SI = H0 * 1.0e3 / Mpc
        self._rhoc0_SI = 3.0 * self._H0_SI ** 2 / (8.0 * math.pi * G)
        #: critical density today [M_sun / Mpc^3]
        self.critical_density0 = self._rhoc0_SI * Mpc ** 3 / M_sun
        self.Ogamma0 = 4.0 * sigma_SB / c_SI ** 3 * Tcmb0 ** 4 / self._rhoc0_SI
        Tnu0 = 0.7137658555036082 * Tcmb0
        self._nu_y = np.array([m / (k_B_eV * Tnu0) for m in m_nu if m > 0])
        self._n_massless = sum(1 for m in m_nu if m == 0)
        self.Onu0 = self.Ogamma0 * self.nu_relative_density(0.0)
        self.Ode0 = 1.0 - Om0 - self.Ogamma0 - self.Onu0

    def nu_relative_density(self, z):
        """Neutrino-to-photon density ratio (Komatsu et al. 2011, eq. 26 fit)."""
        prefac, p, invp, k = 0.22710731766, 1.83, 0.54644808743, 0.3173
        z = np.asarray(z, dtype=np.float64)
        curr = self._nu_y / (1.0 + z[..., None])
        rel_mass = ((1.0 + (k * curr) ** p) ** invp).sum(-1) + self._n_massless
        return prefac * (self.Neff / 3.0) * rel_mass

    def efunc2(self, z):
        z = np.asarray(z, dtype=np.float64)
        zp1 = 1.0 + z
        Or = self.Ogamma0 * (1.0 + self.nu_relative_density(z))
        return zp1 ** 3 * (Or * zp1 + self.Om0) + self.Ode0

    def critical_density(self, z):
        """rho_c(z) in M_sun / Mpc^3."""
        return self.critical_density0 * self.efunc2(z)

    def comoving_distance(self, z):
        """Line-of-sight comoving distance [Mpc]."""
        from scipy import integrate
        dh = c_SI / 1.0e3 / self.H0
        f = lambda zz: 1.0 / math.sqrt(float(self.efunc2(zz)))
        zs = np.atleast_1d(np.asarray(z, dtype=np.float64))
        out = np.array([dh * integrate.quad(f, 0.0, zi, epsrel=1e-10)[0] for zi in zs])
        return out if np.ndim(z) else float(out[0])

    def z_at_comoving_distance(self, D_c, zmax=20.0):
        """Inverse of comoving_distance (replaces astropy z_at_value; reference: lib/transform.py:116-119, called per
        halo by Catalog(calculate_redshifts=True), paint_bucket.py:321-323).  A handful of distances are inverted by
        root finding; a catalog goes through a table: D_c on 8192 nodes (one adaptive quadrature per node interval,
        1e-12), inverted with a cubic spline -- agrees with the root finder to ~1e-11 in z and turns the hours the
        reference's per-halo z_at_value takes for 1e7 halos into a second."""
        from scipy import optimize
        D = np.atleast_1d(np.asarray(D_c, dtype=np.float64))
        if D.size > 32:
            out = self._z_of_D_table(zmax)(np.clip(D, 0.0, None))
            return out if np.ndim(D_c) else float(out[0])
        out = np.array([optimize.brentq(lambda zz: self.comoving_distance(zz) - d, 0.0, zmax, xtol=1e-12)
                        if d > 0 else 0.0 for d in D])
        return out if np.ndim(D_c) else float(out[0])

    def _z_of_D_table(self, zmax):
        cache = self.__dict__.setdefault("_zD_cache", {})
        if zmax not in cache:
            from scipy import integrate
            from scipy.interpolate import CubicSpline
            z = np.concatenate([[0.0], np.geomspace(1e-7, zmax, 8191)])
            dh = c_SI / 1.0e3 / self.H0
            f = lambda zz: 1.0 / math.sqrt(float(self.efunc2(zz)))
            steps = [integrate.quad(f, a, b, epsabs=0.0, epsrel=1e-12)[0] for a, b in zip(z[:-1], z[1:])]
            D = dh * np.concatenate([[0.0], np.cumsum(steps)])
            cache[zmax] = CubicSpline(D, z)
        return cache[zmax]


Planck18_arXiv_v2 = FlatLambdaCDM(H0=67.66, Om0=0.30966, Ob0=0.04897, Tcmb0=2.7255, Neff=3.046,
                                  m_nu=(0.0, 0.0, 0.06))
cosmo = Planck18_arXiv_v2

sigma_T = sigma_T_SI / Mpc ** 2    # [Mpc^2]
m_p = m_p_SI / M_sun               # [M_sun]
