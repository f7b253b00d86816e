"""Simple projected spherical profiles (reference: astropaint/profiles/spherical.py:25-79) plus
the README examples (README.md:99-101 Gaussian, :170-206 top-hat).  Each has a device twin."""
import numpy as np

from .._registry import (device_profile, AP_CONSTANT, AP_LINEAR, AP_SOLID_SPHERE, AP_GAUSSIAN)


@device_profile(AP_SOLID_SPHERE, ["M_200c", "R_200c"])
def solid_sphere_2D(R, M_200c, R_200c):
    """Projected mass density of a uniform sphere of mass M_200c and radius R_200c."""
    return M_200c / 2 / np.pi * np.sqrt(R_200c ** 2 - R ** 2) / R_200c ** 3


#: analytic LOS projection of the README's 3-D top-hat, in units of the sphere's mass
tophat_2D = solid_sphere_2D


@device_profile(AP_CONSTANT, ["constant"])
def constant_density_2D(R, constant):
    """The same value at every R."""
    return constant


@device_profile(AP_LINEAR, ["intercept", "slope"])
def linear_density_2D(R, intercept, slope):
    """intercept + slope * R."""
    return intercept + R * slope


@device_profile(AP_GAUSSIAN, ["R_200c", "x", "y", "z"])
def gaussian_2D(R, R_200c, x, y, z):
    """The README's demonstration template: a Gaussian of width 3 R_200c scaled by x + y + z."""
    return np.exp(-(R / R_200c / 3) ** 2) * (x + y + z)
