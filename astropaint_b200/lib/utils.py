"""Sampling, projection and interpolation decorators (+ a timer).

`@LOS_integrate` and `@interpolate` keep the reference's behaviour when the decorated template
is called directly (astropaint/lib/utils.py:406-521).  In addition they record what they wrap
(`_ap_los`, `_ap_interp`) so that `Painter.spray` can run the same semantics on the GPU: the
per-halo nodes of `@interpolate` (sample_array over the halo's own R range), the not-a-knot
spline and its evaluation on every pixel happen on device; only the node values of arbitrary
Python profiles are evaluated on the host.
"""
import functools
import time
from contextlib import contextmanager

import numpy as np
from decorator import decorate
from scipy import integrate
from scipy.interpolate import InterpolatedUnivariateSpline

from .transform import arcmin2rad, fwhm2sigma  # noqa: F401  (re-exported like the reference)


def sample_array(array, n_samples, method="linspace", eps=0.001):
    """n_samples points between array.min() and array.max() (utils.py:406-425)."""
    assert method in ["linspace", "logspace", "random"]
    lo, hi = array.min(), array.max()
    if method == "linspace":
        return np.linspace(lo, hi, n_samples)
    if method == "logspace":
        if lo < eps:
            lo = eps
            print(f"sample array min set to {lo}")
        return np.logspace(np.log10(lo), np.log10(hi), n_samples)
    return np.random.uniform(lo, hi, size=n_samples)


def _los_caller(profile_3D, *args, **kwargs):
    R, rest = args[0], args[1:]
    scalar = not hasattr(R, "__iter__")
    out = []
    for R_i in ([R] if scalar else R):
        f = lambda r: profile_3D(r, *rest, **kwargs) * 2. * r / np.sqrt(r ** 2 - R_i ** 2)
        out.append(integrate.quad(f, R_i, np.inf, epsabs=0., epsrel=1.e-2)[0])
    return np.asarray(out[0] if scalar else out)


def LOS_integrate(profile_3D):
    """Project a 3-D profile along the line of sight: one QUADPACK integral per R (utils.py:428-454)."""
    fun = decorate(profile_3D, _los_caller)
    fun._ap_los = {"inner": profile_3D}
    return fun


def interpolate(profile=None, n_samples=20, min_frac=None, sampling_method="linspace", k=3,
                interpolator=InterpolatedUnivariateSpline):
    """Evaluate `profile` on n_samples points of the input R and spline in between (utils.py:457-521).

    Usable bare (``@interpolate``) or with arguments (``@interpolate(n_samples=15)``).
    """
    if profile is not None and not callable(profile):  # @interpolate(15)
        n_samples, profile = profile, None
    if profile is None:
        return functools.partial(interpolate, n_samples=n_samples, min_frac=min_frac,
                                 sampling_method=sampling_method, k=k, interpolator=interpolator)
    assert n_samples > 1, "number of samples must be larger than 1"

    def caller(f, *args, **kwargs):
        R, rest = args[0], args[1:]
        n = n_samples
        if min_frac:
            assert 0 <= min_frac <= 1, "min_frac must be between 0 to 1"
            n = max(n, min_frac * len(R))
        if len(R) < n:
            return f(R, *rest, **kwargs)
        R_samples = sample_array(R, n, method=sampling_method)
        values = np.array([f(R_s, *rest, **kwargs) for R_s in R_samples])
        return interpolator(R_samples, values, k=k)(R)

    fun = decorate(profile, caller)
    fun._ap_interp = {"inner": profile, "n_samples": n_samples, "min_frac": min_frac,
                      "sampling_method": sampling_method, "k": k,
                      "default_interpolator": interpolator is InterpolatedUnivariateSpline}
    return fun


@contextmanager
def timeit(process_name="Process"):
    """Print the wall time of the enclosed block in minutes (utils.py:528-542)."""
    print("{:=>50}\n{} started at {}\n".format("", process_name, time.strftime("%H:%M:%S %p")))
    t0 = time.time()
    yield
    print("{} was done in {:.1f} min.\n{:=>50}\n".format(process_name, (time.time() - t0) / 60, ""))
