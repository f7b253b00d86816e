from . import spherical, NFW, Battaglia16  # noqa: F401
