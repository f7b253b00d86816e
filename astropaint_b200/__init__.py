"""astropaint_b200 — B200-native implementation of AstroPaint's painting hot path.

    from astropaint_b200 import Catalog, Canvas, Painter
    from astropaint_b200.profiles import NFW, Battaglia16, spherical

mirrors `from astropaint import ...` of the reference (astropaint/__init__.py:3-5).
"""
__version__ = "0.1.0"

from .paint_bucket import Catalog, Canvas, Painter  # noqa: F401
from .lib import transform, utils  # noqa: F401
from .lib.utils import LOS_integrate, interpolate  # noqa: F401
