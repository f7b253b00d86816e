"""TEST INFRASTRUCTURE — CPU oracle, never imported by the product path.

numpy restatement of the cutout operators the reference applies through `apply_func`
(astropaint/paint_bucket.py:1729-1742):

    transform.rotate_patch   astropaint/lib/transform.py:266-275   scipy.ndimage.rotate(patch, angle, reshape=False)
    transform.taper_patch    astropaint/lib/transform.py:278-283   patch *= outer(hanning(n), hanning(n))

PINNED: scipy (the reference's own dependency, requirements.txt) is installed, and tests/test_oracle.py checks
this restatement against scipy.ndimage.rotate itself.  It exists so that the device kernels
(csrc/patch_ops.cuh) can be read against a plain statement of the algorithm: cubic B-spline prefilter with
mirror extension (ndimage/src/ni_splines.c), then 4 x 4 tap sampling of in = M out + offset with `cval`
outside [0, n - 1] (ni_interpolation.c, mode='constant').
"""
import math

import numpy as np
from scipy import special

POLE = math.sqrt(3.0) - 2.0


def spline_filter_line(c):
    c = np.array(c, dtype=np.float64)
    n = len(c)
    if n < 2:
        return c
    z = POLE
    c *= (1.0 - z) * (1.0 - 1.0 / z)
    z_i = z
    z_n_1 = z ** (n - 1)
    c0 = c[0] + z_n_1 * c[n - 1]
    for i in range(1, n - 1):
        c0 += z_i * (c[i] + z_n_1 * c[n - 1 - i])
        z_i *= z
    c[0] = c0 / (1.0 - z_n_1 * z_n_1)
    for i in range(1, n):
        c[i] += z * c[i - 1]
    c[n - 1] = (z * c[n - 2] + c[n - 1]) * z / (z * z - 1.0)
    for i in range(n - 2, -1, -1):
        c[i] = z * (c[i + 1] - c[i])
    return c


def spline_prefilter(patch):
    out = np.array(patch, dtype=np.float64)
    for j in range(out.shape[1]):           # axis 0 first, as scipy.ndimage.spline_filter
        out[:, j] = spline_filter_line(out[:, j])
    for i in range(out.shape[0]):
        out[i, :] = spline_filter_line(out[i, :])
    return out


def _weights(x):
    x = x - math.floor(x)
    y, z = x, 1.0 - x
    w1 = (y * y * (y - 2.0) * 3.0 + 4.0) / 6.0
    w2 = (z * z * (z - 2.0) * 3.0 + 4.0) / 6.0
    w0 = z * z * z / 6.0
    return w0, w1, w2, 1.0 - w0 - w1 - w2


def _mirror(i, n):
    if n == 1:
        return 0
    period = 2 * (n - 1)
    i = abs(i) % period
    return period - i if i >= n else i


def rotate_patch(patch, angle):
    """scipy.ndimage.rotate(patch, angle, reshape=False): order 3, mode 'constant', cval 0."""
    patch = np.asarray(patch, dtype=np.float64)
    ny, nx = patch.shape
    c, s = float(special.cosdg(angle)), float(special.sindg(angle))
    cy, cx = (ny - 1) / 2.0, (nx - 1) / 2.0
    off_y = cy - (c * cy + s * cx)
    off_x = cx - (-s * cy + c * cx)
    coef = spline_prefilter(patch)
    out = np.zeros_like(patch)
    for i in range(ny):
        for j in range(nx):
            yy = (c * i + s * j) + off_y
            xx = (-s * i + c * j) + off_x
            if yy < 0 or yy > ny - 1 or xx < 0 or xx > nx - 1:
                continue
            wy, wx = _weights(yy), _weights(xx)
            sy, sx = math.floor(yy) - 1, math.floor(xx) - 1
            v = 0.0
            for a in range(4):
                row = coef[_mirror(sy + a, ny)]
                for b in range(4):
                    v += row[_mirror(sx + b, nx)] * (wy[a] * wx[b])
            out[i, j] = v
    return out


def taper_patch(patch):
    han = np.hanning(patch.shape[0])
    return patch * np.outer(han, han)
