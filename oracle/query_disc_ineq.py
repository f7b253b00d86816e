"""A SECOND, independent statement of hp.query_disc(inclusive=False) for the RING scheme -- test infrastructure.

oracle/healpix_np.py restates healpix_cxx's own algorithm (Appendix A of SURVEY.md: per ring,
`floor(nr/2pi (phi +- dphi) - shift)` with dphi = atan2(sqrt(ysq), x)).  healpy is not installable here, so that
restatement is pinned by construction only.  This module narrows what is unverified: it defines the SAME set from
the geometric definition the healpix paper gives for the exclusive query,

    { pixels p : angular distance(centre of p, disc centre) < radius }
    <=>  cos(angdist) = z_p z_0 + sqrt(1 - z_p^2) sqrt(1 - z_0^2) cos(phi_p - phi_0)  >  cos(radius),

evaluated on every pixel centre of every ring that can possibly reach the disc, with the ring / in-ring pixel
numbers enumerated as integers (no floor of an angle anywhere).  The two statements share only the pixel-centre
convention (ring2z, phi_j = (j + shift) 2 pi / nr), which healpy's published docstring examples pin
(tests/golden/healpy_docstring_kat.py).  They can disagree only for a pixel whose centre lies within rounding error
of the disc's edge; `compare` reports every such pixel with its distance from the edge.
"""
import math

import numpy as np

TWOPI = 2.0 * math.pi


def _ring_tables(nside):
    """z, first pixel, pixel count and phi offset (in pixels) of rings 1 .. 4 nside - 1"""
    ring = np.arange(1, 4 * nside, dtype=np.int64)
    npix = 12 * nside * nside
    north = ring < nside
    south = ring > 3 * nside
    rs = 4 * nside - ring
    z = np.where(north, 1.0 - ring.astype(np.float64) ** 2 * (4.0 / npix),
                 np.where(south, rs.astype(np.float64) ** 2 * (4.0 / npix) - 1.0,
                          (2 * nside - ring) * (2.0 * nside * (4.0 / npix))))
    nr = np.where(north, 4 * ring, np.where(south, 4 * rs, 4 * nside))
    start = np.where(north, 2 * ring * (ring - 1),
                     np.where(south, npix - 2 * rs * (rs + 1), 2 * nside * (nside - 1) + (ring - nside) * 4 * nside))
    shifted = np.where(north | south, True, ((ring - nside) & 1) == 0)
    return z, start, nr, np.where(shifted, 0.5, 0.0)


class DiscByInequality:
    def __init__(self, nside):
        self.nside = int(nside)
        self.npix = 12 * self.nside * self.nside
        self.z, self.start, self.nr, self.shift = _ring_tables(self.nside)

    def query(self, vec, radius, with_margin=False):
        """sorted int64 pixel ids with cos(angdist) > cos(radius); optionally also (pixel, cosd - cosr) of every
        candidate pixel within 1e-9 of the edge"""
        x, y, zc = (float(v) for v in vec)
        theta0 = math.atan2(math.sqrt(x * x + y * y), zc)
        phi0 = math.atan2(y, x)
        if radius >= math.pi:
            out = np.arange(self.npix, dtype=np.int64)
            return (out, []) if with_margin else out
        cosr = math.cos(radius)
        z0 = math.cos(theta0)
        s0 = math.sqrt((1.0 - z0) * (1.0 + z0))
        # rings whose colatitude is within radius (+ one ring of slack) of the centre's
        pad = 3.0 / self.nside
        zhi = math.cos(max(0.0, theta0 - radius - pad))
        zlo = math.cos(min(math.pi, theta0 + radius + pad))
        rows = np.nonzero((self.z <= zhi) & (self.z >= zlo))[0]
        pix, near = [], []
        for r in rows:
            zp, nr, st, sh = float(self.z[r]), int(self.nr[r]), int(self.start[r]), float(self.shift[r])
            sp = math.sqrt((1.0 - zp) * (1.0 + zp))
            # cos(dphi_max) the inequality needs on this ring; every pixel if <= -1, none if >= 1
            denom = sp * s0
            if denom == 0.0:
                inside_all = zp * z0 > cosr
                if inside_all:
                    pix.append(np.arange(st, st + nr, dtype=np.int64))
                continue
            c = (cosr - zp * z0) / denom
            if c >= 1.0 + 1e-9:
                continue
            if c <= -1.0 - 1e-9:
                j = np.arange(nr, dtype=np.int64)
            else:
                half = math.acos(min(1.0, max(-1.0, c))) + 4.0 * TWOPI / nr      # generous candidate window
                if half >= math.pi:
                    j = np.arange(nr, dtype=np.int64)
                else:
                    jc = phi0 / TWOPI * nr - sh
                    w = half / TWOPI * nr
                    j = np.arange(int(math.floor(jc - w)) - 1, int(math.ceil(jc + w)) + 2, dtype=np.int64)
            phip = (np.mod(j, nr) + sh) * (TWOPI / nr)
            cosd = zp * z0 + denom * np.cos(phip - phi0)
            keep = cosd > cosr
            p = st + np.mod(j[keep], nr)
            pix.append(p)
            if with_margin:
                close = np.abs(cosd - cosr) < 1e-9
                near += [(int(st + (jj % nr)), float(cd - cosr)) for jj, cd in zip(j[close], cosd[close])]
        out = np.unique(np.concatenate(pix)) if pix else np.empty(0, dtype=np.int64)
        return (out, near) if with_margin else out


def compare(hp, ineq, vec, radius):
    """(pixels only in the floor-form set, pixels only in the inequality set, {pixel: cosd - cosr} for both)"""
    a = hp.query_disc(vec, radius)
    b, near = ineq.query(vec, radius, with_margin=True)
    only_a = np.setdiff1d(a, b)
    only_b = np.setdiff1d(b, a)
    margins = dict(near)
    return only_a, only_b, margins
