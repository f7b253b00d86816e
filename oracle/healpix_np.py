"""TEST INFRASTRUCTURE — CPU oracle, never imported by the product path.

numpy restatement of the healpy 1.13.0 / healpix_cxx (T_Healpix_Base<int64>, RING scheme)
primitives that the reference's hot path calls.  healpy is a third-party dependency of the
reference (requirements.txt:15, pinned ==1.13.0) and is NOT vendored under /root/reference
and NOT installable here (no network) -> **parity unpinned**: this file restates the
published HEALPix algorithms (Gorski et al. 2005; healpix_cxx `healpix_base.cc`) and is
anchored on the reference's call sites:

    hp.query_disc   astropaint/paint_bucket.py:1222
    hp.pix2ang      astropaint/paint_bucket.py:1199,1238
    hp.pix2vec      astropaint/paint_bucket.py:1254
    hp.ang2vec      astropaint/paint_bucket.py:1212
    hp.ang2pix      astropaint/paint_bucket.py:757,1183
    hp.vec2pix      astropaint/paint_bucket.py:1721
    hp.rotator.angdist            astropaint/paint_bucket.py:1264
    hp.projector.CartesianProj    astropaint/paint_bucket.py:1712 (.projmap :1719)

It is pinned only by construction-true tests (tests/test_oracle_healpix.py): brute-force
disc membership over all pixels, pix2vec/vec2pix round trips, ring-structure identities.
"""
import math

import numpy as np

TWOTHIRD = 2.0 / 3.0
TWOPI = 2.0 * math.pi
INV_TWOPI = 1.0 / TWOPI
HALFPI = 0.5 * math.pi


def nside2npix(nside):
    return 12 * int(nside) * int(nside)


class HealpixRing:
    """Constants of T_Healpix_Base<int64>::SetNside for the RING scheme."""

    def __init__(self, nside):
        nside = int(nside)
        assert nside > 0
        self.nside = nside
        self.npix = 12 * nside * nside
        self.ncap = 2 * nside * (nside - 1)
        self.fact2 = 4.0 / self.npix
        self.fact1 = (nside << 1) * self.fact2

    # -- healpix_base.cc: ring_above ------------------------------------------------
    def ring_above(self, z):
        az = abs(z)
        if az <= TWOTHIRD:
            return int(self.nside * (2.0 - 1.5 * z))
        iring = int(self.nside * math.sqrt(3.0 * (1.0 - az)))
        return iring if z > 0 else 4 * self.nside - iring - 1

    # -- healpix_base.h: ring2z -----------------------------------------------------
    def ring2z(self, ring):
        ns = self.nside
        if ring < ns:
            return 1.0 - ring * ring * self.fact2
        if ring <= 3 * ns:
            return (2 * ns - ring) * self.fact1
        ring = 4 * ns - ring
        return ring * ring * self.fact2 - 1.0

    # -- healpix_base.cc: get_ring_info_small -> (startpix, ringpix, shifted) ------
    def ring_info(self, ring):
        ns = self.nside
        if ring < ns:
            return 2 * ring * (ring - 1), 4 * ring, True
        if ring < 3 * ns:
            return self.ncap + (ring - ns) * 4 * ns, 4 * ns, ((ring - ns) & 1) == 0
        nr = 4 * ns - ring
        return self.npix - 2 * nr * (nr + 1), 4 * nr, True

    # -- healpix_base.cc: query_disc_internal, RING, fact == 0 ----------------------
    def query_disc_ranges(self, vec, radius):
        """Half-open pixel ranges [a, b) in append order (already sorted, merged)."""
        x, y, zc = (float(v) for v in vec)
        # pointing(vec3): the vector need not be normalised
        theta = math.atan2(math.sqrt(x * x + y * y), zc)
        phi = math.atan2(y, x)
        if phi < 0.0:
            phi += TWOPI
        ranges = []

        def append(a, b):
            if b <= a:
                return
            if ranges and a <= ranges[-1][1]:
                if b > ranges[-1][1]:
                    ranges[-1][1] = b
            else:
                ranges.append([a, b])

        if radius >= math.pi:
            return [[0, self.npix]]
        cosr = math.cos(radius)
        z0 = math.cos(theta)
        den = math.sqrt((1.0 - z0) * (1.0 + z0))
        xa = 1.0 / den if den != 0.0 else math.inf  # IEEE division as in the C++ (halo on a pole)
        rlat1 = theta - radius
        zmax = math.cos(rlat1)
        irmin = self.ring_above(zmax) + 1
        if rlat1 <= 0 and irmin > 1:  # north pole in the disc
            sp, rp, _ = self.ring_info(irmin - 1)
            append(0, sp + rp)
        rlat2 = theta + radius
        zmin = math.cos(rlat2)
        irmax = self.ring_above(zmin)
        for iz in range(irmin, irmax + 1):
            z = self.ring2z(iz)
            xx = (cosr - z * z0) * xa
            ysq = 1.0 - z * z - xx * xx
            if ysq != ysq:      # NaN propagates as in C++: neither (ysq <= 0) nor (dphi > 0)
                dphi = math.nan
            else:
                dphi = 0.0 if ysq <= 0 else math.atan2(math.sqrt(ysq), xx)
            if dphi > 0:
                ipix1, nr, shifted = self.ring_info(iz)
                shift = 0.5 if shifted else 0.0
                ipix2 = ipix1 + nr - 1
                ip_lo = int(math.floor(nr * INV_TWOPI * (phi - dphi) - shift)) + 1
                ip_hi = int(math.floor(nr * INV_TWOPI * (phi + dphi) - shift))
                if ip_lo <= ip_hi:
                    if ip_hi >= nr:
                        ip_lo -= nr
                        ip_hi -= nr
                    if ip_lo < 0:
                        append(ipix1, ipix1 + ip_hi + 1)
                        append(ipix1 + ip_lo + nr, ipix2 + 1)
                    else:
                        append(ipix1 + ip_lo, ipix1 + ip_hi + 1)
        if rlat2 >= math.pi and irmax + 1 < 4 * self.nside:  # south pole in the disc
            sp, rp, _ = self.ring_info(irmax + 1)
            append(sp, self.npix)
        return ranges

    # -- healpix_base.cc: max_pixrad -------------------------------------------------
    def max_pixrad(self):
        """Largest angular distance between a pixel centre and its corners: v_angle between
        set_z_phi(2/3, pi/(4 nside)) and set_z_phi(1 - (1 - 1/nside)^2 / 3, 0)."""
        def zphi(z, phi):  # vec3::set_z_phi
            st = math.sqrt((1.0 - z) * (1.0 + z))
            return st * math.cos(phi), st * math.sin(phi), z
        ax, ay, az = zphi(2.0 / 3.0, math.pi / (4 * self.nside))
        t1 = 1.0 - 1.0 / self.nside
        t1 *= t1
        bx, by, bz = zphi(1.0 - t1 / 3.0, 0.0)
        # v_angle = atan2(|a x b|, a . b), plain left-to-right double arithmetic as in vec3.h
        cx, cy, cz = ay * bz - az * by, az * bx - ax * bz, ax * by - ay * bx
        return math.atan2(math.sqrt(cx * cx + cy * cy + cz * cz), ax * bx + ay * by + az * bz)

    # -- healpix_base.cc: ring2xyf / xyf2ring ---------------------------------------
    _JRLL = (2, 2, 2, 2, 3, 3, 3, 3, 4, 4, 4, 4)
    _JPLL = (1, 3, 5, 7, 0, 2, 4, 6, 1, 3, 5, 7)

    @staticmethod
    def _special_div(a, b):
        t = 1 if a >= (b << 1) else 0
        a -= t * (b << 1)
        return (t << 1) + (1 if a >= b else 0)

    def ring2xyf(self, pix):
        ns, npix, ncap = self.nside, self.npix, self.ncap
        nl2 = 2 * ns
        if pix < ncap:
            iring = (1 + int(_isqrt(1 + 2 * pix))) >> 1
            iphi = (pix + 1) - 2 * iring * (iring - 1)
            kshift = 0
            nr = iring
            face = self._special_div(iphi - 1, nr)
        elif pix < npix - ncap:
            ip = pix - ncap
            tmp = ip // (4 * ns)
            iring = tmp + ns
            iphi = ip - tmp * 4 * ns + 1
            kshift = (iring + ns) & 1
            nr = ns
            ire = tmp + 1
            irm = nl2 + 1 - tmp
            ifm = (iphi - (ire >> 1) + ns - 1) // ns
            ifp = (iphi - (irm >> 1) + ns - 1) // ns
            face = (ifp | 4) if ifp == ifm else (ifp if ifp < ifm else ifm + 8)
        else:
            ip = npix - pix
            iring = (1 + int(_isqrt(2 * ip - 1))) >> 1
            iphi = 4 * iring + 1 - (ip - 2 * iring * (iring - 1))
            kshift = 0
            nr = iring
            iring = 2 * nl2 - iring
            face = self._special_div(iphi - 1, nr) + 8
        irt = iring - ((2 + (face >> 2)) * ns) + 1
        ipt = 2 * iphi - self._JPLL[face] * nr - kshift - 1
        if ipt >= nl2:
            ipt -= 8 * ns
        return (ipt - irt) >> 1, (-ipt - irt) >> 1, face

    def xyf2ring(self, ix, iy, face):
        ns = self.nside
        nl4 = 4 * ns
        jr = self._JRLL[face] * ns - ix - iy - 1
        n_before, nr, shifted = self.ring_info(jr)
        nr >>= 2
        kshift = 0 if shifted else 1
        jp = (self._JPLL[face] * nr + ix - iy + 1 + kshift) // 2
        assert jp <= 4 * nr
        if jp < 1:
            jp += nl4
        return n_before + jp - 1

    def pix2zphi(self, pix):
        z, phi, _, _ = self.pix2loc(pix)
        return float(z[0]), float(phi[0])

    def zphi2pix(self, z, phi):
        return int(self.loc2pix(np.array([z]), np.array([phi]), np.array([0.0]), False)[0])

    # -- healpix_base.cc: check_pixel_ring (anonymous namespace) --------------------
    def _check_pixel_ring(self, b2, pix, nr, ipix1, fct, cz, cphi, cosrp2, cpix):
        """True when NO boundary sub-pixel centre of `pix` (at resolution b2 = fct * nside) lies within
        the enlarged radius, i.e. the pixel can be trimmed from the end of the span."""
        if pix >= nr:
            pix -= nr
        if pix < 0:
            pix += nr
        pix += ipix1
        if pix == cpix:
            return False  # disc centre in the pixel => overlap
        px, py, pf = self.ring2xyf(pix)

        def cosdist(ix, iy):
            pz, pphi = b2.pix2zphi(b2.xyf2ring(ix, iy, pf))
            return pz * cz + math.cos(pphi - cphi) * math.sqrt((1.0 - pz * pz) * (1.0 - cz * cz))

        ox, oy = fct * px, fct * py
        for i in range(fct - 1):  # go along the 4 edges
            if cosdist(ox + i, oy) > cosrp2:
                return False
            if cosdist(ox + fct - 1, oy + i) > cosrp2:
                return False
            if cosdist(ox + fct - 1 - i, oy + fct - 1) > cosrp2:
                return False
            if cosdist(ox, oy + fct - 1 - i) > cosrp2:
                return False
        return True

    # -- healpix_base.cc: query_disc_internal, RING, fact > 0 (query_disc_inclusive) -
    def query_disc_inclusive_ranges(self, vec, radius, fact=4, fast=False):
        """healpy.query_disc(..., inclusive=True, fact=4): every pixel OVERLAPPING the disc, possibly a
        few more.  Same ring walk as the exclusive query with (a) the radius enlarged by the pixel radius
        at nside (span boundaries) and at fact*nside (ring range, trimming test), (b) one more ring on
        each side, (c) whole rings when the enlarged disc swallows the ring, and (d) each span trimmed
        from both ends while check_pixel_ring proves the end pixel does not overlap.

        fast=True is a TEST SPEED-UP, not part of the restatement: before the two trimming loops it steps over
        the pixels whose centre is farther than rsmall + 1.02 max_pixrad from the disc centre (for those
        check_pixel_ring is certainly true), which is what keeps a ring of 4 nside pixels lying just outside
        the disc from being walked pixel by pixel in Python.  tests/test_oracle.py proves fast == plain."""
        fct = int(fact)
        assert fct >= 1
        x, y, zc = (float(v) for v in vec)
        theta = math.atan2(math.sqrt(x * x + y * y), zc)
        phi = math.atan2(y, x)
        if phi < 0.0:
            phi += TWOPI
        ranges = []

        def append(a, b):
            if b <= a:
                return
            if ranges and a <= ranges[-1][1]:
                if b > ranges[-1][1]:
                    ranges[-1][1] = b
            else:
                ranges.append([a, b])

        b2 = None
        if fct > 1:
            b2 = HealpixRing(fct * self.nside)
            rsmall = radius + b2.max_pixrad()
            rbig = radius + self.max_pixrad()
        else:
            rsmall = rbig = radius + self.max_pixrad()
        if rsmall >= math.pi:
            return [[0, self.npix]]
        rbig = min(math.pi, rbig)
        cosrsmall = math.cos(rsmall)
        cosrbig = math.cos(rbig)
        z0 = math.cos(theta)
        den = math.sqrt((1.0 - z0) * (1.0 + z0))
        xa = 1.0 / den if den != 0.0 else math.inf
        cpix = self.zphi2pix(z0, phi)
        rlat1 = theta - rsmall
        zmax = math.cos(rlat1)
        irmin = self.ring_above(zmax) + 1
        if rlat1 <= 0 and irmin > 1:  # north pole in the disc
            sp, rp, _ = self.ring_info(irmin - 1)
            append(0, sp + rp)
        if fct > 1 and rlat1 > 0:
            irmin = max(1, irmin - 1)
        rlat2 = theta + rsmall
        zmin = math.cos(rlat2)
        irmax = self.ring_above(zmin)
        if fct > 1 and rlat2 < math.pi:
            irmax = min(4 * self.nside - 1, irmax + 1)
        for iz in range(irmin, irmax + 1):
            z = self.ring2z(iz)
            xx = (cosrbig - z * z0) * xa
            ysq = 1.0 - z * z - xx * xx
            if ysq != ysq:
                dphi = math.nan
            elif ysq <= 0:  # no intersection, ring completely inside or outside
                dphi = 0.0 if fct == 1 else math.pi - 1e-15
            else:
                dphi = math.atan2(math.sqrt(ysq), xx)
            if dphi > 0:
                ipix1, nr, shifted = self.ring_info(iz)
                shift = 0.5 if shifted else 0.0
                ipix2 = ipix1 + nr - 1
                ip_lo = int(math.floor(nr * INV_TWOPI * (phi - dphi) - shift)) + 1
                ip_hi = int(math.floor(nr * INV_TWOPI * (phi + dphi) - shift))
                if fct > 1 and fast:
                    pad = 1.02 * self.max_pixrad()
                    if rsmall + pad < math.pi:
                        sring = math.sqrt((1.0 - z) * (1.0 + z))
                        c = (math.cos(rsmall + pad) - z * z0) * xa / sring if sring > 0 else -2.0
                        if c >= 1.0:
                            ip_lo = ip_hi + 1
                        elif c > -1.0:
                            dsafe = math.acos(c) + 1e-9
                            glo = int(math.floor(nr * INV_TWOPI * (phi - dsafe) - shift)) - 1
                            ghi = int(math.floor(nr * INV_TWOPI * (phi + dsafe) - shift)) + 2
                            if glo > ip_lo:
                                ip_lo = min(glo, ip_hi + 1)
                            if ghi < ip_hi:
                                ip_hi = max(ghi, ip_lo - 1)
                if fct > 1:
                    while ip_lo <= ip_hi and self._check_pixel_ring(b2, ip_lo, nr, ipix1, fct, z0, phi, cosrsmall, cpix):
                        ip_lo += 1
                    while ip_hi > ip_lo and self._check_pixel_ring(b2, ip_hi, nr, ipix1, fct, z0, phi, cosrsmall, cpix):
                        ip_hi -= 1
                if ip_lo <= ip_hi:
                    if ip_hi >= nr:
                        ip_lo -= nr
                        ip_hi -= nr
                    if ip_lo < 0:
                        append(ipix1, ipix1 + ip_hi + 1)
                        append(ipix1 + ip_lo + nr, ipix2 + 1)
                    else:
                        append(ipix1 + ip_lo, ipix1 + ip_hi + 1)
        if rlat2 >= math.pi and irmax + 1 < 4 * self.nside:  # south pole in the disc
            sp, rp, _ = self.ring_info(irmax + 1)
            append(sp, self.npix)
        return ranges

    def query_disc(self, vec, radius, inclusive=False, fact=4, fast=False):
        if inclusive:
            ranges = self.query_disc_inclusive_ranges(vec, radius, fact, fast)
        else:
            ranges = self.query_disc_ranges(vec, radius)
        if not ranges:
            return np.empty(0, dtype=np.int64)
        return np.concatenate([np.arange(a, b, dtype=np.int64) for a, b in ranges])

    # -- healpix_base.cc: pix2loc (RING) -> z, phi, sth, have_sth -------------------
    def pix2loc(self, pix):
        pix = np.atleast_1d(np.asarray(pix, dtype=np.int64))
        ns, npix, ncap = self.nside, self.npix, self.ncap
        z = np.empty(pix.shape, dtype=np.float64)
        phi = np.empty(pix.shape, dtype=np.float64)
        sth = np.zeros(pix.shape, dtype=np.float64)
        have = np.zeros(pix.shape, dtype=bool)

        north = pix < ncap
        south = pix >= npix - ncap
        eq = ~(north | south)

        if north.any():
            p = pix[north]
            iring = (1 + _isqrt(1 + 2 * p)) >> 1
            iphi = (p + 1) - 2 * iring * (iring - 1)
            tmp = (iring * iring) * self.fact2
            zz = 1.0 - tmp
            z[north] = zz
            big = zz > 0.99
            s = np.zeros_like(zz)
            s[big] = np.sqrt(tmp[big] * (2.0 - tmp[big]))
            sth[north] = s
            have[north] = big
            phi[north] = (iphi - 0.5) * HALFPI / iring
        if eq.any():
            p = pix[eq]
            nl4 = 4 * ns
            ip = p - ncap
            tmp = ip // nl4
            iring = tmp + ns
            iphi = ip - nl4 * tmp + 1
            fodd = np.where(((iring + ns) & 1) != 0, 1.0, 0.5)
            z[eq] = (2 * ns - iring) * self.fact1
            phi[eq] = (iphi - fodd) * math.pi * 0.75 * self.fact1
        if south.any():
            p = pix[south]
            ip = npix - p
            iring = (1 + _isqrt(2 * ip - 1)) >> 1
            iphi = 4 * iring + 1 - (ip - 2 * iring * (iring - 1))
            tmp = (iring * iring) * self.fact2
            zz = tmp - 1.0
            z[south] = zz
            big = zz < -0.99
            s = np.zeros_like(zz)
            s[big] = np.sqrt(tmp[big] * (2.0 - tmp[big]))
            sth[south] = s
            have[south] = big
            phi[south] = (iphi - 0.5) * HALFPI / iring
        return z, phi, sth, have

    def pix2ang(self, pix):
        z, phi, sth, have = self.pix2loc(pix)
        theta = np.where(have, np.arctan2(sth, z), np.arccos(z))
        return theta, phi

    def pix2vec(self, pix):
        z, phi, sth, have = self.pix2loc(pix)
        s = np.where(have, sth, np.sqrt((1.0 - z) * (1.0 + z)))
        return s * np.cos(phi), s * np.sin(phi), z

    # -- healpix_base.cc: loc2pix (RING) --------------------------------------------
    def loc2pix(self, z, phi, sth, have_sth):
        z = np.asarray(z, dtype=np.float64)
        phi = np.asarray(phi, dtype=np.float64)
        sth = np.asarray(sth, dtype=np.float64)
        have_sth = np.broadcast_to(np.asarray(have_sth, dtype=bool), z.shape)
        ns, npix, ncap = self.nside, self.npix, self.ncap
        za = np.abs(z)
        tt = _fmodulo(phi * (1.0 / HALFPI), 4.0)
        pix = np.empty(z.shape, dtype=np.int64)

        eq = za <= TWOTHIRD
        if eq.any():
            nl4 = 4 * ns
            t1 = ns * (0.5 + tt[eq])
            t2 = ns * z[eq] * 0.75
            jp = (t1 - t2).astype(np.int64)
            jm = (t1 + t2).astype(np.int64)
            ir = ns + 1 + jp - jm
            kshift = 1 - (ir & 1)
            t = jp + jm - ns + kshift + 1 + nl4 + nl4
            ip = (t >> 1) & (nl4 - 1) if (ns & (ns - 1)) == 0 else (t >> 1) % nl4
            pix[eq] = ncap + (ir - 1) * nl4 + ip
        po = ~eq
        if po.any():
            ttp = tt[po]
            zp = z[po]
            zap = za[po]
            tp = ttp - ttp.astype(np.int64)
            use_sth = (zap >= 0.99) & have_sth[po]
            with np.errstate(invalid="ignore"):
                tmp = np.where(use_sth,
                               ns * sth[po] / np.sqrt((1.0 + zap) / 3.0),
                               ns * np.sqrt(3.0 * (1.0 - zap)))
            jp = (tp * tmp).astype(np.int64)
            jm = ((1.0 - tp) * tmp).astype(np.int64)
            ir = jp + jm + 1
            ip = (ttp * ir).astype(np.int64)
            ip = np.where(ip >= 4 * ir, ip - 4 * ir, ip)  # imodulo(ip, 4*ir)
            pix[po] = np.where(zp > 0, 2 * ir * (ir - 1) + ip, npix - 2 * ir * (ir + 1) + ip)
        return pix

    def vec2pix(self, x, y, z):
        x = np.asarray(x, dtype=np.float64)
        y = np.asarray(y, dtype=np.float64)
        z = np.asarray(z, dtype=np.float64)
        xl = 1.0 / np.sqrt(x * x + y * y + z * z)
        phi = np.arctan2(y, x)  # safe_atan2
        nz = z * xl
        sth = np.sqrt(x * x + y * y) * xl
        return self.loc2pix(nz, phi, sth, ~(np.abs(nz) < 0.99))

    def ang2pix(self, theta, phi):
        theta = np.asarray(theta, dtype=np.float64)
        phi = np.asarray(phi, dtype=np.float64)
        z = np.cos(theta)
        near_pole = (theta < 0.01) | (theta > 3.14159 - 0.01)
        sth = np.where(near_pole, np.sin(theta), 0.0)
        return self.loc2pix(z, phi, sth, near_pole)


def _isqrt(v):
    """Integer square root of int64 arrays (healpix isqrt: double sqrt + fix-up)."""
    v = np.asarray(v, dtype=np.int64)
    r = np.sqrt(v.astype(np.float64) + 0.5).astype(np.int64)
    r = np.where(r * r > v, r - 1, r)
    r = np.where((r + 1) * (r + 1) <= v, r + 1, r)
    return r


def _fmodulo(v1, v2):
    """healpix fmodulo: result in [0, v2)."""
    v1 = np.asarray(v1, dtype=np.float64)
    out = np.where(v1 >= 0, np.where(v1 < v2, v1, np.fmod(v1, v2)), 0.0)
    neg = v1 < 0
    if neg.any():
        tmp = np.fmod(v1[neg], v2) + v2
        out = out.copy()
        out[neg] = np.where(tmp == v2, 0.0, tmp)
    return out


# ---------------------------------------------------------------------------------
# healpy/rotator.py
# ---------------------------------------------------------------------------------
def dir2vec(theta, phi):
    """healpy.rotator.dir2vec (lonlat=False)."""
    theta = np.asarray(theta, dtype=np.float64)
    phi = np.asarray(phi, dtype=np.float64)
    st = np.sin(theta)
    return np.array([st * np.cos(phi), st * np.sin(phi), np.cos(theta) + 0.0 * phi])


def ang2vec(theta, phi):
    """healpy.ang2vec: returns shape (3,) for scalars or (n, 3)."""
    v = dir2vec(theta, phi)
    return v.T if v.ndim == 2 else v


def angdist(dir1, dir2):
    """healpy.rotator.angdist for (theta, phi) inputs: atan2(|v1 x v2|, v1 . v2)."""
    dir1 = np.asarray(dir1, dtype=np.float64)
    dir2 = np.asarray(dir2, dtype=np.float64)
    v1 = dir2vec(dir1[0], dir1[1]).reshape(3, -1)
    v2 = dir2vec(dir2[0], dir2[1]).reshape(3, -1)
    cx = v1[1] * v2[2] - v1[2] * v2[1]
    cy = v1[2] * v2[0] - v1[0] * v2[2]
    cz = v1[0] * v2[1] - v1[1] * v2[0]
    vec_prod = np.sqrt(cx * cx + cy * cy + cz * cz)
    scal_prod = (v1 * v2).sum(axis=0)
    return np.arctan2(vec_prod, scal_prod)


def rotator_inverse_matrix(lon_deg, lat_deg):
    """Matrix of healpy Rotator(rot=(lon, lat)).I : maps the image frame (halo on +x) to the sky.

    healpy euler_matrix_new(a1=lon, a2=-lat, a3=0, Z=True) builds M = m3 @ m2 @ m1 taking the
    sky direction (lon, lat) to +x; the inverse is its transpose.
    """
    a1 = math.radians(lon_deg)
    a2 = -math.radians(lat_deg)
    c1, s1 = math.cos(a1), math.sin(a1)
    c2, s2 = math.cos(a2), math.sin(a2)
    m1 = np.array([[c1, -s1, 0.0], [s1, c1, 0.0], [0.0, 0.0, 1.0]])  # rot about Z by a1
    m2 = np.array([[c2, 0.0, s2], [0.0, 1.0, 0.0], [-s2, 0.0, c2]])  # rot about Y by a2
    return m1 @ m2


def cartesian_projmap(hpx, healpix_map, lon_deg, lat_deg, lonra, latra, xsize, ysize=None):
    """healpy.projector.CartesianProj(lonra, latra, xsize, ysize).projmap(map, rot=(lon, lat),
    vec2pix_func=partial(hp.vec2pix, nside)) as called at paint_bucket.py:1712-1721.

    Returns the (ysize, xsize) float64 image (nearest-pixel lookup, 'astro' flip).
    """
    lon0, lon1 = float(lonra[0]), float(lonra[1])
    lat0, lat1 = float(latra[0]), float(latra[1])
    # CartesianProj.set_proj_plane_info: lonra[1] = lonra[0] + (np.mod(lonra[1]-lonra[0], 360) or 360)
    dl = math.fmod(lon1 - lon0, 360.0)
    if dl < 0:
        dl += 360.0
    if dl == 0:
        dl = 360.0
    lon1 = lon0 + dl
    if ysize is None:
        ratio = (lat1 - lat0) / (lon1 - lon0)
        ysize = int(round(xsize * ratio))
    # CartesianProj.ij2xy (healpy/projector.py), as recalled by two independent restatements (round-1 advisor and
    # round-2 builder; SURVEY A.6's half-pixel-centre form was a third recollection and is NOT what healpy does):
    #     lonra = flip * np.float64(lonra)[::flip]            # 'astro' flip = -1  ->  [-lonra[1], -lonra[0]]
    #     y = (latra[1] - latra[0]) / (ysize - 1.) * idx_i + latra[0]
    #     x = (lonra[1] - lonra[0]) / (xsize - 1.) * idx_j + lonra[0]
    # i.e. an ENDPOINT-INCLUSIVE grid; healpy is not installed here, so this stays "unpinned" (DESIGN.md section 5).
    j = np.arange(xsize, dtype=np.float64)
    i = np.arange(ysize, dtype=np.float64)
    fl0, fl1 = -lon1, -lon0
    x = ((fl1 - fl0) / (xsize - 1.0)) * j + fl0 if xsize > 1 else np.full(1, fl0)
    y = ((lat1 - lat0) / (ysize - 1.0)) * i + lat0 if ysize > 1 else np.full(1, lat0)
    X, Y = np.meshgrid(x, y)  # shape (ysize, xsize)
    # xy2vec: flip = -1 ('astro'); theta = pi/2 - y, phi = flip * x
    theta = HALFPI - np.radians(Y)
    phi = -np.radians(X)
    st = np.sin(theta)
    v = np.array([st * np.cos(phi), st * np.sin(phi), np.cos(theta)])
    m = rotator_inverse_matrix(lon_deg, lat_deg)
    vs = np.tensordot(m, v, axes=(1, 0))
    pix = hpx.vec2pix(vs[0], vs[1], vs[2])
    return np.asarray(healpix_map)[pix]


# ---------------------------------------------------------------------------------
# healpy.pixelfunc.ud_grade (RING in, RING out, power-of-two nsides)
# ---------------------------------------------------------------------------------
UNSEEN = -1.6375e30


def ring2nest(hp, pix):
    """T_Healpix_Base::ring2nest = xyf2nest(ring2xyf): nest index = face * nside^2 + Morton(ix, iy)."""
    ix, iy, face = hp.ring2xyf(int(pix))
    m = 0
    for b in range(hp.nside.bit_length()):
        m |= ((ix >> b) & 1) << (2 * b)
        m |= ((iy >> b) & 1) << (2 * b + 1)
    return face * hp.nside * hp.nside + m


def ud_grade(map_in, nside_out, pess=False, power=None):
    """healpy.ud_grade(map_in, nside_out, pess=False, order_in='RING', power=None) as the reference's
    Canvas.ud_grade calls it (paint_bucket.py:2233-2257): reorder to NEST, then (degrade) average the good
    children of every parent -- UNSEEN / non-finite children are skipped, a parent with no good child
    (pess: with any bad child) becomes UNSEEN -- or (upgrade) replicate the parent; reorder back to RING."""
    map_in = np.asarray(map_in, dtype=np.float64)
    nside_in = int(round(math.sqrt(len(map_in) / 12)))
    assert 12 * nside_in * nside_in == len(map_in)
    for ns in (nside_in, nside_out):
        assert ns > 0 and (ns & (ns - 1)) == 0, "ud_grade goes through NEST ordering: power-of-two nside"
    hin, hout = HealpixRing(nside_in), HealpixRing(nside_out)
    r2n_in = np.array([ring2nest(hin, p) for p in range(hin.npix)])
    r2n_out = np.array([ring2nest(hout, p) for p in range(hout.npix)])
    nest_in = np.empty_like(map_in)
    nest_in[r2n_in] = map_in
    if nside_out < nside_in:
        rat2 = hin.npix // hout.npix
        ratio = nside_out / nside_in
        mr = nest_in.reshape(hout.npix, rat2)
        bad = (np.abs(mr - UNSEEN) <= 1e-8 + 1e-5 * abs(UNSEEN)) | ~np.isfinite(mr)
        goods = ~bad
        out = np.sum(np.where(goods, mr, 0.0), axis=1)
        nhit = goods.sum(axis=1).astype(np.float64)
        badout = (nhit != rat2) if pess else (nhit == 0)
        if power:
            nhit = nhit / ratio ** power
        nz = nhit != 0
        out[nz] = out[nz] / nhit[nz]
        out[badout] = UNSEEN
        nest_out = out
    elif nside_out > nside_in:
        rat2 = hout.npix // hin.npix
        fact = (nside_out / nside_in) ** power if power else 1.0
        nest_out = np.repeat(nest_in * fact, rat2)
    else:
        nest_out = nest_in
    return nest_out[r2n_out]

