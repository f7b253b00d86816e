// HEALPix RING-scheme geometry for the disc query and the pixel <-> direction maps.
//
// Replaces, on device, the healpy 1.13.0 (healpix_cxx T_Healpix_Base<int64>) calls made by
// the reference's hot path:
//   hp.query_disc  astropaint/paint_bucket.py:1222   -> disc_header() + ring_span()
//   hp.pix2ang / hp.pix2vec  :1238, :1254            -> ring geometry (z, phi) per ring item
//   hp.vec2pix     :1721                             -> vec2pix_ring()
// The published HEALPix algorithm (Gorski et al. 2005) is followed operation by operation
// for everything that decides an integer (ring numbers, span boundaries, pixel numbers) so
// that pixel sets are bit-exact; see oracle/healpix_np.py for the CPU restatement.
#pragma once
#include "common.cuh"

namespace ap {

struct Hpx {
  int64_t nside, npix, ncap;
  double fact1, fact2;
  // inclusive disc queries (healpy query_disc(inclusive=True, fact=incl_fct)): 0 = exclusive query.
  // rad_small / rad_big = max_pixrad() at incl_fct * nside and at nside, formed on the host (hpx_set_inclusive)
  int32_t incl_fct;
  double rad_small, rad_big;
  double pad_cos, pad_sin;  // cos / sin of 1.02 * max_pixrad(nside): the trimming pre-filter of ring_span
  // Ring-band clip of the painting kernels (ap_band_clip): only rings clip_lo <= iz < clip_hi are painted and
  // the map pointer handed to the kernel addresses pixel clip_pix0 (the first pixel of ring clip_lo) -- the
  // band of the map one rank owns.  Default: every ring, clip_pix0 = 0.
  int32_t clip_lo, clip_hi;
  int64_t clip_pix0, clip_pix1;   // first pixel of ring clip_lo, first pixel past ring clip_hi - 1
};

AP_HD Hpx make_hpx(int64_t nside) {
  Hpx h;
  h.nside = nside;
  h.npix = 12 * nside * nside;
  h.ncap = 2 * nside * (nside - 1);
  h.fact2 = 4.0 / (double)h.npix;
  h.fact1 = (double)(nside << 1) * h.fact2;
  h.incl_fct = 0;
  h.rad_small = h.rad_big = 0.0;
  h.pad_cos = 1.0;
  h.pad_sin = 0.0;
  h.clip_lo = 1;
  h.clip_hi = (int32_t)(4 * nside);
  h.clip_pix0 = 0;
  h.clip_pix1 = h.npix;
  return h;
}

// T_Healpix_Base::max_pixrad(): v_angle(set_z_phi(2/3, pi/(4 nside)), set_z_phi(1 - (1 - 1/nside)^2/3, 0))
AP_HD double max_pixrad(int64_t nside) {
  double za = 2.0 / 3.0, pa = kPi / (double)(4 * nside);
  double sa = sqrt(AP_MUL(AP_SUB(1.0, za), AP_ADD(1.0, za)));
  double ax = AP_MUL(sa, cos(pa)), ay = AP_MUL(sa, sin(pa)), az = za;
  double t1 = AP_SUB(1.0, 1.0 / (double)nside);
  t1 = AP_MUL(t1, t1);
  double zb = AP_SUB(1.0, t1 / 3.0);
  double sb = sqrt(AP_MUL(AP_SUB(1.0, zb), AP_ADD(1.0, zb)));
  double bx = AP_MUL(sb, 1.0), by = AP_MUL(sb, 0.0), bz = zb;  // cos(0), sin(0)
  double cx = AP_SUB(AP_MUL(ay, bz), AP_MUL(az, by));
  double cy = AP_SUB(AP_MUL(az, bx), AP_MUL(ax, bz));
  double cz = AP_SUB(AP_MUL(ax, by), AP_MUL(ay, bx));
  double cl = sqrt(AP_ADD(AP_ADD(AP_MUL(cx, cx), AP_MUL(cy, cy)), AP_MUL(cz, cz)));
  double dot = AP_ADD(AP_ADD(AP_MUL(ax, bx), AP_MUL(ay, by)), AP_MUL(az, bz));
  return atan2(cl, dot);
}

// query_disc_internal's radii for fact > 0 (healpix_base.cc): host only (called by the C ABI entry points)
inline void hpx_set_inclusive(Hpx& h, int fct) {
  h.incl_fct = fct;
  if (fct <= 0) {
    h.incl_fct = 0;
    h.rad_small = h.rad_big = 0.0;
  } else if (fct == 1) {
    h.rad_small = h.rad_big = max_pixrad(h.nside);
  } else {
    h.rad_small = max_pixrad(h.nside * fct);
    h.rad_big = max_pixrad(h.nside);
  }
  const double pad = 1.02 * max_pixrad(h.nside);
  h.pad_cos = cos(pad);
  h.pad_sin = sin(pad);
}

AP_HD int64_t ring_above(const Hpx& h, double z) {
  double az = fabs(z);
  if (az <= kTwoThird) return (int64_t)AP_MUL((double)h.nside, AP_SUB(2.0, AP_MUL(1.5, z)));
  int64_t iring = (int64_t)AP_MUL((double)h.nside, sqrt(AP_MUL(3.0, AP_SUB(1.0, az))));
  return (z > 0) ? iring : 4 * h.nside - iring - 1;
}

AP_HD double ring2z(const Hpx& h, int64_t ring) {
  // (double)(ring * ring) is exact either way; squaring in fp64 keeps the integer pipe out of it
  const int32_t r = (int32_t)ring, ns = (int32_t)h.nside;
  if (r < ns) return AP_SUB(1.0, AP_MUL(AP_MUL((double)r, (double)r), h.fact2));
  if (r <= 3 * ns) return AP_MUL((double)(2 * ns - r), h.fact1);
  const double q = (double)(4 * ns - r);
  return AP_SUB(AP_MUL(AP_MUL(q, q), h.fact2), 1.0);
}

AP_HD void ring_info(const Hpx& h, int64_t ring, int64_t& startpix, int64_t& ringpix, bool& shifted) {
  const int32_t r = (int32_t)ring, ns = (int32_t)h.nside;  // ring numbers fit 32 bits (nside <= 2^18)
  if (r < ns) {
    shifted = true;
    ringpix = 4 * r;
    startpix = (int64_t)(2 * r) * (r - 1);
  } else if (r < 3 * ns) {
    shifted = ((r - ns) & 1) == 0;
    ringpix = 4 * ns;
    startpix = h.ncap + (int64_t)(r - ns) * (4 * ns);
  } else {
    shifted = true;
    const int32_t nr = 4 * ns - r;
    ringpix = 4 * nr;
    startpix = h.npix - (int64_t)(2 * nr) * (nr + 1);
  }
}

// Per-halo part of query_disc_internal (RING, fact == 0).  Ring items are the rings
// ir_first..ir_last; ring iz is a FULL ring (pole cap inside the disc) when iz < irmin or
// iz > irmax, otherwise its span comes from ring_span().
struct DiscHeader {
  double phi, z0, xa, cosr;  // cosr = cos(rbig): the span boundaries
  double cosrs;              // cos(rsmall): check_pixel_ring's threshold (inclusive queries only)
  int64_t irmin, irmax, ir_first, ir_last;
};

// INCL = false compiles the inclusive branches away: inlined next to the exclusive (default) query they cost
// K1 40% and K3 10% (measured), so the painter kernels are instantiated for both and pick at launch.
template <bool INCL>
AP_HD DiscHeader disc_header_t(const Hpx& h, double x, double y, double z, double radius) {
  DiscHeader d;
  // pointing(vec3): theta = atan2(sqrt(x^2+y^2), z); phi = atan2(y, x) (+2pi if negative)
  double theta = atan2(sqrt(AP_ADD(AP_MUL(x, x), AP_MUL(y, y))), z);
  double phi = atan2(y, x);
  if (phi < 0.0) phi = AP_ADD(phi, kTwoPi);
  d.phi = phi;
  const int fct = INCL ? h.incl_fct : 0;
  double rsmall = radius, rbig = radius;
  if (fct) {  // inclusive: radius enlarged by the pixel radius at fct * nside (ring range, trimming) and nside (spans)
    rsmall = AP_ADD(radius, h.rad_small);
    rbig = AP_ADD(radius, h.rad_big);
  }
  d.cosrs = 0.0;
  if (rsmall >= kPi) {  // whole sky
    d.z0 = d.xa = d.cosr = 0.0;
    d.ir_first = 1;
    d.ir_last = 4 * h.nside - 1;
    d.irmin = 4 * h.nside;
    d.irmax = 4 * h.nside;
    return d;
  }
  if (fct) {
    rbig = fmin(kPi, rbig);
    d.cosrs = cos(rsmall);
  }
  d.cosr = cos(rbig);
  d.z0 = cos(theta);
  d.xa = 1.0 / sqrt(AP_MUL(AP_SUB(1.0, d.z0), AP_ADD(1.0, d.z0)));
  double rlat1 = AP_SUB(theta, rsmall);
  double zmax = cos(rlat1);
  d.irmin = ring_above(h, zmax) + 1;
  d.ir_first = d.irmin;
  if (rlat1 <= 0 && d.irmin > 1) d.ir_first = 1;  // north pole in the disc
  if (fct > 1 && rlat1 > 0) {
    d.irmin = d.irmin > 2 ? d.irmin - 1 : 1;
    d.ir_first = d.irmin;
  }
  double rlat2 = AP_ADD(theta, rsmall);
  double zmin = cos(rlat2);
  d.irmax = ring_above(h, zmin);
  if (fct > 1 && rlat2 < kPi) d.irmax = d.irmax + 1 < 4 * h.nside - 1 ? d.irmax + 1 : 4 * h.nside - 1;
  d.ir_last = d.irmax;
  if (rlat2 >= kPi && d.irmax + 1 < 4 * h.nside) d.ir_last = 4 * h.nside - 1;  // south pole
  return d;
}

AP_HD DiscHeader disc_header(const Hpx& h, double x, double y, double z, double radius) {
  return h.incl_fct ? disc_header_t<true>(h, x, y, z, radius) : disc_header_t<false>(h, x, y, z, radius);
}

AP_HD int64_t disc_n_rings(const DiscHeader& d) {
  int64_t n = d.ir_last - d.ir_first + 1;
  return n > 0 ? n : 0;
}

// One ring of the disc.  The pixels are startpix + ((ip_lo + j) mod nr), j = 0..len-1, with
// ip_lo possibly negative (span wrapping through phi = 0).  Returns len (0 = ring not hit).
struct RingSpan {
  int64_t startpix;
  int32_t nr, ip_lo, len;
  bool shifted;
  double z;
};

AP_HD bool check_pixel_ring(const Hpx& b1, const Hpx& b2, int64_t pix, int64_t nr, int64_t ipix1, int fct, double cz,
                            double cphi, double cosrp2, int64_t cpix);
AP_HD int64_t loc2pix_ring(const Hpx& h, double nz, double tt, double sth, bool have_sth = true);
AP_HD double tt_of_phi(double phi);

template <bool INCL>
AP_HD RingSpan ring_span_t(const Hpx& h, const DiscHeader& d, int64_t iz) {
  RingSpan s;
  int64_t nr;
  ring_info(h, iz, s.startpix, nr, s.shifted);
  s.nr = (int32_t)nr;
  s.z = ring2z(h, iz);
  s.ip_lo = 0;
  s.len = 0;
  if (iz < d.irmin || iz > d.irmax) {  // ring inside a pole cap that the disc swallows
    s.len = s.nr;
    return s;
  }
  double z = s.z;
  double x = AP_MUL(AP_SUB(d.cosr, AP_MUL(z, d.z0)), d.xa);
  double ysq = AP_SUB(AP_SUB(1.0, AP_MUL(z, z)), AP_MUL(x, x));
  double dphi;
  if (ysq <= 0) {  // no intersection: ring completely inside or outside the (enlarged) disc
    if (!INCL || h.incl_fct <= 1) return s;
    dphi = kPi - 1e-15;
  } else {
    dphi = atan2(sqrt(ysq), x);
  }
  if (!(dphi > 0)) return s;
  double shift = s.shifted ? 0.5 : 0.0;
  double f = AP_MUL((double)s.nr, kInvTwoPi);
  // |floor(...)| <= 2 nr < 2^21: the conversions stay in 32 bits
  int32_t ip_lo = (int32_t)floor(AP_SUB(AP_MUL(f, AP_SUB(d.phi, dphi)), shift)) + 1;
  int32_t ip_hi = (int32_t)floor(AP_SUB(AP_MUL(f, AP_ADD(d.phi, dphi)), shift));
  if (INCL && h.incl_fct > 1) {  // trim both ends while the end pixel provably does not overlap the disc
    const Hpx b2 = make_hpx(h.nside * h.incl_fct);
    const int64_t cpix = loc2pix_ring(h, d.z0, tt_of_phi(d.phi), 0.0, false);  // zphi2pix(z0, phi)
    // Pre-filter, equivalent to the two loops below: a pixel whose CENTRE is farther from the disc centre
    // than rsafe = rsmall + 1.02 max_pixrad(nside) has every boundary sub-pixel centre farther than rsmall
    // (each lies within max_pixrad of the pixel centre), so check_pixel_ring is certainly true for it and
    // the loops would step over it.  The centres within rsafe form the azimuth range phi +- dsafe; the
    // loops start at its ends (one pixel of slack).  Without this, a ring that the widened ring range adds
    // just outside the disc (ysq <= 0 above) is walked pixel by pixel: 4 nside * 12 sub-pixel tests.
    if (d.cosrs > -h.pad_cos) {  // rsafe < pi
      const double cs = AP_SUB(AP_MUL(d.cosrs, h.pad_cos),
                               AP_MUL(sqrt(fmax(0.0, AP_SUB(1.0, AP_MUL(d.cosrs, d.cosrs)))), h.pad_sin));
      const double sring = sqrt(AP_MUL(AP_SUB(1.0, z), AP_ADD(1.0, z)));
      const double c = AP_MUL(AP_SUB(cs, AP_MUL(z, d.z0)), d.xa) / sring;
      if (c >= 1.0) {
        ip_lo = ip_hi + 1;  // no pixel centre of this ring within rsafe: everything is trimmed
      } else if (c > -1.0) {
        const double dsafe = AP_ADD(acos(c), 1e-9);
        const double glo = floor(AP_SUB(AP_MUL(f, AP_SUB(d.phi, dsafe)), shift)) - 1.0;
        const double ghi = floor(AP_SUB(AP_MUL(f, AP_ADD(d.phi, dsafe)), shift)) + 2.0;
        if (glo > (double)ip_lo) ip_lo = (int32_t)fmin(glo, (double)ip_hi + 1.0);
        if (ghi < (double)ip_hi) ip_hi = (int32_t)fmax(ghi, (double)ip_lo - 1.0);
      }
    }
    while (ip_lo <= ip_hi && check_pixel_ring(h, b2, ip_lo, s.nr, s.startpix, h.incl_fct, d.z0, d.phi, d.cosrs, cpix))
      ++ip_lo;
    while (ip_hi > ip_lo && check_pixel_ring(h, b2, ip_hi, s.nr, s.startpix, h.incl_fct, d.z0, d.phi, d.cosrs, cpix))
      --ip_hi;
  }
  if (ip_lo > ip_hi) return s;
  if (ip_hi >= s.nr) {
    ip_lo -= s.nr;
    ip_hi -= s.nr;
  }
  s.ip_lo = ip_lo;
  s.len = ip_hi - ip_lo + 1;
  return s;
}

// Near-tie audit of the exclusive query: how close the two floor() arguments of ring iz come to an integer
// (in pixels).  A span boundary can differ between two libm implementations (CUDA's atan2 / cos vs glibc's, <= 2 ulp
// apart) only when this margin is of the order of 1e-11; ap_disc_near_ties lists the halos below a tolerance so that
// the host can recompute exactly those with its own libm.  +inf for rings the disc does not cross.
// distance of ring_above()'s truncated argument from the next integer boundary (the ring range irmin / irmax)
AP_HD double ring_above_margin(const Hpx& h, double z) {
  const double az = fabs(z);
  const double a = (az <= kTwoThird) ? AP_MUL((double)h.nside, AP_SUB(2.0, AP_MUL(1.5, z)))
                                     : AP_MUL((double)h.nside, sqrt(AP_MUL(3.0, AP_SUB(1.0, az))));
  return fmin(fabs(a - rint(a)), fabs(az - kTwoThird) * (double)h.nside);
}

AP_HD double ring_tie_margin(const Hpx& h, const DiscHeader& d, int64_t iz) {
  if (iz < d.irmin || iz > d.irmax) return INFINITY;
  int64_t startpix, nr;
  bool shifted;
  ring_info(h, iz, startpix, nr, shifted);
  const double z = ring2z(h, iz);
  const double x = AP_MUL(AP_SUB(d.cosr, AP_MUL(z, d.z0)), d.xa);
  const double ysq = AP_SUB(AP_SUB(1.0, AP_MUL(z, z)), AP_MUL(x, x));
  if (ysq <= 0) return fabs(ysq) * (double)nr;   // the ring grazes the disc: report how close to "no intersection"
  const double dphi = atan2(sqrt(ysq), x);
  if (!(dphi > 0)) return 0.0;
  const double shift = shifted ? 0.5 : 0.0;
  const double f = AP_MUL((double)nr, kInvTwoPi);
  const double a = AP_SUB(AP_MUL(f, AP_SUB(d.phi, dphi)), shift), b = AP_SUB(AP_MUL(f, AP_ADD(d.phi, dphi)), shift);
  return fmin(fabs(a - rint(a)), fabs(b - rint(b)));
}

AP_HD RingSpan ring_span(const Hpx& h, const DiscHeader& d, int64_t iz) {
  return h.incl_fct ? ring_span_t<true>(h, d, iz) : ring_span_t<false>(h, d, iz);
}

// j-th pixel (0 <= j < len) of a span in ASCENDING pixel order (the order healpy returns).
AP_HD int64_t span_pixel_sorted(const RingSpan& s, int32_t j) {
  if (s.ip_lo >= 0) return s.startpix + s.ip_lo + j;
  int32_t ip_hi = s.ip_lo + s.len - 1;  // >= 0 part is [0, ip_hi]
  if (j <= ip_hi) return s.startpix + j;
  return s.startpix + (s.ip_lo + s.nr) + (j - ip_hi - 1);
}

// Ring-local index of the j-th pixel in painting order and its azimuth index before wrapping.
AP_HD int32_t span_ip_raw(const RingSpan& s, int32_t j) { return s.ip_lo + j; }

// Colatitude of a ring as healpy's pix2ang returns it: atan2(sth, z) when |z| > 0.99
// (pix2loc supplies sth there), acos(z) otherwise.  Also returns sin(theta) as pix2vec uses it.
AP_HD void ring_theta(const Hpx& h, int64_t iz, double z, double& theta, double& sth) {
  bool cap = (iz < h.nside) || (iz > 3 * h.nside);
  if (cap && fabs(z) > 0.99) {
    int64_t ir = (iz < h.nside) ? iz : 4 * h.nside - iz;
    double tmp = AP_MUL((double)(ir * ir), h.fact2);
    sth = sqrt(AP_MUL(tmp, AP_SUB(2.0, tmp)));
    theta = atan2(sth, z);
  } else {
    sth = sqrt(AP_MUL(AP_SUB(1.0, z), AP_ADD(1.0, z)));
    theta = acos(z);
  }
}

// sin(theta) of a ring as healpix pix2loc / pix2vec forms it
AP_HD double ring_sth(const Hpx& h, int64_t iz, double z) {
  bool cap = (iz < h.nside) || (iz > 3 * h.nside);
  if (cap && fabs(z) > 0.99) {
    int64_t ir = (iz < h.nside) ? iz : 4 * h.nside - iz;
    double tmp = AP_MUL((double)(ir * ir), h.fact2);
    return sqrt(AP_MUL(tmp, AP_SUB(2.0, tmp)));
  }
  return sqrt(AP_MUL(AP_SUB(1.0, z), AP_ADD(1.0, z)));
}

// loc2pix (RING) from nz = z/|v|, tt = fmodulo(phi / (pi/2), 4) and, for |nz| >= 0.99, sth = sin(theta).
AP_HD int64_t loc2pix_ring(const Hpx& h, double nz, double tt, double sth, bool have_sth) {
  double za = fabs(nz);
  int64_t ns = h.nside;
  if (za <= kTwoThird) {
    // ring / in-ring indices fit 32 bits (nside <= 2^18 is enforced by the C ABI); only the pixel
    // number needs 64
    int32_t n32 = (int32_t)ns, nl4 = 4 * n32;
    double t1 = AP_MUL((double)ns, AP_ADD(0.5, tt));
    double t2 = AP_MUL(AP_MUL((double)ns, nz), 0.75);
    int32_t jp = (int32_t)AP_SUB(t1, t2);
    int32_t jm = (int32_t)AP_ADD(t1, t2);
    int32_t ir = n32 + 1 + jp - jm;
    int32_t kshift = 1 - (ir & 1);
    int32_t t = jp + jm - n32 + kshift + 1 + nl4 + nl4;
    int32_t ip = ((n32 & (n32 - 1)) == 0) ? ((t >> 1) & (nl4 - 1)) : ((t >> 1) % nl4);
    return h.ncap + (int64_t)(ir - 1) * nl4 + ip;
  }
  double tp = AP_SUB(tt, (double)(int64_t)tt);
  double tmp;
  if (za < 0.99 || !have_sth) {
    tmp = AP_MUL((double)ns, sqrt(AP_MUL(3.0, AP_SUB(1.0, za))));
  } else {
    tmp = AP_MUL((double)ns, sth) / sqrt(AP_ADD(1.0, za) / 3.0);
  }
  int64_t jp = (int64_t)AP_MUL(tp, tmp);
  int64_t jm = (int64_t)AP_MUL(AP_SUB(1.0, tp), tmp);
  int64_t ir = jp + jm + 1;
  int64_t ip = (int64_t)AP_MUL(tt, (double)ir);
  if (ip >= 4 * ir) ip -= 4 * ir;
  return (nz > 0) ? 2 * ir * (ir - 1) + ip : h.npix - 2 * ir * (ir + 1) + ip;
}

AP_HD double tt_of_phi(double phi) {  // fmodulo(phi / halfpi, 4)
  double tt = AP_MUL(phi, 1.0 / kHalfPi);
  if (tt < 0) {
    tt = AP_ADD(fmod(tt, 4.0), 4.0);
    if (tt == 4.0) tt = 0.0;
  } else if (tt >= 4.0) {
    tt = fmod(tt, 4.0);
  }
  return tt;
}

// loc2pix / vec2pix, RING scheme (hp.vec2pix, paint_bucket.py:1721).
AP_HD int64_t vec2pix_ring(const Hpx& h, double x, double y, double z) {
  double xl = 1.0 / sqrt(AP_ADD(AP_ADD(AP_MUL(x, x), AP_MUL(y, y)), AP_MUL(z, z)));
  double phi = atan2(y, x);
  double nz = AP_MUL(z, xl);
  double sth = (fabs(nz) < 0.99) ? 0.0 : AP_MUL(sqrt(AP_ADD(AP_MUL(x, x), AP_MUL(y, y))), xl);
  return loc2pix_ring(h, nz, tt_of_phi(phi), sth);
}

// vec2pix for a direction known to lie close in azimuth to a reference azimuth phi_c (a cutout sample
// around its halo): phi = phi_c + atan(u), u = tan(phi - phi_c) from one cross and one dot product,
// atan by its series for |u| <= 1/8; otherwise the general atan2.  Identical to vec2pix_ring except
// where tt lies within ~1e-15 of a pixel boundary (phi carries one extra rounding).
AP_HD int64_t vec2pix_ring_near(const Hpx& h, double x, double y, double z, double phi_c, double cphi, double sphi) {
#if defined(__CUDA_ARCH__)
  double xl = rsqrt(AP_ADD(AP_ADD(AP_MUL(x, x), AP_MUL(y, y)), AP_MUL(z, z)));  // <= 1 ulp from 1/sqrt
#else
  double xl = 1.0 / sqrt(AP_ADD(AP_ADD(AP_MUL(x, x), AP_MUL(y, y)), AP_MUL(z, z)));
#endif
  double dot = x * cphi + y * sphi;
  double crs = y * cphi - x * sphi;
  double phi;
  if (dot > 0.0 && fabs(crs) <= 0.125 * dot) {
#if defined(__CUDA_ARCH__)
    double u = crs * __drcp_rn(dot), u2 = u * u;
#else
    double u = crs / dot, u2 = u * u;
#endif
    // atan(u)/u = 1 - u^2/3 + u^4/5 - ... - u^18/19 (next term 8^-20/21 < 5e-20)
    double p = -1.0 / 19.0;
    p = p * u2 + 1.0 / 17.0;
    p = p * u2 - 1.0 / 15.0;
    p = p * u2 + 1.0 / 13.0;
    p = p * u2 - 1.0 / 11.0;
    p = p * u2 + 1.0 / 9.0;
    p = p * u2 - 1.0 / 7.0;
    p = p * u2 + 1.0 / 5.0;
    p = p * u2 - 1.0 / 3.0;
    phi = phi_c + (u + u * u2 * p);
  } else {
    phi = atan2(y, x);
  }
  double nz = AP_MUL(z, xl);
  double sth = (fabs(nz) < 0.99) ? 0.0 : AP_MUL(sqrt(AP_ADD(AP_MUL(x, x), AP_MUL(y, y))), xl);
  return loc2pix_ring(h, nz, tt_of_phi(phi), sth);
}


// ---- inclusive query helpers (healpix_base.cc: ring2xyf, xyf2ring, pix2loc, check_pixel_ring) ----
AP_HD int64_t hp_isqrt(int64_t v) {
  int64_t r = (int64_t)sqrt((double)v + 0.5);
  if (r * r > v) --r;
  if ((r + 1) * (r + 1) <= v) ++r;
  return r;
}

AP_HD int hp_special_div(int64_t a, int64_t b) {  // a / b for 0 <= a < 4b
  int t = (a >= (b << 1)) ? 1 : 0;
  a -= t * (b << 1);
  return (t << 1) + ((a >= b) ? 1 : 0);
}

AP_HD void ring2xyf(const Hpx& h, int64_t pix, int64_t& ix, int64_t& iy, int& face) {
  const int jpll[12] = {1, 3, 5, 7, 0, 2, 4, 6, 1, 3, 5, 7};
  const int64_t ns = h.nside, nl2 = 2 * ns;
  int64_t iring, iphi, kshift, nr;
  if (pix < h.ncap) {
    iring = (1 + hp_isqrt(1 + 2 * pix)) >> 1;
    iphi = (pix + 1) - 2 * iring * (iring - 1);
    kshift = 0;
    nr = iring;
    face = hp_special_div(iphi - 1, nr);
  } else if (pix < h.npix - h.ncap) {
    int64_t ip = pix - h.ncap;
    int64_t tmp = ip / (4 * ns);
    iring = tmp + ns;
    iphi = ip - tmp * 4 * ns + 1;
    kshift = (iring + ns) & 1;
    nr = ns;
    int64_t ire = tmp + 1, irm = nl2 + 1 - tmp;
    int64_t ifm = (iphi - (ire >> 1) + ns - 1) / ns, ifp = (iphi - (irm >> 1) + ns - 1) / ns;
    face = (ifp == ifm) ? (int)(ifp | 4) : ((ifp < ifm) ? (int)ifp : (int)(ifm + 8));
  } else {
    int64_t ip = h.npix - pix;
    iring = (1 + hp_isqrt(2 * ip - 1)) >> 1;
    iphi = 4 * iring + 1 - (ip - 2 * iring * (iring - 1));
    kshift = 0;
    nr = iring;
    iring = 2 * nl2 - iring;
    face = hp_special_div(iphi - 1, nr) + 8;
  }
  int64_t irt = iring - ((2 + (face >> 2)) * ns) + 1;
  int64_t ipt = 2 * iphi - jpll[face] * nr - kshift - 1;
  if (ipt >= nl2) ipt -= 8 * ns;
  ix = (ipt - irt) >> 1;
  iy = (-ipt - irt) >> 1;
}

AP_HD int64_t xyf2ring(const Hpx& h, int64_t ix, int64_t iy, int face) {
  const int jrll[12] = {2, 2, 2, 2, 3, 3, 3, 3, 4, 4, 4, 4};
  const int jpll[12] = {1, 3, 5, 7, 0, 2, 4, 6, 1, 3, 5, 7};
  const int64_t nl4 = 4 * h.nside;
  int64_t jr = jrll[face] * h.nside - ix - iy - 1;
  int64_t n_before, nr;
  bool shifted;
  ring_info(h, jr, n_before, nr, shifted);
  nr >>= 2;
  int64_t kshift = shifted ? 0 : 1;
  int64_t jp = (jpll[face] * nr + ix - iy + 1 + kshift) / 2;
  if (jp < 1) jp += nl4;
  return n_before + jp - 1;
}

// pix2loc (RING): z and phi of a pixel centre
AP_HD void pix2zphi_ring(const Hpx& h, int64_t pix, double& z, double& phi) {
  const int64_t ns = h.nside;
  if (pix < h.ncap) {
    int64_t iring = (1 + hp_isqrt(1 + 2 * pix)) >> 1;
    int64_t iphi = (pix + 1) - 2 * iring * (iring - 1);
    double tmp = AP_MUL((double)(iring * iring), h.fact2);
    z = AP_SUB(1.0, tmp);
    phi = AP_MUL(AP_SUB((double)iphi, 0.5), kHalfPi) / (double)iring;
  } else if (pix < h.npix - h.ncap) {
    int64_t nl4 = 4 * ns;
    int64_t ip = pix - h.ncap;
    int64_t tmp = ip / nl4;
    int64_t iring = tmp + ns, iphi = ip - nl4 * tmp + 1;
    double fodd = ((iring + ns) & 1) ? 1.0 : 0.5;
    z = AP_MUL((double)(2 * ns - iring), h.fact1);
    phi = AP_MUL(AP_MUL(AP_MUL(AP_SUB((double)iphi, fodd), kPi), 0.75), h.fact1);
  } else {
    int64_t ip = h.npix - pix;
    int64_t iring = (1 + hp_isqrt(2 * ip - 1)) >> 1;
    int64_t iphi = 4 * iring + 1 - (ip - 2 * iring * (iring - 1));
    double tmp = AP_MUL((double)(iring * iring), h.fact2);
    z = AP_SUB(tmp, 1.0);
    phi = AP_MUL(AP_SUB((double)iphi, 0.5), kHalfPi) / (double)iring;
  }
}

// hp.pix2ang (RING; paint_bucket.py:1085, 1199, 1238): pix2loc, then theta = atan2(sth, z) where pix2loc supplies
// sin(theta) (polar-cap pixels with |z| > 0.99) and acos(z) elsewhere -- healpix_base.h pix2ang / pointing(loc).
AP_HD void pix2ang_ring(const Hpx& h, int64_t pix, double& theta, double& phi) {
  double z;
  pix2zphi_ring(h, pix, z, phi);
  const bool north = pix < h.ncap, south = pix >= h.npix - h.ncap;
  if ((north || south) && fabs(z) > 0.99) {
    const int64_t iring = north ? (1 + hp_isqrt(1 + 2 * pix)) >> 1 : (1 + hp_isqrt(2 * (h.npix - pix) - 1)) >> 1;
    const double tmp = AP_MUL((double)(iring * iring), h.fact2);
    theta = atan2(sqrt(AP_MUL(tmp, AP_SUB(2.0, tmp))), z);
  } else {
    theta = acos(z);
  }
}

// hp.ang2pix (RING; paint_bucket.py:996, 1183): healpix_base.h ang2pix -- sin(theta) is handed to loc2pix only
// within 0.01 rad of a pole (the literal 3.14159 is healpix's)
AP_HD int64_t ang2pix_ring(const Hpx& h, double theta, double phi) {
  const bool near_pole = (theta < 0.01) || (theta > 3.14159 - 0.01);
  return loc2pix_ring(h, cos(theta), tt_of_phi(phi), near_pole ? sin(theta) : 0.0, near_pole);
}

AP_HD double cosdist_zphi(double z1, double phi1, double z2, double phi2) {
  return AP_ADD(AP_MUL(z1, z2), AP_MUL(cos(AP_SUB(phi1, phi2)),
                                       sqrt(AP_MUL(AP_SUB(1.0, AP_MUL(z1, z1)), AP_SUB(1.0, AP_MUL(z2, z2))))));
}

// true = none of the 4 (fct - 1) boundary sub-pixel centres (resolution b2 = fct * nside) of ring pixel
// `pix` lies within the enlarged radius: the pixel does not overlap the disc and may be trimmed
AP_HD bool check_pixel_ring(const Hpx& b1, const Hpx& b2, int64_t pix, int64_t nr, int64_t ipix1, int fct, double cz,
                            double cphi, double cosrp2, int64_t cpix) {
  if (pix >= nr) pix -= nr;
  if (pix < 0) pix += nr;
  pix += ipix1;
  if (pix == cpix) return false;  // disc centre in the pixel => overlap
  int64_t px, py;
  int pf;
  ring2xyf(b1, pix, px, py, pf);
  const int64_t ox = fct * px, oy = fct * py;
  double pz, pphi;
  for (int i = 0; i < fct - 1; ++i) {  // go along the 4 edges
    pix2zphi_ring(b2, xyf2ring(b2, ox + i, oy, pf), pz, pphi);
    if (cosdist_zphi(pz, pphi, cz, cphi) > cosrp2) return false;
    pix2zphi_ring(b2, xyf2ring(b2, ox + fct - 1, oy + i, pf), pz, pphi);
    if (cosdist_zphi(pz, pphi, cz, cphi) > cosrp2) return false;
    pix2zphi_ring(b2, xyf2ring(b2, ox + fct - 1 - i, oy + fct - 1, pf), pz, pphi);
    if (cosdist_zphi(pz, pphi, cz, cphi) > cosrp2) return false;
    pix2zphi_ring(b2, xyf2ring(b2, ox, oy + fct - 1 - i, pf), pz, pphi);
    if (cosdist_zphi(pz, pphi, cz, cphi) > cosrp2) return false;
  }
  return true;
}

}  // namespace ap
