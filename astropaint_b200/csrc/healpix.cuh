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
};

AP_HD Hpx make_hpx(int64_t nside) {
  Hpx h;
  h.nside = nside;
  h.npix = 12 * nside * nside;
  h.ncap = 2 * nside * (nside - 1);
  h.fact2 = 4.0 / (double)h.npix;
  h.fact1 = (double)(nside << 1) * h.fact2;
  return h;
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
  double phi, z0, xa, cosr;
  int64_t irmin, irmax, ir_first, ir_last;
};

AP_HD DiscHeader disc_header(const Hpx& h, double x, double y, double z, double radius) {
  DiscHeader d;
  // pointing(vec3): theta = atan2(sqrt(x^2+y^2), z); phi = atan2(y, x) (+2pi if negative)
  double theta = atan2(sqrt(AP_ADD(AP_MUL(x, x), AP_MUL(y, y))), z);
  double phi = atan2(y, x);
  if (phi < 0.0) phi = AP_ADD(phi, kTwoPi);
  d.phi = phi;
  if (radius >= kPi) {  // whole sky
    d.z0 = d.xa = d.cosr = 0.0;
    d.ir_first = 1;
    d.ir_last = 4 * h.nside - 1;
    d.irmin = 4 * h.nside;
    d.irmax = 4 * h.nside;
    return d;
  }
  d.cosr = cos(radius);
  d.z0 = cos(theta);
  d.xa = 1.0 / sqrt(AP_MUL(AP_SUB(1.0, d.z0), AP_ADD(1.0, d.z0)));
  double rlat1 = AP_SUB(theta, radius);
  double zmax = cos(rlat1);
  d.irmin = ring_above(h, zmax) + 1;
  d.ir_first = d.irmin;
  if (rlat1 <= 0 && d.irmin > 1) d.ir_first = 1;  // north pole in the disc
  double rlat2 = AP_ADD(theta, radius);
  double zmin = cos(rlat2);
  d.irmax = ring_above(h, zmin);
  d.ir_last = d.irmax;
  if (rlat2 >= kPi && d.irmax + 1 < 4 * h.nside) d.ir_last = 4 * h.nside - 1;  // south pole
  return d;
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

AP_HD RingSpan ring_span(const Hpx& h, const DiscHeader& d, int64_t iz) {
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
  if (ysq <= 0) return s;
  double dphi = atan2(sqrt(ysq), x);
  if (!(dphi > 0)) return s;
  double shift = s.shifted ? 0.5 : 0.0;
  double f = AP_MUL((double)s.nr, kInvTwoPi);
  // |floor(...)| <= 2 nr < 2^21: the conversions stay in 32 bits
  int32_t ip_lo = (int32_t)floor(AP_SUB(AP_MUL(f, AP_SUB(d.phi, dphi)), shift)) + 1;
  int32_t ip_hi = (int32_t)floor(AP_SUB(AP_MUL(f, AP_ADD(d.phi, dphi)), shift));
  if (ip_lo > ip_hi) return s;
  if (ip_hi >= s.nr) {
    ip_lo -= s.nr;
    ip_hi -= s.nr;
  }
  s.ip_lo = ip_lo;
  s.len = ip_hi - ip_lo + 1;
  return s;
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
AP_HD int64_t loc2pix_ring(const Hpx& h, double nz, double tt, double sth) {
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
  if (za < 0.99) {
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

}  // namespace ap
