// Pixel-centre -> halo-centre separation for one ring of a disc.
//
// The reference forms, per halo, pix2ang -> dir2vec -> atan2(|p x c|, p . c) (gen_cent2pix_rad,
// astropaint/paint_bucket.py:1256-1264; healpy.rotator.angdist) and R = D_a * theta (:1266-1273).
// All pixels of one ring share theta_p, so the separation is evaluated in haversine form with
// the ring-constant parts hoisted:
//   hav(theta) = sin^2((theta_p - theta_c)/2) + sin(theta_p) sin(theta_c) sin^2((phi_p - phi_c)/2)
// which is accurate to a few ulp for every separation (no cancellation at small theta), well
// inside the 1e-6 map tolerance.  R_vec mode (gen_cent2pix_mpc_vec, :1283-1290) needs the
// pixel vector itself: p = (sth cos(phi), sth sin(phi), z), c = dir2vec(theta_c, phi_c).
#pragma once
#include "common.cuh"
#include "healpix.cuh"
#include "fastmath.cuh"

namespace ap {

struct HaloCenter {
  double theta, phi, sin_t, cos_t, D_a;
  double cx, cy, cz;
};

AP_HD HaloCenter make_center(double theta, double phi, double D_a, bool need_vec) {
  HaloCenter c;
  c.theta = theta;
  c.phi = phi;
  c.D_a = D_a;
  c.sin_t = sin(theta);
  c.cos_t = cos(theta);
  c.cx = c.cy = c.cz = 0.0;
  if (need_vec) {  // hp.ang2vec(theta, phi), :1212
    c.cx = c.sin_t * cos(phi);
    c.cy = c.sin_t * sin(phi);
    c.cz = c.cos_t;
  }
  return c;
}

struct RingGeom {
  double A, B;           // hav = A + B sin^2(d / 2)
  double phi_step, d0;   // d(j) = d0 + j * phi_step : azimuth of span pixel j minus phi_c, d0 in [-pi, pi)
  double phi0;           // azimuth of span pixel 0 (R_vec mode)
  double z, sth;
};

AP_HD RingGeom ring_geom(const Hpx& h, int64_t iz, const RingSpan& s, const HaloCenter& c) {
  RingGeom g;
  g.sth = ring_sth(h, iz, s.z);
  g.z = s.z;
  // A = sin^2((theta_p - theta_c)/2) without any inverse trigonometry:
  //   z_p - z_c = -2 sin((theta_p+theta_c)/2) sin((theta_p-theta_c)/2),
  //   sin^2((theta_p+theta_c)/2) = (1 - z_p z_c + s_p s_c)/2
  // (the reference goes through theta_p = acos(z_p), pix2ang; identical to ~1e-12 relative)
  double dz = s.z - c.cos_t;
#if defined(__CUDA_ARCH__)
  g.A = (dz * dz) * __drcp_rn(2.0 * ((1.0 - s.z * c.cos_t) + g.sth * c.sin_t));
  g.phi_step = kTwoPi * __drcp_rn((double)s.nr);
#else
  g.A = (dz * dz) / (2.0 * ((1.0 - s.z * c.cos_t) + g.sth * c.sin_t));
  g.phi_step = kTwoPi / (double)s.nr;
#endif
  g.B = g.sth * c.sin_t;
  g.phi0 = ((double)s.ip_lo + (s.shifted ? 0.5 : 0.0)) * g.phi_step;
  double d = g.phi0 - c.phi;
  g.d0 = d - kTwoPi * rint(d * kInvTwoPi);
  return g;
}

// span-local index (painting order, pixel = startpix + (ip_lo + j) mod nr) of ring-local index ip
AP_HD int32_t span_j_of_ip(const RingSpan& s, int32_t ip) {
  int32_t j = ip - s.ip_lo;
  return j >= s.nr ? j - s.nr : j;
}

// The per-pixel geometry below is CALL-FREE on the device: no libdevice sin / asin / sqrt / rsqrt / division,
// whose out-of-line slow paths put CALLs into the painter's pixel loop -- after every call site ptxas
// rematerialises the loop's fp64 constants (265 UMOVs in the phase-2 loop of paint_kernel<P_TABLE>, SASS in
// profiles/).  Accuracy stays at a few ulp (tests/test_hostcheck.py runs the same code on the host).

// sin(x) for |x| <= pi/4: Taylor series to x^17 (next term (pi/4)^19 / 19! = 8e-20)
AP_HD double sin_quarter(double x) {
  const double x2 = x * x;
  double p = 1.0 / 355687428096000.0;
  p = p * x2 - 1.0 / 1307674368000.0;
  p = p * x2 + 1.0 / 6227020800.0;
  p = p * x2 - 1.0 / 39916800.0;
  p = p * x2 + 1.0 / 362880.0;
  p = p * x2 - 1.0 / 5040.0;
  p = p * x2 + 1.0 / 120.0;
  p = p * x2 - 1.0 / 6.0;
  return x + x * (x2 * p);
}

// sin^2(d / 2) for |d / 2| < 1/8: 5-term series (the branch of sin2_half every disc that is not polar or sky-sized
// takes; the painter's straight-line pixel loop calls it directly when a whole chunk of ring items is known to qualify)
AP_HD double sin2_half_small(double d) {
  const double x = 0.5 * d;
  const double x2 = x * x;  // next term x^10 / 39916800 < 2.4e-17 relative
  const double s = x * (1.0 + x2 * (-1.0 / 6.0 + x2 * (1.0 / 120.0 + x2 * (-1.0 / 5040.0 + x2 * (1.0 / 362880.0)))));
  return s * s;
}

// sin^2(d / 2); small arguments (every disc that is not polar or sky-sized) take a 5-term series
AP_HD double sin2_half(double d) {
  double x = 0.5 * d, s;
  if (fabs(x) < 0.125) return sin2_half_small(d);
  // sin^2 has period pi: fold x into [-pi/2, pi/2], then sin^2(x) = 1 - sin^2(pi/2 - |x|) above pi/4
  x = x - kPi * rint(x * (1.0 / kPi));
  const double ax = fabs(x);
  if (ax > 0.25 * kPi) {
    s = sin_quarter(kHalfPi - ax);
    return 1.0 - s * s;
  }
  s = sin_quarter(ax);
  return s * s;
}

AP_HD double hav_of(const RingGeom& g, int32_t j) { return g.A + g.B * sin2_half(g.d0 + (double)j * g.phi_step); }

// asin(x) for 0 <= x <= 1/2: x + x R(x^2), the rational approximation of fdlibm's e_asin.c (|error| < 2^-58)
AP_HD double asin_half(double x) {
  const double t = x * x;
  const double p = t * (1.66666666666666657415e-01 + t * (-3.25565818622400915405e-01 + t * (2.01212532134862925881e-01 +
                   t * (-4.00555345006794114027e-02 + t * (7.91534994289814532176e-04 + t * 3.47933107596021167570e-05)))));
  const double q = 1.0 + t * (-2.40339491173441421878e+00 + t * (2.02094576023350569471e+00 +
                   t * (-6.88283971605453293030e-01 + t * 7.70381505559019352791e-02)));
  return x + x * (p * fm_rcp(q));
}

// 2 asin(sqrt(u)) for 0 < u <= 1/2 (normal u): direct below u = 1/4, else through
// asin(s) = pi/2 - 2 asin(sqrt((1 - s) / 2)) so that asin_half always sees an argument <= 1/2
AP_HD double two_asin_sqrt(double u) {
  const double s = u * fm_rsqrt(u);
  if (u <= 0.25) return 2.0 * asin_half(s);
  const double w = 0.5 * (1.0 - s);
  return kPi - 4.0 * asin_half(w * fm_rsqrt(w));
}

// separation [rad] for hav < 1e-4 (theta < 0.02 rad), branch-free: the series branch of sep_from_hav with its
// zero guard as a select
AP_HD double sep_from_hav_small(double hav) {
  const bool pos = hav > 1.0e-290;       // hav is 0 or >= ~1e-35 by construction; keeps fm_rsqrt on normal numbers
  const double sq = hav * fm_rsqrt(pos ? hav : 1.0);
  // asin(s)/s = 1 + h/6 + 3h^2/40 + 15h^3/336 + 105h^4/3456 + ..., h = s^2 (next term < 3e-22)
  const double v = 2.0 * sq * (1.0 + hav * (1.0 / 6.0 + hav * (3.0 / 40.0 + hav * (15.0 / 336.0 + hav * (105.0 / 3456.0)))));
  return pos ? v : 0.0;
}

// separation [rad] from its haversine: theta = 2 asin(sqrt(hav)); series below hav = 1e-4
AP_HD double sep_from_hav(double hav) {
  if (hav < 1.0e-4) return sep_from_hav_small(hav);
  if (hav <= 0.5) return two_asin_sqrt(hav);
  double co = 1.0 - hav;
  if (!(co > 1.0e-290)) return kPi;
  return kPi - two_asin_sqrt(co);
}

// angular separation [rad] between span pixel j of the ring and the halo centre
AP_HD double pixel_sep(const RingGeom& g, int32_t j) { return sep_from_hav(hav_of(g, j)); }

// chord vector D_a * (p - c)
AP_HD void pixel_chord(const RingGeom& g, int32_t j, const HaloCenter& c, double& rx, double& ry, double& rz) {
  double sp, cp, phi = g.phi0 + (double)j * g.phi_step;
#if defined(__CUDA_ARCH__)
  sincos(phi, &sp, &cp);
#else
  sp = sin(phi);
  cp = cos(phi);
#endif
  rx = c.D_a * (g.sth * cp - c.cx);
  ry = c.D_a * (g.sth * sp - c.cy);
  rz = c.D_a * (g.z - c.cz);
}

// min / max haversine over the len pixels of one span.  sin^2(d/2) is monotone in the azimuthal
// distance |d mod 2pi|, so only the span ends and the pixels next to d = 0, 2pi (minimum) and
// d = pi (maximum) can be extremal; they are ranked by |d| and only the two winners are evaluated.
// Feeds R.min() / R.max() of @interpolate's sample_array (lib/utils.py:411-412).
AP_HD void span_hav_range(const RingGeom& g, int32_t len, double& hmin, double& hmax) {
  int32_t jmin = 0, jmax = 0;
  double wmin = INFINITY, wmax = -INFINITY;
  const double inv_step = 1.0 / g.phi_step;
  for (int c = 0; c < 8; ++c) {
    int32_t j;
    if (c == 0) j = 0;
    else if (c == 1) j = len - 1;
    else {
      double target = (c < 4) ? 0.0 : (c < 6 ? kTwoPi : kPi);
      j = (int32_t)floor((target - g.d0) * inv_step) + (c & 1);
      if (j < 0) j = 0;
      if (j > len - 1) j = len - 1;
    }
    double d = g.d0 + (double)j * g.phi_step;
    double w = fabs(d - kTwoPi * rint(d * kInvTwoPi));
    if (w < wmin) { wmin = w; jmin = j; }
    if (w > wmax) { wmax = w; jmax = j; }
  }
  hmin = hav_of(g, jmin);
  hmax = hav_of(g, jmax);
}

}  // namespace ap
