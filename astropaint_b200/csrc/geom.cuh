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

namespace ap {

struct HaloCenter {
  double theta, phi, sin_t, D_a;
  double cx, cy, cz;
};

AP_HD HaloCenter make_center(double theta, double phi, double D_a, bool need_vec) {
  HaloCenter c;
  c.theta = theta;
  c.phi = phi;
  c.D_a = D_a;
  c.sin_t = sin(theta);
  c.cx = c.cy = c.cz = 0.0;
  if (need_vec) {  // hp.ang2vec(theta, phi), :1212
    c.cx = c.sin_t * cos(phi);
    c.cy = c.sin_t * sin(phi);
    c.cz = cos(theta);
  }
  return c;
}

struct RingGeom {
  double A, B;           // hav = A + B sin^2(dphi/2)
  double phi_step, phi_off;  // phi_p = (ip + phi_off) * phi_step
  double z, sth;
};

AP_HD RingGeom ring_geom(const Hpx& h, int64_t iz, const RingSpan& s, const HaloCenter& c) {
  RingGeom g;
  double theta_p;
  ring_theta(h, iz, s.z, theta_p, g.sth);
  g.z = s.z;
  double sh = sin(0.5 * (theta_p - c.theta));
  g.A = sh * sh;
  g.B = g.sth * c.sin_t;
  g.phi_step = kTwoPi / (double)s.nr;
  g.phi_off = s.shifted ? 0.5 : 0.0;
  return g;
}

AP_HD double pixel_phi(const RingGeom& g, int32_t ip_raw) { return ((double)ip_raw + g.phi_off) * g.phi_step; }

// angular separation [rad] between pixel ip_raw of the ring and the halo centre
AP_HD double pixel_sep(const RingGeom& g, int32_t ip_raw, double phi_c) {
  double S = sin(0.5 * (pixel_phi(g, ip_raw) - phi_c));
  double hav = g.A + g.B * (S * S);
  if (hav < 0.5) return 2.0 * asin(sqrt(hav));
  double co = 1.0 - hav;
  if (co < 0.0) co = 0.0;
  return kPi - 2.0 * asin(sqrt(co));
}

// chord vector D_a * (p - c)
AP_HD void pixel_chord(const RingGeom& g, int32_t ip_raw, const HaloCenter& c, double& rx, double& ry,
                       double& rz) {
  double sp, cp;
#if defined(__CUDA_ARCH__)
  sincos(pixel_phi(g, ip_raw), &sp, &cp);
#else
  sp = sin(pixel_phi(g, ip_raw));
  cp = cos(pixel_phi(g, ip_raw));
#endif
  rx = c.D_a * (g.sth * cp - c.cx);
  ry = c.D_a * (g.sth * sp - c.cy);
  rz = c.D_a * (g.z - c.cz);
}

// min / max separation over the pixels of one span (sin^2(dphi/2) is monotone in the azimuthal
// distance, so only the span ends, the pixels next to phi_c and the pixels next to phi_c + pi
// can be extremal).  Feeds R.min() / R.max() of @interpolate's sample_array (utils.py:411-412).
AP_HD void span_sep_range(const RingGeom& g, const RingSpan& s, double phi_c, double& smin, double& smax) {
  int32_t lo = s.ip_lo, hi = s.ip_lo + s.len - 1;
  double a = pixel_sep(g, lo, phi_c), b = pixel_sep(g, hi, phi_c);
  smin = fmin(a, b);
  smax = fmax(a, b);
  double t = phi_c / g.phi_step - g.phi_off;
  for (int pass = 0; pass < 2; ++pass) {
    int32_t k0 = (int32_t)floor(t + (pass ? 0.5 * (double)s.nr : 0.0));
    for (int d = 0; d < 2; ++d) {
      int32_t k = k0 + d;
      while (k > hi) k -= s.nr;
      while (k < lo) k += s.nr;
      if (k > hi) continue;
      double v = pixel_sep(g, k, phi_c);
      smin = fmin(smin, v);
      smax = fmax(smax, v);
    }
  }
}

}  // namespace ap
