// TEST-ONLY host build of the device arithmetic headers (g++ -ffp-contract=off).
//
// libap_hostcheck.so lets the CPU test-suite (-m "not gpu") validate healpix.cuh, geom.cuh,
// qagi.cuh, profiles.cuh and spline.cuh against the oracle and scipy in a container with no
// GPU.  It is never loaded by the product path (astropaint_b200/_capi.py loads only
// libastropaint_b200.so and raises if it is missing).
#include <cstdint>
#include <cmath>

#include "common.cuh"
#include "healpix.cuh"
#include "geom.cuh"
#include "profiles.cuh"
#include "spline.cuh"

using namespace ap;

template <int P>
static void eval_R(const double* p, int64_t n, const double* R, double* out) {
  double ctx[kCtx];
  profile_setup<P>(p, ctx);
  for (int64_t i = 0; i < n; ++i) out[i] = profile_eval_R<P>(ctx, R[i]);
}

extern "C" {

int64_t hc_query_disc(int64_t nside, double x, double y, double z, double radius, int64_t* out, int64_t cap) {
  Hpx h = make_hpx(nside);
  DiscHeader d = disc_header(h, x, y, z, radius);
  int64_t n = 0;
  for (int64_t iz = d.ir_first; iz <= d.ir_last; ++iz) {
    RingSpan s = ring_span(h, d, iz);
    for (int32_t j = 0; j < s.len; ++j) {
      if (out && n < cap) out[n] = span_pixel_sorted(s, j);
      ++n;
    }
  }
  return n;
}

// healpy.query_disc(inclusive=True, fact=fact)
int64_t hc_query_disc_inclusive(int64_t nside, double x, double y, double z, double radius, int fact, int64_t* out,
                                int64_t cap) {
  Hpx h = make_hpx(nside);
  hpx_set_inclusive(h, fact);
  DiscHeader d = disc_header(h, x, y, z, radius);
  int64_t n = 0;
  for (int64_t iz = d.ir_first; iz <= d.ir_last; ++iz) {
    RingSpan s = ring_span(h, d, iz);
    for (int32_t j = 0; j < s.len; ++j) {
      if (out && n < cap) out[n] = span_pixel_sorted(s, j);
      ++n;
    }
  }
  return n;
}

double hc_max_pixrad(int64_t nside) { return max_pixrad(nside); }

// the call-free per-pixel geometry of geom.cuh: sin^2(d/2) and the separation 2 asin(sqrt(hav))
void hc_geom_funcs(int64_t n, const double* d, const double* hav, double* s2, double* sep) {
  for (int64_t i = 0; i < n; ++i) {
    if (s2) s2[i] = sin2_half(d[i]);
    if (sep) sep[i] = sep_from_hav(hav[i]);
  }
}

// per-halo pixel count and sum of pixel ids for n halos (near-tie scans against the device)
void hc_query_disc_stats(int64_t nside, int64_t n, const double* x, const double* y, const double* z,
                         const double* radius, int64_t* count, int64_t* pixsum) {
  Hpx h = make_hpx(nside);
  for (int64_t i = 0; i < n; ++i) {
    DiscHeader d = disc_header(h, x[i], y[i], z[i], radius[i]);
    int64_t c = 0, s = 0;
    for (int64_t iz = d.ir_first; iz <= d.ir_last; ++iz) {
      RingSpan sp = ring_span(h, d, iz);
      for (int32_t j = 0; j < sp.len; ++j) s += span_pixel_sorted(sp, j);
      c += sp.len;
    }
    count[i] = c;
    pixsum[i] = s;
  }
}

// R (and R_vec) of every disc pixel, in ascending pixel order; also the span-based R range
int64_t hc_disc_R(int64_t nside, double x, double y, double z, double radius, double theta, double phi, double D_a,
                  double* R, double* Rvec, int64_t cap, double* rmin, double* rmax) {
  Hpx h = make_hpx(nside);
  DiscHeader d = disc_header(h, x, y, z, radius);
  HaloCenter c = make_center(theta, phi, D_a, true);
  int64_t n = 0;
  double hmin = INFINITY, hmax = -INFINITY;
  for (int64_t iz = d.ir_first; iz <= d.ir_last; ++iz) {
    RingSpan s = ring_span(h, d, iz);
    if (s.len <= 0) continue;
    RingGeom g = ring_geom(h, iz, s, c);
    double a, b;
    span_hav_range(g, s.len, a, b);
    hmin = fmin(hmin, a);
    hmax = fmax(hmax, b);
    for (int32_t jj = 0; jj < s.len; ++jj) {
      int64_t p = span_pixel_sorted(s, jj);
      int32_t j = span_j_of_ip(s, (int32_t)(p - s.startpix));
      if (n < cap) {
        if (R) R[n] = c.D_a * pixel_sep(g, j);
        if (Rvec) pixel_chord(g, j, c, Rvec[3 * n], Rvec[3 * n + 1], Rvec[3 * n + 2]);
      }
      ++n;
    }
  }
  if (rmin) *rmin = D_a * sep_from_hav(hmin);
  if (rmax) *rmax = D_a * sep_from_hav(hmax);
  return n;
}

void hc_ang2pix(int64_t nside, int64_t n, const double* theta, const double* phi, int64_t* out) {
  Hpx h = make_hpx(nside);
  for (int64_t i = 0; i < n; ++i) out[i] = ang2pix_ring(h, theta[i], phi[i]);
}

void hc_pix2ang(int64_t nside, int64_t n, const int64_t* pix, double* theta, double* phi) {
  Hpx h = make_hpx(nside);
  for (int64_t i = 0; i < n; ++i) pix2ang_ring(h, pix[i], theta[i], phi[i]);
}

void hc_vec2pix(int64_t nside, int64_t n, const double* x, const double* y, const double* z, int64_t* out) {
  Hpx h = make_hpx(nside);
  for (int64_t i = 0; i < n; ++i) out[i] = vec2pix_ring(h, x[i], y[i], z[i]);
}

void hc_vec2pix_near(int64_t nside, int64_t n, const double* x, const double* y, const double* z, double phi_c,
                     int64_t* out) {
  Hpx h = make_hpx(nside);
  double c = cos(phi_c), s = sin(phi_c);
  for (int64_t i = 0; i < n; ++i) out[i] = vec2pix_ring_near(h, x[i], y[i], z[i], phi_c, c, s);
}

void hc_qagi_b16(double R, double R_200c, double M_200c, double redshift, double* result, double* abserr, int* neval,
                 int* ier) {
  double ctx[kCtx];
  b16_setup(R_200c, M_200c, redshift, ctx);
  QagiResult q;
  b16_tau_2D(ctx, R, &q);
  *result = q.result;
  *abserr = q.abserr;
  *neval = q.neval;
  *ier = q.ier;
}

// generic LOS projection with QUADPACK diagnostics (kind = LosKind)
int hc_los(int kind, double R, double p0, double p1, double p2, double* result, int* neval) {
  double ctx[kCtx];
  QagiResult q;
  switch (kind) {
    case LOS_B16_TAU: los_setup<LOS_B16_TAU>(p0, p1, p2, ctx); los_eval<LOS_B16_TAU>(ctx, R, &q); break;
    case LOS_B16_NE: los_setup<LOS_B16_NE>(p0, p1, p2, ctx); los_eval<LOS_B16_NE>(ctx, R, &q); break;
    case LOS_B16_RHO_GAS: los_setup<LOS_B16_RHO_GAS>(p0, p1, p2, ctx); los_eval<LOS_B16_RHO_GAS>(ctx, R, &q); break;
    case LOS_NFW_RHO: los_setup<LOS_NFW_RHO>(p0, p1, p2, ctx); los_eval<LOS_NFW_RHO>(ctx, R, &q); break;
    default: return 1;
  }
  *result = q.result;
  *neval = q.neval;
  return 0;
}

// the rightmost-bisection fast path alone: returns 1 if it accepted the quadrature (result/abserr/neval/ier
// filled), 0 if it bailed out to the general routine; kind as hc_los
int hc_los_fast_path(int kind, double R, double p0, double p1, double p2, double* result, double* abserr, int* neval,
                     int* ier) {
  double ctx[kCtx];
  QagiResult q;
  bool ok;
  if (kind == LOS_NFW_RHO) {
    los_setup<LOS_NFW_RHO>(p0, p1, p2, ctx);
    NfwRhoIntegrand f;
    f.K8 = ctx[0]; f.rs = ctx[1]; f.RR = R * R;
    ok = qagi_upper_rm(f, R, 0.0, 1.0e-2, q);
  } else {
    b16_setup(p0, p1, p2, ctx);
    B16Integrand f;
    f.inv_R200xc = ctx[0]; f.alpha = ctx[1]; f.expo = ctx[2]; f.RR = R * R;
    ok = qagi_upper_rm(f, R, 0.0, 1.0e-2, q);
    if (ok) { q.result *= 2.0 * ctx[3]; q.abserr *= 2.0 * ctx[3]; }
  }
  if (!ok) return 0;
  *result = q.result; *abserr = q.abserr; *neval = q.neval; *ier = q.ier;
  return 1;
}

double hc_b16_tau_3D(double r, double R_200c, double M_200c, double redshift) {
  double ctx[kCtx];
  b16_setup(R_200c, M_200c, redshift, ctx);
  B16Integrand f;
  f.inv_R200xc = ctx[0]; f.alpha = ctx[1]; f.expo = ctx[2]; f.RR = 0.0;
  return f(r) * ctx[3];  // with R = 0 the LOS weight r/sqrt(r^2) is 1; ctx[3] = the profile's normalisation
}

double hc_critical_density(double z) { return critical_density(z); }

void hc_fast_log_exp(int64_t n, const double* x, double* lg, double* ex) {
  for (int64_t i = 0; i < n; ++i) {
    if (lg) lg[i] = fast_log(x[i]);
    if (ex) ex[i] = fast_exp(x[i]);
  }
}

int hc_profile_R(int profile, const double* p, int64_t n, const double* R, double* out) {
  switch (profile) {
    case P_CONSTANT: eval_R<P_CONSTANT>(p, n, R, out); return 0;
    case P_LINEAR: eval_R<P_LINEAR>(p, n, R, out); return 0;
    case P_SOLID_SPHERE: eval_R<P_SOLID_SPHERE>(p, n, R, out); return 0;
    case P_GAUSSIAN: eval_R<P_GAUSSIAN>(p, n, R, out); return 0;
    case P_NFW_RHO2D: eval_R<P_NFW_RHO2D>(p, n, R, out); return 0;
    case P_NFW_TAU2D: eval_R<P_NFW_TAU2D>(p, n, R, out); return 0;
    case P_NFW_KSZ: eval_R<P_NFW_KSZ>(p, n, R, out); return 0;
    case P_NFW_DEFLECTION: eval_R<P_NFW_DEFLECTION>(p, n, R, out); return 0;
    case P_B16_TAU2D: eval_R<P_B16_TAU2D>(p, n, R, out); return 0;
    case P_B16_RHOGAS2D: eval_R<P_B16_RHOGAS2D>(p, n, R, out); return 0;
    case P_NFW_RHO2D_LOS: eval_R<P_NFW_RHO2D_LOS>(p, n, R, out); return 0;
  }
  return 1;
}

int hc_profile_vec(int profile, const double* p, int64_t n, const double* Rvec, double* out) {
  if (profile != P_NFW_BG) return 1;
  double ctx[kCtx];
  profile_setup<P_NFW_BG>(p, ctx);
  for (int64_t i = 0; i < n; ++i)
    out[i] = profile_eval_vec<P_NFW_BG>(ctx, Rvec[3 * i], Rvec[3 * i + 1], Rvec[3 * i + 2]);
  return 0;
}

int hc_spline(int n, const double* x, const double* y, int64_t m, const double* X, double* out) {
  if (n < 4 || n > kMaxNodes) return 1;
  double c[(kMaxNodes - 1) * 4];
  spline_build_nak(n, x, y, c);
  for (int64_t i = 0; i < m; ++i) out[i] = spline_eval(n, x, c, X[i]);
  return 0;
}

}  // extern "C"
