// astropaint_b200 — sm_100a kernels and the C ABI (include/astropaint_b200.h): translation unit 1 of 4
// (k_core.cu: modes, catalog, flat lists, cutouts, host plumbing; k_paint.cu: K1 / K3; k_paint_los.cu: K3 for
// the per-pixel quadrature templates; k_los.cu: K4 tables, QUADPACK nodes, bypass).
//
// Kernel map (DESIGN.md section 3 has the roofline of each):
//   disc_radius_kernel            R_times * arcmin2rad(R_ang_200c) with numpy's rounding
//   paint_kernel<P, MapT>         K3  fused disc query -> pixel geometry -> template -> RED scatter
//     P = built-in profile id | P_TABLE (per-halo spline) | P_STATS (K1: pixel count + R range)
//   disc_pixels_kernel            K1b flat (halo, pixel) list (+R, +R_vec): parity, Python templates
//   scatter_add_kernel            np.add.at for the flat list
//   table_nodes_kernel, los_nodes_kernel<KIND>, spline_build(_uniform)_kernel   K4 @interpolate tables
//   paint_los_items_kernel<KIND>  the len(R) < n_samples bypass of @interpolate (device-built work list)
//   cutout_kernel<STACK>          K5  CartesianProj cutouts and their stack
//   patch_ops.cuh                 K6  apply_func library: spline prefilter + rotate, taper, weight, stack
#include "internal.cuh"

static unsigned stream_grid_n(int64_t n) {  // grid-stride launch: at most 148 SMs x 8 resident blocks of 256
  int64_t b = (n + 255) / 256;
  return (unsigned)(b < 148 * 8 ? (b < 1 ? 1 : b) : 148 * 8);
}

// ------------------------------------------------------------------------------------------
// error plumbing
// ------------------------------------------------------------------------------------------
static thread_local std::string g_err;
int fail(const std::string& m) {
  g_err = m;
  return 1;
}
// Disc-query mode of the calling thread: 0 = hp.query_disc(inclusive=False) (Canvas default,
// paint_bucket.py:782), f > 0 = inclusive=True with oversampling factor f (healpy's default f is 4).
// Every entry point that runs a disc query (count, pixels, paint, LOS bypass) honours it.
static thread_local int g_incl_fact = 0;
extern "C" int ap_disc_mode(int inclusive_fact) {
  int prev = g_incl_fact;
  if (inclusive_fact >= 0) g_incl_fact = inclusive_fact;  // negative: query only
  return prev;
}
// 1 (default): ap_los_nodes uses the rightmost-bisection fast path of dqagie (qagi.cuh) and falls back to the
// general routine per quadrature; 0: the general routine only (parity tests compare the two bit for bit)
static int g_quad_fast = 1;
int quad_fast_enabled() { return g_quad_fast; }
extern "C" int ap_quad_fast_path(int enable) {
  int prev = g_quad_fast;
  if (enable >= 0) g_quad_fast = enable ? 1 : 0;
  return prev;
}
// Ring band painted by the calling thread's next launches (0, 0 = the whole map): rings [lo, hi) only, and the
// map pointer of every painting entry point addresses the first pixel of ring lo.  One rank of a multi-GPU
// spray owns one band of the map; halos whose disc straddles a boundary are painted by
// both neighbours, each keeping its own rings, so no partial map and no reduction exist.
static thread_local int64_t g_clip_lo = 0, g_clip_hi = 0;
extern "C" int ap_band_clip(int64_t ring_lo, int64_t ring_hi) {
  if (ring_lo <= 0 && ring_hi <= 0) {
    g_clip_lo = g_clip_hi = 0;
    return 0;
  }
  if (ring_lo < 1 || ring_hi <= ring_lo) return fail("band_clip: need 1 <= ring_lo < ring_hi");
  g_clip_lo = ring_lo;
  g_clip_hi = ring_hi;
  return 0;
}
// Deterministic accumulation of the calling thread's next painting launches (ap_emit_begin .. ap_emit_end)
static thread_local EmitArgs g_emit = {nullptr, nullptr, nullptr, 0, nullptr, 0};
EmitArgs emit_args() { return g_emit; }
extern "C" int ap_emit_begin(uint64_t* d_keys, double* d_vals, unsigned long long* d_count, int64_t cap,
                             const int64_t* d_labels, int32_t shift) {
  if (!d_keys || !d_vals || !d_count || cap < 0) return fail("emit_begin: NULL buffer");
  if (shift < 1 || shift > 40) return fail("emit_begin: shift (bits kept for the halo label) must be in [1, 40]");
  g_emit.keys = reinterpret_cast<unsigned long long*>(d_keys);
  g_emit.vals = d_vals;
  g_emit.count = d_count;
  g_emit.cap = cap;
  g_emit.labels = reinterpret_cast<const long long*>(d_labels);
  g_emit.shift = shift;
  return 0;
}
extern "C" int ap_emit_end(void) {
  g_emit = EmitArgs{nullptr, nullptr, nullptr, 0, nullptr, 0};
  return 0;
}
Hpx disc_hpx(int64_t nside) {
  Hpx h = make_hpx(nside);
  hpx_set_inclusive(h, g_incl_fact);
  if (g_clip_hi > 0) {
    const int64_t top = 4 * nside;
    h.clip_lo = (int32_t)(g_clip_lo < top ? g_clip_lo : top);
    h.clip_hi = (int32_t)(g_clip_hi < top ? g_clip_hi : top);
    int64_t nr;
    bool shifted;
    if (h.clip_lo < top) ring_info(h, h.clip_lo, h.clip_pix0, nr, shifted);
    else h.clip_pix0 = h.npix;
    if (h.clip_hi < top) ring_info(h, h.clip_hi, h.clip_pix1, nr, shifted);
    else h.clip_pix1 = h.npix;
  }
  return h;
}AP_TU_BOUNDS_GETTER(ap_bounds_core)
extern "C" long long ap_debug_bounds_errors(void) {
#ifdef AP_DEBUG_BOUNDS
  long long v[4] = {ap_bounds_core(), ap_bounds_paint(), ap_bounds_paint_los(), ap_bounds_los()};
  long long s = 0;
  for (long long x : v) {
    if (x < 0) return -2;
    s += x;
  }
  return s;
#else
  return -1;  // not an instrumented build
#endif
}

extern "C" const char* ap_last_error(void) { return g_err.c_str(); }
extern "C" int ap_version(void) { return 200; }

// ------------------------------------------------------------------------------------------
// disc radius: R_times * np.deg2rad(R_ang_200c / 60), operation for operation
// (Canvas.Disc.gen_pixel_index, paint_bucket.py:1226-1227; transform.arcmin2rad, lib/transform.py:148-150)
// ------------------------------------------------------------------------------------------
__global__ void disc_radius_kernel(int64_t n, const double* __restrict__ r_ang_arcmin, double r_times,
                                   double* __restrict__ radius) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const double deg = __ddiv_rn(r_ang_arcmin[i], 60.0);               // angle / 60
  const double rad = __dmul_rn(deg, 3.14159265358979323846 / 180.0);  // np.deg2rad: x * (pi / 180)
  radius[i] = __dmul_rn(r_times, rad);
}

extern "C" int ap_disc_radius(int64_t n, const double* d_R_ang_200c, double R_times, double* d_radius, void* stream) {
  if (n < 0) return fail("n < 0");
  if (n == 0) return 0;
  if (!d_R_ang_200c || !d_radius) return fail("disc_radius: NULL pointer");
  disc_radius_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(n, d_R_ang_200c, R_times, d_radius);
  AP_LAUNCH_CHECK();
  return 0;
}

// ------------------------------------------------------------------------------------------
// K0: Catalog.build_dataframe (paint_bucket.py:308-364) -- the 17 derived columns of the halo catalog,
// one thread per halo, structure-of-arrays in and out (coalesced).  HBM-bound: 64-72 B read + 136 B written.
// ------------------------------------------------------------------------------------------
struct CatalogIn {
  const double *x, *y, *z, *vx, *vy, *vz, *M, *redshift;  // redshift may be NULL -> default_redshift
};
struct CatalogOut {
  double* col[AP_CATALOG_N_OUT];
};

__global__ void __launch_bounds__(256) catalog_build_kernel(int64_t n, CatalogIn in, double default_redshift,
                                                            double crit_dens_0, double hubble_h, CatalogOut out) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const double x = in.x[i], y = in.y[i], z = in.z[i];
  // astropy cartesian_to_spherical: r, lat in [-pi/2, pi/2], lon wrapped into [0, 2 pi)
  const double s = hypot(x, y);
  const double D_c = hypot(s, z);
  const double lat = atan2(z, s);
  double lon = atan2(y, x);
  if (lon < 0.0) lon = AP_ADD(lon, kTwoPi);
  const double zr = in.redshift ? in.redshift[i] : default_redshift;
  const double theta = AP_SUB(kHalfPi, lat), phi = lon;
  const double zp1 = AP_ADD(1.0, zr);
  const double D_a = D_c / zp1;                                            // transform.D_c_to_D_a
  const double M = in.M[i];
  // transform.M_200c_to_R_200c: (3 M / (800 pi rho_c0 (1+z)^3))^(1/3)
  const double rho = AP_MUL(crit_dens_0, pow(zp1, 3.0));
  const double R200 = pow(AP_MUL(3.0, M) / AP_MUL(800.0 * kPi, rho), 1.0 / 3.0);
  // transform.M_200c_to_c_200c (Coe 2010): A (1+z)^d (M / M_0)^m
  const double c200 = AP_MUL(AP_MUL(5.71, pow(zp1, -0.047)), pow(M / (2.0e12 / hubble_h), -0.097));
  const double R_ang = AP_MUL(AP_MUL(R200 / D_a, 180.0 / kPi), 60.0);      // rad2arcmin(R_200c / D_a)
  const double R_s = R200 / c200;
  const double A_c = AP_SUB(log(AP_ADD(1.0, c200)), c200 / AP_ADD(1.0, c200));
  const double rho_s = M / AP_MUL(AP_MUL(16.0 * kPi, pow(R_s, 3.0)), A_c);   // transform.M_200c_to_rho_s
  // velocities: einsum('ij...,i...->j...', get_cart2sph_jacobian(theta, phi), v_cart)
  double st, ct, sp, cp;
  sincos(theta, &st, &ct);
  sincos(phi, &sp, &cp);
  const double vx = in.vx[i], vy = in.vy[i], vz = in.vz[i];
  const double v_r = AP_ADD(AP_ADD(AP_MUL(AP_MUL(st, cp), vx), AP_MUL(AP_MUL(st, sp), vy)), AP_MUL(ct, vz));
  const double v_th = AP_ADD(AP_ADD(AP_MUL(AP_MUL(ct, cp), vx), AP_MUL(AP_MUL(ct, sp), vy)), AP_MUL(-st, vz));
  const double v_ph = AP_ADD(AP_ADD(AP_MUL(-sp, vx), AP_MUL(cp, vy)), AP_MUL(AP_MUL(0.0, ct), vz));
  double* const* o = out.col;
  o[AP_CAT_D_C][i] = D_c;
  o[AP_CAT_REDSHIFT][i] = zr;
  o[AP_CAT_LAT][i] = AP_MUL(lat, 180.0 / kPi);
  o[AP_CAT_LON][i] = AP_MUL(lon, 180.0 / kPi);
  o[AP_CAT_THETA][i] = theta;
  o[AP_CAT_PHI][i] = phi;
  o[AP_CAT_D_A][i] = D_a;
  o[AP_CAT_R_200C][i] = R200;
  o[AP_CAT_C_200C][i] = c200;
  o[AP_CAT_R_ANG_200C][i] = R_ang;
  o[AP_CAT_RHO_S][i] = rho_s;
  o[AP_CAT_R_S][i] = R_s;
  o[AP_CAT_V_R][i] = v_r;
  o[AP_CAT_V_TH][i] = v_th;
  o[AP_CAT_V_PH][i] = v_ph;
  o[AP_CAT_V_LAT][i] = -v_th;
  o[AP_CAT_V_LON][i] = v_ph;
}

extern "C" int ap_catalog_build(int64_t n, const double* d_x, const double* d_y, const double* d_z, const double* d_vx,
                                const double* d_vy, const double* d_vz, const double* d_M_200c,
                                const double* d_redshift, double default_redshift, double crit_dens_0, double h,
                                double* const* h_out, void* stream) {
  if (n < 0) return fail("n < 0");
  if (n == 0) return 0;
  if (!d_x || !d_y || !d_z || !d_vx || !d_vy || !d_vz || !d_M_200c || !h_out) return fail("catalog_build: NULL pointer");
  if (!(crit_dens_0 > 0.0) || !(h > 0.0)) return fail("catalog_build: cosmology constants must be positive");
  CatalogIn in{d_x, d_y, d_z, d_vx, d_vy, d_vz, d_M_200c, d_redshift};
  CatalogOut out;
  for (int k = 0; k < AP_CATALOG_N_OUT; ++k) {
    if (!h_out[k]) return fail("catalog_build: NULL output column");
    out.col[k] = h_out[k];
  }
  catalog_build_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(n, in, default_redshift, crit_dens_0, h,
                                                                                    out);
  AP_LAUNCH_CHECK();
  return 0;
}

// ------------------------------------------------------------------------------------------
// K1b: flat pixel list (+R, +R_vec); one warp per halo
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) disc_pixels_kernel(Hpx hpx, HalosDev H, const int64_t* pix_off, int64_t* pix,
                                                          double* R, double* Rvec) {
  int64_t h = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  int lane = threadIdx.x & 31;
  if (h >= H.n) return;
  DiscHeader d = disc_header(hpx, H.x[h], H.y[h], H.z[h], H.radius[h]);
  bool want = (R != nullptr) || (Rvec != nullptr);
  HaloCenter c;
  if (want) c = make_center(H.theta[h], H.phi[h], H.D_a[h], Rvec != nullptr);
  int64_t run = pix_off[h];
  for (int64_t iz = d.ir_first; iz <= d.ir_last; ++iz) {
    RingSpan s = ring_span(hpx, d, iz);
    if (s.len <= 0) continue;
    RingGeom g;
    if (want) g = ring_geom(hpx, iz, s, c);
    for (int32_t j = lane; j < s.len; j += 32) {
      int64_t p = span_pixel_sorted(s, j);
      int64_t o = run + j;
      if (!AP_BOUNDS_OK(p >= 0 && p < hpx.npix && o < pix_off[h + 1])) continue;
      pix[o] = p;
      if (want) {
        int32_t js = span_j_of_ip(s, (int32_t)(p - s.startpix));
        if (R) R[o] = c.D_a * pixel_sep(g, js);
        if (Rvec) {
          double rx, ry, rz;
          pixel_chord(g, js, c, rx, ry, rz);
          Rvec[3 * o + 0] = rx;
          Rvec[3 * o + 1] = ry;
          Rvec[3 * o + 2] = rz;
        }
      }
    }
    run += s.len;
  }
}

extern "C" int ap_disc_pixels(int64_t nside, const ap_halos* halos, const int64_t* d_pix_off, int64_t* d_pix,
                              double* d_R, double* d_R_vec, void* stream) {
  if (check_nside(nside)) return 1;
  if (check_halos(halos, d_R || d_R_vec)) return 1;
  if (halos->n == 0) return 0;
  if (!d_pix_off || !d_pix) return fail("d_pix_off / d_pix NULL");
  int threads = 256;
  int64_t blocks = (halos->n * 32 + threads - 1) / threads;
  disc_pixels_kernel<<<(unsigned)blocks, threads, 0, (cudaStream_t)stream>>>(disc_hpx(nside), to_dev(halos), d_pix_off,
                                                                             d_pix, d_R, d_R_vec);
  AP_LAUNCH_CHECK();
  return 0;
}

// ------------------------------------------------------------------------------------------
// Deterministic accumulation, second half: the (pixel << shift | halo) keys arrive SORTED; the thread that finds the
// first entry of a pixel adds that pixel's values to the map one after the other, in halo order -- the order the
// reference's serial loop adds them in (np.add.at over halos 0, 1, 2, ...; paint_bucket.py:2380-2389).  Plain
// loads and stores: every pixel has exactly one writer.
// ------------------------------------------------------------------------------------------
template <class MapT>
__global__ void __launch_bounds__(256) segment_accumulate_kernel(int64_t n, const unsigned long long* __restrict__ keys,
                                                                 const double* __restrict__ vals, int shift, int64_t pix0,
                                                                 MapT* __restrict__ map) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const unsigned long long pix = keys[i] >> shift;
    if (i > 0 && (keys[i - 1] >> shift) == pix) continue;   // not the first entry of its pixel
    MapT acc = map[(int64_t)pix - pix0];
    for (int64_t j = i; j < n && (keys[j] >> shift) == pix; ++j) acc += (MapT)vals[j];
    map[(int64_t)pix - pix0] = acc;
  }
}

extern "C" int ap_segment_accumulate(int64_t n, const uint64_t* d_keys_sorted, const double* d_vals_sorted, int32_t shift,
                                     int64_t pix0, void* d_map, int f32, void* stream) {
  if (n < 0 || shift < 1 || shift > 40) return fail("segment_accumulate: bad arguments");
  if (n == 0) return 0;
  if (!d_keys_sorted || !d_vals_sorted || !d_map) return fail("segment_accumulate: NULL pointer");
  const unsigned long long* k = reinterpret_cast<const unsigned long long*>(d_keys_sorted);
  if (f32)
    segment_accumulate_kernel<float><<<stream_grid_n(n), 256, 0, (cudaStream_t)stream>>>(n, k, d_vals_sorted, shift, pix0, (float*)d_map);
  else
    segment_accumulate_kernel<double><<<stream_grid_n(n), 256, 0, (cudaStream_t)stream>>>(n, k, d_vals_sorted, shift, pix0, (double*)d_map);
  AP_LAUNCH_CHECK();
  return 0;
}

// ------------------------------------------------------------------------------------------
// Near-tie audit of the disc query: the device lists the halos with a span boundary within `tol`
// pixels of an integer (or a ring-range decision within tol of flipping); the host recomputes their pixel sets with
// its own libm (ap_query_disc_host) and the two are compared (astropaint_b200/_engine.py::disc_tie_audit).
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) disc_near_ties_kernel(Hpx hpx, HalosDev H, double tol, unsigned long long* n_flagged,
                                                             int64_t* flagged, int64_t cap, double* min_margin) {
  int64_t h = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (h >= H.n) return;
  DiscHeader d = disc_header_t<false>(hpx, H.x[h], H.y[h], H.z[h], H.radius[h]);
  double m = INFINITY;
  {  // the ring range itself: ring_above(cos(theta -+ radius))
    const double theta = atan2(sqrt(AP_ADD(AP_MUL(H.x[h], H.x[h]), AP_MUL(H.y[h], H.y[h]))), H.z[h]);
    const double r = H.radius[h];
    if (r < kPi) m = fmin(ring_above_margin(hpx, cos(AP_SUB(theta, r))), ring_above_margin(hpx, cos(AP_ADD(theta, r))));
  }
  for (int64_t iz = d.irmin; iz <= d.irmax; ++iz) m = fmin(m, ring_tie_margin(hpx, d, iz));
  if (min_margin) min_margin[h] = m;
  if (m < tol) {
    unsigned long long at = atomicAdd(n_flagged, 1ull);
    if ((int64_t)at < cap) flagged[at] = h;
  }
}

extern "C" int ap_disc_near_ties(int64_t nside, const ap_halos* halos, double tol, unsigned long long* d_n_flagged,
                                 int64_t* d_flagged, int64_t cap, double* d_min_margin, void* stream) {
  if (check_nside(nside)) return 1;
  if (check_halos(halos, false)) return 1;
  if (!d_n_flagged || (cap > 0 && !d_flagged) || cap < 0) return fail("disc_near_ties: NULL pointer");
  if (halos->n == 0) return 0;
  disc_near_ties_kernel<<<(unsigned)((halos->n + 127) / 128), 128, 0, (cudaStream_t)stream>>>(
      make_hpx(nside), to_dev(halos), tol, d_n_flagged, d_flagged, cap, d_min_margin);
  AP_LAUNCH_CHECK();
  return 0;
}

// hp.query_disc on the HOST (this library's ring arithmetic compiled for the CPU, glibc libm): per disc the pixel
// count, the sum of the pixel ids and, optionally, the sorted pixel list of ONE disc.  All pointers are host pointers.
extern "C" int ap_query_disc_host(int64_t nside, int64_t n, const double* h_x, const double* h_y, const double* h_z,
                                  const double* h_radius, int inclusive_fact, int64_t* h_count, int64_t* h_pixsum,
                                  int64_t* h_pix_first, int64_t pix_cap) {
  if (check_nside(nside)) return 1;
  if (n < 0 || inclusive_fact < 0) return fail("query_disc_host: bad arguments");
  if (n > 0 && (!h_x || !h_y || !h_z || !h_radius)) return fail("query_disc_host: NULL pointer");
  Hpx hp = make_hpx(nside);
  hpx_set_inclusive(hp, inclusive_fact);
  for (int64_t i = 0; i < n; ++i) {
    DiscHeader d = disc_header(hp, h_x[i], h_y[i], h_z[i], h_radius[i]);
    int64_t c = 0, sum = 0;
    for (int64_t iz = d.ir_first; iz <= d.ir_last; ++iz) {
      RingSpan sp = ring_span(hp, d, iz);
      for (int32_t j = 0; j < sp.len; ++j) {
        const int64_t p = span_pixel_sorted(sp, j);
        if (i == 0 && h_pix_first && c + j < pix_cap) h_pix_first[c + j] = p;
        sum += p;
      }
      c += sp.len > 0 ? sp.len : 0;
    }
    if (h_count) h_count[i] = c;
    if (h_pixsum) h_pixsum[i] = sum;
  }
  return 0;
}

// ------------------------------------------------------------------------------------------
// Pixel <-> angle maps of the RING scheme for the centre generators (hp.ang2pix / hp.pix2ang:
// paint_bucket.py:996, 1085, 1183, 1199, 1238).  Streaming kernels: 16 B in + 8 B out per element.
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) ang2pix_kernel(Hpx hpx, int64_t n, const double* __restrict__ theta,
                                                      const double* __restrict__ phi, int64_t* __restrict__ pix) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const double t = theta[i];
    pix[i] = (t >= 0.0 && t <= kPi) ? ang2pix_ring(hpx, t, phi[i]) : -1;  // healpy raises on an invalid theta
  }
}

__global__ void __launch_bounds__(256) pix2ang_kernel(Hpx hpx, int64_t n, const int64_t* __restrict__ pix,
                                                      double* __restrict__ theta, double* __restrict__ phi) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t p = pix[i];
    double t = NAN, f = NAN;
    if (p >= 0 && p < hpx.npix) pix2ang_ring(hpx, p, t, f);
    theta[i] = t;
    phi[i] = f;
  }
}

static unsigned stream_grid(int64_t n) {  // grid-stride launch: at most 148 SMs x 8 resident blocks of 256
  int64_t b = (n + 255) / 256;
  return (unsigned)(b < 148 * 8 ? (b < 1 ? 1 : b) : 148 * 8);
}

extern "C" int ap_ang2pix(int64_t nside, int64_t n, const double* d_theta, const double* d_phi, int64_t* d_pix,
                          void* stream) {
  if (check_nside(nside)) return 1;
  if (n < 0) return fail("ang2pix: n < 0");
  if (n == 0) return 0;
  if (!d_theta || !d_phi || !d_pix) return fail("ang2pix: NULL buffer");
  ang2pix_kernel<<<stream_grid(n), 256, 0, (cudaStream_t)stream>>>(make_hpx(nside), n, d_theta, d_phi, d_pix);
  AP_LAUNCH_CHECK();
  return 0;
}

extern "C" int ap_pix2ang(int64_t nside, int64_t n, const int64_t* d_pix, double* d_theta, double* d_phi,
                          void* stream) {
  if (check_nside(nside)) return 1;
  if (n < 0) return fail("pix2ang: n < 0");
  if (n == 0) return 0;
  if (!d_theta || !d_phi || !d_pix) return fail("pix2ang: NULL buffer");
  pix2ang_kernel<<<stream_grid(n), 256, 0, (cudaStream_t)stream>>>(make_hpx(nside), n, d_pix, d_theta, d_phi);
  AP_LAUNCH_CHECK();
  return 0;
}

// ------------------------------------------------------------------------------------------
// scatter-add of a flat work list
// ------------------------------------------------------------------------------------------
template <class MapT>
__global__ void scatter_add_kernel(int64_t n, const int64_t* pix, const double* val, MapT* map) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (; i < n; i += stride) atomicAdd(&map[pix[i]], (MapT)val[i]);  // indices come from ap_disc_pixels
}

static int scatter_add_impl(int64_t n, const int64_t* d_pix, const double* d_val, void* d_map, bool f32, void* stream) {
  if (n < 0) return fail("n < 0");
  if (n == 0) return 0;
  if (!d_pix || !d_val || !d_map) return fail("scatter_add: NULL pointer");
  int threads = 256;
  int64_t blocks = (n + threads - 1) / threads;
  if (blocks > 148 * 32) blocks = 148 * 32;
  if (f32)
    scatter_add_kernel<float><<<(unsigned)blocks, threads, 0, (cudaStream_t)stream>>>(n, d_pix, d_val, (float*)d_map);
  else
    scatter_add_kernel<double><<<(unsigned)blocks, threads, 0, (cudaStream_t)stream>>>(n, d_pix, d_val, (double*)d_map);
  AP_LAUNCH_CHECK();
  return 0;
}
extern "C" int ap_scatter_add(int64_t n, const int64_t* d_pix, const double* d_val, double* d_map, void* stream) {
  return scatter_add_impl(n, d_pix, d_val, d_map, false, stream);
}
extern "C" int ap_scatter_add_f32(int64_t n, const int64_t* d_pix, const double* d_val, float* d_map, void* stream) {
  return scatter_add_impl(n, d_pix, d_val, d_map, true, stream);
}

// ------------------------------------------------------------------------------------------
// K5: cutouts and stack
// ------------------------------------------------------------------------------------------
// healpy CartesianProj.ij2xy (projector.py; reached through projmap at paint_bucket.py:1719): the image grid is
// ENDPOINT-INCLUSIVE, x_j = lonra'[0] + j (lonra'[1] - lonra'[0]) / (xsize - 1), y_i = latra[0] + i (latra[1] -
// latra[0]) / (ysize - 1) [deg], with lonra' = flip * lonra[::flip] = [-lonra[1], -lonra[0]] for the default
// 'astro' flip (-1); xy2vec then takes theta = pi/2 - y, phi = flip * x.  So column 0 looks at longitude
// offset +lonra[1] (east is left) and the last column at lonra[0].
struct CutGrid {
  double x0, dx, y0, dy;  // x_j = dx * j + x0 ; y_i = dy * i + y0  [deg]
  int32_t xsize, ysize;
};

__device__ __forceinline__ void image_vec(const CutGrid& g, int32_t s, double& vx, double& vy, double& vz) {
  int32_t i = s / g.xsize, j = s - i * g.xsize;
  const double dtor = kPi / 180.0;
  double x = g.dx * (double)j + g.x0;
  double y = g.dy * (double)i + g.y0;
  double theta = kHalfPi - y * dtor;  // xy2vec: theta = pi/2 - y, phi = flip * x, flip = -1 (astro)
  double phi = -(x * dtor);
  double st, ct, sp, cp;
  sincos(theta, &st, &ct);
  sincos(phi, &sp, &cp);
  vx = st * cp;
  vy = st * sp;
  vz = ct;
}

// Rotator(rot=(lon, lat)).I = m1(lon) m2(-lat): rows of the 3x3 matrix
__device__ __forceinline__ void rot_matrix(double lon_deg, double lat_deg, double* m) {
  const double dtor = kPi / 180.0;
  double s1, c1, s2, c2;
  sincos(lon_deg * dtor, &s1, &c1);
  sincos(-(lat_deg * dtor), &s2, &c2);
  m[0] = c1 * c2; m[1] = -s1; m[2] = c1 * s2;
  m[3] = s1 * c2; m[4] = c1;  m[5] = s1 * s2;
  m[6] = -s2;     m[7] = 0.0; m[8] = c2;
}

constexpr int kStackThreads = 256;
constexpr int kStackHalos = 128;

template <bool STACK>
__global__ void __launch_bounds__(kStackThreads) cutout_kernel(Hpx hpx, const double* __restrict__ map, int64_t n_sel,
                                                               const double* lon, const double* lat, CutGrid grid,
                                                               const double* weight, double* out) {
  __shared__ double rot[kStackHalos][9];
  __shared__ double wgt[kStackHalos], phic[kStackHalos], cphi[kStackHalos], sphi[kStackHalos];
  const int32_t nsamp = grid.xsize * grid.ysize;
  const int32_t s = blockIdx.x * kStackThreads + threadIdx.x;
  const int64_t hbase = (int64_t)blockIdx.y * kStackHalos;
  const int nh = (int)min((int64_t)kStackHalos, n_sel - hbase);
  for (int i = threadIdx.x; i < nh; i += kStackThreads) {
    rot_matrix(lon[hbase + i], lat[hbase + i], rot[i]);
    wgt[i] = weight ? weight[hbase + i] : 1.0;
    double pc = lon[hbase + i] * (kPi / 180.0);  // the halo's azimuth: every sample of its cutout is close to it
    phic[i] = pc;
    sincos(pc, &sphi[i], &cphi[i]);
  }
  __syncthreads();
  if (s >= nsamp) return;
  double vx, vy, vz;
  image_vec(grid, s, vx, vy, vz);
  double acc = 0.0;
  for (int i = 0; i < nh; ++i) {
    const double* m = rot[i];
    double sx = m[0] * vx + m[1] * vy + m[2] * vz;
    double sy = m[3] * vx + m[4] * vy + m[5] * vz;
    double sz = m[6] * vx + m[7] * vy + m[8] * vz;
    int64_t pix = vec2pix_ring_near(hpx, sx, sy, sz, phic[i], cphi[i], sphi[i]);
    if (!AP_BOUNDS_OK(pix >= 0 && pix < hpx.npix)) continue;
    double v = __ldg(&map[pix]);
    if (STACK)
      acc += wgt[i] * v;
    else
      out[(hbase + i) * (int64_t)nsamp + s] = v;
  }
  if (STACK) atomicAdd(&out[s], acc);
}

static int cut_grid(double lon0, double lon1, double lat0, double lat1, int32_t xsize, int32_t ysize, CutGrid& g) {
  if (xsize < 1 || ysize < 1) return fail("xsize/ysize must be >= 1");
  // CartesianProj.set_proj_plane_info: lonra[1] = lonra[0] + (mod(lonra[1]-lonra[0], 360) or 360)
  double span = fmod(lon1 - lon0, 360.0);
  if (span < 0) span += 360.0;
  if (span == 0) span = 360.0;
  if (!(lat0 >= -90.0 && lat1 <= 90.0 && lat0 < lat1)) return fail("invalid lat_range");
  g.x0 = -(lon0 + span);   // flip * lonra[::flip][0] with lonra[1] = lonra[0] + span
  g.dx = xsize > 1 ? span / (double)(xsize - 1) : 0.0;
  g.y0 = lat0;
  g.dy = ysize > 1 ? (lat1 - lat0) / (double)(ysize - 1) : 0.0;
  g.xsize = xsize;
  g.ysize = ysize;
  return 0;
}

extern "C" int ap_cutouts(int64_t nside, const double* d_map, int64_t n_sel, const double* d_lon_deg,
                          const double* d_lat_deg, double lon0, double lon1, double lat0, double lat1, int32_t xsize,
                          int32_t ysize, double* d_out, void* stream) {
  if (check_nside(nside)) return 1;
  CutGrid g;
  if (cut_grid(lon0, lon1, lat0, lat1, xsize, ysize, g)) return 1;
  if (n_sel < 0) return fail("n_sel < 0");
  if (n_sel == 0) return 0;
  if (!d_map || !d_lon_deg || !d_lat_deg || !d_out) return fail("cutouts: NULL pointer");
  dim3 grid((xsize * ysize + kStackThreads - 1) / kStackThreads, (unsigned)((n_sel + kStackHalos - 1) / kStackHalos));
  if (grid.y > 65535) return fail("too many halos in one ap_cutouts call (max 8388480)");
  cutout_kernel<false><<<grid, kStackThreads, 0, (cudaStream_t)stream>>>(make_hpx(nside), d_map, n_sel, d_lon_deg, d_lat_deg,
                                                                        g, nullptr, d_out);
  AP_LAUNCH_CHECK();
  return 0;
}

extern "C" int ap_stack_cutouts(int64_t nside, const double* d_map, int64_t n_sel, const double* d_lon_deg,
                                const double* d_lat_deg, double lon0, double lon1, double lat0, double lat1,
                                int32_t xsize, int32_t ysize, const double* d_weight, double* d_stack, void* stream) {
  if (check_nside(nside)) return 1;
  CutGrid g;
  if (cut_grid(lon0, lon1, lat0, lat1, xsize, ysize, g)) return 1;
  if (n_sel < 0) return fail("n_sel < 0");
  if (n_sel == 0) return 0;
  if (!d_map || !d_lon_deg || !d_lat_deg || !d_stack) return fail("stack_cutouts: NULL pointer");
  dim3 grid((xsize * ysize + kStackThreads - 1) / kStackThreads, (unsigned)((n_sel + kStackHalos - 1) / kStackHalos));
  if (grid.y > 65535) return fail("too many halos in one ap_stack_cutouts call (max 8388480)");
  cutout_kernel<true><<<grid, kStackThreads, 0, (cudaStream_t)stream>>>(make_hpx(nside), d_map, n_sel, d_lon_deg, d_lat_deg,
                                                                       g, d_weight, d_stack);
  AP_LAUNCH_CHECK();
  return 0;
}

// ------------------------------------------------------------------------------------------
// host-buffer entry
// ------------------------------------------------------------------------------------------
extern "C" int ap_spray_host(int profile, int64_t nside, int64_t n, const double* h_x, const double* h_y,
                             const double* h_z, const double* h_radius, const double* h_theta, const double* h_phi,
                             const double* h_D_a, const double* const* h_params, const int32_t* h_strides,
                             int32_t n_params, double* h_map, uint64_t* h_n_painted) {
  if (check_nside(nside)) return 1;
  if (n < 0) return fail("n < 0");
  if (n_params < 0 || n_params > kMaxParams) return fail("bad n_params");
  if (!h_map) return fail("h_map == NULL");
  const int64_t npix = 12 * nside * nside;
  const double* cols[7] = {h_x, h_y, h_z, h_radius, h_theta, h_phi, h_D_a};
  double* d_cols[7] = {0};
  double* d_par[kMaxParams] = {0};
  double* d_map = nullptr;
  unsigned long long* d_cnt = nullptr;
  int rc = 0;
  cudaStream_t st = nullptr;
  auto cleanup = [&]() {
    for (auto p : d_cols) if (p) cudaFree(p);
    for (auto p : d_par) if (p) cudaFree(p);
    if (d_map) cudaFree(d_map);
    if (d_cnt) cudaFree(d_cnt);
    if (st) cudaStreamDestroy(st);
  };
#define AP_TRY(call)                                                                  \
  do {                                                                                \
    cudaError_t e_ = (call);                                                          \
    if (e_ != cudaSuccess) {                                                          \
      cleanup();                                                                      \
      return fail(std::string(#call) + ": " + cudaGetErrorString(e_));                \
    }                                                                                 \
  } while (0)
  AP_TRY(cudaStreamCreate(&st));
  AP_TRY(cudaMalloc(&d_map, sizeof(double) * npix));
  AP_TRY(cudaMemcpyAsync(d_map, h_map, sizeof(double) * npix, cudaMemcpyHostToDevice, st));
  AP_TRY(cudaMalloc(&d_cnt, sizeof(unsigned long long)));
  AP_TRY(cudaMemsetAsync(d_cnt, 0, sizeof(unsigned long long), st));
  for (int i = 0; i < 7 && n > 0; ++i) {
    if (!cols[i]) { cleanup(); return fail("NULL halo column"); }
    AP_TRY(cudaMalloc(&d_cols[i], sizeof(double) * n));
    AP_TRY(cudaMemcpyAsync(d_cols[i], cols[i], sizeof(double) * n, cudaMemcpyHostToDevice, st));
  }
  for (int i = 0; i < n_params && n > 0; ++i) {
    if (!h_params || !h_params[i]) { cleanup(); return fail("NULL template parameter"); }
    int64_t len = (h_strides && h_strides[i] == 0) ? 1 : n;
    AP_TRY(cudaMalloc(&d_par[i], sizeof(double) * len));
    AP_TRY(cudaMemcpyAsync(d_par[i], h_params[i], sizeof(double) * len, cudaMemcpyHostToDevice, st));
  }
  {
    ap_halos H;
    H.n = n;
    H.d_x = d_cols[0]; H.d_y = d_cols[1]; H.d_z = d_cols[2]; H.d_radius = d_cols[3];
    H.d_theta = d_cols[4]; H.d_phi = d_cols[5]; H.d_D_a = d_cols[6];
    rc = ap_paint_profile(profile, nside, &H, d_par, h_strides, n_params, d_map, d_cnt, st);
  }
  if (rc == 0) {
    AP_TRY(cudaMemcpyAsync(h_map, d_map, sizeof(double) * npix, cudaMemcpyDeviceToHost, st));
    unsigned long long cnt = 0;
    AP_TRY(cudaMemcpyAsync(&cnt, d_cnt, sizeof(cnt), cudaMemcpyDeviceToHost, st));
    AP_TRY(cudaStreamSynchronize(st));
    if (h_n_painted) *h_n_painted = cnt;
  }
  cleanup();
  return rc;
}

// ------------------------------------------------------------------------------------------
// host map plumbing: page-locking of a caller-owned host range (the rank's band of a map shared between the
// processes of one node) and the asynchronous band copy device -> host
// ------------------------------------------------------------------------------------------
extern "C" int ap_host_register(void* h_ptr, uint64_t bytes) {
  if (!h_ptr || bytes == 0) return fail("host_register: empty range");
  cudaError_t e = cudaHostRegister(h_ptr, (size_t)bytes, cudaHostRegisterPortable);
  if (e != cudaSuccess) {
    cudaGetLastError();
    return fail(std::string("cudaHostRegister: ") + cudaGetErrorString(e));
  }
  return 0;
}
extern "C" int ap_host_unregister(void* h_ptr) {
  if (!h_ptr) return fail("host_unregister: NULL");
  cudaError_t e = cudaHostUnregister(h_ptr);
  if (e != cudaSuccess) {
    cudaGetLastError();
    return fail(std::string("cudaHostUnregister: ") + cudaGetErrorString(e));
  }
  return 0;
}
extern "C" int ap_copy_d2h_async(void* h_dst, const void* d_src, uint64_t bytes, void* stream) {
  if (bytes == 0) return 0;
  if (!h_dst || !d_src) return fail("copy_d2h: NULL pointer");
  cudaError_t e = cudaMemcpyAsync(h_dst, d_src, (size_t)bytes, cudaMemcpyDeviceToHost, (cudaStream_t)stream);
  if (e != cudaSuccess) return fail(std::string("cudaMemcpyAsync(D2H): ") + cudaGetErrorString(e));
  return 0;
}
extern "C" int ap_copy_h2d_async(void* d_dst, const void* h_src, uint64_t bytes, void* stream) {
  if (bytes == 0) return 0;
  if (!d_dst || !h_src) return fail("copy_h2d: NULL pointer");
  cudaError_t e = cudaMemcpyAsync(d_dst, h_src, (size_t)bytes, cudaMemcpyHostToDevice, (cudaStream_t)stream);
  if (e != cudaSuccess) return fail(std::string("cudaMemcpyAsync(H2D): ") + cudaGetErrorString(e));
  return 0;
}

// ------------------------------------------------------------------------------------------
// fp64 roofline probe: 8 independent DFMA chains per thread, register operands, 8 blocks of 256 threads per
// SM.  bench.py times it with CUDA events next to the QUADPACK node kernel: the denominator of that kernel's
// roofline fraction is this MEASURED fp64 instruction rate (MEASURED_PEAKS.json has no fp64 entry).
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) fp64_peak_kernel(double* out, int iters, double a, double b) {
  double x[8];
#pragma unroll
  for (int c = 0; c < 8; ++c) x[c] = threadIdx.x * 1e-9 + c;
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int c = 0; c < 8; ++c) x[c] = fma(x[c], a, b);
  }
  double s = 0;
#pragma unroll
  for (int c = 0; c < 8; ++c) s += x[c];
  if (s == 1.2345) out[0] = s;  // never true: keeps the chains alive
}
extern "C" int ap_fp64_peak(int32_t iters, int32_t blocks, double* d_out, uint64_t* h_n_dfma, void* stream) {
  if (iters < 1 || blocks < 1 || !d_out) return fail("fp64_peak: bad arguments");
  fp64_peak_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(d_out, iters, 1.0000001, 1e-9);
  AP_LAUNCH_CHECK();
  if (h_n_dfma) *h_n_dfma = (uint64_t)blocks * 256ull * 8ull * (uint64_t)iters;  // thread-level DFMAs issued
  return 0;
}

// K6: per-cutout operators (rotate / taper / weight / stack) for Canvas.cutouts(apply_func=...)
#include "patch_ops.cuh"
