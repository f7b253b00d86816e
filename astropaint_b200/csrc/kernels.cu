// astropaint_b200 — sm_100a kernels and the C ABI (include/astropaint_b200.h).
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
#include <cuda_runtime.h>
#include <cstdio>
#include <cstring>
#include <string>

#include "../../include/astropaint_b200.h"
#include "common.cuh"
#include "healpix.cuh"
#include "geom.cuh"
#include "profiles.cuh"
#include "spline.cuh"

using namespace ap;

// ------------------------------------------------------------------------------------------
// error plumbing
// ------------------------------------------------------------------------------------------
static thread_local std::string g_err;
static int fail(const std::string& m) {
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
static Hpx disc_hpx(int64_t nside) {
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
  }
  return h;
}
#define AP_LAUNCH_CHECK()                                                               \
  do {                                                                                  \
    cudaError_t e_ = cudaGetLastError();                                                \
    if (e_ != cudaSuccess) return fail(std::string("launch: ") + cudaGetErrorString(e_)); \
  } while (0)

// Optional bounds instrumentation (-DAP_DEBUG_BOUNDS, tools/debug_bounds.sh): compute-sanitizer is not
// available on the B200 pool, so a debug build counts every out-of-range index instead of making the
// access; ap_debug_bounds_errors() must read 0 after the GPU test-suite.
__device__ unsigned long long g_bounds_errors = 0ull;
#ifdef AP_DEBUG_BOUNDS
#define AP_BOUNDS_OK(cond) ((cond) ? true : (atomicAdd(&g_bounds_errors, 1ull), false))
#else
#define AP_BOUNDS_OK(cond) true
#endif
extern "C" long long ap_debug_bounds_errors(void) {
#ifdef AP_DEBUG_BOUNDS
  unsigned long long v = 0;
  if (cudaMemcpyFromSymbol(&v, g_bounds_errors, sizeof(v)) != cudaSuccess) return -2;
  return (long long)v;
#else
  return -1;  // not an instrumented build
#endif
}

extern "C" const char* ap_last_error(void) { return g_err.c_str(); }
extern "C" int ap_version(void) { return 100; }

struct HalosDev {
  int64_t n;
  const double *x, *y, *z, *radius, *theta, *phi, *D_a;
};
static HalosDev to_dev(const ap_halos* h) {
  HalosDev d;
  d.n = h->n;
  d.x = h->d_x; d.y = h->d_y; d.z = h->d_z; d.radius = h->d_radius;
  d.theta = h->d_theta; d.phi = h->d_phi; d.D_a = h->d_D_a;
  return d;
}

static int check_halos(const ap_halos* h, bool need_center) {
  if (!h) return fail("halos == NULL");
  if (h->n < 0) return fail("halos.n < 0");
  if (h->n > 0 && (!h->d_x || !h->d_y || !h->d_z || !h->d_radius)) return fail("halos: x/y/z/radius NULL");
  if (need_center && h->n > 0 && (!h->d_theta || !h->d_phi || !h->d_D_a)) return fail("halos: theta/phi/D_a NULL");
  return 0;
}
static int check_nside(int64_t nside) {
  if (nside < 1 || nside > (int64_t(1) << 29)) return fail("nside out of range");
  // ring pixel counts and ring numbers are held in int32 inside the kernels
  if (nside > 8192 * 32) return fail("nside > 262144 not supported by the int32 ring arithmetic");
  return 0;
}

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
// K3: fused painter
// ------------------------------------------------------------------------------------------
constexpr int kPaintThreads = 256;
constexpr int kMaxHalosPerBlock = 112;
// templates that run a quadrature per PIXEL also hold the log/exp tables in shared memory (fastmath.cuh):
// fewer halos per block keep them inside the 48 KB static limit; their time is all quadrature anyway
template <int P>
constexpr int max_halos_per_block() {
  return (P == P_B16_TAU2D || P == P_B16_RHOGAS2D || P == P_NFW_RHO2D_LOS) ? 48 : kMaxHalosPerBlock;
}
constexpr int kHintMax = 256;     // warp-group hints cover chunks up to 8 Ki pixels
constexpr int kPixMapMax = 4096;  // pixel -> ring-item byte map for the first 4 Ki pixels of a chunk
constexpr int kStageHalos = 28;   // spline coefficients staged in shared memory per chunk (P_TABLE)
constexpr int kStageMaxNodes = 15;
constexpr int P_TABLE = 100;   // tabulated template (spline of per-halo nodes)
constexpr int P_STATS = 101;   // no painting: per-halo pixel count and R range (K1)

struct ParamPtrs {
  const double* p[kMaxParams];
  int32_t stride[kMaxParams];
  int32_t n;
};

struct TableArgs {
  const double* nodes;
  const double* coef;
  const int32_t* n_nodes;
  const double* scale;
  int32_t n_samples;
  int32_t uniform;  // nodes are linspace: interval index by multiplication
};

struct StatsArgs {
  int64_t* n_pix;
  int32_t* n_rings;
  double* rmin;
  double* rmax;
  // optional work list of the `len(R) < n_samples` bypass: halos with 0 < n_pix < item_ns append one packed
  // item (halo * 32 + pixel ordinal) per pixel at items[atomicAdd(item_count, n_pix) ...] (capacity item_cap)
  unsigned long long* item_count;
  int64_t* items;
  int64_t item_cap;
  int32_t item_ns;
};

struct HaloHdr {  // DiscHeader with 32-bit ring numbers (nside <= 2^18)
  double phi, z0, xa, cosr, cosrs;
  int32_t irmin, irmax, ir_first, ir_last;
};

struct RingItem {
  int64_t startpix;
  int32_t nr, ip_lo, len, hl;
  double A, B, phi_step, d0;
};
struct RingItemVec {  // extra per-ring data of R_vec templates
  double phi0, z, sth;
};

template <int P>
struct PaintShared {
  static constexpr int MAXH = max_halos_per_block<P>();
  static constexpr bool VEC = (P == P_NFW_BG);
  static constexpr int NCTX = (P == P_TABLE) ? 4 : ((P == P_STATS) ? 1 : kCtx);
  HaloHdr hdr[MAXH];
  double cen_cos[MAXH], cen_phi[MAXH], cen_sin[MAXH],
      cen_Da[MAXH];
  double cen_vec[VEC ? MAXH : 1][3];
  double ctx[MAXH][NCTX];
  int32_t ring_off[MAXH + 1];
  RingItem item[kPaintThreads];
  RingItemVec itemv[VEC ? kPaintThreads : 1];
  int32_t pix_off[kPaintThreads + 1];
  int32_t warp_sum[kPaintThreads / 32];
  int32_t hint[kHintMax];  // hint[g] = ring item holding pixel 32 g of the current chunk
  uint8_t pixmap[P == P_STATS ? 1 : kPixMapMax];  // ring item of each of the first 4 Ki pixels
  double stage[P == P_TABLE ? kStageHalos * (kStageMaxNodes - 1) * 4 : 1];
  // P_STATS accumulators
  unsigned long long cnt[P == P_STATS ? MAXH : 1];
  unsigned long long hmin[P == P_STATS ? MAXH : 1], hmax[P == P_STATS ? MAXH : 1];
};

// One block paints `halos_per_block` consecutive halos.  Phase 0: per-halo disc header, centre and
// template constants (one thread per halo).  Then, 256 ring items at a time: phase 1 computes one
// ring span per thread (query_disc's per-ring work) and its ring-constant geometry, a block scan of
// the span lengths gives the flat pixel numbering; phase 2 walks the pixels with consecutive lanes
// on consecutive pixels of a ring (sector-coalesced fp64 RED traffic), evaluates separation,
// template and accumulates.  No work list ever touches HBM.
template <int P, class MapT, bool INCL>
__global__ void __launch_bounds__(kPaintThreads) paint_kernel(Hpx hpx, HalosDev H, ParamPtrs pp, TableArgs tab,
                                                               StatsArgs st, int halos_per_block,
                                                               MapT* __restrict__ map,
                                                               unsigned long long* n_painted) {
  constexpr bool VEC = (P == P_NFW_BG);
  constexpr bool STATS = (P == P_STATS);
  __shared__ PaintShared<P> sm;
  if (P == P_B16_TAU2D || P == P_B16_RHOGAS2D || P == P_NFW_RHO2D_LOS) fm_smem_init();
  const int tid = threadIdx.x;
  const int lane = tid & 31, warp = tid >> 5;
  const int64_t h0 = (int64_t)blockIdx.x * halos_per_block;
  const int G = (int)min((int64_t)halos_per_block, H.n - h0);

  // phase 0
  if (tid < G) {
    int64_t h = h0 + tid;
    DiscHeader d = disc_header_t<INCL>(hpx, H.x[h], H.y[h], H.z[h], H.radius[h]);
    if (P == P_TABLE && tab.n_nodes[h] == 0) {  // bypass halo: painted elsewhere
      d.ir_first = 1;
      d.ir_last = 0;
    }
    if (STATS) {  // counts stay those of the WHOLE disc; a disc that misses the rank's band reports 0 pixels
      if (d.ir_last < hpx.clip_lo || d.ir_first >= hpx.clip_hi) {
        d.ir_first = 1;
        d.ir_last = 0;
      }
    } else {      // paint the rings of the rank's own band only
      if (d.ir_first < hpx.clip_lo) d.ir_first = hpx.clip_lo;
      if (d.ir_last >= hpx.clip_hi) d.ir_last = hpx.clip_hi - 1;
    }
    HaloHdr hh;
    hh.phi = d.phi; hh.z0 = d.z0; hh.xa = d.xa; hh.cosr = d.cosr; hh.cosrs = d.cosrs;
    hh.irmin = (int32_t)d.irmin; hh.irmax = (int32_t)d.irmax;
    hh.ir_first = (int32_t)d.ir_first; hh.ir_last = (int32_t)d.ir_last;
    sm.hdr[tid] = hh;
    if (!STATS || st.rmin) {
      HaloCenter c = make_center(H.theta[h], H.phi[h], H.D_a[h], VEC);
      sm.cen_cos[tid] = c.cos_t; sm.cen_phi[tid] = c.phi; sm.cen_sin[tid] = c.sin_t; sm.cen_Da[tid] = c.D_a;
      if (VEC) { sm.cen_vec[tid][0] = c.cx; sm.cen_vec[tid][1] = c.cy; sm.cen_vec[tid][2] = c.cz; }
    }
    if (STATS) {
      sm.cnt[tid] = 0ull;
      sm.hmin[tid] = 0x7ff0000000000000ull;  // +inf
      sm.hmax[tid] = 0ull;                   // haversines are >= 0: their bit patterns order like uint64
    } else if (P == P_TABLE) {
      const double* x = tab.nodes + h * tab.n_samples;
      sm.ctx[tid][0] = tab.scale ? tab.scale[h] : 1.0;
      sm.ctx[tid][1] = x[0];
      sm.ctx[tid][3] = (x[tab.n_samples - 1] - x[0]) / (double)(tab.n_samples - 1);
      sm.ctx[tid][2] = 1.0 / sm.ctx[tid][3];
    } else {
      double p[kMaxParams];
#pragma unroll
      for (int i = 0; i < kMaxParams; ++i)
        if (i < pp.n) p[i] = pp.p[i][h * pp.stride[i]];
      profile_setup<P>(p, sm.ctx[tid]);
    }
  }
  __syncthreads();
  if (tid == 0) {
    int32_t run = 0;
    for (int i = 0; i < G; ++i) {
      sm.ring_off[i] = run;
      int32_t n = sm.hdr[i].ir_last - sm.hdr[i].ir_first + 1;
      run += n > 0 ? n : 0;
    }
    sm.ring_off[G] = run;
  }
  __syncthreads();
  const int32_t T = sm.ring_off[G];
  unsigned long long painted = 0;

  for (int32_t base = 0; base < T; base += kPaintThreads) {
    // P_TABLE: the spline coefficients of the halos this chunk touches are one contiguous run in HBM;
    // start streaming the first kStageHalos of them now (coalesced), park them in shared memory after
    // phase 1 so that phase 2 never waits on a dependent global load
    constexpr int kStageRegs = (P == P_TABLE) ? (kStageHalos * (kStageMaxNodes - 1) * 4 + kPaintThreads - 1) / kPaintThreads : 1;
    double stage_reg[kStageRegs];
    int stage_lo = 0, stage_n = 0;
    const bool do_stage = (P == P_TABLE) && tab.n_samples <= kStageMaxNodes;
    if (do_stage) {
      int lo = 0, hi = G;
      while (hi - lo > 1) {
        int mid = (lo + hi) >> 1;
        if (sm.ring_off[mid] <= base) lo = mid; else hi = mid;
      }
      stage_lo = lo;
      const int nc = (tab.n_samples - 1) * 4;
      stage_n = min(kStageHalos, G - lo) * nc;
      const double* src = tab.coef + (h0 + lo) * (int64_t)nc;
#pragma unroll
      for (int q = 0; q < kStageRegs; ++q) {
        int e = tid + q * kPaintThreads;
        stage_reg[q] = (e < stage_n) ? __ldg(src + e) : 0.0;
      }
    }

    // phase 1: one ring item per thread
    int32_t it = base + tid;
    int32_t len = 0;
    if (it < T) {
      int lo = 0, hi = G;  // largest hl with ring_off[hl] <= it
      while (hi - lo > 1) {
        int mid = (lo + hi) >> 1;
        if (sm.ring_off[mid] <= it) lo = mid; else hi = mid;
      }
      const HaloHdr& hh = sm.hdr[lo];
      DiscHeader d;
      d.phi = hh.phi; d.z0 = hh.z0; d.xa = hh.xa; d.cosr = hh.cosr; d.cosrs = hh.cosrs;
      d.irmin = hh.irmin; d.irmax = hh.irmax; d.ir_first = hh.ir_first; d.ir_last = hh.ir_last;
      int64_t iz = d.ir_first + (it - sm.ring_off[lo]);
      RingSpan s = ring_span_t<INCL>(hpx, d, iz);
      len = s.len;
      RingItem& r = sm.item[tid];
      r.len = s.len;
      if (len > 0 && (!STATS || st.rmin)) {
        HaloCenter c;
        c.cos_t = sm.cen_cos[lo]; c.phi = sm.cen_phi[lo]; c.sin_t = sm.cen_sin[lo]; c.D_a = sm.cen_Da[lo];
        RingGeom g = ring_geom(hpx, iz, s, c);
        if (STATS) {
          double a, b;
          span_hav_range(g, s.len, a, b);
          atomicMin(&sm.hmin[lo], (unsigned long long)__double_as_longlong(a));
          atomicMax(&sm.hmax[lo], (unsigned long long)__double_as_longlong(b));
        } else {
          r.startpix = s.startpix; r.nr = s.nr; r.ip_lo = s.ip_lo; r.hl = lo;
          r.A = g.A; r.B = g.B; r.phi_step = g.phi_step; r.d0 = g.d0;
          if (VEC) { sm.itemv[tid].phi0 = g.phi0; sm.itemv[tid].z = g.z; sm.itemv[tid].sth = g.sth; }
        }
      }
      if (STATS && len > 0) atomicAdd(&sm.cnt[lo], (unsigned long long)len);
    }
    if (STATS) continue;  // counts and ranges only: no pixel walk (no barrier needed either)

    // block exclusive scan of len
    int32_t v = len;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      int32_t t = __shfl_up_sync(0xffffffffu, v, o);
      if (lane >= o) v += t;
    }
    if (lane == 31) sm.warp_sum[warp] = v;
    __syncthreads();
    int32_t woff = 0;
#pragma unroll
    for (int w = 0; w < kPaintThreads / 32; ++w)
      if (w < warp) woff += sm.warp_sum[w];
    {
      const int32_t start = woff + v - len, end = start + len;
      sm.pix_off[tid] = start;
      if (tid == kPaintThreads - 1) sm.pix_off[kPaintThreads] = end;
      // every multiple of 32 inside [start, end) starts a warp-sized pixel group in this ring item
      for (int32_t m = (start + 31) & ~31; m < end && (m >> 5) < kHintMax; m += 32) sm.hint[m >> 5] = tid;
      const int32_t mend = min(end, (int32_t)kPixMapMax);
      for (int32_t m = start; m < mend; ++m) sm.pixmap[m] = (uint8_t)tid;
      if (do_stage) {
#pragma unroll
        for (int q = 0; q < kStageRegs; ++q) {
          int e = tid + q * kPaintThreads;
          if (e < stage_n) sm.stage[e] = stage_reg[q];
        }
      }
    }
    __syncthreads();
    const int32_t total = sm.pix_off[kPaintThreads];
    painted += (tid == 0) ? (unsigned long long)total : 0ull;

    // phase 2: two pixels per thread per iteration (k and k + 256): independent dependency chains
    // (table loads, series) overlap, then both REDs are issued
    auto eval = [&](int32_t k, int64_t& pix) -> double {
      int lo;
      if (k < kPixMapMax) {
        lo = sm.pixmap[k];
        if (!AP_BOUNDS_OK(k >= sm.pix_off[lo] && k < sm.pix_off[lo + 1])) lo = 0;
      } else if ((k >> 5) < kHintMax) {
        lo = sm.hint[k >> 5];
        while (sm.pix_off[lo + 1] <= k) ++lo;
      } else {  // chunk with more than 32 Ki pixels (sky-sized discs): plain binary search
        lo = 0;
        int hi = kPaintThreads;
        while (hi - lo > 1) {
          int mid = (lo + hi) >> 1;
          if (sm.pix_off[mid] <= k) lo = mid; else hi = mid;
        }
      }
      const RingItem& r = sm.item[lo];
      const int32_t j = k - sm.pix_off[lo];
      const int hl = r.hl;
      double val;
      if (VEC) {
        RingGeom g;
        g.phi_step = r.phi_step; g.phi0 = sm.itemv[lo].phi0; g.z = sm.itemv[lo].z; g.sth = sm.itemv[lo].sth;
        HaloCenter c;
        c.D_a = sm.cen_Da[hl]; c.cx = sm.cen_vec[hl][0]; c.cy = sm.cen_vec[hl][1]; c.cz = sm.cen_vec[hl][2];
        double rx, ry, rz;
        pixel_chord(g, j, c, rx, ry, rz);
        val = profile_eval_vec<P>(sm.ctx[hl], rx, ry, rz);
      } else {
        double hav = r.A + r.B * sin2_half(r.d0 + (double)j * r.phi_step);
        double R = sm.cen_Da[hl] * sep_from_hav(hav);
        if (P == P_TABLE) {
          const int64_t h = h0 + hl;
          const int ns = tab.n_samples;
          const double* x = tab.nodes + h * ns;
          int i;
          if (tab.uniform) {  // linspace nodes: the spline is C2, so +-1 ulp at a knot picks an equal piece
            i = (int)((R - sm.ctx[hl][1]) * sm.ctx[hl][2]);
            i = i < 0 ? 0 : (i > ns - 2 ? ns - 2 : i);
          } else {
            i = spline_interval(ns, x, R);
          }
          double t;
          double2 c01, c23;
          const int sh = hl - stage_lo;
          if (do_stage && tab.uniform && sh < kStageHalos &&
              AP_BOUNDS_OK(sh >= 0 && (sh * (ns - 1) + i) * 4 + 3 < stage_n && i >= 0 && i <= ns - 2)) {
            t = R - (sm.ctx[hl][1] + (double)i * sm.ctx[hl][3]);  // node i of np.linspace to 1 ulp
            const double2* c2 = reinterpret_cast<const double2*>(sm.stage + (sh * (ns - 1) + i) * 4);
            c01 = c2[0];
            c23 = c2[1];
          } else {
            t = R - __ldg(x + i);
            const double2* c2 = reinterpret_cast<const double2*>(tab.coef + (h * (int64_t)(ns - 1) + i) * 4);
            c01 = __ldg(c2);
            c23 = __ldg(c2 + 1);
          }
          val = (c01.x + t * (c01.y + t * (c23.x + t * c23.y))) * sm.ctx[hl][0];
        } else {
          val = profile_eval_R<P>(sm.ctx[hl], R);
        }
      }
      int32_t ip = r.ip_lo + j;
      pix = r.startpix + (ip < 0 ? ip + r.nr : ip);
      return val;
    };
    for (int32_t k = tid; k < total; k += 2 * kPaintThreads) {
      int64_t p0, p1 = 0;
      const bool two = (k + kPaintThreads < total);
      double v0 = eval(k, p0);
      double v1 = two ? eval(k + kPaintThreads, p1) : 0.0;
      if (AP_BOUNDS_OK(p0 >= hpx.clip_pix0 && p0 < hpx.npix)) atomicAdd(&map[p0 - hpx.clip_pix0], (MapT)v0);
      if (two && AP_BOUNDS_OK(p1 >= hpx.clip_pix0 && p1 < hpx.npix)) atomicAdd(&map[p1 - hpx.clip_pix0], (MapT)v1);
    }
    __syncthreads();
  }
  if (STATS) {
    __syncthreads();
    if (tid < G) {
      int64_t h = h0 + tid;
      unsigned long long c = sm.cnt[tid];
      st.n_pix[h] = (int64_t)c;
      if (st.n_rings) {
        int32_t n = sm.hdr[tid].ir_last - sm.hdr[tid].ir_first + 1;
        st.n_rings[h] = n > 0 ? n : 0;
      }
      if (st.rmin) {
        double a = __longlong_as_double((long long)sm.hmin[tid]), b = __longlong_as_double((long long)sm.hmax[tid]);
        st.rmin[h] = c ? sm.cen_Da[tid] * sep_from_hav(a) : INFINITY;
        st.rmax[h] = c ? sm.cen_Da[tid] * sep_from_hav(b) : -INFINITY;
      }
      if (st.items && c > 0ull && c < (unsigned long long)st.item_ns) {
        const unsigned long long at = atomicAdd(st.item_count, c);
        for (unsigned long long k = 0; k < c; ++k)
          if ((int64_t)(at + k) < st.item_cap) st.items[at + k] = h * 32 + (int64_t)k;
      }
    }
    return;
  }
  if (n_painted && tid == 0 && painted) atomicAdd(n_painted, painted);
}

static int pick_halos_per_block(int64_t n, int max_halos) {
  // enough blocks to fill 148 SMs several times over; many halos per block amortise phase 0
  int64_t g = n / (148 * 8);
  if (g < 1) g = 1;
  if (g > max_halos) g = max_halos;
  return (int)g;
}

template <int P>
static int launch_paint(int64_t nside, const ap_halos* halos, const ParamPtrs& pp, const TableArgs& tab, void* d_map,
                        unsigned long long* d_n_painted, void* stream, const StatsArgs* stats = nullptr,
                        bool f32 = false) {
  if ((P == P_B16_TAU2D || P == P_B16_RHOGAS2D || P == P_NFW_RHO2D_LOS) && rm_ensure_tables(stream))
    return fail("could not initialise the quadrature tables");
  int g = pick_halos_per_block(halos->n, max_halos_per_block<P>());
  int64_t blocks = (halos->n + g - 1) / g;
  if (blocks > 0x7fffffffLL) return fail("too many blocks");
  StatsArgs st;
  memset(&st, 0, sizeof(st));
  if (stats) st = *stats;
  const Hpx hpx = disc_hpx(nside);
  const unsigned nb = (unsigned)blocks;
  cudaStream_t cs = (cudaStream_t)stream;
  if (f32 && P != P_STATS) {
    if (hpx.incl_fct)
      paint_kernel<P, float, true><<<nb, kPaintThreads, 0, cs>>>(hpx, to_dev(halos), pp, tab, st, g, (float*)d_map, d_n_painted);
    else
      paint_kernel<P, float, false><<<nb, kPaintThreads, 0, cs>>>(hpx, to_dev(halos), pp, tab, st, g, (float*)d_map, d_n_painted);
  } else {
    if (hpx.incl_fct)
      paint_kernel<P, double, true><<<nb, kPaintThreads, 0, cs>>>(hpx, to_dev(halos), pp, tab, st, g, (double*)d_map, d_n_painted);
    else
      paint_kernel<P, double, false><<<nb, kPaintThreads, 0, cs>>>(hpx, to_dev(halos), pp, tab, st, g, (double*)d_map, d_n_painted);
  }
  AP_LAUNCH_CHECK();
  return 0;
}

// K1: hp.query_disc counts (+ R range) with the painter's block structure, phase 2 skipped
static int disc_count_impl(int64_t nside, const ap_halos* halos, int inclusive, int64_t* d_n_pix, int32_t* d_n_rings,
                           double* d_rmin, double* d_rmax, int32_t item_ns, int64_t* d_items, int64_t item_cap,
                           unsigned long long* d_item_count, void* stream) {
  if (check_nside(nside)) return 1;
  bool want_r = d_rmin || d_rmax;
  if (check_halos(halos, want_r)) return 1;
  // `inclusive` != 0 asks for healpy's inclusive query for THIS call (factor 4 unless ap_disc_mode set one)
  struct ModeGuard {
    int prev;
    explicit ModeGuard(int inc) : prev(g_incl_fact) { if (inc && g_incl_fact == 0) g_incl_fact = 4; }
    ~ModeGuard() { g_incl_fact = prev; }
  } guard(inclusive);
  if (want_r && (!d_rmin || !d_rmax)) return fail("d_rmin and d_rmax must both be given");
  if (!d_n_pix) return fail("d_n_pix == NULL");
  if (d_items && (!d_item_count || item_cap < 0 || item_ns < 1 || item_ns > 32))
    return fail("disc_count: item list needs d_item_count, item_cap >= 0 and 1 <= n_samples <= 32");
  if (halos->n == 0) return 0;
  ParamPtrs pp;
  memset(&pp, 0, sizeof(pp));
  TableArgs tab;
  memset(&tab, 0, sizeof(tab));
  StatsArgs st;
  memset(&st, 0, sizeof(st));
  st.n_pix = d_n_pix; st.n_rings = d_n_rings; st.rmin = d_rmin; st.rmax = d_rmax;
  st.items = d_items; st.item_cap = item_cap; st.item_count = d_item_count; st.item_ns = item_ns;
  return launch_paint<P_STATS>(nside, halos, pp, tab, nullptr, nullptr, stream, &st);
}
extern "C" int ap_disc_count(int64_t nside, const ap_halos* halos, int inclusive, int64_t* d_n_pix,
                             int32_t* d_n_rings, double* d_rmin, double* d_rmax, void* stream) {
  return disc_count_impl(nside, halos, inclusive, d_n_pix, d_n_rings, d_rmin, d_rmax, 0, nullptr, 0, nullptr, stream);
}
extern "C" int ap_disc_count_items(int64_t nside, const ap_halos* halos, int64_t* d_n_pix, double* d_rmin, double* d_rmax,
                                   int32_t n_samples, int64_t* d_items, int64_t item_cap,
                                   unsigned long long* d_item_count, void* stream) {
  if (!d_items) return fail("disc_count_items: d_items == NULL");
  return disc_count_impl(nside, halos, 0, d_n_pix, nullptr, d_rmin, d_rmax, n_samples, d_items, item_cap, d_item_count,
                         stream);
}

static int paint_profile_impl(int profile, int64_t nside, const ap_halos* halos, const double* const* h_params,
                              const int32_t* h_strides, int32_t n_params, void* d_map, bool f32,
                              unsigned long long* d_n_painted, void* stream) {
  if (check_nside(nside)) return 1;
  if (check_halos(halos, true)) return 1;
  if (!d_map) return fail("d_map == NULL");
  int want = profile_n_params(profile);
  if (want < 0) return fail("unknown profile id");
  if (n_params != want) return fail("wrong number of template parameters for this profile");
  if (halos->n == 0) return 0;
  ParamPtrs pp;
  memset(&pp, 0, sizeof(pp));
  pp.n = n_params;
  for (int i = 0; i < n_params; ++i) {
    if (!h_params || !h_params[i]) return fail("NULL template parameter pointer");
    pp.p[i] = h_params[i];
    pp.stride[i] = h_strides ? h_strides[i] : 1;
    if (pp.stride[i] != 0 && pp.stride[i] != 1) return fail("stride must be 0 or 1");
  }
  TableArgs tab;
  memset(&tab, 0, sizeof(tab));
  switch (profile) {
    case P_CONSTANT: return launch_paint<P_CONSTANT>(nside, halos, pp, tab, d_map, d_n_painted, stream, nullptr, f32);
    case P_LINEAR: return launch_paint<P_LINEAR>(nside, halos, pp, tab, d_map, d_n_painted, stream, nullptr, f32);
    case P_SOLID_SPHERE: return launch_paint<P_SOLID_SPHERE>(nside, halos, pp, tab, d_map, d_n_painted, stream, nullptr, f32);
    case P_GAUSSIAN: return launch_paint<P_GAUSSIAN>(nside, halos, pp, tab, d_map, d_n_painted, stream, nullptr, f32);
    case P_NFW_RHO2D: return launch_paint<P_NFW_RHO2D>(nside, halos, pp, tab, d_map, d_n_painted, stream, nullptr, f32);
    case P_NFW_TAU2D: return launch_paint<P_NFW_TAU2D>(nside, halos, pp, tab, d_map, d_n_painted, stream, nullptr, f32);
    case P_NFW_KSZ: return launch_paint<P_NFW_KSZ>(nside, halos, pp, tab, d_map, d_n_painted, stream, nullptr, f32);
    case P_NFW_DEFLECTION: return launch_paint<P_NFW_DEFLECTION>(nside, halos, pp, tab, d_map, d_n_painted, stream, nullptr, f32);
    case P_NFW_BG: return launch_paint<P_NFW_BG>(nside, halos, pp, tab, d_map, d_n_painted, stream, nullptr, f32);
    case P_B16_TAU2D: return launch_paint<P_B16_TAU2D>(nside, halos, pp, tab, d_map, d_n_painted, stream, nullptr, f32);
    case P_B16_RHOGAS2D: return launch_paint<P_B16_RHOGAS2D>(nside, halos, pp, tab, d_map, d_n_painted, stream, nullptr, f32);
    case P_NFW_RHO2D_LOS: return launch_paint<P_NFW_RHO2D_LOS>(nside, halos, pp, tab, d_map, d_n_painted, stream, nullptr, f32);
  }
  return fail("unknown profile id");
}
extern "C" int ap_paint_profile(int profile, int64_t nside, const ap_halos* halos, const double* const* h_params,
                                const int32_t* h_strides, int32_t n_params, double* d_map,
                                unsigned long long* d_n_painted, void* stream) {
  return paint_profile_impl(profile, nside, halos, h_params, h_strides, n_params, d_map, false, d_n_painted, stream);
}
extern "C" int ap_paint_profile_f32(int profile, int64_t nside, const ap_halos* halos, const double* const* h_params,
                                    const int32_t* h_strides, int32_t n_params, float* d_map,
                                    unsigned long long* d_n_painted, void* stream) {
  return paint_profile_impl(profile, nside, halos, h_params, h_strides, n_params, d_map, true, d_n_painted, stream);
}

static int paint_table_impl(int64_t nside, const ap_halos* halos, int32_t n_samples, int32_t uniform_nodes,
                            const double* d_nodes, const double* d_coef, const int32_t* d_n_nodes,
                            const double* d_scale, void* d_map, bool f32, unsigned long long* d_n_painted,
                            void* stream) {
  if (check_nside(nside)) return 1;
  if (check_halos(halos, true)) return 1;
  if (n_samples < 4 || n_samples > kMaxNodes) return fail("n_samples must be in [4, 32]");
  if (!d_map || !d_nodes || !d_coef || !d_n_nodes) return fail("paint_table: NULL pointer");
  if (halos->n == 0) return 0;
  ParamPtrs pp;
  memset(&pp, 0, sizeof(pp));
  TableArgs tab;
  tab.nodes = d_nodes; tab.coef = d_coef; tab.n_nodes = d_n_nodes; tab.scale = d_scale; tab.n_samples = n_samples;
  tab.uniform = uniform_nodes ? 1 : 0;
  return launch_paint<P_TABLE>(nside, halos, pp, tab, d_map, d_n_painted, stream, nullptr, f32);
}
extern "C" int ap_paint_table(int64_t nside, const ap_halos* halos, int32_t n_samples, int32_t uniform_nodes,
                              const double* d_nodes, const double* d_coef, const int32_t* d_n_nodes,
                              const double* d_scale, double* d_map, unsigned long long* d_n_painted, void* stream) {
  return paint_table_impl(nside, halos, n_samples, uniform_nodes, d_nodes, d_coef, d_n_nodes, d_scale, d_map, false,
                          d_n_painted, stream);
}
extern "C" int ap_paint_table_f32(int64_t nside, const ap_halos* halos, int32_t n_samples, int32_t uniform_nodes,
                                  const double* d_nodes, const double* d_coef, const int32_t* d_n_nodes,
                                  const double* d_scale, float* d_map, unsigned long long* d_n_painted, void* stream) {
  return paint_table_impl(nside, halos, n_samples, uniform_nodes, d_nodes, d_coef, d_n_nodes, d_scale, d_map, true,
                          d_n_painted, stream);
}

// ------------------------------------------------------------------------------------------
// K4: @interpolate tables
// ------------------------------------------------------------------------------------------
__global__ void table_nodes_kernel(int64_t n, const int64_t* n_pix, const double* rmin, const double* rmax,
                                   int32_t ns, int32_t method, double* nodes, int32_t* n_nodes) {
  int64_t h = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (h >= n) return;
  if (n_pix[h] < ns) {  // len(R) < n_samples -> direct evaluation (utils.py:510-511)
    n_nodes[h] = 0;
    for (int i = 0; i < ns; ++i) nodes[h * ns + i] = 0.0;
    return;
  }
  n_nodes[h] = ns;
  double a = rmin[h], b = rmax[h];
  if (method == 1) {  // logspace: sample_array clamps the minimum to eps = 1e-3 (utils.py:417-421)
    if (a < 1.0e-3) a = 1.0e-3;
    a = log10(a);
    b = log10(b);
  }
  // np.linspace: step = (b - a) / (n - 1); y_i = i * step + a; y_{n-1} = b
  double step = AP_SUB(b, a) / (double)(ns - 1);
  for (int i = 0; i < ns; ++i) {
    double v = (i == ns - 1) ? b : AP_ADD(AP_MUL((double)i, step), a);
    if (method == 1) v = pow(10.0, v);
    nodes[h * ns + i] = v;
  }
}

extern "C" int ap_table_nodes(int64_t n, const int64_t* d_n_pix, const double* d_rmin, const double* d_rmax,
                              int32_t n_samples, int32_t method, double* d_nodes, int32_t* d_n_nodes, void* stream) {
  if (n < 0) return fail("n < 0");
  if (n_samples < 2 || n_samples > kMaxNodes) return fail("n_samples must be in [2, 32]");
  if (method != 0 && method != 1) return fail("method must be 0 (linspace) or 1 (logspace)");
  if (n == 0) return 0;
  if (!d_n_pix || !d_rmin || !d_rmax || !d_nodes || !d_n_nodes) return fail("table_nodes: NULL pointer");
  int threads = 128;
  table_nodes_kernel<<<(unsigned)((n + threads - 1) / threads), threads, 0, (cudaStream_t)stream>>>(
      n, d_n_pix, d_rmin, d_rmax, n_samples, method, d_nodes, d_n_nodes);
  AP_LAUNCH_CHECK();
  return 0;
}

// 6 blocks of 128 threads per SM (80 registers): measured 176.1 / 173.5 / 180.9 / 182.9 ms
// per 1e7 halos at 7 / 6 / 5 / 4 blocks (72 / 80 / 96 / 128 registers); the kernel is fp64-pipe bound
// (66% of the DFMA issue rate, profiles/), not occupancy bound
#ifndef AP_QAGI_THREADS
#define AP_QAGI_THREADS 128
#endif
#ifndef AP_QAGI_MINBLOCKS
#define AP_QAGI_MINBLOCKS 6
#endif
template <int KIND, int FAST>
__global__ void __launch_bounds__(AP_QAGI_THREADS, AP_QAGI_MINBLOCKS) los_nodes_kernel(int64_t n, const double* p0, const double* p1,
                                                                           const double* p2, int32_t ns, const double* nodes,
                                                                           const int32_t* n_nodes, double* values) {
  fm_smem_init();
  // one thread per (halo, node).  A persistent grid-stride form (each resident thread reusing ONE local-memory
  // frame, which removes the ~138 GB of dead local-memory write-back per 1e7 halos) measured 191 ms against
  // 170 ms: resident warps then run in lockstep through the fp64-heavy and the bookkeeping phases.
  int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n * ns) return;
  int64_t h = t / ns;
  if (n_nodes[h] == 0) {
    values[t] = 0.0;
    return;
  }
  double ctx[kCtx];
  los_setup<KIND>(p0[h], p1[h], p2 ? p2[h] : 0.0, ctx);
  values[t] = los_eval<KIND, FAST>(ctx, nodes[t]);
}

extern "C" int ap_los_nodes(int kind, int64_t n, const double* d_p0, const double* d_p1, const double* d_p2,
                            int32_t n_samples, const double* d_nodes, const int32_t* d_n_nodes, double* d_values,
                            void* stream) {
  if (n < 0) return fail("n < 0");
  if (kind < 0 || kind > LOS_NFW_RHO) return fail("unknown LOS integrand kind");
  if (n == 0) return 0;
  if (!d_p0 || !d_p1 || (kind != LOS_NFW_RHO && !d_p2) || !d_nodes || !d_n_nodes || !d_values)
    return fail("los_nodes: NULL pointer");
  if (rm_ensure_tables(stream)) return fail("could not initialise the quadrature tables");
  int threads = AP_QAGI_THREADS;
  int64_t total = n * n_samples;
  unsigned blocks = (unsigned)((total + threads - 1) / threads);
  cudaStream_t st = (cudaStream_t)stream;
#define AP_LOS_NODES(K)                                                                                                  \
  if (g_quad_fast)                                                                                                       \
    los_nodes_kernel<K, 1><<<blocks, threads, 0, st>>>(n, d_p0, d_p1, d_p2, n_samples, d_nodes, d_n_nodes, d_values);    \
  else                                                                                                                   \
    los_nodes_kernel<K, 0><<<blocks, threads, 0, st>>>(n, d_p0, d_p1, d_p2, n_samples, d_nodes, d_n_nodes, d_values)
  switch (kind) {
    case LOS_B16_TAU: AP_LOS_NODES(LOS_B16_TAU); break;
    case LOS_B16_NE: AP_LOS_NODES(LOS_B16_NE); break;
    case LOS_B16_RHO_GAS: AP_LOS_NODES(LOS_B16_RHO_GAS); break;
    case LOS_NFW_RHO: AP_LOS_NODES(LOS_NFW_RHO); break;
  }
#undef AP_LOS_NODES
  AP_LAUNCH_CHECK();
  return 0;
}

// diagnostic: tau_2D(R) with QUADPACK's own error estimate and evaluation count
__global__ void __launch_bounds__(128) b16_tau_eval_kernel(int64_t n, const double* R, const double* R200,
                                                           const double* M200, const double* zred, double* result,
                                                           double* abserr, int32_t* neval) {
  fm_smem_init();
  int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n) return;
  double ctx[kCtx];
  b16_setup(R200[t], M200[t], zred[t], ctx);
  QagiResult q;
  result[t] = b16_tau_2D(ctx, R[t], &q);
  if (abserr) abserr[t] = q.abserr;
  if (neval) neval[t] = q.neval;
}

extern "C" int ap_b16_tau_eval(int64_t n, const double* d_R, const double* d_R_200c, const double* d_M_200c,
                               const double* d_redshift, double* d_result, double* d_abserr, int32_t* d_neval,
                               void* stream) {
  if (n < 0) return fail("n < 0");
  if (n == 0) return 0;
  if (!d_R || !d_R_200c || !d_M_200c || !d_redshift || !d_result) return fail("b16_tau_eval: NULL pointer");
  if (rm_ensure_tables(stream)) return fail("could not initialise the quadrature tables");
  int threads = 128;
  b16_tau_eval_kernel<<<(unsigned)((n + threads - 1) / threads), threads, 0, (cudaStream_t)stream>>>(
      n, d_R, d_R_200c, d_M_200c, d_redshift, d_result, d_abserr, d_neval);
  AP_LAUNCH_CHECK();
  return 0;
}

__global__ void __launch_bounds__(128) spline_build_kernel(int64_t n, int32_t ns, const double* nodes,
                                                           const double* values, const int32_t* n_nodes, double* coef) {
  int64_t h = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (h >= n) return;
  double* c = coef + h * (int64_t)(ns - 1) * 4;
  if (n_nodes[h] == 0) {
    for (int i = 0; i < (ns - 1) * 4; ++i) c[i] = 0.0;
    return;
  }
  double x[kMaxNodes], y[kMaxNodes], cc[(kMaxNodes - 1) * 4];
  for (int i = 0; i < ns; ++i) {
    x[i] = nodes[h * ns + i];
    y[i] = values[h * ns + i];
  }
  spline_build_nak(ns, x, y, cc);
  for (int i = 0; i < (ns - 1) * 4; ++i) c[i] = cc[i];
}

// Uniform (np.linspace) nodes: the not-a-knot system has constant coefficients, so with NS a
// compile-time constant the Thomas multipliers fold to literals and every array lives in registers.
// A block builds 64 consecutive halos: node values are staged through shared memory (coalesced
// 120 B rows in, 448 B rows out) instead of 8 B accesses strided by the row length.
template <int NS>
__global__ void __launch_bounds__(64) spline_build_uniform_kernel(int64_t n, const double* __restrict__ nodes,
                                                                  const double* __restrict__ values,
                                                                  const int32_t* __restrict__ n_nodes,
                                                                  double* __restrict__ coef) {
  constexpr int NC = (NS - 1) * 4;
  __shared__ double sm[64 * NC];
  const int tid = threadIdx.x;
  const int64_t h0 = (int64_t)blockIdx.x * 64;
  const int nh = (int)min((int64_t)64, n - h0);
  for (int i = tid; i < nh * NS; i += 64) sm[i] = values[h0 * NS + i];
  __syncthreads();
  double y[NS], s[NS];
  const bool active = tid < nh && n_nodes[h0 + tid] != 0;
  double hstep = 1.0;
  if (tid < nh) {
#pragma unroll
    for (int i = 0; i < NS; ++i) y[i] = sm[tid * NS + i];
    if (active) hstep = (nodes[(h0 + tid) * NS + NS - 1] - nodes[(h0 + tid) * NS]) / (double)(NS - 1);
  }
  __syncthreads();
  if (tid < nh) {
    const double ih = 1.0 / hstep;
    double dl[NS - 1];
#pragma unroll
    for (int i = 0; i < NS - 1; ++i) dl[i] = (y[i + 1] - y[i]) * ih;
    // rows scaled by 1/h: [1 2 | (5 d0 + d1)/2], [1 4 1 | 3 (d_{i-1} + d_i)], [2 1 | (d_{n-3} + 5 d_{n-2})/2]
    double diag[NS], up[NS];
    s[0] = 0.5 * (5.0 * dl[0] + dl[1]);
    diag[0] = 1.0;
    up[0] = 2.0;
#pragma unroll
    for (int i = 1; i < NS - 1; ++i) {
      const double m = 1.0 / diag[i - 1];  // lower = 1
      diag[i] = 4.0 - m * up[i - 1];
      up[i] = 1.0;
      s[i] = 3.0 * (dl[i - 1] + dl[i]) - m * s[i - 1];
    }
    {
      const double m = 2.0 / diag[NS - 2];
      diag[NS - 1] = 1.0 - m * up[NS - 2];
      s[NS - 1] = 0.5 * (dl[NS - 3] + 5.0 * dl[NS - 2]) - m * s[NS - 2];
    }
    s[NS - 1] = s[NS - 1] / diag[NS - 1];
#pragma unroll
    for (int i = NS - 2; i >= 0; --i) s[i] = (s[i] - up[i] * s[i + 1]) / diag[i];
    double* out = sm + tid * NC;
#pragma unroll
    for (int i = 0; i < NS - 1; ++i) {
      out[4 * i + 0] = active ? y[i] : 0.0;
      out[4 * i + 1] = active ? s[i] : 0.0;
      out[4 * i + 2] = active ? (3.0 * dl[i] - 2.0 * s[i] - s[i + 1]) * ih : 0.0;
      out[4 * i + 3] = active ? (s[i] + s[i + 1] - 2.0 * dl[i]) * ih * ih : 0.0;
    }
  }
  __syncthreads();
  for (int i = tid; i < nh * NC; i += 64) coef[h0 * NC + i] = sm[i];
}

extern "C" int ap_spline_build_uniform(int64_t n, int32_t n_samples, const double* d_nodes, const double* d_values,
                                       const int32_t* d_n_nodes, double* d_coef, void* stream);

extern "C" int ap_spline_build(int64_t n, int32_t n_samples, const double* d_nodes, const double* d_values,
                               const int32_t* d_n_nodes, double* d_coef, void* stream) {
  if (n < 0) return fail("n < 0");
  if (n_samples < 4 || n_samples > kMaxNodes) return fail("n_samples must be in [4, 32]");
  if (n == 0) return 0;
  if (!d_nodes || !d_values || !d_n_nodes || !d_coef) return fail("spline_build: NULL pointer");
  int threads = 128;
  spline_build_kernel<<<(unsigned)((n + threads - 1) / threads), threads, 0, (cudaStream_t)stream>>>(
      n, n_samples, d_nodes, d_values, d_n_nodes, d_coef);
  AP_LAUNCH_CHECK();
  return 0;
}

extern "C" int ap_spline_build_uniform(int64_t n, int32_t n_samples, const double* d_nodes, const double* d_values,
                                       const int32_t* d_n_nodes, double* d_coef, void* stream) {
  if (n < 0) return fail("n < 0");
  if (n == 0) return 0;
  if (!d_nodes || !d_values || !d_n_nodes || !d_coef) return fail("spline_build_uniform: NULL pointer");
  unsigned blocks = (unsigned)((n + 63) / 64);
  if (n_samples == 15)
    spline_build_uniform_kernel<15><<<blocks, 64, 0, (cudaStream_t)stream>>>(n, d_nodes, d_values, d_n_nodes, d_coef);
  else if (n_samples == 20)
    spline_build_uniform_kernel<20><<<blocks, 64, 0, (cudaStream_t)stream>>>(n, d_nodes, d_values, d_n_nodes, d_coef);
  else
    return ap_spline_build(n, n_samples, d_nodes, d_values, d_n_nodes, d_coef, stream);  // generic solver
  AP_LAUNCH_CHECK();
  return 0;
}

// The `len(R) < n_samples` bypass of @interpolate (lib/utils.py:510-511): halos with 1 .. n_samples-1 pixels
// get the projected profile (one QUADPACK quadrature) at every pixel.  Work items come from K1
// (ap_disc_count_items: halo * 32 + pixel ordinal, appended on the device, their number in *n_items), so the
// host never reads a count back; the grid is sized for the SMs and strides over the list, one quadrature per
// lane.  Pixels outside the calling rank's ring band (ap_band_clip) belong to another rank and are skipped.
template <int KIND, class MapT>
__global__ void __launch_bounds__(128) paint_los_items_kernel(Hpx hpx, HalosDev H, const double* p0, const double* p1,
                                                              const double* p2, const double* scale,
                                                              const unsigned long long* n_items, int64_t item_cap,
                                                              const int64_t* items, MapT* map,
                                                              unsigned long long* n_painted) {
  fm_smem_init();
  const int64_t total = min((int64_t)*n_items, item_cap);
  unsigned painted = 0;
  for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x) {
    const int64_t h = items[t] >> 5;
    const int32_t want = (int32_t)(items[t] & 31);
    if (!AP_BOUNDS_OK(h >= 0 && h < H.n)) continue;
    DiscHeader d = disc_header(hpx, H.x[h], H.y[h], H.z[h], H.radius[h]);
    HaloCenter c = make_center(H.theta[h], H.phi[h], H.D_a[h], false);
    int32_t run = 0;
    for (int64_t iz = d.ir_first; iz <= d.ir_last; ++iz) {
      RingSpan s = ring_span(hpx, d, iz);
      if (s.len <= 0) continue;
      if (want < run + s.len) {
        if (iz < hpx.clip_lo || iz >= hpx.clip_hi) break;  // another rank's ring
        RingGeom g = ring_geom(hpx, iz, s, c);
        int32_t js = want - run;
        int32_t ip = s.ip_lo + js;
        double R = c.D_a * pixel_sep(g, js);
        double ctx[kCtx];
        los_setup<KIND>(p0[h], p1[h], p2 ? p2[h] : 0.0, ctx);
        double val = los_eval<KIND>(ctx, R) * (scale ? scale[h] : 1.0);
        const int64_t pix = s.startpix + (ip < 0 ? ip + s.nr : ip);
        if (AP_BOUNDS_OK(ip + s.nr >= 0 && ip < s.nr && pix >= hpx.clip_pix0 && pix < hpx.npix)) {
          atomicAdd(&map[pix - hpx.clip_pix0], (MapT)val);
          ++painted;
        }
        break;
      }
      run += s.len;
    }
  }
  painted = __reduce_add_sync(0xffffffffu, painted);
  if (n_painted && painted && (threadIdx.x & 31) == 0) atomicAdd(n_painted, (unsigned long long)painted);
}

template <int KIND>
static void launch_los_items(bool f32, unsigned blocks, cudaStream_t st, Hpx hpx, HalosDev H, const double* p0,
                             const double* p1, const double* p2, const double* scale, const unsigned long long* n_items,
                             int64_t item_cap, const int64_t* items, void* map, unsigned long long* n_painted) {
  if (f32)
    paint_los_items_kernel<KIND, float><<<blocks, 128, 0, st>>>(hpx, H, p0, p1, p2, scale, n_items, item_cap, items, (float*)map, n_painted);
  else
    paint_los_items_kernel<KIND, double><<<blocks, 128, 0, st>>>(hpx, H, p0, p1, p2, scale, n_items, item_cap, items, (double*)map, n_painted);
}

static int paint_los_items_impl(int kind, int64_t nside, const ap_halos* halos, const double* p0, const double* p1,
                                const double* p2, const double* scale, const unsigned long long* d_n_items,
                                int64_t item_cap, const int64_t* d_items, void* d_map, bool f32,
                                unsigned long long* d_n_painted, void* stream) {
  if (check_nside(nside)) return 1;
  if (check_halos(halos, true)) return 1;
  if (kind < 0 || kind > LOS_NFW_RHO) return fail("unknown LOS integrand kind");
  if (item_cap < 0) return fail("item_cap < 0");
  if (item_cap == 0 || halos->n == 0) return 0;
  if (!p0 || !p1 || (kind != LOS_NFW_RHO && !p2) || !d_n_items || !d_items || !d_map)
    return fail("paint_los_items: NULL pointer");
  if (rm_ensure_tables(stream)) return fail("could not initialise the quadrature tables");
  int64_t nb = (item_cap + 127) / 128;
  if (nb > 148 * 12) nb = 148 * 12;  // 12 resident blocks of 128 threads per SM: the list is walked grid-stride
  unsigned blocks = (unsigned)nb;
  Hpx hpx = disc_hpx(nside);
  HalosDev H = to_dev(halos);
  cudaStream_t st = (cudaStream_t)stream;
  switch (kind) {
    case LOS_B16_TAU: launch_los_items<LOS_B16_TAU>(f32, blocks, st, hpx, H, p0, p1, p2, scale, d_n_items, item_cap, d_items, d_map, d_n_painted); break;
    case LOS_B16_NE: launch_los_items<LOS_B16_NE>(f32, blocks, st, hpx, H, p0, p1, p2, scale, d_n_items, item_cap, d_items, d_map, d_n_painted); break;
    case LOS_B16_RHO_GAS: launch_los_items<LOS_B16_RHO_GAS>(f32, blocks, st, hpx, H, p0, p1, p2, scale, d_n_items, item_cap, d_items, d_map, d_n_painted); break;
    case LOS_NFW_RHO: launch_los_items<LOS_NFW_RHO>(f32, blocks, st, hpx, H, p0, p1, p2, scale, d_n_items, item_cap, d_items, d_map, d_n_painted); break;
  }
  AP_LAUNCH_CHECK();
  return 0;
}
extern "C" int ap_paint_los_items(int kind, int64_t nside, const ap_halos* halos, const double* d_p0, const double* d_p1,
                                  const double* d_p2, const double* d_scale, const unsigned long long* d_n_items,
                                  int64_t item_cap, const int64_t* d_items, double* d_map,
                                  unsigned long long* d_n_painted, void* stream) {
  return paint_los_items_impl(kind, nside, halos, d_p0, d_p1, d_p2, d_scale, d_n_items, item_cap, d_items, d_map, false,
                              d_n_painted, stream);
}
extern "C" int ap_paint_los_items_f32(int kind, int64_t nside, const ap_halos* halos, const double* d_p0,
                                      const double* d_p1, const double* d_p2, const double* d_scale,
                                      const unsigned long long* d_n_items, int64_t item_cap, const int64_t* d_items,
                                      float* d_map, unsigned long long* d_n_painted, void* stream) {
  return paint_los_items_impl(kind, nside, halos, d_p0, d_p1, d_p2, d_scale, d_n_items, item_cap, d_items, d_map, true,
                              d_n_painted, stream);
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
