/* astropaint_b200 — C ABI of the B200-native painting hot path.
 *
 * The reference (syasini/AstroPaint) has no FFI: the path sits behind a plain Python object
 * API (Painter.spray / Canvas.stack_cutouts).  These entry points are what a maintainer binds
 * (ctypes stub in INTEGRATION.md) to replace the per-halo Python/healpy/numpy loop.  Each one
 * cites the reference lines it replaces (paths relative to the reference repo).
 *
 * Conventions: every `const double*` / `int64_t*` argument named d_* is a DEVICE pointer owned
 * by the caller; h_* is a HOST pointer.  `stream` is a cudaStream_t passed as void* (NULL =
 * default stream); all device entry points are asynchronous on it.  Return value 0 = OK,
 * non-zero = error (ap_last_error() gives the message).  Nothing here falls back to the CPU.
 */
#ifndef ASTROPAINT_B200_H
#define ASTROPAINT_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* profile ids for ap_paint_profile (the built-in templates as device functions) */
enum ap_profile {
  AP_CONSTANT = 0,       /* profiles/spherical.py:44-59   constant_density_2D(R, constant) */
  AP_LINEAR = 1,         /* profiles/spherical.py:62-79   linear_density_2D(R, intercept, slope) */
  AP_SOLID_SPHERE = 2,   /* profiles/spherical.py:25-42   solid_sphere_2D(R, M_200c, R_200c) */
  AP_GAUSSIAN = 3,       /* README.md:99-101              gaussian(R, R_200c, x, y, z) */
  AP_NFW_RHO2D = 4,      /* profiles/NFW.py:65-80         rho_2D_bartlemann(R, rho_s, R_s) */
  AP_NFW_TAU2D = 5,      /* profiles/NFW.py:153-169       tau_2D(R, rho_s, R_s) */
  AP_NFW_KSZ = 6,        /* profiles/NFW.py:172-178       kSZ_T(R, rho_s, R_s, v_r, T_cmb) */
  AP_NFW_DEFLECTION = 7, /* profiles/NFW.py:111-150       deflection_angle(R, c_200c, R_200c, M_200c) */
  AP_NFW_BG = 8,         /* profiles/NFW.py:181-193       BG(R_vec, c_200c, R_200c, M_200c, theta, phi, v_th, v_ph, T_cmb) */
  AP_B16_TAU2D = 9,      /* profiles/Battaglia16.py:221-239 tau_2D(R, R_200c, M_200c, redshift) */
  AP_B16_RHOGAS2D = 10,  /* profiles/Battaglia16.py:181-198 rho_gas_2D(R, R_200c, M_200c, redshift) */
  AP_NFW_RHO2D_LOS = 11  /* profiles/NFW.py:83-93         rho_2D(R, rho_s, R_s) = LOS_integrate(rho_3D) */
};

/* 3-D profiles projected by @LOS_integrate on device (one QUADPACK qagie per line of sight) */
enum ap_los_kind {
  AP_LOS_B16_TAU = 0,      /* Battaglia16.tau_3D(r, R_200c, M_200c, redshift)   :132-155 */
  AP_LOS_B16_NE = 1,       /* Battaglia16.ne_3D                                  :97-129  */
  AP_LOS_B16_RHO_GAS = 2,  /* Battaglia16.rho_gas_3D                             :65-94   */
  AP_LOS_NFW_RHO = 3       /* NFW.rho_3D(r, rho_s, r_s)                          NFW.py:38-58 */
};

/* Per-halo geometry inputs shared by all entry points (catalog columns, paint_bucket.py:308-364):
 *   d_x, d_y, d_z   raw Cartesian position [Mpc]  -> disc centre of hp.query_disc (:1222-1225)
 *   d_radius        R_times * arcmin2rad(R_ang_200c) [rad], computed on the host (:1226-1227)
 *   d_theta, d_phi  catalog (theta, phi) -> centre of hp.rotator.angdist (:1202-1203, :1264)
 *   d_D_a           angular-diameter distance [Mpc] (:1273)                                  */
typedef struct ap_halos {
  int64_t n;
  const double *d_x, *d_y, *d_z, *d_radius, *d_theta, *d_phi, *d_D_a;
} ap_halos;

const char* ap_last_error(void);
int ap_version(void);
/* Number of out-of-range indices seen by a -DAP_DEBUG_BOUNDS build (tools/debug_bounds.sh); -1 for a
 * normal build.                                                                               */
long long ap_debug_bounds_errors(void);

/* Catalog.build_dataframe (paint_bucket.py:308-364, calculate_redshifts=False): the derived columns of the
 * halo catalog from x, y, z [Mpc], v_x, v_y, v_z [km/s], M_200c [M_sun] and an optional redshift column
 * (d_redshift == NULL -> default_redshift for every halo).  crit_dens_0 [M_sun / Mpc^3] and h are the
 * cosmology constants the reference's lib/transform.py:30-44 takes from astropy.  h_out is a HOST array of
 * AP_CATALOG_N_OUT DEVICE pointers (n doubles each) in the order of the AP_CAT_* indices below.            */
enum {
  AP_CAT_D_C = 0, AP_CAT_REDSHIFT, AP_CAT_LAT, AP_CAT_LON, AP_CAT_THETA, AP_CAT_PHI, AP_CAT_D_A, AP_CAT_R_200C,
  AP_CAT_C_200C, AP_CAT_R_ANG_200C, AP_CAT_RHO_S, AP_CAT_R_S, AP_CAT_V_R, AP_CAT_V_TH, AP_CAT_V_PH, AP_CAT_V_LAT,
  AP_CAT_V_LON, AP_CATALOG_N_OUT
};
int ap_catalog_build(int64_t n, const double* d_x, const double* d_y, const double* d_z, const double* d_vx,
                     const double* d_vy, const double* d_vz, const double* d_M_200c, const double* d_redshift,
                     double default_redshift, double crit_dens_0, double h, double* const* h_out, void* stream);

/* ap_los_nodes integrates with the rightmost-bisection fast path of dqagie (bit-identical to the general
 * routine whenever it accepts a quadrature, general routine otherwise).  enable = 0 forces the general routine
 * (used by the parity tests), 1 restores the default, negative only queries.  Returns the previous setting.  */
int ap_quad_fast_path(int enable);

/* Disc-query mode of the calling thread, honoured by every entry point below that runs a disc query
 * (count, pixels, paint, LOS bypass): 0 = hp.query_disc(..., inclusive=False), the Canvas default
 * (paint_bucket.py:782, :1228); f > 0 = inclusive=True with healpy's oversampling factor f (healpy's own
 * default is 4, which is what Canvas(inclusive=True) gets).  Returns the previous mode; a negative
 * argument only queries.                                                                        */
int ap_disc_mode(int inclusive_fact);

/* d_radius[i] = R_times * np.deg2rad(d_R_ang_200c[i] / 60): the query_disc radius of
 * Canvas.Disc.gen_pixel_index (paint_bucket.py:1226-1227, lib/transform.py:148-150), rounded exactly
 * as numpy rounds it (IEEE divide, multiply, multiply; no reciprocal, no FMA).                 */
int ap_disc_radius(int64_t n, const double* d_R_ang_200c, double R_times, double* d_radius, void* stream);

/* hp.query_disc per halo (Canvas.Disc.gen_pixel_index, paint_bucket.py:1215-1229), counts only.
 * d_n_pix[n] (required), d_n_rings[n] (optional).  If d_rmin/d_rmax are non-NULL they receive
 * min/max of R = D_a * angdist over each halo's pixels (gen_cent2pix_mpc, :1266-1273), i.e.
 * R.min()/R.max() of utils.sample_array (lib/utils.py:411-412); +inf/-inf for empty discs.
 * inclusive != 0: healpy's inclusive query (factor 4) for this call even if ap_disc_mode is 0. */
int ap_disc_count(int64_t nside, const ap_halos* halos, int inclusive, int64_t* d_n_pix,
                  int32_t* d_n_rings, double* d_rmin, double* d_rmax, void* stream);

/* Ring-band ownership of the map (the multi-GPU form of the reference's shared accumulator, paint_bucket.py:
 * 2414-2433: joblib threads adding into ONE array).  After ap_band_clip(ring_lo, ring_hi) every painting entry
 * point called by this thread (ap_paint_profile, ap_paint_table, ap_paint_los_items and their _f32 forms) paints
 * only rings ring_lo <= iz < ring_hi (1-based HEALPix ring numbers) and treats d_map as the address of the first
 * pixel of ring ring_lo, i.e. of a buffer holding just that band; ap_disc_count keeps whole-disc counts but
 * reports 0 pixels for a halo whose disc misses the band.  A halo that straddles a boundary is painted by both
 * owners, each keeping its own rings, so the bands of all ranks tile the map with no reduction.
 * ap_band_clip(0, 0) restores the whole map.                                                     */
int ap_band_clip(int64_t ring_lo, int64_t ring_hi);

/* ap_disc_count that also builds the work list of the `len(R) < n_samples` bypass of utils.interpolate
 * (lib/utils.py:510-511) on the device: every halo with 0 < n_pix < n_samples appends one item
 * (halo * 32 + pixel ordinal, 0 <= ordinal < n_pix, ring by ring) per pixel at
 * d_items[atomicAdd(d_item_count, n_pix) ...]; items beyond item_cap are dropped (size the list
 * n * (n_samples - 1) to make that impossible).  *d_item_count must be zeroed by the caller.   */
int ap_disc_count_items(int64_t nside, const ap_halos* halos, int64_t* d_n_pix, double* d_rmin, double* d_rmax,
                        int32_t n_samples, int64_t* d_items, int64_t item_cap, unsigned long long* d_item_count,
                        void* stream);

/* Deterministic accumulation (the sort-then-segmented-reduce form of np.add.at, paint_bucket.py:2387-2389).  Between
 * ap_emit_begin and ap_emit_end every painting entry point called by this thread (ap_paint_profile, ap_paint_table,
 * ap_paint_los_items, _f32 forms) appends each painted value to d_keys / d_vals instead of adding it atomically:
 * key = pixel << shift | label, label = d_labels[halo] (or the halo's position when d_labels is NULL; must be below
 * 2^shift).  *d_count (zeroed by the caller) receives the number of pairs; pairs beyond cap are dropped.  The caller
 * sorts the keys (values along) and calls ap_segment_accumulate, which adds every pixel's values to d_map in label
 * order -- the reference's serial summation order -- with plain stores; d_map addresses pixel pix0 (the band's first
 * pixel under ap_band_clip, else 0).  Results are bit-reproducible from run to run.                          */
int ap_emit_begin(uint64_t* d_keys, double* d_vals, unsigned long long* d_count, int64_t cap, const int64_t* d_labels,
                  int32_t shift);
int ap_emit_end(void);
int ap_segment_accumulate(int64_t n, const uint64_t* d_keys_sorted, const double* d_vals_sorted, int32_t shift,
                          int64_t pix0, void* d_map, int f32, void* stream);

/* Near-tie audit of the disc query (hp.query_disc, paint_bucket.py:1222).  The span boundaries are floor() of
 * expressions holding atan2 / cos, where CUDA's libm and the CPU's may differ in the last bit: a boundary can flip
 * only if its floor argument lies within ~1e-11 of an integer.  ap_disc_near_ties appends every halo that has a
 * boundary within `tol` pixels of an integer to d_flagged (their number in *d_n_flagged, which the caller zeroes;
 * d_min_margin[n], optional, receives each halo's smallest margin); ap_query_disc_host recomputes discs on the HOST
 * with the CPU's libm (count and sum of pixel ids per disc; the pixel list of the first disc, optional) so the
 * caller can compare.  inclusive_fact as in ap_disc_mode.                                                    */
int ap_disc_near_ties(int64_t nside, const ap_halos* halos, double tol, unsigned long long* d_n_flagged,
                      int64_t* d_flagged, int64_t cap, double* d_min_margin, void* stream);
int ap_query_disc_host(int64_t nside, int64_t n, const double* h_x, const double* h_y, const double* h_z,
                       const double* h_radius, int inclusive_fact, int64_t* h_count, int64_t* h_pixsum,
                       int64_t* h_pix_first, int64_t pix_cap);

/* Flat (halo, pixel) work list: pixel ids of halo h in ascending order at
 * d_pix[d_pix_off[h] .. d_pix_off[h+1]) — identical to np.concatenate of gen_pixel_index().
 * Optional outputs (NULL to skip): d_R[total] = gen_cent2pix_mpc (:1266-1273),
 * d_R_vec[total*3] = gen_cent2pix_mpc_vec (:1283-1290).                                       */
int ap_disc_pixels(int64_t nside, const ap_halos* halos, const int64_t* d_pix_off, int64_t* d_pix,
                   double* d_R, double* d_R_vec, void* stream);

/* np.add.at(canvas.pixels, pixel_index, values) for a flat work list (paint_bucket.py:2387-2389). */
int ap_scatter_add(int64_t n, const int64_t* d_pix, const double* d_val, double* d_map, void* stream);

/* Painter.spray for a built-in template (paint_bucket.py:2378-2389 fused: disc query, pixel
 * geometry, template, scatter-add).  h_params[i] is the DEVICE pointer of the i-th template
 * argument (reference signature order, R/R_vec excluded); h_strides[i] is 1 for a per-halo
 * array, 0 for a single broadcast value.  d_map is the fp64 HEALPix RING map (12*nside^2),
 * accumulated in place.  d_n_painted (optional, device uint64) is incremented by the number of
 * (halo, pixel) pairs painted.                                                               */
int ap_paint_profile(int profile, int64_t nside, const ap_halos* halos, const double* const* h_params,
                     const int32_t* h_strides, int32_t n_params, double* d_map,
                     unsigned long long* d_n_painted, void* stream);

/* @interpolate support (lib/utils.py:457-521).  Nodes: for halos with n_pix >= n_samples,
 * sample_array(R, n_samples, method) (:406-425; method 0 = linspace, 1 = logspace with
 * eps = 1e-3) into d_nodes[n*n_samples] and d_n_nodes[h] = n_samples; else d_n_nodes[h] = 0
 * (the `len(R) < n_samples` bypass, :510-511).                                               */
int ap_table_nodes(int64_t n, const int64_t* d_n_pix, const double* d_rmin, const double* d_rmax,
                   int32_t n_samples, int32_t method, double* d_nodes, int32_t* d_n_nodes, void* stream);

/* @LOS_integrate at the nodes: one QUADPACK qagie(epsabs=0, epsrel=1e-2) per (halo, node)
 * (lib/utils.py:445-448).  d_p0, d_p1, d_p2 are the 3-D profile's per-halo arguments in reference
 * order (R_200c, M_200c, redshift | rho_s, r_s, NULL).  Serves Battaglia16.tau_2D_interp / kSZ_T (n=15,
 * :242-271), Battaglia16.ne_2D (interpolate(n=20) of LOS_integrate(ne_3D), :201-218) and
 * NFW.rho_2D_interp (n=20, logspace; NFW.py:96-108).  d_halo_scratch (optional, 4 * n doubles, 32-byte
 * aligned): when given, the per-halo constants of the profile are formed once per halo by a small setup
 * kernel instead of once per (halo, node) thread.                                               */
int ap_los_nodes(int kind, int64_t n, const double* d_p0, const double* d_p1, const double* d_p2, int32_t n_samples,
                 const double* d_nodes, const int32_t* d_n_nodes, double* d_values, double* d_halo_scratch,
                 void* stream);

/* Diagnostic: Battaglia16.tau_2D at arbitrary (R, halo) pairs with QUADPACK's abserr and neval
 * (scipy.integrate.quad full_output; lib/utils.py:448).  d_abserr / d_neval may be NULL.       */
int ap_b16_tau_eval(int64_t n, const double* d_R, const double* d_R_200c, const double* d_M_200c,
                    const double* d_redshift, double* d_result, double* d_abserr, int32_t* d_neval,
                    void* stream);

/* InterpolatedUnivariateSpline(nodes, values, k=3) per halo -> piecewise-cubic coefficients
 * d_coef[n*(n_samples-1)*4] (lib/utils.py:519).                                              */
int ap_spline_build(int64_t n, int32_t n_samples, const double* d_nodes, const double* d_values,
                    const int32_t* d_n_nodes, double* d_coef, void* stream);

/* Same result for np.linspace nodes (ap_table_nodes method 0): constant-coefficient solve kept in
 * registers for n_samples 15 and 20, other sizes fall through to ap_spline_build.             */
int ap_spline_build_uniform(int64_t n, int32_t n_samples, const double* d_nodes, const double* d_values,
                            const int32_t* d_n_nodes, double* d_coef, void* stream);

/* Painter.spray for a tabulated template: value = spline_h(R) * d_scale[h] (d_scale optional).
 * uniform_nodes = 1 when the nodes are linspace (method 0 of ap_table_nodes).
* Halos with d_n_nodes[h] == 0 are skipped (bypass halos are painted by ap_paint_los_items or
 * through the flat work list).                                                               */
int ap_paint_table(int64_t nside, const ap_halos* halos, int32_t n_samples, int32_t uniform_nodes,
                   const double* d_nodes, const double* d_coef, const int32_t* d_n_nodes, const double* d_scale,
                   double* d_map, unsigned long long* d_n_painted, void* stream);

/* The `len(R) < n_samples` branch of utils.interpolate (lib/utils.py:510-511) for the @LOS_integrate kinds:
 * item t of the list ap_disc_count_items built paints the (d_items[t] & 31)-th pixel (ring by ring) of halo
 * d_items[t] >> 5 with the projected profile (one QUADPACK quadrature) times d_scale[halo].  The number of items
 * is read on the device (*d_n_items, clamped to item_cap): no host synchronisation.            */
int ap_paint_los_items(int kind, int64_t nside, const ap_halos* halos, const double* d_p0, const double* d_p1,
                       const double* d_p2, const double* d_scale, const unsigned long long* d_n_items,
                       int64_t item_cap, const int64_t* d_items, double* d_map, unsigned long long* d_n_painted,
                       void* stream);

/* Canvas.cutouts / stack_cutouts (paint_bucket.py:1601-1743, 1745-1941): healpy
 * CartesianProj(lonra, latra, xsize, ysize).projmap(map, rot=(lon, lat), vec2pix) per halo; the image
 * grid is healpy's endpoint-inclusive one (ij2xy: x_j = -lon1 + j (lon1 - lon0) / (xsize - 1), astro flip).
 * d_lon_deg/d_lat_deg[n_sel] are the catalog lon/lat (degrees) of the selected halos.
 * ap_cutouts writes d_out[n_sel*ysize*xsize]; ap_stack_cutouts accumulates
 * d_stack[ysize*xsize] += sum_h weight_h * cutout_h (d_weight optional).                    */
int ap_cutouts(int64_t nside, const double* d_map, int64_t n_sel, const double* d_lon_deg,
               const double* d_lat_deg, double lon0, double lon1, double lat0, double lat1,
               int32_t xsize, int32_t ysize, double* d_out, void* stream);
int ap_stack_cutouts(int64_t nside, const double* d_map, int64_t n_sel, const double* d_lon_deg,
                     const double* d_lat_deg, double lon0, double lon1, double lat0, double lat1,
                     int32_t xsize, int32_t ysize, const double* d_weight, double* d_stack, void* stream);

/* fp32-accumulate mode (BASELINE north_star: 1e-4 parity bar): the same kernels accumulating into a
 * float32 map (templates are still evaluated in fp64; the value is rounded once before the RED).  */
int ap_scatter_add_f32(int64_t n, const int64_t* d_pix, const double* d_val, float* d_map, void* stream);
int ap_paint_profile_f32(int profile, int64_t nside, const ap_halos* halos, const double* const* h_params,
                         const int32_t* h_strides, int32_t n_params, float* d_map,
                         unsigned long long* d_n_painted, void* stream);
int ap_paint_table_f32(int64_t nside, const ap_halos* halos, int32_t n_samples, int32_t uniform_nodes,
                       const double* d_nodes, const double* d_coef, const int32_t* d_n_nodes, const double* d_scale,
                       float* d_map, unsigned long long* d_n_painted, void* stream);
int ap_paint_los_items_f32(int kind, int64_t nside, const ap_halos* halos, const double* d_p0, const double* d_p1,
                           const double* d_p2, const double* d_scale, const unsigned long long* d_n_items,
                           int64_t item_cap, const int64_t* d_items, float* d_map, unsigned long long* d_n_painted,
                           void* stream);

/* Host-buffer convenience entry (what bench.py's e2e leg and INTEGRATION.md's stub call):
 * copies the halo columns and parameters host->device, paints into a device map initialised
 * from h_map, and copies the map back.  All pointers are HOST pointers.                      */
int ap_spray_host(int profile, int64_t nside, int64_t n, const double* h_x, const double* h_y,
                  const double* h_z, const double* h_radius, const double* h_theta, const double* h_phi,
                  const double* h_D_a, const double* const* h_params, const int32_t* h_strides,
                  int32_t n_params, double* h_map, uint64_t* h_n_painted);

/* The same for the tabulated templates -- @interpolate(@LOS_integrate(profile_3D)): Battaglia16.tau_2D_interp / kSZ_T
 * (kind AP_LOS_B16_TAU, n_samples 15, method 0), Battaglia16.ne_2D (AP_LOS_B16_NE, 20, 0), NFW.rho_2D_interp
 * (AP_LOS_NFW_RHO, 20, method 1 = logspace) -- i.e. Painter.spray of BASELINE configs 2 and 3 with HOST buffers in and
 * out: K1, nodes, QUADPACK, spline, K3 and the `len(R) < n_samples` bypass.  h_p0..h_p2: the 3-D profile's per-halo
 * arguments (h_p2 NULL for NFW); h_scale (optional): per-halo multiplier (kSZ: -v_r / c * T_cmb * (1 + z)).      */
int ap_spray_table_host(int kind, int64_t nside, int64_t n, const double* h_x, const double* h_y, const double* h_z,
                        const double* h_radius, const double* h_theta, const double* h_phi, const double* h_D_a,
                        const double* h_p0, const double* h_p1, const double* h_p2, const double* h_scale,
                        int32_t n_samples, int32_t method, double* h_map, uint64_t* h_n_painted);

/* Host side of the map (Canvas.pixels is a host ndarray in the reference, paint_bucket.py:798, 858-869): the
 * ranks of one node write their bands into ONE host array.  ap_host_register page-locks a caller-owned host
 * range (e.g. the rank's band of a shared mapping) so that ap_copy_d2h_async / ap_copy_h2d_async on it are
 * asynchronous DMA on `stream`; ap_host_unregister undoes it.                                              */
int ap_host_register(void* h_ptr, uint64_t bytes);
int ap_host_unregister(void* h_ptr);
int ap_copy_d2h_async(void* h_dst, const void* d_src, uint64_t bytes, void* stream);
int ap_copy_h2d_async(void* d_dst, const void* h_src, uint64_t bytes, void* stream);

/* Measurement aid (bench.py): launches `blocks` x 256 threads each issuing 8 x iters dependent-chain DFMAs from
 * registers; *h_n_dfma receives the number of thread-level DFMAs.  Timed with CUDA events it gives the fp64
 * instruction rate this GPU sustains -- the roofline of the QUADPACK node kernel.                           */
int ap_fp64_peak(int32_t iters, int32_t blocks, double* d_out, uint64_t* h_n_dfma, void* stream);

/* ---- per-cutout operators: Canvas.cutouts / stack_cutouts `apply_func` (paint_bucket.py:1729-1742) on a batch
 * of device-resident cutouts d_patches[n][ysize][xsize] (what ap_cutouts writes).                          */

/* In-place cubic B-spline prefilter of every cutout (scipy.ndimage.spline_filter, order 3, the filter
 * scipy.ndimage.rotate applies first for mode='constant': mirror extension, axis 0 then axis 1).           */
int ap_patch_spline_prefilter(int64_t n, int32_t ysize, int32_t xsize, double* d_patches, void* stream);

/* transform.rotate_patch (lib/transform.py:266-275) = scipy.ndimage.rotate(patch, angle, reshape=False) on
 * prefiltered coefficients d_coef: out = spline(M out_index + offset), M = [[cos, sin], [-sin, cos]] with
 * d_cos/d_sin[n] = scipy.special.cosdg/sindg(angle) per cutout, 0 outside the cutout.  Optionally fused:
 * taper (d_win_y[ysize], d_win_x[xsize], both or neither), per-cutout d_weight[n] (NULL = 1), and with
 * stack != 0 the result is ACCUMULATED into d_out[ysize][xsize] instead of stored to d_out[n][ysize][xsize]. */
int ap_patch_rotate(int64_t n, int32_t ysize, int32_t xsize, const double* d_coef, const double* d_cos,
                    const double* d_sin, const double* d_win_y, const double* d_win_x, const double* d_weight,
                    int stack, double* d_out, void* stream);

/* transform.taper_patch (lib/transform.py:278-283: patch *= outer(hanning, hanning)) and/or per-cutout weights
 * without rotation: in place when d_stack == NULL, otherwise accumulated into d_stack[ysize][xsize].          */
int ap_patch_scale(int64_t n, int32_t ysize, int32_t xsize, double* d_patches, const double* d_win_y,
                   const double* d_win_x, const double* d_weight, double* d_stack, void* stream);

/* Canvas.ud_grade (paint_bucket.py:2233-2257) = healpy.ud_grade(map, nside_out, pess, power) for RING maps:
 * degrade = mean over the good (not UNSEEN, finite) NEST children of each parent (UNSEEN if none is good, or if
 * any is bad when pess != 0), upgrade = the parent's value; times (nside_out / nside_in)**power when
 * use_power != 0.  Both nside powers of two.  d_map_out holds 12 nside_out^2 doubles.                        */
int ap_ud_grade(int64_t nside_in, const double* d_map_in, int64_t nside_out, int pess, double power, int use_power,
                double* d_map_out, void* stream);

/* hp.ang2pix(nside, theta, phi) (paint_bucket.py:996 find_centers_indx, :1183 gen_center_ipix) and
 * hp.pix2ang(nside, ipix) (:1085, :1199 gen_center_ang(snap2pixel=True), :1238 gen_pixel_ang), RING scheme.
 * theta outside [0, pi] gives pixel -1 (healpy raises); a pixel outside [0, npix) gives NaN angles.        */
int ap_ang2pix(int64_t nside, int64_t n, const double* d_theta, const double* d_phi, int64_t* d_pix, void* stream);
int ap_pix2ang(int64_t nside, int64_t n, const int64_t* d_pix, double* d_theta, double* d_phi, void* stream);

#ifdef __cplusplus
}
#endif
#endif
