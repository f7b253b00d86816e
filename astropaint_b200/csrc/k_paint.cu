// astropaint_b200 — translation unit 2 of 4: K1 (disc counts / R range / bypass list) and K3 (fused painter) for
// the closed-form templates and the tabulated (@interpolate) template.
#include "paint_kernel.cuh"

AP_TU_BOUNDS_GETTER(ap_bounds_paint)

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
    explicit ModeGuard(int inc) : prev(ap_disc_mode(-1)) { if (inc && prev == 0) ap_disc_mode(4); }
    ~ModeGuard() { ap_disc_mode(prev); }
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
    case P_B16_TAU2D: return launch_paint_los(profile, nside, halos, pp, tab, d_map, d_n_painted, stream, f32);
    case P_B16_RHOGAS2D: return launch_paint_los(profile, nside, halos, pp, tab, d_map, d_n_painted, stream, f32);
    case P_NFW_RHO2D_LOS: return launch_paint_los(profile, nside, halos, pp, tab, d_map, d_n_painted, stream, f32);
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
  tab.stage = (n_samples <= kStageMaxNodes && (reinterpret_cast<uintptr_t>(d_coef) & 15) == 0) ? 1 : 0;
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
