// astropaint_b200 — translation unit 3 of 4: K3 for the templates that run one QUADPACK quadrature per PIXEL
// (Battaglia16.tau_2D, rho_gas_2D, NFW.rho_2D: direct @LOS_integrate templates).
#include "paint_kernel.cuh"

AP_TU_BOUNDS_GETTER(ap_bounds_paint_los)

int launch_paint_los(int profile, int64_t nside, const ap_halos* halos, const ParamPtrs& pp, const TableArgs& tab,
                     void* d_map, unsigned long long* d_n_painted, void* stream, bool f32) {
  switch (profile) {
    case P_B16_TAU2D: return launch_paint<P_B16_TAU2D>(nside, halos, pp, tab, d_map, d_n_painted, stream, nullptr, f32);
    case P_B16_RHOGAS2D: return launch_paint<P_B16_RHOGAS2D>(nside, halos, pp, tab, d_map, d_n_painted, stream, nullptr, f32);
    case P_NFW_RHO2D_LOS: return launch_paint<P_NFW_RHO2D_LOS>(nside, halos, pp, tab, d_map, d_n_painted, stream, nullptr, f32);
  }
  return fail("not a per-pixel quadrature profile");
}
