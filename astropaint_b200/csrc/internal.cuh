// Shared host-side plumbing of the CUDA translation units (k_core.cu, k_paint.cu, k_paint_los.cu, k_los.cu):
// error reporting, the thread-local disc-query / band-clip modes, halo-pointer checks.  The library is built
// from several translation units compiled in parallel (__graft_entry__.build()); no device symbol crosses a
// translation unit.
#pragma once
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

// Deterministic accumulation (ap_emit_begin): instead of an atomic add, every painted (pixel, value) pair is
// appended to a list as key = pixel << shift | halo label; the caller sorts the keys and sums each pixel's values in
// halo order (ap_segment_accumulate).  keys == NULL = atomics.
struct EmitArgs {
  unsigned long long* keys;
  double* vals;
  unsigned long long* count;
  long long cap;
  const long long* labels;
  int32_t shift;
};
EmitArgs emit_args();                    // the calling thread's ap_emit_begin state

int fail(const std::string& m);          // records the message for ap_last_error(), returns 1
Hpx disc_hpx(int64_t nside);             // make_hpx + the calling thread's inclusive / band-clip modes
int quad_fast_enabled();                 // ap_quad_fast_path setting

#define AP_LAUNCH_CHECK()                                                               \
  do {                                                                                  \
    cudaError_t e_ = cudaGetLastError();                                                \
    if (e_ != cudaSuccess) return fail(std::string("launch: ") + cudaGetErrorString(e_)); \
  } while (0)

// Optional bounds instrumentation (-DAP_DEBUG_BOUNDS, tools/debug_bounds.sh): compute-sanitizer is not
// available on the B200 pool, so a debug build counts every out-of-range index instead of making the
// access; ap_debug_bounds_errors() must read 0 after the GPU test-suite.  One counter per translation unit.
static __device__ unsigned long long g_bounds_errors = 0ull;
#ifdef AP_DEBUG_BOUNDS
#define AP_BOUNDS_OK(cond) ((cond) ? true : (atomicAdd(&g_bounds_errors, 1ull), false))
#else
#define AP_BOUNDS_OK(cond) true
#endif
// One painted value: an fp64 (or fp32) RED into the rank's band of the map, or -- in deterministic mode -- an
// append to the (key, value) list with one warp-aggregated slot allocation per converged group of lanes.
#if defined(__CUDACC__)
template <bool EMIT, class MapT>
__device__ __forceinline__ void accumulate(const Hpx& hpx, const EmitArgs& em, MapT* map, int64_t pix, double val,
                                           int64_t halo) {
  if (EMIT) {
    const unsigned mask = __activemask();
    const int lane = threadIdx.x & 31;
    const int leader = __ffs(mask) - 1;
    unsigned long long base = 0;
    if (lane == leader) base = atomicAdd(em.count, (unsigned long long)__popc(mask));
    base = __shfl_sync(mask, base, leader);
    const unsigned long long slot = base + __popc(mask & ((1u << lane) - 1u));
    if ((long long)slot < em.cap) {
      const unsigned long long label = (unsigned long long)(em.labels ? em.labels[halo] : halo);
      em.keys[slot] = ((unsigned long long)pix << em.shift) | label;
      em.vals[slot] = val;
    }
  } else {
    atomicAdd(&map[pix - hpx.clip_pix0], (MapT)val);
  }
}
#endif

#define AP_TU_BOUNDS_GETTER(name)                                                          \
  long long name() {                                                                       \
    unsigned long long v = 0;                                                              \
    if (cudaMemcpyFromSymbol(&v, g_bounds_errors, sizeof(v)) != cudaSuccess) return -2;    \
    return (long long)v;                                                                   \
  }
long long ap_bounds_core();
long long ap_bounds_paint();
long long ap_bounds_paint_los();
long long ap_bounds_los();

struct HalosDev {
  int64_t n;
  const double *x, *y, *z, *radius, *theta, *phi, *D_a;
};
static inline HalosDev to_dev(const ap_halos* h) {
  HalosDev d;
  d.n = h->n;
  d.x = h->d_x; d.y = h->d_y; d.z = h->d_z; d.radius = h->d_radius;
  d.theta = h->d_theta; d.phi = h->d_phi; d.D_a = h->d_D_a;
  return d;
}

static inline int check_halos(const ap_halos* h, bool need_center) {
  if (!h) return fail("halos == NULL");
  if (h->n < 0) return fail("halos.n < 0");
  if (h->n > 0 && (!h->d_x || !h->d_y || !h->d_z || !h->d_radius)) return fail("halos: x/y/z/radius NULL");
  if (need_center && h->n > 0 && (!h->d_theta || !h->d_phi || !h->d_D_a)) return fail("halos: theta/phi/D_a NULL");
  return 0;
}
static inline int check_nside(int64_t nside) {
  if (nside < 1 || nside > (int64_t(1) << 29)) return fail("nside out of range");
  // ring pixel counts and ring numbers are held in int32 inside the kernels
  if (nside > 8192 * 32) return fail("nside > 262144 not supported by the int32 ring arithmetic");
  return 0;
}
