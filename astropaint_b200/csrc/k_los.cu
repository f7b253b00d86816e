// astropaint_b200 — translation unit 4 of 4: K4 (@interpolate tables: nodes, QUADPACK node values, splines) and
// the `len(R) < n_samples` bypass.
#include "internal.cuh"

AP_TU_BOUNDS_GETTER(ap_bounds_los)

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

// 6 blocks of 128 threads per SM (80 registers, 24 warps).  Round 1 (per-thread los_setup inside the kernel) measured
// 176.1 / 173.5 / 180.9 / 182.9 ms per 1e7 halos at 7 / 6 / 5 / 4 blocks (72 / 80 / 96 / 128 registers).  Round 2, with
// the per-halo constants in their own kernel: timed ALONE for 30 ms bursts the 64-register build wins (135.8 ms at 8
// blocks against 138.7 at 6, 137.3 at 9, 136.6 at 10, 135.6 as 4 blocks of 256; profiles/r2_runs/r2_variants*.log), but in the
// sustained step its 32 warps per SM draw enough power for the 1 kW cap to pull the SM clock from 1965 to ~1870 MHz
// (nvidia-smi: sw_power_cap active) and it LOSES: 141.3 ms against 138.5 ms, same box, back to back, twice
// (profiles/r2_runs/r2_mb_ab.log).  (A pure-DFMA probe holds 1965 MHz for seconds, so it is the kernel's memory-side activity --
// table look-ups, local-memory frames, register-file traffic of 32 warps -- that costs the power, not the fp64 pipe.)
#ifndef AP_QAGI_THREADS
#define AP_QAGI_THREADS 128
#endif
#ifndef AP_QAGI_MINBLOCKS
#define AP_QAGI_MINBLOCKS 6
#endif
// per-halo constants of the 3-D profile (los_setup: eight pow() for the Battaglia fit and the critical density):
// once per halo, one halo per lane, instead of once per (halo, node) thread of the node kernel
template <int KIND>
__global__ void __launch_bounds__(128) los_setup_kernel(int64_t n, const double* p0, const double* p1, const double* p2,
                                                        const int32_t* n_nodes, double4* ctx_out) {
  int64_t h = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (h >= n || n_nodes[h] == 0) return;
  double ctx[kCtx];
  los_setup<KIND>(p0[h], p1[h], p2 ? p2[h] : 0.0, ctx);
  ctx_out[h] = make_double4(ctx[0], ctx[1], ctx[2], ctx[3]);
}

template <int KIND, int FAST>
__global__ void __launch_bounds__(AP_QAGI_THREADS, AP_QAGI_MINBLOCKS) los_nodes_kernel(int64_t n, const double* p0, const double* p1,
                                                                           const double* p2, const double4* halo_ctx, int32_t ns,
                                                                           const double* nodes, const int32_t* n_nodes,
                                                                           double* values) {
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
  if (halo_ctx) {
    const double2* c = reinterpret_cast<const double2*>(halo_ctx + h);
    const double2 c01 = __ldg(c), c23 = __ldg(c + 1);
    ctx[0] = c01.x; ctx[1] = c01.y; ctx[2] = c23.x; ctx[3] = c23.y;
  } else {
    los_setup<KIND>(p0[h], p1[h], p2 ? p2[h] : 0.0, ctx);
  }
  values[t] = los_eval<KIND, FAST>(ctx, nodes[t]);
}

extern "C" int ap_los_nodes(int kind, int64_t n, const double* d_p0, const double* d_p1, const double* d_p2,
                            int32_t n_samples, const double* d_nodes, const int32_t* d_n_nodes, double* d_values,
                            double* d_halo_scratch, void* stream) {
  if (n < 0) return fail("n < 0");
  if (kind < 0 || kind > LOS_NFW_RHO) return fail("unknown LOS integrand kind");
  if (d_halo_scratch && (reinterpret_cast<uintptr_t>(d_halo_scratch) & 31)) return fail("los_nodes: d_halo_scratch must be 32-byte aligned");
  if (n == 0) return 0;
  if (!d_p0 || !d_p1 || (kind != LOS_NFW_RHO && !d_p2) || !d_nodes || !d_n_nodes || !d_values)
    return fail("los_nodes: NULL pointer");
  if (rm_ensure_tables(stream)) return fail("could not initialise the quadrature tables");
  int threads = AP_QAGI_THREADS;
  int64_t total = n * n_samples;
  unsigned blocks = (unsigned)((total + threads - 1) / threads);
  cudaStream_t st = (cudaStream_t)stream;
  const double4* hc = reinterpret_cast<const double4*>(d_halo_scratch);
#define AP_LOS_NODES(K)                                                                                                     \
  if (hc)                                                                                                                   \
    los_setup_kernel<K><<<(unsigned)((n + 127) / 128), 128, 0, st>>>(n, d_p0, d_p1, d_p2, d_n_nodes,                        \
                                                                     reinterpret_cast<double4*>(d_halo_scratch));           \
  if (quad_fast_enabled())                                                                                                  \
    los_nodes_kernel<K, 1><<<blocks, threads, 0, st>>>(n, d_p0, d_p1, d_p2, hc, n_samples, d_nodes, d_n_nodes, d_values);   \
  else                                                                                                                      \
    los_nodes_kernel<K, 0><<<blocks, threads, 0, st>>>(n, d_p0, d_p1, d_p2, hc, n_samples, d_nodes, d_n_nodes, d_values)
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
template <int KIND, class MapT, bool EMIT>
__global__ void __launch_bounds__(128, 6) paint_los_items_kernel(Hpx hpx, HalosDev H, const double* p0, const double* p1,
                                                              const double* p2, const double* scale,
                                                              const unsigned long long* n_items, int64_t item_cap,
                                                              const int64_t* items, MapT* map,
                                                              unsigned long long* n_painted, EmitArgs em) {
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
        if (AP_BOUNDS_OK(ip + s.nr >= 0 && ip < s.nr && pix >= hpx.clip_pix0 && pix < hpx.clip_pix1)) {
          accumulate<EMIT>(hpx, em, map, pix, val, h);
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
  const EmitArgs em = emit_args();
  if (em.keys)
    paint_los_items_kernel<KIND, double, true><<<blocks, 128, 0, st>>>(hpx, H, p0, p1, p2, scale, n_items, item_cap, items, (double*)map, n_painted, em);
  else if (f32)
    paint_los_items_kernel<KIND, float, false><<<blocks, 128, 0, st>>>(hpx, H, p0, p1, p2, scale, n_items, item_cap, items, (float*)map, n_painted, em);
  else
    paint_los_items_kernel<KIND, double, false><<<blocks, 128, 0, st>>>(hpx, H, p0, p1, p2, scale, n_items, item_cap, items, (double*)map, n_painted, em);
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
// host-buffer entry for the tabulated templates (@interpolate(@LOS_integrate(profile_3D))): the whole K1 -> nodes ->
// QUADPACK -> spline -> K3 -> bypass pipeline of Painter.spray for Battaglia16.tau_2D_interp / kSZ_T / ne_2D and
// NFW.rho_2D_interp, host pointers in and out (the table counterpart of ap_spray_host)
// ------------------------------------------------------------------------------------------
extern "C" int ap_spray_table_host(int kind, int64_t nside, int64_t n, const double* h_x, const double* h_y,
                                   const double* h_z, const double* h_radius, const double* h_theta, const double* h_phi,
                                   const double* h_D_a, const double* h_p0, const double* h_p1, const double* h_p2,
                                   const double* h_scale, int32_t n_samples, int32_t method, double* h_map,
                                   uint64_t* h_n_painted) {
  if (check_nside(nside)) return 1;
  if (n < 0) return fail("n < 0");
  if (kind < 0 || kind > LOS_NFW_RHO) return fail("unknown LOS integrand kind");
  if (n_samples < 4 || n_samples > kMaxNodes) return fail("n_samples must be in [4, 32]");
  if (method != 0 && method != 1) return fail("method must be 0 (linspace) or 1 (logspace)");
  if (!h_map) return fail("h_map == NULL");
  const int64_t npix = 12 * nside * nside;
  const double* cols[11] = {h_x, h_y, h_z, h_radius, h_theta, h_phi, h_D_a, h_p0, h_p1, h_p2, h_scale};
  const int n_required = (kind == LOS_NFW_RHO) ? 9 : 10;
  for (int i = 0; i < n_required && n > 0; ++i)
    if (!cols[i]) return fail("spray_table_host: NULL halo column");
  const int64_t nn = n > 0 ? n : 1, ns = n_samples;
  // one device allocation carved into 32-byte aligned pieces
  struct Piece { void** ptr; size_t bytes; };
  double *d_cols[11] = {0}, *d_map = nullptr, *d_rmin = nullptr, *d_rmax = nullptr, *d_nodes = nullptr, *d_values = nullptr,
         *d_coef = nullptr, *d_ctx = nullptr;
  int64_t *d_n_pix = nullptr, *d_items = nullptr;
  int32_t* d_n_nodes = nullptr;
  unsigned long long* d_cnt = nullptr;  // [0] painted pairs, [1] bypass items
  const int64_t item_cap = nn * (ns - 1);
  Piece pieces[] = {{(void**)&d_map, sizeof(double) * (size_t)npix}, {(void**)&d_rmin, 8 * (size_t)nn},
                    {(void**)&d_rmax, 8 * (size_t)nn}, {(void**)&d_nodes, 8 * (size_t)(nn * ns)},
                    {(void**)&d_values, 8 * (size_t)(nn * ns)}, {(void**)&d_coef, 8 * (size_t)(nn * (ns - 1) * 4)},
                    {(void**)&d_ctx, 32 * (size_t)nn}, {(void**)&d_n_pix, 8 * (size_t)nn},
                    {(void**)&d_items, 8 * (size_t)item_cap}, {(void**)&d_n_nodes, 4 * (size_t)nn},
                    {(void**)&d_cnt, 16}};
  size_t total = 0;
  for (auto& p : pieces) total += (p.bytes + 255) & ~(size_t)255;
  for (int i = 0; i < 11; ++i) total += (8 * (size_t)nn + 255) & ~(size_t)255;
  char* base = nullptr;
  cudaStream_t st = nullptr;
  auto cleanup = [&]() {
    if (base) cudaFree(base);
    if (st) cudaStreamDestroy(st);
  };
#define AP_TRY(call)                                                     \
  do {                                                                   \
    cudaError_t e_ = (call);                                             \
    if (e_ != cudaSuccess) {                                             \
      cleanup();                                                         \
      return fail(std::string(#call) + ": " + cudaGetErrorString(e_));   \
    }                                                                    \
  } while (0)
#define AP_RC(call)        \
  do {                     \
    if ((call) != 0) {     \
      cleanup();           \
      return 1;            \
    }                      \
  } while (0)
  AP_TRY(cudaStreamCreate(&st));
  AP_TRY(cudaMalloc(&base, total));
  size_t off = 0;
  for (auto& p : pieces) {
    *p.ptr = base + off;
    off += (p.bytes + 255) & ~(size_t)255;
  }
  for (int i = 0; i < 11; ++i) {
    d_cols[i] = reinterpret_cast<double*>(base + off);
    off += (8 * (size_t)nn + 255) & ~(size_t)255;
    if (cols[i] && n > 0) AP_TRY(cudaMemcpyAsync(d_cols[i], cols[i], 8 * (size_t)n, cudaMemcpyHostToDevice, st));
  }
  AP_TRY(cudaMemcpyAsync(d_map, h_map, sizeof(double) * (size_t)npix, cudaMemcpyHostToDevice, st));
  AP_TRY(cudaMemsetAsync(d_cnt, 0, 16, st));
  if (n > 0) {
    ap_halos H;
    H.n = n;
    H.d_x = d_cols[0]; H.d_y = d_cols[1]; H.d_z = d_cols[2]; H.d_radius = d_cols[3];
    H.d_theta = d_cols[4]; H.d_phi = d_cols[5]; H.d_D_a = d_cols[6];
    const double* p2 = h_p2 ? d_cols[9] : nullptr;
    const double* sc = h_scale ? d_cols[10] : nullptr;
    AP_RC(ap_disc_count_items(nside, &H, d_n_pix, d_rmin, d_rmax, n_samples, d_items, item_cap, d_cnt + 1, st));
    AP_RC(ap_table_nodes(n, d_n_pix, d_rmin, d_rmax, n_samples, method, d_nodes, d_n_nodes, st));
    AP_RC(ap_los_nodes(kind, n, d_cols[7], d_cols[8], p2, n_samples, d_nodes, d_n_nodes, d_values, d_ctx, st));
    if (method == 0)
      AP_RC(ap_spline_build_uniform(n, n_samples, d_nodes, d_values, d_n_nodes, d_coef, st));
    else
      AP_RC(ap_spline_build(n, n_samples, d_nodes, d_values, d_n_nodes, d_coef, st));
    AP_RC(ap_paint_table(nside, &H, n_samples, method == 0, d_nodes, d_coef, d_n_nodes, sc, d_map, d_cnt, st));
    AP_RC(ap_paint_los_items(kind, nside, &H, d_cols[7], d_cols[8], p2, sc, d_cnt + 1, item_cap, d_items, d_map, d_cnt, st));
  }
  AP_TRY(cudaMemcpyAsync(h_map, d_map, sizeof(double) * (size_t)npix, cudaMemcpyDeviceToHost, st));
  unsigned long long cnt = 0;
  AP_TRY(cudaMemcpyAsync(&cnt, d_cnt, sizeof(cnt), cudaMemcpyDeviceToHost, st));
  AP_TRY(cudaStreamSynchronize(st));
  if (h_n_painted) *h_n_painted = cnt;
  cleanup();
  return 0;
#undef AP_TRY
#undef AP_RC
}
