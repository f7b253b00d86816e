// K3 / K1: the fused painter kernel template and its launcher (instantiated by k_paint.cu and k_paint_los.cu).
#pragma once
#include "internal.cuh"

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

// 1-D bulk copy (TMA, cp.async.bulk) global -> shared completing on an mbarrier: the spline coefficients of the halos
// a chunk touches are ONE contiguous run in HBM, so one elected thread moves them while the block computes ring spans
// (round 1 staged them through registers: a binary search, 7 LDG and 7 STS per thread and chunk, 13% of K3's stall
// samples in the ncu source page)
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void bulk_copy_g2s(void* smem_dst, const void* gmem_src, uint32_t bytes, unsigned long long* bar) {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // the previous chunk's generic-proxy reads of the buffer
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(smem_u32(smem_dst)), "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, uint32_t parity) {
  asm volatile(
      "{\n\t"
      ".reg .pred P1;\n\t"
      "WAIT_LOOP:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n\t"
      "@P1 bra WAIT_DONE;\n\t"
      "bra WAIT_LOOP;\n\t"
      "WAIT_DONE:\n\t"
      "}" ::"r"(smem_u32(bar)), "r"(parity)
      : "memory");
}

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
  int32_t stage;    // coefficients may be staged with bulk copies (n_samples <= 15, 16-byte aligned d_coef)
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
  __align__(16) double stage[P == P_TABLE ? kStageHalos * (kStageMaxNodes - 1) * 4 : 2];
  unsigned long long stage_bar;       // mbarrier of the bulk copy into `stage`
  int32_t stage_lo, stage_ok;         // first staged halo of the chunk; every halo of the chunk is staged
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
// Four resident blocks per SM (64 registers) for every template whose pixel loop is a few dozen fp64 operations:
// measured on the bench catalog, P_TABLE runs 7.3 ms per 1e7 halos at 64 registers, 8.2 ms at 80 and 10.1 ms at the
// 103 ptxas takes when left alone (the loop is issue-bound and needs the warps).  The per-pixel quadrature templates
// get two blocks (128 registers): they are bound by the QUADPACK rule, not by this kernel's bookkeeping.
template <int P>
constexpr int paint_min_blocks() {
  return (P == P_B16_TAU2D || P == P_B16_RHOGAS2D || P == P_NFW_RHO2D_LOS) ? 2 : 4;
}
// EMIT: deterministic mode -- painted values are appended to the (key, value) list of ap_emit_begin instead of being
// added atomically; a separate instantiation, so that the atomic kernel carries neither the branch nor its registers
// (as a run-time flag it cost K3 0.7 ms per 1e7 halos and a 32-byte spill)
template <int P, class MapT, bool INCL, bool EMIT>
__global__ void __launch_bounds__(kPaintThreads, paint_min_blocks<P>()) paint_kernel(Hpx hpx, HalosDev H, ParamPtrs pp, TableArgs tab,
                                                               StatsArgs st, int halos_per_block,
                                                               MapT* __restrict__ map,
                                                               unsigned long long* n_painted, EmitArgs em) {
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
  if (P == P_TABLE && tid == 0) mbar_init(&sm.stage_bar, 1);
  __syncthreads();
  // exclusive scan of the halos' ring counts (G <= 112: four warps): warp scans, then the warp totals
  int32_t ring_incl = 0;
  if (tid < 128) {
    int32_t n = 0;
    if (tid < G) {
      n = sm.hdr[tid].ir_last - sm.hdr[tid].ir_first + 1;
      n = n > 0 ? n : 0;
    }
    ring_incl = n;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      int32_t t = __shfl_up_sync(0xffffffffu, ring_incl, o);
      if (lane >= o) ring_incl += t;
    }
    if (lane == 31) sm.warp_sum[warp] = ring_incl;
  }
  __syncthreads();
  if (tid < G) {
    int32_t add = 0;
    for (int w = 0; w < warp; ++w) add += sm.warp_sum[w];
    sm.ring_off[tid + 1] = ring_incl + add;
  }
  if (tid == 0) sm.ring_off[0] = 0;
  __syncthreads();
  const int32_t T = sm.ring_off[G];
  unsigned long long painted = 0;
  uint32_t stage_phase = 0;

  // gridDim.y blocks share the halos of blockIdx.x: block y takes every gridDim.y-th chunk of 256 ring items, so a
  // catalog of few, sky-sized discs (BASELINE config 1) is not serialised on the block that owns the largest disc
  for (int32_t base = (int32_t)blockIdx.y * kPaintThreads; base < T; base += kPaintThreads * (int32_t)gridDim.y) {
    // P_TABLE: the spline coefficients of the halos this chunk touches are one contiguous run in HBM: one thread
    // starts a bulk copy of the first kStageHalos of them into shared memory now; it lands while phase 1 runs
    const bool do_stage = (P == P_TABLE) && tab.stage != 0;
    if (do_stage && tid == 0) {
      int lo = 0, hi = G;   // halo of the chunk's first ring item
      while (hi - lo > 1) {
        int mid = (lo + hi) >> 1;
        if (sm.ring_off[mid] <= base) lo = mid; else hi = mid;
      }
      const int32_t last_it = min(base + kPaintThreads, T) - 1;
      int lo2 = lo, hi2 = G;  // halo of its last ring item
      while (hi2 - lo2 > 1) {
        int mid = (lo2 + hi2) >> 1;
        if (sm.ring_off[mid] <= last_it) lo2 = mid; else hi2 = mid;
      }
      sm.stage_lo = lo;
      sm.stage_ok = (lo2 - lo) < kStageHalos;
      const int nc = (tab.n_samples - 1) * 4;
      bulk_copy_g2s(sm.stage, tab.coef + (h0 + lo) * (int64_t)nc, (uint32_t)(min(kStageHalos, G - lo) * nc * 8), &sm.stage_bar);
    }

    // phase 1: one ring item per thread
    int32_t it = base + tid;
    int32_t len = 0;
    // "small" ring item: every pixel has |d / 2| < 1/8 and a haversine below 1e-4 (theta < 0.02 rad) and, for the
    // tabulated template, its halo's spline sits in the staged shared-memory copy.  When ALL items of the chunk
    // are small, phase 2 runs a straight-line pixel loop (no series / table-path branches).
    bool small_ok = !VEC && !STATS;
    int st_key = -1;                                             // K1: halo of this thread's ring item
    unsigned long long st_min = 0x7ff0000000000000ull, st_max = 0ull;   // +inf / 0: neutral for min / max of haversines
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
          st_min = (unsigned long long)__double_as_longlong(a);
          st_max = (unsigned long long)__double_as_longlong(b);
        } else {
          r.startpix = s.startpix; r.nr = s.nr; r.ip_lo = s.ip_lo; r.hl = lo;
          r.A = g.A; r.B = g.B; r.phi_step = g.phi_step; r.d0 = g.d0;
          if (VEC) { sm.itemv[tid].phi0 = g.phi0; sm.itemv[tid].z = g.z; sm.itemv[tid].sth = g.sth; }
          const double dmax = fmax(fabs(g.d0), fabs(g.d0 + (double)(s.len - 1) * g.phi_step));
          small_ok = small_ok && dmax < 0.25 && (g.A + g.B * (0.25 * dmax * dmax)) < 1.0e-4;   // sin^2(x) <= x^2
          if (P == P_TABLE) small_ok = small_ok && do_stage && tab.uniform;
        }
      }
      if (STATS) st_key = lo;
    }
    if (STATS) {
      // The ring items of one halo sit on consecutive lanes: reduce (count, min, max) over each run of equal halo
      // with shuffles and let the run's first lane update the halo's shared accumulators -- one set of atomics per
      // halo and warp instead of three per ring item (the contended 64-bit shared atomics were 23% of K1's samples)
      unsigned long long cnt = len > 0 ? (unsigned long long)len : 0ull;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const int k2 = __shfl_down_sync(0xffffffffu, st_key, o);
        const unsigned long long c2 = __shfl_down_sync(0xffffffffu, cnt, o);
        const unsigned long long mn2 = __shfl_down_sync(0xffffffffu, st_min, o);
        const unsigned long long mx2 = __shfl_down_sync(0xffffffffu, st_max, o);
        if (lane + o < 32 && k2 == st_key) {
          cnt += c2;
          st_min = mn2 < st_min ? mn2 : st_min;
          st_max = mx2 > st_max ? mx2 : st_max;
        }
      }
      const int kprev = __shfl_up_sync(0xffffffffu, st_key, 1);
      if (st_key >= 0 && (lane == 0 || kprev != st_key) && cnt > 0ull) {
        atomicAdd(&sm.cnt[st_key], cnt);
        if (st.rmin) {
          atomicMin(&sm.hmin[st_key], st_min);
          atomicMax(&sm.hmax[st_key], st_max);
        }
      }
      continue;  // counts and ranges only: no pixel walk (no barrier needed either)
    }

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
    }
    // (the barrier inside the scan above made thread 0's stage_lo / stage_ok visible)
    const int stage_lo = do_stage ? sm.stage_lo : 0;
    const int stage_n = do_stage ? min(kStageHalos, G - stage_lo) * (tab.n_samples - 1) * 4 : 0;
    if (P == P_TABLE) small_ok = small_ok && (!do_stage || sm.stage_ok != 0);
    const bool all_small = __syncthreads_and(small_ok ? 1 : 0) != 0;
    if (do_stage) {   // the coefficients have landed (phase flips every chunk)
      mbar_wait(&sm.stage_bar, stage_phase);
      stage_phase ^= 1u;
    }
    const int32_t total = sm.pix_off[kPaintThreads];
    painted += (tid == 0) ? (unsigned long long)total : 0ull;
    const bool fastp = !VEC && all_small && total <= kPixMapMax;

    // phase 2: two pixels per thread per iteration (k and k + 256): independent dependency chains
    // (table loads, series) overlap, then both REDs are issued
    auto eval = [&](int32_t k, int64_t& pix, int& hl_out) -> double {
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
      if (EMIT) hl_out = hl;
      return val;
    };
    // the same arithmetic with every data-dependent branch resolved for the whole chunk (small discs: the common case)
    auto eval_small = [&](int32_t k, int64_t& pix, int& hl_out) -> double {
      const int lo = sm.pixmap[k];
      const RingItem& r = sm.item[lo];
      const int32_t j = k - sm.pix_off[lo];
      const int hl = r.hl;
      const double hav = r.A + r.B * sin2_half_small(r.d0 + (double)j * r.phi_step);
      const double R = sm.cen_Da[hl] * sep_from_hav_small(hav);
      double val;
      if (P == P_TABLE) {
        const int ns = tab.n_samples;
        int i = (int)((R - sm.ctx[hl][1]) * sm.ctx[hl][2]);
        i = i < 0 ? 0 : (i > ns - 2 ? ns - 2 : i);
        const double t = R - (sm.ctx[hl][1] + (double)i * sm.ctx[hl][3]);
        const double2* c2 = reinterpret_cast<const double2*>(sm.stage + ((hl - stage_lo) * (ns - 1) + i) * 4);
        const double2 c01 = c2[0], c23 = c2[1];
        val = (c01.x + t * (c01.y + t * (c23.x + t * c23.y))) * sm.ctx[hl][0];
      } else {
        val = profile_eval_R<P>(sm.ctx[hl], R);
      }
      const int32_t ip = r.ip_lo + j;
      pix = r.startpix + (ip < 0 ? ip + r.nr : ip);
      if (EMIT) hl_out = hl;
      return val;
    };
    for (int32_t k = tid; k < total; k += 2 * kPaintThreads) {
      int64_t p0, p1 = 0;
      int g0 = 0, g1 = 0;
      const bool two = (k + kPaintThreads < total);
      double v0, v1;
      if (fastp) {
        v0 = eval_small(k, p0, g0);
        v1 = two ? eval_small(k + kPaintThreads, p1, g1) : 0.0;
      } else {
        v0 = eval(k, p0, g0);
        v1 = two ? eval(k + kPaintThreads, p1, g1) : 0.0;
      }
      if (AP_BOUNDS_OK(p0 >= hpx.clip_pix0 && p0 < hpx.clip_pix1)) accumulate<EMIT>(hpx, em, map, p0, v0, h0 + g0);
      if (two && AP_BOUNDS_OK(p1 >= hpx.clip_pix0 && p1 < hpx.clip_pix1)) accumulate<EMIT>(hpx, em, map, p1, v1, h0 + g1);
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
  // few blocks (small catalog): split each block's ring items over up to 8 blocks (K1 keeps its per-block sums)
  const dim3 nb((unsigned)blocks, (P != P_STATS && blocks <= 4096) ? 8u : 1u);
  cudaStream_t cs = (cudaStream_t)stream;
  const HalosDev hd = to_dev(halos);
  const EmitArgs em = emit_args();
  if (em.keys && P != P_STATS) {   // deterministic mode: the map is not touched by this launch
    if (hpx.incl_fct)
      paint_kernel<P, double, true, true><<<nb, kPaintThreads, 0, cs>>>(hpx, hd, pp, tab, st, g, (double*)d_map, d_n_painted, em);
    else
      paint_kernel<P, double, false, true><<<nb, kPaintThreads, 0, cs>>>(hpx, hd, pp, tab, st, g, (double*)d_map, d_n_painted, em);
  } else if (f32 && P != P_STATS) {
    if (hpx.incl_fct)
      paint_kernel<P, float, true, false><<<nb, kPaintThreads, 0, cs>>>(hpx, hd, pp, tab, st, g, (float*)d_map, d_n_painted, em);
    else
      paint_kernel<P, float, false, false><<<nb, kPaintThreads, 0, cs>>>(hpx, hd, pp, tab, st, g, (float*)d_map, d_n_painted, em);
  } else {
    if (hpx.incl_fct)
      paint_kernel<P, double, true, false><<<nb, kPaintThreads, 0, cs>>>(hpx, hd, pp, tab, st, g, (double*)d_map, d_n_painted, em);
    else
      paint_kernel<P, double, false, false><<<nb, kPaintThreads, 0, cs>>>(hpx, hd, pp, tab, st, g, (double*)d_map, d_n_painted, em);
  }
  AP_LAUNCH_CHECK();
  return 0;
}

// the per-pixel quadrature templates (P_B16_TAU2D, P_B16_RHOGAS2D, P_NFW_RHO2D_LOS) are instantiated in k_paint_los.cu
int launch_paint_los(int profile, int64_t nside, const ap_halos* halos, const ParamPtrs& pp, const TableArgs& tab,
                     void* d_map, unsigned long long* d_n_painted, void* stream, bool f32);
