// Per-cutout operators of Canvas.cutouts / Canvas.stack_cutouts (`apply_func`, paint_bucket.py:1729-1742)
// on batches of device-resident cutouts [n][ysize][xsize] (fp64):
//
//   transform.rotate_patch  lib/transform.py:266-275  = scipy.ndimage.rotate(patch, angle, reshape=False)
//                           (order 3, mode 'constant', cval 0, prefilter=True)
//   transform.taper_patch   lib/transform.py:278-283  = patch *= outer(hanning(n), hanning(n))
//   per-cutout weights (the sign flip of the oriented-stacking notebooks) and the sum over cutouts.
//
// scipy is a dependency of the reference and IS installed here, so these are pinned against the real
// thing (tests/test_gpu_parity.py::test_patch_ops_match_scipy).  The spline code restates scipy's
// ndimage/src/ni_splines.c + ni_interpolation.c (BSD) algorithm:
//   prefilter: per axis, c *= 6; causal pass c[i] += z c[i-1] with the MIRROR initialisation
//              c0 = (sum_{i<n-1} z^i (c[i] + z^(n-1) c[n-1-i]) ...) / (1 - z^(2n-2)), anticausal pass
//              c[i] = z (c[i+1] - c[i]) from c[n-1] = (z c[n-2] + c[n-1]) z / (z^2 - 1), z = sqrt(3) - 2
//              (mode 'constant' filters with mirror extension and no padding);
//   sampling:  in = M out + offset per axis; outside [0, n-1] -> cval; else 4 x 4 cubic B-spline taps from
//              start = floor(in) - 1 with mirror indexing.
#pragma once

namespace patchops {

constexpr double kPole = -0.267949192431122706472553658494127633;  // sqrt(3) - 2

// one thread per line; `stride` = distance between consecutive elements of a line
__device__ __forceinline__ void spline_filter_line(double* c, int n, int64_t stride) {
  if (n < 2) return;
  const double z = kPole;
  const double gain = (1.0 - z) * (1.0 - 1.0 / z);
  for (int i = 0; i < n; ++i) c[i * stride] *= gain;
  // _init_causal_mirror
  double z_i = z;
  const double z_n_1 = pow(z, (double)(n - 1));
  double c0 = c[0] + z_n_1 * c[(int64_t)(n - 1) * stride];
  for (int i = 1; i < n - 1; ++i) {
    c0 += z_i * (c[i * stride] + z_n_1 * c[(int64_t)(n - 1 - i) * stride]);
    z_i *= z;
  }
  c0 /= 1.0 - z_n_1 * z_n_1;
  c[0] = c0;
  double prev = c0;
  for (int i = 1; i < n; ++i) {
    prev = c[i * stride] + z * prev;
    c[i * stride] = prev;
  }
  // _init_anticausal_mirror
  double last = (z * c[(int64_t)(n - 2) * stride] + c[(int64_t)(n - 1) * stride]) * z / (z * z - 1.0);
  c[(int64_t)(n - 1) * stride] = last;
  for (int i = n - 2; i >= 0; --i) {
    last = z * (last - c[i * stride]);
    c[i * stride] = last;
  }
}

// axis 1 (rows of length xsize, contiguous) then axis 0 (columns): scipy filters axis 0 first, then axis 1;
// the two passes commute only up to rounding, so keep scipy's order: AXIS = 0 first.
template <int AXIS>
__global__ void __launch_bounds__(128) spline_prefilter_kernel(int64_t n, int32_t ysize, int32_t xsize, double* p) {
  const int64_t lines_per = AXIS == 0 ? xsize : ysize;
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n * lines_per) return;
  const int64_t patch = t / lines_per, line = t - patch * lines_per;
  double* base = p + patch * (int64_t)ysize * xsize;
  if (AXIS == 0)
    spline_filter_line(base + line, ysize, xsize);       // a column: consecutive threads, consecutive columns
  else
    spline_filter_line(base + line * (int64_t)xsize, xsize, 1);
}

__device__ __forceinline__ void bspline3_weights(double x, double* w) {
  x -= floor(x);
  const double y = x, z = 1.0 - x;
  w[1] = (y * y * (y - 2.0) * 3.0 + 4.0) / 6.0;
  w[2] = (z * z * (z - 2.0) * 3.0 + 4.0) / 6.0;
  w[0] = z * z * z / 6.0;
  w[3] = 1.0 - w[0] - w[1] - w[2];
}

__device__ __forceinline__ int mirror_index(int i, int n) {
  if (n == 1) return 0;
  const int period = 2 * (n - 1);
  if (i < 0) i = -i;
  i %= period;
  return i >= n ? period - i : i;
}

// out[p][i][j] = spline(coef[p])(M (i, j) + offset), M = [[c, s], [-s, c]] (scipy.ndimage.rotate, axes (1, 0),
// reshape=False): offset = centre - M centre, centre = ((ysize - 1) / 2, (xsize - 1) / 2).  Optionally fused:
// * taper (outer Hanning window), * weight[p], and accumulation into a stack instead of a store.
template <bool STACK>
__global__ void __launch_bounds__(256) patch_rotate_kernel(int64_t n, int32_t ysize, int32_t xsize,
                                                           const double* __restrict__ coef, const double* cosd,
                                                           const double* sind, const double* win_y, const double* win_x,
                                                           const double* weight, double* out) {
  const int32_t npx = ysize * xsize;
  const int32_t s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= npx) return;
  const int i = s / xsize, j = s - i * xsize;
  const double cy = (ysize - 1) / 2.0, cx = (xsize - 1) / 2.0;
  const double taper = (win_y ? win_y[i] * win_x[j] : 1.0);
  const int64_t p0 = (int64_t)blockIdx.y * 64;
  const int64_t p1 = p0 + 64 < n ? p0 + 64 : n;
  double acc = 0.0;
  for (int64_t p = p0; p < p1; ++p) {
    const double c = cosd[p], sn = sind[p];
    // offset = in_center - M out_center (scipy rotate); in = M out + offset, same evaluation order as
    // NI_GeometricTransform: sum_k matrix[h][k] * out[k], then + shift
    const double off_y = cy - (c * cy + sn * cx), off_x = cx - (-sn * cy + c * cx);
    const double yy = (c * (double)i + sn * (double)j) + off_y;
    const double xx = (-sn * (double)i + c * (double)j) + off_x;
    double v = 0.0;
    if (!(yy < 0.0 || yy > (double)(ysize - 1) || xx < 0.0 || xx > (double)(xsize - 1))) {
      double wy[4], wx[4];
      bspline3_weights(yy, wy);
      bspline3_weights(xx, wx);
      const int sy = (int)floor(yy) - 1, sx = (int)floor(xx) - 1;
      const double* cp = coef + p * (int64_t)npx;
      int ix[4];
#pragma unroll
      for (int b = 0; b < 4; ++b) ix[b] = mirror_index(sx + b, xsize);
#pragma unroll
      for (int a = 0; a < 4; ++a) {
        const double* row = cp + (int64_t)mirror_index(sy + a, ysize) * xsize;
#pragma unroll
        for (int b = 0; b < 4; ++b) v += __ldg(row + ix[b]) * (wy[a] * wx[b]);
      }
    }
    v *= taper;
    if (weight) v *= weight[p];
    if (STACK)
      acc += v;
    else
      out[p * (int64_t)npx + s] = v;
  }
  if (STACK) atomicAdd(&out[s], acc);
}

// in place: patch *= taper (outer window, optional) * weight[p] (optional); or, STACK, accumulate the result
template <bool STACK>
__global__ void __launch_bounds__(256) patch_scale_kernel(int64_t n, int32_t ysize, int32_t xsize, double* patches,
                                                          const double* win_y, const double* win_x, const double* weight,
                                                          double* stack) {
  const int32_t npx = ysize * xsize;
  const int32_t s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= npx) return;
  const int i = s / xsize, j = s - i * xsize;
  const double taper = (win_y ? win_y[i] * win_x[j] : 1.0);
  const int64_t p0 = (int64_t)blockIdx.y * 64;
  const int64_t p1 = p0 + 64 < n ? p0 + 64 : n;
  double acc = 0.0;
  for (int64_t p = p0; p < p1; ++p) {
    double v = patches[p * (int64_t)npx + s] * taper;
    if (weight) v *= weight[p];
    if (STACK)
      acc += v;
    else
      patches[p * (int64_t)npx + s] = v;
  }
  if (STACK) atomicAdd(&stack[s], acc);
}

}  // namespace patchops

static int patch_args(int64_t n, int32_t ysize, int32_t xsize, const void* p) {
  if (n < 0) return fail("n < 0");
  if (ysize < 1 || xsize < 1 || (int64_t)ysize * xsize > (int64_t(1) << 28)) return fail("patch size out of range");
  if (n > 0 && !p) return fail("patch buffer NULL");
  return 0;
}

extern "C" int ap_patch_spline_prefilter(int64_t n, int32_t ysize, int32_t xsize, double* d_patches, void* stream) {
  if (patch_args(n, ysize, xsize, d_patches)) return 1;
  if (n == 0) return 0;
  cudaStream_t st = (cudaStream_t)stream;
  int64_t l0 = n * xsize, l1 = n * ysize;
  patchops::spline_prefilter_kernel<0><<<(unsigned)((l0 + 127) / 128), 128, 0, st>>>(n, ysize, xsize, d_patches);
  patchops::spline_prefilter_kernel<1><<<(unsigned)((l1 + 127) / 128), 128, 0, st>>>(n, ysize, xsize, d_patches);
  AP_LAUNCH_CHECK();
  return 0;
}

extern "C" int ap_patch_rotate(int64_t n, int32_t ysize, int32_t xsize, const double* d_coef, const double* d_cos,
                               const double* d_sin, const double* d_win_y, const double* d_win_x, const double* d_weight,
                               int stack, double* d_out, void* stream) {
  if (patch_args(n, ysize, xsize, d_coef)) return 1;
  if (n == 0) return 0;
  if (!d_cos || !d_sin || !d_out) return fail("patch_rotate: NULL pointer");
  if ((d_win_y == nullptr) != (d_win_x == nullptr)) return fail("patch_rotate: give both windows or neither");
  dim3 grid((unsigned)(((int64_t)ysize * xsize + 255) / 256), (unsigned)((n + 63) / 64));
  cudaStream_t st = (cudaStream_t)stream;
  if (stack)
    patchops::patch_rotate_kernel<true><<<grid, 256, 0, st>>>(n, ysize, xsize, d_coef, d_cos, d_sin, d_win_y, d_win_x, d_weight, d_out);
  else
    patchops::patch_rotate_kernel<false><<<grid, 256, 0, st>>>(n, ysize, xsize, d_coef, d_cos, d_sin, d_win_y, d_win_x, d_weight, d_out);
  AP_LAUNCH_CHECK();
  return 0;
}

extern "C" int ap_patch_scale(int64_t n, int32_t ysize, int32_t xsize, double* d_patches, const double* d_win_y,
                              const double* d_win_x, const double* d_weight, double* d_stack, void* stream) {
  if (patch_args(n, ysize, xsize, d_patches)) return 1;
  if (n == 0) return 0;
  if ((d_win_y == nullptr) != (d_win_x == nullptr)) return fail("patch_scale: give both windows or neither");
  dim3 grid((unsigned)(((int64_t)ysize * xsize + 255) / 256), (unsigned)((n + 63) / 64));
  cudaStream_t st = (cudaStream_t)stream;
  if (d_stack)
    patchops::patch_scale_kernel<true><<<grid, 256, 0, st>>>(n, ysize, xsize, d_patches, d_win_y, d_win_x, d_weight, d_stack);
  else
    patchops::patch_scale_kernel<false><<<grid, 256, 0, st>>>(n, ysize, xsize, d_patches, d_win_y, d_win_x, d_weight, nullptr);
  AP_LAUNCH_CHECK();
  return 0;
}

// ------------------------------------------------------------------------------------------
// Map sink: Canvas.ud_grade (paint_bucket.py:2233-2257) = healpy.ud_grade(RING map, nside_out, pess, power).
// healpy reorders to NEST, averages (degrade) the good children of each parent or replicates (upgrade) the
// parent, and reorders back; NEST children of a parent are the (ix, iy) block of its face, so one thread per
// OUTPUT ring pixel walks that block through ring2xyf / xyf2ring (healpix.cuh) -- no reordered copy of the map.
// ------------------------------------------------------------------------------------------
namespace patchops {

constexpr double kUnseen = -1.6375e30;

__device__ __forceinline__ bool ud_bad(double v) {
  return fabs(v - kUnseen) <= 1.0e-8 + 1.0e-5 * 1.6375e30 || !isfinite(v);
}

__global__ void __launch_bounds__(256) ud_grade_kernel(ap::Hpx hin, ap::Hpx hout, const double* __restrict__ map_in,
                                                       double* __restrict__ map_out, int pess, double power_fact) {
  const int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= hout.npix) return;
  int64_t ix, iy;
  int face;
  ap::ring2xyf(hout, p, ix, iy, face);
  if (hin.nside > hout.nside) {  // degrade: children in NEST (Morton) order, as healpy's reshape(npix_out, rat2) sums them
    const int64_t r = hin.nside / hout.nside;
    const int64_t rat2 = r * r;
    double sum = 0.0;
    int64_t nhit = 0;
    for (int64_t c = 0; c < rat2; ++c) {
      int64_t dx = 0, dy = 0;
      for (int b = 0; (int64_t(1) << b) < r; ++b) {
        dx |= ((c >> (2 * b)) & 1) << b;
        dy |= ((c >> (2 * b + 1)) & 1) << b;
      }
      const double v = __ldg(map_in + ap::xyf2ring(hin, ix * r + dx, iy * r + dy, face));
      if (!ud_bad(v)) {
        sum += v;
        ++nhit;
      }
    }
    const bool badout = pess ? (nhit != rat2) : (nhit == 0);
    double hits = (double)nhit;
    if (power_fact != 1.0) hits = hits / power_fact;   // nhit / ratio**power
    map_out[p] = badout ? kUnseen : (nhit ? sum / hits : sum);
  } else {                       // upgrade (or copy): the parent's value times ratio**power
    const int64_t r = hout.nside / hin.nside;
    map_out[p] = __ldg(map_in + ap::xyf2ring(hin, ix / r, iy / r, face)) * power_fact;
  }
}

}  // namespace patchops

extern "C" int ap_ud_grade(int64_t nside_in, const double* d_map_in, int64_t nside_out, int pess, double power,
                           int use_power, double* d_map_out, void* stream) {
  if (check_nside(nside_in) || check_nside(nside_out)) return 1;
  if ((nside_in & (nside_in - 1)) || (nside_out & (nside_out - 1)))
    return fail("ud_grade: healpy degrades through NEST ordering, both nside must be powers of two");
  if (!d_map_in || !d_map_out) return fail("ud_grade: NULL map");
  const ap::Hpx hin = ap::make_hpx(nside_in), hout = ap::make_hpx(nside_out);
  const double ratio = (double)nside_out / (double)nside_in;
  const double pf = use_power ? pow(ratio, power) : 1.0;
  patchops::ud_grade_kernel<<<(unsigned)((hout.npix + 255) / 256), 256, 0, (cudaStream_t)stream>>>(hin, hout, d_map_in,
                                                                                                  d_map_out, pess, pf);
  AP_LAUNCH_CHECK();
  return 0;
}

