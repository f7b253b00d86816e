// Not-a-knot cubic interpolating spline: the device side of the reference's @interpolate
// (astropaint/lib/utils.py:457-521), which fits scipy InterpolatedUnivariateSpline(x, y, k=3)
// (FITPACK interpolating spline with knots x[0], x[2..n-3], x[n-1] == the unique not-a-knot
// cubic) through n_samples nodes and evaluates it on the halo's pixels.
#pragma once
#include "common.cuh"
#include "fastmath.cuh"

namespace ap {

constexpr int kMaxNodes = 32;

// Builds piecewise-polynomial coefficients: on [x_i, x_{i+1}] with t = X - x_i
//   S(X) = c[4i] + c[4i+1] t + c[4i+2] t^2 + c[4i+3] t^3,   i = 0..n-2.   Requires 4 <= n <= 32.
AP_HD_NOINLINE void spline_build_nak(int n, const double* x, const double* y, double* c) {
  double h[kMaxNodes], dl[kMaxNodes], diag[kMaxNodes], up[kMaxNodes], lo[kMaxNodes], b[kMaxNodes];
  for (int i = 0; i < n - 1; ++i) {
    h[i] = x[i + 1] - x[i];
    dl[i] = (y[i + 1] - y[i]) / h[i];
  }
  // slopes s_i: interior rows  h_i s_{i-1} + 2 (h_{i-1} + h_i) s_i + h_{i-1} s_{i+1} = 3 (h_i d_{i-1} + h_{i-1} d_i)
  for (int i = 1; i < n - 1; ++i) {
    lo[i] = h[i];
    diag[i] = 2.0 * (h[i - 1] + h[i]);
    up[i] = h[i - 1];
    b[i] = 3.0 * (h[i] * dl[i - 1] + h[i - 1] * dl[i]);
  }
  {  // not-a-knot at both ends
    double d = x[2] - x[0];
    diag[0] = h[1];
    up[0] = d;
    lo[0] = 0.0;
    b[0] = ((h[0] + 2.0 * d) * h[1] * dl[0] + h[0] * h[0] * dl[1]) / d;
    d = x[n - 1] - x[n - 3];
    diag[n - 1] = h[n - 3];
    lo[n - 1] = d;
    up[n - 1] = 0.0;
    b[n - 1] = (h[n - 2] * h[n - 2] * dl[n - 3] + (2.0 * d + h[n - 2]) * h[n - 3] * dl[n - 2]) / d;
  }
  // Thomas elimination
  for (int i = 1; i < n; ++i) {
    double m = lo[i] / diag[i - 1];
    diag[i] -= m * up[i - 1];
    b[i] -= m * b[i - 1];
  }
  b[n - 1] /= diag[n - 1];
  for (int i = n - 2; i >= 0; --i) b[i] = (b[i] - up[i] * b[i + 1]) / diag[i];
  for (int i = 0; i < n - 1; ++i) {
    c[4 * i + 0] = y[i];
    c[4 * i + 1] = b[i];
    c[4 * i + 2] = (3.0 * dl[i] - 2.0 * b[i] - b[i + 1]) / h[i];
    c[4 * i + 3] = (b[i] + b[i + 1] - 2.0 * dl[i]) / (h[i] * h[i]);
  }
}

// interval i with x_i <= X < x_{i+1} (clamped): uniform guess, then fix up (nodes may be log-spaced)
AP_HD int spline_interval(int n, const double* x, double X) {
  // a guess only (the loops below fix it up): reciprocal instead of a division keeps the painter's pixel loop call-free
  int i = (int)((X - x[0]) * fm_rcp(x[n - 1] - x[0]) * (double)(n - 1));
  if (i < 0) i = 0;
  if (i > n - 2) i = n - 2;
  while (i > 0 && X < x[i]) --i;
  while (i < n - 2 && X >= x[i + 1]) ++i;
  return i;
}

// Evaluate; X outside [x_0, x_{n-1}] extrapolates the end pieces (FITPACK splev, ext=0).
AP_HD double spline_eval(int n, const double* x, const double* c, double X) {
  int i = spline_interval(n, x, X);
  double t = X - x[i];
  const double* k = c + 4 * i;
  return k[0] + t * (k[1] + t * (k[2] + t * k[3]));
}

}  // namespace ap
