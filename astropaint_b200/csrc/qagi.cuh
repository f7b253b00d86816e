// QUADPACK dqagie / dqk15i / dqelg / dqpsrt, restated for one semi-infinite range
// (bound, +inf), as a host/device template.
//
// The reference's @LOS_integrate (astropaint/lib/utils.py:428-454) calls
//   scipy.integrate.quad(f, R_i, np.inf, epsabs=0., epsrel=1.e-2)        (utils.py:448)
// which is QUADPACK dqagie(inf=1, limit=50).  With epsrel=1e-2 its true error on the
// Battaglia16 integrand reaches ~2.5e-3, so reproducing the reference map to 1e-6 means
// reproducing QUADPACK's adaptive decisions, not the integral: this file follows the
// published netlib algorithm (Piessens et al. 1983, public domain) statement by statement.
// csrc/hostcheck.cpp compiles it with g++ and tests/test_hostcheck.py checks it against
// scipy.integrate.quad (value, abserr, neval) on the CPU.
#pragma once
#include "common.cuh"
#include "fastmath.cuh"

namespace ap {

#ifndef AP_QAGI_LIMIT
#define AP_QAGI_LIMIT 50
#endif
constexpr int kQagiLimit = AP_QAGI_LIMIT;  // scipy.integrate.quad default `limit` = 50
constexpr double kEpmach = 2.220446049250313e-16;   // d1mach(4)
constexpr double kUflow = 2.2250738585072014e-308;  // d1mach(1)
constexpr double kOflow = 1.7976931348623157e+308;  // d1mach(2)

struct QagiResult {
  double result, abserr;
  int neval, ier, last;
};

// 15-point transformed Gauss-Kronrod rule on (a, b) subset of (0, 1], x = boun + (1-t)/t.
//
// Code-size note (measured, profiles/): fully inlining the rule at its three call sites with the 15
// integrand evaluations unrolled produced a 270 KB kernel that thrashed the instruction cache.
// The rule is therefore ONE out-of-line device function with a rolled pair loop, and its
// node/weight tables live in __constant__ memory (uniform broadcast loads).  (Round 1, later: giving
// dqagie a single call site of the rule -- first evaluation as iteration 1 of the bisection loop -- and
// inlining it there measured 183.5 ms against 168.0 ms per 1e7 halos for this out-of-line form; a
// shared-memory home for fv1/fv2 and bank-conflict-free replicated log/exp tables measured +-1%.)  An earlier version
// kept the tables as function-local arrays: after inlining, nvcc 12.9 (-O3, sm_100a) overlapped
// their local-memory slots with the caller's epsilon table and the device results diverged from
// the host build — tests/test_gpu_parity.py::test_device_quadpack_matches_scipy guards this.
#if defined(__CUDACC__)
#define AP_TABLE_DECL static __constant__
#define AP_NOINLINE __noinline__
#else
#define AP_TABLE_DECL static const
#define AP_NOINLINE
#endif
#if defined(__CUDA_ARCH__) || !defined(__CUDACC__)
#define AP_GK(t) gk15_##t
#else
#define AP_GK(t) gk15h_##t
#endif
#define AP_GK15_XGK {0.991455371120812639206854697526329, 0.949107912342758524526189684047851, \
                     0.864864423359769072789712788640926, 0.741531185599394439863864773280788, \
                     0.586087235467691130294144838258730, 0.405845151377397166906606412076961, \
                     0.207784955007898467600689403773245, 0.0}
#define AP_GK15_WG {0.0, 0.129484966168869693270611432679082, 0.0, 0.279705391489276667901467771423780, \
                    0.0, 0.381830050505118944950369775488975, 0.0, 0.417959183673469387755102040816327}
#define AP_GK15_WGK {0.022935322010529224963732008058970, 0.063092092629978553290700663189204, \
                     0.104790010322250183839876322541518, 0.140653259715525918745189590510238, \
                     0.169004726639267902826583426598550, 0.190350578064785409913256402421014, \
                     0.204432940075298892414161999234649, 0.209482141084727828012999174891714}
AP_TABLE_DECL double gk15_xgk[8] = AP_GK15_XGK;
AP_TABLE_DECL double gk15_wg[8] = AP_GK15_WG;
AP_TABLE_DECL double gk15_wgk[8] = AP_GK15_WGK;
#if defined(__CUDACC__)
// host pass of nvcc (the host-buffer entry points never integrate, but the templates must compile)
static const double gk15h_xgk[8] = AP_GK15_XGK;
static const double gk15h_wg[8] = AP_GK15_WG;
static const double gk15h_wgk[8] = AP_GK15_WGK;
#endif

template <class F>
AP_NOINLINE AP_HD_NOINLINE void qk15i(const F& f_ref, double boun, double a, double b, double& result,
                                      double& abserr, double& resabs, double& resasc) {
  // the rule is out of line, so the integrand object arrives by address (the caller's local memory):
  // take a register copy once instead of reloading its fields on every evaluation
  const F f = f_ref;
  double fv1[7], fv2[7];
  double centr = AP_MUL(0.5, AP_ADD(a, b));
  double hlgth = AP_MUL(0.5, AP_SUB(b, a));
  // x = boun + (1 - t) / t and the Jacobian 1 / t^2 share one reciprocal (QUADPACK divides three
  // times; the results differ by <= 1 ulp)
  double ic = fm_rcp(centr);
  double fval1 = f(fm_fma(AP_SUB(1.0, centr), ic, boun));
  double fc = AP_MUL(AP_MUL(fval1, ic), ic);
  double resg = AP_MUL(AP_GK(wg)[7], fc);
  double resk = AP_MUL(AP_GK(wgk)[7], fc);
  resabs = fabs(resk);
#if defined(__CUDA_ARCH__)
#pragma unroll 1
#endif
  for (int j = 0; j < 7; ++j) {
    double absc = AP_MUL(hlgth, AP_GK(xgk)[j]);
    double absc1 = AP_SUB(centr, absc);
    double absc2 = AP_ADD(centr, absc);
    // 1/(c - a) and 1/(c + a) from one reciprocal of (c - a)(c + a)
    double ip = fm_rcp(AP_MUL(absc1, absc2));
    double i1 = AP_MUL(absc2, ip), i2 = AP_MUL(absc1, ip);
    double f1 = f(fm_fma(AP_SUB(1.0, absc1), i1, boun));
    double f2 = f(fm_fma(AP_SUB(1.0, absc2), i2, boun));
    f1 = AP_MUL(AP_MUL(f1, i1), i1);
    f2 = AP_MUL(AP_MUL(f2, i2), i2);
    fv1[j] = f1;
    fv2[j] = f2;
    double fsum = AP_ADD(f1, f2);
    resg = AP_ADD(resg, AP_MUL(AP_GK(wg)[j], fsum));
    resk = AP_ADD(resk, AP_MUL(AP_GK(wgk)[j], fsum));
    resabs = AP_ADD(resabs, AP_MUL(AP_GK(wgk)[j], AP_ADD(fabs(f1), fabs(f2))));
  }
  double reskh = AP_MUL(resk, 0.5);
  resasc = AP_MUL(AP_GK(wgk)[7], fabs(AP_SUB(fc, reskh)));
  for (int j = 0; j < 7; ++j)
    resasc = AP_ADD(resasc, AP_MUL(AP_GK(wgk)[j], AP_ADD(fabs(AP_SUB(fv1[j], reskh)),
                                                         fabs(AP_SUB(fv2[j], reskh)))));
  result = AP_MUL(resk, hlgth);
  resasc = AP_MUL(resasc, hlgth);
  resabs = AP_MUL(resabs, hlgth);
  abserr = fabs(AP_MUL(AP_SUB(resk, resg), hlgth));
  if (resasc != 0.0 && abserr != 0.0) {
#if defined(__CUDA_ARCH__) && !defined(AP_QK_IEEE_TAIL)
    // call-free (200 abserr / resasc)^1.5: reciprocal and reciprocal square root by MUFU seed + Newton steps
    // (<= 1 ulp from the IEEE division / sqrt; an overflowing or NaN intermediate ends in min(1, .) = 1 as before)
    double q = AP_MUL(AP_MUL(200.0, abserr), fm_rcp(resasc));
    double p = (q > 1.0e-290) ? AP_MUL(q, AP_MUL(q, fm_rsqrt(q))) : 0.0;
#else
    double q = AP_MUL(200.0, abserr) / resasc;
    double p = AP_MUL(q, sqrt(q));  // q**1.5
#endif
    abserr = AP_MUL(resasc, fmin(1.0, p));
  }
  if (resabs > kUflow / (50.0 * kEpmach)) abserr = fmax(AP_MUL(kEpmach * 50.0, resabs), abserr);
}

// epsilon algorithm (dqelg).  epstab has 52 entries, 1-based in the original.
AP_NOINLINE AP_HD_NOINLINE void qelg(int& n, double* epstab, double& result, double& abserr, double* res3la,
                         int& nres) {
  nres += 1;
  abserr = kOflow;
  result = epstab[n - 1];
  if (n < 3) {
    abserr = fmax(abserr, 5.0 * kEpmach * fabs(result));
    return;
  }
  const int limexp = 50;
  epstab[n + 1] = epstab[n - 1];
  int newelm = (n - 1) / 2;
  epstab[n - 1] = kOflow;
  int num = n;
  int k1 = n;
  bool converged = false;
  for (int i = 1; i <= newelm; ++i) {
    int k2 = k1 - 1, k3 = k1 - 2;
    double res = epstab[k1 + 1];
    double e0 = epstab[k3 - 1], e1 = epstab[k2 - 1], e2 = res;
    double e1abs = fabs(e1);
    double delta2 = AP_SUB(e2, e1), err2 = fabs(delta2);
    double tol2 = AP_MUL(fmax(fabs(e2), e1abs), kEpmach);
    double delta3 = AP_SUB(e1, e0), err3 = fabs(delta3);
    double tol3 = AP_MUL(fmax(e1abs, fabs(e0)), kEpmach);
    if (!(err2 > tol2 || err3 > tol3)) {
      // e0, e1, e2 equal to within machine accuracy: convergence assumed
      result = res;
      abserr = AP_ADD(err2, err3);
      abserr = fmax(abserr, 5.0 * kEpmach * fabs(result));
      converged = true;
      break;
    }
    double e3 = epstab[k1 - 1];
    epstab[k1 - 1] = e1;
    double delta1 = AP_SUB(e1, e3), err1 = fabs(delta1);
    double tol1 = AP_MUL(fmax(e1abs, fabs(e3)), kEpmach);
    if (err1 <= tol1 || err2 <= tol2 || err3 <= tol3) {
      n = i + i - 1;
      break;
    }
    double ss = AP_SUB(AP_ADD(1.0 / delta1, 1.0 / delta2), 1.0 / delta3);
    double epsinf = fabs(AP_MUL(ss, e1));
#ifdef AP_QAGI_TRACE
    printf("   qelg i=%d n=%d e0=%.17g e1=%.17g e2=%.17g e3=%.17g d1=%.17g d2=%.17g d3=%.17g ss=%.17g\n", i, n, e0, e1, e2,
           e3, delta1, delta2, delta3, ss);
#endif
    if (!(epsinf > 1.0e-4)) {
      n = i + i - 1;
      break;
    }
    res = AP_ADD(e1, 1.0 / ss);
    epstab[k1 - 1] = res;
    k1 -= 2;
    double error = AP_ADD(AP_ADD(err2, fabs(AP_SUB(res, e2))), err3);
    if (error > abserr) continue;
    abserr = error;
    result = res;
  }
  if (converged) return;  // label 100 already applied
  // shift the table
  if (n == limexp) n = 2 * (limexp / 2) - 1;
  int ib = ((num / 2) * 2 == num) ? 2 : 1;
  int ie = newelm + 1;
  for (int i = 1; i <= ie; ++i) {
    int ib2 = ib + 2;
    epstab[ib - 1] = epstab[ib2 - 1];
    ib = ib2;
  }
  if (num != n) {
    int indx = num - n + 1;
    for (int i = 1; i <= n; ++i) {
      epstab[i - 1] = epstab[indx - 1];
      indx += 1;
    }
  }
  if (nres < 4) {
    res3la[nres - 1] = result;
    abserr = kOflow;
  } else {
    abserr = AP_ADD(AP_ADD(fabs(AP_SUB(result, res3la[2])), fabs(AP_SUB(result, res3la[1]))),
                    fabs(AP_SUB(result, res3la[0])));
    res3la[0] = res3la[1];
    res3la[1] = res3la[2];
    res3la[2] = result;
  }
  abserr = fmax(abserr, 5.0 * kEpmach * fabs(result));
}

// dqpsrt: keep iord (1-based indices stored in a 0-based array) sorted by decreasing elist.
AP_NOINLINE AP_HD_NOINLINE void qpsrt(int limit, int last, int& maxerr, double& ermax, const double* elist,
                          int* iord, int& nrmax) {
  if (last <= 2) {
    iord[0] = 1;
    iord[1] = 2;
  } else {
    double errmax = elist[maxerr - 1];
    if (nrmax != 1) {
      int ido = nrmax - 1;
      for (int i = 1; i <= ido; ++i) {
        int isucc = iord[nrmax - 2];
        if (errmax <= elist[isucc - 1]) break;
        iord[nrmax - 1] = isucc;
        nrmax -= 1;
      }
    }
    int jupbn = last;
    if (last > (limit / 2 + 2)) jupbn = limit + 3 - last;
    double errmin = elist[last - 1];
    int jbnd = jupbn - 1;
    int ibeg = nrmax + 1;
    bool placed = false;
    int i = ibeg;
    for (; i <= jbnd; ++i) {
      int isucc = iord[i - 1];
      if (errmax >= elist[isucc - 1]) {
        placed = true;
        break;
      }
      iord[i - 2] = isucc;
    }
    if (!placed) {
      iord[jbnd - 1] = maxerr;
      iord[jupbn - 1] = last;
    } else {
      iord[i - 2] = maxerr;
      int k = jbnd;
      bool placed2 = false;
      for (int j = i; j <= jbnd; ++j) {
        int isucc = iord[k - 1];
        if (errmin < elist[isucc - 1]) {
          placed2 = true;
          break;
        }
        iord[k] = isucc;
        k -= 1;
      }
      if (placed2)
        iord[k] = last;
      else
        iord[i - 1] = last;
    }
  }
  maxerr = iord[nrmax - 1];
  ermax = elist[maxerr - 1];
}

// dqagie with inf = 1: integral of f over (bound, +inf).
template <class F>
AP_NOINLINE AP_HD_NOINLINE QagiResult qagi_upper(const F& f, double bound, double epsabs, double epsrel) {
  const int limit = kQagiLimit;
  double alist[kQagiLimit], blist[kQagiLimit], rlist[kQagiLimit], elist[kQagiLimit];
  int iord[kQagiLimit];
  double rlist2[52], res3la[3];
  QagiResult out;
  int ier = 0, last = 0;
  double result = 0.0, abserr = 0.0;
  alist[0] = 0.0;
  blist[0] = 1.0;
  rlist[0] = 0.0;
  elist[0] = 0.0;
  iord[0] = 0;
  if (epsabs <= 0.0 && epsrel < fmax(50.0 * kEpmach, 0.5e-28)) {
    out.result = 0.0; out.abserr = 0.0; out.neval = 0; out.ier = 6; out.last = 0;
    return out;
  }
  double boun = bound;
  double defabs, resabs;
  qk15i(f, boun, 0.0, 1.0, result, abserr, defabs, resabs);
  last = 1;
  rlist[0] = result;
  elist[0] = abserr;
  iord[0] = 1;
  double dres = fabs(result);
  double errbnd = fmax(epsabs, AP_MUL(epsrel, dres));
  if (abserr <= 100.0 * kEpmach * defabs && abserr > errbnd) ier = 2;
  if (limit == 1) ier = 1;
  bool done = (ier != 0 || (abserr <= errbnd && abserr != resabs) || abserr == 0.0);
  if (!done) {
    rlist2[0] = result;
    double errmax = abserr;
    int maxerr = 1;
    double area = result;
    double errsum = abserr;
    abserr = kOflow;
    int nrmax = 1, nres = 0, ktmin = 0, numrl2 = 2;
    bool extrap = false, noext = false;
    int ierro = 0, iroff1 = 0, iroff2 = 0, iroff3 = 0;
    int ksgn = -1;
    if (dres >= (1.0 - 50.0 * kEpmach) * defabs) ksgn = 1;
    double small = 0.0, erlarg = 0.0, ertest = 0.0, correc = 0.0, erlast;
    double reseps, abseps;
    // exit codes of the main loop: 0 = fell off the end (-> 100), 100, 115
    int exit_label = 100;
    for (last = 2; last <= limit; ++last) {
      double a1 = alist[maxerr - 1];
      double b1 = AP_MUL(0.5, AP_ADD(alist[maxerr - 1], blist[maxerr - 1]));
      double a2 = b1;
      double b2 = blist[maxerr - 1];
      erlast = errmax;
      double area1, error1, defab1, area2, error2, defab2, rabs;
      qk15i(f, boun, a1, b1, area1, error1, rabs, defab1);
      qk15i(f, boun, a2, b2, area2, error2, rabs, defab2);
      double area12 = AP_ADD(area1, area2);
      double erro12 = AP_ADD(error1, error2);
      errsum = AP_SUB(AP_ADD(errsum, erro12), errmax);
      area = AP_SUB(AP_ADD(area, area12), rlist[maxerr - 1]);
      if (!(defab1 == error1 || defab2 == error2)) {
        if (!(fabs(AP_SUB(rlist[maxerr - 1], area12)) > AP_MUL(1.0e-5, fabs(area12)) ||
              erro12 < AP_MUL(0.99, errmax))) {
          if (extrap) iroff2 += 1; else iroff1 += 1;
        }
        if (last > 10 && erro12 > errmax) iroff3 += 1;
      }
      rlist[maxerr - 1] = area1;
      rlist[last - 1] = area2;
      errbnd = fmax(epsabs, AP_MUL(epsrel, fabs(area)));
      if (iroff1 + iroff2 >= 10 || iroff3 >= 20) ier = 2;
      if (iroff2 >= 5) ierro = 3;
      if (last == limit) ier = 1;
      if (fmax(fabs(a1), fabs(b2)) <= (1.0 + 100.0 * kEpmach) * (fabs(a2) + 1000.0 * kUflow)) ier = 4;
      if (error2 > error1) {
        alist[maxerr - 1] = a2;
        alist[last - 1] = a1;
        blist[last - 1] = b1;
        rlist[maxerr - 1] = area2;
        rlist[last - 1] = area1;
        elist[maxerr - 1] = error2;
        elist[last - 1] = error1;
      } else {
        alist[last - 1] = a2;
        blist[maxerr - 1] = b1;
        blist[last - 1] = b2;
        elist[maxerr - 1] = error1;
        elist[last - 1] = error2;
      }
      qpsrt(limit, last, maxerr, errmax, elist, iord, nrmax);
#ifdef AP_QAGI_TRACE
      printf(" last=%d a1=%.17g b2=%.17g area1=%.17g e1=%.17g area2=%.17g e2=%.17g errsum=%.17g\n", last, a1, b2, area1,
             error1, area2, error2, errsum);
#endif
      if (errsum <= errbnd) { exit_label = 115; break; }
      if (ier != 0) { exit_label = 100; break; }
      if (last == 2) {
        small = 0.375;
        erlarg = errsum;
        ertest = errbnd;
        rlist2[1] = area;
        continue;
      }
      if (noext) continue;
      erlarg = AP_SUB(erlarg, erlast);
      if (fabs(AP_SUB(b1, a1)) > small) erlarg = AP_ADD(erlarg, erro12);
      if (!extrap) {
        if (fabs(AP_SUB(blist[maxerr - 1], alist[maxerr - 1])) > small) continue;
        extrap = true;
        nrmax = 2;
      }
      if (!(ierro == 3 || erlarg <= ertest)) {
        int id = nrmax;
        int jupbnd = last;
        if (last > (2 + limit / 2)) jupbnd = limit + 3 - last;
        bool cont = false;
        for (int k = id; k <= jupbnd; ++k) {
          maxerr = iord[nrmax - 1];
          errmax = elist[maxerr - 1];
          if (fabs(AP_SUB(blist[maxerr - 1], alist[maxerr - 1])) > small) { cont = true; break; }
          nrmax += 1;
        }
        if (cont) continue;
      }
      // perform extrapolation
      numrl2 += 1;
      rlist2[numrl2 - 1] = area;
      qelg(numrl2, rlist2, reseps, abseps, res3la, nres);
#ifdef AP_QAGI_TRACE
      printf("  last=%d area=%.17g errsum=%.17g numrl2=%d nres=%d reseps=%.17g abseps=%.17g\n", last, area, errsum,
             numrl2, nres, reseps, abseps);
#endif
      ktmin += 1;
      if (ktmin > 5 && abserr < AP_MUL(1.0e-3, errsum)) ier = 5;
      if (!(abseps >= abserr)) {
        ktmin = 0;
        abserr = abseps;
        result = reseps;
        correc = erlarg;
        ertest = fmax(epsabs, AP_MUL(epsrel, fabs(reseps)));
        if (abserr <= ertest) { exit_label = 100; break; }
      }
      if (numrl2 == 1) noext = true;
      if (ier == 5) { exit_label = 100; break; }
      maxerr = iord[0];
      errmax = elist[maxerr - 1];
      nrmax = 1;
      extrap = false;
      small = AP_MUL(small, 0.5);
      erlarg = errsum;
    }
    if (last > limit) last = limit;  // loop ran to completion (Fortran leaves last = limit+1; list holds limit)
    bool sum_list = (exit_label == 115);
    bool finished = false;
    if (!sum_list) {
      // label 100
      if (abserr == kOflow) {
        sum_list = true;
      } else if (ier + ierro == 0) {
        // label 110
        if (!(ksgn == -1 && fmax(fabs(result), fabs(area)) <= AP_MUL(defabs, 0.01))) {
          if (0.01 > result / area || result / area > 100.0 || errsum > fabs(area)) ier = 6;
        }
        finished = true;
      } else {
        if (ierro == 3) abserr = AP_ADD(abserr, correc);
        if (ier == 0) ier = 3;
        bool go110 = false;
        if (result != 0.0 && area != 0.0) {
          if (abserr / fabs(result) > errsum / fabs(area)) sum_list = true; else go110 = true;
        } else if (abserr > errsum) {
          sum_list = true;
        } else if (area == 0.0) {
          finished = true;
        } else {
          go110 = true;
        }
        if (go110) {
          if (!(ksgn == -1 && fmax(fabs(result), fabs(area)) <= AP_MUL(defabs, 0.01))) {
            if (0.01 > result / area || result / area > 100.0 || errsum > fabs(area)) ier = 6;
          }
          finished = true;
        }
      }
    }
    if (sum_list && !finished) {
      result = 0.0;
      for (int k = 0; k < last; ++k) result = AP_ADD(result, rlist[k]);
      abserr = errsum;
    }
  }
  out.neval = 30 * last - 15;
  if (ier > 2) ier -= 1;
  out.result = result;
  out.abserr = abserr;
  out.ier = ier;
  out.last = last;
  return out;
}


// =============================================================================================
// Rightmost-bisection fast path of dqagie.
//
// Every @LOS_integrate integrand ends in an inverse-square-root singularity at r = R, i.e. at t = 1 of
// the transformed range, so dqagie almost always bisects the interval that ends at t = 1 (measured on the
// bench catalog: 100% of 30 000 quadratures; depth 5 for 99.5% of them).  For that tree everything the
// general routine keeps in per-thread work lists (alist/blist/rlist/elist/iord in local memory, dqpsrt)
// collapses to a few scalars, and every abscissa of every rule is a compile-time-known dyadic number:
// (1 - t)/t and the Jacobian factors come from a table instead of a reciprocal per abscissa pair.
//
// qagi_upper_rm() is the SAME algorithm, statement for statement, under the invariant "the interval with
// the largest error is the one ending at t = 1 and it sits in list slot 1"; whenever the general routine
// could do anything else (left child with the larger error, a tie, the search for a larger interval before
// extrapolating, an error return after the first rule, a tree deeper than the table) it returns false and
// the caller runs qagi_upper().  The table is filled with the same device arithmetic qk15i uses, so accepted
// results are bit-identical to the general routine's (tests: host build and device, value/abserr/neval).
// =============================================================================================
#ifndef AP_RM_PAIRS
#define AP_RM_PAIRS 2
#endif
constexpr int kRmLevels = 24;                       // deepest bisection level tabulated
constexpr int kRmIntervals = 1 + 2 * kRmLevels;     // root, then (left, right) child per level

struct RmInterval {
  double hlgth, omc, ic, pad;          // half length; 1 - centre; 1 / centre
  double om1[7], i1[7], om2[7], i2[7];  // 1 - abscissa and 1 / abscissa of the 7 abscissa pairs
};

AP_HD void rm_fill(RmInterval& T, double a, double b) {
  const double centr = AP_MUL(0.5, AP_ADD(a, b));
  T.hlgth = AP_MUL(0.5, AP_SUB(b, a));
  T.ic = fm_rcp(centr);
  T.omc = AP_SUB(1.0, centr);
  T.pad = 0.0;
  for (int j = 0; j < 7; ++j) {
    double absc = AP_MUL(T.hlgth, AP_GK(xgk)[j]);
    double absc1 = AP_SUB(centr, absc);
    double absc2 = AP_ADD(centr, absc);
    double ip = fm_rcp(AP_MUL(absc1, absc2));
    T.i1[j] = AP_MUL(absc2, ip);
    T.i2[j] = AP_MUL(absc1, ip);
    T.om1[j] = AP_SUB(1.0, absc1);
    T.om2[j] = AP_SUB(1.0, absc2);
  }
}

// interval `idx`: 0 = (0, 1); 2k - 1 = (1 - 2^-(k-1), 1 - 2^-k); 2k = (1 - 2^-k, 1)
AP_HD void rm_bounds(int idx, double& a, double& b) {
  if (idx == 0) { a = 0.0; b = 1.0; return; }
  const int k = (idx + 1) >> 1;
  const double hi = AP_SUB(1.0, ldexp(1.0, -k)), lo = AP_SUB(1.0, ldexp(1.0, -(k - 1)));
  if (idx & 1) { a = lo; b = hi; } else { a = hi; b = 1.0; }
}

#if defined(__CUDACC__)
static __constant__ RmInterval rm_tab_c[kRmIntervals];
static __device__ RmInterval rm_tab_g[kRmIntervals];
static __global__ void rm_init_kernel() {
  int i = threadIdx.x;
  if (i >= kRmIntervals) return;
  double a, b;
  rm_bounds(i, a, b);
  rm_fill(rm_tab_g[i], a, b);
}
// fills the __constant__ table on the current device (once per device and process); returns a cudaError_t
static inline int rm_ensure_tables(void* stream) {
  static bool done[64] = {false};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 1;
  if (done[dev]) return 0;
  cudaStream_t st = (cudaStream_t)stream;
  rm_init_kernel<<<1, 64, 0, st>>>();
  void* src = nullptr;
  if (cudaGetSymbolAddress(&src, rm_tab_g) != cudaSuccess) return 1;
  if (cudaMemcpyToSymbolAsync(rm_tab_c, src, sizeof(RmInterval) * kRmIntervals, 0, cudaMemcpyDeviceToDevice, st) !=
      cudaSuccess)
    return 1;
  if (cudaGetLastError() != cudaSuccess) return 1;
  done[dev] = true;
  return 0;
}
#endif
#if !defined(__CUDA_ARCH__)
struct RmHostTable {
  RmInterval t[kRmIntervals];
  RmHostTable() {
    for (int i = 0; i < kRmIntervals; ++i) {
      double a, b;
      rm_bounds(i, a, b);
      rm_fill(t[i], a, b);
    }
  }
};
inline const RmInterval* rm_host_table() {
  static RmHostTable tab;
  return tab.t;
}
#endif
#if defined(__CUDA_ARCH__)
#define AP_RM_TAB(i) rm_tab_c[i]
#else
#define AP_RM_TAB(i) rm_host_table()[i]
#endif

// qk15i on tabulated interval `idx`: the same sums in the same order, abscissae from the table
#ifdef AP_RM_INLINE_RULE
#define AP_RM_RULE_ATTR AP_HD
#else
#define AP_RM_RULE_ATTR AP_NOINLINE AP_HD_NOINLINE
#endif
template <class F>
AP_RM_RULE_ATTR void qk15i_rm(const F& f_ref, double boun, int idx, double& result, double& abserr,
                              double& resabs, double& resasc) {
  const F f = f_ref;
  const RmInterval& T = AP_RM_TAB(idx);
  double fv1[7], fv2[7];
  const double hlgth = T.hlgth;
  const double ic = T.ic;
  double fval1 = f(fm_fma(T.omc, ic, boun));
  double fc = AP_MUL(AP_MUL(fval1, ic), ic);
  double resg = AP_MUL(AP_GK(wg)[7], fc);
  double resk = AP_MUL(AP_GK(wgk)[7], fc);
  resabs = fabs(resk);
#if defined(__CUDA_ARCH__) && AP_RM_PAIRS == 2
  // two abscissa pairs (four independent integrand evaluations) per trip: the evaluations are long serial
  // DFMA chains and the scheduler has only 5-6 warps to hide their latency with; the sums stay in j order.
  // Two bit-identical omissions: (1) the Gauss weights wg[0], wg[2], wg[4], wg[6] are exactly 0 (dqk15i
  // keeps them so that one loop serves both rules): resg + 0 * fsum == resg for finite fsum, so those
  // updates are skipped; (2) for an integrand that is >= 0 everywhere (F::kPositive) resabs repeats resk's
  // operations on the same operands, so it is not accumulated separately.
  {
    const double i1 = T.i1[0], i2 = T.i2[0];
    double f1 = f(fm_fma(T.om1[0], i1, boun));
    double f2 = f(fm_fma(T.om2[0], i2, boun));
    f1 = AP_MUL(AP_MUL(f1, i1), i1);
    f2 = AP_MUL(AP_MUL(f2, i2), i2);
    fv1[0] = f1;
    fv2[0] = f2;
    double fsum = AP_ADD(f1, f2);
    resk = AP_ADD(resk, AP_MUL(AP_GK(wgk)[0], fsum));
    if (!F::kPositive) resabs = AP_ADD(resabs, AP_MUL(AP_GK(wgk)[0], AP_ADD(fabs(f1), fabs(f2))));
  }
#pragma unroll 1
  for (int j = 1; j < 7; j += 2) {
    const double i1 = T.i1[j], i2 = T.i2[j], i3 = T.i1[j + 1], i4 = T.i2[j + 1];
    double f1 = f(fm_fma(T.om1[j], i1, boun));
    double f2 = f(fm_fma(T.om2[j], i2, boun));
    double f3 = f(fm_fma(T.om1[j + 1], i3, boun));
    double f4 = f(fm_fma(T.om2[j + 1], i4, boun));
    f1 = AP_MUL(AP_MUL(f1, i1), i1);
    f2 = AP_MUL(AP_MUL(f2, i2), i2);
    f3 = AP_MUL(AP_MUL(f3, i3), i3);
    f4 = AP_MUL(AP_MUL(f4, i4), i4);
    fv1[j] = f1;
    fv2[j] = f2;
    fv1[j + 1] = f3;
    fv2[j + 1] = f4;
    double fsum = AP_ADD(f1, f2);
    resg = AP_ADD(resg, AP_MUL(AP_GK(wg)[j], fsum));
    resk = AP_ADD(resk, AP_MUL(AP_GK(wgk)[j], fsum));
    if (!F::kPositive) resabs = AP_ADD(resabs, AP_MUL(AP_GK(wgk)[j], AP_ADD(fabs(f1), fabs(f2))));
    fsum = AP_ADD(f3, f4);
    resk = AP_ADD(resk, AP_MUL(AP_GK(wgk)[j + 1], fsum));
    if (!F::kPositive) resabs = AP_ADD(resabs, AP_MUL(AP_GK(wgk)[j + 1], AP_ADD(fabs(f3), fabs(f4))));
  }
  if (F::kPositive) resabs = fabs(resk);
#else
#if defined(__CUDA_ARCH__)
#pragma unroll 1
#endif
  for (int j = 0; j < 7; ++j) {
    const double i1 = T.i1[j], i2 = T.i2[j];
    double f1 = f(fm_fma(T.om1[j], i1, boun));
    double f2 = f(fm_fma(T.om2[j], i2, boun));
    f1 = AP_MUL(AP_MUL(f1, i1), i1);
    f2 = AP_MUL(AP_MUL(f2, i2), i2);
    fv1[j] = f1;
    fv2[j] = f2;
    double fsum = AP_ADD(f1, f2);
    resg = AP_ADD(resg, AP_MUL(AP_GK(wg)[j], fsum));
    resk = AP_ADD(resk, AP_MUL(AP_GK(wgk)[j], fsum));
    resabs = AP_ADD(resabs, AP_MUL(AP_GK(wgk)[j], AP_ADD(fabs(f1), fabs(f2))));
  }
#endif
  double reskh = AP_MUL(resk, 0.5);
  resasc = AP_MUL(AP_GK(wgk)[7], fabs(AP_SUB(fc, reskh)));
  for (int j = 0; j < 7; ++j)
    resasc = AP_ADD(resasc, AP_MUL(AP_GK(wgk)[j], AP_ADD(fabs(AP_SUB(fv1[j], reskh)),
                                                         fabs(AP_SUB(fv2[j], reskh)))));
  result = AP_MUL(resk, hlgth);
  resasc = AP_MUL(resasc, hlgth);
  resabs = AP_MUL(resabs, hlgth);
  abserr = fabs(AP_MUL(AP_SUB(resk, resg), hlgth));
  if (resasc != 0.0 && abserr != 0.0) {
#if defined(__CUDA_ARCH__) && !defined(AP_QK_IEEE_TAIL)
    // call-free (200 abserr / resasc)^1.5: reciprocal and reciprocal square root by MUFU seed + Newton steps
    // (<= 1 ulp from the IEEE division / sqrt; an overflowing or NaN intermediate ends in min(1, .) = 1 as before)
    double q = AP_MUL(AP_MUL(200.0, abserr), fm_rcp(resasc));
    double p = (q > 1.0e-290) ? AP_MUL(q, AP_MUL(q, fm_rsqrt(q))) : 0.0;
#else
    double q = AP_MUL(200.0, abserr) / resasc;
    double p = AP_MUL(q, sqrt(q));  // q**1.5
#endif
    abserr = AP_MUL(resasc, fmin(1.0, p));
  }
  if (resabs > kUflow / (50.0 * kEpmach)) abserr = fmax(AP_MUL(kEpmach * 50.0, resabs), abserr);
}

// dqagie (inf = 1) under the rightmost-bisection invariant.  false = not this tree: run qagi_upper().
template <class F>
AP_NOINLINE AP_HD_NOINLINE bool qagi_upper_rm(const F& f, double bound, double epsabs, double epsrel, QagiResult& out) {
  const int limit = kQagiLimit;
  if (!(AP_RM_TAB(0).hlgth == 0.5)) return false;  // table not initialised on this device
  if (epsabs <= 0.0 && epsrel < fmax(50.0 * kEpmach, 0.5e-28)) return false;
  double rlist2[52], res3la[3];
  double lres[kRmLevels + 2];  // rlist[1 .. last-1]: the left children in list-slot order
  int ier = 0, last = 1;
  double result, abserr, defabs, resabs;
  const double boun = bound;
#ifdef AP_RM_INLINE_RULE
  // ONE call site of the (inlined) rule: rule number r = 0 is the root, r = 2 lev - 1 / 2 lev the left / right
  // child of level lev; the bisection step runs after every even r
  double area1 = 0.0, error1 = 0.0, defab1 = 0.0, area2 = 0.0, error2 = 0.0, defab2 = 0.0;
  double dres = 0.0, errbnd = 0.0, rR = 0.0, errmax = 0.0, aR = 0.0, eLmax = 0.0, area = 0.0, errsum = 0.0;
  int nres = 0, ktmin = 0, numrl2 = 2;
  bool extrap = false, noext = false;
  int ierro = 0, iroff1 = 0, iroff2 = 0, iroff3 = 0;
  int ksgn = -1;
  double small = 0.0, erlarg = 0.0, ertest = 0.0, correc = 0.0, erlast = 0.0;
  double reseps, abseps;
  int exit_label = 100;
  result = abserr = defabs = resabs = 0.0;
#if defined(__CUDA_ARCH__)
#pragma unroll 1
#endif
  for (int r = 0;; ++r) {
    double r_, e_, ra_, rc_;
    qk15i_rm(f, boun, r, r_, e_, ra_, rc_);
    if (r == 0) {
      result = r_; abserr = e_; defabs = ra_; resabs = rc_;
      dres = fabs(result);
      errbnd = fmax(epsabs, AP_MUL(epsrel, dres));
      if (abserr <= 100.0 * kEpmach * defabs && abserr > errbnd) ier = 2;
      if (limit == 1) ier = 1;
      if (ier != 0 || (abserr <= errbnd && abserr != resabs) || abserr == 0.0) return false;
      rlist2[0] = result;
      rR = result; errmax = abserr; area = result; errsum = abserr;
      abserr = kOflow;
      if (dres >= (1.0 - 50.0 * kEpmach) * defabs) ksgn = 1;
      if (1 > kRmLevels) return false;
      continue;
    }
    if (r & 1) {
      area1 = r_; error1 = e_; defab1 = rc_;
      continue;
    }
    area2 = r_; error2 = e_; defab2 = rc_;
    last = (r >> 1) + 1;
    // ---- the bisection step of iteration `last` ----
    bool next = false;   // true: `continue` of the original loop
    do {
      const double a1 = aR;
      const double b1 = AP_MUL(0.5, AP_ADD(aR, 1.0));
      const double a2 = b1;
      const double b2 = 1.0;
      erlast = errmax;
      double area12 = AP_ADD(area1, area2);
      double erro12 = AP_ADD(error1, error2);
      errsum = AP_SUB(AP_ADD(errsum, erro12), errmax);
      area = AP_SUB(AP_ADD(area, area12), rR);
      if (!(defab1 == error1 || defab2 == error2)) {
        if (!(fabs(AP_SUB(rR, area12)) > AP_MUL(1.0e-5, fabs(area12)) || erro12 < AP_MUL(0.99, errmax))) {
          if (extrap) iroff2 += 1; else iroff1 += 1;
        }
        if (last > 10 && erro12 > errmax) iroff3 += 1;
      }
      errbnd = fmax(epsabs, AP_MUL(epsrel, fabs(area)));
      if (iroff1 + iroff2 >= 10 || iroff3 >= 20) ier = 2;
      if (iroff2 >= 5) ierro = 3;
      if (last == limit) ier = 1;
      if (fmax(fabs(a1), fabs(b2)) <= (1.0 + 100.0 * kEpmach) * (fabs(a2) + 1000.0 * kUflow)) ier = 4;
      if (!(error2 > error1) || !(error2 > eLmax)) return false;
      lres[last - 1] = area1;
      rR = area2;
      aR = a2;
      errmax = error2;
      eLmax = fmax(eLmax, error1);
      if (errsum <= errbnd) { exit_label = 115; break; }
      if (ier != 0) { exit_label = 100; break; }
      if (last == 2) {
        small = 0.375;
        erlarg = errsum;
        ertest = errbnd;
        rlist2[1] = area;
        next = true; break;
      }
      if (noext) { next = true; break; }
      erlarg = AP_SUB(erlarg, erlast);
      if (fabs(AP_SUB(b1, a1)) > small) erlarg = AP_ADD(erlarg, erro12);
      if (!extrap) {
        if (fabs(AP_SUB(1.0, aR)) > small) { next = true; break; }
        extrap = true;
      }
      if (!(ierro == 3 || erlarg <= ertest)) return false;
      numrl2 += 1;
      rlist2[numrl2 - 1] = area;
      qelg(numrl2, rlist2, reseps, abseps, res3la, nres);
      ktmin += 1;
      if (ktmin > 5 && abserr < AP_MUL(1.0e-3, errsum)) ier = 5;
      if (!(abseps >= abserr)) {
        ktmin = 0;
        abserr = abseps;
        result = reseps;
        correc = erlarg;
        ertest = fmax(epsabs, AP_MUL(epsrel, fabs(reseps)));
        if (abserr <= ertest) { exit_label = 100; break; }
      }
      if (numrl2 == 1) noext = true;
      if (ier == 5) { exit_label = 100; break; }
      extrap = false;
      small = AP_MUL(small, 0.5);
      erlarg = errsum;
      next = true;
    } while (false);
    if (!next) break;                 // exit_label set
    if (last + 1 > limit) { last = limit + 1; break; }   // the original loop would fall off its end
    if (last > kRmLevels) return false;                 // next level is not tabulated
  }
#else
  qk15i_rm(f, boun, 0, result, abserr, defabs, resabs);
  double dres = fabs(result);
  double errbnd = fmax(epsabs, AP_MUL(epsrel, dres));
  if (abserr <= 100.0 * kEpmach * defabs && abserr > errbnd) ier = 2;
  if (limit == 1) ier = 1;
  if (ier != 0 || (abserr <= errbnd && abserr != resabs) || abserr == 0.0) return false;  // finished after one rule
  rlist2[0] = result;
  double rR = result;       // rlist[0]: the interval ending at t = 1
  double errmax = abserr;   // elist[0]
  double aR = 0.0;          // alist[0] (blist[0] = 1)
  double eLmax = 0.0;       // largest elist[k], k >= 1
  double area = result;
  double errsum = abserr;
  abserr = kOflow;
  int nres = 0, ktmin = 0, numrl2 = 2;
  bool extrap = false, noext = false;
  int ierro = 0, iroff1 = 0, iroff2 = 0, iroff3 = 0;
  int ksgn = -1;
  if (dres >= (1.0 - 50.0 * kEpmach) * defabs) ksgn = 1;
  double small = 0.0, erlarg = 0.0, ertest = 0.0, correc = 0.0, erlast;
  double reseps, abseps;
  int exit_label = 100;
  for (last = 2; last <= limit; ++last) {
    const int lev = last - 1;
    if (lev > kRmLevels) return false;
    const double a1 = aR;
    const double b1 = AP_MUL(0.5, AP_ADD(aR, 1.0));
    const double a2 = b1;
    const double b2 = 1.0;
    erlast = errmax;
    double area1, error1, defab1, area2, error2, defab2, rabs;
    qk15i_rm(f, boun, 2 * lev - 1, area1, error1, rabs, defab1);
    qk15i_rm(f, boun, 2 * lev, area2, error2, rabs, defab2);
    double area12 = AP_ADD(area1, area2);
    double erro12 = AP_ADD(error1, error2);
    errsum = AP_SUB(AP_ADD(errsum, erro12), errmax);
    area = AP_SUB(AP_ADD(area, area12), rR);
    if (!(defab1 == error1 || defab2 == error2)) {
      if (!(fabs(AP_SUB(rR, area12)) > AP_MUL(1.0e-5, fabs(area12)) || erro12 < AP_MUL(0.99, errmax))) {
        if (extrap) iroff2 += 1; else iroff1 += 1;
      }
      if (last > 10 && erro12 > errmax) iroff3 += 1;
    }
    errbnd = fmax(epsabs, AP_MUL(epsrel, fabs(area)));
    if (iroff1 + iroff2 >= 10 || iroff3 >= 20) ier = 2;
    if (iroff2 >= 5) ierro = 3;
    if (last == limit) ier = 1;
    if (fmax(fabs(a1), fabs(b2)) <= (1.0 + 100.0 * kEpmach) * (fabs(a2) + 1000.0 * kUflow)) ier = 4;
    // the invariant: the right child (error2) takes slot 1 and is strictly the largest error of the list, so
    // dqpsrt leaves maxerr = 1 (a tie or a larger left child is the general routine's business)
    if (!(error2 > error1) || !(error2 > eLmax)) return false;
    lres[last - 1] = area1;
    rR = area2;
    aR = a2;
    errmax = error2;
    eLmax = fmax(eLmax, error1);
    if (errsum <= errbnd) { exit_label = 115; break; }
    if (ier != 0) { exit_label = 100; break; }
    if (last == 2) {
      small = 0.375;
      erlarg = errsum;
      ertest = errbnd;
      rlist2[1] = area;
      continue;
    }
    if (noext) continue;
    erlarg = AP_SUB(erlarg, erlast);
    if (fabs(AP_SUB(b1, a1)) > small) erlarg = AP_ADD(erlarg, erro12);
    if (!extrap) {
      if (fabs(AP_SUB(1.0, aR)) > small) continue;
      extrap = true;
    }
    if (!(ierro == 3 || erlarg <= ertest)) return false;  // dqagie would first bisect a larger interval
    numrl2 += 1;
    rlist2[numrl2 - 1] = area;
    qelg(numrl2, rlist2, reseps, abseps, res3la, nres);
    ktmin += 1;
    if (ktmin > 5 && abserr < AP_MUL(1.0e-3, errsum)) ier = 5;
    if (!(abseps >= abserr)) {
      ktmin = 0;
      abserr = abseps;
      result = reseps;
      correc = erlarg;
      ertest = fmax(epsabs, AP_MUL(epsrel, fabs(reseps)));
      if (abserr <= ertest) { exit_label = 100; break; }
    }
    if (numrl2 == 1) noext = true;
    if (ier == 5) { exit_label = 100; break; }
    // maxerr = iord[0] = 1, errmax = elist[0] (unchanged), nrmax = 1
    extrap = false;
    small = AP_MUL(small, 0.5);
    erlarg = errsum;
  }
#endif
  if (last > limit) last = limit;
  bool sum_list = (exit_label == 115);
  bool finished = false;
  if (!sum_list) {
    if (abserr == kOflow) {
      sum_list = true;
    } else if (ier + ierro == 0) {
      if (!(ksgn == -1 && fmax(fabs(result), fabs(area)) <= AP_MUL(defabs, 0.01))) {
        if (0.01 > result / area || result / area > 100.0 || errsum > fabs(area)) ier = 6;
      }
      finished = true;
    } else {
      if (ierro == 3) abserr = AP_ADD(abserr, correc);
      if (ier == 0) ier = 3;
      bool go110 = false;
      if (result != 0.0 && area != 0.0) {
        if (abserr / fabs(result) > errsum / fabs(area)) sum_list = true; else go110 = true;
      } else if (abserr > errsum) {
        sum_list = true;
      } else if (area == 0.0) {
        finished = true;
      } else {
        go110 = true;
      }
      if (go110) {
        if (!(ksgn == -1 && fmax(fabs(result), fabs(area)) <= AP_MUL(defabs, 0.01))) {
          if (0.01 > result / area || result / area > 100.0 || errsum > fabs(area)) ier = 6;
        }
        finished = true;
      }
    }
  }
  if (sum_list && !finished) {
    result = rR;
    for (int k = 1; k < last; ++k) result = AP_ADD(result, lres[k]);
    abserr = errsum;
  }
  out.neval = 30 * last - 15;
  if (ier > 2) ier -= 1;
  out.result = result;
  out.abserr = abserr;
  out.ier = ier;
  out.last = last;
  return true;
}

// dqagie, inf = 1: the fast path when the quadrature follows the rightmost-bisection tree, else the general routine
template <class F, int FAST = 1>
AP_HD QagiResult qagi_upper_auto(const F& f, double bound, double epsabs, double epsrel) {
#ifndef AP_QAGI_NO_FAST_PATH
  if (FAST) {
    QagiResult q;
    if (qagi_upper_rm(f, bound, epsabs, epsrel, q)) return q;
  }
#endif
  return qagi_upper(f, bound, epsabs, epsrel);
}

}  // namespace ap
