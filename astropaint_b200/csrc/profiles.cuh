// Built-in templates as host/device functions.
//
// Each profile P has
//   setup<P>(p, ctx)      once per halo: p[] = the template's per-halo arguments in the order of
//                         the reference signature (without R / R_vec), ctx[] = derived constants
//   eval_R<P>(ctx, R)     R-mode value at projected distance R [Mpc]        (spray, :2386-2389)
//   eval_vec<P>(ctx, rv)  R_vec-mode value at the chord vector rv [Mpc]
// Reference formulas: astropaint/profiles/spherical.py:25-79, NFW.py:65-193,
// Battaglia16.py:49-271, README.md:99-101 (the "nonsense" Gaussian), lib/transform.py:182-225.
#pragma once
#include "common.cuh"
#include "qagi.cuh"
#include "fastmath.cuh"

namespace ap {

enum ProfileId : int {
  P_CONSTANT = 0,       // spherical.constant_density_2D(R, constant)
  P_LINEAR = 1,         // spherical.linear_density_2D(R, intercept, slope)
  P_SOLID_SPHERE = 2,   // spherical.solid_sphere_2D(R, M_200c, R_200c)   (projected top-hat)
  P_GAUSSIAN = 3,       // README gaussian(R, R_200c, x, y, z)
  P_NFW_RHO2D = 4,      // NFW.rho_2D_bartlemann(R, rho_s, R_s)
  P_NFW_TAU2D = 5,      // NFW.tau_2D(R, rho_s, R_s)
  P_NFW_KSZ = 6,        // NFW.kSZ_T(R, rho_s, R_s, v_r, *, T_cmb)
  P_NFW_DEFLECTION = 7, // NFW.deflection_angle(R, c_200c, R_200c, M_200c)
  P_NFW_BG = 8,         // NFW.BG(R_vec, c_200c, R_200c, M_200c, theta, phi, v_th, v_ph, *, T_cmb)
  P_B16_TAU2D = 9,      // Battaglia16.tau_2D(R, R_200c, M_200c, redshift)  (QUADPACK per pixel)
  P_B16_RHOGAS2D = 10,  // Battaglia16.rho_gas_2D(R, R_200c, M_200c, redshift)  (QUADPACK per pixel)
  P_NFW_RHO2D_LOS = 11, // NFW.rho_2D(R, rho_s, R_s) = LOS_integrate(rho_3D)  (QUADPACK per pixel)
  P_COUNT_ = 12
};

// 3-D profiles that @LOS_integrate projects on device (ap_los_nodes / ap_paint_los_small)
enum LosKind : int {
  LOS_B16_TAU = 0,      // Battaglia16.tau_3D(r, R_200c, M_200c, redshift)
  LOS_B16_NE = 1,       // Battaglia16.ne_3D
  LOS_B16_RHO_GAS = 2,  // Battaglia16.rho_gas_3D
  LOS_NFW_RHO = 3       // NFW.rho_3D(r, rho_s, r_s)
};

constexpr int kMaxParams = 8;
constexpr int kCtx = 8;

AP_HD int profile_n_params(int id) {
  switch (id) {
    case P_CONSTANT: return 1;
    case P_LINEAR: return 2;
    case P_SOLID_SPHERE: return 2;
    case P_GAUSSIAN: return 4;
    case P_NFW_RHO2D: return 2;
    case P_NFW_TAU2D: return 2;
    case P_NFW_KSZ: return 4;
    case P_NFW_DEFLECTION: return 3;
    case P_NFW_BG: return 8;
    case P_B16_TAU2D: return 3;
    case P_B16_RHOGAS2D: return 3;
    case P_NFW_RHO2D_LOS: return 2;
    default: return -1;
  }
}
AP_HD bool profile_is_vec(int id) { return id == P_NFW_BG; }

// ---- constants (NFW.py:25-31, Battaglia16.py:32-46; astropy CODATA2018 / Planck18_arXiv_v2) ----
namespace cst {
constexpr double G = 6.6743e-11;
constexpr double c_si = 299792458.0;
constexpr double Mpc = 3.085677581491367e22;
constexpr double GMsun = 1.3271244e20;
constexpr double Msun = GMsun / G;
constexpr double sigma_sb = 5.6703744191844314e-08;
constexpr double kB_eV = 8.617333262145179e-05;
constexpr double H0 = 67.66, Om0 = 0.30966, Ob0 = 0.04897, Tcmb0 = 2.7255, Neff = 3.046;
constexpr double m_nu = 0.06;  // eV, one massive species + two massless
constexpr double h = H0 / 100.0;
constexpr double f_b = Ob0 / Om0;
constexpr double H0_si = H0 * 1.0e3 / Mpc;
constexpr double rhoc0_si = 3.0 * H0_si * H0_si / (8.0 * kPi * G);
constexpr double crit_dens_0 = rhoc0_si * Mpc * Mpc * Mpc / Msun;  // M_sun / Mpc^3
constexpr double Ogamma0 = 4.0 * sigma_sb / (c_si * c_si * c_si) * (Tcmb0 * Tcmb0 * Tcmb0 * Tcmb0) / rhoc0_si;
constexpr double Tnu0 = 0.7137658555036082 * Tcmb0;
constexpr double nu_y = m_nu / (kB_eV * Tnu0);
constexpr double sigma_T = 6.6524587321e-29 / (Mpc * Mpc);  // Mpc^2
constexpr double m_p = 1.67262192369e-27 / Msun;            // M_sun
constexpr double c_kms = 299792.0;
constexpr double T_cmb = 2.7251;
constexpr double Gcm2 = 4.785E-20;
}  // namespace cst

AP_HD double nu_relative_density(double z) {
  const double prefac = 0.22710731766, p = 1.83, invp = 0.54644808743, k = 0.3173;
  double curr = cst::nu_y / (1.0 + z);
  double rel_mass = pow(1.0 + pow(k * curr, p), invp) + 2.0;
  return prefac * (cst::Neff / 3.0) * rel_mass;
}

// cosmo.critical_density(z) in M_sun / Mpc^3 (Battaglia16.py:89)
AP_HD double critical_density(double z) {
  double Onu0 = cst::Ogamma0 * nu_relative_density(0.0);
  double Ode0 = 1.0 - cst::Om0 - cst::Ogamma0 - Onu0;
  double zp1 = 1.0 + z;
  double Or = cst::Ogamma0 * (1.0 + nu_relative_density(z));
  double e2 = zp1 * zp1 * zp1 * (Or * zp1 + cst::Om0) + Ode0;
  return cst::crit_dens_0 * e2;
}

// ---- NFW helper: g(x) = Re[ 2/sqrt(1-x^2) * arctanh(sqrt((1-x)/(1+x))) ]  (NFW.py:76, :140) ----
AP_HD double nfw_g(double x) {
  if (x < 1.0) {
    double s = sqrt((1.0 - x) * (1.0 + x));
    return 2.0 / s * atanh(sqrt((1.0 - x) / (1.0 + x)));
  }
  double s = sqrt((x - 1.0) * (x + 1.0));
  return 2.0 / s * atan(sqrt((x - 1.0) / (1.0 + x)));
}

// ---- Battaglia16 tau_3D(r) * 2r / sqrt(r^2 - R^2): the @LOS_integrate integrand ----
struct B16Integrand {
  static constexpr bool kPositive = true;  // exp(.) * r / sqrt(r^2 - R^2) >= 0: qk15i_rm skips the resabs sums
  double inv_R200xc, alpha, expo, RR;  // RR = R^2
  // fit = (1 + y^alpha)^expo * y^gamma evaluated as exp(expo * log(1 + e^(alpha L)) + gamma L),
  // L = log y: two logs and two exps instead of three pow() calls (each ~4x a log+exp pair on
  // sm_100a).  Relative error ~1e-15, far inside the distance at which QUADPACK's decisions flip
  // (tests compare value AND neval against scipy on host and device).
  // The constant factor K2 = 2 rho0 rho_c(z) f_b / factor * sigma_T multiplies the INTEGRAL, not the
  // integrand (los_eval): with epsabs = 0 every QUADPACK test is homogeneous in the integrand's scale.
  AP_HD double operator()(double r) const {
    // (Round 2: replacing the exp -> log pair of log(1 + y^alpha) by ONE table-driven degree-7 polynomial of
    // softplus(alpha L) -- 12 fp64 instructions instead of 20 -- ran 66% SLOWER: its 64 B of coefficients per
    // evaluation and lane saturate the shared-memory pipe (the pair reads 24 B); profiles/README.md.)
    double L = fast_log(AP_MUL(r, inv_R200xc));        // log((r / R_200c) / xc)
    double ya = fast_exp(AP_MUL(alpha, L));
    double fit = fast_exp(fm_fma(expo, fast_log(AP_ADD(1.0, ya)), AP_MUL(-0.2, L)));
    double w = fm_fma(r, r, -RR);
#if defined(__CUDA_ARCH__)
    return AP_MUL(AP_MUL(fit, r), fm_rsqrt(w));
#else
    return AP_MUL(fit, r) / sqrt(w);
#endif
  }
};
// ctx layout for P_B16_TAU2D: [inv_R200xc, alpha, expo, K]
AP_HD void b16_setup(double R_200c, double M_200c, double z, double* ctx) {
  const double xc = 0.5, gamma = -0.2;
  double m = (M_200c / cst::h) / 1.0e14;
  double rho0 = 4.0e3 * pow(m, 0.29) * pow(1.0 + z, -0.66);
  double alpha = 0.88 * pow(m, -0.03) * pow(1.0 + z, 0.19);
  double beta = 3.83 * pow(m, 0.04) * pow(1.0 + z, -0.025);
  // ne_3D factor (Battaglia16.py:115-124)
  const double me = 9.10938291e-31, mH = 1.67262178e-27, mHe = 4.0 * mH, xH = 0.76;
  const double nH_ne = 2.0 * xH / (xH + 1.0), nHe_ne = (1.0 - xH) / (2.0 * (1.0 + xH));
  double factor = (me + nH_ne * mH + nHe_ne * mHe) / 1.989e30 * cst::h;
  ctx[0] = 1.0 / (R_200c * xc);
  ctx[1] = alpha;
  ctx[2] = -(beta + gamma) / alpha;
  ctx[3] = rho0 * critical_density(z) * (cst::Ob0 / cst::Om0) / factor * cst::sigma_T;
}

// NFW.rho_3D(r) * 2r / sqrt(r^2 - R^2) = 8 rho_s r_s^3 / ((r + r_s)^2 sqrt(r^2 - R^2))   (NFW.py:38-58, 83-93)
struct NfwRhoIntegrand {
  static constexpr bool kPositive = false;  // rho_s is a user column: no sign assumption
  double K8, rs, RR;
  AP_HD double operator()(double r) const {
    double q = AP_ADD(r, rs);
    double w = AP_SUB(AP_MUL(r, r), RR);
#if defined(__CUDA_ARCH__)
    return AP_MUL(K8 * fm_rcp(AP_MUL(q, q)), fm_rsqrt(w));
#else
    return K8 / AP_MUL(q, q) / sqrt(w);
#endif
  }
};

template <int KIND>
AP_HD void los_setup(double p0, double p1, double p2, double* ctx) {
  if (KIND == LOS_NFW_RHO) {
    ctx[0] = 8.0 * p0 * p1 * p1 * p1;
    ctx[1] = p1;
    return;
  }
  b16_setup(p0, p1, p2, ctx);  // ctx[3] = sigma_T * n_e normalisation
  if (KIND != LOS_B16_TAU) {
    ctx[3] /= cst::sigma_T;  // n_e
    if (KIND == LOS_B16_RHO_GAS) {
      const double me = 9.10938291e-31, mH = 1.67262178e-27, mHe = 4.0 * mH, xH = 0.76;
      const double nH_ne = 2.0 * xH / (xH + 1.0), nHe_ne = (1.0 - xH) / (2.0 * (1.0 + xH));
      ctx[3] *= (me + nH_ne * mH + nHe_ne * mHe) / 1.989e30 * cst::h;  // back to rho_gas
    }
  }
}

template <int KIND, int FAST = 1>
AP_HD double los_eval(const double* ctx, double R, QagiResult* info = nullptr) {
  QagiResult q;
  if (KIND == LOS_NFW_RHO) {
    NfwRhoIntegrand f;
    f.K8 = ctx[0];
    f.rs = ctx[1];
    f.RR = AP_MUL(R, R);
    q = qagi_upper_auto<decltype(f), FAST>(f, R, 0.0, 1.0e-2);
  } else {
    B16Integrand f;
    f.inv_R200xc = ctx[0];
    f.alpha = ctx[1];
    f.expo = ctx[2];
    f.RR = AP_MUL(R, R);
    q = qagi_upper_auto<decltype(f), FAST>(f, R, 0.0, 1.0e-2);  // utils.py:448
    const double K2 = 2.0 * ctx[3];
    q.result = AP_MUL(q.result, K2);
    q.abserr = AP_MUL(q.abserr, K2);
  }
  if (info) *info = q;
  return q.result;
}

AP_HD double b16_tau_2D(const double* ctx, double R, QagiResult* info = nullptr) {
  B16Integrand f;
  f.inv_R200xc = ctx[0];
  f.alpha = ctx[1];
  f.expo = ctx[2];
  f.RR = AP_MUL(R, R);
  QagiResult q = qagi_upper_auto(f, R, 0.0, 1.0e-2);  // utils.py:448
  const double K2 = 2.0 * ctx[3];
  q.result = AP_MUL(q.result, K2);
  q.abserr = AP_MUL(q.abserr, K2);
  if (info) *info = q;
  return q.result;
}

// ---------------------------------------------------------------------------------------------
template <int P>
AP_HD void profile_setup(const double* p, double* ctx) {
  if (P == P_CONSTANT) {
    ctx[0] = p[0];
  } else if (P == P_LINEAR) {
    ctx[0] = p[0];
    ctx[1] = p[1];
  } else if (P == P_SOLID_SPHERE) {  // M_200c / 2 / pi * sqrt(R_200c^2 - R^2) / R_200c^3
    ctx[0] = p[0] / 2.0 / kPi;
    ctx[1] = p[1] * p[1];
    ctx[2] = p[1] * p[1] * p[1];
  } else if (P == P_GAUSSIAN) {  // exp(-(R / R_200c / 3)^2) * (x + y + z)
    ctx[0] = p[0];
    ctx[1] = p[1] + p[2] + p[3];
  } else if (P == P_NFW_RHO2D || P == P_NFW_TAU2D || P == P_NFW_KSZ) {
    double rho_s = p[0], R_s = p[1];
    ctx[0] = R_s;
    double amp = 8.0 * rho_s * R_s;
    if (P != P_NFW_RHO2D) {  // NFW.py:162-168
      const double X_H = 0.76, x_e = (X_H + 1.0) / 2.0 * X_H, f_s = 0.02;
      const double mu = 4.0 / (2.0 * X_H + 1.0 + X_H * x_e);
      amp = cst::sigma_T * x_e * X_H * (1.0 - f_s) * cst::f_b * amp / mu / cst::m_p;
    }
    if (P == P_NFW_KSZ) amp = -amp * p[2] / cst::c_kms * p[3];  // -tau * v_r / c * T_cmb
    ctx[1] = amp;
  } else if (P == P_NFW_DEFLECTION || P == P_NFW_BG) {  // NFW.py:133-148
    double c200 = p[0], R200 = p[1], M200 = p[2];
    double A = M200 * c200 * c200 / (log(1.0 + c200) - c200 / (1.0 + c200)) / 4.0 / kPi;
    double C = 16.0 * kPi * cst::Gcm2 * A / c200 / R200;
    ctx[0] = R200 / c200;   // R_s
    ctx[1] = C;
    ctx[2] = 8.0 * R200;    // suppression radius
    if (P == P_NFW_BG) {    // v = J_sph2cart(theta, phi) . (0, v_th, v_ph); transform.py:191-215
      double th = p[3], ph = p[4], v_th = p[5], v_ph = p[6], T = p[7];
      double st = sin(th), ct = cos(th), sp = sin(ph), cp = cos(ph);
      double k = -1.0 / cst::c_kms * T;
      ctx[3] = (v_th * ct * cp - v_ph * sp) * k;
      ctx[4] = (v_th * ct * sp + v_ph * cp) * k;
      ctx[5] = (-v_th * st) * k;
    }
  } else if (P == P_B16_TAU2D) {
    b16_setup(p[0], p[1], p[2], ctx);
  } else if (P == P_B16_RHOGAS2D) {
    los_setup<LOS_B16_RHO_GAS>(p[0], p[1], p[2], ctx);
  } else if (P == P_NFW_RHO2D_LOS) {
    los_setup<LOS_NFW_RHO>(p[0], p[1], 0.0, ctx);
  }
}

AP_HD double nfw_deflection(const double* ctx, double R) {
  double x = R / ctx[0];
  double f = (log(x / 2.0) + nfw_g(x)) / x;
  double t = R / ctx[2];
  return ctx[1] * f * exp(-(t * t * t));
}

template <int P>
AP_HD double profile_eval_R(const double* ctx, double R) {
  if (P == P_CONSTANT) return ctx[0];
  if (P == P_LINEAR) return ctx[0] + R * ctx[1];
  if (P == P_SOLID_SPHERE) return ctx[0] * sqrt(ctx[1] - R * R) / ctx[2];
  if (P == P_GAUSSIAN) {
    double t = R / ctx[0] / 3.0;
    return exp(-(t * t)) * ctx[1];
  }
  if (P == P_NFW_RHO2D || P == P_NFW_TAU2D || P == P_NFW_KSZ) {
    double x = R / ctx[0];
    double f = 1.0 - nfw_g(x);
    return ctx[1] * (f / ((x - 1.0) * (x + 1.0)));
  }
  if (P == P_NFW_DEFLECTION) return nfw_deflection(ctx, R);
  if (P == P_B16_TAU2D) return b16_tau_2D(ctx, R);
  if (P == P_B16_RHOGAS2D) return los_eval<LOS_B16_RHO_GAS>(ctx, R);
  if (P == P_NFW_RHO2D_LOS) return los_eval<LOS_NFW_RHO>(ctx, R);
  return 0.0;
}

template <int P>
AP_HD double profile_eval_vec(const double* ctx, double rx, double ry, double rz) {
  if (P == P_NFW_BG) {  // dT = -alpha * (R_hat . v) / c * T_cmb, NFW.py:186-192
    double R = sqrt(rx * rx + ry * ry + rz * rz);
    double alpha = nfw_deflection(ctx, R);
    double dot = (rx * ctx[3] + ry * ctx[4] + rz * ctx[5]) / R;
    return alpha * dot;
  }
  return 0.0;
}

}  // namespace ap
