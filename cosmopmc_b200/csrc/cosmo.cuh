// cosmo.cuh -- batched cosmology likelihoods (one sample per thread):
// SN Ia chi^2 (replaces nicaea SetDl + chi2_SN behind wrappers/src/sn.c:138-281),
// BAO distance ratios (bao.c:80-184), WMAP distance priors (wmap.c:945-1049).
// FP64 throughout; the Romberg node sequence and stopping rule of the
// reference are reproduced exactly.
#pragma once
#include "common.cuh"

#include "cosmo_types.cuh"
#include "fastmath.cuh"

struct Model {
  pmcb200_cosmo_t c;
  double Theta2[4], stretch, color;
};

// The `switch (like->par[i])` blocks of sn.c:167-224 / bao.c:100-147 /
// wmap.c:966-1019 followed by set_base_parameters (param.c:1544-1661).
// Returns non-zero on the reference's error conditions.
__device__ inline int apply_params(const DevLike &L, const double *__restrict__ x, Model &m) {
  double Omegam = -1, Omegab = -1, Omegac = -1, Omegade = -1, Omeganumass = -1;
  double omegam = -1, omegab = -1, omegac = -1, omegade = -1, omeganumass = -1, h100 = -1;
  double OmegaK = 0, omegaK = 0;
  int iOmegade = 0, iOmegaK = 0, iomegade = 0, iomegaK = 0, bad = 0;
  m.c = L.model;
#pragma unroll
  for (int i = 0; i < 4; i++) m.Theta2[i] = L.Theta2[i];
  m.stretch = 1.0; m.color = 0.0;
  const bool sn = (L.kind == PMCB200_LIKE_SNIa);
  for (int i = 0; i < L.npar; i++) {
    double v = x[i];
    switch (L.par[i]) {
      case PMCB200_P_Omegam: Omegam = v; break;
      case PMCB200_P_Omegab: Omegab = v; break;
      case PMCB200_P_Omegade: Omegade = v; iOmegade = 1; break;
      case PMCB200_P_Omeganumass: Omeganumass = v; break;
      case PMCB200_P_Omegac: Omegac = v; break;
      case PMCB200_P_OmegaK: OmegaK = v; iOmegaK = 1; break;
      case PMCB200_P_omegam: omegam = v; break;
      case PMCB200_P_omegab: omegab = v; break;
      case PMCB200_P_100_omegab: omegab = v / 100.0; break;
      case PMCB200_P_omegade: omegade = v; iomegade = 1; break;
      case PMCB200_P_omeganumass: if (!sn) omeganumass = v; break;
      case PMCB200_P_omegac: omegac = v; break;
      case PMCB200_P_omegaK: omegaK = v; iomegaK = 1; break;
      case PMCB200_P_w0de: m.c.w0_de = v; break;
      case PMCB200_P_w1de: m.c.w1_de = v; break;
      case PMCB200_P_h100: h100 = v; break;
      case PMCB200_P_Neffnumass: if (!sn) m.c.Neff_nu_mass = v; break;
      case PMCB200_P_M: if (sn) m.Theta2[0] = v; break;
      case PMCB200_P_alpha: if (sn) m.Theta2[1] = v; break;
      case PMCB200_P_beta: if (sn) m.Theta2[2] = v; break;
      case PMCB200_P_logbeta: if (sn) m.Theta2[2] = -exp(v); break;
      case PMCB200_P_beta_z: if (sn) m.Theta2[3] = v; break;
      case PMCB200_P_stretch: if (sn) m.stretch = v; break;
      case PMCB200_P_color: if (sn) m.color = v; break;
      default: break;
    }
    if (L.kind == PMCB200_LIKE_CMBDistPrior && !isfinite(v)) bad = 1;
  }
  if (h100 < 0) h100 = m.c.h_100; else m.c.h_100 = h100;
  if (Omegam > 0 || Omegab > 0 || iOmegade == 1 || Omeganumass > 0 || Omegac > 0 || iOmegaK == 1) {
    if (omegam > 0 || omegab > 0 || iomegade == 1 || omeganumass > 0 || omegac > 0 || iomegaK == 1)
      bad = 1;
  } else {
    if (h100 < 0) bad = 1;
    // two divisions each, as param.c:1567-1572 writes them (x/h/h and x/(h*h) differ in the last bit)
    Omegam = omegam / h100 / h100; Omegab = omegab / h100 / h100; Omegade = omegade / h100 / h100;
    Omeganumass = omeganumass / h100 / h100; Omegac = omegac / h100 / h100; OmegaK = omegaK / h100 / h100;
    iOmegade = iomegade; iOmegaK = iomegaK;
  }
  if (Omegam > 0 && iOmegade == 1 && iOmegaK == 1) bad = 1;
  if (Omegam > 0 && Omegab > 0 && Omegac > 0) bad = 1;
  if (Omeganumass < 0) Omeganumass = 0;
  if (Omegam > 0) m.c.Omega_m = Omegam;
  else if (Omegab > 0 && Omegac > 0) m.c.Omega_m = Omegab + Omegac;
  else if (Omegade > 0 && iOmegaK == 1) m.c.Omega_m = 1.0 - Omegade - OmegaK - Omeganumass;
  if (Omegab > 0) m.c.Omega_b = Omegab;
  else if (Omegam > 0 && Omegac > 0) m.c.Omega_b = Omegam - Omegac;
  if (Omegade > 0) m.c.Omega_de = Omegade;
  else if (Omegam > 0) m.c.Omega_de = 1.0 - Omegam - OmegaK - Omeganumass;
  else if (Omegab > 0 && Omegac > 0) m.c.Omega_de = 1.0 - Omegab - Omegac - OmegaK - Omeganumass;
  if (Omeganumass > 0) m.c.Omega_nu_mass = Omeganumass;
  return bad;
}

// nicaea test_range_de_conservative (sn.c:265, bao.c:156, wmap.c:1029): w(a) outside [-1, -1/3]
// at a = 1 or a = a_acc.  A violating model gets log L = 0 from the SN and BAO probes (sn.c:263-274,
// bao.c:154-176); likeli_CMBDistPrior raises wmap_de_prior instead (wmap.c:1041-1044), which drops the point.
__device__ __forceinline__ bool de_conservative_violated(const pmcb200_cosmo_t &c) {
  const double w_now = c.w0_de;
  double w_acc = w_now;
  if (c.de_param == PMCB200_DE_linder) w_acc = c.w0_de + c.w1_de * (1.0 - DE_A_ACC);
  else if (c.de_param == PMCB200_DE_jassal) w_acc = c.w0_de + c.w1_de * DE_A_ACC * (1.0 - DE_A_ACC);
  return w_now < -1.0 || w_now > -1.0 / 3.0 || w_acc < -1.0 || w_acc > -1.0 / 3.0;
}

// SN integrand's 2^s: s = k/1024 + f, |f| <= 1/2048, 2^f by a degree-3 Chebyshev interpolant (max rel.
// error 1.4e-16), 2^(j/1024) from a 1024-entry table staged in shared memory.  The table's high
// words have j << 10 subtracted: adding k32 << 10 (k32 = 1024 n + j) to that high word in ONE
// integer multiply-add lands n in the exponent field and cancels the j part.  Filled once per device
// by pmc_init_sn_tables() (host long-double exp2l, rounded to double).
#define SN_EXP2_N 1024
__device__ double g_sn_exp2[SN_EXP2_N];
__constant__ double EXP2D3[4] = {0x1.62e42fefa39efp-1, 0x1.ebfbe033445b4p-3, 0x1.c6b08d910ecbdp-5,
                                 6597069766656.0};   // [3] 1.5 * 2^42: ulp = 2^-10
// the SN kernel's table: the three above + the pre-biased 1024-entry exp2 table
__device__ __forceinline__ void load_fast_tables_sn(double *T) {
  for (int i = threadIdx.x; i < SN_EXP2_N; i += blockDim.x) T[96 + i] = g_sn_exp2[i];
  load_fast_tables(T);
}


// Coefficients of a^4 E^2(a) = a (Om + OK a) + Or + Ode exp(p ln a + q g(a)),
// g = (1-a) [linder] or (1-a)^2 [jassal].
struct ECoef {
  double Om, OK, Ode, Or, p, q;
  int jassal;
};
__device__ __forceinline__ ECoef make_ecoef(const pmcb200_cosmo_t &c, int wOmegar) {
  ECoef e;
  e.Om = c.Omega_m + c.Omega_nu_mass;
  e.OK = 1.0 - c.Omega_m - c.Omega_de - c.Omega_nu_mass;
  e.Ode = c.Omega_de;
  e.Or = wOmegar ? OMEGA_GAMMA_H2 * (1.0 + 0.2271 * NEFF_NU) / (c.h_100 * c.h_100) : 0.0;
  e.jassal = (c.de_param == PMCB200_DE_jassal);
  if (e.jassal) { e.p = 4.0 - 3.0 * (1.0 + c.w0_de); e.q = 1.5 * c.w1_de; }
  else { e.p = 4.0 - 3.0 * (1.0 + c.w0_de + c.w1_de); e.q = -3.0 * c.w1_de; }
  return e;
}
__device__ __forceinline__ double a4E2(const ECoef &e, double a, double lna) {
  double oma = 1.0 - a;
  double t = e.jassal ? fma(e.q * oma, oma, e.p * lna) : fma(e.q, oma, e.p * lna);
  double de = (a > 0.0) ? e.Ode * exp(t) : 0.0;
  return fma(a, fma(e.OK, a, e.Om), e.Or) + de;
}
// sinh(x)/x for u = x^2 > 0 and sin(x)/x for u = -x^2 < 0: one series in u, |u| < 1
// (truncation 1/21! = 2e-20); replaces the branchy libdevice sinh / sin
__device__ __forceinline__ double sinhc_series(double u) {
  double p = 1.0 / 121645100408832000.0;      // 1/19!
  p = fma(p, u, 1.0 / 355687428096000.0);     // 1/17!
  p = fma(p, u, 1.0 / 1307674368000.0);       // 1/15!
  p = fma(p, u, 1.0 / 6227020800.0);          // 1/13!
  p = fma(p, u, 1.0 / 39916800.0);            // 1/11!
  p = fma(p, u, 1.0 / 362880.0);              // 1/9!
  p = fma(p, u, 1.0 / 5040.0);                // 1/7!
  p = fma(p, u, 1.0 / 120.0);                 // 1/5!
  p = fma(p, u, 1.0 / 6.0);                   // 1/3!
  return fma(p, u, 1.0);
}
// transverse comoving distance from the comoving distance w [Mpc/h]
__device__ __forceinline__ double f_K_from(double OK, double w) {
  if (fabs(OK) < FLAT_EPS) return w;
  const double x = w * (1.0 / R_HUBBLE), u = OK * x * x;
  if (fabs(u) < 1.0) return w * sinhc_series(u);
  const double sk = sqrt(fabs(OK)) / R_HUBBLE;
  return OK > 0.0 ? sinh(sk * w) / sk : sin(sk * w) / sk;
}
__device__ __forceinline__ double f_K(const pmcb200_cosmo_t &c, double w) {
  return f_K_from(1.0 - c.Omega_m - c.Omega_de - c.Omega_nu_mass, w);
}
// a^4 E^2(a) with the table-based primitives: base-2 exponent folded with log2|Ode|;
// outside the fast range (|s| >= 990, Ode == 0, a == 0) the libdevice path is taken
struct ECoefF {
  ECoef e;
  double p2, q2, lg;
  unsigned sgn;
  int ok;       // Ode != 0 and finite
};
__device__ __forceinline__ ECoefF make_ecoef_fast(const pmcb200_cosmo_t &c, int wOmegar) {
  ECoefF f;
  f.e = make_ecoef(c, wOmegar);
  f.p2 = f.e.p * M_LOG2E; f.q2 = f.e.q * M_LOG2E;
  f.lg = log2(fabs(f.e.Ode));
  f.sgn = (f.e.Ode < 0.0) ? 0x80000000u : 0u;
  f.ok = isfinite(f.lg) && isfinite(f.p2) && isfinite(f.q2);
  return f;
}
__device__ __forceinline__ double a4E2_fast(const ECoefF &f, double a, const double *__restrict__ T) {
  if (!(a > 0.0)) return f.e.Or;
  const double oma = 1.0 - a;
  double de;
  if (f.ok) {
    const double lna = fast_log(a, T);
    const double s = f.e.jassal ? fma(f.q2 * oma, oma, fma(f.p2, lna, f.lg)) : fma(f.q2, oma, fma(f.p2, lna, f.lg));
    de = (fabs(s) < 990.0) ? fast_exp2_signed(s, T, f.sgn) : f.e.Ode * exp2(s - f.lg);
  } else {
    const double lna = log(a);
    const double t = f.e.jassal ? fma(f.e.q * oma, oma, f.e.p * lna) : fma(f.e.q, oma, f.e.p * lna);
    de = f.e.Ode * exp(t);
  }
  return fma(a, fma(f.e.OK, a, f.e.Om), f.e.Or) + de;
}
// comoving distance [Mpc/h] with on-the-fly nodes
__device__ inline double w_generic(const pmcb200_cosmo_t &c, double a, int wOmegar, int &err, const double *__restrict__ T) {
  const ECoefF f = make_ecoef_fast(c, wOmegar);
  int bad = 0;
  double r = romberg([&](double x) {
    double dd = a4E2_fast(f, x, T);
    if (!(dd > 0.0)) bad = 1;
    return fast_rsqrt(dd);
  }, a, 1.0, err);
  if (bad) err = 1;
  return R_HUBBLE * r;
}
__device__ inline double r_sound(const pmcb200_cosmo_t &c, double a, int &err, const double *__restrict__ T) {
  const ECoefF f = make_ecoef_fast(c, 1);
  const double Rfac = 0.75 * c.Omega_b * c.h_100 * c.h_100 / OMEGA_GAMMA_H2;
  int bad = 0;
  double r = romberg([&](double x) {
    double dd = a4E2_fast(f, x, T) * 3.0 * fma(Rfac, x, 1.0);
    if (!(dd > 0.0)) bad = 1;
    return fast_rsqrt(dd);
  }, 0.0, a, err);
  if (bad) err = 1;
  return R_HUBBLE * r;
}
__device__ inline double D_V(const pmcb200_cosmo_t &c, double z, int &err, const double *__restrict__ T) {
  double a = 1.0 / (1.0 + z);
  double ww = w_generic(c, a, 0, err, T);
  double fK = f_K(c, ww);
  const ECoefF f = make_ecoef_fast(c, 0);
  double a2 = a * a;
  double EE = a4E2_fast(f, a, T) / (a2 * a2);
  if (!(EE > 0.0)) { err = 1; return NAN; }
  return cbrt(fK * fK * R_HUBBLE * z / sqrt(EE));
}
// The fitting formulae take eleven powers of two bases: x^c = 2^(c log2 x) with ONE logarithm per base (the powers were
// 20 % of k_like_bao's instructions as libdevice pow calls: profiles/r02/c4_k_like_bao_summary.txt).  |c log2 x| < 8, so
// the result carries a few ulp; the callers have checked omega_m > 0 (omega_b = 0: 2^(-inf) = 0 = pow(0, c > 0)).
__device__ __forceinline__ double pow2l(double l2x, double c) { return exp2(c * l2x); }
__device__ inline double z_drag(const pmcb200_cosmo_t &c) {
  const double omm = c.Omega_m * c.h_100 * c.h_100, omb = c.Omega_b * c.h_100 * c.h_100;
  const double lm = log2(omm), lb = log2(omb);
  const double b1 = 0.313 * pow2l(lm, -0.419) * (1.0 + 0.607 * pow2l(lm, 0.674));
  const double b2 = 0.238 * pow2l(lm, 0.223);
  return 1291.0 * pow2l(lm, 0.251) / (1.0 + 0.659 * pow2l(lm, 0.828)) * (1.0 + b1 * pow2l(lb, b2));
}
__device__ inline double z_star(const pmcb200_cosmo_t &c) {
  const double omm = c.Omega_m * c.h_100 * c.h_100, omb = c.Omega_b * c.h_100 * c.h_100;
  const double lm = log2(omm), lb = log2(omb);
  const double g1 = 0.0783 * pow2l(lb, -0.238) / (1.0 + 39.5 * pow2l(lb, 0.763));
  const double g2 = 0.560 / (1.0 + 21.1 * pow2l(lb, 1.81));
  return 1048.0 * (1.0 + 0.00124 * pow2l(lb, -0.738)) * (1.0 + g1 * pow2l(lm, g2));
}
// Gaussian log-pdf of a model vector against data packed as a component
__device__ inline double gauss_comp_logpdf(const double *__restrict__ comp, int n, const double *model) {
  const double *mean = comp + 2, *L = comp + 2 + n, *rd = comp + 2 + n + mix_tri(n);
  double y[4], m = 0.0;
  int off = 0;
  for (int i = 0; i < n; i++) {
    double t = model[i] - mean[i];
    for (int k = 0; k < i; k++) t = fma(-L[off + k], y[k], t);
    y[i] = t * rd[i];
    m = fma(y[i], y[i], m);
    off += i + 1;
  }
  return fma(-0.5, m, comp[1]);
}

// ---- lean integrand of the BAO / CMB integrals (k_like_bao, k_like_cmbdp) ------------------------------
// These integrals have sample-dependent limits (a_star, a_drag, a(z_BAO)), so their Romberg nodes cannot
// be tabulated the way the SN kernel's are: every node needs ln a.  The comoving distance to a_star alone
// takes 11 stages = 1025 evaluations per sample (measured stage histogram), so the
// per-evaluation instruction count IS the kernel.  Per interior node (a > 0):
//   ln a   1024-entry {1/c_i, -ln(1/c_i)} table on the top ten mantissa bits, r = m/c_i - 1 (|r| < 2^-11),
//          degree-4 log1p (truncation r^5/5 < 6e-18): 6 FP64 + 1 I2F (the 32-entry version: 12)
//   2^s    the SN kernel's pre-biased 1024-entry table + degree-3 polynomial: 6 FP64 + 1 IMAD (was 9 + LDS)
//   no per-node range / sign tests: the bound on |s| over the whole interval is checked once per sample,
//   and a non-positive radicand turns into NaN through MUFU.RSQ64H and reaches the Romberg result
#define LOG1K_N 1024
__device__ double2 g_log1k[LOG1K_N];       // filled by pmc_init_sn_tables()
__constant__ double LOG1KP[4] = {-0.25, 1.0 / 3.0, -0.5, 0.693147180559945309417};
struct GLean {
  double Om, OK, Or, p2, q2, lg, lgq;      // lg = log2|Ode|, lgq = lg + q2 (linder: s = p2 ln a - q2 a + lgq)
  unsigned sgn;
  int jassal, ok;
};
__device__ __forceinline__ GLean make_glean(const ECoefF &f) {
  GLean g;
  g.Om = f.e.Om; g.OK = f.e.OK; g.Or = f.e.Or; g.p2 = f.p2; g.q2 = f.q2; g.lg = f.lg; g.lgq = f.lg + f.q2;
  g.sgn = f.sgn; g.jassal = f.e.jassal; g.ok = f.ok;
  return g;
}
// fast path valid for every node a >= amin of the interval: |s| stays inside the table-based exp2's range
__device__ __forceinline__ bool glean_in_range(const GLean &g, double amin) {
  return g.ok && g.sgn == 0u && (fabs(g.p2) * fabs(log(amin)) + 2.0 * fabs(g.q2) + fabs(g.lg) < 990.0);
}
__device__ __forceinline__ double lean_log(double x, const double2 *__restrict__ LT) {
  const int hi = __double2hiint(x);
  const double2 tc = LT[(hi >> 10) & (LOG1K_N - 1)];
  const double m = __hiloint2double((hi & 0x000fffff) | 0x3ff00000, __double2loint(x));
  const double r = fma(m, tc.x, -1.0);
  double q = fma(LOG1KP[0], r, LOG1KP[1]);
  q = fma(q, r, LOG1KP[2]);
  q = fma(q, r, 1.0);
  const double e = (double)((hi >> 20) - 1023);
  return fma(e, LOG1KP[3], fma(q, r, tc.y));
}
// a^4 E^2(a) at an interior node; ET = the pre-biased 2^(j/1024) table (g_sn_exp2)
template <bool HASQ>
__device__ __forceinline__ double a4E2_lean(const GLean &g, const double2 *__restrict__ LT, const double *__restrict__ ET,
                                            double a) {
  const double lna = lean_log(a, LT);
  double s;
  if (HASQ) {
    if (g.jassal) { const double oma = 1.0 - a; s = fma(g.q2 * oma, oma, fma(g.p2, lna, g.lg)); }
    else s = fma(g.p2, lna, fma(-g.q2, a, g.lgq));
  } else s = fma(g.p2, lna, g.lg);
  const double kf = s + EXP2D3[3];
  const int k32 = __double2loint(kf);
  const double f = s - (kf - EXP2D3[3]);
  double p = EXP2D3[2];
  p = fma(p, f, EXP2D3[1]);
  p = fma(p, f, EXP2D3[0]);
  p = fma(p, f, 1.0);
  const double tj = ET[k32 & (SN_EXP2_N - 1)];
  const int hi = __double2hiint(tj) + k32 * SN_EXP2_N;      // Ode > 0 on this path (glean_in_range)
  return fma(__hiloint2double(hi, __double2loint(tj)), p, fma(a, fma(g.OK, a, g.Om), g.Or));
}
// acc + 1/sqrt(v): MUFU.RSQ64H seed + third-order correction (as sn_f); v <= 0 gives NaN
__device__ __forceinline__ double rsqrt_acc(double v, double acc) {
  double y;
  asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(v));
  const double t = v * y;
  const double er = fma(-t, y, 1.0);
  const double c = fma(0.375, er, 0.5);
  const double w = fma(er, c, 1.0);
  return fma(y, w, acc);
}
// One integral of the BAO / CMB kernels: comoving distance (RS = false) or sound horizon (RS = true: the radicand
// is multiplied by 3 (1 + R a), R3 = 3 R_fac)
struct GInt { GLean g; double R3; };
template <bool HASQ, bool RS>
__device__ __forceinline__ double gint_acc(const GInt &q, const double2 *__restrict__ LT, const double *__restrict__ ET,
                                           double x, double acc) {
  double v = a4E2_lean<HASQ>(q.g, LT, ET, x);
  if (RS) v *= fma(q.R3, x, 3.0);
  return rsqrt_acc(v, acc);
}
__device__ __forceinline__ double shfl_d(double v, int src) { return __shfl_sync(0xffffffffu, v, src); }
// the fields gint_acc reads, from lane src
__device__ __forceinline__ GInt gint_from_lane(const GInt &q, int src) {
  GInt r;
  r.g.Om = shfl_d(q.g.Om, src); r.g.OK = shfl_d(q.g.OK, src); r.g.Or = shfl_d(q.g.Or, src);
  r.g.p2 = shfl_d(q.g.p2, src); r.g.q2 = shfl_d(q.g.q2, src); r.g.lg = shfl_d(q.g.lg, src); r.g.lgq = shfl_d(q.g.lgq, src);
  r.g.sgn = 0u; r.g.jassal = __shfl_sync(0xffffffffu, q.g.jassal, src); r.g.ok = 1;
  r.R3 = shfl_d(q.R3, src);
  return r;
}
// NR qromb, warp-synchronous: EVERY lane of the warp calls this together (lanes without an integral pass
// active = false).  Stages with fewer than ROMB_COOP_MIN new nodes: every lane evaluates its own nodes, four at a
// time into four partial sums (independent chains; the reference sums sequentially: the difference is O(1e-16)).
// Deeper stages -- the rare integral that needs them (dark energy dominating the early universe makes the
// sound-horizon integrand singular and the rule runs to 2^19 nodes; round 1 let one lane grind through them while
// 31 waited and the kernel waited for that warp: 137 ms for 1e7 BAO samples that need 2 ms) -- are evaluated by
// the WHOLE WARP for one lane at a time: the owner's coefficients travel by shuffles, lane l takes the nodes
// l, l + 32, ... and a butterfly sum hands the stage sum back (the same bits in every lane, and a function of
// the owner's sample alone: results do not depend on the warp's other samples).
#define ROMB_COOP_MIN 1024
#ifndef ROMB_COOP_MIN_RS
#define ROMB_COOP_MIN_RS ROMB_COOP_MIN      // the same threshold for the sound-horizon integral (build switch for A/B measurements)
#endif
template <bool HASQ, bool RS>
__device__ double romberg_warp(const GInt &q, const double2 *__restrict__ LT, const double *__restrict__ ET, double fa,
                               double fb, double a, double b, bool active, int &err, unsigned &nev,
                               int coop_min = (RS ? ROMB_COOP_MIN_RS : ROMB_COOP_MIN)) {      // >= 128, warp-uniform
  const unsigned FULL = 0xffffffffu;
  const int lane = threadIdx.x & 31;
  double y[5];
  const double h = b - a;
  double st = 0.5 * h * (fa + fb);
  y[0] = st;
  double ss = st, dss;
  bool done = !active;
  if (active) nev += 2;
  for (int j = 1;; j++) {                         // j is warp-uniform: a lane only ever stops early
    const int it = 1 << (j - 1);
    const double tnm = (double)it, del = h / tnm;
    double sum = 0.0;
    if (it < coop_min) {
      if (!done) {
        if (it < 4) {
          for (int i = 0; i < it; i++) sum = gint_acc<HASQ, RS>(q, LT, ET, fma((double)i + 0.5, del, a), sum);
        } else {
          // x += 4 del per chain, as NR's trapzd steps x += del (one rounding per step either way)
          double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
          double x0 = fma(0.5, del, a), x1 = fma(1.5, del, a), x2 = fma(2.5, del, a), x3 = fma(3.5, del, a);
          const double d4 = 4.0 * del;
          for (int i = 0; i < it; i += 4) {
            s0 = gint_acc<HASQ, RS>(q, LT, ET, x0, s0); s1 = gint_acc<HASQ, RS>(q, LT, ET, x1, s1);
            s2 = gint_acc<HASQ, RS>(q, LT, ET, x2, s2); s3 = gint_acc<HASQ, RS>(q, LT, ET, x3, s3);
            x0 += d4; x1 += d4; x2 += d4; x3 += d4;
          }
          sum = (s0 + s1) + (s2 + s3);
        }
      }
    } else {
      unsigned mask = __ballot_sync(FULL, !done);
      while (mask) {
        const int src = __ffs(mask) - 1;
        mask &= mask - 1;
        const GInt qs = gint_from_lane(q, src);
        const double as = shfl_d(a, src), dels = shfl_d(del, src);
        const double step = 32.0 * dels, d4 = 4.0 * step;
        double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
        double x0 = fma((double)lane + 0.5, dels, as), x1 = x0 + step, x2 = x1 + step, x3 = x2 + step;
        for (int k = 0; k < (it >> 5); k += 4) {
          s0 = gint_acc<HASQ, RS>(qs, LT, ET, x0, s0); s1 = gint_acc<HASQ, RS>(qs, LT, ET, x1, s1);
          s2 = gint_acc<HASQ, RS>(qs, LT, ET, x2, s2); s3 = gint_acc<HASQ, RS>(qs, LT, ET, x3, s3);
          x0 += d4; x1 += d4; x2 += d4; x3 += d4;
        }
        const double tot = warp_sum((s0 + s1) + (s2 + s3));
        if (lane == src) sum = tot;
      }
    }
    if (!done) {
      nev += it;
      st = 0.5 * (st + h * sum / tnm);
      if (j < 5) y[j] = st;
      else { y[0] = y[1]; y[1] = y[2]; y[2] = y[3]; y[3] = y[4]; y[4] = st; }
      if (j >= 4) {
        ss = romb_extrap(y, dss);
        if (!isfinite(ss)) { err = 1; done = true; }
        else if (fabs(dss) <= ROMB_EPS * fabs(ss)) done = true;
      }
      if (!done && j + 1 >= ROMB_JMAX) { err = 1; done = true; }      // too many steps
    }
    if (__all_sync(FULL, done)) break;
  }
  return ss;
}
struct LeanTabs { const double *T; const double2 *LT; const double *ET; };
// comoving distance [Mpc/h], a .. 1.  Warp-synchronous: all lanes call it; act = this lane wants the value.
template <bool HASQ>
__device__ __forceinline__ double w_lean(const pmcb200_cosmo_t &c, double a, int wOmegar, bool act, int &err,
                                         const LeanTabs &tb, unsigned &nev, int coop_min = ROMB_COOP_MIN) {
  const ECoefF f = make_ecoef_fast(c, wOmegar);
  GInt q;
  q.g = make_glean(f); q.R3 = 0.0;
  const bool fast = act && (a > 0.0) && glean_in_range(q.g, a);
  double fa = 0.0, fb = 0.0;
  if (fast) { fa = rsqrt_acc(a4E2_fast(f, a, tb.T), 0.0); fb = rsqrt_acc(a4E2_fast(f, 1.0, tb.T), 0.0); }
  double r = R_HUBBLE * romberg_warp<HASQ, false>(q, tb.LT, tb.ET, fa, fb, a, 1.0, fast, err, nev, coop_min);
  if (act && !fast) r = w_generic(c, a, wOmegar, err, tb.T);      // outside the table-based exp2's range: general path
  return r;
}
// comoving sound horizon [Mpc/h], 0 .. a (radiation included); warp-synchronous like w_lean
template <bool HASQ>
__device__ __forceinline__ double r_sound_lean(const pmcb200_cosmo_t &c, double a, bool act, int &err, const LeanTabs &tb,
                                               unsigned &nev) {
  const ECoefF f = make_ecoef_fast(c, 1);
  GInt q;
  q.g = make_glean(f);
  const double Rfac = 0.75 * c.Omega_b * c.h_100 * c.h_100 / OMEGA_GAMMA_H2;
  q.R3 = 3.0 * Rfac;
  // the smallest interior node of the deepest stage (ROMB_JMAX = 20) is a 2^-20
  const bool fast = act && (a > 0.0) && glean_in_range(q.g, a * (1.0 / 1048576.0));
  double fa = 0.0, fb = 0.0;
  if (fast) { fa = rsqrt_acc(f.e.Or * 3.0, 0.0); fb = rsqrt_acc(a4E2_fast(f, a, tb.T) * fma(q.R3, a, 3.0), 0.0); }
  double r = R_HUBBLE * romberg_warp<HASQ, true>(q, tb.LT, tb.ET, fa, fb, 0.0, a, fast, err, nev);
  if (act && !fast) r = r_sound(c, a, err, tb.T);
  return r;
}
template <bool HASQ>
__device__ __forceinline__ double D_V_lean(const pmcb200_cosmo_t &c, double z, bool act, int &err, const LeanTabs &tb,
                                           unsigned &nev) {
  const double a = 1.0 / (1.0 + z);
  const double ww = w_lean<HASQ>(c, a, 0, act, err, tb, nev);
  if (!act) return 1.0;
  const double fK = f_K(c, ww);
  const ECoefF f = make_ecoef_fast(c, 0);
  const double a2 = a * a;
  const double EE = a4E2_fast(f, a, tb.T) / (a2 * a2);
  if (!(EE > 0.0)) { err = 1; return NAN; }
  return cbrt(fK * fK * R_HUBBLE * z / sqrt(EE));
}
// stage the three tables: T[96] (general path), LT[1024] (log), ET[1024] (exp2); blockDim.x >= 96
__device__ __forceinline__ void load_lean_tables(double *T, double2 *LT, double *ET) {
  for (int i = threadIdx.x; i < LOG1K_N; i += blockDim.x) { LT[i] = g_log1k[i]; ET[i] = g_sn_exp2[i]; }
  load_fast_tables(T);
}

// write/accumulate one likelihood term
__device__ __forceinline__ void put_loglike(double *logpi, int32_t *err, int64_t n, int set,
                                            double add_const, double res, int e) {
  if (set) { logpi[n] = res + add_const; if (err) err[n] = e; }
  else { logpi[n] += res; if (err && e) err[n] = 1; }
}

// ---- SN Ia: one sample per thread; the whole warp walks the same redshift
// so node loads are warp-uniform and the adaptive stage count is resolved by
// a warp vote.  HASQ: w1 != 0 or jassal (second exponent term); FLAT: the
// curvature term is identically zero for every sample of the launch; SLOW:
// libdevice exp for warps holding a sample outside the fast path's range. -------
struct SNCoef {
  double Om, OK, Ode, p, q;        // a^3 E^2 = Om + OK a + Ode exp(p ln a + q g(a)),  p = -3 (w0 + w1) [linder]
  double Oms, OKs, p2, q2;         // fast form, scaled by 1/|Ode|: Oms + OKs a + sgn 2^(p2 ln a + q2 g(a))
  double scale;                    // |Ode|^-1/2 (fast): the integral of the scaled integrand times this
  unsigned sgn;                    // sign of Ode as an XOR mask (NEG instantiation: some lane has Ode < 0)
  int jassal, slow;
};
// Integrand of the comoving distance at a tabulated node, 1/sqrt(a^4 E^2) = a^-1/2 / sqrt(Q) with
// Q = a^3 E^2: the a^-1/2 factor is tabulated with the node, so the sample-dependent part is one
// 2^s (its last multiply fused into Q = T 2^f + Om) and one 1/sqrt.  The fast path works on
// Q / |Ode| (the exponent needs no additive constant; |Ode|^-1/2 multiplies the finished integral).
// 1/sqrt: MUFU.RSQ64H seed y (measured |1 - Q y^2| < 2^-19, tools/micro/rsqrt_err.cu) and a
// third-order correction written so that no FP64 instruction reads three varying registers
// except the final accumulate.  Returns acc + a^-1/2 / sqrt(Q).
template <bool HASQ, bool FLAT, bool SLOW, bool NEG>
__device__ __forceinline__ double sn_f(const SNCoef &e, const double *__restrict__ T, double lna, double rsa, double a,
                                       double acc) {
  double Q;
  if (SLOW) {
    double t = e.p * lna;
    if (HASQ) { double oma = 1.0 - a; t = e.jassal ? fma(e.q * oma, oma, t) : fma(e.q, oma, t); }
    Q = fma(e.Ode, exp(t), FLAT ? e.Om : fma(e.OK, a, e.Om));
  } else {
    double kf, f;
    if (HASQ) {
      const double oma = 1.0 - a;
      const double s = e.jassal ? fma(e.q2 * oma, oma, e.p2 * lna) : fma(e.q2, oma, e.p2 * lna);
      kf = s + EXP2D3[3];
      f = s - (kf - EXP2D3[3]);
    } else {      // the product p2 ln a is never formed: both uses are fused
      kf = fma(e.p2, lna, EXP2D3[3]);
      f = fma(e.p2, lna, -(kf - EXP2D3[3]));
    }
    const int k32 = __double2loint(kf);
    double p = EXP2D3[2];
    p = fma(p, f, EXP2D3[1]);
    p = fma(p, f, EXP2D3[0]);
    p = fma(p, f, 1.0);
    const double tj = T[96 + (k32 & (SN_EXP2_N - 1))];
    int hi = __double2hiint(tj) + k32 * SN_EXP2_N;
    if (NEG) hi ^= e.sgn;
    Q = fma(__hiloint2double(hi, __double2loint(tj)), p, FLAT ? e.Oms : fma(e.OKs, a, e.Oms));
  }
  double y;
  asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(Q));
  const double ry = rsa * y;
  const double t = Q * y;
  const double er = fma(-t, y, 1.0);
  const double c = fma(0.375, er, 0.5);
  const double w = fma(er, c, 1.0);
  return fma(ry, w, acc);
}
// 32-byte global load (LDG.256): two {ln a, a^-1/2} nodes, or one {ln a, a^-1/2, a, -} node
struct Ld4 { double x, y, z, w; };
__device__ __forceinline__ Ld4 ld256(const void *p) {
  Ld4 r;
  asm("ld.global.nc.v4.f64 {%0,%1,%2,%3}, [%4];" : "=d"(r.x), "=d"(r.y), "=d"(r.z), "=d"(r.w) : "l"(p));
  return r;
}
// Stages 1..5 of NR qromb at redshift iz in closed form.  With g0 = (f(a_z) + f(1))/2 and the sums
// S_j of the 2^(j-1) new nodes of stage j+1, T_j = h 2^(1-j) (g0 + S_1 + .. + S_(j-1)), so the K = 5
// extrapolant ss = R[5][4] and dss = R[5][4] - R[5][3] are fixed linear combinations (exact
// rational weights of the Neville tableau for step ratios 1/4).  The 16 tabulated nodes are
// independent evaluations (issued together for ILP).
__constant__ double ROMBW[10] = {3937.0 / 103275.0, 3062.0 / 80325.0, 27728.0 / 722925.0, 22016.0 / 722925.0,
                                  65536.0 / 722925.0, -31.0 / 206550.0, -73.0 / 481950.0, -67.0 / 722925.0,
                                  -424.0 / 722925.0, 256.0 / 722925.0};
struct SNSums { double g0, S1, S2, S3, S4; };
// SN_CHAIN_SUMS (build variant, default off): the sums S_j of the new nodes of a stage are formed by chaining
// the integrand's final accumulate (fma(ry, w, acc)) through the stage's nodes, NR's sequential order, instead of
// 15 separate additions per redshift.  Changes the last bits (summation order), not the algorithm.  Measured
// once (profiles/sn_r01_v8_hotspots.txt): 52.45 vs 52.93 ms per 1e7 samples; not adopted in round 1 because the
// parity suite could not be re-run on the GPU afterwards.
#ifndef SN_CHAIN_SUMS
#define SN_CHAIN_SUMS 0
#endif
// accumulate operand of tabulated node i: the running sum of its stage (stages start at i = 1, 2, 4, 8)
#if SN_CHAIN_SUMS
#define SN_ACC(i, fv, f1) ((i) == 0 ? (f1) : (((i) & ((i) - 1)) == 0 ? 0.0 : (fv)[(i) - 1]))
#else
#define SN_ACC(i, fv, f1) ((i) == 0 ? (f1) : 0.0)
#endif
template <bool HASQ, bool FLAT, bool SLOW, bool NEG>
__device__ __forceinline__ void sn_romb5(const SNCoef &ec, const double *__restrict__ T,
                                         const double2 *__restrict__ nd, const double *__restrict__ na,
                                         const double *__restrict__ n4, double f1, SNSums &S, double &ss,
                                         double &dss, double &lnaz, double &h) {
  double fv[16];
  if (HASQ || !FLAT) {          // one 32-byte node {ln a, a^-1/2, a, -} per load
#pragma unroll
    for (int i = 0; i < 16; i++) {
      const Ld4 n = ld256(n4 + 4 * i);
      if (i == 0) { lnaz = n.x; h = 1.0 - n.z; }
      fv[i] = sn_f<HASQ, FLAT, SLOW, NEG>(ec, T, n.x, n.y, n.z, SN_ACC(i, fv, f1));
    }
  } else {                      // two 16-byte nodes {ln a, a^-1/2} per load
#pragma unroll
    for (int i = 0; i < 16; i += 2) {
      const Ld4 n = ld256(nd + i);
      if (i == 0) { lnaz = n.x; h = 1.0 - __ldg(na); }
      fv[i] = sn_f<HASQ, FLAT, SLOW, NEG>(ec, T, n.x, n.y, 0.0, SN_ACC(i, fv, f1));
      fv[i + 1] = sn_f<HASQ, FLAT, SLOW, NEG>(ec, T, n.z, n.w, 0.0, SN_ACC(i + 1, fv, f1));
    }
  }
  S.g0 = 0.5 * fv[0];
#if SN_CHAIN_SUMS
  // build variant: each stage's sum was chained through the integrand's final FMA (NR's sequential order)
  S.S1 = fv[1]; S.S2 = fv[3]; S.S3 = fv[7]; S.S4 = fv[15];
#else
  S.S1 = fv[1];
  S.S2 = fv[2] + fv[3];
  S.S3 = (fv[4] + fv[5]) + (fv[6] + fv[7]);
  S.S4 = ((fv[8] + fv[9]) + (fv[10] + fv[11])) + ((fv[12] + fv[13]) + (fv[14] + fv[15]));
#endif
  ss = h * fma(ROMBW[0], S.g0, fma(ROMBW[1], S.S1, fma(ROMBW[2], S.S2, fma(ROMBW[3], S.S3, ROMBW[4] * S.S4))));
  dss = h * fma(ROMBW[5], S.g0, fma(ROMBW[6], S.S1, fma(ROMBW[7], S.S2, fma(ROMBW[8], S.S3, ROMBW[9] * S.S4))));
}

// The redshift loop: per unique z one adaptive Romberg integral (tabulated
// nodes), then the chi^2 terms of the supernovae at that z.
struct SNPer {           // per-sample constants of the chi^2 part
  double Theta0, Theta3, t1, t2base, d1, d2, stretch, color;
  double base0, k1, k2, k3, k4, k5;   // folded forms for the non-betaz modes
};
template <bool HASQ, bool FLAT, bool SLOW, bool NEG>
__device__ __forceinline__ void sn_zloop(const DevLike &L, const SNCoef &ec, const double *__restrict__ T,
                                         const SNPer &m_, double f1, double rh, double &chi2, double &logdet,
                                         int &e, unsigned &nev) {
  const bool flat = fabs(ec.OK) < FLAT_EPS;
  const int mode = L.sn_chi2mode;
  for (int iz = 0; iz < L.sn_nz; iz++) {
    const double2 *__restrict__ nd = L.nodes + (size_t)iz * SN_NODES;
    const double *__restrict__ na = L.nodes_a + (size_t)iz * SN_NODES;
    SNSums S;
    double ss, dss, lnaz, h;
    sn_romb5<HASQ, FLAT, SLOW, NEG>(ec, T, nd, na, L.nodes4 + (size_t)iz * (4 * 16), f1, S, ss, dss, lnaz, h);
    const double g0 = S.g0, S1 = S.S1, S2 = S.S2, S3 = S.S3, S4 = S.S4;
    bool done = !isfinite(ss) || (fabs(dss) <= ROMB_EPS * fabs(ss));
    nev += 17;
    int j = 5;                      // stages completed
    double R0 = 0.0, R1 = 0.0, R2 = 0.0, R3 = 0.0, st = 0.0;
    if (!__all_sync(0xffffffffu, done)) {   // rare: rebuild the tableau row of stage 5 to continue
      const double T1 = h * g0, T2 = 0.5 * h * (g0 + S1), T3 = 0.25 * h * (g0 + S1 + S2),
                   T4 = 0.125 * h * (g0 + S1 + S2 + S3), T5 = 0.0625 * h * (g0 + S1 + S2 + S3 + S4);
      double a1 = fma(T2 - T1, 1.0 / 3.0, T2);
      double b1 = fma(T3 - T2, 1.0 / 3.0, T3), b2 = fma(b1 - a1, 1.0 / 15.0, b1);
      double c1 = fma(T4 - T3, 1.0 / 3.0, T4), c2 = fma(c1 - b1, 1.0 / 15.0, c1), c3 = fma(c2 - b2, 1.0 / 63.0, c2);
      double d1 = fma(T5 - T4, 1.0 / 3.0, T5), d2 = fma(d1 - c1, 1.0 / 15.0, d1), d3 = fma(d2 - c2, 1.0 / 63.0, d2);
      (void)c3;
      st = T5; R0 = T5; R1 = d1; R2 = d2; R3 = d3;
    }
    while (!__all_sync(0xffffffffu, done)) {
      if (j >= ROMB_JMAX) { if (!done) { ss = NAN; done = true; } break; }
      if (!done) {
        const int it = 1 << (j - 1);
        double s = 0.0;
        if (2 * it <= SN_NODES) {
#pragma unroll 4
          for (int i = it; i < 2 * it; i++) {
            const double2 n = __ldg(&nd[i]);
            s = sn_f<HASQ, FLAT, SLOW, NEG>(ec, T, n.x, n.y, (HASQ || !FLAT) ? __ldg(&na[i]) : 0.0, s);
          }
        } else {
          const double del = h / (double)it, az = __ldg(na);
          for (int i = 0; i < it; i++) {
            double a = fma((double)i + 0.5, del, az);
            s = sn_f<HASQ, FLAT, SLOW, NEG>(ec, T, log(a), rsqrt(a), a, s);
          }
        }
        nev += it;
        st = 0.5 * (st + h * s / (double)it);
        double n1 = fma(st - R0, 1.0 / 3.0, st), n2 = fma(n1 - R1, 1.0 / 15.0, n1),
               n3 = fma(n2 - R2, 1.0 / 63.0, n2);
        dss = (n3 - R3) * (1.0 / 255.0);
        ss = n3 + dss;
        R0 = st; R1 = n1; R2 = n2; R3 = n3;
        done = !isfinite(ss) || (fabs(dss) <= ROMB_EPS * fabs(ss));
      }
      j++;
    }
    // luminosity distance [Mpc/h] and distance modulus
    double ww = rh * ss;               // rh = c/H0 [Mpc/h] (times |Ode|^-1/2 on the fast path)
    double fk = (FLAT || flat) ? ww : f_K_from(ec.OK, ww);
    if (!(fk > 0.0)) e = 1;         // also catches NaN
    // mu_th = 5 log10(fk / (az H_fid)) + 25; the az part is tabulated (ln az)
    const double mu_th = fma(5.0 / M_LN10, (SLOW ? log(fk) : fast_log(fk, T)) - lnaz, SN_MU0);
    const int i0 = __ldg(&L.first[iz]), i1 = __ldg(&L.first[iz + 1]);
    for (int i = i0; i < i1; i++) {
      const double2 *__restrict__ r = reinterpret_cast<const double2 *>(L.sn + (size_t)i * SN_ROW);
      const double2 ms = __ldg(&r[0]), cz = __ldg(&r[1]), w01 = __ldg(&r[2]), w23 = __ldg(&r[3]),
                    w45 = __ldg(&r[4]);
      double mu_obs, sig2;
      if (mode == PMCB200_CHI2_betaz) {
        const double t2 = fma(m_.Theta3, cz.y, m_.t2base);
        mu_obs = ms.x + m_.Theta0 + m_.t1 * (ms.y - m_.stretch) + t2 * (cz.x - m_.color);
        sig2 = w01.x + m_.d1 * m_.d1 * w01.y + t2 * t2 * w23.x
               + 2.0 * (m_.d1 * w23.y + t2 * w45.x + m_.d1 * t2 * w45.y);
      } else {
        // mu_obs = m + Theta0 + t1 (s - stretch) + t2 (c - color); sigma^2 = theta^T W theta + const
        mu_obs = fma(m_.t1, ms.y, fma(m_.t2base, cz.x, ms.x + m_.base0));
        sig2 = fma(m_.k1, w01.y, fma(m_.k2, w23.x, fma(m_.k3, w23.y, fma(m_.k4, w45.x, fma(m_.k5, w45.y, w01.x)))));
      }
      const double res = mu_obs - mu_th;
      chi2 = fma(res * res, fast_rcp(sig2), chi2);
      if (L.sn_add_logdetCov) logdet += log(sig2);
    }
  }
}

// One redshift for ONE lane (the warp-per-sample kernel below): sn_zloop's body with lane-local control flow --
// the rare integral that does not converge at stage 5 continues in its own lane.
template <bool HASQ, bool FLAT, bool SLOW, bool NEG>
__device__ __forceinline__ void sn_zone(const DevLike &L, const SNCoef &ec, const double *__restrict__ T,
                                        const SNPer &m_, double f1, double rh, int iz, double &chi2, double &logdet,
                                        int &e, unsigned &nev) {
  const bool flat = fabs(ec.OK) < FLAT_EPS;
  const int mode = L.sn_chi2mode;
  const double2 *__restrict__ nd = L.nodes + (size_t)iz * SN_NODES;
  const double *__restrict__ na = L.nodes_a + (size_t)iz * SN_NODES;
  SNSums S;
  double ss, dss, lnaz, h;
  sn_romb5<HASQ, FLAT, SLOW, NEG>(ec, T, nd, na, L.nodes4 + (size_t)iz * (4 * 16), f1, S, ss, dss, lnaz, h);
  bool done = !isfinite(ss) || (fabs(dss) <= ROMB_EPS * fabs(ss));
  nev += 17;
  if (!done) {        // rare: rebuild the tableau row of stage 5 and continue (as sn_zloop)
    const double g0 = S.g0, S1 = S.S1, S2 = S.S2, S3 = S.S3, S4 = S.S4;
    const double T1 = h * g0, T2 = 0.5 * h * (g0 + S1), T3 = 0.25 * h * (g0 + S1 + S2),
                 T4 = 0.125 * h * (g0 + S1 + S2 + S3), T5 = 0.0625 * h * (g0 + S1 + S2 + S3 + S4);
    double a1 = fma(T2 - T1, 1.0 / 3.0, T2);
    double b1 = fma(T3 - T2, 1.0 / 3.0, T3), b2 = fma(b1 - a1, 1.0 / 15.0, b1);
    double c1 = fma(T4 - T3, 1.0 / 3.0, T4), c2 = fma(c1 - b1, 1.0 / 15.0, c1);
    double d1 = fma(T5 - T4, 1.0 / 3.0, T5), d2 = fma(d1 - c1, 1.0 / 15.0, d1), d3 = fma(d2 - c2, 1.0 / 63.0, d2);
    (void)b2;
    double st = T5, R0 = T5, R1 = d1, R2 = d2, R3 = d3;
    int j = 5;
    while (!done) {
      if (j >= ROMB_JMAX) { ss = NAN; break; }
      const int it = 1 << (j - 1);
      double s = 0.0;
      if (2 * it <= SN_NODES) {
        for (int i = it; i < 2 * it; i++) {
          const double2 n = __ldg(&nd[i]);
          s = sn_f<HASQ, FLAT, SLOW, NEG>(ec, T, n.x, n.y, (HASQ || !FLAT) ? __ldg(&na[i]) : 0.0, s);
        }
      } else {
        const double del = h / (double)it, az = __ldg(na);
        for (int i = 0; i < it; i++) {
          double a = fma((double)i + 0.5, del, az);
          s = sn_f<HASQ, FLAT, SLOW, NEG>(ec, T, log(a), rsqrt(a), a, s);
        }
      }
      nev += it;
      st = 0.5 * (st + h * s / (double)it);
      double n1 = fma(st - R0, 1.0 / 3.0, st), n2 = fma(n1 - R1, 1.0 / 15.0, n1), n3 = fma(n2 - R2, 1.0 / 63.0, n2);
      dss = (n3 - R3) * (1.0 / 255.0);
      ss = n3 + dss;
      R0 = st; R1 = n1; R2 = n2; R3 = n3;
      done = !isfinite(ss) || (fabs(dss) <= ROMB_EPS * fabs(ss));
      j++;
    }
  }
  double ww = rh * ss;
  double fk = (FLAT || flat) ? ww : f_K_from(ec.OK, ww);
  if (!(fk > 0.0)) e = 1;
  const double mu_th = fma(5.0 / M_LN10, (SLOW ? log(fk) : fast_log(fk, T)) - lnaz, SN_MU0);
  const int i0 = __ldg(&L.first[iz]), i1 = __ldg(&L.first[iz + 1]);
  for (int i = i0; i < i1; i++) {
    const double2 *__restrict__ r = reinterpret_cast<const double2 *>(L.sn + (size_t)i * SN_ROW);
    const double2 ms = __ldg(&r[0]), cz = __ldg(&r[1]), w01 = __ldg(&r[2]), w23 = __ldg(&r[3]), w45 = __ldg(&r[4]);
    double mu_obs, sig2;
    if (mode == PMCB200_CHI2_betaz) {
      const double t2 = fma(m_.Theta3, cz.y, m_.t2base);
      mu_obs = ms.x + m_.Theta0 + m_.t1 * (ms.y - m_.stretch) + t2 * (cz.x - m_.color);
      sig2 = w01.x + m_.d1 * m_.d1 * w01.y + t2 * t2 * w23.x + 2.0 * (m_.d1 * w23.y + t2 * w45.x + m_.d1 * t2 * w45.y);
    } else {
      mu_obs = fma(m_.t1, ms.y, fma(m_.t2base, cz.x, ms.x + m_.base0));
      sig2 = fma(m_.k1, w01.y, fma(m_.k2, w23.x, fma(m_.k3, w23.y, fma(m_.k4, w45.x, fma(m_.k5, w45.y, w01.x)))));
    }
    const double res = mu_obs - mu_th;
    chi2 = fma(res * res, fast_rcp(sig2), chi2);
    if (L.sn_add_logdetCov) logdet += log(sig2);
  }
}

#ifndef SN_MIN_BLOCKS
#define SN_MIN_BLOCKS 2
#endif
#ifndef SN_BLOCK
#define SN_BLOCK 256
#endif
// per-sample constants of the SN likelihood: integrand coefficients in both normalisations, chi^2 terms
__device__ __forceinline__ void sn_setup(const DevLike &L, const Model &m, int force_slow, SNCoef &ec, SNPer &pm,
                                         double &f1, double &f1s) {
  {
    const ECoef g = make_ecoef(m.c, 0);
    ec.Om = g.Om; ec.OK = g.OK; ec.Ode = g.Ode; ec.p = g.p - 1.0; ec.q = g.q; ec.jassal = g.jassal;
    ec.p2 = ec.p * M_LOG2E; ec.q2 = g.q * M_LOG2E;
    const double aO = fabs(g.Ode);
    const double lg = log2(aO);
    ec.sgn = (g.Ode < 0.0) ? 0x80000000u : 0u;
    ec.Oms = g.Om / aO; ec.OKs = g.OK / aO; ec.scale = rsqrt(aO);
    // bound on |s| over a in [a_min, 1]; outside the fast path's range (or Ode = 0,
    // non-finite input) the warp takes the libdevice path
    const double lna_min = -__ldg(&L.nodes[(size_t)(L.sn_nz - 1) * SN_NODES]).x;
    ec.slow = force_slow || !(fabs(ec.p2) * lna_min + fabs(ec.q2) + fabs(lg) < 990.0);
  }
  // integrand at a = 1 (a^-1/2 = 1, 2^0 = 1), in both normalisations
  f1 = rsqrt(ec.Om + ec.OK + ec.Ode);
  f1s = rsqrt(ec.Oms + ec.OKs + (ec.sgn ? -1.0 : 1.0));
  pm.Theta0 = m.Theta2[0]; pm.Theta3 = m.Theta2[3]; pm.t1 = m.Theta2[1]; pm.t2base = m.Theta2[2];
  pm.d1 = pm.t1; pm.d2 = pm.t2base; pm.stretch = m.stretch; pm.color = m.color;
  if (L.sn_chi2mode == PMCB200_CHI2_no_sc) { pm.d1 = 0.0; pm.d2 = 0.0; pm.t1 = 0.0; pm.t2base = 0.0; }
  if (L.sn_chi2mode == PMCB200_CHI2_Theta2_denom_fixed) { pm.d1 = L.Theta2_denom[1]; pm.d2 = L.Theta2_denom[2]; }
  pm.base0 = pm.Theta0 - pm.t1 * pm.stretch - pm.t2base * pm.color;
  pm.k1 = pm.d1 * pm.d1; pm.k2 = pm.d2 * pm.d2; pm.k3 = 2.0 * pm.d1; pm.k4 = 2.0 * pm.d2;
  pm.k5 = 2.0 * pm.d1 * pm.d2;
}

// ---- SN Ia for small batches: ONE SAMPLE PER WARP, lanes across the redshifts ------------------------------
// The thread-per-sample kernel needs N >= 2 x 148 x 256 samples to fill the machine; the reference's own demo
// draws 10^4 per iteration (Demo/MC_Demo/SN/config_pmc), where it occupies 40 of 148 SMs with 8 warps each.
// Here lane l of the warp owns the redshifts l, l + 32, ...: the 16 tabulated nodes, the Romberg combination, the
// distance modulus and the chi^2 terms of the supernovae at that redshift are all lane-local (sn_zone); the
// sample's chi^2 is one butterfly sum.  Same per-redshift arithmetic as k_like_sn; the sum over redshifts is
// associated differently (per lane, then across lanes), an O(1e-16) relative difference.
// one sample, evaluated by the calling warp (flg already checked by the caller)
template <bool HASQ, bool FLAT>
__device__ __forceinline__ void sn_warp_sample(const DevLike &L, const double *__restrict__ T, int64_t n,
                                               const double *__restrict__ X, int d, double *__restrict__ logpi,
                                               int32_t *__restrict__ err, int set, double add_const, DevCount *cnt,
                                               int force_slow) {
  const int lane = threadIdx.x & 31;
  Model m;
  int e = apply_params(L, X + n * d, m);
  const bool cut = !e && L.special == PMCB200_SPECIAL_de_conservative && de_conservative_violated(m.c);
  unsigned nev = 0;
  double chi2 = 0.0, logdet = 0.0;
  if (!e) {
    SNCoef ec;
    SNPer pm;
    double f1, f1s;
    sn_setup(L, m, force_slow, ec, pm, f1, f1s);
    for (int iz = lane; iz < L.sn_nz; iz += 32) {
      if (ec.slow) sn_zone<HASQ, FLAT, true, false>(L, ec, T, pm, f1, R_HUBBLE, iz, chi2, logdet, e, nev);
      else if (ec.sgn != 0u) sn_zone<HASQ, FLAT, false, true>(L, ec, T, pm, f1s, R_HUBBLE * ec.scale, iz, chi2, logdet, e, nev);
      else sn_zone<HASQ, FLAT, false, false>(L, ec, T, pm, f1s, R_HUBBLE * ec.scale, iz, chi2, logdet, e, nev);
    }
    chi2 = warp_sum(chi2);
    if (L.sn_add_logdetCov) logdet = warp_sum(logdet);
    e = __any_sync(0xffffffffu, e);
  }
  double res = -0.5 * chi2;
  if (L.sn_add_logdetCov) res -= 0.5 * logdet;
  if (cut) res = 0.0;                      // sn.c:260-274: SetDl ran, chi2_SN did not
  else if (!isfinite(res)) e = 1;
  if (lane == 0) put_loglike(logpi, err, n, set, add_const, e ? 0.0 : res, e);
  if (cnt) {
    unsigned tot = nev;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) tot += __shfl_xor_sync(0xffffffffu, tot, o);
    if (lane == 0) {
      atomicAdd(&cnt->sn_evals, (unsigned long long)tot);
      atomicAdd(&cnt->sn_zsteps, (unsigned long long)L.sn_nz);
      atomicAdd(&cnt->sn_exact, 1ull);
    }
  }
}
template <bool HASQ, bool FLAT>
__global__ void __launch_bounds__(SN_BLOCK, 2)
k_like_sn_warp(const DevLike L, int64_t N, const double *__restrict__ X, int d,
               const int16_t *__restrict__ flg, double *__restrict__ logpi,
               int32_t *__restrict__ err, int set, double add_const, DevCount *cnt, int force_slow) {
  __shared__ double T[96 + SN_EXP2_N];
  load_fast_tables_sn(T);
  const int64_t n = (int64_t)blockIdx.x * (SN_BLOCK / 32) + (threadIdx.x >> 5);     // this warp's sample
  if (n >= N) return;
  if (flg && !flg[n]) { if ((threadIdx.x & 31) == 0 && set) { logpi[n] = 0.0; if (err) err[n] = 0; } return; }
  sn_warp_sample<HASQ, FLAT>(L, T, n, X, d, logpi, err, set, add_const, cnt, force_slow);
}
// the same for a LIST of samples (the ones k_like_sn_spec could not certify): a fixed grid, warps stride over the list
template <bool HASQ, bool FLAT>
__global__ void __launch_bounds__(SN_BLOCK, 2)
k_like_sn_warp_list(const DevLike L, const double *__restrict__ X, int d, double *__restrict__ logpi,
                    int32_t *__restrict__ err, int set, double add_const, DevCount *cnt, int force_slow,
                    const uint32_t *__restrict__ list, const unsigned *__restrict__ count) {
  __shared__ double T[96 + SN_EXP2_N];
  const unsigned cntv = *count;
  if (cntv == 0u) return;
  load_fast_tables_sn(T);
  const unsigned nw = gridDim.x * (SN_BLOCK / 32);
  for (unsigned i = blockIdx.x * (SN_BLOCK / 32) + (threadIdx.x >> 5); i < cntv; i += nw)
    sn_warp_sample<HASQ, FLAT>(L, T, (int64_t)list[i], X, d, logpi, err, set, add_const, cnt, force_slow);
}

template <bool HASQ, bool FLAT>
__global__ void __launch_bounds__(SN_BLOCK, SN_MIN_BLOCKS)
k_like_sn(const DevLike L, int64_t N, const double *__restrict__ X, int d,
          const int16_t *__restrict__ flg, double *__restrict__ logpi,
          int32_t *__restrict__ err, int set, double add_const, DevCount *cnt, int force_slow) {
  __shared__ double T[96 + SN_EXP2_N];   // [32 exp2 | 32 log reciprocals | 32 log offsets | 2^(j/1024), pre-biased]
  load_fast_tables_sn(T);
  int64_t n = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  bool active = (n < N) && (!flg || flg[n]);
  unsigned nev = 0;
  Model m;
  int e = 0;
  if (active) e = apply_params(L, X + n * d, m);
  const bool cut = active && !e && L.special == PMCB200_SPECIAL_de_conservative && de_conservative_violated(m.c);
  if (!active || e) {   // keep the warp's control flow uniform on a benign model
    m.c = L.model;
#pragma unroll
    for (int i = 0; i < 4; i++) m.Theta2[i] = L.Theta2[i];
    m.stretch = 1.0; m.color = 0.0;
  }
  SNCoef ec;
  SNPer pm;
  double f1, f1s;
  sn_setup(L, m, force_slow, ec, pm, f1, f1s);
  double chi2 = 0.0, logdet = 0.0;
  if (__any_sync(0xffffffffu, ec.slow)) sn_zloop<HASQ, FLAT, true, false>(L, ec, T, pm, f1, R_HUBBLE, chi2, logdet, e, nev);
  else if (__any_sync(0xffffffffu, ec.sgn != 0u))
    sn_zloop<HASQ, FLAT, false, true>(L, ec, T, pm, f1s, R_HUBBLE * ec.scale, chi2, logdet, e, nev);
  else sn_zloop<HASQ, FLAT, false, false>(L, ec, T, pm, f1s, R_HUBBLE * ec.scale, chi2, logdet, e, nev);
  double res = -0.5 * chi2;
  if (L.sn_add_logdetCov) res -= 0.5 * logdet;
  // de_conservative: log L = 0 without evaluating chi2_SN (sn.c:263-274); SetDl has already run (sn.c:260), so a
  // distance error of a violating model is still an error
  if (cut) res = 0.0;
  else if (!isfinite(res)) e = 1;
  if (active) put_loglike(logpi, err, n, set, add_const, e ? 0.0 : res, e);
  else if (n < N && set) { logpi[n] = 0.0; if (err) err[n] = 0; }
  if (cnt) {   // measurement only: one atomic pair per warp
    unsigned tot = active ? nev : 0u, nact = active ? 1u : 0u;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      tot += __shfl_xor_sync(0xffffffffu, tot, o);
      nact += __shfl_xor_sync(0xffffffffu, nact, o);
    }
    if ((threadIdx.x & 31) == 0 && nact) {
      atomicAdd(&cnt->sn_evals, (unsigned long long)tot);
      atomicAdd(&cnt->sn_zsteps, (unsigned long long)nact * L.sn_nz);
    }
  }
}

// ---- BAO / CMB distance priors: one sample per thread, lean integrand (HASQ: w1 != 0 or jassal possible) -----
// cnt->gen_evals / gen_integrals: integrand evaluations and integrals of these two kernels (measurement)
__device__ __forceinline__ void count_gen(DevCount *cnt, unsigned nev, unsigned nint) {
  if (!cnt) return;
  unsigned long long tot = nev, ni = nint;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) { tot += __shfl_xor_sync(0xffffffffu, tot, o); ni += __shfl_xor_sync(0xffffffffu, ni, o); }
  if ((threadIdx.x & 31) == 0 && ni) { atomicAdd(&cnt->gen_evals, tot); atomicAdd(&cnt->gen_integrals, ni); }
}
template <bool HASQ>
__global__ void __launch_bounds__(PMC_BLOCK)
k_like_bao(const DevLike L, int64_t N, const double *__restrict__ X, int d,
           const int16_t *__restrict__ flg, double *__restrict__ logpi,
           int32_t *__restrict__ err, int set, double add_const, DevCount *cnt) {
  __shared__ double T[96];
  __shared__ double2 LT[LOG1K_N];
  __shared__ double ET[SN_EXP2_N];
  load_lean_tables(T, LT, ET);
  const LeanTabs tb{T, LT, ET};
  const int64_t n = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  unsigned nev = 0, nint = 0;
  const bool live = (n < N) && (!flg || flg[n]);
  Model m;
  int e = 0;
  if (live) e = apply_params(L, X + n * d, m);
  const bool cut = live && !e && L.special == PMCB200_SPECIAL_de_conservative && de_conservative_violated(m.c);
  const int nd = L.g_ndim, method = L.bao_method;
  // the integrals are warp-synchronous (romberg_warp): every lane walks the same calls, act says whose count
  bool act = live && !e && !cut;
  if (act && method != PMCB200_BAO_distance_D_V_ratio && !(m.c.Omega_m > 0.0)) { e = 1; act = false; }
  if (act && method == PMCB200_BAO_distance_d_z && !(m.c.Omega_b > 0.0)) { e = 1; act = false; }
  if (!act) m.c = L.model;
  const pmcb200_cosmo_t &c = m.c;
  double model[4] = {0.0, 0.0, 0.0, 0.0};
  if (method == PMCB200_BAO_distance_A) {
    for (int i = 0; i < nd; i++) {
      model[i] = D_V_lean<HASQ>(c, L.g_z[i], act, e, tb, nev) * sqrt(c.Omega_m) / (L.g_z[i] * R_HUBBLE);
      nint += act;
    }
  } else if (method == PMCB200_BAO_distance_d_z) {
    const double rs = r_sound_lean<HASQ>(c, 1.0 / (1.0 + z_drag(c)), act, e, tb, nev);
    nint += act;
    for (int i = 0; i < nd; i++) { model[i] = rs / D_V_lean<HASQ>(c, L.g_z[i], act, e, tb, nev); nint += act; }
  } else {
    for (int i = 0; i < nd; i++) {
      const double num = D_V_lean<HASQ>(c, L.g_z[2 * i], act, e, tb, nev);
      model[i] = num / D_V_lean<HASQ>(c, L.g_z[2 * i + 1], act, e, tb, nev);
      nint += 2 * act;
    }
  }
  double res = 0.0;
  if (act) {
    if (!e) res = gauss_comp_logpdf(L.g_comp, nd, model);
    if (!isfinite(res)) e = 1;
  }
  if (live) put_loglike(logpi, err, n, set, add_const, e ? 0.0 : res, e);      // cut: log L = 0, no error (bao.c:154-176)
  else if (n < N && set) { logpi[n] = 0.0; if (err) err[n] = 0; }
  count_gen(cnt, nev, nint);
}

// ---- comoving distance to a*: the reference's 11-stage Romberg value as a linear functional (round 2, session AC) ---
// likeli_CMBDistPrior (wmap.c:945-1049) needs w(a*) = R_H int_(a*)^1 da / sqrt(a^4 E^2), and NR qromb takes 11 stages = 1025
// equidistant nodes in a for it (the integrand varies over three decades of a): 95 % of k_like_cmbdp's evaluations.  The value
// it returns is NOT the integral -- its truncation error is 2e-5, measured against Gauss-Legendre in ln a
// (tools/proto/cmb_spectral_proto.py) -- so the rule itself has to be reproduced.  But when the stopping test fails at
// stages 5..10 and holds at stage 11, what qromb returns is a FIXED linear functional of the integrand at the nodes
// a_i = a* + (1 - a*) i / 1024: ss_11 = h sum_i omega_i f(a_i) (trapezoid stages 7..11 combined by the Neville weights), and
// so is every stage's error estimate dss_j.  With t = (a - a*) / (1 - a*) in [0, 1] and the FIXED variable v = ln(t + tau)
// (tau = a*_fid / (1 - a*_fid): v is ln a up to 1 % distortion at the lower end) the integrand is smooth in v, and the
// functionals applied to its Chebyshev interpolant on CMB_M fixed points v_k become fixed weights on the node VALUES:
//     ss_j = h sum_k Theta_j[k] f(a* + h t_k),   dss_j likewise          (tables: pmc_init_sn_tables, long double)
// 56 evaluations + 18 x 56 FMA instead of 1025 evaluations; measured on the CPU: |ss_11 / orc_w - 1| <= 2.4e-15.
// Certificates per sample (all from the same 56 values): the interpolant's last three coefficients <= CMB_TAIL_TOL |c_0|; the
// stopping test FAILS at stages 5..10 and HOLDS at stage 11, each with a 1e-4 margin (the reference's |dss_11| / |ss_11| sits at
// 0.77 .. 0.99 of EPS: the test is evaluated, not assumed).  A sample that fails any of them takes the node-by-node path.
#define CMB_M 56
#define CMB_NROW 18          // ss_11, dss_11, ss_5..10, dss_5..10, c_0, c_(M-3), c_(M-2), c_(M-1)
#define CMB_TAIL_TOL 5.0e-12
#ifndef CMB_UNROLL
#define CMB_UNROLL 4       // evaluations per loop trip (C5 per 1e7 samples: 56 = fully unrolled, 5k instructions, 28.1 ms; 8: 27.2; 4: 26.6; 2: 26.6; 1: 27.0)
#endif
#define CMB_STR2(x) #x
#define CMB_PRAGMA_UNROLL(n) _Pragma(CMB_STR2(unroll n))
#define CMB_TAU (1.0 / 1090.0)      // (1/1091) / (1 - 1/1091)
__constant__ double CMB_T[CMB_M];                 // t_k = exp(v_k) - tau
__constant__ double CMB_TH[CMB_NROW * CMB_M];
template <bool HASQ>
__device__ __forceinline__ bool cmb_spec_w(const pmcb200_cosmo_t &c, double as, const LeanTabs &tb, double &w) {
  const ECoefF f = make_ecoef_fast(c, 1);
  GInt q;
  q.g = make_glean(f); q.R3 = 0.0;
  if (!((as > 0.0) && (as < 1.0) && glean_in_range(q.g, as))) return false;
  const double h = 1.0 - as;
  double acc[CMB_NROW];
#pragma unroll
  for (int r = 0; r < CMB_NROW; r++) acc[r] = 0.0;
  CMB_PRAGMA_UNROLL(CMB_UNROLL)
  for (int k = 0; k < CMB_M; k++) {
    const double fv = gint_acc<HASQ, false>(q, tb.LT, tb.ET, fma(h, CMB_T[k], as), 0.0);
#pragma unroll
    for (int r = 0; r < CMB_NROW; r++) acc[r] = fma(CMB_TH[r * CMB_M + k], fv, acc[r]);
  }
  bool ok = (fabs(acc[15]) + fabs(acc[16]) + fabs(acc[17]) <= CMB_TAIL_TOL * fabs(acc[14])) && (acc[0] > 0.0);
  ok = ok && (fabs(acc[1]) <= (1.0 - 1.0e-4) * ROMB_EPS * fabs(acc[0]));
#pragma unroll
  for (int j = 0; j < 6; j++) ok = ok && (fabs(acc[8 + j]) > (1.0 + 1.0e-4) * ROMB_EPS * fabs(acc[2 + j]));
  w = R_HUBBLE * h * acc[0];
  return ok;      // a NaN anywhere fails the comparisons
}

template <bool HASQ>
__global__ void __launch_bounds__(PMC_BLOCK)
k_like_cmbdp(const DevLike L, int64_t N, const double *__restrict__ X, int d,
             const int16_t *__restrict__ flg, double *__restrict__ logpi,
             int32_t *__restrict__ err, int set, double add_const, DevCount *cnt, int spectral) {
  __shared__ double T[96];
  __shared__ double2 LT[LOG1K_N];
  __shared__ double ET[SN_EXP2_N];
  load_lean_tables(T, LT, ET);
  const LeanTabs tb{T, LT, ET};
  const int64_t n = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  unsigned nev = 0, nint = 0;
  const bool live = (n < N) && (!flg || flg[n]);
  Model m;
  int e = 0;
  if (live) e = apply_params(L, X + n * d, m);
  if (live && !e && L.special == PMCB200_SPECIAL_de_conservative && de_conservative_violated(m.c))
    e = 1;               // wmap.c:1041-1044: wmap_de_prior error, the point gets zero weight
  if (live && !e && (!(m.c.Omega_m > 0.0) || !(m.c.Omega_b > 0.0))) e = 1;
  const bool act = live && !e;
  if (!act) m.c = L.model;
  const pmcb200_cosmo_t &c = m.c;
  const double zs = z_star(c), as = 1.0 / (1.0 + zs);
  // distance to a*: spectral form where it is certified; the warp-synchronous Romberg runs for the other lanes only (and
  // returns at once when no lane of the warp needs it)
  double wsp = 0.0;
  const bool cert = spectral && act && cmb_spec_w<HASQ>(c, as, tb, wsp);
  if (cert) nev += CMB_M;
  // (with the spectral form on, the few lanes left share their stages from 128 new nodes on with the whole warp: the warp would
  // otherwise wait for one lane's 1025 or 2049 serial evaluations)
  const double wex = w_lean<HASQ>(c, as, 1, act && !cert, e, tb, nev, spectral ? 128 : ROMB_COOP_MIN);
  const double ww = cert ? wsp : wex;
  const double rs = r_sound_lean<HASQ>(c, as, act, e, tb, nev);
  nint += 2 * act;
  double res = 0.0;
  if (act) {
    double model[4];
    const double fK = f_K(c, ww);
    model[0] = M_PI * fK / rs;
    model[1] = sqrt(c.Omega_m) * fK / R_HUBBLE;
    model[2] = zs;
    model[3] = 100.0 * c.Omega_b * c.h_100 * c.h_100;
    if (!e) res = gauss_comp_logpdf(L.g_comp, L.g_ndim, model);
    if (!isfinite(res)) e = 1;
  }
  if (live) put_loglike(logpi, err, n, set, add_const, e ? 0.0 : res, e);
  else if (n < N && set) { logpi[n] = 0.0; if (err) err[n] = 0; }
  count_gen(cnt, nev, nint);
  if (cnt) {
    const unsigned nc = __popc(__ballot_sync(0xffffffffu, cert));
    if ((threadIdx.x & 31) == 0 && nc) atomicAdd(&cnt->cmb_spec, (unsigned long long)nc);
  }
}

// round-1 versions (general integrand with per-node checks, sequential sums): kept for A/B measurements
// (PMCB200_LIKE_V1=1) and as the cross-check of the lean path in the parity suite
__global__ void __launch_bounds__(PMC_BLOCK)
k_like_bao_v1(const DevLike L, int64_t N, const double *__restrict__ X, int d,
              const int16_t *__restrict__ flg, double *__restrict__ logpi,
              int32_t *__restrict__ err, int set, double add_const) {
  __shared__ double T[96];
  load_fast_tables(T);
  int64_t n = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= N) return;
  if (flg && !flg[n]) { if (set) { logpi[n] = 0.0; if (err) err[n] = 0; } return; }
  Model m;
  int e = apply_params(L, X + n * d, m);
  double res = 0.0;
  if (!e && !(L.special == PMCB200_SPECIAL_de_conservative && de_conservative_violated(m.c))) {
    double model[4];
    const int nd = L.g_ndim;
    const pmcb200_cosmo_t &c = m.c;
    if (L.bao_method == PMCB200_BAO_distance_A) {
      if (!(c.Omega_m > 0.0)) e = 1;
      else for (int i = 0; i < nd; i++)
        model[i] = D_V(c, L.g_z[i], e, T) * sqrt(c.Omega_m) / (L.g_z[i] * R_HUBBLE);
    } else if (L.bao_method == PMCB200_BAO_distance_d_z) {
      if (!(c.Omega_m > 0.0) || !(c.Omega_b > 0.0)) e = 1;
      else {
        double rs = r_sound(c, 1.0 / (1.0 + z_drag(c)), e, T);
        for (int i = 0; i < nd; i++) model[i] = rs / D_V(c, L.g_z[i], e, T);
      }
    } else {
      for (int i = 0; i < nd; i++) model[i] = D_V(c, L.g_z[2 * i], e, T) / D_V(c, L.g_z[2 * i + 1], e, T);
    }
    if (!e) res = gauss_comp_logpdf(L.g_comp, nd, model);
    if (!isfinite(res)) e = 1;
  }
  put_loglike(logpi, err, n, set, add_const, e ? 0.0 : res, e);
}

__global__ void __launch_bounds__(PMC_BLOCK)
k_like_cmbdp_v1(const DevLike L, int64_t N, const double *__restrict__ X, int d,
                const int16_t *__restrict__ flg, double *__restrict__ logpi,
                int32_t *__restrict__ err, int set, double add_const) {
  __shared__ double T[96];
  load_fast_tables(T);
  int64_t n = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= N) return;
  if (flg && !flg[n]) { if (set) { logpi[n] = 0.0; if (err) err[n] = 0; } return; }
  Model m;
  int e = apply_params(L, X + n * d, m);
  double res = 0.0;
  const pmcb200_cosmo_t &c = m.c;
  const bool cut = !e && L.special == PMCB200_SPECIAL_de_conservative && de_conservative_violated(c);
  if (cut) e = 1;
  if (!e && (!(c.Omega_m > 0.0) || !(c.Omega_b > 0.0))) e = 1;
  if (!e) {
    double model[4];
    double zs = z_star(c), as = 1.0 / (1.0 + zs);
    double ww = w_generic(c, as, 1, e, T);
    double fK = f_K(c, ww);
    double rs = r_sound(c, as, e, T);
    model[0] = M_PI * fK / rs;
    model[1] = sqrt(c.Omega_m) * fK / R_HUBBLE;
    model[2] = zs;
    model[3] = 100.0 * c.Omega_b * c.h_100 * c.h_100;
    if (!e) res = gauss_comp_logpdf(L.g_comp, L.g_ndim, model);
    if (!isfinite(res)) e = 1;
  }
  put_loglike(logpi, err, n, set, add_const, e ? 0.0 : res, e);
}

// Parity probe: the mapped model of one data set per sample (the device's apply_params), so that the parameter
// mapping can be compared bit for bit with the reference's own switch + set_base_parameters.
// out[n*16 ..] = Omega_m Omega_de w0 w1 h_100 Omega_b Omega_nu_mass Neff_nu_mass de_param Theta2[4] stretch color 0
__global__ void __launch_bounds__(PMC_BLOCK)
k_map_params(const DevLike L, int64_t N, const double *__restrict__ X, int d, double *__restrict__ out,
             int32_t *__restrict__ err) {
  int64_t n = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= N) return;
  Model m;
  const int e = apply_params(L, X + n * d, m);
  double *o = out + n * 16;
  o[0] = m.c.Omega_m; o[1] = m.c.Omega_de; o[2] = m.c.w0_de; o[3] = m.c.w1_de; o[4] = m.c.h_100;
  o[5] = m.c.Omega_b; o[6] = m.c.Omega_nu_mass; o[7] = m.c.Neff_nu_mass; o[8] = (double)m.c.de_param;
#pragma unroll
  for (int i = 0; i < 4; i++) o[9 + i] = m.Theta2[i];
  o[13] = m.stretch; o[14] = m.color; o[15] = 0.0;
  if (err) err[n] = e;
}

__global__ void __launch_bounds__(PMC_BLOCK)
k_like_banana(const DevLike L, int64_t N, const double *__restrict__ X, int d,
              const int16_t *__restrict__ flg, double *__restrict__ logpi,
              int32_t *__restrict__ err, int set, double add_const) {
  int64_t n = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= N) return;
  if (flg && !flg[n]) { if (set) { logpi[n] = 0.0; if (err) err[n] = 0; } return; }
  const double *x = X + n * d;
  double s1 = L.banana_sigma1sq, b = L.banana_b;
  double x0 = x[0], y2 = x[1] + b * (x0 * x0 - s1);
  double q = x0 * x0 / s1 + y2 * y2;
  for (int j = 2; j < d; j++) q = fma(x[j], x[j], q);
  double res = -0.5 * q - 0.5 * (d * LN2PI + log(s1));
  int e = !isfinite(res);
  put_loglike(logpi, err, n, set, add_const, e ? 0.0 : res, e);
}
