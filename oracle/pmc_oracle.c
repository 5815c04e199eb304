/* ============================================================================
 * pmc_oracle.c -- CPU restatement ("oracle") of CosmoPMC's PMC iteration.
 *
 * TEST INFRASTRUCTURE ONLY (see pmc_oracle.h).  PARITY UNPINNED: pmclib and
 * nicaea are external, un-pinned and absent; every function below cites the
 * reference call site it serves and the published algorithm it restates.
 *
 * Plain C99, IEEE double, no dependencies beyond libm (+ optional OpenMP for
 * the batch loops, which only parallelise over independent samples).
 * ========================================================================== */
#include "pmc_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

/* ---- named conventions that are [UPSTREAM-RECALL] ------ */
#define ORC_R_HUBBLE      2997.92458          /* c/(100 km/s/Mpc) in Mpc/h     */
#define ORC_C_KMS         299792.458
#define ORC_LN2PI         1.8378770664093454836
#define ORC_ROMB_EPS      1.0e-6              /* nicaea w(): qromberg EPS      */
#define ORC_ROMB_JMAX     20
#define ORC_ROMB_K        5
#define ORC_SN_H_FID      0.7                 /* M = Mbar - 5 log10 h70, manual.tex:1321 */
#define ORC_FLAT_EPS      1.0e-8              /* |Omega_K| below this => flat  */
#define ORC_OMEGA_GAMMA_H2 2.469e-5           /* photons, Komatsu et al. 2009  */
#define ORC_NEFF_NU       3.04
#define ORC_PERP_DENOM_ALL 1                  /* perplexity, evidence: divide by
                                                 psim->nsamples (all draws)    */

/* ==========================================================================
 * Linear algebra
 * ========================================================================== */

/* In-place lower Cholesky (row-major), strict upper triangle zeroed.
 * mvdens_cholesky_decomp, call site wrappers/src/param.c:700. */
int orc_cholesky(int d, double *A)
{
   for (int j = 0; j < d; j++) {
      double s = A[j * d + j];
      for (int k = 0; k < j; k++) s -= A[j * d + k] * A[j * d + k];
      if (!(s > 0.0) || !isfinite(s)) return -1;
      double ljj = sqrt(s);
      A[j * d + j] = ljj;
      for (int i = j + 1; i < d; i++) {
         double t = A[i * d + j];
         for (int k = 0; k < j; k++) t -= A[i * d + k] * A[j * d + k];
         A[i * d + j] = t / ljj;
      }
   }
   for (int i = 0; i < d; i++)
      for (int j = i + 1; j < d; j++) A[i * d + j] = 0.0;
   return 0;
}

/* ==========================================================================
 * Densities.  pmclib mvdens_log_pdf / mix_mvdens_log_pdf, passed as
 * mix_mvdens_log_pdf_void at exec/cosmo_pmc.c:343; mvdens_log_pdf used at
 * wrappers/src/param.c:1023.  SURVEY.md 8a row a4.
 * ========================================================================== */
double orc_mvdens_log_pdf(int d, int df, const double *mean, const double *chol,
                          const double *x)
{
   double y[PMCB200_MAX_DIM];
   double m = 0.0, logdet = 0.0;
   /* forward substitution y = L^-1 (x - mean)  (dtrsv, lower, non-unit) */
   for (int i = 0; i < d; i++) {
      double t = x[i] - mean[i];
      for (int k = 0; k < i; k++) t -= chol[i * d + k] * y[k];
      y[i] = t / chol[i * d + i];
      m += y[i] * y[i];
      logdet += log(chol[i * d + i]);
   }
   if (df <= 0) return -0.5 * (m + d * ORC_LN2PI) - logdet;
   /* multivariate Student-t, nu = df (manual.tex:444-450) */
   double nu = (double)df;
   return lgamma(0.5 * (nu + d)) - lgamma(0.5 * nu) - 0.5 * d * log(nu * M_PI)
          - logdet - 0.5 * (nu + d) * log1p(m / nu);
}

/* log sum_k alpha_k exp(log phi_k), skipping alpha_k == 0, NO max-shift
 * (far tails give log 0 = -inf, the sample is then dropped by the weight
 * stage exactly as in the reference).  */
double orc_mix_log_pdf(int K, int d, int df, const double *wght,
                       const double *mean, const double *chol, const double *x)
{
   double s = 0.0;
   for (int k = 0; k < K; k++) {
      if (wght[k] == 0.0) continue;
      s += wght[k] * exp(orc_mvdens_log_pdf(d, df, mean + k * d,
                                            chol + (size_t)k * d * d, x));
   }
   return log(s);
}

void orc_mix_log_pdf_batch(int64_t N, int K, int d, int df, const double *wght,
                           const double *mean, const double *chol,
                           const double *X, double *out)
{
   for (int64_t n = 0; n < N; n++)
      out[n] = orc_mix_log_pdf(K, d, df, wght, mean, chol, X + n * d);
}

/* ==========================================================================
 * Sampler.  pmclib simulate_mix_mvdens / mix_mvdens_ran / mvdens_ran, call
 * site exec/cosmo_pmc.c:320.  SURVEY.md 8a row a3.
 *
 * The reference draws from GSL mt19937 serially on rank 0; the RNG stream is
 * not part of the parity contract (GSL row).  What is: given
 * the uniform u the component index is bit-exact, and given the normals z the
 * point is x = mu_k + L_k z.  The oracle also restates the product's counter
 * based stream (Philox4x32-10, Salmon et al. 2011) so whole shards can be
 * compared.
 * ========================================================================== */
static inline uint32_t mulhilo32(uint32_t a, uint32_t b, uint32_t *hi)
{
   uint64_t p = (uint64_t)a * (uint64_t)b;
   *hi = (uint32_t)(p >> 32);
   return (uint32_t)p;
}

void orc_philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4])
{
   uint32_t c0 = ctr[0], c1 = ctr[1], c2 = ctr[2], c3 = ctr[3];
   uint32_t k0 = key[0], k1 = key[1];
   for (int r = 0; r < 10; r++) {
      uint32_t hi0, hi1;
      uint32_t lo0 = mulhilo32(0xD2511F53u, c0, &hi0);
      uint32_t lo1 = mulhilo32(0xCD9E8D57u, c2, &hi1);
      uint32_t n0 = hi1 ^ c1 ^ k0, n1 = lo1, n2 = hi0 ^ c3 ^ k1, n3 = lo0;
      c0 = n0; c1 = n1; c2 = n2; c3 = n3;
      k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
   }
   out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

/* inverse CDF over the running sum of the weights: first k with u < cw_k;
 * zero-weight (dead) components can never be chosen; fallback = last live
 * component (round-off in the running sum). */
int orc_select_component(int K, const double *wght, double u)
{
   double cw = 0.0;
   int last = 0;
   for (int k = 0; k < K; k++) {
      if (wght[k] == 0.0) continue;
      cw += wght[k];
      last = k;
      if (u < cw) return k;
   }
   return last;
}

static inline double u53(uint32_t hi, uint32_t lo)
{  /* (0,1]: 53 random bits + 1 ulp */
   uint64_t b = (((uint64_t)hi << 32) | lo) >> 11;
   return ((double)b + 1.0) * (1.0 / 9007199254740992.0);
}

/* draws of global sample g: counter = (g_lo, g_hi, call, iter), key = seed.
 * call 0: u = r0 * 2^-32 in [0,1) (component); call 1+p: Box-Muller pair p
 * (z[2p], z[2p+1]); Student-t: chi^2_df from df further normals. */
void orc_sample_draws(uint64_t seed, uint32_t iter, int64_t g, int d, int df,
                      double *u, double *z, double *tscale)
{
   uint32_t key[2] = {(uint32_t)seed, (uint32_t)(seed >> 32)};
   uint32_t ctr[4] = {(uint32_t)(uint64_t)g, (uint32_t)((uint64_t)g >> 32), 0, iter};
   uint32_t r[4];
   orc_philox4x32_10(ctr, key, r);
   *u = (double)r[0] * (1.0 / 4294967296.0);
   int nz = d + (df > 0 ? df : 0);
   double chi2 = 0.0;
   for (int p = 0; 2 * p < nz; p++) {
      ctr[2] = 1 + p;
      orc_philox4x32_10(ctr, key, r);
      double u1 = u53(r[0], r[1]), u2 = u53(r[2], r[3]);
      double rad = sqrt(-2.0 * log(u1));
      double ang = 2.0 * M_PI * u2;
      double zz[2] = {rad * cos(ang), rad * sin(ang)};
      for (int q = 0; q < 2; q++) {
         int i = 2 * p + q;
         if (i < d) z[i] = zz[q];
         else if (i < nz) chi2 += zz[q] * zz[q];
      }
   }
   *tscale = (df > 0) ? sqrt((double)df / chi2) : 1.0;
}

static int in_box(int d, const double *bmin, const double *bmax, const double *x)
{
   if (!bmin || !bmax) return 1;
   for (int j = 0; j < d; j++)
      if (!(x[j] >= bmin[j] && x[j] <= bmax[j])) return 0;
   return 1;
}

static void transform(int d, const double *mean, const double *chol,
                      const double *z, double scale, double *x)
{  /* x = mu + scale * L z   (dtrmv lower) */
   for (int i = 0; i < d; i++) {
      double t = 0.0;
      for (int k = 0; k <= i; k++) t += chol[i * d + k] * z[k];
      x[i] = mean[i] + scale * t;
   }
}

int64_t orc_simulate(int64_t N, uint64_t seed, uint32_t iter, int64_t offset,
                     int K, int d, int df, const double *wght,
                     const double *mean, const double *chol, const double *bmin,
                     const double *bmax, double *X, int32_t *idx, int16_t *flg)
{
   int64_t nok = 0;
   for (int64_t n = 0; n < N; n++) {
      double u, z[PMCB200_MAX_DIM], ts;
      orc_sample_draws(seed, iter, offset + n, d, df, &u, z, &ts);
      int k = orc_select_component(K, wght, u);
      transform(d, mean + k * d, chol + (size_t)k * d * d, z, ts, X + n * d);
      idx[n] = k;
      flg[n] = (int16_t)in_box(d, bmin, bmax, X + n * d);
      nok += flg[n];
   }
   return nok;
}

int64_t orc_simulate_from_draws(int64_t N, const double *u, const double *z,
                     int K, int d, const double *wght, const double *mean,
                     const double *chol, const double *bmin, const double *bmax,
                     double *X, int32_t *idx, int16_t *flg)
{
   int64_t nok = 0;
   for (int64_t n = 0; n < N; n++) {
      int k = orc_select_component(K, wght, u[n]);
      transform(d, mean + k * d, chol + (size_t)k * d * d, z + n * d, 1.0, X + n * d);
      idx[n] = k;
      flg[n] = (int16_t)in_box(d, bmin, bmax, X + n * d);
      nok += flg[n];
   }
   return nok;
}

/* ==========================================================================
 * Romberg quadrature.  pmclib maths.c sm2_qromberg = Numerical Recipes
 * `qromb` (trapzd stages, K=5 point polynomial extrapolation to h=0, stop
 * when |dss| <= eps |ss|).  The truncation error (~1e-6) is far above the
 * parity tolerance, so the node sequence and the stopping rule ARE the
 * algorithm.
 * ========================================================================== */
static void polint0(const double *xa, const double *ya, int n, double *y, double *dy)
{  /* NR polint evaluated at x = 0; arrays 0-based */
   double c[ORC_ROMB_K], dd[ORC_ROMB_K];
   int ns = 0;
   double dif = fabs(xa[0]);
   for (int i = 0; i < n; i++) {
      double dift = fabs(xa[i]);
      if (dift < dif) { ns = i; dif = dift; }
      c[i] = ya[i]; dd[i] = ya[i];
   }
   *y = ya[ns--];
   for (int m = 1; m < n; m++) {
      for (int i = 0; i < n - m; i++) {
         double ho = xa[i], hp = xa[i + m];
         double w = c[i + 1] - dd[i];
         double den = w / (ho - hp);
         dd[i] = hp * den;
         c[i] = ho * den;
      }
      *dy = (2 * (ns + 1) < (n - m)) ? c[ns + 1] : dd[ns--];
      *y += *dy;
   }
}

double orc_qromberg(double (*f)(double, void *), void *p, double a, double b,
                    double eps, int *nstage, int *err)
{
   double s[ORC_ROMB_JMAX + 2], h[ORC_ROMB_JMAX + 2];
   double st = 0.0, ss = 0.0, dss;
   h[0] = 1.0;
   for (int j = 0; j < ORC_ROMB_JMAX; j++) {
      /* trapzd stage j+1 */
      if (j == 0) {
         st = 0.5 * (b - a) * (f(a, p) + f(b, p));
      } else {
         long it = 1L << (j - 1);
         double tnm = (double)it, del = (b - a) / tnm, x = a + 0.5 * del, sum = 0.0;
         for (long i = 0; i < it; i++, x += del) sum += f(x, p);
         st = 0.5 * (st + (b - a) * sum / tnm);
      }
      s[j] = st;
      if (j + 1 >= ORC_ROMB_K) {
         polint0(&h[j + 1 - ORC_ROMB_K], &s[j + 1 - ORC_ROMB_K], ORC_ROMB_K, &ss, &dss);
         if (!isfinite(ss)) { if (err) *err = 1; if (nstage) *nstage = j + 1; return ss; }
         if (fabs(dss) <= eps * fabs(ss)) { if (nstage) *nstage = j + 1; return ss; }
      }
      h[j + 1] = 0.25 * h[j];
   }
   if (err) *err = 1;           /* "too many steps" (pmc_tooManySteps) */
   if (nstage) *nstage = ORC_ROMB_JMAX;
   return ss;
}

/* ==========================================================================
 * Cosmology.  nicaea cosmo.c (Esqr, w, f_K, D_lum), restated from
 * Manual/manual.tex:1290-1325 and par_files/cosmo.par:39-44 (dark-energy
 * parametrisations).  SURVEY.md 8a row a6, App. A.
 * ========================================================================== */
static double f_de(const pmcb200_cosmo_t *c, double a)
{
   if (c->de_param == PMCB200_DE_jassal)   /* w(a) = w0 + w1 a (1-a) */
      return pow(a, -3.0 * (1.0 + c->w0_de)) * exp(1.5 * c->w1_de * (1.0 - a) * (1.0 - a));
   /* linder: w(a) = w0 + w1 (1-a) */
   return pow(a, -3.0 * (1.0 + c->w0_de + c->w1_de)) * exp(-3.0 * c->w1_de * (1.0 - a));
}

static double Omega_r(const pmcb200_cosmo_t *c)
{
   return ORC_OMEGA_GAMMA_H2 * (1.0 + 0.2271 * ORC_NEFF_NU) / (c->h_100 * c->h_100);
}

double orc_Esqr(const pmcb200_cosmo_t *c, double a, int wOmegar)
{
   double OK = 1.0 - c->Omega_m - c->Omega_de - c->Omega_nu_mass;
   double a2 = a * a;
   double EE = (c->Omega_m + c->Omega_nu_mass) / (a2 * a) + OK / a2 + c->Omega_de * f_de(c, a);
   if (wOmegar) EE += Omega_r(c) / (a2 * a2);
   return EE;
}

typedef struct { const pmcb200_cosmo_t *c; int wOmegar; int bad; } wint_t;

static double int_for_w(double a, void *p)
{
   wint_t *q = (wint_t *)p;
   double a2 = a * a;
   double dd = a2 * a2 * orc_Esqr(q->c, a, q->wOmegar);
   if (!(dd > 0.0)) { q->bad = 1; return NAN; }
   return 1.0 / sqrt(dd);
}

/* comoving distance [Mpc/h] to scale factor a */
double orc_w(const pmcb200_cosmo_t *c, double a, int wOmegar, int *nstage, int *err)
{
   wint_t q = {c, wOmegar, 0};
   int e = 0;
   double r = orc_qromberg(int_for_w, &q, a, 1.0, ORC_ROMB_EPS, nstage, &e);
   if ((e || q.bad) && err) *err = 1;
   return ORC_R_HUBBLE * r;
}

double orc_f_K(const pmcb200_cosmo_t *c, double w, int wOmegar)
{
   (void)wOmegar;
   double OK = 1.0 - c->Omega_m - c->Omega_de - c->Omega_nu_mass;
   if (fabs(OK) < ORC_FLAT_EPS) return w;
   double sk = sqrt(fabs(OK)) / ORC_R_HUBBLE;
   return OK > 0.0 ? sinh(sk * w) / sk : sin(sk * w) / sk;
}

double orc_D_lum(const pmcb200_cosmo_t *c, double a, int *err)
{
   double ww = orc_w(c, a, 0, NULL, err);
   return orc_f_K(c, ww, 0) / a;
}

/* ---- parameter mapping: the `switch (like->par[i])` blocks of
 * wrappers/src/sn.c:167-224, bao.c:100-147, wmap.c:966-1019 followed by
 * set_base_parameters (wrappers/src/param.c:1544-1661). ---------------------- */
typedef struct {
   pmcb200_cosmo_t c;
   double Theta2[4], stretch, color;
} model_t;

static int apply_params(const pmcb200_like_t *L, const double *x, model_t *m)
{
   double Omegam = -1, Omegab = -1, Omegac = -1, Omegade = -1, Omeganumass = -1;
   double omegam = -1, omegab = -1, omegac = -1, omegade = -1, omeganumass = -1, h100 = -1;
   double OmegaK = 0, omegaK = 0;
   int iOmegade = 0, iOmegaK = 0, iomegade = 0, iomegaK = 0;

   m->c = L->model;
   for (int i = 0; i < 4; i++) m->Theta2[i] = L->sn_Theta2[i];
   m->stretch = 1.0; m->color = 0.0;

   for (int i = 0; i < L->npar; i++) {
      double v = x[i];
      switch (L->par[i]) {
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
         case PMCB200_P_omeganumass:
            if (L->kind != PMCB200_LIKE_SNIa) omeganumass = v;   /* sn.c has no such case */
            break;
         case PMCB200_P_omegac: omegac = v; break;
         case PMCB200_P_omegaK: omegaK = v; iomegaK = 1; break;
         case PMCB200_P_w0de: m->c.w0_de = v; break;
         case PMCB200_P_w1de: m->c.w1_de = v; break;
         case PMCB200_P_h100: h100 = v; break;
         case PMCB200_P_Neffnumass:
            if (L->kind != PMCB200_LIKE_SNIa) m->c.Neff_nu_mass = v;
            break;
         default: break;
      }
      if (L->kind == PMCB200_LIKE_SNIa) {
         switch (L->par[i]) {
            case PMCB200_P_M: m->Theta2[0] = v; break;
            case PMCB200_P_alpha: m->Theta2[1] = v; break;
            case PMCB200_P_beta: m->Theta2[2] = v; break;
            case PMCB200_P_logbeta: m->Theta2[2] = -exp(v); break;
            case PMCB200_P_beta_z: m->Theta2[3] = v; break;
            case PMCB200_P_stretch: m->stretch = v; break;
            case PMCB200_P_color: m->color = v; break;
            default: break;
         }
      }
      if (L->kind == PMCB200_LIKE_CMBDistPrior && !isfinite(v)) return 1;  /* wmap.c:965 */
   }
   if (h100 < 0) h100 = m->c.h_100; else m->c.h_100 = h100;

   /* set_base_parameters, param.c:1544-1593 */
   if (Omegam > 0 || Omegab > 0 || iOmegade == 1 || Omeganumass > 0 || Omegac > 0 || iOmegaK == 1) {
      if (omegam > 0 || omegab > 0 || iomegade == 1 || omeganumass > 0 || omegac > 0 || iomegaK == 1)
         return 1;                         /* mixing physical / non-physical */
   } else {
      if (h100 < 0) return 1;
      /* param.c:1567-1572 divides twice (x/h100/h100) */
      Omegam = omegam / h100 / h100; Omegab = omegab / h100 / h100; Omegade = omegade / h100 / h100;
      Omeganumass = omeganumass / h100 / h100; Omegac = omegac / h100 / h100; OmegaK = omegaK / h100 / h100;
      iOmegade = iomegade; iOmegaK = iomegaK;
   }
   if (Omegam > 0 && iOmegade == 1 && iOmegaK == 1) return 1;  /* overdetermined */
   if (Omegam > 0 && Omegab > 0 && Omegac > 0) return 1;
   if (Omeganumass < 0) Omeganumass = 0;
   /* set_base_Omegam, param.c:1595-1615 */
   if (Omegam > 0) m->c.Omega_m = Omegam;
   else if (Omegab > 0 && Omegac > 0) m->c.Omega_m = Omegab + Omegac;
   else if (Omegade > 0 && iOmegaK == 1) m->c.Omega_m = 1.0 - Omegade - OmegaK - Omeganumass;
   /* set_base_Omegab, param.c:1617-1632 */
   if (Omegab > 0) m->c.Omega_b = Omegab;
   else if (Omegam > 0 && Omegac > 0) m->c.Omega_b = Omegam - Omegac;
   /* set_base_Omegade, param.c:1634-1657 */
   if (Omegade > 0) m->c.Omega_de = Omegade;
   else if (Omegam > 0) m->c.Omega_de = 1.0 - Omegam - OmegaK - Omeganumass;
   else if (Omegab > 0 && Omegac > 0) m->c.Omega_de = 1.0 - Omegab - Omegac - OmegaK - Omeganumass;
   /* set_base_Omeganumass, param.c:1659-1662 */
   if (Omeganumass > 0) m->c.Omega_nu_mass = Omeganumass;
   return 0;
}

/* ---- SN Ia: nicaea SetDl + chi2_SN (call sites sn.c:260,270); formula
 * Manual/manual.tex:1290-1325.  Returns log L = -chi2/2 [- sum log sigma^2 /2]. */
static double loglike_sn(const pmcb200_like_t *L, const model_t *m, int *err,
                         double *mean_stage)
{
   const double pv_fac = 5.0 / M_LN10 * L->sn_v_pec / ORC_C_KMS;
   double chi2 = 0.0, logdet = 0.0, stages = 0.0;
   for (int i = 0; i < L->sn_n; i++) {
      double z = L->sn_z[i], a = 1.0 / (1.0 + z);
      int ns = 0, e = 0;
      double ww = orc_w(&m->c, a, 0, &ns, &e);
      if (e) { *err = 1; return 0.0; }
      stages += ns;
      double dl = orc_f_K(&m->c, ww, 0) / a;
      if (!(dl > 0.0)) { *err = 1; return 0.0; }
      double mu_th = 5.0 * log10(dl / ORC_SN_H_FID) + 25.0;
      const double *W = L->sn_cov + 6 * i;   /* Vmm Vss Vcc Cms Cmc Csc */
      double t1 = m->Theta2[1], t2 = m->Theta2[2];
      if (L->sn_chi2mode == PMCB200_CHI2_betaz) t2 += m->Theta2[3] * z;
      double mu_obs, d1, d2;
      if (L->sn_chi2mode == PMCB200_CHI2_no_sc) {
         mu_obs = L->sn_m[i] + m->Theta2[0];
         d1 = d2 = 0.0;
      } else {
         mu_obs = L->sn_m[i] + m->Theta2[0] + t1 * (L->sn_s[i] - m->stretch)
                  + t2 * (L->sn_c[i] - m->color);
         d1 = t1; d2 = t2;
         if (L->sn_chi2mode == PMCB200_CHI2_Theta2_denom_fixed) {
            d1 = L->sn_Theta2_denom[1]; d2 = L->sn_Theta2_denom[2];
         }
      }
      double spv = pv_fac / z;
      double sig2 = W[0] + d1 * d1 * W[1] + d2 * d2 * W[2]
                    + 2.0 * (d1 * W[3] + d2 * W[4] + d1 * d2 * W[5])
                    + spv * spv + L->sn_sig_int * L->sn_sig_int;
      double r = mu_obs - mu_th;
      chi2 += r * r / sig2;
      logdet += log(sig2);
   }
   if (mean_stage) *mean_stage = stages / L->sn_n;
   double res = -0.5 * chi2;
   if (L->sn_add_logdetCov) res -= 0.5 * logdet;
   return res;
}

/* ---- BAO / CMB helpers: nicaea cmb_bao.c (call sites bao.c:163-171,
 * wmap.c:1034); formulas Manual/manual.tex:1707-1735, Eisenstein & Hu 1998
 * (z_drag), Hu & Sugiyama 1996 (z_star), Komatsu et al. 2009 (l_A, R). ------- */
typedef struct { const pmcb200_cosmo_t *c; int bad; } rsint_t;

static double int_for_r_sound(double a, void *p)
{
   rsint_t *q = (rsint_t *)p;
   const pmcb200_cosmo_t *c = q->c;
   double a2 = a * a;
   /* a^4 E^2 written so that a = 0 is regular */
   double OK = 1.0 - c->Omega_m - c->Omega_de - c->Omega_nu_mass;
   double a4E2 = (c->Omega_m + c->Omega_nu_mass) * a + OK * a2 + Omega_r(c);
   if (a > 0.0) a4E2 += c->Omega_de * a2 * a2 * f_de(c, a);
   double R = 0.75 * c->Omega_b * c->h_100 * c->h_100 / ORC_OMEGA_GAMMA_H2 * a;
   double dd = a4E2 * 3.0 * (1.0 + R);
   if (!(dd > 0.0)) { q->bad = 1; return NAN; }
   return 1.0 / sqrt(dd);
}

double orc_r_sound(const pmcb200_cosmo_t *c, double a, int *nstage, int *err)
{  /* comoving sound horizon [Mpc/h] at scale factor a */
   rsint_t q = {c, 0};
   int e = 0;
   double r = orc_qromberg(int_for_r_sound, &q, 0.0, a, ORC_ROMB_EPS, nstage, &e);
   if (e || q.bad) *err = 1;
   return ORC_R_HUBBLE * r;
}
static double r_sound(const pmcb200_cosmo_t *c, double a, int *err) { return orc_r_sound(c, a, NULL, err); }
double orc_z_star(const pmcb200_cosmo_t *c);
double orc_z_drag(const pmcb200_cosmo_t *c);

static double z_drag(const pmcb200_cosmo_t *c)
{
   double omm = c->Omega_m * c->h_100 * c->h_100, omb = c->Omega_b * c->h_100 * c->h_100;
   double b1 = 0.313 * pow(omm, -0.419) * (1.0 + 0.607 * pow(omm, 0.674));
   double b2 = 0.238 * pow(omm, 0.223);
   return 1291.0 * pow(omm, 0.251) / (1.0 + 0.659 * pow(omm, 0.828)) * (1.0 + b1 * pow(omb, b2));
}

static double z_star(const pmcb200_cosmo_t *c)
{
   double omm = c->Omega_m * c->h_100 * c->h_100, omb = c->Omega_b * c->h_100 * c->h_100;
   double g1 = 0.0783 * pow(omb, -0.238) / (1.0 + 39.5 * pow(omb, 0.763));
   double g2 = 0.560 / (1.0 + 21.1 * pow(omb, 1.81));
   return 1048.0 * (1.0 + 0.00124 * pow(omb, -0.738)) * (1.0 + g1 * pow(omm, g2));
}

double orc_z_star(const pmcb200_cosmo_t *c) { return z_star(c); }
double orc_z_drag(const pmcb200_cosmo_t *c) { return z_drag(c); }

static double D_V(const pmcb200_cosmo_t *c, double z, int *err)
{  /* [f_K^2(w) c z / H(z)]^(1/3), Mpc/h */
   double a = 1.0 / (1.0 + z);
   double ww = orc_w(c, a, 0, NULL, err);
   double fK = orc_f_K(c, ww, 0);
   double EE = orc_Esqr(c, a, 0);
   if (!(EE > 0.0)) { *err = 1; return NAN; }
   return cbrt(fK * fK * ORC_R_HUBBLE * z / sqrt(EE));
}

static double gauss_data_log_pdf(const pmcb200_like_t *L, const double *model)
{
   return orc_mvdens_log_pdf(L->g_ndim, -1, L->g_mean, L->g_chol, model);
}

static double loglike_bao(const pmcb200_like_t *L, const model_t *m, int *err)
{
   double model[PMCB200_MAX_DIM];
   const pmcb200_cosmo_t *c = &m->c;
   int n = L->g_ndim;
   switch (L->bao_method) {
      case PMCB200_BAO_distance_A:
         if (!(c->Omega_m > 0.0)) { *err = 1; return 0.0; }
         for (int i = 0; i < n; i++)
            model[i] = D_V(c, L->g_z[i], err) * sqrt(c->Omega_m) / (L->g_z[i] * ORC_R_HUBBLE);
         break;
      case PMCB200_BAO_distance_d_z: {
         if (!(c->Omega_m > 0.0) || !(c->Omega_b > 0.0)) { *err = 1; return 0.0; }
         double rs = r_sound(c, 1.0 / (1.0 + z_drag(c)), err);
         for (int i = 0; i < n; i++) model[i] = rs / D_V(c, L->g_z[i], err);
         break;
      }
      case PMCB200_BAO_distance_D_V_ratio:
         for (int i = 0; i < n; i++)
            model[i] = D_V(c, L->g_z[2 * i], err) / D_V(c, L->g_z[2 * i + 1], err);
         break;
      default: *err = 1; return 0.0;
   }
   if (*err) return 0.0;
   return gauss_data_log_pdf(L, model);
}

static double loglike_cmbdp(const pmcb200_like_t *L, const model_t *m, int *err)
{  /* model vector (l_A, R, z_star [, 100 omega_b]) */
   const pmcb200_cosmo_t *c = &m->c;
   if (!(c->Omega_m > 0.0) || !(c->Omega_b > 0.0)) { *err = 1; return 0.0; }
   double model[4];
   double zs = z_star(c), as = 1.0 / (1.0 + zs);
   double ww = orc_w(c, as, 1, NULL, err);
   double fK = orc_f_K(c, ww, 1);
   double rs = r_sound(c, as, err);
   if (*err) return 0.0;
   model[0] = M_PI * fK / rs;
   model[1] = sqrt(c->Omega_m) * fK / ORC_R_HUBBLE;
   model[2] = zs;
   model[3] = 100.0 * c->Omega_b * c->h_100 * c->h_100;
   return gauss_data_log_pdf(L, model);
}

/* banana: Wraith et al. 2009, twisted Gaussian */
static double loglike_banana(const pmcb200_like_t *L, const double *x)
{
   int d = L->npar;
   double s1 = L->banana_sigma1sq, b = L->banana_b;
   double y2 = x[1] + b * (x[0] * x[0] - s1);
   double q = x[0] * x[0] / s1 + y2 * y2;
   for (int j = 2; j < d; j++) q += x[j] * x[j];
   return -0.5 * q - 0.5 * (d * ORC_LN2PI + log(s1));
}

/* nicaea test_range_de_conservative (call sites sn.c:265, bao.c:156, wmap.c:1029) [UPSTREAM-RECALL]:
 * 1 if w(a) leaves [-1, -1/3] at a = 1 or at a = ORC_A_ACC (the a_acc of param.c:1091). */
#define ORC_A_ACC (2.0 / 3.0)
static int de_conservative_violated(const pmcb200_cosmo_t *c)
{
   double w_now = c->w0_de, w_acc = w_now;
   if (c->de_param == PMCB200_DE_linder) w_acc = c->w0_de + c->w1_de * (1.0 - ORC_A_ACC);
   else if (c->de_param == PMCB200_DE_jassal) w_acc = c->w0_de + c->w1_de * ORC_A_ACC * (1.0 - ORC_A_ACC);
   return (w_now < -1.0 || w_now > -1.0 / 3.0 || w_acc < -1.0 || w_acc > -1.0 / 3.0);
}

double orc_loglike(const pmcb200_like_t *L, const double *x, int *err)
{
   model_t m;
   /* hard cut of the de_conservative prior inside each probe: SN and BAO return log L = 0
    * (not -inf) for a violating model, sn.c:263-274, bao.c:154-176; likeli_CMBDistPrior computes 0
    * and then raises wmap_de_prior (wmap.c:1027-1044), so the point is dropped */
   if (L->special == PMCB200_SPECIAL_de_conservative &&
       (L->kind == PMCB200_LIKE_SNIa || L->kind == PMCB200_LIKE_BAO || L->kind == PMCB200_LIKE_CMBDistPrior)) {
      if (apply_params(L, x, &m)) { *err = 1; return 0.0; }
      if (de_conservative_violated(&m.c)) {
         if (L->kind == PMCB200_LIKE_CMBDistPrior) *err = 1;
         /* sn.c:260 runs SetDl before the test of sn.c:263-274: a distance error still counts */
         if (L->kind == PMCB200_LIKE_SNIa) loglike_sn(L, &m, err, NULL);
         return 0.0;
      }
   }
   switch (L->kind) {
      case PMCB200_LIKE_Mvdens:
         return orc_mvdens_log_pdf(L->mix_ndim, L->mix_df, L->mix_mean, L->mix_chol, x);
      case PMCB200_LIKE_MixMvdens:
         return orc_mix_log_pdf(L->mix_ncomp, L->mix_ndim, L->mix_df, L->mix_wght,
                                L->mix_mean, L->mix_chol, x);
      case PMCB200_LIKE_BANANA:
         return loglike_banana(L, x);
      case PMCB200_LIKE_SNIa:
         if (apply_params(L, x, &m)) { *err = 1; return 0.0; }
         return loglike_sn(L, &m, err, NULL);
      case PMCB200_LIKE_BAO:
         if (apply_params(L, x, &m)) { *err = 1; return 0.0; }
         return loglike_bao(L, &m, err);
      case PMCB200_LIKE_CMBDistPrior:
         if (apply_params(L, x, &m)) { *err = 1; return 0.0; }
         return loglike_cmbdp(L, &m, err);
      default: *err = 1; return 0.0;
   }
}

/* the mapped model of one probe, for the bit-level comparison with the reference's own
 * switch + set_base_parameters (oracle/_ref/libref_param.so):
 * out = Omega_m Omega_de w0 w1 h_100 Omega_b Omega_nu_mass Neff_nu_mass de_param Theta2[4] stretch color */
int orc_map_params(const pmcb200_like_t *L, const double *x, double out[16])
{
   model_t m;
   int e = apply_params(L, x, &m);
   out[0] = m.c.Omega_m; out[1] = m.c.Omega_de; out[2] = m.c.w0_de; out[3] = m.c.w1_de; out[4] = m.c.h_100;
   out[5] = m.c.Omega_b; out[6] = m.c.Omega_nu_mass; out[7] = m.c.Neff_nu_mass; out[8] = (double)m.c.de_param;
   for (int i = 0; i < 4; i++) out[9 + i] = m.Theta2[i];
   out[13] = m.stretch; out[14] = m.color; out[15] = 0.0;
   return e;
}

double orc_sn_mean_stages(const pmcb200_like_t *L, const double *x)
{
   model_t m; int err = 0; double ms = 0.0;
   if (apply_params(L, x, &m)) return -1.0;
   loglike_sn(L, &m, &err, &ms);
   return err ? -1.0 : ms;
}

/* posterior_log_pdf_common, wrappers/src/param.c:958-1041:
 * sum_i log L_i + logpr_default (param.c:124-129) + special prior term
 * (param.c:1055-1101) + optional Gaussian prior (param.c:1009-1026). */
double orc_posterior_log_pdf(const pmcb200_target_t *t, const double *x, int *err)
{
   double logpost = 0.0;
   for (int i = 0; i < t->ndata; i++) {
      int e = 0;
      logpost += orc_loglike(&t->like[i], x, &e);
      if (e) { *err = 1; return 0.0; }
   }
   double logpr = 0.0;
   for (int j = 0; j < t->npar; j++) logpr -= log(t->max[j] - t->min[j]);
   int special = t->like[0].special;
   if (special == PMCB200_SPECIAL_unity)
      for (int j = 0; j < t->npar; j++) logpr += log(t->max[j] - t->min[j]);
   else if (special == PMCB200_SPECIAL_de_conservative) {   /* param.c:1072-1094 */
      int iw0 = -1, iw1 = -1;
      for (int j = 0; j < t->npar; j++) {
         if (t->like[0].par[j] == PMCB200_P_w0de) iw0 = j;
         if (t->like[0].par[j] == PMCB200_P_w1de) iw1 = j;
      }
      if (iw0 >= 0) {
         if (t->min[iw0] > -1.0 || t->max[iw0] < -1.0 / 3.0) { *err = 1; return 0.0; }
         if (iw1 < 0) logpr += log(t->max[iw0] - t->min[iw0]) - log(2.0 / 3.0);
         else logpr += log(t->max[iw0] - t->min[iw0]) + log(t->max[iw1] - t->min[iw1])
                       - log(0.5 * 2.0 / 3.0 * 2.0 / 3.0 / (1.0 - ORC_A_ACC)) - log(0.5 * 2.0 / 3.0 * 2.0 / 3.0);
      }
   }
   else if (special != PMCB200_SPECIAL_none) { *err = 1; return 0.0; }
   logpost += logpr;
   if (t->prior_mean) {
      double xp[PMCB200_MAX_DIM];
      const double *xx = x;
      if (t->nprior > 0) {
         int j = 0;
         for (int i = 0; i < t->npar; i++) if (t->indprior[i] == 1) xp[j++] = x[i];
         xx = xp;
      }
      logpost += orc_mvdens_log_pdf(t->prior_ndim, -1, t->prior_mean, t->prior_chol, xx);
   }
   return logpost;
}

void orc_posterior_log_pdf_batch(const pmcb200_target_t *t, int64_t N,
                                 const double *X, double *out, int32_t *err,
                                 int nthreads)
{
   (void)nthreads;
#ifdef _OPENMP
   if (nthreads > 0) omp_set_num_threads(nthreads);
#pragma omp parallel for schedule(dynamic, 16)
#endif
   for (int64_t n = 0; n < N; n++) {
      int e = 0;
      out[n] = orc_posterior_log_pdf(t, X + n * t->npar, &e);
      if (err) err[n] = e;
   }
}

/* ==========================================================================
 * Importance weights.  pmclib generic_get_importance_weight_and_deduced_verb
 * (call site cosmo_pmc.c:343-345), normalize_importance_weight (:378),
 * perplexity_and_ess (:46), evidence (:62), effective_number_of_components
 * (:84).  SURVEY.md 8a rows a10, a11, a13.
 * ========================================================================== */
int64_t orc_importance_weights(const pmcb200_target_t *t, int64_t N,
                     const double *X, int K, int d, int df, const double *wght,
                     const double *mean, const double *chol, double beta,
                     int16_t *flg, double *logw, double *maxW, int nthreads)
{
   (void)nthreads;
#ifdef _OPENMP
   if (nthreads > 0) omp_set_num_threads(nthreads);
#pragma omp parallel for schedule(dynamic, 16)
#endif
   for (int64_t n = 0; n < N; n++) {
      if (!flg[n]) { logw[n] = 0.0; continue; }
      const double *x = X + n * d;
      double lq = orc_mix_log_pdf(K, d, df, wght, mean, chol, x);
      int e = 0;
      double lp = orc_posterior_log_pdf(t, x, &e);
      double lw = beta * lp - lq;
      if (e || !isfinite(lq) || !isfinite(lw)) { flg[n] = 0; logw[n] = 0.0; continue; }
      logw[n] = lw;
   }
   int64_t nok = 0;
   double mx = -INFINITY;
   for (int64_t n = 0; n < N; n++)
      if (flg[n]) { nok++; if (logw[n] > mx) mx = logw[n]; }
   *maxW = mx;
   return nok;
}

/* in place log w -> wbar; returns sum exp(log w - maxW); logSum = log sum + maxW */
double orc_normalize_weights(int64_t N, const int16_t *flg, double *w,
                             double maxW, double *logSum)
{
   double sum = 0.0;
   for (int64_t n = 0; n < N; n++) {
      if (flg[n]) { w[n] = exp(w[n] - maxW); sum += w[n]; }
      else w[n] = 0.0;
   }
   for (int64_t n = 0; n < N; n++) w[n] /= sum;
   if (logSum) *logSum = log(sum) + maxW;
   return sum;
}

/* perplexity = exp(-sum wbar log wbar)/N (Wraith et al. 2009 eq. 18,
 * manual.tex:559), ESS = 1/sum wbar^2 */
double orc_perplexity_and_ess(int64_t N, const int16_t *flg, const double *wbar,
                              double *ess)
{
   double H = 0.0, s2 = 0.0;
   int64_t nok = 0;
   for (int64_t n = 0; n < N; n++) {
      if (!flg[n]) continue;
      nok++;
      if (wbar[n] > 0.0) { H -= wbar[n] * log(wbar[n]); s2 += wbar[n] * wbar[n]; }
   }
   if (ess) *ess = 1.0 / s2;
   return exp(H) / (double)(ORC_PERP_DENOM_ALL ? N : nok);
}

/* ENC = 1/sum alpha^2 (manual.tex:599-603; bin/neff_proposal.pl:30-34) */
double orc_enc(int K, const double *wght)
{
   double s = 0.0;
   for (int k = 0; k < K; k++) s += wght[k] * wght[k];
   return 1.0 / s;
}

/* ==========================================================================
 * Rao-Blackwellised EM update.  pmclib update_prop_rb (call site
 * cosmo_pmc.c:247); Cappe et al. 2008 (arXiv:0710.4242) sect. 3-4; Wraith et
 * al. 2009 eqs. 12-14; dead components Manual/manual.tex:482-490.
 * SURVEY.md 8a row a12.  wbar = normalised weights.  Updates wght/mean/chol
 * in place (chol = Cholesky of the new covariance); cov_out (may be NULL)
 * receives the new covariances.  Returns the number of components killed.
 * ========================================================================== */
static int update_prop_rb_impl(int64_t N, const double *X, const int32_t *idx,
                       const int16_t *flg, const double *wbar, int K, int d,
                       int df, double *wght, double *mean, double *chol,
                       double *cov_out, int nthreads);
int orc_update_prop_rb(int64_t N, const double *X, const int32_t *idx,
                       const int16_t *flg, const double *wbar, int K, int d,
                       int df, double *wght, double *mean, double *chol,
                       double *cov_out)
{
   return update_prop_rb_impl(N, X, idx, flg, wbar, K, d, df, wght, mean, chol, cov_out, 1);
}

/* nthreads == 1: the serial loops of the reference's rank 0 (what the parity tests check).  nthreads > 1
 * ("mode B" of the CPU baseline): contiguous chunks of samples per thread with thread-private sums, combined in
 * thread order -- same sums up to the association of the additions. */
static int update_prop_rb_impl(int64_t N, const double *X, const int32_t *idx,
                       const int16_t *flg, const double *wbar, int K, int d,
                       int df, double *wght, double *mean, double *chol,
                       double *cov_out, int nthreads)
{
   const size_t dd = (size_t)d * d;
   int nt = nthreads > 1 ? nthreads : 1;
   double *A = calloc((size_t)nt * K, sizeof(double)), *G = calloc((size_t)nt * K, sizeof(double));
   double *B = calloc((size_t)nt * K * d, sizeof(double));
   double *C = calloc((size_t)nt * K * dd, sizeof(double));
   double *rho = malloc(((size_t)N * K) * sizeof(double));   /* rho*gamma cached */
   int64_t *count = calloc((size_t)nt * K, sizeof(int64_t));
   int64_t Nall = N;
   int ndead = 0;

   /* E-step + first moments */
#ifdef _OPENMP
#pragma omp parallel num_threads(nt)
#endif
   {
#ifdef _OPENMP
   const int th = omp_get_thread_num();
#else
   const int th = 0;
#endif
   const int64_t n0 = N * th / nt, n1 = N * (th + 1) / nt;
   double *At = A + (size_t)th * K, *Gt = G + (size_t)th * K, *Bt = B + (size_t)th * K * d;
   int64_t *ct = count + (size_t)th * K;
   for (int64_t n = n0; n < n1; n++) {
      if (!flg[n]) continue;
      const double *x = X + n * d;
      double r[PMCB200_MAX_COMP], gam[PMCB200_MAX_COMP], rt = 0.0;
      for (int k = 0; k < K; k++) {
         r[k] = 0.0; gam[k] = 1.0;
         if (wght[k] == 0.0) continue;
         r[k] = wght[k] * exp(orc_mvdens_log_pdf(d, df, mean + k * d, chol + k * dd, x));
         rt += r[k];
         if (df > 0) {
            double y[PMCB200_MAX_DIM], m = 0.0;
            for (int i = 0; i < d; i++) {
               double t = x[i] - mean[k * d + i];
               for (int j = 0; j < i; j++) t -= chol[k * dd + i * d + j] * y[j];
               y[i] = t / chol[k * dd + i * d + i];
               m += y[i] * y[i];
            }
            gam[k] = (df + d) / (df + m);
         }
      }
      ct[idx[n]]++;
      for (int k = 0; k < K; k++) {
         double rk = r[k] / rt;
         rho[n * K + k] = rk * gam[k];
         double wr = wbar[n] * rk;
         At[k] += wr;
         Gt[k] += wr * gam[k];
         for (int i = 0; i < d; i++) Bt[k * d + i] += wr * gam[k] * x[i];
      }
   }
   }
   for (int th = 1; th < nt; th++)
      for (int k = 0; k < K; k++) {
         A[k] += A[(size_t)th * K + k]; G[k] += G[(size_t)th * K + k]; count[k] += count[(size_t)th * K + k];
         for (int i = 0; i < d; i++) B[k * d + i] += B[((size_t)th * K + k) * d + i];
      }
   /* M-step: alpha' = A, mu' = B/G, Sigma' = sum w rho gamma (x-mu')(x-mu')^T / A */
   for (int k = 0; k < K; k++) {
      if (wght[k] == 0.0 || !(A[k] > 0.0)) continue;
      for (int i = 0; i < d; i++) B[k * d + i] /= G[k];
   }
#ifdef _OPENMP
#pragma omp parallel num_threads(nt)
#endif
   {
#ifdef _OPENMP
   const int th = omp_get_thread_num();
#else
   const int th = 0;
#endif
   const int64_t n0 = N * th / nt, n1 = N * (th + 1) / nt;
   double *Ct = C + (size_t)th * K * dd;
   for (int64_t n = n0; n < n1; n++) {
      if (!flg[n]) continue;
      const double *x = X + n * d;
      for (int k = 0; k < K; k++) {
         if (wght[k] == 0.0 || !(A[k] > 0.0)) continue;
         double wr = wbar[n] * rho[n * K + k];
         if (wr == 0.0) continue;
         for (int i = 0; i < d; i++) {
            double di = x[i] - B[k * d + i];
            for (int j = 0; j <= i; j++)
               Ct[k * dd + i * d + j] += wr * di * (x[j] - B[k * d + j]);
         }
      }
   }
   }
   for (int th = 1; th < nt; th++)
      for (size_t e = 0; e < (size_t)K * dd; e++) C[e] += C[(size_t)th * K * dd + e];
   /* install + cleanup_after_update: dead if alpha < 1/N or fewer than
    * MINCOUNT points sampled from it, or the new covariance is not PD */
   double wsum = 0.0;
   for (int k = 0; k < K; k++) {
      int was_alive = wght[k] != 0.0;
      int dead = !was_alive || !(A[k] >= 1.0 / (double)Nall) || count[k] < PMCB200_MINCOUNT;
      double cov[PMCB200_MAX_DIM * PMCB200_MAX_DIM];
      if (!dead) {
         for (int i = 0; i < d; i++)
            for (int j = 0; j <= i; j++)
               cov[i * d + j] = cov[j * d + i] = C[k * dd + i * d + j] / A[k];
         double Lk[PMCB200_MAX_DIM * PMCB200_MAX_DIM];
         memcpy(Lk, cov, dd * sizeof(double));
         if (orc_cholesky(d, Lk) != 0) dead = 1;
         else {
            memcpy(chol + k * dd, Lk, dd * sizeof(double));
            memcpy(mean + k * d, B + k * d, d * sizeof(double));
            if (cov_out) memcpy(cov_out + k * dd, cov, dd * sizeof(double));
            wght[k] = A[k];
         }
      }
      if (dead) {
         if (was_alive) ndead++;
         wght[k] = 0.0;
         if (cov_out) {   /* keep old covariance L L^T */
            for (int i = 0; i < d; i++)
               for (int j = 0; j < d; j++) {
                  double s = 0.0;
                  for (int q = 0; q < d; q++) s += chol[k * dd + i * d + q] * chol[k * dd + j * d + q];
                  cov_out[k * dd + i * d + j] = s;
               }
         }
      }
      wsum += wght[k];
   }
   if (wsum > 0.0) for (int k = 0; k < K; k++) wght[k] /= wsum;
   free(A); free(G); free(B); free(C); free(rho); free(count);
   return ndead;
}

/* ==========================================================================
 * One whole iteration (the body of run_pmc_iteration_MPI, cosmo_pmc.c:293-402,
 * plus the diagnostics of post_processing :441-461), used as the CPU baseline.
 * nthreads > 1 parallelises the weight stage over samples like the
 * reference's MPI scatter; sampling, normalisation and EM stay serial as on
 * the reference's rank 0.
 * ========================================================================== */
/* "Mode B" of the CPU baseline: every stage parallel over samples -- what a CPU port
 * would do given the freedom this repository took on the GPU.  Sampling is the same Philox stream (identical
 * draws); the EM sums are combined per thread. */
int orc_iteration_mode_b(const pmcb200_target_t *t, int64_t N, uint64_t seed,
                  uint32_t iter, double beta, int K, int d, int df,
                  double *wght, double *mean, double *chol, double *X,
                  int32_t *idx, int16_t *flg, double *w,
                  pmcb200_stats_t *st, int nthreads)
{
   memset(st, 0, sizeof(*st));
   st->nsamples = N;
   int nt = nthreads > 0 ? nthreads : 1;
#ifdef _OPENMP
   if (nthreads <= 0) nt = omp_get_max_threads();
#endif
   int64_t nok = 0;
#ifdef _OPENMP
#pragma omp parallel for num_threads(nt) reduction(+ : nok) schedule(static)
#endif
   for (int64_t c = 0; c < (N + 4095) / 4096; c++) {
      int64_t n0 = c * 4096, n1 = n0 + 4096 < N ? n0 + 4096 : N;
      nok += orc_simulate(n1 - n0, seed, iter, n0, K, d, df, wght, mean, chol, t->min, t->max, X + n0 * d, idx + n0, flg + n0);
   }
   st->nok_box = nok;
   if (st->nok_box == 0) return PMCB200_ERR_NOSAMPLE;
   st->nok = orc_importance_weights(t, N, X, K, d, df, wght, mean, chol, beta, flg, w, &st->maxW, nt);
   if (st->nok == 0) return PMCB200_ERR_NOSAMPLE;
   st->sum_shift = orc_normalize_weights(N, flg, w, st->maxW, &st->logSum);
   st->perplexity = orc_perplexity_and_ess(N, flg, w, &st->ess);
   st->ln_evidence = st->logSum - log((double)N);
   st->ndead = update_prop_rb_impl(N, X, idx, flg, w, K, d, df, wght, mean, chol, NULL, nt);
   st->enc = orc_enc(K, wght);
   return 0;
}

int orc_iteration(const pmcb200_target_t *t, int64_t N, uint64_t seed,
                  uint32_t iter, double beta, int K, int d, int df,
                  double *wght, double *mean, double *chol, double *X,
                  int32_t *idx, int16_t *flg, double *w,
                  pmcb200_stats_t *st, int nthreads)
{
   memset(st, 0, sizeof(*st));
   st->nsamples = N;
   st->nok_box = orc_simulate(N, seed, iter, 0, K, d, df, wght, mean, chol,
                              t->min, t->max, X, idx, flg);
   if (st->nok_box == 0) return PMCB200_ERR_NOSAMPLE;
   st->nok = orc_importance_weights(t, N, X, K, d, df, wght, mean, chol, beta,
                                    flg, w, &st->maxW, nthreads);
   if (st->nok == 0) return PMCB200_ERR_NOSAMPLE;
   st->sum_shift = orc_normalize_weights(N, flg, w, st->maxW, &st->logSum);
   st->perplexity = orc_perplexity_and_ess(N, flg, w, &st->ess);
   st->ln_evidence = st->logSum - log((double)N);
   st->ndead = orc_update_prop_rb(N, X, idx, flg, w, K, d, df, wght, mean, chol, NULL);
   st->enc = orc_enc(K, wght);
   return 0;
}
