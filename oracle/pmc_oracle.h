/* ============================================================================
 * pmc_oracle.h -- CPU restatement ("oracle") of CosmoPMC's PMC iteration.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing in the product path (cosmopmc_b200/,
 * include/) may call into this directory; only tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline / --impl reference legs do.
 *
 * PARITY UNPINNED: the arithmetic of this path lives in two external,
 * un-vendored, un-pinned libraries (github.com/CosmoStat/pmclib and
 * github.com/CosmoStat/nicaea, cloned at HEAD by install_CosmoPMC.sh:247,261)
 * that are absent from the reference tree and from this container, and the
 * reference stores no golden vectors for the path.
 * The oracle therefore restates the published algorithms (Numerical-Recipes
 * Romberg, Cappe et al. 2008 Rao-Blackwellised EM, Wraith et al. 2009,
 * Manual/manual.tex formulas) anchored on the reference's call sites, and is
 * cross-checked against scipy (tests/test_oracle_vs_scipy.py) and against the
 * reference's own perl restatements of evidence/ENC.
 * ========================================================================== */
#ifndef PMC_ORACLE_H
#define PMC_ORACLE_H

#include <stdint.h>
#include "../include/pmcb200.h"   /* POD descriptors only (no product code) */

#ifdef __cplusplus
extern "C" {
#endif

/* ---- linear algebra ------------------------------------------------------- */
int    orc_cholesky(int d, double *A);                 /* in place, lower; 0 ok */
/* ---- densities (pmclib mvdens.c; call sites cosmo_pmc.c:343, param.c:1023) */
double orc_mvdens_log_pdf(int d, int df, const double *mean, const double *chol,
                          const double *x);
double orc_mix_log_pdf(int K, int d, int df, const double *wght,
                       const double *mean, const double *chol, const double *x);
void   orc_mix_log_pdf_batch(int64_t N, int K, int d, int df, const double *wght,
                             const double *mean, const double *chol,
                             const double *X, double *out);
/* ---- sampler (pmclib simulate_mix_mvdens; call site cosmo_pmc.c:320) ------ */
void   orc_philox4x32_10(const uint32_t ctr[4], const uint32_t key[2],
                         uint32_t out[4]);
int    orc_select_component(int K, const double *wght, double u);
void   orc_sample_draws(uint64_t seed, uint32_t iter, int64_t g, int d, int df,
                        double *u, double *z, double *tscale);
int64_t orc_simulate(int64_t N, uint64_t seed, uint32_t iter, int64_t offset,
                     int K, int d, int df, const double *wght,
                     const double *mean, const double *chol, const double *bmin,
                     const double *bmax, double *X, int32_t *idx, int16_t *flg);
int64_t orc_simulate_from_draws(int64_t N, const double *u, const double *z,
                     int K, int d, const double *wght, const double *mean,
                     const double *chol, const double *bmin, const double *bmax,
                     double *X, int32_t *idx, int16_t *flg);
/* ---- cosmology (nicaea cosmo.c / sn1a.c / cmb_bao.c; call sites sn.c:260,270,
 *      bao.c:163-171, wmap.c:1034) ------------------------------------------- */
double orc_qromberg(double (*f)(double, void *), void *p, double a, double b,
                    double eps, int *nstage, int *err);
double orc_Esqr(const pmcb200_cosmo_t *c, double a, int wOmegar);
double orc_w(const pmcb200_cosmo_t *c, double a, int wOmegar, int *nstage, int *err);
double orc_f_K(const pmcb200_cosmo_t *c, double w, int wOmegar);
double orc_D_lum(const pmcb200_cosmo_t *c, double a, int *err);
double orc_r_sound(const pmcb200_cosmo_t *c, double a, int *nstage, int *err);
double orc_z_star(const pmcb200_cosmo_t *c);
double orc_z_drag(const pmcb200_cosmo_t *c);
double orc_loglike(const pmcb200_like_t *L, const double *x, int *err);
double orc_posterior_log_pdf(const pmcb200_target_t *t, const double *x, int *err);
void   orc_posterior_log_pdf_batch(const pmcb200_target_t *t, int64_t N,
                                   const double *X, double *out, int32_t *err,
                                   int nthreads);
double orc_sn_mean_stages(const pmcb200_like_t *L, const double *x);
int    orc_map_params(const pmcb200_like_t *L, const double *x, double out[16]);
/* ---- weights (pmclib pmc.c; call sites cosmo_pmc.c:343,378,46,62,84) ------ */
int64_t orc_importance_weights(const pmcb200_target_t *t, int64_t N,
                     const double *X, int K, int d, int df, const double *wght,
                     const double *mean, const double *chol, double beta,
                     int16_t *flg, double *logw, double *maxW, int nthreads);
double orc_normalize_weights(int64_t N, const int16_t *flg, double *w,
                             double maxW, double *logSum);
double orc_perplexity_and_ess(int64_t N, const int16_t *flg, const double *wbar,
                              double *ess);
double orc_enc(int K, const double *wght);
/* ---- EM (pmclib update_prop_rb; call site cosmo_pmc.c:247) ---------------- */
int    orc_update_prop_rb(int64_t N, const double *X, const int32_t *idx,
                          const int16_t *flg, const double *wbar, int K, int d,
                          int df, double *wght, double *mean, double *chol,
                          double *cov_out);
/* ---- whole iteration, for the CPU baseline -------------------------------- */
int    orc_iteration(const pmcb200_target_t *t, int64_t N, uint64_t seed,
                     uint32_t iter, double beta, int K, int d, int df,
                     double *wght, double *mean, double *chol, double *X,
                     int32_t *idx, int16_t *flg, double *w,
                     pmcb200_stats_t *st, int nthreads);

int    orc_iteration_mode_b(const pmcb200_target_t *t, int64_t N, uint64_t seed,
                     uint32_t iter, double beta, int K, int d, int df,
                     double *wght, double *mean, double *chol, double *X,
                     int32_t *idx, int16_t *flg, double *w,
                     pmcb200_stats_t *st, int nthreads);

#ifdef __cplusplus
}
#endif
#endif
