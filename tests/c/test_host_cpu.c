/* test_host_cpu.c -- CPU-only checks of the pmclib-named host API: error stack,
 * mvdens / mix_mvdens allocation, text formats (Manual/manual.tex:3204-3255),
 * Cholesky / inverse, single-point log-pdf, parabox, gsl shim, and that the
 * batched entry points fail loudly (no CPU path) without a CUDA device.
 * Prints "ok <n>" on success; exits non-zero at the first failed check. */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "pmclib/pmc.h"
#include "gsl/gsl_randist.h"

static int nchecks = 0;
#define CHECK(c) do { nchecks++; if (!(c)) { fprintf(stderr, "FAILED %s:%d: %s\n", __FILE__, __LINE__, #c); exit(1); } } while (0)

static double fails(int code, error **err)
{
   testErrorRetVA(code != 0, mv_dimension, "code %d", *err, __LINE__, -1.0, code);
   return 1.0;
}
static double forwards(int code, error **err)
{
   double r = fails(code, err);
   forwardError(*err, __LINE__, -2.0);
   return r;
}

int main(int argc, char **argv)
{
   error *myerr = NULL, **err = &myerr;
   const char *tmp = argc > 1 ? argv[1] : "/tmp";
   char name[1024];

   /* error stack: callee appends, caller forwards, value of the originating error survives */
   CHECK(forwards(0, err) == 1.0 && !isError(*err));
   CHECK(forwards(7, err) == -2.0 && isError(*err));
   CHECK(getErrorValue(*err) == mv_dimension);
   CHECK(strstr((*err)->next->errText, "code 7") != NULL);
   purgeError(err);
   CHECK(!isError(*err));
   FILE *F = fopen_err("/nonexistent/dir/file", "r", err);
   CHECK(F == NULL && getErrorValue(*err) == io_file);
   purgeError(err);

   /* mvdens: the manual's example (manual.tex:3226-3234) */
   const char *example =
      "5 -1 5 0\n0.38559 -1.5238 19.338 1.3692 -2.4358 \n"
      "0.0053677 -0.025608 0.00066748 -0.0011893 0.00087517 \n-0.025608 0.16837 -0.0079163 0.0027364 -0.0035709 \n"
      "0.00066748 -0.0079163 0.0011077 0.0010986 -0.00067815 \n-0.0011893 0.0027364 0.0010986 0.016716 0.0026266 \n"
      "0.00087517 -0.0035709 -0.00067815 0.0026266 0.014881 \n";
   sprintf(name, "%s/mvd_example", tmp);
   F = fopen(name, "w"); fputs(example, F); fclose(F);
   F = fopen_err(name, "r", err);
   mvdens *g = mvdens_dwnp(F, err); fclose(F);
   CHECK(!isError(*err) && g->ndim == 5 && g->df == -1 && g->band_limit == 5 && g->chol == 0);
   CHECK(g->mean[2] == 19.338 && g->std[1 * 5 + 0] == -0.025608 && g->std[4 * 5 + 4] == 0.014881);
   /* dump is byte-identical to the example for %g-representable input */
   sprintf(name, "%s/mvd_dump", tmp);
   F = fopen(name, "w"); mvdens_dump(F, g); fclose(F);
   { char buf[4096]; F = fopen(name, "r"); size_t n = fread(buf, 1, sizeof(buf) - 1, F); buf[n] = 0; fclose(F);
     CHECK(strcmp(buf, example) == 0); }
   /* log-pdf at the mean = -(d/2) ln 2pi - (1/2) ln det; Cholesky on demand */
   double cov[25]; memcpy(cov, g->std, sizeof(cov));
   double lp = mvdens_log_pdf(g, g->mean, err);
   CHECK(!isError(*err) && g->chol == 1);
   double logdet = 0.0; for (int i = 0; i < 5; i++) logdet += 2.0 * log(g->std[i * 5 + i]);
   CHECK(fabs(lp - (-2.5 * log(2 * M_PI) - 0.5 * logdet)) < 1e-12);
   CHECK(fabs(g->detL - exp(0.5 * logdet)) < 1e-12 * g->detL);
   /* L L^T == covariance; dump after Cholesky still writes the covariance with c = 0 */
   for (int i = 0; i < 5; i++) for (int j = 0; j < 5; j++) {
      double s = 0; for (int k = 0; k < 5; k++) s += g->std[i * 5 + k] * g->std[j * 5 + k];
      CHECK(fabs(s - cov[i * 5 + j]) < 1e-15);
   }
   F = fopen(name, "w"); mvdens_dump(F, g); fclose(F);
   F = fopen(name, "r"); mvdens *g2 = mvdens_dwnp(F, err); fclose(F);
   CHECK(g2->chol == 0);
   for (int i = 0; i < 25; i++) CHECK(fabs(g2->std[i] - cov[i]) <= 1e-5 * fabs(cov[i]) + 1e-12);
   /* inverse twice = identity operation; returns det */
   double det = mvdens_inverse(g2, err);
   CHECK(!isError(*err) && fabs(det - exp(logdet)) < 1e-4 * det);
   { double id = 0; for (int k = 0; k < 5; k++) id += g2->std[0 * 5 + k] * cov[k * 5 + 0]; CHECK(fabs(id - 1.0) < 1e-4); }
   mvdens_inverse(g2, err);
   for (int i = 0; i < 25; i++) CHECK(fabs(g2->std[i] - cov[i]) <= 2e-5 * fabs(cov[i]) + 1e-11);
   /* not positive definite -> mv_cholesky */
   mvdens *bad = mvdens_alloc(2, err);
   double bm[2] = {0, 0}, bv[4] = {1, 2, 2, 1};
   mvdens_from_meanvar(bad, bm, bv, 1.0);
   mvdens_cholesky_decomp(bad, err);
   CHECK(isError(*err) && getErrorValue(*err) == mv_cholesky);
   purgeError(err);
   /* from_meanvar scale (fvar, param.c:536-551) */
   mvdens_from_meanvar(bad, bm, NULL, 1.8);
   CHECK(bad->std[0] == 1.8 && bad->std[1] == 0.0 && bad->chol == 0);

   /* mix_mvdens: alloc (one lump), weights view, dump / dwnp round trip, ENC, single-point log-pdf */
   mix_mvdens *m = mix_mvdens_alloc(3, 2, err);
   CHECK(!isError(*err) && m->ncomp == 3 && m->ndim == 2 && m->wght_view->size == 3);
   double w[3] = {0.2, 0.3, 0.5};
   for (int k = 0; k < 3; k++) {
      double mu[2] = {0.1 * k, -0.2 * k}, v[4] = {1.0 + k, 0.3, 0.3, 2.0};
      m->wght[k] = w[k];
      mvdens_from_meanvar(m->comp[k], mu, v, 1.0);
   }
   CHECK(fabs(effective_number_of_components(m, err) - 1.0 / 0.38) < 1e-12);
   gsl_vector_scale(m->wght_view, 2.0);                       /* cosmo_pmc.c:275 */
   CHECK(m->wght[2] == 1.0);
   gsl_vector_scale(m->wght_view, 0.5);
   sprintf(name, "%s/mix_dump", tmp);
   F = fopen(name, "w"); mix_mvdens_dump(F, m); fclose(F);
   F = fopen(name, "r"); mix_mvdens *m2 = mix_mvdens_dwnp(F, err); fclose(F);
   CHECK(!isError(*err) && m2->ncomp == 3 && m2->ndim == 2);
   for (int k = 0; k < 3; k++) {
      CHECK(fabs(m2->wght[k] - w[k]) < 1e-12 && m2->comp[k]->chol == 0);
      for (int i = 0; i < 4; i++) CHECK(m2->comp[k]->std[i] == m->comp[k]->std[i]);
   }
   mix_mvdens_cholesky_decomp(m2, err);
   double x[2] = {0.3, 0.1};
   double direct = 0.0;
   for (int k = 0; k < 3; k++) direct += w[k] * exp(mvdens_log_pdf(m2->comp[k], x, err));
   CHECK(fabs(mix_mvdens_log_pdf(m2, x, err) - log(direct)) < 1e-14);
   m2->wght[1] = 0.0;                                          /* dead component is skipped */
   CHECK(fabs(mix_mvdens_log_pdf_void(m2, x, err) - log(direct - w[1] * exp(mvdens_log_pdf(m2->comp[1], x, err)))) < 1e-13);
   mix_mvdens *m3 = mix_mvdens_alloc(3, 2, err);
   mix_mvdens_copy(m3, m2, err);
   CHECK(!isError(*err) && m3->comp[2]->chol == 1 && m3->comp[2]->std[3] == m2->comp[2]->std[3]);
   mix_mvdens *m4 = mix_mvdens_alloc(2, 2, err);
   mix_mvdens_copy(m4, m2, err);
   CHECK(isError(*err) && getErrorValue(*err) == mv_dimension);
   purgeError(err);

   /* parabox (param.c:915-926) */
   parabox *pb = init_parabox(2, err);
   add_slab(pb, 0, 0.0, 1.2, err); add_slab(pb, 1, -3.5, 0.5, err);
   double in[2] = {1.2, -3.5}, out[2] = {1.2000001, 0.0};
   CHECK(isinBox(pb, in, err) == 1 && isinBox(pb, out, err) == 0);
   add_slab(pb, 2, 0, 1, err);
   CHECK(isError(*err)); purgeError(err);

   /* gsl shim: seeded, reproducible, in range */
   gsl_rng *r = gsl_rng_alloc(gsl_rng_default);
   gsl_rng_set(r, 42); double a1 = gsl_rng_uniform(r), a2 = gsl_ran_flat(r, -1, 1);
   gsl_rng_set(r, 42); CHECK(a1 == gsl_rng_uniform(r) && a2 == gsl_ran_flat(r, -1, 1));
   CHECK(a1 >= 0 && a1 < 1 && a2 >= -1 && a2 < 1);
   double s1 = 0, s2 = 0; for (int i = 0; i < 200000; i++) { double gsn = gsl_ran_gaussian(r, 2.0); s1 += gsn; s2 += gsn * gsn; }
   CHECK(fabs(s1 / 200000) < 0.03 && fabs(s2 / 200000 - 4.0) < 0.08);

   /* pmc_simu container: one lump, realloc, field semantics */
   pmc_simu *psim = pmc_simu_init_mpi(100, 5, 1, err);
   CHECK(!isError(*err) && psim->nsamples == 100 && psim->ndim == 5 && psim->n_ded == 1);
   CHECK((char *)psim->X == (char *)psim->buf && psim->X_ded == psim->X + 500 && psim->weights == psim->X_ded + 100);
   pmc_simu_realloc(psim, 250, err);
   CHECK(!isError(*err) && psim->nsamples == 250);
   psim->flg[249] = 1; psim->indices[249] = 7; psim->X[249 * 5 + 4] = 1.0;
   for (int i = 0; i < 250; i++) { psim->flg[i] = 1; psim->weights[i] = 1.0 / 250; psim->X[i * 5] = i; }
   CHECK(fabs(mean_from_psim(psim->X, psim->weights, psim->flg, 250, 5, 0) - 124.5) < 1e-10);
   double pm[5], pv[25];
   estimate_param_covar_weight(5, 250, 0, psim->X, psim->weights, pm, pv, err);
   CHECK(fabs(pm[0] - 124.5) < 1e-10 && fabs(pv[0] - (250.0 * 250.0 - 1) / 12.0) < 1e-8);
   psim->isLog = 0;
   clip_weights(psim, 2, NULL, err);
   { int nf = 0; double ws = 0; for (int i = 0; i < 250; i++) { nf += psim->flg[i]; ws += psim->weights[i]; }
     CHECK(nf == 248 && fabs(ws - 1.0) < 1e-12); }

   /* batched entry points: loud failure without a device, never a CPU path */
   if (pmcb200_device_count() == 0) {
      simulate_mix_mvdens(psim, m3, r, NULL, err);
      CHECK(isError(*err) && getErrorValue(*err) == pmc_undef);
      purgeError(err);
      psim->isLog = 1;
      normalize_importance_weight(psim, err);
      CHECK(isError(*err)); purgeError(err);
   }
   pmc_simu_free(&psim); CHECK(psim == NULL);
   mvdens_free(&g); mvdens_free(&g2); mvdens_free(&bad);
   mix_mvdens_free(&m); mix_mvdens_free(&m2); mix_mvdens_free(&m3); mix_mvdens_free(&m4);
   free_parabox(&pb); gsl_rng_free(r);
   printf("ok %d\n", nchecks);
   return 0;
}
