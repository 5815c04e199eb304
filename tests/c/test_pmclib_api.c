/* test_pmclib_api.c -- drives ONE PMC iteration through the pmclib-named host
 * API exactly in the order of run_pmc_iteration_MPI (reference
 * exec/cosmo_pmc.c:305-401, myid == 0, nproc == 1) and then the same iteration
 * through the fused pmc_b200_iteration; prints the results as `key value...`
 * lines for tests/test_gpu_host_c.py.
 *
 * usage: test_pmclib_api <sn_table> <proposal_in> <N> <seed> <beta> <outdir>      */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "pmclib/pmc.h"

static double my_posterior(void *data, const double *x, error **err)
{  /* stands for posterior_log_pdf_common_void; never called: the device target is */
   (void)data; (void)x;
   *err = addError(pmc_undef, "the scalar host callback must not be called", *err, __LINE__);
   return 0.0;
}

static int read_sn(const char *name, pmcb200_like_t *L)
{
   FILE *F = fopen(name, "r");
   if (!F) return -1;
   static double z[1024], m[1024], s[1024], c[1024], cov[6 * 1024];
   char line[4096];
   int n = 0;
   while (fgets(line, sizeof(line), F)) {
      if (line[0] == '#' || line[0] == '\n') continue;
      if (line[0] == '@') {
         char key[64]; double v;
         sscanf(line, "%63s %lg", key, &v);
         if (!strcmp(key, "@sig_int")) L->sn_sig_int = v;
         if (!strcmp(key, "@v_pec")) L->sn_v_pec = v;
         continue;
      }
      double *q = cov + 6 * n;
      if (sscanf(line, "%lg %lg %lg %lg %lg %lg %lg %lg %lg %lg", &z[n], &m[n], &s[n], &c[n], q, q + 1, q + 2, q + 3,
                 q + 4, q + 5) == 10) n++;
   }
   fclose(F);
   L->sn_n = n; L->sn_z = z; L->sn_m = m; L->sn_s = s; L->sn_c = c; L->sn_cov = cov;
   return n;
}

static void print_prop(const char *tag, mix_mvdens *p)
{
   printf("%s_wght", tag);
   for (size_t k = 0; k < p->ncomp; k++) printf(" %.17g", p->wght[k]);
   printf("\n%s_mean", tag);
   for (size_t k = 0; k < p->ncomp; k++) for (size_t i = 0; i < p->ndim; i++) printf(" %.17g", p->comp[k]->mean[i]);
   printf("\n%s_chol", tag);
   for (size_t k = 0; k < p->ncomp; k++) for (size_t i = 0; i < p->ndim * p->ndim; i++) printf(" %.17g", p->comp[k]->std[i]);
   printf("\n");
}

int main(int argc, char **argv)
{
   if (argc < 7) { fprintf(stderr, "usage: %s sn_table proposal N seed beta outdir\n", argv[0]); return 2; }
   error *myerr = NULL, **err = &myerr;
   long N = atol(argv[3]);
   unsigned long seed = strtoul(argv[4], NULL, 10);
   double beta = atof(argv[5]);
   char name[1024];

   /* target = the SN demo (Demo/MC_Demo/SN/config_pmc): what INTEGRATION.md's glue builds from config_base */
   static pmcb200_target_t t;
   memset(&t, 0, sizeof(t));
   const double lo[5] = {0.0, -3.5, 19.1, 0.5, -3.5}, hi[5] = {1.2, 0.5, 19.8, 2.6, -0.8};
   const int par[5] = {PMCB200_P_Omegam, PMCB200_P_w0de, PMCB200_P_M, PMCB200_P_alpha, PMCB200_P_beta};
   t.npar = 5; t.ndata = 1;
   pmcb200_like_t *L = &t.like[0];
   L->kind = PMCB200_LIKE_SNIa; L->npar = 5; L->special = PMCB200_SPECIAL_none;
   for (int j = 0; j < 5; j++) { t.min[j] = lo[j]; t.max[j] = hi[j]; L->par[j] = par[j]; }
   pmcb200_cosmo_t c0 = {0.27, 0.73, -1.0, 0.0, 0.73, 0.049, 0.0, 0.0, PMCB200_DE_linder, 0};
   L->model = c0;
   L->sn_chi2mode = PMCB200_CHI2_simple; L->sn_add_logdetCov = 0;
   L->sn_Theta2[0] = 19.31; L->sn_Theta2[1] = 1.6; L->sn_Theta2[2] = -1.8; L->sn_Theta2[3] = 0.0;
   if (read_sn(argv[1], L) < 1) { fprintf(stderr, "cannot read %s\n", argv[1]); return 2; }
   int dummy_config = 0;
   pmc_b200_register_target(my_posterior, &dummy_config, &t, err);           quitOnError(*err, __LINE__, stderr);

   /* proposal: read from the mix_mvdens text format, as `sinitial file` does (param.c:625-640) */
   FILE *F = fopen_err(argv[2], "r", err);                                    quitOnError(*err, __LINE__, stderr);
   mix_mvdens *proposal = mix_mvdens_dwnp(F, err);                            quitOnError(*err, __LINE__, stderr);
   fclose(F);
   mix_mvdens_cholesky_decomp(proposal, err);                                 quitOnError(*err, __LINE__, stderr);
   mix_mvdens *proposal2 = mix_mvdens_alloc(proposal->ncomp, proposal->ndim, err);
   mix_mvdens_copy(proposal2, proposal, err);                                 quitOnError(*err, __LINE__, stderr);
   printf("enc0 %.17g\n", effective_number_of_components(proposal, err));

   parabox *pb = init_parabox(5, err);
   for (int j = 0; j < 5; j++) add_slab(pb, j, lo[j], hi[j], err);
   quitOnError(*err, __LINE__, stderr);
   gsl_rng *rng = gsl_rng_alloc(gsl_rng_default);
   gsl_rng_set(rng, seed);

   /* ---- the body of run_pmc_iteration_MPI, cosmo_pmc.c:313-399 ---- */
   pmc_simu *psim = pmc_simu_init_mpi(N, 5, 0, err);                          quitOnError(*err, __LINE__, stderr);
   pmc_simu_realloc(psim, N, err);                                            quitOnError(*err, __LINE__, stderr);
   sprintf(name, "%s/proposal", argv[6]);
   F = fopen_err(name, "w", err);                                             quitOnError(*err, __LINE__, stderr);
   mix_mvdens_dump(F, proposal); fclose(F);
   size_t nok = simulate_mix_mvdens(psim, proposal, rng, pb, err);            quitOnError(*err, __LINE__, stderr);
   printf("nok_box %zu\nnshards %d\n", nok, pmc_b200_nshards());
   nok = generic_get_importance_weight_and_deduced_verb(psim, proposal, mix_mvdens_log_pdf_void, my_posterior, NULL,
                                                        &dummy_config, beta, 1, err);
   quitOnError(*err, __LINE__, stderr);
   printf("nok %zu\nisLog %d\nmaxW %.17g\n", nok, psim->isLog, psim->maxW);
   double ln_evi_log;
   evidence(psim, &ln_evi_log, err);                                          quitOnError(*err, __LINE__, stderr);
   printf("ln_evidence_from_log %.17g\n", ln_evi_log);
   double norm = normalize_importance_weight(psim, err);                      quitOnError(*err, __LINE__, stderr);
   printf("norm %.17g\nlogSum %.17g\nisLog_after %d\n", norm, psim->logSum, psim->isLog);
   update_prop_rb(proposal, psim, err);                                       quitOnError(*err, __LINE__, stderr);
   /* pmclib semantics: the updated components hold covariances (chol = 0) */
   for (size_t k = 0; k < proposal->ncomp; k++)
      if (proposal->wght[k] > 0 && proposal->comp[k]->chol != 0) { printf("MISMATCH chol flag\n"); return 1; }
   mix_mvdens_cholesky_decomp(proposal, err);                                 quitOnError(*err, __LINE__, stderr);
   /* ---- post_processing, cosmo_pmc.c:441-461 ---- */
   double ess, ln_evi;
   double perp = perplexity_and_ess(psim, MC_UNORM, &ess, err);               quitOnError(*err, __LINE__, stderr);
   evidence(psim, &ln_evi, err);                                              quitOnError(*err, __LINE__, stderr);
   printf("perplexity %.17g\ness %.17g\nln_evidence %.17g\n", perp, ess, ln_evi);
   printf("enc %.17g\n", effective_number_of_components(proposal, err));
   print_prop("staged", proposal);
   double wsum = 0.0, xw0 = 0.0; long nflag = 0;
   for (long i = 0; i < psim->nsamples; i++) if (psim->flg[i]) { wsum += psim->weights[i]; xw0 += psim->weights[i] * psim->X[i * 5]; nflag++; }
   printf("wsum %.17g\nmean0 %.17g\nmean0_lib %.17g\nnflag %ld\n", wsum, xw0, mean_from_psim(psim->X, psim->weights, psim->flg, psim->nsamples, 5, 0), nflag);
   printf("idx_first %zu %zu %zu %zu\n", psim->indices[0], psim->indices[1], psim->indices[2], psim->indices[3]);
   printf("x_first %.17g %.17g %.17g %.17g %.17g\n", psim->X[0], psim->X[1], psim->X[2], psim->X[3], psim->X[4]);
   sprintf(name, "%s/proposal_updated", argv[6]);
   F = fopen_err(name, "w", err);                                             quitOnError(*err, __LINE__, stderr);
   mix_mvdens_dump(F, proposal); fclose(F);

   /* ---- binary pmcsim sidecar: dump, read back through pmc_simu_from_file, compare with the live psim ---- */
   {
      sprintf(name, "%s/pmcsim.bin", argv[6]);
      F = fopen_err(name, "wb", err);                                         quitOnError(*err, __LINE__, stderr);
      pmc_simu_dump_binary(F, psim, err);                                     quitOnError(*err, __LINE__, stderr);
      fclose(F);
      F = fopen_err(name, "rb", err);                                         quitOnError(*err, __LINE__, stderr);
      pmc_simu *psim3 = pmc_simu_from_file(F, N, 5, 0, NULL, 0, err);         quitOnError(*err, __LINE__, stderr);
      fclose(F);
      long k = 0; double maxrel = 0.0; int bad = 0;
      for (long i = 0; i < psim->nsamples; i++) {
         if (!psim->flg[i]) continue;
         if (!psim3->flg[k] || psim3->indices[k] != psim->indices[i] ||
             memcmp(psim3->X + 5 * k, psim->X + 5 * i, 5 * sizeof(double)) != 0) bad++;
         if (psim->weights[i] > 0) {
            double r = fabs(psim3->weights[k] - psim->weights[i]) / psim->weights[i];
            if (r > maxrel) maxrel = r;
         }
         k++;
      }
      printf("bin_roundtrip %ld %d %.3g %.17g %ld\n", k, bad, maxrel, psim3->logSum, psim3->nsamples);
      pmc_simu_free(&psim3);
   }

   /* ---- the same iteration through the fused call ---- */
   gsl_rng_set(rng, seed);
   pmc_simu *psim2 = pmc_simu_init_mpi(N, 5, 0, err);
   pmcb200_stats_t st;
   size_t nok2 = pmc_b200_iteration(psim2, proposal2, rng, beta, &st, err);   quitOnError(*err, __LINE__, stderr);
   printf("fused_nok %zu\nfused_perplexity %.17g\nfused_ess %.17g\nfused_logSum %.17g\nfused_enc %.17g\n", nok2,
          st.perplexity, st.ess, st.logSum, st.enc);
   mix_mvdens_cholesky_decomp(proposal2, err);                                quitOnError(*err, __LINE__, stderr);
   print_prop("fused", proposal2);
   double dmax = 0.0;
   for (long i = 0; i < N; i++) {
      double dw = fabs(psim->weights[i] - psim2->weights[i]);
      if (dw > dmax) dmax = dw;
      if (psim->flg[i] != psim2->flg[i] || psim->indices[i] != psim2->indices[i]) { printf("MISMATCH at %ld\n", i); return 1; }
   }
   printf("max_abs_dw %.3g\n", dmax);

   /* error behaviour: an unregistered callback must fail loudly, not fall back */
   int other = 0;
   generic_get_importance_weight_and_deduced_verb(psim, proposal, mix_mvdens_log_pdf_void, my_posterior, NULL, &other,
                                                  1.0, 1, err);
   printf("unregistered_is_error %d %d\n", isError(*err), getErrorValue(*err));
   purgeError(err);

   {  /* mirror bookkeeping: X went up at most once in the staged iteration above (never, in fact: simulate leaves it
         on the device), and a host-side edit of single elements followed by pmc_b200_invalidate_mirror is seen */
      long up = 0, skipped = 0;
      pmc_b200_mirror_traffic(&up, &skipped);
      printf("mirror_uploaded %ld\nmirror_skipped %ld\n", up, skipped);
      double before, ess_b;
      before = perplexity_and_ess(psim, MC_UNORM, &ess_b, err);               quitOnError(*err, __LINE__, stderr);
      long imax = 0;
      for (long i = 1; i < psim->nsamples; i++) if (psim->weights[i] > psim->weights[imax]) imax = i;
      psim->weights[imax] = 0.0;                                              /* single-element edit */
      pmc_b200_invalidate_mirror(psim);
      double after = perplexity_and_ess(psim, MC_UNORM, &ess_b, err);         quitOnError(*err, __LINE__, stderr);
      printf("perplexity_edit_seen %d\n", after != before);
   }
   pmc_simu_free(&psim); pmc_simu_free(&psim2);
   mix_mvdens_free(&proposal); mix_mvdens_free(&proposal2);
   free_parabox(&pb); gsl_rng_free(rng);
   pmc_b200_shutdown();
   printf("done 1\n");
   return 0;
}
