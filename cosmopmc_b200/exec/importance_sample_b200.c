/* importance_sample_b200.c -- batched counterpart of the reference's exec/importance_sample.c.
 *
 * Same command line, same input and output files (pmcsim text format, exec_helper.c:351-424),
 * same arithmetic: every point of an existing PMC sample is re-weighted under the posterior of
 * a second config file,
 *      log w_new(x) = log pi_2(x) + log w_prev(x)          (importance_sample.c:228-283)
 * and the result is normalised and written out.  The reference evaluates pi_2 through one scalar
 * host callback per point (importance_sample.c:61) and spreads the points over MPI ranks; this
 * driver hands the whole sample to the device posterior in one call
 * (pmc_b200_importance_sample: one kernel launch per data set and GPU shard).
 *
 * Compiled with the reference's own headers next to its unchanged wrappers/ and tools/ sources
 * (tools/build_ref_cosmo_pmc.py), like the glue unit. */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>
#include <unistd.h>

#include "param.h"
#include "exec_helper.h"
#include "pmclib/pmc.h"

static void usage(int ex)
{
   fprintf(stderr, "Usage: importance_sample_b200 [OPTIONS] INSAMPLE\n"
                   "Importance run on a PMC sample; the posterior is evaluated on the GPU(s)\n"
                   "(PMCB200_NGPU shards the sample over several devices).\n"
                   "OPTIONS:\n"
                   "  -c CONFIG        Configuration file (default: config_pmc)\n"
                   "  -o OUTSAMPLE     Output sample name (default: 'INSAMPLE.out')\n"
                   "  -q               Quiet mode\n"
                   "  -h               This message\n");
   exit(ex);
}

int main(int argc, char *argv[])
{
   error *myerr = NULL, **err = &myerr;
   const char *cname = "config_pmc";
   char *outname = NULL;
   int quiet = 0, c;

   while ((c = getopt(argc, argv, ":c:o:qh")) != -1) {
      switch (c) {
         case 'c': cname = optarg; break;
         case 'o': outname = optarg; break;
         case 'q': quiet = 1; break;
         case 'h': usage(0); break;
         default:  usage(2);
      }
   }
   if (argc - optind != 1) usage(argc - optind > 1 ? 6 : 7);
   const char *inname = argv[optind];
   if (!outname) {
      outname = (char *)malloc_err(strlen(inname) + 8, err);          quitOnError(*err, __LINE__, stderr);
      sprintf(outname, "%s.out", inname);
   }
   time_t t_start;
   time(&t_start);

   config_pmc config;
   read_config_pmc_file(&config, cname, NULL, NULL, 0, err);          quitOnError(*err, __LINE__, stderr);

   /* the stored sample: normalised weights + logSum (pmc_simu_from_file normalises) */
   FILE *F = fopen_err(inname, "r", err);                             quitOnError(*err, __LINE__, stderr);
   pmc_simu *psim = pmc_simu_from_file(F, config.nsamples * config.fsfinal, config.base.npar, config.base.n_ded,
                                       NULL, config.nclipw, err);
   quitOnError(*err, __LINE__, stderr);
   fclose(F);
   if (psim->ndim != config.base.npar) {
      fprintf(stderr, "Number of parameters differs between %s (npar=%d) and %s (npar=%d)\n", cname, config.base.npar,
              inname, psim->ndim);
      return 3;
   }
   long n = psim->nsamples;
   double *lw_prev = (double *)malloc_err(sizeof(double) * (size_t)n, err);
   short *flg_prev = (short *)malloc_err(sizeof(short) * (size_t)n, err);
   quitOnError(*err, __LINE__, stderr);
   for (long i = 0; i < n; i++) {
      flg_prev[i] = psim->flg[i];
      lw_prev[i] = (psim->flg[i] && psim->weights[i] > 0.0) ? log(psim->weights[i]) + psim->logSum : -HUGE_VAL;
   }

   /* log pi_2 at every point, batched on the device */
   size_t nok = pmc_b200_importance_sample(psim, posterior_log_pdf_common_void, &config.base, err);
   quitOnError(*err, __LINE__, stderr);
   if (!quiet) fprintf(stderr, "importance weights for %ld points on %d GPU shard(s), nok=%zu\n", n, pmc_b200_nshards(), nok);

   /* log w_new = log pi_2 + log w_prev for the points the input file held */
   double MW = -HUGE_VAL;
   for (long i = 0; i < n; i++) {
      if (!flg_prev[i] || !psim->flg[i] || !(lw_prev[i] > -HUGE_VAL)) { psim->flg[i] = 0; psim->weights[i] = 0.0; continue; }
      psim->weights[i] += lw_prev[i];
      if (psim->weights[i] > MW) MW = psim->weights[i];
   }
   psim->maxW = MW;
   psim->isLog = 1;
   double norm = normalize_importance_weight(psim, err);              quitOnError(*err, __LINE__, stderr);

   out_pmc_simu_cosmo_pmc(outname, psim, config.base.par, norm, err); quitOnError(*err, __LINE__, stderr);

   free(lw_prev); free(flg_prev);
   pmc_simu_free(&psim);
   pmc_b200_shutdown();
   end_time(t_start, stderr);
   fprintf(stderr, "importance_sample_b200 finished\n");
   return 0;
}
