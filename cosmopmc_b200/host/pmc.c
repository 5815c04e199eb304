/* pmc.c -- pmclib-named host shell of the PMC iteration (include/pmclib/pmc.h).
 * Marshals `pmc_simu` / `mix_mvdens` to the C-ABI of include/pmcb200.h; all
 * batched arithmetic happens in CUDA kernels.  No CPU path for the batched
 * functions. */
#include "pmclib/pmc.h"
#include <math.h>
#include <stdint.h>

/* ---- parabox ------------------------------------------------------------------ */
parabox *init_parabox(int ndim, error **err)
{
   parabox *pb = (parabox *)calloc_err(1, sizeof(parabox), err);
   forwardError(*err, __LINE__, NULL);
   pb->ndim = ndim;
   pb->min = (double *)calloc_err(ndim, sizeof(double), err);  forwardError(*err, __LINE__, NULL);
   pb->max = (double *)calloc_err(ndim, sizeof(double), err);  forwardError(*err, __LINE__, NULL);
   pb->set = (int *)calloc_err(ndim, sizeof(int), err);        forwardError(*err, __LINE__, NULL);
   for (int i = 0; i < ndim; i++) { pb->min[i] = -HUGE_VAL; pb->max[i] = HUGE_VAL; }
   return pb;
}

void add_slab(parabox *pb, int idim, double sinf, double ssup, error **err)
{
   testErrorRetVA(idim < 0 || idim >= pb->ndim, pb_outOfBound, "Dimension %d out of range [0,%d)", *err, __LINE__, ,
                  idim, pb->ndim);
   testErrorRetVA(!(ssup > sinf), pb_outOfBound, "Empty slab [%g,%g] for dimension %d", *err, __LINE__, , sinf, ssup, idim);
   pb->min[idim] = sinf; pb->max[idim] = ssup; pb->set[idim] = 1;
}

void free_parabox(parabox **pb)
{
   if (!pb || !*pb) return;
   free((*pb)->min); free((*pb)->max); free((*pb)->set); free(*pb);
   *pb = NULL;
}

int isinBox(const parabox *pb, const double *pos, error **err)
{
   (void)err;
   for (int i = 0; i < pb->ndim; i++)
      if (!(pos[i] >= pb->min[i] && pos[i] <= pb->max[i])) return 0;
   return 1;
}

/* ---- pmc_simu --------------------------------------------------------------------- */
static void psim_carve(pmc_simu *p, long n)
{
   /* lump: X, X_ded, weights, indices, flg */
   char *b = (char *)p->buf;
   p->X = (double *)b;                 b += sizeof(double) * (size_t)n * p->ndim;
   p->X_ded = (double *)b;             b += sizeof(double) * (size_t)n * p->n_ded;
   p->weights = (double *)b;           b += sizeof(double) * (size_t)n;
   p->indices = (size_t *)b;           b += sizeof(size_t) * (size_t)n;
   p->flg = (short *)b;
}
static size_t psim_bytes(long n, int ndim, int n_ded)
{
   return (size_t)n * (sizeof(double) * (ndim + n_ded + 1) + sizeof(size_t) + sizeof(short)) + 64;
}

pmc_simu *pmc_simu_init_plus_ded(long nsamples, int ndim, int n_ded, error **err)
{
   testErrorRetVA(nsamples < 1 || ndim < 1 || n_ded < 0, pmc_dimension, "Invalid pmc_simu size (%ld,%d,%d)", *err,
                  __LINE__, NULL, nsamples, ndim, n_ded);
   pmc_simu *p = (pmc_simu *)calloc_err(1, sizeof(pmc_simu), err);
   forwardError(*err, __LINE__, NULL);
   p->nsamples = p->nsamples_alloc = nsamples; p->ndim = ndim; p->n_ded = n_ded;
   p->buf = calloc_err(1, psim_bytes(nsamples, ndim, n_ded), err);
   forwardError(*err, __LINE__, NULL);
   psim_carve(p, nsamples);
   p->isLog = 0; p->logSum = 0.0; p->maxW = 0.0; p->mpi_rank = 0; p->mpi_size = 1;
   return p;
}
pmc_simu *pmc_simu_init(long nsamples, int ndim, error **err) { return pmc_simu_init_plus_ded(nsamples, ndim, 0, err); }
pmc_simu *pmc_simu_init_mpi(long nsamples, int ndim, int n_ded, error **err)
{
   return pmc_simu_init_plus_ded(nsamples, ndim, n_ded, err);
}

void pmc_simu_realloc(pmc_simu *p, long nsamples, error **err)
{
   if (nsamples <= p->nsamples_alloc && nsamples == p->nsamples) return;
   testErrorRetVA(nsamples < 1, pmc_dimension, "Invalid number of samples %ld", *err, __LINE__, , nsamples);
   free(p->buf);
   p->buf = calloc_err(1, psim_bytes(nsamples, p->ndim, p->n_ded), err);
   forwardError(*err, __LINE__, );
   p->nsamples = p->nsamples_alloc = nsamples;
   psim_carve(p, nsamples);
   p->isLog = 0;
}

void pmc_simu_free(pmc_simu **p)
{
   if (!p || !*p) return;
   free((*p)->buf); free(*p); *p = NULL;
}

/* ---- device context + target registry ------------------------------------------------
 * One process drives one GPU and, like the reference (single-threaded, no locks anywhere,
 * SURVEY.md 8b), this layer keeps process-wide state: it is not re-entrant. */
#define MAX_TARGETS 8
static pmcb200_ctx *g_ctx = NULL;
static struct { posterior_log_pdf_func *f; void *data; pmcb200_target_t t; int used; } g_targets[MAX_TARGETS];
static const void *g_active_target = NULL;
/* device mirrors of the last psim (grow-only) */
static struct { void *X, *idx, *flg, *w, *block; long cap; int d; long blen; } g_dev;

#define B200_OK(ctx, call, errcode, ret)                                                          \
   do { int rc__ = (call);                                                                        \
        if (rc__ != 0) { *err = addErrorVA((errcode), "%s (pmcb200 code %d)", *err, __LINE__,     \
                                           pmcb200_last_error(ctx), rc__); return ret; } } while (0)

pmcb200_ctx *pmc_b200_context(error **err)
{
   if (g_ctx) return g_ctx;
   const char *e = getenv("PMCB200_DEVICE");
   int dev = e ? atoi(e) : 0;
   int rc = pmcb200_create(dev, NULL, &g_ctx);
   if (rc != 0) {
      g_ctx = NULL;
      *err = addErrorVA(pmc_undef, "No usable CUDA device %d (pmcb200 code %d); the PMC iteration has no CPU path",
                        *err, __LINE__, dev, rc);
      return NULL;
   }
   return g_ctx;
}

void pmc_b200_shutdown(void)
{
   if (!g_ctx) return;
   if (g_dev.X) pmcb200_dev_free(g_ctx, g_dev.X);
   if (g_dev.idx) pmcb200_dev_free(g_ctx, g_dev.idx);
   if (g_dev.flg) pmcb200_dev_free(g_ctx, g_dev.flg);
   if (g_dev.w) pmcb200_dev_free(g_ctx, g_dev.w);
   if (g_dev.block) pmcb200_dev_free(g_ctx, g_dev.block);
   memset(&g_dev, 0, sizeof(g_dev));
   pmcb200_destroy(g_ctx);
   g_ctx = NULL; g_active_target = NULL;
}

void pmc_b200_register_target(posterior_log_pdf_func *f, void *data, const pmcb200_target_t *t, error **err)
{
   for (int i = 0; i < MAX_TARGETS; i++)
      if (!g_targets[i].used || (g_targets[i].f == f && g_targets[i].data == data)) {
         g_targets[i].f = f; g_targets[i].data = data; g_targets[i].t = *t; g_targets[i].used = 1;
         if (g_active_target == &g_targets[i].t) g_active_target = NULL;
         return;
      }
   *err = addError(pmc_outOfBound, "Too many registered device targets", *err, __LINE__);
}

static void activate_target(pmcb200_ctx *ctx, posterior_log_pdf_func *f, void *data, error **err)
{
   for (int i = 0; i < MAX_TARGETS; i++)
      if (g_targets[i].used && g_targets[i].f == f && g_targets[i].data == data) {
         if (g_active_target != &g_targets[i].t) {
            B200_OK(ctx, pmcb200_set_target(ctx, &g_targets[i].t), pmc_incompat, );
            g_active_target = &g_targets[i].t;
         }
         return;
      }
   {  /* not registered: give the caller-side glue a chance to flatten its own config */
      static pmcb200_target_t t;
      memset(&t, 0, sizeof(t));
      int bound = pmc_b200_autobind(f, data, &t, err);
      forwardError(*err, __LINE__, );
      if (bound) {
         pmc_b200_register_target(f, data, &t, err);
         forwardError(*err, __LINE__, );
         activate_target(ctx, f, data, err);
         forwardError(*err, __LINE__, );
         return;
      }
   }
   *err = addError(pmc_undef, "No device target registered for this posterior callback "
                   "(pmc_b200_register_target); the scalar host callback cannot be batched and there is no CPU path",
                   *err, __LINE__);
}

static void ensure_dev(pmcb200_ctx *ctx, long n, int d, long blen, error **err)
{
   if (n > g_dev.cap || d > g_dev.d) {
      if (g_dev.X) { pmcb200_dev_free(ctx, g_dev.X); pmcb200_dev_free(ctx, g_dev.idx);
                     pmcb200_dev_free(ctx, g_dev.flg); pmcb200_dev_free(ctx, g_dev.w); }
      long cap = n > g_dev.cap ? n : g_dev.cap;
      int dd = d > g_dev.d ? d : g_dev.d;
      B200_OK(ctx, pmcb200_dev_alloc(ctx, sizeof(double) * (size_t)cap * dd, &g_dev.X), pmc_allocate, );
      B200_OK(ctx, pmcb200_dev_alloc(ctx, sizeof(int32_t) * (size_t)cap, &g_dev.idx), pmc_allocate, );
      B200_OK(ctx, pmcb200_dev_alloc(ctx, sizeof(int16_t) * (size_t)cap, &g_dev.flg), pmc_allocate, );
      B200_OK(ctx, pmcb200_dev_alloc(ctx, sizeof(double) * (size_t)cap, &g_dev.w), pmc_allocate, );
      g_dev.cap = cap; g_dev.d = dd;
   }
   if (blen > g_dev.blen) {
      if (g_dev.block) pmcb200_dev_free(ctx, g_dev.block);
      B200_OK(ctx, pmcb200_dev_alloc(ctx, sizeof(double) * (size_t)blen, &g_dev.block), pmc_allocate, );
      g_dev.blen = blen;
   }
}

/* proposal (Cholesky-decomposed mix_mvdens) -> device */
static void push_proposal(pmcb200_ctx *ctx, mix_mvdens *m, error **err)
{
   size_t K = m->ncomp, d = m->ndim;
   mix_mvdens_cholesky_decomp(m, err);
   forwardError(*err, __LINE__, );
   double *mean = (double *)malloc_err(sizeof(double) * K * d * (d + 1), err);
   forwardError(*err, __LINE__, );
   double *chol = mean + K * d;
   for (size_t k = 0; k < K; k++) {
      memcpy(mean + k * d, m->comp[k]->mean, d * sizeof(double));
      memcpy(chol + k * d * d, m->comp[k]->std, d * d * sizeof(double));
   }
   int rc = pmcb200_set_proposal(ctx, (int)K, (int)d, m->comp[0]->df, m->wght, mean, chol);
   free(mean);
   if (rc != 0) *err = addErrorVA(rc == PMCB200_ERR_CHOLESKY ? pmc_cholesky : pmc_incompat, "%s", *err, __LINE__,
                                  pmcb200_last_error(ctx));
}

/* updated proposal <- device.  As in pmclib, update_prop_rb leaves the COVARIANCE in comp[k]->std
 * (chol = 0; the Cholesky factor is recomputed on demand by the next use): the reference copies
 * std between components as a covariance (revive_comp, exec/cosmo_pmc.c:225).  Components that died
 * in the update (weight 0, cleanup_after_update) keep their previous mean and matrix. */
static void pull_proposal(pmcb200_ctx *ctx, mix_mvdens *m, error **err)
{
   size_t K = m->ncomp, d = m->ndim;
   double *mean = (double *)malloc_err(sizeof(double) * K * d * (2 * d + 1), err);
   forwardError(*err, __LINE__, );
   double *chol = mean + K * d, *cov = chol + K * d * d;
   int rc = pmcb200_get_proposal(ctx, m->wght, mean, chol, cov);
   if (rc == 0)
      for (size_t k = 0; k < K; k++) {
         if (m->wght[k] == 0.0) continue;
         memcpy(m->comp[k]->mean, mean + k * d, d * sizeof(double));
         memcpy(m->comp[k]->std, cov + k * d * d, d * d * sizeof(double));
         m->comp[k]->chol = 0;
         m->comp[k]->detL = 0.0;
      }
   free(mean);
   m->init_cwght = 0;
   if (rc != 0) *err = addErrorVA(pmc_incompat, "%s", *err, __LINE__, pmcb200_last_error(ctx));
}

static void push_samples(pmcb200_ctx *ctx, pmc_simu *p, int with_idx, int with_w, error **err)
{
   long n = p->nsamples;
   B200_OK(ctx, pmcb200_h2d(ctx, g_dev.X, p->X, sizeof(double) * (size_t)n * p->ndim), pmc_badComm, );
   B200_OK(ctx, pmcb200_h2d(ctx, g_dev.flg, p->flg, sizeof(short) * (size_t)n), pmc_badComm, );
   if (with_w) B200_OK(ctx, pmcb200_h2d(ctx, g_dev.w, p->weights, sizeof(double) * (size_t)n), pmc_badComm, );
   if (with_idx) {
      int32_t *t = (int32_t *)malloc_err(sizeof(int32_t) * (size_t)n, err);
      forwardError(*err, __LINE__, );
      for (long i = 0; i < n; i++) t[i] = (int32_t)p->indices[i];
      int rc = pmcb200_h2d(ctx, g_dev.idx, t, sizeof(int32_t) * (size_t)n);
      free(t);
      B200_OK(ctx, rc, pmc_badComm, );
   }
}

__attribute__((weak)) int pmc_b200_autobind(posterior_log_pdf_func *f, void *data, pmcb200_target_t *t, error **err)
{
   (void)f; (void)data; (void)t; (void)err;
   return 0;
}

/* one log-likelihood on the device with N = 1 (own context, so the PMC run's
 * target stays resident).  Box = [x-1/2, x+1/2] so the flat-prior term is 0. */
double pmc_b200_single_loglike(const pmcb200_like_t *like, const double *x, error **err)
{
   static pmcb200_ctx *ctx1 = NULL;
   static void *dx = NULL, *dlp = NULL, *derr = NULL;
   if (!ctx1) {
      const char *e = getenv("PMCB200_DEVICE");
      int rc = pmcb200_create(e ? atoi(e) : 0, NULL, &ctx1);
      if (rc != 0) { ctx1 = NULL;
         *err = addErrorVA(pmc_undef, "No usable CUDA device (pmcb200 code %d); no CPU path", *err, __LINE__, rc);
         return 0.0; }
      pmcb200_dev_alloc(ctx1, sizeof(double) * PMCB200_MAX_DIM, &dx);
      pmcb200_dev_alloc(ctx1, sizeof(double), &dlp);
      pmcb200_dev_alloc(ctx1, sizeof(int32_t), &derr);
   }
   static pmcb200_target_t t;
   memset(&t, 0, sizeof(t));
   t.npar = like->npar; t.ndata = 1; t.like[0] = *like;
   for (int j = 0; j < like->npar; j++) { t.min[j] = x[j] - 0.5; t.max[j] = x[j] + 0.5; }
   B200_OK(ctx1, pmcb200_set_target(ctx1, &t), pmc_incompat, 0.0);
   B200_OK(ctx1, pmcb200_h2d(ctx1, dx, x, sizeof(double) * like->npar), pmc_badComm, 0.0);
   B200_OK(ctx1, pmcb200_posterior_log_pdf(ctx1, 1, (double *)dx, (double *)dlp, (int32_t *)derr), pmc_undef, 0.0);
   double lp = 0.0; int32_t e1 = 0;
   B200_OK(ctx1, pmcb200_d2h(ctx1, &lp, dlp, sizeof(double)), pmc_badComm, 0.0);
   B200_OK(ctx1, pmcb200_d2h(ctx1, &e1, derr, sizeof(int32_t)), pmc_badComm, 0.0);
   testErrorRet(e1 != 0, pmc_infnan, "Likelihood could not be evaluated for this model (unphysical distance integral)",
                *err, __LINE__, 0.0);
   return lp;
}

/* ---- simulate_mix_mvdens (cosmo_pmc.c:320) -------------------------------------------- */
size_t simulate_mix_mvdens(pmc_simu *psim, mix_mvdens *proposal, gsl_rng *r, parabox *pb, error **err)
{
   pmcb200_ctx *ctx = pmc_b200_context(err);
   forwardError(*err, __LINE__, 0);
   testErrorRetVA((size_t)psim->ndim != proposal->ndim, pmc_dimension, "psim ndim %d != proposal ndim %zu", *err,
                  __LINE__, 0, psim->ndim, proposal->ndim);
   testErrorRet(pb != NULL && pb->ndim != psim->ndim, pmc_dimension, "parabox dimension mismatch", *err, __LINE__, 0);
   long n = psim->nsamples;
   int d = psim->ndim;
   push_proposal(ctx, proposal, err);                       forwardError(*err, __LINE__, 0);
   ensure_dev(ctx, n, d, 0, err);                           forwardError(*err, __LINE__, 0);
   {
      double lo[PMCB200_MAX_DIM], hi[PMCB200_MAX_DIM];
      for (int j = 0; j < d; j++) { lo[j] = pb ? pb->min[j] : -HUGE_VAL; hi[j] = pb ? pb->max[j] : HUGE_VAL; }
      B200_OK(ctx, pmcb200_set_box(ctx, d, lo, hi), pmc_incompat, 0);
   }
   uint64_t seed = r ? r->seed : 0;
   uint32_t stream = r ? r->stream++ : 0;
   B200_OK(ctx, pmcb200_simulate(ctx, n, seed, stream, 0, (double *)g_dev.X, (int32_t *)g_dev.idx, (int16_t *)g_dev.flg),
           pmc_undef, 0);
   int64_t nok = 0;
   B200_OK(ctx, pmcb200_read_counts(ctx, &nok, NULL, NULL), pmc_badComm, 0);
   B200_OK(ctx, pmcb200_d2h(ctx, psim->X, g_dev.X, sizeof(double) * (size_t)n * d), pmc_badComm, 0);
   B200_OK(ctx, pmcb200_d2h(ctx, psim->flg, g_dev.flg, sizeof(short) * (size_t)n), pmc_badComm, 0);
   {
      int32_t *t = (int32_t *)malloc_err(sizeof(int32_t) * (size_t)n, err);
      forwardError(*err, __LINE__, 0);
      int rc = pmcb200_d2h(ctx, t, g_dev.idx, sizeof(int32_t) * (size_t)n);
      for (long i = 0; i < n; i++) psim->indices[i] = (size_t)t[i];
      free(t);
      B200_OK(ctx, rc, pmc_badComm, 0);
   }
   psim->isLog = 0;
   return (size_t)nok;
}

/* ---- generic_get_importance_weight_and_deduced_verb (cosmo_pmc.c:343-345) ------------ */
size_t generic_get_importance_weight_and_deduced_verb(pmc_simu *psim, const void *proposal_data,
          posterior_log_pdf_func *proposal_log_pdf, posterior_log_pdf_func *posterior_log_pdf,
          retrieve_ded_func *retrieve_ded, void *target_data, double beta, int quiet, error **err)
{
   (void)retrieve_ded; (void)quiet;
   pmcb200_ctx *ctx = pmc_b200_context(err);
   forwardError(*err, __LINE__, 0);
   testErrorRet(proposal_log_pdf != mix_mvdens_log_pdf_void, pmc_undef,
                "Only mix_mvdens_log_pdf_void proposals have a device path", *err, __LINE__, 0);
   testErrorRet(psim->n_ded > 0, pmc_undef, "Deduced parameters (n_ded > 0) are not supported on the device path",
                *err, __LINE__, 0);
   mix_mvdens *proposal = (mix_mvdens *)proposal_data;
   long n = psim->nsamples;
   push_proposal(ctx, proposal, err);                              forwardError(*err, __LINE__, 0);
   activate_target(ctx, posterior_log_pdf, target_data, err);      forwardError(*err, __LINE__, 0);
   ensure_dev(ctx, n, psim->ndim, 0, err);                         forwardError(*err, __LINE__, 0);
   push_samples(ctx, psim, 0, 0, err);                             forwardError(*err, __LINE__, 0);
   B200_OK(ctx, pmcb200_importance_weights(ctx, n, (double *)g_dev.X, beta, (int16_t *)g_dev.flg, (double *)g_dev.w),
           pmc_undef, 0);
   int64_t nok = 0; double maxW = 0.0;
   B200_OK(ctx, pmcb200_read_counts(ctx, NULL, &nok, &maxW), pmc_badComm, 0);
   B200_OK(ctx, pmcb200_d2h(ctx, psim->weights, g_dev.w, sizeof(double) * (size_t)n), pmc_badComm, 0);
   B200_OK(ctx, pmcb200_d2h(ctx, psim->flg, g_dev.flg, sizeof(short) * (size_t)n), pmc_badComm, 0);
   psim->isLog = 1;
   psim->maxW = maxW;
   return (size_t)nok;
}

size_t generic_get_importance_weight_and_deduced(pmc_simu *psim, const void *proposal_data,
          posterior_log_pdf_func *proposal_log_pdf, posterior_log_pdf_func *posterior_log_pdf,
          retrieve_ded_func *retrieve_ded, void *target_data, error **err)
{
   return generic_get_importance_weight_and_deduced_verb(psim, proposal_data, proposal_log_pdf, posterior_log_pdf,
                                                         retrieve_ded, target_data, 1.0, 1, err);
}

/* ---- normalize_importance_weight (cosmo_pmc.c:378) --------------------------------------- */
double normalize_importance_weight(pmc_simu *psim, error **err)
{
   pmcb200_ctx *ctx = pmc_b200_context(err);
   forwardError(*err, __LINE__, 0.0);
   testErrorRet(psim->isLog != 1, pmc_isLog, "Weights are not in log form", *err, __LINE__, 0.0);
   long n = psim->nsamples;
   ensure_dev(ctx, n, psim->ndim, 0, err);                         forwardError(*err, __LINE__, 0.0);
   B200_OK(ctx, pmcb200_h2d(ctx, g_dev.flg, psim->flg, sizeof(short) * (size_t)n), pmc_badComm, 0.0);
   B200_OK(ctx, pmcb200_h2d(ctx, g_dev.w, psim->weights, sizeof(double) * (size_t)n), pmc_badComm, 0.0);
   double sum = 0.0, logSum = 0.0, maxW = 0.0;
   int rc = pmcb200_normalize_log_weights(ctx, n, (int16_t *)g_dev.flg, (double *)g_dev.w, &sum, &logSum, &maxW);
   testErrorRetVA(rc == PMCB200_ERR_NOSAMPLE, pmc_nosamplep, "%s", *err, __LINE__, 0.0, pmcb200_last_error(ctx));
   B200_OK(ctx, rc, pmc_undef, 0.0);
   B200_OK(ctx, pmcb200_d2h(ctx, psim->weights, g_dev.w, sizeof(double) * (size_t)n), pmc_badComm, 0.0);
   psim->isLog = 0; psim->logSum = logSum; psim->maxW = maxW;
   return sum;
}

/* ---- update_prop_rb (cosmo_pmc.c:247) ------------------------------------------------------ */
void update_prop_rb(mix_mvdens *proposal, pmc_simu *psim, error **err)
{
   pmcb200_ctx *ctx = pmc_b200_context(err);
   forwardError(*err, __LINE__, );
   testErrorRet(psim->isLog != 0, pmc_isLog, "update_prop_rb needs normalised (non-log) weights", *err, __LINE__, );
   long n = psim->nsamples;
   push_proposal(ctx, proposal, err);                              forwardError(*err, __LINE__, );
   long blen = (long)pmcb200_stat_block_len(ctx);
   ensure_dev(ctx, n, psim->ndim, blen, err);                      forwardError(*err, __LINE__, );
   push_samples(ctx, psim, 1, 1, err);                             forwardError(*err, __LINE__, );
   B200_OK(ctx, pmcb200_em_local_linear(ctx, n, (double *)g_dev.X, (int32_t *)g_dev.idx, (int16_t *)g_dev.flg,
                                        (double *)g_dev.w, (double *)g_dev.block), pmc_undef, );
   pmcb200_stats_t st;
   int rc = pmcb200_em_finish(ctx, 1, (double *)g_dev.block, n, &st);
   testErrorRetVA(rc == PMCB200_ERR_NOSAMPLE, pmc_nosamplep, "%s", *err, __LINE__, , pmcb200_last_error(ctx));
   B200_OK(ctx, rc, pmc_undef, );
   pull_proposal(ctx, proposal, err);
   forwardError(*err, __LINE__, );
}
void update_prop_rb_void(void *proposal, pmc_simu *psim, error **err) { update_prop_rb((mix_mvdens *)proposal, psim, err); }

/* ---- diagnostics ------------------------------------------------------------------------------ */
static void weight_stats(pmc_simu *psim, double o[8], error **err)
{
   pmcb200_ctx *ctx = pmc_b200_context(err);
   forwardError(*err, __LINE__, );
   long n = psim->nsamples;
   ensure_dev(ctx, n, psim->ndim, 0, err);                         forwardError(*err, __LINE__, );
   B200_OK(ctx, pmcb200_h2d(ctx, g_dev.flg, psim->flg, sizeof(short) * (size_t)n), pmc_badComm, );
   B200_OK(ctx, pmcb200_h2d(ctx, g_dev.w, psim->weights, sizeof(double) * (size_t)n), pmc_badComm, );
   B200_OK(ctx, pmcb200_weight_stats(ctx, n, (int16_t *)g_dev.flg, (double *)g_dev.w, psim->isLog, o), pmc_undef, );
}

/* perplexity = exp(-sum wbar log wbar)/N, ESS = 1/sum wbar^2 (manual.tex:555-590) */
double perplexity_and_ess(pmc_simu *psim, int normalize, double *ess, error **err)
{
   (void)normalize;      /* the statistics are scale-free: wbar = w/S is applied analytically */
   double o[8];
   weight_stats(psim, o, err);
   forwardError(*err, __LINE__, 0.0);
   double S = o[1], S2 = o[2], T = o[3];
   testErrorRet(!(S > 0.0), pmc_negWeight, "Sum of weights is not positive", *err, __LINE__, 0.0);
   if (ess) *ess = S * S / S2;
   return exp(log(S) - T / S) / (double)psim->nsamples;
}

/* E = (1/N) sum_n w_n over ALL N draws (cf. bin/evidence.pl:16-28) */
double evidence(pmc_simu *psim, double *ln_evi, error **err)
{
   double le;
   if (psim->isLog) {
      double o[8];
      weight_stats(psim, o, err);
      forwardError(*err, __LINE__, 0.0);
      le = log(o[1]) + o[0] - log((double)psim->nsamples);
   } else {
      le = psim->logSum - log((double)psim->nsamples);
   }
   if (ln_evi) *ln_evi = le;
   return exp(le);
}

/* zero the nclipw largest weights (manual.tex:639-641).  nclipw is a handful
 * (Demo/.../WMAP_Distance_Priors/config_pmc:34), so this is nclipw scans. */
void clip_weights(pmc_simu *psim, int nclipw, FILE *OUT, error **err)
{
   testErrorRet(psim->isLog != 0, pmc_isLog, "clip_weights needs normalised weights", *err, __LINE__, );
   for (int c = 0; c < nclipw; c++) {
      long imax = -1; double wmax = 0.0;
      for (long i = 0; i < psim->nsamples; i++)
         if (psim->flg[i] && psim->weights[i] > wmax) { wmax = psim->weights[i]; imax = i; }
      if (imax < 0) break;
      if (OUT) fprintf(OUT, "Clipping point %ld with weight %g\n", imax, wmax);
      psim->weights[imax] = 0.0; psim->flg[imax] = 0;
   }
   double s = 0.0;
   for (long i = 0; i < psim->nsamples; i++) if (psim->flg[i]) s += psim->weights[i];
   testErrorRet(!(s > 0.0), pmc_negWeight, "All weights clipped", *err, __LINE__, );
   for (long i = 0; i < psim->nsamples; i++) psim->weights[i] = psim->flg[i] ? psim->weights[i] / s : 0.0;
   psim->logSum += log(s);
}

double mean_from_psim(const double *X, const double *weights, const short *flg, long nsamples, int ndim, int a)
{
   double m = 0.0, s = 0.0;
   for (long i = 0; i < nsamples; i++)
      if (flg[i]) { m += weights[i] * X[i * ndim + a]; s += weights[i]; }
   return s > 0.0 ? m / s : 0.0;
}

void estimate_param_covar_weight(size_t ndim, size_t nsamples, size_t nskip, const double *X, const double *weight,
                                 double *pmean, double *pvar, error **err)
{
   double s = 0.0;
   memset(pmean, 0, ndim * sizeof(double));
   memset(pvar, 0, ndim * ndim * sizeof(double));
   for (size_t i = nskip; i < nsamples; i++) {
      s += weight[i];
      for (size_t a = 0; a < ndim; a++) pmean[a] += weight[i] * X[i * ndim + a];
   }
   testErrorRet(!(s > 0.0), pmc_negWeight, "Sum of weights is not positive", *err, __LINE__, );
   for (size_t a = 0; a < ndim; a++) pmean[a] /= s;
   for (size_t i = nskip; i < nsamples; i++)
      for (size_t a = 0; a < ndim; a++)
         for (size_t b = 0; b <= a; b++)
            pvar[a * ndim + b] += weight[i] * (X[i * ndim + a] - pmean[a]) * (X[i * ndim + b] - pmean[b]);
   for (size_t a = 0; a < ndim; a++)
      for (size_t b = 0; b <= a; b++) pvar[b * ndim + a] = pvar[a * ndim + b] = pvar[a * ndim + b] / s;
}

/* ---- pmcsim reader (restart path, cosmo_pmc.c:404-439; format exec_helper.c:351-424) ------- */
pmc_simu *pmc_simu_from_file(FILE *F, int nsamples, int npar, int n_ded, mix_mvdens *proposal, int nclipw, error **err)
{
   (void)proposal;
   pmc_simu *p = pmc_simu_init_plus_ded(nsamples, npar, n_ded, err);
   forwardError(*err, __LINE__, NULL);
   char line[16384];
   long n = 0;
   double maxW = -HUGE_VAL;
   while (fgets(line, sizeof(line), F)) {
      char *s = line;
      while (*s == ' ' || *s == '\t') s++;
      if (*s == '#' || *s == '\n' || *s == 0) continue;
      if (n >= nsamples) break;
      char *end;
      double lw = strtod(s, &end);  s = end;
      double comp = strtod(s, &end); s = end;
      for (int j = 0; j < npar; j++) { p->X[n * npar + j] = strtod(s, &end); s = end; }
      for (int j = 0; j < n_ded; j++) { p->X_ded[n * n_ded + j] = strtod(s, &end); s = end; }
      p->weights[n] = lw; p->indices[n] = (size_t)(-comp + 0.5); p->flg[n] = 1;
      if (lw > maxW) maxW = lw;
      n++;
   }
   testErrorRet(n == 0, pmc_nosamplep, "No sample point in pmcsim file", *err, __LINE__, NULL);
   /* nsamples stays the number of DRAWS (the file holds the flagged points only; the tail has flg = 0),
      so perplexity / evidence keep their denominators across a restart */
   for (long i = n; i < nsamples; i++) { p->flg[i] = 0; p->weights[i] = 0.0; p->indices[i] = 0; }
   p->isLog = 1; p->maxW = maxW;
   normalize_importance_weight(p, err);
   forwardError(*err, __LINE__, NULL);
   if (nclipw > 0) { clip_weights(p, nclipw, NULL, err); forwardError(*err, __LINE__, NULL); }
   return p;
}

/* ---- whole iteration in one call ----------------------------------------- */
size_t pmc_b200_iteration(pmc_simu *psim, mix_mvdens *proposal, gsl_rng *r, double beta,
                          pmcb200_stats_t *stats, error **err)
{
   pmcb200_ctx *ctx = pmc_b200_context(err);
   forwardError(*err, __LINE__, 0);
   testErrorRet(g_active_target == NULL, pmc_undef, "Activate a device target first (a weight call, or "
                "pmc_b200_register_target + generic_get_importance_weight...)", *err, __LINE__, 0);
   long n = psim->nsamples;
   push_proposal(ctx, proposal, err);                              forwardError(*err, __LINE__, 0);
   int32_t *idx = (int32_t *)malloc_err(sizeof(int32_t) * (size_t)n, err);
   forwardError(*err, __LINE__, 0);
   pmcb200_stats_t st;
   int rc = pmcb200_iteration_host(ctx, n, r ? r->seed : 0, r ? r->stream++ : 0, beta, psim->X, idx,
                                   (int16_t *)psim->flg, psim->weights, &st);
   for (long i = 0; i < n; i++) psim->indices[i] = (size_t)idx[i];
   free(idx);
   testErrorRetVA(rc == PMCB200_ERR_NOSAMPLE, pmc_nosamplep, "%s", *err, __LINE__, 0, pmcb200_last_error(ctx));
   B200_OK(ctx, rc, pmc_undef, 0);
   psim->isLog = 0; psim->logSum = st.logSum; psim->maxW = st.maxW;
   pull_proposal(ctx, proposal, err);
   forwardError(*err, __LINE__, 0);
   if (stats) *stats = st;
   return (size_t)st.nok;
}
