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

/* The lump is page-locked (when a device exists) so that the slices of several shards move
 * concurrently and the device-to-host copies overlap the likelihood kernel. */
static void *psim_lump(size_t bytes, error **err)
{
   void *b = NULL;
   testErrorRetVA(pmcb200_host_alloc(bytes, &b) != 0 || !b, pmc_allocate, "Cannot allocate %zu bytes for a pmc_simu", *err,
                  __LINE__, NULL, bytes);
   memset(b, 0, bytes);
   return b;
}

pmc_simu *pmc_simu_init_plus_ded(long nsamples, int ndim, int n_ded, error **err)
{
   testErrorRetVA(nsamples < 1 || ndim < 1 || n_ded < 0, pmc_dimension, "Invalid pmc_simu size (%ld,%d,%d)", *err,
                  __LINE__, NULL, nsamples, ndim, n_ded);
   pmc_simu *p = (pmc_simu *)calloc_err(1, sizeof(pmc_simu), err);
   forwardError(*err, __LINE__, NULL);
   p->nsamples = p->nsamples_alloc = nsamples; p->ndim = ndim; p->n_ded = n_ded;
   p->buf = psim_lump(psim_bytes(nsamples, ndim, n_ded), err);
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
   void *nb = psim_lump(psim_bytes(nsamples, p->ndim, p->n_ded), err);     /* the old lump survives a failed allocation */
   forwardError(*err, __LINE__, );
   pmc_b200_invalidate_mirror(p);
   pmcb200_host_free(p->buf);
   p->buf = nb;
   p->nsamples = p->nsamples_alloc = nsamples;
   psim_carve(p, nsamples);
   p->isLog = 0;
}

void pmc_simu_free(pmc_simu **p)
{
   if (!p || !*p) return;
   pmc_b200_invalidate_mirror(*p);
   pmcb200_host_free((*p)->buf); free(*p); *p = NULL;
}

/* ---- device contexts (shards) + target registry -------------------------------------
 * One process drives G >= 1 GPUs: shard r of G owns the contiguous sample range
 * [r*ceil(n/G), ...) of every psim (the reference's MPI ranks, cosmo_pmc.c:323-376, become
 * contexts of one process).  G = $PMCB200_NGPU ("all" = every visible device; default 1);
 * $PMCB200_DEVICES = comma-separated ordinals, used round-robin (an ordinal may repeat:
 * several shards on one device, which is how the sharded path is tested on a one-GPU box);
 * default list = $PMCB200_DEVICE (or 0), +1, ... modulo the device count.
 * Like the reference (single-threaded, no locks anywhere) this layer keeps
 * process-wide state: it is not re-entrant. */
#define MAX_TARGETS 8
#define MAX_SHARDS 64
static int g_ns = 0;
static pmcb200_ctx *g_ctxs[MAX_SHARDS];
static struct { posterior_log_pdf_func *f; void *data; pmcb200_target_t t; int used; } g_targets[MAX_TARGETS];
static const void *g_active_target = NULL;
/* device mirrors of each shard's slice of the last psim (grow-only) */
typedef struct { void *X, *idx, *flg, *w, *block, *all; long cap; int d; long blen; } dev_mirror;
static dev_mirror g_devs[MAX_SHARDS];
/* page-locked staging of pmc_simu->indices (size_t on the host, int32 on the device) */
static int32_t *g_idx32 = NULL;
static long g_idx32_cap = 0;

/* ---- what the device mirrors hold --------------------------------------------
 * The reference calls simulate / weights / normalise / update one after the other on the same psim and never
 * writes the sample arrays in between (exec/cosmo_pmc.c:320-399 only pokes psim->nsamples).  Round 1 re-uploaded
 * every array on every call (2 x 400 MB of X per iteration at 1e7 samples).  Now each entry point records which
 * arrays it left identical on host and device, with a fingerprint (MIRROR_NFP sampled elements + size) of the host
 * copy; the next entry point skips the upload of an array whose record is intact AND whose fingerprint still
 * matches the host array.  A caller that rewrites an array wholesale (pmc_simu_from_file, a copy, a new draw) is
 * caught by the fingerprint; one that edits single elements must call pmc_b200_invalidate_mirror(psim).
 * PMCB200_ALWAYS_UPLOAD=1 restores the unconditional uploads. */
#define MV_X 1u
#define MV_IDX 2u
#define MV_FLG 4u
#define MV_W 8u
#define MIRROR_NFP 256
static struct {
   const pmc_simu *owner; long n; int d, ns; unsigned valid;
   double fpX[MIRROR_NFP], fpW[MIRROR_NFP]; short fpF[MIRROR_NFP]; size_t fpI[MIRROR_NFP];
} g_mir;
static long g_mir_skipped_bytes = 0, g_mir_uploaded_bytes = 0;
static inline long fp_pos(long j, long n) { return (long)(((unsigned long long)j * 2654435761ull + (j == 1 ? (unsigned long long)(n - 1) : 0ull)) % (unsigned long long)n); }
static void mirror_mark(const pmc_simu *p, unsigned bits)
{
   if (g_mir.owner != p || g_mir.n != p->nsamples || g_mir.d != p->ndim || g_mir.ns != g_ns) {
      g_mir.owner = p; g_mir.n = p->nsamples; g_mir.d = p->ndim; g_mir.ns = g_ns; g_mir.valid = 0;
   }
   long n = p->nsamples;
   size_t nx = (size_t)n * p->ndim;
   for (long j = 0; j < MIRROR_NFP && n > 0; j++) {
      long i = fp_pos(j, n);
      if (bits & MV_X) g_mir.fpX[j] = p->X[fp_pos(j, (long)nx)];
      if (bits & MV_W) g_mir.fpW[j] = p->weights[i];
      if (bits & MV_FLG) g_mir.fpF[j] = p->flg[i];
      if (bits & MV_IDX) g_mir.fpI[j] = p->indices[i];
   }
   g_mir.valid |= bits;
}
static void mirror_clear(const pmc_simu *p, unsigned bits) { if (g_mir.owner == p) g_mir.valid &= ~bits; }
void pmc_b200_invalidate_mirror(const pmc_simu *p) { if (!p || g_mir.owner == p) g_mir.valid = 0; }
/* arrays whose device mirror may be trusted for this call */
static unsigned mirror_fresh(const pmc_simu *p)
{
   static int always = -1;
   if (always < 0) { const char *e = getenv("PMCB200_ALWAYS_UPLOAD"); always = e && *e && *e != '0'; }
   if (always || g_mir.owner != p || g_mir.n != p->nsamples || g_mir.d != p->ndim || g_mir.ns != g_ns) return 0;
   unsigned ok = g_mir.valid;
   long n = p->nsamples;
   size_t nx = (size_t)n * p->ndim;
   for (long j = 0; j < MIRROR_NFP && n > 0 && ok; j++) {
      long i = fp_pos(j, n);
      /* bit patterns, not values: NaN == NaN here */
      if ((ok & MV_X) && memcmp(&g_mir.fpX[j], &p->X[fp_pos(j, (long)nx)], sizeof(double)) != 0) ok &= ~MV_X;
      if ((ok & MV_W) && memcmp(&g_mir.fpW[j], &p->weights[i], sizeof(double)) != 0) ok &= ~MV_W;
      if ((ok & MV_FLG) && g_mir.fpF[j] != p->flg[i]) ok &= ~MV_FLG;
      if ((ok & MV_IDX) && g_mir.fpI[j] != p->indices[i]) ok &= ~MV_IDX;
   }
   g_mir.valid = ok;
   return ok;
}
void pmc_b200_mirror_traffic(long *uploaded_bytes, long *skipped_bytes)
{
   if (uploaded_bytes) *uploaded_bytes = g_mir_uploaded_bytes;
   if (skipped_bytes) *skipped_bytes = g_mir_skipped_bytes;
}

/* ---- PMCB200_TIMING=1: wall-clock per pmclib-named entry point, printed to stderr when the process exits
 * (how INTEGRATION.md's per-phase table of the unchanged driver at 10^6 - 10^7 samples was taken) ---------------- */
#include <time.h>
enum { TM_SIMULATE, TM_WEIGHTS, TM_NORMALIZE, TM_UPDATE, TM_STATS, TM_CLIP, TM_N };
static const char *tm_name[TM_N] = { "simulate_mix_mvdens", "generic_get_importance_weight", "normalize_importance_weight",
                                     "update_prop_rb", "perplexity/ess/evidence", "clip_weights" };
static double tm_sec[TM_N]; static long tm_calls[TM_N]; static int tm_on = -1;
static double tm_now(void) { struct timespec t; clock_gettime(CLOCK_MONOTONIC, &t); return t.tv_sec + 1e-9 * t.tv_nsec; }
static void tm_report(void)
{
   long up = g_mir_uploaded_bytes, sk = g_mir_skipped_bytes;
   fprintf(stderr, "[pmcb200 timing] %-32s %8s %12s\n", "entry point", "calls", "seconds");
   for (int i = 0; i < TM_N; i++)
      if (tm_calls[i]) fprintf(stderr, "[pmcb200 timing] %-32s %8ld %12.4f\n", tm_name[i], tm_calls[i], tm_sec[i]);
   fprintf(stderr, "[pmcb200 timing] host->device sample traffic: %.1f MB uploaded, %.1f MB avoided by the mirror bookkeeping\n",
           up / 1e6, sk / 1e6);
}
static double tm_begin(void)
{
   if (tm_on < 0) { const char *e = getenv("PMCB200_TIMING"); tm_on = e && *e && *e != '0'; if (tm_on) atexit(tm_report); }
   return tm_on ? tm_now() : 0.0;
}
static void tm_end(int which, double t0) { if (tm_on > 0) { tm_sec[which] += tm_now() - t0; tm_calls[which]++; } }

#define B200_OK(ctx, call, errcode, ret)                                                          \
   do { int rc__ = (call);                                                                        \
        if (rc__ != 0) { *err = addErrorVA((errcode), "%s (pmcb200 code %d)", *err, __LINE__,     \
                                           pmcb200_last_error(ctx), rc__); return ret; } } while (0)

pmcb200_ctx *pmc_b200_context(error **err)
{
   if (g_ns > 0) return g_ctxs[0];
   int ndev = pmcb200_device_count();
   if (ndev < 1) {
      *err = addError(pmc_undef, "No usable CUDA device; the PMC iteration has no CPU path", *err, __LINE__);
      return NULL;
   }
   int ns = 1;
   const char *e = getenv("PMCB200_NGPU");
   if (e && *e) ns = (strcmp(e, "all") == 0) ? ndev : atoi(e);
   testErrorRetVA(ns < 1 || ns > MAX_SHARDS, pmc_undef, "PMCB200_NGPU = %s (1..%d or 'all')", *err, __LINE__, NULL, e, MAX_SHARDS);
   int list[MAX_SHARDS], nl = 0;
   const char *dl = getenv("PMCB200_DEVICES");
   if (dl && *dl) {
      const char *q = dl;
      while (*q && nl < MAX_SHARDS) {
         char *end;
         long v = strtol(q, &end, 10);
         if (end == q) break;
         list[nl++] = (int)v;
         q = (*end == ',') ? end + 1 : end;
      }
   }
   if (nl == 0) {
      const char *d0 = getenv("PMCB200_DEVICE");
      int first = d0 ? atoi(d0) : 0;
      for (int r = 0; r < ns; r++) list[nl++] = (first + r) % ndev;
   }
   for (int r = 0; r < ns; r++) {
      int dev = list[r % nl];
      int rc = pmcb200_create(dev, NULL, &g_ctxs[r]);
      if (rc != 0) {
         for (int q = 0; q < r; q++) { pmcb200_destroy(g_ctxs[q]); g_ctxs[q] = NULL; }
         *err = addErrorVA(pmc_undef, "No usable CUDA device %d (pmcb200 code %d); the PMC iteration has no CPU path",
                           *err, __LINE__, dev, rc);
         return NULL;
      }
   }
   g_ns = ns;
   return g_ctxs[0];
}

int pmc_b200_nshards(void) { return g_ns; }

void pmc_b200_shutdown(void)
{
   for (int r = 0; r < g_ns; r++) {
      dev_mirror *m = &g_devs[r];
      void *bufs[6] = { m->X, m->idx, m->flg, m->w, m->block, m->all };
      for (int i = 0; i < 6; i++) if (bufs[i]) pmcb200_dev_free(g_ctxs[r], bufs[i]);
      pmcb200_destroy(g_ctxs[r]);
      g_ctxs[r] = NULL;
   }
   memset(g_devs, 0, sizeof(g_devs));
   if (g_idx32) pmcb200_host_free(g_idx32);
   g_idx32 = NULL; g_idx32_cap = 0;
   g_ns = 0; g_active_target = NULL;
   memset(&g_mir, 0, sizeof(g_mir));
}

/* slice of an n-sample psim owned by shard r (same rule as pmcb200_iteration_host_multi) */
static void shard_range(long n, int r, long *off, long *nr)
{
   long per = (n + g_ns - 1) / g_ns;
   long o = (long)r * per;
   if (o > n) o = n;
   long m = n - o;
   if (m > per) m = per;
   *off = o; *nr = m;
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

static void activate_target(posterior_log_pdf_func *f, void *data, error **err)
{
   for (int i = 0; i < MAX_TARGETS; i++)
      if (g_targets[i].used && g_targets[i].f == f && g_targets[i].data == data) {
         if (g_active_target != &g_targets[i].t) {
            for (int r = 0; r < g_ns; r++)
               B200_OK(g_ctxs[r], pmcb200_set_target(g_ctxs[r], &g_targets[i].t), pmc_incompat, );
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
         activate_target(f, data, err);
         forwardError(*err, __LINE__, );
         return;
      }
   }
   *err = addError(pmc_undef, "No device target registered for this posterior callback "
                   "(pmc_b200_register_target); the scalar host callback cannot be batched and there is no CPU path",
                   *err, __LINE__);
}

/* mirrors of every shard's slice of an n-sample, d-dimensional psim; blen > 0 also sizes the
 * statistics block and the gathered blocks of all shards */
static void ensure_dev(long n, int d, long blen, error **err)
{
   for (int r = 0; r < g_ns; r++) {
      pmcb200_ctx *ctx = g_ctxs[r];
      dev_mirror *m = &g_devs[r];
      long off, nr;
      shard_range(n, r, &off, &nr);
      if (nr < 1) nr = 1;
      if (nr > m->cap || d > m->d) {
         g_mir.valid = 0;                           /* the mirrors are being replaced */
         if (m->X) { pmcb200_dev_free(ctx, m->X); pmcb200_dev_free(ctx, m->idx);
                     pmcb200_dev_free(ctx, m->flg); pmcb200_dev_free(ctx, m->w); }
         long cap = nr > m->cap ? nr : m->cap;
         int dd = d > m->d ? d : m->d;
         B200_OK(ctx, pmcb200_dev_alloc(ctx, sizeof(double) * (size_t)cap * dd, &m->X), pmc_allocate, );
         B200_OK(ctx, pmcb200_dev_alloc(ctx, sizeof(int32_t) * (size_t)cap, &m->idx), pmc_allocate, );
         B200_OK(ctx, pmcb200_dev_alloc(ctx, sizeof(int16_t) * (size_t)cap, &m->flg), pmc_allocate, );
         B200_OK(ctx, pmcb200_dev_alloc(ctx, sizeof(double) * (size_t)cap, &m->w), pmc_allocate, );
         m->cap = cap; m->d = dd;
      }
      if (blen > m->blen) {
         if (m->block) { pmcb200_dev_free(ctx, m->block); pmcb200_dev_free(ctx, m->all); }
         B200_OK(ctx, pmcb200_dev_alloc(ctx, sizeof(double) * (size_t)blen, &m->block), pmc_allocate, );
         B200_OK(ctx, pmcb200_dev_alloc(ctx, sizeof(double) * (size_t)blen * g_ns, &m->all), pmc_allocate, );
         m->blen = blen;
      }
   }
}

static int32_t *idx_staging(long n, error **err)
{
   if (n > g_idx32_cap) {
      if (g_idx32) pmcb200_host_free(g_idx32);
      g_idx32 = NULL; g_idx32_cap = 0;
      void *p = NULL;
      testErrorRet(pmcb200_host_alloc(sizeof(int32_t) * (size_t)n, &p) != 0, pmc_allocate, "Cannot allocate the index staging buffer",
                   *err, __LINE__, NULL);
      g_idx32 = (int32_t *)p; g_idx32_cap = n;
   }
   return g_idx32;
}

static void sync_all(error **err)
{
   for (int r = 0; r < g_ns; r++) B200_OK(g_ctxs[r], pmcb200_sync(g_ctxs[r]), pmc_badComm, );
}

/* proposal (Cholesky-decomposed mix_mvdens) -> every shard */
static void push_proposal(mix_mvdens *m, error **err)
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
   for (int r = 0; r < g_ns; r++) {
      int rc = pmcb200_set_proposal(g_ctxs[r], (int)K, (int)d, m->comp[0]->df, m->wght, mean, chol);
      if (rc != 0) {
         *err = addErrorVA(rc == PMCB200_ERR_CHOLESKY ? pmc_cholesky : pmc_incompat, "%s", *err, __LINE__,
                           pmcb200_last_error(g_ctxs[r]));
         break;
      }
   }
   free(mean);
}

/* updated proposal <- device (shard 0; every shard holds the identical update).  As in pmclib,
 * update_prop_rb leaves the COVARIANCE in comp[k]->std (chol = 0; the Cholesky factor is recomputed
 * on demand by the next use): the reference copies std between components as a covariance
 * (revive_comp, exec/cosmo_pmc.c:225).  Components that died in the update (weight 0,
 * cleanup_after_update) keep their previous mean and matrix. */
static void pull_proposal(mix_mvdens *m, error **err)
{
   pmcb200_ctx *ctx = g_ctxs[0];
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

/* every shard's slice of the psim arrays -> its device mirror (queued, not waited for) */
static void push_samples(pmc_simu *p, int with_X, int with_idx, int with_w, error **err)
{
   long n = p->nsamples;
   int d = p->ndim;
   int32_t *t = NULL;
   int with_flg = 1;
   const unsigned fresh = mirror_fresh(p);
   if (with_X && (fresh & MV_X)) { with_X = 0; g_mir_skipped_bytes += (long)sizeof(double) * n * d; }
   if (with_idx && (fresh & MV_IDX)) { with_idx = 0; g_mir_skipped_bytes += (long)sizeof(int32_t) * n; }
   if (with_w && (fresh & MV_W)) { with_w = 0; g_mir_skipped_bytes += (long)sizeof(double) * n; }
   if (fresh & MV_FLG) { with_flg = 0; g_mir_skipped_bytes += (long)sizeof(short) * n; }
   g_mir_uploaded_bytes += (with_X ? (long)sizeof(double) * n * d : 0) + (with_idx ? (long)sizeof(int32_t) * n : 0) +
                           (with_w ? (long)sizeof(double) * n : 0) + (with_flg ? (long)sizeof(short) * n : 0);
   if (with_idx) {
      t = idx_staging(n, err);
      forwardError(*err, __LINE__, );
      for (long i = 0; i < n; i++) t[i] = (int32_t)p->indices[i];
   }
   for (int r = 0; r < g_ns; r++) {
      pmcb200_ctx *ctx = g_ctxs[r];
      dev_mirror *m = &g_devs[r];
      long off, nr;
      shard_range(n, r, &off, &nr);
      if (nr == 0) continue;
      if (with_X) B200_OK(ctx, pmcb200_h2d_async(ctx, m->X, p->X + (size_t)off * d, sizeof(double) * (size_t)nr * d), pmc_badComm, );
      if (with_flg) B200_OK(ctx, pmcb200_h2d_async(ctx, m->flg, p->flg + off, sizeof(short) * (size_t)nr), pmc_badComm, );
      if (with_w) B200_OK(ctx, pmcb200_h2d_async(ctx, m->w, p->weights + off, sizeof(double) * (size_t)nr), pmc_badComm, );
      if (with_idx) B200_OK(ctx, pmcb200_h2d_async(ctx, m->idx, t + off, sizeof(int32_t) * (size_t)nr), pmc_badComm, );
   }
   mirror_mark(p, (with_X ? MV_X : 0u) | (with_idx ? MV_IDX : 0u) | (with_w ? MV_W : 0u) | (with_flg ? MV_FLG : 0u));
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
   const double tm0 = tm_begin();
   pmc_b200_context(err);
   forwardError(*err, __LINE__, 0);
   testErrorRetVA((size_t)psim->ndim != proposal->ndim, pmc_dimension, "psim ndim %d != proposal ndim %zu", *err,
                  __LINE__, 0, psim->ndim, proposal->ndim);
   testErrorRet(pb != NULL && pb->ndim != psim->ndim, pmc_dimension, "parabox dimension mismatch", *err, __LINE__, 0);
   long n = psim->nsamples;
   int d = psim->ndim;
   push_proposal(proposal, err);                            forwardError(*err, __LINE__, 0);
   ensure_dev(n, d, 0, err);                                forwardError(*err, __LINE__, 0);
   int32_t *t = idx_staging(n, err);                        forwardError(*err, __LINE__, 0);
   double lo[PMCB200_MAX_DIM], hi[PMCB200_MAX_DIM];
   for (int j = 0; j < d; j++) { lo[j] = pb ? pb->min[j] : -HUGE_VAL; hi[j] = pb ? pb->max[j] : HUGE_VAL; }
   uint64_t seed = r ? r->seed : 0;
   uint32_t stream = r ? r->stream++ : 0;
   for (int s = 0; s < g_ns; s++) {      /* queue every shard, then wait */
      pmcb200_ctx *ctx = g_ctxs[s];
      dev_mirror *m = &g_devs[s];
      long off, ns;
      shard_range(n, s, &off, &ns);
      B200_OK(ctx, pmcb200_set_box(ctx, d, lo, hi), pmc_incompat, 0);
      B200_OK(ctx, pmcb200_simulate(ctx, ns, seed, stream, off, (double *)m->X, (int32_t *)m->idx, (int16_t *)m->flg),
              pmc_undef, 0);
      if (ns == 0) continue;
      B200_OK(ctx, pmcb200_d2h_async(ctx, psim->X + (size_t)off * d, m->X, sizeof(double) * (size_t)ns * d), pmc_badComm, 0);
      B200_OK(ctx, pmcb200_d2h_async(ctx, psim->flg + off, m->flg, sizeof(short) * (size_t)ns), pmc_badComm, 0);
      B200_OK(ctx, pmcb200_d2h_async(ctx, t + off, m->idx, sizeof(int32_t) * (size_t)ns), pmc_badComm, 0);
   }
   int64_t nok = 0;
   for (int s = 0; s < g_ns; s++) {
      int64_t k = 0;
      B200_OK(g_ctxs[s], pmcb200_read_counts(g_ctxs[s], &k, NULL, NULL), pmc_badComm, 0);    /* waits for the shard */
      nok += k;
   }
   for (long i = 0; i < n; i++) psim->indices[i] = (size_t)t[i];
   psim->isLog = 0;
   tm_end(TM_SIMULATE, tm0);
   pmc_b200_invalidate_mirror(psim);
   mirror_mark(psim, MV_X | MV_IDX | MV_FLG);       /* the draw is identical on host and device; the weights are stale */
   return (size_t)nok;
}

/* ---- generic_get_importance_weight_and_deduced_verb (cosmo_pmc.c:343-345) ------------ */
size_t generic_get_importance_weight_and_deduced_verb(pmc_simu *psim, const void *proposal_data,
          posterior_log_pdf_func *proposal_log_pdf, posterior_log_pdf_func *posterior_log_pdf,
          retrieve_ded_func *retrieve_ded, void *target_data, double beta, int quiet, error **err)
{
   (void)retrieve_ded; (void)quiet;
   const double tm0 = tm_begin();
   pmc_b200_context(err);
   forwardError(*err, __LINE__, 0);
   testErrorRet(proposal_log_pdf != mix_mvdens_log_pdf_void, pmc_undef,
                "Only mix_mvdens_log_pdf_void proposals have a device path", *err, __LINE__, 0);
   testErrorRet(psim->n_ded > 0, pmc_undef, "Deduced parameters (n_ded > 0) are not supported on the device path",
                *err, __LINE__, 0);
   mix_mvdens *proposal = (mix_mvdens *)proposal_data;
   long n = psim->nsamples;
   push_proposal(proposal, err);                                   forwardError(*err, __LINE__, 0);
   activate_target(posterior_log_pdf, target_data, err);           forwardError(*err, __LINE__, 0);
   ensure_dev(n, psim->ndim, 0, err);                              forwardError(*err, __LINE__, 0);
   push_samples(psim, 1, 0, 0, err);                               forwardError(*err, __LINE__, 0);
   for (int s = 0; s < g_ns; s++) {
      pmcb200_ctx *ctx = g_ctxs[s];
      dev_mirror *m = &g_devs[s];
      long off, ns;
      shard_range(n, s, &off, &ns);
      B200_OK(ctx, pmcb200_importance_weights(ctx, ns, (double *)m->X, beta, (int16_t *)m->flg, (double *)m->w), pmc_undef, 0);
      if (ns == 0) continue;
      B200_OK(ctx, pmcb200_d2h_async(ctx, psim->weights + off, m->w, sizeof(double) * (size_t)ns), pmc_badComm, 0);
      B200_OK(ctx, pmcb200_d2h_async(ctx, psim->flg + off, m->flg, sizeof(short) * (size_t)ns), pmc_badComm, 0);
   }
   int64_t nok = 0;
   double maxW = -HUGE_VAL;
   for (int s = 0; s < g_ns; s++) {
      int64_t k = 0; double mw = -HUGE_VAL;
      B200_OK(g_ctxs[s], pmcb200_read_counts(g_ctxs[s], NULL, &k, &mw), pmc_badComm, 0);
      nok += k;
      if (mw > maxW) maxW = mw;
   }
   psim->isLog = 1;
   psim->maxW = maxW;
   tm_end(TM_WEIGHTS, tm0);
   mirror_mark(psim, MV_W | MV_FLG);                /* log weights and the cleared flags, both sides */
   return (size_t)nok;
}

size_t generic_get_importance_weight_and_deduced(pmc_simu *psim, const void *proposal_data,
          posterior_log_pdf_func *proposal_log_pdf, posterior_log_pdf_func *posterior_log_pdf,
          retrieve_ded_func *retrieve_ded, void *target_data, error **err)
{
   return generic_get_importance_weight_and_deduced_verb(psim, proposal_data, proposal_log_pdf, posterior_log_pdf,
                                                         retrieve_ded, target_data, 1.0, 1, err);
}

/* ---- importance_sample (exec/importance_sample.c:26-91), batched ---------------------------
 * The reference re-weights a stored sample under a second posterior with one host callback
 * per point.  Here every point of psim goes through the device posterior in one launch per
 * shard: weights[i] <- log pi(x_i), maxW, isLog = 1.  flg[i] = 1 for every evaluated point as in
 * the reference, except that a point whose posterior raises an error or is not finite gets
 * flg = 0 (the reference purges the error and keeps the callback's dummy return value; dropping
 * the point is its documented policy, manual.tex:507-512).  Returns the number of flagged points. */
size_t pmc_b200_importance_sample(pmc_simu *psim, posterior_log_pdf_func *posterior_log_pdf, void *target_data,
                                  error **err)
{
   pmc_b200_context(err);
   forwardError(*err, __LINE__, 0);
   testErrorRet(psim->n_ded > 0, pmc_undef, "Deduced parameters (n_ded > 0) are not supported on the device path",
                *err, __LINE__, 0);
   long n = psim->nsamples;
   activate_target(posterior_log_pdf, target_data, err);           forwardError(*err, __LINE__, 0);
   ensure_dev(n, psim->ndim, 0, err);                              forwardError(*err, __LINE__, 0);
   int32_t *e32 = idx_staging(n, err);                             forwardError(*err, __LINE__, 0);
   push_samples(psim, 1, 0, 0, err);                               forwardError(*err, __LINE__, 0);
   for (int s = 0; s < g_ns; s++) {
      pmcb200_ctx *ctx = g_ctxs[s];
      dev_mirror *m = &g_devs[s];
      long off, ns;
      shard_range(n, s, &off, &ns);
      if (ns == 0) continue;
      B200_OK(ctx, pmcb200_posterior_log_pdf(ctx, ns, (double *)m->X, (double *)m->w, (int32_t *)m->idx), pmc_undef, 0);
      B200_OK(ctx, pmcb200_d2h_async(ctx, psim->weights + off, m->w, sizeof(double) * (size_t)ns), pmc_badComm, 0);
      B200_OK(ctx, pmcb200_d2h_async(ctx, e32 + off, m->idx, sizeof(int32_t) * (size_t)ns), pmc_badComm, 0);
   }
   sync_all(err);                                                  forwardError(*err, __LINE__, 0);
   size_t count = 0;
   double MW = -1.0e30;
   for (long i = 0; i < n; i++) {
      int ok = (e32[i] == 0) && isfinite(psim->weights[i]);
      psim->flg[i] = (short)ok;
      if (!ok) { psim->weights[i] = 0.0; continue; }
      if (count == 0 || psim->weights[i] > MW) MW = psim->weights[i];
      count++;
   }
   psim->maxW = MW;
   psim->isLog = 1;
   mirror_clear(psim, MV_IDX | MV_FLG | MV_W);      /* the index mirror served as the error array; host flags / weights were edited */
   return count;
}

/* {M, S, S2, T, n_flagged} of the weights held in the device mirrors (pmcb200_weight_stats per
 * shard), combined over shards in shard order: log weights are re-based on the global maximum. */
static void weight_stats_dev(long n, int is_log, double o[8], error **err)
{
   double part[MAX_SHARDS][8];
   int have[MAX_SHARDS];
   double M = -HUGE_VAL;
   for (int s = 0; s < g_ns; s++) {
      long off, ns;
      shard_range(n, s, &off, &ns);
      have[s] = 0;
      if (ns == 0) continue;
      B200_OK(g_ctxs[s], pmcb200_weight_stats(g_ctxs[s], ns, (int16_t *)g_devs[s].flg, (double *)g_devs[s].w, is_log, part[s]),
              pmc_undef, );
      if (!(part[s][1] > 0.0)) continue;          /* no flagged sample with a finite weight in this shard */
      have[s] = 1;
      if (part[s][0] > M) M = part[s][0];
   }
   memset(o, 0, 8 * sizeof(double));
   o[0] = is_log ? M : 0.0;
   for (int s = 0; s < g_ns; s++) {
      long off, ns;
      shard_range(n, s, &off, &ns);
      if (ns > 0) o[4] += part[s][4];
      if (!have[s]) continue;
      const double *q = part[s];
      if (is_log) {
         double dM = q[0] - M, f = exp(dM);       /* dM = 0, f = 1 exactly for the shard holding the maximum */
         o[1] += f * q[1];
         o[2] += f * f * q[2];
         o[3] += f * (q[3] + dM * q[1]);
      } else {
         o[1] += q[1]; o[2] += q[2]; o[3] += q[3];
      }
   }
}

/* ---- normalize_importance_weight (cosmo_pmc.c:378) --------------------------------------- */
double normalize_importance_weight(pmc_simu *psim, error **err)
{
   const double tm0 = tm_begin();
   pmc_b200_context(err);
   forwardError(*err, __LINE__, 0.0);
   testErrorRet(psim->isLog != 1, pmc_isLog, "Weights are not in log form", *err, __LINE__, 0.0);
   long n = psim->nsamples;
   ensure_dev(n, psim->ndim, 0, err);                              forwardError(*err, __LINE__, 0.0);
   push_samples(psim, 0, 0, 1, err);                               forwardError(*err, __LINE__, 0.0);
   double o[8];
   weight_stats_dev(n, 1, o, err);                                 forwardError(*err, __LINE__, 0.0);
   testErrorRet(!(o[1] > 0.0), pmc_nosamplep, "normalize: no sample with finite weight", *err, __LINE__, 0.0);
   for (int s = 0; s < g_ns; s++) {
      pmcb200_ctx *ctx = g_ctxs[s];
      long off, ns;
      shard_range(n, s, &off, &ns);
      if (ns == 0) continue;
      B200_OK(ctx, pmcb200_normalize_with(ctx, ns, (int16_t *)g_devs[s].flg, (double *)g_devs[s].w, o[0], o[1]), pmc_undef, 0.0);
      B200_OK(ctx, pmcb200_d2h_async(ctx, psim->weights + off, g_devs[s].w, sizeof(double) * (size_t)ns), pmc_badComm, 0.0);
   }
   sync_all(err);                                                  forwardError(*err, __LINE__, 0.0);
   psim->isLog = 0; psim->logSum = log(o[1]) + o[0]; psim->maxW = o[0];
   tm_end(TM_NORMALIZE, tm0);
   mirror_mark(psim, MV_W);
   return o[1];
}

/* ---- update_prop_rb (cosmo_pmc.c:247) ------------------------------------------------------ */
void update_prop_rb(mix_mvdens *proposal, pmc_simu *psim, error **err)
{
   const double tm0 = tm_begin();
   pmc_b200_context(err);
   forwardError(*err, __LINE__, );
   testErrorRet(psim->isLog != 0, pmc_isLog, "update_prop_rb needs normalised (non-log) weights", *err, __LINE__, );
   long n = psim->nsamples;
   push_proposal(proposal, err);                                   forwardError(*err, __LINE__, );
   long blen = (long)pmcb200_stat_block_len(g_ctxs[0]);
   ensure_dev(n, psim->ndim, blen, err);                           forwardError(*err, __LINE__, );
   push_samples(psim, 1, 1, 1, err);                               forwardError(*err, __LINE__, );
   double *blk[MAX_SHARDS], *all[MAX_SHARDS];
   for (int s = 0; s < g_ns; s++) {
      pmcb200_ctx *ctx = g_ctxs[s];
      dev_mirror *m = &g_devs[s];
      long off, ns;
      shard_range(n, s, &off, &ns);
      blk[s] = (double *)m->block; all[s] = (double *)m->all;
      B200_OK(ctx, pmcb200_em_local_linear(ctx, ns, (double *)m->X, (int32_t *)m->idx, (int16_t *)m->flg,
                                           (double *)m->w, blk[s]), pmc_undef, );
   }
   /* the one exchange per iteration (the reference gathers the weights on the master, cosmo_pmc.c:353-376):
      statistics blocks to every shard by peer copies, then the identical combine + M-step everywhere */
   if (g_ns > 1) B200_OK(g_ctxs[0], pmcb200_allgather_blocks(g_ctxs, g_ns, blk, all, blen), pmc_badComm, );
   for (int s = 0; s < g_ns; s++) {
      pmcb200_stats_t st;
      int rc = pmcb200_em_finish(g_ctxs[s], g_ns, g_ns > 1 ? all[s] : blk[s], n, &st);
      testErrorRetVA(rc == PMCB200_ERR_NOSAMPLE, pmc_nosamplep, "%s", *err, __LINE__, , pmcb200_last_error(g_ctxs[s]));
      B200_OK(g_ctxs[s], rc, pmc_undef, );
   }
   pull_proposal(proposal, err);
   tm_end(TM_UPDATE, tm0);
   forwardError(*err, __LINE__, );
}
void update_prop_rb_void(void *proposal, pmc_simu *psim, error **err) { update_prop_rb((mix_mvdens *)proposal, psim, err); }

/* ---- diagnostics ------------------------------------------------------------------------------ */
static void weight_stats(pmc_simu *psim, double o[8], error **err)
{
   const double tm0 = tm_begin();
   pmc_b200_context(err);
   forwardError(*err, __LINE__, );
   long n = psim->nsamples;
   ensure_dev(n, psim->ndim, 0, err);                              forwardError(*err, __LINE__, );
   push_samples(psim, 0, 0, 1, err);                               forwardError(*err, __LINE__, );
   weight_stats_dev(n, psim->isLog, o, err);                       forwardError(*err, __LINE__, );
   tm_end(TM_STATS, tm0);
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
   const double tm0 = tm_begin();
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
   mirror_clear(psim, MV_W | MV_FLG);               /* host-side edit of weights and flags */
   for (long i = 0; i < psim->nsamples; i++) psim->weights[i] = psim->flg[i] ? psim->weights[i] / s : 0.0;
   psim->logSum += log(s);
   tm_end(TM_CLIP, tm0);
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

/* ---- pmcsim files (restart path, cosmo_pmc.c:404-439; text format exec_helper.c:351-424) -----
 * Text rows are `%16.9g` x (log w_unnormalised, -component, x[0..npar), x_ded[..]) for the flagged
 * points.  At 1e7-1e8 samples the text file is the bottleneck of a restart and keeps only nine
 * digits, so this layer also reads and writes a binary sidecar with the same content in exact
 * doubles:
 *   header  char magic[8] = "PMCSIMB1"; int32 npar, n_ded; int64 nsamples (draws), nrows; double logSum
 *   rows    nrows x { double logw_unnormalised; double component; double x[npar]; double x_ded[n_ded] }
 * pmc_simu_from_file recognises the magic, so a caller that opens either file gets the same psim. */
static const char PMCSIM_MAGIC[8] = { 'P', 'M', 'C', 'S', 'I', 'M', 'B', '1' };
typedef struct { char magic[8]; int32_t npar, n_ded; int64_t nsamples, nrows; double logSum; } pmcsim_bin_header;

void pmc_simu_dump_binary(FILE *F, const pmc_simu *psim, error **err)
{
   pmcsim_bin_header h;
   memcpy(h.magic, PMCSIM_MAGIC, 8);
   h.npar = psim->ndim; h.n_ded = psim->n_ded; h.nsamples = psim->nsamples; h.nrows = 0; h.logSum = psim->logSum;
   for (long i = 0; i < psim->nsamples; i++) h.nrows += psim->flg[i] != 0;
   testErrorRet(fwrite(&h, sizeof(h), 1, F) != 1, io_file, "Cannot write the pmcsim header", *err, __LINE__, );
   int nc = 2 + psim->ndim + psim->n_ded;
   const long CH = 65536;
   double *row = (double *)malloc_err(sizeof(double) * (size_t)nc * CH, err);
   forwardError(*err, __LINE__, );
   long k = 0;
   for (long i = 0; i < psim->nsamples; i++) {
      if (psim->flg[i]) {
         double *q = row + (size_t)k * nc;
         double lw = psim->weights[i];
         q[0] = (psim->isLog ? lw : log(lw)) + psim->logSum;    /* as out_pmc_simu_cosmo_pmc, exec_helper.c:408-420: logSum is always added */
         q[1] = (double)psim->indices[i];
         memcpy(q + 2, psim->X + (size_t)i * psim->ndim, sizeof(double) * psim->ndim);
         if (psim->n_ded) memcpy(q + 2 + psim->ndim, psim->X_ded + (size_t)i * psim->n_ded, sizeof(double) * psim->n_ded);
         k++;
      }
      if (k == CH || (i == psim->nsamples - 1 && k > 0)) {
         if (fwrite(row, sizeof(double) * nc, (size_t)k, F) != (size_t)k) {
            free(row);
            *err = addError(io_file, "Cannot write the pmcsim rows", *err, __LINE__);
            return;
         }
         k = 0;
      }
   }
   free(row);
}

static long read_pmcsim_binary(FILE *F, pmc_simu *p, long nsamples, int npar, int n_ded, double *maxW, error **err)
{
   pmcsim_bin_header h;
   testErrorRet(fread(&h, sizeof(h), 1, F) != 1, io_eof, "Truncated binary pmcsim header", *err, __LINE__, 0);
   testErrorRetVA(h.npar != npar || h.n_ded != n_ded, pmc_dimension, "Binary pmcsim holds npar=%d n_ded=%d, expected %d %d",
                  *err, __LINE__, 0, h.npar, h.n_ded, npar, n_ded);
   int nc = 2 + npar + n_ded;
   long n = h.nrows < nsamples ? (long)h.nrows : nsamples;
   double *row = (double *)malloc_err(sizeof(double) * (size_t)nc * 65536, err);
   forwardError(*err, __LINE__, 0);
   for (long i0 = 0; i0 < n; i0 += 65536) {
      long m = n - i0 < 65536 ? n - i0 : 65536;
      if (fread(row, sizeof(double) * nc, (size_t)m, F) != (size_t)m) {
         free(row);
         *err = addError(io_eof, "Truncated binary pmcsim file", *err, __LINE__);
         return 0;
      }
      for (long k = 0; k < m; k++) {
         const double *q = row + (size_t)k * nc;
         long i = i0 + k;
         p->weights[i] = q[0]; p->indices[i] = (size_t)q[1]; p->flg[i] = 1;
         memcpy(p->X + (size_t)i * npar, q + 2, sizeof(double) * npar);
         if (n_ded) memcpy(p->X_ded + (size_t)i * n_ded, q + 2 + npar, sizeof(double) * n_ded);
         if (q[0] > *maxW) *maxW = q[0];
      }
   }
   free(row);
   return n;
}

pmc_simu *pmc_simu_from_file(FILE *F, int nsamples, int npar, int n_ded, mix_mvdens *proposal, int nclipw, error **err)
{
   (void)proposal;
   pmc_simu *p = pmc_simu_init_plus_ded(nsamples, npar, n_ded, err);
   forwardError(*err, __LINE__, NULL);
   char line[16384];
   long n = 0;
   double maxW = -HUGE_VAL;
   int c0 = fgetc(F);
   if (c0 != EOF) ungetc(c0, F);
   if (c0 == PMCSIM_MAGIC[0]) {          /* a text pmcsim starts with '#', a digit, a sign or blank */
      n = read_pmcsim_binary(F, p, nsamples, npar, n_ded, &maxW, err);
      forwardError(*err, __LINE__, NULL);
   } else while (fgets(line, sizeof(line), F)) {
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
   pmc_b200_context(err);
   forwardError(*err, __LINE__, 0);
   testErrorRet(g_active_target == NULL, pmc_undef, "Activate a device target first (a weight call, or "
                "pmc_b200_register_target + generic_get_importance_weight...)", *err, __LINE__, 0);
   long n = psim->nsamples;
   push_proposal(proposal, err);                                   forwardError(*err, __LINE__, 0);
   int32_t *idx = idx_staging(n, err);                             forwardError(*err, __LINE__, 0);
   pmcb200_stats_t st;
   int rc = pmcb200_iteration_host_multi(g_ctxs, g_ns, n, r ? r->seed : 0, r ? r->stream++ : 0, beta, psim->X, idx,
                                         (int16_t *)psim->flg, psim->weights, &st);
   testErrorRetVA(rc == PMCB200_ERR_NOSAMPLE, pmc_nosamplep, "%s", *err, __LINE__, 0, pmcb200_last_error(g_ctxs[0]));
   B200_OK(g_ctxs[0], rc, pmc_undef, 0);
   for (long i = 0; i < n; i++) psim->indices[i] = (size_t)idx[i];
   psim->isLog = 0; psim->logSum = st.logSum; psim->maxW = st.maxW;
   pmc_b200_invalidate_mirror(psim);               /* the fused call works on library scratch, not on the mirrors */
   pull_proposal(proposal, err);
   forwardError(*err, __LINE__, 0);
   if (stats) *stats = st;
   return (size_t)st.nok;
}
