/* pmclib/pmc.h -- the PMC host API of pmclib as CosmoPMC calls it
 * (exec/cosmo_pmc.c:293-402).  Same names, argument meaning
 * and error behaviour; every batched computation runs on the GPU through
 * include/pmcb200.h.  There is no CPU path for the batched functions: they
 * raise pmc_undef if no CUDA device / no device target is available. */
#ifndef PMCLIB_PMC_H
#define PMCLIB_PMC_H

#include <stdio.h>
#include <stddef.h>
#include "pmctools/errorlist.h"
#include "pmctools/mvdens.h"
#include "pmclib/parabox.h"
#include "pmcb200.h"

#ifdef __cplusplus
extern "C" {
#endif

#define MC_AF_ACCEPT  1
#define MC_AF_REJECT  0
#define MC_NORM       0     /* perplexity_and_ess: normalise the weights first   */
#define MC_UNORM      1     /* weights are already normalised (cosmo_pmc.c:46)   */
#define MINCOUNT      PMCB200_MINCOUNT

/* The sample container.  Field names and meanings are ABI: the reference reads
 * and writes them directly (exec/exec_helper.c:408-423, exec/cosmo_pmc.c:328,372,
 * exec/importance_sample.c:34-40,265-289). */
typedef struct {
  long    nsamples;      /* N */
  int     ndim, n_ded;
  double *X;             /* [N*ndim] */
  double *X_ded;         /* [N*n_ded] deduced parameters */
  double *weights;       /* [N] log w while isLog, normalised w afterwards */
  short  *flg;           /* [N] 1 = usable sample */
  size_t *indices;       /* [N] proposal component the sample was drawn from */
  double  logSum;        /* log sum_n w_n (unnormalised) */
  double  maxW;          /* max log w */
  int     isLog;
  int     mpi_rank, mpi_size;
  long    nsamples_alloc;
  void   *buf;           /* one contiguous lump (exec/add_pmc_proposal.c:139) */
} pmc_simu;

/* pmclib's generic density object.  CosmoPMC's sources use the explicit
 * (pointer + callback) API below and never name this type; it is
 * kept for header compatibility. */
typedef struct _distribution_struct_ {
  int ndim, n_ded;
  void *data;
  posterior_log_pdf_func *log_pdf;
  retrieve_ded_func *retrieve;
  void (*free)(void **);
  long (*simulate)(void *, void *, void *, void *, error **);
  void *broadcast_mpi;
} distribution;

pmc_simu *pmc_simu_init(long nsamples, int ndim, error **err);
pmc_simu *pmc_simu_init_plus_ded(long nsamples, int ndim, int n_ded, error **err);
pmc_simu *pmc_simu_init_mpi(long nsamples, int ndim, int n_ded, error **err);
void      pmc_simu_realloc(pmc_simu *psim, long nsamples, error **err);
void      pmc_simu_free(pmc_simu **psim);
pmc_simu *pmc_simu_from_file(FILE *PMCSIM, int nsamples, int npar, int n_ded, mix_mvdens *proposal,
                             int nclipw, error **err);

/* binary sidecar of the pmcsim text file (same rows in exact doubles; pmc_simu_from_file reads
 * either format): for samples of 1e7-1e8 points, where the %16.9g text file dominates a restart */
void      pmc_simu_dump_binary(FILE *F, const pmc_simu *psim, error **err);

/* ---- the four hot calls of run_pmc_iteration_MPI ----------------------------- */
size_t simulate_mix_mvdens(pmc_simu *psim, mix_mvdens *proposal, gsl_rng *r, parabox *pb, error **err);
size_t generic_get_importance_weight_and_deduced_verb(pmc_simu *psim, const void *proposal_data,
          posterior_log_pdf_func *proposal_log_pdf, posterior_log_pdf_func *posterior_log_pdf,
          retrieve_ded_func *retrieve_ded, void *target_data, double beta, int quiet, error **err);
size_t generic_get_importance_weight_and_deduced(pmc_simu *psim, const void *proposal_data,
          posterior_log_pdf_func *proposal_log_pdf, posterior_log_pdf_func *posterior_log_pdf,
          retrieve_ded_func *retrieve_ded, void *target_data, error **err);
double normalize_importance_weight(pmc_simu *psim, error **err);
void   update_prop_rb(mix_mvdens *proposal, pmc_simu *psim, error **err);
void   update_prop_rb_void(void *proposal, pmc_simu *psim, error **err);

/* ---- diagnostics ---------------------------------------------------------------- */
double perplexity_and_ess(pmc_simu *psim, int normalize, double *ess, error **err);
double evidence(pmc_simu *psim, double *ln_evi, error **err);
void   clip_weights(pmc_simu *psim, int nclipw, FILE *OUT, error **err);
double mean_from_psim(const double *X, const double *weights, const short *flg, long nsamples, int ndim, int a);
void   estimate_param_covar_weight(size_t ndim, size_t nsamples, size_t nskip, const double *X,
                                   const double *weight, double *pmean, double *pvar, error **err);

/* ---- B200 binding: device context and target registry ----------------------------
 * The scalar host callback (posterior_log_pdf_common_void, param.c:948-954)
 * cannot be batched; a caller registers the flattened device target that stands
 * for a (callback, data) pair once after reading the config.
 * generic_get_importance_weight_and_deduced_verb raises pmc_undef for a callback
 * that has no registered device target. */
pmcb200_ctx *pmc_b200_context(error **err);          /* lazily created; shard 0 (device = $PMCB200_DEVICE or 0) */
void pmc_b200_shutdown(void);
/* number of shards (contexts) the host layer drives: $PMCB200_NGPU (default 1, "all" = every
 * visible device), devices from $PMCB200_DEVICES (comma-separated, round-robin) */
int pmc_b200_nshards(void);
/* Device mirrors of a psim's arrays are reused between the pmclib-named calls of one iteration instead of being
 * re-uploaded (host/pmc.c "what the device mirrors hold").  A caller that edits single elements of X / weights /
 * flg / indices between two calls must say so; wholesale rewrites are detected.  psim == NULL: forget everything. */
void pmc_b200_invalidate_mirror(const pmc_simu *psim);
/* bytes uploaded / uploads avoided by the mirror bookkeeping since start (diagnostic) */
void pmc_b200_mirror_traffic(long *uploaded_bytes, long *skipped_bytes);
void pmc_b200_register_target(posterior_log_pdf_func *posterior_log_pdf, void *target_data,
                              const pmcb200_target_t *t, error **err);
/* Auto-binding hook: called for a (callback, data) pair that has no registered
 * target.  The library's weak default returns 0 (=> pmc_undef).  A glue unit
 * compiled against the caller's headers (cosmopmc_b200/glue/pmcb200_glue.c for
 * CosmoPMC's config_base) overrides it, fills *t and returns 1, which lets the
 * UNCHANGED exec/cosmo_pmc.c run on the device path. */
int pmc_b200_autobind(posterior_log_pdf_func *posterior_log_pdf, void *target_data, pmcb200_target_t *t,
                      error **err);
/* one log-likelihood on the device (N = 1), used by the scalar nicaea-named
 * functions (chi2_SN, chi2_bao_*, chi2_cmbDP); x has like->npar entries */
double pmc_b200_single_loglike(const pmcb200_like_t *like, const double *x, error **err);
/* importance_sample() of exec/importance_sample.c:26-91 as one batched device evaluation:
 * weights[i] <- log posterior(x_i) for every point of psim (isLog = 1, maxW set); flg = 0 for
 * points whose posterior raised an error.  Returns the number of flagged points. */
size_t pmc_b200_importance_sample(pmc_simu *psim, posterior_log_pdf_func *posterior_log_pdf, void *target_data,
                                  error **err);
/* whole iteration on a host psim + proposal in one call (the fast path used by
 * a binding that replaces the body of run_pmc_iteration_MPI) */
size_t pmc_b200_iteration(pmc_simu *psim, mix_mvdens *proposal, gsl_rng *r, double beta,
                          pmcb200_stats_t *stats, error **err);

#ifdef __cplusplus
}
#endif
#endif
