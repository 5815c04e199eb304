/* ============================================================================
 * pmcb200.h -- C-ABI of the B200-native Population Monte Carlo iteration.
 *
 * This is the drop-in boundary for CosmoPMC's hot path (one PMC iteration:
 * sample -> log-posterior -> importance weights -> Rao-Blackwellised EM
 * update), `run_pmc_iteration_MPI`, reference exec/cosmo_pmc.c:293-402.
 *
 * Plain C: pointers, sizes and POD structs only.  No torch / C++ types.
 * Every entry point returns 0 on success or a negative pmclib-style error
 * code (see PMCB200_ERR_*); pmcb200_last_error() returns the message.
 *
 * There is no CPU fallback: every compute entry point launches sm_100a CUDA
 * kernels and fails with PMCB200_ERR_CUDA when no device is usable.
 *
 * Each declaration cites the reference interface (file:line under the
 * reference tree) that it replaces.  pmclib / nicaea are external to the
 * reference tree (install_CosmoPMC.sh:247,261), so for those the citation is
 * the reference's *call site*.
 * ========================================================================== */
#ifndef PMCB200_H
#define PMCB200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PMCB200_VERSION      100
#define PMCB200_MAX_DIM      32   /* max. parameter dimension d            */
#define PMCB200_MAX_COMP     64   /* max. mixture components K             */
#define PMCB200_MAX_DATA     4    /* max. data sets in one posterior       */
#define PMCB200_MINCOUNT     20   /* dead-component rule, manual.tex:482-490 */

/* ---- error codes (negative, pmclib convention: errorlist.h) -------------- */
#define PMCB200_OK            0
#define PMCB200_ERR_CUDA     (-9001)  /* CUDA runtime / no device           */
#define PMCB200_ERR_ARG      (-9002)  /* bad argument                        */
#define PMCB200_ERR_DIM      (-9003)  /* dimension mismatch (mv_dimension)   */
#define PMCB200_ERR_CHOLESKY (-9004)  /* not positive definite (mv_cholesky) */
#define PMCB200_ERR_NOSAMPLE (-9005)  /* no point simulated (pmc_nosamplep)  */
#define PMCB200_ERR_UNSUP    (-9006)  /* unsupported mode (e.g. chi2_dust)   */
#define PMCB200_ERR_STATE    (-9007)  /* call order (proposal/target unset)  */

/* ---- parameter roles ------------------------------------------------------
 * The integer values are *identical* to the reference's `par_t` enum
 * (tools/include/par.h:11-34) so a binding passes `like->par[]` straight in.
 * Only the roles consumed by the SNIa / BAO / CMBDistPrior wrappers
 * (wrappers/src/sn.c:167-224, bao.c:100-147, wmap.c:966-1019) are named. */
enum {
  PMCB200_P_Omegam = 0, PMCB200_P_Omegab = 1, PMCB200_P_Omegade = 2,
  PMCB200_P_h100 = 3, PMCB200_P_Omeganumass = 4, PMCB200_P_Omegac = 5,
  PMCB200_P_OmegaK = 6, PMCB200_P_omegam = 7, PMCB200_P_omegab = 8,
  PMCB200_P_100_omegab = 9, PMCB200_P_omegade = 10, PMCB200_P_omeganumass = 11,
  PMCB200_P_omegac = 12, PMCB200_P_omegaK = 13, PMCB200_P_w0de = 14,
  PMCB200_P_w1de = 15, PMCB200_P_Neffnumass = 20,
  PMCB200_P_M = 37, PMCB200_P_alpha = 38, PMCB200_P_beta = 39,
  PMCB200_P_beta_z = 40, PMCB200_P_logbeta = 41,
  PMCB200_P_stretch = 42, PMCB200_P_color = 43
};

/* ---- likelihood kinds: values match `data_t` (wrappers/include/types.h:5-24)
 * for the in-tree kinds; PMCB200_LIKE_BANANA is this repo's C3 stress target
 * (Wraith et al. 2009), plugged in as a posterior_log_pdf_func. */
enum {
  PMCB200_LIKE_Mvdens = 0, PMCB200_LIKE_MixMvdens = 1, PMCB200_LIKE_SNIa = 3,
  PMCB200_LIKE_CMBDistPrior = 6, PMCB200_LIKE_BAO = 7,
  PMCB200_LIKE_BANANA = 100
};

/* special_t, wrappers/include/init_wrappers.h (none, unity, de_conservative) */
enum { PMCB200_SPECIAL_none = 0, PMCB200_SPECIAL_unity = 1,
       PMCB200_SPECIAL_de_conservative = 2 };

/* chi2mode_t of nicaea sn1a.h as used at wrappers/src/sn.c:33-44,238-257 */
enum { PMCB200_CHI2_simple = 0, PMCB200_CHI2_Theta2_denom_fixed = 1,
       PMCB200_CHI2_no_sc = 2, PMCB200_CHI2_betaz = 3 };

/* method_t, wrappers/include/bao.h:31 */
enum { PMCB200_BAO_distance_A = 0, PMCB200_BAO_distance_d_z = 1,
       PMCB200_BAO_distance_D_V_ratio = 2 };

/* de_param_t of nicaea cosmo.h (par_files/cosmo.par:39-44) */
enum { PMCB200_DE_jassal = 0, PMCB200_DE_linder = 1 };

/* Default cosmological model: the fields of nicaea's `cosmo` that the three
 * distance likelihoods read (par_files/cosmo.par:3-13,44). */
typedef struct {
  double Omega_m, Omega_de, w0_de, w1_de, h_100, Omega_b, Omega_nu_mass,
         Neff_nu_mass;
  int    de_param;
  int    _pad;
} pmcb200_cosmo_t;

/* One data set of the posterior: replaces `common_like` + the plug-in state
 * (wrappers/include/init_wrappers.h:15-22; Sn_state sn.h:31-47; bao_state
 * bao.h:41-48; cmbDP_state wmap.h:55-60).  All pointers are HOST pointers;
 * pmcb200_set_target deep-copies the arrays to the device. */
typedef struct {
  int kind;                        /* PMCB200_LIKE_*                          */
  int npar;                        /* = like->npar                            */
  int par[PMCB200_MAX_DIM];        /* = like->par[] (par_t values)            */
  int special;                     /* state->special                          */
  pmcb200_cosmo_t model;           /* state->model (default cosmology)        */

  /* SNIa (sn.c:138-281; formula Manual/manual.tex:1290-1325) */
  int    sn_chi2mode, sn_add_logdetCov;
  double sn_Theta2[4];             /* (-M, alpha, -beta, beta_z) cosmo_SN.par:9 */
  double sn_Theta2_denom[3];
  double sn_sig_int, sn_v_pec;     /* @INTRINSIC_DISPERSION, @PECULIAR_VELOCITY */
  int    sn_n;                     /* number of supernovae                    */
  const double *sn_z, *sn_m, *sn_s, *sn_c;  /* each [sn_n]                   */
  const double *sn_cov;            /* [sn_n*6]: Vmm Vss Vcc Cms Cmc Csc       */

  /* BAO (bao.c:80-184) and CMBDistPrior (wmap.c:945-1049): Gaussian data */
  int    bao_method;
  int    g_ndim;                   /* dimension of the data mvdens            */
  const double *g_z;               /* BAO redshifts [g_ndim or 2*g_ndim]      */
  const double *g_mean;            /* data vector [g_ndim]                    */
  const double *g_chol;            /* lower Cholesky of data covariance
                                      [g_ndim*g_ndim] row-major               */

  /* Mvdens / MixMvdens analytic targets (param.c:1485-1537), BANANA */
  int    mix_ncomp, mix_ndim, mix_df;
  const double *mix_wght, *mix_mean, *mix_chol; /* [K], [K*d], [K*d*d]        */
  double banana_b, banana_sigma1sq;
} pmcb200_like_t;

/* The posterior: replaces `config_base` as consumed by
 * posterior_log_pdf_common (wrappers/src/param.c:958-1041). */
typedef struct {
  int    npar;                     /* config->npar                            */
  int    ndata;                    /* config->ndata                           */
  double min[PMCB200_MAX_DIM];     /* config->min (flat box prior, parabox)   */
  double max[PMCB200_MAX_DIM];     /* config->max                             */
  pmcb200_like_t like[PMCB200_MAX_DATA];
  /* optional Gaussian prior (param.c:1009-1026): nprior==0 && prior_mean!=NULL
   * means "all parameters"; indprior[i]==1 selects parameter i otherwise */
  int    nprior;
  int    indprior[PMCB200_MAX_DIM];
  int    prior_ndim;
  const double *prior_mean, *prior_chol;
} pmcb200_target_t;

/* Per-iteration summary (host struct).  perplexity/ess: perplexity_and_ess
 * (cosmo_pmc.c:46); ln_evidence: evidence() (cosmo_pmc.c:62); enc:
 * effective_number_of_components (cosmo_pmc.c:84); logSum/maxW: fields of
 * pmc_simu read at exec_helper.c:408-420. */
typedef struct {
  int64_t nsamples;      /* N (global over all ranks)                        */
  int64_t nok_box;       /* samples inside the box (simulate_mix_mvdens nok) */
  int64_t nok;           /* samples with finite weight (importance nok)      */
  double  maxW;          /* max log w                                        */
  double  logSum;        /* log sum_n w_n (unnormalised)                     */
  double  sum_shift;     /* sum exp(log w - maxW) (normalize_... return)     */
  double  perplexity;    /* exp(-sum wbar log wbar)/N                        */
  double  ess;           /* 1/sum wbar^2                                     */
  double  ln_evidence;   /* logSum - log N                                   */
  double  enc;           /* 1/sum alpha_d^2 of the UPDATED proposal          */
  int32_t ndead;         /* components killed by the update                  */
  int32_t _pad;
} pmcb200_stats_t;

typedef struct pmcb200_ctx pmcb200_ctx;

/* ---- life cycle ----------------------------------------------------------- */
/* device: CUDA ordinal.  stream: a cudaStream_t (as void*); NULL = the legacy
 * default stream; (void*)-1 = a private non-blocking stream owned by the
 * library.  Replaces pmc_simu_init_mpi (cosmo_pmc.c:633). */
int  pmcb200_create(int device, void *stream, pmcb200_ctx **out);
void pmcb200_destroy(pmcb200_ctx *ctx);
const char *pmcb200_last_error(const pmcb200_ctx *ctx);
int  pmcb200_version(void);
int  pmcb200_device_count(void);
int  pmcb200_sync(pmcb200_ctx *ctx);
void *pmcb200_stream(pmcb200_ctx *ctx);

/* ---- proposal: replaces the `mix_mvdens` handed to every pmclib call
 * (cosmo_pmc.c:320,343,247).  HOST arrays: wght[K], mean[K*d],
 * chol[K*d*d] = lower Cholesky factors, row-major (std with chol==1).
 * df = -1 Gaussian, >0 Student-t (manual.tex:444-450). */
int pmcb200_set_proposal(pmcb200_ctx *ctx, int ncomp, int ndim, int df,
                         const double *wght, const double *mean,
                         const double *chol);
/* same from covariances (does the Cholesky, mix_mvdens_cholesky_decomp
 * param.c:700); returns PMCB200_ERR_CHOLESKY if a component fails */
int pmcb200_set_proposal_cov(pmcb200_ctx *ctx, int ncomp, int ndim, int df,
                             const double *wght, const double *mean,
                             const double *cov);
/* read back the (updated) proposal; cov may be NULL */
int pmcb200_get_proposal(pmcb200_ctx *ctx, double *wght, double *mean,
                         double *chol, double *cov);

/* ---- target: replaces (posterior_log_pdf_common_void, &config->base)
 * passed at cosmo_pmc.c:343-345 */
int pmcb200_set_target(pmcb200_ctx *ctx, const pmcb200_target_t *t);

/* ---- stage entry points on DEVICE arrays --------------------------------- *
 * X[N*d] row-major (pmc_simu->X), idx[N] int32 (pmc_simu->indices),
 * flg[N] int16 (pmc_simu->flg), logw[N] (pmc_simu->weights while isLog).    */

/* simulate_mix_mvdens, cosmo_pmc.c:320.  Sample index g = offset+n is the
 * Philox counter, so a shard's draws do not depend on the rank count. */
int pmcb200_simulate(pmcb200_ctx *ctx, int64_t N, uint64_t seed, uint32_t iter,
                     int64_t offset, double *dX, int32_t *didx, int16_t *dflg);
/* same transform from caller-supplied draws (parity: component selection
 * bit-exact for identical uniforms): u[N] uniforms, z[N*d] normals */
int pmcb200_simulate_from_draws(pmcb200_ctx *ctx, int64_t N, const double *du,
                                const double *dz, double *dX, int32_t *didx,
                                int16_t *dflg);
/* mix_mvdens_log_pdf_void, cosmo_pmc.c:343 */
int pmcb200_proposal_log_pdf(pmcb200_ctx *ctx, int64_t N, const double *dX,
                             double *dlogq);
/* posterior_log_pdf_common_void, param.c:948-1041.  derr[n] != 0 marks a
 * sample whose likelihood raised an error (manual.tex:507-512). derr may be
 * NULL. */
int pmcb200_posterior_log_pdf(pmcb200_ctx *ctx, int64_t N, const double *dX,
                              double *dlogpi, int32_t *derr);
/* The parameter mapping alone, for parity checks against the reference's compiled code: the
 * `switch (like->par[i])` of likeli_SNIa / likeli_BAO / likeli_CMBDistPrior (sn.c:167-224,
 * bao.c:100-147, wmap.c:966-1019) followed by set_base_parameters (param.c:1544-1661) for data
 * set idata.  dout[n*16 ..] = Omega_m Omega_de w0_de w1_de h_100 Omega_b Omega_nu_mass
 * Neff_nu_mass de_param Theta2[0..3] stretch color 0; derr[n] != 0 where the reference raises
 * tls_cosmo_par / ce_infnan (derr may be NULL). */
int pmcb200_map_params(pmcb200_ctx *ctx, int idata, int64_t N, const double *dX,
                       double *dout, int32_t *derr);
/* generic_get_importance_weight_and_deduced_verb, cosmo_pmc.c:343-345:
 * log w = beta*log pi - log q for flagged samples; clears flg on error or
 * non-finite weight; tracks max log w and nok on the device. */
int pmcb200_importance_weights(pmcb200_ctx *ctx, int64_t N, const double *dX,
                               double beta, int16_t *dflg, double *dlogw);
/* normalize_importance_weight, cosmo_pmc.c:378 (in place: logw -> wbar) */
int pmcb200_normalize_weights(pmcb200_ctx *ctx, int64_t N, const int16_t *dflg,
                              double *dw);

/* ---- EM sufficient statistics (update_prop_rb, cosmo_pmc.c:247) -----------
 * Local step: accumulates this rank's block of pmcb200_stat_block_len()
 * doubles into dblock (device).  The caller all-gathers the blocks of all
 * ranks (NCCL over NVLink; one collective per iteration) into dall
 * [nranks * len] and calls pmcb200_em_finish on every rank, which combines
 * them in rank order (bit-identical on all ranks), performs the M-step, the
 * dead-component rule and the Cholesky on the device, installs the new
 * proposal in ctx and fills *stats.  nranks==1: dall == dblock.
 * E-step cache: for ndim >= 10 and a Gaussian proposal pmcb200_importance_weights
 * leaves alpha_k phi_k(x_n) of its samples in a context-owned buffer, and the NEXT
 * pmcb200_em_local on the same dX, N and proposal reads it instead of repeating the
 * K whitenings per sample (bit-identical statistics).  The caller must therefore not
 * modify dX between the two calls (the reference's iteration never does:
 * cosmo_pmc.c:343-378 only reads psim->X).  Any other call order recomputes. */
int64_t pmcb200_stat_block_len(const pmcb200_ctx *ctx);
int pmcb200_em_local(pmcb200_ctx *ctx, int64_t N, const double *dX,
                     const int32_t *didx, const int16_t *dflg,
                     const double *dlogw, double *dblock);
int pmcb200_em_finish(pmcb200_ctx *ctx, int nranks, const double *dall,
                      int64_t N_global, pmcb200_stats_t *stats);

/* ---- pieces for callers that drive the stages on a host pmc_simu ------------ */
/* flat box prior without a target (parabox, cosmo_pmc.c:624) */
int pmcb200_set_box(pmcb200_ctx *ctx, int ndim, const double *min, const double *max);
/* device counters of the last simulate / importance_weights call */
int pmcb200_read_counts(pmcb200_ctx *ctx, int64_t *nok_box, int64_t *nok, double *maxW);
/* out = {M, S, S2, T, n_flagged}: is_log: M = max log w, S = sum e^(lw-M),
 * S2 = sum e^2(lw-M), T = sum e^(lw-M)(lw-M); linear (normalised) weights: M = 0,
 * S = sum w, S2 = sum w^2, T = sum w log w.  perplexity_and_ess (cosmo_pmc.c:46)
 * and evidence (cosmo_pmc.c:62) follow from these. */
int pmcb200_weight_stats(pmcb200_ctx *ctx, int64_t N, const int16_t *dflg, const double *dw,
                         int is_log, double out[8]);
/* normalize_importance_weight (cosmo_pmc.c:378) without a preceding em_finish */
int pmcb200_normalize_log_weights(pmcb200_ctx *ctx, int64_t N, const int16_t *dflg, double *dw,
                                  double *sum_shift, double *logSum, double *maxW);
/* update_prop_rb (cosmo_pmc.c:247) on NORMALISED weights (isLog = 0), as the
 * reference calls it after normalize_importance_weight */
int pmcb200_em_local_linear(pmcb200_ctx *ctx, int64_t N, const double *dX, const int32_t *didx,
                            const int16_t *dflg, const double *dwbar, double *dblock);

/* ---- whole iteration ------------------------------------------------------ */
/* Device-resident shard: simulate + weights + em_local on N samples starting
 * at global index `offset`; leaves the stat block in dblock.  Any of
 * dX/didx/dflg/dlogw may be NULL (library scratch is used). */
int pmcb200_iteration_local(pmcb200_ctx *ctx, int64_t N, uint64_t seed,
                            uint32_t iter, int64_t offset, double beta,
                            double *dX, int32_t *didx, int16_t *dflg,
                            double *dlogw, double *dblock);
/* Multi-GPU with HOST buffers: one rank's shard; X / idx / flg are copied to the
 * host on a copy stream while the likelihood kernel runs.  The caller then
 * all-gathers dblock, calls pmcb200_em_finish, and fetches the shard's
 * NORMALISED weights with pmcb200_shard_weights_host (which also waits for the
 * outstanding copies). */
int pmcb200_iteration_shard_host(pmcb200_ctx *ctx, int64_t N, uint64_t seed, uint32_t iter,
                                 int64_t offset, double beta, double *hX, int32_t *hidx,
                                 int16_t *hflg, double *dblock);
int pmcb200_shard_weights_host(pmcb200_ctx *ctx, int64_t N, double *hw);
/* Single-GPU, HOST buffers, what the pmclib-named shims call: proposal is
 * taken from ctx; fills the pmc_simu arrays on the host (any may be NULL to
 * skip the copy) and updates the proposal.  hw receives NORMALISED weights
 * (isLog = 0) as after normalize_importance_weight. */
int pmcb200_iteration_host(pmcb200_ctx *ctx, int64_t N, uint64_t seed,
                           uint32_t iter, double beta, double *hX,
                           int32_t *hidx, int16_t *hflg, double *hw,
                           pmcb200_stats_t *stats);

/* Pipelined delivery of the host arrays: _begin returns as soon as the update is done (stats filled, new
 * proposal installed) with the device-to-host copies of X / idx / flg / w still draining on the library's
 * copy stream; the NEXT iteration may be begun at once (the sample arrays exist twice on the device) and its
 * kernels overlap those copies.  pmcb200_host_wait(ctx, 1) returns when the arrays of every iteration but the
 * most recent one are complete on the host, pmcb200_host_wait(ctx, 0) when all are.  Consecutive _begin calls
 * must be given different host arrays.  This is what lets the driver write iteration i's pmcsim file
 * (cosmo_pmc.c:392) while iteration i + 1 computes; pmcb200_iteration_host == _begin + pmcb200_host_wait(ctx, 0). */
int pmcb200_iteration_host_begin(pmcb200_ctx *ctx, int64_t N, uint64_t seed,
                                 uint32_t iter, double beta, double *hX,
                                 int32_t *hidx, int16_t *hflg, double *hw,
                                 pmcb200_stats_t *stats);
int pmcb200_shard_weights_host_begin(pmcb200_ctx *ctx, int64_t N, double *hw);
int pmcb200_host_wait(pmcb200_ctx *ctx, int lag);
/* Lazy sample delivery: any of hX / hidx / hflg may be NULL in the iteration calls above, the array then stays on
 * the device (the weighted post-processing of section "post" works there).  The reference touches psim->X on the
 * host only to dump it together with the component indices (out_pmc_simu_cosmo_pmc, cosmo_pmc.c:392) and to
 * post-process the final sample; this call copies the sample array (N x ndim doubles) and / or the component indices
 * of the most recent iteration to hX / hidx on request (either may be NULL), queued like the other host arrays
 * (complete after pmcb200_host_wait(ctx, 0)). */
int pmcb200_samples_host_begin(pmcb200_ctx *ctx, int64_t N, double *hX, int32_t *hidx);

/* ---- several GPUs in ONE process (the reference shards the sample
 * over MPI ranks, cosmo_pmc.c:323-376: send_simulation / receive_importance_weight)
 * One context per shard; contexts may sit on different devices (or, for tests,
 * on the same one).  Shard r of n owns the global sample indices
 * [r*per, min(N,(r+1)*per)), per = ceil(N/n), and the Philox counter is the global
 * index, so the result does not depend on n.  The only exchange per iteration is
 * the statistics block, moved by peer copies (NVLink P2P between peer devices). */

/* asynchronous copies on the context's stream (order them with pmcb200_sync);
 * hptr should be page-locked (pmcb200_host_alloc) for the copy to overlap */
int pmcb200_h2d_async(pmcb200_ctx *ctx, void *dptr, const void *hptr, size_t bytes);
int pmcb200_d2h_async(pmcb200_ctx *ctx, void *hptr, const void *dptr, size_t bytes);
/* page-locked host memory, portable across devices (pmc_simu_init_plus_ded's
 * lump: X, X_ded, weights, indices, flg); falls back to malloc without a device */
int pmcb200_host_alloc(size_t bytes, void **hptr);
int pmcb200_host_free(void *hptr);
/* all-gather between the contexts of one process: for every q,
 * dall[q][r*len .. (r+1)*len) = dblock[r][0..len), ordered after the work already
 * queued on ctx[r]'s stream and before later work on ctx[q]'s stream.  Replaces
 * the MPI gather of cosmo_pmc.c:353-376. */
int pmcb200_allgather_blocks(pmcb200_ctx *const *ctx, int n, double *const *dblock,
                             double *const *dall, int64_t len);
/* pmcb200_iteration_host sharded over n contexts that hold the same proposal
 * and target: every shard is queued before any is waited for; all contexts end
 * with the identical updated proposal (fixed-order combine in pmcb200_em_finish). */
int pmcb200_iteration_host_multi(pmcb200_ctx *const *ctx, int n, int64_t N, uint64_t seed,
                                 uint32_t iter, double beta, double *hX, int32_t *hidx,
                                 int16_t *hflg, double *hw, pmcb200_stats_t *stats);
/* normalize_importance_weight (cosmo_pmc.c:378) with a max / sum that the caller
 * combined over shards: w <- exp(w - maxW) / sum_shift for flagged samples */
int pmcb200_normalize_with(pmcb200_ctx *ctx, int64_t N, const int16_t *dflg, double *dw,
                           double maxW, double sum_shift);

/* ---- weighted post-processing of a stored sample on the device --
 * DEVICE inputs X[N*d], flg[N] (NULL = all samples), w[N] (NULL = unit weights);
 * HOST outputs.  These replace the host loops and the per-parameter qsort that
 * follow the last iteration (cosmo_pmc.c:441-461, meanvar_sample, histograms_sample). */
/* mean_from_psim (exec_helper.c:79) and estimate_param_covar_weight
 * (exec_helper.c:320-349): weighted mean[d] and covariance[d*d] (second moments
 * about the mean, two passes); cov may be NULL */
int pmcb200_post_moments(pmcb200_ctx *ctx, int64_t N, int d, const double *dX, const int16_t *dflg,
                         const double *dw, double *mean, double *cov);
/* sigma_from_psim (exec_helper.c:201-275) for parameter a around `center`:
 * conf[3] = half confidence volumes (conf_68/2, conf_95/2, conf_99/2);
 * sigma[0..3) upper, sigma[3..6) lower half-widths, -1 where the sample ends
 * before the volume is reached.  *median = median_from_psim (exec_helper.c:164-199).
 * The weights must be normalised (isLog = 0). */
int pmcb200_post_sigma(pmcb200_ctx *ctx, int64_t N, int d, const double *dX, const int16_t *dflg,
                       const double *dw, int a, double center, const double conf[3], double sigma[6],
                       double *median, int64_t *nflagged);
/* acc_histogram (tools/src/nhist.c:87-162) for nhdim = 1 or 2 parameters pidx[],
 * nbins[] bins between limits[2*i], limits[2*i+1] (samples on or outside a limit
 * are dropped; the last axis runs fastest): per bin the number of samples, sum w
 * and sum w^2, from which the reference's data[] and var[] follow. */
int pmcb200_post_histogram(pmcb200_ctx *ctx, int64_t N, int d, const double *dX, const int16_t *dflg,
                           const double *dw, int nhdim, const int *pidx, const int *nbins,
                           const double *limits, double *count, double *sumw, double *sumw2);

/* ---- Fisher matrix at a point ------------------------------
 * go_fishing.c:37-85 fisher_element / :124-156 fisher_matrix_part: the second
 * derivatives of -log posterior by Numerical Recipes (5.7.10),
 *   F_ab = -(P(+a,+b) - P(+a,-b) - P(-a,+b) + P(-a,-b)) / (4 h_a h_b),
 * h_a = fh (max_a - min_a) supplied by the caller (mkmax.h:33), with the
 * reference's diagonal shortcut (the two mixed points coincide with the centre
 * and are evaluated once).  The reference evaluates the 4 d(d+1)/2 stencil
 * points one host call at a time (MPI-split over elements); here they are ONE
 * batched posterior launch.  pos, h: host [npar]; F: host [npar*npar], symmetric;
 * diag_only != 0 computes the diagonal only (mcmcini_fisher_diag,
 * go_fishing.c:100,134) and zeroes the rest.  A likelihood error at any stencil
 * point fails the call (the reference forwards the error), *nbad = their number. */
int pmcb200_fisher_host(pmcb200_ctx *ctx, const double *pos, const double *h, int diag_only,
                        double *F, int *nbad);

/* number of kernels launched by this context since creation (bench's
 * gpu_launches claim) */
/* Column plan of the tensor-core SN kernel for a list of redshifts (host only: diagnostics, CPU tests).  Replaces nothing
 * in the reference: nicaea's chi2_SN (called at wrappers/src/sn.c:262-275) loops over the supernovae one by one; here 8 of
 * them are the columns of a tile, a PRIMARY tile holds the first supernova of 8 distinct redshifts and the s-th further
 * supernova of a redshift sits in the same column of the s-th SECONDARY tile behind it (it reuses the primary column's
 * distance).  tile_col[8 t + j] = index into z[] of the supernova in column j of tile t, or -1 - index for an empty
 * column; tile_sec[t] = 1 for a secondary tile.  Returns the number of tiles (-needed if cap is too small). */
int pmcb200_sn_tile_plan(int n, const double *z, int cap, int *tile_sec, int *tile_col);

int64_t pmcb200_launch_count(const pmcb200_ctx *ctx);
/* measurement helpers (not part of the reference's API):
 * counters[0] = SN integrand evaluations summed over all posterior launches
 * since the last call (resets on read), counters[1] = samples x redshifts
 * walked by the SN kernel, counters[2] / [3] = integrand evaluations / integrals of
 * the BAO and CMB kernels; fp64_peak runs a DFMA-only kernel and returns the
 * measured vector-FP64 peak in TFLOP/s (the roofline denominator that
 * MEASURED_PEAKS.json does not carry). */
int pmcb200_counters(pmcb200_ctx *ctx, int64_t out[4]);
/* the same with the SN kernel split: out[4] = samples evaluated by the spectral SN kernel, out[5] = samples
 * evaluated node by node by the warp-per-sample kernel (small batches, and the samples the spectral kernel could
 * not certify); out[6] = CMB samples whose distance to a* came from the spectral form; n <= 7 entries are written,
 * further ones zeroed */
int pmcb200_counters_ex(pmcb200_ctx *ctx, int64_t *out, int n);
/* Tables of the spectral form of the comoving distance to a* in likeli_CMBDistPrior (wrappers/src/wmap.c:1027-1035 ->
 * nicaea w(a*) -> pmclib sm2_qromberg: 11 stages, 1025 equidistant nodes): m fixed nodes t_k in [0, 1] (a = a* + (1 - a*) t)
 * and nrow rows of m weights on the integrand's values there: row 0 = the stage-11 Romberg value, row 1 = its error
 * estimate, rows 2..7 / 8..13 = the same for stages 5..10, rows 14..17 = Chebyshev coefficients 0, m-3, m-2, m-1 of the
 * interpolant (host only; returns m). */
int pmcb200_cmb_spectral_tables(double *tk, double *theta, int *nrow);
int pmcb200_fp64_peak(pmcb200_ctx *ctx, double *tflops);

/* raw device helpers so C hosts need not link the CUDA runtime */
int pmcb200_dev_alloc(pmcb200_ctx *ctx, size_t bytes, void **dptr);
int pmcb200_dev_free(pmcb200_ctx *ctx, void *dptr);
int pmcb200_h2d(pmcb200_ctx *ctx, void *dptr, const void *hptr, size_t bytes);
int pmcb200_d2h(pmcb200_ctx *ctx, void *hptr, const void *dptr, size_t bytes);

#ifdef __cplusplus
}
#endif
#endif /* PMCB200_H */
