// launch.h -- host-side launch interface between the C-ABI translation unit
// (pmcb200.cu) and the kernel translation units (k_mix.cu x4 dimension groups,
// k_cosmo.cu).  Split so the kernels compile in parallel.
#pragma once
#include "common.cuh"

struct DevLike;

enum MixOp { OP_SIMULATE, OP_SIMULATE_DRAWS, OP_LOGQ, OP_LIKE_MIX, OP_WEIGHTS, OP_EM };

struct MixArgs {
  const double *mix = nullptr; MixHdr h{};
  int64_t N = 0;
  // sampler
  const double *box = nullptr; uint64_t seed = 0; uint32_t iter = 0; int64_t offset = 0;
  double *X = nullptr; int32_t *idx = nullptr; int16_t *flg = nullptr; DevScal *scal = nullptr;
  const double *U = nullptr, *Z = nullptr;
  // log-pdf / weights / EM inputs
  const double *Xc = nullptr; const int32_t *idxc = nullptr; const int16_t *flgc = nullptr;
  double *out = nullptr;
  int is_mixture = 0, dX = 0; const int *sel = nullptr;
  double *logpi = nullptr; const double *logpic = nullptr; int32_t *err = nullptr; const int32_t *errc = nullptr;
  int set = 0; double add_const = 0.0, beta = 1.0;
  double *logw = nullptr; const double *logwc = nullptr;
  double *partials = nullptr; int blocks = 0; size_t smem = 0; int linear = 0; int *nblocks_out = nullptr; int k0 = 0, Kg = 0;
  int em_mma = 0;      // OP_EM: FP64 tensor-core kernel (host checked em_mma_ok)
  // E-step cache: OP_WEIGHTS stores r_k(x_n) = alpha_k phi_k(x_n) at rho[k * N + n] (and sets *rho_written when the
  // kernel it chose does so); OP_EM reads them (rho_in) instead of repeating the K whitenings
  double *rho = nullptr; int *rho_written = nullptr; const double *rho_in = nullptr;
};

bool pmc_mix_launch_g0(int op, const MixArgs &a, cudaStream_t s, cudaError_t *e);
bool pmc_mix_launch_g1(int op, const MixArgs &a, cudaStream_t s, cudaError_t *e);
bool pmc_mix_launch_g2(int op, const MixArgs &a, cudaStream_t s, cudaError_t *e);
bool pmc_mix_launch_g3(int op, const MixArgs &a, cudaStream_t s, cudaError_t *e);
static inline cudaError_t pmc_mix_launch(int op, const MixArgs &a, cudaStream_t s) {
  cudaError_t e = cudaSuccess;
  if (pmc_mix_launch_g0(op, a, s, &e) || pmc_mix_launch_g1(op, a, s, &e) ||
      pmc_mix_launch_g2(op, a, s, &e) || pmc_mix_launch_g3(op, a, s, &e))
    return e;
  return cudaErrorInvalidValue;
}

// cosmology / analytic likelihood kernels (k_cosmo.cu)
// fb_list [>= N] / fb_count: work list of the SN samples the spectral kernel hands to the exact kernel (both may be
// null: the exact kernels run alone); pmc_sn_spectral_wanted says whether a launch of N samples would use them
void pmc_launch_like(const DevLike &L, int64_t N, const double *X, int d, const int16_t *flg, double *logpi,
                     int32_t *err, int set, double add_const, DevCount *cnt, uint32_t *fb_list, unsigned *fb_count,
                     cudaStream_t s);
bool pmc_sn_spectral_wanted(const DevLike &L, int64_t N);
int pmc_sn_spectral_M();
void pmc_cmb_spectral_tables(double *tk, double *th);      // [CMB_M], [CMB_NROW][CMB_M] (cosmo.cuh); host only
void pmc_launch_map_params(const DevLike &L, int64_t N, const double *X, int d, double *out, int32_t *err, cudaStream_t s);
// small kernels (k_cosmo.cu)
void pmc_launch_normalize(int64_t N, const int16_t *flg, double *w, double M, double invS, cudaStream_t s);
void pmc_launch_em_reduce(const double *partials, int nblocks, int64_t len, const DevScal *scal, int64_t N_local,
                          double *block, cudaStream_t s);
void pmc_launch_em_finish(const double *mix, MixHdr h, int nranks, const double *all, int64_t N_global,
                          double *work, double *result, unsigned *done_cnt, cudaStream_t s);
int pmc_init_sn_tables();   // per device, at context creation
void pmc_launch_fp64_peak(double *out, const double *in, int blocks, int iters, cudaStream_t s);
void pmc_launch_wstat(int64_t N, const int16_t *flg, const double *w, int is_log, int blocks, double *maxpart,
                      double *part, double *out8, cudaStream_t s);

// weighted post-processing of a stored sample (k_post.cu)
void pmc_launch_post_moments(int64_t N, int d, const double *X, const int16_t *flg, const double *w,
                             const double *pivot, int blocks, double *partials, double *out, cudaStream_t s);
int pmc_launch_post_hist(int64_t N, int d, const double *X, const int16_t *flg, const double *w, int nd,
                         const int *pidx, const int *nbins, const double *limits, int blocks, double *out,
                         cudaStream_t s);
size_t pmc_post_sigma_temp_bytes(int64_t N);
void pmc_launch_post_sigma(int64_t N, int d, int a, const double *X, const int16_t *flg, const double *w,
                           double center, const double conf[3], double *work, void *temp, size_t temp_bytes,
                           unsigned long long *nflag, double *out8, cudaStream_t s);
