// k_cosmo.cu -- the cosmology / analytic likelihood kernels and the small
// non-templated kernels, with their host launch wrappers.
#include "cosmo.cuh"
#include "sn_spectral.cuh"
#include "small_kernels.cuh"
#include "launch.h"
#include <cstdlib>
#include <cstring>
#include <cmath>
#include <algorithm>
#include <vector>

// Tables of cmb_spec_w (cosmo.cuh), long double.  NR qromb with K = 5 (pmclib sm2_qromberg): stage j's value
// is the degree-4 extrapolation to h^2 -> 0 of the trapezoid sums s_(j-4..j) (abscissae 4^-q), its error estimate the
// difference to the degree-3 extrapolation of s_(j-3..j) (the last Neville correction on polint's path); s_k is the
// trapezoid sum over 2^(k-1) intervals, whose nodes are every (1024 / 2^(k-1))-th of the 1025 stage-11 nodes t_i = i / 1024.
// Node i in the Chebyshev variable: x_i = affine(ln(t_i + tau)); cardinal functions of the CMB_M Chebyshev points
// l_k(x) = sum'_m (2/M) T_m(x_k) T_m(x).  Row r of the table: sum_i weight_r[i] l_k(x_i).
void pmc_cmb_spectral_tables(double *tk, double *th) {
  typedef long double ld;
  const int M = CMB_M, NN = 1024, K = 5;
  ld lam5[K], lam4[K], xq[K];
  for (int q = 0; q < K; q++) xq[q] = powl(4.0L, -(ld)q);
  for (int k = 0; k < K; k++) {
    lam5[k] = 1.0L; lam4[k] = (k == 0) ? 0.0L : 1.0L;
    for (int m = 0; m < K; m++) {
      if (m == k) continue;
      lam5[k] *= (0.0L - xq[m]) / (xq[k] - xq[m]);
      if (k > 0 && m > 0) lam4[k] *= (0.0L - xq[m]) / (xq[k] - xq[m]);
    }
  }
  const ld tau = (ld)CMB_TAU, vlo = logl(tau), vhi = logl(1.0L + tau);
  std::vector<ld> xk(M), Tk((size_t)M * M);
  for (int k = 0; k < M; k++) {
    const ld th_k = M_PIl * (k + 0.5L) / M;
    xk[k] = cosl(th_k);
    tk[k] = (double)(expl((vhi - vlo) / 2 * xk[k] + (vhi + vlo) / 2) - tau);
    for (int m = 0; m < M; m++) Tk[(size_t)k * M + m] = cosl(m * th_k);
  }
  // cardinal values card[i][k] at the 1025 nodes
  std::vector<ld> card((size_t)(NN + 1) * M);
  std::vector<ld> B(M);
  for (int i = 0; i <= NN; i++) {
    ld x = (2.0L * logl((ld)i / NN + tau) - (vhi + vlo)) / (vhi - vlo);
    x = std::max<ld>(-1.0L, std::min<ld>(1.0L, x));
    const ld thx = acosl(x);
    for (int m = 0; m < M; m++) B[m] = cosl(m * thx) * ((m == 0 ? 1.0L : 2.0L) / M);
    for (int k = 0; k < M; k++) {
      ld v = 0.0L;
      for (int m = 0; m < M; m++) v += B[m] * Tk[(size_t)k * M + m];
      card[(size_t)i * M + k] = v;
    }
  }
  auto stage_row = [&](int j, const ld *lam, double *out) {      // out[k] = sum_i weight[i] card[i][k]
    std::vector<ld> wgt(NN + 1, 0.0L);
    for (int q = 0; q < K; q++) {
      const int kk = j - 4 + q, nk = 1 << (kk - 1), step = NN / nk;
      for (int i = 0; i <= NN; i += step) wgt[i] += lam[q] * ((i == 0 || i == NN) ? 0.5L : 1.0L) / nk;
    }
    for (int k = 0; k < M; k++) {
      ld v = 0.0L;
      for (int i = 0; i <= NN; i++) v += wgt[i] * card[(size_t)i * M + k];
      out[k] = (double)v;
    }
  };
  ld mu5[K];
  for (int q = 0; q < K; q++) mu5[q] = lam5[q] - lam4[q];
  stage_row(11, lam5, th + 0 * M);
  stage_row(11, mu5, th + 1 * M);
  for (int j = 5; j <= 10; j++) { stage_row(j, lam5, th + (2 + j - 5) * M); stage_row(j, mu5, th + (8 + j - 5) * M); }
  const int cm[4] = {0, M - 3, M - 2, M - 1};
  for (int r = 0; r < 4; r++)
    for (int k = 0; k < M; k++) th[(14 + r) * M + k] = (double)((cm[r] == 0 ? 1.0L : 2.0L) / M * Tk[(size_t)k * M + cm[r]]);
}

// nodes t_k [m] and value-space functionals [nrow][m] of the spectral form of the distance to a* (host only: the CPU tests check
// them against a node-by-node Romberg); returns m, *nrow = rows; a null pointer skips that table
extern "C" int pmcb200_cmb_spectral_tables(double *tk, double *theta, int *nrow) {
  std::vector<double> t(CMB_M), th((size_t)CMB_NROW * CMB_M);
  pmc_cmb_spectral_tables(t.data(), th.data());
  if (tk) memcpy(tk, t.data(), sizeof(double) * CMB_M);
  if (theta) memcpy(theta, th.data(), sizeof(double) * CMB_NROW * CMB_M);
  if (nrow) *nrow = CMB_NROW;
  return CMB_M;
}

// Fill the SN kernel's 2^(j/1024) table on the current device (pre-biased high words, cosmo.cuh).
int pmc_init_sn_tables() {
  static double tab[SN_EXP2_N];
  static bool have = false;
  if (!have) {
    for (int j = 0; j < SN_EXP2_N; j++) {
      const double v = (double)exp2l((long double)j / SN_EXP2_N);
      unsigned long long b;
      memcpy(&b, &v, 8);
      b -= (unsigned long long)j << (32 + 10);     // high word -= j << 10
      memcpy(&tab[j], &b, 8);
    }
    have = true;
  }
  // lean_log's table: c_i = 1 + (i + 1/2)/1024, {rc = double(1/c_i), -ln(rc)} (of the ROUNDED reciprocal)
  static double ltab[2 * LOG1K_N];
  if (ltab[0] == 0.0)
    for (int i = 0; i < LOG1K_N; i++) {
      const double rc = (double)(1.0L / (1.0L + ((long double)i + 0.5L) / LOG1K_N));
      ltab[2 * i] = rc;
      ltab[2 * i + 1] = (double)(-logl((long double)rc));
    }
  // the spectral SN kernel's folded DCT: [m][j < M/2] = (2/M) cos(pi m (j + 1/2) / M), row 0 halved
  static double dct[SNS_M * SNS_M / 2];
  if (dct[0] == 0.0)
    for (int m = 0; m < SNS_M; m++)
      for (int j = 0; j < SNS_M / 2; j++)
        dct[m * (SNS_M / 2) + j] = (double)((m == 0 ? 1.0L : 2.0L) / SNS_M * cosl(M_PIl * m * (j + 0.5L) / SNS_M));
  // spectral form of the distance to a* (cmb_spec_w): nodes and value-space functionals
  static double cmb_t[CMB_M], cmb_th[CMB_NROW * CMB_M];
  if (cmb_th[0] == 0.0) pmc_cmb_spectral_tables(cmb_t, cmb_th);
  if (cudaMemcpyToSymbol(CMB_T, cmb_t, sizeof(cmb_t)) != cudaSuccess) return 1;
  if (cudaMemcpyToSymbol(CMB_TH, cmb_th, sizeof(cmb_th)) != cudaSuccess) return 1;
  if (cudaMemcpyToSymbol(SNS_DCT, dct, sizeof(dct)) != cudaSuccess) return 1;
  if (cudaMemcpyToSymbol(g_log1k, ltab, sizeof(ltab)) != cudaSuccess) return 1;
  return cudaMemcpyToSymbol(g_sn_exp2, tab, sizeof(tab)) == cudaSuccess ? 0 : 1;
}

static inline int nblk(int64_t N) { return (int)((N + PMC_BLOCK - 1) / PMC_BLOCK); }

static int64_t sn_warp_max() {
  const char *evw = getenv("PMCB200_SN_WARP_MAX");
  return evw && *evw ? atoll(evw) : 4096;
}
// PMCB200_SN_EXACT=1: node-by-node kernels only (A/B measurements, cross-check of the spectral form); read per call
bool pmc_sn_spectral_wanted(const DevLike &L, int64_t N) {
  const char *ev = getenv("PMCB200_SN_EXACT");
  if (ev && *ev && *ev != '0') return false;
  return L.kind == PMCB200_LIKE_SNIa && L.cheb_W && N > sn_warp_max() && N < 4294967296ll;
}
int pmc_sn_spectral_M() { return SNS_M; }

// spectral SN kernel + the exact kernel over the list of samples it could not certify
template <bool H, bool F>
static void launch_sn_spectral(const DevLike &L, int64_t N, const double *X, int d, const int16_t *flg, double *logpi,
                               int32_t *err, int set, double add, DevCount *cnt, uint32_t *fb_list, unsigned *fb_count,
                               int force_slow, cudaStream_t s) {
  // tensor-core form unless the chi^2 mode has redshift-dependent coefficients (chi2_betaz) or wants log sigma^2; PMCB200_SN_SPEC_V1=1
  // keeps the one-sample-per-thread form (A/B measurements); read per call
  const char *ev = getenv("PMCB200_SN_SPEC_V1");
  const bool mma = L.sn_chi2mode != PMCB200_CHI2_betaz && !L.sn_add_logdetCov && !(ev && *ev && *ev != '0');
  if (mma) {
    // PMCB200_SN_TAIL32=1: coefficient tail m >= 16 on the TF32 path (measured equal to the all-FP64 kernel, 12.29 against
    // 12.22 ms per 1e7 samples -- the legacy HMMA path shares the dispatch with the FP64 pipe -- so not the default; read per call)
    const char *et = getenv("PMCB200_SN_TAIL32");
    const bool t32 = (et && *et == '1') && L.cheb_Wt;
    // persistent: one block per SM (register-bound), each warp walks its own 32-sample tasks
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    // small batches: fewer warps per block, so that the 32-sample tasks spread over all SMs (1e4 samples = 313 tasks: 40 blocks
    // of 8 warps would leave 108 SMs idle and put two warps on every busy scheduler)
    const int64_t ntask = (N + 31) / 32;
    const int wpb = (int)std::max<int64_t>(1, std::min<int64_t>(SNS2_BLOCK / 32, (ntask + sms - 1) / sms));
    const int gm = (int)std::min<int64_t>((ntask + wpb - 1) / wpb, sms);
    if (t32) {
      cudaFuncSetAttribute(k_like_sn_spec_mma<H, F, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SNS2_SMEM);   // per device
      k_like_sn_spec_mma<H, F, true><<<gm, 32 * wpb, SNS2_SMEM, s>>>(L, N, X, d, flg, logpi, err, set, add, cnt, fb_list, fb_count);
    } else {
      cudaFuncSetAttribute(k_like_sn_spec_mma<H, F, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SNS2_SMEM);
      k_like_sn_spec_mma<H, F, false><<<gm, 32 * wpb, SNS2_SMEM, s>>>(L, N, X, d, flg, logpi, err, set, add, cnt, fb_list, fb_count);
    }
  } else {
    k_like_sn_spec<H, F><<<(int)((N + SNS_BLOCK - 1) / SNS_BLOCK), SNS_BLOCK, 0, s>>>(L, N, X, d, flg, logpi, err, set, add, cnt,
                                                                                     fb_list, fb_count);
  }
  const int gl = (int)std::min<int64_t>(148 * 8, (N + SN_BLOCK / 32 - 1) / (SN_BLOCK / 32));
  k_like_sn_warp_list<H, F><<<gl, SN_BLOCK, 0, s>>>(L, X, d, logpi, err, set, add, cnt, force_slow, fb_list, fb_count);
}

void pmc_launch_like(const DevLike &L, int64_t N, const double *X, int d, const int16_t *flg, double *logpi,
                     int32_t *err, int set, double add, DevCount *cnt, uint32_t *fb_list, unsigned *fb_count,
                     cudaStream_t s) {
  const int g = nblk(N);
  const int gs = (int)((N + SN_BLOCK - 1) / SN_BLOCK);
  // PMCB200_SN_FORCE_SLOW=1 routes every warp through libdevice exp (used by the
  // tests to validate the fast path against it)
  static const int force_slow = getenv("PMCB200_SN_FORCE_SLOW") ? atoi(getenv("PMCB200_SN_FORCE_SLOW")) : 0;
  // PMCB200_LIKE_V1=1: round-1 BAO / CMB kernels (A/B measurements, cross-check of the lean integrand); read per call
  // crossover of the SN kernels (measured, 1e4 samples: warp-per-sample node-by-node 0.20 ms, spectral tensor-core kernel
  // 0.105 ms; 5e4: 0.74 / 0.16 ms): the spectral kernel from 4096 samples (one block per SM on 16 SMs) on
  const int64_t sn_warp_max = ::sn_warp_max();
  const char *ev1 = getenv("PMCB200_LIKE_V1");
  const int like_v1 = ev1 && *ev1 && *ev1 != '0';
  switch (L.kind) {
    case PMCB200_LIKE_SNIa:
      // small batches: one sample per warp (fills the machine from ~600 samples on); PMCB200_SN_WARP_MAX overrides
      if (N <= sn_warp_max) {
        const int gw = (int)((N + SN_BLOCK / 32 - 1) / (SN_BLOCK / 32));
        if (L.sn_hasq && L.sn_flat) k_like_sn_warp<true, true><<<gw, SN_BLOCK, 0, s>>>(L, N, X, d, flg, logpi, err, set, add, cnt, force_slow);
        else if (L.sn_hasq) k_like_sn_warp<true, false><<<gw, SN_BLOCK, 0, s>>>(L, N, X, d, flg, logpi, err, set, add, cnt, force_slow);
        else if (L.sn_flat) k_like_sn_warp<false, true><<<gw, SN_BLOCK, 0, s>>>(L, N, X, d, flg, logpi, err, set, add, cnt, force_slow);
        else k_like_sn_warp<false, false><<<gw, SN_BLOCK, 0, s>>>(L, N, X, d, flg, logpi, err, set, add, cnt, force_slow);
        break;
      }
      // large batches: spectral form of the quadrature; the samples it cannot certify go to the exact kernel by list
      if (fb_list && fb_count && pmc_sn_spectral_wanted(L, N)) {
        cudaMemsetAsync(fb_count, 0, sizeof(unsigned), s);
        if (L.sn_hasq && L.sn_flat) launch_sn_spectral<true, true>(L, N, X, d, flg, logpi, err, set, add, cnt, fb_list, fb_count, force_slow, s);
        else if (L.sn_hasq) launch_sn_spectral<true, false>(L, N, X, d, flg, logpi, err, set, add, cnt, fb_list, fb_count, force_slow, s);
        else if (L.sn_flat) launch_sn_spectral<false, true>(L, N, X, d, flg, logpi, err, set, add, cnt, fb_list, fb_count, force_slow, s);
        else launch_sn_spectral<false, false>(L, N, X, d, flg, logpi, err, set, add, cnt, fb_list, fb_count, force_slow, s);
        break;
      }
      if (L.sn_hasq && L.sn_flat) k_like_sn<true, true><<<gs, SN_BLOCK, 0, s>>>(L, N, X, d, flg, logpi, err, set, add, cnt, force_slow);
      else if (L.sn_hasq) k_like_sn<true, false><<<gs, SN_BLOCK, 0, s>>>(L, N, X, d, flg, logpi, err, set, add, cnt, force_slow);
      else if (L.sn_flat) k_like_sn<false, true><<<gs, SN_BLOCK, 0, s>>>(L, N, X, d, flg, logpi, err, set, add, cnt, force_slow);
      else k_like_sn<false, false><<<gs, SN_BLOCK, 0, s>>>(L, N, X, d, flg, logpi, err, set, add, cnt, force_slow);
      break;
    case PMCB200_LIKE_BAO:
      if (like_v1) k_like_bao_v1<<<g, PMC_BLOCK, 0, s>>>(L, N, X, d, flg, logpi, err, set, add);
      else if (L.sn_hasq) k_like_bao<true><<<g, PMC_BLOCK, 0, s>>>(L, N, X, d, flg, logpi, err, set, add, cnt);
      else k_like_bao<false><<<g, PMC_BLOCK, 0, s>>>(L, N, X, d, flg, logpi, err, set, add, cnt);
      break;
    case PMCB200_LIKE_CMBDistPrior:
      if (like_v1) k_like_cmbdp_v1<<<g, PMC_BLOCK, 0, s>>>(L, N, X, d, flg, logpi, err, set, add);
      else {
        // PMCB200_CMB_EXACT=1: the distance to a* node by node for every sample (A/B measurements, cross-check); read per call
        const char *ec = getenv("PMCB200_CMB_EXACT");
        const int spec = !(ec && *ec && *ec != '0');
        if (L.sn_hasq) k_like_cmbdp<true><<<g, PMC_BLOCK, 0, s>>>(L, N, X, d, flg, logpi, err, set, add, cnt, spec);
        else k_like_cmbdp<false><<<g, PMC_BLOCK, 0, s>>>(L, N, X, d, flg, logpi, err, set, add, cnt, spec);
      }
      break;
    case PMCB200_LIKE_BANANA:
      k_like_banana<<<g, PMC_BLOCK, 0, s>>>(L, N, X, d, flg, logpi, err, set, add);
      break;
    default:
      break;
  }
}

void pmc_launch_map_params(const DevLike &L, int64_t N, const double *X, int d, double *out, int32_t *err, cudaStream_t s) {
  k_map_params<<<nblk(N), PMC_BLOCK, 0, s>>>(L, N, X, d, out, err);
}
void pmc_launch_normalize(int64_t N, const int16_t *flg, double *w, double M, double invS, cudaStream_t s) {
  k_normalize<<<nblk(N), PMC_BLOCK, 0, s>>>(N, flg, w, M, invS);
}
void pmc_launch_em_reduce(const double *partials, int nblocks, int64_t len, const DevScal *scal, int64_t N_local,
                          double *block, cudaStream_t s) {
  k_em_reduce<<<(int)((len + PMC_BLOCK - 1) / PMC_BLOCK), PMC_BLOCK, 0, s>>>(partials, nblocks, len, scal, N_local, block);
}
void pmc_launch_em_finish(const double *mix, MixHdr h, int nranks, const double *all, int64_t N_global,
                          double *work, double *result, unsigned *done_cnt, cudaStream_t s) {
  k_em_finish<<<h.K, 32, 0, s>>>(mix, h, nranks, all, N_global, work, result, done_cnt);
}
void pmc_launch_fp64_peak(double *out, const double *in, int blocks, int iters, cudaStream_t s) {
  k_fp64_peak<<<blocks, 256, 0, s>>>(out, in, iters);
}
void pmc_launch_wstat(int64_t N, const int16_t *flg, const double *w, int is_log, int blocks, double *maxpart,
                      double *part, double *out8, cudaStream_t s) {
  if (is_log) k_wstat_max<<<blocks, PMC_BLOCK, 0, s>>>(N, flg, w, maxpart);
  k_wstat_sums<<<blocks, PMC_BLOCK, 0, s>>>(N, flg, w, is_log, maxpart, blocks, part);
  k_wstat_final<<<1, 32, 0, s>>>(part, blocks, out8);
}
