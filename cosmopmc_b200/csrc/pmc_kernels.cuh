// pmc_kernels.cuh -- sampler, mixture log-pdf, importance weights, EM
// sufficient statistics and the M-step.  One sample per thread; D (padded
// dimension) is a template parameter so x/y live in registers.
#pragma once
#include "common.cuh"
#include "stat_layout.cuh"
#include "fastmath.cuh"

// ---- K1: mixture sampler (replaces simulate_mix_mvdens, cosmo_pmc.c:320) ------
// Philox counter = (g_lo, g_hi, call, iter), key = seed; g = global sample
// index, so a shard's draws are independent of the number of ranks.
// Box-Muller pair from one Philox call: rad = sqrt(-2 ln u1) with the table-based log (fastmath.cuh, T = the
// block's shared fast tables), angle by sincospi(2 u2) (no reduction by pi).  The published Box-Muller transform;
// against libm's log / cos / sin the variates agree to ~1e-16.
__device__ __forceinline__ void box_muller(const uint32_t (&r)[4], const double *__restrict__ T, double &z0, double &z1) {
  const double u1 = u53(r[0], r[1]), u2 = u53(r[2], r[3]);
  const double rad = sqrt(-2.0 * fast_log(u1, T));
  double sn, cs;
  sincospi(2.0 * u2, &sn, &cs);
  z0 = rad * cs; z1 = rad * sn;
}
template <int D>
__device__ __forceinline__ void draw_normals(uint64_t seed, uint32_t iter, uint64_t g, int d, int df,
                                             double &u, double (&z)[D], double &tscale, const double *__restrict__ T) {
  uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
  uint32_t g0 = (uint32_t)g, g1 = (uint32_t)(g >> 32), r[4];
  philox4x32_10(g0, g1, 0u, iter, k0, k1, r);
  u = (double)r[0] * (1.0 / 4294967296.0);
  // the d coordinates: pairs with compile-time indices (no selects over the register array)
#pragma unroll
  for (int p = 0; 2 * p < D; p++) {
    double zz0 = 0.0, zz1 = 0.0;
    if (2 * p < d) {
      philox4x32_10(g0, g1, 1u + p, iter, k0, k1, r);
      box_muller(r, T, zz0, zz1);
    }
    z[2 * p] = zz0;
    if (2 * p + 1 < D) z[2 * p + 1] = (2 * p + 1 < d) ? zz1 : 0.0;
  }
  tscale = 1.0;
  if (df > 0) {      // Student-t: df further variates for the chi^2 (continuing the pair sequence)
    double chi2 = 0.0;
    const int nz = d + df;
    for (int p = d / 2; 2 * p < nz; p++) {
      philox4x32_10(g0, g1, 1u + p, iter, k0, k1, r);
      double zz0, zz1;
      box_muller(r, T, zz0, zz1);
      const int i0 = 2 * p, i1 = 2 * p + 1;
      if (i0 >= d && i0 < nz) chi2 = fma(zz0, zz0, chi2);
      if (i1 >= d && i1 < nz) chi2 = fma(zz1, zz1, chi2);
    }
    tscale = sqrt((double)df / chi2);
  }
}

template <int D>
__device__ __forceinline__ void transform_store(const double *__restrict__ comp, int d,
                                                const double (&z)[D], double scale,
                                                const double *__restrict__ bmin,
                                                const double *__restrict__ bmax,
                                                double *__restrict__ xout, int &inbox) {
  const double *mean = comp + 2, *L = comp + 2 + D;      // layout padded to D
  int ok = 1;
#pragma unroll
  for (int i = 0; i < D; i++) {
    if (i < d) {
      double t = 0.0;
#pragma unroll
      for (int k = 0; k <= i; k++) t = fma(L[i * (i + 1) / 2 + k], z[k], t);
      double x = fma(scale, t, mean[i]);
      xout[i] = x;
      if (!(x >= bmin[i] && x <= bmax[i])) ok = 0;
    }
  }
  inbox = ok;
}

template <int D>
__global__ void __launch_bounds__(PMC_BLOCK)
k_simulate(const double *__restrict__ mix, const MixHdr h, const double *__restrict__ box,
           int64_t N, uint64_t seed, uint32_t iter, int64_t offset,
           double *__restrict__ X, int32_t *__restrict__ idx, int16_t *__restrict__ flg,
           DevScal *scal) {
  __shared__ double T[96];
  load_fast_tables(T);
  int64_t n = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  int inbox = 0;
  if (n < N) {
    double u, z[D], ts;
    draw_normals<D>(seed, iter, (uint64_t)(offset + n), h.d, h.df, u, z, ts, T);
    int k = select_component(mix, h, u);
    transform_store<D>(mix + (size_t)k * h.stride, h.d, z, ts, box, box + h.d, X + n * h.d, inbox);
    idx[n] = k;
    flg[n] = (int16_t)inbox;
  }
  unsigned b = __ballot_sync(0xffffffffu, inbox);
  if ((threadIdx.x & 31) == 0 && b) atomicAdd(&scal->nok_box, (unsigned long long)__popc(b));
}

// Staged variant (the default): every lane transforms with ITS OWN component's Cholesky factor, so the factor
// loads are a gather over up to K addresses, and every lane stores its own 8d-byte row (32 sectors per store).
// Here the packed mixture sits in shared memory with an odd component stride (lanes of different components
// fall on different banks) and the tile of samples is assembled in shared memory (odd row stride) and written
// out with fully coalesced stores.  Same arithmetic, same results.
template <int D>
__global__ void __launch_bounds__(PMC_BLOCK)
k_simulate_staged(const double *__restrict__ mixg, const MixHdr h, const double *__restrict__ box,
                  int64_t N, uint64_t seed, uint32_t iter, int64_t offset,
                  double *__restrict__ X, int32_t *__restrict__ idx, int16_t *__restrict__ flg,
                  DevScal *scal) {
  extern __shared__ double s_sim[];
  const int sstride = h.stride | 1, xs = h.d | 1;
  double *s_out = s_sim;                                   // [PMC_BLOCK][xs]
  double *s_mix = s_sim + (size_t)PMC_BLOCK * xs;          // [K][sstride]
  for (int i = threadIdx.x; i < h.K * h.stride; i += PMC_BLOCK) {
    const int k = i / h.stride;
    s_mix[k * sstride + (i - k * h.stride)] = mixg[i];
  }
  __shared__ double T[96];
  load_fast_tables(T);      // ends with __syncthreads()
  const int64_t base = (int64_t)blockIdx.x * PMC_BLOCK;
  const int64_t n = base + threadIdx.x;
  int inbox = 0;
  if (n < N) {
    double u, z[D], ts;
    draw_normals<D>(seed, iter, (uint64_t)(offset + n), h.d, h.df, u, z, ts, T);
    MixHdr hs = h; hs.stride = sstride;
    int k = select_component(s_mix, hs, u);
    transform_store<D>(s_mix + (size_t)k * sstride, h.d, z, ts, box, box + h.d, s_out + (size_t)threadIdx.x * xs, inbox);
    idx[n] = k;
    flg[n] = (int16_t)inbox;
  }
  unsigned b = __ballot_sync(0xffffffffu, inbox);
  if ((threadIdx.x & 31) == 0 && b) atomicAdd(&scal->nok_box, (unsigned long long)__popc(b));
  __syncthreads();
  const int rows = (int)min((int64_t)PMC_BLOCK, N - base);
  double *Xt = X + base * h.d;
  // the tile is rows * d contiguous doubles of X: element e = r d + c, advanced by PMC_BLOCK per step
  const int tot = rows * h.d, qd = PMC_BLOCK / h.d, rd = PMC_BLOCK - qd * h.d;
  int r = threadIdx.x / h.d, c = threadIdx.x - r * h.d;
  for (int e = threadIdx.x; e < tot; e += PMC_BLOCK) {
    Xt[e] = s_out[(size_t)r * xs + c];
    r += qd; c += rd;
    if (c >= h.d) { c -= h.d; r++; }
  }
}

template <int D>
__global__ void __launch_bounds__(PMC_BLOCK)
k_simulate_from_draws(const double *__restrict__ mix, const MixHdr h,
                      const double *__restrict__ box, int64_t N,
                      const double *__restrict__ U, const double *__restrict__ Z,
                      double *__restrict__ X, int32_t *__restrict__ idx,
                      int16_t *__restrict__ flg) {
  int64_t n = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= N) return;
  double z[D];
#pragma unroll
  for (int i = 0; i < D; i++) z[i] = (i < h.d) ? Z[n * h.d + i] : 0.0;
  int k = select_component(mix, h, U[n]);
  int inbox;
  transform_store<D>(mix + (size_t)k * h.stride, h.d, z, 1.0, box, box + h.d, X + n * h.d, inbox);
  idx[n] = k;
  flg[n] = (int16_t)inbox;
}

// ---- K2: batched mixture log-pdf (mix_mvdens_log_pdf_void, cosmo_pmc.c:343) ---
template <int D>
__device__ __forceinline__ void load_x(const double *__restrict__ X, int64_t n, int d, double (&x)[D]) {
#pragma unroll
  for (int i = 0; i < D; i++) x[i] = (i < d) ? X[n * d + i] : 0.0;
}

template <int D>
__global__ void __launch_bounds__(PMC_BLOCK)
k_logq(const double *__restrict__ mix, const MixHdr h, int64_t N,
       const double *__restrict__ X, double *__restrict__ logq) {
  int64_t n = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= N) return;
  double x[D];
  load_x<D>(X, n, h.d, x);
  logq[n] = mix_logpdf<D>(mix, h, x);
}

// Mixture / Gaussian used as a *target* (likeli_Mvdens / likeli_MixMvdens,
// param.c:1485-1537) or as the Gaussian prior (param.c:1009-1026, sel = the
// indprior gather map, nsel = its length).
template <int D>
__global__ void __launch_bounds__(PMC_BLOCK)
k_like_mix(const double *__restrict__ mix, const MixHdr h, int is_mixture, int64_t N,
           const double *__restrict__ X, int dX, const int *__restrict__ sel,
           const int16_t *__restrict__ flg, double *__restrict__ logpi,
           int32_t *__restrict__ err, int set, double add_const) {
  int64_t n = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= N) return;
  if (flg && !flg[n]) { if (set) { logpi[n] = 0.0; if (err) err[n] = 0; } return; }
  double x[D];
#pragma unroll
  for (int i = 0; i < D; i++) x[i] = (i < h.d) ? X[n * dX + (sel ? sel[i] : i)] : 0.0;
  double res;
  if (is_mixture) res = mix_logpdf<D>(mix, h, x);
  else {
    double y[D];
    double m = comp_maha<D>(mix, h.d, x, y);
    res = comp_logpdf_from_maha(mix, h.d, h.df, m);
  }
  if (set) { logpi[n] = res + add_const; if (err) err[n] = 0; }
  else logpi[n] += res;
}

// ---- K4: importance weights (generic_get_importance_weight_and_deduced_verb,
// cosmo_pmc.c:343-345): log w = beta log pi - log q, flag clearing, running
// max (warp shuffle + one atomic per block) and nok ----------------------------
template <int D>
__global__ void __launch_bounds__(PMC_BLOCK)
k_weights(const double *__restrict__ mix, const MixHdr h, int64_t N,
          const double *__restrict__ X, const double *__restrict__ logpi,
          const int32_t *__restrict__ err, double beta, int16_t *__restrict__ flg,
          double *__restrict__ logw, DevScal *scal) {
  __shared__ double red[32];
  __shared__ int cnt[32];
  int64_t n = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  double lw = -INFINITY;
  int ok = 0;
  if (n < N) {
    if (flg[n]) {
      double x[D];
      load_x<D>(X, n, h.d, x);
      double lq = mix_logpdf<D>(mix, h, x);
      double v = beta * logpi[n] - lq;
      ok = (err[n] == 0) && isfinite(lq) && isfinite(v);
      if (ok) lw = v; else flg[n] = 0;
    }
    logw[n] = ok ? lw : 0.0;
  }
  double wm = warp_max(lw);
  unsigned b = __ballot_sync(0xffffffffu, ok);
  int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  if (lane == 0) { red[w] = wm; cnt[w] = __popc(b); }
  __syncthreads();
  if (w == 0) {
    double v = (lane < (blockDim.x >> 5)) ? red[lane] : -INFINITY;
    int c = (lane < (blockDim.x >> 5)) ? cnt[lane] : 0;
    v = warp_max(v);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
    if (lane == 0 && c > 0) {
      atomicMax(&scal->max_key, dkey(v));
      atomicAdd(&scal->nok, (unsigned long long)c);
    }
  }
}

// K4 for d >= 10, where the K log-pdfs per sample make the kernel LSU-bound (one warp-uniform mixture load
// per FMA; tools/micro/dmma_probe.cu: a DFMA fed by a broadcast LDG runs at 25 % of the FP64 peak, by a
// broadcast LDS at 49 %).  The packed mixture is staged in shared memory; every thread owns S samples
// (tile_base + s * PMC_BLOCK + tid: loads stay coalesced) which share each mixture load; the samples sit in
// shared memory too ([i][s][tid], conflict-free) so that only the S x D working rows of the column-oriented
// substitution live in registers.
// CP (experimental, PMCB200_ESTEP=4): the staged mixture is the column-packed copy (CpLayout, common.cuh), read
// two doubles per LDS; same operations in the same order.
template <int D, int S, bool CP = false>
__global__ void __launch_bounds__(PMC_BLOCK, (S * D <= 48) ? 2 : 1)
k_weights_multi(const double *__restrict__ mixg, const MixHdr h, int64_t N,
                const double *__restrict__ X, const double *__restrict__ logpi,
                const int32_t *__restrict__ err, double beta, int16_t *__restrict__ flg,
                double *__restrict__ logw, DevScal *scal, double *__restrict__ rho) {
  __shared__ double red[32];
  __shared__ int cnt[32];
  extern __shared__ double s_buf[];
  double *s_xt = s_buf;                                    // [D][S][PMC_BLOCK]
  double *s_mix = s_buf + (size_t)D * S * PMC_BLOCK;       // packed mixture
  if constexpr (CP) {
    const int ncp = h.K * CpLayout<D>::stride;
    for (int i = threadIdx.x; i < ncp; i += PMC_BLOCK) s_mix[i] = 0.0;
    __syncthreads();
    for (int k = 0; k < h.K; k++)
      cp_fill_component<D>(mixg + (size_t)k * h.stride, s_mix + (size_t)k * CpLayout<D>::stride, threadIdx.x, PMC_BLOCK);
  } else {
    const int nmix = h.K * h.stride;
    for (int i = threadIdx.x; i < nmix; i += PMC_BLOCK) s_mix[i] = mixg[i];
  }
  const int64_t base = (int64_t)blockIdx.x * (PMC_BLOCK * S) + threadIdx.x;
  bool live[S];
#pragma unroll
  for (int s = 0; s < S; s++) {
    const int64_t n = base + (int64_t)s * PMC_BLOCK;
    live[s] = (n < N) && flg[n];
#pragma unroll
    for (int i = 0; i < D; i++)
      s_xt[((size_t)i * S + s) * PMC_BLOCK + threadIdx.x] = (live[s] && i < h.d) ? X[n * h.d + i] : 0.0;
  }
  __syncthreads();
  double acc[S], m[S], t[S][D];
#pragma unroll
  for (int s = 0; s < S; s++) acc[s] = 0.0;
  for (int k = 0; k < h.K; k++) {
    const double *comp = s_mix + (size_t)k * (CP ? CpLayout<D>::stride : h.stride);
    const double w = comp[0];
    if (w == 0.0) {
      if (rho) {
#pragma unroll
        for (int s = 0; s < S; s++) if (live[s]) rho[(size_t)k * N + base + (int64_t)s * PMC_BLOCK] = 0.0;
      }
      continue;
    }
#pragma unroll
    for (int i = 0; i < D; i++)
#pragma unroll
      for (int s = 0; s < S; s++) t[s][i] = s_xt[((size_t)i * S + s) * PMC_BLOCK + threadIdx.x];
    if constexpr (CP) comp_maha_cols_cp<D, S>(comp, t, m);
    else comp_maha_cols<D, S>(comp, t, m);
#pragma unroll
    for (int s = 0; s < S; s++) {
      const double e = exp(comp_logpdf_from_maha(comp, h.d, h.df, m[s]));
      acc[s] = fma(w, e, acc[s]);
      // E-step cache: alpha_k phi_k(x) exactly as the EM kernel forms it (k_em_stats_mma, phase 1)
      if (rho && live[s]) rho[(size_t)k * N + base + (int64_t)s * PMC_BLOCK] = w * e;
    }
  }
  double lwmax = -INFINITY;
  int nok = 0;
#pragma unroll
  for (int s = 0; s < S; s++) {
    const int64_t n = base + (int64_t)s * PMC_BLOCK;
    if (n < N) {
      int ok = 0;
      double lw = 0.0;
      if (live[s]) {
        const double lq = log(acc[s]);
        const double v = beta * logpi[n] - lq;
        ok = (err[n] == 0) && isfinite(lq) && isfinite(v);
        if (ok) { lw = v; lwmax = fmax(lwmax, v); nok++; } else flg[n] = 0;
      }
      logw[n] = ok ? lw : 0.0;
    }
  }
  double wm = warp_max(lwmax);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) nok += __shfl_xor_sync(0xffffffffu, nok, o);
  int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  if (lane == 0) { red[w] = wm; cnt[w] = nok; }
  __syncthreads();
  if (w == 0) {
    double v = (lane < (PMC_BLOCK >> 5)) ? red[lane] : -INFINITY;
    int c = (lane < (PMC_BLOCK >> 5)) ? cnt[lane] : 0;
    v = warp_max(v);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
    if (lane == 0 && c > 0) {
      atomicMax(&scal->max_key, dkey(v));
      atomicAdd(&scal->nok, (unsigned long long)c);
    }
  }
}

// ---- K5: Rao-Blackwellised EM sufficient statistics (update_prop_rb,
// cosmo_pmc.c:247).  Stat block layout (doubles):
//   [0] M = local max log w  [1] S = sum e^(lw-M)  [2] S2 = sum e^2(lw-M)
//   [3] T = sum e^(lw-M) (lw-M)  [4] nok  [5] nok_box  [6] N_local  [7] -
//   then per component k at 8 + k*cs, cs = 3 + d + tri(d):
//   A = sum w rho, G = sum w rho gamma, count (points drawn from k),
//   B[d] = sum w rho gamma (x - p), C[tri] = sum w rho gamma (x-p)(x-p)^T (lower)
// with w = e^(lw-M) and p the common pivot (weighted mean of the old means).
// Phase 1 (per tile of PMC_BLOCK samples): each thread computes its sample's w and
// responsibilities into shared memory (and bumps the per-component draw counter).
// Phase 2: the K x M' statistics (M' = 2 + d + d(d+1)/2 weighted features per
// component) are a K x tile x M' contraction.  Thread (g, f) owns feature f for
// a contiguous sample chunk g of the tile (G = 256 / M' chunks when M' < 256,
// else several features per thread); it forms the feature once per sample and
// 8-component chunk and accumulates 8 components in registers, so each FMA
// costs one broadcast LDS.  The per-thread partials live in shared memory across
// the tiles of a persistent block and are combined in a fixed order.
#define EM_KCHUNK 8
__host__ __device__ inline int em_nfeat(int d) { return 2 + d + mix_tri(d); }             // A, G, B[d], C[tri]
__host__ __device__ inline int em_nf(int Mp) { return (Mp + PMC_BLOCK - 1) / PMC_BLOCK; }  // features per thread
__host__ __device__ inline int em_xs(int d) { return pmc_pad_dim(d) | 1; }                // smem row stride of x
#define EM_RACC 32        // register accumulators per thread (one feature x up to 32 components)
// register partials pay off only where the shared-memory partials would cost the second
// resident block per SM (measured: C5 K=30,d=8 13.8 -> 10.7 ms; C2 and C3 are faster with s_acc)
__host__ __device__ inline int em_use_reg(int K, int d, int student) {
  if (em_nf(em_nfeat(d)) != 1 || K > EM_RACC) return 0;
  const int KP = (K + EM_KCHUNK - 1) / EM_KCHUNK * EM_KCHUNK;
  const size_t with_acc = ((size_t)KP * PMC_BLOCK * (student ? 2 : 1) + (size_t)PMC_BLOCK * em_xs(d) +
                           (size_t)K * PMC_BLOCK) * sizeof(double);
  return with_acc > 112 * 1024;
}
__host__ __device__ inline size_t em_smem_bytes(int K, int d, int student) {
  const int KP = (K + EM_KCHUNK - 1) / EM_KCHUNK * EM_KCHUNK;   // zero-padded rows
  const size_t nacc = em_use_reg(K, d, student) ? 0 : (size_t)em_nf(em_nfeat(d)) * K * PMC_BLOCK;   // s_acc
  return ((size_t)KP * PMC_BLOCK * (student ? 2 : 1)           // s_wr (, s_wg)
          + (size_t)PMC_BLOCK * em_xs(d)                        // s_x
          + nacc) * sizeof(double)
         + (size_t)K * sizeof(unsigned long long);              // s_cnt
}

// weighted feature f: 0 -> 1 with weight w rho (A), 1 -> 1 with weight w rho gamma (G),
// 2..2+d -> x_i, then x_i x_j (lower triangle), all with weight w rho gamma
__device__ __forceinline__ void em_decode(int f, int d, int &type, int &i, int &j) {
  i = 0; j = 0;
  if (f < 2) { type = f; return; }
  if (f < 2 + d) { type = 3; i = f - 2; return; }
  type = 4;
  int q = f - 2 - d;
  while ((i + 1) * (i + 2) / 2 <= q) i++;
  j = q - i * (i + 1) / 2;
}

template <int KC, int XS>
__device__ __forceinline__ void em_chunk_acc(const double *__restrict__ wp, const double *__restrict__ xi,
                                             const double *__restrict__ xj, int type, int t0, int t1,
                                             double (&acc)[KC]) {
  const double *w = wp + t0;
  const double *pi = xi + (size_t)t0 * XS, *pj = xj + (size_t)t0 * XS;
  if (type == 4) {
#pragma unroll 4
    for (int t = t0; t < t1; t++, w++, pi += XS, pj += XS) {
      const double feat = pi[0] * pj[0];
#pragma unroll
      for (int q = 0; q < KC; q++) acc[q] = fma(w[q * PMC_BLOCK], feat, acc[q]);
    }
  } else if (type == 3) {
#pragma unroll 4
    for (int t = t0; t < t1; t++, w++, pi += XS) {
      const double feat = pi[0];
#pragma unroll
      for (int q = 0; q < KC; q++) acc[q] = fma(w[q * PMC_BLOCK], feat, acc[q]);
    }
  } else {
#pragma unroll 4
    for (int t = t0; t < t1; t++, w++) {
#pragma unroll
      for (int q = 0; q < KC; q++) acc[q] += w[q * PMC_BLOCK];
    }
  }
}
// shared-memory accumulator variant (several features per thread)
template <int KC, int XS>
__device__ __forceinline__ void em_chunk(const double *__restrict__ wp, const double *__restrict__ xi,
                                         const double *__restrict__ xj, int type, int t0, int t1,
                                         double *__restrict__ accp, int kleft) {
  double acc[KC];
#pragma unroll
  for (int q = 0; q < KC; q++) acc[q] = 0.0;
  em_chunk_acc<KC, XS>(wp, xi, xj, type, t0, t1, acc);
#pragma unroll
  for (int q = 0; q < KC; q++) if (q < kleft) accp[(size_t)q * PMC_BLOCK] += acc[q];
}

template <int D, bool REG>
__global__ void __launch_bounds__(PMC_BLOCK, 2)
k_em_stats(const double *__restrict__ mix, const MixHdr h, int64_t N,
           const double *__restrict__ X, const int32_t *__restrict__ idx,
           const int16_t *__restrict__ flg, const double *__restrict__ logw,
           const DevScal *__restrict__ scal, double *__restrict__ partials, int linear, int k0, int Kg) {
  // Kall components in the mixture; this launch accumulates the group [k0, k0 + K)
  // (one launch unless K x 256 doubles of shared memory per array is too much)
  extern __shared__ double sm[];
  constexpr int XS = D | 1;                   // padded row stride (bank spread), compile-time
  const int Kall = h.K, K = Kg, d = h.d, M = stat_cs(d), Mp = em_nfeat(d);
  const bool student = h.df > 0;
  const int NF = em_nf(Mp);
  const int KP = (K + EM_KCHUNK - 1) / EM_KCHUNK * EM_KCHUNK;
  double *s_wr = sm;                          // [KP][PMC_BLOCK]  w*rho (rows >= K stay zero)
  double *s_wg = student ? s_wr + (size_t)KP * PMC_BLOCK : s_wr;   // w*rho*gamma (Gaussian: gamma = 1)
  double *s_x = s_wg + (size_t)KP * PMC_BLOCK;  // [PMC_BLOCK][XS] x - pivot
  unsigned long long *s_cnt = (unsigned long long *)(s_x + (size_t)PMC_BLOCK * XS);        // [K] draws per component
  double *s_acc = (double *)(s_cnt + K);        // [NF][K][PMC_BLOCK] per-thread partials (absent on the register path)
  __shared__ double red[32];
  const double *pivot = mix + (size_t)Kall * h.stride;
  // linear != 0: logw holds normalised (linear) weights, as after
  // normalize_importance_weight; the shift is then 0
  const double M0 = linear ? 0.0 : dunkey(scal->max_key);
  double tS = 0.0, tS2 = 0.0, tT = 0.0, tN = 0.0;
  const int tid = threadIdx.x;
  // thread -> (sample chunk g, feature f)
  const int G = (Mp < PMC_BLOCK) ? PMC_BLOCK / Mp : 1;
  const int g = (Mp < PMC_BLOCK) ? tid / Mp : 0;
  const int f0 = (Mp < PMC_BLOCK) ? tid - g * Mp : tid;
  const bool worker = (Mp < PMC_BLOCK) ? (g < G) : true;
  const int TS = (PMC_BLOCK + G - 1) / G;
  const int t0 = g * TS, t1 = min(PMC_BLOCK, t0 + TS);
  constexpr bool use_reg = REG;     // one feature per thread: partials stay in registers (host: em_use_reg)
  double racc[REG ? EM_RACC : 1];
#pragma unroll
  for (int q = 0; q < (REG ? EM_RACC : 1); q++) racc[q] = 0.0;
  if (!use_reg)
    for (int p = 0; p < NF; p++)
      for (int k = 0; k < K; k++) s_acc[((size_t)p * K + k) * PMC_BLOCK + tid] = 0.0;
  for (int k = K; k < KP; k++) { s_wr[k * PMC_BLOCK + tid] = 0.0; if (student) s_wg[k * PMC_BLOCK + tid] = 0.0; }
  if (tid < K) s_cnt[tid] = 0ull;
  const int64_t ntiles = (N + PMC_BLOCK - 1) / PMC_BLOCK;

  for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const int64_t n = tile * PMC_BLOCK + tid;
    __syncthreads();
    // ---- phase 1
    const bool ok = (n < N) && flg[n] && (!linear || logw[n] > 0.0);
    double w = 0.0;
    if (ok) {
      double x[D], y[D];
      load_x<D>(X, n, d, x);
      double lw;
      if (linear) { w = logw[n]; lw = log(w); }
      else { lw = logw[n] - M0; w = exp(lw); }
      tS += w; tS2 = fma(w, w, tS2); tT = fma(w, lw, tT); tN += 1.0;
      double rt = 0.0;
      for (int ka = 0; ka < Kall; ka++) {
        const double *comp = mix + (size_t)ka * h.stride;
        const double a = comp[0];
        double r = 0.0, gam = 1.0;
        if (a != 0.0) {
          double m = comp_maha<D>(comp, d, x, y);
          r = a * exp(comp_logpdf_from_maha(comp, d, h.df, m));
          if (student) gam = (double)(h.df + d) / ((double)h.df + m);
        }
        rt += r;
        const int k = ka - k0;
        if (k >= 0 && k < K) {
          s_wr[k * PMC_BLOCK + tid] = r;
          if (student) s_wg[k * PMC_BLOCK + tid] = gam;
        }
      }
      const double sc = w / rt;
      for (int k = 0; k < K; k++) {
        double r = s_wr[k * PMC_BLOCK + tid] * sc;
        s_wr[k * PMC_BLOCK + tid] = r;
        if (student) s_wg[k * PMC_BLOCK + tid] *= r;
      }
#pragma unroll
      for (int i = 0; i < D; i++) s_x[tid * XS + i] = x[i] - pivot[i];      // padded: 0 - 0
    } else {
      for (int k = 0; k < K; k++) { s_wr[k * PMC_BLOCK + tid] = 0.0; if (student) s_wg[k * PMC_BLOCK + tid] = 0.0; }
#pragma unroll
      for (int i = 0; i < D; i++) s_x[tid * XS + i] = 0.0;
    }
    // draws per component count every flagged sample (also zero-weight ones)
    if ((n < N) && flg[n]) { const int c = idx[n] - k0; if (c >= 0 && c < K) atomicAdd(&s_cnt[c], 1ull); }
    __syncthreads();
    // ---- phase 2 (rows k >= K of s_wr/s_wg are zero padding up to KP, so the
    // chunk loops need no predicates and every LDS has an immediate offset)
    if (worker && use_reg) {
      if (f0 < Mp) {
        int type, fi, fj;
        em_decode(f0, d, type, fi, fj);
        const double *wsrc = (type == 0) ? s_wr : s_wg;
#pragma unroll
        for (int c = 0; c < (REG ? EM_RACC / EM_KCHUNK : 0); c++) {
          if (c * EM_KCHUNK < K) {
            double (&acc8)[EM_KCHUNK] = *reinterpret_cast<double (*)[EM_KCHUNK]>(&racc[c * EM_KCHUNK]);
            em_chunk_acc<EM_KCHUNK, XS>(wsrc + (size_t)c * EM_KCHUNK * PMC_BLOCK, s_x + fi, s_x + fj, type, t0, t1, acc8);
          }
        }
      }
    } else if (worker) {
      for (int p = 0; p < NF; p++) {
        const int f = f0 + p * PMC_BLOCK;
        if (f >= Mp) break;
        int type, fi, fj;
        em_decode(f, d, type, fi, fj);
        const double *wsrc = (type == 0) ? s_wr : s_wg;
        double *accp = s_acc + (size_t)p * K * PMC_BLOCK + tid;
        for (int kc = 0; kc < K; kc += EM_KCHUNK) {
          const double *wp = wsrc + (size_t)kc * PMC_BLOCK;
          if (K - kc > EM_KCHUNK / 2)
            em_chunk<EM_KCHUNK, XS>(wp, s_x + fi, s_x + fj, type, t0, t1, accp + (size_t)kc * PMC_BLOCK, K - kc);
          else
            em_chunk<EM_KCHUNK / 2, XS>(wp, s_x + fi, s_x + fj, type, t0, t1, accp + (size_t)kc * PMC_BLOCK, K - kc);
        }
      }
    }
  }
  __syncthreads();
  if (use_reg) {        // park the register partials in the (now free) s_wr rows: [k][tid]
    s_acc = s_wr;
#pragma unroll
    for (int q = 0; q < (REG ? EM_RACC : 0); q++) if (q < K) s_acc[(size_t)q * PMC_BLOCK + tid] = racc[q];
    __syncthreads();
  }
  // ---- this block's partial: combine the G sample chunks in fixed order
  double *P = partials + (size_t)blockIdx.x * stat_len(Kall, d) + (size_t)k0 * M;
  double bS = block_sum(tS, red), bS2 = block_sum(tS2, red), bT = block_sum(tT, red), bN = block_sum(tN, red);
  if (tid == 0) {
    double *P0 = partials + (size_t)blockIdx.x * stat_len(Kall, d);
    P0[0] = M0; P0[1] = bS; P0[2] = bS2; P0[3] = bT; P0[4] = bN; P0[5] = 0; P0[6] = 0; P0[7] = 0;
  }
  for (int out = tid; out < K * M; out += PMC_BLOCK) {
    const int k = out / M, fo = out - k * M;       // fo: position in the stat block (count at 2)
    double s = 0.0;
    if (fo == 2) s = (double)s_cnt[k];
    else {
      const int f = fo < 2 ? fo : fo - 1;
      if (Mp < PMC_BLOCK) { for (int gg = 0; gg < G; gg++) s += s_acc[(size_t)k * PMC_BLOCK + gg * Mp + f]; }
      else s = s_acc[((size_t)(f / PMC_BLOCK) * K + k) * PMC_BLOCK + (f % PMC_BLOCK)];
    }
    P[STAT_HDR + out] = s;
  }
}

// ---- K5, FP64 tensor-core variant (K <= 32; the default) ------------------------
// ncu on C3 (d = 20, K = 10): k_em_stats is bound by the LSU data pipe (80 % of its
// wavefront peak; one broadcast LDS per FMA of the K x tile x M' contraction) with the
// FP64 pipe below a quarter busy.  tools/micro/dmma_probe.cu: DMMA.8x8x4 sustains 99.8 %
// of the vector FP64 peak and a DFMA fed by one LDS runs at 49 %.  So phase 2 is issued
// as mma.sync.m8n8k4.f64: M = 8 components, N = 8 features, K = 4 samples -- 256 FMA per
// warp instruction for 1 + 2/MT shared-memory loads per lane.
//   A fragment  a  = w rho gamma [component 8 mt + lane/4][sample s0 + lane%4]
//   B fragment  b  = feature [sample s0 + lane%4][feature 8 tile + lane/4] = x_i x_j, with
//                    a constant 1 stored in column D of the staged row, so that the G
//                    statistic (1*1), the first moments (x_i * 1) and the second moments
//                    are one expression
//   C fragment  c0,c1 = statistic [component 8 mt + lane/4][feature 8 tile + 2 (lane%4) + {0,1}]
// Warp w owns the feature tiles w, w + 8, ... (NT per warp) for all MT component tiles;
// the accumulators stay in registers across the tiles of the persistent block, and every
// (component, feature) sum is owned by exactly one thread: no combine step, fixed order.
// A = sum w rho equals G for a Gaussian proposal; for Student-t warp 0 accumulates it with
// a second set of A fragments (w rho without gamma) against the ones column.
// RHO: phase 1 reads alpha_k phi_k(x_n) from the E-step cache the weight kernel of the same iteration
// filled (same expression, same operands: bit-identical statistics) instead of repeating the whitenings.
#define EM_WSTRIDE (PMC_BLOCK + 4)     // row stride of s_wr / s_wg: the 8 rows of an A fragment fall on distinct banks
#define EM_MMA_MAXACC 24               // MT x NT accumulator pairs per thread
__host__ __device__ inline int em_mma_nt(int D) { return ((1 + D + mix_tri(D) + 7) / 8 + 7) / 8; }
__host__ __device__ inline int em_mma_mt(int K) { return (K + 7) / 8; }
__host__ __device__ inline size_t em_mma_smem_bytes(int K, int d, int student) {
  const int D = pmc_pad_dim(d), KP = 8 * em_mma_mt(K);
  return ((size_t)KP * EM_WSTRIDE * (student ? 2 : 1) + (size_t)PMC_BLOCK * ((D + 1) | 1)) * sizeof(double) +
         (size_t)K * sizeof(unsigned long long) + ((size_t)K * mix_stride(d) + D) * sizeof(double);   // + staged mixture, pivot
}
// accumulators within the register budget and everything staged within the shared-memory budget;
// otherwise the shared-memory kernel (component groups) takes the update
__host__ __device__ inline bool em_mma_ok(int K, int d, int student) {
  const int D = pmc_pad_dim(d);
  return K <= 32 && em_mma_mt(K) * em_mma_nt(D) <= EM_MMA_MAXACC && em_mma_smem_bytes(K, d, student) <= 200 * 1024;
}
__device__ __forceinline__ void dmma884(double &c0, double &c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
               : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

template <int D, int MT, bool STUDENT, bool RHO>
__global__ void __launch_bounds__(PMC_BLOCK, 2)
k_em_stats_mma(const double *__restrict__ mix, const MixHdr h, int64_t N,
               const double *__restrict__ X, const int32_t *__restrict__ idx,
               const int16_t *__restrict__ flg, const double *__restrict__ logw,
               const DevScal *__restrict__ scal, double *__restrict__ partials, int linear,
               const double *__restrict__ rho) {
  static_assert(!(RHO && STUDENT), "the E-step cache holds no Mahalanobis distances (Student-t gamma)");
  extern __shared__ double sm[];
  constexpr int XS = (D + 1) | 1;           // odd row stride (conflict-free rows); column D holds the constant 1
  constexpr int NT = ((1 + D + D * (D + 1) / 2 + 7) / 8 + 7) / 8;
  constexpr int KP = 8 * MT;
  const int K = h.K, d = h.d, M = stat_cs(d);
  const int nfeat = 1 + d + mix_tri(d);       // G, B[d], C[tri] (stat block position: f = 0 -> 1, f >= 1 -> f + 2)
  double *s_wr = sm;                                                        // [KP][EM_WSTRIDE] w rho
  double *s_wg = STUDENT ? s_wr + (size_t)KP * EM_WSTRIDE : s_wr;           // w rho gamma
  double *s_x = s_wg + (size_t)KP * EM_WSTRIDE;                             // [PMC_BLOCK][XS] x - pivot | 1
  unsigned long long *s_cnt = (unsigned long long *)(s_x + (size_t)PMC_BLOCK * XS);   // [K]
  double *s_mix = (double *)(s_cnt + K);      // packed mixture + pivot: broadcast LDS instead of LDG in phase 1
  __shared__ double red[32];
  for (int i = threadIdx.x; i < K * h.stride + D; i += PMC_BLOCK) s_mix[i] = mix[i];
  const double *pivot = s_mix + (size_t)K * h.stride;
  const double M0 = linear ? 0.0 : dunkey(scal->max_key);
  double tS = 0.0, tS2 = 0.0, tT = 0.0, tN = 0.0;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, g = lane >> 2, t = lane & 3;
  // this lane's B-fragment column of each owned feature tile: staged-row offsets (oi, oj)
  int oi[NT], oj[NT];
  int ntv = 0;                                 // owned tiles that hold a feature (warp-uniform)
#pragma unroll
  for (int q = 0; q < NT; q++) {
    const int tile = warp + 8 * q;
    if (tile * 8 < nfeat) ntv = q + 1;
    const int f = tile * 8 + g;
    int a = D, b = D;
    if (f >= 1 && f < 1 + d) a = f - 1;
    else if (f >= 1 + d && f < nfeat) {
      const int qq = f - 1 - d;
      int i = 0;
      while ((i + 1) * (i + 2) / 2 <= qq) i++;
      a = i; b = qq - i * (i + 1) / 2;
    }
    oi[q] = a; oj[q] = b;
  }
  double acc[MT][NT][2], accA[MT][2];
#pragma unroll
  for (int m = 0; m < MT; m++) {
    accA[m][0] = 0.0; accA[m][1] = 0.0;
#pragma unroll
    for (int q = 0; q < NT; q++) { acc[m][q][0] = 0.0; acc[m][q][1] = 0.0; }
  }
  for (int k = K; k < KP; k++) { s_wr[k * EM_WSTRIDE + tid] = 0.0; if (STUDENT) s_wg[k * EM_WSTRIDE + tid] = 0.0; }
  if (tid < K) s_cnt[tid] = 0ull;
  const int64_t ntiles = (N + PMC_BLOCK - 1) / PMC_BLOCK;

  for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const int64_t n = tile * PMC_BLOCK + tid;
    __syncthreads();
    // ---- phase 1: weight and responsibilities of this thread's sample (as k_em_stats; the sample sits in
    // its s_x row -- odd stride, conflict-free -- and each component runs the column-oriented substitution)
    const bool ok = (n < N) && flg[n] && (!linear || logw[n] > 0.0);
    double *xrow = s_x + (size_t)tid * XS;
    if (ok) {
#pragma unroll
      for (int i = 0; i < D; i++) xrow[i] = (i < d) ? X[n * d + i] : 0.0;
      double lw, w;
      if (linear) { w = logw[n]; lw = log(w); }
      else { lw = logw[n] - M0; w = exp(lw); }
      tS += w; tS2 = fma(w, w, tS2); tT = fma(w, lw, tT); tN += 1.0;
      double rt = 0.0;
      for (int k = 0; k < K; k++) {
        const double *comp = s_mix + (size_t)k * h.stride;
        const double a = comp[0];
        double r = 0.0, gam = 1.0;
        if (RHO) r = rho[(size_t)k * N + n];      // alpha_k phi_k(x_n) left by the weight kernel of this iteration
        else if (a != 0.0) {
          double tt[1][D], m1[1];
#pragma unroll
          for (int i = 0; i < D; i++) tt[0][i] = xrow[i];
          comp_maha_cols<D, 1>(comp, tt, m1);
          const double m = m1[0];
          r = a * exp(comp_logpdf_from_maha(comp, d, h.df, m));
          if (STUDENT) gam = (double)(h.df + d) / ((double)h.df + m);
        }
        rt += r;
        s_wr[k * EM_WSTRIDE + tid] = r;
        if (STUDENT) s_wg[k * EM_WSTRIDE + tid] = gam;
      }
      const double sc = w / rt;
      for (int k = 0; k < K; k++) {
        double r = s_wr[k * EM_WSTRIDE + tid] * sc;
        s_wr[k * EM_WSTRIDE + tid] = r;
        if (STUDENT) s_wg[k * EM_WSTRIDE + tid] *= r;
      }
#pragma unroll
      for (int i = 0; i < D; i++) xrow[i] -= pivot[i];      // padded: 0 - 0
    } else {
      for (int k = 0; k < K; k++) { s_wr[k * EM_WSTRIDE + tid] = 0.0; if (STUDENT) s_wg[k * EM_WSTRIDE + tid] = 0.0; }
#pragma unroll
      for (int i = 0; i < D; i++) xrow[i] = 0.0;
    }
    s_x[tid * XS + D] = 1.0;
    if ((n < N) && flg[n]) { const int c = idx[n]; if (c >= 0 && c < K) atomicAdd(&s_cnt[c], 1ull); }
    __syncthreads();
    // ---- phase 2: K x 256 x nfeat contraction on the FP64 tensor cores, 4 samples per step
    const double *wa = s_wg + (size_t)g * EM_WSTRIDE + t;
    const double *wr = s_wr + (size_t)g * EM_WSTRIDE + t;
    const double *xr = s_x + (size_t)t * XS;
#pragma unroll 2
    for (int s0 = 0; s0 < PMC_BLOCK; s0 += 4, wa += 4, wr += 4, xr += 4 * XS) {
      double a[MT];
#pragma unroll
      for (int m = 0; m < MT; m++) a[m] = wa[(size_t)m * 8 * EM_WSTRIDE];
      double b0 = 0.0;
#pragma unroll
      for (int q = 0; q < NT; q++) {
        if (q < ntv) {
          const double b = xr[oi[q]] * xr[oj[q]];
          if (q == 0) b0 = b;
#pragma unroll
          for (int m = 0; m < MT; m++) dmma884(acc[m][q][0], acc[m][q][1], a[m], b);
        }
      }
      if (STUDENT && warp == 0) {
#pragma unroll
        for (int m = 0; m < MT; m++) dmma884(accA[m][0], accA[m][1], wr[(size_t)m * 8 * EM_WSTRIDE], b0);
      }
    }
  }
  __syncthreads();
  // ---- this block's partial
  double *P0 = partials + (size_t)blockIdx.x * stat_len(K, d);
  double bS = block_sum(tS, red), bS2 = block_sum(tS2, red), bT = block_sum(tT, red), bN = block_sum(tN, red);
  if (tid == 0) { P0[0] = M0; P0[1] = bS; P0[2] = bS2; P0[3] = bT; P0[4] = bN; P0[5] = 0; P0[6] = 0; P0[7] = 0; }
  if (tid < K) P0[STAT_HDR + (size_t)tid * M + 2] = (double)s_cnt[tid];
#pragma unroll
  for (int m = 0; m < MT; m++) {
    const int k = m * 8 + g;
    if (k >= K) continue;
    double *Pk = P0 + STAT_HDR + (size_t)k * M;
#pragma unroll
    for (int q = 0; q < NT; q++) {
      if (q >= ntv) continue;
#pragma unroll
      for (int c = 0; c < 2; c++) {
        const int f = (warp + 8 * q) * 8 + 2 * t + c;
        if (f >= nfeat) continue;
        const double v = acc[m][q][c];
        if (f == 0) { Pk[1] = v; if (!STUDENT) Pk[0] = v; }
        else Pk[f + 2] = v;
      }
    }
    if (STUDENT && warp == 0 && t == 0) Pk[0] = accA[m][0];
  }
}

// ---- K5, FP64 tensor cores, warp-sliced (few feature tiles: d <= 10) ---------------------------------------------
// ncu on k_em_stats_mma at C2 (d = 5, K = 10; profiles/r02/c2_k_em_stats_mma_summary.txt): 21 features are 3 tiles, so 3 of
// the block's 8 warps own all of phase 2 (64 dependent k-steps each per 256-sample tile, two accumulator chains),
// the other 5 walk an empty loop between two block barriers: 77 warp instructions per sample, DMMA sub-pipe 35 %, 1.6 ms
// per 1e7 samples where the contraction itself is worth 0.2 ms.  Here the SAMPLES are sliced over the warps instead of
// the feature tiles: a warp takes 32 samples per step, does phase 1 for them (one sample per lane, its own shared-memory
// rows), and contracts them against ALL feature tiles (8 k-steps x MT x NT independent accumulator chains) -- no block
// barrier inside the sample loop, every warp issues DMMA, every shared-memory row is read by the warp that wrote it.
// The 8 warps' accumulators are summed in warp order at the end (fixed order => deterministic for a given grid).
#define EM_WS_STRIDE 36                // row stride of the per-warp w rho rows (4 mod 16: the 8 rows of an A fragment on distinct banks)
__host__ __device__ constexpr int em_ws_nt(int D) { return (1 + D + D * (D + 1) / 2 + 7) / 8; }
__host__ __device__ constexpr int em_ws_perwarp(int D, int MT, bool student) {
  const int work = 8 * MT * EM_WS_STRIDE * (student ? 2 : 1) + 8 * em_ws_nt(D) * EM_WS_STRIDE, accn = MT * em_ws_nt(D) * 64;
  return work > accn ? work : accn;
}
__host__ __device__ inline size_t em_ws_smem_bytes(int K, int d, int student) {
  const int D = pmc_pad_dim(d);
  return ((size_t)(PMC_BLOCK / 32) * em_ws_perwarp(D, em_mma_mt(K), student != 0) + (size_t)K * mix_stride(d) + D) * sizeof(double) +
         (size_t)K * sizeof(unsigned long long);
}
__host__ __device__ inline bool em_ws_ok(int K, int d, int student) {
  const int D = pmc_pad_dim(d);
  return K <= 32 && em_ws_nt(D) <= 9 && em_mma_mt(K) * em_ws_nt(D) <= 24 && em_ws_smem_bytes(K, d, student) <= 200 * 1024;
}

// resident blocks per SM the register budget allows: accumulators + prefetch registers (+ ~50 for the rest)
__host__ __device__ constexpr int em_ws_minb(int D, int MT, bool rho) {
  const int est = 2 * (2 * MT * em_ws_nt(D) + (rho ? 8 * MT : 0) + D) + 50, b = 65536 / (PMC_BLOCK * est);
  return b < 1 ? 1 : (b > 3 ? 3 : b);
}
template <int D, int MT, bool STUDENT, bool RHO>
__global__ void __launch_bounds__(PMC_BLOCK, em_ws_minb(D, MT, RHO))
k_em_stats_mma_ws(const double *__restrict__ mix, const MixHdr h, int64_t N,
                  const double *__restrict__ X, const int32_t *__restrict__ idx,
                  const int16_t *__restrict__ flg, const double *__restrict__ logw,
                  const DevScal *__restrict__ scal, double *__restrict__ partials, int linear,
                  const double *__restrict__ rho) {
  static_assert(!(RHO && STUDENT), "the E-step cache holds no Mahalanobis distances (Student-t gamma)");
  extern __shared__ double sm[];
  constexpr int NT = em_ws_nt(D);
  constexpr int KP = 8 * MT;
  constexpr int WS = EM_WS_STRIDE;
  constexpr int PERW = em_ws_perwarp(D, MT, STUDENT);
  constexpr int NW = PMC_BLOCK / 32;
  const int K = h.K, d = h.d, M = stat_cs(d);
  const int nfeat = 1 + d + mix_tri(d);       // G, B[d], C[tri] (stat block position: f = 0 -> 1, f >= 1 -> f + 2)
  const int ntv = (nfeat + 7) / 8;            // feature tiles in use (d < D leaves the last ones empty)
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, g = lane >> 2, t = lane & 3;
  double *s_wr = sm + (size_t)warp * PERW;                                  // [KP][WS] w rho of this warp's 32 samples
  double *s_wg = STUDENT ? s_wr + KP * WS : s_wr;                           // w rho gamma
  double *s_f = s_wg + KP * WS;                                             // [8 NT][WS] features 1, x_i, x_i x_j (x - pivot) of the 32 samples
  unsigned long long *s_cnt = (unsigned long long *)(sm + (size_t)NW * PERW);   // [K]
  double *s_mix = (double *)(s_cnt + K);
  __shared__ double red[32];
  for (int i = tid; i < K * h.stride + D; i += PMC_BLOCK) s_mix[i] = mix[i];
  const double *pivot = s_mix + (size_t)K * h.stride;
  const double M0 = linear ? 0.0 : dunkey(scal->max_key);
  double tS = 0.0, tS2 = 0.0, tT = 0.0, tN = 0.0;
  double acc[MT][NT][2], accA[MT][2];
#pragma unroll
  for (int m = 0; m < MT; m++) {
    accA[m][0] = 0.0; accA[m][1] = 0.0;
#pragma unroll
    for (int q = 0; q < NT; q++) { acc[m][q][0] = 0.0; acc[m][q][1] = 0.0; }
  }
  for (int k = K; k < KP; k++) { s_wr[k * WS + lane] = 0.0; if (STUDENT) s_wg[k * WS + lane] = 0.0; }
  for (int f = 0; f < 8 * NT; f++) s_f[f * WS + lane] = 0.0;      // a sample without weight keeps the previous (finite) features
  if (tid < K) s_cnt[tid] = 0ull;
  __syncthreads();
  const int64_t nsteps = (N + 31) / 32, stride = (int64_t)gridDim.x * NW;
  // this lane's sample of the NEXT step travels while the tensor cores work on the current one (ncu on the first version:
  // 8.9 long-scoreboard stall cycles per issue, 0.35 of the HBM peak -- load, compute, load without overlap)
  double px[D], plw = 0.0, prh[RHO ? KP : 1];
  int pfl = 0, pix = 0;
  auto fetch = [&](int64_t stf) {
    const int64_t nf = stf * 32 + lane;
    pfl = 0;
    if (nf < N) {
      pfl = flg[nf]; plw = logw[nf]; pix = idx[nf];
#pragma unroll
      for (int i = 0; i < D; i++) px[i] = (i < d) ? X[nf * d + i] : 0.0;
      if (RHO) {
#pragma unroll
        for (int k = 0; k < KP; k++) if (k < K) prh[k] = rho[(size_t)k * N + nf];      // alpha_k phi_k(x_n) left by the weight kernel
      }
    }
  };
  int64_t st = (int64_t)blockIdx.x * NW + warp;
  if (st < nsteps) fetch(st);

  for (; st < nsteps; st += stride) {
    __syncwarp();
    // ---- phase 1: weight and responsibilities of this lane's sample (the arithmetic of k_em_stats_mma)
    const bool fl = pfl != 0;
    const bool ok = fl && (!linear || plw > 0.0);
    if (ok) {
      double lw, w;
      if (linear) { w = plw; lw = log(w); }
      else { lw = plw - M0; w = exp(lw); }
      tS += w; tS2 = fma(w, w, tS2); tT = fma(w, lw, tT); tN += 1.0;
      double rt = 0.0;
      if (RHO) {
#pragma unroll
        for (int k = 0; k < KP; k++) if (k < K) rt += prh[k];
        const double sc = w / rt;
#pragma unroll
        for (int k = 0; k < KP; k++) if (k < K) s_wr[k * WS + lane] = prh[k] * sc;
      } else {
#pragma unroll 5
        for (int k = 0; k < K; k++) {
          const double *comp = s_mix + (size_t)k * h.stride;
          double r = 0.0, gam = 1.0;
          if (comp[0] != 0.0) {
            double tt[1][D], m1[1];
#pragma unroll
            for (int i = 0; i < D; i++) tt[0][i] = px[i];
            comp_maha_cols<D, 1>(comp, tt, m1);
            const double m = m1[0];
            r = comp[0] * exp(comp_logpdf_from_maha(comp, d, h.df, m));
            if (STUDENT) gam = (double)(h.df + d) / ((double)h.df + m);
          }
          rt += r;
          s_wr[k * WS + lane] = r;
          if (STUDENT) s_wg[k * WS + lane] = gam;
        }
        const double sc = w / rt;
#pragma unroll 5
        for (int k = 0; k < K; k++) {
          const double r = s_wr[k * WS + lane] * sc;
          s_wr[k * WS + lane] = r;
          if (STUDENT) s_wg[k * WS + lane] *= r;
        }
      }
      // this sample's feature column: 1, x_i, x_i x_j about the pivot (the lane holds x in registers; the warp's B fragments
      // are then ONE conflict-free load each -- the gather x[oi] * x[oj] from staged rows kept the LSU wavefronts at 70 %)
      double xc[D];
#pragma unroll
      for (int i = 0; i < D; i++) xc[i] = px[i] - pivot[i];      // padded: 0 - 0
      s_f[lane] = 1.0;
#pragma unroll
      for (int i = 0; i < D; i++) if (i < d) s_f[(1 + i) * WS + lane] = xc[i];
#pragma unroll
      for (int i = 0; i < D; i++)
#pragma unroll
        for (int j = 0; j <= i; j++) if (i < d) s_f[(1 + d + i * (i + 1) / 2 + j) * WS + lane] = xc[i] * xc[j];
    } else {
      for (int k = 0; k < K; k++) { s_wr[k * WS + lane] = 0.0; if (STUDENT) s_wg[k * WS + lane] = 0.0; }
    }
    if (fl) { const int c = pix; if (c >= 0 && c < K) atomicAdd(&s_cnt[c], 1ull); }
    if (st + stride < nsteps) fetch(st + stride);
    __syncwarp();
    // ---- phase 2: K x 32 x nfeat on the FP64 tensor cores, 4 samples per k-step, all feature tiles
    const double *wa = s_wg + g * WS + t;
    const double *wr = s_wr + g * WS + t;
    const double *fb = s_f + g * WS + t;
#pragma unroll
    for (int s0 = 0; s0 < 32; s0 += 4) {
      double a[MT];
#pragma unroll
      for (int m = 0; m < MT; m++) a[m] = wa[m * 8 * WS + s0];
      double b0 = 0.0;
#pragma unroll
      for (int q = 0; q < NT; q++) {
        if (q < ntv) {
          const double b = fb[q * 8 * WS + s0];
          if (q == 0) b0 = b;
#pragma unroll
          for (int m = 0; m < MT; m++) dmma884(acc[m][q][0], acc[m][q][1], a[m], b);
        }
      }
      if (STUDENT) {
#pragma unroll
        for (int m = 0; m < MT; m++) dmma884(accA[m][0], accA[m][1], wr[m * 8 * WS + s0], b0);
      }
    }
  }
  // ---- this block's partial: the warps' accumulators summed in warp order
  __syncwarp();
  double *s_acc = sm + (size_t)warp * PERW;      // [(m NT + q) 2 + c][lane], over this warp's own (now idle) rows
#pragma unroll
  for (int m = 0; m < MT; m++)
#pragma unroll
    for (int q = 0; q < NT; q++) {
      s_acc[((m * NT + q) * 2 + 0) * 32 + lane] = acc[m][q][0];
      s_acc[((m * NT + q) * 2 + 1) * 32 + lane] = acc[m][q][1];
    }
  __syncthreads();
  double *P0 = partials + (size_t)blockIdx.x * stat_len(K, d);
  for (int slot = tid; slot < MT * NT * 64; slot += PMC_BLOCK) {
    const int ln = slot & 31, c = (slot >> 5) & 1, mq = slot >> 6, m = mq / NT, q = mq - m * NT;
    const int k = m * 8 + (ln >> 2), f = q * 8 + 2 * (ln & 3) + c;
    if (k >= K || f >= nfeat) continue;
    double v = 0.0;
#pragma unroll
    for (int w = 0; w < NW; w++) v += sm[(size_t)w * PERW + slot];
    double *Pk = P0 + STAT_HDR + (size_t)k * M;
    if (f == 0) { Pk[1] = v; if (!STUDENT) Pk[0] = v; }
    else Pk[f + 2] = v;
  }
  if (STUDENT) {      // A = sum w rho (without gamma): column 0 of the ones tile, rows 8 m + g, held by the lanes t = 0
    __syncthreads();
#pragma unroll
    for (int m = 0; m < MT; m++) if (t == 0) s_acc[m * 8 + g] = accA[m][0];
    __syncthreads();
    if (tid < K) {
      double v = 0.0;
#pragma unroll
      for (int w = 0; w < NW; w++) v += sm[(size_t)w * PERW + tid];
      P0[STAT_HDR + (size_t)tid * M] = v;
    }
  }
  double bS = block_sum(tS, red), bS2 = block_sum(tS2, red), bT = block_sum(tT, red), bN = block_sum(tN, red);
  if (tid == 0) { P0[0] = M0; P0[1] = bS; P0[2] = bS2; P0[3] = bT; P0[4] = bN; P0[5] = 0; P0[6] = 0; P0[7] = 0; }
  if (tid < K) P0[STAT_HDR + (size_t)tid * M + 2] = (double)s_cnt[tid];
}
