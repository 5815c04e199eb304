// small_kernels.cuh -- non-templated kernels: weight normalisation, fixed-order
// reduction of the EM block partials, the M-step, and the DFMA peak probe.
#pragma once
#include "pmc_kernels.cuh"

// normalize_importance_weight (cosmo_pmc.c:378): wbar = exp(lw - M)/S
__global__ void __launch_bounds__(PMC_BLOCK)
k_normalize(int64_t N, const int16_t *__restrict__ flg, double *__restrict__ w, double M, double invS) {
  int64_t n = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= N) return;
  w[n] = flg[n] ? exp(w[n] - M) * invS : 0.0;
}

// fixed-order sum of the block partials -> this rank's stat block.  One thread per entry of the block, every
// thread sums its entry over the blocks in ascending order (the order the one-block version of round 1 used:
// identical results), with eight loads in flight; adjacent threads read adjacent addresses.
__global__ void __launch_bounds__(PMC_BLOCK)
k_em_reduce(const double *__restrict__ partials, int nblocks, int64_t len,
            const DevScal *__restrict__ scal, int64_t N_local, double *__restrict__ block) {
  const int64_t o = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (o >= len) return;
  double s = 0.0;
  if (o >= 1 && o != 5 && o != 6 && o != 7) {
    const double *p = partials + o;
    int b = 0;
    for (; b + 8 <= nblocks; b += 8) {
      double v[8];
#pragma unroll
      for (int u = 0; u < 8; u++) v[u] = p[(size_t)(b + u) * len];
#pragma unroll
      for (int u = 0; u < 8; u++) s += v[u];
    }
    for (; b < nblocks; b++) s += p[(size_t)b * len];
  }
  if (o == 0) s = partials[0];          // the shift every block used
  if (o == 5) s = (double)scal->nok_box;
  if (o == 6) s = (double)N_local;
  block[o] = s;
}

// ---- M-step: combine the rank blocks in rank order, update alpha/mu/Sigma,
// dead-component rule (manual.tex:482-490), Cholesky, diagnostics ---------------
// result layout (doubles): [0..16) stats, then wght[K], mean[K*d], chol[K*d*d]
// One block (one warp) per component: lanes over the entries of the component's statistics for the
// combine, over the rows of the covariance for the Cholesky (left-looking; row i's inner product runs over
// q < j in ascending order, as the one-thread version did).  The block that finishes last (device counter)
// normalises the weights and writes the diagnostics.  Every rank runs this on identical input, so the
// updated proposal is bitwise identical across ranks.
// work: [0..K) unnormalised alpha, [K..2K) newly-dead flags (scratch between the blocks)
__global__ void __launch_bounds__(32)
k_em_finish(const double *__restrict__ mix, const MixHdr h, int nranks,
            const double *__restrict__ all, int64_t N_global, double *work,
            double *__restrict__ result, unsigned *done_cnt) {
  const int K = h.K, d = h.d, tri = mix_tri(d), cs = stat_cs(d);
  const int64_t len = stat_len(K, d);
  const double *pivot = mix + (size_t)K * h.stride;
  __shared__ double s_scale[64];
  __shared__ double s_hdr[8];
  __shared__ double s_st[3 + PMCB200_MAX_DIM + PMCB200_MAX_DIM * (PMCB200_MAX_DIM + 1) / 2];
  __shared__ double s_L[PMCB200_MAX_DIM][PMCB200_MAX_DIM + 1];
  __shared__ int s_last;
  const int lane = threadIdx.x, k = blockIdx.x;
  // global max and per-rank rescale
  double M = -INFINITY;
  for (int g = 0; g < nranks; g++) M = fmax(M, all[g * len]);
  for (int g = lane; g < nranks; g += 32) {
    const double Mg = all[g * len];
    s_scale[g] = (Mg == -INFINITY) ? 0.0 : exp(Mg - M);
  }
  __syncwarp();
  // header sums (fixed rank order; every block computes the same values)
  if (lane < 7) {
    const int o = lane;
    double s = 0.0;
    if (o == 0) s = M;
    else if (o == 1) { for (int g = 0; g < nranks; g++) s += all[g * len + 1] * s_scale[g]; }
    else if (o == 2) { for (int g = 0; g < nranks; g++) s += all[g * len + 2] * s_scale[g] * s_scale[g]; }
    else if (o == 3) {   // T_g shifts: sum w_g (lw - M) = e^(Mg-M) [T_g + (Mg-M) S_g]
      for (int g = 0; g < nranks; g++) {
        const double Mg = all[g * len];
        if (s_scale[g] > 0.0) s += s_scale[g] * (all[g * len + 3] + (Mg - M) * all[g * len + 1]);
      }
    } else { for (int g = 0; g < nranks; g++) s += all[g * len + o]; }
    s_hdr[o] = s;
  }
  // this component's statistics, combined over the ranks
  for (int o = lane; o < cs; o += 32) {
    const int64_t off = STAT_HDR + (int64_t)k * cs + o;
    double s = 0.0;
    for (int g = 0; g < nranks; g++) s += all[g * len + off] * (o == 2 ? 1.0 : s_scale[g]);
    s_st[o] = s;
  }
  __syncwarp();
  const double S = s_hdr[1];
  const double *comp = mix + (size_t)k * h.stride;
  double *o_mean = result + RES_HDR + K + (size_t)k * d;
  double *o_chol = result + RES_HDR + K + (size_t)K * d + (size_t)k * d * d;
  const double A = s_st[0], G = s_st[1], count = s_st[2];
  const double alpha = A / S;
  const int was_alive = comp[0] != 0.0;
  int dead = !was_alive || !(alpha >= 1.0 / (double)N_global) || count < (double)PMCB200_MINCOUNT;
  if (!dead) {
    // delta = B/G, mu' = p + delta, Sigma' = (C - G delta delta^T)/A; Cholesky in shared memory
    const double *B = s_st + 3, *Cc = s_st + 3 + d;
    for (int e = lane; e < tri; e += 32) {
      int i = 0;
      while ((i + 1) * (i + 2) / 2 <= e) i++;
      const int j = e - i * (i + 1) / 2;
      const double di = B[i] / G, dj = B[j] / G;
      s_L[i][j] = (Cc[e] - G * di * dj) / A;
    }
    __syncwarp();
    for (int j = 0; j < d; j++) {
      double sjj = s_L[j][j];
      for (int q = 0; q < j; q++) sjj -= s_L[j][q] * s_L[j][q];
      if (!(sjj > 0.0) || !isfinite(sjj)) { dead = 1; break; }       // warp-uniform
      const double ljj = sqrt(sjj);
      for (int i = j + 1 + lane; i < d; i += 32) {
        double t = s_L[i][j];
        for (int q = 0; q < j; q++) t -= s_L[i][q] * s_L[j][q];
        s_L[i][j] = t / ljj;
      }
      __syncwarp();
      if (lane == 0) s_L[j][j] = ljj;
      __syncwarp();
    }
  }
  if (!dead) {
    const double *B = s_st + 3;
    for (int i = lane; i < d; i += 32) o_mean[i] = pivot[i] + B[i] / G;
    for (int e = lane; e < d * d; e += 32) { const int i = e / d, j = e - i * d; o_chol[e] = (j <= i) ? s_L[i][j] : 0.0; }
  } else {   // keep the old mean / factor, weight 0
    const int Dp = pmc_pad_dim(d);
    const double *mean = comp + 2, *L = comp + 2 + Dp;
    for (int i = lane; i < d; i += 32) o_mean[i] = mean[i];
    for (int e = lane; e < d * d; e += 32) { const int i = e / d, j = e - i * d; o_chol[e] = (j <= i) ? L[i * (i + 1) / 2 + j] : 0.0; }
  }
  if (lane == 0) { work[k] = dead ? 0.0 : alpha; work[K + k] = (dead && was_alive) ? 1.0 : 0.0; }
  // ---- last block: normalise the weights, diagnostics
  __threadfence();
  if (lane == 0) s_last = (atomicAdd(done_cnt, 1u) == (unsigned)(K - 1));
  __syncwarp();
  if (!s_last) return;
  __threadfence();
  if (lane == 0) {
    *done_cnt = 0u;                                  // ready for the next launch
    const volatile double *wk = work;
    double wsum = 0.0, enc = 0.0;
    int ndead = 0;
    for (int q = 0; q < K; q++) { wsum += wk[q]; ndead += (int)wk[K + q]; }
    for (int q = 0; q < K; q++) {
      const double a = (wsum > 0.0) ? wk[q] / wsum : 0.0;
      result[RES_HDR + q] = a;
      enc = fma(a, a, enc);
    }
    const double Ng = (double)N_global, S2 = s_hdr[2], T = s_hdr[3];
    result[0] = M;                                   // maxW
    result[1] = S;                                   // sum_shift
    result[2] = log(S) + M;                          // logSum
    result[3] = exp(log(S) - T / S) / Ng;            // perplexity = exp(-sum wbar log wbar)/N
    result[4] = S * S / S2;                          // ESS
    result[5] = log(S) + M - log(Ng);                // ln evidence
    result[6] = 1.0 / enc;                           // ENC (updated proposal)
    result[7] = (double)ndead;
    result[8] = s_hdr[4];                            // nok
    result[9] = s_hdr[5];                            // nok_box
    result[10] = s_hdr[6];                           // N summed over ranks
  }
}

// DFMA-only kernel: 8 independent chains per thread, 2 register operands + 1
// constant per DFMA (roofline denominator probe; see tools/micro/fp64_operands.cu)
__global__ void __launch_bounds__(256) k_fp64_peak(double *out, const double *in, int iters) {
  const double y = in[threadIdx.x & 7];
  double x0 = threadIdx.x, x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3, x4 = x0 + 4, x5 = x0 + 5, x6 = x0 + 6, x7 = x0 + 7;
  for (int i = 0; i < iters; i++) {
#pragma unroll
    for (int u = 0; u < 16; u++) {
      x0 = fma(x0, y, 1e-9); x1 = fma(x1, y, 1e-9); x2 = fma(x2, y, 1e-9); x3 = fma(x3, y, 1e-9);
      x4 = fma(x4, y, 1e-9); x5 = fma(x5, y, 1e-9); x6 = fma(x6, y, 1e-9); x7 = fma(x7, y, 1e-9);
    }
  }
  double s = x0 + x1 + x2 + x3 + x4 + x5 + x6 + x7;
  if (s == 123.456) out[0] = s;
}

// ---- stand-alone weight statistics (normalize_importance_weight, perplexity_and_ess,
// evidence on a pmc_simu's weight column): pass 1 max, pass 2 sums relative to it.
// Persistent blocks, fixed-order final reduction.  out8: M, S, S2, T, nok.
__global__ void __launch_bounds__(PMC_BLOCK)
k_wstat_max(int64_t N, const int16_t *__restrict__ flg, const double *__restrict__ w, double *__restrict__ part) {
  __shared__ double red[32];
  double m = -INFINITY;
  for (int64_t n = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; n < N; n += (int64_t)gridDim.x * blockDim.x)
    if (flg[n]) m = fmax(m, w[n]);
  m = warp_max(m);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = m;
  __syncthreads();
  if (threadIdx.x < 32) {
    double v = (threadIdx.x < (blockDim.x >> 5)) ? red[threadIdx.x] : -INFINITY;
    v = warp_max(v);
    if (threadIdx.x == 0) part[blockIdx.x] = v;
  }
}
__global__ void __launch_bounds__(PMC_BLOCK)
k_wstat_sums(int64_t N, const int16_t *__restrict__ flg, const double *__restrict__ w, int is_log,
             const double *__restrict__ maxpart, int nmax, double *__restrict__ part) {
  __shared__ double red[32];
  double M = 0.0;
  if (is_log) { M = -INFINITY; for (int b = 0; b < nmax; b++) M = fmax(M, maxpart[b]); }
  double tS = 0.0, tS2 = 0.0, tT = 0.0, tN = 0.0;
  for (int64_t n = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; n < N; n += (int64_t)gridDim.x * blockDim.x) {
    if (!flg[n]) continue;
    double v, lw;
    // a flagged sample whose weight is zero / whose log weight is not finite (a stored log(0) read back by
    // pmc_simu_from_file) counts as a draw but contributes nothing: 0 * -inf must not reach the sums
    if (is_log) { lw = w[n] - M; v = exp(lw); if (!(v > 0.0) || !isfinite(lw)) { tN += 1.0; continue; } }
    else { v = w[n]; if (!(v > 0.0)) { tN += 1.0; continue; } lw = log(v); }
    tS += v; tS2 = fma(v, v, tS2); tT = fma(v, lw, tT); tN += 1.0;
  }
  double bS = block_sum(tS, red), bS2 = block_sum(tS2, red), bT = block_sum(tT, red), bN = block_sum(tN, red);
  if (threadIdx.x == 0) {
    double *P = part + (size_t)blockIdx.x * 8;
    P[0] = M; P[1] = bS; P[2] = bS2; P[3] = bT; P[4] = bN;
  }
}
__global__ void k_wstat_final(const double *__restrict__ part, int nblocks, double *__restrict__ out8) {
  if (threadIdx.x == 0 && blockIdx.x == 0) {
    double S = 0, S2 = 0, T = 0, Nn = 0;
    for (int b = 0; b < nblocks; b++) { S += part[b * 8 + 1]; S2 += part[b * 8 + 2]; T += part[b * 8 + 3]; Nn += part[b * 8 + 4]; }
    out8[0] = part[0]; out8[1] = S; out8[2] = S2; out8[3] = T; out8[4] = Nn; out8[5] = 0; out8[6] = 0; out8[7] = 0;
  }
}
