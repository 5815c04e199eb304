// k_post.cu -- weighted post-processing of a PMC sample on the device.
//
// What the reference does on the host after the last iteration, O(N log N) per parameter:
//   mean_from_psim / estimate_param_covar_weight      exec/exec_helper.c:63-119, 320-349
//   sigma_from_psim, median_from_psim (qsort)         exec/exec_helper.c:164-275
//   acc_histogram (1-D / 2-D marginals)               tools/src/nhist.c:87-162
// All HBM-bound streaming passes over X[N][d] (8d + 10 B per sample and pass); the sort of the
// confidence intervals is CUB's radix sort (library code, off the iteration's hot path).
#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_scan.cuh>

#include <algorithm>

#include "common.cuh"
#include "launch.h"

#define POST_BLOCK 256

// ---- moments: S0 = sum w, S1[j] = sum w (x_j - p_j), S2[a][b] = sum w (x_a - p_a)(x_b - p_b), b <= a --------
// Feature-owner threads: a block stages a tile of 256 samples (x - pivot, w) in shared memory;
// thread (g, f) accumulates feature f over the samples n = g (mod G) of every tile in a register.
// Block partials are written per block and summed in block order by k_post_reduce (deterministic).
__global__ void __launch_bounds__(POST_BLOCK)
k_post_moments(int64_t N, int d, const double *__restrict__ X, const int16_t *__restrict__ flg,
               const double *__restrict__ w, const double *__restrict__ pivot, int M, int G,
               double *__restrict__ partials) {
  extern __shared__ double sm[];
  double *sx = sm;                          // [POST_BLOCK][d + 1]: x - pivot, then w (0 when unflagged)
  double *sp = sm + POST_BLOCK * (d + 1);   // [d] pivot, then [M] block totals
  double *tot = sp + d;
  const int tid = threadIdx.x, ld = d + 1;
  for (int j = tid; j < d; j += blockDim.x) sp[j] = pivot ? pivot[j] : 0.0;
  for (int j = tid; j < M; j += blockDim.x) tot[j] = 0.0;
  // feature f -> (a, b): f = 0: S0; 1..d: S1[f-1]; then the lower triangle by rows.
  // M <= 256: G = 256 / M sample groups, one feature per thread; M > 256 (d >= 22): one group,
  // thread t owns features t, t + 256, t + 512.
  const int g = (M <= POST_BLOCK) ? tid / M : 0;
  int ff[3], fa[3], fb[3];
  double acc[3] = {0.0, 0.0, 0.0};
#pragma unroll
  for (int i = 0; i < 3; i++) {
    int f = (M <= POST_BLOCK) ? (i == 0 && g < G ? tid % M : -1) : (tid + i * POST_BLOCK < M ? tid + i * POST_BLOCK : -1);
    ff[i] = f; fa[i] = -1; fb[i] = -1;
    if (f >= 1 && f <= d) fa[i] = f - 1;
    else if (f > d) {
      int t = f - d - 1, a = 0;
      while ((a + 1) * (a + 2) / 2 <= t) a++;
      fa[i] = a; fb[i] = t - a * (a + 1) / 2;
    }
  }
  const int64_t ntiles = (N + POST_BLOCK - 1) / POST_BLOCK;
  for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    __syncthreads();
    const int64_t n = tile * POST_BLOCK + tid;
    double wn = 0.0;
    if (n < N && (!flg || flg[n])) wn = w ? w[n] : 1.0;
    sx[tid * ld + d] = wn;
    // coalesced staging of the tile's rows
    const int64_t base = tile * POST_BLOCK * (int64_t)d;
    const int lim = (int)min((int64_t)POST_BLOCK, N - tile * POST_BLOCK) * d;
    for (int o = tid; o < POST_BLOCK * d; o += POST_BLOCK) {
      const int r = o / d, c = o - r * d;
      sx[r * ld + c] = (o < lim) ? X[base + o] - sp[c] : 0.0;
    }
    __syncthreads();
    if (ff[0] >= 0) {
      for (int r = g; r < POST_BLOCK; r += G) {
        const double *row = sx + r * ld;
        const double ww = row[d];
        if (ww == 0.0) continue;              // unflagged rows may hold anything
#pragma unroll
        for (int i = 0; i < 3; i++) {
          if (ff[i] < 0) continue;
          double v = ww;
          if (fa[i] >= 0) v *= row[fa[i]];
          if (fb[i] >= 0) v *= row[fb[i]];
          acc[i] += v;
        }
      }
    }
  }
#pragma unroll
  for (int i = 0; i < 3; i++)
    if (ff[i] >= 0) atomicAdd(&tot[ff[i]], acc[i]);   // at most 256 / M addends per feature
  __syncthreads();
  for (int j = tid; j < M; j += blockDim.x) partials[(size_t)blockIdx.x * M + j] = tot[j];
}

// d <= 8: one sample per thread, the M = 1 + D + D(D+1)/2 running sums in registers (compile-time D),
// grid-stride over the sample, warp-shuffle + shared-memory reduction per block.  HBM-bound: 8 D + 10
// bytes per sample against ~3 M FP64 operations.
template <int D>
__global__ void __launch_bounds__(POST_BLOCK)
k_post_moments_reg(int64_t N, const double *__restrict__ X, const int16_t *__restrict__ flg,
                   const double *__restrict__ w, const double *__restrict__ pivot, double *__restrict__ partials) {
  constexpr int M = 1 + D + D * (D + 1) / 2;
  double acc[M];
#pragma unroll
  for (int j = 0; j < M; j++) acc[j] = 0.0;
  double piv[D];
#pragma unroll
  for (int j = 0; j < D; j++) piv[j] = pivot ? pivot[j] : 0.0;
  // a warp's 32 RPT rows are 32 RPT D contiguous doubles: loaded coalesced (lane + 32 k) into the
  // warp's shared-memory slab (all loads of the step in flight together), then every lane reads
  // its own rows (odd row stride: conflict-free)
  constexpr int RPT = 2;                        // rows per thread and step
  constexpr int DS = D | 1;
  __shared__ double slab[POST_BLOCK / 32][32 * RPT * DS];
  double *my = slab[threadIdx.x >> 5];
  const int ln = threadIdx.x & 31;
  const int64_t wstep = (int64_t)gridDim.x * (blockDim.x >> 5) * (32 * RPT);
  for (int64_t n0 = ((int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5)) * (32 * RPT); n0 < N; n0 += wstep) {
    const int64_t rem = (N - n0) * D;           // doubles left from the warp's first row
    double v[RPT * D];
#pragma unroll
    for (int k = 0; k < RPT * D; k++) {
      const int o = ln + 32 * k;
      v[k] = (o < rem) ? X[n0 * D + o] : 0.0;
    }
    double ww[RPT];
#pragma unroll
    for (int r = 0; r < RPT; r++) {
      const int64_t n = n0 + ln + 32 * r;
      ww[r] = (n < N && (!flg || flg[n])) ? (w ? w[n] : 1.0) : 0.0;
    }
    __syncwarp();
#pragma unroll
    for (int k = 0; k < RPT * D; k++) {
      const int o = ln + 32 * k;
      my[(o / D) * DS + (o % D)] = v[k];
    }
    __syncwarp();
#pragma unroll
    for (int r = 0; r < RPT; r++) {
      if (ww[r] == 0.0) continue;               // unflagged rows may hold anything
      double x[D];
#pragma unroll
      for (int j = 0; j < D; j++) x[j] = my[(ln + 32 * r) * DS + j] - piv[j];
      acc[0] += ww[r];
#pragma unroll
      for (int a = 0, t = 1 + D; a < D; a++) {
        const double wa = ww[r] * x[a];
        acc[1 + a] += wa;
#pragma unroll
        for (int bb = 0; bb <= a; bb++, t++) acc[t] = fma(wa, x[bb], acc[t]);
      }
    }
  }
  __shared__ double red[POST_BLOCK / 32][M];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll
  for (int j = 0; j < M; j++) {
    double v = acc[j];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if (lane == 0) red[wid][j] = v;
  }
  __syncthreads();
  for (int j = threadIdx.x; j < M; j += blockDim.x) {
    double v = 0.0;
#pragma unroll
    for (int q = 0; q < POST_BLOCK / 32; q++) v += red[q][j];
    partials[(size_t)blockIdx.x * M + j] = v;
  }
}

// block partials -> totals: one warp per feature, lanes stride over the blocks, fixed shuffle tree
// (deterministic; a serial walk over ~1000 partials would cost more than the streaming pass itself)
__global__ void k_post_reduce(const double *__restrict__ partials, int nblocks, int M, double *__restrict__ out) {
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
  for (int j = wid; j < M; j += nw) {
    double s = 0.0;
    for (int b = lane; b < nblocks; b += 32) s += partials[(size_t)b * M + j];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (lane == 0) out[j] = s;
  }
}

// launches the two kernels; out[M] on the device, M = 1 + d + d(d+1)/2
void pmc_launch_post_moments(int64_t N, int d, const double *X, const int16_t *flg, const double *w,
                             const double *pivot, int blocks, double *partials, double *out, cudaStream_t s) {
  const int M = 1 + d + d * (d + 1) / 2;
  if (d <= 8) {
    switch (d) {
#define POST_CASE(D) case D: k_post_moments_reg<D><<<blocks, POST_BLOCK, 0, s>>>(N, X, flg, w, pivot, partials); break;
      POST_CASE(1) POST_CASE(2) POST_CASE(3) POST_CASE(4) POST_CASE(5) POST_CASE(6) POST_CASE(7) POST_CASE(8)
#undef POST_CASE
    }
    k_post_reduce<<<1, 1024, 0, s>>>(partials, blocks, M, out);
    return;
  }
  const int G = M > POST_BLOCK ? 1 : POST_BLOCK / M;
  const size_t smem = sizeof(double) * (POST_BLOCK * (d + 1) + d + M);
  cudaFuncSetAttribute(k_post_moments, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  k_post_moments<<<blocks, POST_BLOCK, smem, s>>>(N, d, X, flg, w, pivot, M, G, partials);
  k_post_reduce<<<1, 1024, 0, s>>>(partials, blocks, M, out);
}

// ---- histogram: acc_histogram, tools/src/nhist.c:87-162 --------------------------------------------------
// bins of 1 or 2 parameters; per bin {count, sum w, sum w^2}.  A sample on or outside a limit is
// dropped (vp <= lo || vp >= hi), bin = (int)((vp - lo) / step) with the top bin protected.
struct HistSpec { int nd; int pidx[2]; int nb[2]; double lo[2], hi[2], stp[2]; };
// R private copies of the histogram per block (lane l adds into copy l % R): with few bins most
// lanes of a warp hit the same bin and shared-memory atomics on one address serialise.
__global__ void __launch_bounds__(POST_BLOCK)
k_post_hist(int64_t N, int d, const double *__restrict__ X, const int16_t *__restrict__ flg,
            const double *__restrict__ w, HistSpec h, int tdim, int R, double *__restrict__ out /* [3][tdim] */) {
  extern __shared__ double sh[];            // [R][3][tdim]
  for (int i = threadIdx.x; i < R * 3 * tdim; i += blockDim.x) sh[i] = 0.0;
  __syncthreads();
  double *mine = sh + (size_t)(threadIdx.x % R) * 3 * tdim;
  constexpr int U = 4;                      // rows per thread and step: U independent strided loads in flight
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t n0 = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; n0 < N; n0 += U * stride) {
    double v0[U], v1[U], wg[U];
    bool ok[U];
#pragma unroll
    for (int u = 0; u < U; u++) {
      const int64_t n = n0 + u * stride;
      ok[u] = n < N && (!flg || flg[n]);
      v0[u] = ok[u] ? X[n * d + h.pidx[0]] : 0.0;
      v1[u] = (ok[u] && h.nd > 1) ? X[n * d + h.pidx[1]] : 0.0;
      wg[u] = ok[u] ? (w ? w[n] : 1.0) : 0.0;
    }
#pragma unroll
    for (int u = 0; u < U; u++) {
      if (!ok[u]) continue;
      int pos = 0, mul = 1;
      bool valid = true;
      for (int ap = 0; ap < h.nd; ap++) {     // nhist.c:118-131: last axis first, it runs fastest
        const int ip = h.nd - ap - 1;
        const double vp = ip ? v1[u] : v0[u];
        if (vp <= h.lo[ip] || vp >= h.hi[ip]) { valid = false; break; }
        int nb = (int)((vp - h.lo[ip]) / h.stp[ip]);
        if (nb == h.nb[ip]) nb--;
        pos += nb * mul;
        mul *= h.nb[ip];
      }
      if (!valid) continue;
      atomicAdd(&mine[pos], 1.0);
      atomicAdd(&mine[tdim + pos], wg[u]);
      atomicAdd(&mine[2 * tdim + pos], wg[u] * wg[u]);
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < 3 * tdim; i += blockDim.x) {
    double v = 0.0;
    for (int r = 0; r < R; r++) v += sh[(size_t)r * 3 * tdim + i];
    if (v != 0.0) atomicAdd(&out[i], v);
  }
}

int pmc_launch_post_hist(int64_t N, int d, const double *X, const int16_t *flg, const double *w, int nd,
                         const int *pidx, const int *nbins, const double *limits, int blocks, double *out,
                         cudaStream_t s) {
  HistSpec h;
  h.nd = nd;
  int tdim = 1;
  for (int i = 0; i < nd; i++) {
    h.pidx[i] = pidx[i]; h.nb[i] = nbins[i]; h.lo[i] = limits[2 * i]; h.hi[i] = limits[2 * i + 1];
    h.stp[i] = (limits[2 * i + 1] - limits[2 * i]) / nbins[i];
    tdim *= nbins[i];
  }
  const size_t one = sizeof(double) * 3 * (size_t)tdim;
  if (one > 200 * 1024) return -1;
  int R = (int)std::min<size_t>(32, (48 * 1024) / one);      // copies within 48 KB
  if (R < 1) R = 1;
  while (R & (R - 1)) R &= R - 1;                             // power of two
  const size_t smem = one * R;
  cudaMemsetAsync(out, 0, one, s);
  cudaFuncSetAttribute(k_post_hist, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  k_post_hist<<<blocks, POST_BLOCK, smem, s>>>(N, d, X, flg, w, h, tdim, R, out);
  return 0;
}

// ---- confidence intervals and median: sort (x_a, w) by x_a, prefix-sum the weights, search -----------
// keys of unflagged samples are +inf with weight 0 (they sort behind the n flagged ones)
__global__ void __launch_bounds__(POST_BLOCK)
k_post_gather(int64_t N, int d, int a, const double *__restrict__ X, const int16_t *__restrict__ flg,
              const double *__restrict__ w, double *__restrict__ key, double *__restrict__ val,
              unsigned long long *__restrict__ nflag) {
  constexpr int U = 4;                      // rows per thread: U independent strided loads in flight
  const int64_t base = ((int64_t)blockIdx.x * U) * blockDim.x + threadIdx.x;
  double xv[U], wv[U];
  bool ok[U];
#pragma unroll
  for (int u = 0; u < U; u++) {
    const int64_t n = base + (int64_t)u * blockDim.x;
    ok[u] = (n < N) && (!flg || flg[n]);
    xv[u] = ok[u] ? X[n * d + a] : INFINITY;
    wv[u] = ok[u] ? (w ? w[n] : 1.0) : 0.0;
  }
  unsigned cnt = 0;
#pragma unroll
  for (int u = 0; u < U; u++) {
    const int64_t n = base + (int64_t)u * blockDim.x;
    if (n < N) { key[n] = xv[u]; val[n] = wv[u]; }
    cnt += __popc(__ballot_sync(0xffffffffu, ok[u]));
  }
  if ((threadIdx.x & 31) == 0 && cnt) atomicAdd(nflag, (unsigned long long)cnt);
}

// sigma_from_psim (exec_helper.c:201-275) and median_from_psim (:164-199) on the sorted sample.
// cw = inclusive prefix sums of the sorted weights.  One thread: a handful of binary searches.
__device__ inline double cw_before(const double *cw, int64_t i) { return i > 0 ? cw[i - 1] : 0.0; }
__global__ void k_post_sigma(const double *__restrict__ key, const double *__restrict__ cw,
                             const unsigned long long *__restrict__ nflag, double center, double c0, double c1,
                             double c2, double *__restrict__ out /* [8] */) {
  if (threadIdx.x || blockIdx.x) return;
  const int64_t n = (int64_t)*nflag;
  for (int j = 0; j < 8; j++) out[j] = -1.0;
  out[7] = (double)n;
  if (n < 1) return;
  // imean = first i with key[i] >= center, but at most n-1 (the reference's loop stops at n-1)
  int64_t lo = 0, hi = n;
  while (lo < hi) { const int64_t mid = (lo + hi) >> 1; if (key[mid] < center) lo = mid + 1; else hi = mid; }
  const int64_t imean = lo < n - 1 ? lo : n - 1;
  const double conf[3] = {c0, c1, c2};
  const double base = cw_before(cw, imean);
  for (int j = 0; j < 3; j++) {
    // right: the reference adds w[imean], w[imean+1], .. while sum <= conf and reports the element
    // after the last one added: the smallest i in [imean+1, n-1] with cw[i-1] - base > conf; none: -1
    int64_t a = imean + 1, b = n - 1, ans = -1;
    while (a <= b) {
      const int64_t mid = (a + b) >> 1;
      if (cw[mid - 1] - base > conf[j]) { ans = mid; b = mid - 1; } else a = mid + 1;
    }
    out[j] = ans < 0 ? -1.0 : key[ans] - center;
    // left: adds w[imean-1], w[imean-2], .. and reports the element below the last one added:
    // the largest i in [0, imean-2] with base - cw[i] > conf; none: -1 (boundary hit)
    a = 0; b = imean - 2; ans = -1;
    while (a <= b) {
      const int64_t mid = (a + b) >> 1;
      if (base - cw[mid] > conf[j]) { ans = mid; a = mid + 1; } else b = mid - 1;
    }
    out[3 + j] = ans < 0 ? -1.0 : center - key[ans];
  }
  // median: first i with cw[i] >= 0.5 (== 0.5 returns key[i]; > 0.5 returns the mean of key[i], key[i-1])
  {
    int64_t a = 0, b = n;
    while (a < b) { const int64_t mid = (a + b) >> 1; if (cw[mid] >= 0.5) b = mid; else a = mid + 1; }
    if (a < n) out[6] = (cw[a] == 0.5 || a == 0) ? key[a] : 0.5 * (key[a] + key[a - 1]);
    else out[6] = NAN;                        // err_median: the weights do not sum beyond 0.5
  }
}

size_t pmc_post_sigma_temp_bytes(int64_t N) {
  size_t a = 0, b = 0;
  cub::DeviceRadixSort::SortPairs(nullptr, a, (const double *)nullptr, (double *)nullptr, (const double *)nullptr,
                                  (double *)nullptr, N);
  cub::DeviceScan::InclusiveSum(nullptr, b, (const double *)nullptr, (double *)nullptr, N);
  return a > b ? a : b;
}

// work: 4 N doubles (key, val, sorted key, sorted val -> prefix sums in place) + 1 counter + temp
void pmc_launch_post_sigma(int64_t N, int d, int a, const double *X, const int16_t *flg, const double *w,
                           double center, const double conf[3], double *work, void *temp, size_t temp_bytes,
                           unsigned long long *nflag, double *out8, cudaStream_t s) {
  double *key = work, *val = work + N, *skey = work + 2 * N, *sval = work + 3 * N;
  cudaMemsetAsync(nflag, 0, sizeof(unsigned long long), s);
  k_post_gather<<<(unsigned)((N + 4 * POST_BLOCK - 1) / (4 * POST_BLOCK)), POST_BLOCK, 0, s>>>(N, d, a, X, flg, w, key, val, nflag);
  cub::DeviceRadixSort::SortPairs(temp, temp_bytes, key, skey, val, sval, N, 0, 64, s);
  cub::DeviceScan::InclusiveSum(temp, temp_bytes, sval, val, N, s);     // val <- prefix sums of the sorted weights
  k_post_sigma<<<1, 32, 0, s>>>(skey, val, nflag, center, conf[0], conf[1], conf[2], out8);
}
