// common.cuh -- shared device helpers: Philox4x32-10, warp/block reductions,
// the packed device layout of the Gaussian/Student-t mixture, and the
// Romberg quadrature that the distance integrals use.  sm_100a only.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <math.h>
#include <type_traits>
#include "../../include/pmcb200.h"

#define PMC_BLOCK 256
#define LN2PI 1.8378770664093454836
#define R_HUBBLE 2997.92458
#define C_KMS 299792.458
#define ROMB_EPS 1.0e-6
#define ROMB_JMAX 20
#define SN_H_FID 0.7
#define DE_A_ACC (2.0 / 3.0)   // a_acc of the de_conservative prior (param.c:1091)
#define SN_MU0 25.774509799928715845   // 25 - 5 log10(SN_H_FID); a literal: nvcc does not fold log10() of a constant
#define FLAT_EPS 1.0e-8
#define OMEGA_GAMMA_H2 2.469e-5
#define NEFF_NU 3.04

// ---- packed mixture ---------------------------------------------------------
// One contiguous device buffer of doubles (mirrors pmclib keeping a mix_mvdens
// in one lump), laid out for the PADDED dimension D = pmc_pad_dim(d)
// so that every offset is a compile-time constant in the D-templated kernels
// (padded coordinates: mean 0, L = identity, so they contribute nothing):
// for component k at mix + k*stride
//   [0] wght   [1] lognorm = log-pdf constant (-d/2 ln2pi - log det L, or the
//   Student-t lgamma form)   [2 .. 2+D) mean   [2+D .. 2+D+T) lower Cholesky
//   factor packed by rows (row i holds L[i][0..i]), T = D(D+1)/2
//   [2+D+T .. 2+2D+T) 1/L[i][i]
// followed after K*stride by the common EM pivot p[D].
struct MixHdr {
  int K, d, df, stride, tri;
};
__host__ __device__ inline int mix_tri(int d) { return d * (d + 1) / 2; }

// padded template dimension for a runtime dimension d
__host__ __device__ inline int pmc_pad_dim(int d) {
  const int list[13] = {2, 3, 4, 5, 6, 7, 8, 10, 12, 16, 20, 24, 32};
  for (int i = 0; i < 13; i++) if (d <= list[i]) return list[i];
  return 32;
}

// stride of one packed component for runtime dimension d (layout uses D = pmc_pad_dim(d))
__host__ __device__ inline int mix_stride(int d) { const int D = pmc_pad_dim(d); return 2 + 2 * D + mix_tri(D); }

// ---- per-iteration device scalars --------------------------------------------
struct DevScal {
  unsigned long long max_key;   // order-preserving key of max log w (0 = none)
  unsigned long long nok;       // samples with finite weight
  unsigned long long nok_box;   // samples inside the box
  unsigned long long pad;
};
// measurement counters (never reset by the iteration itself)
struct DevCount {
  unsigned long long sn_evals;  // SN integrand evaluations
  unsigned long long sn_zsteps; // (sample, redshift) pairs integrated
  unsigned long long gen_evals;      // integrand evaluations of the BAO / CMB kernels (on-the-fly nodes)
  unsigned long long gen_integrals;  // their integrals
  unsigned long long sn_spec;        // SN samples evaluated by the spectral kernel (k_like_sn_spec)
  unsigned long long sn_exact;       // SN samples evaluated by the exact warp-per-sample kernel
  unsigned long long cmb_spec;       // CMB samples whose distance to a* came from the spectral form (cmb_spec_w)
};

__device__ __forceinline__ unsigned long long dkey(double x) {
  unsigned long long b = (unsigned long long)__double_as_longlong(x);
  return (b & 0x8000000000000000ull) ? ~b : (b | 0x8000000000000000ull);
}
__device__ __forceinline__ double dunkey(unsigned long long k) {
  if (k == 0ull) return -INFINITY;
  unsigned long long b = (k & 0x8000000000000000ull) ? (k & 0x7fffffffffffffffull) : ~k;
  return __longlong_as_double((long long)b);
}

// ---- Philox4x32-10 (Salmon et al. 2011) ---------------------------------------
__device__ __forceinline__ void philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3,
                                              uint32_t k0, uint32_t k1, uint32_t out[4]) {
#pragma unroll
  for (int r = 0; r < 10; r++) {
    uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
    uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
    uint32_t n0 = hi1 ^ c1 ^ k0, n2 = hi0 ^ c3 ^ k1;
    c0 = n0; c1 = lo1; c2 = n2; c3 = lo0;
    k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
  }
  out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}
__device__ __forceinline__ double u53(uint32_t hi, uint32_t lo) {   // (0,1]
  unsigned long long b = (((unsigned long long)hi << 32) | lo) >> 11;
  return ((double)b + 1.0) * (1.0 / 9007199254740992.0);
}

// ---- reductions ---------------------------------------------------------------
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ double warp_max(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
// block-wide sum in a fixed order (warp shuffles, then warp 0 over the warp
// partials): deterministic for a given block size.  red: >= 32 doubles smem.
__device__ __forceinline__ double block_sum(double v, double *red) {
  int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
  v = warp_sum(v);
  __syncthreads();
  if (lane == 0) red[w] = v;
  __syncthreads();
  double r = (threadIdx.x < nw) ? red[threadIdx.x] : 0.0;
  if (w == 0) r = warp_sum(r);
  if (threadIdx.x == 0) red[0] = r;
  __syncthreads();
  return red[0];
}

// ---- component selection (bit-exact mirror of the reference's inverse-CDF
// scan; oracle orc_select_component) ------------------------------------------
__device__ __forceinline__ int select_component(const double *__restrict__ mix, const MixHdr h,
                                                double u) {
  double cw = 0.0;
  int last = 0;
  for (int k = 0; k < h.K; k++) {
    double w = mix[(size_t)k * h.stride];
    if (w == 0.0) continue;
    cw = __dadd_rn(cw, w);
    last = k;
    if (u < cw) return k;
  }
  return last;
}

// ---- Gaussian / Student-t component log-pdf: forward substitution
// y = L^-1 (x - mu), m = y.y (the "batched triangular contraction") ------------
template <int D>
__host__ __device__ __forceinline__ double comp_maha(const double *__restrict__ comp, int d,
                                            const double (&x)[D], double (&y)[D]) {
  (void)d;       // the packed layout is padded to D: all offsets are compile-time constants
  const double *mean = comp + 2, *L = comp + 2 + D, *rd = comp + 2 + D + D * (D + 1) / 2;
  double m = 0.0;
#pragma unroll
  for (int i = 0; i < D; i++) {
    double t = x[i] - mean[i];
#pragma unroll
    for (int k = 0; k < i; k++) t = fma(-L[i * (i + 1) / 2 + k], y[k], t);
    y[i] = t * rd[i];
    m = fma(y[i], y[i], m);
  }
  return m;
}
__host__ __device__ __forceinline__ double comp_logpdf_from_maha(const double *__restrict__ comp, int d,
                                                        int df, double m) {
  if (df <= 0) return fma(-0.5, m, comp[1]);
  return comp[1] - 0.5 * (double)(df + d) * log1p(m / (double)df);
}

// log q(x) = log sum_k alpha_k exp(log phi_k(x)), NO max shift (reference
// semantics)
template <int D>
__device__ __forceinline__ double mix_logpdf(const double *__restrict__ mix, const MixHdr h,
                                             const double (&x)[D]) {
  double s = 0.0, y[D];
  for (int k = 0; k < h.K; k++) {
    const double *comp = mix + (size_t)k * h.stride;
    double w = comp[0];
    if (w == 0.0) continue;
    double m = comp_maha<D>(comp, h.d, x, y);
    s = fma(w, exp(comp_logpdf_from_maha(comp, h.d, h.df, m)), s);
  }
  return log(s);
}

// Compile-time loop: the body sees its index as a constant, so the register arrays it indexes can
// never be demoted to local memory (nvcc gave up unrolling the triangular nest with #pragma unroll).
template <int I, int N, class F>
__host__ __device__ __forceinline__ void static_for(F &&f) {
  if constexpr (I < N) {
    f(std::integral_constant<int, I>{});
    static_for<I + 1, N>(f);
  }
}
// S samples per thread, column-oriented: after y_k = t_k / L_kk every remaining row takes its update
// t_i -= L_ik y_k at once.  Each t_i still receives its updates in ascending k and m its squares in
// ascending i, so the result is bit-identical to comp_maha; but the D - k - 1 updates of a step are
// independent (no serial FMA chain), and each L_ik fetched (a warp-uniform load: the LSU data pipe, not
// the FP64 pipe, bounds these kernels -- tools/micro/dmma_probe.cu) feeds S FMAs.
// t[s][i] holds x_i of sample s on entry and is destroyed.
template <int D, int S>
__host__ __device__ __forceinline__ void comp_maha_cols(const double *__restrict__ comp, double (&t)[S][D], double (&m)[S]) {
  const double *mean = comp + 2, *L = comp + 2 + D, *rd = comp + 2 + D + D * (D + 1) / 2;
  static_for<0, D>([&](auto ii) {
    constexpr int i = decltype(ii)::value;
    const double mu = mean[i];
#pragma unroll
    for (int s = 0; s < S; s++) t[s][i] -= mu;
  });
#pragma unroll
  for (int s = 0; s < S; s++) m[s] = 0.0;
  static_for<0, D>([&](auto kk) {
    constexpr int k = decltype(kk)::value;
    const double r = rd[k];
    double y[S];
#pragma unroll
    for (int s = 0; s < S; s++) { y[s] = t[s][k] * r; m[s] = fma(y[s], y[s], m[s]); }
    static_for<k + 1, D>([&](auto ii) {
      constexpr int i = decltype(ii)::value;
      const double l = L[i * (i + 1) / 2 + k];
#pragma unroll
      for (int s = 0; s < S; s++) t[s][i] = fma(-l, y[s], t[s][i]);
    });
  });
}
// Column-packed copy of one component for the staged E-step (built by cp_fill_component): every section
// starts on a 16-byte boundary so that two doubles travel per shared-memory load (LDS.128) --
//   [0] wght  [1] lognorm | mean[Dp] | 1/L_ii [Dp] | for k = 0 .. D-2: column k of L below the diagonal,
//   L[k+1..D-1][k], padded to an even length;  Dp = D rounded up to even.
template <int D> struct CpLayout {
  static constexpr int Dp = (D + 1) & ~1;
  static constexpr int o_mean = 2, o_rd = 2 + Dp, o_L = 2 + 2 * Dp;
  __host__ __device__ static constexpr int coff(int k) {      // offset of column k inside the L section
    int o = 0;
    for (int j = 0; j < k; j++) o += ((D - 1 - j) + 1) & ~1;
    return o;
  }
  static constexpr int stride = o_L + coff(D - 1);            // column D-1 is empty
};
// dst (CpLayout<D>::stride doubles, zeroed by the caller: padding must be finite) from the row-packed component src
template <int D>
__host__ __device__ inline void cp_fill_component(const double *__restrict__ src, double *__restrict__ dst, int lane,
                                                  int nlanes) {
  using CL = CpLayout<D>;
  const double *mean = src + 2, *L = src + 2 + D, *rd = src + 2 + D + D * (D + 1) / 2;
  if (lane == 0) { dst[0] = src[0]; dst[1] = src[1]; }
  for (int i = lane; i < D; i += nlanes) { dst[CL::o_mean + i] = mean[i]; dst[CL::o_rd + i] = rd[i]; }
  for (int e = lane; e < D * (D + 1) / 2; e += nlanes) {      // e = i (i + 1) / 2 + k, k <= i
    int i = 0;
    while ((i + 1) * (i + 2) / 2 <= e) i++;
    const int k = e - i * (i + 1) / 2;
    if (k < i) dst[CL::o_L + CL::coff(k) + (i - k - 1)] = L[e];
  }
}
// comp_maha_cols on the column-packed copy: identical operations in identical order (bit-identical result),
// half the shared-memory load instructions.
template <int D, int S>
__host__ __device__ __forceinline__ void comp_maha_cols_cp(const double *__restrict__ cp, double (&t)[S][D],
                                                           double (&m)[S]) {
  using CL = CpLayout<D>;
  const double2 *mean2 = reinterpret_cast<const double2 *>(cp + CL::o_mean);
  const double *rd = cp + CL::o_rd;
  static_for<0, CL::Dp / 2>([&](auto jj) {
    constexpr int i = 2 * decltype(jj)::value;
    const double2 mu = mean2[i / 2];
#pragma unroll
    for (int s = 0; s < S; s++) {
      t[s][i] -= mu.x;
      if constexpr (i + 1 < D) t[s][i + 1] -= mu.y;
    }
  });
#pragma unroll
  for (int s = 0; s < S; s++) m[s] = 0.0;
  static_for<0, D>([&](auto kk) {
    constexpr int k = decltype(kk)::value;
    const double r = rd[k];
    double y[S];
#pragma unroll
    for (int s = 0; s < S; s++) { y[s] = t[s][k] * r; m[s] = fma(y[s], y[s], m[s]); }
    constexpr int n = D - 1 - k;                               // rows below the diagonal
    const double2 *col = reinterpret_cast<const double2 *>(cp + CL::o_L + CL::coff(k));
    static_for<0, (n + 1) / 2>([&](auto jj) {
      constexpr int j = decltype(jj)::value;
      const double2 l = col[j];
      constexpr int i0 = k + 1 + 2 * j;
#pragma unroll
      for (int s = 0; s < S; s++) {
        t[s][i0] = fma(-l.x, y[s], t[s][i0]);
        if constexpr (i0 + 1 < D) t[s][i0 + 1] = fma(-l.y, y[s], t[s][i0 + 1]);
      }
    });
  });
}

// ---- Romberg (Numerical Recipes qromb, K = 5) ---------------------------------
// Window y[0..4] of the last five trapezoid values (step ratio 1/4): Neville at
// h = 0 with the constant ratios folded in.  Returns ss, sets dss.
__device__ __forceinline__ double romb_extrap(const double (&y)[5], double &dss) {
  double c[5], d[5];
#pragma unroll
  for (int i = 0; i < 5; i++) { c[i] = y[i]; d[i] = y[i]; }
  double ss = y[4], dy = 0.0;
  const double r1[4] = {1.0 / 3.0, 1.0 / 15.0, 1.0 / 63.0, 1.0 / 255.0};
  const double r4[4] = {4.0 / 3.0, 16.0 / 15.0, 64.0 / 63.0, 256.0 / 255.0};
#pragma unroll
  for (int m = 1; m < 5; m++) {
#pragma unroll
    for (int i = 0; i < 5 - m; i++) {
      double w = c[i + 1] - d[i];
      d[i] = w * r1[m - 1];
      c[i] = w * r4[m - 1];
    }
    dy = d[4 - m];
    ss += dy;
  }
  dss = dy;
  return ss;
}

// Generic adaptive Romberg of f over [a,b] with on-the-fly nodes (used for the
// few BAO/CMB integrals per sample; the SN kernel has its own tabulated-node
// version).  err is set on non-finite result or too many stages.
template <class F>
__device__ double romberg(F f, double a, double b, int &err) {
  double y[5];
  double st = 0.5 * (b - a) * (f(a) + f(b));
  y[0] = st;
  double ss = st, dss;
  for (int j = 1; j < ROMB_JMAX; j++) {
    long it = 1L << (j - 1);
    double tnm = (double)it, del = (b - a) / tnm, sum = 0.0;
    for (long i = 0; i < it; i++) sum += f(fma((double)i + 0.5, del, a));
    st = 0.5 * (st + (b - a) * sum / tnm);
    if (j < 5) y[j] = st;
    else { y[0] = y[1]; y[1] = y[2]; y[2] = y[3]; y[3] = y[4]; y[4] = st; }
    if (j >= 4) {
      ss = romb_extrap(y, dss);
      if (!isfinite(ss)) { err = 1; return ss; }
      if (fabs(dss) <= ROMB_EPS * fabs(ss)) return ss;
    }
  }
  err = 1;
  return ss;
}
