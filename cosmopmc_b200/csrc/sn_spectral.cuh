// sn_spectral.cuh -- SN Ia likelihood, spectral form of the reference's quadrature (replaces nicaea SetDl + chi2_SN
// behind wrappers/src/sn.c:138-281 for large batches; the node-by-node kernel k_like_sn / k_like_sn_warp of
// cosmo.cuh stays the exact path).
//
// What the reference computes per (sample, redshift) is NR qromb on 1/sqrt(a^4 E^2) over [a_z, 1]: when it stops
// at stage 5 -- it always does on a smooth integrand, see below -- the result ss_z and its error estimate dss_z
// are two FIXED linear functionals of the integrand at 17 nodes (ROMBW, cosmo.cuh).  The 241 redshifts of the
// Union sample therefore ask for 4097 evaluations of ONE smooth function of a per sample, all inside
// [a(z_max), 1].  This kernel evaluates the sample-dependent factor q(a) = Q(a)^-1/2 (Q = a^3 E^2 / |Omega_de|)
// at SNS_M Chebyshev points of that interval, turns the values into Chebyshev coefficients c_m (an SNS_M-point DCT,
// folded), and applies the reference's functional to the interpolant:
//     ss_z = sum_m W[z][m] c_m,      W[z][m] = h_z sum_i w_i a_i^-1/2 T_m(x(a_i))   (host, long double)
// i.e. the Romberg rule -- nodes, weights, truncation error and all -- acts on a polynomial that agrees with the
// integrand to ~1e-15.  Per sample: SNS_M = 28 integrand evaluations + a 241 x 28 matrix-vector product instead of 4097
// evaluations (13 FP64 instructions each).
//
// Guarantees (per sample, checked in the kernel; a sample that fails ANY of them is handed to the exact kernel
// through a list, so the fast path never decides anything the reference would decide differently):
//   * interpolation: |c_(M-1)| + |c_(M-2)| + |c_(M-3)| <= SNS_TAIL_TOL |c_0|  (measured: the error of ss_z is
//     below 1e-13 relative whenever this holds, tools/proto/sn_spectral_proto.py)
//   * stopping rule: sum_m dmax_m |c_m| <= 0.25e-6 min_j q_j, with dmax_m = max_z |D[z][m]| / h_z the error
//     functional of stage 5 -- a sufficient condition for |dss_z| <= EPS |ss_z| at EVERY redshift, so the
//     reference stops at stage 5 everywhere and ss_z above is what it returns
//   * range: Omega_de > 0, exponent inside the table-based 2^s (the exact kernel has the general paths)
#pragma once
#include "cosmo.cuh"

#ifndef SNS_M
#define SNS_M 28       // Chebyshev points per sample (a multiple of 4)
#endif
#define SNS_TAIL_TOL 1.0e-12
#ifndef SNS_BLOCK
#define SNS_BLOCK 256
#endif
#ifndef SNS_MIN_BLOCKS
#define SNS_MIN_BLOCKS 2
#endif

// [m][j < M/2]: (2/M) cos(pi m (j + 1/2) / M), row 0 halved (filled by pmc_init_sn_tables, long double)
__constant__ double SNS_DCT[SNS_M * SNS_M / 2];

// chi^2 terms of the supernovae at redshift iz (the body of sn_zloop's inner loop)
__device__ __forceinline__ void sn_chi2_terms(const DevLike &L, const SNPer &m_, int mode, int iz, double mu_th,
                                              double &chi2, double &logdet) {
  const int i0 = __ldg(&L.first[iz]), i1 = __ldg(&L.first[iz + 1]);
  for (int i = i0; i < i1; i++) {
    const double2 *__restrict__ r = reinterpret_cast<const double2 *>(L.sn + (size_t)i * SN_ROW);
    const double2 ms = __ldg(&r[0]), cz = __ldg(&r[1]), w01 = __ldg(&r[2]), w23 = __ldg(&r[3]), w45 = __ldg(&r[4]);
    double mu_obs, sig2;
    if (mode == PMCB200_CHI2_betaz) {
      const double t2 = fma(m_.Theta3, cz.y, m_.t2base);
      mu_obs = ms.x + m_.Theta0 + m_.t1 * (ms.y - m_.stretch) + t2 * (cz.x - m_.color);
      sig2 = w01.x + m_.d1 * m_.d1 * w01.y + t2 * t2 * w23.x + 2.0 * (m_.d1 * w23.y + t2 * w45.x + m_.d1 * t2 * w45.y);
    } else {
      mu_obs = fma(m_.t1, ms.y, fma(m_.t2base, cz.x, ms.x + m_.base0));
      sig2 = fma(m_.k1, w01.y, fma(m_.k2, w23.x, fma(m_.k3, w23.y, fma(m_.k4, w45.x, fma(m_.k5, w45.y, w01.x)))));
    }
    const double res = mu_obs - mu_th;
    chi2 = fma(res * res, fast_rcp(sig2), chi2);
    if (L.sn_add_logdetCov) logdet += log(sig2);
  }
}

// Chebyshev coefficients c[0..M) of q(a) = Q(a)^-1/2 on [a(z_max), 1] and the smallest node value.  Four nodes from
// each end per step (u_j = q_j + q_(M-1-j) feeds the even coefficients, v_j = q_j - q_(M-1-j) the odd ones), so that
// only the 32 running sums and eight node values are live at a time.
template <bool HASQ, bool FLAT>
__device__ __forceinline__ void sn_cheb_coeffs(const DevLike &L, const SNCoef &ec, const double *__restrict__ T,
                                               double (&c)[SNS_M], double &qmin) {
#pragma unroll
  for (int mm = 0; mm < SNS_M; mm++) c[mm] = 0.0;
  qmin = INFINITY;
#pragma unroll
  for (int j0 = 0; j0 < SNS_M / 2; j0 += 4) {
    double u[4] = {0.0, 0.0, 0.0, 0.0}, v[4] = {0.0, 0.0, 0.0, 0.0};
#pragma unroll
    for (int i = 0; i < 4; i++) {
      if (j0 + i >= SNS_M / 2) continue;
      const Ld4 na = ld256(L.cheb_nodes4 + 4 * (j0 + i)), nb = ld256(L.cheb_nodes4 + 4 * (SNS_M - 1 - j0 - i));
      const double qa = sn_f<HASQ, FLAT, false, false>(ec, T, na.x, 1.0, na.z, 0.0);
      const double qb = sn_f<HASQ, FLAT, false, false>(ec, T, nb.x, 1.0, nb.z, 0.0);
      qmin = fmin(qmin, fmin(qa, qb));
      u[i] = qa + qb; v[i] = qa - qb;
    }
#pragma unroll
    for (int mm = 0; mm < SNS_M; mm++) {
#pragma unroll
      for (int i = 0; i < 4; i++)
        if (j0 + i < SNS_M / 2) c[mm] = fma(SNS_DCT[mm * (SNS_M / 2) + j0 + i], (mm & 1) ? v[i] : u[i], c[mm]);
    }
  }
}
// the two certificates of the header (a NaN anywhere fails them)
__device__ __forceinline__ bool sn_spec_certified(const DevLike &L, const double (&c)[SNS_M], double qmin) {
  const double tail = fabs(c[SNS_M - 1]) + fabs(c[SNS_M - 2]) + fabs(c[SNS_M - 3]);
  double B = 0.0;
#pragma unroll
  for (int mm = 0; mm < SNS_M; mm++) B = fma(__ldg(&L.cheb_dmax[mm]), fabs(c[mm]), B);
  return (tail <= SNS_TAIL_TOL * fabs(c[0])) && (B <= 0.25 * ROMB_EPS * qmin);
}

template <bool HASQ, bool FLAT>
__global__ void __launch_bounds__(SNS_BLOCK, SNS_MIN_BLOCKS)
k_like_sn_spec(const DevLike L, int64_t N, const double *__restrict__ X, int d,
               const int16_t *__restrict__ flg, double *__restrict__ logpi,
               int32_t *__restrict__ err, int set, double add_const, DevCount *cnt,
               uint32_t *__restrict__ fb_list, unsigned *__restrict__ fb_count) {
  __shared__ double T[96 + SN_EXP2_N];     // general tables + pre-biased 2^(j/1024)
  __shared__ double2 LT[LOG1K_N];          // lean_log's {1/c_i, -ln(1/c_i)}
  for (int i = threadIdx.x; i < LOG1K_N; i += blockDim.x) LT[i] = g_log1k[i];
  load_fast_tables_sn(T);
  const int64_t n = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const bool active = (n < N) && (!flg || flg[n]);
  Model m;
  int e = 0;
  if (active) e = apply_params(L, X + n * d, m);
  const bool cut = active && !e && L.special == PMCB200_SPECIAL_de_conservative && de_conservative_violated(m.c);
  if (!active || e) {   // benign model: the lane walks the loops, its result is not used
    m.c = L.model;
#pragma unroll
    for (int i = 0; i < 4; i++) m.Theta2[i] = L.Theta2[i];
    m.stretch = 1.0; m.color = 0.0;
  }
  SNCoef ec;
  SNPer pm;
  double f1, f1s;
  sn_setup(L, m, 0, ec, pm, f1, f1s);
  bool ok = !(ec.slow || ec.sgn != 0u);

  // --- Chebyshev coefficients of q(a) = Q(a)^-1/2 on [a(z_max), 1], and the two guarantees
  double c[SNS_M];
  {
    double qmin;
    sn_cheb_coeffs<HASQ, FLAT>(L, ec, T, c, qmin);
    if (!sn_spec_certified(L, c, qmin)) ok = false;
  }

  // --- redshift loop: ss_z = W[z] . c, then the distance modulus and the chi^2 terms as in sn_zloop
  const double rh = R_HUBBLE * ec.scale;
  const bool flat = fabs(ec.OK) < FLAT_EPS;
  const int mode = L.sn_chi2mode;
  double chi2 = 0.0, logdet = 0.0;
  const int nz = L.sn_nz;
  for (int iz = 0; iz < nz; iz++) {
    const double *__restrict__ w = L.cheb_W + (size_t)iz * SNS_M;
    double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
#pragma unroll
    for (int mm = 0; mm < SNS_M; mm += 4) {
      const Ld4 ww4 = ld256(w + mm);
      s0 = fma(ww4.x, c[mm], s0); s1 = fma(ww4.y, c[mm + 1], s1);
      s2 = fma(ww4.z, c[mm + 2], s2); s3 = fma(ww4.w, c[mm + 3], s3);
    }
    const double ss = (s0 + s1) + (s2 + s3);
    const double ww = rh * ss;
    const double fk = (FLAT || flat) ? ww : f_K_from(ec.OK, ww);
    if (!(fk > 0.0)) e = 1;
    const double lnaz = __ldg(&L.nodes[(size_t)iz * SN_NODES]).x;
    const double mu_th = fma(5.0 / M_LN10, lean_log(fk, LT) - lnaz, SN_MU0);
    sn_chi2_terms(L, pm, mode, iz, mu_th, chi2, logdet);
  }
  double res = -0.5 * chi2;
  if (L.sn_add_logdetCov) res -= 0.5 * logdet;
  if (cut) res = 0.0;
  else if (!isfinite(res)) e = 1;
  if (active && ok) put_loglike(logpi, err, n, set, add_const, e ? 0.0 : res, e);
  else if (active) fb_list[atomicAdd(fb_count, 1u)] = (uint32_t)n;      // the exact kernel evaluates this sample
  else if (n < N && set) { logpi[n] = 0.0; if (err) err[n] = 0; }
  if (cnt) {
    unsigned nsp = (active && ok) ? 1u : 0u;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) nsp += __shfl_xor_sync(0xffffffffu, nsp, o);
    if ((threadIdx.x & 31) == 0 && nsp) {
      atomicAdd(&cnt->sn_evals, (unsigned long long)nsp * SNS_M);
      atomicAdd(&cnt->sn_zsteps, (unsigned long long)nsp * nz);
      atomicAdd(&cnt->sn_spec, (unsigned long long)nsp);
    }
  }
}

// ---- the same on the FP64 tensor cores ---------------------------------------------------------------------------
// ncu on k_like_sn_spec (profiles/r02/sn_spec_v1_summary.txt): FP64 pipe 26 % busy, 9.5 long-scoreboard stall cycles
// per issue, LSU write-back 63 % -- the warp-uniform 32-byte loads that feed W[z][m] (and the supernova rows) to
// one-sample-per-thread DFMAs deliver 1 KB into the register file per instruction.  Here the three contractions of
// the redshift loop are mma.sync.m8n8k4.f64 tiles (8 samples x 8 supernovae x 4 terms), whose B operands are ONE
// double per lane from a fragment-ordered table (256 coalesced bytes per instruction, 8 DFMA-equivalents each, shared
// by the warp's four 8-sample row tiles):
//     ss[s][i]   = sum_m  c_s[m]              W[z(i)][m]                              (M/4 k-steps)
//     mu'[s][i]  = (1, base0, t1, t2)_s     . (m_i + 5/ln10 ln a_i - mu0, 1, s_i, c_i)     (1 k-step)
//     sig2[s][i] = (1, k1..k5, 0, 0)_s      . (V0, Vss, Vcc, Cms, Cmc, Csc, 0, 0)_i        (2 k-steps)
// (chi2_betaz, whose coefficients depend on the redshift, and add_logdetCov stay with k_like_sn_spec.)
// Columns are SUPERNOVAE (sorted by redshift; two supernovae at one redshift repeat its W row), so that ss, mu' and
// sig2 of a (sample, supernova) pair land in the same accumulator slot of the same lane and the chi^2 term is formed
// right there: res = mu' - 5/ln10 ln f_K(rh ss), chi2 += res^2 / sig2.  Phase 1 (one sample per lane: parameters,
// integrand at the Chebyshev points, DCT, certificates) is k_like_sn_spec's; the coefficients and the per-sample
// chi^2 constants reach the A fragments through a per-warp shared-memory transposition.
#ifndef SNS2_BLOCK
#define SNS2_BLOCK 256      // 2 warps per scheduler: the A fragments alone are 88 registers
#endif
#define SNS_KS (SNS_M / 4 + 3)        // k-steps per supernova tile: coefficients, mu', sig2 (2)
#define SNS_TRS 12                    // row stride of the transposition buffer (12 r mod 16 distinct for 4 rows)
#define SNS2_SMEM (sizeof(double2) * LOG1K_N + sizeof(double) * (96 + SN_EXP2_N + (SNS2_BLOCK / 32) * 32 * SNS_TRS))

__device__ __forceinline__ void sn_dmma(double &c0, double &c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
               : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}
// 1/s, s > 0: MUFU.RCP64H seed y (rel. error e < 2^-20), y (1 + e + e^2): error e^3
__device__ __forceinline__ double sn_rcp3(double s) {
  double y;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(s));
  const double e = fma(-s, y, 1.0);
  return fma(y, fma(e, e, e), y);
}

// ---- mixed-precision tail (template flag T32) ----------------------------------------------------------------------
// The Chebyshev coefficients of an analytic integrand decay geometrically: beyond m = SNS_M64 = 16 they are below
// ~1e-8 |c_0| for every sample that passes certificate (i), so their contribution to ss needs 5-6 digits, not 16.  With
// T32 the terms m >= 16 of ss = sum_m c_m W[z][m] leave the FP64 pipe: c_m and W are split into two TF32 numbers each
// (hi + lo carries 22 bits) and the three leading products hi*hi + hi*lo + lo*hi are accumulated by
// mma.sync.m16n8k8.tf32 with FP32 accumulators, whose accumulator layout (row lane/4 [+8], columns 2 (lane%4) [+1])
// coincides with the FP64 tiles' -- the tail lands in the lane that owns the pair and seeds its FP64 accumulator.
// 3 of the 7 coefficient k-steps (12 of 40 DMMA per tile) disappear.  Error: representation 3 * 2^-22 per product,
// accumulation of 48 terms in FP32 <= 48 * 2^-23, together below 8e-6 of sum_m |c_m W_m|; with |W[z][m]| <= 0.62 h_z and
// ss_z >= h_z min_j q_j the relative error of ss_z is below 5e-6 * S / min_j q_j, S = sum_(m>=16) |c_m|.
// Certificate (iii): S <= SNS_T32_TOL * min_j q_j  =>  relative error of ss_z below 2.5e-13 (worst case; measured
// against the node-by-node kernel over 1e7 samples: see DESIGN.md).  A sample that fails it goes to the exact kernel
// like the others; in practice (iii) is implied by (i), which already demands the same decay rate.
#define SNS_M64 16
#define SNS_T32_TOL 5.0e-8
__device__ __forceinline__ void sn_split_tf32(double v, uint32_t &hi, uint32_t &lo) {
  const float f = (float)v;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(hi) : "f"(f));
  const float r = (float)(v - (double)__uint_as_float(hi));
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(lo) : "f"(r));
}
__device__ __forceinline__ void sn_mma_tf32(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
struct Ld8u { uint32_t v[8]; };
__device__ __forceinline__ Ld8u ld256u(const void *p) {
  Ld8u r;
  asm("ld.global.nc.v8.u32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
      : "=r"(r.v[0]), "=r"(r.v[1]), "=r"(r.v[2]), "=r"(r.v[3]), "=r"(r.v[4]), "=r"(r.v[5]), "=r"(r.v[6]), "=r"(r.v[7]) : "l"(p));
  return r;
}

template <bool HASQ, bool FLAT, bool T32>
__global__ void __launch_bounds__(SNS2_BLOCK, 1)
k_like_sn_spec_mma(const DevLike L, int64_t N, const double *__restrict__ X, int d,
                   const int16_t *__restrict__ flg, double *__restrict__ logpi,
                   int32_t *__restrict__ err, int set, double add_const, DevCount *cnt,
                   uint32_t *__restrict__ fb_list, unsigned *__restrict__ fb_count) {
  extern __shared__ double2 sns2_smem[];      // SNS2_SMEM bytes: LT | T | per-warp transposition buffers
  double2 *LT = sns2_smem;
  double *T = reinterpret_cast<double *>(sns2_smem + LOG1K_N);
  for (int i = threadIdx.x; i < LOG1K_N; i += blockDim.x) LT[i] = g_log1k[i];
  load_fast_tables_sn(T);
  const int lane = threadIdx.x & 31;
  double *__restrict__ tr = T + (96 + SN_EXP2_N) + (threadIdx.x >> 5) * (32 * SNS_TRS);
  // Persistent block (one per SM: the tables are staged once), every warp walks its own sequence of 32-sample tasks;
  // nothing below is block-wide.  11.93 -> 11.5 ms per 1e7 samples.  (Measured and not adopted, session O / P: the second
  // warp of each scheduler started half a task late -- 11.56 / 11.45 / 11.42 ms for 0 / 20 / 40 us; phase 1 as a kernel of
  // its own at 12 warps per SM with the A fragments through HBM, tools/micro/sn_two_kernel_split_experiment.patch -- 0.29 +
  // 2.11 ms per 2e6 samples = 12.0 ms: the tile loop alone keeps the FP64 pipe 78 % busy with two warps per scheduler.)
  const int64_t ntask = (N + 31) / 32;
  for (int64_t task = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); task < ntask;      // blockDim <= SNS2_BLOCK:
       task += (int64_t)gridDim.x * (blockDim.x >> 5)) {                                                 // small batches launch fewer warps per block
  __syncwarp();
  const int64_t n = task * 32 + lane;
  const bool active = (n < N) && (!flg || flg[n]);
  Model m;
  int e = 0;
  if (active) e = apply_params(L, X + n * d, m);
  const bool cut = active && !e && L.special == PMCB200_SPECIAL_de_conservative && de_conservative_violated(m.c);
  if (!active || e) {
    m.c = L.model;
#pragma unroll
    for (int i = 0; i < 4; i++) m.Theta2[i] = L.Theta2[i];
    m.stretch = 1.0; m.color = 0.0;
  }
  bool ok;
  const int qrow = lane >> 2, qcol = lane & 3;
  // A fragments: [row tile][k-step] = value [sample 8 rt + lane/4][4 ks + lane%4]; k-steps: KC coefficient ones
  // (all SNS_M / 4, or the first SNS_M64 / 4 with the TF32 tail), then mu', sig2, sig2
  constexpr int KC = T32 ? SNS_M64 / 4 : SNS_M / 4;
  constexpr int NT32 = (SNS_M - SNS_M64 + 7) / 8;      // TF32 k-steps of 8 coefficients
  double A[4][KC + 3];
  uint32_t A32[2][NT32 > 0 ? NT32 : 1][2][4];            // [16-sample tile][k-step][hi, lo][a0..a3]
  double rhv[4], OKv[4];
  {
    SNCoef ec;
    SNPer pm;
    double f1, f1s;
    sn_setup(L, m, 0, ec, pm, f1, f1s);
    ok = !(ec.slow || ec.sgn != 0u);
    double c[SNS_M];
    {
      double qmin;
      sn_cheb_coeffs<HASQ, FLAT>(L, ec, T, c, qmin);
      if (!sn_spec_certified(L, c, qmin)) ok = false;
      if (T32) {      // certificate (iii)
        double S = 0.0;
#pragma unroll
        for (int mm = SNS_M64; mm < SNS_M; mm++) S += fabs(c[mm]);
        if (!(S <= SNS_T32_TOL * qmin)) ok = false;
      }
    }
    // lane-owned values -> A fragments, eight values per trip through the warp's buffer
#pragma unroll
    for (int ch = 0; ch < (SNS_M + 7) / 8; ch++) {
      __syncwarp();
#pragma unroll
      for (int j = 0; j < 8; j++) tr[lane * SNS_TRS + j] = (8 * ch + j < SNS_M) ? c[8 * ch + j] : 0.0;
      __syncwarp();
      if (!T32 || 8 * ch < SNS_M64) {
#pragma unroll
        for (int rt = 0; rt < 4; rt++) {
          A[rt][2 * ch] = tr[(8 * rt + qrow) * SNS_TRS + qcol];
          if (2 * ch + 1 < KC) A[rt][2 * ch + 1] = tr[(8 * rt + qrow) * SNS_TRS + 4 + qcol];
        }
      } else {      // TF32 tail: m16n8k8 A fragment a0 (row g, col t), a1 (row g + 8, col t), a2 (row g, col t + 4), a3 (row g + 8, col t + 4)
        const int k2 = (8 * ch - SNS_M64) / 8;
#pragma unroll
        for (int mt = 0; mt < 2; mt++) {
          const double v[4] = {tr[(16 * mt + qrow) * SNS_TRS + qcol], tr[(16 * mt + 8 + qrow) * SNS_TRS + qcol],
                               tr[(16 * mt + qrow) * SNS_TRS + 4 + qcol], tr[(16 * mt + 8 + qrow) * SNS_TRS + 4 + qcol]};
#pragma unroll
          for (int i = 0; i < 4; i++) sn_split_tf32(v[i], A32[mt][k2][0][i], A32[mt][k2][1][i]);
        }
      }
    }
    __syncwarp();
    // flat: ln f_K = ln rh + ln ss, the per-sample part joins the constant of mu'
    tr[lane * SNS_TRS + 0] = 1.0;
    tr[lane * SNS_TRS + 1] = FLAT ? fma(-5.0 / M_LN10, log(R_HUBBLE * ec.scale), pm.base0) : pm.base0;
    tr[lane * SNS_TRS + 2] = pm.t1;
    tr[lane * SNS_TRS + 3] = pm.t2base;
    tr[lane * SNS_TRS + 4] = 1.0; tr[lane * SNS_TRS + 5] = pm.k1; tr[lane * SNS_TRS + 6] = pm.k2;
    tr[lane * SNS_TRS + 7] = pm.k3;
    __syncwarp();
#pragma unroll
    for (int rt = 0; rt < 4; rt++) {
      A[rt][KC] = tr[(8 * rt + qrow) * SNS_TRS + qcol];
      A[rt][KC + 1] = tr[(8 * rt + qrow) * SNS_TRS + 4 + qcol];
    }
    __syncwarp();
    tr[lane * SNS_TRS + 0] = pm.k4; tr[lane * SNS_TRS + 1] = pm.k5; tr[lane * SNS_TRS + 2] = 0.0;
    tr[lane * SNS_TRS + 3] = 0.0;
    tr[lane * SNS_TRS + 4] = R_HUBBLE * ec.scale; tr[lane * SNS_TRS + 5] = (fabs(ec.OK) < FLAT_EPS) ? 0.0 : ec.OK;
    __syncwarp();
#pragma unroll
    for (int rt = 0; rt < 4; rt++) {
      A[rt][KC + 2] = tr[(8 * rt + qrow) * SNS_TRS + qcol];
      rhv[rt] = tr[(8 * rt + qrow) * SNS_TRS + 4];
      OKv[rt] = tr[(8 * rt + qrow) * SNS_TRS + 5];
    }
  }

  // --- supernova tiles.  A PRIMARY tile holds 8 distinct redshifts (its columns = the first supernova at each): all k-steps,
  // ln f_K per pair.  The Union sample has 307 supernovae at 241 redshifts; the further supernovae of a redshift sit in the
  // SAME column of a SECONDARY tile that follows its primary tile directly: ss, hence ln f_K, is the primary column's (same
  // lane, same accumulator slot, kept in lnf), so a secondary tile only runs the mu' and sig2 k-steps (3 of 10) and no
  // logarithm.  Empty columns carry sigma^2 = 1e300: their terms vanish.  31 primary + 12 secondary tiles instead of 39 full ones.
  const double *__restrict__ wf = L.cheb_Wf + lane;
  const int *__restrict__ tsec = L.sn_tile_sec;
  const int ntile = L.sn_ntile;
  double chi[4] = {0.0, 0.0, 0.0, 0.0};
  double lnf[4][2] = {{0.0, 0.0}, {0.0, 0.0}, {0.0, 0.0}, {0.0, 0.0}};
  unsigned ebits = 0u, ubits = 0u;      // per row tile: distance error; curvature argument outside the series
  // fragment table k-steps this kernel reads: the KC coefficient ones, then mu', sig2, sig2 (table positions M/4 ..)
  double b[KC + 3], bn[KC + 3];
  Ld8u bt, btn;      // TF32 tail fragments of the tile: [k-step][hi b0, hi b1, lo b0, lo b1]
  const uint32_t *__restrict__ wt = L.cheb_Wt + (size_t)lane * 8;
#pragma unroll
  for (int ks = 0; ks < KC + 3; ks++) bn[ks] = __ldg(wf + (size_t)(ks < KC ? ks : ks - KC + SNS_M / 4) * 32);
  if (T32) btn = ld256u(wt);
  int secn = 0;      // tile 0 is a primary tile
  for (int t = 0; t < ntile; t++) {
#pragma unroll
    for (int ks = 0; ks < KC + 3; ks++) b[ks] = bn[ks];
    if (T32) bt = btn;
    const int sec = secn;
    const int tn = min(t + 1, ntile - 1);       // next tile's fragments travel while this one computes
#pragma unroll
    for (int ks = 0; ks < KC + 3; ks++) bn[ks] = __ldg(wf + ((size_t)tn * SNS_KS + (ks < KC ? ks : ks - KC + SNS_M / 4)) * 32);
    if (T32) btn = ld256u(wt + (size_t)tn * 256);
    secn = __ldg(tsec + tn);
    if (sec) {      // further supernovae at the primary tile's redshifts: mu', sig2, and the stored ln f_K
#pragma unroll
      for (int rt = 0; rt < 4; rt++) {
        double m0 = 0.0, m1 = 0.0, g0 = 0.0, g1 = 0.0;
        sn_dmma(m0, m1, A[rt][KC], b[KC]);
        sn_dmma(g0, g1, A[rt][KC + 1], b[KC + 1]);
        sn_dmma(g0, g1, A[rt][KC + 2], b[KC + 2]);
#pragma unroll
        for (int h = 0; h < 2; h++) {
          const double mu = h ? m1 : m0, sg = h ? g1 : g0;
          const double res = fma(-5.0 / M_LN10, lnf[rt][h], mu);
          chi[rt] = fma(res * res, sn_rcp3(sg), chi[rt]);
        }
      }
      continue;
    }
    float t32[2][4] = {{0.f, 0.f, 0.f, 0.f}, {0.f, 0.f, 0.f, 0.f}};
    if (T32) {
#pragma unroll
      for (int mt = 0; mt < 2; mt++)
#pragma unroll
        for (int k2 = 0; k2 < NT32; k2++) {
          sn_mma_tf32(t32[mt], A32[mt][k2][1], bt.v[4 * k2], bt.v[4 * k2 + 1]);          // lo * hi
          sn_mma_tf32(t32[mt], A32[mt][k2][0], bt.v[4 * k2 + 2], bt.v[4 * k2 + 3]);      // hi * lo
          sn_mma_tf32(t32[mt], A32[mt][k2][0], bt.v[4 * k2], bt.v[4 * k2 + 1]);          // hi * hi
        }
    }
#pragma unroll
    for (int rt = 0; rt < 4; rt++) {
      // the tail (accumulator rows lane/4 and lane/4 + 8 of the 16-sample tile = row tiles 2 mt and 2 mt + 1) seeds ss
      double s0 = T32 ? (double)t32[rt >> 1][2 * (rt & 1)] : 0.0, s1 = T32 ? (double)t32[rt >> 1][2 * (rt & 1) + 1] : 0.0;
      double m0 = 0.0, m1 = 0.0, g0 = 0.0, g1 = 0.0;
#pragma unroll
      for (int ks = 0; ks < KC; ks++) sn_dmma(s0, s1, A[rt][ks], b[ks]);      // (two independent half chains: 11.73 against 11.34 ms)
      sn_dmma(m0, m1, A[rt][KC], b[KC]);
      sn_dmma(g0, g1, A[rt][KC + 1], b[KC + 1]);
      sn_dmma(g0, g1, A[rt][KC + 2], b[KC + 2]);
#pragma unroll
      for (int h = 0; h < 2; h++) {
        const double ss = h ? s1 : s0, mu = h ? m1 : m0, sg = h ? g1 : g0;
        double fk;
        if (FLAT) fk = ss;
        else {      // f_K_from without branches: the series covers |u| < 1, beyond it the sample goes to the exact kernel
          const double ww = rhv[rt] * ss, x = ww * (1.0 / R_HUBBLE), u = OKv[rt] * x * x;
          fk = ww * sinhc_series(u);      // OKv = 0 for |Omega_K| < FLAT_EPS: the series is exactly 1, f_K = w as in f_K_from
          if (!(fabs(u) < 1.0)) ubits |= 1u << rt;
        }
        if (!(fk > 0.0)) ebits |= 1u << rt;
        const double lg = lean_log(fk, LT);
        lnf[rt][h] = lg;
        const double res = fma(-5.0 / M_LN10, lg, mu);
        chi[rt] = fma(res * res, sn_rcp3(sg), chi[rt]);
      }
    }
  }
  // --- per-sample sums: the four lanes of a quad hold the columns of one sample row; then to the owner lane
  double chi2 = 0.0;
  ebits |= __shfl_xor_sync(0xffffffffu, ebits, 1);
  ebits |= __shfl_xor_sync(0xffffffffu, ebits, 2);
#pragma unroll
  for (int rt = 0; rt < 4; rt++) {
    double v = chi[rt];
    v += __shfl_xor_sync(0xffffffffu, v, 1); v += __shfl_xor_sync(0xffffffffu, v, 2);
    const double vo = __shfl_sync(0xffffffffu, v, 4 * (lane & 7));
    if ((lane >> 3) == rt) chi2 = vo;
  }
  {
    const unsigned eo = __shfl_sync(0xffffffffu, ebits, 4 * (lane & 7));
    if ((eo >> (lane >> 3)) & 1u) e = 1;
  }
  if (!FLAT) {
    ubits |= __shfl_xor_sync(0xffffffffu, ubits, 1);
    ubits |= __shfl_xor_sync(0xffffffffu, ubits, 2);
    const unsigned uo = __shfl_sync(0xffffffffu, ubits, 4 * (lane & 7));
    if ((uo >> (lane >> 3)) & 1u) ok = false;
  }
  double res = -0.5 * chi2;
  if (cut) res = 0.0;
  else if (!isfinite(res)) e = 1;
  if (active && ok) put_loglike(logpi, err, n, set, add_const, e ? 0.0 : res, e);
  else if (active) fb_list[atomicAdd(fb_count, 1u)] = (uint32_t)n;      // the exact kernel evaluates this sample
  else if (n < N && set) { logpi[n] = 0.0; if (err) err[n] = 0; }
  if (cnt) {
    unsigned nsp = (active && ok) ? 1u : 0u;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) nsp += __shfl_xor_sync(0xffffffffu, nsp, o);
    if (lane == 0 && nsp) {
      atomicAdd(&cnt->sn_evals, (unsigned long long)nsp * SNS_M);
      atomicAdd(&cnt->sn_zsteps, (unsigned long long)nsp * L.sn_nz);
      atomicAdd(&cnt->sn_spec, (unsigned long long)nsp);
    }
  }
  }      // task loop
}
