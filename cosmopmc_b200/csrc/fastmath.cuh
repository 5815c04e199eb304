// fastmath.cuh -- table-based FP64 primitives shared by the likelihood kernels (cosmo.cuh) and the sampler
// (pmc_kernels.cuh): 2^s, ln x, 1/sqrt(v), 1/s with MUFU seeds, and the loaders of their shared-memory tables.
#pragma once
#include "common.cuh"

// ---- fast FP64 primitives (SN hot loop and the BAO / CMB integrals) ---------------------------------
// 2^s for |s| < 1000: s = k/32 + f with |f| <= 1/64 (one magic-number add, the
// remainder is exact), 2^f by a degree-5 near-minimax polynomial (Chebyshev
// interpolant, max rel. error 1.4e-16 before rounding), 2^(j/32) from a
// 32-entry shared-memory table, power of two by exponent-field addition.
// Branch-free: 9 FP64 ops + 1 LDS.
// Constants live in constant memory so DFMA takes them as c[bank][offset]
// operands (no per-use UMOV/IMAD materialisation).
__constant__ double EXP2C[8] = {
    0x1.62e42fefa39efp-1, 0x1.ebfbdff7fee6fp-3, 0x1.c6b08d70380bfp-5, 0x1.3b2b30255298ap-7,
    0x1.5d885e73db266p-10,
    211106232532992.0,              // [5] 1.5 * 2^47: ulp = 2^-5
    0.0, 0.0};
__constant__ double EXP2T[32] = {
    0x1.0000000000000p+0, 0x1.059b0d3158574p+0, 0x1.0b5586cf9890fp+0, 0x1.11301d0125b51p+0, 0x1.172b83c7d517bp+0,
    0x1.1d4873168b9aap+0, 0x1.2387a6e756238p+0, 0x1.29e9df51fdee1p+0, 0x1.306fe0a31b715p+0, 0x1.371a7373aa9cbp+0,
    0x1.3dea64c123422p+0, 0x1.44e086061892dp+0, 0x1.4bfdad5362a27p+0, 0x1.5342b569d4f82p+0, 0x1.5ab07dd485429p+0,
    0x1.6247eb03a5585p+0, 0x1.6a09e667f3bcdp+0, 0x1.71f75e8ec5f74p+0, 0x1.7a11473eb0187p+0, 0x1.82589994cce13p+0,
    0x1.8ace5422aa0dbp+0, 0x1.93737b0cdc5e5p+0, 0x1.9c49182a3f090p+0, 0x1.a5503b23e255dp+0, 0x1.ae89f995ad3adp+0,
    0x1.b7f76f2fb5e47p+0, 0x1.c199bdd85529cp+0, 0x1.cb720dcef9069p+0, 0x1.d5818dcfba487p+0, 0x1.dfc97337b9b5fp+0,
    0x1.ea4afa2a490dap+0, 0x1.f50765b6e4540p+0};
// ln(x) for positive normal x: x = 2^e m, m in [1,2); the top five mantissa bits
// pick c_i = 1 + (i + 1/2)/32, r = m/c_i - 1 (|r| < 1/64, one FMA with the tabulated
// rounded reciprocal), ln m = ln c_i + log1p(r) with a degree-8 Taylor polynomial
// (abs. error 9e-18).  12 FP64 ops instead of libdevice's ~28.
__constant__ double LOGRC[32] = {
    0x1.f81f81f81f820p-1, 0x1.e9131abf0b767p-1, 0x1.dae6076b981dbp-1, 0x1.cd85689039b0bp-1, 0x1.c0e070381c0e0p-1,
    0x1.b4e81b4e81b4fp-1, 0x1.a98ef606a63bep-1, 0x1.9ec8e951033d9p-1, 0x1.948b0fcd6e9e0p-1, 0x1.8acb90f6bf3aap-1,
    0x1.8181818181818p-1, 0x1.78a4c8178a4c8p-1, 0x1.702e05c0b8170p-1, 0x1.6816816816817p-1, 0x1.6058160581606p-1,
    0x1.58ed2308158edp-1, 0x1.51d07eae2f815p-1, 0x1.4afd6a052bf5bp-1, 0x1.446f86562d9fbp-1, 0x1.3e22cbce4a902p-1,
    0x1.3813813813814p-1, 0x1.323e34a2b10bfp-1, 0x1.2c9fb4d812ca0p-1, 0x1.27350b8812735p-1, 0x1.21fb78121fb78p-1,
    0x1.1cf06ada2811dp-1, 0x1.1811811811812p-1, 0x1.135c81135c811p-1, 0x1.0ecf56be69c90p-1, 0x1.0a6810a6810a7p-1,
    0x1.0624dd2f1a9fcp-1, 0x1.0204081020408p-1};
__constant__ double LOGLC[32] = {   // -ln(LOGRC[i]) of the ROUNDED reciprocals
    0x1.fc0a8b0fc03c4p-7, 0x1.77458f632dcffp-5, 0x1.341d7961bd1d0p-4, 0x1.a926d3a4ad562p-4, 0x1.0d77e7cd08e5bp-3,
    0x1.44d2b6ccb7d1cp-3, 0x1.7ab890210d907p-3, 0x1.af3c94e80bff3p-3, 0x1.e27076e2af2e8p-3, 0x1.0a324e27390e2p-2,
    0x1.22941fbcf7966p-2, 0x1.3a64c556945eap-2, 0x1.51aad872df82ep-2, 0x1.686c81e9b14adp-2, 0x1.7eaf83b82afc2p-2,
    0x1.947941c2116fbp-2, 0x1.a9cec9a9a084ap-2, 0x1.beb4d9da71b7ap-2, 0x1.d32fe7e00ebd5p-2, 0x1.e744261d68789p-2,
    0x1.faf588f78f31dp-2, 0x1.0723e5c1cdf41p-1, 0x1.109f39e2d4c96p-1, 0x1.19ee6b467c96fp-1, 0x1.23130d7bebf43p-1,
    0x1.2c0e9ed448e8cp-1, 0x1.34e289d9ce1d2p-1, 0x1.3d9026a7156fbp-1, 0x1.4618bc21c5ec2p-1, 0x1.4e7d811b75bb0p-1,
    0x1.56bf9d5b3f399p-1, 0x1.5ee02a9241676p-1};
__constant__ double LOGP[8] = {-1.0 / 8.0, 1.0 / 7.0, -1.0 / 6.0, 1.0 / 5.0, -1.0 / 4.0, 1.0 / 3.0, 0.693147180559945309417, 0.0};
// T: shared table [32 exp2 | 32 LOGRC | 32 LOGLC]
__device__ __forceinline__ double fast_log(double x, const double *__restrict__ T) {
  const int hi = __double2hiint(x);
  const int i = (hi >> 15) & 31;
  const double m = __hiloint2double((hi & 0x000fffff) | 0x3ff00000, __double2loint(x));
  const double r = fma(m, T[32 + i], -1.0);
  double q = LOGP[0];
  q = fma(q, r, LOGP[1]);
  q = fma(q, r, LOGP[2]);
  q = fma(q, r, LOGP[3]);
  q = fma(q, r, LOGP[4]);
  q = fma(q, r, LOGP[5]);
  q = fma(q, r, -0.5);
  q = fma(q, r, 1.0);
  const double e = (double)((hi >> 20) - 1023);
  return fma(e, LOGP[6], fma(q, r, T[64 + i]));
}
// returns sign * 2^s, sign given as an XOR mask for the high word
__device__ __forceinline__ double fast_exp2_signed(double s, const double *__restrict__ T, unsigned sgn) {
  double kf = s + EXP2C[5];
  const int k32 = __double2loint(kf);
  kf -= EXP2C[5];
  const double f = s - kf;
  double p = EXP2C[4];
  p = fma(p, f, EXP2C[3]);
  p = fma(p, f, EXP2C[2]);
  p = fma(p, f, EXP2C[1]);
  p = fma(p, f, EXP2C[0]);
  p = fma(p, f, 1.0);
  p *= T[k32 & 31];
  return __hiloint2double((__double2hiint(p) + ((k32 >> 5) << 20)) ^ sgn, __double2loint(p));
}
// 1/sqrt(v): MUFU.RSQ64H seed (rel. error 2^-22) + one third-order step -> 2^-66.
// v < 0 -> NaN, v = 0 -> NaN (both are the reference's error condition).
__device__ __forceinline__ double fast_rsqrt(double v) {
  double y;
  asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(v));
  double t = v * y;
  double e = fma(-t, y, 1.0);
  double c = fma(0.375, e, 0.5);
  return fma(y * e, c, y);
}
// 1/s for s > 0: MUFU.RCP64H seed + two Newton steps
__device__ __forceinline__ double fast_rcp(double s) {
  double y;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(s));
  double e = fma(-s, y, 1.0);
  y = fma(y, e, y);
  e = fma(-s, y, 1.0);
  return fma(y, e, y);
}

// load the shared table [32 exp2 | 32 log reciprocals | 32 log offsets]; blockDim.x >= 96
__device__ __forceinline__ void load_fast_tables(double *T) {
  for (int i = threadIdx.x; i < 96; i += blockDim.x)      // any block size (the SN tensor-core kernel launches 32 .. 256 threads)
    T[i] = i < 32 ? EXP2T[i] : (i < 64 ? LOGRC[i - 32] : LOGLC[i - 64]);
  __syncthreads();
}
