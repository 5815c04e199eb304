// k_mix.cu -- instantiates the dimension-templated kernels (sampler, mixture
// log-pdf, weights, EM statistics) for one group of padded dimensions.
// Compiled four times with -DMIX_GROUP=0..3 (parallel build).
#include "pmc_kernels.cuh"
#include "launch.h"
#include <algorithm>
#include <cstdlib>
#ifndef ESTEP_DEFAULT
#define ESTEP_DEFAULT 2
#endif

#if MIX_GROUP == 0
#define DLIST(X) X(2) X(3) X(4) X(5)
#define FN pmc_mix_launch_g0
#elif MIX_GROUP == 1
#define DLIST(X) X(6) X(7) X(8) X(10)
#define FN pmc_mix_launch_g1
#elif MIX_GROUP == 2
#define DLIST(X) X(12) X(16) X(20)
#define FN pmc_mix_launch_g2
#else
#define DLIST(X) X(24) X(32)
#define FN pmc_mix_launch_g3
#endif

static inline int nblk(int64_t N) { return (int)((N + PMC_BLOCK - 1) / PMC_BLOCK); }

// FP64 tensor-core EM statistics (d >= 10): persistent grid sized by the occupancy API
template <int DD, int MT, bool STUDENT, bool RHO>
static cudaError_t launch_em_mma(const MixArgs &a, cudaStream_t s) {
  auto kern = k_em_stats_mma<DD, MT, STUDENT, RHO>;
  const size_t smem = em_mma_smem_bytes(a.h.K, a.h.d, STUDENT);
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  int per_sm = 1, dev = 0, sms = 148;
  cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, PMC_BLOCK, smem);
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  int blocks = std::max(1, std::min(a.blocks, std::max(1, per_sm) * sms));
  if (a.nblocks_out) *a.nblocks_out = blocks;
  kern<<<blocks, PMC_BLOCK, smem, s>>>(a.mix, a.h, a.N, a.Xc, a.idxc, a.flgc, a.logwc, a.scal, a.partials, a.linear,
                                       a.rho_in);
  return cudaGetLastError();
}
// warp-sliced form (d <= 10: few feature tiles): samples sliced over the warps, up to 4 resident blocks per SM
template <int DD, int MT, bool STUDENT, bool RHO>
static cudaError_t launch_em_ws(const MixArgs &a, cudaStream_t s) {
  auto kern = k_em_stats_mma_ws<DD, MT, STUDENT, RHO>;
  const size_t smem = em_ws_smem_bytes(a.h.K, a.h.d, STUDENT);
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  int per_sm = 1, dev = 0, sms = 148;
  cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, PMC_BLOCK, smem);
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const int64_t nsteps = (a.N + 31) / 32, want = (nsteps + PMC_BLOCK / 32 - 1) / (PMC_BLOCK / 32);
  int blocks = (int)std::max<int64_t>(1, std::min<int64_t>(std::min<int64_t>(a.blocks, want), (int64_t)std::max(1, per_sm) * sms));
  if (a.nblocks_out) *a.nblocks_out = blocks;
  kern<<<blocks, PMC_BLOCK, smem, s>>>(a.mix, a.h, a.N, a.Xc, a.idxc, a.flgc, a.logwc, a.scal, a.partials, a.linear,
                                       a.rho_in);
  return cudaGetLastError();
}
template <int DD>
static cudaError_t run_em_mma(const MixArgs &a, cudaStream_t s) {
  if constexpr (em_ws_nt(DD) <= 9) {
    const char *ew = getenv("PMCB200_EM_NO_WS");      // A/B measurements, cross-check in the tests; read per call
    const int no_ws = ew && *ew && *ew != '0';
    const int mt = em_mma_mt(a.h.K);
    const bool st = a.h.df > 0;
    if (!no_ws && em_ws_ok(a.h.K, a.h.d, st)) {
#define EMW(MTV) if (mt == MTV) { if constexpr (MTV * em_ws_nt(DD) <= 24) \
      { if (a.rho_in && !st) return launch_em_ws<DD, MTV, false, true>(a, s); \
        return st ? launch_em_ws<DD, MTV, true, false>(a, s) : launch_em_ws<DD, MTV, false, false>(a, s); } }
      EMW(1) EMW(2) EMW(3) EMW(4)
#undef EMW
    }
  }
  {
    const int mt = em_mma_mt(a.h.K);
    const bool st = a.h.df > 0;
#define EMM(MTV) if (mt == MTV) { if constexpr (MTV * (((1 + DD + DD * (DD + 1) / 2 + 7) / 8 + 7) / 8) <= EM_MMA_MAXACC) \
      { if (a.rho_in && !st) return launch_em_mma<DD, MTV, false, true>(a, s); \
        return st ? launch_em_mma<DD, MTV, true, false>(a, s) : launch_em_mma<DD, MTV, false, false>(a, s); } }
    EMM(1) EMM(2) EMM(3) EMM(4)
#undef EMM
  }
  return cudaErrorInvalidValue;
}

template <int DD>
static cudaError_t run(int op, const MixArgs &a, cudaStream_t s) {
  switch (op) {
    case OP_SIMULATE: {
      // staged kernel where the Cholesky-factor gather weighs (measured per 1e7 samples: d = 20 2.56 -> 2.22 ms,
      // d = 5 0.47 -> 0.55 ms, where the Philox / Box-Muller arithmetic dominates); PMCB200_SIM_STAGED=0/1 forces
      const char *es = getenv("PMCB200_SIM_STAGED");      // read per call
      const int staged_env = es && *es ? atoi(es) : -1;
      const int staged = staged_env >= 0 ? staged_env : (DD >= 10);
      const size_t sm = ((size_t)PMC_BLOCK * (a.h.d | 1) + (size_t)a.h.K * (a.h.stride | 1)) * sizeof(double);
      if (staged && sm <= 100 * 1024) {
        auto kern = k_simulate_staged<DD>;
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm);
        if (e != cudaSuccess) return e;
        kern<<<nblk(a.N), PMC_BLOCK, sm, s>>>(a.mix, a.h, a.box, a.N, a.seed, a.iter, a.offset, a.X, a.idx, a.flg, a.scal);
        break;
      }
      k_simulate<DD><<<nblk(a.N), PMC_BLOCK, 0, s>>>(a.mix, a.h, a.box, a.N, a.seed, a.iter, a.offset, a.X, a.idx,
                                                     a.flg, a.scal);
      break;
    }
    case OP_SIMULATE_DRAWS:
      k_simulate_from_draws<DD><<<nblk(a.N), PMC_BLOCK, 0, s>>>(a.mix, a.h, a.box, a.N, a.U, a.Z, a.X, a.idx, a.flg);
      break;
    case OP_LOGQ:
      k_logq<DD><<<nblk(a.N), PMC_BLOCK, 0, s>>>(a.mix, a.h, a.N, a.Xc, a.out);
      break;
    case OP_LIKE_MIX:
      k_like_mix<DD><<<nblk(a.N), PMC_BLOCK, 0, s>>>(a.mix, a.h, a.is_mixture, a.N, a.Xc, a.dX, a.sel, a.flgc,
                                                     a.logpi, a.err, a.set, a.add_const);
      break;
    case OP_WEIGHTS:
      if constexpr (DD >= 5) {
        // PMCB200_ESTEP: 0 = one sample per thread, row-oriented, mixture through L1 (k_weights);
        // 1, 3 = k_weights_multi with that many samples per thread; 2 (default) = 4 samples for d <= 8, else 2
        // (measured, C3 d = 20 K = 10, 1e7 samples: 5.78 / 3.89 / 3.00 / 3.33 ms for modes 0 / 1 / 2 / 3)
        const char *em = getenv("PMCB200_ESTEP");      // read per call
        const int mode = em && *em ? atoi(em) : ESTEP_DEFAULT;
        const size_t mixbytes = (size_t)a.h.K * a.h.stride * sizeof(double);
#define WM(SV) { auto kern = k_weights_multi<DD, SV>; const size_t sm = mixbytes + (size_t)DD * SV * PMC_BLOCK * sizeof(double); \
          if (sm <= 200 * 1024) { \
            cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm); if (e != cudaSuccess) return e; \
            kern<<<(int)((a.N + PMC_BLOCK * SV - 1) / (PMC_BLOCK * SV)), PMC_BLOCK, sm, s>>>(a.mix, a.h, a.N, a.Xc, a.logpic, a.errc, a.beta, a.flg, a.logw, a.scal, a.rho); \
            if (a.rho && a.rho_written) *a.rho_written = 1; \
            return cudaGetLastError(); } }
        if (mode == 1) WM(1)
        if (mode == 2) { if constexpr (DD <= 8) WM(4) else WM(2) }
        if (mode == 3) WM(3)
        if (mode == 4) {      // experimental: column-packed staged mixture (two doubles per LDS)
          auto kern = k_weights_multi<DD, (DD <= 8 ? 4 : 2), true>;
          constexpr int SV = (DD <= 8 ? 4 : 2);
          const size_t sm = (size_t)a.h.K * CpLayout<DD>::stride * sizeof(double) + (size_t)DD * SV * PMC_BLOCK * sizeof(double);
          if (sm <= 200 * 1024) {
            cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm); if (e != cudaSuccess) return e;
            kern<<<(int)((a.N + PMC_BLOCK * SV - 1) / (PMC_BLOCK * SV)), PMC_BLOCK, sm, s>>>(a.mix, a.h, a.N, a.Xc, a.logpic, a.errc, a.beta, a.flg, a.logw, a.scal, a.rho);
            if (a.rho && a.rho_written) *a.rho_written = 1;
            return cudaGetLastError();
          }
        }
#undef WM
      }
      k_weights<DD><<<nblk(a.N), PMC_BLOCK, 0, s>>>(a.mix, a.h, a.N, a.Xc, a.logpic, a.errc, a.beta, a.flg, a.logw,
                                                    a.scal);
      break;
    case OP_EM: {
      {
        if (a.em_mma && em_mma_ok(a.h.K, a.h.d, a.h.df > 0) && a.k0 == 0 && a.Kg == a.h.K) return run_em_mma<DD>(a, s);
      }
      const bool reg = em_use_reg(a.Kg, a.h.d, a.h.df > 0);
      auto kern = reg ? k_em_stats<DD, true> : k_em_stats<DD, false>;
      cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)a.smem);
      if (e != cudaSuccess) return e;
      // persistent grid: as many blocks as are resident (a.blocks = buffer capacity)
      int per_sm = 1, dev = 0, sms = 148;
      cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, PMC_BLOCK, a.smem);
      cudaGetDevice(&dev);
      cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
      int blocks = std::max(1, std::min(a.blocks, std::max(1, per_sm) * sms));
      if (a.nblocks_out) *a.nblocks_out = blocks;
      kern<<<blocks, PMC_BLOCK, a.smem, s>>>(a.mix, a.h, a.N, a.Xc, a.idxc, a.flgc, a.logwc, a.scal,
                                             a.partials, a.linear, a.k0, a.Kg);
      break;
    }
    default:
      return cudaErrorInvalidValue;
  }
  return cudaGetLastError();
}

bool FN(int op, const MixArgs &a, cudaStream_t s, cudaError_t *e) {
  const int dd = pmc_pad_dim(a.h.d);
#define X(D) if (dd == D) { *e = run<D>(op, a, s); return true; }
  DLIST(X)
#undef X
  return false;
}
