// pmcb200.cu -- C-ABI of the B200-native PMC iteration (include/pmcb200.h).
// Host side: context, packing of the proposal/target into device layouts,
// kernel launches.  No CPU compute path: every entry point launches kernels.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cstdarg>
#include <cmath>
#include <vector>
#include <algorithm>

#include <nvtx3/nvToolsExt.h>     // header-only; the ranges cost nothing unless a profiler injects itself

#include "common.cuh"
#include "cosmo_types.cuh"
#include "pmc_kernels.cuh"
#include "launch.h"

// ---- context ----------------------------------------------------------------
struct DevBuf {
  void *p = nullptr;
  size_t cap = 0;
};

struct pmcb200_ctx {
  int device = 0;
  cudaStream_t stream = nullptr;
  bool own_stream = false;
  cudaStream_t copy_stream = nullptr;     // D2H of finished arrays overlaps the likelihood kernel
  cudaEvent_t ev_a = nullptr, ev_b = nullptr;
  cudaEvent_t ev_blk = nullptr;           // statistics block ready (all-gather between contexts)
  char errmsg[512] = {0};
  int64_t launches = 0;
  // proposal
  bool have_prop = false;
  MixHdr h{};
  std::vector<double> wght, mean, chol;   // host mirror: [K], [K*d], [K*d*d]
  double *d_mix = nullptr;                // packed device mixture (+ pivot)
  size_t mix_cap = 0;
  // target
  bool have_target = false;
  pmcb200_target_t tgt{};
  DevLike like[PMCB200_MAX_DATA];
  std::vector<void *> tgt_allocs;
  double *d_box = nullptr;                // [2*d]: min, max
  int box_d = 0;
  double logpr_const = 0.0;
  MixHdr prior_h{};
  double *d_prior = nullptr;
  int *d_prior_sel = nullptr;
  // per-iteration device state
  DevScal *d_scal = nullptr;
  DevCount *d_cnt = nullptr;
  unsigned *d_fin_cnt = nullptr;          // blocks of k_em_finish that are done (self-resetting)
  double *d_partials = nullptr; size_t partials_cap = 0;
  double *d_work = nullptr, *d_result = nullptr; size_t work_cap = 0;
  double *h_result = nullptr;             // pinned
  int em_blocks = 0;
  int em_no_mma = 0;        // PMCB200_EM_NO_MMA=1: keep the shared-memory EM kernel for d >= 10 (A/B measurements)
  int sm_count = 148;
  // E-step cache (d >= 5, Gaussian proposal): the weight kernel leaves alpha_k phi_k(x_n) here and the EM statistics
  // kernel of the same iteration reads it back instead of repeating the K whitenings per sample.  Valid for exactly
  // one (sample array, N, proposal version); consumed by the next em_local.  PMCB200_EM_NO_RHO=1 disables it.
  DevBuf sRho;
  const double *rho_X = nullptr; int64_t rho_N = 0; uint64_t rho_ver = 0, prop_ver = 0; bool rho_valid = false;
  int em_no_rho = 0;
  int rho_min_dim = 5;      // smallest padded dimension that uses the cache (PMCB200_RHO_MIN_DIM; the multi-sample weight
                            // kernel that fills it exists from d = 5).  Measured per 1e7 samples, cache from d = 10 / from d = 5:
                            // C2 14.86 / 14.56 ms, C4 23.68 / 23.19, C5 (K = 30, d = 8) 44.19 / 42.46
  // scratch for the host-buffer API.  The sample arrays exist twice so that the device-to-host copies of one
  // iteration can drain while the next one computes (pmcb200_iteration_host_begin / pmcb200_host_wait); the
  // second set is only allocated if a caller actually leaves copies in flight.
  struct ScratchSet {
    DevBuf X, Idx, Flg, Logw;
    cudaEvent_t copied = nullptr;         // recorded on copy_stream after the set's last queued copy
    bool busy = false;                    // copies queued and not yet waited for
    uint64_t seq = 0;                     // order of the iterations that used the set
  } set[2];
  int cur = 0;
  uint64_t seq = 0;
  DevBuf sFish, sLogpi, sErr, sBlock, sAll;
  DevBuf sFb;                             // [N] uint32 work list + counter (at the end): SN samples for the exact kernel
  unsigned *d_fb_cnt = nullptr;
  DevBuf sPost, sPostTmp;                 // post-processing work space
};

// NVTX range per phase of the iteration (nsys / ncu --nvtx timelines: sample, likelihood, weights, EM, M-step, copies)
struct NvtxRange {
  explicit NvtxRange(const char *name) { nvtxRangePushA(name); }
  ~NvtxRange() { nvtxRangePop(); }
};

static int fail(pmcb200_ctx *c, int code, const char *fmt, ...) {
  if (c) {
    va_list ap; va_start(ap, fmt);
    vsnprintf(c->errmsg, sizeof(c->errmsg), fmt, ap);
    va_end(ap);
  }
  return code;
}
#define CUDA_OK(c, call)                                                            \
  do {                                                                              \
    cudaError_t e__ = (call);                                                       \
    if (e__ != cudaSuccess)                                                         \
      return fail(c, PMCB200_ERR_CUDA, "%s: %s (%s:%d)", #call, cudaGetErrorString(e__), \
                  __FILE__, __LINE__);                                              \
  } while (0)
#define LAUNCH_OK(c)                                                                \
  do {                                                                              \
    (c)->launches++;                                                                \
    cudaError_t e__ = cudaGetLastError();                                           \
    if (e__ != cudaSuccess)                                                         \
      return fail(c, PMCB200_ERR_CUDA, "kernel launch: %s (%s:%d)",                 \
                  cudaGetErrorString(e__), __FILE__, __LINE__);                     \
  } while (0)

#define MIX_OK(c, op, a)                                                            \
  do {                                                                              \
    (c)->launches++;                                                                \
    cudaError_t e__ = pmc_mix_launch(op, a, (c)->stream);                           \
    if (e__ != cudaSuccess)                                                         \
      return fail(c, PMCB200_ERR_CUDA, "kernel launch (op %d): %s (%s:%d)", (int)(op), \
                  cudaGetErrorString(e__), __FILE__, __LINE__);                     \
  } while (0)


static int ensure(pmcb200_ctx *c, DevBuf &b, size_t bytes) {
  if (b.cap >= bytes) return 0;
  if (b.p) CUDA_OK(c, cudaFree(b.p));
  b.p = nullptr; b.cap = 0;
  CUDA_OK(c, cudaMalloc(&b.p, bytes));
  b.cap = bytes;
  return 0;
}

// ---- supernova tiles of the tensor-core SN kernel (sn_spectral.cuh) ------------------------------------------------
// nz redshifts, the supernovae sorted by redshift: first[z] .. first[z + 1] are the rows at redshift z.  Per tile 8 columns:
// tcol = supernova row, or -(row) - 1 for an empty column (repeats that row with sigma^2 = 1e300); tz = redshift index of
// the column (its W row); tsec = 1 for a secondary tile.  A primary tile takes the first supernova of 8 redshifts
// (ordered by multiplicity, so that the repeated ones share tiles), the s-th further supernova of each sits in the same
// column of the s-th secondary tile behind it.
static void sn_tile_plan(int nz, const std::vector<int> &first, std::vector<int> &tcol, std::vector<int> &tsec,
                         std::vector<int> &tz) {
  std::vector<int> zord(nz);
  for (int z = 0; z < nz; z++) zord[z] = z;
  std::stable_sort(zord.begin(), zord.end(), [&](int a, int b) { return first[a + 1] - first[a] > first[b + 1] - first[b]; });
  tcol.clear(); tsec.clear(); tz.clear();
  for (int p0 = 0; p0 < nz; p0 += 8) {
    int maxm = 1;
    for (int j = 0; j < 8 && p0 + j < nz; j++) maxm = std::max(maxm, first[zord[p0 + j] + 1] - first[zord[p0 + j]]);
    for (int sidx = 0; sidx < maxm; sidx++) {
      tsec.push_back(sidx > 0);
      for (int j = 0; j < 8; j++) {
        const int z = zord[std::min(p0 + j, nz - 1)], mult = first[z + 1] - first[z];
        const bool real = p0 + j < nz && sidx < mult;
        tz.push_back(z);
        tcol.push_back(real ? first[z] + sidx : -first[z] - 1);
      }
    }
  }
}
// the plan for a list of redshifts (host only, no device needed: diagnostics and the CPU tests).  tile_col[8 t + j] = index
// into z[] of the supernova in column j of tile t, or -1 - index for an empty column; returns the number of tiles, or
// a negative error code (cap too small: -needed)
extern "C" int pmcb200_sn_tile_plan(int n, const double *z, int cap, int *tile_sec, int *tile_col) {
  if (n < 1 || !z || cap < 0 || (cap > 0 && (!tile_sec || !tile_col))) return PMCB200_ERR_ARG;
  std::vector<int> ord(n);
  for (int i = 0; i < n; i++) ord[i] = i;
  std::stable_sort(ord.begin(), ord.end(), [&](int a, int b) { return z[a] < z[b]; });
  std::vector<int> first;
  for (int r = 0; r < n; r++) if (r == 0 || z[ord[r]] != z[ord[r - 1]]) first.push_back(r);
  const int nz = (int)first.size();
  first.push_back(n);
  std::vector<int> tcol, tsec, tz;
  sn_tile_plan(nz, first, tcol, tsec, tz);
  const int ntile = (int)tsec.size();
  if (ntile > cap) return -ntile;
  for (int t = 0; t < ntile; t++) {
    tile_sec[t] = tsec[t];
    for (int j = 0; j < 8; j++) {
      const int col = tcol[(size_t)t * 8 + j];
      tile_col[t * 8 + j] = col >= 0 ? ord[col] : -1 - ord[-col - 1];
    }
  }
  return ntile;
}

// ---- host-side packing ---------------------------------------------------------
static int host_cholesky(int d, double *A) {
  for (int j = 0; j < d; j++) {
    double s = A[j * d + j];
    for (int k = 0; k < j; k++) s -= A[j * d + k] * A[j * d + k];
    if (!(s > 0.0) || !std::isfinite(s)) return -1;
    double ljj = std::sqrt(s);
    A[j * d + j] = ljj;
    for (int i = j + 1; i < d; i++) {
      double t = A[i * d + j];
      for (int k = 0; k < j; k++) t -= A[i * d + k] * A[j * d + k];
      A[i * d + j] = t / ljj;
    }
  }
  for (int i = 0; i < d; i++)
    for (int j = i + 1; j < d; j++) A[i * d + j] = 0.0;
  return 0;
}

// pack one component (wght, mean[d], chol[d*d] lower) at dst[2 + 2D + D(D+1)/2], layout padded
// to D >= d (padded coordinates: mean 0, unit diagonal)
static void pack_comp(double *dst, int d, int D, int df, double w, const double *mean, const double *chol) {
  const int tri = mix_tri(D);
  double logdet = 0.0;
  dst[0] = w;
  for (int i = 0; i < D; i++) {
    dst[2 + i] = i < d ? mean[i] : 0.0;
    for (int j = 0; j <= i; j++)
      dst[2 + D + i * (i + 1) / 2 + j] = (i < d) ? chol[i * d + j] : (i == j ? 1.0 : 0.0);
    dst[2 + D + tri + i] = i < d ? 1.0 / chol[i * d + i] : 1.0;
    if (i < d) logdet += std::log(chol[i * d + i]);
  }
  if (df <= 0) dst[1] = -0.5 * d * LN2PI - logdet;
  else {
    double nu = (double)df;
    dst[1] = std::lgamma(0.5 * (nu + d)) - std::lgamma(0.5 * nu) - 0.5 * d * std::log(nu * M_PI) - logdet;
  }
}

static int upload_mix(pmcb200_ctx *c, int K, int d, int df, const double *w, const double *mean,
                      const double *chol, MixHdr &h, double **dbuf, size_t *cap) {
  const int D = pmc_pad_dim(d);
  h.K = K; h.d = d; h.df = df; h.stride = mix_stride(d); h.tri = mix_tri(D);
  size_t n = (size_t)K * h.stride + D;
  std::vector<double> buf(n, 0.0);
  double wsum = 0.0;
  for (int k = 0; k < K; k++) {
    pack_comp(buf.data() + (size_t)k * h.stride, d, D, df, w[k], mean + (size_t)k * d, chol + (size_t)k * d * d);
    wsum += w[k];
  }
  for (int i = 0; i < d; i++) {   // common EM pivot: weighted mean of component means
    double p = 0.0;
    for (int k = 0; k < K; k++) p += w[k] * mean[(size_t)k * d + i];
    buf[(size_t)K * h.stride + i] = (wsum > 0.0) ? p / wsum : 0.0;
  }
  if (!*dbuf || *cap < n) {
    if (*dbuf) CUDA_OK(c, cudaFree(*dbuf));
    *dbuf = nullptr;
    CUDA_OK(c, cudaMalloc((void **)dbuf, n * sizeof(double)));
    *cap = n;
  }
  CUDA_OK(c, cudaMemcpyAsync(*dbuf, buf.data(), n * sizeof(double), cudaMemcpyHostToDevice, c->stream));
  CUDA_OK(c, cudaStreamSynchronize(c->stream));   // buf is a stack-lifetime staging vector
  return 0;
}

template <class T>
static int dev_copy(pmcb200_ctx *c, const T *src, size_t n, const T **out) {
  T *p = nullptr;
  CUDA_OK(c, cudaMalloc((void **)&p, std::max<size_t>(n, 1) * sizeof(T)));
  c->tgt_allocs.push_back(p);
  if (n) CUDA_OK(c, cudaMemcpy(p, src, n * sizeof(T), cudaMemcpyHostToDevice));
  *out = p;
  return 0;
}

static void free_target(pmcb200_ctx *c) {
  for (void *p : c->tgt_allocs) cudaFree(p);
  c->tgt_allocs.clear();
  c->have_target = false;
  c->d_box = nullptr; c->box_d = 0; c->d_prior = nullptr; c->d_prior_sel = nullptr;
}

// ---- life cycle ------------------------------------------------------------------
extern "C" int pmcb200_version(void) { return PMCB200_VERSION; }

extern "C" int pmcb200_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) return 0;
  return n;
}

extern "C" void pmcb200_destroy(pmcb200_ctx *c);

static int create_impl(pmcb200_ctx *c, int device, void *stream) {
  CUDA_OK(c, cudaSetDevice(device));
  // NULL = the legacy default stream (orders with torch's default stream and
  // with plain cudaMemcpy in C hosts); (void*)-1 = a private non-blocking stream
  if (stream == (void *)-1) { CUDA_OK(c, cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking)); c->own_stream = true; }
  else c->stream = (cudaStream_t)stream;
  CUDA_OK(c, cudaStreamCreateWithFlags(&c->copy_stream, cudaStreamNonBlocking));
  CUDA_OK(c, cudaEventCreateWithFlags(&c->ev_a, cudaEventDisableTiming));
  CUDA_OK(c, cudaEventCreateWithFlags(&c->ev_b, cudaEventDisableTiming));
  CUDA_OK(c, cudaEventCreateWithFlags(&c->ev_blk, cudaEventDisableTiming));
  for (auto &t : c->set) CUDA_OK(c, cudaEventCreateWithFlags(&t.copied, cudaEventDisableTiming));
  cudaDeviceProp prop;
  CUDA_OK(c, cudaGetDeviceProperties(&prop, device));
  c->sm_count = prop.multiProcessorCount;
  CUDA_OK(c, cudaMalloc((void **)&c->d_scal, sizeof(DevScal)));
  CUDA_OK(c, cudaMemset(c->d_scal, 0, sizeof(DevScal)));
  if (pmc_init_sn_tables()) return fail(c, PMCB200_ERR_CUDA, "SN table upload failed");
  CUDA_OK(c, cudaMalloc((void **)&c->d_cnt, sizeof(DevCount)));
  CUDA_OK(c, cudaMemset(c->d_cnt, 0, sizeof(DevCount)));
  CUDA_OK(c, cudaMalloc((void **)&c->d_fb_cnt, sizeof(unsigned)));
  CUDA_OK(c, cudaMemset(c->d_fb_cnt, 0, sizeof(unsigned)));
  CUDA_OK(c, cudaMalloc((void **)&c->d_fin_cnt, sizeof(unsigned)));
  CUDA_OK(c, cudaMemset(c->d_fin_cnt, 0, sizeof(unsigned)));
  CUDA_OK(c, cudaMallocHost((void **)&c->h_result, sizeof(double) * (RES_HDR + PMCB200_MAX_COMP * (1 + PMCB200_MAX_DIM + PMCB200_MAX_DIM * PMCB200_MAX_DIM))));
  return 0;
}

extern "C" int pmcb200_create(int device, void *stream, pmcb200_ctx **out) {
  if (!out) return PMCB200_ERR_ARG;
  *out = nullptr;
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess || n <= 0 || device < 0 || device >= n) {
    fprintf(stderr, "pmcb200_create: no usable CUDA device %d (%s); there is no CPU fallback\n",
            device, e != cudaSuccess ? cudaGetErrorString(e) : "device count");
    return PMCB200_ERR_CUDA;
  }
  pmcb200_ctx *c = new pmcb200_ctx();
  c->device = device;
  int rc = create_impl(c, device, stream);
  if (rc) {
    fprintf(stderr, "pmcb200_create: %s\n", c->errmsg);
    pmcb200_destroy(c);
    return rc;
  }
  *out = c;
  return 0;
}

extern "C" void pmcb200_destroy(pmcb200_ctx *c) {
  if (!c) return;
  cudaSetDevice(c->device);
  cudaStreamSynchronize(c->stream);
  free_target(c);
  if (c->copy_stream) cudaStreamSynchronize(c->copy_stream);
  for (auto &t : c->set) {
    for (DevBuf *b : {&t.X, &t.Idx, &t.Flg, &t.Logw}) if (b->p) cudaFree(b->p);
    if (t.copied) cudaEventDestroy(t.copied);
  }
  for (DevBuf *b : {&c->sFish, &c->sLogpi, &c->sErr, &c->sBlock, &c->sAll, &c->sPost, &c->sPostTmp, &c->sRho, &c->sFb})
    if (b->p) cudaFree(b->p);
  if (c->d_fb_cnt) cudaFree(c->d_fb_cnt);
  if (c->d_mix) cudaFree(c->d_mix);
  if (c->d_scal) cudaFree(c->d_scal);
  if (c->d_cnt) cudaFree(c->d_cnt);
  if (c->d_fin_cnt) cudaFree(c->d_fin_cnt);
  if (c->d_partials) cudaFree(c->d_partials);
  if (c->d_work) cudaFree(c->d_work);
  if (c->d_result) cudaFree(c->d_result);
  if (c->h_result) cudaFreeHost(c->h_result);
  if (c->copy_stream) cudaStreamDestroy(c->copy_stream);
  if (c->ev_a) cudaEventDestroy(c->ev_a);
  if (c->ev_b) cudaEventDestroy(c->ev_b);
  if (c->ev_blk) cudaEventDestroy(c->ev_blk);
  if (c->own_stream) cudaStreamDestroy(c->stream);
  delete c;
}

extern "C" const char *pmcb200_last_error(const pmcb200_ctx *c) { return c ? c->errmsg : "null context"; }
extern "C" void *pmcb200_stream(pmcb200_ctx *c) { return c ? (void *)c->stream : nullptr; }
extern "C" int64_t pmcb200_launch_count(const pmcb200_ctx *c) { return c ? c->launches : 0; }

extern "C" int pmcb200_sync(pmcb200_ctx *c) {
  if (!c) return PMCB200_ERR_ARG;
  CUDA_OK(c, cudaStreamSynchronize(c->stream));
  return 0;
}

extern "C" int pmcb200_dev_alloc(pmcb200_ctx *c, size_t bytes, void **dptr) {
  if (!c || !dptr) return PMCB200_ERR_ARG;
  CUDA_OK(c, cudaSetDevice(c->device));
  CUDA_OK(c, cudaMalloc(dptr, std::max<size_t>(bytes, 1)));
  return 0;
}
extern "C" int pmcb200_dev_free(pmcb200_ctx *c, void *dptr) {
  if (!c) return PMCB200_ERR_ARG;
  CUDA_OK(c, cudaFree(dptr));
  return 0;
}
extern "C" int pmcb200_h2d(pmcb200_ctx *c, void *dptr, const void *hptr, size_t bytes) {
  if (!c) return PMCB200_ERR_ARG;
  c->rho_valid = false;      // the caller may be rewriting the sample or flag array the E-step cache was computed from
  CUDA_OK(c, cudaMemcpyAsync(dptr, hptr, bytes, cudaMemcpyHostToDevice, c->stream));
  CUDA_OK(c, cudaStreamSynchronize(c->stream));
  return 0;
}
extern "C" int pmcb200_d2h(pmcb200_ctx *c, void *hptr, const void *dptr, size_t bytes) {
  if (!c) return PMCB200_ERR_ARG;
  CUDA_OK(c, cudaMemcpyAsync(hptr, dptr, bytes, cudaMemcpyDeviceToHost, c->stream));
  CUDA_OK(c, cudaStreamSynchronize(c->stream));
  return 0;
}

// ---- proposal ----------------------------------------------------------------------
static int install_proposal(pmcb200_ctx *c, int K, int d, int df, const double *w, const double *mean,
                            const double *chol) {
  if (K < 1 || K > PMCB200_MAX_COMP || d < 1 || d > PMCB200_MAX_DIM)
    return fail(c, PMCB200_ERR_DIM, "proposal: ncomp=%d (max %d), ndim=%d (max %d)", K, PMCB200_MAX_COMP, d,
                PMCB200_MAX_DIM);
  if (em_smem_bytes(1, d, df > 0) > 200 * 1024)
    return fail(c, PMCB200_ERR_UNSUP, "proposal: d=%d needs %zu B of shared memory per component in the EM kernel",
                d, em_smem_bytes(1, d, df > 0));
  for (int k = 0; k < K; k++) {
    if (!(w[k] >= 0.0) || !std::isfinite(w[k])) return fail(c, PMCB200_ERR_ARG, "proposal weight %d = %g", k, w[k]);
    for (int i = 0; i < d; i++)
      if (!(chol[(size_t)k * d * d + i * d + i] > 0.0))
        return fail(c, PMCB200_ERR_CHOLESKY, "component %d: Cholesky diagonal %d not positive", k, i);
  }
  c->wght.assign(w, w + K);
  c->mean.assign(mean, mean + (size_t)K * d);
  c->chol.assign(chol, chol + (size_t)K * d * d);
  for (int k = 0; k < K; k++)   // strict upper triangle is ignored
    for (int i = 0; i < d; i++)
      for (int j = i + 1; j < d; j++) c->chol[(size_t)k * d * d + i * d + j] = 0.0;
  int rc = upload_mix(c, K, d, df, c->wght.data(), c->mean.data(), c->chol.data(), c->h, &c->d_mix, &c->mix_cap);
  if (rc) return rc;
  // EM work buffers
  int64_t len = stat_len(K, d);
  { const char *ev = getenv("PMCB200_EM_NO_MMA"); c->em_no_mma = ev && *ev && *ev != '0'; }
  { const char *ev = getenv("PMCB200_EM_NO_RHO"); c->em_no_rho = ev && *ev && *ev != '0'; }
  { const char *ev = getenv("PMCB200_RHO_MIN_DIM"); if (ev && *ev) c->rho_min_dim = atoi(ev); }
  c->prop_ver++; c->rho_valid = false;
  c->em_blocks = 4 * c->sm_count;     // capacity of the partials buffer; the launch uses the resident count
  size_t need = (size_t)c->em_blocks * len;
  if (c->partials_cap < need) {
    if (c->d_partials) CUDA_OK(c, cudaFree(c->d_partials));
    CUDA_OK(c, cudaMalloc((void **)&c->d_partials, need * sizeof(double)));
    c->partials_cap = need;
  }
  size_t rlen = RES_HDR + (size_t)K * (1 + d + (size_t)d * d);
  if (c->work_cap < std::max<size_t>(len, rlen)) {
    if (c->d_work) CUDA_OK(c, cudaFree(c->d_work));
    if (c->d_result) CUDA_OK(c, cudaFree(c->d_result));
    c->work_cap = std::max<size_t>(len, rlen);
    CUDA_OK(c, cudaMalloc((void **)&c->d_work, c->work_cap * sizeof(double)));
    CUDA_OK(c, cudaMalloc((void **)&c->d_result, c->work_cap * sizeof(double)));
  }
  c->have_prop = true;
  return 0;
}

extern "C" int pmcb200_set_proposal(pmcb200_ctx *c, int K, int d, int df, const double *w,
                                    const double *mean, const double *chol) {
  if (!c || !w || !mean || !chol) return PMCB200_ERR_ARG;
  CUDA_OK(c, cudaSetDevice(c->device));
  return install_proposal(c, K, d, df, w, mean, chol);
}

extern "C" int pmcb200_set_proposal_cov(pmcb200_ctx *c, int K, int d, int df, const double *w,
                                        const double *mean, const double *cov) {
  if (!c || !w || !mean || !cov) return PMCB200_ERR_ARG;
  if (K < 1 || K > PMCB200_MAX_COMP || d < 1 || d > PMCB200_MAX_DIM)
    return fail(c, PMCB200_ERR_DIM, "proposal: ncomp=%d, ndim=%d out of range", K, d);
  std::vector<double> L(cov, cov + (size_t)K * d * d);
  for (int k = 0; k < K; k++)
    if (host_cholesky(d, L.data() + (size_t)k * d * d))
      return fail(c, PMCB200_ERR_CHOLESKY, "component %d: covariance not positive definite", k);
  CUDA_OK(c, cudaSetDevice(c->device));
  return install_proposal(c, K, d, df, w, mean, L.data());
}

extern "C" int pmcb200_get_proposal(pmcb200_ctx *c, double *w, double *mean, double *chol, double *cov) {
  if (!c) return PMCB200_ERR_ARG;
  if (!c->have_prop) return fail(c, PMCB200_ERR_STATE, "no proposal set");
  const int K = c->h.K, d = c->h.d;
  if (w) memcpy(w, c->wght.data(), K * sizeof(double));
  if (mean) memcpy(mean, c->mean.data(), (size_t)K * d * sizeof(double));
  if (chol) memcpy(chol, c->chol.data(), (size_t)K * d * d * sizeof(double));
  if (cov)
    for (int k = 0; k < K; k++) {
      const double *L = c->chol.data() + (size_t)k * d * d;
      for (int i = 0; i < d; i++)
        for (int j = 0; j < d; j++) {
          double s = 0.0;
          for (int q = 0; q <= std::min(i, j); q++) s += L[i * d + q] * L[j * d + q];
          cov[(size_t)k * d * d + i * d + j] = s;
        }
    }
  return 0;
}

// ---- target ------------------------------------------------------------------------
static int pack_gauss(pmcb200_ctx *c, int n, const double *mean, const double *chol, const double **out) {
  std::vector<double> buf(2 + 2 * n + mix_tri(n));     // unpadded: read with runtime n (gauss_comp_logpdf)
  pack_comp(buf.data(), n, n, -1, 1.0, mean, chol);
  return dev_copy<double>(c, buf.data(), buf.size(), out);
}

static int build_sn(pmcb200_ctx *c, const pmcb200_like_t &L, DevLike &D) {
  const int n = L.sn_n;
  if (n < 1 || !L.sn_z || !L.sn_m || !L.sn_s || !L.sn_c || !L.sn_cov)
    return fail(c, PMCB200_ERR_ARG, "SNIa: empty sample");
  if (L.sn_chi2mode < 0 || L.sn_chi2mode > PMCB200_CHI2_betaz)
    return fail(c, PMCB200_ERR_UNSUP, "SNIa: chi2mode %d not supported (chi2_Theta1 / chi2_dust)", L.sn_chi2mode);
  // sort by redshift; supernovae sharing a redshift share one distance integral
  std::vector<int> ord(n);
  for (int i = 0; i < n; i++) ord[i] = i;
  std::stable_sort(ord.begin(), ord.end(), [&](int a, int b) { return L.sn_z[a] < L.sn_z[b]; });
  std::vector<double> rows((size_t)n * SN_ROW, 0.0);
  std::vector<int> first;
  std::vector<double2> nodes;
  std::vector<double> nodes_a;
  auto node = [](double x) { return make_double2(std::log(x), 1.0 / std::sqrt(x)); };
  const double pv_fac = 5.0 / M_LN10 * L.sn_v_pec / C_KMS;
  for (int r = 0; r < n; r++) {
    int i = ord[r];
    double z = L.sn_z[i];
    if (!(z > 0.0)) return fail(c, PMCB200_ERR_ARG, "SNIa: redshift[%d] = %g", i, z);
    if (r == 0 || z != L.sn_z[ord[r - 1]]) {
      first.push_back(r);
      double a = 1.0 / (1.0 + z), hh = 1.0 - a;
      size_t base = nodes.size();
      nodes.resize(base + SN_NODES);
      nodes_a.resize(base + SN_NODES, 1.0);
      nodes[base] = node(a); nodes_a[base] = a;
      for (int j = 2; j <= 7; j++) {
        int it = 1 << (j - 2);
        double del = hh / (double)it, x = a + 0.5 * del;   // NR trapzd: x += del
        for (int q = 0; q < it; q++, x += del) { nodes[base + it + q] = node(x); nodes_a[base + it + q] = x; }
      }
    }
    double *row = rows.data() + (size_t)r * SN_ROW;
    row[0] = L.sn_m[i]; row[1] = L.sn_s[i]; row[2] = L.sn_c[i]; row[3] = z;
    for (int q = 0; q < 6; q++) row[4 + q] = L.sn_cov[(size_t)i * 6 + q];
    double spv = pv_fac / z;
    row[4] += spv * spv + L.sn_sig_int * L.sn_sig_int;   // Vmm + sigma_pv^2 + sigma_int^2
  }
  first.push_back(n);
  D.sn_n = n;
  D.sn_nz = (int)first.size() - 1;
  D.sn_chi2mode = L.sn_chi2mode;
  D.sn_add_logdetCov = L.sn_add_logdetCov;
  for (int i = 0; i < 4; i++) D.Theta2[i] = L.sn_Theta2[i];
  for (int i = 0; i < 3; i++) D.Theta2_denom[i] = L.sn_Theta2_denom[i];
  D.sig_int2 = L.sn_sig_int * L.sn_sig_int;
  D.pv_fac = pv_fac;
  // launch-uniform specialisation: second exponent term only if w1 can be non-zero;
  // curvature term dropped only if flatness is structural (Omega_de = 1 - Omega_m - 0 - Omega_nu)
  bool has_w1 = L.model.w1_de != 0.0 || L.model.de_param == PMCB200_DE_jassal;
  bool s_Om = false, s_other = false;
  for (int j = 0; j < L.npar; j++) {
    int pj = L.par[j];
    if (pj == PMCB200_P_w1de) has_w1 = true;
    if (pj == PMCB200_P_Omegam) s_Om = true;
    if (pj == PMCB200_P_Omegade || pj == PMCB200_P_OmegaK || pj == PMCB200_P_Omegab || pj == PMCB200_P_Omegac ||
        (pj >= PMCB200_P_omegam && pj <= PMCB200_P_omegaK))
      s_other = true;
  }
  double OK0 = 1.0 - L.model.Omega_m - L.model.Omega_de - L.model.Omega_nu_mass;
  D.sn_hasq = has_w1 ? 1 : 0;
  D.sn_flat = (!s_other && (s_Om || std::fabs(OK0) < 1e-15)) ? 1 : 0;
  int rc;
  if ((rc = dev_copy<double2>(c, nodes.data(), nodes.size(), &D.nodes))) return rc;
  if ((rc = dev_copy<double>(c, nodes_a.data(), nodes_a.size(), &D.nodes_a))) return rc;
  std::vector<double> nodes4((size_t)D.sn_nz * 64, 0.0);
  for (int z = 0; z < D.sn_nz; z++)
    for (int i = 0; i < 16; i++) {
      double *q = &nodes4[((size_t)z * 16 + i) * 4];
      q[0] = nodes[(size_t)z * SN_NODES + i].x; q[1] = nodes[(size_t)z * SN_NODES + i].y;
      q[2] = nodes_a[(size_t)z * SN_NODES + i];
    }
  if ((rc = dev_copy<double>(c, nodes4.data(), nodes4.size(), &D.nodes4))) return rc;
  if ((rc = dev_copy<int>(c, first.data(), first.size(), &D.first))) return rc;
  if ((rc = dev_copy<double>(c, rows.data(), rows.size(), &D.sn))) return rc;
  // Spectral form (sn_spectral.cuh).  Stage 5 of NR qromb at redshift z is ss = h sum_i w_i f(a_i), dss = h sum_i e_i
  // f(a_i) over the 17 nodes tabulated above (+ the end point a = 1), f(a) = a^-1/2 q(a).  With q = sum_m c_m T_m(x(a))
  // on [a(z_max), 1]:  ss = sum_m W[z][m] c_m,  dss = sum_m D[z][m] c_m.  Long double throughout.
  {
    const int M = pmc_sn_spectral_M(), nz = D.sn_nz;
    static const long double RW[10] = {3937.0L / 103275.0L, 3062.0L / 80325.0L, 27728.0L / 722925.0L, 22016.0L / 722925.0L,
                                       65536.0L / 722925.0L, -31.0L / 206550.0L, -73.0L / 481950.0L, -67.0L / 722925.0L,
                                       -424.0L / 722925.0L, 256.0L / 722925.0L};      // = ROMBW (cosmo.cuh)
    long double alo = 1.0L;
    for (int z = 0; z < nz; z++) alo = std::min<long double>(alo, (long double)nodes_a[(size_t)z * SN_NODES]);
    if (alo > 0.999L) alo = 0.999L;     // keep the map x(a) well conditioned for a sample of tiny redshifts
    const long double mid = 0.5L * (1.0L + alo), half = 0.5L * (1.0L - alo);
    std::vector<double> cn((size_t)M * 4, 0.0), W((size_t)nz * M), dmax(M, 0.0);
    for (int j = 0; j < M; j++) {
      const long double aj = mid + half * cosl(M_PIl * (j + 0.5L) / M);
      cn[4 * j + 1] = 1.0; cn[4 * j + 2] = (double)aj;
      cn[4 * j] = (double)logl((long double)cn[4 * j + 2]);      // ln of the ROUNDED node (what the kernel evaluates at)
    }
    std::vector<long double> Tm(M);
    for (int z = 0; z < nz; z++) {
      const size_t base = (size_t)z * SN_NODES;
      const long double az = nodes_a[base], h = 1.0L - az;
      std::vector<long double> w(M, 0.0L), dd(M, 0.0L);
      for (int i = 0; i <= 16; i++) {       // i < 16: tabulated node i; i = 16: the end point a = 1
        const long double a = (i < 16) ? (long double)nodes_a[base + i] : 1.0L;
        int st = 0;                         // weight class: 0 = end points (half weight), j = new nodes of stage j + 1
        if (i >= 1 && i < 16) { st = 1; while ((2 << (st - 1)) <= i) st++; }
        const long double wi = (st == 0 ? 0.5L : 1.0L) * RW[st], ei = (st == 0 ? 0.5L : 1.0L) * RW[5 + st];
        long double x = (a - mid) / half;
        x = std::max<long double>(-1.0L, std::min<long double>(1.0L, x));
        const long double th = acosl(x), rsa = 1.0L / sqrtl(a);
        for (int m = 0; m < M; m++) {
          const long double t = cosl(m * th) * rsa;
          w[m] += wi * t; dd[m] += ei * t;
        }
      }
      for (int m = 0; m < M; m++) {
        W[(size_t)z * M + m] = (double)(h * w[m]);
        dmax[m] = std::max(dmax[m], (double)fabsl(dd[m]));      // |D[z][m]| / h_z
      }
    }
    if ((rc = dev_copy<double>(c, cn.data(), cn.size(), &D.cheb_nodes4))) return rc;
    if ((rc = dev_copy<double>(c, W.data(), W.size(), &D.cheb_W))) return rc;
    if ((rc = dev_copy<double>(c, dmax.data(), dmax.size(), &D.cheb_dmax))) return rc;
    // B fragments of the tensor-core kernel.  Columns are supernovae; a PRIMARY tile holds the first supernova of 8 distinct
    // redshifts, the further supernovae of those redshifts sit in the same column of the SECONDARY tiles that follow it
    // (they reuse the primary column's ss: only the chi^2 k-steps are read).  Redshifts ordered by multiplicity so that the
    // repeated ones share primary tiles (Union: 241 redshifts, 307 supernovae -> 31 primary + 12 secondary tiles).
    // Lane l of k-step ks holds B[k = l % 4][column = l / 4]: W of the column's redshift (ks < M / 4), then the chi^2
    // features.  An empty column repeats a supernova with sigma^2 = 1e300: its term vanishes.
    const int KS = M / 4 + 3;
    std::vector<int> tcol, tsec, tz;
    sn_tile_plan(nz, first, tcol, tsec, tz);
    const int ntile = (int)tsec.size();
    std::vector<double> Wf((size_t)ntile * KS * 32, 0.0);
    for (int t = 0; t < ntile; t++)
      for (int ks = 0; ks < KS; ks++)
        for (int l = 0; l < 32; l++) {
          const int col = tcol[(size_t)t * 8 + l / 4], k = l % 4, z = tz[(size_t)t * 8 + l / 4];
          const bool padcol = col < 0;
          const int r = padcol ? -col - 1 : col;
          const double *row = rows.data() + (size_t)r * SN_ROW;      // m s | c z | V0 Vss | Vcc Cms | Cmc Csc
          double v;
          if (ks < M / 4) v = W[(size_t)z * M + 4 * ks + k];
          else if (ks == M / 4) {
            const double lnaz = nodes[(size_t)z * SN_NODES].x;
            const double f[4] = {row[0] + (5.0 / M_LN10) * lnaz - SN_MU0, 1.0, row[1], row[2]};
            v = f[k];
          } else if (ks == M / 4 + 1) v = padcol ? (k == 0 ? 1.0e300 : 0.0) : row[4 + k];
          else v = (k < 2 && !padcol) ? row[8 + k] : 0.0;
          Wf[((size_t)t * KS + ks) * 32 + l] = v;
        }
    if ((rc = dev_copy<double>(c, Wf.data(), Wf.size(), &D.cheb_Wf))) return rc;
    if ((rc = dev_copy<int>(c, tsec.data(), tsec.size(), &D.sn_tile_sec))) return rc;
    D.sn_ntile = ntile;
    // TF32 tail of the coefficient contraction (m >= 16): W split into hi + lo TF32 numbers (cvt.rna: nearest, ties away),
    // in the B-fragment order of mma.m16n8k8: lane l holds b0 = B[k = l % 4][column l / 4], b1 = B[k = l % 4 + 4][column l / 4]
    // per k-step of 8 coefficients; per lane and tile: [k-step][hi b0, hi b1, lo b0, lo b1]
    {
      auto tf32 = [](float f) { uint32_t u; memcpy(&u, &f, 4); u = (u + 0x1000u) & 0xffffe000u; return u; };
      auto asf = [](uint32_t u) { float f; memcpy(&f, &u, 4); return f; };
      const int M64 = 16, nk2 = (M - M64 + 7) / 8;
      std::vector<uint32_t> Wt((size_t)ntile * 32 * 8, 0u);
      for (int t = 0; t < ntile; t++)
        for (int l = 0; l < 32; l++) {
          const int zc = tz[(size_t)t * 8 + l / 4];
          for (int k2 = 0; k2 < nk2 && k2 < 2; k2++)
            for (int j = 0; j < 2; j++) {
              const int mm = M64 + 8 * k2 + l % 4 + 4 * j;
              const double w = mm < M ? W[(size_t)zc * M + mm] : 0.0;
              const uint32_t hi = tf32((float)w), lo = tf32((float)(w - (double)asf(hi)));
              Wt[((size_t)t * 32 + l) * 8 + 4 * k2 + j] = hi;
              Wt[((size_t)t * 32 + l) * 8 + 4 * k2 + 2 + j] = lo;
            }
        }
      if ((rc = dev_copy<uint32_t>(c, Wt.data(), Wt.size(), &D.cheb_Wt))) return rc;
    }
  }
  return 0;
}

extern "C" int pmcb200_set_target(pmcb200_ctx *c, const pmcb200_target_t *t) {
  if (!c || !t) return PMCB200_ERR_ARG;
  CUDA_OK(c, cudaSetDevice(c->device));
  CUDA_OK(c, cudaStreamSynchronize(c->stream));
  free_target(c);
  if (t->npar < 1 || t->npar > PMCB200_MAX_DIM || t->ndata < 1 || t->ndata > PMCB200_MAX_DATA)
    return fail(c, PMCB200_ERR_DIM, "target: npar=%d ndata=%d out of range", t->npar, t->ndata);
  c->tgt = *t;
  const int d = t->npar;
  int rc;
  // flat box prior: logpr_default = -sum log(max-min), param.c:124-129
  double logpr = 0.0;
  std::vector<double> box(2 * d);
  for (int j = 0; j < d; j++) {
    if (!(t->max[j] > t->min[j])) return fail(c, PMCB200_ERR_ARG, "target: max[%d] <= min[%d]", j, j);
    logpr -= std::log(t->max[j] - t->min[j]);
    box[j] = t->min[j]; box[d + j] = t->max[j];
  }
  // special prior of data set 0 only, param.c:998-1001,1055-1101
  int special = t->like[0].special;
  if (special == PMCB200_SPECIAL_unity) for (int j = 0; j < d; j++) logpr += std::log(t->max[j] - t->min[j]);
  else if (special == PMCB200_SPECIAL_de_conservative) {   // volume term, param.c:1072-1094
    int iw0 = -1, iw1 = -1;
    for (int j = 0; j < d; j++) {
      if (t->like[0].par[j] == PMCB200_P_w0de) iw0 = j;
      if (t->like[0].par[j] == PMCB200_P_w1de) iw1 = j;
    }
    if (iw0 >= 0) {
      if (t->min[iw0] > -1.0 || t->max[iw0] < -1.0 / 3.0)
        return fail(c, PMCB200_ERR_ARG, "Range of w0_de [%g;%g] should not be smaller than de_conservative prior [-1;-1/3]",
                    t->min[iw0], t->max[iw0]);
      if (iw1 < 0) logpr += std::log(t->max[iw0] - t->min[iw0]) - std::log(2.0 / 3.0);
      else logpr += std::log(t->max[iw0] - t->min[iw0]) + std::log(t->max[iw1] - t->min[iw1])
                    - std::log(0.5 * 2.0 / 3.0 * 2.0 / 3.0 / (1.0 - DE_A_ACC)) - std::log(0.5 * 2.0 / 3.0 * 2.0 / 3.0);
    }
  } else if (special != PMCB200_SPECIAL_none)
    return fail(c, PMCB200_ERR_UNSUP, "target: special prior %d not supported", special);
  c->logpr_const = logpr;
  const double *dbox;
  if ((rc = dev_copy<double>(c, box.data(), box.size(), &dbox))) return rc;
  c->d_box = (double *)dbox;
  c->box_d = d;

  for (int i = 0; i < t->ndata; i++) {
    const pmcb200_like_t &L = t->like[i];
    DevLike &D = c->like[i];
    memset(&D, 0, sizeof(D));
    D.kind = L.kind; D.npar = L.npar; D.special = L.special; D.model = L.model;
    if (L.npar != d) return fail(c, PMCB200_ERR_DIM, "like %d: npar %d != %d", i, L.npar, d);
    for (int j = 0; j < d; j++) D.par[j] = L.par[j];
    {   // launch-uniform specialisation: the second exponent term exists only if w1 can be non-zero
      bool has_w1 = L.model.w1_de != 0.0 || L.model.de_param == PMCB200_DE_jassal;
      for (int j = 0; j < L.npar; j++) if (L.par[j] == PMCB200_P_w1de) has_w1 = true;
      D.sn_hasq = has_w1 ? 1 : 0;
    }
    switch (L.kind) {
      case PMCB200_LIKE_SNIa:
        if ((rc = build_sn(c, L, D))) return rc;
        break;
      case PMCB200_LIKE_BAO: {
        if (L.g_ndim < 1 || L.g_ndim > 4 || !L.g_z || !L.g_mean || !L.g_chol)
          return fail(c, PMCB200_ERR_ARG, "BAO: bad data (ndim=%d)", L.g_ndim);
        D.bao_method = L.bao_method; D.g_ndim = L.g_ndim;
        int nz = L.g_ndim * (L.bao_method == PMCB200_BAO_distance_D_V_ratio ? 2 : 1);
        if ((rc = dev_copy<double>(c, L.g_z, nz, &D.g_z))) return rc;
        if ((rc = pack_gauss(c, L.g_ndim, L.g_mean, L.g_chol, &D.g_comp))) return rc;
        break;
      }
      case PMCB200_LIKE_CMBDistPrior:
        if (L.g_ndim < 3 || L.g_ndim > 4 || !L.g_mean || !L.g_chol)
          return fail(c, PMCB200_ERR_ARG, "CMBDistPrior: ndim=%d (3 or 4)", L.g_ndim);
        D.g_ndim = L.g_ndim;
        if ((rc = pack_gauss(c, L.g_ndim, L.g_mean, L.g_chol, &D.g_comp))) return rc;
        break;
      case PMCB200_LIKE_Mvdens:
      case PMCB200_LIKE_MixMvdens: {
        int K = L.kind == PMCB200_LIKE_Mvdens ? 1 : L.mix_ncomp;
        if (K < 1 || K > PMCB200_MAX_COMP || L.mix_ndim != d || !L.mix_mean || !L.mix_chol)
          return fail(c, PMCB200_ERR_DIM, "Mvdens target: ncomp=%d ndim=%d (npar=%d)", K, L.mix_ndim, d);
        std::vector<double> w(K, 1.0);
        if (L.mix_wght) w.assign(L.mix_wght, L.mix_wght + K);
        double *dm = nullptr; size_t cap = 0;
        if ((rc = upload_mix(c, K, d, L.mix_df, w.data(), L.mix_mean, L.mix_chol, D.mixh, &dm, &cap))) return rc;
        c->tgt_allocs.push_back(dm);
        D.mix = dm;
        break;
      }
      case PMCB200_LIKE_BANANA:
        if (d < 2) return fail(c, PMCB200_ERR_DIM, "banana target needs d >= 2");
        D.banana_b = L.banana_b; D.banana_sigma1sq = L.banana_sigma1sq;
        break;
      default:
        return fail(c, PMCB200_ERR_UNSUP, "likelihood kind %d not supported on the device path", L.kind);
    }
  }
  // Gaussian prior, param.c:1009-1026
  if (t->prior_mean) {
    int np = t->prior_ndim;
    if (np < 1 || np > d || !t->prior_chol) return fail(c, PMCB200_ERR_DIM, "prior: ndim=%d", np);
    std::vector<int> sel;
    if (t->nprior > 0) { for (int j = 0; j < d; j++) if (t->indprior[j] == 1) sel.push_back(j); }
    else for (int j = 0; j < np; j++) sel.push_back(j);
    if ((int)sel.size() != np) return fail(c, PMCB200_ERR_DIM, "prior: %d selected parameters != ndim %d", (int)sel.size(), np);
    double one = 1.0, *dm = nullptr; size_t cap = 0;
    if ((rc = upload_mix(c, 1, np, -1, &one, t->prior_mean, t->prior_chol, c->prior_h, &dm, &cap))) return rc;
    c->tgt_allocs.push_back(dm);
    c->d_prior = dm;
    const int *dsel;
    if ((rc = dev_copy<int>(c, sel.data(), sel.size(), &dsel))) return rc;
    c->d_prior_sel = (int *)dsel;
  }
  c->have_target = true;
  return 0;
}

// ---- stage launches ------------------------------------------------------------------
static int need(pmcb200_ctx *c, bool prop, bool tgt) {
  if (!c) return PMCB200_ERR_ARG;
  if (prop && !c->have_prop) return fail(c, PMCB200_ERR_STATE, "proposal not set (pmcb200_set_proposal)");
  if (tgt && !c->have_target) return fail(c, PMCB200_ERR_STATE, "target not set (pmcb200_set_target)");
  if (prop && tgt && c->h.d != c->tgt.npar)
    return fail(c, PMCB200_ERR_DIM, "proposal ndim %d != target npar %d", c->h.d, c->tgt.npar);
  cudaError_t e = cudaSetDevice(c->device);
  if (e != cudaSuccess) return fail(c, PMCB200_ERR_CUDA, "cudaSetDevice: %s", cudaGetErrorString(e));
  return 0;
}

static int reset_scal(pmcb200_ctx *c) {
  CUDA_OK(c, cudaMemsetAsync(c->d_scal, 0, sizeof(DevScal), c->stream));
  return 0;
}

static int launch_simulate(pmcb200_ctx *c, int64_t N, uint64_t seed, uint32_t iter, int64_t offset,
                           double *dX, int32_t *didx, int16_t *dflg) {
  NvtxRange nvtx_("pmc:simulate");
  if (N <= 0) return 0;
  if (!c->d_box) return fail(c, PMCB200_ERR_STATE, "simulate needs the target's box (pmcb200_set_target)");
  MixArgs a; a.mix = c->d_mix; a.h = c->h; a.box = c->d_box; a.N = N; a.seed = seed; a.iter = iter;
  a.offset = offset; a.X = dX; a.idx = didx; a.flg = dflg; a.scal = c->d_scal;
  c->rho_valid = false;      // the sample array is being rewritten
  MIX_OK(c, OP_SIMULATE, a);
  return 0;
}

static int launch_posterior(pmcb200_ctx *c, int64_t N, const double *dX, const int16_t *dflg,
                            double *dlogpi, int32_t *derr) {
  NvtxRange nvtx_("pmc:posterior");
  if (N <= 0) return 0;
  const int d = c->tgt.npar;
  for (int i = 0; i < c->tgt.ndata; i++) {
    const DevLike &L = c->like[i];
    const int set = (i == 0);
    const double add = set ? c->logpr_const : 0.0;
    if (L.kind == PMCB200_LIKE_Mvdens || L.kind == PMCB200_LIKE_MixMvdens) {
      MixArgs a; a.mix = L.mix; a.h = L.mixh; a.is_mixture = (L.kind == PMCB200_LIKE_MixMvdens); a.N = N;
      a.Xc = dX; a.dX = d; a.sel = nullptr; a.flgc = dflg; a.logpi = dlogpi; a.err = derr; a.set = set;
      a.add_const = add;
      MIX_OK(c, OP_LIKE_MIX, a);
      continue;
    }
    uint32_t *fb = nullptr;
    if (pmc_sn_spectral_wanted(L, N)) {       // work list for the samples the spectral kernel hands to the exact one
      int rc = ensure(c, c->sFb, (size_t)N * sizeof(uint32_t));
      if (rc) return rc;
      fb = (uint32_t *)c->sFb.p;
    }
    pmc_launch_like(L, N, dX, d, dflg, dlogpi, derr, set, add, c->d_cnt, fb, c->d_fb_cnt, c->stream);
    if (fb) c->launches += 2;
    LAUNCH_OK(c);
  }
  if (c->d_prior) {
    MixArgs a; a.mix = c->d_prior; a.h = c->prior_h; a.is_mixture = 0; a.N = N; a.Xc = dX; a.dX = d;
    a.sel = c->d_prior_sel; a.flgc = dflg; a.logpi = dlogpi; a.err = derr; a.set = 0; a.add_const = 0.0;
    MIX_OK(c, OP_LIKE_MIX, a);
  }
  return 0;
}

static int launch_weights(pmcb200_ctx *c, int64_t N, const double *dX, const double *dlogpi,
                          const int32_t *derr, double beta, int16_t *dflg, double *dlogw) {
  NvtxRange nvtx_("pmc:weights");
  if (N <= 0) return 0;
  MixArgs a; a.mix = c->d_mix; a.h = c->h; a.N = N; a.Xc = dX; a.logpic = dlogpi; a.errc = derr; a.beta = beta;
  a.flg = dflg; a.logw = dlogw; a.scal = c->d_scal;
  // E-step cache: 16 K bytes of HBM traffic per sample instead of the K whitenings + exps of the EM kernel's phase 1
  c->rho_valid = false;
  int written = 0;
  const int K = c->h.K, d = c->h.d;
  if (!c->em_no_rho && !c->em_no_mma && pmc_pad_dim(d) >= c->rho_min_dim && c->h.df <= 0 && em_mma_ok(K, d, 0)) {
    const size_t bytes = (size_t)N * K * sizeof(double);
    if (c->sRho.cap < bytes) {        // soft: without the buffer the EM kernel simply recomputes
      if (c->sRho.p) cudaFree(c->sRho.p);
      c->sRho.p = nullptr; c->sRho.cap = 0;
      if (cudaMalloc(&c->sRho.p, bytes) == cudaSuccess) c->sRho.cap = bytes;
      else { c->sRho.p = nullptr; cudaGetLastError(); }
    }
    if (c->sRho.p) { a.rho = (double *)c->sRho.p; a.rho_written = &written; }
  }
  MIX_OK(c, OP_WEIGHTS, a);
  if (written) { c->rho_valid = true; c->rho_X = dX; c->rho_N = N; c->rho_ver = c->prop_ver; }
  return 0;
}

static int launch_em_local(pmcb200_ctx *c, int64_t N, const double *dX, const int32_t *didx,
                           const int16_t *dflg, const double *dlogw, double *dblock, int linear = 0) {
  NvtxRange nvtx_("pmc:em_local");
  const int K = c->h.K, d = c->h.d;
  const int64_t len = stat_len(K, d);
  int64_t ntiles = (N + PMC_BLOCK - 1) / PMC_BLOCK;
  int blocks = (int)std::max<int64_t>(1, std::min<int64_t>(c->em_blocks, ntiles));
  // component groups: one launch unless the K x 256 shared-memory arrays exceed the budget
  int Kg = K;
  const bool mma = em_mma_ok(K, d, c->h.df > 0) && !c->em_no_mma;     // FP64 tensor-core kernel: all components in one launch
  while (!mma && Kg > 1 && em_smem_bytes(Kg, d, c->h.df > 0) > 200 * 1024) Kg = (Kg + 1) / 2;
  MixArgs a; a.mix = c->d_mix; a.h = c->h; a.N = N; a.Xc = dX; a.idxc = didx; a.flgc = dflg; a.logwc = dlogw;
  a.scal = c->d_scal; a.partials = c->d_partials; a.blocks = blocks; a.linear = linear; a.em_mma = mma;
  if (mma && !linear && c->rho_valid && c->rho_X == dX && c->rho_N == N && c->rho_ver == c->prop_ver)
    a.rho_in = (const double *)c->sRho.p;
  c->rho_valid = false;      // one weight pass feeds one update
  int used = blocks;
  a.nblocks_out = &used;
  for (int k0 = 0; k0 < K; k0 += Kg) {
    a.k0 = k0; a.Kg = std::min(Kg, K - k0);
    a.smem = mma ? em_mma_smem_bytes(K, d, c->h.df > 0) : em_smem_bytes(a.Kg, d, c->h.df > 0);
    a.blocks = (k0 == 0) ? blocks : used;     // every group must use the same grid (shared partials layout)
    MIX_OK(c, OP_EM, a);
  }
  pmc_launch_em_reduce(c->d_partials, used, len, c->d_scal, N, dblock, c->stream);
  LAUNCH_OK(c);
  return 0;
}

// ---- C-ABI stage entry points ------------------------------------------------------------
extern "C" int pmcb200_simulate(pmcb200_ctx *c, int64_t N, uint64_t seed, uint32_t iter, int64_t offset,
                                double *dX, int32_t *didx, int16_t *dflg) {
  int rc = need(c, true, false);
  if (rc) return rc;
  if (N < 0 || (N > 0 && (!dX || !didx || !dflg))) return fail(c, PMCB200_ERR_ARG, "simulate: bad arguments");
  if ((rc = reset_scal(c))) return rc;
  return launch_simulate(c, N, seed, iter, offset, dX, didx, dflg);
}

extern "C" int pmcb200_simulate_from_draws(pmcb200_ctx *c, int64_t N, const double *du, const double *dz,
                                           double *dX, int32_t *didx, int16_t *dflg) {
  int rc = need(c, true, false);
  if (rc) return rc;
  if (N < 0 || (N > 0 && (!du || !dz || !dX || !didx || !dflg))) return fail(c, PMCB200_ERR_ARG, "simulate_from_draws: bad arguments");
  if (!c->d_box) return fail(c, PMCB200_ERR_STATE, "simulate needs the target's box (pmcb200_set_target)");
  if (c->h.df > 0) return fail(c, PMCB200_ERR_UNSUP, "simulate_from_draws: Gaussian proposals only");
  if (N == 0) return 0;
  MixArgs a; a.mix = c->d_mix; a.h = c->h; a.box = c->d_box; a.N = N; a.U = du; a.Z = dz; a.X = dX; a.idx = didx;
  a.flg = dflg;
  c->rho_valid = false;
  MIX_OK(c, OP_SIMULATE_DRAWS, a);
  return 0;
}

extern "C" int pmcb200_proposal_log_pdf(pmcb200_ctx *c, int64_t N, const double *dX, double *dlogq) {
  int rc = need(c, true, false);
  if (rc) return rc;
  if (N < 0 || (N > 0 && (!dX || !dlogq))) return fail(c, PMCB200_ERR_ARG, "proposal_log_pdf: bad arguments");
  if (N == 0) return 0;
  MixArgs a; a.mix = c->d_mix; a.h = c->h; a.N = N; a.Xc = dX; a.out = dlogq;
  MIX_OK(c, OP_LOGQ, a);
  return 0;
}

extern "C" int pmcb200_posterior_log_pdf(pmcb200_ctx *c, int64_t N, const double *dX, double *dlogpi,
                                         int32_t *derr) {
  int rc = need(c, false, true);
  if (rc) return rc;
  if (N < 0 || (N > 0 && (!dX || !dlogpi))) return fail(c, PMCB200_ERR_ARG, "posterior_log_pdf: bad arguments");
  return launch_posterior(c, N, dX, nullptr, dlogpi, derr);
}

extern "C" int pmcb200_map_params(pmcb200_ctx *c, int idata, int64_t N, const double *dX, double *dout, int32_t *derr) {
  int rc = need(c, false, true);
  if (rc) return rc;
  if (idata < 0 || idata >= c->tgt.ndata || N < 0 || (N > 0 && (!dX || !dout)))
    return fail(c, PMCB200_ERR_ARG, "map_params: bad arguments");
  const int kind = c->like[idata].kind;
  if (kind != PMCB200_LIKE_SNIa && kind != PMCB200_LIKE_BAO && kind != PMCB200_LIKE_CMBDistPrior)
    return fail(c, PMCB200_ERR_UNSUP, "map_params: data set %d has no cosmological parameters", idata);
  if (N == 0) return 0;
  pmc_launch_map_params(c->like[idata], N, dX, c->tgt.npar, dout, derr, c->stream);
  LAUNCH_OK(c);
  return 0;
}

// Fisher matrix at a point: go_fishing.c:37-85, all stencil points in one posterior launch.
extern "C" int pmcb200_fisher_host(pmcb200_ctx *c, const double *pos, const double *h, int diag_only,
                                   double *F, int *nbad) {
  int rc = need(c, false, true);
  if (rc) return rc;
  if (!pos || !h || !F) return fail(c, PMCB200_ERR_ARG, "fisher_host: bad arguments");
  const int d = c->tgt.npar;
  for (int a = 0; a < d; a++)
    if (!(h[a] >= 1.0e-20) || !std::isfinite(h[a])) return fail(c, PMCB200_ERR_ARG, "fisher_host: h[%d] too small", a);
  static const int diff[4][2] = {{+1, +1}, {+1, -1}, {-1, +1}, {-1, -1}};
  // stencil: per element (a, b >= a) the points j = 0..3 (j = 2 of a diagonal element repeats j = 1)
  std::vector<double> pts;
  std::vector<int> first;           // index of point j = 0 of each element, elements in (a, b) order
  for (int a = 0; a < d; a++)
    for (int b = a; b < d; b++) {
      if (diag_only && a != b) continue;
      first.push_back((int)(pts.size() / d));
      for (int j = 0; j < 4; j++) {
        if (j == 2 && a == b) continue;
        const size_t o = pts.size();
        pts.insert(pts.end(), pos, pos + d);
        pts[o + a] += diff[j][0] * h[a];        // two separate additions, as the reference's loop over k
        pts[o + b] += diff[j][1] * h[b];
      }
    }
  const int64_t np = (int64_t)(pts.size() / d);
  if ((rc = ensure(c, c->sFish, (size_t)np * d * sizeof(double)))) return rc;
  if ((rc = ensure(c, c->sLogpi, (size_t)np * sizeof(double)))) return rc;
  if ((rc = ensure(c, c->sErr, (size_t)np * sizeof(int32_t)))) return rc;
  CUDA_OK(c, cudaMemcpyAsync(c->sFish.p, pts.data(), (size_t)np * d * sizeof(double), cudaMemcpyHostToDevice, c->stream));
  c->rho_valid = false;
  if ((rc = launch_posterior(c, np, (const double *)c->sFish.p, nullptr, (double *)c->sLogpi.p, (int32_t *)c->sErr.p))) return rc;
  std::vector<double> lp(np);
  std::vector<int32_t> er(np);
  CUDA_OK(c, cudaMemcpyAsync(lp.data(), c->sLogpi.p, (size_t)np * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
  CUDA_OK(c, cudaMemcpyAsync(er.data(), c->sErr.p, (size_t)np * sizeof(int32_t), cudaMemcpyDeviceToHost, c->stream));
  CUDA_OK(c, cudaStreamSynchronize(c->stream));
  int bad = 0;
  for (int64_t i = 0; i < np; i++) bad += (er[i] != 0) || !std::isfinite(lp[i]);
  if (nbad) *nbad = bad;
  if (bad) return fail(c, PMCB200_ERR_ARG, "fisher_host: likelihood error at %d of %lld stencil points", bad, (long long)np);
  for (int i = 0; i < d * d; i++) F[i] = 0.0;
  size_t e = 0;
  for (int a = 0; a < d; a++)
    for (int b = a; b < d; b++) {
      if (diag_only && a != b) continue;
      const double *q = lp.data() + first[e++];
      const double c0 = q[0], c1 = q[1], c2 = (a == b) ? q[1] : q[2], c3 = (a == b) ? q[2] : q[3];
      const double f = -(c0 - c1 - c2 + c3) / (4.0 * h[a] * h[b]);
      F[a * d + b] = f;
      F[b * d + a] = f;
    }
  return 0;
}

extern "C" int pmcb200_importance_weights(pmcb200_ctx *c, int64_t N, const double *dX, double beta,
                                          int16_t *dflg, double *dlogw) {
  int rc = need(c, true, true);
  if (rc) return rc;
  if (N < 0 || (N > 0 && (!dX || !dflg || !dlogw))) return fail(c, PMCB200_ERR_ARG, "importance_weights: bad arguments");
  if ((rc = ensure(c, c->sLogpi, (size_t)std::max<int64_t>(N, 1) * sizeof(double)))) return rc;
  if ((rc = ensure(c, c->sErr, (size_t)std::max<int64_t>(N, 1) * sizeof(int32_t)))) return rc;
  // keep nok_box of a preceding simulate; reset max / nok
  CUDA_OK(c, cudaMemsetAsync(c->d_scal, 0, 2 * sizeof(unsigned long long), c->stream));
  if ((rc = launch_posterior(c, N, dX, dflg, (double *)c->sLogpi.p, (int32_t *)c->sErr.p))) return rc;
  return launch_weights(c, N, dX, (double *)c->sLogpi.p, (int32_t *)c->sErr.p, beta, dflg, dlogw);
}

extern "C" int pmcb200_normalize_weights(pmcb200_ctx *c, int64_t N, const int16_t *dflg, double *dw) {
  int rc = need(c, false, false);
  if (rc) return rc;
  if (N < 0 || (N > 0 && (!dflg || !dw))) return fail(c, PMCB200_ERR_ARG, "normalize_weights: bad arguments");
  // M and S come from the last em_finish (they are global over all ranks)
  double M = c->h_result[0], S = c->h_result[1];
  if (!(S > 0.0)) return fail(c, PMCB200_ERR_STATE, "normalize_weights: call pmcb200_em_finish first");
  if (N == 0) return 0;
  pmc_launch_normalize(N, dflg, dw, M, 1.0 / S, c->stream);
  LAUNCH_OK(c);
  return 0;
}

extern "C" int64_t pmcb200_stat_block_len(const pmcb200_ctx *c) {
  if (!c || !c->have_prop) return 0;
  return stat_len(c->h.K, c->h.d);
}

extern "C" int pmcb200_em_local(pmcb200_ctx *c, int64_t N, const double *dX, const int32_t *didx,
                                const int16_t *dflg, const double *dlogw, double *dblock) {
  int rc = need(c, true, false);
  if (rc) return rc;
  if (N < 0 || !dblock || (N > 0 && (!dX || !didx || !dflg || !dlogw)))
    return fail(c, PMCB200_ERR_ARG, "em_local: bad arguments");
  return launch_em_local(c, N, dX, didx, dflg, dlogw, dblock);
}

extern "C" int pmcb200_em_finish(pmcb200_ctx *c, int nranks, const double *dall, int64_t N_global,
                                 pmcb200_stats_t *stats) {
  int rc = need(c, true, false);
  if (rc) return rc;
  if (nranks < 1 || nranks > 64 || !dall || N_global < 1) return fail(c, PMCB200_ERR_ARG, "em_finish: bad arguments");
  NvtxRange nvtx_("pmc:em_finish");
  const int K = c->h.K, d = c->h.d;
  pmc_launch_em_finish(c->d_mix, c->h, nranks, dall, N_global, c->d_work, c->d_result, c->d_fin_cnt, c->stream);
  LAUNCH_OK(c);
  size_t rlen = RES_HDR + (size_t)K * (1 + d + (size_t)d * d);
  CUDA_OK(c, cudaMemcpyAsync(c->h_result, c->d_result, rlen * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
  CUDA_OK(c, cudaStreamSynchronize(c->stream));
  const double *r = c->h_result;
  pmcb200_stats_t st;
  memset(&st, 0, sizeof(st));
  st.nsamples = N_global;
  st.maxW = r[0]; st.sum_shift = r[1]; st.logSum = r[2]; st.perplexity = r[3]; st.ess = r[4];
  st.ln_evidence = r[5]; st.enc = r[6]; st.ndead = (int32_t)r[7];
  st.nok = (int64_t)r[8]; st.nok_box = (int64_t)r[9];
  if (stats) *stats = st;
  if (st.nok == 0 || !(st.sum_shift > 0.0)) {
    c->h_result[1] = 0.0;
    return fail(c, PMCB200_ERR_NOSAMPLE, "no sample with finite importance weight (nok_box=%lld)", (long long)st.nok_box);
  }
  const double *w = r + RES_HDR, *mean = w + K, *chol = mean + (size_t)K * d;
  double wsum = 0.0;
  for (int k = 0; k < K; k++) wsum += w[k];
  if (!(wsum > 0.0)) return fail(c, PMCB200_ERR_NOSAMPLE, "all proposal components died in the update");
  std::vector<double> w2(w, w + K), m2(mean, mean + (size_t)K * d), c2(chol, chol + (size_t)K * d * d);
  double M = r[0], S = r[1];
  rc = install_proposal(c, K, d, c->h.df, w2.data(), m2.data(), c2.data());
  c->h_result[0] = M; c->h_result[1] = S;
  return rc;
}

// ---- whole iteration -----------------------------------------------------------------------
// shard iteration with the D2H of the finished arrays overlapped with the
// likelihood kernel (copy stream); leaves the statistics block in dblock
static int iteration_core(pmcb200_ctx *c, int64_t N, uint64_t seed, uint32_t iter, int64_t offset, double beta,
                          double *dX, int32_t *didx, int16_t *dflg, double *dlogw, double *dblock,
                          double *hX, int32_t *hidx, int16_t *hflg) {
  int rc;
  const int d = c->h.d;
  const size_t n1 = (size_t)std::max<int64_t>(N, 1), n = (size_t)N;
  if (!dX || !didx || !dflg || !dlogw) {
    // library scratch: stay on the current set unless an earlier iteration's copies are still draining from it
    if (c->set[c->cur].busy) {
      if (cudaEventQuery(c->set[c->cur].copied) == cudaSuccess) c->set[c->cur].busy = false;
      else {
        c->cur ^= 1;
        if (c->set[c->cur].busy) {            // both in flight: the kernels wait for the older set's copies
          CUDA_OK(c, cudaStreamWaitEvent(c->stream, c->set[c->cur].copied, 0));
          c->set[c->cur].busy = false;
        }
      }
      cudaGetLastError();
    }
    c->set[c->cur].seq = ++c->seq;
  }
  pmcb200_ctx::ScratchSet &t = c->set[c->cur];
  if (!dX) { if ((rc = ensure(c, t.X, n1 * d * sizeof(double)))) return rc; dX = (double *)t.X.p; }
  if (!didx) { if ((rc = ensure(c, t.Idx, n1 * sizeof(int32_t)))) return rc; didx = (int32_t *)t.Idx.p; }
  if (!dflg) { if ((rc = ensure(c, t.Flg, n1 * sizeof(int16_t)))) return rc; dflg = (int16_t *)t.Flg.p; }
  if (!dlogw) { if ((rc = ensure(c, t.Logw, n1 * sizeof(double)))) return rc; dlogw = (double *)t.Logw.p; }
  if ((rc = ensure(c, c->sLogpi, n1 * sizeof(double)))) return rc;
  if ((rc = ensure(c, c->sErr, n1 * sizeof(int32_t)))) return rc;
  if ((rc = reset_scal(c))) return rc;
  if ((rc = launch_simulate(c, N, seed, iter, offset, dX, didx, dflg))) return rc;
  if (N > 0 && (hX || hidx)) {   // X and idx are final after the sampler
    CUDA_OK(c, cudaEventRecord(c->ev_a, c->stream));
    CUDA_OK(c, cudaStreamWaitEvent(c->copy_stream, c->ev_a, 0));
    if (hX) CUDA_OK(c, cudaMemcpyAsync(hX, dX, n * d * sizeof(double), cudaMemcpyDeviceToHost, c->copy_stream));
    if (hidx) CUDA_OK(c, cudaMemcpyAsync(hidx, didx, n * sizeof(int32_t), cudaMemcpyDeviceToHost, c->copy_stream));
  }
  if ((rc = launch_posterior(c, N, dX, dflg, (double *)c->sLogpi.p, (int32_t *)c->sErr.p))) return rc;
  if ((rc = launch_weights(c, N, dX, (double *)c->sLogpi.p, (int32_t *)c->sErr.p, beta, dflg, dlogw))) return rc;
  if (N > 0 && hflg) {           // flags are final after the weight stage
    CUDA_OK(c, cudaEventRecord(c->ev_b, c->stream));
    CUDA_OK(c, cudaStreamWaitEvent(c->copy_stream, c->ev_b, 0));
    CUDA_OK(c, cudaMemcpyAsync(hflg, dflg, n * sizeof(int16_t), cudaMemcpyDeviceToHost, c->copy_stream));
  }
  return launch_em_local(c, N, dX, didx, dflg, dlogw, dblock);
}

extern "C" int pmcb200_iteration_local(pmcb200_ctx *c, int64_t N, uint64_t seed, uint32_t iter,
                                       int64_t offset, double beta, double *dX, int32_t *didx,
                                       int16_t *dflg, double *dlogw, double *dblock) {
  int rc = need(c, true, true);
  if (rc) return rc;
  if (N < 0 || !dblock) return fail(c, PMCB200_ERR_ARG, "iteration_local: bad arguments");
  return iteration_core(c, N, seed, iter, offset, beta, dX, didx, dflg, dlogw, dblock, nullptr, nullptr, nullptr);
}

extern "C" int pmcb200_iteration_shard_host(pmcb200_ctx *c, int64_t N, uint64_t seed, uint32_t iter,
                                            int64_t offset, double beta, double *hX, int32_t *hidx,
                                            int16_t *hflg, double *dblock) {
  int rc = need(c, true, true);
  if (rc) return rc;
  if (N < 0 || !dblock) return fail(c, PMCB200_ERR_ARG, "iteration_shard_host: bad arguments");
  return iteration_core(c, N, seed, iter, offset, beta, nullptr, nullptr, nullptr, nullptr, dblock, hX, hidx, hflg);
}

// normalised weights of the current scratch set to the host, queued behind the set's other copies; marks the
// set as draining (pmcb200_host_wait, or the next-but-one iteration, waits for it)
extern "C" int pmcb200_shard_weights_host_begin(pmcb200_ctx *c, int64_t N, double *hw) {
  int rc = need(c, true, false);
  if (rc) return rc;
  NvtxRange nvtx_("pmc:weights_to_host");
  pmcb200_ctx::ScratchSet &t = c->set[c->cur];
  if (N < 0 || (hw && (size_t)N * sizeof(double) > t.Logw.cap)) return fail(c, PMCB200_ERR_ARG, "shard_weights_host: bad arguments");
  if (N > 0 && hw) {
    if ((rc = pmcb200_normalize_weights(c, N, (int16_t *)t.Flg.p, (double *)t.Logw.p))) return rc;
    CUDA_OK(c, cudaEventRecord(c->ev_b, c->stream));
    CUDA_OK(c, cudaStreamWaitEvent(c->copy_stream, c->ev_b, 0));
    CUDA_OK(c, cudaMemcpyAsync(hw, t.Logw.p, (size_t)N * sizeof(double), cudaMemcpyDeviceToHost, c->copy_stream));
  }
  CUDA_OK(c, cudaEventRecord(t.copied, c->copy_stream));
  t.busy = true;
  return 0;
}

// The sample array of the most recent iteration that used the library's scratch (hX was NULL in the iteration call:
// X and the component indices stay in HBM for the device post-processing) to the host, on request -- the reference needs it only for the pmcsim
// dump (cosmo_pmc.c:392) and the host post-processing.  Queued on the copy stream like the other arrays.
extern "C" int pmcb200_samples_host_begin(pmcb200_ctx *c, int64_t N, double *hX, int32_t *hidx) {
  int rc = need(c, true, false);
  if (rc) return rc;
  NvtxRange nvtx_("pmc:samples_to_host");
  pmcb200_ctx::ScratchSet &t = c->set[c->cur];
  const size_t n = (size_t)std::max<int64_t>(N, 0), bytes = n * c->h.d * sizeof(double);
  if (N < 0 || (!hX && !hidx) || (hX && (!t.X.p || bytes > t.X.cap)) || (hidx && (!t.Idx.p || n * sizeof(int32_t) > t.Idx.cap)))
    return fail(c, PMCB200_ERR_ARG, "samples_host: no such sample array in the scratch set");
  if (N > 0) {
    CUDA_OK(c, cudaEventRecord(c->ev_a, c->stream));
    CUDA_OK(c, cudaStreamWaitEvent(c->copy_stream, c->ev_a, 0));
    if (hX) CUDA_OK(c, cudaMemcpyAsync(hX, t.X.p, bytes, cudaMemcpyDeviceToHost, c->copy_stream));
    if (hidx) CUDA_OK(c, cudaMemcpyAsync(hidx, t.Idx.p, n * sizeof(int32_t), cudaMemcpyDeviceToHost, c->copy_stream));
  }
  CUDA_OK(c, cudaEventRecord(t.copied, c->copy_stream));
  t.busy = true;
  return 0;
}

// wait for the host arrays: lag = 0 all iterations begun so far, lag = 1 all but the most recent one
extern "C" int pmcb200_host_wait(pmcb200_ctx *c, int lag) {
  if (!c || lag < 0) return PMCB200_ERR_ARG;
  CUDA_OK(c, cudaSetDevice(c->device));
  for (auto &t : c->set)
    if (t.busy && t.seq + (uint64_t)lag <= c->seq) {
      CUDA_OK(c, cudaEventSynchronize(t.copied));
      t.busy = false;
    }
  if (lag == 0) {
    CUDA_OK(c, cudaStreamSynchronize(c->copy_stream));
    CUDA_OK(c, cudaStreamSynchronize(c->stream));
  }
  return 0;
}

extern "C" int pmcb200_shard_weights_host(pmcb200_ctx *c, int64_t N, double *hw) {
  if (!hw) return fail(c, PMCB200_ERR_ARG, "shard_weights_host: bad arguments");
  int rc = pmcb200_shard_weights_host_begin(c, N, hw);
  if (rc) return rc;
  return pmcb200_host_wait(c, 0);
}

extern "C" int pmcb200_iteration_host_begin(pmcb200_ctx *c, int64_t N, uint64_t seed, uint32_t iter, double beta,
                                            double *hX, int32_t *hidx, int16_t *hflg, double *hw,
                                            pmcb200_stats_t *stats) {
  int rc = need(c, true, true);
  if (rc) return rc;
  if (N < 1) return fail(c, PMCB200_ERR_ARG, "iteration_host: N = %lld", (long long)N);
  if ((rc = ensure(c, c->sBlock, (size_t)stat_len(c->h.K, c->h.d) * sizeof(double)))) return rc;
  double *dblock = (double *)c->sBlock.p;
  if ((rc = iteration_core(c, N, seed, iter, 0, beta, nullptr, nullptr, nullptr, nullptr, dblock, hX, hidx, hflg))) return rc;
  rc = pmcb200_em_finish(c, 1, dblock, N, stats);
  int rc2 = pmcb200_shard_weights_host_begin(c, N, rc == 0 ? hw : nullptr);
  return rc ? rc : rc2;
}

extern "C" int pmcb200_iteration_host(pmcb200_ctx *c, int64_t N, uint64_t seed, uint32_t iter, double beta,
                                      double *hX, int32_t *hidx, int16_t *hflg, double *hw,
                                      pmcb200_stats_t *stats) {
  int rc = pmcb200_iteration_host_begin(c, N, seed, iter, beta, hX, hidx, hflg, hw, stats);
  int rc2 = c ? pmcb200_host_wait(c, 0) : 0;
  return rc ? rc : rc2;
}


// ---- several contexts in one process --------------------------------------------
extern "C" int pmcb200_h2d_async(pmcb200_ctx *c, void *dptr, const void *hptr, size_t bytes) {
  if (!c) return PMCB200_ERR_ARG;
  CUDA_OK(c, cudaSetDevice(c->device));
  c->rho_valid = false;
  if (bytes) CUDA_OK(c, cudaMemcpyAsync(dptr, hptr, bytes, cudaMemcpyHostToDevice, c->stream));
  return 0;
}
extern "C" int pmcb200_d2h_async(pmcb200_ctx *c, void *hptr, const void *dptr, size_t bytes) {
  if (!c) return PMCB200_ERR_ARG;
  CUDA_OK(c, cudaSetDevice(c->device));
  if (bytes) CUDA_OK(c, cudaMemcpyAsync(hptr, dptr, bytes, cudaMemcpyDeviceToHost, c->stream));
  return 0;
}

// Page-locked host lumps are remembered so that pmcb200_host_free can tell them from the
// malloc fallback (no device: host-only callers such as the CPU tests still get memory).
static std::vector<void *> g_pinned;
extern "C" int pmcb200_host_alloc(size_t bytes, void **hptr) {
  if (!hptr) return PMCB200_ERR_ARG;
  *hptr = nullptr;
  int n = 0;
  if (cudaGetDeviceCount(&n) == cudaSuccess && n > 0 &&
      cudaHostAlloc(hptr, std::max<size_t>(bytes, 1), cudaHostAllocPortable) == cudaSuccess) {
    g_pinned.push_back(*hptr);
    return 0;
  }
  cudaGetLastError();
  *hptr = malloc(std::max<size_t>(bytes, 1));
  return *hptr ? 0 : PMCB200_ERR_ARG;
}
extern "C" int pmcb200_host_free(void *hptr) {
  if (!hptr) return 0;
  for (size_t i = 0; i < g_pinned.size(); i++)
    if (g_pinned[i] == hptr) {
      g_pinned.erase(g_pinned.begin() + i);
      return cudaFreeHost(hptr) == cudaSuccess ? 0 : PMCB200_ERR_CUDA;
    }
  free(hptr);
  return 0;
}

extern "C" int pmcb200_allgather_blocks(pmcb200_ctx *const *ctx, int n, double *const *dblock,
                                        double *const *dall, int64_t len) {
  if (!ctx || n < 1 || n > 64 || !dblock || !dall || len < 1) return PMCB200_ERR_ARG;
  for (int r = 0; r < n; r++) {
    if (!ctx[r] || !dblock[r] || !dall[r]) return PMCB200_ERR_ARG;
    pmcb200_ctx *c = ctx[r];
    CUDA_OK(c, cudaSetDevice(c->device));
    CUDA_OK(c, cudaEventRecord(c->ev_blk, c->stream));
  }
  const size_t bytes = (size_t)len * sizeof(double);
  for (int q = 0; q < n; q++) {
    pmcb200_ctx *c = ctx[q];
    CUDA_OK(c, cudaSetDevice(c->device));
    for (int r = 0; r < n; r++) {
      if (r != q) CUDA_OK(c, cudaStreamWaitEvent(c->stream, ctx[r]->ev_blk, 0));
      if (ctx[r]->device == c->device)
        CUDA_OK(c, cudaMemcpyAsync(dall[q] + (size_t)r * len, dblock[r], bytes, cudaMemcpyDeviceToDevice, c->stream));
      else   // direct over NVLink when peer access is possible, staged by the driver otherwise
        CUDA_OK(c, cudaMemcpyPeerAsync(dall[q] + (size_t)r * len, c->device, dblock[r], ctx[r]->device, bytes, c->stream));
    }
  }
  return 0;
}

extern "C" int pmcb200_normalize_with(pmcb200_ctx *c, int64_t N, const int16_t *dflg, double *dw,
                                      double maxW, double sum_shift) {
  int rc = need(c, false, false);
  if (rc) return rc;
  if (N < 0 || (N > 0 && (!dflg || !dw)) || !(sum_shift > 0.0)) return fail(c, PMCB200_ERR_ARG, "normalize_with: bad arguments");
  if (N == 0) return 0;
  pmc_launch_normalize(N, dflg, dw, maxW, 1.0 / sum_shift, c->stream);
  LAUNCH_OK(c);
  return 0;
}

extern "C" int pmcb200_iteration_host_multi(pmcb200_ctx *const *ctx, int n, int64_t N, uint64_t seed,
                                            uint32_t iter, double beta, double *hX, int32_t *hidx,
                                            int16_t *hflg, double *hw, pmcb200_stats_t *stats) {
  if (!ctx || n < 1 || n > 64 || !ctx[0]) return PMCB200_ERR_ARG;
  if (n == 1) return pmcb200_iteration_host(ctx[0], N, seed, iter, beta, hX, hidx, hflg, hw, stats);
  pmcb200_ctx *c0 = ctx[0];
  if (N < 1) return fail(c0, PMCB200_ERR_ARG, "iteration_host_multi: N = %lld", (long long)N);
  int rc;
  for (int r = 0; r < n; r++) {
    if (!ctx[r]) return fail(c0, PMCB200_ERR_ARG, "iteration_host_multi: null context %d", r);
    if ((rc = need(ctx[r], true, true))) { if (r) fail(c0, rc, "shard %d: %s", r, ctx[r]->errmsg); return rc; }
    if (ctx[r]->h.K != c0->h.K || ctx[r]->h.d != c0->h.d)
      return fail(c0, PMCB200_ERR_DIM, "iteration_host_multi: shard %d holds another proposal shape", r);
  }
  const int d = c0->h.d;
  const int64_t len = stat_len(c0->h.K, d), per = (N + n - 1) / n;
  double *blk[64], *all[64];
  // 1. queue every shard (sampler, likelihood, weights, local EM statistics, overlapped D2H)
  for (int r = 0; r < n; r++) {
    pmcb200_ctx *c = ctx[r];
    CUDA_OK(c, cudaSetDevice(c->device));
    if ((rc = ensure(c, c->sBlock, (size_t)len * sizeof(double))) || (rc = ensure(c, c->sAll, (size_t)len * n * sizeof(double)))) {
      if (r) fail(c0, rc, "shard %d: %s", r, c->errmsg);
      return rc;
    }
    blk[r] = (double *)c->sBlock.p; all[r] = (double *)c->sAll.p;
    const int64_t off = std::min<int64_t>(N, r * per), nr = std::max<int64_t>(0, std::min<int64_t>(per, N - off));
    rc = iteration_core(c, nr, seed, iter, off, beta, nullptr, nullptr, nullptr, nullptr, blk[r],
                        hX ? hX + (size_t)off * d : nullptr, hidx ? hidx + off : nullptr, hflg ? hflg + off : nullptr);
    if (rc) { if (r) fail(c0, rc, "shard %d: %s", r, c->errmsg); return rc; }
  }
  // 2. the one exchange: statistics blocks to every context
  if ((rc = pmcb200_allgather_blocks(ctx, n, blk, all, len))) return rc;
  // 3. identical fixed-order combine + M-step on every context
  pmcb200_stats_t st0;
  int rc_fin = 0;
  for (int r = 0; r < n; r++) {
    pmcb200_stats_t st;
    rc = pmcb200_em_finish(ctx[r], n, all[r], N, &st);
    if (r == 0) { st0 = st; rc_fin = rc; }
    else if (rc != rc_fin) return fail(c0, rc ? rc : PMCB200_ERR_STATE, "shard %d finished differently (%d vs %d): %s", r, rc, rc_fin, ctx[r]->errmsg);
  }
  if (stats) *stats = st0;
  if (rc_fin) return rc_fin;
  // 4. normalised weights of every shard to the host: queue all, then wait
  for (int r = 0; r < n && hw; r++) {
    pmcb200_ctx *c = ctx[r];
    const int64_t off = std::min<int64_t>(N, r * per), nr = std::max<int64_t>(0, std::min<int64_t>(per, N - off));
    if (nr == 0) continue;
    if ((rc = pmcb200_normalize_weights(c, nr, (int16_t *)c->set[c->cur].Flg.p, (double *)c->set[c->cur].Logw.p))) { if (r) fail(c0, rc, "shard %d: %s", r, c->errmsg); return rc; }
    CUDA_OK(c, cudaMemcpyAsync(hw + off, c->set[c->cur].Logw.p, (size_t)nr * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
  }
  for (int r = 0; r < n; r++) {
    pmcb200_ctx *c = ctx[r];
    CUDA_OK(c, cudaSetDevice(c->device));
    CUDA_OK(c, cudaStreamSynchronize(c->copy_stream));
    CUDA_OK(c, cudaStreamSynchronize(c->stream));
  }
  return 0;
}

// ---- entry points used by the pmclib-named host shims -------------------------------------
extern "C" int pmcb200_set_box(pmcb200_ctx *c, int d, const double *bmin, const double *bmax) {
  if (!c || !bmin || !bmax || d < 1 || d > PMCB200_MAX_DIM) return PMCB200_ERR_ARG;
  CUDA_OK(c, cudaSetDevice(c->device));
  std::vector<double> box(2 * d);
  for (int j = 0; j < d; j++) { box[j] = bmin[j]; box[d + j] = bmax[j]; }
  if (c->d_box && c->box_d == d) {      // reuse the buffer (called once per iteration by the host shims)
    CUDA_OK(c, cudaMemcpyAsync(c->d_box, box.data(), box.size() * sizeof(double), cudaMemcpyHostToDevice, c->stream));
    CUDA_OK(c, cudaStreamSynchronize(c->stream));
    return 0;
  }
  const double *dbox;
  int rc = dev_copy<double>(c, box.data(), box.size(), &dbox);
  if (rc) return rc;
  c->d_box = (double *)dbox;
  c->box_d = d;
  return 0;
}

extern "C" int pmcb200_read_counts(pmcb200_ctx *c, int64_t *nok_box, int64_t *nok, double *maxW) {
  if (!c) return PMCB200_ERR_ARG;
  DevScal h;
  CUDA_OK(c, cudaMemcpyAsync(&h, c->d_scal, sizeof(h), cudaMemcpyDeviceToHost, c->stream));
  CUDA_OK(c, cudaStreamSynchronize(c->stream));
  if (nok_box) *nok_box = (int64_t)h.nok_box;
  if (nok) *nok = (int64_t)h.nok;
  if (maxW) {
    unsigned long long k = h.max_key;
    if (k == 0ull) *maxW = -INFINITY;
    else {
      unsigned long long b = (k & 0x8000000000000000ull) ? (k & 0x7fffffffffffffffull) : ~k;
      memcpy(maxW, &b, 8);
    }
  }
  return 0;
}

extern "C" int pmcb200_weight_stats(pmcb200_ctx *c, int64_t N, const int16_t *dflg, const double *dw,
                                    int is_log, double out[8]) {
  int rc = need(c, false, false);
  if (rc) return rc;
  if (N < 1 || !dflg || !dw || !out) return fail(c, PMCB200_ERR_ARG, "weight_stats: bad arguments");
  const int blocks = (int)std::max<int64_t>(1, std::min<int64_t>(2 * c->sm_count, (N + PMC_BLOCK - 1) / PMC_BLOCK));
  if ((rc = ensure(c, c->sBlock, (size_t)std::max<int64_t>(stat_len(PMCB200_MAX_COMP, 1), 16 + 9 * blocks) * sizeof(double)))) return rc;
  double *buf = (double *)c->sBlock.p;      // [0..8) result, then max partials, then sum partials
  pmc_launch_wstat(N, dflg, dw, is_log, blocks, buf + 8, buf + 8 + blocks, buf, c->stream);
  c->launches += is_log ? 3 : 2;
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return fail(c, PMCB200_ERR_CUDA, "weight_stats launch: %s", cudaGetErrorString(e));
  CUDA_OK(c, cudaMemcpyAsync(out, buf, 8 * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
  CUDA_OK(c, cudaStreamSynchronize(c->stream));
  return 0;
}

extern "C" int pmcb200_normalize_log_weights(pmcb200_ctx *c, int64_t N, const int16_t *dflg, double *dw,
                                             double *sum_shift, double *logSum, double *maxW) {
  double o[8];
  int rc = pmcb200_weight_stats(c, N, dflg, dw, 1, o);
  if (rc) return rc;
  if (!(o[1] > 0.0)) return fail(c, PMCB200_ERR_NOSAMPLE, "normalize: no sample with finite weight");
  pmc_launch_normalize(N, dflg, dw, o[0], 1.0 / o[1], c->stream);
  LAUNCH_OK(c);
  if (sum_shift) *sum_shift = o[1];
  if (logSum) *logSum = std::log(o[1]) + o[0];
  if (maxW) *maxW = o[0];
  return 0;
}

extern "C" int pmcb200_em_local_linear(pmcb200_ctx *c, int64_t N, const double *dX, const int32_t *didx,
                                       const int16_t *dflg, const double *dwbar, double *dblock) {
  int rc = need(c, true, false);
  if (rc) return rc;
  if (N < 0 || !dblock || (N > 0 && (!dX || !didx || !dflg || !dwbar)))
    return fail(c, PMCB200_ERR_ARG, "em_local_linear: bad arguments");
  return launch_em_local(c, N, dX, didx, dflg, dwbar, dblock, 1);
}

// ---- weighted post-processing of a stored sample -----------------------------
extern "C" int pmcb200_post_moments(pmcb200_ctx *c, int64_t N, int d, const double *dX, const int16_t *dflg,
                                    const double *dw, double *mean, double *cov) {
  int rc = need(c, false, false);
  if (rc) return rc;
  if (N < 1 || d < 1 || d > PMCB200_MAX_DIM || !dX || !mean) return fail(c, PMCB200_ERR_ARG, "post_moments: bad arguments");
  const int M = 1 + d + d * (d + 1) / 2;
  // d <= 8: register kernel, enough resident threads to cover the HBM latency; larger d: tile kernel
  const int blocks = (int)std::max<int64_t>(1, std::min<int64_t>((d <= 8 ? 8 : 2) * c->sm_count, (N + 255) / 256));
  if ((rc = ensure(c, c->sPost, sizeof(double) * ((size_t)blocks * M + 2 * M + d)))) return rc;
  double *part = (double *)c->sPost.p, *out = part + (size_t)blocks * M, *piv = out + 2 * M;
  std::vector<double> h(M);
  // pass 1 (pivot 0): S0 and the mean
  pmc_launch_post_moments(N, d, dX, dflg, dw, nullptr, blocks, part, out, c->stream);
  c->launches += 2;
  CUDA_OK(c, cudaMemcpyAsync(h.data(), out, sizeof(double) * M, cudaMemcpyDeviceToHost, c->stream));
  CUDA_OK(c, cudaStreamSynchronize(c->stream));
  if (!(h[0] > 0.0)) return fail(c, PMCB200_ERR_NOSAMPLE, "post_moments: sum of weights is not positive");
  for (int j = 0; j < d; j++) mean[j] = h[1 + j] / h[0];
  if (!cov) return 0;
  // pass 2: second moments about the mean (as estimate_param_covar_weight does)
  CUDA_OK(c, cudaMemcpyAsync(piv, mean, sizeof(double) * d, cudaMemcpyHostToDevice, c->stream));
  pmc_launch_post_moments(N, d, dX, dflg, dw, piv, blocks, part, out, c->stream);
  c->launches += 2;
  CUDA_OK(c, cudaMemcpyAsync(h.data(), out, sizeof(double) * M, cudaMemcpyDeviceToHost, c->stream));
  CUDA_OK(c, cudaStreamSynchronize(c->stream));
  for (int a = 0, t = 1 + d; a < d; a++)
    for (int b = 0; b <= a; b++, t++) {
      const double da = h[1 + a] / h[0], db = h[1 + b] / h[0];     // residual of the mean: O(1e-16)
      cov[a * d + b] = cov[b * d + a] = h[t] / h[0] - da * db;
    }
  return 0;
}

extern "C" int pmcb200_post_sigma(pmcb200_ctx *c, int64_t N, int d, const double *dX, const int16_t *dflg,
                                  const double *dw, int a, double center, const double conf[3], double sigma[6],
                                  double *median, int64_t *nflagged) {
  int rc = need(c, false, false);
  if (rc) return rc;
  if (N < 1 || d < 1 || a < 0 || a >= d || !dX || !conf || !sigma) return fail(c, PMCB200_ERR_ARG, "post_sigma: bad arguments");
  const size_t tb = pmc_post_sigma_temp_bytes(N);
  if ((rc = ensure(c, c->sPost, sizeof(double) * (4 * (size_t)N + 16)))) return rc;
  if ((rc = ensure(c, c->sPostTmp, tb))) return rc;
  double *work = (double *)c->sPost.p, *out8 = work + 4 * (size_t)N;
  unsigned long long *nf = (unsigned long long *)(out8 + 8);
  pmc_launch_post_sigma(N, d, a, dX, dflg, dw, center, conf, work, c->sPostTmp.p, tb, nf, out8, c->stream);
  c->launches += 2;          // own kernels: gather + search (the sort and the scan are CUB's)
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return fail(c, PMCB200_ERR_CUDA, "post_sigma launch: %s", cudaGetErrorString(e));
  double h[8];
  CUDA_OK(c, cudaMemcpyAsync(h, out8, sizeof(h), cudaMemcpyDeviceToHost, c->stream));
  CUDA_OK(c, cudaStreamSynchronize(c->stream));
  for (int j = 0; j < 6; j++) sigma[j] = h[j];
  if (median) *median = h[6];
  if (nflagged) *nflagged = (int64_t)h[7];
  return 0;
}

extern "C" int pmcb200_post_histogram(pmcb200_ctx *c, int64_t N, int d, const double *dX, const int16_t *dflg,
                                      const double *dw, int nhdim, const int *pidx, const int *nbins,
                                      const double *limits, double *count, double *sumw, double *sumw2) {
  int rc = need(c, false, false);
  if (rc) return rc;
  if (N < 1 || d < 1 || !dX || (nhdim != 1 && nhdim != 2) || !pidx || !nbins || !limits || !count || !sumw || !sumw2)
    return fail(c, PMCB200_ERR_ARG, "post_histogram: bad arguments");
  size_t tdim = 1;
  for (int i = 0; i < nhdim; i++) {
    if (pidx[i] < 0 || pidx[i] >= d || nbins[i] < 1 || !(limits[2 * i + 1] > limits[2 * i]))
      return fail(c, PMCB200_ERR_ARG, "post_histogram: bad axis %d", i);
    tdim *= (size_t)nbins[i];
  }
  if (tdim * 3 * sizeof(double) > 200 * 1024) return fail(c, PMCB200_ERR_UNSUP, "post_histogram: %zu bins exceed the shared-memory copy (max 8533)", tdim);
  if ((rc = ensure(c, c->sPost, sizeof(double) * 3 * tdim))) return rc;
  const int blocks = (int)std::max<int64_t>(1, std::min<int64_t>((tdim <= 1024 ? 8 : 2) * c->sm_count, (N + 255) / 256));
  if (pmc_launch_post_hist(N, d, dX, dflg, dw, nhdim, pidx, nbins, limits, blocks, (double *)c->sPost.p, c->stream))
    return fail(c, PMCB200_ERR_UNSUP, "post_histogram: too many bins");
  LAUNCH_OK(c);
  std::vector<double> h(3 * tdim);
  CUDA_OK(c, cudaMemcpyAsync(h.data(), c->sPost.p, sizeof(double) * 3 * tdim, cudaMemcpyDeviceToHost, c->stream));
  CUDA_OK(c, cudaStreamSynchronize(c->stream));
  memcpy(count, h.data(), sizeof(double) * tdim);
  memcpy(sumw, h.data() + tdim, sizeof(double) * tdim);
  memcpy(sumw2, h.data() + 2 * tdim, sizeof(double) * tdim);
  return 0;
}

// ---- measurement helpers -------------------------------------------------------------------
extern "C" int pmcb200_counters(pmcb200_ctx *c, int64_t out[4]) {
  if (!c || !out) return PMCB200_ERR_ARG;
  DevCount h;
  CUDA_OK(c, cudaMemcpyAsync(&h, c->d_cnt, sizeof(h), cudaMemcpyDeviceToHost, c->stream));
  CUDA_OK(c, cudaMemsetAsync(c->d_cnt, 0, sizeof(DevCount), c->stream));
  CUDA_OK(c, cudaStreamSynchronize(c->stream));
  out[0] = (int64_t)h.sn_evals; out[1] = (int64_t)h.sn_zsteps; out[2] = (int64_t)h.gen_evals; out[3] = (int64_t)h.gen_integrals;
  return 0;
}

extern "C" int pmcb200_counters_ex(pmcb200_ctx *c, int64_t *out, int n) {
  if (!c || !out || n < 1) return PMCB200_ERR_ARG;
  DevCount h;
  CUDA_OK(c, cudaMemcpyAsync(&h, c->d_cnt, sizeof(h), cudaMemcpyDeviceToHost, c->stream));
  CUDA_OK(c, cudaMemsetAsync(c->d_cnt, 0, sizeof(DevCount), c->stream));
  CUDA_OK(c, cudaStreamSynchronize(c->stream));
  const int64_t v[7] = {(int64_t)h.sn_evals, (int64_t)h.sn_zsteps, (int64_t)h.gen_evals, (int64_t)h.gen_integrals,
                        (int64_t)h.sn_spec, (int64_t)h.sn_exact, (int64_t)h.cmb_spec};
  for (int i = 0; i < n; i++) out[i] = i < 7 ? v[i] : 0;
  return 0;
}

extern "C" int pmcb200_fp64_peak(pmcb200_ctx *c, double *tflops) {
  if (!c || !tflops) return PMCB200_ERR_ARG;
  CUDA_OK(c, cudaSetDevice(c->device));
  double *d = nullptr;
  CUDA_OK(c, cudaMalloc((void **)&d, 8 + 64));
  {
    double h[8];
    for (int i = 0; i < 8; i++) h[i] = 0.999999 + 1e-9 * i;
    CUDA_OK(c, cudaMemcpy(d + 1, h, 64, cudaMemcpyHostToDevice));
  }
  cudaEvent_t e0, e1;
  CUDA_OK(c, cudaEventCreate(&e0));
  CUDA_OK(c, cudaEventCreate(&e1));
  const int blocks = c->sm_count * 8, iters = 2048;
  double best = 0.0;
  for (int rep = 0; rep < 6; rep++) {
    CUDA_OK(c, cudaEventRecord(e0, c->stream));
    pmc_launch_fp64_peak(d, d + 1, blocks, iters, c->stream);
    CUDA_OK(c, cudaEventRecord(e1, c->stream));
    CUDA_OK(c, cudaEventSynchronize(e1));
    float ms = 0.f;
    CUDA_OK(c, cudaEventElapsedTime(&ms, e0, e1));
    double fl = 2.0 * 8 * 16 * (double)iters * 256.0 * blocks;
    if (rep > 0) best = std::max(best, fl / (ms * 1e-3) * 1e-12);
  }
  cudaEventDestroy(e0); cudaEventDestroy(e1); cudaFree(d);
  *tflops = best;
  return 0;
}
