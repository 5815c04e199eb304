// Host-side check of the pure-arithmetic device helpers (compiled by nvcc, runs on the CPU: the helpers are
// __host__ __device__ and use only IEEE add / mul / fma, so the host results equal the device's bit for bit).
//  1. comp_maha_cols<D, S> (column-oriented, S samples per thread: k_weights_multi, k_em_stats_mma phase 1) is
//     BIT-IDENTICAL to comp_maha<D> (row-oriented forward substitution: k_weights, k_logq, k_em_stats), and both
//     agree with a long-double forward substitution; so is comp_maha_cols_cp on the column-packed copy built by
//     cp_fill_component (experimental variant).
//  2. the packed-mixture layout helpers (mix_stride, pmc_pad_dim) and the EM feature enumeration used by the
//     tensor-core kernel (f = 0 -> G, 1..d -> B, then the lower triangle by rows) cover the statistics block once.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cmath>
#include <vector>
#include "../../cosmopmc_b200/csrc/common.cuh"
#include "../../cosmopmc_b200/csrc/stat_layout.cuh"

static double urand() { return (double)rand() / RAND_MAX; }

template <int D, int S>
static int check_maha(int d) {
  const int stride = mix_stride(d);
  if (pmc_pad_dim(d) != D) { printf("pad_dim(%d) != %d\n", d, D); return 1; }
  std::vector<double> comp(stride, 0.0);
  double *mean = comp.data() + 2, *L = comp.data() + 2 + D, *rd = comp.data() + 2 + D + D * (D + 1) / 2;
  for (int i = 0; i < D; i++) {
    mean[i] = (i < d) ? 2.0 * urand() - 1.0 : 0.0;
    for (int k = 0; k <= i; k++)
      L[i * (i + 1) / 2 + k] = (i < d && k < d) ? (k == i ? 0.5 + urand() : 0.6 * (urand() - 0.5)) : (k == i ? 1.0 : 0.0);
    rd[i] = 1.0 / L[i * (i + 1) / 2 + i];
  }
  int bad = 0;
  for (int trial = 0; trial < 200; trial++) {
    double x[S][D], t[S][D], m_cols[S], y[D];
    for (int s = 0; s < S; s++)
      for (int i = 0; i < D; i++) { x[s][i] = (i < d) ? 6.0 * urand() - 3.0 : 0.0; t[s][i] = x[s][i]; }
    comp_maha_cols<D, S>(comp.data(), t, m_cols);
    // column-packed copy (experimental k_weights_multi<.., CP = true>): same operations, same order
    alignas(16) double cp[CpLayout<D>::stride + 2];
    for (int e = 0; e < CpLayout<D>::stride; e++) cp[e] = 0.0;
    for (int lane = 0; lane < 7; lane++) cp_fill_component<D>(comp.data(), cp, lane, 7);
    double t2[S][D], m_cp[S];
    for (int s = 0; s < S; s++) for (int i = 0; i < D; i++) t2[s][i] = x[s][i];
    comp_maha_cols_cp<D, S>(cp, t2, m_cp);
    if (cp[0] != comp[0] || cp[1] != comp[1]) { printf("D=%d: cp header\n", D); bad++; }
    for (int s = 0; s < S; s++) {
      const double m_row = comp_maha<D>(comp.data(), d, x[s], y);
      if (memcmp(&m_row, &m_cols[s], 8) != 0) { printf("D=%d S=%d: row %.17g != cols %.17g\n", D, S, m_row, m_cols[s]); bad++; }
      if (memcmp(&m_row, &m_cp[s], 8) != 0) { printf("D=%d S=%d: row %.17g != column-packed %.17g\n", D, S, m_row, m_cp[s]); bad++; }
      long double yl[D], ml = 0.0L;       // reference
      for (int i = 0; i < D; i++) {
        long double tt = (long double)x[s][i] - mean[i];
        for (int k = 0; k < i; k++) tt -= (long double)L[i * (i + 1) / 2 + k] * yl[k];
        yl[i] = tt / L[i * (i + 1) / 2 + i];
        ml += yl[i] * yl[i];
      }
      if (fabsl(ml - m_row) > 1e-12L * fabsl(ml)) { printf("D=%d: m %.17g vs long double %.17Lg\n", D, m_row, ml); bad++; }
    }
  }
  return bad;
}

static int check_features(int d) {
  // enumeration of k_em_stats_mma: feature f -> position in a component's statistics block
  const int nfeat = 1 + d + mix_tri(d), M = stat_cs(d);
  std::vector<int> hit(M, 0);
  for (int f = 0; f < nfeat; f++) {
    int a = -1, b = -1;      // staged-row columns (D = ones column)
    if (f >= 1 && f < 1 + d) a = f - 1;
    else if (f >= 1 + d) {
      const int qq = f - 1 - d;
      int i = 0;
      while ((i + 1) * (i + 2) / 2 <= qq) i++;
      a = i; b = qq - i * (i + 1) / 2;
      if (b > a || a >= d) return 1;
      if (3 + d + a * (a + 1) / 2 + b != f + 2) return 1;   // C[tri] sits behind A, G, count, B[d]
    }
    const int fo = (f == 0) ? 1 : f + 2;
    if (fo >= M) return 1;
    hit[fo]++;
  }
  hit[0]++; hit[2]++;        // A (= G for a Gaussian proposal) and the draw count are written separately
  for (int i = 0; i < M; i++) if (hit[i] != 1) return 1;
  return 0;
}

int main() {
  srand(12345);
  int bad = 0;
  bad += check_maha<5, 1>(5) + check_maha<5, 4>(5) + check_maha<8, 4>(8) + check_maha<12, 2>(11) +
         check_maha<20, 1>(20) + check_maha<20, 2>(18) + check_maha<20, 3>(20) + check_maha<32, 2>(32) +
         check_maha<2, 4>(1);
  for (int d = 1; d <= 32; d++) bad += check_features(d);
  if (bad) { printf("FAILED: %d\n", bad); return 1; }
  printf("device math host check ok\n");
  return 0;
}
