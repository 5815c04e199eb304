// stat_layout.cuh -- layout of the EM sufficient-statistics block and of the
// M-step result buffer (shared by host API and kernels).
#pragma once
#include "common.cuh"

#define STAT_HDR 8
__host__ __device__ inline int stat_cs(int d) { return 3 + d + mix_tri(d); }
__host__ __device__ inline int64_t stat_len(int K, int d) { return STAT_HDR + (int64_t)K * stat_cs(d); }

#define RES_HDR 16       // M-step result: [0..16) stats, then wght[K], mean[K*d], chol[K*d*d]
