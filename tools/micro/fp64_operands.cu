// FP64 pipe microbenchmark: DFMA throughput vs number of distinct register operands.
// nvcc -O3 -gencode arch=compute_100a,code=sm_100a fp64_operands.cu -o fp64_operands
#include <cstdio>
#include <cuda_runtime.h>
template <int MODE>
__global__ void __launch_bounds__(256) k(double *out, const double *in, int iters) {
  double y = in[threadIdx.x & 7], z = in[8 + (threadIdx.x & 7)];
  double x0 = threadIdx.x, x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3, x4 = x0 + 4, x5 = x0 + 5, x6 = x0 + 6, x7 = x0 + 7;
  for (int i = 0; i < iters; i++) {
#pragma unroll
    for (int u = 0; u < 16; u++) {
      if (MODE == 0) { x0 = fma(x0, 0.999, 1e-9); x1 = fma(x1, 0.999, 1e-9); x2 = fma(x2, 0.999, 1e-9); x3 = fma(x3, 0.999, 1e-9); x4 = fma(x4, 0.999, 1e-9); x5 = fma(x5, 0.999, 1e-9); x6 = fma(x6, 0.999, 1e-9); x7 = fma(x7, 0.999, 1e-9); }
      if (MODE == 1) { x0 = fma(x0, y, 1e-9); x1 = fma(x1, y, 1e-9); x2 = fma(x2, y, 1e-9); x3 = fma(x3, y, 1e-9); x4 = fma(x4, y, 1e-9); x5 = fma(x5, y, 1e-9); x6 = fma(x6, y, 1e-9); x7 = fma(x7, y, 1e-9); }
      if (MODE == 2) { x0 = fma(x0, y, z); x1 = fma(x1, y, z); x2 = fma(x2, y, z); x3 = fma(x3, y, z); x4 = fma(x4, y, z); x5 = fma(x5, y, z); x6 = fma(x6, y, z); x7 = fma(x7, y, z); }
      if (MODE == 3) { x0 = fma(x0, x1, x2); x1 = fma(x1, x2, x3); x2 = fma(x2, x3, x4); x3 = fma(x3, x4, x5); x4 = fma(x4, x5, x6); x5 = fma(x5, x6, x7); x6 = fma(x6, x7, x0); x7 = fma(x7, x0, x1); }
      if (MODE == 4) { x0 += y; x1 += y; x2 += y; x3 += y; x4 += y; x5 += y; x6 += y; x7 += y; }
      if (MODE == 5) { x0 *= y; x1 *= y; x2 *= y; x3 *= y; x4 *= y; x5 *= y; x6 *= y; x7 *= y; }
    }
  }
  double s = x0 + x1 + x2 + x3 + x4 + x5 + x6 + x7;
  if (s == 123.456) out[0] = s;
}
template <int MODE> void run(const char *name, int blocks_per_sm) {
  double *d, *in; cudaMalloc(&d, 8); cudaMalloc(&in, 128);
  double h[16]; for (int i = 0; i < 16; i++) h[i] = 0.999 + 1e-6 * i; cudaMemcpy(in, h, 128, cudaMemcpyHostToDevice);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  int blocks = 148 * blocks_per_sm, iters = 2048; float best = 1e30f;
  for (int r = 0; r < 5; r++) { cudaEventRecord(e0); k<MODE><<<blocks, 256>>>(d, in, iters); cudaEventRecord(e1); cudaEventSynchronize(e1); float ms; cudaEventElapsedTime(&ms, e0, e1); if (r && ms < best) best = ms; }
  double ops = 8.0 * 16 * iters * 256.0 * blocks;
  printf("%-28s blocks/SM %d: %.2f Tinstr-lanes/s  (%.1f%% of 148*64*1.965e9)\n", name, blocks_per_sm, ops / best * 1e-9, 100 * ops / (best * 1e-3) / (148 * 64 * 1.965e9));
}
int main() {
  for (int b : {2, 4, 8}) {
    run<0>("DFMA 1 reg + 2 imm", b); run<1>("DFMA 2 regs + imm", b); run<2>("DFMA 3 regs (2 invariant)", b);
    run<3>("DFMA 3 regs (all varying)", b); run<4>("DADD 2 regs", b); run<5>("DMUL 2 regs", b);
  }
  return 0;
}
