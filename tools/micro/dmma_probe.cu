// FP64 tensor-core (DMMA m8n8k4) probe for sm_100a: fragment layout check, throughput, and how
// much a second instruction stream (integer ALU, shared-memory broadcast loads, indexed constant
// loads) costs next to a saturated DFMA stream.
// nvcc -O3 -gencode arch=compute_100a,code=sm_100a dmma_probe.cu -o dmma_probe
#include <cstdio>
#include <cmath>
#include <cuda_runtime.h>

__device__ __forceinline__ void dmma(double &c0, double &c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
               : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

// C[8][8] = A[8][4] B[4][8] with the fragment layout of the PTX ISA:
// a: A[lane/4][lane%4]; b: B[lane%4][lane/4]; c0,c1: C[lane/4][2 (lane%4) + {0,1}]
__global__ void k_layout(const double *A, const double *B, double *C) {
  const int lane = threadIdx.x, g = lane >> 2, t = lane & 3;
  double c0 = 0.0, c1 = 0.0;
  dmma(c0, c1, A[g * 4 + t], B[t * 8 + g]);
  C[g * 8 + 2 * t] = c0;
  C[g * 8 + 2 * t + 1] = c1;
}

template <int NACC>
__global__ void __launch_bounds__(256) k_dmma(double *out, const double *in, int iters) {
  double a = in[threadIdx.x & 7], b = in[8 + (threadIdx.x & 7)];
  double c[NACC][2];
#pragma unroll
  for (int i = 0; i < NACC; i++) { c[i][0] = i; c[i][1] = -i; }
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int i = 0; i < NACC; i++) dmma(c[i][0], c[i][1], a, b);
  }
  double s = 0.0;
#pragma unroll
  for (int i = 0; i < NACC; i++) s += c[i][0] + c[i][1];
  if (s == 123.456) out[0] = s;
}

// DFMA and DMMA interleaved (NF DFMAs per DMMA): do the vector FP64 lanes and the FP64 tensor path share one
// datapath (combined rate <= 100 % of the vector peak) or run side by side?
template <int NF>
__global__ void __launch_bounds__(256) k_both(double *out, const double *in, int iters) {
  double a = in[threadIdx.x & 7], b = in[8 + (threadIdx.x & 7)], y = in[threadIdx.x & 3];
  double c[4][2], x[8];
#pragma unroll
  for (int i = 0; i < 4; i++) { c[i][0] = i; c[i][1] = -i; }
#pragma unroll
  for (int i = 0; i < 8; i++) x[i] = threadIdx.x + i;
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int i = 0; i < 4; i++) {
      dmma(c[i][0], c[i][1], a, b);
#pragma unroll
      for (int j = 0; j < NF; j++) x[(i * NF + j) & 7] = fma(x[(i * NF + j) & 7], y, 1e-9);
    }
  }
  double s = 0.0;
#pragma unroll
  for (int i = 0; i < 4; i++) s += c[i][0] + c[i][1];
#pragma unroll
  for (int i = 0; i < 8; i++) s += x[i];
  if (s == 123.456) out[0] = s;
}

__constant__ double CT[1024];
// MODE 0: 8 DFMA; 1: 8 DFMA + 8 IMAD; 2: 8 DFMA + 8 broadcast LDS feeding them; 3: 8 DFMA + 8 uniform-indexed
// constant loads feeding them; 4: 8 DFMA + 4 IMAD; 5: 8 DFMA fed by 8 broadcast LDG (L1-resident)
template <int MODE>
__global__ void __launch_bounds__(256) k_mix(double *out, const double *in, int iters, int kvar) {
  __shared__ double sh[1024];
  for (int i = threadIdx.x; i < 1024; i += 256) sh[i] = in[i & 15];
  __syncthreads();
  double y = in[threadIdx.x & 7];
  double x0 = threadIdx.x, x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3, x4 = x0 + 4, x5 = x0 + 5, x6 = x0 + 6, x7 = x0 + 7;
  int j0 = threadIdx.x, j1 = 1, j2 = 2, j3 = 3, j4 = 4, j5 = 5, j6 = 6, j7 = 7;
  for (int i = 0; i < iters; i++) {
    const int base = ((i * kvar) & 63) * 8;     // warp-uniform, varies per iteration
#pragma unroll
    for (int u = 0; u < 8; u++) {
      if (MODE == 0 || MODE == 1 || MODE == 4) {
        x0 = fma(x0, y, 1e-9); x1 = fma(x1, y, 1e-9); x2 = fma(x2, y, 1e-9); x3 = fma(x3, y, 1e-9);
        x4 = fma(x4, y, 1e-9); x5 = fma(x5, y, 1e-9); x6 = fma(x6, y, 1e-9); x7 = fma(x7, y, 1e-9);
      }
      if (MODE == 1) {
        j0 = j0 * 3 + j1; j1 = j1 * 5 + j2; j2 = j2 * 7 + j3; j3 = j3 * 9 + j4;
        j4 = j4 * 11 + j5; j5 = j5 * 13 + j6; j6 = j6 * 15 + j7; j7 = j7 * 17 + j0;
      }
      if (MODE == 4) { j0 = j0 * 3 + j1; j1 = j1 * 5 + j2; j2 = j2 * 7 + j3; j3 = j3 * 9 + j0; }
      if (MODE == 2) {
        const double *s = sh + base + 8 * u * 0;
        x0 = fma(x0, s[u * 8 + 0], 1e-9); x1 = fma(x1, s[u * 8 + 1], 1e-9); x2 = fma(x2, s[u * 8 + 2], 1e-9);
        x3 = fma(x3, s[u * 8 + 3], 1e-9); x4 = fma(x4, s[u * 8 + 4], 1e-9); x5 = fma(x5, s[u * 8 + 5], 1e-9);
        x6 = fma(x6, s[u * 8 + 6], 1e-9); x7 = fma(x7, s[u * 8 + 7], 1e-9);
      }
      if (MODE == 3) {
        const double *s = CT + base;
        x0 = fma(x0, s[u * 8 + 0], 1e-9); x1 = fma(x1, s[u * 8 + 1], 1e-9); x2 = fma(x2, s[u * 8 + 2], 1e-9);
        x3 = fma(x3, s[u * 8 + 3], 1e-9); x4 = fma(x4, s[u * 8 + 4], 1e-9); x5 = fma(x5, s[u * 8 + 5], 1e-9);
        x6 = fma(x6, s[u * 8 + 6], 1e-9); x7 = fma(x7, s[u * 8 + 7], 1e-9);
      }
      if (MODE == 5) {
        const double *s = in + 16 + base;
        x0 = fma(x0, __ldg(s + u * 8 + 0), 1e-9); x1 = fma(x1, __ldg(s + u * 8 + 1), 1e-9);
        x2 = fma(x2, __ldg(s + u * 8 + 2), 1e-9); x3 = fma(x3, __ldg(s + u * 8 + 3), 1e-9);
        x4 = fma(x4, __ldg(s + u * 8 + 4), 1e-9); x5 = fma(x5, __ldg(s + u * 8 + 5), 1e-9);
        x6 = fma(x6, __ldg(s + u * 8 + 6), 1e-9); x7 = fma(x7, __ldg(s + u * 8 + 7), 1e-9);
      }
    }
  }
  double s = x0 + x1 + x2 + x3 + x4 + x5 + x6 + x7 + (double)(j0 + j1 + j2 + j3 + j4 + j5 + j6 + j7);
  if (s == 123.456) out[0] = s;
}

template <class K, class... A> float best_ms(K kern, int blocks, A... args) {
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  float best = 1e30f;
  for (int r = 0; r < 5; r++) {
    cudaEventRecord(e0); kern<<<blocks, 256>>>(args...); cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1); if (r && ms < best) best = ms;
  }
  return best;
}

int main() {
  double *d, *in; cudaMalloc(&d, 4096); cudaMalloc(&in, 8192 + 128);
  double h[1040]; for (int i = 0; i < 1040; i++) h[i] = 0.999 + 1e-6 * (i & 15);
  cudaMemcpy(in, h, sizeof h, cudaMemcpyHostToDevice);
  cudaMemcpyToSymbol(CT, h, 8192);
  {  // layout
    double A[32], B[32], C[64], R[64], *dA, *dB, *dC;
    for (int i = 0; i < 32; i++) { A[i] = 1.0 + 0.37 * i; B[i] = -2.0 + 0.11 * i * i; }
    for (int i = 0; i < 8; i++) for (int j = 0; j < 8; j++) { double s = 0; for (int k = 0; k < 4; k++) s += A[i * 4 + k] * B[k * 8 + j]; R[i * 8 + j] = s; }
    cudaMalloc(&dA, 256); cudaMalloc(&dB, 256); cudaMalloc(&dC, 512);
    cudaMemcpy(dA, A, 256, cudaMemcpyHostToDevice); cudaMemcpy(dB, B, 256, cudaMemcpyHostToDevice);
    k_layout<<<1, 32>>>(dA, dB, dC);
    cudaMemcpy(C, dC, 512, cudaMemcpyDeviceToHost);
    double e = 0; for (int i = 0; i < 64; i++) e = fmax(e, fabs(C[i] - R[i]) / fabs(R[i]));
    printf("dmma m8n8k4 layout check: max rel err %.3g (%s)\n", e, e < 1e-14 ? "OK" : "MISMATCH");
  }
  const double peak = 148.0 * 64 * 1.965e9;   // FP64 FMA lanes per second
  for (int b : {2, 4, 8}) {
    const int blocks = 148 * b, iters = 4096;
    float ms = best_ms(k_dmma<8>, blocks, d, in, iters);
    double fma = 256.0 * 8 * iters * 8.0 * blocks;   // 256 FMA per warp-level DMMA, 8 warps per block
    printf("DMMA m8n8k4 x8 acc   blocks/SM %d: %.2f TFLOP/s (%.1f%% of vector FP64 peak)\n", b, 2 * fma / ms * 1e-9, 100 * fma / (ms * 1e-3) / peak);
    ms = best_ms(k_dmma<4>, blocks, d, in, iters);
    fma = 256.0 * 4 * iters * 8.0 * blocks;
    printf("DMMA m8n8k4 x4 acc   blocks/SM %d: %.2f TFLOP/s (%.1f%%)\n", b, 2 * fma / ms * 1e-9, 100 * fma / (ms * 1e-3) / peak);
  }
  for (int b : {2, 4}) {     // 1 DMMA = 8 warp-DFMAs of work: NF = 8 asks for equal shares
    const int blocks = 148 * b, iters = 4096;
    float ms = best_ms(k_both<8>, blocks, d, in, iters);
    double fma = (256.0 + 32.0 * 8) * 4 * iters * 8.0 * blocks;
    printf("4 x (1 DMMA + 8 DFMA)  blocks/SM %d: %.3f ms, combined %.1f%% of the vector FP64 peak\n", b, ms, 100 * fma / (ms * 1e-3) / peak);
    ms = best_ms(k_both<2>, blocks, d, in, iters);
    fma = (256.0 + 32.0 * 2) * 4 * iters * 8.0 * blocks;
    printf("4 x (1 DMMA + 2 DFMA)  blocks/SM %d: %.3f ms, combined %.1f%% of the vector FP64 peak\n", b, ms, 100 * fma / (ms * 1e-3) / peak);
  }
  const char *names[6] = {"8 DFMA", "8 DFMA + 8 IMAD", "8 DFMA + 8 LDS.64 bcast", "8 DFMA + 8 const idx", "8 DFMA + 4 IMAD", "8 DFMA + 8 LDG bcast"};
  for (int b : {2, 4}) {
    const int blocks = 148 * b, iters = 2048;
    float ms[6];
    ms[0] = best_ms(k_mix<0>, blocks, d, in, iters, 3); ms[1] = best_ms(k_mix<1>, blocks, d, in, iters, 3);
    ms[2] = best_ms(k_mix<2>, blocks, d, in, iters, 3); ms[3] = best_ms(k_mix<3>, blocks, d, in, iters, 3);
    ms[4] = best_ms(k_mix<4>, blocks, d, in, iters, 3); ms[5] = best_ms(k_mix<5>, blocks, d, in, iters, 3);
    for (int m = 0; m < 6; m++) {
      double ops = 8.0 * 8 * iters * 256.0 * blocks;
      printf("%-26s blocks/SM %d: %.3f ms, DFMA at %.1f%% of peak\n", names[m], b, ms[m], 100 * ops / (ms[m] * 1e-3) / peak);
    }
  }
  cudaError_t e = cudaDeviceSynchronize();
  printf("status: %s\n", cudaGetErrorString(e));
  return 0;
}
