// rsqrt_err.cu -- accuracy of the MUFU.RSQ64H seed (PTX rsqrt.approx.ftz.f64) and of the one- and
// two-term corrections used by the SN integrand (cosmopmc_b200/csrc/cosmo.cuh, sn_f / fast_rsqrt).
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o rsqrt_err rsqrt_err.cu && ./rsqrt_err
// Prints max |e| with e = 1 - x y0^2, and the max relative error of y0 (1 + e/2) [one Newton step]
// and of y0 (1 + e/2 + 3 e^2/8) [third order] against 1/sqrt(x) evaluated in IEEE double.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ unsigned long long g_emax, g_err1, g_err3;

__device__ __forceinline__ uint64_t mix64(uint64_t z) {
  z += 0x9e3779b97f4a7c15ull;
  z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ull;
  z = (z ^ (z >> 27)) * 0x94d049bb133111ebull;
  return z ^ (z >> 31);
}

__global__ void k(uint64_t n_per_thread, int wide) {
  uint64_t tid = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
  double emax = 0.0, e1 = 0.0, e3 = 0.0;
  for (uint64_t i = 0; i < n_per_thread; i++) {
    uint64_t r = mix64(tid * n_per_thread + i);
    // mantissa random; exponent in [1,4) or (wide) in 2^[-200, 200]
    int ex = wide ? (int)((r >> 52) % 401) - 200 : (int)((r >> 52) & 1);
    double x = __longlong_as_double(((uint64_t)(1023 + ex) << 52) | (r & 0x000fffffffffffffull));
    double y;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
    double t = x * y, e = fma(-t, y, 1.0);
    double h = __hiloint2double(__double2hiint(y) - 0x00100000, __double2loint(y));
    double y1 = fma(h, e, y);
    double y3 = fma(y * e, fma(0.375, e, 0.5), y);
    double ref = 1.0 / sqrt(x);
    emax = fmax(emax, fabs(e));
    e1 = fmax(e1, fabs(y1 - ref) / ref);
    e3 = fmax(e3, fabs(y3 - ref) / ref);
  }
  atomicMax(&g_emax, (unsigned long long)__double_as_longlong(emax));
  atomicMax(&g_err1, (unsigned long long)__double_as_longlong(e1));
  atomicMax(&g_err3, (unsigned long long)__double_as_longlong(e3));
}

int main() {
  for (int wide = 0; wide < 2; wide++) {
    unsigned long long z = 0;
    cudaMemcpyToSymbol(g_emax, &z, 8); cudaMemcpyToSymbol(g_err1, &z, 8); cudaMemcpyToSymbol(g_err3, &z, 8);
    k<<<148 * 8, 256>>>(1 << 14, wide);
    if (cudaDeviceSynchronize() != cudaSuccess) { printf("kernel failed\n"); return 1; }
    double a, b, c;
    cudaMemcpyFromSymbol(&a, g_emax, 8); cudaMemcpyFromSymbol(&b, g_err1, 8); cudaMemcpyFromSymbol(&c, g_err3, 8);
    printf("%s: samples %.3g  max|1 - x y0^2| = %.4g (2^%.2f)  one-step rel err = %.4g  third-order rel err = %.4g\n",
           wide ? "x in 2^[-200,200]" : "x in [1,4)", 148.0 * 8 * 256 * (1 << 14), a, log2(a), b, c);
  }
  return 0;
}
