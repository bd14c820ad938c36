// Hardware experiment (not product code): FP64 / conversion throughput per SM on B200 (sizes the per-channel BatchNorm
// finalize arithmetic of the deferred-finalize path).
#include <stdio.h>
#include <cuda_runtime.h>
template <int MODE>
__global__ void k(double* out, int iters, double seed) {
  double a = seed + threadIdx.x, b = 1.0000001, c = 0.5;
  float fa = (float)a, fb = 1.0000001f, fc = 0.5f;
  long long ia = (long long)a;
  for (int i = 0; i < iters; ++i) {
    if (MODE == 0) {
#pragma unroll
      for (int j = 0; j < 8; ++j) a = fma(a, b, c);
    } else if (MODE == 1) {
#pragma unroll
      for (int j = 0; j < 8; ++j) fa = fmaf(fa, fb, fc);
    } else if (MODE == 2) {
#pragma unroll
      for (int j = 0; j < 8; ++j) { fa = (float)a; a = a + (double)fa; }   // D2F + F2D + DADD
    } else {
#pragma unroll
      for (int j = 0; j < 8; ++j) { fa += (float)ia; ia += (long long)fa; }  // I2F.S64 + F2I.S64
    }
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = a + fa + (double)ia;
}
template <int MODE>
void run(const char* name, int ops_per_iter) {
  double* d; cudaMalloc(&d, 148 * 8 * 256 * 8);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  const int iters = 4096;
  k<MODE><<<148 * 8, 256>>>(d, 16, 1.0);
  cudaEventRecord(e0);
  k<MODE><<<148 * 8, 256>>>(d, iters, 1.0);
  cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  const double ops = 148.0 * 8 * 256 * (double)iters * ops_per_iter;
  printf("%-28s %.3f ms  %.1f Gop/s  = %.2f ops/clk/SM at 1.9 GHz\n", name, ms, ops / ms * 1e-6, ops / (ms * 1e-3) / 148 / 1.9e9);
}
int main() {
  run<0>("DFMA", 8);
  run<1>("FFMA", 8);
  run<2>("D2F+F2D+DADD (x3)", 24);
  run<3>("I2F.S64+F2I.S64+FADD (x3)", 24);
  return 0;
}
