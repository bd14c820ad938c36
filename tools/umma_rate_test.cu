// Hardware experiment (not product code): issue rate / execution time of tcgen05.mma cta_group::1 kind::f16, M=128,
// for N in {64,128,256}, K-major SW128 operands, A descriptor aligned vs row-shifted; one CTA per SM on `grid` SMs.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O2 -std=c++17 -I cilrs-autonomous-driving-carla_b200/csrc tools/umma_rate_test.cu -o tools/umma_rate_test
#include "common.cuh"
#include <stdio.h>
#include <vector>
using namespace cilrs;

__global__ void __launch_bounds__(128, 1) rate_kernel(int N, int shift, int reps, int mode, long long* out) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint8_t* sA = smem;                 // 320 rows x 128 B
  uint8_t* sB = smem + 320 * 128;     // 9 x 256 rows x 128 B... (mode 3: nine weight tiles)
  uint64_t* done = (uint64_t*)(sB + 9 * 128 * 128);
  uint32_t* slot = (uint32_t*)(done + 1);
  for (int i = threadIdx.x; i < (320 + 9 * 128) * 128 / 4; i += 128) ((uint32_t*)smem)[i] = 0x3c003c00u + (i & 3);
  fence_proxy_async();
  if (threadIdx.x == 0) { mbar_init(done, 1); fence_barrier_init(); }
  if (threadIdx.x < 32) { tmem_alloc(slot, 512); tmem_relinquish(); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *slot;
  if (threadIdx.x == 0) {
    const uint32_t idesc = umma_idesc_bf16(128, N, 0, 0);
    const uint32_t a0 = smem_u32(sA) + shift * 128, b0 = smem_u32(sB);
    long long t0 = clock64();
    for (int r = 0; r < reps; ++r) {
      // mode 0: same accumulator every time; mode 1: alternate two accumulators; mode 2: 4 sub-tiles (different A rows) like mt=4
      const uint32_t d = tmem + (mode == 1 ? (r & 1) * 256 : (mode == 2 ? (r & 3) * 64 : 0));
      uint32_t a = a0 + (mode == 2 ? (r & 1) * 128 * 128 : 0);
      uint32_t bb = b0;
      if (mode >= 3) {
        // like the conv kernel: nine taps = nine weight tiles and nine row shifts of the activation slab
        const int t = r % 9;
        a = a0 + (uint32_t)(((t / 3) * 51 + (t % 3)) * 128);
        bb = b0 + (uint32_t)(t * (N <= 128 ? N : 128) * 128);
      }
#pragma unroll
      for (int kk = 0; kk < 4; ++kk) {
        const uint64_t da = umma_desc_sw128(a + kk * 32, 16, 1024);
        const uint64_t db = umma_desc_sw128(bb + kk * 32, 16, 1024);
        umma_bf16(d, da, db, idesc, 1u);
      }
    }
    long long t1 = clock64();
    umma_commit(done);
    mbar_wait(done, 0);
    long long t2 = clock64();
    if (blockIdx.x == 0) { out[0] = t1 - t0; out[1] = t2 - t0; }
  }
  __syncthreads();
  tc_fence_before();
  __syncthreads();
  if (threadIdx.x < 32) tmem_dealloc(tmem, 512);
}

int main() {
  long long* dout;
  cudaMalloc(&dout, 16);
  const int smem_bytes = (320 + 9 * 128) * 128 + 1024 + 64;
  cudaFuncSetAttribute(rate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes);
  const int reps = 256;
  for (int grid : {148}) {
    for (int mode = 0; mode < 4; ++mode) {
      for (int N : {64, 128, 256}) {
        if (mode == 2 && N != 64) continue;
        if (mode == 3 && N == 256) continue;
        for (int shift : {0, 53}) {
          long long h[2];
          rate_kernel<<<grid, 128, smem_bytes>>>(N, shift, reps, mode, dout);
          cudaError_t e = cudaDeviceSynchronize();
          if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
          cudaMemcpy(h, dout, 16, cudaMemcpyDeviceToHost);
          printf("grid %3d mode %d N %3d shift %2d: issue %.1f clk/MMA, complete %.1f clk/MMA (ideal %d)\n", grid, mode, N, shift,
                 (double)h[0] / (reps * 4), (double)h[1] / (reps * 4), N / 2);
        }
      }
    }
  }
  return 0;
}
