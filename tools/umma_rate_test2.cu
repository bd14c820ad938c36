// Hardware experiment (not product code): how lean must the tcgen05.mma issue loop be? Compares descriptor handling
// variants for M=128, N in {64,128,256}, nine taps x 4 K-steps with distinct operand addresses (like the conv kernel).
#include "common.cuh"
#include <stdio.h>
using namespace cilrs;

CILRS_DEVINL void mma_acc(uint32_t d, uint64_t da, uint64_t db, uint32_t idesc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, 1, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
               ::"r"(d), "l"(da), "l"(db), "r"(idesc) : "memory");
}

__global__ void __launch_bounds__(128, 1) rate_kernel(int N, int reps, int variant, const int* __restrict__ shifts, long long* out) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint8_t* sA = smem;                 // 320 rows x 128 B
  uint8_t* sB = smem + 320 * 128;     // 9 tiles of up to 128 rows
  uint64_t* done = (uint64_t*)(sB + 9 * 128 * 128);
  uint32_t* slot = (uint32_t*)(done + 1);
  for (int i = threadIdx.x; i < (320 + 9 * 128) * 128 / 4; i += 128) ((uint32_t*)smem)[i] = 0x3c003c00u + (i & 3);
  fence_proxy_async();
  if (threadIdx.x == 0) { mbar_init(done, 1); fence_barrier_init(); }
  if (threadIdx.x < 32) { tmem_alloc(slot, 512); tmem_relinquish(); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_ld = *slot;
  if (tmem_ld != 0) __trap();
  if (threadIdx.x < 32) {
    const uint32_t idesc = umma_idesc_bf16(128, N, 0, 0);
    const uint32_t a0 = smem_u32(sA), b0 = smem_u32(sB);
    const int nb = N <= 128 ? N : 128;
    long long t0 = clock64();
    if (variant == 0) {
      // as in conv_flat.cuh today: lane 0 only, descriptors rebuilt per MMA, TMEM address from shared memory
      if (threadIdx.x == 0) {
        for (int r = 0; r < reps; ++r)
          for (int t = 0; t < 9; ++t) {
            const uint32_t a = a0 + (uint32_t)(shifts[t] * 128), bb = b0 + (uint32_t)((N == 256 ? (t & 7) : t) * nb * 128);
#pragma unroll
            for (int kk = 0; kk < 4; ++kk)
              umma_bf16(tmem_ld, umma_desc_sw128(a + kk * 32, 16, 1024), umma_desc_sw128(bb + kk * 32, 16, 1024), idesc, 1u);
          }
      }
    } else if (variant == 1) {
      // lane 0 only, TMEM address a constant, descriptors advanced by 64-bit adds
      if (threadIdx.x == 0) {
        const uint64_t dA0 = umma_desc_sw128(a0, 16, 1024), dB0 = umma_desc_sw128(b0, 16, 1024);
        for (int r = 0; r < reps; ++r)
          for (int t = 0; t < 9; ++t) {
            const uint64_t da = dA0 + (uint64_t)(shifts[t] * 8), db = dB0 + (uint64_t)((N == 256 ? (t & 7) : t) * nb * 8);
#pragma unroll
            for (int kk = 0; kk < 4; ++kk) mma_acc(0u, da + kk * 2, db + kk * 2, idesc);
          }
      }
    } else {
      // whole warp runs the loop (uniform control flow), one elected lane issues
      const uint64_t dA0 = umma_desc_sw128(a0, 16, 1024), dB0 = umma_desc_sw128(b0, 16, 1024);
      const bool leader = elect_one();
      for (int r = 0; r < reps; ++r)
        for (int t = 0; t < 9; ++t) {
          const uint64_t da = dA0 + (uint64_t)(shifts[t] * 8), db = dB0 + (uint64_t)((N == 256 ? (t & 7) : t) * nb * 8);
          if (leader) {
#pragma unroll
            for (int kk = 0; kk < 4; ++kk) mma_acc(0u, da + kk * 2, db + kk * 2, idesc);
          }
          __syncwarp();
        }
    }
    long long t1 = clock64();
    if (threadIdx.x == 0) {
      umma_commit(done);
      mbar_wait(done, 0);
      long long t2 = clock64();
      if (blockIdx.x == 0) { out[0] = t1 - t0; out[1] = t2 - t0; }
    }
  }
  __syncthreads();
  tc_fence_before();
  __syncthreads();
  if (threadIdx.x < 32) tmem_dealloc(0u, 512);
}

int main() {
  long long* dout;
  int* dsh;
  cudaMalloc(&dout, 16);
  cudaMalloc(&dsh, 9 * 4);
  int sh[9];
  for (int t = 0; t < 9; ++t) sh[t] = (t / 3) * 51 + (t % 3);
  cudaMemcpy(dsh, sh, sizeof(sh), cudaMemcpyHostToDevice);
  const int smem_bytes = (320 + 9 * 128) * 128 + 1024 + 64;
  cudaFuncSetAttribute(rate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes);
  const int reps = 32;
  for (int variant = 0; variant < 3; ++variant)
    for (int N : {64, 128, 256}) {
      long long h[2];
      rate_kernel<<<148, 128, smem_bytes>>>(N, reps, variant, dsh, dout);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
      cudaMemcpy(h, dout, 16, cudaMemcpyDeviceToHost);
      printf("variant %d N %3d: issue %.1f clk/MMA, complete %.1f clk/MMA (ideal %d)\n", variant, N, (double)h[0] / (reps * 36),
             (double)h[1] / (reps * 36), N / 2);
    }
  return 0;
}
