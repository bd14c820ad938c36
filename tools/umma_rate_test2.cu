// Hardware experiment (not product code): how lean must the tcgen05.mma issue loop be? Compares descriptor handling
// variants for M=128, N in {64,128,256}, nine taps x 4 K-steps with distinct operand addresses (like the conv kernel).
#include "common.cuh"
#include <stdio.h>
using namespace cilrs;

CILRS_DEVINL bool try_wait_nohint(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile("{\n\t.reg .pred P1;\n\tmbarrier.try_wait.parity.shared::cta.b64 P1, [%1], %2;\n\tselp.b32 %0, 1, 0, P1;\n\t}"
               : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
  return ok != 0;
}
CILRS_DEVINL bool test_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile("{\n\t.reg .pred P1;\n\tmbarrier.test_wait.parity.shared::cta.b64 P1, [%1], %2;\n\tselp.b32 %0, 1, 0, P1;\n\t}"
               : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
  return ok != 0;
}
CILRS_DEVINL void mma_acc(uint32_t d, uint64_t da, uint64_t db, uint32_t idesc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, 1, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
               ::"r"(d), "l"(da), "l"(db), "r"(idesc) : "memory");
}

__global__ void __launch_bounds__(128, 1) rate_kernel(int N, int reps, int variant, const int* __restrict__ shifts, long long* out) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint8_t* sA = smem;                 // 640 rows x 128 B
  uint8_t* sB = smem + 640 * 128;     // 9 tiles of up to 128 rows
  uint64_t* done = (uint64_t*)(sB + 5 * 128 * 128);
  uint32_t* slot = (uint32_t*)(done + 4);
  for (int i = threadIdx.x; i < (640 + 5 * 128) * 128 / 4; i += 128) ((uint32_t*)smem)[i] = 0x3c003c00u + (i & 3);
  fence_proxy_async();
  if (threadIdx.x == 0) { mbar_init(done, 1); mbar_init(done + 2, 1 << 20); mbar_init(done + 3, 1); fence_barrier_init(); mbar_arrive(done + 3); }
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
            const uint32_t a = a0 + (uint32_t)(shifts[t] * 128), bb = b0 + (uint32_t)((t & 3) * nb * 128);
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
            const uint64_t da = dA0 + (uint64_t)(shifts[t] * 8), db = dB0 + (uint64_t)((t & 3) * nb * 8);
#pragma unroll
            for (int kk = 0; kk < 4; ++kk) mma_acc(0u, da + kk * 2, db + kk * 2, idesc);
          }
      }
    } else if (variant >= 6) {
      // per-tap synchronisation overheads with one sub-tile (4 MMAs per tap), barriers that are already complete:
      //  6 = mbar_wait + fence + MMAs + commit + syncwarp (the conv kernel's tap loop)   7 = no commit
      //  8 = wait only (no fence, no commit)   9 = probe the next tap's barrier before issuing this tap's MMAs
      const uint64_t dA0 = umma_desc_sw128(a0, 16, 1024), dB0 = umma_desc_sw128(b0, 16, 1024);
      const bool leader = elect_one();
      bool ok_next = mbar_try_wait(done + 3, 0);
      for (int r = 0; r < reps; ++r)
        for (int t = 0; t < 9; ++t) {
          if (variant == 9) {
            if (!ok_next) mbar_wait(done + 3, 0);
            tc_fence_after();
          } else if (variant == 10) {
            while (!try_wait_nohint(done + 3, 0)) { }
            tc_fence_after();
          } else if (variant == 11) {
            while (!test_wait(done + 3, 0)) { }
            tc_fence_after();
          } else if (variant == 12) {
            if (leader) { while (!try_wait_nohint(done + 3, 0)) { } }
            __syncwarp();
            tc_fence_after();
          } else if (variant == 13) {
            tc_fence_after();   // no wait at all: cost of the rest of the tap loop
          } else {
            mbar_wait(done + 3, 0);
            if (variant != 8) tc_fence_after();
          }
          const uint64_t da = dA0 + (uint64_t)(shifts[t] * 8), db = dB0 + (uint64_t)((t & 3) * nb * 8);
          if (variant == 9) ok_next = mbar_try_wait(done + 3, 0);
          if (leader) {
#pragma unroll
            for (int kk = 0; kk < 4; ++kk) mma_acc(0u, da + kk * 2, db + kk * 2, idesc);
            if (variant == 6 || variant == 9) umma_commit(done + 2);
          }
          __syncwarp();
        }
    } else if (variant >= 3) {
      // variant 2 + what the conv kernel adds: 3 = a tcgen05.commit per tap, 4 = four sub-tiles (A rows m*128 apart, own
      // accumulators) per tap, 5 = both
      const uint64_t dA0 = umma_desc_sw128(a0, 16, 1024), dB0 = umma_desc_sw128(b0, 16, 1024);
      const bool leader = elect_one();
      const int mt = variant >= 4 ? 4 : 1;
      for (int r = 0; r < reps / mt; ++r)
        for (int t = 0; t < 9; ++t) {
          const uint64_t da = dA0 + (uint64_t)(shifts[t] * 8), db = dB0 + (uint64_t)((t & 3) * nb * 8);
          if (leader) {
            for (int m = 0; m < mt; ++m) {
#pragma unroll
              for (int kk = 0; kk < 4; ++kk) mma_acc((uint32_t)(m * N) & 511u, da + (uint64_t)(m * 1024) + kk * 2, db + kk * 2, idesc);
            }
            if (variant == 3 || variant == 5) umma_commit(done + 2);
          }
          __syncwarp();
        }
    } else {
      // whole warp runs the loop (uniform control flow), one elected lane issues
      const uint64_t dA0 = umma_desc_sw128(a0, 16, 1024), dB0 = umma_desc_sw128(b0, 16, 1024);
      const bool leader = elect_one();
      for (int r = 0; r < reps; ++r)
        for (int t = 0; t < 9; ++t) {
          const uint64_t da = dA0 + (uint64_t)(shifts[t] * 8), db = dB0 + (uint64_t)((t & 3) * nb * 8);
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
  const int smem_bytes = (640 + 5 * 128) * 128 + 1024 + 64;
  cudaFuncSetAttribute(rate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes);
  const int reps = 32;
  for (int variant = 6; variant < 14; ++variant)
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
