// Hardware experiment (not product code): can a SWIZZLE_128B UMMA shared-memory descriptor start at a row that is not a
// multiple of 8 (i.e. not 1024-byte aligned)?  The padded-flat conv design reads the 3x3 taps as row-shifted views of one
// TMA-loaded activation slab, which needs exactly that.
//   mode 0: K-major A (rows = pixels), start = slab + s*128, base_offset field = 0
//   mode 1: K-major A, base_offset = (start >> 7) & 7
//   mode 2: MN-major A and B (rows = K = pixels), both shifted by s rows, base_offset = 0
//   mode 3: MN-major, base_offset = (start >> 7) & 7
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O2 -std=c++17 -I cilrs-autonomous-driving-carla_b200/csrc \
//        tools/umma_shift_test.cu -o tools/umma_shift_test
#include "common.cuh"
#include <cudaTypedefs.h>
#include <stdio.h>
#include <stdlib.h>
#include <math.h>
#include <vector>

using namespace cilrs;

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static int encode2d(CUtensorMap* m, const void* base, int inner, int rows, int box_inner, int box_rows) {
  void* f = nullptr;
  cudaDriverEntryPointQueryResult q;
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &q) != cudaSuccess) return 1;
  cuuint64_t gdim[2] = {(cuuint64_t)inner, (cuuint64_t)rows};
  cuuint64_t gstr[1] = {(cuuint64_t)inner * 2};
  cuuint32_t box[2] = {(cuuint32_t)box_inner, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  return ((EncodeTiledFn)f)(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, (void*)base, gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                            CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS;
}

constexpr int ROWS = 160;  // slab rows (pixels)

__device__ __forceinline__ uint64_t desc_bo(uint32_t saddr, uint32_t lbo, uint32_t sbo, int use_bo) {
  uint64_t d = umma_desc_sw128(saddr, lbo, sbo);
  if (use_bo) d |= (uint64_t)((saddr >> 7) & 7) << 49;
  return d;
}

// A: [ROWS][128] bf16 (two 64-channel slabs in smem), B: [ROWS][64] (MN-major) or [64][64] (K-major: n rows x k cols)
__global__ void __launch_bounds__(128, 1) test_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                                                      int shift, int mode, float* out) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint8_t* sA0 = smem;                    // ROWS x 128 B (channels 0..63)
  uint8_t* sA1 = smem + ROWS * 128;       // channels 64..127
  uint8_t* sB = smem + 2 * ROWS * 128;    // ROWS x 128 B
  uint64_t* bar = (uint64_t*)(smem + 3 * ROWS * 128);
  uint64_t* done = bar + 1;
  uint32_t* slot = (uint32_t*)(bar + 2);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const bool mn = mode >= 2;
  const int use_bo = mode & 1;
  if (threadIdx.x == 0) {
    mbar_init(bar, 1);
    mbar_init(done, 1);
    fence_barrier_init();
  }
  if (warp == 0) {
    tmem_alloc(slot, 64);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *slot;
  if (threadIdx.x == 0) {
    const uint32_t bytes = (uint32_t)(2 * ROWS * 128 + (mn ? ROWS * 128 : 64 * 128));
    mbar_arrive_expect_tx(bar, bytes);
    tma_load_2d(&tmA, bar, sA0, 0, 0);
    tma_load_2d(&tmA, bar, sA1, 64, 0);
    tma_load_2d(&tmB, bar, sB, 0, 0);
    mbar_wait(bar, 0);
    tc_fence_after();
    if (!mn) {
      // D[128 pixels][64 n] = sum_k A[shift + pixel][k] * B[n][k], k = 0..63 (channels 0..63 of A)
      const uint32_t idesc = umma_idesc_bf16(128, 64, 0, 0);
      for (int kk = 0; kk < 4; ++kk) {
        const uint64_t da = desc_bo(smem_u32(sA0) + shift * 128 + kk * 32, 16, 1024, use_bo);
        const uint64_t db = umma_desc_sw128(smem_u32(sB) + kk * 32, 16, 1024);
        umma_bf16(tmem, da, db, idesc, kk ? 1u : 0u);
      }
    } else {
      // D[128 ch of A][64 ch of B] = sum over 128 rows r of A[shift + r][m] * B[shift + r][n]
      const uint32_t idesc = umma_idesc_bf16(128, 64, 1, 1);
      for (int kk = 0; kk < 8; ++kk) {
        const uint64_t da = desc_bo(smem_u32(sA0) + shift * 128 + kk * 2048, ROWS * 128, 1024, use_bo);
        const uint64_t db = desc_bo(smem_u32(sB) + shift * 128 + kk * 2048, ROWS * 128, 1024, use_bo);
        umma_bf16(tmem, da, db, idesc, kk ? 1u : 0u);
      }
    }
    umma_commit(done);
  }
  __syncwarp();
  mbar_wait(done, 0);
  tc_fence_after();
  uint32_t v[32];
  for (int c0 = 0; c0 < 64; c0 += 32) {
    tmem_ld_32x32(tmem + ((uint32_t)(warp * 32) << 16) + c0, v);
    tmem_ld_wait();
    for (int e = 0; e < 32; ++e) out[(warp * 32 + lane) * 64 + c0 + e] = __uint_as_float(v[e]);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 64);
}

static float bf(float x) {
  __nv_bfloat16 b = __float2bfloat16(x);
  return __bfloat162float(b);
}

int main() {
  std::vector<float> A(ROWS * 128), Bk(64 * 64), Bm(ROWS * 64);
  srand(1);
  auto rnd = []() { return (float)((rand() % 17) - 8) / 8.f; };
  for (auto& x : A) x = bf(rnd());
  for (auto& x : Bk) x = bf(rnd());
  for (auto& x : Bm) x = bf(rnd());
  std::vector<__nv_bfloat16> hA(A.size()), hBk(Bk.size()), hBm(Bm.size());
  for (size_t i = 0; i < A.size(); ++i) hA[i] = __float2bfloat16(A[i]);
  for (size_t i = 0; i < Bk.size(); ++i) hBk[i] = __float2bfloat16(Bk[i]);
  for (size_t i = 0; i < Bm.size(); ++i) hBm[i] = __float2bfloat16(Bm[i]);
  __nv_bfloat16 *dA, *dBk, *dBm;
  float* dout;
  cudaMalloc(&dA, hA.size() * 2); cudaMalloc(&dBk, hBk.size() * 2); cudaMalloc(&dBm, hBm.size() * 2);
  cudaMalloc(&dout, 128 * 64 * 4);
  cudaMemcpy(dA, hA.data(), hA.size() * 2, cudaMemcpyHostToDevice);
  cudaMemcpy(dBk, hBk.data(), hBk.size() * 2, cudaMemcpyHostToDevice);
  cudaMemcpy(dBm, hBm.data(), hBm.size() * 2, cudaMemcpyHostToDevice);
  CUtensorMap tmA, tmBk, tmBm;
  if (encode2d(&tmA, dA, 128, ROWS, 64, ROWS) || encode2d(&tmBk, dBk, 64, 64, 64, 64) || encode2d(&tmBm, dBm, 64, ROWS, 64, ROWS)) {
    printf("encode failed\n");
    return 1;
  }
  const int smem_bytes = 3 * ROWS * 128 + 1024 + 256;
  cudaFuncSetAttribute(test_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes);
  std::vector<float> out(128 * 64);
  for (int mode = 0; mode < 4; ++mode) {
    for (int shift = 0; shift <= 27; ++shift) {
      cudaMemset(dout, 0, 128 * 64 * 4);
      test_kernel<<<1, 128, smem_bytes>>>(tmA, mode >= 2 ? tmBm : tmBk, shift, mode, dout);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) {
        printf("mode %d shift %d: CUDA error %s\n", mode, shift, cudaGetErrorString(e));
        return 2;
      }
      cudaMemcpy(out.data(), dout, 128 * 64 * 4, cudaMemcpyDeviceToHost);
      double maxerr = 0;
      for (int m = 0; m < 128; ++m)
        for (int n = 0; n < 64; ++n) {
          double ref = 0;
          if (mode < 2) {
            for (int k = 0; k < 64; ++k) ref += (double)A[(shift + m) * 128 + k] * Bk[n * 64 + k];
          } else {
            for (int r = 0; r < 128; ++r) ref += (double)A[(shift + r) * 128 + m] * Bm[(shift + r) * 64 + n];
          }
          maxerr = fmax(maxerr, fabs(ref - out[m * 64 + n]));
        }
      printf("mode %d shift %2d maxerr %.4f %s\n", mode, shift, maxerr, maxerr < 1e-2 ? "OK" : "WRONG");
    }
  }
  return 0;
}
