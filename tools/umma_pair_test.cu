// Hardware experiment (not product code): tcgen05.mma.cta_group::2 on a cluster of two CTAs, as the conv kernel would use
// it. D[256 x N] = A[256 x 64] * B[N x 64]^T, both K-major SWIZZLE_128B. CTA r of the pair holds A rows r*128.. and B rows
// r*N/2.. in ITS shared memory; the leader (rank 0) issues the MMAs; each CTA reads its 128 accumulator rows from its own
// TMEM. Also exercises what the pipeline needs: TMA loads of both CTAs completing on the leader's mbarrier, the multicast
// commit, a remote mbarrier arrive from the peer, cluster barriers, 2-CTA TMEM allocation.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O2 -std=c++17 -I cilrs-autonomous-driving-carla_b200/csrc \
//        tools/umma_pair_test.cu -o tools/umma_pair_test
#include "common.cuh"
#include "pair.cuh"
#include <cudaTypedefs.h>
#include <stdio.h>
#include <stdlib.h>
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

__global__ void __launch_bounds__(128, 1) pair_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                                                      int N, float* out) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* sA = smem;                 // 128 rows x 128 B
  uint8_t* sB = smem + 128 * 128;     // N/2 rows x 128 B
  uint64_t* full = (uint64_t*)(smem + 128 * 128 + 128 * 128);
  uint64_t* done = full + 1;
  uint64_t* ready = full + 2;         // leader only: both CTAs are ready (count 2: one local, one remote arrive)
  uint32_t* slot = (uint32_t*)(full + 3);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  if (threadIdx.x == 0) {
    mbar_init(full, 1);
    mbar_init(done, 1);
    mbar_init(ready, 2);
    fence_barrier_init();
  }
  if (warp == 0) {
    tmem_alloc_pair(slot, 128);
    tmem_relinquish_pair();
  }
  tc_fence_before();
  cluster_sync_all();   // both CTAs' barriers are initialised before any remote access
  tc_fence_after();
  const uint32_t tmem = *slot;
  if (threadIdx.x == 0) {
    // both CTAs: their halves of the operands, completing on the LEADER's barrier
    const uint32_t half_bytes = (uint32_t)(128 * 128 + (N / 2) * 128);
    if (rank == 0) mbar_arrive_expect_tx(full, 2 * half_bytes);
    tma_load_2d_pair(&tmA, full, sA, 0, (int)rank * 128);
    tma_load_2d_pair(&tmB, full, sB, 0, (int)rank * (N / 2));
    mbar_arrive_remote(ready, 0);  // "my accumulator is free" handshake: the leader waits for both
    if (rank == 0) {
      mbar_wait(ready, 0);
      mbar_wait(full, 0);
      tc_fence_after();
      const uint32_t idesc = umma_idesc_bf16(256, N, 0, 0);
      for (int kk = 0; kk < 4; ++kk) {
        const uint64_t da = umma_desc_sw128(smem_u32(sA) + kk * 32, 16, 1024);
        const uint64_t db = umma_desc_sw128(smem_u32(sB) + kk * 32, 16, 1024);
        umma_bf16_pair(tmem, da, db, idesc, kk ? 1u : 0u);
      }
      umma_commit_pair(done, 3);   // arrives on `done` of both CTAs
    }
  }
  __syncwarp();
  mbar_wait(done, 0);
  tc_fence_after();
  uint32_t v[32];
  for (int c0 = 0; c0 < N; c0 += 32) {
    tmem_ld_32x32(tmem + ((uint32_t)(warp * 32) << 16) + c0, v);
    tmem_ld_wait();
    for (int e = 0; e < 32; ++e) out[((int)rank * 128 + warp * 32 + lane) * N + c0 + e] = __uint_as_float(v[e]);
  }
  tc_fence_before();
  cluster_sync_all();   // nobody exits while the peer may still touch its shared memory / TMEM
  if (warp == 0) tmem_dealloc_pair(tmem, 128);
}

static float bf(float x) { return __bfloat162float(__float2bfloat16(x)); }

int main() {
  int fails = 0;
  for (int N : {64, 128, 256}) {
    if (N > 128) break;  // this test allocates 128 TMEM columns
    std::vector<float> A(256 * 64), B(N * 64);
    srand(7 + N);
    auto rnd = []() { return (float)((rand() % 17) - 8) / 8.f; };
    for (auto& x : A) x = bf(rnd());
    for (auto& x : B) x = bf(rnd());
    std::vector<__nv_bfloat16> hA(A.size()), hB(B.size());
    for (size_t i = 0; i < A.size(); ++i) hA[i] = __float2bfloat16(A[i]);
    for (size_t i = 0; i < B.size(); ++i) hB[i] = __float2bfloat16(B[i]);
    __nv_bfloat16 *dA, *dB;
    float* dout;
    cudaMalloc(&dA, hA.size() * 2); cudaMalloc(&dB, hB.size() * 2); cudaMalloc(&dout, 256 * N * 4);
    cudaMemcpy(dA, hA.data(), hA.size() * 2, cudaMemcpyHostToDevice);
    cudaMemcpy(dB, hB.data(), hB.size() * 2, cudaMemcpyHostToDevice);
    cudaMemset(dout, 0, 256 * N * 4);
    CUtensorMap tmA, tmB;
    if (encode2d(&tmA, dA, 64, 256, 64, 128) || encode2d(&tmB, dB, 64, N, 64, N / 2)) { printf("encode failed\n"); return 1; }
    const int smem_bytes = 2 * 128 * 128 + 1024 + 256;
    cudaFuncSetAttribute(pair_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes);
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(2); cfg.blockDim = dim3(128); cfg.dynamicSmemBytes = smem_bytes;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
    cudaError_t e = cudaLaunchKernelEx(&cfg, pair_kernel, tmA, tmB, N, dout);
    if (e == cudaSuccess) e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("N %d: CUDA error %s\n", N, cudaGetErrorString(e)); return 2; }
    std::vector<float> out(256 * N);
    cudaMemcpy(out.data(), dout, out.size() * 4, cudaMemcpyDeviceToHost);
    double maxerr = 0;
    for (int m = 0; m < 256; ++m)
      for (int n = 0; n < N; ++n) {
        double r = 0;
        for (int k = 0; k < 64; ++k) r += (double)A[m * 64 + k] * B[n * 64 + k];
        const double d = fabs(r - out[m * N + n]);
        if (d > maxerr) maxerr = d;
      }
    printf("cta_group::2 M=256 N=%d: max |err| = %.3g %s\n", N, maxerr, maxerr < 1e-3 ? "OK" : "MISMATCH");
    if (maxerr >= 1e-3) ++fails;
  }
  return fails;
}
