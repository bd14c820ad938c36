// Weight gradient of a 3x3 stride-1 convolution on the padded-flat layout (conv_params.h: PadGeom).
//
//   dW[co, tap, ci] = sum over flat pixels f of  dY[f, co] * X[f + shift_tap, ci]
//
// dY is zero on the padding pixels of the layout, so the sum may run over every flat pixel. Both operands are
// "MN-major" (one pixel per 128-byte shared-memory row = the K index). Per 128-pixel K tile ONE slab of
// (128 + Wp + 1) pixels of X is loaded; the taps of the CTA's tap group are row-shifted descriptors into it (the old
// kernel loaded one X box per tap). A CTA owns [128 co] x [64 ci] x [4 or 5 taps] and a slice of the K tiles; the
// accumulator stays in TMEM for the whole slice and is added to the fp32 OIHW gradient with red.global.add.
//   warp 0 : TMA producer      warp 1 : MMA issuer      warps 2..5 : epilogue
#pragma once
#include "common.cuh"
#include "conv_params.h"

namespace cilrs {

__global__ void __launch_bounds__(WF_THREADS, 1) wgrad_flat_kernel(const __grid_constant__ WgradFlatParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  const int x_bytes = p.x_boxes * p.x_box_rows * 128;
  const int stage_bytes = 2 * WG_SLAB + x_bytes;
  uint64_t* bars = (uint64_t*)(smem + (size_t)p.num_stages * stage_bytes);
  uint64_t* full_bar = bars;
  uint64_t* empty_bar = bars + WF_MAX_STAGES;
  uint64_t* done_bar = bars + 2 * WF_MAX_STAGES;
  uint32_t* tmem_slot = (uint32_t*)(done_bar + 1);

  if (p.m_halves == 1) {
    // the A side always spans two 64-channel slabs; with 64 output channels the second one stays zero
    const uint4 z = make_uint4(0, 0, 0, 0);
    for (int st = 0; st < p.num_stages; ++st) {
      uint4* q = (uint4*)(smem + (size_t)st * stage_bytes + WG_SLAB);
      for (int i = threadIdx.x; i < WG_SLAB / 16; i += WF_THREADS) q[i] = z;
    }
    fence_proxy_async();
  }
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&p.tmDY);
    tma_prefetch_desc(&p.tmX);
    for (int i = 0; i < p.num_stages; ++i) {
      mbar_init(&full_bar[i], 1);
      mbar_init(&empty_bar[i], 1);
    }
    mbar_init(done_bar, 1);
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  // work item of this CTA
  int wi = blockIdx.x;
  const int z = wi % p.split_z; wi /= p.split_z;
  const int tg = wi % p.tap_groups; wi /= p.tap_groups;
  const int cic = wi % p.ci_chunks; wi /= p.ci_chunks;
  const int cob = wi;
  const int per = (p.k_tiles + p.split_z - 1) / p.split_z;
  const int kt_begin = z * per;
  const int kt_end = min(p.k_tiles, kt_begin + per);
  const int t_first = p.group_first[tg], t_count = p.group_count[tg];
  const int min_shift = p.tap_shift[t_first];  // tap shifts increase with the tap id
  const uint32_t tx_bytes = (uint32_t)(p.m_halves * WG_SLAB + x_bytes);

  if (warp == 0) {
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int kt = kt_begin; kt < kt_end; ++kt) {
        const int k0 = kt * 128;
        mbar_wait(&empty_bar[stage], phase ^ 1);
        uint8_t* s = smem + (size_t)stage * stage_bytes;
        mbar_arrive_expect_tx(&full_bar[stage], tx_bytes);
        for (int mh = 0; mh < p.m_halves; ++mh) tma_load_2d(&p.tmDY, &full_bar[stage], s + mh * WG_SLAB, cob * 128 + mh * 64, k0);
        for (int bx = 0; bx < p.x_boxes; ++bx)
          tma_load_2d(&p.tmX, &full_bar[stage], s + 2 * WG_SLAB + (size_t)bx * p.x_box_rows * 128, cic * 64,
                      k0 + min_shift + bx * p.x_box_rows);
        if (++stage == p.num_stages) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      const uint32_t idesc = umma_idesc_bf16(128, 64, 1, 1);
      int stage = 0;
      uint32_t phase = 0;
      uint32_t first = 1;
      for (int kt = kt_begin; kt < kt_end; ++kt) {
        mbar_wait(&full_bar[stage], phase);
        tc_fence_after();
        const uint32_t a_addr = smem_u32(smem + (size_t)stage * stage_bytes);
        const uint32_t x_addr = a_addr + 2 * WG_SLAB;
        for (int j = 0; j < t_count; ++j) {
          const uint32_t b_addr = x_addr + (uint32_t)((p.tap_shift[t_first + j] - min_shift) * 128);
#pragma unroll
          for (int kk = 0; kk < 8; ++kk) {  // 8 x (K = 16 pixels = 16 rows of 128 bytes)
            const uint64_t da = umma_desc_sw128(a_addr + kk * 2048, WG_SLAB, 1024);
            const uint64_t db = umma_desc_sw128(b_addr + kk * 2048, WG_SLAB, 1024);
            umma_bf16(tmem_base + (uint32_t)(j * 64), da, db, idesc, (first && kk == 0) ? 0u : 1u);
          }
        }
        first = 0;
        umma_commit(&empty_bar[stage]);
        if (++stage == p.num_stages) { stage = 0; phase ^= 1; }
      }
      umma_commit(done_bar);
    }
  } else {
    const int q = warp & 3;
    const int row = q * 32 + lane;
    const int co = cob * 128 + row;
    if (kt_end > kt_begin) {
      mbar_wait(done_bar, 0);
      tc_fence_after();
      const int ncols = 64 * t_count;
      for (int c0 = 0; c0 < ncols; c0 += 32) {
        uint32_t v[32];
        __syncwarp();
        tmem_ld_32x32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)c0, v);
        tmem_ld_wait();
        if (row < p.m_halves * 64 && co < p.cout) {
          const int t = t_first + (c0 >> 6);
          float* gp = p.grad + ((size_t)co * p.cin + cic * 64 + (c0 & 63)) * 9 + t;
#pragma unroll
          for (int e = 0; e < 32; ++e) atomicAdd(gp + e * 9, __uint_as_float(v[e]));
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    __syncwarp();
    tmem_dealloc(tmem_base, 512);
  }
}

}  // namespace cilrs
