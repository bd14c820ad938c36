// Weight gradient of a 3x3 stride-1 convolution on the padded-flat layout (conv_params.h: PadGeom).
//
//   dW[co, tap, ci] = sum over flat pixels f of  dY[f, co] * X[f + shift_tap, ci]
//
// dY is zero on the padding pixels of the layout, so the sum may run over every flat pixel. Both operands are
// "MN-major" (one pixel per 128-byte shared-memory row = the K index). A CTA owns [128 co] x [64 ci] x [one filter row =
// 3 taps] and a slice of the 128-pixel K tiles. Per K tile ONE slab of 130 pixels of X is loaded; the three taps
// (dw = -1, 0, +1) are ONE tcgen05.mma with N = 192 whose B descriptor walks its three 64-column slabs with a leading
// byte offset of 128 = one pixel row, i.e. the slabs are the same shared-memory data shifted by one pixel each.
// The accumulator stays in TMEM for the whole slice. Split-K partials are NOT combined with atomics (3.5 M scattered
// red.global.add per conv cost 25 us, twice the MMA time): every CTA stores its 128 x 192 fp32 tile with coalesced 256-bit
// stores to a scratch slot, and wgrad_reduce_kernel sums the slots in a fixed order into the OIHW gradient (deterministic).
//   warp 0 : TMA producer      warp 1 : MMA issuer (whole warp, one elected lane issues)      warps 2..5 : epilogue
#pragma once
#include "common.cuh"
#include "conv_params.h"

namespace cilrs {

__global__ void __launch_bounds__(WF_THREADS, 1) wgrad_flat_kernel(const __grid_constant__ WgradFlatParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  const int x_bytes = p.x_boxes * p.x_box_rows * 128;
  const int stage_bytes = 2 * WG_SLAB + x_bytes;
  uint64_t* bars = (uint64_t*)(smem + (size_t)p.num_stages * stage_bytes);
  uint64_t* full_bar = bars;
  uint64_t* empty_bar = bars + WF_MAX_STAGES;
  uint64_t* done_bar = bars + 2 * WF_MAX_STAGES;
  uint32_t* tmem_slot = (uint32_t*)(done_bar + 1);

  pdl_launch_dependents();
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
    tmem_alloc(tmem_slot, 256);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait();

  // work item of this CTA
  int wi = blockIdx.x;
  const int z = wi % p.split_z; wi /= p.split_z;
  const int tg = wi % 3; wi /= 3;   // filter row dh = tg - 1
  const int cic = wi % p.ci_chunks; wi /= p.ci_chunks;
  const int cob = wi;
  const int per = (p.k_tiles + p.split_z - 1) / p.split_z;
  const int kt_begin = z * per;
  const int kt_end = min(p.k_tiles, kt_begin + per);
  const int min_shift = p.tap_shift[tg * 3];  // dw = -1 of this filter row
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
    // whole warp in uniform control flow (descriptors stay in uniform registers), one elected lane issues.
    // One CTA per SM (shared memory) and the first allocation of the SM: the accumulator starts at TMEM column 0.
    if (tmem_base != 0) __trap();
    const bool leader = elect_one();
    const uint32_t idesc = umma_idesc_bf16(128, 192, 1, 1);
    const uint64_t descA0 = umma_desc_sw128(smem_u32(smem), WG_SLAB, 1024);             // two 64-co slabs 16 KB apart
    const uint64_t descB0 = umma_desc_sw128(smem_u32(smem) + 2 * WG_SLAB, 128, 1024);   // three 64-ci slabs one pixel row apart
    const uint32_t stage_units = (uint32_t)(stage_bytes >> 4);
    int stage = 0;
    uint32_t phase = 0;
    uint32_t first = 1;
    for (int kt = kt_begin; kt < kt_end; ++kt) {
      mbar_wait(&full_bar[stage], phase);
      tc_fence_after();
      const uint64_t da = descA0 + (uint64_t)((uint32_t)stage * stage_units);
      const uint64_t db = descB0 + (uint64_t)((uint32_t)stage * stage_units);
      if (leader) {
#pragma unroll
        for (int kk = 0; kk < 8; ++kk)  // 8 x (K = 16 pixels = 16 rows of 128 bytes = 128 sixteen-byte units)
          umma_bf16(0u, da + kk * 128, db + kk * 128, idesc, (first && kk == 0) ? 0u : 1u);
        umma_commit(&empty_bar[stage]);
      }
      __syncwarp();
      first = 0;
      if (++stage == p.num_stages) { stage = 0; phase ^= 1; }
    }
    if (leader) umma_commit(done_bar);
    __syncwarp();
  } else {
    const int q = warp & 3;
    const int row = q * 32 + lane;
    // scratch slot of this CTA: [tile = (cob, cic, tg)][z][128 rows][192 columns = 3 taps x 64 ci] fp32
    const int tile = (cob * p.ci_chunks + cic) * 3 + tg;
    float* dst = p.scratch + (((size_t)tile * p.split_z + z) * 128 + row) * 192;
    if (kt_end > kt_begin) {
      mbar_wait(done_bar, 0);
      tc_fence_after();
      for (int c0 = 0; c0 < 192; c0 += 32) {
        uint32_t v[32];
        __syncwarp();
        tmem_ld_32x32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)c0, v);
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 4; ++j)
          asm volatile("st.global.v8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(dst + c0 + j * 8), "r"(v[j * 8]), "r"(v[j * 8 + 1]),
                       "r"(v[j * 8 + 2]), "r"(v[j * 8 + 3]), "r"(v[j * 8 + 4]), "r"(v[j * 8 + 5]), "r"(v[j * 8 + 6]), "r"(v[j * 8 + 7])
                       : "memory");
      }
    } else {
      // a K slice past the end (split_z does not divide the tile count): the reducer still reads this slot
      const uint4 zz = make_uint4(0, 0, 0, 0);
      for (int c0 = 0; c0 < 192; c0 += 4) *(uint4*)(dst + c0) = zz;
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    __syncwarp();
    tmem_dealloc(tmem_base, 256);
  }
}

// ------------------------------------------------------------------------------------------------------------------
// Split-K reduction of the flat wgrad partials into the fp32 OIHW gradient (accumulating: grad += sum over z).
// One CTA per (output channel co, 64-channel input chunk): its 192 threads sum, for each of the three filter rows, the
// z partial rows (coalesced 768-byte reads), permute through shared memory into the OIHW order ci*9 + tap and add the
// 576 contiguous floats to the gradient. All flat convolutions of a backward part are served by one launch (job table).
// ------------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(768) wgrad_reduce_kernel(const __grid_constant__ WgradReduceJobs jobs, float* __restrict__ grads) {
  __shared__ float s_part[8][576];   // [K-slice chain][576] when a CTA serves one output channel, [output channel][576] when it serves several
  pdl_entry();
  int j = 0;
  while (j + 1 < jobs.n && (int)blockIdx.x >= jobs.job[j + 1].first_block) ++j;
  const WgradReduceJob& jb = jobs.job[j];
  const int u = blockIdx.x - jb.first_block;
  const int cgrp = u / jb.ci_chunks, cic = u - cgrp * jb.ci_chunks;
  const int t = threadIdx.x % 192;   // column of the 192-wide tile: tap-in-row (t >> 6), input channel (t & 63)
  const int zs = threadIdx.x / 192;  // this thread sums the K slices z = zs, zs + nz, ... (nz = blockDim.x / 192 chains per column)
  const int nz = blockDim.x / 192;
  if (jb.rows > 1) {
    // Shallow splits (layers 3-4: 6 / 1 K slices): one CTA per output channel was ~20 000 CTAs of three 768-byte reads each -
    // pure CTA-launch overhead (61-90 us per launch measured). A CTA now serves `rows` (4 / 8) output channels: every load is
    // issued before the first use, one barrier, `rows` x 2304 contiguous bytes out.
    const int co0 = cgrp * jb.rows;
    if (zs == 0) {
      float acc[8][3];
#pragma unroll
      for (int r = 0; r < 8; ++r) {
        if (r < jb.rows) {
          const int co = co0 + r, cob = co >> 7, row = co & 127;
#pragma unroll
          for (int tg = 0; tg < 3; ++tg) {
            const int tile = (cob * jb.ci_chunks + cic) * 3 + tg;
            const float* src = jb.scratch + (((size_t)tile * jb.split_z) * 128 + row) * 192 + t;
            float a = 0.f;
            for (int z = 0; z < jb.split_z; ++z) a += __ldcg(src + (size_t)z * 128 * 192);
            acc[r][tg] = a;
          }
        }
      }
#pragma unroll
      for (int r = 0; r < 8; ++r)
        if (r < jb.rows) {
#pragma unroll
          for (int tg = 0; tg < 3; ++tg) s_part[r][(t & 63) * 9 + tg * 3 + (t >> 6)] = acc[r][tg];
        }
    }
    __syncthreads();
    for (int r = 0; r < jb.rows; ++r) {
      float* g = grads + jb.grad_off + ((size_t)(co0 + r) * jb.cin + cic * 64) * 9;
      for (int o = threadIdx.x; o < 576; o += blockDim.x) g[o] += s_part[r][o];
    }
    return;
  }
  const int co = cgrp;
  const int cob = co >> 7, row = co & 127;
#pragma unroll
  for (int tg = 0; tg < 3; ++tg) {
    const int tile = (cob * jb.ci_chunks + cic) * 3 + tg;
    const float* src = jb.scratch + (((size_t)tile * jb.split_z) * 128 + row) * 192 + t;
    float acc = 0.f;
#pragma unroll 4
    for (int z = zs; z < jb.split_z; z += nz) acc += __ldcg(src + (size_t)z * 128 * 192);
    s_part[zs][(t & 63) * 9 + tg * 3 + (t >> 6)] = acc;
  }
  __syncthreads();
  // 576 contiguous floats of the OIHW gradient: ci*9 + tap
  float* g = grads + jb.grad_off + ((size_t)co * jb.cin + cic * 64) * 9;
  for (int o = threadIdx.x; o < 576; o += blockDim.x) {
    float v = s_part[0][o];
    for (int k = 1; k < nz; ++k) v += s_part[k][o];
    g[o] += v;
  }
}

}  // namespace cilrs
