// Weight gradient of a 3x3 stride-1 convolution on the padded-flat layout (conv_params.h: PadGeom).
//
//   dW[co, tap, ci] = sum over flat pixels f of  dY[f, co] * X[f + shift_tap, ci]
//
// dY is zero on the padding pixels of the layout, so the sum may run over every flat pixel. Both operands are
// "MN-major" (one pixel per 128-byte shared-memory row = the K index). A CTA owns [128 co] x [64 ci] x [one filter row =
// 3 taps] and a slice of the 128-pixel K tiles. Per K tile ONE slab of 130 pixels of X is loaded; the three taps
// (dw = -1, 0, +1) are ONE tcgen05.mma with N = 192 whose B descriptor walks its three 64-column slabs with a leading
// byte offset of 128 = one pixel row, i.e. the slabs are the same shared-memory data shifted by one pixel each.
// The accumulator stays in TMEM for the whole slice. Split-K: every K slice ADDS its 128 x 192 fp32 tile into ONE accumulator
// tile per (co block, ci chunk, filter row) with bulk asynchronous reductions (cp.reduce.async.bulk ... .add.f32: the tile is
// staged row by row in the shared memory the finished main loop no longer needs, one 768-byte reduction per row, performed
// by the L2). Round 1 stored every slice's tile to its own scratch slot and summed the slots in a separate launch: up to
// 49 x the weight bytes per conv written and read again (420 MB per step, 0.22 ms of wgrad_reduce_kernel); now that launch
// reads each accumulator once, clears it and permutes it into the OIHW gradient. (Scalar red.global.add straight into the OIHW
// gradient - 3.5 M scattered atomics per conv - cost twice the MMA time; the OIHW strides (36 B) also rule out a TMA tensor
// reduction.) The price: the order in which slices are added is not fixed, so weight gradients are reproducible to fp32
// rounding (~1e-7), not bit for bit.
//   warp 0 : TMA producer      warp 1 : MMA issuer (whole warp, one elected lane issues)      warps 2..5 : epilogue
#pragma once
#include "common.cuh"
#include "pair.cuh"
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
    // accumulator tile of this CTA's (cob, cic, tg): [128 rows][192 columns = 3 taps x 64 ci] fp32, shared by all K slices
    const int tile = (cob * p.ci_chunks + cic) * 3 + tg;
    float* dst = p.scratch + ((size_t)tile * 128 + row) * 192;
    if (kt_end > kt_begin) {
      mbar_wait(done_bar, 0);   // every MMA has completed: the stage buffers are free
      tc_fence_after();
      // stage this thread's row (768 B) at a 784-byte pitch: 16-byte stores of a quarter warp then hit 8 different bank groups
      uint8_t* srow = smem + (size_t)row * WF_STAGE_PITCH;
      for (int c0 = 0; c0 < 192; c0 += 32) {
        uint32_t v[32];
        __syncwarp();
        tmem_ld_32x32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)c0, v);
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 8; ++j) *(uint4*)(srow + (c0 + 4 * j) * 4) = make_uint4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
      }
      fence_proxy_async();   // the generic-proxy stores above are read by the async proxy
      asm volatile("cp.reduce.async.bulk.global.shared::cta.bulk_group.add.f32 [%0], [%1], %2;"
                   ::"l"(dst), "r"(smem_u32(srow)), "r"(768) : "memory");
      asm volatile("cp.async.bulk.commit_group;" ::: "memory");
      // (the staging rows have been read; the additions themselves complete before the grid does, which is what orders them
      //  before the fold launch)
      asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
    }
    // (a K slice past the end - split_z does not divide the tile count - has nothing to add)
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    __syncwarp();
    tmem_dealloc(tmem_base, 256);
  }
}

// ------------------------------------------------------------------------------------------------------------------
// The same contraction with the three filter rows of a (co block, ci chunk, K slice) as ONE CLUSTER of three CTAs.
// wgrad_flat_kernel is bound by L2 -> SM traffic, not by the tensor core: every dY tile is fetched by 3 x ci_chunks CTAs and
// every X slab by co_blocks of them (layer3: 115 MB through the crossbar for 12.8 MB of operands, 64 B per clock and SM
// asked of the L2 by 144 SMs at once). The three CTAs of a cluster need the SAME dY tile and X slabs that are the same pixels
// shifted by one padded row each (130 rows at -Wp, 0, +Wp): here they load the dY tile and ONE union slab of 130 + 2 Wp rows
// once per cluster - CTA 0 issues dY half 0, CTA 1 dY half 1, CTA 2 the X slab - with cp.async.bulk.tensor ...
// .multicast::cluster into all three shared memories (same offsets; every CTA's own `full` barrier counts all bytes), and
// CTA r's B descriptor starts r * Wp rows into the union slab. A stage may be refilled when all three CTAs are done with
// it: each MMA warp's tcgen05.commit arrives on the `empty` barrier of all three (multicast, count 3).
// L2 -> SM bytes per K tile and cluster: 32 KB + (130 + 2 Wp) * 128 B instead of 3 * 49 KB.
// ------------------------------------------------------------------------------------------------------------------
CILRS_DEVINL void tma_load_2d_mc(const CUtensorMap* m, uint64_t* bar, void* dst, int c0, int c1, uint16_t cta_mask) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1, {%3, %4}], [%2], %5;"
      ::"r"(smem_u32(dst)), "l"((uint64_t)m), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "h"(cta_mask)
      : "memory");
}
CILRS_DEVINL void umma_commit_mc(uint64_t* bar, uint16_t cta_mask) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
               ::"r"(smem_u32(bar)), "h"(cta_mask) : "memory");
}

__global__ void __launch_bounds__(WF_THREADS, 1) wgrad_flat3_kernel(const __grid_constant__ WgradFlatParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();   // = filter row: dh = rank - 1

  const int x_bytes = p.xu_rows * 128;
  const int stage_bytes = 2 * WG_SLAB + x_bytes;
  uint64_t* bars = (uint64_t*)(smem + (size_t)p.stages3 * stage_bytes);
  uint64_t* full_bar = bars;
  uint64_t* empty_bar = bars + WF_MAX_STAGES;
  uint64_t* done_bar = bars + 2 * WF_MAX_STAGES;
  uint32_t* tmem_slot = (uint32_t*)(done_bar + 1);

  pdl_launch_dependents();
  if (p.m_halves == 1) {
    const uint4 z = make_uint4(0, 0, 0, 0);
    for (int st = 0; st < p.stages3; ++st) {
      uint4* q = (uint4*)(smem + (size_t)st * stage_bytes + WG_SLAB);
      for (int i = threadIdx.x; i < WG_SLAB / 16; i += WF_THREADS) q[i] = z;
    }
    fence_proxy_async();
  }
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&p.tmDY);
    tma_prefetch_desc(&p.tmXU);
    for (int i = 0; i < p.stages3; ++i) {
      mbar_init(&full_bar[i], 1);
      mbar_init(&empty_bar[i], 3);   // the MMA warps of the three CTAs
    }
    mbar_init(done_bar, 1);
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, 256);
    tmem_relinquish();
  }
  tc_fence_before();
  cluster_sync_all();   // the peers' barriers are initialised (and their zero slabs written) before anything lands in them
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait();

  int wi = blockIdx.x / 3;
  const int z = wi % p.split_z; wi /= p.split_z;
  const int cic = wi % p.ci_chunks; wi /= p.ci_chunks;
  const int cob = wi;
  const int tg = (int)rank;
  const int per = (p.k_tiles + p.split_z - 1) / p.split_z;
  const int kt_begin = z * per;
  const int kt_end = min(p.k_tiles, kt_begin + per);
  const uint32_t tx_bytes = (uint32_t)(p.m_halves * WG_SLAB + x_bytes);   // every CTA receives every item
  // who issues what: dY half 0 -> CTA 0, dY half 1 (128 output channels) -> CTA 1, the X slab -> the next CTA
  const int x_issuer = p.m_halves;   // 2 (cout >= 128) or 1

  if (warp == 0) {
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int kt = kt_begin; kt < kt_end; ++kt) {
        const int k0 = kt * 128;
        mbar_wait(&empty_bar[stage], phase ^ 1);   // all three CTAs have consumed this stage
        uint8_t* s = smem + (size_t)stage * stage_bytes;
        mbar_arrive_expect_tx(&full_bar[stage], tx_bytes);
        if ((int)rank < p.m_halves) tma_load_2d_mc(&p.tmDY, &full_bar[stage], s + rank * WG_SLAB, cob * 128 + (int)rank * 64, k0, 7);
        if ((int)rank == x_issuer) tma_load_2d_mc(&p.tmXU, &full_bar[stage], s + 2 * WG_SLAB, cic * 64, k0 - p.wp - 1, 7);
        if (++stage == p.stages3) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1) {
    if (tmem_base != 0) __trap();
    const bool leader = elect_one();
    const uint32_t idesc = umma_idesc_bf16(128, 192, 1, 1);
    const uint64_t descA0 = umma_desc_sw128(smem_u32(smem), WG_SLAB, 1024);
    // filter row `rank`: its dw = -1 tap starts rank * Wp pixel rows into the union slab; the three taps are one pixel row apart
    const uint64_t descB0 = umma_desc_sw128(smem_u32(smem) + 2 * WG_SLAB + rank * (uint32_t)p.wp * 128u, 128, 1024);
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
        for (int kk = 0; kk < 8; ++kk)
          umma_bf16(0u, da + kk * 128, db + kk * 128, idesc, (first && kk == 0) ? 0u : 1u);
        umma_commit_mc(&empty_bar[stage], 7);
      }
      __syncwarp();
      first = 0;
      if (++stage == p.stages3) { stage = 0; phase ^= 1; }
    }
    if (leader) umma_commit(done_bar);
    __syncwarp();
  } else {
    const int q = warp & 3;
    const int row = q * 32 + lane;
    const int tile = (cob * p.ci_chunks + cic) * 3 + tg;
    float* dst = p.scratch + ((size_t)tile * 128 + row) * 192;
    if (kt_end > kt_begin) {
      mbar_wait(done_bar, 0);   // every MMA of THIS CTA has completed
      tc_fence_after();
    }
    // The staging below overwrites stage buffers the peers may still be multicasting into (they can be one K tile behind):
    // wait until all three CTAs are past their main loops. (All warps of the CTA join the cluster barrier; see below.)
    cluster_sync_all();
    if (kt_end > kt_begin) {
      uint8_t* srow = smem + (size_t)row * WF_STAGE_PITCH;
      for (int c0 = 0; c0 < 192; c0 += 32) {
        uint32_t v[32];
        __syncwarp();
        tmem_ld_32x32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)c0, v);
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 8; ++j) *(uint4*)(srow + (c0 + 4 * j) * 4) = make_uint4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
      }
      fence_proxy_async();
      asm volatile("cp.reduce.async.bulk.global.shared::cta.bulk_group.add.f32 [%0], [%1], %2;"
                   ::"l"(dst), "r"(smem_u32(srow)), "r"(768) : "memory");
      asm volatile("cp.async.bulk.commit_group;" ::: "memory");
      // (the staging rows have been read; the additions themselves complete before the grid does, which is what orders them
      //  before the fold launch)
      asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
    }
  }
  if (warp < 2) {   // the producer / MMA warps' half of the barrier the epilogue warps wait on above
    __syncwarp();
    cluster_sync_all();
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    __syncwarp();
    tmem_dealloc(tmem_base, 256);
  }
  cluster_sync_all();   // nobody leaves while a peer's commit may still be arriving on its barriers
}

// ------------------------------------------------------------------------------------------------------------------
// Split-K reduction of the flat wgrad partials into the fp32 OIHW gradient (accumulating: grad += sum over z).
// One CTA per (output channel co, 64-channel input chunk): its 192 threads sum, for each of the three filter rows, the
// z partial rows (coalesced 768-byte reads), permute through shared memory into the OIHW order ci*9 + tap and add the
// 576 contiguous floats to the gradient. All flat convolutions of a backward part are served by one launch (job table).
// ------------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(768) wgrad_reduce_kernel(const __grid_constant__ WgradReduceJobs jobs, float* __restrict__ grads) {
  __shared__ __align__(16) float s_part[8][576];   // [K-slice chain][576] when a CTA serves one output channel, [output channel][576] when it serves several
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
    // One accumulator tile per (co block, ci chunk, filter row) (the wgrad kernel's K slices add into it): a CTA serves 8 output
    // channels = 24 rows of 768 B in, 8 x 2304 contiguous bytes of the OIHW gradient out. Every load (accumulators AND the
    // gradient it is added to) is a 16-byte vector issued before the first use; the zeroing stores come after ALL loads - placed
    // between them they alias the next load for the compiler and the kernel degenerated into 24 dependent round trips per
    // thread (0.8 TB/s, 340 us of kernel time at the end of every step).
    const int co0 = cgrp * 8;
    const int tid = threadIdx.x;                 // 192 threads: 6 vectors each of the 24 x 48 accumulator vectors
    float4 v[6], gv[6];
    float4* srcp[6];
    float4* gp[6];
#pragma unroll
    for (int i = 0; i < 6; ++i) {
      const int q = tid + i * 192;               // 0..1151
      const int rt = q / 48, c4 = q - rt * 48;   // (row r, filter row tg) = rt / 3, rt % 3 ; vector inside the 192-wide row
      const int r = rt / 3, tg = rt - r * 3;
      const int co = co0 + r, cob = co >> 7, row = co & 127;
      const int tile = (cob * jb.ci_chunks + cic) * 3 + tg;
      srcp[i] = (float4*)(jb.scratch + ((size_t)tile * 128 + row) * 192) + c4;
      v[i] = __ldcg(srcp[i]);
      // output vector q: row r2 = q / 144, 16-byte vector o4 of its 576 floats
      const int r2 = q / 144, o4 = q - r2 * 144;
      gp[i] = (float4*)(grads + jb.grad_off + ((size_t)(co0 + r2) * jb.cin + cic * 64) * 9) + o4;
      gv[i] = __ldcg(gp[i]);
    }
    const float4 zero4 = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int i = 0; i < 6; ++i) {
      const int q = tid + i * 192;
      const int rt = q / 48, c4 = q - rt * 48;
      const int r = rt / 3, tg = rt - r * 3;
      const int t0 = c4 * 4;                     // column: tap-in-row (t >> 6), input channel (t & 63); 4 consecutive channels
      float* d = &s_part[r][(t0 & 63) * 9 + tg * 3 + (t0 >> 6)];
      d[0] = v[i].x; d[9] = v[i].y; d[18] = v[i].z; d[27] = v[i].w;
      if (jb.zero_src) __stcg(srcp[i], zero4);   // the accumulator is ready for the next backward
    }
    __syncthreads();
#pragma unroll
    for (int i = 0; i < 6; ++i) {
      const int q = tid + i * 192;
      const int r2 = q / 144, o4 = q - r2 * 144;
      const float4 a = *(const float4*)&s_part[r2][o4 * 4];
      gv[i].x += a.x; gv[i].y += a.y; gv[i].z += a.z; gv[i].w += a.w;
      *gp[i] = gv[i];
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
