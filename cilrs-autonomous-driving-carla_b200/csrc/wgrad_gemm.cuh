// Weight-gradient GEMM on tcgen05 tensor cores.
//
//   dW[co, tap, ci] = sum over output pixels p of  dY[p, co] * X[p (*stride) + tap offset, ci]
//
// The contraction runs over pixels, so both operands are "MN-major" for UMMA: a TMA box of 128 pixels x 64
// channels lands in shared memory as 128-byte rows (one pixel per row, 128B swizzle) and is consumed as a
// K(=pixel)-by-64 operand slab. One CTA owns a [128 co] x [g taps x 64 ci] block of dW and a slice of the pixel
// tiles (split-K over pixels); its accumulator stays in TMEM for the whole slice and is added to the fp32
// OIHW gradient with red.global.add at the end.
//   warp 0 : TMA producer      warp 1 : MMA issuer      warps 2..5 : epilogue
#pragma once
#include "common.cuh"
#include "conv_params.h"

namespace cilrs {

__global__ void __launch_bounds__(WG_THREADS, 1) wgrad_gemm_kernel(const __grid_constant__ WgradParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);  // keeps the shared address space (LDS/STS)
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  const int slabs = 2 + p.g;  // the A side always spans two 64-channel slabs (the second stays zero if m_halves == 1)
  const int stage_bytes = slabs * WG_SLAB;
  uint64_t* bars = (uint64_t*)(smem + (size_t)p.num_stages * stage_bytes);
  uint64_t* full_bar = bars;
  uint64_t* empty_bar = bars + WG_MAX_STAGES;
  uint64_t* done_bar = bars + 2 * WG_MAX_STAGES;
  uint32_t* tmem_slot = (uint32_t*)(done_bar + 1);

  pdl_launch_dependents();
  // pixel rows no TMA box ever writes (valid_rows..127 of every slab) and the unused second dY slab must read as 0
  {
    const int vr = p.BW * p.BH * p.BN;
    const uint4 z = make_uint4(0, 0, 0, 0);
    for (int st = 0; st < p.num_stages; ++st) {
      for (int sl = 0; sl < slabs; ++sl) {
        const bool whole = (sl == 1 && p.m_halves == 1);
        const int first = whole ? 0 : vr;
        uint4* q = (uint4*)(smem + (size_t)st * stage_bytes + (size_t)sl * WG_SLAB + (size_t)first * 128);
        const int n16 = (128 - first) * 8;
        for (int i = threadIdx.x; i < n16; i += WG_THREADS) q[i] = z;
      }
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
  const int tg = wi % p.tap_groups; wi /= p.tap_groups;
  const int cic = wi % p.ci_chunks; wi /= p.ci_chunks;
  const int cob = wi;
  const int pix_tiles = p.tiles_w * p.tiles_h * p.tiles_n;
  const int per = (pix_tiles + p.split_z - 1) / p.split_z;
  const int pt_begin = z * per;
  const int pt_end = min(pix_tiles, pt_begin + per);
  const int valid_rows = p.BW * p.BH * p.BN;
  const uint32_t tx_bytes = (uint32_t)((p.m_halves + p.g) * valid_rows * 128);

  if (warp == 0) {
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int pt = pt_begin; pt < pt_end; ++pt) {
        const int tw = pt % p.tiles_w;
        const int th = (pt / p.tiles_w) % p.tiles_h;
        const int tn = pt / (p.tiles_w * p.tiles_h);
        const int w0 = tw * p.BW, h0 = th * p.BH, n0 = tn * p.BN;
        mbar_wait(&empty_bar[stage], phase ^ 1);
        uint8_t* s = smem + (size_t)stage * stage_bytes;
        mbar_arrive_expect_tx(&full_bar[stage], tx_bytes);
        for (int mh = 0; mh < p.m_halves; ++mh)
          tma_load_4d(&p.tmDY, &full_bar[stage], s + mh * WG_SLAB, cob * 128 + mh * 64, w0, h0, n0);
        for (int j = 0; j < p.g; ++j) {
          const int t = tg * p.g + j;
          tma_load_4d(&p.tmX, &full_bar[stage], s + (2 + j) * WG_SLAB, cic * 64, w0 * p.in_sw + p.tap_dw[t],
                      h0 * p.in_sh + p.tap_dh[t], n0);
        }
        if (++stage == p.num_stages) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1) {
    // whole warp in uniform control flow, one elected lane issues (see conv_flat.cuh); first TMEM allocation of the SM
    if (tmem_base != 0) __trap();
    const bool leader = elect_one();
    const uint32_t idesc = umma_idesc_bf16(128, 64 * p.g, 1, 1);
    const uint64_t descA0 = umma_desc_sw128(smem_u32(smem), WG_SLAB, 1024);
    const uint64_t descB0 = umma_desc_sw128(smem_u32(smem) + 2 * WG_SLAB, WG_SLAB, 1024);
    const uint32_t stage_units = (uint32_t)(stage_bytes >> 4);
    int stage = 0;
    uint32_t phase = 0;
    uint32_t first = 1;
    for (int pt = pt_begin; pt < pt_end; ++pt) {
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
    const int co = cob * 128 + row;
    if (pt_end > pt_begin) {
      mbar_wait(done_bar, 0);
      tc_fence_after();
      const int ncols = 64 * p.g;
      for (int c0 = 0; c0 < ncols; c0 += 32) {
        uint32_t v[32];
        __syncwarp();
        tmem_ld_32x32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)c0, v);
        tmem_ld_wait();
        if (row < p.m_halves * 64 && co < p.cout) {
          const int j = c0 >> 6;        // tap within the group
          const int t = tg * p.g + j;
          float* gp = p.grad + (size_t)co * p.co_stride;
#pragma unroll
          for (int e = 0; e < 32; ++e) {
            const int cc = (c0 & 63) + e;  // column inside the 64-wide slab
            int off;
            if (p.col_mode == WG_COL_REGULAR) {
              off = (cic * 64 + cc) * p.ci_stride + p.tap_id[t];
            } else {
              // conv1 space-to-depth packing: cc = s'*16 + dy*8 + dx*4 + c, tap t = r'
              const int sp = cc >> 4, dy = (cc >> 3) & 1, dx = (cc >> 2) & 1, c = cc & 3;
              const int ky = 2 * t + dy, kx = 2 * sp + dx;
              off = (ky < 7 && kx < 7 && c < 3) ? c * 49 + ky * 7 + kx : -1;
            }
            if (off >= 0) atomicAdd(gp + off, __uint_as_float(v[e]));
          }
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    __syncwarp();
    tmem_dealloc(tmem_base, 256);
  }
}

}  // namespace cilrs
