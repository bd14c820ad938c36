// Implicit-GEMM convolution on tcgen05 tensor cores (fprop and dgrad share this kernel).
//
//   D[pixel, n] = sum over taps t, channel chunks c of  A_t[pixel (+tap offset), c*64..] * W[slab(t)][n, c*64..]^T
//
// A (activations, NHWC bf16) is never materialised as an im2col matrix: for every filter tap the TMA unit
// loads a 4-D box (64 channels x BW x BH x BN pixels) whose origin is shifted by the tap offset; out-of-range
// coordinates (the conv's zero padding, the ragged last tile, images past the batch) are zero-filled by TMA.
// The box lands in shared memory as 128-byte rows with the 128B swizzle = the canonical K-major UMMA operand.
// Strided convs use the tensor map's element strides. Accumulators live in TMEM (2 x 256 columns, so the
// epilogue of tile i overlaps the MMAs of tile i+1). Persistent CTAs, warp-specialised:
//   warp 0 : TMA producer      warp 1 : MMA issuer (one lane)      warps 2..5 : epilogue (TMEM -> regs -> smem -> HBM)
#pragma once
#include "common.cuh"
#include "conv_params.h"

namespace cilrs {

CILRS_DEVINL void conv_gemm_body(const ConvGemmParams& p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);  // keeps the shared address space (LDS/STS)

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int stage_bytes = CG_A_BYTES + p.block_n * 128;
  uint8_t* staging = smem + (size_t)p.num_stages * stage_bytes;
  long long* s_rowoff = (long long*)(staging + CG_STAGING_BYTES);  // [128]
  float* s_stat = (float*)(s_rowoff + CG_BLOCK_M);                 // [4][64][2]
  float* s_acc = s_stat + 4 * 64 * 2;                              // [2][512] per-CTA running (sum, sumsq) per channel
  uint64_t* bars = (uint64_t*)(s_acc + 2 * 512);
  uint64_t* full_bar = bars;
  uint64_t* empty_bar = bars + CG_MAX_STAGES;
  uint64_t* tfull_bar = bars + 2 * CG_MAX_STAGES;
  uint64_t* tempty_bar = tfull_bar + 2;
  uint32_t* tmem_slot = (uint32_t*)(tempty_bar + 2);

  pdl_launch_dependents();
  // rows of an A tile that no TMA box ever writes (valid_rows..127) must read as 0: zero them once per stage
  {
    const int vr = p.BW * p.BH * p.BN;
    const int n16 = (CG_BLOCK_M - vr) * 8;  // 16-byte units per stage
    const uint4 z = make_uint4(0, 0, 0, 0);
    for (int st = 0; st < p.num_stages; ++st) {
      uint4* q = (uint4*)(smem + (size_t)st * stage_bytes + (size_t)vr * 128);
      for (int i = threadIdx.x; i < n16; i += CG_THREADS) q[i] = z;
    }
    fence_proxy_async();
  }
  if (warp == 0 && lane == 0) {
    for (int i = 0; i < 4; ++i) tma_prefetch_desc(&p.tmA[i]);
    tma_prefetch_desc(&p.tmB[0]);
    tma_prefetch_desc(&p.tmB[1]);
    for (int i = 0; i < p.num_stages; ++i) {
      mbar_init(&full_bar[i], 1);
      mbar_init(&empty_bar[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tfull_bar[i], 1);
      mbar_init(&tempty_bar[i], 4);  // one arrive per epilogue warp
    }
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
  pdl_wait();

  const int m_tiles = p.tiles_w * p.tiles_h * p.tiles_n;
  const int total_tiles = m_tiles * p.n_blocks;
  const int k_iters = p.num_taps * p.chunks;
  const int valid_rows = p.BW * p.BH * p.BN;
  const uint32_t tx_bytes = (uint32_t)(valid_rows * 128 + p.block_n * 128);

  if (warp == 0) {
    // ================= TMA producer =================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
        const int m_tile = tile / p.n_blocks;
        const int n_blk = tile - m_tile * p.n_blocks;
        const int tw = m_tile % p.tiles_w;
        const int th = (m_tile / p.tiles_w) % p.tiles_h;
        const int tn = m_tile / (p.tiles_w * p.tiles_h);
        const int w0 = tw * p.BW * p.in_sw, h0 = th * p.BH * p.in_sh, n0 = tn * p.BN;
        for (int t = 0; t < p.num_taps; ++t) {
          const CUtensorMap* ma = &p.tmA[p.tap_a[t]];
          const CUtensorMap* mb = &p.tmB[p.tap_b[t]];
          const int cw = w0 + p.tap_dw[t], ch = h0 + p.tap_dh[t];
          const int brow = p.tap_slab[t] * p.slab_rows + n_blk * p.block_n;
          for (int c = 0; c < p.chunks; ++c) {
            mbar_wait(&empty_bar[stage], phase ^ 1);
            uint8_t* sa = smem + (size_t)stage * stage_bytes;
            mbar_arrive_expect_tx(&full_bar[stage], tx_bytes);
            tma_load_4d(ma, &full_bar[stage], sa, c * 64, cw, ch, n0);
            tma_load_2d(mb, &full_bar[stage], sa + CG_A_BYTES, c * 64, brow);
            if (++stage == p.num_stages) { stage = 0; phase ^= 1; }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ================= MMA issuer =================
    // whole warp in uniform control flow (descriptors stay in uniform registers), one elected lane issues; the CTA owns
    // all 512 TMEM columns, so the allocation starts at column 0 (see conv_flat.cuh)
    if (tmem_base != 0) __trap();
    const bool leader = elect_one();
    const uint32_t idesc = umma_idesc_bf16(CG_BLOCK_M, p.block_n, 0, 0);
    const uint64_t descA0 = umma_desc_sw128(smem_u32(smem), 16, 1024);
    const uint64_t descB0 = umma_desc_sw128(smem_u32(smem) + CG_A_BYTES, 16, 1024);
    const uint32_t stage_units = (uint32_t)(stage_bytes >> 4);
    int stage = 0;
    uint32_t phase = 0;
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
      mbar_wait(&tempty_bar[acc], acc_phase ^ 1);
      tc_fence_after();
      const uint32_t d_tmem = (uint32_t)acc * 256;
      for (int k = 0; k < k_iters; ++k) {
        mbar_wait(&full_bar[stage], phase);
        tc_fence_after();
        const uint64_t da = descA0 + (uint64_t)((uint32_t)stage * stage_units);
        const uint64_t db = descB0 + (uint64_t)((uint32_t)stage * stage_units);
        if (leader) {
#pragma unroll
          for (int kk = 0; kk < 4; ++kk)  // 4 x (K = 16 bf16 = 32 bytes = 2 sixteen-byte units) inside the 128-byte swizzled row
            umma_bf16(d_tmem, da + kk * 2, db + kk * 2, idesc, (k | kk) != 0 ? 1u : 0u);
          umma_commit(&empty_bar[stage]);
        }
        __syncwarp();
        if (++stage == p.num_stages) { stage = 0; phase ^= 1; }
      }
      if (leader) umma_commit(&tfull_bar[acc]);
      __syncwarp();
      acc ^= 1;
      if (acc == 0) acc_phase ^= 1;
    }
  } else {
    // ================= epilogue (4 warps = 128 TMEM lanes) =================
    const int q = warp & 3;            // TMEM lane quadrant this warp may access
    const int row = q * 32 + lane;     // accumulator row = pixel within the tile
    const int etid = threadIdx.x - 64; // 0..127
    int acc = 0;
    uint32_t acc_phase = 0;
    if (p.flags & CG_STATS) {
      for (int i = etid; i < 2 * 512; i += 128) s_acc[i] = 0.f;  // thread etid owns channels == etid (mod 64) from here on
    }
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
      const int m_tile = tile / p.n_blocks;
      const int n_blk = tile - m_tile * p.n_blocks;
      const int tw = m_tile % p.tiles_w;
      const int th = (m_tile / p.tiles_w) % p.tiles_h;
      const int tn = m_tile / (p.tiles_w * p.tiles_h);
      {
        // element offset of the pixel held by row `etid`
        const int r = etid;
        const int bw = r % p.BW, bh = (r / p.BW) % p.BH, bn = r / (p.BW * p.BH);
        const int w = tw * p.BW + bw, h = th * p.BH + bh, n = tn * p.BN + bn;
        long long off = -1;
        if (r < valid_rows && w < p.ow && h < p.oh && n < p.n_img)
          off = p.out_off + (long long)n * p.out_sn + (long long)h * p.out_sh + (long long)w * p.out_sw;
        bar_sync_named(2, 128);  // previous tile's readers of s_rowoff are done
        s_rowoff[r] = off;
        bar_sync_named(2, 128);
      }
      const long long my_off = s_rowoff[row];
      mbar_wait(&tfull_bar[acc], acc_phase);
      tc_fence_after();
      const int n_chunks = p.block_n >> 6;
      for (int chunk = 0; chunk < n_chunks; ++chunk) {
        uint32_t v[64];
        const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * 256 + chunk * 64);
        tmem_ld_32x32(taddr, v);
        tmem_ld_32x32(taddr + 32, v + 32);
        tmem_ld_wait();
        if (chunk == n_chunks - 1) {
          // all of this warp's reads of the accumulator are complete: hand it back to the MMA warp
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&tempty_bar[acc]);
        }
        const int n_base = n_blk * p.block_n + chunk * 64;
        if (p.flags & CG_SCALE_BIAS) {
#pragma unroll
          for (int j = 0; j < 64; ++j)
            v[j] = __float_as_uint(fmaf(__uint_as_float(v[j]), __ldg(p.scale + n_base + j), __ldg(p.bias + n_base + j)));
        }
        if ((p.flags & CG_RESIDUAL) && my_off >= 0) {
          const uint4* rp = (const uint4*)(p.residual + my_off + n_base);
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const uint4 rv = __ldg(rp + j);
            const uint32_t rr[4] = {rv.x, rv.y, rv.z, rv.w};
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              v[j * 8 + e * 2] = __float_as_uint(__uint_as_float(v[j * 8 + e * 2]) + bf16lo(rr[e]));
              v[j * 8 + e * 2 + 1] = __float_as_uint(__uint_as_float(v[j * 8 + e * 2 + 1]) + bf16hi(rr[e]));
            }
          }
        }
        if (p.flags & CG_RELU) {
#pragma unroll
          for (int j = 0; j < 64; ++j) v[j] = __float_as_uint(fmaxf(__uint_as_float(v[j]), 0.f));
        }
        if (my_off < 0) {
          // a row outside the output image (ragged tile): it may hold a partial sum of real inputs, and must
          // not leak into the statistics
#pragma unroll
          for (int j = 0; j < 64; ++j) v[j] = 0u;
        }
        __syncwarp();
        // stage the bf16 chunk: row-major 128-byte rows, 16-byte units XOR-swizzled by (row & 7)
        uint8_t* sbuf = staging + (chunk & 1) * CG_A_BYTES;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          uint4 o;
          o.x = pack_bf16x2(__uint_as_float(v[j * 8 + 0]), __uint_as_float(v[j * 8 + 1]));
          o.y = pack_bf16x2(__uint_as_float(v[j * 8 + 2]), __uint_as_float(v[j * 8 + 3]));
          o.z = pack_bf16x2(__uint_as_float(v[j * 8 + 4]), __uint_as_float(v[j * 8 + 5]));
          o.w = pack_bf16x2(__uint_as_float(v[j * 8 + 6]), __uint_as_float(v[j * 8 + 7]));
          *(uint4*)(sbuf + row * 128 + ((j ^ (row & 7)) << 4)) = o;
        }
        bar_sync_named(1, 128);
        // coalesced write-out: 8 consecutive threads cover one 128-byte row
#pragma unroll
        for (int it = 0; it < 8; ++it) {
          const int u = it * 128 + etid;
          const int r = u >> 3, c16 = u & 7;
          const long long off = s_rowoff[r];
          if (off >= 0) {
            const uint4 o = *(const uint4*)(sbuf + r * 128 + ((c16 ^ (r & 7)) << 4));
            *(uint4*)(p.out + off + n_base + c16 * 8) = o;
          }
        }
        if (p.flags & CG_STATS) {
          // per-channel sum and sum of squares over the tile's rows (invalid rows hold exact zeros)
          const int cp = etid & 31, g = etid >> 5;
          float s0 = 0.f, s1 = 0.f, q0 = 0.f, q1 = 0.f;
          const int c16 = cp >> 2, sub = (cp & 3) * 4;
#pragma unroll 8
          for (int rr = 0; rr < 32; ++rr) {
            const int r = g * 32 + rr;
            const uint32_t u = *(const uint32_t*)(sbuf + r * 128 + ((c16 ^ (r & 7)) << 4) + sub);
            const float a = bf16lo(u), b = bf16hi(u);
            s0 += a; s1 += b;
            q0 = fmaf(a, a, q0); q1 = fmaf(b, b, q1);
          }
          float* st = s_stat + (g * 64 + cp * 2) * 2;
          st[0] = s0; st[1] = q0; st[2] = s1; st[3] = q1;
          bar_sync_named(3, 128);
          if (etid < 64) {
            float s = 0.f, qq = 0.f;
#pragma unroll
            for (int gg = 0; gg < 4; ++gg) {
              s += s_stat[(gg * 64 + etid) * 2];
              qq += s_stat[(gg * 64 + etid) * 2 + 1];
            }
            s_acc[n_base + etid] += s;          // tiles are visited in a fixed order: deterministic
            s_acc[512 + n_base + etid] += qq;
          }
        }
      }
      acc ^= 1;
      if (acc == 0) acc_phase ^= 1;
    }
    if (p.flags & CG_STATS) {
      // one partial per CTA: stats[blockIdx.x][2][n_total]
      bar_sync_named(3, 128);
      float* gp = p.stats + (size_t)blockIdx.x * 2 * p.n_total;
      for (int i = etid; i < p.n_total; i += 128) {
        gp[i] = s_acc[i];
        gp[p.n_total + i] = s_acc[512 + i];
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

__global__ void __launch_bounds__(CG_THREADS, 1) conv_gemm_kernel(const __grid_constant__ ConvGemmParams p) { conv_gemm_body(p); }

// The four output-parity problems of a stride-2 dgrad in ONE launch: blockIdx.y selects the problem, every problem gets
// gridDim.x persistent CTAs. (Four separate launches of ~40..100 tiles each left most of the SMs idle four times in a row.)
struct ConvGemmParams4 {
  ConvGemmParams p[4];
};
__global__ void __launch_bounds__(CG_THREADS, 1) conv_gemm_multi_kernel(const __grid_constant__ ConvGemmParams4 pp) {
  conv_gemm_body(pp.p[blockIdx.y]);
}

}  // namespace cilrs
