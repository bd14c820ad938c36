// 3x3 stride-1 convolution (fprop and dgrad) on the padded-flat activation layout (conv_params.h: PadGeom).
//
//   D[f, n] = sum over taps t, channel chunks c of   A[f + shift_t, c*64..] * W[slab_t][n, c*64..]^T
//
// Per (tile, channel chunk) ONE slab of (mt*128 + 2*halo) flat pixels x 64 channels is loaded by TMA; the nine taps are
// row-shifted UMMA descriptors into that slab, so the activations cross L2->SMEM ~1.3x instead of 9x. `mt` 128-row
// sub-tiles share every weight tile (weights cross 1/mt as often); with 64x64 layers the whole 3x3 filter stays
// resident in shared memory. Accumulators live in TMEM (acc_sets x mt x block_n columns) so the epilogue of one tile
// overlaps the MMAs of the next. Persistent CTAs, warp-specialised:
//   warp 0 : TMA producer     warp 1 : MMA issuer (one lane)     warps 2..9 : two epilogue groups (TMEM -> regs -> smem -> HBM)
// Fused epilogues: folded BN / residual / ReLU (inference), BN batch statistics + finalize (training forward),
// ReLU mask + BatchNorm-backward reductions + finalize (dgrad). Reductions are deterministic: per-CTA partials in a
// fixed order, folded by the last CTA to finish.
#pragma once
#include "common.cuh"
#include "conv_params.h"

namespace cilrs {

// optional event trace of CTA 0 (tools/trace_flat.py; compiled in only with -DCF_TRACE)
#ifdef CF_TRACE
__device__ unsigned long long g_cf_trace[3][2048];
__device__ int g_cf_trace_n[3];
#define CF_EVENT(role, code)                                                              \
  do {                                                                                    \
    if (blockIdx.x == 0 && cf_idx[role] < 2048)                                           \
      g_cf_trace[role][cf_idx[role]++] = ((unsigned long long)clock64() << 16) | (unsigned long long)((code) & 0xFFFF); \
  } while (0)
#else
#define CF_EVENT(role, code) do { } while (0)
#endif

CILRS_DEVINL void cf_unpack8(const uint4 u, float* f) {
  f[0] = bf16lo(u.x); f[1] = bf16hi(u.x); f[2] = bf16lo(u.y); f[3] = bf16hi(u.y);
  f[4] = bf16lo(u.z); f[5] = bf16hi(u.z); f[6] = bf16lo(u.w); f[7] = bf16hi(u.w);
}

__global__ void __launch_bounds__(CF_THREADS, 1) conv_flat_kernel(const __grid_constant__ FlatConvParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  const int a_stage_bytes = p.a_boxes * p.a_box_rows * 128;
  const int b_stage_bytes = p.block_n * 128;
  uint8_t* sA = smem;
  uint8_t* sB = sA + (size_t)p.a_stages * a_stage_bytes;
  uint8_t* staging = sB + (size_t)p.b_stages * b_stage_bytes;
  float* s_stat = (float*)(staging + CF_STAGING_BYTES);  // [2 groups][4 warps][64 ch][3]
  float* s_acc = s_stat + 2 * 4 * 64 * 3;                // [2 groups][3][n_total]
  uint64_t* bars = (uint64_t*)(s_acc + 2 * 3 * p.n_total);
  uint64_t* full_a = bars;
  uint64_t* empty_a = full_a + CF_MAX_A_STAGES;
  uint64_t* full_b = empty_a + CF_MAX_A_STAGES;
  uint64_t* empty_b = full_b + CF_MAX_B_STAGES;
  uint64_t* tfull = empty_b + CF_MAX_B_STAGES;
  uint64_t* tempty = tfull + CF_MAX_ACC;
  uint32_t* tmem_slot = (uint32_t*)(tempty + CF_MAX_ACC);
  uint32_t* s_flag = tmem_slot + 1;
#ifdef CF_TRACE
  uint32_t* cf_idx = s_flag + 1;  // event counters of the three traced roles (shared memory: cheap to bump)
  if (threadIdx.x < 3) cf_idx[threadIdx.x] = 0;
#endif

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&p.tmA);
    tma_prefetch_desc(&p.tmB);
    for (int i = 0; i < p.a_stages; ++i) { mbar_init(&full_a[i], 1); mbar_init(&empty_a[i], 1); }
    for (int i = 0; i < p.b_stages; ++i) { mbar_init(&full_b[i], 1); mbar_init(&empty_b[i], 1); }
    for (int i = 0; i < p.acc_sets; ++i) { mbar_init(&tfull[i], 1); mbar_init(&tempty[i], 8); }  // 8 epilogue warps
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

  const int total_tiles = p.m_tiles * p.n_blocks;
  const int acc_stride = p.mt * p.block_n;  // TMEM columns per accumulator set

  if (warp == 0) {
    // ================= TMA producer =================
    if (lane == 0) {
      int as = 0, bs = 0;
      uint32_t aph = 0, bph = 0;
      bool first = true;
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
        const int m_tile = tile / p.n_blocks;
        const int n_blk = tile - m_tile * p.n_blocks;
        const int row0 = m_tile * p.mt * 128;
        for (int c = 0; c < p.chunks; ++c) {
          mbar_wait(&empty_a[as], aph ^ 1);
          CF_EVENT(0, 0x100 + c);
          mbar_arrive_expect_tx(&full_a[as], (uint32_t)a_stage_bytes);
          uint8_t* dst = sA + (size_t)as * a_stage_bytes;
          for (int bx = 0; bx < p.a_boxes; ++bx)
            tma_load_2d(&p.tmA, &full_a[as], dst + (size_t)bx * p.a_box_rows * 128, c * 64, row0 - p.halo + bx * p.a_box_rows);
          if (++as == p.a_stages) { as = 0; aph ^= 1; }
          for (int t = 0; t < p.num_taps; ++t) {
            if (!p.b_resident || first) {
              if (!p.b_resident) mbar_wait(&empty_b[bs], bph ^ 1);
              CF_EVENT(0, 0x200 + t);
              mbar_arrive_expect_tx(&full_b[bs], (uint32_t)b_stage_bytes);
              tma_load_2d(&p.tmB, &full_b[bs], sB + (size_t)bs * b_stage_bytes, c * 64, p.tap_slab[t] * p.n_total + n_blk * p.block_n);
            }
            if (++bs == p.b_stages) { bs = 0; bph ^= 1; }
          }
        }
        first = false;
      }
    }
  } else if (warp == 1) {
    // ================= MMA issuer =================
    // The whole warp runs the loop so every address / descriptor stays in uniform registers and an MMA costs ~8
    // instructions; one elected lane issues. (With a lone lane inside a divergent branch each tcgen05.mma cost ~90
    // clocks of instruction issue - more than the 32..64 clocks the tensor core needs for N = 64..128;
    // tools/umma_rate_test2.cu.) The CTA owns all 512 TMEM columns, so the allocation starts at column 0.
    if (tmem_base != 0) __trap();
    const bool leader = elect_one();
    const uint32_t idesc = umma_idesc_bf16(128, p.block_n, 0, 0);
    const uint64_t descA0 = umma_desc_sw128(smem_u32(sA), 16, 1024);
    const uint64_t descB0 = umma_desc_sw128(smem_u32(sB), 16, 1024);
    const uint32_t a_stage_units = (uint32_t)(a_stage_bytes >> 4), b_stage_units = (uint32_t)(b_stage_bytes >> 4);
    int as = 0, bs = 0, acc = 0;
    uint32_t aph = 0, bph = 0, accph = 0;
    bool first = true;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
      mbar_wait(&tempty[acc], accph ^ 1);
      tc_fence_after();
      if (leader) CF_EVENT(1, 0x300);
      const uint32_t d_base = (uint32_t)(acc * acc_stride);
      for (int c = 0; c < p.chunks; ++c) {
        mbar_wait(&full_a[as], aph);
        tc_fence_after();
        if (leader) CF_EVENT(1, 0x100 + c);
        const uint64_t da_stage = descA0 + (uint64_t)((uint32_t)as * a_stage_units);
        for (int t = 0; t < p.num_taps; ++t) {
          if (!p.b_resident || first) {
            mbar_wait(&full_b[bs], bph);
            tc_fence_after();
          }
          if (leader) CF_EVENT(1, 0x200 + t);
          const uint64_t db = descB0 + (uint64_t)((uint32_t)bs * b_stage_units);
          const uint64_t da_tap = da_stage + (uint64_t)((uint32_t)(p.halo + p.tap_shift[t]) * 8u);  // 128-byte rows, in 16-byte units
          if (leader) {
            for (int m = 0; m < p.mt; ++m) {
              const uint64_t da = da_tap + (uint64_t)((uint32_t)m * 1024u);
              const uint32_t d = d_base + (uint32_t)(m * p.block_n);
#pragma unroll
              for (int kk = 0; kk < 4; ++kk) umma_bf16(d, da + kk * 2, db + kk * 2, idesc, (c | t | kk) != 0 ? 1u : 0u);
            }
            if (!p.b_resident) umma_commit(&empty_b[bs]);
          }
          __syncwarp();
          if (++bs == p.b_stages) { bs = 0; bph ^= 1; }
        }
        if (leader) umma_commit(&empty_a[as]);
        __syncwarp();
        if (++as == p.a_stages) { as = 0; aph ^= 1; }
      }
      if (leader) {
        umma_commit(&tfull[acc]);
        CF_EVENT(1, 0x400);
      }
      __syncwarp();
      if (++acc == p.acc_sets) { acc = 0; accph ^= 1; }
      first = false;
    }
  } else {
    // ================= epilogue: 2 groups x 4 warps; a group handles every other 128 x 64 unit =================
    const int ew = warp - 2;
    const int grp = ew >> 2;
    const int q = warp & 3;                 // TMEM lane quadrant this warp may access
    const int row = q * 32 + lane;          // accumulator row
    const int etid = (ew & 3) * 32 + lane;  // 0..127 inside the group
    const int bar_id = 1 + grp;
    uint8_t* sbuf = staging + grp * (CF_STAGING_BYTES / 2);
    float* g_stat = s_stat + grp * (4 * 64 * 3);
    float* g_acc = s_acc + grp * (3 * p.n_total);
    const bool do_stats = (p.flags & (CF_STATS | CF_BNBWD)) != 0;
    const bool bwd = (p.flags & CF_BNBWD) != 0;
    const bool bwd2 = (p.flags & CF_BNBWD2) != 0;
    const int nq = bwd2 ? 3 : 2;
    if (do_stats) {
      for (int i = etid; i < 3 * p.n_total; i += 128) g_acc[i] = 0.f;
    }
    int acc = 0;
    uint32_t accph = 0;
    uint32_t uc = 0;
    const int n_chunks = p.block_n >> 6;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
      const int m_tile = tile / p.n_blocks;
      const int n_blk = tile - m_tile * p.n_blocks;
      const int row0 = m_tile * p.mt * 128;
      mbar_wait(&tfull[acc], accph);
      tc_fence_after();
      if (ew == 0 && lane == 0) CF_EVENT(2, 0x500);
      for (int m = 0; m < p.mt; ++m) {
        const int f = row0 + m * 128 + row;
        bool valid = f < p.total_rows;
        if (valid) {
          const unsigned int uf = (unsigned int)f;
          const unsigned int wq = uf / (unsigned int)p.g.Wp;
          const unsigned int w = uf - wq * (unsigned int)p.g.Wp;
          const unsigned int h = wq % (unsigned int)p.g.Hp;
          valid = (w < (unsigned int)p.g.W) && (h < (unsigned int)p.g.H);
        }
        for (int chunk = 0; chunk < n_chunks; ++chunk) {
          if (((uc++) & 1u) != (uint32_t)grp) continue;
          const int n_base = n_blk * p.block_n + chunk * 64;
          const long long goff = (long long)f * p.n_total + n_base;
          uint32_t v[64];
          const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * acc_stride + m * p.block_n + chunk * 64);
          tmem_ld_32x32(taddr, v);
          tmem_ld_32x32(taddr + 32, v + 32);
          tmem_ld_wait();
          if (ew == 0 && lane == 0) CF_EVENT(2, 0x601);
          if (p.flags & CF_SCALE_BIAS) {
#pragma unroll
            for (int j = 0; j < 64; ++j)
              v[j] = __float_as_uint(fmaf(__uint_as_float(v[j]), __ldg(p.scale + n_base + j), __ldg(p.bias + n_base + j)));
          }
          if ((p.flags & CF_RESIDUAL) && valid) {
            const uint4* rp = (const uint4*)(p.residual + goff);
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              float r[8];
              cf_unpack8(__ldg(rp + j), r);
#pragma unroll
              for (int e = 0; e < 8; ++e) v[j * 8 + e] = __float_as_uint(__uint_as_float(v[j * 8 + e]) + r[e]);
            }
          }
          if ((p.flags & CF_MASK) && valid) {
            const uint4* mp = (const uint4*)(p.mask + goff);
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              float r[8];
              cf_unpack8(__ldg(mp + j), r);
#pragma unroll
              for (int e = 0; e < 8; ++e)
                if (!(r[e] > 0.f)) v[j * 8 + e] = 0u;
            }
          }
          if (p.flags & CF_RELU) {
#pragma unroll
            for (int j = 0; j < 64; ++j) v[j] = __float_as_uint(fmaxf(__uint_as_float(v[j]), 0.f));
          }
          if (!valid) {
            // padding pixels of the layout (and rows past the end): they stay exact zeros in memory and in the statistics
#pragma unroll
            for (int j = 0; j < 64; ++j) v[j] = 0u;
          }
          if (ew == 0 && lane == 0) CF_EVENT(2, 0x602);
          bar_sync_named(bar_id, 128);  // the group's previous unit no longer reads the staging buffer
          if (ew == 0 && lane == 0) CF_EVENT(2, 0x603);
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            uint4 o;
            o.x = pack_bf16x2(__uint_as_float(v[j * 8 + 0]), __uint_as_float(v[j * 8 + 1]));
            o.y = pack_bf16x2(__uint_as_float(v[j * 8 + 2]), __uint_as_float(v[j * 8 + 3]));
            o.z = pack_bf16x2(__uint_as_float(v[j * 8 + 4]), __uint_as_float(v[j * 8 + 5]));
            o.w = pack_bf16x2(__uint_as_float(v[j * 8 + 6]), __uint_as_float(v[j * 8 + 7]));
            *(uint4*)(sbuf + row * 128 + ((j ^ (row & 7)) << 4)) = o;
          }
          bar_sync_named(bar_id, 128);
          if (ew == 0 && lane == 0) CF_EVENT(2, 0x604);
          // coalesced write-out: 8 consecutive threads cover one 128-byte row
          const int fb = row0 + m * 128;
#pragma unroll
          for (int it = 0; it < 8; ++it) {
            const int u = it * 128 + etid;
            const int r = u >> 3, c16 = u & 7;
            if (fb + r < p.total_rows) {
              const uint4 o = *(const uint4*)(sbuf + r * 128 + ((c16 ^ (r & 7)) << 4));
              *(uint4*)(p.out + (long long)(fb + r) * p.n_total + n_base + c16 * 8) = o;
            }
          }
          if (ew == 0 && lane == 0) CF_EVENT(2, 0x605);
          if (do_stats) {
            // column phase: thread (cp, rg) owns channel pair cp over the 32 rows of row group rg
            const int cp = etid & 31, rg = etid >> 5;
            const int c16 = cp >> 2, sub = (cp & 3) * 4;
            float s0 = 0.f, s1 = 0.f, q0 = 0.f, q1 = 0.f, t0 = 0.f, t1 = 0.f;
            if (!bwd) {
#pragma unroll 8
              for (int rr = 0; rr < 32; ++rr) {
                const int r = rg * 32 + rr;
                const uint32_t u = *(const uint32_t*)(sbuf + r * 128 + ((c16 ^ (r & 7)) << 4) + sub);
                const float a = bf16lo(u), b = bf16hi(u);
                s0 += a; s1 += b;
                q0 = fmaf(a, a, q0); q1 = fmaf(b, b, q1);
              }
            } else {
              const int ch = n_base + cp * 2;
              const float m1a = __ldg(p.stat1 + 2 * p.n_total + ch), m1b = __ldg(p.stat1 + 2 * p.n_total + ch + 1);
              const float r1a = __ldg(p.stat1 + 3 * p.n_total + ch), r1b = __ldg(p.stat1 + 3 * p.n_total + ch + 1);
              float m2a = 0.f, m2b = 0.f, r2a = 0.f, r2b = 0.f;
              if (bwd2) {
                m2a = __ldg(p.stat2 + 2 * p.n_total + ch); m2b = __ldg(p.stat2 + 2 * p.n_total + ch + 1);
                r2a = __ldg(p.stat2 + 3 * p.n_total + ch); r2b = __ldg(p.stat2 + 3 * p.n_total + ch + 1);
              }
              const int last = p.total_rows - 1;
#pragma unroll 8
              for (int rr = 0; rr < 32; ++rr) {
                const int r = rg * 32 + rr;
                const uint32_t u = *(const uint32_t*)(sbuf + r * 128 + ((c16 ^ (r & 7)) << 4) + sub);
                const float a = bf16lo(u), b = bf16hi(u);
                const long long fo = (long long)min(fb + r, last) * p.n_total + ch;  // rows past the end hold dz = 0
                const uint32_t yu = __ldg((const unsigned int*)(p.y1 + fo));
                s0 += a; s1 += b;
                q0 = fmaf(a, (bf16lo(yu) - m1a) * r1a, q0);
                q1 = fmaf(b, (bf16hi(yu) - m1b) * r1b, q1);
                if (bwd2) {
                  const uint32_t y2u = __ldg((const unsigned int*)(p.y2 + fo));
                  t0 = fmaf(a, (bf16lo(y2u) - m2a) * r2a, t0);
                  t1 = fmaf(b, (bf16hi(y2u) - m2b) * r2b, t1);
                }
              }
            }
            if (ew == 0 && lane == 0) CF_EVENT(2, 0x606);
            float* st = g_stat + (rg * 64 + cp * 2) * 3;
            st[0] = s0; st[1] = q0; st[2] = t0;
            st[3] = s1; st[4] = q1; st[5] = t1;
            bar_sync_named(bar_id, 128);
            if (etid < 64) {
              float s = 0.f, qq = 0.f, tt = 0.f;
#pragma unroll
              for (int gg = 0; gg < 4; ++gg) {
                s += g_stat[(gg * 64 + etid) * 3];
                qq += g_stat[(gg * 64 + etid) * 3 + 1];
                tt += g_stat[(gg * 64 + etid) * 3 + 2];
              }
              g_acc[n_base + etid] += s;  // units are visited in a fixed order: deterministic
              g_acc[p.n_total + n_base + etid] += qq;
              g_acc[2 * p.n_total + n_base + etid] += tt;
            }
          }
        }
      }
      if (ew == 0 && lane == 0) CF_EVENT(2, 0x600);
      // all of this warp's reads of the accumulator set are complete: hand it back to the MMA warp
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tempty[acc]);
      if (++acc == p.acc_sets) { acc = 0; accph ^= 1; }
    }

    if (do_stats) {
      // ---- per-CTA partial, then the last CTA to finish folds all partials and finalizes ----
      const int tid = ew * 32 + lane;  // 0..255
      if (tid == 0) CF_EVENT(2, 0x700);
      bar_sync_named(3, 256);
      float* gp = p.partials + (size_t)blockIdx.x * nq * p.n_total;
      for (int i = tid; i < nq * p.n_total; i += 256) gp[i] = s_acc[i] + s_acc[3 * p.n_total + i];
      __threadfence();
      bar_sync_named(3, 256);
      if (tid == 0) {
        const unsigned int done = atomicAdd(p.counter, 1u);
        *s_flag = (done == gridDim.x - 1) ? 1u : 0u;
      }
      bar_sync_named(3, 256);
      if (tid == 0) CF_EVENT(2, 0x701);
      if (*s_flag) {
        __threadfence();
        double* s_fold = (double*)staging;  // [slices][nq][n_total] doubles <= 24 KB
        const int quads = p.n_total >> 2;
        const int slices = 256 / quads;     // n_total 64 -> 16, 512 -> 2
        {
          const int qd = tid % quads, sl = tid / quads;
          double a[3][4];
#pragma unroll
          for (int k = 0; k < 3; ++k)
#pragma unroll
            for (int e = 0; e < 4; ++e) a[k][e] = 0.0;
          if (sl < slices) {
#pragma unroll 4
            for (unsigned int b = sl; b < gridDim.x; b += slices) {
              const float* bp = p.partials + (size_t)b * nq * p.n_total + qd * 4;
#pragma unroll
              for (int k = 0; k < 3; ++k) {
                if (k < nq) {
                  const float4 x = __ldcg((const float4*)(bp + k * p.n_total));
                  a[k][0] += (double)x.x; a[k][1] += (double)x.y; a[k][2] += (double)x.z; a[k][3] += (double)x.w;
                }
              }
            }
#pragma unroll
            for (int k = 0; k < 3; ++k)
#pragma unroll
              for (int e = 0; e < 4; ++e) s_fold[((size_t)sl * 3 + k) * p.n_total + qd * 4 + e] = a[k][e];
          }
        }
        bar_sync_named(3, 256);
        if (tid == 0) CF_EVENT(2, 0x702);
        for (int c = tid; c < p.n_total; c += 256) {
          double S[3] = {0.0, 0.0, 0.0};
          for (int sl = 0; sl < slices; ++sl)
#pragma unroll
            for (int k = 0; k < 3; ++k) S[k] += s_fold[((size_t)sl * 3 + k) * p.n_total + c];
          if (!bwd) {
            const double mean_d = S[0] / p.count;
            double var_d = S[1] / p.count - mean_d * mean_d;
            if (var_d < 0.0) var_d = 0.0;
            const float mean = (float)mean_d, var = (float)var_d;
            if (p.update_running) {
              const double unbiased = p.count > 1.0 ? var_d * p.count / (p.count - 1.0) : var_d;
              p.running_mean[c] = (1.f - p.momentum) * p.running_mean[c] + p.momentum * mean;
              p.running_var[c] = (1.f - p.momentum) * p.running_var[c] + p.momentum * (float)unbiased;
            }
            const float rstd = 1.0f / sqrtf(var + p.eps);
            const float sc = p.gamma[c] * rstd;
            p.vec[c] = sc;
            p.vec[p.n_total + c] = p.beta[c] - mean * sc;
            p.vec[2 * p.n_total + c] = mean;
            p.vec[3 * p.n_total + c] = rstd;
          } else {
            const float bs = (float)S[0], bd1 = (float)S[1];
            p.bred1[c] = bs;
            p.bred1[p.n_total + c] = bd1;
            if (p.dgamma1) p.dgamma1[c] += bd1;
            if (p.dbeta1) p.dbeta1[c] += bs;
            if (bwd2) {
              const float bd2 = (float)S[2];
              p.bred2[c] = bs;
              p.bred2[p.n_total + c] = bd2;
              if (p.dgamma2) p.dgamma2[c] += bd2;
              if (p.dbeta2) p.dbeta2[c] += bs;
            }
          }
        }
        if (tid == 0) {
          CF_EVENT(2, 0x703);
          *p.counter = 0u;  // ready for the next launch / graph replay
          if (!bwd && p.update_running && p.nbt) *p.nbt += 1;
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
#ifdef CF_TRACE
  if (blockIdx.x == 0 && threadIdx.x < 3) g_cf_trace_n[threadIdx.x] = (int)cf_idx[threadIdx.x];
#endif
  if (warp == 1) {
    __syncwarp();
    tmem_dealloc(tmem_base, 512);
  }
}

}  // namespace cilrs
