// 3x3 stride-1 convolution (fprop and dgrad) on the padded-flat activation layout (conv_params.h: PadGeom).
//
//   D[f, n] = sum over taps t, channel chunks c of   A[f + shift_t, c*64..] * W[slab_t][n, c*64..]^T
//
// Per (tile, channel chunk) ONE slab of (mt*128 + 2*halo) flat pixels x 64 channels is loaded by TMA; the nine taps are
// row-shifted UMMA descriptors into that slab, so the activations cross L2->SMEM ~1.3x instead of 9x. `mt` 128-row
// sub-tiles share every weight tile (weights cross 1/mt as often); with 64x64 layers the whole 3x3 filter stays
// resident in shared memory. Accumulators live in TMEM (acc_sets x mt x block_n columns) so the epilogue of one tile
// overlaps the MMAs of the next. Persistent CTAs, warp-specialised:
//   warp 0 : TMA producer     warp 1 : MMA issuer     warps 4..11 : two epilogue groups (TMEM -> registers -> HBM)
// Fused epilogues: folded BN / residual / ReLU (inference), BN batch statistics (training forward), ReLU mask +
// BatchNorm-backward reductions (dgrad). The per-channel sums go to fp64 global accumulators (order-independent); they are
// finalized either by the last CTA to finish or, with CF_DEFER, by the elementwise kernel that consumes them.
#pragma once
#include "common.cuh"
#include "pair.cuh"
#include "conv_params.h"
#include "bn_math.cuh"

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

// 256-bit global accesses (one full 32-byte sector per thread)
CILRS_DEVINL void ldg256(const void* p, uint32_t* r) {
  asm volatile("ld.global.nc.v8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]) : "l"(p));
}
CILRS_DEVINL void stg256(void* p, const uint32_t* r) {
  asm volatile("st.global.v8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};"
               ::"l"(p), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]) : "memory");
}
// one row (64 bf16 = 128 bytes) of a padded-flat tensor -> 32 packed registers
CILRS_DEVINL void load_row64(const __nv_bfloat16* p, uint32_t* r) {
#pragma unroll
  for (int j = 0; j < 4; ++j) ldg256(p + j * 16, r + j * 8);
}

// PAIR: the grid is made of clusters of two CTAs that share every tcgen05.mma (cta_group::2, M = 256): CTA `rank` of a pair
// owns the rows [(2*m_tile + rank) * MT*128, +MT*128) of a pair tile - its own activation slab, its own 128-row accumulators
// in its own TMEM, its own epilogue - and HALF of every weight tile (block_n/2 rows), so each weight byte crosses L2->SMEM
// once per pair and the tensor core fetches 4096 + 16 N instead of 4096 + 32 N bytes of shared memory per instruction and
// CTA (the single-CTA form is bound by exactly that: conv_flat.cu). The leader (rank 0) issues all MMAs; both CTAs' TMA
// loads complete on the leader's `full` barriers, the leader's commits arrive on the `empty` / `tfull` barriers of both.
template <int MT, bool PAIR>
__global__ void __launch_bounds__(CF_THREADS, 1) conv_flat_kernel(const __grid_constant__ FlatConvParams p) {
  extern __shared__ uint8_t smem_raw[];
  // aligned by pointer arithmetic on the __shared__ array (not through an integer cast) so that the compiler keeps the
  // shared address space and emits LDS/STS instead of generic loads/stores
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  const uint32_t rank = PAIR ? cluster_ctarank() : 0u;
  const int pair_id = PAIR ? (int)(blockIdx.x >> 1) : (int)blockIdx.x;   // tile iterator of this CTA (pair)
  const int n_pairs = PAIR ? (int)(gridDim.x >> 1) : (int)gridDim.x;
  const int a_stage_bytes = p.a_boxes * p.a_box_rows * 128;
  const int b_rows = PAIR ? (p.block_n >> 1) : p.block_n;       // weight-tile rows held by this CTA
  const int b_tap_bytes = b_rows * 128;
  const int b_stage_bytes = p.tap_group * b_tap_bytes;  // one stage = the weight tiles of `tap_group` consecutive taps
  uint8_t* sA = smem;
  uint8_t* sB = sA + (size_t)p.a_stages * a_stage_bytes;
  uint8_t* s_out = sB + (size_t)p.b_stages * b_stage_bytes;            // [8 epilogue warps][32 rows][128 B] output staging
  uint8_t* s_yin = s_out + CF_STAGING_BYTES;                           // [8][32][128 B] y tiles of the BN-backward dot (CF_BNBWD only)
  float* s_wacc = (float*)(s_yin + ((p.flags & CF_BNBWD) ? CF_STAGING_BYTES : 0));  // [8 epilogue warps][3][block_n] running statistics
  uint64_t* bars = (uint64_t*)(s_wacc + 8 * 3 * p.block_n);
  uint64_t* full_a = bars;
  uint64_t* empty_a = full_a + CF_MAX_A_STAGES;
  uint64_t* full_b = empty_a + CF_MAX_A_STAGES;
  uint64_t* empty_b = full_b + CF_MAX_B_STAGES;
  uint64_t* tfull = empty_b + CF_MAX_B_STAGES;
  uint64_t* tempty = tfull + CF_MAX_ACC;
  uint64_t* ebar = tempty + CF_MAX_ACC;                                // [8 epilogue warps][2]: residual tile, y tile landed
  uint32_t* tmem_slot = (uint32_t*)(ebar + 16);
  uint32_t* s_flag = tmem_slot + 1;
#ifdef CF_TRACE
  uint32_t* cf_idx = s_flag + 1;  // event counters of the three traced roles (shared memory: cheap to bump)
  if (threadIdx.x < 3) cf_idx[threadIdx.x] = 0;
#endif

  pdl_launch_dependents();
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&p.tmA);
    tma_prefetch_desc(&p.tmB);
    tma_prefetch_desc(&p.tmOut);
    if (p.operand_maps) { tma_prefetch_desc(&p.tmRes); tma_prefetch_desc(&p.tmY1); }
    if (p.flags & CF_FUSE) tma_prefetch_desc(&p.tmOut2);
    for (int i = 0; i < p.a_stages; ++i) { mbar_init(&full_a[i], 1); mbar_init(&empty_a[i], 1); }
    for (int i = 0; i < p.b_stages; ++i) { mbar_init(&full_b[i], 1); mbar_init(&empty_b[i], 1); }
    for (int i = 0; i < p.acc_sets; ++i) { mbar_init(&tfull[i], 1); mbar_init(&tempty[i], PAIR ? 16 : 8); }  // 8 epilogue warps (per CTA)
    for (int i = 0; i < 16; ++i) mbar_init(&ebar[i], 1);
    fence_barrier_init();
  }
  if (warp == 1) {
    if (PAIR) { tmem_alloc_pair(tmem_slot, 512); tmem_relinquish_pair(); }
    else { tmem_alloc(tmem_slot, 512); tmem_relinquish(); }
  }
  tc_fence_before();
  if (PAIR) cluster_sync_all();  // the peer's barriers are initialised before anything arrives on them
  else __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait();  // everything above is local to the CTA; from here on the predecessor's results are read

  const int total_tiles = p.m_tiles * p.n_blocks;
  const int acc_stride = MT * p.block_n;  // TMEM columns per accumulator set
  const int n_groups = (p.num_taps + p.tap_group - 1) / p.tap_group;

  // 384 threads x 168 registers at launch: warpgroup 0 keeps 64 per thread, the two epilogue warpgroups grow to 216 so that
  // a unit's operand rows (residual, y) can be prefetched into registers without spilling
  if (warp < 4) {
  setmaxnreg_dec<64>();
  if (warp == 0) {
    // ================= TMA producer =================
    if (lane == 0) {
      int as = 0, bs = 0;
      uint32_t aph = 0, bph = 0;
      bool first = true;
      for (int tile = pair_id; tile < total_tiles; tile += n_pairs) {
        const int m_tile = tile / p.n_blocks;
        const int n_blk = tile - m_tile * p.n_blocks;
        const int row0 = (m_tile * (PAIR ? 2 : 1) + (int)rank) * MT * 128;
        for (int c = 0; c < p.chunks; ++c) {
          mbar_wait(&empty_a[as], aph ^ 1);
          CF_EVENT(0, 0x100 + c);
          // (pair: the leader's barrier counts the bytes of both CTAs; the peer only issues its loads)
          if (rank == 0) mbar_arrive_expect_tx(&full_a[as], (uint32_t)a_stage_bytes * (PAIR ? 2u : 1u));
          uint8_t* dst = sA + (size_t)as * a_stage_bytes;
          for (int bx = 0; bx < p.a_boxes; ++bx) {
            if (PAIR) tma_load_2d_pair(&p.tmA, &full_a[as], dst + (size_t)bx * p.a_box_rows * 128, c * 64, row0 - p.halo + bx * p.a_box_rows);
            else tma_load_2d(&p.tmA, &full_a[as], dst + (size_t)bx * p.a_box_rows * 128, c * 64, row0 - p.halo + bx * p.a_box_rows);
          }
          if (++as == p.a_stages) { as = 0; aph ^= 1; }
          for (int gi = 0; gi < n_groups; ++gi) {
            const int t0 = gi * p.tap_group;
            const int cnt = min(p.tap_group, p.num_taps - t0);
            if (!p.b_resident || first) {
              if (!p.b_resident) mbar_wait(&empty_b[bs], bph ^ 1);
              CF_EVENT(0, 0x200 + gi);
              if (rank == 0) mbar_arrive_expect_tx(&full_b[bs], (uint32_t)(cnt * b_tap_bytes) * (PAIR ? 2u : 1u));
              for (int j = 0; j < cnt; ++j) {
                uint8_t* dstb = sB + (size_t)bs * b_stage_bytes + (size_t)j * b_tap_bytes;
                const int wrow = p.tap_slab[t0 + j] * p.n_total + n_blk * p.block_n + (int)rank * b_rows;
                if (PAIR) tma_load_2d_pair(&p.tmB, &full_b[bs], dstb, c * 64, wrow);
                else tma_load_2d(&p.tmB, &full_b[bs], dstb, c * 64, wrow);
              }
            }
            if (++bs == p.b_stages) { bs = 0; bph ^= 1; }
          }
        }
        first = false;
      }
    }
  } else if (warp == 1 && rank == 0) {
    // ================= MMA issuer (pair: the leader CTA only) =================
    // The whole warp runs the loop so every address / descriptor stays in uniform registers and an MMA costs ~8
    // instructions; one elected lane issues. (With a lone lane inside a divergent branch each tcgen05.mma cost ~90
    // clocks of instruction issue - more than the 32..64 clocks the tensor core needs for N = 64..128;
    // tools/umma_rate_test2.cu.) The CTA owns all 512 TMEM columns, so the allocation starts at column 0.
    if (tmem_base != 0) __trap();
    const bool leader = elect_one();
    const uint32_t idesc = umma_idesc_bf16(PAIR ? 256 : 128, p.block_n, 0, 0);
    const uint64_t descA0 = umma_desc_sw128(smem_u32(sA), 16, 1024);
    const uint64_t descB0 = umma_desc_sw128(smem_u32(sB), 16, 1024);
    const uint32_t a_stage_units = (uint32_t)(a_stage_bytes >> 4), b_stage_units = (uint32_t)(b_stage_bytes >> 4);
    const uint32_t b_tap_units = (uint32_t)(b_tap_bytes >> 4);
    int as = 0, bs = 0, acc = 0;
    uint32_t aph = 0, bph = 0, accph = 0;
    bool first = true;
    for (int tile = pair_id; tile < total_tiles; tile += n_pairs) {
      if (PAIR) mbar_wait_cluster(&tempty[acc], accph ^ 1);  // (the peer's epilogue warps arrive remotely)
      else mbar_wait(&tempty[acc], accph ^ 1);
      tc_fence_after();
      if (leader) CF_EVENT(1, 0x300);
      const uint32_t d_base = (uint32_t)(acc * acc_stride);
      for (int c = 0; c < p.chunks; ++c) {
        mbar_wait(&full_a[as], aph);
        tc_fence_after();
        if (leader) CF_EVENT(1, 0x100 + c);
        const uint64_t da_stage = descA0 + (uint64_t)((uint32_t)as * a_stage_units);
        for (int gi = 0; gi < n_groups; ++gi) {
          const int t0 = gi * p.tap_group;
          const int cnt = min(p.tap_group, p.num_taps - t0);
          if (!p.b_resident || first) {
            mbar_wait(&full_b[bs], bph);
            tc_fence_after();
          }
          if (leader) CF_EVENT(1, 0x200 + gi);
          const uint64_t db_stage = descB0 + (uint64_t)((uint32_t)bs * b_stage_units);
          if (leader) {
            for (int j = 0; j < cnt; ++j) {
              const uint64_t db = db_stage + (uint64_t)((uint32_t)j * b_tap_units);
              const uint64_t da_tap = da_stage + (uint64_t)((uint32_t)(p.halo + p.tap_shift[t0 + j]) * 8u);  // 128-byte rows, in 16-byte units
#pragma unroll
              for (int m = 0; m < MT; ++m) {
                const uint64_t da = da_tap + (uint64_t)((uint32_t)m * 1024u);
                const uint32_t d = d_base + (uint32_t)(m * p.block_n);
#pragma unroll
                for (int kk = 0; kk < 4; ++kk) {
                  if (PAIR) umma_bf16_pair(d, da + kk * 2, db + kk * 2, idesc, (c | (t0 + j) | kk) != 0 ? 1u : 0u);
                  else umma_bf16(d, da + kk * 2, db + kk * 2, idesc, (c | (t0 + j) | kk) != 0 ? 1u : 0u);
                }
              }
            }
            if (!p.b_resident) { if (PAIR) umma_commit_pair(&empty_b[bs], 3); else umma_commit(&empty_b[bs]); }
          }
          __syncwarp();
          if (++bs == p.b_stages) { bs = 0; bph ^= 1; }
        }
        if (leader) { if (PAIR) umma_commit_pair(&empty_a[as], 3); else umma_commit(&empty_a[as]); }
        __syncwarp();
        if (++as == p.a_stages) { as = 0; aph ^= 1; }
      }
      if (leader) {
        if (PAIR) umma_commit_pair(&tfull[acc], 3); else umma_commit(&tfull[acc]);
        CF_EVENT(1, 0x400);
      }
      __syncwarp();
      if (++acc == p.acc_sets) { acc = 0; accph ^= 1; }
      first = false;
    }
  }
  } else {
    setmaxnreg_inc<216>();
    // ================= epilogue: 2 groups x 4 warps; a group handles every other 128 x 64 unit =================
    // TMEM -> registers -> (scale/bias, residual, ReLU / ReLU-mask) -> bf16 -> this warp's swizzled staging tile -> one TMA
    // store. The operand tiles (residual, y of the BatchNorm-backward dot) come in by TMA as well: per-lane row loads make
    // every instruction touch 32 different lines. The per-channel sums are a COLUMN pass over the staged tiles: lane l reads
    // the 4 bytes of columns 2l, 2l+1 of each of the 32 rows (conflict-free, 32 shared-memory cycles) - a third of the
    // instructions of a shuffle butterfly over the registers, whose 62 shuffles per quantity also each take a cycle of the
    // shared-memory pipe the tensor core's operand fetch is saturating.
    const int ew = warp - 4;
    const int grp = ew >> 2;
    const int q = warp & 3;                 // TMEM lane quadrant this warp may access
    const int row = q * 32 + lane;          // accumulator row
    const bool do_stats = (p.flags & (CF_STATS | CF_BNBWD)) != 0;
    const bool bwd = (p.flags & CF_BNBWD) != 0;
    const bool bwd2 = (p.flags & CF_BNBWD2) != 0;
    const int nq = bwd2 ? 3 : 2;
    const bool res_tma = (p.flags & CF_RESIDUAL) && p.operand_maps;
    float* w_acc = s_wacc + ew * (3 * p.block_n);  // this warp's running sums [3][block_n]
    uint8_t* w_out = s_out + ew * (32 * 128);      // this warp's output staging (1024-byte aligned, 128B-swizzled rows)
    uint8_t* w_y = s_yin + ew * (32 * 128);        // this warp's y tile
    uint64_t* bar_res = &ebar[ew * 2];
    uint64_t* bar_y = &ebar[ew * 2 + 1];
    uint32_t ph_res = 0, ph_y = 0;
    // byte offset of (row r, columns 2*lane, 2*lane+1) inside a swizzled tile, without the row term
    const uint32_t col_chunk = (uint32_t)(lane >> 2), col_in = (uint32_t)(lane & 3) * 4u;
    if (do_stats) {
      for (int i = lane; i < 3 * p.block_n; i += 32) w_acc[i] = 0.f;
      __syncwarp();
    }
    int acc = 0;
    uint32_t accph = 0;
    uint32_t uc = 0;
    const int n_chunks = p.block_n >> 6;
    for (int tile = pair_id; tile < total_tiles; tile += n_pairs) {
      const int m_tile = tile / p.n_blocks;
      const int n_blk = tile - m_tile * p.n_blocks;
      const int row0 = (m_tile * (PAIR ? 2 : 1) + (int)rank) * MT * 128;
      mbar_wait(&tfull[acc], accph);
      tc_fence_after();
      if (ew == 0 && lane == 0) CF_EVENT(2, 0x500);
#pragma unroll 1
      for (int m = 0; m < MT; ++m) {
        const int f = row0 + m * 128 + row;
        const bool in_range = f < p.total_rows;
        bool valid = in_range;
        if (valid) {
          const unsigned int uf = (unsigned int)f;
          const unsigned int wq = uf / (unsigned int)p.g.Wp;
          const unsigned int w = uf - wq * (unsigned int)p.g.Wp;
          const unsigned int h = wq % (unsigned int)p.g.Hp;
          valid = (w < (unsigned int)p.g.W) && (h < (unsigned int)p.g.H);
        }
#pragma unroll 1
        for (int chunk = 0; chunk < n_chunks; ++chunk) {
          if (((uc++) & 1u) != (uint32_t)grp) continue;
          const int n_base = n_blk * p.block_n + chunk * 64;
          const long long goff = (long long)f * p.n_total + n_base;
          const int row_first = row0 + m * 128 + q * 32;   // first row of this warp's 32 x 64 tile
          // ---- operand tiles of this unit: issued first so that they land while the accumulator is read ----
          // (the residual tile lands in the output staging tile: the previous unit's store must have read it; the previous
          //  unit's column pass over both tiles is complete on every lane)
          tma_store_wait_read();
          __syncwarp();
          if (lane == 0) {
            if (res_tma) {
              mbar_arrive_expect_tx(bar_res, 32 * 128);
              tma_load_2d(&p.tmRes, bar_res, w_out, n_base, row_first);
            }
            if (bwd) {
              mbar_arrive_expect_tx(bar_y, 32 * 128);
              tma_load_2d(&p.tmY1, bar_y, w_y, n_base, row_first);
            }
          }
          uint2 mbits = make_uint2(0xffffffffu, 0xffffffffu);
          const bool use_bits = (p.flags & CF_MASK) && p.mask_bits != nullptr;
          if (use_bits && valid) mbits = __ldg(reinterpret_cast<const uint2*>(p.mask_bits + (goff >> 3)));
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
          if (p.flags & CF_RESIDUAL) {
            uint32_t rres[32];
            if (res_tma) {
              mbar_wait(bar_res, ph_res);
              ph_res ^= 1;
#pragma unroll
              for (int j = 0; j < 8; ++j) {
                const uint4 t = *(const uint4*)(w_out + lane * 128 + ((j ^ (lane & 7)) << 4));
                rres[4 * j] = t.x; rres[4 * j + 1] = t.y; rres[4 * j + 2] = t.z; rres[4 * j + 3] = t.w;
              }
            } else if (valid) {
              load_row64(p.residual + goff, rres);
            }
            if (res_tma || valid) {
#pragma unroll
              for (int j = 0; j < 32; ++j) {
                v[2 * j] = __float_as_uint(__uint_as_float(v[2 * j]) + bf16lo(rres[j]));
                v[2 * j + 1] = __float_as_uint(__uint_as_float(v[2 * j + 1]) + bf16hi(rres[j]));
              }
            }
          }
          if (use_bits) {
#pragma unroll
            for (int j = 0; j < 64; ++j)
              if (!(((j < 32 ? mbits.x : mbits.y) >> (j & 31)) & 1u)) v[j] = 0u;
          } else if ((p.flags & CF_MASK) && valid) {
            uint32_t r[32];
            load_row64(p.mask + goff, r);
#pragma unroll
            for (int j = 0; j < 32; ++j) {
              if (!(bf16lo(r[j]) > 0.f)) v[2 * j] = 0u;
              if (!(bf16hi(r[j]) > 0.f)) v[2 * j + 1] = 0u;
            }
          }
          if (p.flags & CF_RELU) {
#pragma unroll
            for (int j = 0; j < 64; ++j) v[j] = __float_as_uint(fmaxf(__uint_as_float(v[j]), 0.f));
          }
          uint32_t u[32];
#pragma unroll
          for (int j = 0; j < 32; ++j)
            u[j] = valid ? pack_bf16x2(__uint_as_float(v[2 * j]), __uint_as_float(v[2 * j + 1])) : 0u;  // padding pixels stay exact zeros
          // Output: the warp's 32 rows x 128 B go through the swizzled staging tile and ONE TMA store (full 128-byte lines).
          // Per-thread 32-byte stores of a row each cost the LSU 32 sector requests per instruction: 1.5 k clocks per unit.
          __syncwarp();  // every lane has read its residual row out of the staging tile
#pragma unroll
          for (int j = 0; j < 8; ++j)
            *(uint4*)(w_out + lane * 128 + ((j ^ (lane & 7)) << 4)) = make_uint4(u[4 * j], u[4 * j + 1], u[4 * j + 2], u[4 * j + 3]);
          fence_proxy_async();
          __syncwarp();
          if (lane == 0 && !(p.flags & CF_NO_STORE)) {   // (CF_NO_STORE: only the sums of this tile are needed; pass 2 recomputes it)
            tma_store_2d(&p.tmOut, w_out, n_base, row_first);
            tma_store_commit();
          }
          if (ew == 0 && lane == 0) CF_EVENT(2, 0x605);
          if (do_stats) {
            // column pass over the staged (bf16-rounded, zero on padding rows) tile: lane l owns columns 2l, 2l+1
            float* wa = w_acc + chunk * 64 + 2 * lane;
            float s0x = 0.f, s0y = 0.f, s1x = 0.f, s1y = 0.f;
            if (!bwd) {
#pragma unroll
              for (int r = 0; r < 32; ++r) {
                const uint32_t t = *(const uint32_t*)(w_out + r * 128 + ((col_chunk ^ (uint32_t)(r & 7)) << 4) + col_in);
                const float a = bf16lo(t), b = bf16hi(t);
                s0x += a; s0y += b;
                s1x = fmaf(a, a, s1x); s1y = fmaf(b, b, s1y);
              }
            } else {
              // sum dz and sum dz * y (raw); the finalize turns the latter into sum dz * xhat = rstd * (sum dz*y - mean * sum dz)
              mbar_wait(bar_y, ph_y);
              ph_y ^= 1;
#pragma unroll
              for (int r = 0; r < 32; ++r) {
                const uint32_t off = (uint32_t)(r * 128) + ((col_chunk ^ (uint32_t)(r & 7)) << 4) + col_in;
                const uint32_t t = *(const uint32_t*)(w_out + off);
                const uint32_t y = *(const uint32_t*)(w_y + off);
                const float a = bf16lo(t), b = bf16hi(t);
                s0x += a; s0y += b;
                s1x = fmaf(a, bf16lo(y), s1x); s1y = fmaf(b, bf16hi(y), s1y);
              }
            }
            wa[0] += s0x; wa[1] += s0y;
            wa[p.block_n] += s1x; wa[p.block_n + 1] += s1y;
            if (bwd2) {
              // second BatchNorm fed by the same gradient (downsample branch): its y tile replaces the first one
              __syncwarp();
              if (lane == 0) {
                mbar_arrive_expect_tx(bar_y, 32 * 128);
                tma_load_2d(&p.tmY2, bar_y, w_y, n_base, row_first);
              }
              mbar_wait(bar_y, ph_y);
              ph_y ^= 1;
              float s2x = 0.f, s2y = 0.f;
#pragma unroll
              for (int r = 0; r < 32; ++r) {
                const uint32_t off = (uint32_t)(r * 128) + ((col_chunk ^ (uint32_t)(r & 7)) << 4) + col_in;
                const uint32_t t = *(const uint32_t*)(w_out + off);
                const uint32_t y = *(const uint32_t*)(w_y + off);
                s2x = fmaf(bf16lo(t), bf16lo(y), s2x); s2y = fmaf(bf16hi(t), bf16hi(y), s2y);
              }
              wa[2 * p.block_n] += s2x; wa[2 * p.block_n + 1] += s2y;
            }
          }
          if (ew == 0 && lane == 0) CF_EVENT(2, 0x606);
        }
      }
      if (ew == 0 && lane == 0) CF_EVENT(2, 0x600);
      // all of this warp's reads of the accumulator set are complete: hand it back to the MMA warp
      tc_fence_before();
      __syncwarp();
      if (lane == 0) { if (PAIR) mbar_arrive_remote(&tempty[acc], 0); else mbar_arrive(&tempty[acc]); }
      if (++acc == p.acc_sets) { acc = 0; accph ^= 1; }
    }

    const bool fuse = (p.flags & CF_FUSE) != 0;
    // The staging tiles have been read by every store; the writes themselves only have to be complete (~1.3 k clocks later in a
    // CTA trace) when the second pass of CF_FUSE reads them back - otherwise the end of the grid orders them before any consumer
    if (fuse) tma_store_wait_all(); else tma_store_wait_read();

    if (do_stats) {
      // ---- per-CTA partial (the eight warps' sums in a fixed order), then the last CTA to finish folds all partials ----
      const int tid = ew * 32 + lane;  // 0..255
      if (tid == 0) CF_EVENT(2, 0x700);
      bar_sync_named(3, 256);
      // this CTA's sums (the eight warps in a fixed order) go into the global per-channel accumulators [3][n_total] with
      // red.global.add; the last CTA to arrive reads them, finalizes and re-zeroes them for the next launch
      const int nb = pair_id % p.n_blocks;  // the number of CTAs (pairs) is a multiple of n_blocks: one channel block per CTA
      for (int i = tid; i < nq * p.block_n; i += 256) {
        float t = 0.f;
#pragma unroll
        for (int wv = 0; wv < 8; ++wv) t += s_wacc[wv * 3 * p.block_n + i];
        const int k = i / p.block_n, c = i - k * p.block_n;
        // fp64 accumulation of fp32 partials is exact unless their exponents span more than 2^22, so the result does not
        // depend on the order in which the CTAs arrive (reproducible statistics without a serial fold)
        atomicAdd(p.partials + k * p.n_total + nb * p.block_n + c, (double)t);
      }
      bool last_cta = false;
      if (!(p.flags & CF_DEFER)) {  // (deferred: the kernel boundary orders the sums before the consumer's reads)
        bar_sync_named(3, 256);
        if (tid == 0) {
          // release: the barrier ordered every thread's atomics before this fence (cumulativity); acquire after the count
          __threadfence();
          const unsigned int done = atomicAdd(p.counter, 1u);
          const bool last = (done == gridDim.x - 1);
          if (last) __threadfence();
          *s_flag = last ? 1u : 0u;
        }
        bar_sync_named(3, 256);
        last_cta = *s_flag != 0u;
      }
      if (tid == 0) CF_EVENT(2, 0x701);
      if (last_cta) {
        // per-channel inputs of the finalize (loads issued together, one L2 round trip)
        float pre[2][4];
#pragma unroll
        for (int it = 0; it < 2; ++it) {
          const int c = tid + it * 256;
          if (c < p.n_total) {
            if (!bwd) {
              pre[it][0] = __ldg(p.gamma + c); pre[it][1] = __ldg(p.beta + c);
              pre[it][2] = p.update_running ? p.running_mean[c] : 0.f; pre[it][3] = p.update_running ? p.running_var[c] : 0.f;
            } else {
              pre[it][0] = __ldg(p.stat1 + 2 * p.n_total + c); pre[it][1] = __ldg(p.stat1 + 3 * p.n_total + c);
              pre[it][2] = bwd2 ? __ldg(p.stat2 + 2 * p.n_total + c) : 0.f; pre[it][3] = bwd2 ? __ldg(p.stat2 + 3 * p.n_total + c) : 0.f;
            }
          }
        }
        const double inv_count = 1.0 / p.count;
#pragma unroll
        for (int it = 0; it < 2; ++it) {
          const int c = tid + it * 256;
          if (c >= p.n_total) break;
          double S[3] = {0.0, 0.0, 0.0};
#pragma unroll
          for (int k = 0; k < 3; ++k) {
            if (k < nq) {
              S[k] = __ldcg(p.partials + k * p.n_total + c);
              p.partials[k * p.n_total + c] = 0.0;
            }
          }
          if (!bwd) {
            const double mean_d = S[0] * inv_count;
            double var_d = S[1] * inv_count - mean_d * mean_d;
            if (var_d < 0.0) var_d = 0.0;
            const float mean = (float)mean_d, var = (float)var_d;
            if (p.update_running) {
              const float unbiased = p.count > 1.0 ? (float)(var_d * (p.count / (p.count - 1.0))) : var;
              p.running_mean[c] = (1.f - p.momentum) * pre[it][2] + p.momentum * mean;
              p.running_var[c] = (1.f - p.momentum) * pre[it][3] + p.momentum * unbiased;
            }
            const float rstd = 1.0f / sqrtf(var + p.eps);
            const float sc = pre[it][0] * rstd;
            p.vec[c] = sc;
            p.vec[p.n_total + c] = pre[it][1] - mean * sc;
            p.vec[2 * p.n_total + c] = mean;
            p.vec[3 * p.n_total + c] = rstd;
          } else {
            // sum dz * xhat = rstd * (sum dz * y - mean * sum dz)
            const float bs = (float)S[0];
            const float bd1 = (float)((double)pre[it][1] * (S[1] - (double)pre[it][0] * S[0]));
            p.bred1[c] = bs;
            p.bred1[p.n_total + c] = bd1;
            if (p.dgamma1) p.dgamma1[c] += bd1;
            if (p.dbeta1) p.dbeta1[c] += bs;
            if (bwd2) {
              const float bd2 = (float)((double)pre[it][3] * (S[2] - (double)pre[it][2] * S[0]));
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

    if (fuse) {
      // ================= grid-synchronous BatchNorm (CF_FUSE) =================
      // Every accumulator of this CTA is still in TMEM (the host guarantees tiles per CTA <= acc_sets) and every CTA's
      // per-channel sums are on their way to `partials`: meet all CTAs of the launch, derive this CTA's channel constants
      // exactly as bn_apply / bn_bwd_apply do in their deferred prologue, then run the epilogue a second time.
      const int tid = ew * 32 + lane;
      bar_sync_named(3, 256);   // every epilogue thread of this CTA has issued its atomics / finished its stores
      if (tid == 0) {
        __threadfence();
        atomicAdd(p.grid_bar, 1u);
        const uint64_t t0 = globaltimer_ns();
        unsigned int seen;
        do {
          asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(seen) : "l"(p.grid_bar) : "memory");
          if (seen < gridDim.x && globaltimer_ns() - t0 > 2000000000ull) __trap();   // a CTA of the launch never arrived
        } while (seen < gridDim.x);
        __threadfence();
      }
      bar_sync_named(3, 256);
      const int bn_ = p.block_n;
      float* s_par = s_wacc;    // [10][block_n] per-channel constants (the running sums are not needed any more)
      const int nb = pair_id % p.n_blocks;
      const bool writer = pair_id < p.n_blocks && rank == 0;   // the CTA that owns m_tile 0 of this channel block publishes
      for (int c = tid; c < bn_; c += 256) {
        const int cg = nb * bn_ + c;
        if (!bwd) {
          const BnStat st = bn_stat_from_sums(__ldcg(p.partials + cg), __ldcg(p.partials + p.n_total + cg), p.inv_count, p.unbias, p.eps,
                                              __ldg(p.gamma + cg), __ldg(p.beta + cg));
          s_par[c] = st.scale; s_par[bn_ + c] = st.shift;
          if (p.fuse_rvec) { s_par[2 * bn_ + c] = __ldg(p.fuse_rvec + cg); s_par[3 * bn_ + c] = __ldg(p.fuse_rvec + p.n_total + cg); }
          if (writer) {
            p.vec[cg] = st.scale; p.vec[p.n_total + cg] = st.shift; p.vec[2 * p.n_total + cg] = st.mean; p.vec[3 * p.n_total + cg] = st.rstd;
            if (p.update_running) {
              p.running_mean[cg] = (1.f - p.momentum) * p.running_mean[cg] + p.momentum * st.mean;
              p.running_var[cg] = (1.f - p.momentum) * p.running_var[cg] + p.momentum * st.unbiased_var;
            }
          }
        } else {
          const float invc = (float)p.inv_count;
          const double S0 = __ldcg(p.partials + cg);
#pragma unroll
          for (int k2 = 0; k2 < 2; ++k2) {
            if (k2 == 1 && !bwd2) break;
            const float* stat = k2 ? p.stat2 : p.stat1;
            const float mean = __ldg(stat + 2 * p.n_total + cg), rstd = __ldg(stat + 3 * p.n_total + cg);
            const double S1 = __ldcg(p.partials + (1 + k2) * p.n_total + cg);
            const float bs = (float)S0, bd = bn_bdot_from_sums(S0, S1, mean, rstd);
            float* sp = s_par + 5 * k2 * bn_;
            sp[c] = mean; sp[bn_ + c] = rstd; sp[2 * bn_ + c] = __ldg((k2 ? p.gamma2 : p.gamma1) + cg);
            sp[3 * bn_ + c] = bs * invc; sp[4 * bn_ + c] = bd * invc;
            if (writer) {
              float* bred = k2 ? p.bred2 : p.bred1;
              float* dg = k2 ? p.dgamma2 : p.dgamma1;
              float* db = k2 ? p.dbeta2 : p.dbeta1;
              bred[cg] = bs; bred[p.n_total + cg] = bd;
              if (dg) dg[cg] += bd;
              if (db) db[cg] += bs;
            }
          }
        }
      }
      if (writer && nb == 0 && tid == 0 && !bwd && p.update_running && p.nbt) *p.nbt += 1;
      bar_sync_named(3, 256);
      tc_fence_after();
      const bool no_store = (p.flags & CF_NO_STORE) != 0;
      const bool fres = (p.flags & CF_FUSE_RES) != 0;
      int acc2 = 0;
      uint32_t uc2 = 0;
      for (int tile = pair_id; tile < total_tiles; tile += n_pairs, ++acc2) {
        const int m_tile = tile / p.n_blocks;
        const int n_blk = tile - m_tile * p.n_blocks;
        const int row0 = (m_tile * (PAIR ? 2 : 1) + (int)rank) * MT * 128;
#pragma unroll 1
        for (int m = 0; m < MT; ++m) {
          const int f = row0 + m * 128 + row;
          const bool in_range = f < p.total_rows;
          bool valid = in_range;
          if (valid) {
            const unsigned int uf = (unsigned int)f;
            const unsigned int wq = uf / (unsigned int)p.g.Wp;
            const unsigned int w = uf - wq * (unsigned int)p.g.Wp;
            const unsigned int h = wq % (unsigned int)p.g.Hp;
            valid = (w < (unsigned int)p.g.W) && (h < (unsigned int)p.g.H);
          }
#pragma unroll 1
          for (int chunk = 0; chunk < n_chunks; ++chunk) {
            if (((uc2++) & 1u) != (uint32_t)grp) continue;
            const int n_base = n_blk * p.block_n + chunk * 64;
            const int cl = chunk * 64;   // channel offset inside this CTA's block
            const long long goff = (long long)f * p.n_total + n_base;
            const int row_first = row0 + m * 128 + q * 32;
            tma_store_wait_read();
            __syncwarp();
            if (lane == 0) {
              if (!bwd && fres) {
                mbar_arrive_expect_tx(bar_res, 32 * 128);
                tma_load_2d(&p.tmRes, bar_res, w_out, n_base, row_first);
              }
              if (bwd) {
                mbar_arrive_expect_tx(bar_y, 32 * 128);
                tma_load_2d(&p.tmY1, bar_y, w_y, n_base, row_first);
                if (!no_store) {   // the dz tile this warp stored in pass 1 (complete: tma_store_wait_all above)
                  mbar_arrive_expect_tx(bar_res, 32 * 128);
                  tma_load_2d(&p.tmOut, bar_res, w_out, n_base, row_first);
                }
              }
            }
            float x[64];   // forward: bf16-rounded conv output; backward: dz
            if (!bwd || no_store) {
              uint32_t v[64];
              const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc2 * acc_stride + m * p.block_n + chunk * 64);
              tmem_ld_32x32(taddr, v);
              tmem_ld_32x32(taddr + 32, v + 32);
              tmem_ld_wait();
              if (bwd) {
                uint2 mbits = make_uint2(0xffffffffu, 0xffffffffu);
                if ((p.flags & CF_MASK) && p.mask_bits != nullptr && valid) mbits = __ldg(reinterpret_cast<const uint2*>(p.mask_bits + (goff >> 3)));
#pragma unroll
                for (int j = 0; j < 64; ++j)
                  if (!(((j < 32 ? mbits.x : mbits.y) >> (j & 31)) & 1u)) v[j] = 0u;
              }
#pragma unroll
              for (int j = 0; j < 64; ++j) x[j] = valid ? __bfloat162float(__float2bfloat16(__uint_as_float(v[j]))) : 0.f;
            } else {
              mbar_wait(bar_res, ph_res);
              ph_res ^= 1;
#pragma unroll
              for (int j = 0; j < 8; ++j) {
                const uint4 t4 = *(const uint4*)(w_out + lane * 128 + ((j ^ (lane & 7)) << 4));
                x[8 * j] = bf16lo(t4.x); x[8 * j + 1] = bf16hi(t4.x); x[8 * j + 2] = bf16lo(t4.y); x[8 * j + 3] = bf16hi(t4.y);
                x[8 * j + 4] = bf16lo(t4.z); x[8 * j + 5] = bf16hi(t4.z); x[8 * j + 6] = bf16lo(t4.w); x[8 * j + 7] = bf16hi(t4.w);
              }
            }
            uint32_t u[32];
            if (!bwd) {
              // out2 = relu( x * scale + shift [+ res | + res * rscale + rshift] ), and its ReLU bits
#pragma unroll
              for (int j = 0; j < 64; ++j) x[j] = fmaf(x[j], s_par[cl + j], s_par[bn_ + cl + j]);
              if (fres) {
                mbar_wait(bar_res, ph_res);
                ph_res ^= 1;
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                  const uint4 t4 = *(const uint4*)(w_out + lane * 128 + ((j ^ (lane & 7)) << 4));
                  float r[8] = {bf16lo(t4.x), bf16hi(t4.x), bf16lo(t4.y), bf16hi(t4.y), bf16lo(t4.z), bf16hi(t4.z), bf16lo(t4.w), bf16hi(t4.w)};
#pragma unroll
                  for (int e = 0; e < 8; ++e) {
                    if (p.fuse_rvec) r[e] = fmaf(r[e], s_par[2 * bn_ + cl + 8 * j + e], s_par[3 * bn_ + cl + 8 * j + e]);
                    x[8 * j + e] += r[e];
                  }
                }
              }
              uint2 ob = make_uint2(0u, 0u);
#pragma unroll
              for (int j = 0; j < 32; ++j) {
                const float a = fmaxf(x[2 * j], 0.f), b = fmaxf(x[2 * j + 1], 0.f);
                u[j] = valid ? pack_bf16x2(a, b) : 0u;
                const uint32_t bits2 = ((u[j] & 0xFFFFu) != 0u && !(u[j] & 0x8000u) ? 1u : 0u) | ((u[j] >> 16) != 0u && !(u[j] & 0x80000000u) ? 2u : 0u);
                if (j < 16) ob.x |= bits2 << (2 * j); else ob.y |= bits2 << (2 * (j - 16));
              }
              if (in_range && p.bits_out) *reinterpret_cast<uint2*>(p.bits_out + (goff >> 3)) = ob;
              __syncwarp();  // every lane has read its residual row out of the staging tile
#pragma unroll
              for (int j = 0; j < 8; ++j)
                *(uint4*)(w_out + lane * 128 + ((j ^ (lane & 7)) << 4)) = make_uint4(u[4 * j], u[4 * j + 1], u[4 * j + 2], u[4 * j + 3]);
              fence_proxy_async();
              __syncwarp();
              if (lane == 0) {
                tma_store_2d(&p.tmOut2, w_out, n_base, row_first);
                tma_store_commit();
              }
            } else {
              // dy = gamma * rstd * (dz - k0 - xhat * k1), xhat = (y - mean) * rstd, for BN 1 (and BN 2 on the same dz)
#pragma unroll 1
              for (int k2 = 0; k2 < 2; ++k2) {
                if (k2 == 1 && !bwd2) break;
                if (k2 == 1) {
                  __syncwarp();   // every lane has read its y1 row
                  if (lane == 0) {
                    mbar_arrive_expect_tx(bar_y, 32 * 128);
                    tma_load_2d(&p.tmY2, bar_y, w_y, n_base, row_first);
                  }
                }
                mbar_wait(bar_y, ph_y);
                ph_y ^= 1;
                const float* sp = s_par + 5 * k2 * bn_ + cl;
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                  const uint4 t4 = *(const uint4*)(w_y + lane * 128 + ((j ^ (lane & 7)) << 4));
                  const float yv[8] = {bf16lo(t4.x), bf16hi(t4.x), bf16lo(t4.y), bf16hi(t4.y), bf16lo(t4.z), bf16hi(t4.z), bf16lo(t4.w), bf16hi(t4.w)};
                  float o[8];
#pragma unroll
                  for (int e = 0; e < 8; ++e) {
                    const int cc = 8 * j + e;
                    const float mean = sp[cc], rstd = sp[bn_ + cc], gam = sp[2 * bn_ + cc], k0 = sp[3 * bn_ + cc], k1 = sp[4 * bn_ + cc];
                    const float xhat = (yv[e] - mean) * rstd;
                    o[e] = gam * rstd * (x[cc] - k0 - xhat * k1);
                  }
                  u[4 * j] = valid ? pack_bf16x2(o[0], o[1]) : 0u; u[4 * j + 1] = valid ? pack_bf16x2(o[2], o[3]) : 0u;
                  u[4 * j + 2] = valid ? pack_bf16x2(o[4], o[5]) : 0u; u[4 * j + 3] = valid ? pack_bf16x2(o[6], o[7]) : 0u;
                }
                if (k2 == 1) tma_store_wait_read();   // the dy1 store has read the staging tile
                __syncwarp();  // every lane has read its dz row out of the staging tile
#pragma unroll
                for (int j = 0; j < 8; ++j)
                  *(uint4*)(w_out + lane * 128 + ((j ^ (lane & 7)) << 4)) = make_uint4(u[4 * j], u[4 * j + 1], u[4 * j + 2], u[4 * j + 3]);
                fence_proxy_async();
                __syncwarp();
                if (lane == 0) {
                  tma_store_2d(k2 ? &p.tmOut3 : &p.tmOut2, w_out, n_base, row_first);
                  tma_store_commit();
                }
              }
            }
          }
        }
      }
      tma_store_wait_all();
    }
  }

  tc_fence_before();
  if (PAIR) cluster_sync_all();  // nobody leaves while the peer may still signal its barriers or the pair's MMAs run
  else __syncthreads();
#ifdef CF_TRACE
  if (blockIdx.x == 0 && threadIdx.x < 3) g_cf_trace_n[threadIdx.x] = (int)cf_idx[threadIdx.x];
#endif
  if (warp == 1) {
    __syncwarp();
    if (PAIR) tmem_dealloc_pair(tmem_base, 512); else tmem_dealloc(tmem_base, 512);
  }
}

}  // namespace cilrs
