// Host side of the padded-flat 3x3 stride-1 convolution kernels (conv_flat.cuh, wgrad_flat.cuh): tile-shape choice,
// tensor maps, launches and their C-ABI entry points (include/cilrs_b200.h).
#include "conv_flat.cuh"
#include "wgrad_flat.cuh"
#include "conv_host.h"
#include <string.h>
#include <stdio.h>
#include <stdlib.h>

namespace cilrs {

static int sm_count() {
  static int n = 0;
  if (!n) {
    int dev = 0;
    cudaGetDevice(&dev);
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
  }
  return n;
}

static int round_up(int x, int a) { return (x + a - 1) / a * a; }

// slab of `rows` pixel rows as `boxes` TMA boxes of `box_rows` (multiple of 8, <= 256) rows each
static void split_boxes(int rows, int* box_rows, int* boxes) {
  const int n = (rows + 255) / 256;
  *box_rows = round_up((rows + n - 1) / n, 8);
  *boxes = n;
}

static long long flat_smem_fixed(int block_n, int flags) {
  return 1024 /* alignment slack */ + CF_STAGING_BYTES + ((flags & CF_BNBWD) ? CF_STAGING_BYTES : 0) /* y tiles */ +
         8LL * 3 * block_n * 4 /* warp-private statistics */ +
         (2 * CF_MAX_A_STAGES + 2 * CF_MAX_B_STAGES + 2 * CF_MAX_ACC + 16) * 8 + 64 + 64;
}

// Tile-shape choice by a small cost model in clocks per CTA, calibrated on B200 traces (tools/trace_flat.py):
//  * a tcgen05.mma 128 x N x 16 takes max(N/2, shared-memory operand fetch (4096 + 32 N bytes at 128 B/clk)) clocks:
//    48 / 64 / 128 for N = 64 / 128 / 256;
//  * every mbarrier round trip of the issuing warp (one per weight stage = `tap_group` taps) costs ~550 clocks of which the
//    tensor-core queue hides ~3.3 MMAs;
//  * a weight stage comes around every b_stages groups and is busy for its MMAs + ~700 clocks of TMA latency + its transfer:
//    two big stages stall the issuing warp where four small ones do not (layer3: tap_group 1 x 4 stages beats 2 x 2);
//  * a 128 x 64 epilogue unit costs ~2000 clocks and is only hidden behind the MMAs of a following tile;
//  * activations + weights cross L2->SMEM at ~min(64, 6300 / active CTAs) bytes per clock.
struct FlatShape {
  int mt, block_n, resident, tap_group, a_stages, b_stages, a_box_rows, a_boxes;
  double cost;
  int pair;
};

template <int MT, bool PAIR> static void (*flat_kernel_ptr())(FlatConvParams) { return conv_flat_kernel<MT, PAIR>; }
static void (*flat_kernel(int mt, int pair))(FlatConvParams) {
  if (pair) return mt == 1 ? flat_kernel_ptr<1, true>() : (mt == 2 ? flat_kernel_ptr<2, true>() : (mt == 4 ? flat_kernel_ptr<4, true>() : nullptr));
  return mt == 1 ? flat_kernel_ptr<1, false>() : (mt == 2 ? flat_kernel_ptr<2, false>() : (mt == 4 ? flat_kernel_ptr<4, false>() : nullptr));
}
static int flat_kernel_attrs() {
  static bool done = false;
  if (done) return OK;
  const int mts[3] = {1, 2, 4};
  for (int pr = 0; pr < 2; ++pr)
    for (int i = 0; i < 3; ++i) {
      cudaError_t e = cudaFuncSetAttribute(flat_kernel(mts[i], pr), cudaFuncAttributeMaxDynamicSharedMemorySize, CG_SMEM_TOTAL);
      if (e != cudaSuccess) return cuda_status(e);
    }
  done = true;
  return OK;
}
// CTA pairs (clusters of 2) that can be resident at once; 0 = pairs unavailable / switched off (CILRS_FLAT_PAIR=0)
static int max_pairs() {
  static int n = -1;
  if (n >= 0) return n;
  n = 0;
  const char* env = getenv("CILRS_FLAT_PAIR");
  if (env && env[0] == '0') return n;
  if (flat_kernel_attrs() != OK) return n;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(2 * sm_count()); cfg.blockDim = dim3(CF_THREADS); cfg.dynamicSmemBytes = CG_SMEM_TOTAL;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr; cfg.numAttrs = 1;
  int clusters = 0;
  if (cudaOccupancyMaxActiveClusters(&clusters, flat_kernel(1, 1), &cfg) == cudaSuccess && clusters > 0) n = clusters;
  else cudaGetLastError();
  if (n > sm_count() / 2) n = sm_count() / 2;
  return n;
}

// bn_cta = weight-tile rows held by one CTA (block_n, or block_n / 2 for a CTA pair)
static bool flat_stages(long long budget, int chunks, int n_blocks, int bn_cta, long long a_stage, int resident, int G, int* a_st, int* b_st) {
  const long long b_stage = (long long)G * bn_cta * 128;
  if (resident) {
    (void)n_blocks;
    if (G != 9 || chunks > CF_MAX_B_STAGES) return false;
    const long long left = budget - chunks * b_stage;
    if (left < 2 * a_stage) return false;
    *b_st = chunks;
    *a_st = (int)(left / a_stage) > CF_MAX_A_STAGES ? CF_MAX_A_STAGES : (int)(left / a_stage);
    return true;
  }
  long long left = budget - 2 * a_stage;
  if (left < 2 * b_stage) return false;
  int nb = (int)(left / b_stage);
  if (nb > CF_MAX_B_STAGES) nb = CF_MAX_B_STAGES;
  if (nb * G > 18) nb = (18 + G - 1) / G;  // more than two chunks of weights in flight buys nothing
  if (nb < 2) nb = 2;
  left -= nb * b_stage;
  *b_st = nb;
  *a_st = left >= a_stage ? 3 : 2;
  return true;
}

static FlatShape choose_flat_shape(int total_rows, int k_channels, int n_total, const PadGeom& g, int flags) {
  const int sms = sm_count();
  const int chunks = k_channels / 64;
  const int halo = g.Wp + 1;
  FlatShape best{};
  best.cost = 1e30;
  const int bns[3] = {256, 128, 64};
  const int mts[3] = {1, 2, 4};
  const int groups[4] = {9, 3, 2, 1};
  const int pairs_avail = max_pairs();
  for (int pr = 0; pr < 2; ++pr) {
    if (pr && pairs_avail < 1) continue;
    for (int bi = 0; bi < 3; ++bi) {
      const int bn = bns[bi];
      if (n_total % bn) continue;
      const int n_blocks = n_total / bn;
      const int bn_cta = pr ? bn / 2 : bn;  // weight-tile rows per CTA
      const long long budget = CG_SMEM_TOTAL - flat_smem_fixed(bn, flags);
      // tcgen05.mma 128(x2) x N x 16: max(N/2, shared-memory operand fetch per CTA at 128 B/clk)
      const double fetch = (4096.0 + 32.0 * bn_cta) / 128.0;
      const double mma_clk = bn / 2.0 > fetch ? bn / 2.0 : fetch;
      for (int mi = 0; mi < 3; ++mi) {
        const int mt = mts[mi];
        if (mt * bn > 512) continue;
        int box_rows, boxes;
        split_boxes(mt * 128 + 2 * halo, &box_rows, &boxes);
        const long long a_stage = (long long)box_rows * boxes * 128;
        for (int gi = 0; gi < 4; ++gi) {
          const int res = gi == 0;
          const int G = res ? 9 : groups[gi];
          int a_st, b_st;
          if (!flat_stages(budget, chunks, n_blocks, bn_cta, a_stage, res, G, &a_st, &b_st)) continue;
          const int tile_rows = (pr ? 2 : 1) * mt * 128;
          const int m_tiles = (total_rows + tile_rows - 1) / tile_rows;
          const long long tiles = (long long)m_tiles * n_blocks;
          const int slots = pr ? pairs_avail : sms;          // CTAs or CTA pairs that work on tiles concurrently
          int active = tiles < slots ? (int)tiles : slots;
          active -= active % n_blocks;  // the grid is a multiple of n_blocks (one channel block per CTA)
          if (active < 1) continue;
          const int ctas = active * (pr ? 2 : 1);
          const double bw = 6300.0 / ctas < 64.0 ? 6300.0 / ctas : 64.0;
          const double exposed = 550.0 - 3.3 * mma_clk > 60.0 ? 550.0 - 3.3 * mma_clk : 60.0;
          double chunk_clk = 0.0;
          for (int t0 = 0; t0 < 9; t0 += G) {
            const int cnt = 9 - t0 < G ? 9 - t0 : G;
            double group = cnt * mt * 4 * mma_clk + exposed;
            if (!res) {
              // weight-stage pipeline: a stage is busy for its own MMAs plus its refill (~700 clocks of TMA latency + the
              // transfer at ~64 B/clk), and comes around every b_st groups
              const double refill = 700.0 + cnt * bn_cta * 128.0 / 64.0;
              const double ring = (refill + cnt * mt * 4 * mma_clk) / b_st;
              if (ring > group) group = ring;
            }
            chunk_clk += group;
          }
          const double mma = chunks * chunk_clk;
          const double bytes = (double)chunks * a_stage + (res ? 0.0 : 9.0 * chunks * bn_cta * 128.0);
          const double tile_clk = mma > bytes / bw ? mma : bytes / bw;
          const long long rounds = (tiles + active - 1) / active;
          const double units = mt * (bn / 64) * 0.5;                      // per epilogue group (and CTA)
          const int acc_sets = 512 / (mt * bn);
          // the epilogue of a tile hides behind the next tile's MMAs only with >= 2 accumulator sets and a next tile
          const double epi_tail = units * 2000.0 * (acc_sets >= 2 ? 1.0 : (double)rounds);
          const double cost = rounds * tile_clk + 1500.0 + epi_tail;
          if (cost < best.cost * 0.99) best = FlatShape{mt, bn, res, G, a_st, b_st, box_rows, boxes, cost, pr};
        }
      }
    }
  }
  // measurement aid: CILRS_FLAT_SHAPE="mt,block_n,resident,tap_group[,pair]" overrides the model
  if (const char* env = getenv("CILRS_FLAT_SHAPE")) {
    int mt = 0, bn = 0, res = 0, G = 0, pr = 0;
    const int nf = sscanf(env, "%d,%d,%d,%d,%d", &mt, &bn, &res, &G, &pr);
    if (nf >= 4 && (mt == 1 || mt == 2 || mt == 4) && bn >= 64 && n_total % bn == 0 && mt * bn <= 512 && G >= 1 && G <= 9 &&
        (!pr || pairs_avail > 0)) {
      int box_rows, boxes, a_st, b_st;
      split_boxes(mt * 128 + 2 * halo, &box_rows, &boxes);
      const long long a_stage = (long long)box_rows * boxes * 128;
      const long long budget = CG_SMEM_TOTAL - flat_smem_fixed(bn, flags);
      if (res) G = 9;
      if (flat_stages(budget, chunks, n_total / bn, pr ? bn / 2 : bn, a_stage, res, G, &a_st, &b_st))
        best = FlatShape{mt, bn, res, G, a_st, b_st, box_rows, boxes, 0.0, pr ? 1 : 0};
      else
        best.cost = 1e30;  // the requested shape does not fit: fail instead of silently measuring the model's choice
    } else {
      best.cost = 1e30;
    }
  }
  if (getenv("CILRS_FLAT_DEBUG"))
    fprintf(stderr, "[cilrs flat] rows=%d K=%d N=%d Wp=%d flags=%d -> pair=%d mt=%d block_n=%d resident=%d tap_group=%d a_stages=%d b_stages=%d a_box=%dx%d cost=%.0f\n",
            total_rows, k_channels, n_total, g.Wp, flags, best.pair, best.mt, best.block_n, best.resident, best.tap_group, best.a_stages,
            best.b_stages, best.a_boxes, best.a_box_rows, best.cost);
  return best;
}

int flat_total_rows(int batch, const PadGeom& g) { return batch * g.Hp * g.Wp; }

// taps of a 3x3 filter as flat shifts; dgrad reads dy at the mirrored position
static void flat_taps(FlatConvParams* p, int dgrad) {
  p->num_taps = 9;
  for (int r = 0; r < 3; ++r)
    for (int s = 0; s < 3; ++s) {
      const int t = r * 3 + s;
      const int sh = (r - 1) * p->g.Wp + (s - 1);
      p->tap_shift[t] = dgrad ? -sh : sh;
      p->tap_slab[t] = t;
    }
}

// x: padded-flat [rows][k_channels] bf16, w: bf16 [9][n_total][k_channels], out: padded-flat [rows][n_total]
int build_flat_conv(FlatConvParams* p, int batch, const PadGeom& g, int k_channels, int n_total, int dgrad, const void* x,
                    const void* w, void* out, int flags) {
  if (batch < 1 || g.H < 1 || g.W < 1 || g.Hp < g.H || g.Wp <= g.W || g.Wp < 3) return ERR_INVALID;
  if (k_channels % 64 || n_total % 64 || k_channels < 64 || n_total < 64 || n_total > 512) return ERR_UNSUPPORTED;
  memset(p, 0, sizeof(*p));
  p->g = g;
  p->total_rows = flat_total_rows(batch, g);
  const FlatShape sh = choose_flat_shape(p->total_rows, k_channels, n_total, g, flags);
  if (sh.cost >= 1e30) return ERR_UNSUPPORTED;
  p->pair = sh.pair;
  p->mt = sh.mt; p->block_n = sh.block_n; p->n_blocks = n_total / sh.block_n; p->n_total = n_total;
  p->chunks = k_channels / 64;
  p->halo = g.Wp + 1;
  p->a_box_rows = sh.a_box_rows; p->a_boxes = sh.a_boxes;
  p->a_stages = sh.a_stages; p->b_stages = sh.b_stages; p->b_resident = sh.resident; p->tap_group = sh.tap_group;
  int sets = 512 / (sh.mt * sh.block_n);
  p->acc_sets = sets > CF_MAX_ACC ? CF_MAX_ACC : sets;
  const int tile_rows = (sh.pair ? 2 : 1) * sh.mt * 128;
  p->m_tiles = (p->total_rows + tile_rows - 1) / tile_rows;
  flat_taps(p, dgrad);
  p->flags = flags;
  p->out = (__nv_bfloat16*)out;
  int st = encode_2d_map(&p->tmA, x, k_channels, p->total_rows, 64, p->a_box_rows);
  if (st) return st;
  st = encode_2d_map(&p->tmOut, out, n_total, p->total_rows, 64, 32);
  if (st) return st;
  return encode_2d_map(&p->tmB, w, k_channels, 9 * n_total, 64, sh.pair ? p->block_n / 2 : p->block_n);
}

// tensor maps of the epilogue's operand tiles (residual, y1, y2): call after the pointers are set, before the launch
int flat_conv_bind_operands(FlatConvParams* p) {
  p->operand_maps = 0;
  if ((p->flags & (CF_RESIDUAL | CF_FUSE_RES)) && p->residual) {
    int st = encode_2d_map(&p->tmRes, p->residual, p->n_total, p->total_rows, 64, 32);
    if (st) return st;
  }
  if ((p->flags & CF_BNBWD) && p->y1) {
    int st = encode_2d_map(&p->tmY1, p->y1, p->n_total, p->total_rows, 64, 32);
    if (st) return st;
  }
  if ((p->flags & CF_BNBWD2) && p->y2) {
    int st = encode_2d_map(&p->tmY2, p->y2, p->n_total, p->total_rows, 64, 32);
    if (st) return st;
  }
  p->operand_maps = 1;
  return OK;
}

// CF_FUSE (grid-synchronous BatchNorm): possible when every tile's accumulator can stay in TMEM until the grid barrier, i.e. no
// CTA gets more tiles than it has accumulator sets. CILRS_NO_FUSE=1 switches it off (A/B measurements, equivalence tests).
// OFF by default: measured on B200 at batch 128 the fused step is 3.45 ms against 3.05 ms - the second pass costs the conv
// launch about what the separate elementwise launch cost (it is serial per 128 x 64 unit: TMA operand load -> TMEM read ->
// store, on 8 warps per SM), and the elementwise launches it removes were hiding the weight-gradient stream. Kept as an
// option (CILRS_BN_FUSION=1 / cilrs_set_bn_fusion) and covered by the parity tests; see DESIGN.md.
static int g_fuse_enabled = -1;   // -1: not decided yet
// 0 = off, 3 / 4 = forward / data-gradient convolutions only, 1 = every flat convolution that fits tensor memory, 2 = only the 256- / 512-channel layers (layers 3-4: small tensors,
// 98 / 128-CTA grids that leave SMs to the weight-gradient stream, one or two units per CTA in the second pass)
static int fuse_default() { const char* e = getenv("CILRS_BN_FUSION"); return (e && e[0] >= '0' && e[0] <= '4') ? e[0] - '0' : 0; }
int flat_conv_fuse_ok(const FlatConvParams* p) {
  if (g_fuse_enabled < 0) g_fuse_enabled = fuse_default();
  if (!g_fuse_enabled) return 0;
  if (g_fuse_enabled == 2 && p->n_total < 256) return 0;
  if (g_fuse_enabled == 3 && !(p->flags & CF_STATS)) return 0;   // 3: forward convolutions only
  if (g_fuse_enabled == 4 && !(p->flags & CF_BNBWD)) return 0;   // 4: data-gradient convolutions only
  const long long total = (long long)p->m_tiles * p->n_blocks;
  const int grid = flat_conv_grid(p);
  const int n = p->pair ? grid / 2 : grid;
  if (n < 1) return 0;
  return (total + n - 1) / n <= p->acc_sets ? 1 : 0;
}
// tensor maps of the second-pass outputs (same [rows][n_total] geometry as `out`)
int flat_conv_bind_fuse(FlatConvParams* p, void* out2, void* out3) {
  if (!out2) return ERR_INVALID;
  int st = encode_2d_map(&p->tmOut2, out2, p->n_total, p->total_rows, 64, 32);
  if (st) return st;
  if (out3) st = encode_2d_map(&p->tmOut3, out3, p->n_total, p->total_rows, 64, 32);
  return st;
}

// CTAs of the launch: one per tile up to the SM count (pairs: two per tile up to the resident pair count), a multiple of
// n_blocks tiles wide so that every CTA only ever sees one channel block
int flat_conv_grid(const FlatConvParams* p) {
  const long long total = (long long)p->m_tiles * p->n_blocks;
  const int slots = p->pair ? max_pairs() : sm_count();
  int n = total < slots ? (int)total : slots;
  n -= n % p->n_blocks;
  return p->pair ? 2 * n : n;
}

int launch_flat_conv(const FlatConvParams* p, cudaStream_t s) {
  int st = flat_kernel_attrs();
  if (st) return st;
  if ((p->flags & CF_BNBWD) && (!p->operand_maps || !p->y1 || ((p->flags & CF_BNBWD2) && !p->y2))) return ERR_INVALID;
  if ((p->flags & (CF_STATS | CF_BNBWD)) && (!p->partials || (!(p->flags & CF_DEFER) && !p->counter))) return ERR_INVALID;
  if (p->flags & CF_FUSE) {
    if (!(p->flags & CF_DEFER) || !(p->flags & (CF_STATS | CF_BNBWD)) || !p->grid_bar || !flat_conv_fuse_ok(p)) return ERR_INVALID;
    if ((p->flags & CF_STATS) && (!p->gamma || !p->beta || !p->vec)) return ERR_INVALID;
    if ((p->flags & CF_BNBWD) && (!p->stat1 || !p->gamma1 || !p->bred1 || ((p->flags & CF_BNBWD2) && (!p->stat2 || !p->gamma2 || !p->bred2)))) return ERR_INVALID;
    if ((p->flags & CF_NO_STORE) && ((p->flags & CF_RESIDUAL) || !(p->flags & CF_BNBWD) || ((p->flags & CF_MASK) && !p->mask_bits))) return ERR_INVALID;
    if ((p->flags & CF_FUSE_RES) && (!p->operand_maps || !p->residual)) return ERR_INVALID;
  }
  const int grid = flat_conv_grid(p);
  void (*kernel)(FlatConvParams) = flat_kernel(p->mt, p->pair);
  if (!kernel || grid < 1) return ERR_INVALID;
  ++g_cilrs_launches;
  if (p->pair) return cuda_status(launch_pdl_cluster(kernel, dim3(grid), dim3(CF_THREADS), CG_SMEM_TOTAL, s, 2, *p));
  return cuda_status(launch_pdl(kernel, dim3(grid), dim3(CF_THREADS), CG_SMEM_TOTAL, s, *p));
}

// ------------------------------------------------------------------------------------------------
// wgrad
// ------------------------------------------------------------------------------------------------
static int g_wgrad_cluster = -1;   // -1: not decided yet (environment), 0 / 1
// clusters of three wgrad_flat3_kernel CTAs that can be resident at once (0: unavailable)
static int max_clusters3() {
  static int n = -1;
  if (n >= 0) return n;
  n = 0;
  if (cudaFuncSetAttribute(wgrad_flat3_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, CG_SMEM_TOTAL) != cudaSuccess) { cudaGetLastError(); return n; }
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(3 * sm_count()); cfg.blockDim = dim3(WF_THREADS); cfg.dynamicSmemBytes = CG_SMEM_TOTAL;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 3; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr; cfg.numAttrs = 1;
  int clusters = 0;
  if (cudaOccupancyMaxActiveClusters(&clusters, wgrad_flat3_kernel, &cfg) == cudaSuccess && clusters > 0) n = clusters;
  else cudaGetLastError();
  if (getenv("CILRS_FLAT_DEBUG")) fprintf(stderr, "[cilrs wgrad] resident clusters of 3: %d\n", n);
  return n;
}

int build_wgrad_flat(WgradFlatParams* p, int batch, const PadGeom& g, int cin, int cout, const void* dy, const void* x, float* scratch) {
  if (batch < 1 || g.H < 1 || g.W < 1 || g.Hp < g.H || g.Wp <= g.W || g.Wp < 3) return ERR_INVALID;
  if (cin % 64 || cout % 64 || cin < 64 || cout < 64) return ERR_UNSUPPORTED;
  memset(p, 0, sizeof(*p));
  p->total_rows = flat_total_rows(batch, g);
  p->k_tiles = (p->total_rows + 127) / 128;
  p->co_blocks = (cout + 127) / 128;
  p->m_halves = cout >= 128 ? 2 : 1;
  p->ci_chunks = cin / 64;
  p->tap_groups = 3;  // one filter row (dw = -1, 0, +1) per CTA: a single N = 192 MMA over a 130-pixel slab
  for (int r = 0; r < 3; ++r)
    for (int s = 0; s < 3; ++s) p->tap_shift[r * 3 + s] = (r - 1) * g.Wp + (s - 1);
  split_boxes(128 + 2, &p->x_box_rows, &p->x_boxes);
  const int base = p->co_blocks * p->ci_chunks * p->tap_groups;
  // CTAs of the launch = base * z, the depth of the K split. One CTA per SM is the fastest launch in isolation, but in the step
  // the weight gradients of layers 1-2 share the GPU with the dgrad / BatchNorm chain, which is the critical path there: with
  // at most HALF the SMs (74) the step is 40 us shorter (2.71 -> 2.67 ms at batch 128; 66..74 CTAs measure the same, 78 and
  // more lose it again; layers 3-4 do not care). Measurement aid: CILRS_WGRAD_CTAS="l1,l2,l3,l4" overrides the targets.
  static int targets[4] = {-1, -1, -1, -1};   // by input channels 64 / 128 / 256 / 512
  if (targets[0] < 0) {
    int t[4] = {0, 0, 0, 0};
    const char* env = getenv("CILRS_WGRAD_CTAS");
    const int nf = env ? sscanf(env, "%d,%d,%d,%d", &t[0], &t[1], &t[2], &t[3]) : 0;
    for (int i = 0; i < 4; ++i) {
      const int v = i < nf ? t[i] : (nf == 1 ? t[0] : 0);
      targets[i] = v >= 1 ? v : (i < 2 ? sm_count() / 2 : sm_count());
    }
  }
  const int target = targets[cin <= 64 ? 0 : (cin <= 128 ? 1 : (cin <= 256 ? 2 : 3))];
  int z = target / base;
  if (z < 1) z = 1;
  if (z > p->k_tiles) z = p->k_tiles;
  p->split_z = z;
  const long long stage = 2LL * WG_SLAB + (long long)p->x_boxes * p->x_box_rows * 128;
  int st = (int)((CG_SMEM_TOTAL - 1024 - 256) / stage);
  if (st > WF_MAX_STAGES) st = WF_MAX_STAGES;
  if (st < 2) return ERR_UNSUPPORTED;
  p->num_stages = st;
  p->cout = cout; p->cin = cin;
  p->scratch = scratch;
  int e = encode_2d_map(&p->tmDY, dy, cout, p->total_rows, 64, 128);
  if (e) return e;
  e = encode_2d_map(&p->tmX, x, cin, p->total_rows, 64, p->x_box_rows);
  if (e) return e;
  // cluster form (wgrad_flat3_kernel): one union slab of 130 + 2 Wp rows serves the three filter rows. OFF by default
  // (CILRS_WGRAD_CLUSTER=1 / cilrs_set_wgrad_cluster): measured on B200 at batch 128 it is slower - 0.99 ms against 0.93 ms for
  // the 40 weight-gradient launches of a step, 2.775 against 2.709 ms per step. The kernel was not bound by the crossbar after
  // all, three CTAs in lock-step lose more than the 2.5x smaller L2 -> SM traffic wins, and only 45 clusters of three SMs are
  // resident (135 of 148 SMs). Kept as a tested option and as evidence.
  p->cluster3 = 0;
  p->wp = g.Wp;
  p->xu_rows = round_up(130 + 2 * g.Wp, 8);
  if (g_wgrad_cluster < 0) { const char* env = getenv("CILRS_WGRAD_CLUSTER"); g_wgrad_cluster = (env && env[0] == '1') ? 1 : 0; }
  const int want = g_wgrad_cluster;
  const int maxc = want ? max_clusters3() : 0;
  if (want && p->xu_rows <= 256 && maxc >= p->co_blocks * p->ci_chunks) {
    // one wave: the K split is what fits the resident clusters (a cluster takes three SMs of ONE GPC, so fewer than
    // sm_count / 3 fit: with the plain split a few clusters ran as a second wave and the launch took 8 us longer)
    int z3 = maxc / (p->co_blocks * p->ci_chunks);
    if (z3 > p->k_tiles) z3 = p->k_tiles;
    const long long stage3 = 2LL * WG_SLAB + (long long)p->xu_rows * 128;
    int st3 = (int)((CG_SMEM_TOTAL - 1024 - 256) / stage3);
    if (st3 > WF_MAX_STAGES) st3 = WF_MAX_STAGES;
    if (st3 >= 2 && st3 * stage3 >= 128LL * WF_STAGE_PITCH && encode_2d_map(&p->tmXU, x, cin, p->total_rows, 64, p->xu_rows) == OK) {
      p->stages3 = st3;
      p->cluster3 = 1;
      if (z3 < p->split_z) p->split_z = z3;
    }
  }
  return OK;
}

int launch_wgrad_flat(const WgradFlatParams* p, cudaStream_t s) {
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(wgrad_flat_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, CG_SMEM_TOTAL);
    if (e != cudaSuccess) return cuda_status(e);
    attr_set = true;
  }
  const int grid = p->co_blocks * p->ci_chunks * p->tap_groups * p->split_z;
  if ((long long)p->co_blocks * p->ci_chunks * p->tap_groups * 128 * 192 * 4 > WF_SCRATCH_BYTES || !p->scratch) return ERR_WORKSPACE;
  if ((long long)p->num_stages * (2LL * WG_SLAB + (long long)p->x_boxes * p->x_box_rows * 128) < 128LL * WF_STAGE_PITCH) return ERR_UNSUPPORTED;
  ++g_cilrs_launches;
  if (p->cluster3) {
    static bool attr3_set = false;
    if (!attr3_set) {
      cudaError_t e = cudaFuncSetAttribute(wgrad_flat3_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, CG_SMEM_TOTAL);
      if (e != cudaSuccess) return cuda_status(e);
      attr3_set = true;
    }
    // (the same grid: the three filter rows of a work item are the three consecutive CTAs of a cluster)
    return cuda_status(launch_pdl_cluster(wgrad_flat3_kernel, dim3(grid), dim3(WF_THREADS), CG_SMEM_TOTAL, s, 3, *p));
  }
  return cuda_status(launch_pdl(wgrad_flat_kernel, dim3(grid), dim3(WF_THREADS), CG_SMEM_TOTAL, s, *p));
}

// append the reduction of one flat wgrad to a job list (the list is a kernel parameter, see wgrad_reduce_kernel)
int add_wgrad_reduce_job(WgradReduceJobs* jobs, const WgradFlatParams* p, long long grad_off) {
  if (jobs->n >= 16) return ERR_INVALID;
  WgradReduceJob& j = jobs->job[jobs->n++];
  // (the K slices of the wgrad kernel add into ONE accumulator tile: the reduce sees a single "slice" and clears it)
  j.scratch = p->scratch; j.grad_off = grad_off; j.cout = p->cout; j.cin = p->cin; j.ci_chunks = p->ci_chunks; j.split_z = 1;
  j.rows = 8; j.zero_src = 1;
  j.first_block = jobs->total_blocks;
  jobs->total_blocks += p->cout / j.rows * p->ci_chunks;
  return OK;
}

int launch_wgrad_reduce(const WgradReduceJobs* jobs, float* grads, cudaStream_t s) {
  if (jobs->n == 0) return OK;
  // 192 threads per K-slice chain; more chains when the split is deep (layer1: 49 slices)
  int max_z = 1;
  for (int i = 0; i < jobs->n; ++i) max_z = jobs->job[i].split_z > max_z ? jobs->job[i].split_z : max_z;
  int nz = (max_z + 11) / 12;
  nz = nz < 1 ? 1 : (nz > 4 ? 4 : nz);
  for (int i = 0; i < jobs->n; ++i)   // the 8-channel path is written for one accumulator tile and 192 threads
    if (jobs->job[i].rows > 1 && (jobs->job[i].rows != 8 || jobs->job[i].split_z != 1 || nz != 1 || jobs->job[i].cout % 8)) return ERR_INVALID;
  ++g_cilrs_launches;
  return cuda_status(launch_pdl(wgrad_reduce_kernel, dim3(jobs->total_blocks), dim3(192 * nz), 0, s, *jobs, grads));
}

}  // namespace cilrs

using namespace cilrs;

extern "C" {

long long cilrs_flat_rows(int batch, int H, int W) { return (long long)batch * (H + 1) * (W + 1); }

// Grid-synchronous BatchNorm inside the flat convolutions (conv_params.h: CF_FUSE) on / off for plans built AFTER the call
// (a cilrs_model rebuilds its plans when the batch size or mode changes, or via cilrs_model_invalidate_plans). Returns the
// previous setting. Measurement / test aid: both settings compute the same network.
int cilrs_set_bn_fusion(int enable) {
  const int prev = g_fuse_enabled < 0 ? fuse_default() : g_fuse_enabled;
  g_fuse_enabled = enable < 0 ? 0 : (enable > 4 ? 1 : enable);
  return prev;
}

// The weight-gradient kernel as clusters of three CTAs that multicast their operand loads (wgrad_flat3_kernel) on / off for plans
// built AFTER the call; returns the previous setting. Measurement / test aid: both kernels compute the same sums.
int cilrs_set_wgrad_cluster(int enable) {
  const int prev = g_wgrad_cluster < 0 ? 0 : g_wgrad_cluster;
  g_wgrad_cluster = enable ? 1 : 0;
  return prev;
}

int cilrs_conv_flat(const cilrs_flat_conv_args* a, void* stream) {
  if (!a || !a->x || !a->w || !a->y) return ERR_INVALID;
  const PadGeom g{a->H, a->W, a->H + 1, a->W + 1};
  FlatConvParams p;
  int flags = 0;
  if (a->flags & CILRS_EPI_STATS) flags |= CF_STATS;
  if (a->flags & CILRS_EPI_SCALE_BIAS) flags |= CF_SCALE_BIAS;
  if (a->flags & CILRS_EPI_RESIDUAL) flags |= CF_RESIDUAL;
  if (a->flags & CILRS_EPI_RELU) flags |= CF_RELU;
  if (a->flags & CILRS_EPI_MASK) flags |= CF_MASK;
  if (a->flags & CILRS_EPI_BNBWD) flags |= CF_BNBWD;
  if (a->flags & CILRS_EPI_BNBWD2) flags |= CF_BNBWD | CF_BNBWD2;
  if ((a->flags & CILRS_EPI_DEFER) && (flags & (CF_STATS | CF_BNBWD))) flags |= CF_DEFER;
  if ((flags & CF_STATS) && (flags & CF_BNBWD)) return ERR_INVALID;
  if ((flags & CF_SCALE_BIAS) && (!a->scale || !a->bias)) return ERR_INVALID;
  if ((flags & CF_RESIDUAL) && !a->residual) return ERR_INVALID;
  if ((flags & CF_MASK) && !a->mask) return ERR_INVALID;
  if ((flags & CF_STATS) && !(flags & CF_DEFER) && (!a->gamma || !a->beta || !a->running_mean || !a->running_var || !a->vec)) return ERR_INVALID;
  if ((flags & CF_BNBWD) && (!a->y1 || (!(flags & CF_DEFER) && (!a->vec1 || !a->bred1)))) return ERR_INVALID;
  if ((flags & CF_BNBWD2) && (!a->y2 || (!(flags & CF_DEFER) && (!a->vec2 || !a->bred2)))) return ERR_INVALID;
  if ((flags & (CF_STATS | CF_BNBWD)) && (!a->partials_ws || (!(flags & CF_DEFER) && !a->counter_ws))) return ERR_INVALID;
  int st = build_flat_conv(&p, a->batch, g, a->in_c, a->out_c, a->dgrad, a->x, a->w, a->y, flags);
  if (st) return st;
  p.residual = (const __nv_bfloat16*)a->residual; p.mask = (const __nv_bfloat16*)a->mask; p.mask_bits = (const uint8_t*)a->mask_bits;
  p.scale = a->scale; p.bias = a->bias;
  p.partials = (double*)a->partials_ws; p.counter = a->counter_ws;
  p.gamma = a->gamma; p.beta = a->beta; p.running_mean = a->running_mean; p.running_var = a->running_var;
  p.nbt = a->num_batches_tracked; p.vec = a->vec; p.count = (double)a->batch * a->H * a->W; p.momentum = a->momentum; p.eps = a->eps;
  p.update_running = a->update_running;
  p.y1 = (const __nv_bfloat16*)a->y1; p.stat1 = a->vec1; p.bred1 = a->bred1; p.dgamma1 = a->dgamma1; p.dbeta1 = a->dbeta1;
  p.y2 = (const __nv_bfloat16*)a->y2; p.stat2 = a->vec2; p.bred2 = a->bred2; p.dgamma2 = a->dgamma2; p.dbeta2 = a->dbeta2;
  st = flat_conv_bind_operands(&p);
  if (st) return st;
  return launch_flat_conv(&p, (cudaStream_t)stream);
}

#ifdef CF_TRACE
// debug build only: copy out (and reset) the event trace of CTA 0; out = 3 x 2048 uint64, counts = 3 ints
int cilrs_conv_flat_trace(unsigned long long* out, int* counts) {
  cudaDeviceSynchronize();
  cudaMemcpyFromSymbol(out, g_cf_trace, sizeof(unsigned long long) * 3 * 2048);
  cudaMemcpyFromSymbol(counts, g_cf_trace_n, sizeof(int) * 3);
  int zero[3] = {0, 0, 0};
  cudaMemcpyToSymbol(g_cf_trace_n, zero, sizeof(zero));
  return 0;
}
#endif

size_t cilrs_conv_flat_workspace_floats(int out_c) { return (size_t)256 * 3 * out_c; }

size_t cilrs_wgrad_flat_workspace_bytes(void) { return (size_t)WF_SCRATCH_BYTES; }

int cilrs_wgrad_flat(int batch, int H, int W, int in_c, int out_c, const void* dy, const void* x, float* dw_oihw, float* scratch_ws,
                     void* stream) {
  if (!dy || !x || !dw_oihw || !scratch_ws) return ERR_INVALID;
  const PadGeom g{H, W, H + 1, W + 1};
  WgradFlatParams p;
  int st = build_wgrad_flat(&p, batch, g, in_c, out_c, dy, x, scratch_ws);
  if (st) return st;
  st = launch_wgrad_flat(&p, (cudaStream_t)stream);
  if (st) return st;
  WgradReduceJobs jobs;
  jobs.n = 0; jobs.total_blocks = 0;
  st = add_wgrad_reduce_job(&jobs, &p, 0);
  if (st) return st;
  return launch_wgrad_reduce(&jobs, dw_oihw, (cudaStream_t)stream);
}

}  // extern "C"
