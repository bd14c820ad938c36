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

static long long flat_smem_fixed(int n_total) {
  return 1024 /* alignment slack */ + CF_STAGING_BYTES + 2 * 4 * 64 * 3 * 4 + 2LL * 3 * n_total * 4 +
         (2 * CF_MAX_A_STAGES + 2 * CF_MAX_B_STAGES + 2 * CF_MAX_ACC) * 8 + 64 + 64;
}

// Tile-shape choice by a small cost model (clocks per CTA): a (128*mt) x block_n tile issues 9*chunks*mt*4 MMAs of
// block_n*2/4... clocks and pulls one activation slab per chunk plus (unless resident) one weight tile per tap and chunk
// through L2->SMEM at ~min(64, 6300/active CTAs) bytes per clock; tiles are dealt round-robin to <= #SM persistent CTAs.
struct FlatShape {
  int mt, block_n, resident, a_stages, b_stages, a_box_rows, a_boxes;
  double cost;
};

static FlatShape choose_flat_shape(int total_rows, int k_channels, int n_total, const PadGeom& g) {
  const int sms = sm_count();
  const int chunks = k_channels / 64;
  const int halo = g.Wp + 1;
  const long long budget = CG_SMEM_TOTAL - flat_smem_fixed(n_total);
  FlatShape best{};
  best.cost = 1e30;
  const int bns[3] = {256, 128, 64};
  const int mts[3] = {1, 2, 4};
  for (int bi = 0; bi < 3; ++bi) {
    const int bn = bns[bi];
    if (n_total % bn) continue;
    const int n_blocks = n_total / bn;
    for (int mi = 0; mi < 3; ++mi) {
      const int mt = mts[mi];
      if (mt * bn > 256) continue;
      int box_rows, boxes;
      split_boxes(mt * 128 + 2 * halo, &box_rows, &boxes);
      const long long a_stage = (long long)box_rows * boxes * 128;
      const long long b_stage = (long long)bn * 128;
      for (int res = 0; res < 2; ++res) {
        if (res && n_blocks != 1) continue;
        int a_st, b_st;
        if (res) {
          b_st = 9 * chunks;
          if (b_st > CF_MAX_B_STAGES) continue;
          const long long left = budget - b_st * b_stage;
          if (left < 2 * a_stage) continue;
          a_st = (int)(left / a_stage);
          if (a_st > CF_MAX_A_STAGES) a_st = CF_MAX_A_STAGES;
        } else {
          a_st = 2;
          long long left = budget - a_st * a_stage;
          if (left < 3 * b_stage) continue;
          b_st = (int)(left / b_stage);
          if (b_st > 16) b_st = 16;
          left -= b_st * b_stage;
          if (left >= a_stage) a_st = 3;
        }
        const int m_tiles = (total_rows + mt * 128 - 1) / (mt * 128);
        const long long tiles = (long long)m_tiles * n_blocks;
        const int active = tiles < sms ? (int)tiles : sms;
        const double bw = 6300.0 / active < 64.0 ? 6300.0 / active : 64.0;
        const double mma = 9.0 * chunks * mt * bn * 2.0;
        const double bytes = (double)chunks * a_stage + (res ? 0.0 : 9.0 * chunks * b_stage);
        const double tile_clk = (mma > bytes / bw ? mma : bytes / bw) + 600.0;  // + pipeline fill / epilogue tail
        const long long rounds = (tiles + active - 1) / active;
        const double cost = rounds * tile_clk;
        if (cost < best.cost * 0.98) {
          best = FlatShape{mt, bn, res, a_st, b_st, box_rows, boxes, cost};
        }
      }
    }
  }
  // measurement aid: CILRS_FLAT_SHAPE="mt,block_n,resident,a_stages,b_stages" overrides the model (0 stages = keep the model's)
  if (const char* env = getenv("CILRS_FLAT_SHAPE")) {
    int mt = 0, bn = 0, res = 0, a_st = 0, b_st = 0;
    if (sscanf(env, "%d,%d,%d,%d,%d", &mt, &bn, &res, &a_st, &b_st) >= 3 && mt >= 1 && bn >= 64 && n_total % bn == 0 && mt * bn <= 512) {
      int box_rows, boxes;
      split_boxes(mt * 128 + 2 * halo, &box_rows, &boxes);
      const long long a_stage = (long long)box_rows * boxes * 128, b_stage = (long long)bn * 128;
      if (res) b_st = 9 * chunks;
      if (a_st < 1) a_st = 2;
      if (b_st < 1) b_st = (int)((budget - a_st * a_stage) / b_stage);
      if (b_st > CF_MAX_B_STAGES) b_st = CF_MAX_B_STAGES;
      if (a_st <= CF_MAX_A_STAGES && b_st >= 1 && a_st * a_stage + b_st * b_stage <= budget && (!res || n_total == bn))
        best = FlatShape{mt, bn, res, a_st, b_st, box_rows, boxes, 0.0};
    }
  }
  if (getenv("CILRS_FLAT_DEBUG"))
    fprintf(stderr, "[cilrs flat] rows=%d K=%d N=%d Wp=%d -> mt=%d block_n=%d resident=%d a_stages=%d b_stages=%d a_box=%dx%d cost=%.0f\n",
            total_rows, k_channels, n_total, g.Wp, best.mt, best.block_n, best.resident, best.a_stages, best.b_stages, best.a_boxes,
            best.a_box_rows, best.cost);
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
  const FlatShape sh = choose_flat_shape(p->total_rows, k_channels, n_total, g);
  if (sh.cost >= 1e30) return ERR_UNSUPPORTED;
  p->mt = sh.mt; p->block_n = sh.block_n; p->n_blocks = n_total / sh.block_n; p->n_total = n_total;
  p->chunks = k_channels / 64;
  p->halo = g.Wp + 1;
  p->a_box_rows = sh.a_box_rows; p->a_boxes = sh.a_boxes;
  p->a_stages = sh.a_stages; p->b_stages = sh.b_stages; p->b_resident = sh.resident;
  int sets = 512 / (sh.mt * sh.block_n);
  p->acc_sets = sets > CF_MAX_ACC ? CF_MAX_ACC : sets;
  p->m_tiles = (p->total_rows + sh.mt * 128 - 1) / (sh.mt * 128);
  flat_taps(p, dgrad);
  p->flags = flags;
  p->out = (__nv_bfloat16*)out;
  int st = encode_2d_map(&p->tmA, x, k_channels, p->total_rows, 64, p->a_box_rows);
  if (st) return st;
  return encode_2d_map(&p->tmB, w, k_channels, 9 * n_total, 64, p->block_n);
}

int flat_conv_grid(const FlatConvParams* p) {
  const long long total = (long long)p->m_tiles * p->n_blocks;
  return total < sm_count() ? (int)total : sm_count();
}

int launch_flat_conv(const FlatConvParams* p, cudaStream_t s) {
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(conv_flat_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, CG_SMEM_TOTAL);
    if (e != cudaSuccess) return cuda_status(e);
    attr_set = true;
  }
  if ((p->flags & (CF_STATS | CF_BNBWD)) && (!p->partials || !p->counter)) return ERR_INVALID;
  conv_flat_kernel<<<flat_conv_grid(p), CF_THREADS, CG_SMEM_TOTAL, s>>>(*p); ++g_cilrs_launches;
  return cuda_status(cudaGetLastError());
}

// ------------------------------------------------------------------------------------------------
// wgrad
// ------------------------------------------------------------------------------------------------
int build_wgrad_flat(WgradFlatParams* p, int batch, const PadGeom& g, int cin, int cout, const void* dy, const void* x, float* dw) {
  if (batch < 1 || g.H < 1 || g.W < 1 || g.Hp < g.H || g.Wp <= g.W || g.Wp < 3) return ERR_INVALID;
  if (cin % 64 || cout % 64 || cin < 64 || cout < 64) return ERR_UNSUPPORTED;
  memset(p, 0, sizeof(*p));
  p->total_rows = flat_total_rows(batch, g);
  p->k_tiles = (p->total_rows + 127) / 128;
  p->co_blocks = (cout + 127) / 128;
  p->m_halves = cout >= 128 ? 2 : 1;
  p->ci_chunks = cin / 64;
  p->tap_groups = 2;
  p->group_first[0] = 0; p->group_count[0] = 5;
  p->group_first[1] = 5; p->group_count[1] = 4;
  for (int r = 0; r < 3; ++r)
    for (int s = 0; s < 3; ++s) p->tap_shift[r * 3 + s] = (r - 1) * g.Wp + (s - 1);
  split_boxes(128 + g.Wp + 1, &p->x_box_rows, &p->x_boxes);
  const int base = p->co_blocks * p->ci_chunks * p->tap_groups;
  int z = sm_count() / base;
  if (z < 1) z = 1;
  if (z > p->k_tiles) z = p->k_tiles;
  p->split_z = z;
  const long long stage = 2LL * WG_SLAB + (long long)p->x_boxes * p->x_box_rows * 128;
  int st = (int)((CG_SMEM_TOTAL - 1024 - 256) / stage);
  if (st > WF_MAX_STAGES) st = WF_MAX_STAGES;
  if (st < 2) return ERR_UNSUPPORTED;
  p->num_stages = st;
  p->cout = cout; p->cin = cin;
  p->grad = dw;
  int e = encode_2d_map(&p->tmDY, dy, cout, p->total_rows, 64, 128);
  if (e) return e;
  return encode_2d_map(&p->tmX, x, cin, p->total_rows, 64, p->x_box_rows);
}

int launch_wgrad_flat(const WgradFlatParams* p, cudaStream_t s) {
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(wgrad_flat_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, CG_SMEM_TOTAL);
    if (e != cudaSuccess) return cuda_status(e);
    attr_set = true;
  }
  const int grid = p->co_blocks * p->ci_chunks * p->tap_groups * p->split_z;
  wgrad_flat_kernel<<<grid, WF_THREADS, CG_SMEM_TOTAL, s>>>(*p); ++g_cilrs_launches;
  return cuda_status(cudaGetLastError());
}

}  // namespace cilrs

using namespace cilrs;

extern "C" {

long long cilrs_flat_rows(int batch, int H, int W) { return (long long)batch * (H + 1) * (W + 1); }

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
  if ((flags & CF_STATS) && (flags & CF_BNBWD)) return ERR_INVALID;
  if ((flags & CF_SCALE_BIAS) && (!a->scale || !a->bias)) return ERR_INVALID;
  if ((flags & CF_RESIDUAL) && !a->residual) return ERR_INVALID;
  if ((flags & CF_MASK) && !a->mask) return ERR_INVALID;
  if ((flags & CF_STATS) && (!a->gamma || !a->beta || !a->running_mean || !a->running_var || !a->vec)) return ERR_INVALID;
  if ((flags & CF_BNBWD) && (!a->y1 || !a->vec1 || !a->bred1)) return ERR_INVALID;
  if ((flags & CF_BNBWD2) && (!a->y2 || !a->vec2 || !a->bred2)) return ERR_INVALID;
  if ((flags & (CF_STATS | CF_BNBWD)) && (!a->partials_ws || !a->counter_ws)) return ERR_INVALID;
  int st = build_flat_conv(&p, a->batch, g, a->in_c, a->out_c, a->dgrad, a->x, a->w, a->y, flags);
  if (st) return st;
  p.residual = (const __nv_bfloat16*)a->residual; p.mask = (const __nv_bfloat16*)a->mask; p.scale = a->scale; p.bias = a->bias;
  p.partials = a->partials_ws; p.counter = a->counter_ws;
  p.gamma = a->gamma; p.beta = a->beta; p.running_mean = a->running_mean; p.running_var = a->running_var;
  p.nbt = a->num_batches_tracked; p.vec = a->vec; p.count = (double)a->batch * a->H * a->W; p.momentum = a->momentum; p.eps = a->eps;
  p.update_running = a->update_running;
  p.y1 = (const __nv_bfloat16*)a->y1; p.stat1 = a->vec1; p.bred1 = a->bred1; p.dgamma1 = a->dgamma1; p.dbeta1 = a->dbeta1;
  p.y2 = (const __nv_bfloat16*)a->y2; p.stat2 = a->vec2; p.bred2 = a->bred2; p.dgamma2 = a->dgamma2; p.dbeta2 = a->dbeta2;
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

int cilrs_wgrad_flat(int batch, int H, int W, int in_c, int out_c, const void* dy, const void* x, float* dw_oihw, void* stream) {
  if (!dy || !x || !dw_oihw) return ERR_INVALID;
  const PadGeom g{H, W, H + 1, W + 1};
  WgradFlatParams p;
  int st = build_wgrad_flat(&p, batch, g, in_c, out_c, dy, x, dw_oihw);
  if (st) return st;
  return launch_wgrad_flat(&p, (cudaStream_t)stream);
}

}  // extern "C"
