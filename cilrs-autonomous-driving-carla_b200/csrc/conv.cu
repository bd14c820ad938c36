// Host side of the tcgen05 convolution kernels + their C-ABI entry points (include/cilrs_b200.h).
#include "conv_gemm.cuh"
#include "wgrad_gemm.cuh"
#include "conv_params.h"
#include "conv_host.h"
#include <stdlib.h>
#include <cudaTypedefs.h>
#include <string.h>

namespace cilrs {

// ------------------------------------------------------------------------------------------------
// driver entry point for tensor-map encoding, fetched through the runtime (no link-time libcuda dependency)
// ------------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn get_encode() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* f = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = (EncodeTiledFn)f;
  }
  return fn;
}

int encode_nhwc_map(CUtensorMap* m, const void* base, int C, int W, int H, int N, long long sw_b, long long sh_b,
                    long long sn_b, int box_c, int bw, int bh, int bn, int sw, int sh) {
  EncodeTiledFn enc = get_encode();
  if (!enc) return ERR_DRIVER;
  cuuint64_t gdim[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)N};
  cuuint64_t gstr[3] = {(cuuint64_t)sw_b, (cuuint64_t)sh_b, (cuuint64_t)sn_b};
  cuuint32_t box[4] = {(cuuint32_t)box_c, (cuuint32_t)((bw - 1) * sw + 1), (cuuint32_t)((bh - 1) * sh + 1), (cuuint32_t)bn};
  cuuint32_t estr[4] = {1, (cuuint32_t)sw, (cuuint32_t)sh, 1};
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, (void*)base, gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? OK : ERR_INVALID;
}

int encode_2d_map(CUtensorMap* m, const void* base, int inner, int rows, int box_inner, int box_rows) {
  EncodeTiledFn enc = get_encode();
  if (!enc) return ERR_DRIVER;
  cuuint64_t gdim[2] = {(cuuint64_t)inner, (cuuint64_t)rows};
  cuuint64_t gstr[1] = {(cuuint64_t)inner * 2};
  cuuint32_t box[2] = {(cuuint32_t)box_inner, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, (void*)base, gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? OK : ERR_INVALID;
}

int conv_out_dim(int in, int k, int stride, int pad) { return (in + 2 * pad - k) / stride + 1; }

static int num_sms() {
  static int n = 0;
  if (!n) {
    int dev = 0;
    cudaGetDevice(&dev);
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
  }
  return n;
}

BoxShape choose_box(int ow, int oh, int batch) {
  BoxShape best{};
  double best_eff = -1.0;
  for (int kw = 1; kw <= ow; ++kw) {
    const int bw = (ow + kw - 1) / kw;
    if (bw > 128) continue;
    if (kw > 1 && bw == (ow + kw - 2) / (kw - 1)) continue;
    for (int kh = 1; kh <= oh; ++kh) {
      const int bh = (oh + kh - 1) / kh;
      if (bw * bh > 128) continue;
      if (kh > 1 && bh == (oh + kh - 2) / (kh - 1)) continue;
      int bn = 128 / (bw * bh);
      if (bn > batch) bn = batch;
      if (bn < 1) continue;
      // keep tiles spatially compact (>= 8 pixels per image) so a box is a few long runs, not 128 scattered lines
      const int min_sp = ow * oh < 8 ? ow * oh : 8;
      if (bw * bh < min_sp) continue;
      const int tw = (ow + bw - 1) / bw, th = (oh + bh - 1) / bh, tn = (batch + bn - 1) / bn;
      const double eff = (double)ow * oh * batch / ((double)tw * th * tn * 128.0);
      if (eff > best_eff + 1e-9) {
        best_eff = eff;
        best = BoxShape{bw, bh, bn, tw, th, tn};
      }
    }
  }
  return best;
}

static int pick_block_n(int n_total, int m_tiles) {
  const int sms = num_sms();
  const int cands[3] = {256, 128, 64};
  for (int i = 0; i < 3; ++i) {
    const int bn = cands[i];
    if (n_total % bn) continue;
    if (m_tiles * (n_total / bn) >= sms - 8) return bn;
  }
  return n_total >= 128 ? 128 : 64;
}

static int stages_for(int block_n) {
  const int avail = CG_SMEM_TOTAL - 1024 - CG_STAGING_BYTES - 1024 - 2048 - 4096 - 256;
  int s = avail / (CG_A_BYTES + block_n * 128);
  if (s > CG_MAX_STAGES) s = CG_MAX_STAGES;
  return s;
}

static void set_tiles(ConvGemmParams* p, const BoxShape& b) {
  p->tiles_w = b.tiles_w; p->tiles_h = b.tiles_h; p->tiles_n = b.tiles_n;
  p->BW = b.BW; p->BH = b.BH; p->BN = b.BN;
}

static int check_desc(const cilrs_conv_desc* d) {
  if (!d) return ERR_INVALID;
  if (d->batch < 1 || d->in_h < 1 || d->in_w < 1) return ERR_INVALID;
  if (d->in_c % 64 || d->out_c % 64 || d->in_c < 64 || d->out_c < 64) return ERR_UNSUPPORTED;
  const bool k3 = d->kh == 3 && d->kw == 3 && d->pad == 1;
  const bool k1 = d->kh == 1 && d->kw == 1 && d->pad == 0;
  if (!k3 && !k1) return ERR_UNSUPPORTED;
  if (d->stride != 1 && d->stride != 2) return ERR_UNSUPPORTED;
  return OK;
}

// pixel pitches of a tensor: dense [H][W] or the padded-flat layout of conv_params.h
static void pitches(const PadGeom* g, int H, int W, int* hp, int* wp) {
  *hp = g ? g->Hp : H;
  *wp = g ? g->Wp : W;
}

int build_fprop(ConvGemmParams* p, const cilrs_conv_desc* d, const void* x, const void* w, void* y, const float* scale,
                const float* bias, const void* residual, float* stats, int flags, const PadGeom* gin, const PadGeom* gout) {
  int st = check_desc(d);
  if (st) return st;
  memset(p, 0, sizeof(*p));
  const int OH = conv_out_dim(d->in_h, d->kh, d->stride, d->pad), OW = conv_out_dim(d->in_w, d->kw, d->stride, d->pad);
  int ihp, iwp, ohp, owp;
  pitches(gin, d->in_h, d->in_w, &ihp, &iwp);
  pitches(gout, OH, OW, &ohp, &owp);
  const BoxShape b = choose_box(OW, OH, d->batch);
  set_tiles(p, b);
  p->block_n = pick_block_n(d->out_c, b.m_tiles());
  p->n_blocks = d->out_c / p->block_n;
  p->num_stages = stages_for(p->block_n);
  p->num_taps = d->kh * d->kw;
  p->chunks = d->in_c / 64;
  p->in_sw = p->in_sh = d->stride;
  for (int r = 0; r < d->kh; ++r)
    for (int s = 0; s < d->kw; ++s) {
      const int t = r * d->kw + s;
      p->tap_dh[t] = (int8_t)(r - d->pad);
      p->tap_dw[t] = (int8_t)(s - d->pad);
      p->tap_slab[t] = (int16_t)t;
    }
  p->slab_rows = d->out_c;
  p->n_img = d->batch; p->oh = OH; p->ow = OW;
  p->out_sw = d->out_c; p->out_sh = (long long)owp * d->out_c; p->out_sn = (long long)ohp * owp * d->out_c; p->out_off = 0;
  p->n_total = d->out_c;
  p->out = (__nv_bfloat16*)y; p->residual = (const __nv_bfloat16*)residual; p->scale = scale; p->bias = bias; p->stats = stats;
  p->flags = flags;
  const long long cb = 2LL * d->in_c;
  st = encode_nhwc_map(&p->tmA[0], x, d->in_c, d->in_w, d->in_h, d->batch, cb, cb * iwp, cb * iwp * ihp, 64, b.BW,
                       b.BH, b.BN, d->stride, d->stride);
  if (st) return st;
  p->tmA[1] = p->tmA[2] = p->tmA[3] = p->tmA[0];
  st = encode_2d_map(&p->tmB[0], w, d->in_c, p->num_taps * d->out_c, 64, p->block_n);
  if (st) return st;
  p->tmB[1] = p->tmB[0];
  return OK;
}

// stem geometry: image 88x200, 7x7/2 pad 3 -> 44x100; space-to-depth input [N,47,103,16]
static const int STEM_SH = 47, STEM_SW = 103, STEM_OH = 44, STEM_OW = 100;

static int encode_stem_map(CUtensorMap* m, const void* x_s2d, int batch, const BoxShape& b) {
  // overlapping windows: pixel X exposes the 64 contiguous values of s2d pixels X..X+3 (4 horizontal taps x 16)
  return encode_nhwc_map(m, x_s2d, 64, STEM_OW, STEM_SH, batch, 32, 32LL * STEM_SW, 32LL * STEM_SW * STEM_SH, 64, b.BW, b.BH,
                         b.BN, 1, 1);
}

int build_stem_fprop(ConvGemmParams* p, int batch, const void* x_s2d, const void* w, void* y, const float* scale,
                     const float* bias, float* stats, int flags) {
  if (batch < 1) return ERR_INVALID;
  memset(p, 0, sizeof(*p));
  const BoxShape b = choose_box(STEM_OW, STEM_OH, batch);
  set_tiles(p, b);
  p->block_n = 64; p->n_blocks = 1; p->num_stages = stages_for(64);
  p->num_taps = 4; p->chunks = 1; p->in_sw = p->in_sh = 1;
  for (int r = 0; r < 4; ++r) { p->tap_dh[r] = (int8_t)r; p->tap_dw[r] = 0; p->tap_slab[r] = (int16_t)r; }
  p->slab_rows = 64;
  p->n_img = batch; p->oh = STEM_OH; p->ow = STEM_OW;
  p->out_sw = 64; p->out_sh = 64LL * STEM_OW; p->out_sn = 64LL * STEM_OW * STEM_OH; p->out_off = 0;
  p->n_total = 64;
  p->out = (__nv_bfloat16*)y; p->scale = scale; p->bias = bias; p->stats = stats; p->flags = flags;
  int st = encode_stem_map(&p->tmA[0], x_s2d, batch, b);
  if (st) return st;
  p->tmA[1] = p->tmA[2] = p->tmA[3] = p->tmA[0];
  st = encode_2d_map(&p->tmB[0], w, 64, 4 * 64, 64, 64);
  if (st) return st;
  p->tmB[1] = p->tmB[0];
  return OK;
}

int build_dgrad(ConvGemmParams* p, const cilrs_conv_desc* d, int ph, int pw, const void* dy, const void* wd, void* dx,
                const void* residual, const void* dy2, const void* w2d, const PadGeom* gin, const PadGeom* gout) {
  int st = check_desc(d);
  if (st) return st;
  memset(p, 0, sizeof(*p));
  const int OH = conv_out_dim(d->in_h, d->kh, d->stride, d->pad), OW = conv_out_dim(d->in_w, d->kw, d->stride, d->pad);
  int ihp, iwp, ohp, owp;
  pitches(gin, d->in_h, d->in_w, &ihp, &iwp);
  pitches(gout, OH, OW, &ohp, &owp);
  const int s = d->stride;
  // output of this launch: input pixels (h, w) with h % s == ph, w % s == pw
  const int TH = (d->in_h - ph + s - 1) / s, TW = (d->in_w - pw + s - 1) / s;
  if (TH <= 0 || TW <= 0) return ERR_INVALID;
  const BoxShape b = choose_box(TW, TH, d->batch);
  set_tiles(p, b);
  p->block_n = pick_block_n(d->in_c, b.m_tiles());
  p->n_blocks = d->in_c / p->block_n;
  p->num_stages = stages_for(p->block_n);
  p->chunks = d->out_c / 64;
  p->in_sw = p->in_sh = 1;  // dy is read densely
  int nt = 0;
  for (int r = 0; r < d->kh; ++r) {
    // h = s*ho + r - pad  ->  ho = (h + pad - r)/s with h = s*j + ph
    const int num_h = ph + d->pad - r;
    if (((num_h % s) + s) % s) continue;
    for (int c = 0; c < d->kw; ++c) {
      const int num_w = pw + d->pad - c;
      if (((num_w % s) + s) % s) continue;
      p->tap_dh[nt] = (int8_t)(num_h >= 0 ? num_h / s : -((-num_h) / s));
      p->tap_dw[nt] = (int8_t)(num_w >= 0 ? num_w / s : -((-num_w) / s));
      p->tap_slab[nt] = (int16_t)(r * d->kw + c);
      p->tap_a[nt] = 0; p->tap_b[nt] = 0;
      ++nt;
    }
  }
  if (dy2 && w2d) {
    if (!(s == 2 && ph == 0 && pw == 0)) return ERR_INVALID;
    p->tap_dh[nt] = 0; p->tap_dw[nt] = 0; p->tap_slab[nt] = 0; p->tap_a[nt] = 1; p->tap_b[nt] = 1;
    ++nt;
  }
  if (nt == 0) return ERR_UNSUPPORTED;  // (1x1 stride-2 has no taps on odd parities; callers zero those)
  p->num_taps = nt;
  p->slab_rows = d->in_c;
  p->n_img = d->batch; p->oh = TH; p->ow = TW;
  const long long C = d->in_c;
  p->out_sw = s * C; p->out_sh = (long long)s * iwp * C; p->out_sn = (long long)ihp * iwp * C;
  p->out_off = ((long long)ph * iwp + pw) * C;
  p->n_total = d->in_c;
  p->out = (__nv_bfloat16*)dx; p->residual = (const __nv_bfloat16*)residual;
  p->flags = residual ? CG_RESIDUAL : 0;
  const long long cb = 2LL * d->out_c;
  st = encode_nhwc_map(&p->tmA[0], dy, d->out_c, OW, OH, d->batch, cb, cb * owp, cb * owp * ohp, 64, b.BW, b.BH, b.BN, 1, 1);
  if (st) return st;
  p->tmA[1] = p->tmA[0];
  if (dy2) {
    st = encode_nhwc_map(&p->tmA[1], dy2, d->out_c, OW, OH, d->batch, cb, cb * owp, cb * owp * ohp, 64, b.BW, b.BH, b.BN, 1, 1);
    if (st) return st;
  }
  p->tmA[2] = p->tmA[3] = p->tmA[0];
  st = encode_2d_map(&p->tmB[0], wd, d->out_c, d->kh * d->kw * d->in_c, 64, p->block_n);
  if (st) return st;
  p->tmB[1] = p->tmB[0];
  if (w2d) {
    st = encode_2d_map(&p->tmB[1], w2d, d->out_c, d->in_c, 64, p->block_n);
    if (st) return st;
  }
  return OK;
}

static int wgrad_stages(int g) {
  const int avail = CG_SMEM_TOTAL - 1024 - 256;
  int s = avail / ((2 + g) * WG_SLAB);
  if (s > WG_MAX_STAGES) s = WG_MAX_STAGES;
  return s;
}

int build_wgrad(WgradParams* p, const cilrs_conv_desc* d, const void* dy, const void* x, float* dw, const PadGeom* gin,
                const PadGeom* gout) {
  int st = check_desc(d);
  if (st) return st;
  memset(p, 0, sizeof(*p));
  const int OH = conv_out_dim(d->in_h, d->kh, d->stride, d->pad), OW = conv_out_dim(d->in_w, d->kw, d->stride, d->pad);
  int ihp, iwp, ohp, owp;
  pitches(gin, d->in_h, d->in_w, &ihp, &iwp);
  pitches(gout, OH, OW, &ohp, &owp);
  const BoxShape b = choose_box(OW, OH, d->batch);
  p->tiles_w = b.tiles_w; p->tiles_h = b.tiles_h; p->tiles_n = b.tiles_n;
  p->BW = b.BW; p->BH = b.BH; p->BN = b.BN;
  p->in_sw = p->in_sh = d->stride;
  p->co_blocks = (d->out_c + 127) / 128;
  p->m_halves = d->out_c >= 128 ? 2 : 1;
  p->ci_chunks = d->in_c / 64;
  p->num_taps = d->kh * d->kw;
  p->g = p->num_taps == 9 ? 3 : 1;
  p->tap_groups = p->num_taps / p->g;
  for (int r = 0; r < d->kh; ++r)
    for (int s = 0; s < d->kw; ++s) {
      const int t = r * d->kw + s;
      p->tap_dh[t] = (int8_t)(r - d->pad);
      p->tap_dw[t] = (int8_t)(s - d->pad);
      p->tap_id[t] = (int16_t)t;
    }
  const int base = p->co_blocks * p->ci_chunks * p->tap_groups;
  // CTAs of the launch = base * z. Three quarters of the SMs: these launches (stride-2 and 1x1 convolutions at the stage
  // transitions) run beside the BatchNorm-backward reduce / apply of the transition, which needs SMs of its own (measured at
  // batch 128: 2.669 ms per step with 148, 2.653 with 111, 2.660 with 74 or 48). CILRS_WGRAD_GEMM_CTAS overrides the target.
  static int target = -1;
  if (target < 0) { const char* env = getenv("CILRS_WGRAD_GEMM_CTAS"); target = env ? atoi(env) : 0; if (target < 1) target = num_sms() * 3 / 4; }
  int z = target / base;
  if (z < 1) z = 1;
  if (z > b.m_tiles()) z = b.m_tiles();
  p->split_z = z;
  p->num_stages = wgrad_stages(p->g);
  p->cout = d->out_c; p->cin = d->in_c;
  p->ci_stride = d->kh * d->kw; p->co_stride = d->in_c * d->kh * d->kw;
  p->col_mode = WG_COL_REGULAR;
  p->grad = dw;
  const long long cb = 2LL * d->out_c;
  st = encode_nhwc_map(&p->tmDY, dy, d->out_c, OW, OH, d->batch, cb, cb * owp, cb * owp * ohp, 64, b.BW, b.BH, b.BN, 1, 1);
  if (st) return st;
  const long long xb = 2LL * d->in_c;
  st = encode_nhwc_map(&p->tmX, x, d->in_c, d->in_w, d->in_h, d->batch, xb, xb * iwp, xb * iwp * ihp, 64, b.BW, b.BH,
                       b.BN, d->stride, d->stride);
  return st;
}

int build_stem_wgrad(WgradParams* p, int batch, const void* dy, const void* x_s2d, float* dw) {
  if (batch < 1) return ERR_INVALID;
  memset(p, 0, sizeof(*p));
  const BoxShape b = choose_box(STEM_OW, STEM_OH, batch);
  p->tiles_w = b.tiles_w; p->tiles_h = b.tiles_h; p->tiles_n = b.tiles_n;
  p->BW = b.BW; p->BH = b.BH; p->BN = b.BN;
  p->in_sw = p->in_sh = 1;
  p->co_blocks = 1; p->m_halves = 1; p->ci_chunks = 1;
  p->num_taps = 4; p->g = 4; p->tap_groups = 1;
  for (int r = 0; r < 4; ++r) { p->tap_dh[r] = (int8_t)r; p->tap_dw[r] = 0; p->tap_id[r] = (int16_t)r; }
  int z = num_sms();
  if (z > b.m_tiles()) z = b.m_tiles();
  p->split_z = z;
  p->num_stages = wgrad_stages(4);
  p->cout = 64; p->cin = 64; p->ci_stride = 0; p->co_stride = 147;
  p->col_mode = WG_COL_CONV1_S2D;
  p->grad = dw;
  int st = encode_nhwc_map(&p->tmDY, dy, 64, STEM_OW, STEM_OH, batch, 128, 128LL * STEM_OW, 128LL * STEM_OW * STEM_OH, 64, b.BW,
                           b.BH, b.BN, 1, 1);
  if (st) return st;
  return encode_stem_map(&p->tmX, x_s2d, batch, b);
}

int conv_gemm_grid(const ConvGemmParams* p) {
  const int total = p->tiles_w * p->tiles_h * p->tiles_n * p->n_blocks;
  return total < num_sms() ? total : num_sms();
}

int launch_conv_gemm(const ConvGemmParams* p, cudaStream_t s) {
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(conv_gemm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, CG_SMEM_TOTAL);
    if (e != cudaSuccess) return cuda_status(e);
    attr_set = true;
  }
  const int grid = conv_gemm_grid(p);
  ++g_cilrs_launches;
  return cuda_status(launch_pdl(conv_gemm_kernel, dim3(grid), dim3(CG_THREADS), CG_SMEM_TOTAL, s, *p));
}

// `count` (<= 4) independent problems in one launch (conv_gemm_multi_kernel): the parity plans of a stride-2 dgrad
int launch_conv_gemm_multi(const ConvGemmParams* plans, int count, cudaStream_t s) {
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(conv_gemm_multi_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, CG_SMEM_TOTAL);
    if (e != cudaSuccess) return cuda_status(e);
    attr_set = true;
  }
  if (count < 1 || count > 4) return ERR_INVALID;
  ConvGemmParams4 pp;
  int most = 1;
  for (int i = 0; i < 4; ++i) {
    pp.p[i] = plans[i < count ? i : 0];
    if (i < count) {
      if (plans[i].flags & CG_STATS) return ERR_INVALID;  // the per-CTA statistics slots are indexed by blockIdx.x alone
      const int total = plans[i].tiles_w * plans[i].tiles_h * plans[i].tiles_n * plans[i].n_blocks;
      most = total > most ? total : most;
    }
  }
  int per = num_sms() / count;
  if (per < 1) per = 1;
  const int gx = most < per ? most : per;
  ++g_cilrs_launches;
  return cuda_status(launch_pdl(conv_gemm_multi_kernel, dim3(gx, count), dim3(CG_THREADS), CG_SMEM_TOTAL, s, pp));
}

int launch_wgrad(const WgradParams* p, cudaStream_t s) {
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(wgrad_gemm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, CG_SMEM_TOTAL);
    if (e != cudaSuccess) return cuda_status(e);
    attr_set = true;
  }
  const int grid = p->co_blocks * p->ci_chunks * p->tap_groups * p->split_z;
  ++g_cilrs_launches;
  return cuda_status(launch_pdl(wgrad_gemm_kernel, dim3(grid), dim3(WG_THREADS), CG_SMEM_TOTAL, s, *p));
}

// ------------------------------------------------------------------------------------------------
// weight packing (fp32 OIHW master -> bf16 operand layouts)
// ------------------------------------------------------------------------------------------------
__global__ void pack_weight_kernel(const float* __restrict__ w, __nv_bfloat16* __restrict__ wf, __nv_bfloat16* __restrict__ wd,
                                   int cout, int cin, int kk) {
  const long long total = (long long)cout * cin * kk;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    // i enumerates the fprop layout [t][co][ci] so the bf16 stores are coalesced
    const int ci = (int)(i % cin);
    const int co = (int)((i / cin) % cout);
    const int t = (int)(i / ((long long)cin * cout));
    const __nv_bfloat16 v = __float2bfloat16(w[((long long)co * cin + ci) * kk + t]);
    if (wf) wf[i] = v;
    if (wd) wd[((long long)t * cin + ci) * cout + co] = v;
  }
}

// all layers in one launch: a CTA transposes a 32(co) x 32(ci) x kk block through shared memory, so the fp32 reads
// (runs of 32*kk floats) and both bf16 writes (32 consecutive ci resp. co) are coalesced
__global__ void __launch_bounds__(256) pack_all_kernel(const float* __restrict__ params, const PackJob* __restrict__ jobs, int njobs,
                                                       int block_base) {
  __shared__ float tile[32][32 * 9 + 1];
  const int bidx = (int)blockIdx.x + block_base;   // (a launch may cover a sub-range of the job table: one backward part)
  // the job of this block: the last one whose first block is <= bidx (first_block increases). One round trip for up to 64 jobs
  // (every warp looks for itself) instead of a chain of dependent loads
  int j = 0;
  if (njobs <= 64) {
    const int l = threadIdx.x & 31;
    const int fb0 = l < njobs ? jobs[l].first_block : 0x7FFFFFFF, fb1 = l + 32 < njobs ? jobs[l + 32].first_block : 0x7FFFFFFF;
    j = __popc(__ballot_sync(0xFFFFFFFFu, fb0 <= bidx)) + __popc(__ballot_sync(0xFFFFFFFFu, fb1 <= bidx)) - 1;
    if (j < 0) j = 0;
  } else {
    while (j + 1 < njobs && bidx >= jobs[j + 1].first_block) ++j;
  }
  const PackJob jb = jobs[j];
  const int t_idx = bidx - jb.first_block;
  const int ci_tiles = jb.cin >> 5;
  const int co0 = (t_idx / ci_tiles) * 32, ci0 = (t_idx % ci_tiles) * 32;
  const int kk = jb.kk, run = 32 * kk;
  // (ncu, round 2: the first version spent 52 instructions per weight - a division per element in the load loop, 64-bit index
  //  arithmetic in the store loop - and ran at 50 % issue utilisation, 60 us for 85 MB in + 85 MB out. Loops without divisions:)
  // load: warp w copies the rows co = w, w + 8, ... (runs of 32 * kk consecutive floats, coalesced)
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const float* w = params + jb.w_off + ((size_t)co0 * jb.cin + ci0) * kk;
  const unsigned int row_pitch = (unsigned int)jb.cin * kk;
  for (int co = warp; co < 32; co += 8) {
    const float* src = w + (size_t)co * row_pitch;
    for (int r = lane; r < run; r += 32) tile[co][r] = src[r];
  }
  __syncthreads();
  // store: thread (a2 = idx & 15, b = idx >> 4 & 31) writes the element pair (2 a2, 2 a2 + 1) of row b for every tap
  const int a = (threadIdx.x & 15) * 2, b0 = threadIdx.x >> 4;   // b0 in 0..15; rows b0 and b0 + 16
  const unsigned int f_pitch = (unsigned int)jb.cout * jb.cin;    // elements per tap in either layout
#pragma unroll 1
  for (int t = 0; t < kk; ++t) {
#pragma unroll
    for (int hb = 0; hb < 2; ++hb) {
      const int b = b0 + 16 * hb;
      // fprop layout [t][co][ci]: a = ci (fastest), b = co
      *reinterpret_cast<uint32_t*>(jb.wf + (size_t)t * f_pitch + (unsigned int)(co0 + b) * jb.cin + ci0 + a) =
          pack_bf16x2(tile[b][a * kk + t], tile[b][(a + 1) * kk + t]);
      // dgrad layout [t][ci][co]: a = co (fastest), b = ci
      *reinterpret_cast<uint32_t*>(jb.wd + (size_t)t * f_pitch + (unsigned int)(ci0 + b) * jb.cout + co0 + a) =
          pack_bf16x2(tile[a][b * kk + t], tile[a + 1][b * kk + t]);
    }
  }
}

int launch_pack_all(const float* params, const PackJob* jobs_dev, int njobs, int total_blocks, cudaStream_t s) {
  pack_all_kernel<<<total_blocks, 256, 0, s>>>(params, jobs_dev, njobs, 0); ++g_cilrs_launches;
  return cuda_status(cudaGetLastError());
}
// the blocks [block_lo, block_hi) of the job table only
int launch_pack_range(const float* params, const PackJob* jobs_dev, int njobs, int block_lo, int block_hi, cudaStream_t s) {
  if (block_hi <= block_lo) return OK;
  pack_all_kernel<<<block_hi - block_lo, 256, 0, s>>>(params, jobs_dev, njobs, block_lo); ++g_cilrs_launches;
  return cuda_status(cudaGetLastError());
}

__global__ void pack_stem_kernel(const float* __restrict__ w, __nv_bfloat16* __restrict__ wp) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;  // [r'][co][j]
  if (i >= 4 * 64 * 64) return;
  const int j = i & 63, co = (i >> 6) & 63, r = i >> 12;
  const int sp = j >> 4, dy = (j >> 3) & 1, dx = (j >> 2) & 1, c = j & 3;
  const int ky = 2 * r + dy, kx = 2 * sp + dx;
  float v = 0.f;
  if (ky < 7 && kx < 7 && c < 3) v = w[((co * 3 + c) * 7 + ky) * 7 + kx];
  wp[i] = __float2bfloat16(v);
}

}  // namespace cilrs

using namespace cilrs;

extern "C" {

size_t cilrs_conv_packed_weight_bytes(const cilrs_conv_desc* d) {
  if (check_desc(d)) return 0;
  return (size_t)d->kh * d->kw * d->in_c * d->out_c * 2;
}
int cilrs_conv_stats_tiles(const cilrs_conv_desc* d) {
  if (check_desc(d)) return 0;
  const int OH = conv_out_dim(d->in_h, d->kh, d->stride, d->pad), OW = conv_out_dim(d->in_w, d->kw, d->stride, d->pad);
  const int m_tiles = choose_box(OW, OH, d->batch).m_tiles();
  const int total = m_tiles * (d->out_c / pick_block_n(d->out_c, m_tiles));
  return total < num_sms() ? total : num_sms();  // one (sum, sumsq) partial per persistent CTA
}
size_t cilrs_conv_stats_bytes(const cilrs_conv_desc* d) { return (size_t)cilrs_conv_stats_tiles(d) * 2 * d->out_c * sizeof(float); }

int cilrs_conv_pack_weight(const cilrs_conv_desc* d, const float* w, void* wf, void* wd, void* stream) {
  int st = check_desc(d);
  if (st) return st;
  if (!w) return ERR_INVALID;
  const long long total = (long long)d->out_c * d->in_c * d->kh * d->kw;
  int blocks = (int)((total + 255) / 256);
  if (blocks > 2048) blocks = 2048;
  pack_weight_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(w, (__nv_bfloat16*)wf, (__nv_bfloat16*)wd, d->out_c, d->in_c,
                                                               d->kh * d->kw); ++g_cilrs_launches;
  return cuda_status(cudaGetLastError());
}

int cilrs_conv_fprop(const cilrs_conv_desc* d, const void* x, const void* w, void* y, const float* scale, const float* bias,
                     const void* residual, float* stats, int flags, void* stream) {
  if (!x || !w || !y) return ERR_INVALID;
  if ((flags & CILRS_EPI_STATS) && !stats) return ERR_INVALID;
  if ((flags & CILRS_EPI_SCALE_BIAS) && (!scale || !bias)) return ERR_INVALID;
  if ((flags & CILRS_EPI_RESIDUAL) && !residual) return ERR_INVALID;
  ConvGemmParams p;
  int st = build_fprop(&p, d, x, w, y, scale, bias, residual, stats, flags, nullptr, nullptr);
  if (st) return st;
  return launch_conv_gemm(&p, (cudaStream_t)stream);
}

int cilrs_conv_dgrad(const cilrs_conv_desc* d, const void* dy, const void* wd, void* dx, const void* residual, void* stream) {
  if (!dy || !wd || !dx) return ERR_INVALID;
  int st = check_desc(d);
  if (st) return st;
  for (int ph = 0; ph < d->stride; ++ph)
    for (int pw = 0; pw < d->stride; ++pw) {
      ConvGemmParams p;
      st = build_dgrad(&p, d, ph, pw, dy, wd, dx, residual, nullptr, nullptr, nullptr, nullptr);
      if (st == ERR_UNSUPPORTED && d->kh == 1) continue;  // 1x1/2: odd parities receive no gradient
      if (st) return st;
      st = launch_conv_gemm(&p, (cudaStream_t)stream);
      if (st) return st;
    }
  return OK;
}

int cilrs_conv_wgrad(const cilrs_conv_desc* d, const void* dy, const void* x, float* dw, void* stream) {
  if (!dy || !x || !dw) return ERR_INVALID;
  WgradParams p;
  int st = build_wgrad(&p, d, dy, x, dw, nullptr, nullptr);
  if (st) return st;
  return launch_wgrad(&p, (cudaStream_t)stream);
}

size_t cilrs_stem_packed_weight_bytes(void) { return 4 * 64 * 64 * 2; }
int cilrs_stem_stats_tiles(int batch) {
  if (batch < 1) return 0;
  const int t = choose_box(STEM_OW, STEM_OH, batch).m_tiles();
  return t < num_sms() ? t : num_sms();
}

int cilrs_stem_pack_weight(const float* w, void* wp, void* stream) {
  if (!w || !wp) return ERR_INVALID;
  pack_stem_kernel<<<64, 256, 0, (cudaStream_t)stream>>>(w, (__nv_bfloat16*)wp); ++g_cilrs_launches;
  return cuda_status(cudaGetLastError());
}
int cilrs_stem_fprop(int batch, const void* x, const void* w, void* y, const float* scale, const float* bias, float* stats,
                     int flags, void* stream) {
  if (!x || !w || !y) return ERR_INVALID;
  if ((flags & CILRS_EPI_STATS) && !stats) return ERR_INVALID;
  if ((flags & CILRS_EPI_SCALE_BIAS) && (!scale || !bias)) return ERR_INVALID;
  if (flags & CILRS_EPI_RESIDUAL) return ERR_INVALID;
  ConvGemmParams p;
  int st = build_stem_fprop(&p, batch, x, w, y, scale, bias, stats, flags);
  if (st) return st;
  return launch_conv_gemm(&p, (cudaStream_t)stream);
}
int cilrs_stem_wgrad(int batch, const void* dy, const void* x, float* dw, void* stream) {
  if (!dy || !x || !dw) return ERR_INVALID;
  WgradParams p;
  int st = build_stem_wgrad(&p, batch, dy, x, dw);
  if (st) return st;
  return launch_wgrad(&p, (cudaStream_t)stream);
}

}  // extern "C"
