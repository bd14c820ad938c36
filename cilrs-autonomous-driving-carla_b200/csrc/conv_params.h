// Kernel parameter blocks of the tcgen05 conv kernels (shared by the kernels in conv.cu and the network plan in model.cu).
#pragma once
#include <cuda.h>
#include <cuda_bf16.h>
#include <stdint.h>

namespace cilrs {

constexpr int CG_MAX_TAPS = 32;
constexpr int CG_BLOCK_M = 128;
constexpr int CG_A_BYTES = CG_BLOCK_M * 128;  // 128 pixel rows x 64 bf16
constexpr int CG_THREADS = 192;
constexpr int CG_MAX_STAGES = 8;
constexpr int CG_STAGING_BYTES = 2 * CG_A_BYTES;  // two 128x64 bf16 output chunks
constexpr int CG_SMEM_TOTAL = 227 * 1024;

enum ConvEpilogueFlags : int {
  CG_STATS = 1,       // write per-tile per-channel sum / sum-of-squares of the (bf16-rounded) output
  CG_SCALE_BIAS = 2,  // y = acc * scale[n] + bias[n]   (folded eval-mode BatchNorm)
  CG_RESIDUAL = 4,    // y += residual[pixel, n]
  CG_RELU = 8,        // y = max(y, 0)
};

struct ConvGemmParams {
  CUtensorMap tmA[4];
  CUtensorMap tmB[2];
  // tiling
  int tiles_w, tiles_h, tiles_n, n_blocks;
  int BW, BH, BN;
  int block_n;
  int num_taps, chunks;
  int in_sw, in_sh;
  int num_stages;
  int8_t tap_dw[CG_MAX_TAPS], tap_dh[CG_MAX_TAPS];
  int8_t tap_a[CG_MAX_TAPS], tap_b[CG_MAX_TAPS];
  int16_t tap_slab[CG_MAX_TAPS];
  int slab_rows;
  // output geometry (tile coordinates -> element offset)
  int n_img, oh, ow;
  long long out_sn, out_sh, out_sw, out_off;
  int n_total;
  __nv_bfloat16* out;
  const __nv_bfloat16* residual;
  const float* scale;
  const float* bias;
  float* stats;  // [m_tiles][2][n_total]
  int flags;
};

constexpr int WG_MAX_TAPS = 16;
constexpr int WG_THREADS = 192;
constexpr int WG_SLAB = 128 * 128;  // 128 pixel rows x 64 bf16
constexpr int WG_MAX_STAGES = 4;

enum WgradColMode : int { WG_COL_REGULAR = 0, WG_COL_CONV1_S2D = 1 };

struct WgradParams {
  CUtensorMap tmDY;  // (Cout, OW, OH, N), box (64, BW, BH, BN)
  CUtensorMap tmX;   // (Cin,  W,  H,  N), box (64, BW, BH, BN) with the conv stride as element stride
  int tiles_w, tiles_h, tiles_n;
  int BW, BH, BN;
  int in_sw, in_sh;
  int co_blocks, m_halves;   // M = 128 rows of the accumulator = m_halves x 64 output channels
  int ci_chunks;
  int tap_groups, g;         // g taps per CTA, N = 64 g
  int split_z;
  int num_stages;
  int8_t tap_dw[WG_MAX_TAPS], tap_dh[WG_MAX_TAPS];
  int16_t tap_id[WG_MAX_TAPS];  // position of the tap inside the kh*kw plane of the OIHW gradient
  int num_taps;
  int cout, cin;
  int co_stride, ci_stride;  // element strides of the fp32 gradient tensor
  int col_mode;
  float* grad;
};

}  // namespace cilrs
