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

// ------------------------------------------------------------------------------------------------------------------
// "padded-flat" kernels (3x3 stride-1 convs, the bulk of ResNet-34).
// Activations live in a padded NHWC layout: image b, row h, column w, channel c sits at flat pixel
//   f = (b*Hp + h)*Wp + w,   Hp = H + 1, Wp = W + 1,
// and the extra column / row of every image is all zeros. The left neighbour of column 0 is then the zero column of
// the previous row, the row above row 0 is the zero row of the previous image, and a 3x3 tap (dh, dw) is the constant
// flat shift dh*Wp + dw. One TMA load of a slab of (tile rows + 2*(Wp+1)) pixels therefore serves all nine taps: the
// taps are row-shifted UMMA descriptors into the same shared-memory slab (SWIZZLE_128B is a function of the absolute
// shared-memory address, so a descriptor may start at any 128-byte row; tools/umma_shift_test.cu).
// ------------------------------------------------------------------------------------------------------------------
struct PadGeom {
  int H, W, Hp, Wp;
};

constexpr int CF_THREADS = 384;       // warpgroup 0: TMA (warp 0), MMA (warp 1), two idle warps; warpgroups 1, 2: the two epilogue groups
constexpr int CF_MAX_A_STAGES = 4;
constexpr int CF_MAX_B_STAGES = 16;
constexpr int CF_MAX_ACC = 8;
constexpr int CF_STAGING_BYTES = 8 * 32 * 128;   // 32 rows x 64 bf16 per epilogue warp: staging of the TMA output stores

enum FlatConvFlags : int {
  CF_STATS = 1,       // per-channel sum / sum of squares of the bf16 output; the last CTA folds them into the BN vectors
  CF_SCALE_BIAS = 2,  // y = acc * scale[n] + bias[n]
  CF_RESIDUAL = 4,    // y += residual[f, n]
  CF_RELU = 8,        // y = max(y, 0)
  CF_MASK = 16,       // y = mask[f, n] > 0 ? y : 0          (ReLU backward, fused into the dgrad epilogue)
  CF_BNBWD = 32,      // per-channel sum(y), sum(y * xhat1) with xhat1 = (y1 - mean1) * rstd1   (BatchNorm backward reduce)
  CF_BNBWD2 = 64,     // ... and sum(y * xhat2) for a second BatchNorm fed by the same gradient (downsample branch)
  CF_DEFER = 128,     // CF_STATS / CF_BNBWD: only add the per-channel sums to `partials` (a per-BatchNorm accumulator that the
                      // caller zeroed); the elementwise kernel that consumes them finalizes (elementwise.cuh: BnDefer / BnBwdDefer).
                      // Saves the counter round trip and the last-CTA tail (~6 k clocks per launch).
  CF_FUSE = 256,      // (with CF_DEFER) grid-synchronous BatchNorm: every tile's accumulator STAYS in TMEM (the launch has at most
                      // acc_sets tiles per CTA - 148 SMs x 512 columns hold a whole layer's output at batch 128), the CTAs meet at
                      // a grid barrier once the per-channel sums are complete, and a second epilogue pass applies the
                      // BatchNorm the sums belong to:
                      //   CF_STATS: out2 = relu(bn(y) [+ residual | + bn2(residual)]) + its ReLU bit mask (what bn_apply did)
                      //   CF_BNBWD: dy1 = BatchNorm-backward(dz, y1) -> out2 [, dy2 -> out3 with CF_BNBWD2] (what bn_bwd_apply did)
                      // One launch instead of two, and the elementwise pass's re-read of the conv output disappears.
  CF_FUSE_RES = 512,  // CF_STATS | CF_FUSE: the second pass adds `fuse_res` (through tmRes), scaled by fuse_rvec if given
  CF_NO_STORE = 1024, // CF_BNBWD | CF_FUSE: pass 1 does not store dz (nobody else reads it); pass 2 recomputes it from TMEM
};

struct FlatConvParams {
  CUtensorMap tmA;  // 2-D (channels, flat pixels), box (64, a_box_rows)
  CUtensorMap tmB;  // 2-D (channels, taps * n_total), box (64, block_n)
  CUtensorMap tmOut;  // 2-D (n_total, flat pixels) view of `out`, box (64, 32): one epilogue warp's 32 rows x 64 channels
  CUtensorMap tmRes;  // the same view of `residual`, `y1`, `y2` (operand tiles of the epilogue; conv_flat.cu: flat_conv_bind_operands)
  CUtensorMap tmY1;
  CUtensorMap tmY2;
  int operand_maps;   // tmRes / tmY1 / tmY2 are valid for the current residual / y1 / y2 pointers
  int total_rows;
  PadGeom g;
  int mt;                      // 128-row sub-tiles per tile and CTA (they share every weight tile)
  int pair;                    // 1: clusters of two CTAs share every MMA (cta_group::2); a tile then covers 2*mt*128 rows
  int block_n, n_blocks, n_total;
  int chunks, num_taps;
  int tap_shift[9];
  int tap_slab[9];
  int halo;                    // Wp + 1
  int a_box_rows, a_boxes;     // slab = a_boxes boxes of a_box_rows rows
  int tap_group;               // taps per weight stage (one mbarrier round trip per group); 9 with resident weights
  int a_stages, b_stages, b_resident, acc_sets;
  int m_tiles;
  int flags;
  __nv_bfloat16* out;
  const __nv_bfloat16* residual;
  const __nv_bfloat16* mask;
  const uint8_t* mask_bits;    // optional: the same mask as one bit per element ([rows][n_total/8] bytes); used instead of `mask`
  const float* scale;
  const float* bias;
  double* partials;            // [3][n_total] global fp64 accumulators (zero on entry, left zero by the kernel)
  unsigned int* counter;
  // CF_STATS: BatchNorm forward finalize (by the last CTA)
  const float* gamma;
  const float* beta;
  float* running_mean;
  float* running_var;
  long long* nbt;
  float* vec;                  // [4][n_total]: scale, shift, mean, rstd
  double count;
  float momentum, eps;
  int update_running;
  // CF_BNBWD / CF_BNBWD2
  const __nv_bfloat16* y1;
  const float* stat1;          // vec of BN 1 ([4][n_total]; mean at 2n, rstd at 3n)
  float* bred1;                // [2][n_total]: bsum, bdot
  float* dgamma1;
  float* dbeta1;
  const __nv_bfloat16* y2;
  const float* stat2;
  float* bred2;
  float* dgamma2;
  float* dbeta2;
  // CF_FUSE
  unsigned int* grid_bar;      // zeroed device word: the launch's grid barrier
  CUtensorMap tmOut2;          // second-pass output (activation / dy1), box (64, 32)
  CUtensorMap tmOut3;          // dy2 (CF_BNBWD2)
  uint8_t* bits_out;           // CF_STATS: ReLU bit mask of out2 ([rows][n_total / 8] bytes)
  const float* fuse_rvec;      // CF_FUSE_RES: [4][n_total] vectors of the BatchNorm applied to the residual (downsample branch) or null
  const float* gamma1;         // CF_BNBWD: BatchNorm weights of BN 1 / BN 2
  const float* gamma2;
  double inv_count;            // 1 / (batch * H * W)
  double unbias;               // count / (count - 1)
};

constexpr int WF_THREADS = 192;
constexpr int WF_MAX_STAGES = 4;
constexpr int WF_MAX_TAPS = 5;

struct WgradFlatParams {
  CUtensorMap tmDY;  // 2-D (cout, flat pixels), box (64, 128)
  CUtensorMap tmX;   // 2-D (cin, flat pixels), box (64, x_box_rows)
  int total_rows, k_tiles;
  int co_blocks, m_halves, ci_chunks;
  int tap_groups;                      // 3: one filter row per CTA
  int tap_shift[9];
  int x_box_rows, x_boxes;             // slab rows per stage = x_boxes * x_box_rows
  int split_z, num_stages;
  int cout, cin;
  float* scratch;                      // split-K partial tiles [tile][z][128][192] fp32 (wgrad_reduce_kernel folds them)
  // cluster form (wgrad_flat3_kernel): the three filter rows of a (co block, ci chunk, K slice) are one cluster of three CTAs
  // that multicast their loads to each other
  int cluster3;                        // 1: launch wgrad_flat3_kernel
  CUtensorMap tmXU;                    // 2-D (cin, flat pixels), box (64, xu_rows): the union of the three filter rows' slabs
  int xu_rows, wp;                     // 130 + 2 * Wp rounded up to 8; padded row pitch Wp
  int stages3;
};

struct WgradReduceJob {
  const float* scratch;
  long long grad_off;   // float offset of the OIHW gradient inside the gradient arena
  int cout, cin, ci_chunks, split_z;
  int rows;             // output channels served by one CTA
  int zero_src;         // clear the accumulator tile after reading it (it is ADDED to by the next backward's K slices)
  int first_block;      // CTAs [first_block, first_block + cout / rows * ci_chunks) belong to this job
};
struct WgradReduceJobs {  // passed by value as the kernel parameter: no device-side table
  int n, total_blocks;
  WgradReduceJob job[16];
};

constexpr long long WF_SCRATCH_BYTES = 160LL * 128 * 192 * 4;  // >= the (co block, ci chunk, filter row) accumulator tiles of 128 x 192 fp32 of one convolution
constexpr int WF_STAGE_PITCH = 784;                            // shared-memory pitch of a 768-byte accumulator row in the epilogue staging

}  // namespace cilrs
