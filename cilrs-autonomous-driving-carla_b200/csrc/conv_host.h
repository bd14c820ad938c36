// Host-side builders for the tcgen05 conv kernels: tensor maps, tile shapes, tap tables, launches.
#pragma once
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include "../../include/cilrs_b200.h"

namespace cilrs {

struct ConvGemmParams;
struct WgradParams;
struct FlatConvParams;
struct WgradFlatParams;
struct PadGeom;

struct BoxShape {
  int BW, BH, BN;
  int tiles_w, tiles_h, tiles_n;
  int m_tiles() const { return tiles_w * tiles_h * tiles_n; }
};
// tile of <= 128 output pixels as a (BW x BH x BN) box of the (OW, OH, batch) output volume
BoxShape choose_box(int ow, int oh, int batch);

int encode_nhwc_map(CUtensorMap* m, const void* base, int C, int W, int H, int N, long long stride_w_bytes,
                    long long stride_h_bytes, long long stride_n_bytes, int box_c, int bw, int bh, int bn, int sw, int sh);
int encode_2d_map(CUtensorMap* m, const void* base, int inner, int rows, int box_inner, int box_rows);

// ---- plan builders (fill kernel parameter blocks; no launches) ----
// gin / gout: padded-flat geometry (conv_params.h) of the input / output tensor, nullptr = dense NHWC
int build_fprop(ConvGemmParams* p, const cilrs_conv_desc* d, const void* x, const void* w, void* y, const float* scale,
                const float* bias, const void* residual, float* stats, int flags, const PadGeom* gin, const PadGeom* gout);
int build_stem_fprop(ConvGemmParams* p, int batch, const void* x_s2d, const void* w, void* y, const float* scale,
                     const float* bias, float* stats, int flags);
// stride-1 dgrad: one plan. stride-2: parity (ph, pw) plan; `dy2/w2` optionally fuse the 1x1/2 downsample dgrad
// into parity (0,0) as an extra tap.
int build_dgrad(ConvGemmParams* p, const cilrs_conv_desc* d, int ph, int pw, const void* dy, const void* w_dgrad,
                void* dx, const void* residual, const void* dy2, const void* w2_dgrad, const PadGeom* gin, const PadGeom* gout);
int build_wgrad(WgradParams* p, const cilrs_conv_desc* d, const void* dy, const void* x, float* dw, const PadGeom* gin,
                const PadGeom* gout);
int build_stem_wgrad(WgradParams* p, int batch, const void* dy, const void* x_s2d, float* dw);

// one-launch repack of every 3x3 / 1x1 conv weight of the network (fp32 OIHW masters -> both bf16 operand layouts)
struct PackJob {
  long long w_off;  // float offset of the OIHW master inside the parameter arena
  __nv_bfloat16* wf;
  __nv_bfloat16* wd;
  int cout, cin, kk, first_block;
};
int launch_pack_all(const float* params, const PackJob* jobs_dev, int njobs, int total_blocks, cudaStream_t s);
int launch_pack_range(const float* params, const PackJob* jobs_dev, int njobs, int block_lo, int block_hi, cudaStream_t s);

// ---- padded-flat 3x3 stride-1 kernels (conv_flat.cu) ----
int flat_total_rows(int batch, const PadGeom& g);
// fprop: x [rows][k_channels] -> out [rows][n_total] with w = bf16 [9][n_total][k_channels];
// dgrad: the same call with dy as x, the dgrad weight pack and dgrad = 1. Epilogue pointers are filled in by the caller.
int build_flat_conv(FlatConvParams* p, int batch, const PadGeom& g, int k_channels, int n_total, int dgrad, const void* x,
                    const void* w, void* out, int flags);
int flat_conv_bind_operands(FlatConvParams* p);  // after residual / y1 / y2 are set
int flat_conv_grid(const FlatConvParams* p);
int flat_conv_fuse_ok(const FlatConvParams* p);                          // CF_FUSE possible for this plan (tiles per CTA <= accumulator sets)
int flat_conv_bind_fuse(FlatConvParams* p, void* out2, void* out3);      // tensor maps of the second-pass outputs
int launch_flat_conv(const FlatConvParams* p, cudaStream_t s);
struct WgradReduceJobs;
// scratch: WF_SCRATCH_BYTES of split-K partial tiles, folded into the OIHW gradient by the reduce launch
int build_wgrad_flat(WgradFlatParams* p, int batch, const PadGeom& g, int cin, int cout, const void* dy, const void* x, float* scratch);
int launch_wgrad_flat(const WgradFlatParams* p, cudaStream_t s);
int add_wgrad_reduce_job(WgradReduceJobs* jobs, const WgradFlatParams* p, long long grad_off);
int launch_wgrad_reduce(const WgradReduceJobs* jobs, float* grads, cudaStream_t s);

int launch_conv_gemm(const ConvGemmParams* p, cudaStream_t s);
int launch_conv_gemm_multi(const ConvGemmParams* plans, int count, cudaStream_t s);  // <= 4 independent problems, one launch
int conv_gemm_grid(const ConvGemmParams* p);  // CTAs launched = number of stats partials
int launch_wgrad(const WgradParams* p, cudaStream_t s);
int conv_out_dim(int in, int k, int stride, int pad);

}  // namespace cilrs
