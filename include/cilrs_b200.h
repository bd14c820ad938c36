/* cilrs_b200 — C-ABI of the B200-native CILRS hot path.
 *
 * The reference (rohithr87/CILRS-Autonomous-Driving-CARLA) is pure Python and has no FFI of its own; its hot
 * path is reached through three Python surfaces, and every entry point below replaces the library call made
 * underneath one of them (citations are into /root/reference):
 *
 *   - AutonomousDriver.preprocess_image      model/autonomous_drive.py:897-902   (cv2.resize + /255 + Normalize)
 *     prepare_dataset.process_session        model/prepare_dataset.py:56         (cv2.resize)
 *   - CILRS.forward / loss.backward()        model/autonomous_drive.py:361-399, notebook/notebook.ipynb:549-552
 *   - optim.Adam(...).step()                 notebook/notebook.ipynb:533-534,555
 *
 * Conventions (all entry points):
 *   - plain C types, raw DEVICE pointers, explicit sizes; `stream` is a cudaStream_t passed as void*.
 *   - return 0 on success; 1 = invalid argument, 2 = unsupported shape, 3 = workspace too small,
 *     4 = driver entry point unavailable, >= 1000 = 1000 + cudaError_t of the failed CUDA call.
 *   - never allocate device memory, never synchronise, never throw. Memory is owned by the caller.
 *   - activations are NHWC bf16 ("channels-last") inside the conv stack; the public tensors of the
 *     reference's interface (image f32 NCHW, controls f32 [B,3], pred_speed f32 [B]) keep their layout.
 */
#ifndef CILRS_B200_H_
#define CILRS_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CILRS_ABI_VERSION 4

int cilrs_abi_version(void);
/* human-readable text for a status code returned by any entry point (static storage) */
const char* cilrs_status_string(int status);
/* kernels launched through this library since it was loaded (host-side counter) */
long long cilrs_launch_count(void);

/* ---------------------------------------------------------------------------------------------------------
 * K0  frame preprocessing.  Replaces cv2.resize(image,(200,88)) [INTER_LINEAR, 11-bit fixed point] and
 *     img/255 -> permute(2,0,1) -> Normalize(mean,std)   (model/autonomous_drive.py:898-901, :481-482, :502-504)
 *
 *  src      : uint8 [batch, src_h, src_w, src_c]  (src_c = 3 RGB/BGR or 4 BGRA/RGBA; alpha ignored)
 *  reverse  : 0 keep channel order, 1 reverse the first three channels (BGR(A) -> RGB; autonomous_drive.py:1551)
 *  dst_u8   : optional uint8 [batch, dst_h, dst_w, 3]         — bit-exact cv2.resize result (prepare_dataset.py:56)
 *  dst_f32  : optional float [batch, 3, dst_h, dst_w]         — the tensor preprocess_image returns
 *  dst_s2d  : optional bf16  [batch, 47, 103, 16]             — conv1-ready space-to-depth layout (only 88x200)
 *  Only the (src 600x800 -> dst 88x200) geometry has to be fast; any size with src >= dst is accepted.
 * --------------------------------------------------------------------------------------------------------- */
int cilrs_preprocess_u8(const uint8_t* src, int batch, int src_h, int src_w, int src_c, int reverse,
                        int dst_h, int dst_w, uint8_t* dst_u8, float* dst_f32, void* dst_s2d, void* stream);

/* image f32 NCHW [batch,3,88,200] (the tensor CILRS.forward receives) -> conv1-ready bf16 [batch,47,103,16] */
int cilrs_image_to_s2d(const float* image_nchw, int batch, void* dst_s2d, void* stream);

/* ---------------------------------------------------------------------------------------------------------
 * K1/K2  convolution primitives (replace the cuDNN calls under nn.Conv2d in torchvision resnet34,
 *        model/autonomous_drive.py:365-369). bf16 NHWC activations, fp32 accumulate on tcgen05 tensor cores.
 * --------------------------------------------------------------------------------------------------------- */
typedef struct cilrs_conv_desc {
  int batch;
  int in_h, in_w, in_c; /* input NHWC; in_c multiple of 64 (the 7x7 stem uses the *_stem entry points) */
  int out_c;            /* multiple of 64 */
  int kh, kw;           /* 3x3 or 1x1 */
  int stride;           /* 1 or 2 */
  int pad;              /* 1 for 3x3, 0 for 1x1 */
} cilrs_conv_desc;

enum {
  CILRS_EPI_STATS = 1,      /* also emit per-CTA per-channel (sum, sum of squares) partials for train-mode BatchNorm */
  CILRS_EPI_SCALE_BIAS = 2, /* y = conv * scale[c] + bias[c]  (folded eval-mode BatchNorm) */
  CILRS_EPI_RESIDUAL = 4,   /* y += residual */
  CILRS_EPI_RELU = 8,       /* y = max(y, 0) */
  /* padded-flat kernels only (cilrs_conv_flat): */
  CILRS_EPI_MASK = 16,      /* y = mask > 0 ? y : 0  (ReLU backward fused into the dgrad epilogue) */
  CILRS_EPI_BNBWD = 32,     /* also reduce sum(y), sum(y * xhat1) per channel = the BatchNorm-backward reductions */
  CILRS_EPI_BNBWD2 = 64,    /* ... and sum(y * xhat2) for a second BatchNorm fed by the same gradient */
  CILRS_EPI_DEFER = 128     /* with STATS / BNBWD: only ADD the per-channel fp64 sums to partials_ws ([3][out_c] doubles, zeroed by the
                               caller: sum, sum of squares | sum dz, sum dz*y1, sum dz*y2) and leave the finalize to the consumer -
                               what the network plan does (the BatchNorm apply kernels finalize in their prologue) */
};

/* bytes of packed bf16 weights for fprop / dgrad, and of the stats scratch for a given desc */
size_t cilrs_conv_packed_weight_bytes(const cilrs_conv_desc* d);
size_t cilrs_conv_stats_bytes(const cilrs_conv_desc* d); /* floats [partials][2][out_c], one partial per persistent CTA */
int cilrs_conv_stats_tiles(const cilrs_conv_desc* d);

/* fp32 OIHW master weight -> bf16 [tap][out_c][in_c] (fprop) and bf16 [tap][in_c][out_c] (dgrad); either may be NULL */
int cilrs_conv_pack_weight(const cilrs_conv_desc* d, const float* w_oihw, void* w_fprop, void* w_dgrad, void* stream);

int cilrs_conv_fprop(const cilrs_conv_desc* d, const void* x, const void* w_fprop, void* y, const float* scale,
                     const float* bias, const void* residual, float* stats, int flags, void* stream);
/* dx = conv_transpose(dy, w) (+ residual). For stride 2 the four output parities are separate launches. */
int cilrs_conv_dgrad(const cilrs_conv_desc* d, const void* dy, const void* w_dgrad, void* dx, const void* residual,
                     void* stream);
/* dw_oihw (fp32) += dy^T * im2col(x). The caller zeroes dw first. */
int cilrs_conv_wgrad(const cilrs_conv_desc* d, const void* dy, const void* x, float* dw_oihw, void* stream);

/* 3x3 stride-1 pad-1 convolutions on the PADDED-FLAT activation layout: a tensor [batch, H+1, W+1, C] bf16 whose last
 * row and last column of every image are zero, seen as a flat list of (H+1)*(W+1)*batch pixels. A 3x3 tap is then a
 * constant shift of the flat pixel index, one TMA-loaded slab serves all nine taps, and fprop / dgrad / wgrad of the 29
 * stride-1 3x3 convolutions of ResNet-34 (torchvision/models/resnet.py BasicBlock.conv1/conv2) run on it.
 * dgrad = 1: x is dy (in_c = conv output channels), w the dgrad pack, y = dx (out_c = conv input channels).
 * CILRS_EPI_STATS here also FINALIZES train-mode BatchNorm inside the kernel (last CTA): vec [4][out_c] = scale, shift,
 * mean, rstd, running statistics updated like torch.nn.BatchNorm2d. CILRS_EPI_BNBWD writes bred [2][out_c] = (sum dz,
 * sum dz*xhat) and accumulates dgamma / dbeta. */
typedef struct cilrs_flat_conv_args {
  int batch, H, W, in_c, out_c;
  int dgrad;
  int flags;
  const void* x;
  const void* w;
  void* y;
  const float* scale;
  const float* bias;
  const void* residual; /* padded-flat [rows][out_c] */
  const void* mask;     /* padded-flat [rows][out_c] */
  const void* mask_bits; /* optional uint8 [rows][out_c / 8]: the same mask, one bit per element (see cilrs_bn_apply) */
  const float* gamma;
  const float* beta;
  float* running_mean;
  float* running_var;
  long long* num_batches_tracked;
  float* vec;
  float momentum, eps;
  int update_running;
  const void* y1;
  const float* vec1;
  float* bred1;
  float* dgamma1;
  float* dbeta1;
  const void* y2;
  const float* vec2;
  float* bred2;
  float* dgamma2;
  float* dbeta2;
  float* partials_ws;       /* cilrs_conv_flat_workspace_floats(out_c) floats, ZEROED; left zero by the kernel */
  unsigned int* counter_ws; /* one zeroed uint32, left zero by the kernel */
} cilrs_flat_conv_args;
long long cilrs_flat_rows(int batch, int H, int W);
/* Train-mode BatchNorm fused into the flat convolutions ("grid-synchronous BatchNorm": the accumulators of a whole layer stay in
 * tensor memory across a grid barrier and a second epilogue pass applies the BatchNorm + ReLU / the BatchNorm backward) on or
 * off for plans built after the call; returns the previous setting. Test / measurement aid - both settings compute the same
 * network (replaces the ATen BatchNorm launches under torchvision/models/resnet.py:89-105 either way). */
int cilrs_set_bn_fusion(int enable);
/* The flat weight-gradient kernel as clusters of three CTAs (one per filter row) that multicast their TMA loads to each other:
 * measurement / test aid, off by default (measured slower on B200); applies to plans built after the call; returns the previous
 * setting (ABI 4). */
int cilrs_set_wgrad_cluster(int enable);
size_t cilrs_conv_flat_workspace_floats(int out_c);
int cilrs_conv_flat(const cilrs_flat_conv_args* a, void* stream);
/* dw_oihw (fp32 [out_c,in_c,3,3]) += dy^T * shifted(x), both padded-flat. Two launches: the split-K slices add their tiles into
 * accumulator tiles in scratch_ws with bulk asynchronous fp32 reductions, then one pass permutes the accumulators into dw and
 * clears them. scratch_ws: cilrs_wgrad_flat_workspace_bytes() bytes, ZERO on entry, left zero. */
size_t cilrs_wgrad_flat_workspace_bytes(void);
int cilrs_wgrad_flat(int batch, int H, int W, int in_c, int out_c, const void* dy, const void* x, float* dw_oihw,
                     float* scratch_ws, void* stream);

/* the 7x7/2 stem on the space-to-depth input [batch,47,103,16] -> [batch,44,100,64] */
size_t cilrs_stem_packed_weight_bytes(void);
int cilrs_stem_pack_weight(const float* w_oihw /* [64,3,7,7] */, void* w_packed, void* stream);
int cilrs_stem_fprop(int batch, const void* x_s2d, const void* w_packed, void* y, const float* scale, const float* bias,
                     float* stats, int flags, void* stream);
int cilrs_stem_wgrad(int batch, const void* dy, const void* x_s2d, float* dw_oihw, void* stream);
int cilrs_stem_stats_tiles(int batch);

/* ---------------------------------------------------------------------------------------------------------
 * M1-M5 / B1  the whole CILRS network behind one handle.  Replaces CILRS.forward (model/autonomous_drive.py:389-399)
 * and the autograd backward the reference gets from loss.backward() (notebook/notebook.ipynb:552).
 *
 * Memory model: the caller owns three arenas and a workspace, all device memory:
 *   params  : fp32, every nn.Parameter of the reference module in named_parameters() order, each tensor starting
 *             at the offset cilrs_model_param_layout reports (64-byte aligned), OIHW / [out,in] as in the state_dict
 *   grads   : fp32, same layout; backward ACCUMULATES into it (zero it like optimizer.zero_grad())
 *   buffers : fp32, for each of the 36 BatchNorms in order: running_mean[C] then running_var[C];
 *             num_batches_tracked: int64[36]
 *   workspace : cilrs_model_workspace_bytes(max_batch) bytes, 1024-byte aligned
 * --------------------------------------------------------------------------------------------------------- */
typedef struct cilrs_model cilrs_model;

/* returns the number of parameter tensors (142); fills offsets/sizes (in floats) for the first `capacity` */
int cilrs_model_param_layout(long long* offsets, long long* sizes, int capacity, long long* total_floats,
                             long long* buffer_floats, int* num_bn);
size_t cilrs_model_workspace_bytes(int max_batch);
/* creation is the only call that synchronises the stream (it uploads a few constants) */
int cilrs_model_create(cilrs_model** out, int max_batch, void* workspace, size_t workspace_bytes, void* stream);
void cilrs_model_destroy(cilrs_model* m);
int cilrs_model_bind(cilrs_model* m, float* params, float* grads, float* buffers, long long* num_batches_tracked);
/* call after parameter values changed. what: bit 0 = repack the bf16 conv operands from the fp32 masters,
 * bit 1 = fold eval-mode BatchNorm (running statistics) into per-channel scale/shift for CILRS_MODE_INFER */
int cilrs_model_refresh(cilrs_model* m, int what, void* stream);
/* bit-0 refresh off the critical path: the stem's operand is packed on `stream`, the trunk's on the model's gradient stream
 * (forked behind the work already queued on `stream`); the next cilrs_model_forward* / _backward / _refresh call on `stream`
 * waits for it - the forward only after its stem. Must be followed by such a call before a stream capture of `stream` ends. */
int cilrs_model_refresh_async(cilrs_model* m, void* stream);
/* bit-0 refresh restricted to the convolutions whose gradients backward part `part` completes (0 = layer4 ... 3 = layer1,
 * 4 = stem): the optimizer step + repack of a finished part can run under the rest of the backward */
int cilrs_model_refresh_part(cilrs_model* m, int part, void* stream);

enum { CILRS_MODE_TRAIN = 0,  /* BN uses batch statistics (module.train()) */
       CILRS_MODE_FROZEN = 1, /* BN uses running statistics, activations kept: eval() with autograd */
       CILRS_MODE_INFER = 2   /* eval(), no_grad: BN folded into the conv epilogues */ };

/* image: either f32 NCHW [batch,3,88,200] (image_nchw) or the bf16 space-to-depth tensor K0 produced (image_s2d).
 * speed f32 [batch], command int64 [batch] in {0..3}; controls f32 [batch,3], pred_speed f32 [batch]. */
int cilrs_model_forward(cilrs_model* m, int batch, int mode, const float* image_nchw, const void* image_s2d,
                        const float* speed, const long long* command, float* controls, float* pred_speed,
                        int update_running_stats, int keep_for_backward, float dropout_p, unsigned long long seed,
                        void* stream);
/* The same forward (keep_for_backward = 1, mode TRAIN or FROZEN) with the training loss of cilrs_loss fused into the heads kernel:
 * out6, dcontrols [batch,3], dspeed [batch] as cilrs_loss writes them; the speed head's target is the speed input
 * (criterion(pred_ctrl, tgts, pred_spd, speeds), notebook/notebook.ipynb:550). */
int cilrs_model_forward_loss(cilrs_model* m, int batch, int mode, const float* image_nchw, const void* image_s2d,
                             const float* speed, const long long* command, float* controls, float* pred_speed,
                             int update_running_stats, float dropout_p, unsigned long long seed, const float* targets,
                             int loss_mode, float w_steer, float w_throttle, float w_brake, float w_speed, float grad_scale,
                             float* out6, float* dcontrols, float* dspeed, void* stream);
/* gradients of a scalar w.r.t. controls / pred_speed in, parameter gradients accumulated into `grads`.
 * part = -1 runs the whole backward; parts 0..4 (heads+layer4, layer3, layer2, layer1, stem; in that order) let the
 * host start the allreduce of the gradient range a part completed while the next part runs (data parallelism).
 * Parts 5 and 6 are the two halves of part 4 (5: the stem's max-pool / ReLU / BatchNorm backward, 6: conv1's weight
 * gradient), for a caller that wants to start the optimizer between them (ABI 4). */
int cilrs_model_backward(cilrs_model* m, int batch, int mode, int part, const float* dcontrols, const float* dspeed,
                         const float* speed, const long long* command, float dropout_p, void* stream);
/* first parameter-tensor index (into cilrs_model_param_layout) whose gradient backward part `part` completes;
 * part p completes tensors [first(p), first(p-1)) with first(-1) = number of tensors */
int cilrs_model_backward_part_first_tensor(int part);
/* Asynchronous form of a backward part (part = 0..4): the caller's stream is NOT made to wait for the part's weight
 * gradients (they run on the model's internal gradient stream). Instead everything the part wrote is complete in the ORDER OF
 * cilrs_model_gradient_stream(): enqueue the part's allreduce there (or behind an event recorded there), call the next
 * part right away, and join once after the last part with cilrs_model_backward_join(m, stream) before the optimizer.
 * gradient_stream returns NULL when the model runs everything on the caller's stream (then the three calls degrade to the
 * synchronous behaviour). */
int cilrs_model_backward_part_async(cilrs_model* m, int batch, int mode, int part, const float* dcontrols, const float* dspeed,
                                    const float* speed, const long long* command, float dropout_p, void* stream);
void* cilrs_model_gradient_stream(cilrs_model* m);
int cilrs_model_backward_join(cilrs_model* m, void* stream);
/* measurement aid (bench.py roofline): CUDA events around every launch of the plan, summed per kernel class
 * {conv fprop, conv dgrad, conv wgrad, BN forward + pooling, BN backward, heads, other}. collect() synchronises. */
int cilrs_model_profile(cilrs_model* m, int enable);
int cilrs_model_profile_collect(cilrs_model* m, float* out_ms7, int* out_launches7);
void* cilrs_model_input_s2d(cilrs_model* m); /* where K0 may write the conv1-ready frames directly */
int cilrs_model_invalidate_plans(cilrs_model* m); /* rebuild tensor maps / tile shapes / fusion decisions at the next forward */
/* test hook: bf16 activation of the last forward. which: 0 = max-pool output, 1..16 = BasicBlock outputs,
 * 17 = raw stem conv output; dims receives {H, W, C, Hp, Wp}: the tensor is [batch, Hp, Wp, C] with the real pixels in
 * [:, :H, :W] (padded-flat layout; the stem output is dense, Hp = H, Wp = W) */
void* cilrs_model_debug_activation(cilrs_model* m, int which, int* dims);
int* cilrs_model_error_flag(cilrs_model* m); /* device int, set to 1 when a command was outside [0,4) */
/* test hook: backward of BasicBlocks hi..max(lo,0) only (hi < 0: none; blocks are numbered 0..15 in forward order), plus the stem
 * (max-pool / BN / conv1 backward) when lo < 0, from a given gradient. g_out: bf16 padded-flat gradient w.r.t. the OUTPUT of block
 * hi before its final ReLU mask (stem only: w.r.t. the max-pool output [batch,23,51,64]); padding pixels zero. Must follow a
 * forward(keep_for_backward) of the same batch / mode. Parameter gradients accumulate into the bound arena; the gradient w.r.t.
 * the input of block max(lo,0) is left in cilrs_model_debug_gradient() (padded-flat, geometry of that input). */
int cilrs_model_debug_backward(cilrs_model* m, int batch, int mode, int hi, int lo, const void* g_out, void* stream);
void* cilrs_model_debug_gradient(cilrs_model* m);
/* test hook: fp32 [batch, width] head activations the last forward(keep_for_backward) kept (post-ReLU, post-Dropout).
 * which: 0 speed_encoder.0 (128), 1 speed_encoder.3 (128), 2 branch.0 (256), 3 branch.3 (256), 4 speed_predictor.0 (256),
 * 5 speed_predictor.3 (256) */
float* cilrs_model_debug_heads_saved(cilrs_model* m, int which, int* width);
/* Dropout(p) of the heads under a captured CUDA graph (notebook/notebook.ipynb:480 trains with dropout=0.5): the mask seed of a
 * forward becomes seed + c * (*counter_dev + 1); counter_dev is a device int64 that changes between replays. NULL = off. */
int cilrs_model_set_dropout_counter(cilrs_model* m, const long long* counter_dev);

/* ---------------------------------------------------------------------------------------------------------
 * fp32 compute mode of the same network (csrc/fp32_path.cu): fp32 storage and fp32 FMA end to end, the precision the
 * reference runs in (configs/train_config.json:54 "mixed_precision": false; model/autonomous_drive.py:495). Same arenas
 * (cilrs_model_param_layout), same modes, same call pattern; image_nchw only (no bf16 input). Parity 1e-4 vs the fp64 oracle.
 * --------------------------------------------------------------------------------------------------------- */
typedef struct cilrs_model32 cilrs_model32;
size_t cilrs_model32_workspace_bytes(int max_batch);
int cilrs_model32_create(cilrs_model32** out, int max_batch, void* workspace, size_t workspace_bytes, void* stream);
void cilrs_model32_destroy(cilrs_model32* m);
int cilrs_model32_bind(cilrs_model32* m, float* params, float* grads, float* buffers, long long* num_batches_tracked);
int cilrs_model32_forward(cilrs_model32* m, int batch, int mode, const float* image_nchw, const float* speed,
                          const long long* command, float* controls, float* pred_speed, int update_running_stats,
                          int keep_for_backward, float dropout_p, unsigned long long seed, void* stream);
int cilrs_model32_backward(cilrs_model32* m, int batch, int mode, const float* dcontrols, const float* dspeed,
                           const float* speed, const long long* command, float dropout_p, void* stream);
int* cilrs_model32_error_flag(cilrs_model32* m);

/* heads only (speed encoder + selected branch + speed predictor; model/autonomous_drive.py:371-398) on given
 * features f32 [batch,512]; the backward also accumulates the head parameter gradients and returns d(features) */
int cilrs_model_heads_forward(cilrs_model* m, int batch, const float* feat, const float* speed, const long long* command,
                              float* controls, float* pred_speed, int keep_for_backward, float dropout_p,
                              unsigned long long seed, void* stream);
int cilrs_model_heads_backward(cilrs_model* m, int batch, const float* dcontrols, const float* dspeed, const float* speed,
                               const long long* command, float dropout_p, float* dfeat_out, void* stream);

/* ---------------------------------------------------------------------------------------------------------
 * BatchNorm / ReLU / pooling kernels (replace cuDNN/ATen BatchNorm2d, ReLU, MaxPool2d under torchvision resnet34,
 * torchvision/models/resnet.py BasicBlock.forward; train-mode statistics come from the conv epilogue partials).
 * vec = [4][C] fp32 (scale, shift, mean, rstd) written by cilrs_bn_finalize.
 * --------------------------------------------------------------------------------------------------------- */
int cilrs_bn_finalize(const float* partials, int tiles, int C, double count, const float* gamma, const float* beta,
                      float* running_mean, float* running_var, long long* num_batches_tracked, float momentum, float eps,
                      int training, int update_running, float* vec, void* stream);
/* pad_h, pad_w > 0 (both): every activation argument is in the padded-flat layout [batch, pad_h+1, pad_w+1, C] (see
 * cilrs_conv_flat) and `elems` counts the padding pixels too; they are written as zeros and never read. 0, 0 = dense NHWC. */
/* relu_bits (optional): uint8 [elems / 8], bit k of byte i = (out[8 i + k] > 0): the ReLU mask as a bit tensor, which the
 * flat dgrad epilogue (cilrs_flat_conv_args.mask_bits) reads instead of the bf16 activation */
int cilrs_bn_apply(const void* x, const float* vec, const void* residual, const void* x2, const float* vec2, void* out,
                   long long elems, int C, int relu, int pad_h, int pad_w, uint8_t* relu_bits, void* stream);
/* padded_out != 0: out is padded-flat [batch, (H+1)/2+1, (W+1)/2+1, C] (padding pixels untouched); argmax stays dense */
int cilrs_bn_relu_maxpool(const void* y, const float* vec, void* out, uint8_t* argmax, int batch, int H, int W, int C,
                          int padded_out, void* stream);
/* The same pool that also keeps `ysel`, the RAW y at the arg-max of every window (laid out like `out`; needs argmax): with it
 * the backward of torchvision's bn1 -> relu -> maxpool (model/autonomous_drive.py:365-370) takes its BatchNorm sums over the
 * pool outputs instead of the 4x larger conv1 output (ABI 4). */
int cilrs_bn_relu_maxpool_sel(const void* y, const float* vec, void* out, uint8_t* argmax, void* ysel, int batch, int H, int W,
                              int C, int padded_out, void* stream);
/* Backward of the stem's bn1 -> relu -> maxpool as the training plan runs it: sums over (g, act = pooled activations, ysel)
 * - which also overwrite g with g * [act > 0], the ReLU mask where it matters - then one pass that routes it through the
 * arg-max codes and applies the BatchNorm backward.
 * g / act / ysel: [batch, H/2, W/2, C] (padded != 0: padded-flat [batch, H/2+1, W/2+1, C]); y, dy: dense [batch, H, W, C];
 * H, W even. workspace / counter as for cilrs_bn_backward. dgamma / dbeta (optional) are accumulated (+=). (ABI 4) */
int cilrs_stem_bn_backward(void* g, const void* act, const void* ysel, const uint8_t* argmax, const void* y,
                           const float* vec, const float* gamma, int batch, int H, int W, int C, int padded, int frozen,
                           void* dy, float* dgamma, float* dbeta, float* workspace, unsigned int* counter, void* stream);
/* stem variant (argmax != NULL): pad_h > 0 means the pooled gradient g is padded-flat.
 * workspace: cilrs_bn_backward_workspace_floats(C) floats, 8-byte aligned, the first 4*C ZERO on entry (fp64 accumulators,
 * left zero); counter: zeroed uint32 */
int cilrs_bn_backward(const void* g, const void* act, const void* y, const float* vec, const float* gamma, long long elems,
                      int C, double count, int frozen, void* dy, void* dz, float* dgamma, float* dbeta, float* workspace,
                      unsigned int* counter, const uint8_t* argmax, int H, int W, int pad_h, int pad_w, void* stream);
size_t cilrs_bn_backward_workspace_floats(int C);

/* ---------------------------------------------------------------------------------------------------------
 * L1/L2  losses + their gradients (CILRSLoss.forward notebook/notebook.ipynb:514-527; MSE recipe
 *        configs/train_config.json:30-32).  mode 0: MSE(controls)+w_speed*MSE(speed); mode 1: weighted L1 + w_speed*MSE.
 *        out6 = total, control, steer, throttle, brake, speed. dcontrols/dspeed optional.
 * --------------------------------------------------------------------------------------------------------- */
int cilrs_loss(const float* controls, const float* pred_speed, const float* targets, const float* speed_target,
               int batch, int mode, float w_steer, float w_throttle, float w_brake, float w_speed, float grad_scale,
               float* out6, float* dcontrols, float* dspeed, void* stream);

/* validate() on the device (notebook/notebook.ipynb:563-585): ACCUMULATES into acc16 (device double[16], zeroed by the caller
 * before a validation pass): [0..5] the six loss scalars of this batch (same order as cilrs_loss), [6..9] sum of
 * |pred_steer - target_steer| over the samples of command 0..3, [10..13] their counts, [14] += 1 (batches). One read of acc16 at
 * the end of the pass replaces the reference's 6 .item() + up to 4 masked .cpu() copies per batch. */
int cilrs_validate_accumulate(const float* controls, const float* pred_speed, const float* targets, const float* speed_target,
                              const long long* command, int batch, int mode, float w_steer, float w_throttle, float w_brake,
                              float w_speed, double* acc16, void* stream);

/* ---------------------------------------------------------------------------------------------------------
 * O1/O2  optimiser.  torch.optim.Adam(lr, weight_decay) single step over a flat arena (notebook/notebook.ipynb:533-534,555)
 *        and the squared gradient norm + clip coefficient of clip_grad_norm_ (notebook/notebook.ipynb:553-554).
 * --------------------------------------------------------------------------------------------------------- */
/* step: 1-based step number used for the bias corrections; if step_dev (device int64) is given it is incremented on
 * the stream first and used instead, so a captured CUDA graph of the step stays valid for every step number. */
int cilrs_adam_step(float* p, const float* g, float* m, float* v, long long n, float lr, float beta1, float beta2,
                    float eps, float weight_decay, long long step, long long* step_dev, float grad_scale,
                    const float* grad_scale_dev, void* stream);
int cilrs_grad_sumsq(const float* g, long long n, double* partial_ws, unsigned int* counter_ws, float max_norm,
                     float* out2, void* stream);
/* CUDA-graph form of the step: hyper_dev = device float[8] {lr, beta1, beta2, eps, weight_decay, grad_scale, -, -} read by the
 * kernel at run time (StepLR, notebook/notebook.ipynb:535-536,604, reaches a captured graph through a 32-byte copy);
 * step_dev (device int64, required) is incremented on the stream first. g_bf16 (optional): bf16 gradients used instead of g
 * (the all-reduced exchange buffer of data-parallel training). zero_grad is a bit mask: bit 0 also clears the fp32 arena g
 * (optimizer.zero_grad(), notebook/notebook.ipynb:551, of the next step); bit 1 = do NOT advance step_dev (a further range of
 * the arena within the same optimizer step: data-parallel training updates the already-exchanged ranges while the last
 * allreduce is still in flight). */
int cilrs_adam_step_ex(float* p, float* g, const void* g_bf16, float* m, float* v, long long n, const float* hyper_dev,
                       long long* step_dev, const float* grad_scale_dev, int zero_grad, void* stream);
int cilrs_grad_sumsq_bf16(const void* g_bf16, long long n, double* partial_ws, unsigned int* counter_ws, float max_norm,
                          float* out2, void* stream);
/* fp32 gradient range -> bf16 exchange buffer (round to nearest even); zero_source != 0 also clears the fp32 range */
int cilrs_grad_to_bf16(float* g, void* out_bf16, long long n, int zero_source, void* stream);


/* ---------------------------------------------------------------------------------------------------------------------------
 * N3 - input pipeline: baseline JPEG decode on the GPU.
 * Replaces cv2.imread + cv2.cvtColor(BGR2RGB) of CILRSDataset.__getitem__ (notebook/notebook.ipynb:404-405) for the collector's
 * frames (model/collect_data.py:685-716: 200 x 88, quality 95, YCbCr 4:2:0, no restart markers); the output is bit-identical
 * to OpenCV's (libjpeg-turbo defaults: islow inverse DCT, fancy chroma upsampling). Supported: baseline sequential Huffman,
 * 8 bit, grey / YCbCr 4:4:4 / 4:2:0, Huffman and quantisation table ids 0-1 / 0-3, no restart intervals.
 * ------------------------------------------------------------------------------------------------------------------------- */
enum { CILRS_JPEG_OK = 0, CILRS_JPEG_NOT_JPEG = 1, CILRS_JPEG_UNSUPPORTED = 2, CILRS_JPEG_CORRUPT = 3, CILRS_JPEG_SIZE_MISMATCH = 4 };
size_t cilrs_jpeg_desc_bytes(void);        /* bytes of one image descriptor */
size_t cilrs_jpeg_table_set_bytes(void);   /* bytes of one derived Huffman table set */
size_t cilrs_jpeg_plane_bytes(int height, int width);   /* component-plane scratch per image */
/* HOST: parse n streams that lie at bytes[offsets[i] .. offsets[i + 1]) (offsets 4-byte aligned; a shorter true length is
 * fine - trailing bytes after EOI are ignored). Writes n descriptors and up to max_sets distinct table sets (host memory; copy
 * both to the device). A stream that cannot be decoded gets a non-zero status in its descriptor, not an error return. */
int cilrs_jpeg_prepare(const unsigned char* bytes, const long long* offsets, int n, void* descs_out, void* sets_out, int max_sets,
                       int* n_sets_out);
/* HOST: read n files into dst with `threads` threads. offsets must hold 2n + 1 entries: [0..n] = 16-byte aligned starts (and the
 * total), [n + 1 .. 2n] = the exact end of each stream. */
int cilrs_jpeg_read_files(const char* const* paths, int n, unsigned char* dst, long long capacity, long long* offsets, int threads);
/* DEVICE: decode the batch into out_rgb uint8 [n, height, width, 3] (reverse != 0: BGR as cv2.imread returns it). status_dev[i]
 * receives CILRS_JPEG_* per image (a failed image is written as zeros). Two launches, no synchronisation. */
int cilrs_jpeg_decode(const unsigned char* bytes_dev, const void* descs_dev, const void* sets_dev, int n, int height, int width,
                      unsigned char* planes_dev, size_t plane_bytes_per_image, unsigned char* out_rgb, int reverse,
                      unsigned int* status_dev, void* stream);

/* WeightedRandomSampler(weights, num_samples, replacement=True) (notebook/notebook.ipynb:411-413) on the device: out[i] = the
 * first row whose cumulative weight exceeds u_i * cdf[n - 1]; u_i = Philox4x32-10(seed, offset + i), 53 bits. cdf_dev: inclusive
 * prefix sums of the (non-negative) weights, fp64, device. Same distribution as torch.multinomial, not the same random stream. */
int cilrs_weighted_sample(const double* cdf_dev, long long n, long long num_samples, unsigned long long seed, unsigned long long offset,
                          long long* out_dev, void* stream);

/* N4 - the training augmentations of notebook/notebook.ipynb:387-394 on uint8 RGB frames [n, height, width, 3] (in == out is
 * allowed): RandomBrightnessContrast -> HueSaturationValue -> GaussianBlur -> GaussNoise -> CoarseDropout, one CTA per frame,
 * per-frame parameters (cilrs_augment_param_bytes() each; layout documented in cilrs_b200/augment.py) drawn by the caller.
 * Point-wise colour transforms and the blur are bit-identical to the cv2 calls albumentations makes with the same parameters;
 * the noise is N(0, noise_std^2) per value from Philox(seed, offset + ...). Frames of up to 200 KB / 3 bytes. */
size_t cilrs_augment_param_bytes(void);
int cilrs_augment_u8(const unsigned char* in, unsigned char* out, const void* params_dev, int n, int height, int width,
                     unsigned long long seed, unsigned long long offset, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* CILRS_B200_H_ */
