// Host-side glue of the heads kernels (heads.cuh), shared by the bf16 plan (model.cu) and the fp32 plan (fp32_path.cu):
// both keep the features / saved activations / deltas in their own workspace and describe them with a HeadsCtx.
#pragma once
#include "heads.cuh"
#include <string.h>

namespace cilrs {

constexpr int HD_NUM_SLOTS = 34;  // 4 speed-encoder + 4 x 6 branch + 6 speed-predictor parameter tensors

struct HeadsCtx {
  const float* params;            // parameter arena
  float* grads;                   // gradient arena (backward only)
  long long off[HD_NUM_SLOTS];    // float offsets of the head tensors inside the arenas, state_dict order
  float* feat;                    // [B,512] trunk features
  float* dfeat;                   // [B,512] gradient through the command branch
  float* dfeat2;                  // [B,512] gradient through the speed predictor
  float* head_comb;               // [B,640] = [feat | sfeat], staged for the first branch layer's weight gradient
  HeadsSaved hs;
  int* err_flag;
  const long long* drop_counter;
  unsigned int* loss_counter;     // zeroed device word for the fused loss
};

// optional loss fused into the forward: filled by the caller (controls / pred_speed / batch are taken from the forward)
struct HeadsLossArgs {
  const float* targets;
  const float* speed_target;
  int mode;
  float w_steer, w_throttle, w_brake, w_speed, grad_scale;
  float* out6;
  float* dcontrols;
  float* dspeed;
};

inline HeadsWeights heads_weights(const HeadsCtx& c, const float* base) {
  HeadsWeights w;
  int s = 0;
  auto P = [&](int i) { return base + c.off[i]; };
  w.se0_w = P(s); w.se0_b = P(s + 1); w.se3_w = P(s + 2); w.se3_b = P(s + 3);
  s += 4;
  for (int k = 0; k < 4; ++k) {
    w.br0_w[k] = P(s); w.br0_b[k] = P(s + 1); w.br3_w[k] = P(s + 2); w.br3_b[k] = P(s + 3); w.br6_w[k] = P(s + 4); w.br6_b[k] = P(s + 5);
    s += 6;
  }
  w.sp0_w = P(s); w.sp0_b = P(s + 1); w.sp3_w = P(s + 2); w.sp3_b = P(s + 3); w.sp5_w = P(s + 4); w.sp5_b = P(s + 5);
  return w;
}

inline int heads_forward_run(const HeadsCtx& c, int B, const float* speed, const long long* command, float* controls, float* pred_speed,
                             int keep_for_backward, float dropout_p, unsigned long long seed, const HeadsLossArgs* loss, cudaStream_t s) {
  HeadsFwdParams hp;
  memset(&hp, 0, sizeof(hp));
  hp.w = heads_weights(c, c.params);
  if (keep_for_backward) hp.sv = c.hs;
  hp.feat = c.feat; hp.speed = speed; hp.command = command; hp.controls = controls; hp.pred_speed = pred_speed;
  hp.batch = B; hp.dropout_p = dropout_p; hp.seed = seed; hp.seed_counter = c.drop_counter; hp.error_flag = c.err_flag;
  if (loss) {
    hp.loss.enabled = 1;
    hp.loss.counter = c.loss_counter;
    LossParams& lp = hp.loss.lp;
    lp.targets = loss->targets; lp.speed_target = loss->speed_target; lp.mode = loss->mode;
    lp.w_steer = loss->w_steer; lp.w_throttle = loss->w_throttle; lp.w_brake = loss->w_brake; lp.w_speed = loss->w_speed;
    lp.grad_scale = loss->grad_scale; lp.out = loss->out6; lp.dcontrols = loss->dcontrols; lp.dspeed = loss->dspeed;
  }
  ++g_cilrs_launches;
  return cuda_status(heads_launch_cluster(heads_fwd_kernel, B, s, hp));
}

// deltas + d(features) on stream s; the weight / bias gradients on stream ws (the side stream of the backward when it is in
// use: nothing on the trunk's chain depends on them). ev: event used to order ws behind s (may be null when ws == s).
inline int heads_backward_run(const HeadsCtx& c, int B, const float* dcontrols, const float* dspeed, const float* speed,
                              const long long* command, float dropout_p, cudaStream_t s, cudaStream_t ws, cudaEvent_t ev) {
  HeadsBwdParams bp;
  bp.w = heads_weights(c, c.params); bp.sv = c.hs; bp.dcontrols = dcontrols; bp.dspeed = dspeed; bp.command = command;
  bp.dfeat = c.dfeat; bp.dfeat2 = c.dfeat2; bp.batch = B; bp.dropout_p = dropout_p;
  ++g_cilrs_launches;
  int st = cuda_status(heads_launch_cluster(heads_bwd_kernel, B, s, bp));
  if (st) return st;
  if (ws != s) {
    st = cuda_status(cudaEventRecord(ev, s));
    if (st) return st;
    st = cuda_status(cudaStreamWaitEvent(ws, ev, 0));
    if (st) return st;
  }
  HeadsWgradParams wp;
  memset(&wp, 0, sizeof(wp));
  wp.batch = B; wp.command = command; wp.speed = speed;
  int sidx = 0, tiles = 0, nj = 0;
  auto G = [&](int i) { return c.grads + c.off[i]; };
  auto add = [&](const float* delta, int ldd, const float* x, int ldx, int out, int in, int slot_w, int branch) {
    HeadsWgradJob& j = wp.job[nj++];
    j.delta = delta; j.x = x; j.dw = G(slot_w); j.db = G(slot_w + 1); j.out = out; j.in = in; j.ld_delta = ldd; j.ld_x = ldx;
    j.branch = branch; j.tile_begin = tiles; j.tiles_i = (in + 63) / 64;
    tiles += j.tiles_i * ((out + 15) / 16);
  };
  add(c.hs.d_se0, 128, speed, 1, 128, 1, sidx, -1);
  add(c.hs.d_se3, 128, c.hs.s1, 128, 128, 128, sidx + 2, -1);
  sidx += 4;
  for (int k = 0; k < 4; ++k) {
    add(c.hs.d_br0, 256, c.head_comb, 640, 256, 640, sidx, k);   // x of the first branch layer is [feat | sfeat]
    add(c.hs.d_br3, 256, c.hs.b1, 256, 256, 256, sidx + 2, k);
    add(c.hs.d_br6, 4, c.hs.b2, 256, 3, 256, sidx + 4, k);
    sidx += 6;
  }
  add(c.hs.d_sp0, 256, c.feat, 512, 256, 512, sidx, -1);
  add(c.hs.d_sp3, 256, c.hs.p1, 256, 256, 256, sidx + 2, -1);
  add(c.hs.d_sp5, 1, c.hs.p2, 256, 1, 256, sidx + 4, -1);
  wp.num_jobs = nj;
  st = cuda_status(cudaMemcpy2DAsync(c.head_comb, 640 * 4, c.feat, 512 * 4, 512 * 4, B, cudaMemcpyDeviceToDevice, ws));
  if (st) return st;
  st = cuda_status(cudaMemcpy2DAsync(c.head_comb + 512, 640 * 4, c.hs.sfeat, 128 * 4, 128 * 4, B, cudaMemcpyDeviceToDevice, ws));
  if (st) return st;
  heads_wgrad_kernel<<<tiles, 256, 0, ws>>>(wp); ++g_cilrs_launches;
  return cuda_status(cudaGetLastError());
}

}  // namespace cilrs
