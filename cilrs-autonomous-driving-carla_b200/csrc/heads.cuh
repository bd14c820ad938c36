// K3: the CILRS heads in fp32 on CUDA cores (they are < 0.1 % of the model FLOPs and latency-bound):
//   speed encoder 1->128->128, command-selected control branch 640->256->256->3, speed predictor 512->256->256->1,
//   the two training losses and their gradients, and the head backward pass.
// Reference: CILRS.forward model/autonomous_drive.py:389-399 (all 4 branches evaluated then gather(0, command));
// only the selected branch is computed here — the others receive exactly zero gradient.
#pragma once
#include "common.cuh"

namespace cilrs {

constexpr int HD_THREADS = 256;

struct HeadsWeights {
  const float *se0_w, *se0_b, *se3_w, *se3_b;
  const float *br0_w[4], *br0_b[4], *br3_w[4], *br3_b[4], *br6_w[4], *br6_b[4];
  const float *sp0_w, *sp0_b, *sp3_w, *sp3_b, *sp5_w, *sp5_b;
};
struct HeadsGrads {
  float *se0_w, *se0_b, *se3_w, *se3_b;
  float *br0_w[4], *br0_b[4], *br3_w[4], *br3_b[4], *br6_w[4], *br6_b[4];
  float *sp0_w, *sp0_b, *sp3_w, *sp3_b, *sp5_w, *sp5_b;
};
// activations kept for the backward pass (post-ReLU, post-dropout), all fp32 row-major [B, width]
struct HeadsSaved {
  float *s1, *sfeat;   // [B,128] each
  float *b1, *b2;      // [B,256]
  float *p1, *p2;      // [B,256]
  // deltas written by the backward pass, consumed by the weight-gradient kernel
  float *d_se0, *d_se3;          // [B,128]
  float *d_br0, *d_br3, *d_br6;  // [B,256],[B,256],[B,4] (3 used)
  float *d_sp0, *d_sp3, *d_sp5;  // [B,256],[B,256],[B,1]
};

CILRS_DEVINL float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// counter-based dropout mask (training with p > 0; parity tests use p = 0): keep iff hash >= p * 2^32
CILRS_DEVINL bool drop_keep(unsigned long long seed, int sample, int site, int j, float p) {
  unsigned long long z = seed + 0x9E3779B97F4A7C15ull * ((unsigned long long)sample * 4096ull + (unsigned long long)site * 1024ull + (unsigned long long)j + 1ull);
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  z = z ^ (z >> 31);
  return (float)(z >> 40) * (1.0f / 16777216.0f) >= p;
}

// y[o] = act( W[o,:] . x + b[o] ),  W row-major [out, in], in % 128 == 0 or in == 1; x and y in shared memory
CILRS_DEVINL void gemv_rows(const float* __restrict__ W, const float* __restrict__ b, const float* x, int in, int out, float* y,
                            bool relu) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = HD_THREADS / 32;
  const int n4 = in >> 2;
  // eight output rows per warp iteration: 8 independent load/FMA chains per lane hide the L2 latency of the weight rows
  constexpr int R = 8;
  for (int o0 = warp * R; o0 < out; o0 += nwarps * R) {
    float acc[R];
#pragma unroll
    for (int q = 0; q < R; ++q) acc[q] = 0.f;
    for (int i = lane; i < n4; i += 32) {
      const float4 x4 = *reinterpret_cast<const float4*>(x + 4 * i);
      float4 w4[R];
#pragma unroll
      for (int q = 0; q < R; ++q)
        w4[q] = (o0 + q < out) ? __ldg(reinterpret_cast<const float4*>(W + (size_t)(o0 + q) * in) + i) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
      for (int q = 0; q < R; ++q) {
        acc[q] = fmaf(w4[q].x, x4.x, acc[q]); acc[q] = fmaf(w4[q].y, x4.y, acc[q]);
        acc[q] = fmaf(w4[q].z, x4.z, acc[q]); acc[q] = fmaf(w4[q].w, x4.w, acc[q]);
      }
    }
#pragma unroll
    for (int q = 0; q < R; ++q) {
      const float r = warp_sum(acc[q]);
      if (lane == 0 && o0 + q < out) {
        float v = r + b[o0 + q];
        if (relu) v = fmaxf(v, 0.f);
        y[o0 + q] = v;
      }
    }
  }
}

struct HeadsFwdParams {
  HeadsWeights w;
  HeadsSaved sv;       // pointers may be null when nothing has to be kept (inference)
  const float* feat;   // [B,512]
  const float* speed;  // [B]
  const long long* command;  // [B] int64
  float* controls;     // [B,3]
  float* pred_speed;   // [B]
  int batch;
  float dropout_p;
  unsigned long long seed;
  const long long* seed_counter;  // optional device counter mixed into the seed (a new mask on every CUDA-graph replay)
  int* error_flag;     // set to 1 if a command is outside [0,4)
};

__global__ void __launch_bounds__(HD_THREADS) heads_fwd_kernel(const HeadsFwdParams p) {
  __shared__ __align__(16) float x[640];
  __shared__ __align__(16) float h1[256];
  __shared__ __align__(16) float h2[256];
  __shared__ __align__(16) float o3[4];
  // two CTAs per sample: role 0 = speed encoder + command branch -> controls, role 1 = speed predictor -> pred_speed
  // (the two chains are independent; one CTA per sample ran seven dependent GEMVs back to back)
  const int b = blockIdx.x >> 1;
  const int role = blockIdx.x & 1;
  const int t = threadIdx.x;
  long long cmd = p.command[b];
  if (cmd < 0 || cmd > 3) {
    if (t == 0 && p.error_flag) *p.error_flag = 1;
    cmd = cmd < 0 ? 0 : 3;
  }
  const int k = (int)cmd;
  const float keep_scale = p.dropout_p > 0.f ? 1.f / (1.f - p.dropout_p) : 1.f;
  const unsigned long long seed = p.seed + (p.seed_counter ? 0x9E3779B97F4A7C15ull * (unsigned long long)(*p.seed_counter + 1) : 0ull);
  for (int i = t; i < 512; i += HD_THREADS) x[i] = p.feat[(size_t)b * 512 + i];
  if (role == 0) {
  // speed encoder: Linear(1,128) + ReLU + Dropout, Linear(128,128) + ReLU
  if (t < 128) {
    float v = fmaxf(fmaf(p.w.se0_w[t], p.speed[b], p.w.se0_b[t]), 0.f);
    if (p.dropout_p > 0.f) v = drop_keep(seed, b, 0, t, p.dropout_p) ? v * keep_scale : 0.f;
    h1[t] = v;
    if (p.sv.s1) p.sv.s1[(size_t)b * 128 + t] = v;
  }
  __syncthreads();
  gemv_rows(p.w.se3_w, p.w.se3_b, h1, 128, 128, x + 512, true);
  __syncthreads();
  if (p.sv.sfeat && t < 128) p.sv.sfeat[(size_t)b * 128 + t] = x[512 + t];
  // control branch k: Linear(640,256)+ReLU+Drop, Linear(256,256)+ReLU+Drop, Linear(256,3)
  gemv_rows(p.w.br0_w[k], p.w.br0_b[k], x, 640, 256, h1, true);
  __syncthreads();
  if (p.dropout_p > 0.f) h1[t] = drop_keep(seed, b, 1, t, p.dropout_p) ? h1[t] * keep_scale : 0.f;
  if (p.sv.b1) p.sv.b1[(size_t)b * 256 + t] = h1[t];
  __syncthreads();
  gemv_rows(p.w.br3_w[k], p.w.br3_b[k], h1, 256, 256, h2, true);
  __syncthreads();
  if (p.dropout_p > 0.f) h2[t] = drop_keep(seed, b, 2, t, p.dropout_p) ? h2[t] * keep_scale : 0.f;
  if (p.sv.b2) p.sv.b2[(size_t)b * 256 + t] = h2[t];
  __syncthreads();
  gemv_rows(p.w.br6_w[k], p.w.br6_b[k], h2, 256, 3, o3, false);
  __syncthreads();
  if (t < 3) p.controls[(size_t)b * 3 + t] = o3[t];
  return;
  }
  // speed predictor on the visual features only: Linear(512,256)+ReLU+Drop, Linear(256,256)+ReLU, Linear(256,1)
  __syncthreads();
  gemv_rows(p.w.sp0_w, p.w.sp0_b, x, 512, 256, h1, true);
  __syncthreads();
  if (p.dropout_p > 0.f) h1[t] = drop_keep(seed, b, 3, t, p.dropout_p) ? h1[t] * keep_scale : 0.f;
  if (p.sv.p1) p.sv.p1[(size_t)b * 256 + t] = h1[t];
  __syncthreads();
  gemv_rows(p.w.sp3_w, p.w.sp3_b, h1, 256, 256, h2, true);
  __syncthreads();
  if (p.sv.p2) p.sv.p2[(size_t)b * 256 + t] = h2[t];
  gemv_rows(p.w.sp5_w, p.w.sp5_b, h2, 256, 1, o3, false);
  __syncthreads();
  if (t == 0) p.pred_speed[b] = o3[0];
}

// ---------------------------------------------------------------------------------------------
// losses (+ gradients w.r.t. the predictions). mode 0: MSE(controls) + w_speed*MSE(speed)  (README/config recipe)
//                                               mode 1: w0*L1(steer)+w1*L1(thr)+w2*L1(brk) + w_speed*MSE(speed)  (notebook)
// out[0..5] = total, control, steer, throttle, brake, speed   (mode 0 leaves steer/throttle/brake = per-column MSE)
// single CTA, deterministic tree reduction
// ---------------------------------------------------------------------------------------------
struct LossParams {
  const float* controls;  // [B,3]
  const float* pred_speed;
  const float* targets;   // [B,3]
  const float* speed_target;
  int batch;
  int mode;
  float w_steer, w_throttle, w_brake, w_speed;
  float grad_scale;       // upstream d(total) (1.0), also carries 1/world_size if wanted
  float* out;             // [6]
  float* dcontrols;       // [B,3] may be null
  float* dspeed;          // [B]   may be null
};

__global__ void __launch_bounds__(256) loss_kernel(const LossParams p) {
  __shared__ float red[4][256];
  float a[4] = {0.f, 0.f, 0.f, 0.f};
  const float invB = 1.f / (float)p.batch;
  for (int b = threadIdx.x; b < p.batch; b += 256) {
    const float ds = p.pred_speed[b] - p.speed_target[b];
    a[3] += ds * ds;
    if (p.dspeed) p.dspeed[b] = p.grad_scale * p.w_speed * 2.f * ds * invB;
#pragma unroll
    for (int j = 0; j < 3; ++j) {
      const float d = p.controls[b * 3 + j] - p.targets[b * 3 + j];
      const float w = j == 0 ? p.w_steer : (j == 1 ? p.w_throttle : p.w_brake);
      if (p.mode == 0) {
        a[j] += d * d;
        if (p.dcontrols) p.dcontrols[b * 3 + j] = p.grad_scale * 2.f * d * invB * (1.f / 3.f);
      } else {
        a[j] += fabsf(d);
        if (p.dcontrols) p.dcontrols[b * 3 + j] = p.grad_scale * w * (d > 0.f ? 1.f : (d < 0.f ? -1.f : 0.f)) * invB;
      }
    }
  }
#pragma unroll
  for (int j = 0; j < 4; ++j) red[j][threadIdx.x] = a[j];
  __syncthreads();
  for (int s = 128; s > 0; s >>= 1) {
    if (threadIdx.x < s) {
#pragma unroll
      for (int j = 0; j < 4; ++j) red[j][threadIdx.x] += red[j][threadIdx.x + s];
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    const float steer = red[0][0] * invB, thr = red[1][0] * invB, brk = red[2][0] * invB, spd = red[3][0] * invB;
    float control;
    if (p.mode == 0) control = (steer + thr + brk) * (1.f / 3.f);
    else control = p.w_steer * steer + p.w_throttle * thr + p.w_brake * brk;
    p.out[0] = control + p.w_speed * spd;
    p.out[1] = control; p.out[2] = steer; p.out[3] = thr; p.out[4] = brk; p.out[5] = spd;
  }
}

// ---------------------------------------------------------------------------------------------
// validate() on the device (notebook/notebook.ipynb:563-585): per batch the six loss scalars (as loss_kernel) and the
// per-command steer absolute error, ACCUMULATED into acc[16] (fp64):
//   acc[0..5]  += total, control, steer, throttle, brake, speed of this batch      (the reference sums .item()s per batch)
//   acc[6..9]  += sum over samples with command k of |pred_steer - target_steer|    (the reference extends per-command lists)
//   acc[10..13] += number of samples with command k;   acc[14] += 1 (batches)
// One CTA, fixed-order tree reduction: deterministic. The host reads acc once per validation pass instead of
// (6 + up to 4) synchronising copies per batch.
// ---------------------------------------------------------------------------------------------
struct ValidateParams {
  const float* controls;
  const float* pred_speed;
  const float* targets;
  const float* speed_target;
  const long long* command;
  int batch, mode;
  float w_steer, w_throttle, w_brake, w_speed;
  double* acc;
};

__global__ void __launch_bounds__(256) validate_kernel(const ValidateParams p) {
  __shared__ float red[12][256];
  float a[12];
#pragma unroll
  for (int j = 0; j < 12; ++j) a[j] = 0.f;
  for (int b = threadIdx.x; b < p.batch; b += 256) {
    const float ds = p.pred_speed[b] - p.speed_target[b];
    a[3] += ds * ds;
#pragma unroll
    for (int j = 0; j < 3; ++j) {
      const float d = p.controls[b * 3 + j] - p.targets[b * 3 + j];
      a[j] += p.mode == 0 ? d * d : fabsf(d);
    }
    const long long c = p.command[b];
    const float se = fabsf(p.controls[b * 3] - p.targets[b * 3]);
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      if (c == k) { a[4 + k] += se; a[8 + k] += 1.f; }
    }
  }
#pragma unroll
  for (int j = 0; j < 12; ++j) red[j][threadIdx.x] = a[j];
  __syncthreads();
  for (int s = 128; s > 0; s >>= 1) {
    if (threadIdx.x < s) {
#pragma unroll
      for (int j = 0; j < 12; ++j) red[j][threadIdx.x] += red[j][threadIdx.x + s];
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    const float invB = 1.f / (float)p.batch;
    const float steer = red[0][0] * invB, thr = red[1][0] * invB, brk = red[2][0] * invB, spd = red[3][0] * invB;
    const float control = p.mode == 0 ? (steer + thr + brk) * (1.f / 3.f) : p.w_steer * steer + p.w_throttle * thr + p.w_brake * brk;
    p.acc[0] += (double)(control + p.w_speed * spd);
    p.acc[1] += (double)control; p.acc[2] += (double)steer; p.acc[3] += (double)thr; p.acc[4] += (double)brk; p.acc[5] += (double)spd;
#pragma unroll
    for (int k = 0; k < 4; ++k) { p.acc[6 + k] += (double)red[4 + k][0]; p.acc[10 + k] += (double)red[8 + k][0]; }
    p.acc[14] += 1.0;
  }
}

// ---------------------------------------------------------------------------------------------
// heads backward, part 1: per-sample deltas (one CTA per sample) and d(features)
// ---------------------------------------------------------------------------------------------
struct HeadsBwdParams {
  HeadsWeights w;
  HeadsSaved sv;
  const float* dcontrols;  // [B,3]
  const float* dspeed;     // [B]
  const long long* command;
  float* dfeat;            // [B,512]  gradient of the features through the command branch
  float* dfeat2;           // [B,512]  ... through the speed predictor (the consumer adds the two)
  int batch;
  float dropout_p;
};

// y[i] = sum_o W[o*ld + i] * d[o], i < in  (transposed GEMV). Threads = (column quad, row slice): float4 loads of four
// consecutive columns (coalesced across the quad index), `slices` = 256 / (in/4) independent row slices summed through
// `scratch` (>= 1024 floats of shared memory), so the serial chain per thread is out / slices rows.
// Contains __syncthreads(): every thread of the CTA must call it; in % 4 == 0, in <= 512.
CILRS_DEVINL void gemv_cols(const float* __restrict__ W, int ld, const float* d, int in, int out, float* y, float* scratch) {
  const int ncg = in >> 2;
  int slices = HD_THREADS / ncg;
  if (slices > 8) slices = 8;
  const int cg = threadIdx.x % ncg, sl = threadIdx.x / ncg;
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  if (sl < slices) {
    const float4* wp = reinterpret_cast<const float4*>(W) + cg;
    const int ld4 = ld >> 2;
#pragma unroll 8
    for (int o = sl; o < out; o += slices) {
      const float4 w = __ldg(wp + (size_t)o * ld4);
      const float dv = d[o];
      acc.x = fmaf(w.x, dv, acc.x); acc.y = fmaf(w.y, dv, acc.y); acc.z = fmaf(w.z, dv, acc.z); acc.w = fmaf(w.w, dv, acc.w);
    }
    *reinterpret_cast<float4*>(scratch + sl * in + cg * 4) = acc;
  }
  __syncthreads();
  for (int i = threadIdx.x; i < in; i += HD_THREADS) {
    float v = scratch[i];
    for (int k = 1; k < slices; ++k) v += scratch[k * in + i];
    y[i] = v;
  }
  __syncthreads();  // scratch is reused by the next call
}

__global__ void __launch_bounds__(HD_THREADS) heads_bwd_kernel(const HeadsBwdParams p) {
  __shared__ __align__(16) float d_a[256], d_b[256], dx[640], dx2[512], d3[4], scratch[1024];
  // two CTAs per sample (see heads_fwd_kernel): role 0 = command branch + speed encoder, role 1 = speed predictor
  const int b = blockIdx.x >> 1, role = blockIdx.x & 1, t = threadIdx.x;
  long long cmd = p.command[b];
  const int k = cmd < 0 ? 0 : (cmd > 3 ? 3 : (int)cmd);
  const float ks = p.dropout_p > 0.f ? 1.f / (1.f - p.dropout_p) : 1.f;
  if (role == 1) {
    // ---- speed predictor ----
    if (t == 0) {
      d3[0] = p.dspeed[b];
      p.sv.d_sp5[b] = d3[0];
    }
    __syncthreads();
    {
      const float v = p.sv.p2[(size_t)b * 256 + t] > 0.f ? p.w.sp5_w[t] * d3[0] : 0.f;
      d_a[t] = v;
      p.sv.d_sp3[(size_t)b * 256 + t] = v;
    }
    __syncthreads();
    gemv_cols(p.w.sp3_w, 256, d_a, 256, 256, d_b, scratch);
    __syncthreads();
    {
      const float v = p.sv.p1[(size_t)b * 256 + t] > 0.f ? d_b[t] * ks : 0.f;
      d_b[t] = v;
      p.sv.d_sp0[(size_t)b * 256 + t] = v;
    }
    __syncthreads();
    gemv_cols(p.w.sp0_w, 512, d_b, 512, 256, dx2, scratch);
    __syncthreads();
    for (int i = t; i < 512; i += HD_THREADS) p.dfeat2[(size_t)b * 512 + i] = dx2[i];
    return;
  }
  // ---- control branch ----
  if (t < 4) {
    const float v = t < 3 ? p.dcontrols[(size_t)b * 3 + t] : 0.f;
    d3[t] = v;
    p.sv.d_br6[(size_t)b * 4 + t] = v;
  }
  __syncthreads();
  gemv_cols(p.w.br6_w[k], 256, d3, 256, 3, d_a, scratch);
  __syncthreads();
  {
    const float v = p.sv.b2[(size_t)b * 256 + t] > 0.f ? d_a[t] * ks : 0.f;
    d_a[t] = v;
    p.sv.d_br3[(size_t)b * 256 + t] = v;
  }
  __syncthreads();
  gemv_cols(p.w.br3_w[k], 256, d_a, 256, 256, d_b, scratch);
  __syncthreads();
  {
    const float v = p.sv.b1[(size_t)b * 256 + t] > 0.f ? d_b[t] * ks : 0.f;
    d_b[t] = v;
    p.sv.d_br0[(size_t)b * 256 + t] = v;
  }
  __syncthreads();
  gemv_cols(p.w.br0_w[k], 640, d_b, 512, 256, dx, scratch);
  gemv_cols(p.w.br0_w[k] + 512, 640, d_b, 128, 256, dx + 512, scratch);
  __syncthreads();
  for (int i = t; i < 512; i += HD_THREADS) p.dfeat[(size_t)b * 512 + i] = dx[i];
  // ---- speed encoder ----
  if (t < 128) {
    const float v = p.sv.sfeat[(size_t)b * 128 + t] > 0.f ? dx[512 + t] : 0.f;
    d_a[t] = v;
    p.sv.d_se3[(size_t)b * 128 + t] = v;
  }
  __syncthreads();
  gemv_cols(p.w.se3_w, 128, d_a, 128, 128, d_b, scratch);
  __syncthreads();
  if (t < 128) {
    const float v = p.sv.s1[(size_t)b * 128 + t] > 0.f ? d_b[t] * ks : 0.f;
    p.sv.d_se0[(size_t)b * 128 + t] = v;
  }
}

// ---------------------------------------------------------------------------------------------
// heads backward, part 2: weight / bias gradients.  dW[o,i] += sum_b m_b delta[b,o] x[b,i],  db[o] += sum_b m_b delta[b,o]
// m_b = (command[b] == branch) for branch layers, 1 otherwise. One CTA per 16 x 64 tile of one layer's dW.
// ---------------------------------------------------------------------------------------------
struct HeadsWgradJob {
  const float* delta;  // [B, ld_delta]
  const float* x;      // [B, ld_x]
  float* dw;           // [out, in]
  float* db;           // [out]
  int out, in, ld_delta, ld_x;
  int branch;          // -1: all samples
  int tile_begin;      // first CTA index of this job
  int tiles_i;         // number of 64-wide column tiles
};
constexpr int HD_MAX_JOBS = 17;
struct HeadsWgradParams {
  HeadsWgradJob job[HD_MAX_JOBS];
  int num_jobs;
  int batch;
  const long long* command;
  const float* speed;  // x of the first speed-encoder layer (ld 1)
};

__global__ void __launch_bounds__(256) heads_wgrad_kernel(const HeadsWgradParams p) {
  __shared__ float ds[32][17];
  __shared__ float xs[32][65];
  int j = 0;
  while (j + 1 < p.num_jobs && (int)blockIdx.x >= p.job[j + 1].tile_begin) ++j;
  const HeadsWgradJob jb = p.job[j];
  const int tile = blockIdx.x - jb.tile_begin;
  const int ti = tile % jb.tiles_i, to = tile / jb.tiles_i;
  const int o0 = to * 16, i0 = ti * 64;
  const int il = threadIdx.x & 63, og = threadIdx.x >> 6;  // 4 output rows per thread: o0 + og*4 + q
  float acc[4] = {0.f, 0.f, 0.f, 0.f};
  float bacc = 0.f;  // bias partial (threads with il == 0 .. handled below)
  for (int b0 = 0; b0 < p.batch; b0 += 32) {
    // stage deltas (32 x 16) and inputs (32 x 64), masked by the branch selector
    for (int e = threadIdx.x; e < 32 * 16; e += 256) {
      const int bb = e >> 4, oo = e & 15;
      const int b = b0 + bb, o = o0 + oo;
      float v = 0.f;
      if (b < p.batch && o < jb.out && (jb.branch < 0 || p.command[b] == jb.branch)) v = jb.delta[(size_t)b * jb.ld_delta + o];
      ds[bb][oo] = v;
    }
    for (int e = threadIdx.x; e < 32 * 64; e += 256) {
      const int bb = e >> 6, ii = e & 63;
      const int b = b0 + bb, i = i0 + ii;
      xs[bb][ii] = (b < p.batch && i < jb.in) ? jb.x[(size_t)b * jb.ld_x + i] : 0.f;
    }
    __syncthreads();
#pragma unroll 8
    for (int bb = 0; bb < 32; ++bb) {
      const float xv = xs[bb][il];
#pragma unroll
      for (int q = 0; q < 4; ++q) acc[q] = fmaf(ds[bb][og * 4 + q], xv, acc[q]);
    }
    if (ti == 0 && threadIdx.x < 16) {
      for (int bb = 0; bb < 32; ++bb) bacc += ds[bb][threadIdx.x];
    }
    __syncthreads();
  }
  const int i = i0 + il;
  if (i < jb.in) {
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const int o = o0 + og * 4 + q;
      if (o < jb.out) jb.dw[(size_t)o * jb.in + i] += acc[q];
    }
  }
  if (ti == 0 && threadIdx.x < 16 && o0 + (int)threadIdx.x < jb.out) jb.db[o0 + threadIdx.x] += bacc;
}

}  // namespace cilrs
