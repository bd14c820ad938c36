// K3: the CILRS heads in fp32 on CUDA cores (< 0.1 % of the model FLOPs, latency-bound):
//   speed encoder 1->128->128, command-selected control branch 640->256->256->3, speed predictor 512->256->256->1,
//   the two training losses and their gradients (fused into the forward), and the head backward pass.
// Reference: CILRS.forward model/autonomous_drive.py:389-399 (all 4 branches evaluated then gather(0, command));
// only the selected branch is computed here - the others receive exactly zero gradient.
//
// Batched form (round 2): a thread-block CLUSTER of 8 CTAs owns a group of up to 16 samples that share their weights (the
// samples of one command for the branch role, 16 consecutive samples for the speed-predictor role). Every CTA computes a
// 1/8 column slice of each layer for all 16 samples, so each weight byte is read ONCE per sample group (round 1: once per
// sample - 228 MB of L2->SM traffic per forward at B = 128), and hands its slice of the activations to the other seven CTAs
// through distributed shared memory (st.shared::cluster) before a cluster barrier starts the next layer. One launch per
// direction, no global-memory round trips between layers. The layers are latency-bound (few CTAs, dependent chain), so every
// layer issues ALL its weight loads before the first FMA (one L2 round trip per layer).
#pragma once
#include "common.cuh"
#include "pair.cuh"

namespace cilrs {

constexpr int HD_THREADS = 256;
constexpr int HD_G = 16;    // samples per cluster
constexpr int HD_CL = 8;    // CTAs per cluster
// X | H1 | H2 (forward)  /  dA | dB | dC | cross-warp partials [8][HD_G][64] (backward)
constexpr int HD_SMEM_FLOATS = HD_G * 640 + 2 * HD_G * 256;
constexpr int HD_SMEM_BYTES = HD_SMEM_FLOATS * 4;
static_assert(HD_G * 640 + 2 * HD_G * 256 >= 2 * HD_G * 256 + HD_G * 128 + 8 * HD_G * 64, "backward layout fits the forward's allocation");

struct HeadsWeights {
  const float *se0_w, *se0_b, *se3_w, *se3_b;
  const float *br0_w[4], *br0_b[4], *br3_w[4], *br3_b[4], *br6_w[4], *br6_b[4];
  const float *sp0_w, *sp0_b, *sp3_w, *sp3_b, *sp5_w, *sp5_b;
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

// the whole CTA (256 threads) calls this; red = 4 x 256 floats of shared memory
CILRS_DEVINL void loss_body(const LossParams& p, float (*red)[256]) {
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

static __global__ void __launch_bounds__(256) loss_kernel(const LossParams p) {
  __shared__ float red[4][256];
  loss_body(p, red);
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

static __global__ void __launch_bounds__(256) validate_kernel(const ValidateParams p) {
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
// cluster helpers
// ---------------------------------------------------------------------------------------------
CILRS_DEVINL void st_cluster_f32(uint32_t addr, float v) { asm volatile("st.shared::cluster.f32 [%0], %1;" ::"r"(addr), "f"(v) : "memory"); }
CILRS_DEVINL void st_cluster_v4(uint32_t addr, float4 v) {
  asm volatile("st.shared::cluster.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}

// The sample list of a cluster. Branch role: branch cluster c (0 .. groups(batch) + 2) owns the c-th non-empty group when the
// groups of 16 samples with equal (clamped) command are numbered command-major - at most groups(batch) + 3 such groups exist,
// so the grid does not need 4 x groups(batch) clusters; k_out receives the command. Speed-predictor role (branch = false):
// samples [c*16, c*16+16). Every CTA of the cluster computes the same list. Returns the number of samples (0 = nothing to do;
// uniform over the cluster). s_tmp: 12 ints of shared memory. Contains __syncthreads().
CILRS_DEVINL int heads_sample_list(const long long* __restrict__ command, int batch, bool branch, int c, int* s_list, int* s_tmp,
                                   int* error_flag, int& k_out) {
  const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
  int* s_wcnt = s_tmp;        // [8]
  int* s_cnt4 = s_tmp + 8;    // [4]
  if (t < HD_G) s_list[t] = -1;
  if (t < 4) s_cnt4[t] = 0;
  __syncthreads();
  k_out = -1;
  if (!branch) {
    const int b = c * HD_G + t;
    if (t < HD_G && b < batch) s_list[t] = b;
    __syncthreads();
    return max(0, min(HD_G, batch - c * HD_G));
  }
  // pass 1: samples per command
  for (int b0 = 0; b0 < batch; b0 += HD_THREADS) {
    const int b = b0 + t;
    int cm = -1;
    if (b < batch) {
      long long cc = command[b];
      if (cc < 0 || cc > 3) {
        if (error_flag) *error_flag = 1;   // the reference's gather(0, command) would raise; clamp and flag
        cc = cc < 0 ? 0 : 3;
      }
      cm = (int)cc;
    }
#pragma unroll
    for (int kk = 0; kk < 4; ++kk) {
      const unsigned bal = __ballot_sync(0xffffffffu, cm == kk);
      if (lane == 0 && bal) atomicAdd(&s_cnt4[kk], __popc(bal));
    }
  }
  __syncthreads();
  int k = -1, j = 0, pref = 0;
#pragma unroll
  for (int kk = 0; kk < 4; ++kk) {
    const int gk = (s_cnt4[kk] + HD_G - 1) / HD_G;
    if (k < 0 && c < pref + gk) { k = kk; j = c - pref; }
    pref += gk;
  }
  if (k < 0) return 0;
  k_out = k;
  // pass 2: the samples of command k with rank [j*16, j*16+16) in batch order
  int base = 0;
  const int lo = j * HD_G, hi = lo + HD_G;
  for (int b0 = 0; b0 < batch; b0 += HD_THREADS) {
    const int b = b0 + t;
    bool match = false;
    if (b < batch) {
      long long cc = command[b];
      cc = cc < 0 ? 0 : (cc > 3 ? 3 : cc);
      match = (int)cc == k;
    }
    const unsigned bal = __ballot_sync(0xffffffffu, match);
    __syncthreads();   // s_wcnt of the previous iteration has been read
    if (lane == 0) s_wcnt[warp] = __popc(bal);
    __syncthreads();
    int off = base, tot = 0;
#pragma unroll
    for (int w = 0; w < HD_THREADS / 32; ++w) {
      if (w < warp) off += s_wcnt[w];
      tot += s_wcnt[w];
    }
    const int r = off + __popc(bal & ((1u << lane) - 1u));
    if (match && r >= lo && r < hi) s_list[r - lo] = b;
    base += tot;
    if (base >= hi) break;
  }
  __syncthreads();
  return max(0, min(HD_G, base - lo));
}

// v[16] per lane -> v[0] = sum over the 32 lanes of their v[lane & 15]: one full-width add across the two half-warps, then
// recursive halving (31 shuffles instead of 80 for sixteen butterflies)
CILRS_DEVINL void reduce_transpose16(float (&v)[HD_G], int lane) {
#pragma unroll
  for (int i = 0; i < HD_G; ++i) v[i] += __shfl_xor_sync(0xffffffffu, v[i], 16);
#pragma unroll
  for (int s = 8; s >= 1; s >>= 1) {
    const bool up = (lane & s) != 0;
#pragma unroll
    for (int i = 0; i < s; ++i) {
      const float send = up ? v[i] : v[i + s];
      const float keep = up ? v[i + s] : v[i];
      v[i] = keep + __shfl_xor_sync(0xffffffffu, send, s);
    }
  }
}

// One layer slice, forward: this warp computes output rows o0..o0+3 of W [out, in] (row-major, in = nit * 128) for the 16
// samples x[16][ldx] (shared memory). Lanes split K (float4 per lane and step); the weights of step it+1 are in flight while
// step it is multiplied. Afterwards lane l holds the four sums of sample (l & 15).
// ONE copy of this code serves every layer (__noinline__, the K loop is not unrolled): the first version inlined and fully
// unrolled it per call site - 124 KB of straight-line SASS executed once per launch, i.e. an instruction-fetch-bound kernel.
static __device__ __noinline__ float4 rows_dot4(const float* __restrict__ W, int nit, int o0, int out, const float* x, int ldx, int lane) {
  const int in = nit * 128;
  const float4 zero = make_float4(0.f, 0.f, 0.f, 0.f);
  const float4* wrow[4];
#pragma unroll
  for (int q = 0; q < 4; ++q) wrow[q] = reinterpret_cast<const float4*>(W + (size_t)(o0 + q < out ? o0 + q : 0) * in) + lane;
  float4 wn[4];
#pragma unroll
  for (int q = 0; q < 4; ++q) wn[q] = (o0 + q < out) ? __ldg(wrow[q]) : zero;
  float acc[4][HD_G];
#pragma unroll
  for (int q = 0; q < 4; ++q)
#pragma unroll
    for (int t = 0; t < HD_G; ++t) acc[q][t] = 0.f;
#pragma unroll 1
  for (int it = 0; it < nit; ++it) {
    float4 wc[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) wc[q] = wn[q];
    if (it + 1 < nit) {
#pragma unroll
      for (int q = 0; q < 4; ++q) wn[q] = (o0 + q < out) ? __ldg(wrow[q] + 32 * (it + 1)) : zero;
    }
    const float* xp = x + 4 * (lane + 32 * it);
#pragma unroll
    for (int t = 0; t < HD_G; ++t) {
      const float4 x4 = *reinterpret_cast<const float4*>(xp + t * ldx);
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        acc[q][t] = fmaf(wc[q].x, x4.x, acc[q][t]); acc[q][t] = fmaf(wc[q].y, x4.y, acc[q][t]);
        acc[q][t] = fmaf(wc[q].z, x4.z, acc[q][t]); acc[q][t] = fmaf(wc[q].w, x4.w, acc[q][t]);
      }
    }
  }
  float r[4];
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    reduce_transpose16(acc[q], lane);
    r[q] = acc[q][0];
  }
  return make_float4(r[0], r[1], r[2], r[3]);
}

// write this lane's (sample = lane) four consecutive outputs [col, col+4) into row `lane` of the buffer `dst` (row length ld)
// of EVERY CTA of the cluster
CILRS_DEVINL void bcast_row4(float* dst, int ld, int row, int col, float4 v) {
  float* p = dst + row * ld + col;
#pragma unroll
  for (int rk = 0; rk < HD_CL; ++rk) st_cluster_v4(mapa_u32(p, (uint32_t)rk), v);
}

struct HeadsLossFuse {
  int enabled;             // 0: the forward only produces controls / pred_speed
  LossParams lp;           // controls / pred_speed / batch are filled from the forward's own arguments
  unsigned int* counter;   // zeroed device counter, left zero
};

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
  HeadsLossFuse loss;
};

__host__ __device__ inline int heads_groups(int batch) { return (batch + HD_G - 1) / HD_G; }
__host__ __device__ inline int heads_branch_clusters(int batch) { return heads_groups(batch) + 3; }   // >= non-empty (command, group) pairs
inline int heads_grid(int batch) { return (heads_branch_clusters(batch) + heads_groups(batch)) * HD_CL; }   // + the speed-predictor groups

// generic hidden layer of a role: out columns [32*rank + 4*warp, +4) (or 16 per CTA when out == 128), ReLU, optional dropout,
// broadcast into `dst` of every CTA, optional global save. in = nit * 128.
struct HeadsLayerArgs {
  const float* W; const float* bias; int nit, out;
  const float* x; int ldx;
  float* dst; int ldd, dst_col0;
  float* save; int save_ld;
  float dropout_p; int site;
};
static __device__ __noinline__ void heads_layer(const HeadsLayerArgs a, const int* s_list, int cnt, uint32_t rank, float keep_scale,
                                         unsigned long long seed) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int per_cta = a.out / HD_CL;           // 32 or 16
  if (warp * 4 >= per_cta) return;
  const int o0 = (int)rank * per_cta + warp * 4;
  const float4 r4 = rows_dot4(a.W, a.nit, o0, a.out, a.x, a.ldx, lane);
  if (lane >= HD_G) return;                  // lanes 16..31 hold duplicates
  const float r[4] = {r4.x, r4.y, r4.z, r4.w};
  const int b = s_list[lane];
  float v[4];
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    float u = fmaxf(r[q] + __ldg(a.bias + o0 + q), 0.f);
    if (a.dropout_p > 0.f && b >= 0) u = drop_keep(seed, b, a.site, o0 + q, a.dropout_p) ? u * keep_scale : 0.f;
    v[q] = (lane < cnt) ? u : 0.f;
  }
  const float4 v4 = make_float4(v[0], v[1], v[2], v[3]);
  bcast_row4(a.dst, a.ldd, lane, a.dst_col0 + o0, v4);
  if (a.save && lane < cnt) *reinterpret_cast<float4*>(a.save + (size_t)b * a.save_ld + o0) = v4;
}

// optional phase trace of one CTA (tools/heads_trace.cu; compiled in only with -DHD_TRACE=<blockIdx>)
#ifdef HD_TRACE
__device__ unsigned long long g_hd_trace[32];
#define HD_STAMP(i) do { if (blockIdx.x == (HD_TRACE) && threadIdx.x == 0) g_hd_trace[i] = globaltimer_ns(); } while (0)
#else
#define HD_STAMP(i) do { } while (0)
#endif

static __global__ void __launch_bounds__(HD_THREADS, 1) heads_fwd_kernel(const HeadsFwdParams p) {
  extern __shared__ __align__(16) float hd_smem[];
  __shared__ int s_list[HD_G];
  __shared__ int s_tmp[12];
  __shared__ float s_red[4][256];
  __shared__ int s_last;
  float* X = hd_smem;                    // [G][640]  features | speed features
  float* H1 = X + HD_G * 640;            // [G][256]
  float* H2 = H1 + HD_G * 256;           // [G][256]  (its first half doubles as the speed encoder's hidden layer [G][128])
  const int t = threadIdx.x, warp = t >> 5, lane = t & 31;
  const uint32_t rank = cluster_ctarank();
  const int cluster = blockIdx.x / HD_CL;
  const int nbr = heads_branch_clusters(p.batch);
  const bool branch_role = cluster < nbr;
  int k;
  HD_STAMP(0);
  const int cnt = heads_sample_list(p.command, p.batch, branch_role, branch_role ? cluster : cluster - nbr, s_list, s_tmp, p.error_flag, k);
  HD_STAMP(1);
  const float keep_scale = p.dropout_p > 0.f ? 1.f / (1.f - p.dropout_p) : 1.f;
  const unsigned long long seed = p.seed + (p.seed_counter ? 0x9E3779B97F4A7C15ull * (unsigned long long)(*p.seed_counter + 1) : 0ull);
  if (cnt > 0) {   // (uniform over the cluster: empty clusters skip every cluster barrier together)
    // features of the group's samples (rows of absent samples are zero)
#pragma unroll
    for (int e = t; e < HD_G * 128; e += HD_THREADS) {
      const int g = e >> 7, c4 = e & 127;
      const int b = s_list[g];
      const float4 v = b >= 0 ? __ldg(reinterpret_cast<const float4*>(p.feat + (size_t)b * 512) + c4) : make_float4(0.f, 0.f, 0.f, 0.f);
      *reinterpret_cast<float4*>(X + g * 640 + 4 * c4) = v;
    }
    if (branch_role) {
      // speed encoder layer 0: Linear(1,128) + ReLU + Dropout - cheap, every CTA computes all of it
      for (int e = t; e < HD_G * 128; e += HD_THREADS) {
        const int g = e >> 7, c = e & 127;
        const int b = s_list[g];
        float v = 0.f;
        if (b >= 0) {
          v = fmaxf(fmaf(__ldg(p.w.se0_w + c), __ldg(p.speed + b), __ldg(p.w.se0_b + c)), 0.f);
          if (p.dropout_p > 0.f) v = drop_keep(seed, b, 0, c, p.dropout_p) ? v * keep_scale : 0.f;
          if (p.sv.s1 && rank == 0) p.sv.s1[(size_t)b * 128 + c] = v;
        }
        H2[g * 128 + c] = v;
      }
      __syncthreads();
      HD_STAMP(2);
      cluster_sync_all();   // every CTA's X / H2 are written before peers start to store speed features into them
      HD_STAMP(3);
      // speed encoder layer 3: Linear(128,128) + ReLU -> X[:, 512:640] of every CTA
      heads_layer(HeadsLayerArgs{p.w.se3_w, p.w.se3_b, 1, 128, H2, 128, X, 640, 512, p.sv.sfeat, 128, 0.f, 0}, s_list, cnt, rank, 1.f, seed);
      HD_STAMP(4);
      cluster_sync_all();
      HD_STAMP(5);
      // branch k: Linear(640,256)+ReLU+Drop, Linear(256,256)+ReLU+Drop, Linear(256,3)
      heads_layer(HeadsLayerArgs{p.w.br0_w[k], p.w.br0_b[k], 5, 256, X, 640, H1, 256, 0, p.sv.b1, 256, p.dropout_p, 1}, s_list, cnt, rank, keep_scale, seed);
      HD_STAMP(6);
      cluster_sync_all();
      HD_STAMP(7);
      heads_layer(HeadsLayerArgs{p.w.br3_w[k], p.w.br3_b[k], 2, 256, H1, 256, H2, 256, 0, p.sv.b2, 256, p.dropout_p, 2}, s_list, cnt, rank, keep_scale, seed);
      HD_STAMP(8);
      cluster_sync_all();
      HD_STAMP(9);
      if (rank == 0 && warp == 0) {
        const float4 r4 = rows_dot4(p.w.br6_w[k], 2, 0, 3, H2, 256, lane);
        if (lane < cnt) {
          const int b = s_list[lane];
          p.controls[(size_t)b * 3 + 0] = r4.x + __ldg(p.w.br6_b[k] + 0);
          p.controls[(size_t)b * 3 + 1] = r4.y + __ldg(p.w.br6_b[k] + 1);
          p.controls[(size_t)b * 3 + 2] = r4.z + __ldg(p.w.br6_b[k] + 2);
        }
      }
    } else {
      __syncthreads();
      cluster_sync_all();
      // speed predictor on the visual features only: Linear(512,256)+ReLU+Drop, Linear(256,256)+ReLU, Linear(256,1)
      heads_layer(HeadsLayerArgs{p.w.sp0_w, p.w.sp0_b, 4, 256, X, 640, H1, 256, 0, p.sv.p1, 256, p.dropout_p, 3}, s_list, cnt, rank, keep_scale, seed);
      cluster_sync_all();
      heads_layer(HeadsLayerArgs{p.w.sp3_w, p.w.sp3_b, 2, 256, H1, 256, H2, 256, 0, p.sv.p2, 256, 0.f, 0}, s_list, cnt, rank, 1.f, seed);
      cluster_sync_all();
      if (rank == 0 && warp == 0) {
        const float4 r4 = rows_dot4(p.w.sp5_w, 2, 0, 1, H2, 256, lane);
        if (lane < cnt) p.pred_speed[s_list[lane]] = r4.x + __ldg(p.w.sp5_b);
      }
    }
    HD_STAMP(10);
    cluster_sync_all();   // nobody leaves while a peer may still write into its shared memory
    HD_STAMP(11);
  }
  // ---- fused loss: the last cluster to finish (rank 0 CTAs count) reduces over the whole batch ----
  if (p.loss.enabled && rank == 0) {
    __syncthreads();
    if (t == 0) {
      __threadfence();
      const unsigned int done = atomicAdd(p.loss.counter, 1u);
      const int last = done == gridDim.x / HD_CL - 1;
      if (last) __threadfence();
      s_last = last;
    }
    __syncthreads();
    if (s_last) {
      LossParams lp = p.loss.lp;
      lp.controls = p.controls; lp.pred_speed = p.pred_speed; lp.batch = p.batch;
      loss_body(lp, s_red);
      if (t == 0) *p.loss.counter = 0u;   // ready for the next launch / graph replay
    }
  }
}

// ---------------------------------------------------------------------------------------------
// heads backward, part 1: per-sample deltas and d(features), same cluster decomposition.
// A transposed layer y[g][c] = sum_o W[o][c] d[g][o] is split by COLUMNS c over the CTAs; thread = (column, sample subgroup),
// so the weight loads of a warp are contiguous and every output is one thread's fixed-order sum (deterministic).
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

// Transposed layer slice: y[g][c] = sum_{o < K} W[o*ld + c0 + c] * d[g*ldd + o] for c < WD, g < HD_G.
// The K rows are split over the 8 warps and, inside a warp, over RPW = 128 / WD lane groups (a warp reads RPW whole row slices
// per step: contiguous float4 loads), the next step's weights in flight while the current ones are multiplied. The partial
// sums are combined by shuffles inside the warp and across warps through `part` ([8][HD_G][WD] floats) in a FIXED order
// (deterministic). On return thread t < HD_G * WD / 4 holds y[g][c..c+3] with g = t / (WD/4), c = 4 * (t % (WD/4)) in `res`.
// Contains __syncthreads(): the whole CTA calls it. Not inlined, K loop not unrolled (code size, see rows_dot4).
template <int WD>
__device__ __noinline__ bool cols_dot(const float* __restrict__ W, int ld, int c0, int K, const float* d, int ldd, float* part, float4& res,
                                      int& g_out, int& c_out) {
  constexpr int QW = WD / 4, RPW = 32 / QW;
  const int nstep = K / (8 * RPW);
  const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
  const int q = lane % QW, rg = lane / QW;
  const float4* wp = reinterpret_cast<const float4*>(W + (size_t)(warp * RPW + rg) * ld + c0) + q;
  const size_t wstep = (size_t)8 * RPW * ld / 4;   // float4 elements between this thread's consecutive rows
  float4 acc[HD_G];
#pragma unroll
  for (int g = 0; g < HD_G; ++g) acc[g] = make_float4(0.f, 0.f, 0.f, 0.f);
  float4 wn = __ldg(wp);
#pragma unroll 1
  for (int i = 0; i < nstep; ++i) {
    const float4 wc = wn;
    if (i + 1 < nstep) wn = __ldg(wp + (size_t)(i + 1) * wstep);
    const int o = (i * 8 + warp) * RPW + rg;
#pragma unroll
    for (int g = 0; g < HD_G; ++g) {
      const float dv = d[g * ldd + o];
      acc[g].x = fmaf(wc.x, dv, acc[g].x); acc[g].y = fmaf(wc.y, dv, acc[g].y);
      acc[g].z = fmaf(wc.z, dv, acc[g].z); acc[g].w = fmaf(wc.w, dv, acc[g].w);
    }
  }
  // lanes with the same column quad (different row group) -> every lane
#pragma unroll
  for (int sft = QW; sft < 32; sft <<= 1) {
#pragma unroll
    for (int g = 0; g < HD_G; ++g) {
      acc[g].x += __shfl_xor_sync(0xffffffffu, acc[g].x, sft); acc[g].y += __shfl_xor_sync(0xffffffffu, acc[g].y, sft);
      acc[g].z += __shfl_xor_sync(0xffffffffu, acc[g].z, sft); acc[g].w += __shfl_xor_sync(0xffffffffu, acc[g].w, sft);
    }
  }
  __syncthreads();   // `part` is free (previous call's readers are done)
  if (rg == 0) {
#pragma unroll
    for (int g = 0; g < HD_G; ++g) *reinterpret_cast<float4*>(part + (warp * HD_G + g) * WD + 4 * q) = acc[g];
  }
  __syncthreads();
  const bool active = t < HD_G * QW;
  g_out = t / QW; c_out = 4 * (t % QW);
  res = make_float4(0.f, 0.f, 0.f, 0.f);
  if (active) {
#pragma unroll
    for (int w = 0; w < 8; ++w) {
      const float4 v = *reinterpret_cast<const float4*>(part + (w * HD_G + g_out) * WD + c_out);
      res.x += v.x; res.y += v.y; res.z += v.z; res.w += v.w;
    }
  }
  return active;
}

// store four consecutive values of (sample g, columns col..col+3) into `dst[g*ld + col]` of every CTA of the cluster
CILRS_DEVINL void bcast4(float* dst, int ld, int g, int col, float4 v) { bcast_row4(dst, ld, g, col, v); }

CILRS_DEVINL float4 mask4(const float* act, float4 v, float scale) {
  const float4 a = *reinterpret_cast<const float4*>(act);
  return make_float4(a.x > 0.f ? v.x * scale : 0.f, a.y > 0.f ? v.y * scale : 0.f, a.z > 0.f ? v.z * scale : 0.f, a.w > 0.f ? v.w * scale : 0.f);
}

static __global__ void __launch_bounds__(HD_THREADS, 1) heads_bwd_kernel(const HeadsBwdParams p) {
  extern __shared__ __align__(16) float hd_smem[];
  __shared__ int s_list[HD_G];
  __shared__ int s_tmp[12];
  __shared__ float s_d3[HD_G][4];
  float* dA = hd_smem;                 // [G][256]
  float* dB = dA + HD_G * 256;         // [G][256]
  float* dC = dB + HD_G * 256;         // [G][128]
  float* part = dC + HD_G * 128;       // [8][G][64]
  const int t = threadIdx.x;
  const uint32_t rank = cluster_ctarank();
  const int cluster = blockIdx.x / HD_CL;
  const int nbr = heads_branch_clusters(p.batch);
  const bool branch_role = cluster < nbr;
  int k;
  const int cnt = heads_sample_list(p.command, p.batch, branch_role, branch_role ? cluster : cluster - nbr, s_list, s_tmp, nullptr, k);
  if (cnt == 0) return;   // uniform over the cluster
  const float ks = p.dropout_p > 0.f ? 1.f / (1.f - p.dropout_p) : 1.f;
  const float4 zero4 = make_float4(0.f, 0.f, 0.f, 0.f);
  float4 res;
  int g, c;
  if (branch_role) {
    // ---- control branch ----
    if (t < HD_G * 4) {
      const int gg = t >> 2, q = t & 3;
      const int b = s_list[gg];
      const float v = (b >= 0 && q < 3) ? p.dcontrols[(size_t)b * 3 + q] : 0.f;
      s_d3[gg][q] = v;
      if (b >= 0 && rank == 0) p.sv.d_br6[(size_t)b * 4 + q] = v;
    }
    __syncthreads();
    {  // delta of layer .3's output: every CTA needs all 256 columns, three terms each - computed redundantly
      const float w0 = __ldg(p.w.br6_w[k] + t), w1 = __ldg(p.w.br6_w[k] + 256 + t), w2 = __ldg(p.w.br6_w[k] + 512 + t);
      for (int gg = 0; gg < HD_G; ++gg) {
        const int b = s_list[gg];
        float v = 0.f;
        if (b >= 0 && p.sv.b2[(size_t)b * 256 + t] > 0.f) v = (w0 * s_d3[gg][0] + w1 * s_d3[gg][1] + w2 * s_d3[gg][2]) * ks;
        dA[gg * 256 + t] = v;
        if (b >= 0 && (t >> 5) == (int)rank) p.sv.d_br3[(size_t)b * 256 + t] = v;
      }
    }
    __syncthreads();
    cluster_sync_all();   // peers' dB / dC are free to be written
    // delta of layer .0's output, this CTA's 32 columns
    if (cols_dot<32>(p.w.br3_w[k], 256, 32 * (int)rank, 256, dA, 256, part, res, g, c)) {
      const int b = s_list[g], col = 32 * (int)rank + c;
      const float4 v = b >= 0 ? mask4(p.sv.b1 + (size_t)b * 256 + col, res, ks) : zero4;
      bcast4(dB, 256, g, col, v);
      if (b >= 0) *reinterpret_cast<float4*>(p.sv.d_br0 + (size_t)b * 256 + col) = v;
    }
    cluster_sync_all();
    // d(features): 64 of the 512 feature columns of W0
    if (cols_dot<64>(p.w.br0_w[k], 640, 64 * (int)rank, 256, dB, 256, part, res, g, c)) {
      const int b = s_list[g];
      if (b >= 0) *reinterpret_cast<float4*>(p.dfeat + (size_t)b * 512 + 64 * (int)rank + c) = res;
    }
    // d(speed features): 16 of the 128 columns 512..639 of W0, through the speed encoder's last ReLU
    if (cols_dot<16>(p.w.br0_w[k], 640, 512 + 16 * (int)rank, 256, dB, 256, part, res, g, c)) {
      const int b = s_list[g], col = 16 * (int)rank + c;
      const float4 v = b >= 0 ? mask4(p.sv.sfeat + (size_t)b * 128 + col, res, 1.f) : zero4;
      bcast4(dC, 128, g, col, v);
      if (b >= 0) *reinterpret_cast<float4*>(p.sv.d_se3 + (size_t)b * 128 + col) = v;
    }
    cluster_sync_all();
    // speed encoder layer 0 delta
    if (cols_dot<16>(p.w.se3_w, 128, 16 * (int)rank, 128, dC, 128, part, res, g, c)) {
      const int b = s_list[g], col = 16 * (int)rank + c;
      if (b >= 0) *reinterpret_cast<float4*>(p.sv.d_se0 + (size_t)b * 128 + col) = mask4(p.sv.s1 + (size_t)b * 128 + col, res, ks);
    }
  } else {
    // ---- speed predictor ----
    if (t < HD_G) {
      const int b = s_list[t];
      const float v = b >= 0 ? p.dspeed[b] : 0.f;
      s_d3[t][0] = v;
      if (b >= 0 && rank == 0) p.sv.d_sp5[b] = v;
    }
    __syncthreads();
    {
      const float w = __ldg(p.w.sp5_w + t);
      for (int gg = 0; gg < HD_G; ++gg) {
        const int b = s_list[gg];
        float v = 0.f;
        if (b >= 0 && p.sv.p2[(size_t)b * 256 + t] > 0.f) v = w * s_d3[gg][0];
        dA[gg * 256 + t] = v;
        if (b >= 0 && (t >> 5) == (int)rank) p.sv.d_sp3[(size_t)b * 256 + t] = v;
      }
    }
    __syncthreads();
    cluster_sync_all();
    if (cols_dot<32>(p.w.sp3_w, 256, 32 * (int)rank, 256, dA, 256, part, res, g, c)) {
      const int b = s_list[g], col = 32 * (int)rank + c;
      const float4 v = b >= 0 ? mask4(p.sv.p1 + (size_t)b * 256 + col, res, ks) : zero4;
      bcast4(dB, 256, g, col, v);
      if (b >= 0) *reinterpret_cast<float4*>(p.sv.d_sp0 + (size_t)b * 256 + col) = v;
    }
    cluster_sync_all();
    if (cols_dot<64>(p.w.sp0_w, 512, 64 * (int)rank, 256, dB, 256, part, res, g, c)) {
      const int b = s_list[g];
      if (b >= 0) *reinterpret_cast<float4*>(p.dfeat2 + (size_t)b * 512 + 64 * (int)rank + c) = res;
    }
  }
  cluster_sync_all();   // nobody leaves while a peer may still write into its shared memory
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

static __global__ void __launch_bounds__(256) heads_wgrad_kernel(const HeadsWgradParams p) {
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
      if (b < p.batch && o < jb.out) {
        long long c = p.command[b];
        c = c < 0 ? 0 : (c > 3 ? 3 : c);   // same clamp as the forward (which also raised the error flag)
        if (jb.branch < 0 || c == jb.branch) v = jb.delta[(size_t)b * jb.ld_delta + o];
      }
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

// launch with a cluster of HD_CL CTAs and the dynamic shared memory both kernels need
template <typename P>
inline cudaError_t heads_launch_cluster(void (*kernel)(P), int batch, cudaStream_t s, const P& prm) {
  // The kernels are `static` (one copy per translation unit that includes this header: the bf16 plan and the fp32 plan) while
  // this template has ONE instance per parameter type across the library, so the attribute is tracked per kernel pointer.
  static const void* attr_done[4] = {nullptr, nullptr, nullptr, nullptr};
  bool done = false;
  for (int i = 0; i < 4; ++i) done = done || attr_done[i] == (const void*)kernel;
  if (!done) {
    cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, HD_SMEM_BYTES);
    if (e != cudaSuccess) return e;
    for (int i = 0; i < 4; ++i)
      if (!attr_done[i]) { attr_done[i] = (const void*)kernel; break; }
  }
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)heads_grid(batch)); cfg.blockDim = dim3(HD_THREADS); cfg.dynamicSmemBytes = HD_SMEM_BYTES; cfg.stream = s;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = HD_CL; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, kernel, prm);
}

}  // namespace cilrs
