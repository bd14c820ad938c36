// HBM-bound kernels around the conv stack: BatchNorm statistics / apply / backward, ReLU, residual add,
// max-pool (3x3/2, stem) and global average pool. All activations are NHWC bf16, C a multiple of 64; every
// thread moves 16-byte vectors (8 channels) and keeps its channel group fixed so per-channel parameters live in
// registers. Reductions are deterministic (fixed partial order, no float atomics).
#pragma once
#include "common.cuh"
#include "conv_params.h"
#include "bn_math.cuh"

namespace cilrs {

constexpr int EW_THREADS = 256;
constexpr int EW_MAX_BLOCKS = 148 * 4;
constexpr int EW_DEFER_MAX_C = 512;  // deferred BatchNorm finalize stages the per-channel values of a CTA in shared memory

struct Vec8 {
  float v[8];
};
CILRS_DEVINL Vec8 load8(const __nv_bfloat16* p) {
  const uint4 u = *reinterpret_cast<const uint4*>(p);
  Vec8 r;
  r.v[0] = bf16lo(u.x); r.v[1] = bf16hi(u.x); r.v[2] = bf16lo(u.y); r.v[3] = bf16hi(u.y);
  r.v[4] = bf16lo(u.z); r.v[5] = bf16hi(u.z); r.v[6] = bf16lo(u.w); r.v[7] = bf16hi(u.w);
  return r;
}
CILRS_DEVINL Vec8 unpack8(const uint4& u) {
  Vec8 r;
  r.v[0] = bf16lo(u.x); r.v[1] = bf16hi(u.x); r.v[2] = bf16lo(u.y); r.v[3] = bf16hi(u.y);
  r.v[4] = bf16lo(u.z); r.v[5] = bf16hi(u.z); r.v[6] = bf16lo(u.w); r.v[7] = bf16hi(u.w);
  return r;
}
CILRS_DEVINL void store8(__nv_bfloat16* p, const Vec8& r) {
  uint4 u;
  u.x = pack_bf16x2(r.v[0], r.v[1]); u.y = pack_bf16x2(r.v[2], r.v[3]);
  u.z = pack_bf16x2(r.v[4], r.v[5]); u.w = pack_bf16x2(r.v[6], r.v[7]);
  *reinterpret_cast<uint4*>(p) = u;
}
CILRS_DEVINL Vec8 loadf8(const float* p) {
  Vec8 r;
  const float4 a = *reinterpret_cast<const float4*>(p), b = *reinterpret_cast<const float4*>(p + 4);
  r.v[0] = a.x; r.v[1] = a.y; r.v[2] = a.z; r.v[3] = a.w; r.v[4] = b.x; r.v[5] = b.y; r.v[6] = b.z; r.v[7] = b.w;
  return r;
}

// Walks the pixels a thread visits (constant pixel stride) through the padded-flat layout (conv_params.h: PadGeom) and tells
// whether the current one is a real pixel or a padding pixel, which must be written as zero and never read.
// A dense tensor is the geometry {1, 1, 1, 1}: every pixel is valid.
struct PadWalk {
  unsigned int w, h, sw, sh, Wp, Hp, W, H;
  CILRS_DEVINL void init(long long pix, long long stride, const PadGeom& g) {
    Wp = (unsigned int)g.Wp; Hp = (unsigned int)g.Hp; W = (unsigned int)g.W; H = (unsigned int)g.H;
    w = (unsigned int)(pix % Wp);
    h = (unsigned int)((pix / Wp) % Hp);
    sw = (unsigned int)(stride % Wp);
    sh = (unsigned int)((stride / Wp) % Hp);
  }
  CILRS_DEVINL bool valid() const { return w < W && h < H; }
  CILRS_DEVINL void next() {
    w += sw; h += sh;
    if (w >= Wp) { w -= Wp; ++h; }
    if (h >= Hp) h -= Hp;
  }
};
CILRS_DEVINL void store8_zero(__nv_bfloat16* p) { *reinterpret_cast<uint4*>(p) = make_uint4(0, 0, 0, 0); }

// per-BatchNorm derived vectors, each [C]: scale, shift (forward apply), mean, rstd (backward)
struct BnVectors {
  float* scale;
  float* shift;
  float* mean;
  float* rstd;
};

// Deferred finalize (conv_params.h: CF_DEFER). The fused flat-conv epilogues only add their per-channel fp64 sums to a
// per-BatchNorm accumulator; the elementwise kernel that consumes the statistics derives its own channels' values from the
// sums in its prologue (every thread runs the same arithmetic, so all copies agree bit for bit), and CTA 0 also writes the
// vectors / running statistics / parameter gradients that later kernels read. acc == nullptr: nothing is deferred.
struct BnDefer {
  const double* acc;  // [2][C]: sum, sum of squares of the conv output
  const float* gamma;
  const float* beta;
  float* running_mean;
  float* running_var;
  long long* nbt;
  float* vec;         // [4][C] out: scale, shift, mean, rstd
  double inv_count;
  double unbias;      // count / (count - 1)
  float momentum, eps;
  int update_running;
};
struct BnBwdDefer {
  const double* acc_sum;  // [C] sum dz
  const double* acc_dot;  // [C] sum dz * y (raw conv output); bdot = rstd * (acc_dot - mean * acc_sum)
  float* bred;            // [2][C] out (bsum, bdot)
  float* dgamma;          // += bdot (may be null)
  float* dbeta;           // += bsum
};
// ---------------------------------------------------------------------------------------------
// bn_finalize: per-tile (sum, sumsq) partials -> batch statistics -> scale/shift (+ running stats update)
//   training: mean = S/n, var_b = Q/n - mean^2 (biased, used to normalise), running_var gets the unbiased one,
//             momentum 0.1, num_batches_tracked += 1  (torch.nn.BatchNorm2d semantics, torchvision resnet BN layers)
//   frozen  : statistics come from running_mean / running_var (eval mode)
// one CTA per 32 channels, 32 tile-slices per channel, double accumulation
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(1024) bn_finalize_kernel(const float* __restrict__ partials, int tiles, int C, double count,
                                                           const float* __restrict__ gamma, const float* __restrict__ beta,
                                                           float* running_mean, float* running_var, long long* nbt,
                                                           float momentum, float eps, int training, int update_running,
                                                           BnVectors out) {
  __shared__ double s_sum[32][33], s_sq[32][33];
  pdl_entry();
  const int cl = threadIdx.x & 31, slice = threadIdx.x >> 5;
  const int c = blockIdx.x * 32 + cl;
  double s = 0.0, q = 0.0;
  if (training && c < C) {
    for (int t = slice; t < tiles; t += 32) {
      s += (double)partials[(size_t)t * 2 * C + c];
      q += (double)partials[(size_t)t * 2 * C + C + c];
    }
  }
  s_sum[slice][cl] = s;
  s_sq[slice][cl] = q;
  __syncthreads();
  if (slice == 0 && c < C) {
    float mean, var;
    if (training) {
      for (int i = 1; i < 32; ++i) { s += s_sum[i][cl]; q += s_sq[i][cl]; }
      const double m = s / count;
      double v = q / count - m * m;
      if (v < 0.0) v = 0.0;
      mean = (float)m;
      var = (float)v;
      if (update_running) {
        const double unbiased = count > 1.0 ? v * count / (count - 1.0) : v;
        running_mean[c] = (1.f - momentum) * running_mean[c] + momentum * mean;
        running_var[c] = (1.f - momentum) * running_var[c] + momentum * (float)unbiased;
      }
    } else {
      mean = running_mean[c];
      var = running_var[c];
    }
    const float rstd = 1.0f / sqrtf(var + eps);
    const float sc = gamma[c] * rstd;
    out.scale[c] = sc;
    out.shift[c] = beta[c] - mean * sc;
    out.mean[c] = mean;
    out.rstd[c] = rstd;
  }
  if (training && update_running && nbt && blockIdx.x == 0 && threadIdx.x == 0) *nbt += 1;
}

// ---------------------------------------------------------------------------------------------
// bn_apply: out = relu?( x*scale + shift  [+ res]  [+ x2*scale2 + shift2] )
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(EW_THREADS) bn_apply_kernel(const __nv_bfloat16* __restrict__ x, const float* __restrict__ scale,
                                                              const float* __restrict__ shift, const __nv_bfloat16* __restrict__ res,
                                                              const __nv_bfloat16* __restrict__ x2, const float* __restrict__ scale2,
                                                              const float* __restrict__ shift2, __nv_bfloat16* __restrict__ out,
                                                              long long nvec, int C, int relu, const PadGeom g,
                                                              uint8_t* __restrict__ bits, const BnDefer d) {
  // bits (optional): one byte per 8-channel vector, bit k = (out[channel k] > 0): the ReLU mask the flat dgrad epilogue
  // applies, 1/16 of the bytes of the activation itself
  pdl_entry();
  const int groups = C >> 3;
  const long long stride = (long long)gridDim.x * EW_THREADS;  // multiple of groups (host guarantees)
  long long i = (long long)blockIdx.x * EW_THREADS + threadIdx.x;
  const int cg = (int)(i % groups) * 8;
  Vec8 sc2, sh2;
  if (x2) { sc2 = loadf8(scale2 + cg); sh2 = loadf8(shift2 + cg); }
  PadWalk walk;
  walk.init(i / groups, stride / groups, g);
  // Four vectors per iteration with every load issued before the first use: ~3 CTAs x 256 threads x 4 x (1..2) x 16 B keep
  // the ~45 KB per SM in flight that HBM latency x bandwidth asks for. The first batch is requested BEFORE the deferred
  // finalize below, whose chain (sums from L2 -> fp64 arithmetic -> shared memory -> barrier) then overlaps its latency.
  bool in[4], ok[4];
  uint4 xq[4], rq[4], yq[4];
#define CILRS_BN_APPLY_LOAD()                                                     \
  {                                                                               \
    _Pragma("unroll") for (int u = 0; u < 4; ++u) {                               \
      in[u] = i + u * stride < nvec;                                              \
      ok[u] = in[u] && walk.valid();                                              \
      walk.next();                                                                \
    }                                                                             \
    _Pragma("unroll") for (int u = 0; u < 4; ++u) {                               \
      if (ok[u]) {                                                                \
        const long long o = (i + u * stride) * 8;                                 \
        xq[u] = *reinterpret_cast<const uint4*>(x + o);                           \
        if (res) rq[u] = *reinterpret_cast<const uint4*>(res + o);                \
        if (x2) yq[u] = *reinterpret_cast<const uint4*>(x2 + o);                  \
      }                                                                           \
    }                                                                             \
  }
  CILRS_BN_APPLY_LOAD()
  Vec8 sc, sh;
  __shared__ __align__(16) float s_par[2][EW_DEFER_MAX_C];
  if (d.acc) {
    // scale / shift straight from the conv epilogue's sums: each CTA derives every channel once (coalesced loads of the
    // sums; per-thread strided 8-byte loads cost 256 L1 line lookups per warp) and threads pick their channels from shared memory
    for (int c = threadIdx.x; c < C; c += EW_THREADS) {
      const BnStat st = bn_stat_from_sums(d.acc[c], d.acc[C + c], d.inv_count, d.unbias, d.eps, d.gamma[c], d.beta[c]);
      s_par[0][c] = st.scale; s_par[1][c] = st.shift;
      if (blockIdx.x == 0) {
        d.vec[c] = st.scale; d.vec[C + c] = st.shift; d.vec[2 * C + c] = st.mean; d.vec[3 * C + c] = st.rstd;
        if (d.update_running) {
          d.running_mean[c] = (1.f - d.momentum) * d.running_mean[c] + d.momentum * st.mean;
          d.running_var[c] = (1.f - d.momentum) * d.running_var[c] + d.momentum * st.unbiased_var;
        }
      }
    }
    if (blockIdx.x == 0 && threadIdx.x == 0 && d.update_running && d.nbt) *d.nbt += 1;
    __syncthreads();
    sc = loadf8(&s_par[0][cg]); sh = loadf8(&s_par[1][cg]);
  } else {
    sc = loadf8(scale + cg); sh = loadf8(shift + cg);
  }
  for (;;) {
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const long long iv = i + u * stride;
      if (!ok[u]) {
        if (in[u]) {
          store8_zero(out + iv * 8);
          if (bits) bits[iv] = 0;
        }
        continue;
      }
      Vec8 a = unpack8(xq[u]);
#pragma unroll
      for (int k = 0; k < 8; ++k) a.v[k] = fmaf(a.v[k], sc.v[k], sh.v[k]);
      if (res) {
        const Vec8 r = unpack8(rq[u]);
#pragma unroll
        for (int k = 0; k < 8; ++k) a.v[k] += r.v[k];
      }
      if (x2) {
        const Vec8 r = unpack8(yq[u]);
#pragma unroll
        for (int k = 0; k < 8; ++k) a.v[k] += fmaf(r.v[k], sc2.v[k], sh2.v[k]);
      }
      if (relu) {
#pragma unroll
        for (int k = 0; k < 8; ++k) a.v[k] = fmaxf(a.v[k], 0.f);
      }
      store8(out + iv * 8, a);
      if (bits) {
        unsigned int m = 0;
#pragma unroll
        for (int k = 0; k < 8; ++k) m |= (__bfloat162float(__float2bfloat16(a.v[k])) > 0.f ? 1u : 0u) << k;
        bits[iv] = (uint8_t)m;
      }
    }
    i += 4 * stride;
    if (i >= nvec) break;
    CILRS_BN_APPLY_LOAD()
  }
#undef CILRS_BN_APPLY_LOAD
}

// (the stem's max-pool forward lives in stem_pool.cuh)

// global average pool: x padded-flat [B,Hp,Wp,C] bf16 -> feat [B,C] fp32 (padding pixels are zero, so the sum runs over all of them)
__global__ void avgpool_kernel(const __nv_bfloat16* __restrict__ x, float* __restrict__ feat, int B, int C, const PadGeom g) {
  pdl_entry();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= B * C) return;
  const int c = i % C, n = i / C;
  const int PP = g.Hp * g.Wp;
  float s = 0.f;
  for (int p = 0; p < PP; ++p) s += __bfloat162float(x[((long long)n * PP + p) * C + c]);
  feat[i] = s / (float)(g.H * g.W);
}
// backward: g padded-flat [B,Hp,Wp,C] bf16 = dfeat[b,c] / (H*W) on real pixels, 0 on padding
// (dfeat2: optional second addend - the heads backward produces the feature gradient in two parts)
__global__ void avgpool_bwd_kernel(const float* __restrict__ dfeat, const float* __restrict__ dfeat2, __nv_bfloat16* __restrict__ gout, int B,
                                   int C, const PadGeom g) {
  pdl_entry();
  // 8 channels (one 16-byte store) per thread; C is a multiple of 8 (512)
  const int PP = g.Hp * g.Wp, C8 = C >> 3;
  const long long total = (long long)B * PP * C8;
  const float inv = 1.f / (float)(g.H * g.W);
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % C8) * 8;
    const long long pix = i / C8;
    const int n = (int)(pix / PP);
    const int q = (int)(pix - (long long)n * PP);
    const bool valid = (q % g.Wp) < g.W && (q / g.Wp) < g.H;
    uint4 o = make_uint4(0u, 0u, 0u, 0u);
    if (valid) {
      const float4* a = (const float4*)(dfeat + (long long)n * C + c);
      float4 lo = a[0], hi = a[1];
      if (dfeat2) {
        const float4* a2 = (const float4*)(dfeat2 + (long long)n * C + c);
        const float4 l2 = a2[0], h2 = a2[1];
        lo.x += l2.x; lo.y += l2.y; lo.z += l2.z; lo.w += l2.w;
        hi.x += h2.x; hi.y += h2.y; hi.z += h2.z; hi.w += h2.w;
      }
      __nv_bfloat162 p0 = __floats2bfloat162_rn(lo.x * inv, lo.y * inv), p1 = __floats2bfloat162_rn(lo.z * inv, lo.w * inv);
      __nv_bfloat162 p2 = __floats2bfloat162_rn(hi.x * inv, hi.y * inv), p3 = __floats2bfloat162_rn(hi.z * inv, hi.w * inv);
      o.x = *(uint32_t*)&p0; o.y = *(uint32_t*)&p1; o.z = *(uint32_t*)&p2; o.w = *(uint32_t*)&p3;
    }
    ((uint4*)gout)[i] = o;
  }
}

// ---------------------------------------------------------------------------------------------
// BatchNorm backward, pass 1 (reduce):  dz = g * (act > 0)   [act == nullptr: no ReLU in between]
//   bsum[c] = sum dz,  bdot[c] = sum dz * xhat,  xhat = (y - mean) * rstd
// per-CTA partials in fixed order, the last CTA to finish folds them (threadfence reduction) and also accumulates
// dgamma += bdot, dbeta += bsum into the fp32 gradient arena.
// ---------------------------------------------------------------------------------------------
struct BnBwdReduceParams {
  const __nv_bfloat16* g;
  const __nv_bfloat16* act;
  const __nv_bfloat16* y;
  const float* mean;
  const float* rstd;
  long long nvec;
  int C;
  double* partial;      // [2][C] global fp64 accumulators, zero on entry and left zero
  unsigned int* counter;
  float* bsum;          // [C]
  float* bdot;          // [C]
  float* dgamma;        // += (may be null)
  float* dbeta;         // +=
  __nv_bfloat16* dz_out;  // optional: store dz (stem variant: the routed + masked gradient, reused by the apply pass)
  // stem variant: g is the pooled gradient [B,OH,OW,C] gathered through the arg-max of the 3x3/2 max-pool, and the ReLU
  // mask is recomputed from y*scale+shift
  const uint8_t* argmax;
  const float* scale;
  const float* shift;
  int H, W, OH, OW;
  int OHp, OWp;  // stem variant: padded-flat geometry of the pooled gradient g
  PadGeom geom;  // regular variant: padded-flat geometry of g / act / y / dz_out ({1,1,1,1} = dense)
};

template <bool STEM>
__global__ void __launch_bounds__(EW_THREADS, 2) bn_bwd_reduce_kernel(const BnBwdReduceParams p) {
  __shared__ float s_red[2 * EW_THREADS][8];  // 16 KB: per-thread partial (sum | dot) vectors
  __shared__ bool s_last;
  pdl_entry();
  const int groups = p.C >> 3;
  const long long stride = (long long)gridDim.x * EW_THREADS;
  long long i = (long long)blockIdx.x * EW_THREADS + threadIdx.x;
  const int cg = (int)(i % groups) * 8;
  const Vec8 mean = loadf8(p.mean + cg), rstd = loadf8(p.rstd + cg);
  Vec8 sc, sh;
  if (STEM) { sc = loadf8(p.scale + cg); sh = loadf8(p.shift + cg); }
  float a_sum[8], a_dot[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) { a_sum[k] = 0.f; a_dot[k] = 0.f; }
  if (!STEM) {
    // four vectors per iteration, all loads issued before the first use (12 x 16 bytes in flight per thread)
    PadWalk walk;
    walk.init(i / groups, stride / groups, p.geom);
    for (; i < p.nvec; i += 4 * stride) {
      bool ok[4], in[4];
      uint4 yq[4], gq[4], aq[4];  // kept packed (48 registers) until used: two CTAs per SM stay resident
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        in[u] = i + u * stride < p.nvec;
        ok[u] = in[u] && walk.valid();
        walk.next();
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        if (ok[u]) {
          const long long o = (i + u * stride) * 8;
          yq[u] = *reinterpret_cast<const uint4*>(p.y + o);
          gq[u] = *reinterpret_cast<const uint4*>(p.g + o);
          if (p.act) aq[u] = *reinterpret_cast<const uint4*>(p.act + o);
        }
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const long long o = (i + u * stride) * 8;
        if (ok[u]) {
          const uint32_t yw[4] = {yq[u].x, yq[u].y, yq[u].z, yq[u].w};
          uint32_t gw[4] = {gq[u].x, gq[u].y, gq[u].z, gq[u].w};
          const uint32_t aw[4] = {aq[u].x, aq[u].y, aq[u].z, aq[u].w};
#pragma unroll
          for (int k = 0; k < 8; ++k) {
            const int wi = k >> 1;
            float gk = (k & 1) ? bf16hi(gw[wi]) : bf16lo(gw[wi]);
            if (p.act) {
              const float ak = (k & 1) ? bf16hi(aw[wi]) : bf16lo(aw[wi]);
              if (!(ak > 0.f)) {
                gk = 0.f;
                gw[wi] &= (k & 1) ? 0x0000FFFFu : 0xFFFF0000u;  // dz keeps the bf16 bits of g where the ReLU was active
              }
            }
            const float yk = (k & 1) ? bf16hi(yw[wi]) : bf16lo(yw[wi]);
            a_sum[k] += gk;
            a_dot[k] = fmaf(gk, (yk - mean.v[k]) * rstd.v[k], a_dot[k]);
          }
          if (p.dz_out) *reinterpret_cast<uint4*>(p.dz_out + o) = make_uint4(gw[0], gw[1], gw[2], gw[3]);
        } else if (in[u] && p.dz_out) {
          store8_zero(p.dz_out + o);  // padding pixel: g may hold stale data there; it contributes nothing and dz stays zero
        }
      }
    }
  } else {
    // stem: one thread per 2x2 block of conv1-output pixels (h0 = 2a, w0 = 2b) x 8 channels. The four 3x3/2 pool windows
    // that can select a pixel of the block are (a, b), (a, b+1), (a+1, b), (a+1, b+1): each is loaded once for the whole
    // block (the per-pixel form loaded up to four windows per pixel). H and W are even (44 x 100).
    const int HB = p.H >> 1, WB = p.W >> 1;
    const long long nblk = (long long)(p.nvec / groups / 4) * groups;  // = batch * HB * WB * groups
    for (long long bi = (long long)blockIdx.x * EW_THREADS + threadIdx.x; bi < nblk; bi += stride) {
      const long long blk = bi / groups;
      const int b = (int)(blk % WB);
      const int a = (int)((blk / WB) % HB);
      const int n = (int)(blk / ((long long)WB * HB));
      uint2 am[4];
      uint4 gq[4];
      bool okw[4];
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const int oh = a + (q >> 1), ow = b + (q & 1);
        okw[q] = oh < p.OH && ow < p.OW;
        if (okw[q]) {
          am[q] = *reinterpret_cast<const uint2*>(p.argmax + (((long long)n * p.OH + oh) * p.OW + ow) * p.C + cg);  // dense codes
          gq[q] = *reinterpret_cast<const uint4*>(p.g + (((long long)n * p.OHp + oh) * p.OWp + ow) * p.C + cg);
        }
      }
      Vec8 yv[4];
#pragma unroll
      for (int e = 0; e < 4; ++e)
        yv[e] = load8(p.y + ((((long long)n * p.H + 2 * a + (e >> 1)) * p.W + 2 * b + (e & 1)) * groups) * 8 + cg);
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int dh = e >> 1, dw = e & 1;  // pixel (2a + dh, 2b + dw)
        Vec8 gv;
#pragma unroll
        for (int k = 0; k < 8; ++k) gv.v[k] = 0.f;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          // window (a + qh, b + qw) covers rows 2(a+qh)-1 .. 2(a+qh)+1: the pixel's row inside it is r = dh + 1 - 2 qh
          const int r = dh + 1 - 2 * (q >> 1), sx = dw + 1 - 2 * (q & 1);
          if (r < 0 || sx < 0) continue;  // (compile-time after unrolling)
          if (okw[q]) {
            const uint32_t cc = (uint32_t)(r * 3 + sx) * 0x01010101u;
            const uint32_t m_lo = __vcmpeq4(am[q].x, cc), m_hi = __vcmpeq4(am[q].y, cc);
            const uint32_t gw[4] = {gq[q].x, gq[q].y, gq[q].z, gq[q].w};
#pragma unroll
            for (int k = 0; k < 8; ++k) {
              const uint32_t mk = ((k < 4 ? m_lo : m_hi) >> ((k & 3) * 8)) & 1u;
              const float gvk = (k & 1) ? bf16hi(gw[k >> 1]) : bf16lo(gw[k >> 1]);
              gv.v[k] = fmaf((float)mk, gvk, gv.v[k]);
            }
          }
        }
#pragma unroll
        for (int k = 0; k < 8; ++k)
          if (!(fmaf(yv[e].v[k], sc.v[k], sh.v[k]) > 0.f)) gv.v[k] = 0.f;
        if (p.dz_out) store8(p.dz_out + ((((long long)n * p.H + 2 * a + dh) * p.W + 2 * b + dw) * groups) * 8 + cg, gv);
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          a_sum[k] += gv.v[k];
          a_dot[k] = fmaf(gv.v[k], (yv[e].v[k] - mean.v[k]) * rstd.v[k], a_dot[k]);
        }
      }
    }
  }
  // threads with the same channel group are EW_THREADS/groups apart in steps of `groups`
#pragma unroll
  for (int k = 0; k < 8; ++k) { s_red[threadIdx.x][k] = a_sum[k]; s_red[EW_THREADS + threadIdx.x][k] = a_dot[k]; }
  __syncthreads();
  if (threadIdx.x < groups) {
    float t_sum[8], t_dot[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) { t_sum[k] = 0.f; t_dot[k] = 0.f; }
    for (int t = threadIdx.x; t < EW_THREADS; t += groups) {
#pragma unroll
      for (int k = 0; k < 8; ++k) { t_sum[k] += s_red[t][k]; t_dot[k] += s_red[EW_THREADS + t][k]; }
    }
    // global per-channel accumulators [2][C] (zero on entry, re-zeroed by the last CTA)
    // (fp64 adds of fp32 partials: exact, hence order-independent, unless their exponents span more than 2^22)
    double* pp = p.partial + threadIdx.x * 8;
#pragma unroll
    for (int k = 0; k < 8; ++k) { atomicAdd(pp + k, (double)t_sum[k]); atomicAdd(pp + p.C + k, (double)t_dot[k]); }
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    const unsigned int done = atomicAdd(p.counter, 1u);
    s_last = (done == gridDim.x - 1);
    if (s_last) __threadfence();
  }
  __syncthreads();
  if (s_last) {
    for (int c = threadIdx.x; c < p.C; c += EW_THREADS) {
      const float s = (float)__ldcg(p.partial + c), d = (float)__ldcg(p.partial + p.C + c);
      p.partial[c] = 0.0;
      p.partial[p.C + c] = 0.0;
      p.bsum[c] = s;
      p.bdot[c] = d;
      if (p.dgamma) p.dgamma[c] += d;
      if (p.dbeta) p.dbeta[c] += s;
    }
    if (threadIdx.x == 0) *p.counter = 0u;  // ready for the next launch / graph replay
  }
}

// ---------------------------------------------------------------------------------------------
// BatchNorm backward, pass 2 (apply):
//   training: dy = gamma*rstd * ( dz - bsum/n - xhat * bdot/n )
//   frozen  : dy = gamma*rstd * dz
// optionally also stores dz (the gradient that continues down the identity branch)
// ---------------------------------------------------------------------------------------------
struct BnBwdApplyParams {
  const __nv_bfloat16* g;
  const __nv_bfloat16* act;
  const __nv_bfloat16* y;
  const float* mean;
  const float* rstd;
  const float* gamma;
  const float* bsum;
  const float* bdot;
  float inv_count;
  int frozen;
  long long nvec;
  int C;
  __nv_bfloat16* dy;
  __nv_bfloat16* dz;  // may be null
  const uint8_t* argmax;
  const float* scale;
  const float* shift;
  int H, W, OH, OW;
  PadGeom geom;       // padded-flat geometry of g / act / y / dy / dz ({1,1,1,1} = dense)
  BnBwdDefer defer;   // acc_sum != nullptr: bsum / bdot come from the dgrad epilogue's sums (and CTA 0 accumulates dgamma / dbeta)
};

template <bool STEM>
__global__ void __launch_bounds__(EW_THREADS, 2) bn_bwd_apply_kernel(const BnBwdApplyParams p) {
  pdl_entry();
  const int groups = p.C >> 3;
  const long long stride = (long long)gridDim.x * EW_THREADS;
  long long i = (long long)blockIdx.x * EW_THREADS + threadIdx.x;
  const int cg = (int)(i % groups) * 8;
  // (the first batch of the main loop is requested before the deferred finalize, whose latency chain it overlaps)
  PadWalk walk;
  bool in[4], ok[4];
  uint4 yq[4], gq[4], aq[4];
#define CILRS_BN_BWD_LOAD()                                                       \
  {                                                                               \
    _Pragma("unroll") for (int u = 0; u < 4; ++u) {                               \
      in[u] = i + u * stride < p.nvec;                                            \
      ok[u] = in[u] && walk.valid();                                              \
      walk.next();                                                                \
    }                                                                             \
    _Pragma("unroll") for (int u = 0; u < 4; ++u) {                               \
      if (ok[u]) {                                                                \
        const long long o = (i + u * stride) * 8;                                 \
        yq[u] = *reinterpret_cast<const uint4*>(p.y + o);                         \
        gq[u] = *reinterpret_cast<const uint4*>(p.g + o);                         \
        if (p.act) aq[u] = *reinterpret_cast<const uint4*>(p.act + o);            \
      }                                                                           \
    }                                                                             \
  }
  if (!STEM) {
    walk.init(i / groups, stride / groups, p.geom);
    CILRS_BN_BWD_LOAD()
  }
  const Vec8 mean = loadf8(p.mean + cg), rstd = loadf8(p.rstd + cg), gamma = loadf8(p.gamma + cg);
  Vec8 k0, k1;
  __shared__ __align__(16) float s_par[2][EW_DEFER_MAX_C];
  if (p.defer.acc_sum) {
    // every CTA derives all channels once from the dgrad epilogue's sums (coalesced), threads pick theirs from shared memory
    for (int c = threadIdx.x; c < p.C; c += EW_THREADS) {
      const double s0 = p.defer.acc_sum[c], s1 = p.defer.acc_dot[c];
      const float bs = (float)s0, bd = bn_bdot_from_sums(s0, s1, p.mean[c], p.rstd[c]);
      s_par[0][c] = p.frozen ? 0.f : bs * p.inv_count;
      s_par[1][c] = p.frozen ? 0.f : bd * p.inv_count;
      if (blockIdx.x == 0) {
        p.defer.bred[c] = bs; p.defer.bred[p.C + c] = bd;
        if (p.defer.dgamma) p.defer.dgamma[c] += bd;
        if (p.defer.dbeta) p.defer.dbeta[c] += bs;
      }
    }
    __syncthreads();
    k0 = loadf8(&s_par[0][cg]); k1 = loadf8(&s_par[1][cg]);
  } else {
    const Vec8 bs = loadf8(p.bsum + cg), bd = loadf8(p.bdot + cg);
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      k0.v[k] = p.frozen ? 0.f : bs.v[k] * p.inv_count;
      k1.v[k] = p.frozen ? 0.f : bd.v[k] * p.inv_count;
    }
  }
  if (!STEM) {
    // four vectors per iteration, all loads issued before the first use (see bn_apply_kernel)
    for (;;) {
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const long long o = (i + u * stride) * 8;
        if (!ok[u]) {
          if (in[u]) {
            store8_zero(p.dy + o);
            if (p.dz) store8_zero(p.dz + o);
          }
          continue;
        }
        const Vec8 yv = unpack8(yq[u]);
        Vec8 gv = unpack8(gq[u]);
        if (p.act) {
          const Vec8 av = unpack8(aq[u]);
#pragma unroll
          for (int k = 0; k < 8; ++k)
            if (!(av.v[k] > 0.f)) gv.v[k] = 0.f;
        }
        if (p.dz) store8(p.dz + o, gv);
        Vec8 ov;
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          const float xhat = (yv.v[k] - mean.v[k]) * rstd.v[k];
          ov.v[k] = gamma.v[k] * rstd.v[k] * (gv.v[k] - k0.v[k] - xhat * k1.v[k]);
        }
        store8(p.dy + o, ov);
      }
      i += 4 * stride;
      if (i >= p.nvec) break;
      CILRS_BN_BWD_LOAD()
    }
  }
}

#undef CILRS_BN_BWD_LOAD

// out = a + b (fp32; the two parts of the feature gradient for callers that want it as one tensor)
__global__ void add2_kernel(const float* __restrict__ a, const float* __restrict__ b, float* __restrict__ out, long long n) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) out[i] = a[i] + b[i];
}

// grid size for the vector kernels: a multiple of the channel-group count keeps every thread on one channel group
// reductions: ~1 CTA per SM is enough with the 4-way unrolled loop, and keeps the partials the last CTA folds small
inline int ew_reduce_grid(long long nvec, int C) {
  (void)C;
  long long blocks = (nvec + (long long)EW_THREADS * 4 - 1) / ((long long)EW_THREADS * 4);
  if (blocks > 148 * 2) blocks = 148 * 2;
  if (blocks < 1) blocks = 1;
  return (int)blocks;
}

// (EW_THREADS = 256 is a multiple of every group count 8..64, so any block count keeps that property.)
// `per_thread` = vectors each thread should get at least: streaming kernels use 2, reductions 8 (fewer partials to fold).
inline int ew_grid(long long nvec, int C, int per_thread = 2, int ctas_per_sm = 3) {
  (void)C;
  long long blocks = (nvec + (long long)EW_THREADS * per_thread - 1) / ((long long)EW_THREADS * per_thread);
  // the 4-way unrolled streaming kernels: exactly the resident CTAs (bn_apply: 70 registers -> 3 per SM, bn_bwd_apply: 123 -> 2)
  if (per_thread == 4) { if (blocks > 148 * ctas_per_sm) blocks = 148 * ctas_per_sm; }
  else if (blocks > EW_MAX_BLOCKS) blocks = EW_MAX_BLOCKS;
  if (blocks < 1) blocks = 1;
  return (int)blocks;
}

}  // namespace cilrs
