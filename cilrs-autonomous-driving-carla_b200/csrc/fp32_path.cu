// fp32 compute mode of the CILRS hot path (BASELINE.json north_star: "fp32 mode: controls, speed prediction, losses and
// gradients match within 1e-4 relative"). The reference trains and infers in fp32 (configs/train_config.json:54
// "mixed_precision": false; model/autonomous_drive.py:495), so this is the mode that reproduces ITS numbers; the bf16 tcgen05
// plan (model.cu) is the throughput mode.
//
// Everything here is fp32 storage + fp32 FMA on the CUDA cores with fp64 reductions for the BatchNorm statistics - no tensor
// cores: a bf16x3 split on tcgen05 reaches ~1e-5 per contraction, which the 36 chained train-mode BatchNorm backwards amplify
// past the bar (the reference's own fp32 is 2e-3..8e-3 from fp64 on the trunk gradients), so the parity mode
// keeps true fp32 operands end to end. Same boundary as the bf16 plan: caller-owned arenas in cilrs_model_param_layout order,
// one C-ABI call per forward / backward, no allocation, no synchronisation; the heads are the same fp32 kernels (heads.cuh).
//
// Layout: activations NHWC fp32, dense. Weights are repacked at the start of every forward from the OIHW masters into
// [tap][Cin][Cout] (fprop) and [tap][Cout][Cin] (dgrad) so the GEMM B-tiles are contiguous.
//   conv32_kernel<DGRAD>  implicit GEMM, 64x64x16 tiles, 4x4 register micro-tiles (fprop: M = output pixels, K = taps x Cin;
//                         dgrad: M = input pixels, K = taps x Cout, stride-2 handled by the divisibility test of the gather)
//   wgrad32_kernel        [Cout x Cin] per tap, split-K over pixels into scratch + fixed-order reduce (deterministic)
//   bn32_*                batch statistics (fp64 sums), finalize (running statistics as torch.nn.BatchNorm2d), apply
//                         (+ residual [+ its own BN], ReLU), backward reduce / finalize / apply
//   maxpool32 / avgpool32 forward and backward
#include "common.cuh"
#include "heads_run.cuh"
#include "../../include/cilrs_b200.h"
#include <new>
#include <vector>

namespace cilrs {
namespace f32 {

#define CK(call)            \
  do {                      \
    int _st = (call);       \
    if (_st) return _st;    \
  } while (0)
#define CKL() CK(cuda_status(cudaGetLastError()))

static inline long long align_up(long long x, long long a) { return (x + a - 1) / a * a; }

// ------------------------------------------------------------------------------------------------
// kernels
// ------------------------------------------------------------------------------------------------
static __global__ void nchw_to_nhwc_kernel(const float* __restrict__ in, float* __restrict__ out, int B, int C, int H, int W) {
  const long long n = (long long)B * C * H * W;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % C);
    const long long p = i / C;
    const int w = (int)(p % W);
    const int h = (int)((p / W) % H);
    const int b = (int)(p / ((long long)W * H));
    out[i] = in[(((long long)b * C + c) * H + h) * W + w];
  }
}

// OIHW -> wf [tap][ci][co] and wd [tap][co][ci] (wd may be null)
static __global__ void pack_w32_kernel(const float* __restrict__ w, float* __restrict__ wf, float* __restrict__ wd, int Cout, int Cin, int taps) {
  const long long n = (long long)Cout * Cin * taps;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const int tap = (int)(i % taps);
    const int ci = (int)((i / taps) % Cin);
    const int co = (int)(i / ((long long)taps * Cin));
    const float v = w[i];
    wf[((long long)tap * Cin + ci) * Cout + co] = v;
    if (wd) wd[((long long)tap * Cout + co) * Cin + ci] = v;
  }
}

struct Conv32Params {
  const float* in;      // fprop: x [B,H,W,Cin]          dgrad: dy [B,OH,OW,Cout]
  const float* w;       // fprop: [tap][Cin][Cout]       dgrad: [tap][Cout][Cin]
  float* out;           // fprop: y [B,OH,OW,Cout]       dgrad: dx [B,H,W,Cin]
  const float* addend;  // optional, same shape as out (may alias out)
  int B, H, W, Cin, OH, OW, Cout, KH, KW, stride, pad;
  int M, N, Kdim, Cred; // GEMM dims; Cred = channels of the reduction (Cin for fprop, Cout for dgrad)
};

constexpr int C32_BM = 64, C32_BN = 64, C32_BK = 16;

template <bool DGRAD>
static __global__ void __launch_bounds__(256) conv32_kernel(const Conv32Params p) {
  __shared__ __align__(16) float As[C32_BK][C32_BM + 4];
  __shared__ __align__(16) float Bs[C32_BK][C32_BN];
  const int t = threadIdx.x;
  const int m0 = blockIdx.x * C32_BM, n0 = blockIdx.y * C32_BN;
  // ---- A-load role: k_local = t & 15, rows m_local = (t >> 4) + 16 j ----
  const int kl = t & 15;
  int rb[4], r0[4], r1[4];   // image index and the two base coordinates of this thread's four rows
  bool rok[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const int m = m0 + (t >> 4) + 16 * j;
    rok[j] = m < p.M;
    const int mm = rok[j] ? m : 0;
    if (!DGRAD) {
      const int ow = mm % p.OW, oh = (mm / p.OW) % p.OH;
      rb[j] = mm / (p.OW * p.OH);
      r0[j] = oh * p.stride - p.pad;
      r1[j] = ow * p.stride - p.pad;
    } else {
      const int w = mm % p.W, h = (mm / p.W) % p.H;
      rb[j] = mm / (p.W * p.H);
      r0[j] = h + p.pad;
      r1[j] = w + p.pad;
    }
  }
  // reduction index of this thread: k = k0 + kl -> (kh, kw, c), advanced by 16 per step
  int c = kl, kh = 0, kw = 0;
  while (c >= p.Cred) { c -= p.Cred; if (++kw == p.KW) { kw = 0; ++kh; } }
  // ---- B-load role: n_local = t & 63, k rows (t >> 6) + 4 j ----
  const int nl = t & 63, kb = t >> 6;
  // ---- compute role: 4 x 4 micro-tile ----
  const int tx = t & 15, ty = t >> 4;
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  for (int k0 = 0; k0 < p.Kdim; k0 += C32_BK) {
    float av[4], bv[4];
    const bool kok = (k0 + kl) < p.Kdim;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      float v = 0.f;
      if (rok[j] && kok) {
        if (!DGRAD) {
          const int ih = r0[j] + kh, iw = r1[j] + kw;
          if (ih >= 0 && ih < p.H && iw >= 0 && iw < p.W) v = __ldg(p.in + (((long long)rb[j] * p.H + ih) * p.W + iw) * p.Cred + c);
        } else {
          const int th = r0[j] - kh, tw = r1[j] - kw;
          if (th >= 0 && tw >= 0 && (th % p.stride) == 0 && (tw % p.stride) == 0) {
            const int oh = th / p.stride, ow = tw / p.stride;
            if (oh < p.OH && ow < p.OW) v = __ldg(p.in + (((long long)rb[j] * p.OH + oh) * p.OW + ow) * p.Cred + c);
          }
        }
      }
      av[j] = v;
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int k = k0 + kb + 4 * j;
      bv[j] = k < p.Kdim ? __ldg(p.w + (long long)k * p.N + n0 + nl) : 0.f;
    }
    __syncthreads();   // the previous step's reads of As / Bs are done
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      As[kl][(t >> 4) + 16 * j] = av[j];
      Bs[kb + 4 * j][nl] = bv[j];
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < C32_BK; ++kk) {
      const float4 a4 = *reinterpret_cast<const float4*>(&As[kk][ty * 4]);
      const float4 b4 = *reinterpret_cast<const float4*>(&Bs[kk][tx * 4]);
      const float a[4] = {a4.x, a4.y, a4.z, a4.w}, b[4] = {b4.x, b4.y, b4.z, b4.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    // advance (kh, kw, c) by 16 reduction elements
    c += C32_BK;
    while (c >= p.Cred) { c -= p.Cred; if (++kw == p.KW) { kw = 0; ++kh; } }
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int m = m0 + ty * 4 + i;
    if (m >= p.M) continue;
    float* o = p.out + (long long)m * p.N + n0 + tx * 4;
    float4 v = make_float4(acc[i][0], acc[i][1], acc[i][2], acc[i][3]);
    if (p.addend) {
      const float4 r = *reinterpret_cast<const float4*>(p.addend + (long long)m * p.N + n0 + tx * 4);
      v.x += r.x; v.y += r.y; v.z += r.z; v.w += r.w;
    }
    *reinterpret_cast<float4*>(o) = v;
  }
}

struct Wgrad32Params {
  const float* dy;   // [B,OH,OW,Cout]
  const float* x;    // [B,H,W,Cin]
  float* scratch;    // [kslices][taps][Cout][Cin]
  int B, H, W, Cin, OH, OW, Cout, KH, KW, stride, pad;
  int P;             // B*OH*OW
  int kslices, chunk;  // pixels per slice
};

// grid (Cout/64, ceil(Cin/64), taps * kslices)
static __global__ void __launch_bounds__(256) wgrad32_kernel(const Wgrad32Params p) {
  __shared__ __align__(16) float As[C32_BK][C32_BM + 4];   // [pixel][co]
  __shared__ __align__(16) float Bs[C32_BK][C32_BN];       // [pixel][ci]
  const int t = threadIdx.x;
  const int co0 = blockIdx.x * 64, ci0 = blockIdx.y * 64;
  const int taps = p.KH * p.KW;
  const int tap = blockIdx.z % taps, ks = blockIdx.z / taps;
  const int kh = tap / p.KW, kw = tap % p.KW;
  const int p_begin = ks * p.chunk, p_end = min(p.P, p_begin + p.chunk);
  const int cl = t & 63, pl = t >> 6;   // channel, pixel rows pl + 4 j
  const int tx = t & 15, ty = t >> 4;
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
  for (int q0 = p_begin; q0 < p_end; q0 += C32_BK) {
    float av[4], bv[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int q = q0 + pl + 4 * j;
      float a = 0.f, b = 0.f;
      if (q < p_end) {
        a = __ldg(p.dy + (long long)q * p.Cout + co0 + cl);
        if (ci0 + cl < p.Cin) {
          const int ow = q % p.OW, oh = (q / p.OW) % p.OH, n = q / (p.OW * p.OH);
          const int ih = oh * p.stride + kh - p.pad, iw = ow * p.stride + kw - p.pad;
          if (ih >= 0 && ih < p.H && iw >= 0 && iw < p.W) b = __ldg(p.x + (((long long)n * p.H + ih) * p.W + iw) * p.Cin + ci0 + cl);
        }
      }
      av[j] = a; bv[j] = b;
    }
    __syncthreads();
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      As[pl + 4 * j][cl] = av[j];
      Bs[pl + 4 * j][cl] = bv[j];
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < C32_BK; ++kk) {
      const float4 a4 = *reinterpret_cast<const float4*>(&As[kk][ty * 4]);
      const float4 b4 = *reinterpret_cast<const float4*>(&Bs[kk][tx * 4]);
      const float a[4] = {a4.x, a4.y, a4.z, a4.w}, b[4] = {b4.x, b4.y, b4.z, b4.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
  }
  float* dst = p.scratch + ((long long)ks * taps + tap) * p.Cout * p.Cin;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int co = co0 + ty * 4 + i;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int ci = ci0 + tx * 4 + j;
      if (ci < p.Cin) dst[(long long)co * p.Cin + ci] = acc[i][j];
    }
  }
}

// dw[co][ci][tap] += sum over slices (fixed order) of scratch[slice][tap][co][ci]
static __global__ void wgrad32_reduce_kernel(const float* __restrict__ scratch, float* __restrict__ dw, int Cout, int Cin, int taps, int kslices) {
  const long long n = (long long)Cout * Cin * taps;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const int tap = (int)(i % taps);
    const long long cc = i / taps;   // co * Cin + ci
    float s = 0.f;
    for (int k = 0; k < kslices; ++k) s += scratch[((long long)k * taps + tap) * Cout * Cin + cc];
    dw[i] += s;
  }
}

// per-channel sum and sum of squares over P rows of x [P][C] (fp64), added to acc [2][C] (zeroed by the caller)
static __global__ void __launch_bounds__(256) bn32_stats_kernel(const float* __restrict__ x, long long P, int C, double* __restrict__ acc) {
  __shared__ double red[256][8];
  const int C4 = C >> 2;
  const int rows = 256 / C4;              // rows handled per block iteration (C <= 1024)
  const int c4 = threadIdx.x % C4, r = threadIdx.x / C4;
  double s[4] = {0, 0, 0, 0}, q[4] = {0, 0, 0, 0};
  if (r < rows) {
    for (long long row = (long long)blockIdx.x * rows + r; row < P; row += (long long)gridDim.x * rows) {
      const float4 v = *reinterpret_cast<const float4*>(x + row * C + 4 * c4);
      s[0] += v.x; s[1] += v.y; s[2] += v.z; s[3] += v.w;
      q[0] += (double)v.x * v.x; q[1] += (double)v.y * v.y; q[2] += (double)v.z * v.z; q[3] += (double)v.w * v.w;
    }
  }
#pragma unroll
  for (int k = 0; k < 4; ++k) { red[threadIdx.x][k] = s[k]; red[threadIdx.x][4 + k] = q[k]; }
  __syncthreads();
  if (r == 0) {
    for (int rr = 1; rr < rows; ++rr)
#pragma unroll
      for (int k = 0; k < 8; ++k) red[threadIdx.x][k] += red[rr * C4 + c4][k];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      atomicAdd(acc + 4 * c4 + k, red[threadIdx.x][k]);
      atomicAdd(acc + C + 4 * c4 + k, red[threadIdx.x][4 + k]);
    }
  }
}

// vec [4][C] = scale, shift, mean, rstd; train: from the batch sums (and running statistics updated as torch.nn.BatchNorm2d does),
// else from the running statistics
static __global__ void bn32_finalize_kernel(const double* __restrict__ acc, int C, double count, const float* __restrict__ gamma,
                                            const float* __restrict__ beta, float* running_mean, float* running_var, long long* nbt,
                                            float momentum, float eps, int training, int update_running, float* __restrict__ vec) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  float mean, var;
  if (training) {
    const double m = acc[c] / count;
    double v = acc[C + c] / count - m * m;
    if (v < 0.0) v = 0.0;
    mean = (float)m; var = (float)v;
    if (update_running) {
      const float unbiased = count > 1.0 ? (float)(v * (count / (count - 1.0))) : var;
      running_mean[c] = (1.f - momentum) * running_mean[c] + momentum * mean;
      running_var[c] = (1.f - momentum) * running_var[c] + momentum * unbiased;
      if (c == 0 && nbt) *nbt += 1;
    }
  } else {
    mean = running_mean[c]; var = running_var[c];
  }
  const float rstd = (float)(1.0 / sqrt((double)var + (double)eps));
  const float sc = gamma[c] * rstd;
  vec[c] = sc; vec[C + c] = beta[c] - mean * sc; vec[2 * C + c] = mean; vec[3 * C + c] = rstd;
}

// out = relu?( x*scale + shift [+ res (*rscale + rshift)] )
static __global__ void bn32_apply_kernel(const float* __restrict__ x, const float* __restrict__ vec, const float* __restrict__ res,
                                         const float* __restrict__ rvec, float* __restrict__ out, long long n4, int C, int relu) {
  const int C4 = C >> 2;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % C4) * 4;
    const float4 v = reinterpret_cast<const float4*>(x)[i];
    const float4 sc = *reinterpret_cast<const float4*>(vec + c), sh = *reinterpret_cast<const float4*>(vec + C + c);
    float4 o = make_float4(fmaf(v.x, sc.x, sh.x), fmaf(v.y, sc.y, sh.y), fmaf(v.z, sc.z, sh.z), fmaf(v.w, sc.w, sh.w));
    if (res) {
      float4 r = reinterpret_cast<const float4*>(res)[i];
      if (rvec) {
        const float4 rs = *reinterpret_cast<const float4*>(rvec + c), rh = *reinterpret_cast<const float4*>(rvec + C + c);
        r = make_float4(fmaf(r.x, rs.x, rh.x), fmaf(r.y, rs.y, rh.y), fmaf(r.z, rs.z, rh.z), fmaf(r.w, rs.w, rh.w));
      }
      o.x += r.x; o.y += r.y; o.z += r.z; o.w += r.w;
    }
    if (relu) { o.x = fmaxf(o.x, 0.f); o.y = fmaxf(o.y, 0.f); o.z = fmaxf(o.z, 0.f); o.w = fmaxf(o.w, 0.f); }
    reinterpret_cast<float4*>(out)[i] = o;
  }
}

// dz = g * (act > 0) (act may be null), written to dz_out (may alias g; may be null); acc[0][C] += sum dz, acc[1][C] += sum dz * xhat
static __global__ void __launch_bounds__(256) bn32_bwd_reduce_kernel(const float* __restrict__ g, const float* __restrict__ act,
                                                                     const float* __restrict__ y, const float* __restrict__ vec, long long P,
                                                                     int C, float* dz_out, double* __restrict__ acc) {
  __shared__ double red[256][8];
  const int C4 = C >> 2;
  const int rows = 256 / C4;
  const int c4 = threadIdx.x % C4, r = threadIdx.x / C4;
  double s[4] = {0, 0, 0, 0}, q[4] = {0, 0, 0, 0};
  if (r < rows) {
    const float4 mean = *reinterpret_cast<const float4*>(vec + 2 * C + 4 * c4), rstd = *reinterpret_cast<const float4*>(vec + 3 * C + 4 * c4);
    for (long long row = (long long)blockIdx.x * rows + r; row < P; row += (long long)gridDim.x * rows) {
      const long long o = row * C + 4 * c4;
      float4 gv = *reinterpret_cast<const float4*>(g + o);
      if (act) {
        const float4 a = *reinterpret_cast<const float4*>(act + o);
        if (!(a.x > 0.f)) gv.x = 0.f;
        if (!(a.y > 0.f)) gv.y = 0.f;
        if (!(a.z > 0.f)) gv.z = 0.f;
        if (!(a.w > 0.f)) gv.w = 0.f;
      }
      if (dz_out) *reinterpret_cast<float4*>(dz_out + o) = gv;
      const float4 yv = *reinterpret_cast<const float4*>(y + o);
      s[0] += gv.x; s[1] += gv.y; s[2] += gv.z; s[3] += gv.w;
      q[0] += (double)gv.x * (double)((yv.x - mean.x) * rstd.x); q[1] += (double)gv.y * (double)((yv.y - mean.y) * rstd.y);
      q[2] += (double)gv.z * (double)((yv.z - mean.z) * rstd.z); q[3] += (double)gv.w * (double)((yv.w - mean.w) * rstd.w);
    }
  }
#pragma unroll
  for (int k = 0; k < 4; ++k) { red[threadIdx.x][k] = s[k]; red[threadIdx.x][4 + k] = q[k]; }
  __syncthreads();
  if (r == 0) {
    for (int rr = 1; rr < rows; ++rr)
#pragma unroll
      for (int k = 0; k < 8; ++k) red[threadIdx.x][k] += red[rr * C4 + c4][k];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      atomicAdd(acc + 4 * c4 + k, red[threadIdx.x][k]);
      atomicAdd(acc + C + 4 * c4 + k, red[threadIdx.x][4 + k]);
    }
  }
}

// bred [2][C] = (sum dz, sum dz*xhat) as fp32; dgamma += sum dz*xhat, dbeta += sum dz
static __global__ void bn32_bwd_finalize_kernel(const double* __restrict__ acc, int C, float* __restrict__ bred, float* dgamma, float* dbeta) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  const float s = (float)acc[c], d = (float)acc[C + c];
  bred[c] = s; bred[C + c] = d;
  dgamma[c] += d;
  dbeta[c] += s;
}

// dy = gamma*rstd * (dz - bsum/n - xhat*bdot/n)   (frozen: gamma*rstd*dz)
static __global__ void bn32_bwd_apply_kernel(const float* __restrict__ dz, const float* __restrict__ y, const float* __restrict__ vec,
                                             const float* __restrict__ gamma, const float* __restrict__ bred, float inv_count, int frozen,
                                             float* __restrict__ dy, long long n4, int C) {
  const int C4 = C >> 2;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % C4) * 4;
    const float4 d = reinterpret_cast<const float4*>(dz)[i];
    const float4 yv = reinterpret_cast<const float4*>(y)[i];
    const float dd[4] = {d.x, d.y, d.z, d.w}, yy[4] = {yv.x, yv.y, yv.z, yv.w};
    float o[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const float rstd = vec[3 * C + c + k], mean = vec[2 * C + c + k];
      const float xhat = (yy[k] - mean) * rstd;
      const float k0 = frozen ? 0.f : bred[c + k] * inv_count, k1 = frozen ? 0.f : bred[C + c + k] * inv_count;
      o[k] = gamma[c + k] * rstd * (dd[k] - k0 - xhat * k1);
    }
    reinterpret_cast<float4*>(dy)[i] = make_float4(o[0], o[1], o[2], o[3]);
  }
}

// 3x3 / stride 2 / pad 1 max-pool, NHWC; arg = r*3+s of the (first) maximum
static __global__ void maxpool32_fwd_kernel(const float* __restrict__ x, float* __restrict__ out, uint8_t* __restrict__ arg, int B, int H, int W,
                                            int C, int OH, int OW) {
  const long long n = (long long)B * OH * OW * C;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % C);
    const long long p = i / C;
    const int ow = (int)(p % OW), oh = (int)((p / OW) % OH), b = (int)(p / ((long long)OW * OH));
    float best = -INFINITY;
    int code = 0;
    for (int r = 0; r < 3; ++r)
      for (int s = 0; s < 3; ++s) {
        const int h = oh * 2 - 1 + r, w = ow * 2 - 1 + s;
        if (h < 0 || h >= H || w < 0 || w >= W) continue;
        const float v = x[(((long long)b * H + h) * W + w) * C + c];
        if (v > best) { best = v; code = r * 3 + s; }
      }
    out[i] = best;
    arg[i] = (uint8_t)code;
  }
}

// dx[b,h,w,c] = sum of g[b,oh,ow,c] over the windows whose maximum is (h,w)
static __global__ void maxpool32_bwd_kernel(const float* __restrict__ g, const uint8_t* __restrict__ arg, float* __restrict__ dx, int B, int H,
                                            int W, int C, int OH, int OW) {
  const long long n = (long long)B * H * W * C;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % C);
    const long long p = i / C;
    const int w = (int)(p % W), h = (int)((p / W) % H), b = (int)(p / ((long long)W * H));
    float s = 0.f;
    // windows (oh, ow) with oh*2-1 <= h <= oh*2+1  <=>  oh in [ceil((h-1)/2), floor((h+1)/2)]
    const int oh_lo = h >= 1 ? (h - 1 + 1) / 2 : 0, oh_hi = min(OH - 1, (h + 1) / 2);
    const int ow_lo = w >= 1 ? (w - 1 + 1) / 2 : 0, ow_hi = min(OW - 1, (w + 1) / 2);
    for (int oh = oh_lo; oh <= oh_hi; ++oh)
      for (int ow = ow_lo; ow <= ow_hi; ++ow) {
        const int r = h - (oh * 2 - 1), sx = w - (ow * 2 - 1);
        if (r < 0 || r > 2 || sx < 0 || sx > 2) continue;
        const long long o = (((long long)b * OH + oh) * OW + ow) * C + c;
        if ((int)arg[o] == r * 3 + sx) s += g[o];
      }
    dx[i] = s;
  }
}

static __global__ void avgpool32_fwd_kernel(const float* __restrict__ x, float* __restrict__ feat, int B, int HW, int C) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= B * C) return;
  const int c = i % C, b = i / C;
  float s = 0.f;
  for (int q = 0; q < HW; ++q) s += x[((long long)b * HW + q) * C + c];
  feat[i] = s / (float)HW;
}

static __global__ void avgpool32_bwd_kernel(const float* __restrict__ d1, const float* __restrict__ d2, float* __restrict__ g, int B, int HW, int C) {
  const long long n = (long long)B * HW * C;
  const float inv = 1.f / (float)HW;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % C);
    const int b = (int)(i / ((long long)HW * C));
    g[i] = (d1[b * C + c] + d2[b * C + c]) * inv;
  }
}

// ------------------------------------------------------------------------------------------------
// plan
// ------------------------------------------------------------------------------------------------
struct Bn32 {
  int C;
  int gamma, beta;         // parameter slots
  long long rm_off, rv_off;
  int nbt_idx;
  float* vec;              // [4][C]
  float* bred;             // [2][C]
  double* acc;             // [2][C] forward sums
  double* bacc;            // [2][C] backward sums
};

struct Conv32 {
  int H, W, Cin, Cout, K, stride, pad, OH, OW;
  int w;                   // parameter slot
  float *wf, *wd;
  float* y;                // raw output [B,OH,OW,Cout]
  Bn32 bn;
};

struct Block32 {
  Conv32 a, b, ds;
  bool has_ds;
  const float* in;
  float* act_a;
  float* out;
};

struct Model32 {
  int maxB = 0;
  std::vector<long long> off;   // parameter tensor offsets (cilrs_model_param_layout)
  int nslots = 0;
  long long buffer_floats = 0;
  int num_bn = 0;
  Conv32 stem;
  std::vector<Block32> blocks;
  int head_slot0 = 0;
  float* params = nullptr; float* grads = nullptr; float* buffers = nullptr; long long* nbt = nullptr;
  // workspace
  float* x0; float* act0; float* pool_out; uint8_t* pool_arg;
  double* acc_fwd; double* acc_bwd; long long acc_fwd_bytes = 0, acc_bwd_bytes = 0;
  float* wscratch; long long wscratch_floats = 0;
  float* g[4];
  float *feat, *dfeat, *dfeat2, *head_comb;
  HeadsSaved hs;
  int* err_flag; unsigned int* counters;
  const long long* drop_counter = nullptr;
  int lastB = 0, lastMode = -1;
};

static Bn32 make_bn(Model32& m, int C, int& slot) {
  Bn32 b{};
  b.C = C; b.gamma = slot++; b.beta = slot++;
  b.rm_off = m.buffer_floats; b.rv_off = m.buffer_floats + C;
  m.buffer_floats += 2 * C;
  b.nbt_idx = m.num_bn++;
  return b;
}
static Conv32 make_conv(Model32& m, int h, int w, int cin, int cout, int k, int stride, int pad, int& slot) {
  Conv32 c{};
  c.H = h; c.W = w; c.Cin = cin; c.Cout = cout; c.K = k; c.stride = stride; c.pad = pad;
  c.OH = (h + 2 * pad - k) / stride + 1; c.OW = (w + 2 * pad - k) / stride + 1;
  c.w = slot++;
  c.bn = make_bn(m, cout, slot);
  return c;
}

// same order as model.cu build_topology == the reference's named_parameters()
static void build_topology(Model32& m) {
  m.buffer_floats = 0; m.num_bn = 0; m.blocks.clear();
  int slot = 0;
  m.stem = make_conv(m, 88, 200, 3, 64, 7, 2, 3, slot);
  const int chans[4] = {64, 128, 256, 512}, nblk[4] = {3, 4, 6, 3};
  int h = 22, w = 50, cin = 64;
  for (int s = 0; s < 4; ++s)
    for (int b = 0; b < nblk[s]; ++b) {
      Block32 blk{};
      const int stride = (b == 0 && s > 0) ? 2 : 1, c = chans[s];
      blk.a = make_conv(m, h, w, cin, c, 3, stride, 1, slot);
      blk.b = make_conv(m, blk.a.OH, blk.a.OW, c, c, 3, 1, 1, slot);
      blk.has_ds = (stride != 1 || cin != c);
      if (blk.has_ds) blk.ds = make_conv(m, h, w, cin, c, 1, stride, 0, slot);
      h = blk.a.OH; w = blk.a.OW; cin = c;
      m.blocks.push_back(blk);
    }
  m.head_slot0 = slot;
}

struct Bump {
  char* base; long long off;
  void* take(long long bytes) { off = align_up(off, 256); void* p = base ? (void*)(base + off) : nullptr; off += bytes; return p; }
};

static void carve_conv(Bump& bp, Conv32& c, int B) {
  const long long wn = (long long)c.K * c.K * c.Cin * c.Cout;
  c.wf = (float*)bp.take(wn * 4);
  c.wd = (float*)bp.take(wn * 4);
  c.y = (float*)bp.take((long long)B * c.OH * c.OW * c.Cout * 4);
  c.bn.vec = (float*)bp.take(4LL * c.bn.C * 4);
  c.bn.bred = (float*)bp.take(2LL * c.bn.C * 4);
}

static long long carve(Model32& m, char* base) {
  Bump bp{base, 0};
  const int B = m.maxB;
  m.x0 = (float*)bp.take((long long)B * 88 * 200 * 3 * 4);
  carve_conv(bp, m.stem, B);
  const long long stem_out = (long long)B * 44 * 100 * 64;
  m.act0 = (float*)bp.take(stem_out * 4);
  m.pool_out = (float*)bp.take((long long)B * 22 * 50 * 64 * 4);
  m.pool_arg = (uint8_t*)bp.take((long long)B * 22 * 50 * 64);
  const float* prev = m.pool_out;
  long long ch = m.stem.bn.C;
  for (auto& blk : m.blocks) {
    blk.in = prev;
    carve_conv(bp, blk.a, B);
    carve_conv(bp, blk.b, B);
    if (blk.has_ds) carve_conv(bp, blk.ds, B);
    const long long n = (long long)B * blk.b.OH * blk.b.OW * blk.b.Cout;
    blk.act_a = (float*)bp.take(n * 4);
    blk.out = (float*)bp.take(n * 4);
    prev = blk.out;
    ch += blk.a.bn.C + blk.b.bn.C + (blk.has_ds ? blk.ds.bn.C : 0);
  }
  m.acc_fwd_bytes = ch * 2 * 8; m.acc_bwd_bytes = ch * 2 * 8;
  m.acc_fwd = (double*)bp.take(m.acc_fwd_bytes);
  m.acc_bwd = (double*)bp.take(m.acc_bwd_bytes);
  {
    long long o = 0;
    auto give = [&](Bn32& bn) { bn.acc = m.acc_fwd ? m.acc_fwd + 2 * o : nullptr; bn.bacc = m.acc_bwd ? m.acc_bwd + 2 * o : nullptr; o += bn.C; };
    give(m.stem.bn);
    for (auto& blk : m.blocks) { give(blk.a.bn); give(blk.b.bn); if (blk.has_ds) give(blk.ds.bn); }
  }
  m.wscratch_floats = 16LL * 1024 * 1024;   // 64 MB of split-K partials (layer4: 2 slices x 9 x 512 x 512)
  m.wscratch = (float*)bp.take(m.wscratch_floats * 4);
  for (int i = 0; i < 4; ++i) m.g[i] = (float*)bp.take(stem_out * 4);
  m.feat = (float*)bp.take((long long)B * 512 * 4);
  m.dfeat = (float*)bp.take((long long)B * 512 * 4);
  m.dfeat2 = (float*)bp.take((long long)B * 512 * 4);
  m.head_comb = (float*)bp.take((long long)B * 640 * 4);
  float** hp[14] = {&m.hs.s1, &m.hs.sfeat, &m.hs.b1, &m.hs.b2, &m.hs.p1, &m.hs.p2, &m.hs.d_se0, &m.hs.d_se3,
                    &m.hs.d_br0, &m.hs.d_br3, &m.hs.d_br6, &m.hs.d_sp0, &m.hs.d_sp3, &m.hs.d_sp5};
  const int hw[14] = {128, 128, 256, 256, 256, 256, 128, 128, 256, 256, 4, 256, 256, 1};
  for (int i = 0; i < 14; ++i) *hp[i] = (float*)bp.take((long long)B * hw[i] * 4);
  m.err_flag = (int*)bp.take(64);
  m.counters = (unsigned int*)bp.take(64);
  return align_up(bp.off, 1024);
}

static inline int grid_for(long long n, int per_block = 256) {
  long long b = (n + per_block - 1) / per_block;
  if (b > 148 * 16) b = 148 * 16;
  if (b < 1) b = 1;
  return (int)b;
}

static int run_conv(const Conv32& c, int B, const float* x, float* y, cudaStream_t s) {
  Conv32Params p{};
  p.in = x; p.w = c.wf; p.out = y; p.addend = nullptr;
  p.B = B; p.H = c.H; p.W = c.W; p.Cin = c.Cin; p.OH = c.OH; p.OW = c.OW; p.Cout = c.Cout; p.KH = c.K; p.KW = c.K; p.stride = c.stride; p.pad = c.pad;
  p.M = B * c.OH * c.OW; p.N = c.Cout; p.Kdim = c.K * c.K * c.Cin; p.Cred = c.Cin;
  conv32_kernel<false><<<dim3((p.M + 63) / 64, p.N / 64), 256, 0, s>>>(p); ++g_cilrs_launches;
  return cuda_status(cudaGetLastError());
}
static int run_dgrad(const Conv32& c, int B, const float* dy, float* dx, const float* addend, cudaStream_t s) {
  Conv32Params p{};
  p.in = dy; p.w = c.wd; p.out = dx; p.addend = addend;
  p.B = B; p.H = c.H; p.W = c.W; p.Cin = c.Cin; p.OH = c.OH; p.OW = c.OW; p.Cout = c.Cout; p.KH = c.K; p.KW = c.K; p.stride = c.stride; p.pad = c.pad;
  p.M = B * c.H * c.W; p.N = c.Cin; p.Kdim = c.K * c.K * c.Cout; p.Cred = c.Cout;
  conv32_kernel<true><<<dim3((p.M + 63) / 64, p.N / 64), 256, 0, s>>>(p); ++g_cilrs_launches;
  return cuda_status(cudaGetLastError());
}
static int run_wgrad(Model32& m, const Conv32& c, int B, const float* dy, const float* x, cudaStream_t s) {
  Wgrad32Params p{};
  p.dy = dy; p.x = x; p.scratch = m.wscratch;
  p.B = B; p.H = c.H; p.W = c.W; p.Cin = c.Cin; p.OH = c.OH; p.OW = c.OW; p.Cout = c.Cout; p.KH = c.K; p.KW = c.K; p.stride = c.stride; p.pad = c.pad;
  p.P = B * c.OH * c.OW;
  const int taps = c.K * c.K;
  const int tiles = (c.Cout / 64) * ((c.Cin + 63) / 64) * taps;
  int ks = (592 + tiles - 1) / tiles;
  const long long per_slice = (long long)taps * c.Cout * c.Cin;
  if ((long long)ks * per_slice > m.wscratch_floats) ks = (int)(m.wscratch_floats / per_slice);
  if (ks > (p.P + 255) / 256) ks = (p.P + 255) / 256;
  if (ks < 1) ks = 1;
  p.kslices = ks;
  p.chunk = ((p.P + ks - 1) / ks + 15) / 16 * 16;
  wgrad32_kernel<<<dim3(c.Cout / 64, (c.Cin + 63) / 64, taps * ks), 256, 0, s>>>(p); ++g_cilrs_launches;
  CKL();
  wgrad32_reduce_kernel<<<grid_for(per_slice), 256, 0, s>>>(m.wscratch, m.grads + m.off[c.w], c.Cout, c.Cin, taps, ks); ++g_cilrs_launches;
  return cuda_status(cudaGetLastError());
}

// raw conv output -> vec (train: batch statistics) ; mode: 0 train, else running statistics
static int run_bn_vec(Model32& m, const Conv32& c, int B, int training, int update_running, cudaStream_t s) {
  const Bn32& bn = c.bn;
  const long long P = (long long)B * c.OH * c.OW;
  if (training) {
    bn32_stats_kernel<<<(int)((P + 15) / 16 < 296 ? (P + 15) / 16 : 296), 256, 0, s>>>(c.y, P, bn.C, bn.acc); ++g_cilrs_launches;
    CKL();
  }
  bn32_finalize_kernel<<<(bn.C + 255) / 256, 256, 0, s>>>(bn.acc, bn.C, (double)P, m.params + m.off[bn.gamma], m.params + m.off[bn.beta],
                                                          m.buffers + bn.rm_off, m.buffers + bn.rv_off, m.nbt ? m.nbt + bn.nbt_idx : nullptr,
                                                          0.1f, 1e-5f, training, update_running, bn.vec); ++g_cilrs_launches;
  return cuda_status(cudaGetLastError());
}
static int run_bn_apply(const Conv32& c, int B, const float* res, const float* rvec, float* out, int relu, cudaStream_t s) {
  const long long n4 = (long long)B * c.OH * c.OW * c.Cout / 4;
  bn32_apply_kernel<<<grid_for(n4), 256, 0, s>>>(c.y, c.bn.vec, res, rvec, out, n4, c.Cout, relu); ++g_cilrs_launches;
  return cuda_status(cudaGetLastError());
}
// dz (g masked by act > 0, written to dz_out) -> dy through the BatchNorm of conv c
static int run_bn_bwd(Model32& m, const Conv32& c, int B, const float* g, const float* act, float* dz_out, float* dy, int frozen, cudaStream_t s) {
  const Bn32& bn = c.bn;
  const long long P = (long long)B * c.OH * c.OW;
  bn32_bwd_reduce_kernel<<<(int)((P + 15) / 16 < 296 ? (P + 15) / 16 : 296), 256, 0, s>>>(g, act, c.y, bn.vec, P, bn.C, dz_out, bn.bacc); ++g_cilrs_launches;
  CKL();
  bn32_bwd_finalize_kernel<<<(bn.C + 255) / 256, 256, 0, s>>>(bn.bacc, bn.C, bn.bred, m.grads + m.off[bn.gamma], m.grads + m.off[bn.beta]); ++g_cilrs_launches;
  CKL();
  const long long n4 = P * bn.C / 4;
  bn32_bwd_apply_kernel<<<grid_for(n4), 256, 0, s>>>(dz_out ? dz_out : g, c.y, bn.vec, m.params + m.off[bn.gamma], bn.bred, (float)(1.0 / (double)P),
                                                      frozen, dy, n4, bn.C); ++g_cilrs_launches;
  return cuda_status(cudaGetLastError());
}

static HeadsCtx heads_ctx(const Model32& m) {
  HeadsCtx c;
  c.params = m.params; c.grads = m.grads;
  for (int i = 0; i < HD_NUM_SLOTS; ++i) c.off[i] = m.off[m.head_slot0 + i];
  c.feat = m.feat; c.dfeat = m.dfeat; c.dfeat2 = m.dfeat2; c.head_comb = m.head_comb; c.hs = m.hs; c.err_flag = m.err_flag;
  c.drop_counter = m.drop_counter; c.loss_counter = m.counters + 8;
  return c;
}

static int pack_conv(Model32& m, const Conv32& c, cudaStream_t s) {
  const long long n = (long long)c.K * c.K * c.Cin * c.Cout;
  pack_w32_kernel<<<grid_for(n), 256, 0, s>>>(m.params + m.off[c.w], c.wf, c.wd, c.Cout, c.Cin, c.K * c.K); ++g_cilrs_launches;
  return cuda_status(cudaGetLastError());
}

enum { MODE_TRAIN = 0, MODE_FROZEN = 1, MODE_INFER = 2 };

static int forward(Model32& m, int B, int mode, const float* image, const float* speed, const long long* command, float* controls,
                   float* pred_speed, int update_running, int keep, float dropout_p, unsigned long long seed, cudaStream_t s) {
  if (B < 1 || B > m.maxB || !m.params || !m.buffers || !image) return ERR_INVALID;
  const int training = mode == MODE_TRAIN;
  // operands from the fp32 masters (cheap next to the fp32 convolutions; always current)
  CK(pack_conv(m, m.stem, s));
  for (auto& blk : m.blocks) { CK(pack_conv(m, blk.a, s)); CK(pack_conv(m, blk.b, s)); if (blk.has_ds) CK(pack_conv(m, blk.ds, s)); }
  if (training) CK(cuda_status(cudaMemsetAsync(m.acc_fwd, 0, (size_t)m.acc_fwd_bytes, s)));
  nchw_to_nhwc_kernel<<<grid_for((long long)B * 3 * 88 * 200), 256, 0, s>>>(image, m.x0, B, 3, 88, 200); ++g_cilrs_launches;
  CKL();
  CK(run_conv(m.stem, B, m.x0, m.stem.y, s));
  CK(run_bn_vec(m, m.stem, B, training, update_running, s));
  CK(run_bn_apply(m.stem, B, nullptr, nullptr, m.act0, 1, s));
  maxpool32_fwd_kernel<<<grid_for((long long)B * 22 * 50 * 64), 256, 0, s>>>(m.act0, m.pool_out, m.pool_arg, B, 44, 100, 64, 22, 50); ++g_cilrs_launches;
  CKL();
  for (auto& blk : m.blocks) {
    CK(run_conv(blk.a, B, blk.in, blk.a.y, s));
    CK(run_bn_vec(m, blk.a, B, training, update_running, s));
    CK(run_bn_apply(blk.a, B, nullptr, nullptr, blk.act_a, 1, s));
    if (blk.has_ds) {
      CK(run_conv(blk.ds, B, blk.in, blk.ds.y, s));
      CK(run_bn_vec(m, blk.ds, B, training, update_running, s));
    }
    CK(run_conv(blk.b, B, blk.act_a, blk.b.y, s));
    CK(run_bn_vec(m, blk.b, B, training, update_running, s));
    if (blk.has_ds) CK(run_bn_apply(blk.b, B, blk.ds.y, blk.ds.bn.vec, blk.out, 1, s));
    else CK(run_bn_apply(blk.b, B, blk.in, nullptr, blk.out, 1, s));
  }
  const Block32& last = m.blocks.back();
  avgpool32_fwd_kernel<<<(B * 512 + 255) / 256, 256, 0, s>>>(last.out, m.feat, B, last.b.OH * last.b.OW, 512); ++g_cilrs_launches;
  CKL();
  CK(heads_forward_run(heads_ctx(m), B, speed, command, controls, pred_speed, keep, dropout_p, seed, nullptr, s));
  m.lastB = B; m.lastMode = mode;
  return OK;
}

static int backward(Model32& m, int B, int mode, const float* dcontrols, const float* dspeed, const float* speed, const long long* command,
                    float dropout_p, cudaStream_t s) {
  if (mode == MODE_INFER || B != m.lastB || mode != m.lastMode || !m.grads) return ERR_INVALID;
  const int frozen = mode == MODE_FROZEN;
  CK(cuda_status(cudaMemsetAsync(m.acc_bwd, 0, (size_t)m.acc_bwd_bytes, s)));
  CK(heads_backward_run(heads_ctx(m), B, dcontrols, dspeed, speed, command, dropout_p, s, s, nullptr));
  float *gA = m.g[0], *gB = m.g[1], *gC = m.g[2], *gD = m.g[3];
  {
    const Block32& last = m.blocks.back();
    const int HW = last.b.OH * last.b.OW;
    avgpool32_bwd_kernel<<<grid_for((long long)B * HW * 512), 256, 0, s>>>(m.dfeat, m.dfeat2, gA, B, HW, 512); ++g_cilrs_launches;
    CKL();
  }
  for (int bi = (int)m.blocks.size() - 1; bi >= 0; --bi) {
    Block32& blk = m.blocks[bi];
    // gA = gradient w.r.t. the block output (before its ReLU mask)
    CK(run_bn_bwd(m, blk.b, B, gA, blk.out, gA, gB, frozen, s));                    // gA <- dz, gB <- dy_b
    if (blk.has_ds) CK(run_bn_bwd(m, blk.ds, B, gA, nullptr, nullptr, gC, frozen, s));  // gC <- dy_ds (from the masked dz)
    CK(run_wgrad(m, blk.b, B, gB, blk.act_a, s));
    CK(run_dgrad(blk.b, B, gB, gD, nullptr, s));                                      // gD <- d act_a
    CK(run_bn_bwd(m, blk.a, B, gD, blk.act_a, gD, gB, frozen, s));                    // gB <- dy_a
    CK(run_wgrad(m, blk.a, B, gB, blk.in, s));
    if (blk.has_ds) {
      CK(run_wgrad(m, blk.ds, B, gC, blk.in, s));
      CK(run_dgrad(blk.a, B, gB, gD, nullptr, s));
      CK(run_dgrad(blk.ds, B, gC, gD, gD, s));                                        // gD <- dgrad_a + dgrad_ds
    } else {
      CK(run_dgrad(blk.a, B, gB, gD, gA, s));                                         // gD <- dgrad_a + identity (dz)
    }
    float* t = gA; gA = gD; gD = t;
  }
  // stem: max-pool backward, ReLU + BatchNorm backward, weight gradient (the image needs no gradient)
  maxpool32_bwd_kernel<<<grid_for((long long)B * 44 * 100 * 64), 256, 0, s>>>(gA, m.pool_arg, gD, B, 44, 100, 64, 22, 50); ++g_cilrs_launches;
  CKL();
  CK(run_bn_bwd(m, m.stem, B, gD, m.act0, gD, gB, frozen, s));
  CK(run_wgrad(m, m.stem, B, gB, m.x0, s));
  return OK;
}

}  // namespace f32
}  // namespace cilrs

using namespace cilrs;

extern "C" {

struct cilrs_model32 {
  f32::Model32 m;
};

static int init_layout(f32::Model32& m) {
  long long offs[256], sizes[256], tot = 0, buf = 0;
  int nbn = 0;
  const int n = cilrs_model_param_layout(offs, sizes, 256, &tot, &buf, &nbn);
  f32::build_topology(m);
  if (m.head_slot0 + HD_NUM_SLOTS != n || m.buffer_floats != buf || m.num_bn != nbn) return ERR_INVALID;
  m.off.assign(offs, offs + n);
  m.nslots = n;
  return OK;
}

size_t cilrs_model32_workspace_bytes(int max_batch) {
  if (max_batch < 1) return 0;
  f32::Model32 m;
  m.maxB = max_batch;
  if (init_layout(m)) return 0;
  return (size_t)f32::carve(m, nullptr);
}

int cilrs_model32_create(cilrs_model32** out, int max_batch, void* workspace, size_t workspace_bytes, void* stream) {
  if (!out || max_batch < 1 || !workspace || (((uintptr_t)workspace) & 255)) return ERR_INVALID;
  cilrs_model32* h = new (std::nothrow) cilrs_model32();
  if (!h) return ERR_INVALID;
  h->m.maxB = max_batch;
  int st = init_layout(h->m);
  if (!st && (size_t)f32::carve(h->m, (char*)workspace) > workspace_bytes) st = ERR_WORKSPACE;
  if (!st) st = cuda_status(cudaMemsetAsync(h->m.err_flag, 0, 64, (cudaStream_t)stream));
  if (!st) st = cuda_status(cudaMemsetAsync(h->m.counters, 0, 64, (cudaStream_t)stream));
  if (st) { delete h; return st; }
  *out = h;
  return OK;
}

void cilrs_model32_destroy(cilrs_model32* h) { delete h; }

int cilrs_model32_bind(cilrs_model32* h, float* params, float* grads, float* buffers, long long* num_batches_tracked) {
  if (!h || !params || !buffers) return ERR_INVALID;
  if ((((uintptr_t)params) | ((uintptr_t)grads) | ((uintptr_t)buffers)) & 15) return ERR_INVALID;
  h->m.params = params; h->m.grads = grads; h->m.buffers = buffers; h->m.nbt = num_batches_tracked;
  return OK;
}

int cilrs_model32_forward(cilrs_model32* h, int batch, int mode, const float* image_nchw, const float* speed, const long long* command,
                          float* controls, float* pred_speed, int update_running_stats, int keep_for_backward, float dropout_p,
                          unsigned long long seed, void* stream) {
  if (!h || !image_nchw || !speed || !command || !controls || !pred_speed || mode < 0 || mode > 2) return ERR_INVALID;
  return f32::forward(h->m, batch, mode, image_nchw, speed, command, controls, pred_speed, update_running_stats, keep_for_backward,
                      dropout_p, seed, (cudaStream_t)stream);
}

int cilrs_model32_backward(cilrs_model32* h, int batch, int mode, const float* dcontrols, const float* dspeed, const float* speed,
                           const long long* command, float dropout_p, void* stream) {
  if (!h || !dcontrols || !dspeed || !speed || !command) return ERR_INVALID;
  return f32::backward(h->m, batch, mode, dcontrols, dspeed, speed, command, dropout_p, (cudaStream_t)stream);
}

int* cilrs_model32_error_flag(cilrs_model32* h) { return h ? h->m.err_flag : nullptr; }

}  // extern "C"
