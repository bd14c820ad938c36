// K4: fused flat-buffer Adam with L2-coupled weight decay (torch.optim.Adam semantics; reference:
// notebook/notebook.ipynb:533-534,555) + a deterministic sum-of-squares reduction for gradient-norm clipping
// (notebook/notebook.ipynb:553-554). HBM-bound: 16 B read + 12 B written per parameter.
#pragma once
#include "common.cuh"

namespace cilrs {

struct AdamParams {
  float* p;
  float* g;                     // fp32 gradient arena (read; zeroed afterwards when zero_grad is set)
  const __nv_bfloat16* g16;     // optional: bf16 gradients (the all-reduced communication buffer) used INSTEAD of g
  float* m;
  float* v;
  long long n;  // multiple of 4
  float lr, beta1, beta2, eps, weight_decay;
  float bias_correction1;       // 1 - beta1^t
  float bias_correction2_sqrt;  // sqrt(1 - beta2^t)
  float grad_scale;             // applied to g first (clipping coefficient and/or 1/world_size)
  const float* grad_scale_dev;  // optional device scalar multiplied in as well (clip coefficient computed on device)
  const long long* step_dev;    // optional device step counter: bias corrections are then computed on the device, which
                                // keeps a captured CUDA graph of the training step valid for every step number
  const float* hyper_dev;       // optional device float[8] = {lr, beta1, beta2, eps, weight_decay, grad_scale, -, -} read at run
                                // time INSTEAD of the by-value fields: an LR schedule (StepLR, notebook/notebook.ipynb:535-536,604)
                                // then reaches a captured graph by a 32-byte copy into this buffer
  int zero_grad;                // also write zeros to g (optimizer.zero_grad() of the NEXT step fused here)
};

__global__ void step_increment_kernel(long long* step) { *step += 1; }

__global__ void __launch_bounds__(256) adam_kernel(const AdamParams a) {
  const long long n4 = a.n >> 2;
  float lr = a.lr, beta1 = a.beta1, beta2 = a.beta2, eps = a.eps, wd = a.weight_decay, gs = a.grad_scale;
  if (a.hyper_dev) {
    lr = __ldg(a.hyper_dev); beta1 = __ldg(a.hyper_dev + 1); beta2 = __ldg(a.hyper_dev + 2); eps = __ldg(a.hyper_dev + 3);
    wd = __ldg(a.hyper_dev + 4); gs = __ldg(a.hyper_dev + 5);
  }
  if (a.grad_scale_dev) gs *= *a.grad_scale_dev;
  float bc1 = a.bias_correction1, bc2s = a.bias_correction2_sqrt;
  if (a.step_dev) {
    const double t = (double)*a.step_dev;
    bc1 = (float)(1.0 - pow((double)beta1, t));
    bc2s = (float)sqrt(1.0 - pow((double)beta2, t));
  }
  const float step_size = lr / bc1;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
    float4 p4 = reinterpret_cast<float4*>(a.p)[i];
    float4 g4;
    if (a.g16) {
      const uint2 r = reinterpret_cast<const uint2*>(a.g16)[i];
      g4 = make_float4(bf16lo(r.x), bf16hi(r.x), bf16lo(r.y), bf16hi(r.y));
    } else {
      g4 = reinterpret_cast<const float4*>(a.g)[i];
    }
    float4 m4 = reinterpret_cast<float4*>(a.m)[i];
    float4 v4 = reinterpret_cast<float4*>(a.v)[i];
    float* pp = &p4.x; const float* gg = &g4.x; float* mm = &m4.x; float* vv = &v4.x;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const float g = fmaf(wd, pp[k], gg[k] * gs);
      mm[k] = mm[k] + (1.f - beta1) * (g - mm[k]);                // exp_avg.lerp_(grad, 1 - beta1)
      vv[k] = beta2 * vv[k] + (1.f - beta2) * g * g;              // exp_avg_sq.mul_(beta2).addcmul_(g, g, 1 - beta2)
      const float denom = sqrtf(vv[k]) / bc2s + eps;
      pp[k] = pp[k] - step_size * (mm[k] / denom);                // param.addcdiv_(exp_avg, denom, -step_size)
    }
    reinterpret_cast<float4*>(a.p)[i] = p4;
    reinterpret_cast<float4*>(a.m)[i] = m4;
    reinterpret_cast<float4*>(a.v)[i] = v4;
    // (a store to g issued right behind the load of the same address serialised the LSU: 431 us instead of 108 us measured on
    //  B200 - the zeroing has to come after the loaded value has been consumed)
    if (a.zero_grad) reinterpret_cast<float4*>(a.g)[i] = make_float4(0.f, 0.f, 0.f, 0.f);
  }
}

// out[0] = sum x^2 ; out[1] = clip coefficient min(1, max_norm / (sqrt(sum) + 1e-6)) (clip_grad_norm_ semantics)
// x16 != nullptr: the gradients are the bf16 communication buffer (data-parallel training with bf16 gradient exchange)
__global__ void __launch_bounds__(256) sumsq_kernel(const float* __restrict__ x, const __nv_bfloat16* __restrict__ x16, long long n,
                                                    double* partial, unsigned int* counter, float* out, float max_norm) {
  __shared__ double red[256];
  __shared__ bool last;
  double s = 0.0;
  const long long n4 = n >> 2;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
    float4 a;
    if (x16) {
      const uint2 r = reinterpret_cast<const uint2*>(x16)[i];
      a = make_float4(bf16lo(r.x), bf16hi(r.x), bf16lo(r.y), bf16hi(r.y));
    } else {
      a = reinterpret_cast<const float4*>(x)[i];
    }
    s += (double)(a.x * a.x + a.y * a.y) + (double)(a.z * a.z + a.w * a.w);
  }
  red[threadIdx.x] = s;
  __syncthreads();
  for (int k = 128; k > 0; k >>= 1) {
    if (threadIdx.x < k) red[threadIdx.x] += red[threadIdx.x + k];
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    partial[blockIdx.x] = red[0];
    __threadfence();
    last = atomicAdd(counter, 1u) == gridDim.x - 1;
  }
  __syncthreads();
  if (last) {
    __threadfence();
    double t = 0.0;
    for (unsigned int b = threadIdx.x; b < gridDim.x; b += 256) t += partial[b];
    red[threadIdx.x] = t;
    __syncthreads();
    for (int k = 128; k > 0; k >>= 1) {
      if (threadIdx.x < k) red[threadIdx.x] += red[threadIdx.x + k];
      __syncthreads();
    }
    if (threadIdx.x == 0) {
      out[0] = (float)red[0];
      const float norm = (float)sqrt(red[0]);
      const float c = max_norm / (norm + 1e-6f);
      out[1] = c < 1.f ? c : 1.f;
      *counter = 0u;
    }
  }
}

// fp32 gradient range -> bf16 communication buffer (round to nearest even), optionally zeroing the fp32 source so that the
// next backward can accumulate into it again. 6 (10 with zeroing) bytes per element, HBM-bound.
__global__ void __launch_bounds__(256) grad_to_bf16_kernel(float* __restrict__ g, __nv_bfloat16* __restrict__ out, long long n, int zero) {
  const long long n4 = n >> 2;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
    const float4 a = reinterpret_cast<const float4*>(g)[i];
    reinterpret_cast<uint2*>(out)[i] = make_uint2(pack_bf16x2(a.x, a.y), pack_bf16x2(a.z, a.w));
    if (zero) reinterpret_cast<float4*>(g)[i] = make_float4(0.f, 0.f, 0.f, 0.f);
  }
}

}  // namespace cilrs
