// N3 / N4 - the device side of the reference's input pipeline (notebook/notebook.ipynb:361-431) around the JPEG decoder:
//   weighted_sample_kernel   WeightedRandomSampler(weights, num_samples, replacement=True) (:411-413): inverse-CDF draws
//   augment_kernel           the albumentations Compose of :387-394 on uint8 RGB frames, one CTA per frame, parameters drawn
//                            on the host (a dozen scalars per frame), the per-pixel Gaussian noise drawn on the device (Philox)
#include "../../include/cilrs_b200.h"
#include "common.cuh"

namespace cilrs {

// ---- Philox4x32-10 (Salmon et al., SC'11): counter-based, so every draw is addressable by (seed, index) -----------------------
__device__ __forceinline__ uint4 philox4x32_10(uint4 ctr, uint2 key) {
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const uint32_t hi0 = __umulhi(0xD2511F53u, ctr.x), lo0 = 0xD2511F53u * ctr.x;
    const uint32_t hi1 = __umulhi(0xCD9E8D57u, ctr.z), lo1 = 0xCD9E8D57u * ctr.z;
    ctr = make_uint4(hi1 ^ ctr.y ^ key.x, lo1, hi0 ^ ctr.w ^ key.y, lo0);
    key.x += 0x9E3779B9u;
    key.y += 0xBB67AE85u;
  }
  return ctr;
}

// out[i] = smallest j with cdf[j] > u_i * cdf[n - 1],  u_i uniform in [0, 1) from 53 random bits
__global__ void weighted_sample_kernel(const double* __restrict__ cdf, long long n, long long num_samples, unsigned long long seed,
                                       unsigned long long offset, long long* __restrict__ out) {
  const double total = cdf[n - 1];
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < num_samples; i += (long long)gridDim.x * blockDim.x) {
    const unsigned long long c = offset + (unsigned long long)i;
    const uint4 r = philox4x32_10(make_uint4((uint32_t)c, (uint32_t)(c >> 32), 0x57a3u, 0u), make_uint2((uint32_t)seed, (uint32_t)(seed >> 32)));
    const unsigned long long bits = (((unsigned long long)r.x << 32) | r.y) >> 11;
    const double u = (double)bits * (1.0 / 9007199254740992.0) * total;
    long long lo = 0, hi = n - 1;   // invariant: answer in [lo, hi]
    while (lo < hi) {
      const long long mid = (lo + hi) >> 1;
      if (cdf[mid] > u) hi = mid; else lo = mid + 1;
    }
    out[i] = lo;
  }
}

// ---- augmentations -----------------------------------------------------------------------------------------------------------
struct AugmentParams {     // one frame (64 bytes); a zero flag skips the transform
  uint32_t flags;          // bit 0 brightness/contrast, 1 hue/saturation/value, 2 blur, 3 noise, 4 coarse dropout
  float alpha, beta;       // out = clip(v * alpha + beta * 255)   (RandomBrightnessContrast, brightness_by_max)
  int16_t hue, sat, val;   // HueSaturationValue shifts (OpenCV 8-bit HSV: H in [0, 180))
  int16_t ksize;           // GaussianBlur kernel size: 3 or 5 (sigma 0: OpenCV's fixed small kernels)
  float noise_std;         // GaussNoise standard deviation in [0, 255] units
  uint32_t n_holes;        // CoarseDropout: up to 3 rectangles [y0, y1) x [x0, x1), fill 0
  int16_t hole[3][4];
  uint32_t pad[3];
};
static_assert(sizeof(AugmentParams) == 64, "AugmentParams layout");

// OpenCV's 8-bit RGB -> HSV (imgproc color_hsv: RGB2HSV_b, hrange 180, hsv_shift 12), integer arithmetic throughout
__device__ __forceinline__ void rgb2hsv_u8(int r, int g, int b, int& h, int& s, int& v) {
  v = max(r, max(g, b));
  const int vmin = min(r, min(g, b));
  const int diff = v - vmin;
  const int vr = v == r ? -1 : 0, vg = v == g ? -1 : 0;
  // sdiv_table[v] = saturate_cast<int>((255 << 12) / (double)v), hdiv_table180[diff] = saturate_cast<int>((180 << 12) / (6. * diff)); [0] = 0
  const int sdiv = v ? __double2int_rn((double)(255 << 12) / (double)v) : 0;
  const int hdiv = diff ? __double2int_rn((double)(180 << 12) / (6.0 * (double)diff)) : 0;
  s = (diff * sdiv + (1 << 11)) >> 12;
  h = (vr & (g - b)) + (~vr & ((vg & (b - r + 2 * diff)) + ((~vg) & (r - g + 4 * diff))));
  h = (h * hdiv + (1 << 11)) >> 12;
  h += h < 0 ? 180 : 0;
}
// OpenCV's 8-bit HSV -> RGB (HSV2RGB_b: float HSV2RGB then saturate_cast<uchar>(x * 255)). The rounding order below - products
// rounded, the inner 1 - s * h as ONE fused multiply-add - is what OpenCV 4.13's x86 build executes; it was found by comparing
// all 180 x 256 x 256 inputs with cv2.cvtColor (0 mismatches; plain fp32 differs on 429 half-way cases, oracle/augment_oracle.py).
__device__ __forceinline__ void hsv2rgb_u8(int hi, int si, int vi, int& r, int& g, int& b) {
  const float s = __fmul_rn((float)si, 1.f / 255.f), v = __fmul_rn((float)vi, 1.f / 255.f);
  float fr, fg, fb;
  if (si == 0) {
    fr = fg = fb = v;
  } else {
    float h = __fmul_rn((float)hi, 6.f / 180.f);   // hi in [0, 180): h in [0, 6)
    int sector = (int)h;
    h = __fsub_rn(h, (float)sector);
    if ((unsigned)sector >= 6u) { sector = 0; h = 0.f; }
    const float t0 = v, t1 = __fmul_rn(v, __fsub_rn(1.f, s)), t2 = __fmul_rn(v, __fmaf_rn(-s, h, 1.f)),
                t3 = __fmul_rn(v, __fmaf_rn(-s, __fsub_rn(1.f, h), 1.f));
    // sector_data[][3] = {{1,3,0},{1,0,2},{3,0,1},{0,2,1},{0,1,3},{2,1,0}} -> (b, g, r)
    switch (sector) {
      case 0: fb = t1; fg = t3; fr = t0; break;
      case 1: fb = t1; fg = t0; fr = t2; break;
      case 2: fb = t3; fg = t0; fr = t1; break;
      case 3: fb = t0; fg = t2; fr = t1; break;
      case 4: fb = t0; fg = t1; fr = t3; break;
      default: fb = t2; fg = t1; fr = t0; break;
    }
  }
  r = min(max(__float2int_rn(__fmul_rn(fr, 255.f)), 0), 255);
  g = min(max(__float2int_rn(__fmul_rn(fg, 255.f)), 0), 255);
  b = min(max(__float2int_rn(__fmul_rn(fb, 255.f)), 0), 255);
}

__device__ __forceinline__ int reflect101(int i, int n) {   // BORDER_REFLECT_101
  if (i < 0) i = -i;
  if (i >= n) i = 2 * n - 2 - i;
  return i;
}

// One CTA per frame. Shared memory: the frame after the point-wise transforms (H*W*3 bytes), and - when the frame is blurred -
// the horizontally filtered rows as 16-bit fixed point (8 fractional bits, the intermediate OpenCV's 8-bit Gaussian blur keeps).
__global__ void __launch_bounds__(256) augment_kernel(const uint8_t* __restrict__ in, uint8_t* __restrict__ out, const AugmentParams* __restrict__ params,
                                                      int H, int W, unsigned long long seed, unsigned long long offset) {
  extern __shared__ __align__(16) uint8_t sm[];
  const int img = blockIdx.x;
  const AugmentParams p = params[img];
  const int npx = H * W, nb = npx * 3;
  const uint8_t* src = in + (long long)img * nb;
  uint8_t* dst = out + (long long)img * nb;
  uint8_t* s_img = sm;
  uint16_t* s_h = (uint16_t*)(sm + ((nb + 15) & ~15));
  if ((p.flags & 31u) == 0u) {   // untouched frame: plain copy
    if (src != dst) for (int i = threadIdx.x; i < nb; i += blockDim.x) dst[i] = src[i];
    return;
  }
  // ---- 1. point-wise: brightness / contrast, then hue / saturation / value ----
  for (int i = threadIdx.x; i < npx; i += blockDim.x) {
    int r = src[3 * i], g = src[3 * i + 1], b = src[3 * i + 2];
    if (p.flags & 1u) {
      const float add = __fmul_rn(p.beta, 255.f);   // lut = clip(arange(256, float32) * alpha + beta * 255, 0, 255).astype(uint8)
      r = (int)fminf(fmaxf(__fadd_rn(__fmul_rn((float)r, p.alpha), add), 0.f), 255.f);
      g = (int)fminf(fmaxf(__fadd_rn(__fmul_rn((float)g, p.alpha), add), 0.f), 255.f);
      b = (int)fminf(fmaxf(__fadd_rn(__fmul_rn((float)b, p.alpha), add), 0.f), 255.f);
    }
    if (p.flags & 2u) {
      int h, s, v;
      rgb2hsv_u8(r, g, b, h, s, v);
      h = (h + p.hue) % 180;
      if (h < 0) h += 180;
      s = min(max(s + p.sat, 0), 255);
      v = min(max(v + p.val, 0), 255);
      hsv2rgb_u8(h, s, v, r, g, b);
    }
    s_img[3 * i] = (uint8_t)r; s_img[3 * i + 1] = (uint8_t)g; s_img[3 * i + 2] = (uint8_t)b;
  }
  __syncthreads();
  // ---- 2. Gaussian blur, sigma 0 -> OpenCV's fixed kernels [1 2 1] / 4 and [1 4 6 4 1] / 16, separable, 8.8 fixed point ----
  const bool blur = (p.flags & 4u) && (p.ksize == 3 || p.ksize == 5);
  if (blur) {
    const int k3 = p.ksize == 3;
    for (int i = threadIdx.x; i < nb; i += blockDim.x) {
      const int c = i % 3, x = (i / 3) % W, y = i / (3 * W);
      const uint8_t* row = s_img + (long long)y * W * 3 + c;
      int acc;   // sum of pixel * weight * 256
      if (k3) acc = 64 * row[3 * reflect101(x - 1, W)] + 128 * row[3 * x] + 64 * row[3 * reflect101(x + 1, W)];
      else acc = 16 * (row[3 * reflect101(x - 2, W)] + row[3 * reflect101(x + 2, W)]) + 64 * (row[3 * reflect101(x - 1, W)] + row[3 * reflect101(x + 1, W)]) +
                 96 * row[3 * x];
      s_h[i] = (uint16_t)acc;
    }
    __syncthreads();
    for (int i = threadIdx.x; i < nb; i += blockDim.x) {
      const int y = i / (3 * W), col = i - y * 3 * W;
      uint32_t acc;   // 8.16 fixed point
      if (k3) acc = 64u * s_h[reflect101(y - 1, H) * 3 * W + col] + 128u * s_h[i] + 64u * s_h[reflect101(y + 1, H) * 3 * W + col];
      else acc = 16u * ((uint32_t)s_h[reflect101(y - 2, H) * 3 * W + col] + s_h[reflect101(y + 2, H) * 3 * W + col]) +
                 64u * ((uint32_t)s_h[reflect101(y - 1, H) * 3 * W + col] + s_h[reflect101(y + 1, H) * 3 * W + col]) + 96u * s_h[i];
      s_img[i] = (uint8_t)min((acc + (1u << 15)) >> 16, 255u);
    }
    __syncthreads();
  }
  // ---- 3. Gaussian noise (per pixel and channel), 4. coarse dropout, store ----
  for (int i4 = threadIdx.x; i4 * 4 < nb; i4 += blockDim.x) {
    float z[4] = {0.f, 0.f, 0.f, 0.f};
    if (p.flags & 8u) {
      const unsigned long long c = offset + (unsigned long long)img * (unsigned long long)((nb + 3) / 4) + (unsigned long long)i4;
      const uint4 rr = philox4x32_10(make_uint4((uint32_t)c, (uint32_t)(c >> 32), 0xa06eu, 0u), make_uint2((uint32_t)seed, (uint32_t)(seed >> 32)));
      // Box-Muller on two pairs of uniforms in (0, 1]
      const float u0 = ((float)(rr.x >> 8) + 1.f) * (1.f / 16777216.f), u1 = (float)(rr.y >> 8) * (1.f / 16777216.f);
      const float u2 = ((float)(rr.z >> 8) + 1.f) * (1.f / 16777216.f), u3 = (float)(rr.w >> 8) * (1.f / 16777216.f);
      const float m0 = sqrtf(-2.f * logf(u0)), m1 = sqrtf(-2.f * logf(u2));
      float s0, c0, s1, c1;
      sincospif(2.f * u1, &s0, &c0);
      sincospif(2.f * u3, &s1, &c1);
      z[0] = m0 * c0; z[1] = m0 * s0; z[2] = m1 * c1; z[3] = m1 * s1;
    }
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int i = i4 * 4 + k;
      if (i >= nb) break;
      int v = s_img[i];
      if (p.flags & 8u) v = (int)fminf(fmaxf(__fadd_rn((float)v, __fmul_rn(z[k], p.noise_std)), 0.f), 255.f);
      if (p.flags & 16u) {
        const int x = (i / 3) % W, y = i / (3 * W);
        for (uint32_t hidx = 0; hidx < p.n_holes && hidx < 3u; ++hidx)
          if (y >= p.hole[hidx][0] && y < p.hole[hidx][1] && x >= p.hole[hidx][2] && x < p.hole[hidx][3]) v = 0;
      }
      dst[i] = (uint8_t)v;
    }
  }
}

}  // namespace cilrs

using namespace cilrs;

extern "C" {

int cilrs_weighted_sample(const double* cdf_dev, long long n, long long num_samples, unsigned long long seed, unsigned long long offset,
                          long long* out_dev, void* stream) {
  if (!cdf_dev || n < 1 || num_samples < 0 || !out_dev) return ERR_INVALID;
  if (num_samples == 0) return OK;
  long long grid = (num_samples + 255) / 256;
  if (grid > 148 * 8) grid = 148 * 8;
  ++g_cilrs_launches;
  weighted_sample_kernel<<<(int)grid, 256, 0, (cudaStream_t)stream>>>(cdf_dev, n, num_samples, seed, offset, out_dev);
  return cuda_status(cudaGetLastError());
}

size_t cilrs_augment_param_bytes(void) { return sizeof(AugmentParams); }

int cilrs_augment_u8(const unsigned char* in, unsigned char* out, const void* params_dev, int n, int height, int width,
                     unsigned long long seed, unsigned long long offset, void* stream) {
  if (!in || !out || !params_dev || n < 0 || height < 1 || width < 1) return ERR_INVALID;
  if (n == 0) return OK;
  const size_t nb = (size_t)height * width * 3;
  const size_t smem = ((nb + 15) & ~(size_t)15) + nb * 2;
  if (smem > 200 * 1024) return ERR_UNSUPPORTED;   // frames are 88 x 200 (158 KB); larger ones would need row bands
  static bool attr_done = false;
  if (!attr_done) {
    cudaError_t e = cudaFuncSetAttribute(augment_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    if (e != cudaSuccess) return cuda_status(e);
    attr_done = true;
  }
  ++g_cilrs_launches;
  augment_kernel<<<n, 256, smem, (cudaStream_t)stream>>>(in, out, (const AugmentParams*)params_dev, height, width, seed, offset);
  return cuda_status(cudaGetLastError());
}

}  // extern "C"
