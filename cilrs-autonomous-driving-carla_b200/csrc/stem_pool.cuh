// The stem's max-pool, forward and backward, in the forms the training plan uses (round 2). Reference: torchvision resnet34's
// `bn1 -> relu -> maxpool(3, 2, 1)` between conv1 and layer1 (model/autonomous_drive.py:365-370) and their autograd.
//
//  bn_relu_maxpool_sel_kernel   out = maxpool3x3/2( relu( y*scale + shift ) ), the arg-max code (0..8, first maximum in scan order)
//                               and - for the backward - `ysel`, the RAW conv1 output at the arg-max position.
//                               The old kernel (elementwise.cuh) was instruction-bound (ncu: 67 % issue slots, 38 M warp
//                               instructions for 72 MB): nine windows x eight channels of unpack / fma / max / round / compare /
//                               select. Here two channels travel in one register: `cvt.rn.relu.bf16x2.f32` rounds and clamps a
//                               pair in one instruction, and because the activations are non-negative bf16 their bit patterns
//                               order like unsigned integers, so value and (15 - index) share one 20-bit key and a single
//                               integer max per channel does compare + select + first-index tie-break.
//  (reduce)                     With `ysel` the BatchNorm-backward sums of the stem need no pass over the 4x larger conv1 output:
//                               every pooled gradient lands on exactly one conv1 pixel, so  sum dz = sum over pool outputs of
//                               g * [act > 0]  and  sum dz * xhat = sum g * [act > 0] * (ysel - mean) * rstd  - the ordinary
//                               bn_bwd_reduce_kernel<false> on (g, pool_out, ysel): 57 MB instead of 172 MB.
//                               The same launch stores g * [act > 0] (its dz_out, in place): the ReLU of a conv1 pixel only
//                               matters where a window selected it, and there it equals "the pooled activation is positive".
//  stem_bwd_apply_kernel        routes that masked pooled gradient through the arg-max codes and applies the BatchNorm
//                               backward in one pass: one thread per 2x2 block of conv1 pixels x 8 channels (the four windows
//                               that can select them are loaded once), the routed sum in packed bf16 arithmetic,
//                               dy = A*dz + (C*y + B) with per-channel constants.
//                               Replaces "reduce<STEM> writes dz (72 MB) + apply reads it back".
#pragma once
#include "elementwise.cuh"

namespace cilrs {

// bf16( relu( x * scale + shift ) ) for the two channels of one packed word (round to nearest even, negative -> +0)
CILRS_DEVINL uint32_t bn_relu_bf16x2(uint32_t w, float sc0, float sh0, float sc1, float sh1) {
  const float a = fmaf(__uint_as_float(w << 16), sc0, sh0), b = fmaf(__uint_as_float(w & 0xFFFF0000u), sc1, sh1);
  uint32_t r;
  asm("cvt.rn.relu.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(b), "f"(a));   // first source -> upper half
  return r;
}

// y dense [B,H,W,C] -> out padded-flat [B,OHp,OWp,C] (padding pixels untouched), argmax dense [B,OH,OW,C] (may be null),
// ysel padded-flat like out (may be null)
__global__ void __launch_bounds__(EW_THREADS) bn_relu_maxpool_sel_kernel(const __nv_bfloat16* __restrict__ y, const float* __restrict__ scale,
                                                                         const float* __restrict__ shift, __nv_bfloat16* __restrict__ out,
                                                                         uint8_t* __restrict__ argmax, __nv_bfloat16* __restrict__ ysel, int B,
                                                                         int H, int W, int C, int OH, int OW, int OHp, int OWp) {
  pdl_entry();
  // (32-bit index arithmetic: the host checks that every tensor has fewer than 2^31 elements)
  const unsigned int groups = (unsigned int)C >> 3;
  const unsigned int nvec = (unsigned int)B * OH * OW * groups;
  const unsigned int stride = gridDim.x * EW_THREADS;
  unsigned int i = blockIdx.x * EW_THREADS + threadIdx.x;
  const int cg = (int)(i % groups) * 8;
  const Vec8 sc = loadf8(scale + cg), sh = loadf8(shift + cg);
  for (; i < nvec; i += stride) {
    const unsigned int pix = i / groups;
    const unsigned int t1 = pix / (unsigned int)OW;
    const int ow = (int)(pix - t1 * OW);
    const int n = (int)(t1 / (unsigned int)OH);
    const int oh = (int)(t1 - (unsigned int)n * OH);
    const int h0 = oh * 2 - 1, w0 = ow * 2 - 1;
    const __nv_bfloat16* base = y + ((n * H + h0) * W + w0) * C + cg;   // window origin (may lie outside: never dereferenced there)
    uint4 q[9];
    bool okq[9];
#pragma unroll
    for (int r = 0; r < 3; ++r) {
#pragma unroll
      for (int s = 0; s < 3; ++s) {
        okq[r * 3 + s] = h0 + r >= 0 && h0 + r < H && w0 + s >= 0 && w0 + s < W;
        if (okq[r * 3 + s]) q[r * 3 + s] = *reinterpret_cast<const uint4*>(base + (r * W + s) * C);
      }
    }
    uint32_t key[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) key[k] = 0u;
#pragma unroll
    for (int e = 0; e < 9; ++e) {
      if (!okq[e]) continue;
      const uint32_t w4[4] = {q[e].x, q[e].y, q[e].z, q[e].w};
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const uint32_t t = bn_relu_bf16x2(w4[j], sc.v[2 * j], sh.v[2 * j], sc.v[2 * j + 1], sh.v[2 * j + 1]);
        // key = value bits (non-negative bf16: ordered like unsigned integers) << 4 | (15 - e): the first maximum wins ties
        key[2 * j] = max(key[2 * j], ((t << 4) & 0xFFFF0u) | (uint32_t)(15 - e));
        key[2 * j + 1] = max(key[2 * j + 1], ((t >> 12) & 0xFFFF0u) | (uint32_t)(15 - e));
      }
    }
    uint4 o;
    o.x = (key[0] >> 4) | ((key[1] << 12) & 0xFFFF0000u); o.y = (key[2] >> 4) | ((key[3] << 12) & 0xFFFF0000u);
    o.z = (key[4] >> 4) | ((key[5] << 12) & 0xFFFF0000u); o.w = (key[6] >> 4) | ((key[7] << 12) & 0xFFFF0000u);
    const unsigned int po = (((unsigned int)n * OHp + oh) * OWp + ow) * (unsigned int)C + cg;
    *reinterpret_cast<uint4*>(out + po) = o;
    if (argmax) {
      uint32_t idx[8];
#pragma unroll
      for (int k = 0; k < 8; ++k) idx[k] = 15u - (key[k] & 15u);
      uint2 u;
      u.x = idx[0] | (idx[1] << 8) | (idx[2] << 16) | (idx[3] << 24);
      u.y = idx[4] | (idx[5] << 8) | (idx[6] << 16) | (idx[7] << 24);
      *reinterpret_cast<uint2*>(argmax + (size_t)i * 8) = u;
      if (ysel) {
        // the raw conv1 output at the arg-max: re-read from the lines the window loads just brought into L1
        const unsigned short* yb = reinterpret_cast<const unsigned short*>(base);
        uint32_t v[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          const uint32_t r = (idx[k] * 11u) >> 5, s = idx[k] - 3u * r;   // idx / 3, idx % 3 for idx in 0..8
          v[k] = yb[(r * W + s) * C + k];
        }
        *reinterpret_cast<uint4*>(ysel + po) = make_uint4(v[0] | (v[1] << 16), v[2] | (v[3] << 16), v[4] | (v[5] << 16), v[6] | (v[7] << 16));
      }
    }
  }
}

struct StemBwdApplyParams {
  const __nv_bfloat16* g;       // pooled gradient ALREADY masked by the ReLU (g * [pooled activation > 0]), padded-flat [B,OHp,OWp,C]
  const uint8_t* argmax;        // dense [B,OH,OW,C]
  const __nv_bfloat16* y;       // raw conv1 output, dense [B,H,W,C]
  const float* mean;
  const float* rstd;
  const float* gamma;
  const float* bsum;            // [C] sum dz, [C] sum dz * xhat (written by the reduce pass)
  const float* bdot;
  float inv_count;
  int frozen;
  int B, H, W, OH, OW, OHp, OWp, C;
  __nv_bfloat16* dy;            // dense [B,H,W,C]
};

// 0xFFFF in the half of the word whose byte of `m` (a __vcmpeq4 result: 0xFF / 0x00 per byte) is set: bytes (2j, 2j+1) of m
CILRS_DEVINL uint32_t halves_from_bytes(uint32_t m, int pair) { return __byte_perm(m, 0u, pair ? 0x3322u : 0x1100u); }
CILRS_DEVINL uint32_t add_bf16x2(uint32_t a, uint32_t b) {
  uint32_t r;
  asm("add.rn.bf16x2 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b));
  return r;
}

// H and W even. One thread per (2x2 block of conv1 pixels, 8 channels).
__global__ void __launch_bounds__(EW_THREADS, 2) stem_bwd_apply_kernel(const StemBwdApplyParams p) {
  pdl_entry();
  const unsigned int groups = (unsigned int)p.C >> 3;
  const unsigned int HB = (unsigned int)p.H >> 1, WB = (unsigned int)p.W >> 1;
  const unsigned int nblk = (unsigned int)p.B * HB * WB * groups;
  const unsigned int stride = gridDim.x * EW_THREADS;
  unsigned int bi = blockIdx.x * EW_THREADS + threadIdx.x;
  const int cg = (int)(bi % groups) * 8;
  // dy = gamma*rstd * (dz - bsum/n - xhat * bdot/n), xhat = (y - mean) * rstd   ==   A*dz + (Cy*y + B0)
  Vec8 A, Cy, B0;
  {
    const Vec8 mean = loadf8(p.mean + cg), rstd = loadf8(p.rstd + cg), gamma = loadf8(p.gamma + cg);
    const Vec8 bs = loadf8(p.bsum + cg), bd = loadf8(p.bdot + cg);
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const float k0 = p.frozen ? 0.f : bs.v[k] * p.inv_count, k1 = p.frozen ? 0.f : bd.v[k] * p.inv_count;
      A.v[k] = gamma.v[k] * rstd.v[k];
      Cy.v[k] = -A.v[k] * k1 * rstd.v[k];
      B0.v[k] = -A.v[k] * k0 - Cy.v[k] * mean.v[k];
    }
  }
  for (; bi < nblk; bi += stride) {
    const unsigned int blk = bi / groups;
    const unsigned int t1 = blk / WB;
    const int b = (int)(blk - t1 * WB);
    const int n = (int)(t1 / HB);
    const int a = (int)(t1 - (unsigned int)n * HB);
    // the four pool windows (a + qh, b + qw) that can have selected a pixel of this block
    const uint8_t* am0 = p.argmax + (size_t)(((unsigned int)(n * p.OH + a) * p.OW + b) * (unsigned int)p.C + cg);
    const __nv_bfloat16* g0 = p.g + (size_t)(((unsigned int)(n * p.OHp + a) * p.OWp + b) * (unsigned int)p.C + cg);
    uint2 am[4];
    uint4 gq[4];
    bool okw[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      okw[q] = a + (q >> 1) < p.OH && b + (q & 1) < p.OW;
      if (okw[q]) {
        am[q] = *reinterpret_cast<const uint2*>(am0 + ((q >> 1) * p.OW + (q & 1)) * p.C);
        gq[q] = *reinterpret_cast<const uint4*>(g0 + ((q >> 1) * p.OWp + (q & 1)) * p.C);
      }
    }
    const unsigned int yo = ((unsigned int)(n * p.H + 2 * a) * p.W + 2 * b) * (unsigned int)p.C + cg;   // pixel (2a, 2b)
    uint4 yq[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) yq[e] = *reinterpret_cast<const uint4*>(p.y + (size_t)yo + ((e >> 1) * p.W + (e & 1)) * p.C);
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const int dh = e >> 1, dw = e & 1;   // pixel (2a + dh, 2b + dw)
      uint32_t dz[4] = {0u, 0u, 0u, 0u};   // routed gradient, packed bf16 (a pixel selected by one window - the usual case - is exact)
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        // window (a + qh, b + qw) covers rows 2(a+qh)-1 .. 2(a+qh)+1: the pixel's row inside it is r = dh + 1 - 2 qh
        const int r = dh + 1 - 2 * (q >> 1), sx = dw + 1 - 2 * (q & 1);
        if (r < 0 || sx < 0) continue;      // (compile-time after unrolling)
        if (okw[q]) {
          const uint32_t cc = (uint32_t)(r * 3 + sx) * 0x01010101u;
          const uint32_t m_lo = __vcmpeq4(am[q].x, cc), m_hi = __vcmpeq4(am[q].y, cc);
          dz[0] = add_bf16x2(dz[0], gq[q].x & halves_from_bytes(m_lo, 0));
          dz[1] = add_bf16x2(dz[1], gq[q].y & halves_from_bytes(m_lo, 1));
          dz[2] = add_bf16x2(dz[2], gq[q].z & halves_from_bytes(m_hi, 0));
          dz[3] = add_bf16x2(dz[3], gq[q].w & halves_from_bytes(m_hi, 1));
        }
      }
      const uint32_t yw[4] = {yq[e].x, yq[e].y, yq[e].z, yq[e].w};
      uint32_t ow4[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int c0 = 2 * j, c1 = 2 * j + 1;
        const float o0 = fmaf(A.v[c0], bf16lo(dz[j]), fmaf(Cy.v[c0], bf16lo(yw[j]), B0.v[c0]));
        const float o1 = fmaf(A.v[c1], bf16hi(dz[j]), fmaf(Cy.v[c1], bf16hi(yw[j]), B0.v[c1]));
        ow4[j] = pack_bf16x2(o0, o1);
      }
      *reinterpret_cast<uint4*>(p.dy + (size_t)yo + (dh * p.W + dw) * p.C) = make_uint4(ow4[0], ow4[1], ow4[2], ow4[3]);
    }
  }
}

}  // namespace cilrs
