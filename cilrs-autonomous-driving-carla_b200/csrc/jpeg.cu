// N3 - the input pipeline's decode step on the GPU: baseline JPEG (the collector's 200 x 88 quality-95 frames,
// model/collect_data.py:685-716) -> uint8 RGB [B, H, W, 3], bit-identical to the reference loader's
// cv2.imread + cvtColor(BGR2RGB) (notebook/notebook.ipynb:404-405), i.e. to libjpeg-turbo at its defaults:
// Huffman decode, jidctint.c's jpeg_idct_islow, jdsample.c's h2v2_fancy_upsample, jdcolor.c's ycc_rgb_convert.
// (oracle/jpeg_oracle.py restates the same algorithms in numpy and is pinned against cv2.imdecode.)
//
//   host   cilrs_jpeg_prepare   marker parsing, quantisation tables, derived Huffman lookup tables (per distinct DHT set)
//   kernel jpeg_entropy_idct    one warp per image: lane 0 walks the bit stream (entropy coding is sequential by construction
//                               and the collector writes no restart markers), after every MCU all 32 lanes de-quantise and run
//                               the two 1-D passes of the inverse DCT and store the 8 x 8 sample blocks into component planes
//   kernel jpeg_color           fancy chroma upsampling + YCbCr -> RGB, four pixels (12 bytes) per thread
// A batch of 128 frames is 128 independent warps (~0.4 ms of latency, far below one SM-second): it is meant to run on a side
// stream under the previous training step, not to be fast by itself.
#include <cstddef>
#include <cstdio>
#include <cstring>
#include <string>
#include <thread>
#include <vector>

#include "../../include/cilrs_b200.h"
#include "common.cuh"

namespace cilrs {

struct JpegHuffTable {
  uint16_t look[512];   // 9-bit look-ahead: (code length << 8) | symbol, 0 = longer than 9 bits
  int32_t maxcode[18];  // largest code of each length (-1: none); [17] = sentinel
  int32_t valoff[18];   // huffval index of the first code of each length minus that code
  uint8_t huffval[256];
};
struct JpegHuffSet { JpegHuffTable t[4]; };  // index = class * 2 + id (DC0, DC1, AC0, AC1)

struct JpegDesc {          // one image of the batch (48 + 256 bytes)
  uint32_t data_off;       // byte offset of the stream inside the batch buffer
  uint32_t data_len;
  uint32_t scan_off;       // offset of the entropy-coded segment inside the stream
  uint16_t width, height;
  uint8_t mode;            // 0 = YCbCr 4:2:0 (h2v2), 1 = YCbCr 4:4:4, 2 = grey
  uint8_t huff_set;
  uint8_t dc_id[3], ac_id[3], qt_id[3];
  uint8_t pad;
  uint16_t mcus_x, mcus_y;
  uint32_t status;         // 0 ok, else CILRS_JPEG_* (host-side parse verdict)
  uint32_t reserved[3];
  uint16_t qt[2][64];      // de-quantisation tables in natural order (slot 0: the one component 0 uses, slot 1: chroma)
};
static_assert(sizeof(JpegDesc) == 304 && offsetof(JpegDesc, qt) == 48, "descriptor layout");

__constant__ uint8_t c_zigzag[64] = {0,  1,  8,  16, 9,  2,  3,  10, 17, 24, 32, 25, 18, 11, 4,  5,  12, 19, 26, 33, 40, 48,
                                     41, 34, 27, 20, 13, 6,  7,  14, 21, 28, 35, 42, 49, 56, 57, 50, 43, 36, 29, 22, 15, 23,
                                     30, 37, 44, 51, 58, 59, 52, 45, 38, 31, 39, 46, 53, 60, 61, 54, 47, 55, 62, 63};

// ---- bit reader (one lane) -------------------------------------------------------------------------------------------------
struct BitReader {
  const uint8_t* p;
  const uint8_t* end;
  uint64_t acc;   // the next bits, MSB first
  int n;          // valid bits in acc
  bool marker;    // a marker was reached: zeros are fed from here on (libjpeg's behaviour on truncated data)
  __device__ void fill() {
    while (n <= 32) {
      if (!marker && (((uintptr_t)p & 3) == 0) && p + 4 <= end) {
        const uint32_t w = __ldg((const uint32_t*)p);
        if (__vcmpeq4(w, 0xFFFFFFFFu) == 0) {   // no 0xFF byte: four stream bytes at once
          acc |= (uint64_t)__byte_perm(w, 0, 0x0123) << (32 - n);
          n += 32;
          p += 4;
          continue;
        }
      }
      if (n > 56) break;
      uint32_t c = 0;
      if (!marker && p < end) {
        c = __ldg(p++);
        if (c == 0xFF) {
          const uint32_t c2 = p < end ? __ldg(p) : 0xD9u;
          if (c2 == 0) ++p;                       // stuffed zero
          else { marker = true; --p; c = 0; }
        }
      } else {
        marker = true;
      }
      acc |= (uint64_t)c << (56 - n);
      n += 8;
    }
  }
  __device__ uint32_t peek(int k) const { return (uint32_t)(acc >> (64 - k)); }
  __device__ void skip(int k) { acc <<= k; n -= k; }
};

__device__ __forceinline__ int jpeg_decode_symbol(BitReader& br, const JpegHuffTable* t, bool& bad) {
  if (br.n < 16) br.fill();
  const uint32_t e = t->look[br.peek(9)];
  if (e) {
    br.skip((int)(e >> 8));
    return (int)(e & 255u);
  }
  const uint32_t code16 = br.peek(16);
#pragma unroll 1
  for (int l = 10; l <= 16; ++l) {
    const int code = (int)(code16 >> (16 - l));
    if (code <= t->maxcode[l]) {
      br.skip(l);
      return t->huffval[(t->valoff[l] + code) & 255];
    }
  }
  bad = true;
  br.skip(16);
  return 0;
}
__device__ __forceinline__ int jpeg_receive_extend(BitReader& br, int s) {
  if (br.n < s) br.fill();
  const int v = (int)br.peek(s);
  br.skip(s);
  return v < (1 << (s - 1)) ? v - (1 << s) + 1 : v;
}

// ---- jidctint.c (jpeg_idct_islow), one 1-D pass ------------------------------------------------------------------------------
#define JF_0_298631336 2446
#define JF_0_390180644 3196
#define JF_0_541196100 4433
#define JF_0_765366865 6270
#define JF_0_899976223 7373
#define JF_1_175875602 9633
#define JF_1_501321110 12299
#define JF_1_847759065 15137
#define JF_1_961570560 16069
#define JF_2_053119869 16819
#define JF_2_562915447 20995
#define JF_3_072711026 25172
template <int SHIFT>
__device__ __forceinline__ void idct_islow_1d(const int x[8], int o[8]) {
  int z1 = (x[2] + x[6]) * JF_0_541196100;
  const int tmp2 = z1 + x[6] * (-JF_1_847759065);
  const int tmp3 = z1 + x[2] * JF_0_765366865;
  const int tmp0 = (x[0] + x[4]) << 13;
  const int tmp1 = (x[0] - x[4]) << 13;
  const int tmp10 = tmp0 + tmp3, tmp13 = tmp0 - tmp3, tmp11 = tmp1 + tmp2, tmp12 = tmp1 - tmp2;
  int t0 = x[7], t1 = x[5], t2 = x[3], t3 = x[1];
  z1 = t0 + t3;
  int z2 = t1 + t2, z3 = t0 + t2, z4 = t1 + t3;
  const int z5 = (z3 + z4) * JF_1_175875602;
  t0 *= JF_0_298631336; t1 *= JF_2_053119869; t2 *= JF_3_072711026; t3 *= JF_1_501321110;
  z1 *= -JF_0_899976223; z2 *= -JF_2_562915447;
  z3 = z3 * (-JF_1_961570560) + z5;
  z4 = z4 * (-JF_0_390180644) + z5;
  t0 += z1 + z3; t1 += z2 + z4; t2 += z2 + z3; t3 += z1 + z4;
  const int r = 1 << (SHIFT - 1);
  o[0] = (tmp10 + t3 + r) >> SHIFT; o[7] = (tmp10 - t3 + r) >> SHIFT;
  o[1] = (tmp11 + t2 + r) >> SHIFT; o[6] = (tmp11 - t2 + r) >> SHIFT;
  o[2] = (tmp12 + t1 + r) >> SHIFT; o[5] = (tmp12 - t1 + r) >> SHIFT;
  o[3] = (tmp13 + t0 + r) >> SHIFT; o[4] = (tmp13 - t0 + r) >> SHIFT;
}

constexpr int JPEG_WARPS = 4;

// planes of image i: Y [mcus_y * vy * 8][mcus_x * hy * 8] then Cb, Cr [mcus_y * 8][mcus_x * 8]; plane_stride bytes per image
__global__ void __launch_bounds__(JPEG_WARPS * 32) jpeg_entropy_idct_kernel(const uint8_t* __restrict__ bytes, const JpegDesc* __restrict__ descs,
                                                                           const JpegHuffSet* __restrict__ sets, int n, int H, int W,
                                                                           uint8_t* __restrict__ planes, long long plane_stride,
                                                                           uint32_t* __restrict__ status) {
  __shared__ JpegHuffSet s_set;                           // the table set of this CTA's first image (normally the batch's only one)
  __shared__ __align__(16) int16_t s_coef[JPEG_WARPS][6 * 64];
  __shared__ __align__(16) int s_ws[JPEG_WARPS][6 * 64];
  __shared__ uint16_t s_qt[JPEG_WARPS][2][64];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int img0 = blockIdx.x * JPEG_WARPS;
  const int set0 = descs[img0].huff_set;
  {
    const uint32_t* src = (const uint32_t*)&sets[set0];
    uint32_t* dst = (uint32_t*)&s_set;
    for (int i = threadIdx.x; i < (int)(sizeof(JpegHuffSet) / 4); i += blockDim.x) dst[i] = src[i];
  }
  const int img = img0 + warp;
  const bool active = img < n;
  JpegDesc d{};
  if (active) {
    const uint32_t* src = (const uint32_t*)&descs[img];
    // (header fields only; the tables go to shared memory)
    uint32_t* dd = (uint32_t*)&d;
    for (int i = 0; i < 12; ++i) dd[i] = src[i];
    for (int i = lane; i < 128; i += 32) (&s_qt[warp][0][0])[i] = ((const uint16_t*)descs[img].qt)[i];
  }
  __syncthreads();
  if (!active) return;
  if (d.status != 0 || d.width != W || d.height != H) {   // (the plane scratch is sized for H x W)
    if (lane == 0) status[img] = d.status ? d.status : (uint32_t)CILRS_JPEG_SIZE_MISMATCH;
    return;
  }
  const JpegHuffSet* hs = d.huff_set == set0 ? &s_set : &sets[d.huff_set];
  const int ncomp = d.mode == 2 ? 1 : 3;
  const int hy = d.mode == 0 ? 2 : 1;                      // luma blocks per MCU side
  const int nblk = hy * hy + (ncomp - 1);                  // blocks per MCU: 6, 3 or 1
  const int y_pitch = d.mcus_x * hy * 8, c_pitch = d.mcus_x * 8;
  uint8_t* py = planes + (long long)img * plane_stride;
  uint8_t* pcb = py + (long long)y_pitch * d.mcus_y * hy * 8;
  uint8_t* pcr = pcb + (long long)c_pitch * d.mcus_y * 8;

  BitReader br;
  br.p = bytes + d.data_off + d.scan_off;
  br.end = bytes + d.data_off + d.data_len;
  br.acc = 0; br.n = 0; br.marker = false;
  int pred[3] = {0, 0, 0};
  bool bad = false;
  int16_t* coef = s_coef[warp];
  int* ws = s_ws[warp];
  const int total_mcus = d.mcus_x * d.mcus_y;
  for (int mcu = 0; mcu < total_mcus; ++mcu) {
    for (int i = lane; i < nblk * 32; i += 32) ((uint32_t*)coef)[i] = 0u;
    __syncwarp();
    if (lane == 0) {
#pragma unroll 1
      for (int b = 0; b < nblk; ++b) {
        const int ci = b < hy * hy ? 0 : b - hy * hy + 1;
        const JpegHuffTable* dct = &hs->t[d.dc_id[ci]];
        const JpegHuffTable* act = &hs->t[2 + d.ac_id[ci]];
        int16_t* blk = coef + b * 64;
        const int t = jpeg_decode_symbol(br, dct, bad);
        if (t > 11) bad = true;
        if (t) pred[ci] += jpeg_receive_extend(br, t & 15);
        blk[0] = (int16_t)pred[ci];
        int k = 1;
#pragma unroll 1
        while (k < 64) {
          const int rs = jpeg_decode_symbol(br, act, bad);
          const int r = rs >> 4, s = rs & 15;
          if (s == 0) {
            if (r != 15) break;
            k += 16;
            continue;
          }
          k += r;
          if (k > 63) { bad = true; break; }
          blk[c_zigzag[k]] = (int16_t)jpeg_receive_extend(br, s);
          ++k;
        }
        if (bad) break;
      }
    }
    __syncwarp();
    // ---- de-quantise + column pass (results scaled by 2^PASS1_BITS) ----
    for (int task = lane; task < nblk * 8; task += 32) {
      const int b = task >> 3, c = task & 7;
      const uint16_t* q = s_qt[warp][b < hy * hy ? 0 : 1];
      int x[8], o[8];
#pragma unroll
      for (int r = 0; r < 8; ++r) x[r] = (int)coef[b * 64 + r * 8 + c] * (int)q[r * 8 + c];
      idct_islow_1d<13 - 2>(x, o);
#pragma unroll
      for (int r = 0; r < 8; ++r) ws[b * 64 + r * 8 + c] = o[r];
    }
    __syncwarp();
    // ---- row pass, level shift, range limit, 8 samples = one 8-byte store ----
    const int mx = mcu % d.mcus_x, my = mcu / d.mcus_x;
    for (int task = lane; task < nblk * 8; task += 32) {
      const int b = task >> 3, r = task & 7;
      int x[8], o[8];
#pragma unroll
      for (int c = 0; c < 8; ++c) x[c] = ws[b * 64 + r * 8 + c];
      idct_islow_1d<13 + 2 + 3>(x, o);
      uint32_t lo = 0, hi = 0;
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        lo |= (uint32_t)min(max(o[c] + 128, 0), 255) << (8 * c);
        hi |= (uint32_t)min(max(o[c + 4] + 128, 0), 255) << (8 * c);
      }
      uint8_t* dst;
      if (b < hy * hy) {
        const int by = b / hy, bx = b - by * hy;
        dst = py + (long long)((my * hy + by) * 8 + r) * y_pitch + (mx * hy + bx) * 8;
      } else {
        dst = (b == hy * hy ? pcb : pcr) + (long long)(my * 8 + r) * c_pitch + mx * 8;
      }
      *(uint2*)dst = make_uint2(lo, hi);
    }
    __syncwarp();
    if (__shfl_sync(0xffffffffu, (int)bad, 0)) break;
  }
  if (lane == 0) status[img] = bad ? (uint32_t)CILRS_JPEG_CORRUPT : 0u;
}

// ---- jdsample.c h2v2_fancy_upsample + jdcolor.c ycc_rgb_convert -------------------------------------------------------------
__device__ __forceinline__ int fancy_h2v2(const uint8_t* __restrict__ pl, int pitch, int cw, int chh, int x, int y) {
  const int cy = y >> 1, cx = x >> 1;
  const int fy = (y & 1) ? min(cy + 1, chh - 1) : max(cy - 1, 0);
  const uint8_t* rn = pl + (long long)cy * pitch;
  const uint8_t* rf = pl + (long long)fy * pitch;
  const int cs = 3 * rn[cx] + rf[cx];
  if (x & 1) {
    if (cx == cw - 1) return (4 * cs + 7) >> 4;
    return (3 * cs + 3 * rn[cx + 1] + rf[cx + 1] + 7) >> 4;
  }
  if (cx == 0) return (4 * cs + 8) >> 4;
  return (3 * cs + 3 * rn[cx - 1] + rf[cx - 1] + 8) >> 4;
}
__device__ __forceinline__ uint32_t ycc_rgb(int y, int cb, int cr, int reverse) {
  const int xb = cb - 128, xr = cr - 128;
  const int r = y + ((91881 * xr + 32768) >> 16);                     // FIX(1.40200)
  const int g = y + ((-22554 * xb + 32768 - 46802 * xr) >> 16);       // FIX(0.34414), FIX(0.71414)
  const int b = y + ((116130 * xb + 32768) >> 16);                    // FIX(1.77200)
  const uint32_t R = (uint32_t)min(max(r, 0), 255), G = (uint32_t)min(max(g, 0), 255), B = (uint32_t)min(max(b, 0), 255);
  return reverse ? (B | (G << 8) | (R << 16)) : (R | (G << 8) | (B << 16));
}

__global__ void __launch_bounds__(256) jpeg_color_kernel(const JpegDesc* __restrict__ descs, const uint8_t* __restrict__ planes, long long plane_stride,
                                                         int n, int H, int W, uint8_t* __restrict__ out, int reverse, const uint32_t* __restrict__ status) {
  const int groups = (W + 3) >> 2;
  const long long total = (long long)n * H * groups;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int gx = (int)(i % groups);
    const int y = (int)((i / groups) % H);
    const int img = (int)(i / ((long long)groups * H));
    const JpegDesc& d = descs[img];
    uint32_t px[4] = {0, 0, 0, 0};
    if (status[img] == 0) {
      const int hy = d.mode == 0 ? 2 : 1;
      const int y_pitch = d.mcus_x * hy * 8, c_pitch = d.mcus_x * 8;
      const uint8_t* py = planes + (long long)img * plane_stride;
      const uint8_t* pcb = py + (long long)y_pitch * d.mcus_y * hy * 8;
      const uint8_t* pcr = pcb + (long long)c_pitch * d.mcus_y * 8;
      const int cw = (W + 1) >> 1, chh = (H + 1) >> 1;
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const int x = gx * 4 + k;
        if (x >= W) break;
        const int Y = py[(long long)y * y_pitch + x];
        if (d.mode == 2) px[k] = (uint32_t)Y * 0x010101u;
        else if (d.mode == 1) px[k] = ycc_rgb(Y, pcb[(long long)y * c_pitch + x], pcr[(long long)y * c_pitch + x], reverse);
        else px[k] = ycc_rgb(Y, fancy_h2v2(pcb, c_pitch, cw, chh, x, y), fancy_h2v2(pcr, c_pitch, cw, chh, x, y), reverse);
      }
    }
    uint8_t* o = out + (((long long)img * H + y) * W + gx * 4) * 3;
    if (gx * 4 + 4 <= W && ((W * 3) & 3) == 0) {   // 12 bytes = three aligned words
      uint32_t* o32 = (uint32_t*)o;
      o32[0] = px[0] | (px[1] << 24);
      o32[1] = (px[1] >> 8) | (px[2] << 16);
      o32[2] = (px[2] >> 16) | (px[3] << 8);
    } else {
      for (int k = 0; k < 4 && gx * 4 + k < W; ++k) {
        o[3 * k] = (uint8_t)px[k]; o[3 * k + 1] = (uint8_t)(px[k] >> 8); o[3 * k + 2] = (uint8_t)(px[k] >> 16);
      }
    }
  }
}

// ---- host side: marker parsing and derived tables ---------------------------------------------------------------------------
static void build_huff_table(const uint8_t* counts, const uint8_t* symbols, int nsym, JpegHuffTable* t) {
  memset(t, 0, sizeof(*t));
  memcpy(t->huffval, symbols, (size_t)nsym);
  int code = 0, k = 0;
  for (int l = 1; l <= 16; ++l) {
    t->valoff[l] = k - code;
    if (counts[l - 1]) {
      for (int i = 0; i < counts[l - 1]; ++i, ++code, ++k) {
        if (l <= 9) {
          const int first = code << (9 - l), cnt = 1 << (9 - l);
          for (int j = 0; j < cnt; ++j) t->look[(first + j) & 511] = (uint16_t)((l << 8) | symbols[k]);
        }
      }
      t->maxcode[l] = code - 1;
    } else {
      t->maxcode[l] = -1;
    }
    code <<= 1;
  }
  t->maxcode[17] = 0x7fffffff;
  t->maxcode[0] = -1;
}

struct DhtRaw {
  bool present = false;
  uint8_t counts[16] = {};
  uint8_t symbols[256] = {};
  int nsym = 0;
};

// parse one stream; returns the CILRS_JPEG_* verdict and fills d (except data_off / huff_set) and raw[4] (class * 2 + id)
static uint32_t parse_stream(const uint8_t* b, size_t len, JpegDesc* d, DhtRaw raw[4]) {
  if (len < 4 || b[0] != 0xFF || b[1] != 0xD8) return CILRS_JPEG_NOT_JPEG;
  uint16_t qt[4][64];
  bool have_qt[4] = {false, false, false, false};
  int comp_id[3] = {0, 0, 0}, comp_h[3] = {0, 0, 0}, comp_v[3] = {0, 0, 0}, comp_tq[3] = {0, 0, 0}, ncomp = 0;
  bool have_sof = false;
  size_t i = 2;
  while (i + 4 <= len) {
    if (b[i] != 0xFF) return CILRS_JPEG_CORRUPT;
    const int m = b[i + 1];
    if (m == 0xFF) { ++i; continue; }
    const size_t L = ((size_t)b[i + 2] << 8) | b[i + 3];
    if (L < 2 || i + 2 + L > len) return CILRS_JPEG_CORRUPT;
    const uint8_t* seg = b + i + 4;
    const size_t sl = L - 2;
    if (m == 0xDB) {
      size_t j = 0;
      while (j + 65 <= sl) {
        const int pq = seg[j] >> 4, tq = seg[j] & 15;
        if (pq != 0 || tq > 3) return CILRS_JPEG_UNSUPPORTED;
        for (int k = 0; k < 64; ++k) {
          static const uint8_t zz[64] = {0,  1,  8,  16, 9,  2,  3,  10, 17, 24, 32, 25, 18, 11, 4,  5,  12, 19, 26, 33, 40, 48,
                                         41, 34, 27, 20, 13, 6,  7,  14, 21, 28, 35, 42, 49, 56, 57, 50, 43, 36, 29, 22, 15, 23,
                                         30, 37, 44, 51, 58, 59, 52, 45, 38, 31, 39, 46, 53, 60, 61, 54, 47, 55, 62, 63};
          qt[tq][zz[k]] = seg[j + 1 + k];
        }
        have_qt[tq] = true;
        j += 65;
      }
    } else if (m == 0xC0 || m == 0xC1) {
      if (sl < 6 || seg[0] != 8) return CILRS_JPEG_UNSUPPORTED;
      d->height = (uint16_t)((seg[1] << 8) | seg[2]);
      d->width = (uint16_t)((seg[3] << 8) | seg[4]);
      ncomp = seg[5];
      if ((ncomp != 1 && ncomp != 3) || sl < (size_t)(6 + 3 * ncomp)) return CILRS_JPEG_UNSUPPORTED;
      for (int k = 0; k < ncomp; ++k) {
        comp_id[k] = seg[6 + 3 * k]; comp_h[k] = seg[7 + 3 * k] >> 4; comp_v[k] = seg[7 + 3 * k] & 15; comp_tq[k] = seg[8 + 3 * k];
        if (comp_tq[k] > 3) return CILRS_JPEG_UNSUPPORTED;
      }
      have_sof = true;
    } else if ((m >= 0xC2 && m <= 0xCF) && m != 0xC4 && m != 0xC8 && m != 0xCC) {
      return CILRS_JPEG_UNSUPPORTED;   // progressive, lossless, arithmetic coding
    } else if (m == 0xC4) {
      size_t j = 0;
      while (j + 17 <= sl) {
        const int tc = seg[j] >> 4, th = seg[j] & 15;
        if (tc > 1 || th > 1) return CILRS_JPEG_UNSUPPORTED;
        DhtRaw& r = raw[tc * 2 + th];
        int ns = 0;
        for (int k = 0; k < 16; ++k) { r.counts[k] = seg[j + 1 + k]; ns += r.counts[k]; }
        if (ns > 256 || j + 17 + ns > sl) return CILRS_JPEG_CORRUPT;
        memcpy(r.symbols, seg + j + 17, (size_t)ns);
        r.nsym = ns;
        r.present = true;
        j += 17 + (size_t)ns;
      }
    } else if (m == 0xDD) {
      if (sl >= 2 && ((seg[0] << 8) | seg[1]) != 0) return CILRS_JPEG_UNSUPPORTED;   // restart intervals: the collector writes none
    } else if (m == 0xDA) {
      if (!have_sof || sl < 1 || seg[0] != ncomp || sl < (size_t)(1 + 2 * ncomp)) return CILRS_JPEG_UNSUPPORTED;
      for (int k = 0; k < ncomp; ++k) {
        int ci = -1;
        for (int c = 0; c < ncomp; ++c) if (comp_id[c] == seg[1 + 2 * k]) ci = c;
        if (ci != k) return CILRS_JPEG_UNSUPPORTED;   // interleaved scan in component order
        d->dc_id[k] = seg[2 + 2 * k] >> 4;
        d->ac_id[k] = seg[2 + 2 * k] & 15;
        if (d->dc_id[k] > 1 || d->ac_id[k] > 1 || !raw[d->dc_id[k]].present || !raw[2 + d->ac_id[k]].present) return CILRS_JPEG_UNSUPPORTED;
      }
      if (ncomp == 1) {
        d->mode = 2;
      } else {
        if (comp_h[1] != 1 || comp_v[1] != 1 || comp_h[2] != 1 || comp_v[2] != 1 || comp_tq[1] != comp_tq[2]) return CILRS_JPEG_UNSUPPORTED;
        if (comp_h[0] == 2 && comp_v[0] == 2) d->mode = 0;
        else if (comp_h[0] == 1 && comp_v[0] == 1) d->mode = 1;
        else return CILRS_JPEG_UNSUPPORTED;
      }
      if (!have_qt[comp_tq[0]] || (ncomp == 3 && !have_qt[comp_tq[1]])) return CILRS_JPEG_CORRUPT;
      memcpy(d->qt[0], qt[comp_tq[0]], 128);
      memcpy(d->qt[1], qt[ncomp == 3 ? comp_tq[1] : comp_tq[0]], 128);
      const int mcu = d->mode == 0 ? 16 : 8;
      d->mcus_x = (uint16_t)((d->width + mcu - 1) / mcu);
      d->mcus_y = (uint16_t)((d->height + mcu - 1) / mcu);
      d->scan_off = (uint32_t)(i + 2 + L);
      if (d->width == 0 || d->height == 0) return CILRS_JPEG_CORRUPT;
      return 0;
    }
    i += 2 + L;
  }
  return CILRS_JPEG_CORRUPT;
}

}  // namespace cilrs

using namespace cilrs;

extern "C" {

size_t cilrs_jpeg_desc_bytes(void) { return sizeof(JpegDesc); }
size_t cilrs_jpeg_table_set_bytes(void) { return sizeof(JpegHuffSet); }
// bytes of component-plane scratch one image of height x width needs (worst case over the supported samplings)
size_t cilrs_jpeg_plane_bytes(int height, int width) {
  if (height < 1 || width < 1) return 0;
  const size_t y420 = (size_t)((height + 15) / 16 * 16) * ((width + 15) / 16 * 16);
  const size_t y444 = (size_t)((height + 7) / 8 * 8) * ((width + 7) / 8 * 8);
  const size_t a = y420 + y420 / 2, b = 3 * y444;
  return ((a > b ? a : b) + 255) & ~(size_t)255;
}

int cilrs_jpeg_prepare(const unsigned char* bytes, const long long* offsets, int n, void* descs_out, void* sets_out, int max_sets,
                       int* n_sets_out) {
  if (!bytes || !offsets || n < 0 || !descs_out || !sets_out || max_sets < 1 || !n_sets_out) return ERR_INVALID;
  JpegDesc* descs = (JpegDesc*)descs_out;
  JpegHuffSet* sets = (JpegHuffSet*)sets_out;
  std::vector<std::string> keys;   // raw DHT bytes of each distinct table set
  for (int i = 0; i < n; ++i) {
    JpegDesc& d = descs[i];
    memset(&d, 0, sizeof(d));
    const long long lo = offsets[i], hi = offsets[i + 1];
    if (lo < 0 || hi < lo || (lo & 3)) return ERR_INVALID;   // streams start 4-byte aligned inside the batch buffer
    d.data_off = (uint32_t)lo;
    d.data_len = (uint32_t)(hi - lo);
    DhtRaw raw[4];
    d.status = parse_stream(bytes + lo, (size_t)(hi - lo), &d, raw);
    if (d.status) continue;
    std::string key;
    for (int t = 0; t < 4; ++t) {
      key.push_back((char)raw[t].present);
      key.append((const char*)raw[t].counts, 16);
      key.append((const char*)raw[t].symbols, (size_t)raw[t].nsym);
    }
    int idx = -1;
    for (size_t k = 0; k < keys.size(); ++k) if (keys[k] == key) idx = (int)k;
    if (idx < 0) {
      if ((int)keys.size() >= max_sets) { d.status = CILRS_JPEG_UNSUPPORTED; continue; }
      idx = (int)keys.size();
      keys.push_back(key);
      for (int t = 0; t < 4; ++t) {
        if (raw[t].present) build_huff_table(raw[t].counts, raw[t].symbols, raw[t].nsym, &sets[idx].t[t]);
        else memset(&sets[idx].t[t], 0, sizeof(JpegHuffTable));
      }
    }
    d.huff_set = (uint8_t)idx;
  }
  *n_sets_out = (int)keys.size();
  return OK;
}

// read n files into `dst` (streams 16-byte aligned, offsets[n + 1]) with `threads` host threads; returns ERR_WORKSPACE when
// `capacity` is too small, ERR_INVALID when a file cannot be read (offsets[i + 1] == offsets[i] marks it)
int cilrs_jpeg_read_files(const char* const* paths, int n, unsigned char* dst, long long capacity, long long* offsets, int threads) {
  if (!paths || n < 0 || !dst || !offsets || capacity < 0) return ERR_INVALID;
  std::vector<long long> sizes((size_t)n, -1);
  auto stat_range = [&](int lo, int hi) {
    for (int i = lo; i < hi; ++i) {
      FILE* f = fopen(paths[i], "rb");
      if (!f) continue;
      if (fseek(f, 0, SEEK_END) == 0) sizes[(size_t)i] = ftell(f);
      fclose(f);
    }
  };
  if (threads < 1) threads = 1;
  if (threads > n) threads = n > 0 ? n : 1;
  auto run = [&](auto fn) {
    if (threads == 1) { fn(0, n); return; }
    std::vector<std::thread> pool;
    for (int t = 0; t < threads; ++t) pool.emplace_back(fn, (int)((long long)n * t / threads), (int)((long long)n * (t + 1) / threads));
    for (auto& th : pool) th.join();
  };
  run(stat_range);
  long long off = 0;
  bool missing = false;
  for (int i = 0; i < n; ++i) {
    offsets[i] = off;
    if (sizes[(size_t)i] < 0) { missing = true; continue; }
    off += (sizes[(size_t)i] + 15) & ~15LL;
  }
  offsets[n] = off;
  if (off > capacity) return ERR_WORKSPACE;
  std::vector<long long> ends((size_t)n + 1);
  auto read_range = [&](int lo, int hi) {
    for (int i = lo; i < hi; ++i) {
      if (sizes[(size_t)i] < 0) continue;
      FILE* f = fopen(paths[i], "rb");
      if (!f) { sizes[(size_t)i] = -1; continue; }
      const size_t got = fread(dst + offsets[i], 1, (size_t)sizes[(size_t)i], f);
      fclose(f);
      if ((long long)got != sizes[(size_t)i]) sizes[(size_t)i] = -1;
    }
  };
  run(read_range);
  // callers pass (offsets[i], offsets[i] + size): report exact ends through a second array layout [n + 1 .. 2n]
  for (int i = 0; i < n; ++i) {
    if (sizes[(size_t)i] < 0) { missing = true; offsets[n + 1 + i] = offsets[i]; }
    else offsets[n + 1 + i] = offsets[i] + sizes[(size_t)i];
  }
  return missing ? ERR_INVALID : OK;
}

int cilrs_jpeg_decode(const unsigned char* bytes_dev, const void* descs_dev, const void* sets_dev, int n, int height, int width,
                      unsigned char* planes_dev, size_t plane_bytes_per_image, unsigned char* out_rgb, int reverse,
                      unsigned int* status_dev, void* stream) {
  if (!bytes_dev || !descs_dev || !sets_dev || n < 0 || height < 1 || width < 1 || !planes_dev || !out_rgb || !status_dev) return ERR_INVALID;
  if (plane_bytes_per_image < cilrs_jpeg_plane_bytes(height, width)) return ERR_WORKSPACE;
  if (n == 0) return OK;
  cudaStream_t s = (cudaStream_t)stream;
  ++g_cilrs_launches;
  jpeg_entropy_idct_kernel<<<(n + JPEG_WARPS - 1) / JPEG_WARPS, JPEG_WARPS * 32, 0, s>>>(bytes_dev, (const JpegDesc*)descs_dev, (const JpegHuffSet*)sets_dev,
                                                                                      n, height, width, planes_dev, (long long)plane_bytes_per_image, status_dev);
  int st = cuda_status(cudaGetLastError());
  if (st) return st;
  const long long groups = (long long)n * height * ((width + 3) / 4);
  long long grid = (groups + 255) / 256;
  if (grid > 148LL * 16) grid = 148LL * 16;
  ++g_cilrs_launches;
  jpeg_color_kernel<<<(int)grid, 256, 0, s>>>((const JpegDesc*)descs_dev, planes_dev, (long long)plane_bytes_per_image, n, height, width, out_rgb,
                                              reverse, status_dev);
  return cuda_status(cudaGetLastError());
}

}  // extern "C"
