/* ORACLE — TEST INFRASTRUCTURE ONLY. Not part of the product path; only tests/, __graft_entry__.smoke() and
 * bench.py's cpu_baseline / --impl reference legs may load this.
 *
 * Plain-C restatement of the reference's frame preprocessing
 *     img = cv2.resize(image, (IMG_WIDTH, IMG_HEIGHT))            model/autonomous_drive.py:898, model/prepare_dataset.py:56
 *     img = img.astype(np.float32) / 255.0; permute(2,0,1)        model/autonomous_drive.py:899-900
 *     img = Normalize(mean, std)(img)                             model/autonomous_drive.py:901, :481-482, :502-504
 * The arithmetic lives in un-vendored third-party code: opencv-python==4.6.0.66 (requirements.txt:3; 4.13.0 is what
 * this image has) — cv::resize, INTER_LINEAR, CV_8U: the fixed-point path (INTER_RESIZE_COEF_BITS = 11) of
 * modules/imgproc/src/resize.cpp, restated here from its published algorithm:
 *     fx = (float)((dx + 0.5) * scale_x - 0.5); sx = floor(fx); fx -= sx; clamp (sx<0 -> 0,fx=0; sx>=w-1 -> w-1,fx=0)
 *     ialpha = { saturate_cast<short>((1-fx)*2048), saturate_cast<short>(fx*2048) }   (same per row with fy/ibeta)
 *     H[dx]  = S[sx]*ialpha0 + S[sx+1]*ialpha1                                        (int32, scaled 2^11)
 *     dst    = (((ibeta0 * (H0 >> 4)) >> 16) + ((ibeta1 * (H1 >> 4)) >> 16) + 2) >> 2
 * Pinned against cv2.resize itself (tests/test_oracle_cpu.py, bit-exact) and against committed golden frames
 * produced by the reference call (tests/golden/make_golden.py).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>

static void axis_coef(int d, int ssize, double scale, int* s0, int* s1, int* a0, int* a1) {
  float f = (float)(((double)d + 0.5) * scale - 0.5);
  int s = (int)floorf(f);
  f -= (float)s;
  if (s < 0) { s = 0; f = 0.f; }
  if (s >= ssize - 1) { s = ssize - 1; f = 0.f; }
  *s0 = s;
  *s1 = s + 1 < ssize ? s + 1 : ssize - 1;
  *a0 = (int)lrintf((1.f - f) * 2048.f);
  *a1 = (int)lrintf(f * 2048.f);
}

/* src: uint8 [batch, sh, sw, sc] (sc = 3 or 4), dst_u8: uint8 [batch, dh, dw, 3] or NULL, dst_f32: float [batch,3,dh,dw] or NULL.
 * reverse != 0 flips the first three channels (BGR -> RGB, model/autonomous_drive.py:1551). */
int oracle_preprocess(const uint8_t* src, int batch, int sh, int sw, int sc, int reverse, int dh, int dw, uint8_t* dst_u8,
                      float* dst_f32) {
  static const float mean[3] = {0.485f, 0.456f, 0.406f};
  static const float stdv[3] = {0.229f, 0.224f, 0.225f};
  const double scale_x = (double)sw / dw, scale_y = (double)sh / dh;
  int* xs = (int*)malloc(sizeof(int) * 4 * (size_t)dw);
  if (!xs) return 1;
  for (int dx = 0; dx < dw; ++dx) axis_coef(dx, sw, scale_x, &xs[4 * dx], &xs[4 * dx + 1], &xs[4 * dx + 2], &xs[4 * dx + 3]);
  for (int n = 0; n < batch; ++n) {
    const uint8_t* img = src + (size_t)n * sh * sw * sc;
    for (int dy = 0; dy < dh; ++dy) {
      int y0, y1, b0, b1;
      axis_coef(dy, sh, scale_y, &y0, &y1, &b0, &b1);
      const uint8_t* r0 = img + (size_t)y0 * sw * sc;
      const uint8_t* r1 = img + (size_t)y1 * sw * sc;
      for (int dx = 0; dx < dw; ++dx) {
        const int x0 = xs[4 * dx], x1 = xs[4 * dx + 1], a0 = xs[4 * dx + 2], a1 = xs[4 * dx + 3];
        for (int c = 0; c < 3; ++c) {
          const int cs = reverse ? 2 - c : c;
          const int h0 = r0[x0 * sc + cs] * a0 + r0[x1 * sc + cs] * a1;
          const int h1 = r1[x0 * sc + cs] * a0 + r1[x1 * sc + cs] * a1;
          const int v = (((b0 * (h0 >> 4)) >> 16) + ((b1 * (h1 >> 4)) >> 16) + 2) >> 2;
          if (dst_u8) dst_u8[(((size_t)n * dh + dy) * dw + dx) * 3 + c] = (uint8_t)v;
          if (dst_f32) {
            float f = (float)v / 255.0f;
            f = (f - mean[c]) / stdv[c];
            dst_f32[(((size_t)n * 3 + c) * dh + dy) * dw + dx] = f;
          }
        }
      }
    }
  }
  free(xs);
  return 0;
}
