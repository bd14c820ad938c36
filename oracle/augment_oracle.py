"""TEST INFRASTRUCTURE - CPU restatement of the reference's augmentation arithmetic (notebook/notebook.ipynb:387-394).

albumentations (absent from /root/reference and from this image; the notebook pins no version, its call signatures are the 2.x
API) implements the five transforms on uint8 images with numpy and OpenCV calls; this file makes the same calls for GIVEN
parameters: a 256-entry look-up table for brightness/contrast, cv2.cvtColor RGB<->HSV around integer channel shifts,
cv2.GaussianBlur(ksize, sigma 0), additive Gaussian noise (the noise field is an input here), zero-filled rectangles.
"parity unpinned" against albumentations itself (it cannot be run here); pinned against OpenCV, which does the arithmetic.
Only tests/ may import this module.

HSV -> RGB on 8-bit images is the one call whose OpenCV result depends on the CPU it runs on: the library's scalar code
(`saturate_cast<uchar>(x * 255.f)`: round to nearest; used for rows narrower than one vector and for every row's tail) and its
AVX2 vector code (truncation: measured on this image's cv2 4.13 - all 180 x 256 x 256 inputs equal floor(x * 255), a third of
the components one below the scalar result) disagree by one level. `apply(..., vector_path=False)` (the default) evaluates
cv2.cvtColor through its SCALAR path (one-pixel-wide rows), which is the same on every host and is what the CUDA kernel
restates bit for bit; `vector_path=True` is the live wide-row call, for the "within one level" check.
"""
import cv2
import numpy as np


def apply(img, p, noise=None, vector_path=False):
    """img uint8 [H, W, 3] RGB, p one record of cilrs_b200.augment.PARAM_DTYPE, noise float32 [H, W, 3] standard normal."""
    img = img.copy()
    flags = int(p["flags"])
    if flags & 1:
        lut = np.arange(256, dtype=np.float32) * np.float32(p["alpha"]) + np.float32(p["beta"]) * np.float32(255.0)
        img = cv2.LUT(img, np.clip(lut, 0, 255).astype(np.uint8))
    if flags & 2:
        hsv = cv2.cvtColor(img, cv2.COLOR_RGB2HSV).astype(np.int32)
        hsv[..., 0] = np.mod(hsv[..., 0] + int(p["hue"]), 180)
        hsv[..., 1] = np.clip(hsv[..., 1] + int(p["sat"]), 0, 255)
        hsv[..., 2] = np.clip(hsv[..., 2] + int(p["val"]), 0, 255)
        hsv = hsv.astype(np.uint8)
        if vector_path:
            img = cv2.cvtColor(hsv, cv2.COLOR_HSV2RGB)
        else:
            img = cv2.cvtColor(np.ascontiguousarray(hsv.reshape(-1, 1, 3)), cv2.COLOR_HSV2RGB).reshape(hsv.shape)
    if flags & 4 and int(p["ksize"]) in (3, 5):
        k = int(p["ksize"])
        img = cv2.GaussianBlur(img, (k, k), 0)
    if flags & 8:
        if noise is None:
            raise ValueError("noise field needed")
        img = np.clip(img.astype(np.float32) + noise.astype(np.float32) * np.float32(p["noise_std"]), 0, 255).astype(np.uint8)
    if flags & 16:
        for h in range(min(int(p["n_holes"]), 3)):
            y0, y1, x0, x1 = [int(v) for v in p["hole"][h]]
            img[max(y0, 0):max(y1, 0), max(x0, 0):max(x1, 0)] = 0
    return img


def hsv2rgb_scalar_np(hsv):
    """OpenCV's scalar 8-bit HSV -> RGB (imgproc color_hsv: HSV2RGB_native + saturate_cast<uchar>(x * 255.f)) restated in numpy with
    the rounding order the library's x86 build executes: every product rounded to fp32, the inner `1 - s * h` and `1 - s * (1 - h)`
    as ONE fused multiply-add (emulated in fp64: the product of two fp32 values is exact there), round-to-nearest-even at the end.
    This is the arithmetic `augment_kernel` (csrc/pipeline.cu: hsv2rgb_u8) implements; tests/test_augment_cpu.py pins it against
    cv2.cvtColor's scalar path for all 180 x 256 x 256 inputs."""
    f32 = np.float32
    h, s, v = hsv[..., 0], hsv[..., 1], hsv[..., 2]
    sf = s.astype(f32) * f32(1.0 / 255.0)
    vf = v.astype(f32) * f32(1.0 / 255.0)
    hh = h.astype(f32) * f32(6.0 / 180.0)
    sector = hh.astype(np.int32)
    hf = hh - sector.astype(f32)

    def fnma1(a, b):   # fp32(1 - a * b) with a single rounding
        return (1.0 - a.astype(np.float64) * b.astype(np.float64)).astype(f32)
    t = [vf, vf * (f32(1) - sf), vf * fnma1(sf, hf), vf * fnma1(sf, f32(1) - hf)]
    table = ((1, 3, 0), (1, 0, 2), (3, 0, 1), (0, 2, 1), (0, 1, 3), (2, 1, 0))   # OpenCV's sector_data: (b, g, r)
    out = np.zeros(hsv.shape[:-1] + (3,), dtype=f32)
    for k, (ib, ig, ir) in enumerate(table):
        m = sector == k
        out[..., 0][m] = t[ir][m]; out[..., 1][m] = t[ig][m]; out[..., 2][m] = t[ib][m]
    grey = s == 0
    for c in range(3):
        out[..., c][grey] = vf[grey]
    return np.clip(np.rint(out * f32(255)), 0, 255).astype(np.uint8)
