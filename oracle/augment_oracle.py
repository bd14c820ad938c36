"""TEST INFRASTRUCTURE - CPU restatement of the reference's augmentation arithmetic (notebook/notebook.ipynb:387-394).

albumentations (absent from /root/reference and from this image; the notebook pins no version, its call signatures are the 2.x
API) implements the five transforms on uint8 images with numpy and OpenCV calls; this file makes the same calls for GIVEN
parameters: a 256-entry look-up table for brightness/contrast, cv2.cvtColor RGB<->HSV around integer channel shifts,
cv2.GaussianBlur(ksize, sigma 0), additive Gaussian noise (the noise field is an input here), zero-filled rectangles.
"parity unpinned" against albumentations itself (it cannot be run here); pinned against OpenCV, which does the arithmetic.
Only tests/ may import this module.
"""
import cv2
import numpy as np


def apply(img, p, noise=None):
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
        img = cv2.cvtColor(hsv.astype(np.uint8), cv2.COLOR_HSV2RGB)
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
