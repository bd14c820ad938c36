"""N4 - the reference's training augmentations (notebook/notebook.ipynb:387-394) on the device.

    train_augmentation = A.Compose([
        A.RandomBrightnessContrast(brightness_limit=0.2, contrast_limit=0.2, p=0.5),
        A.HueSaturationValue(hue_shift_limit=10, sat_shift_limit=20, val_shift_limit=15, p=0.3),
        A.GaussianBlur(blur_limit=(3, 5), p=0.2),
        A.GaussNoise(std_range=(0.02, 0.06), p=0.3),
        A.CoarseDropout(num_holes_range=(1, 3), hole_height_range=(4, 10), hole_width_range=(8, 20), fill=0, p=0.2)])

`albumentations` is a third-party dependency that is absent from /root/reference and from this image. What is restated is its
documented behaviour on uint8 images: which transform fires with which probability, the ranges its parameters are drawn from, and
the arithmetic each one applies (uint8 look-up table for brightness/contrast, `cv2.cvtColor` RGB<->HSV with shifted channels,
`cv2.GaussianBlur` with sigma 0, additive Gaussian noise of standard deviation std * 255, zero-filled rectangles). The parameter
draws (a dozen scalars per frame) happen here on the host with a numpy Generator; `cilrs_augment_u8` applies them to the whole
batch in one launch. Parity: the deterministic transforms are bit-identical to the cv2 / numpy calls with the same parameters
(tests/test_pipeline_gpu.py against oracle/augment_oracle.py); the random draws follow the same distributions but not
albumentations' random stream, and the Gaussian-blur sigma follows OpenCV's ksize rule (sigma 0) - parity for the draws is
statistical ("parity unpinned" for N4's random stream, as SURVEY.md section 8 row N4 anticipates).
"""
import ctypes

import numpy as np
import torch

from . import _lib

PARAM_DTYPE = np.dtype([("flags", "<u4"), ("alpha", "<f4"), ("beta", "<f4"), ("hue", "<i2"), ("sat", "<i2"), ("val", "<i2"),
                        ("ksize", "<i2"), ("noise_std", "<f4"), ("n_holes", "<u4"), ("hole", "<i2", (3, 4)), ("pad", "<u4", (3,))])
assert PARAM_DTYPE.itemsize == 64

F_BRIGHTNESS_CONTRAST, F_HSV, F_BLUR, F_NOISE, F_DROPOUT = 1, 2, 4, 8, 16


def draw_params(n, height, width, rng, p_scale=1.0):
    """Per-frame parameters of the reference's Compose (probabilities and ranges of notebook/notebook.ipynb:387-394)."""
    out = np.zeros(n, dtype=PARAM_DTYPE)
    u = rng.random((n, 5))
    fire = u < np.array([0.5, 0.3, 0.2, 0.3, 0.2]) * p_scale
    out["flags"] = (fire * np.array([F_BRIGHTNESS_CONTRAST, F_HSV, F_BLUR, F_NOISE, F_DROPOUT])).sum(axis=1).astype(np.uint32)
    out["alpha"] = 1.0 + rng.uniform(-0.2, 0.2, n)          # contrast_limit
    out["beta"] = rng.uniform(-0.2, 0.2, n)                 # brightness_limit (brightness_by_max: times 255)
    out["hue"] = np.rint(rng.uniform(-10, 10, n)).astype(np.int16)
    out["sat"] = np.rint(rng.uniform(-20, 20, n)).astype(np.int16)
    out["val"] = np.rint(rng.uniform(-15, 15, n)).astype(np.int16)
    out["ksize"] = rng.choice(np.array([3, 5], dtype=np.int16), n)
    out["noise_std"] = rng.uniform(0.02, 0.06, n) * 255.0
    holes = rng.integers(1, 4, n)
    out["n_holes"] = holes
    hh = rng.integers(4, 11, (n, 3))
    hw = rng.integers(8, 21, (n, 3))
    y0 = (rng.random((n, 3)) * (height - hh + 1)).astype(np.int64)
    x0 = (rng.random((n, 3)) * (width - hw + 1)).astype(np.int64)
    out["hole"] = np.stack([y0, y0 + hh, x0, x0 + hw], axis=-1).astype(np.int16)
    return out


class DeviceAugmenter:
    """augmenter(frames_u8 [n, H, W, 3] on the device) -> augmented frames (same tensor unless out= is given)."""

    def __init__(self, max_batch, height=88, width=200, seed=0, device=None, p_scale=1.0):
        if _lib.lib().cilrs_augment_param_bytes() != PARAM_DTYPE.itemsize:
            raise RuntimeError("cilrs_b200.augment: parameter layout mismatch with the library")
        self.dev = torch.device(device if device is not None else "cuda")
        self.h, self.w, self.max_batch = height, width, max_batch
        self.rng = np.random.default_rng(seed)
        self.seed = int(seed) & 0xFFFFFFFFFFFFFFFF
        self.p_scale = p_scale
        self._offset = 0
        self._host = [torch.zeros(max_batch * 64, dtype=torch.uint8).pin_memory() for _ in range(2)]
        self._events = [None, None]
        self._dev = torch.zeros(max_batch * 64, dtype=torch.uint8, device=self.dev)
        self._slot = 0
        self.last_params = None

    def __call__(self, frames, params=None, out=None):
        n = frames.shape[0]
        if frames.dtype != torch.uint8 or tuple(frames.shape[1:]) != (self.h, self.w, 3) or not frames.is_cuda or not frames.is_contiguous():
            raise ValueError("DeviceAugmenter: frames must be a contiguous CUDA uint8 tensor [n, %d, %d, 3]" % (self.h, self.w))
        if n > self.max_batch:
            raise ValueError("DeviceAugmenter: batch larger than max_batch")
        if params is None:
            params = draw_params(n, self.h, self.w, self.rng, self.p_scale)
        self.last_params = params
        slot = self._slot
        self._slot ^= 1
        if self._events[slot] is not None:
            self._events[slot].synchronize()    # the pinned staging buffer is free again
        host = self._host[slot]
        host[:n * 64].numpy()[:] = np.frombuffer(params.tobytes(), dtype=np.uint8)
        self._dev[:n * 64].copy_(host[:n * 64], non_blocking=True)
        ev = torch.cuda.Event()
        ev.record()
        self._events[slot] = ev
        if out is None:
            out = frames
        _lib.call("cilrs_augment_u8", frames, out, self._dev, n, self.h, self.w, ctypes.c_ulonglong(self.seed),
                  ctypes.c_ulonglong(self._offset), _lib.stream_ptr())
        self._offset += n * ((self.h * self.w * 3 + 3) // 4)
        return out
