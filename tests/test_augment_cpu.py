"""N4 on the CPU: the augmentation oracle's HSV arithmetic pinned against OpenCV (no GPU). The CUDA kernel is checked against
the same oracle in tests/test_pipeline_gpu.py."""
import numpy as np
import pytest

cv2 = pytest.importorskip("cv2")
from oracle import augment_oracle as AO   # noqa: E402


def _all_hsv():
    h, s, v = np.meshgrid(np.arange(180), np.arange(256), np.arange(256), indexing="ij")
    return np.ascontiguousarray(np.stack([h, s, v], -1).astype(np.uint8).reshape(-1, 1, 3))


def test_hsv2rgb_restatement_equals_opencv_scalar_path_for_every_input():
    hsv = _all_hsv()                                              # one-pixel rows: OpenCV's scalar code on every host
    want = cv2.cvtColor(hsv, cv2.COLOR_HSV2RGB)
    got = AO.hsv2rgb_scalar_np(hsv)
    assert int((got != want).sum()) == 0


def test_opencv_vector_path_is_within_one_level_of_its_scalar_path():
    """Why the bit-for-bit pin is on the scalar path: a wide row goes through OpenCV's vector code, which on AVX2 hosts truncates
    where the scalar code rounds. Whatever this host's cv2 does, it stays within one level (and never above the scalar result)."""
    hsv = _all_hsv()
    scalar = cv2.cvtColor(hsv, cv2.COLOR_HSV2RGB).reshape(-1, 256, 3).astype(np.int32)
    wide = cv2.cvtColor(np.ascontiguousarray(hsv.reshape(-1, 256, 3)), cv2.COLOR_HSV2RGB).astype(np.int32)
    d = scalar - wide
    assert d.min() >= -1 and d.max() <= 1


def test_rgb2hsv_is_host_independent():
    rgb = np.random.default_rng(0).integers(0, 256, (2048, 256, 3), dtype=np.uint8)
    a = cv2.cvtColor(rgb, cv2.COLOR_RGB2HSV)
    b = cv2.cvtColor(np.ascontiguousarray(rgb.reshape(-1, 1, 3)), cv2.COLOR_RGB2HSV).reshape(a.shape)
    assert np.array_equal(a, b)


def test_apply_paths_agree_without_the_hsv_transform():
    import os
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden"))
    from make_jpeg_golden import synthetic_frame
    img = synthetic_frame(11, 88, 200)
    p = np.zeros(1, dtype=[("flags", "<u4"), ("alpha", "<f4"), ("beta", "<f4"), ("hue", "<i2"), ("sat", "<i2"), ("val", "<i2"),
                           ("ksize", "<i2"), ("noise_std", "<f4"), ("n_holes", "<u4"), ("hole", "<i2", (3, 4)), ("pad", "<u4", (3,))])[0]
    p["flags"], p["alpha"], p["beta"], p["ksize"], p["n_holes"] = 1 | 4 | 16, 1.1, 0.05, 5, 1
    p["hole"][0] = [10, 18, 30, 45]
    assert np.array_equal(AO.apply(img, p), AO.apply(img, p, vector_path=True))
    p["flags"] |= 2
    p["hue"], p["sat"], p["val"] = 7, -12, 9
    a, b = AO.apply(img, p).astype(np.int32), AO.apply(img, p, vector_path=True).astype(np.int32)
    assert np.abs(a - b).max() <= 1
