"""Writes tests/golden/jpeg/*.jpg (frames encoded the way the collector does, model/collect_data.py:697-701: cv2.imwrite /
imencode at JPEG quality 95) and jpeg_golden.npz = what the reference's loader reads from them
(cv2.cvtColor(cv2.imread(path), COLOR_BGR2RGB), notebook/notebook.ipynb:404-405). Run in this container: python tests/golden/make_jpeg_golden.py"""
import os

import cv2
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "jpeg")


def synthetic_frame(seed, h=88, w=200):
    """A road-scene-like frame: sky gradient, textured road, lane markings, a few saturated objects, sensor noise."""
    r = np.random.default_rng(seed)
    y, x = np.mgrid[0:h, 0:w].astype(np.float64)
    img = np.zeros((h, w, 3))
    horizon = h * (0.35 + 0.1 * r.random())
    sky = np.clip((horizon - y) / horizon, 0, 1)[..., None] * np.array([90, 140, 230]) + 25
    road = np.clip((y - horizon) / (h - horizon), 0, 1)[..., None] * np.array([70, 70, 75]) + 40
    img = np.where((y < horizon)[..., None], sky, road)
    img += 12 * np.sin(x / (5 + seed % 7) + y / 3.0)[..., None]
    lane = np.abs(x - w / 2 - (y - horizon) * (0.8 * r.random() - 0.4)) < 1.5 + (y - horizon) / 25
    img[lane & (y > horizon)] = [235, 235, 210]
    for _ in range(3):
        y0, x0 = int(r.integers(0, h - 12)), int(r.integers(0, w - 24))
        img[y0:y0 + int(r.integers(4, 12)), x0:x0 + int(r.integers(6, 24))] = r.integers(0, 256, 3)
    img += r.normal(0, 2 + 2 * (seed % 4), img.shape)
    return np.clip(img, 0, 255).astype(np.uint8)   # RGB


def main():
    os.makedirs(OUT, exist_ok=True)
    golden = {}
    cases = [("frame_%08d" % i, synthetic_frame(i), [cv2.IMWRITE_JPEG_QUALITY, 95]) for i in range(6)]
    cases.append(("odd_87x199", synthetic_frame(7, 87, 199), [cv2.IMWRITE_JPEG_QUALITY, 95]))
    cases.append(("q60_444", synthetic_frame(8), [cv2.IMWRITE_JPEG_QUALITY, 60, cv2.IMWRITE_JPEG_SAMPLING_FACTOR, cv2.IMWRITE_JPEG_SAMPLING_FACTOR_444]))
    for name, rgb, params in cases:
        path = os.path.join(OUT, name + ".jpg")
        cv2.imwrite(path, cv2.cvtColor(rgb, cv2.COLOR_RGB2BGR), params)
        golden[name] = cv2.cvtColor(cv2.imread(path), cv2.COLOR_BGR2RGB)
    grey = synthetic_frame(9).mean(axis=2).astype(np.uint8)
    cv2.imwrite(os.path.join(OUT, "grey.jpg"), grey, [cv2.IMWRITE_JPEG_QUALITY, 95])
    golden["grey"] = cv2.cvtColor(cv2.imread(os.path.join(OUT, "grey.jpg")), cv2.COLOR_BGR2RGB)
    np.savez_compressed(os.path.join(HERE, "jpeg_golden.npz"), **golden)
    print({k: v.shape for k, v in golden.items()}, "opencv", cv2.__version__)


if __name__ == "__main__":
    main()
