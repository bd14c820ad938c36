"""N3 / N4 on the GPU, through the C-ABI: JPEG decode bit-exact against the reference loader's decode (cv2.imread + BGR2RGB,
notebook/notebook.ipynb:404-405: committed golden arrays, the numpy oracle, and OpenCV live), the weighted sampler's distribution,
the augmentations against the cv2 / numpy calls albumentations makes, and the whole loader feeding a training step."""
import glob
import os

import numpy as np
import pytest
import torch

from oracle import augment_oracle as AO
from oracle import jpeg_oracle as J

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))
JPEGS = sorted(glob.glob(os.path.join(HERE, "golden", "jpeg", "*.jpg")))
GOLDEN = np.load(os.path.join(HERE, "golden", "jpeg_golden.npz"))


def _name(p):
    return os.path.splitext(os.path.basename(p))[0]


def _frames(n, seed=0, h=88, w=200):
    import sys
    sys.path.insert(0, os.path.join(HERE, "golden"))
    from make_jpeg_golden import synthetic_frame
    return [synthetic_frame(seed * 1000 + i, h, w) for i in range(n)]


def test_decode_equals_the_reference_loader_golden():
    from cilrs_b200.data import JpegDecoder
    names = [p for p in JPEGS if GOLDEN[_name(p)].shape == (88, 200, 3)]
    dec = JpegDecoder(16)
    out = dec.decode([open(p, "rb").read() for p in names])
    dec.check()
    out = out.cpu().numpy()
    for i, p in enumerate(names):
        assert np.array_equal(out[i], GOLDEN[_name(p)]), _name(p)   # 4:2:0 frames, the 4:4:4 one and the grey one, bit for bit
    # BGR order (what cv2.imread itself returns) and decode_files
    out2 = dec.decode_files(names, reverse=True).cpu().numpy()
    dec.check()
    assert np.array_equal(out2[..., ::-1], out)


def test_decode_odd_size_uses_edge_replication_like_libjpeg():
    from cilrs_b200.data import JpegDecoder
    p = [q for q in JPEGS if _name(q) == "odd_87x199"][0]
    dec = JpegDecoder(4, height=87, width=199)
    out = dec.decode([open(p, "rb").read()] * 3).cpu().numpy()
    dec.check()
    assert all(np.array_equal(out[i], GOLDEN["odd_87x199"]) for i in range(3))


def test_decode_batch_128_equals_opencv_and_the_oracle():
    cv2 = pytest.importorskip("cv2")
    from cilrs_b200.data import JpegDecoder
    streams, refs = [], []
    for i, rgb in enumerate(_frames(128, seed=3)):
        ok, buf = cv2.imencode(".jpg", cv2.cvtColor(rgb, cv2.COLOR_RGB2BGR), [cv2.IMWRITE_JPEG_QUALITY, 95 if i % 8 else 50 + i % 40])
        streams.append(buf.tobytes())
        refs.append(cv2.cvtColor(cv2.imdecode(buf, cv2.IMREAD_COLOR), cv2.COLOR_BGR2RGB))
    dec = JpegDecoder(128)
    out = dec.decode(streams).cpu().numpy()
    dec.check()
    assert np.array_equal(out, np.stack(refs))
    for i in (0, 17, 127):
        assert np.array_equal(out[i], J.decode_rgb(streams[i]))
    # the staging buffers are double-buffered: decoding again (other order) while the first result is alive stays correct
    out2 = dec.decode(streams[::-1]).cpu().numpy()
    assert np.array_equal(out2, np.stack(refs[::-1]))


def test_decode_noise_frames_exercise_long_codes_and_stuffing():
    cv2 = pytest.importorskip("cv2")
    from cilrs_b200.data import JpegDecoder
    rng = np.random.default_rng(11)
    streams, refs = [], []
    for q in (100, 95, 90, 20):
        img = rng.integers(0, 256, (88, 200, 3), dtype=np.uint8)
        ok, buf = cv2.imencode(".jpg", img, [cv2.IMWRITE_JPEG_QUALITY, q])
        streams.append(buf.tobytes())
        refs.append(cv2.cvtColor(cv2.imdecode(buf, cv2.IMREAD_COLOR), cv2.COLOR_BGR2RGB))
    assert any(b"\xff\x00" in s for s in streams)   # byte stuffing present
    dec = JpegDecoder(4, max_bytes_per_image=1 << 17)
    out = dec.decode(streams).cpu().numpy()
    dec.check()
    assert np.array_equal(out, np.stack(refs))


def test_decode_reports_bad_frames_per_image():
    cv2 = pytest.importorskip("cv2")
    from cilrs_b200.data import JpegDecoder, JpegError
    good = open(JPEGS[0], "rb").read()
    ok, prog = cv2.imencode(".jpg", np.zeros((88, 200, 3), np.uint8), [cv2.IMWRITE_JPEG_PROGRESSIVE, 1])
    ok, small = cv2.imencode(".jpg", np.zeros((32, 32, 3), np.uint8))
    dec = JpegDecoder(8)
    out = dec.decode([good, b"garbage" * 10, prog.tobytes(), small.tobytes(), good])
    st = dec.d_status[:5].tolist()
    assert st == [0, 1, 2, 4, 0]
    o = out.cpu().numpy()
    assert np.array_equal(o[0], GOLDEN[_name(JPEGS[0])]) and np.array_equal(o[4], o[0]) and not o[1:4].any()
    with pytest.raises(JpegError):
        dec.check()
    # truncated entropy data: libjpeg pads with zero bits; here it must at least not read out of bounds and finish
    out = dec.decode([good[:len(good) // 2]])
    torch.cuda.synchronize()
    assert dec.d_status[0].item() in (0, 3)


def test_weighted_sampler_distribution_and_determinism():
    from cilrs_b200.data import DeviceSampler, class_weights
    rng = np.random.default_rng(0)
    cmd = rng.choice(4, size=20000, p=[0.7, 0.12, 0.1, 0.08])
    cw, w = class_weights(cmd)
    s = DeviceSampler(w, seed=7)
    idx = s.draw().cpu().numpy()
    assert idx.shape == (20000,) and idx.min() >= 0 and idx.max() < 20000
    freq = np.bincount(cmd[idx], minlength=4) / len(idx)
    assert np.abs(freq - 0.25).max() < 0.015          # the oversampling balances the four commands (sigma ~ 0.003)
    # rows of one class are drawn uniformly: chi-square over 50 bins of the LEFT rows
    rows = np.where(cmd == 1)[0]
    pos = np.searchsorted(rows, idx[cmd[idx] == 1])
    hist = np.bincount(pos * 50 // len(rows), minlength=50).astype(np.float64)
    chi2 = ((hist - hist.mean()) ** 2 / hist.mean()).sum()
    assert chi2 < 100                                  # 49 degrees of freedom: P(chi2 > 100) ~ 2e-5
    # same seed -> same draws; the second epoch continues the stream instead of repeating it
    s2 = DeviceSampler(w, seed=7)
    assert np.array_equal(s2.draw().cpu().numpy(), idx)
    assert not np.array_equal(s2.draw().cpu().numpy(), idx)
    # zero-weight rows are never drawn
    w0 = w.copy()
    w0[::2] = 0
    assert (DeviceSampler(w0, seed=1).draw(5000).cpu().numpy() % 2 == 1).all()


def _params(n, **kw):
    from cilrs_b200.augment import PARAM_DTYPE
    p = np.zeros(n, dtype=PARAM_DTYPE)
    for k, v in kw.items():
        p[k] = v
    return p


def test_augment_deterministic_transforms_equal_the_cv2_calls():
    from cilrs_b200 import augment
    frames = np.stack(_frames(32, seed=5))
    rng = np.random.default_rng(3)
    p = augment.draw_params(32, 88, 200, rng, p_scale=2.0)       # every transform fires often
    p["flags"] &= ~np.uint32(augment.F_NOISE)                    # (the noise draw is checked separately)
    p["flags"][0] = 0                                            # an untouched frame
    p["flags"][1] = 1 | 2 | 4 | 16                               # everything at once
    p["flags"][2:8] = [1, 2, 4, 16, 2 | 4, 1 | 2]
    aug = augment.DeviceAugmenter(32, seed=1)
    x = torch.from_numpy(frames).cuda()
    out = aug(x.clone(), params=p).cpu().numpy()
    bad = []
    for i in range(32):
        want = AO.apply(frames[i], p[i])          # cv2's scalar HSV->RGB path (the same on every host CPU): bit for bit
        if not np.array_equal(out[i], want):
            bad.append((i, int(p["flags"][i]), int(np.abs(out[i].astype(int) - want.astype(int)).max())))
        # OpenCV's vector path (what a wide row gets on an AVX2 host) truncates where its scalar code rounds: one level at most
        live = AO.apply(frames[i], p[i], vector_path=True)
        assert np.abs(out[i].astype(int) - live.astype(int)).max() <= 1, i
        if not (int(p["flags"][i]) & 2):
            assert np.array_equal(out[i], live), i   # without the HSV transform there is no host dependence at all
    assert not bad, bad
    assert np.array_equal(out[0], frames[0])
    # out-of-place form leaves the input alone
    y = torch.empty_like(x)
    aug(x, params=p, out=y)
    assert np.array_equal(x.cpu().numpy(), frames) and np.array_equal(y.cpu().numpy(), out)


def test_augment_hsv_round_trip_all_hues():
    """every hue shift on saturated colours, where OpenCV's float HSV->RGB lands on half-way cases"""
    from cilrs_b200 import augment
    rng = np.random.default_rng(9)
    frames = rng.integers(0, 256, (21, 88, 200, 3), dtype=np.uint8)
    p = _params(21, flags=2)
    p["hue"] = np.arange(-10, 11)
    p["sat"] = rng.integers(-20, 21, 21)
    p["val"] = rng.integers(-15, 16, 21)
    out = augment.DeviceAugmenter(21)(torch.from_numpy(frames).cuda(), params=p).cpu().numpy()
    for i in range(21):
        assert np.array_equal(out[i], AO.apply(frames[i], p[i])), i
        assert np.abs(out[i].astype(int) - AO.apply(frames[i], p[i], vector_path=True).astype(int)).max() <= 1, i


def test_augment_noise_statistics_and_dropout():
    from cilrs_b200 import augment
    frames = np.full((8, 88, 200, 3), 128, dtype=np.uint8)
    p = _params(8, flags=8)
    p["noise_std"] = np.linspace(0.02, 0.06, 8) * 255
    aug = augment.DeviceAugmenter(8, seed=4)
    a = aug(torch.from_numpy(frames).cuda(), params=p).cpu().numpy().astype(np.float64)
    for i in range(8):
        d = a[i] - 128
        n = d.size
        # truncation toward zero after the add (astype(uint8)) biases the mean by about -0.5
        assert abs(d.mean() + 0.5) < 4 * p["noise_std"][i] / np.sqrt(n) + 0.02
        assert abs(d.std() - np.sqrt(p["noise_std"][i] ** 2 + 1 / 12)) < 0.03 * p["noise_std"][i]
        assert abs(np.corrcoef(d[:, :-1].ravel(), d[:, 1:].ravel())[0, 1]) < 0.02   # neighbouring pixels independent
    b = aug(torch.from_numpy(frames).cuda(), params=p).cpu().numpy()
    assert not np.array_equal(a, b)                      # the next batch continues the noise stream
    assert np.array_equal(augment.DeviceAugmenter(8, seed=4)(torch.from_numpy(frames).cuda(), params=p).cpu().numpy(), a)
    # drawn parameters respect the reference's ranges and probabilities
    q = augment.draw_params(20000, 88, 200, np.random.default_rng(0))
    fl = q["flags"]
    for bit, prob in ((1, 0.5), (2, 0.3), (4, 0.2), (8, 0.3), (16, 0.2)):
        assert abs(((fl & bit) != 0).mean() - prob) < 0.012
    assert q["alpha"].min() >= 0.8 and q["alpha"].max() <= 1.2 and np.abs(q["beta"]).max() <= 0.2
    assert np.abs(q["hue"]).max() <= 10 and np.abs(q["sat"]).max() <= 20 and np.abs(q["val"]).max() <= 15
    assert set(np.unique(q["ksize"])) == {3, 5} and q["noise_std"].min() >= 0.02 * 255 and q["noise_std"].max() <= 0.06 * 255
    hh = q["hole"][..., 1] - q["hole"][..., 0]
    hw = q["hole"][..., 3] - q["hole"][..., 2]
    assert hh.min() >= 4 and hh.max() <= 10 and hw.min() >= 8 and hw.max() <= 20 and q["hole"].min() >= 0
    assert q["hole"][..., 1].max() <= 88 and q["hole"][..., 3].max() <= 200 and set(np.unique(q["n_holes"])) == {1, 2, 3}


def _write_dataset(root, n_per_session=40):
    import cv2
    names = ["LANEFOLLOW", "LEFT", "RIGHT", "STRAIGHT"]
    frames = {}
    k = 0
    for s in ("session1_town01", "session2_town02"):
        d = os.path.join(root, s)
        os.makedirs(os.path.join(d, "images"))
        with open(os.path.join(d, "measurements.csv"), "w") as f:
            f.write("frame,image_filename,steer,throttle,brake,speed_kmh,speed_normalized,high_level_command,command_name,"
                    "position_x,position_y,position_z,yaw,timestamp\n")
            for i, rgb in enumerate(_frames(n_per_session, seed=20 + k)):
                fn = "frame_%08d.jpg" % i
                path = os.path.join(d, "images", fn)
                cv2.imwrite(path, cv2.cvtColor(rgb, cv2.COLOR_RGB2BGR), [cv2.IMWRITE_JPEG_QUALITY, 95])
                frames[path] = cv2.cvtColor(cv2.imread(path), cv2.COLOR_BGR2RGB)
                cmd = 0 if i % 5 else 1 + (i // 5) % 3
                f.write("%d,%s,%.6f,%.6f,%.6f,%.2f,%.6f,%d,%s,0,0,0,0,%.3f\n" % (i, fn, 0.01 * i - 0.2, 0.5 + 0.001 * i, 0.0, 30.0, 30.0 / 90.0 + 0.001 * i + 0.05 * k,
                                                                               cmd, names[cmd], 0.05 * i))
        k += 1
    return frames


def test_device_loader_yields_the_reference_loaders_batches_and_trains(tmp_path):
    from cilrs_b200 import data
    from cilrs_b200.model import CILRS
    from cilrs_b200.train import FusedTrainer
    ref_frames = _write_dataset(str(tmp_path))
    table = data.load_sessions(str(tmp_path))
    assert len(table["image_path"]) == 80
    cw, w = data.class_weights(table["command_idx"])
    loader = data.DeviceLoader(table, batch=16, weights=w, seed=3)
    assert len(loader) == 5
    path_index = {p: i for i, p in enumerate(table["image_path"])}
    seen_cmd = []
    batches = 0
    for frames, speed, command, targets in loader:
        loader.decoder.check()
        f = frames.cpu().numpy()
        sp, cm, tg = speed.cpu().numpy(), command.cpu().numpy(), targets.cpu().numpy()
        for j in range(16):
            # identify the row by its (unique) speed label, then the frame must be that row's file as the reference decodes it
            row = int(np.argmin(np.abs(table["speed_normalized"].astype(np.float64) - sp[j]) + 10.0 * (table["command_idx"] != cm[j])))
            assert np.array_equal(f[j], ref_frames[table["image_path"][row]])
            assert np.allclose(tg[j], [table["steer"][row], table["throttle"][row], table["brake"][row]])
        seen_cmd.extend(cm.tolist())
        batches += 1
    assert batches == 5
    assert len(set(seen_cmd)) == 4          # the rare commands are oversampled into the epoch
    # one epoch through the trainer: u8 frames straight from the loader (with the augmentations on), finite decreasing-ish loss
    from cilrs_b200.augment import DeviceAugmenter
    torch.manual_seed(0)
    model = CILRS(num_commands=4, dropout=0.0).cuda()
    tr = FusedTrainer(model, 16, lr=1e-3, loss="l1", speed_w=0.5, frames="u8")
    loader = data.DeviceLoader(table, batch=16, weights=w, seed=4, augment=DeviceAugmenter(16, seed=2))
    losses = []
    for epoch in range(3):
        for batch in loader:
            tr.load_batch(*batch)
            tr.step()
            losses.append(tr.read_loss()["total"])
    assert all(np.isfinite(losses)) and np.mean(losses[-5:]) < np.mean(losses[:5])
