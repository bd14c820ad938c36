"""N3 - the reference's input pipeline (notebook/notebook.ipynb:361-431) with its per-frame work on the device.

    reference (per frame, CPU workers)                      here
    ------------------------------------------------------  ---------------------------------------------------------------
    pd.read_csv(measurements.csv) per session (:361-376)    load_sessions(): same columns, numpy arrays (host bookkeeping)
    class weights + WeightedRandomSampler (:384-385,411)    class_weights() + DeviceSampler (inverse-CDF draws on the device)
    cv2.imread + cvtColor(BGR2RGB) (:404-405)               JpegDecoder: threaded file reads, Huffman + IDCT + upsampling +
                                                            colour conversion on the GPU, bit-identical to OpenCV's result
    albumentations Compose (:387-394,406-407)               augment.DeviceAugmenter
    /255, Normalize, ToTensor (:408-410)                    K0 inside FusedTrainer (frames="u8")
    DataLoader(batch, sampler, drop_last) (:415-420)        DeviceLoader: yields device batches, next batch decoded on a side
                                                            stream while the current one trains
"""
import csv
import ctypes
import os

import numpy as np
import torch

from . import _lib

COMMAND_MAP = {"LANEFOLLOW": 0, "LEFT": 1, "RIGHT": 2, "STRAIGHT": 3}   # notebook/notebook.ipynb:359


def load_sessions(data_dir):
    """All `session*` directories of data_dir (sorted), their measurements.csv rows concatenated (model/collect_data.py:549-564
    writes the 14 columns). Returns a dict of numpy arrays: image_path (object), speed_normalized, steer, throttle, brake
    (float32), command_idx (int64), session (object)."""
    sessions = sorted(d for d in os.listdir(data_dir) if os.path.isdir(os.path.join(data_dir, d)) and "session" in d)
    cols = {k: [] for k in ("image_path", "speed_normalized", "steer", "throttle", "brake", "command_idx", "session")}
    for s in sessions:
        with open(os.path.join(data_dir, s, "measurements.csv"), newline="") as f:
            for row in csv.DictReader(f):
                cols["image_path"].append(os.path.join(data_dir, s, "images", row["image_filename"]))
                cols["speed_normalized"].append(float(row["speed_normalized"]))
                cols["steer"].append(float(row["steer"]))
                cols["throttle"].append(float(row["throttle"]))
                cols["brake"].append(float(row["brake"]))
                name = row["command_name"]
                if name not in COMMAND_MAP:
                    raise ValueError("measurements.csv: unknown command_name %r" % name)
                cols["command_idx"].append(COMMAND_MAP[name])
                cols["session"].append(s)
    out = {k: np.asarray(v, dtype=object) for k, v in cols.items() if k in ("image_path", "session")}
    for k in ("speed_normalized", "steer", "throttle", "brake"):
        out[k] = np.asarray(cols[k], dtype=np.float32)
    out["command_idx"] = np.asarray(cols["command_idx"], dtype=np.int64)
    return out


def class_weights(command_idx):
    """{c: N / (4 * count_c)} and the per-row sample weights (notebook/notebook.ipynb:384-385,411)."""
    n = len(command_idx)
    counts = np.bincount(command_idx, minlength=4)
    cw = {i: n / (4 * (int(counts[i]) if counts[i] else 1)) for i in range(4)}
    return cw, np.asarray([cw[int(c)] for c in command_idx], dtype=np.float64)


class DeviceSampler:
    """WeightedRandomSampler(weights, num_samples, replacement=True) with the draws made on the device."""

    def __init__(self, weights, num_samples=None, seed=0, device=None):
        dev = torch.device(device if device is not None else "cuda")
        w = torch.as_tensor(np.asarray(weights, dtype=np.float64))
        if w.numel() == 0 or bool((w < 0).any()) or float(w.sum()) <= 0:
            raise ValueError("DeviceSampler: weights must be non-negative with a positive sum")
        self.n = w.numel()
        self.num_samples = int(num_samples) if num_samples is not None else self.n
        self.cdf = torch.cumsum(w.to(dev), 0)    # one-time set-up
        self.seed = int(seed) & 0xFFFFFFFFFFFFFFFF
        self._offset = 0

    def draw(self, num_samples=None):
        k = int(num_samples) if num_samples is not None else self.num_samples
        out = torch.empty(k, dtype=torch.long, device=self.cdf.device)
        _lib.call("cilrs_weighted_sample", self.cdf, ctypes.c_longlong(self.n), ctypes.c_longlong(k), ctypes.c_ulonglong(self.seed),
                  ctypes.c_ulonglong(self._offset), out, _lib.stream_ptr())
        self._offset += k
        return out


class JpegError(RuntimeError):
    pass


_STATUS = {1: "not a JPEG stream", 2: "unsupported JPEG (only baseline Huffman, 8 bit, grey / 4:4:4 / 4:2:0, no restart markers)",
           3: "corrupt JPEG stream", 4: "frame size differs from the decoder's"}


class JpegDecoder:
    """decoder.decode(list of bytes) / decoder.decode_files(list of paths) -> uint8 [n, H, W, 3] RGB on the device, equal to
    cv2.cvtColor(cv2.imread(path), cv2.COLOR_BGR2RGB) bit for bit. Work is queued on the current stream; `check()` (one D2H)
    raises JpegError for a frame that could not be decoded."""

    def __init__(self, max_batch, height=88, width=200, device=None, max_bytes_per_image=65536, threads=8, max_sets=4):
        lib = _lib.lib()
        self.dev = torch.device(device if device is not None else "cuda")
        self.h, self.w, self.max_batch, self.threads, self.max_sets = height, width, max_batch, threads, max_sets
        self.desc_bytes = lib.cilrs_jpeg_desc_bytes()
        self.set_bytes = lib.cilrs_jpeg_table_set_bytes()
        self.plane_bytes = lib.cilrs_jpeg_plane_bytes(height, width)
        self.capacity = max_batch * ((max_bytes_per_image + 15) // 16 * 16)
        self._slots = []
        for _ in range(2):   # double-buffered pinned staging: the next batch is assembled while the previous H2D is in flight
            self._slots.append(dict(bytes=torch.zeros(self.capacity, dtype=torch.uint8).pin_memory(),
                                    descs=torch.zeros(max_batch * self.desc_bytes, dtype=torch.uint8).pin_memory(),
                                    sets=torch.zeros(max_sets * self.set_bytes, dtype=torch.uint8).pin_memory(),
                                    offsets=np.zeros(2 * max_batch + 1, dtype=np.int64), event=None))
        self._slot = 0
        self.d_bytes = torch.zeros(self.capacity, dtype=torch.uint8, device=self.dev)
        self.d_descs = torch.zeros(max_batch * self.desc_bytes, dtype=torch.uint8, device=self.dev)
        self.d_sets = torch.zeros(max_sets * self.set_bytes, dtype=torch.uint8, device=self.dev)
        self.d_planes = torch.zeros(max_batch * self.plane_bytes, dtype=torch.uint8, device=self.dev)
        self.d_status = torch.zeros(max_batch, dtype=torch.int32, device=self.dev)
        self._last_n = 0

    def _next_slot(self):
        slot = self._slots[self._slot]
        self._slot ^= 1
        if slot["event"] is not None:
            slot["event"].synchronize()
        return slot

    def _launch(self, slot, n, ends, out, reverse):
        lib = _lib.lib()
        offs = slot["offsets"]
        total = int(offs[n])
        # descriptor offsets: starts + exact ends (a stream's own length, not the padded slot)
        span = np.empty(n + 1, dtype=np.int64)
        span[:n] = offs[:n]
        span[n] = total
        descs, sets = slot["descs"], slot["sets"]
        n_sets = ctypes.c_int(0)
        # cilrs_jpeg_prepare takes [offsets[i], offsets[i + 1]): trailing padding after EOI is ignored by the parser / decoder
        st = lib.cilrs_jpeg_prepare(ctypes.c_void_p(slot["bytes"].data_ptr()), span.ctypes.data_as(ctypes.c_void_p), n,
                                    ctypes.c_void_p(descs.data_ptr()), ctypes.c_void_p(sets.data_ptr()), self.max_sets, ctypes.byref(n_sets))
        if st != 0:
            raise RuntimeError("cilrs_b200.cilrs_jpeg_prepare failed: %s" % lib.cilrs_status_string(st).decode())
        self.d_bytes[:total].copy_(slot["bytes"][:total], non_blocking=True)
        self.d_descs[:n * self.desc_bytes].copy_(descs[:n * self.desc_bytes], non_blocking=True)
        ns = max(1, n_sets.value)
        self.d_sets[:ns * self.set_bytes].copy_(sets[:ns * self.set_bytes], non_blocking=True)
        ev = torch.cuda.Event()
        ev.record()
        slot["event"] = ev
        if out is None:
            out = torch.empty(n, self.h, self.w, 3, dtype=torch.uint8, device=self.dev)
        elif out.dtype != torch.uint8 or tuple(out.shape) != (n, self.h, self.w, 3) or not out.is_contiguous():
            raise ValueError("JpegDecoder: out must be a contiguous uint8 [n, H, W, 3] tensor")
        _lib.call("cilrs_jpeg_decode", self.d_bytes, self.d_descs, self.d_sets, n, self.h, self.w, self.d_planes,
                  ctypes.c_size_t(self.plane_bytes), out, int(bool(reverse)), self.d_status, _lib.stream_ptr())
        self._last_n = n
        return out

    def decode(self, streams, out=None, reverse=False):
        n = len(streams)
        if n > self.max_batch:
            raise ValueError("JpegDecoder: batch larger than max_batch")
        slot = self._next_slot()
        host = slot["bytes"].numpy()
        offs = slot["offsets"]
        off = 0
        for i, b in enumerate(streams):
            a = np.frombuffer(b, dtype=np.uint8)
            if off + a.size > self.capacity:
                raise ValueError("JpegDecoder: staging buffer too small (raise max_bytes_per_image)")
            offs[i] = off
            host[off:off + a.size] = a
            pad = (-a.size) % 16
            host[off + a.size:off + a.size + pad] = 0
            off += a.size + pad
        offs[n] = off
        return self._launch(slot, n, None, out, reverse)

    def decode_files(self, paths, out=None, reverse=False):
        n = len(paths)
        if n > self.max_batch:
            raise ValueError("JpegDecoder: batch larger than max_batch")
        slot = self._next_slot()
        arr = (ctypes.c_char_p * n)(*[os.fsencode(p) for p in paths])
        offs = slot["offsets"]
        lib = _lib.lib()
        st = lib.cilrs_jpeg_read_files(arr, n, ctypes.c_void_p(slot["bytes"].data_ptr()), ctypes.c_longlong(self.capacity),
                                       offs.ctypes.data_as(ctypes.c_void_p), self.threads)
        if st == 3:
            raise ValueError("JpegDecoder: staging buffer too small (raise max_bytes_per_image)")
        if st != 0:
            missing = [paths[i] for i in range(n) if offs[n + 1 + i] == offs[i]]
            raise FileNotFoundError("JpegDecoder: cannot read %s" % (missing[:3],))
        return self._launch(slot, n, None, out, reverse)

    def check(self):
        st = self.d_status[:self._last_n].tolist()
        bad = [(i, s) for i, s in enumerate(st) if s != 0]
        if bad:
            i, s = bad[0]
            raise JpegError("frame %d of the batch: %s (%d frames failed)" % (i, _STATUS.get(s, "status %d" % s), len(bad)))


class DeviceLoader:
    """Iterating yields (frames_u8 [B, H, W, 3], speed [B], command [B], targets [B, 3]) device tensors - what
    FusedTrainer(frames="u8").load_batch takes - for one epoch of `num_samples` weighted draws (drop_last, as the reference's
    train loader). The NEXT batch's file reads, H2D copy, decode and augmentation are queued on a side stream before the
    current batch is handed out, so they run under the training step."""

    def __init__(self, table, batch, weights=None, num_samples=None, augment=None, seed=0, device=None, height=88, width=200,
                 shuffle=True, threads=8):
        self.dev = torch.device(device if device is not None else "cuda")
        self.table, self.batch = table, batch
        self.n = len(table["image_path"])
        self.paths = [str(p) for p in table["image_path"]]
        self.speed = torch.as_tensor(table["speed_normalized"], dtype=torch.float32).to(self.dev)
        self.command = torch.as_tensor(table["command_idx"], dtype=torch.long).to(self.dev)
        self.targets = torch.as_tensor(np.stack([table["steer"], table["throttle"], table["brake"]], axis=1), dtype=torch.float32).to(self.dev)
        self.sampler = DeviceSampler(weights, num_samples, seed, self.dev) if weights is not None else None
        self.num_samples = self.sampler.num_samples if self.sampler is not None else self.n
        self.shuffle = shuffle
        self.gen = torch.Generator(device="cpu").manual_seed(seed)
        self.decoder = JpegDecoder(batch, height, width, self.dev, threads=threads)
        self.augment = augment
        self.stream = torch.cuda.Stream(device=self.dev)
        self.frames = [torch.zeros(batch, height, width, 3, dtype=torch.uint8, device=self.dev) for _ in range(2)]

    def __len__(self):
        return self.num_samples // self.batch

    def _indices(self):
        if self.sampler is not None:
            return self.sampler.draw()
        if self.shuffle:
            return torch.randperm(self.n, generator=self.gen).to(self.dev)
        return torch.arange(self.n, device=self.dev)

    def __iter__(self):
        idx_dev = self._indices()
        idx_host = idx_dev.cpu().numpy()     # one D2H per epoch: the host needs the file names
        nb = len(self)
        cur = torch.cuda.current_stream(self.dev)

        def issue(k):
            sl = slice(k * self.batch, (k + 1) * self.batch)
            buf = self.frames[k & 1]
            self.stream.wait_stream(cur)     # the consumer is done with this buffer (it was handed out two batches ago)
            with torch.cuda.stream(self.stream):
                self.decoder.decode_files([self.paths[i] for i in idx_host[sl]], out=buf)
                if self.augment is not None:
                    self.augment(buf)
                ids = idx_dev[sl]
                batch = (buf, self.speed[ids], self.command[ids], self.targets[ids])
                ev = torch.cuda.Event()
                ev.record(self.stream)
            return batch, ev

        pending = issue(0) if nb > 0 else None
        for k in range(nb):
            batch, ev = pending
            pending = issue(k + 1) if k + 1 < nb else None
            cur.wait_event(ev)
            for t in batch:
                t.record_stream(cur)
            yield batch
