#!/usr/bin/env python
"""bench.py — CILRS training-step throughput on N B200s (BASELINE.json: "train frames/s (bf16, bs128/GPU)").

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

N > 1 is launched by torchrun (one rank per GPU, NCCL); rank 0 prints ONE JSON line.

Workload (BASELINE.json configs[1] / [3]): CILRS train step, batch 128 per GPU, 200x88 synthetic frames, MSE(controls) +
0.05 MSE(speed), Adam lr 2e-4 wd 1e-4, train-mode BatchNorm, dropout 0, random-init weights (reference initialisers), bf16
tensor-core convs with fp32 accumulate, fp32 master weights / heads / optimizer. A step = forward + loss + zero_grad + backward
(+ gradient allreduce) + Adam + bf16 operand repack; nothing is skipped or cached.

  value : frames/s, inputs already in HBM (a rotating pool of device batches), CUDA events over K steps, max over ranks
  e2e   : same steps through the public API with HOST (pinned) uint8 frames: H2D copy (prefetched one step ahead on a copy
          stream) + on-device normalise (K0) + step + D2H read of the loss scalars every step
  roofline     : the implicit-GEMM conv kernels (conv_flat_kernel: 58 of the 77 fprop+dgrad launches, conv_gemm_kernel: stem /
                 stride-2 / 1x1): algorithmic conv FLOPs / time inside those launches (CUDA events around every launch of
                 one eager step; the launches also carry the fused BN statistics / ReLU mask / BN-backward reductions)
  cpu_baseline : the reference's own `class CILRS` train step (oracle/_ref, extracted at build() time; the oracle port only if that
                 file is absent) in torch fp32 on ALL host cores, bounded sample, rank 0 / N=1 only
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

FLOP_FWD = 2795915264.0
FLOP_DGRAD = 2713124864.0
FLOP_WGRAD = 2795915264.0
FLOP_TRAIN = FLOP_FWD + FLOP_DGRAD + FLOP_WGRAD  # 8 304 955 392 conv FLOP / frame
BATCH = 128
METRIC = "train frames/s (bf16, bs128/GPU)"
WORKLOAD = ("CILRS training step (ResNet-34 + speed encoder + 4 command branches), batch 128 per GPU, 200x88 synthetic frames, "
            "MSE+0.05*MSE loss, Adam lr 2e-4 wd 1e-4, train-mode BN, bf16 tcgen05 convs / fp32 master weights")


def peaks():
    p = {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0, "source": "fallback (B200_PROFILING.md)"}
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        try:
            j = json.load(open(path))
            p.update({k: j[k] for k in ("hbm_gbs", "bf16_tflops", "bf16_tflops_sustained") if k in j})
            p["source"] = "MEASURED_PEAKS.json"
        except Exception:
            pass
    return p


class ClockSampler:
    """SM clock / throttle reasons sampled DURING the timed region: NVML every ~20 ms from a thread (the timed region of the
    default run is ~0.1 s - an `nvidia-smi -lms 200` child would not deliver a single sample); nvidia-smi as the fallback."""
    REASONS = ((0x8, "hw_slowdown"), (0x40, "hw_thermal_slowdown"), (0x20, "sw_thermal_slowdown"), (0x4, "sw_power_cap"))

    def __init__(self, index):
        self.index = index
        self.sm, self.mask = [], 0
        self.max_mhz = None
        self.handle = None
        self.stop_flag = False
        self.thread = None
        self.source = None

    def _nvml_handle(self):
        import pynvml
        pynvml.nvmlInit()
        try:
            import torch
            uuid = str(torch.cuda.get_device_properties(self.index).uuid)
            return pynvml, pynvml.nvmlDeviceGetHandleByUUID(("GPU-" + uuid).encode())
        except Exception:
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            idx = self.index
            if vis:
                try:
                    idx = int(vis.split(",")[self.index])
                except Exception:
                    pass
            return pynvml, pynvml.nvmlDeviceGetHandleByIndex(idx)

    def _sample_nvml(self):
        nv = self.nv
        self.sm.append(float(nv.nvmlDeviceGetClockInfo(self.handle, nv.NVML_CLOCK_SM)))
        try:
            self.mask |= int(nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.handle))
        except Exception:
            pass

    def _loop(self):
        while not self.stop_flag:
            try:
                self._sample_nvml()
            except Exception:
                break
            time.sleep(0.02)   # (a 2 ms loop measurably slowed the multi-rank step: it competes with the launch / NCCL threads)

    def start(self):
        try:
            self.nv, self.handle = self._nvml_handle()
            self.max_mhz = float(self.nv.nvmlDeviceGetMaxClockInfo(self.handle, self.nv.NVML_CLOCK_SM))
            self._sample_nvml()
            self.source = "nvml"
            self.thread = threading.Thread(target=self._loop, daemon=True)
            self.thread.start()
        except Exception:
            self.handle = None

    def _smi_once(self):
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        out = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + q, "--format=csv,noheader,nounits"],
                             capture_output=True, text=True, timeout=10).stdout.strip().splitlines()[0]
        r = [c.strip() for c in out.split(",")]
        self.sm.append(float(r[0]))
        self.max_mhz = float(r[1])
        for (bit, _), v in zip(self.REASONS, r[2:6]):
            if v.lower().startswith("active"):
                self.mask |= bit

    def stop(self):
        self.stop_flag = True
        if self.thread is not None:
            self.thread.join(timeout=1)
        if self.handle is None:
            # fallback: one nvidia-smi reading right after the timed region (the GPU is still under load from the e2e leg)
            try:
                self._smi_once()
                self.source = "nvidia-smi (single reading after the timed region)"
            except Exception:
                return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvml and nvidia-smi unavailable"], "samples": 0}
        sm = sorted(self.sm)
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": self.max_mhz,
                "reasons": [name for bit, name in self.REASONS if self.mask & bit], "samples": len(sm), "source": self.source}


def _reference_module():
    """oracle/_ref/reference_hotpath.py (the reference's own classes, extracted by oracle/extract_ref.py at build() time and
    shipped with the snapshot), or None -> the oracle port is used instead."""
    try:
        from oracle import extract_ref
        return extract_ref.load()
    except Exception:
        return None


def synthetic_host_batches(batch, pool, seed):
    """The bench's synthetic data: uint8 200x88 frames (what the reference's dataset stores after prepare_dataset.py),
    speed U[0,1), command randint(0,4), targets U[0,1)^3. Same generator for our arm and the reference arm."""
    import torch
    g = torch.Generator().manual_seed(seed)
    out = []
    for _ in range(pool):
        out.append((torch.randint(0, 256, (batch, 88, 200, 3), generator=g, dtype=torch.uint8), torch.rand(batch, generator=g),
                    torch.randint(0, 4, (batch,), generator=g), torch.rand(batch, 3, generator=g)))
    return out


def reference_initial_state_dict():
    """Random-init weights the way the reference constructs them: `CILRS(num_commands=4, dropout=0.0)` under
    torch.manual_seed(0) (kaiming fan_out convs, BN 1/0, default Linear init). Falls back to the oracle's synthetic state dict
    (same initialiser scales) when oracle/_ref is absent. Returns (state_dict, kind)."""
    import warnings
    import torch
    R = _reference_module()
    if R is not None:
        torch.manual_seed(0)
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            ref = R.CILRS(num_commands=4, dropout=0.0)
        return {k: v.detach().clone() for k, v in ref.state_dict().items()}, "reference"
    from oracle import cilrs_oracle as O
    return O.synthetic_state_dict(0, perturb_bn=False), "port"


def own_initial_state_dict():
    """Random-init weights for OUR arm, made by our own module's constructor under torch.manual_seed(0) (the reference's
    initialisers: kaiming fan_out convs, BN 1/0, default Linear init - tests/test_host_cpu.py) - nothing under oracle/ is touched.
    The CPU leg loads this very state_dict into the reference's class (strict), so both start from the same weights."""
    import torch
    from cilrs_b200.model import CILRS
    torch.manual_seed(0)
    m = CILRS(num_commands=4, dropout=0.0)
    return {k: v.detach().clone() for k, v in m.state_dict().items()}, "cilrs_b200.CILRS ctor"


def cpu_reference_step_rate(batch, steps, warmup, state_dict=None, batches=None):
    """The reference's train step on the host cores, ALL of them (torch.set_num_threads(os.cpu_count()): torchrun exports
    OMP_NUM_THREADS=1, which must not apply to this arm): `model(imgs, speeds, cmds)` -> loss -> zero_grad -> backward ->
    optimizer.step(), i.e. the body of train_one_epoch (notebook/notebook.ipynb:545-555) with the BASELINE recipe (MSE + 0.05 MSE,
    torch.optim.Adam lr 2e-4 wd 1e-4). Runs the reference's own `class CILRS` from oracle/_ref when it was built (kind
    "reference"), else the oracle's functional port (kind "port"). Returns (frames/s, s/step, threads, kind, first-step loss)."""
    import warnings
    import torch
    from oracle import cilrs_oracle as O
    torch.set_num_threads(os.cpu_count() or 1)
    R = _reference_module()
    if state_dict is None:
        state_dict, _ = reference_initial_state_dict()
    if batches is None:
        batches = synthetic_host_batches(batch, 2, 100)
    data = []
    for fr, sp, cm, tg in batches:
        img = torch.from_numpy(O.normalise_np(fr.numpy()))   # the loader's /255 + Normalize (notebook/notebook.ipynb:412-414)
        data.append((img, sp.clone(), cm.clone(), tg.clone()))
    times, first_loss = [], None
    if R is not None:
        kind = "reference"
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            model = R.CILRS(num_commands=4, dropout=0.0)
        model.load_state_dict(state_dict)
        model.train()
        optimizer = torch.optim.Adam(model.parameters(), lr=2e-4, weight_decay=1e-4)
        for i in range(warmup + steps):
            imgs, speeds, cmds, tgts = data[i % len(data)]
            t0 = time.perf_counter()
            pred_ctrl, pred_spd = model(imgs, speeds, cmds)
            loss, _ = O.loss_mse(pred_ctrl, tgts, pred_spd, speeds)   # README / train_config recipe; the reference has no code for it
            optimizer.zero_grad()
            loss.backward()
            optimizer.step()
            lv = loss.item()
            if first_loss is None:
                first_loss = lv
            if i >= warmup:
                times.append(time.perf_counter() - t0)
    else:
        kind = "port"
        params = {k: v.clone().requires_grad_(True) for k, v in state_dict.items()
                  if v.is_floating_point() and not k.endswith(("running_mean", "running_var"))}
        state = dict(state_dict)
        state.update(params)
        opt = torch.optim.Adam(list(params.values()), lr=2e-4, weight_decay=1e-4)
        for i in range(warmup + steps):
            imgs, speeds, cmds, tgts = data[i % len(data)]
            t0 = time.perf_counter()
            upd = {}
            c, p = O.forward(state, imgs, speeds, cmds, training=True, update=upd)
            loss, _ = O.loss_mse(c, tgts, p, speeds)
            opt.zero_grad()
            loss.backward()
            opt.step()
            state.update(upd)
            lv = loss.item()
            if first_loss is None:
                first_loss = lv
            if i >= warmup:
                times.append(time.perf_counter() - t0)
    total = sum(times)
    return batch * len(times) / total, total / len(times), torch.get_num_threads(), kind, first_loss


def run_reference(args, rank):
    """--impl reference: the reference's own CPU implementation of the path, on the stated config (batch 128), with every host
    thread: oracle/_ref (the reference's class CILRS, extracted at build() time) when present, else the oracle port."""
    if rank != 0:
        return
    warm = max(1, min(args.warmup, 2))
    fps, sec, cores, kind, _ = cpu_reference_step_rate(BATCH, max(1, args.steps), warm)
    line = {
        "impl": "reference", "metric": METRIC, "value": fps, "unit": "frames/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": warm, "ms_per_step": sec * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD, "per_gpu_batch": BATCH, "global_batch": BATCH * args.gpus, "parallelism": "cpu"},
        "cpu_baseline": {"value": fps, "unit": "frames/s", "cores": cores, "kind": kind,
                         "sample": "each step = one full train step (fwd + MSE loss + bwd + torch.optim.Adam) on a %d-frame batch, "
                                   "torch fp32 on the host, %d threads" % (BATCH, cores)},
        "e2e": {"value": fps, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


def _timeit(torch, fn, reps=10, flush=None):
    fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        if flush is not None:
            flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    ts.sort()
    return ts[len(ts) // 2] * 1e-3


def extra_kernel_lines(torch, model, pk):
    """The HBM-bound kernels of the path, timed alone (CUDA events, L2 flushed by a 256 MB memset between repetitions):
    K0 preprocessing on configs[2] (1024 frames), the fused Adam, the heads at batch 128."""
    import ctypes
    from cilrs_b200 import _lib
    out = {}
    sp = _lib.stream_ptr()
    hbm = pk["hbm_gbs"]
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    try:
        B = 1024
        frames = torch.randint(0, 256, (B, 600, 800, 3), dtype=torch.uint8, device="cuda")
        f32 = torch.empty(B, 3, 88, 200, device="cuda")
        t = _timeit(torch, lambda: _lib.call("cilrs_preprocess_u8", frames, B, 600, 800, 3, 0, 88, 200, None, f32, None, sp), flush=flush)
        alg = 633600.0
        out["c3_preprocess"] = {"workload": "1024 x (600x800x3 u8 -> 88x200 f32 NCHW), K0", "ms": t * 1e3, "frames_per_s": B / t,
                                "algorithmic_GBps": B * alg / t / 1e9, "hbm_frac": B * alg / t / 1e9 / hbm, "bytes_per_frame": alg}
        del frames, f32
    except Exception as ex:
        out["c3_preprocess"] = {"error": repr(ex)}
    try:
        n = model.flat_parameters().numel()
        p = torch.randn(n, device="cuda")
        g = torch.randn(n, device="cuda") * 1e-3
        mm, vv = torch.zeros(n, device="cuda"), torch.zeros(n, device="cuda")
        hyper = torch.tensor([2e-4, 0.9, 0.999, 1e-8, 1e-4, 1.0, 0, 0], device="cuda")
        step = torch.zeros(1, dtype=torch.long, device="cuda")
        t = _timeit(torch, lambda: _lib.call("cilrs_adam_step_ex", p, g, None, mm, vv, ctypes.c_longlong(n), hyper, step, None, 1, sp),
                    flush=flush)
        out["adam"] = {"workload": "fused Adam + zero_grad over the 22.4 M-parameter arena (32 B/param)", "ms": t * 1e3,
                       "GBps": n * 32 / t / 1e9, "hbm_frac": n * 32 / t / 1e9 / hbm}
        del p, g, mm, vv
    except Exception as ex:
        out["adam"] = {"error": repr(ex)}
    try:
        B = 128
        model.train()
        model._ensure(B)
        feat = torch.randn(B, 512, device="cuda")
        speed = torch.rand(B, device="cuda")
        cmd = torch.randint(0, 4, (B,), device="cuda")
        c, ps = torch.empty(B, 3, device="cuda"), torch.empty(B, device="cuda")
        dc, dsp, df = torch.randn(B, 3, device="cuda"), torch.randn(B, device="cuda"), torch.empty(B, 512, device="cuda")
        model.flat_gradients()
        t = _timeit(torch, lambda: _lib.call("cilrs_model_heads_forward", model._handle, B, feat, speed, cmd, c, ps, 1, ctypes.c_float(0.0),
                                             ctypes.c_ulonglong(1), sp))
        t2 = _timeit(torch, lambda: _lib.call("cilrs_model_heads_backward", model._handle, B, dc, dsp, speed, cmd, ctypes.c_float(0.0), df, sp))
        out["heads"] = {"workload": "heads at batch 128", "forward_ms": t * 1e3, "backward_incl_wgrad_ms": t2 * 1e3}
        model.flat_gradients().zero_()
    except Exception as ex:
        out["heads"] = {"error": repr(ex)}
    try:
        # N3 / N4: the reference loader's per-frame work (cv2.imread + albumentations, notebook/notebook.ipynb:404-407) for one
        # batch of 128 collector-format frames (200 x 88 JPEG, quality 95): host bytes -> decoded + augmented u8 frames in HBM
        import time
        import cv2
        from cilrs_b200.augment import DeviceAugmenter
        from cilrs_b200.data import JpegDecoder
        rng = __import__("numpy").random.default_rng(0)
        streams = []
        for i in range(128):
            img = cv2.GaussianBlur(rng.integers(0, 256, (88, 200, 3), dtype="uint8"), (5, 5), 1.5)
            img[20:40, 30 + i % 50:90 + i % 50] = rng.integers(0, 256, 3)
            streams.append(cv2.imencode(".jpg", img, [cv2.IMWRITE_JPEG_QUALITY, 95])[1].tobytes())
        dec, aug = JpegDecoder(128), DeviceAugmenter(128, seed=0)
        buf = torch.empty(128, 88, 200, 3, dtype=torch.uint8, device="cuda")

        def run():
            dec.decode(streams, out=buf)
            aug(buf)
        t = _timeit(torch, run)
        t0 = time.perf_counter()
        for s_ in streams:
            cv2.cvtColor(cv2.imdecode(__import__("numpy").frombuffer(s_, dtype="uint8"), cv2.IMREAD_COLOR), cv2.COLOR_BGR2RGB)
        t_cpu = time.perf_counter() - t0
        dec.check()
        out["n3_decode_augment"] = {"workload": "128 JPEG frames (200x88, q95, %d B avg): host bytes -> decoded + augmented u8 frames" % (sum(map(len, streams)) // 128),
                                    "ms": t * 1e3, "frames_per_s": 128 / t, "cv2_imdecode_1_thread_frames_per_s": 128 / t_cpu}
    except Exception as ex:
        out["n3_decode_augment"] = {"error": repr(ex)}
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-graph", action="store_true")
    ap.add_argument("--no-extras", action="store_true")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        run_reference(args, rank)
        return
    if args.warmup < 3:
        args.warmup = 3

    import ctypes
    import torch
    import torch.distributed as dist
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (the CUDA path has no CPU fallback); use --impl reference for the CPU arm")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        # measurement aid: cap NCCL's persistent CTAs (they take SMs from the 148-CTA convolution grids while an allreduce overlaps
        # the backward); reported in config.nccl_max_ctas
        if os.environ.get("CILRS_BENCH_NCCL_MAX_CTAS"):
            os.environ["NCCL_MAX_CTAS"] = os.environ["CILRS_BENCH_NCCL_MAX_CTAS"]
        dist.init_process_group("nccl", device_id=dev)
    import cilrs_b200  # noqa: F401
    from cilrs_b200 import _lib
    from cilrs_b200.model import CILRS
    from cilrs_b200.train import FusedTrainer

    lib = _lib.lib()
    lib.cilrs_launch_count.restype = ctypes.c_longlong
    # random-init weights from our own constructor, loaded through the checkpoint interface (the CPU leg loads the same ones)
    init_sd, init_kind = own_initial_state_dict()
    model = CILRS(num_commands=4, dropout=0.0)
    model.load_state_dict(init_sd, strict=True)
    model = model.to(dev)
    trainer = FusedTrainer(model, BATCH, lr=2e-4, weight_decay=1e-4, loss="mse", speed_w=0.05, frames="u8",
                           use_graph=(not args.no_graph),
                           overlap_allreduce=os.environ.get("CILRS_BENCH_ALLREDUCE", "two"),       # measurement aids
                           grad_comm=os.environ.get("CILRS_BENCH_GRAD_COMM", "bf16"),
                           async_parts=os.environ.get("CILRS_BENCH_ASYNC_PARTS", "0") == "1",
                           adam_beside_stem=os.environ.get("CILRS_BENCH_ADAM_BESIDE_STEM", "0") == "1")

    # synthetic batches, per-rank seed (rank 0's are the ones the CPU arm sees)
    POOL = 4
    host = [tuple(t.pin_memory() for t in b) for b in synthetic_host_batches(BATCH, POOL, 100 + rank)]
    devb = [tuple(t.to(dev) for t in b) for b in host]
    loss_host = torch.zeros(6).pin_memory()
    h2d_bytes = sum(t.numel() * t.element_size() for t in host[0])
    d2h_bytes = loss_host.numel() * 4

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    # launches per step (counted on the first step of the run; a graph replays exactly the same kernels). Its loss - the very
    # first step from the initial weights on batch 0 - is what the CPU arm's first step must reproduce (loss_check below).
    trainer.load_batch(*devb[0])
    c0 = lib.cilrs_launch_count()
    if trainer.graph is None:
        trainer._device_step()
        launches_per_step = int(lib.cilrs_launch_count() - c0)
    else:
        trainer.step()
        launches_per_step = None
    torch.cuda.synchronize(dev)
    loss_first = trainer.read_loss()["total"]

    # W warm-up steps as asked, plus (multi-rank only) a few settle steps: the first replays after start-up pay NCCL's lazy
    # channel / buffer set-up and showed up as +0.3 ms/step in 20-step runs on 2 GPUs
    settle = 10 if world > 1 else 0
    for i in range(args.warmup + settle):
        trainer.load_batch(*devb[i % POOL])
        trainer.step()
    # ---------------- device-resident timing ----------------
    sampler = ClockSampler(local_rank)
    barrier()
    sampler.start()   # every rank samples its own GPU; rank 0 reports the slowest one and the union of the throttle reasons
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    trainer.prefetch_batch(*devb[0])
    for i in range(args.steps):
        trainer.load_prefetched()                        # one D2D copy into the captured step's static inputs
        if i + 1 < args.steps:
            trainer.prefetch_batch(*devb[(i + 1) % POOL])    # (device-resident pool -> staging buffer, on the copy stream)
        trainer.step()
    e1.record()
    barrier()
    ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms_total = float(ms)
    clocks = sampler.stop()
    if world > 1:
        allc = [None] * world
        dist.all_gather_object(allc, clocks)
        if rank == 0:
            ok = [c for c in allc if c and c.get("sm_mhz") is not None]
            if ok:
                slow = min(ok, key=lambda c: c["sm_mhz"])
                clocks = dict(slow)
                clocks["reasons"] = sorted(set(r for c in allc if c for r in c.get("reasons", [])))
                clocks["per_rank_sm_mhz"] = [c.get("sm_mhz") if c else None for c in allc]
    loss_last = trainer.read_loss()["total"]

    # ---------------- end-to-end timing: host frames in, loss out, every step ----------------
    # The input pipeline a caller builds with the public API: while step i runs, the H2D copies of batch i + 1 are already in
    # flight on the trainer's copy stream (prefetch_batch), the way the reference's DataLoader(num_workers=2, pin_memory=True)
    # prefetches. Every timed step still contains one full H2D batch copy and the D2H read of its own loss.
    for i in range(2):
        trainer.load_batch(*host[i % POOL])
        trainer.step()
    barrier()
    e2, e3 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e2.record()
    trainer.prefetch_batch(*host[0])                  # H2D from pinned memory (inside the timed region)
    for i in range(args.steps):
        trainer.load_prefetched()                     # batch i: staged -> step inputs (one D2D copy)
        if i + 1 < args.steps:
            trainer.prefetch_batch(*host[(i + 1) % POOL])   # batch i + 1 crosses PCIe under step i
        loss6 = trainer.step()
        loss_host.copy_(loss6, non_blocking=True)     # D2H of the step's result
        torch.cuda.current_stream().synchronize()     # the caller reads the loss every step (reference: .item())
    e3.record()
    barrier()
    ms2 = torch.tensor([e2.elapsed_time(e3)], device=dev)
    if world > 1:
        dist.all_reduce(ms2, op=dist.ReduceOp.MAX)
    ms_e2e = float(ms2)

    frames = BATCH * world * args.steps
    value = frames / (ms_total * 1e-3)
    e2e_value = frames / (ms_e2e * 1e-3)

    # ---- one profiled eager step (CUDA events around every launch of the plan). With N > 1 the step contains the gradient
    # allreduce, so EVERY rank runs it (a collective issued by rank 0 alone would hang); only rank 0 collects the timings.
    torch.cuda.synchronize(dev)
    if rank == 0:
        _lib.call("cilrs_model_profile", model._handle, 1)
    trainer.load_batch(*devb[0])
    c0 = lib.cilrs_launch_count()
    trainer._device_step()
    eager_launches = int(lib.cilrs_launch_count() - c0)
    if launches_per_step is None:
        launches_per_step = eager_launches
    torch.cuda.synchronize(dev)
    out_ms = (ctypes.c_float * 7)()
    out_n = (ctypes.c_int * 7)()
    if rank == 0:
        _lib.call("cilrs_model_profile_collect", model._handle, out_ms, out_n)
        _lib.call("cilrs_model_profile", model._handle, 0)
    # the inference measurements below use their OWN module, loaded from the trained state_dict like a rollout worker would
    # (a 512-frame session on the training module would rebuild its plan under the trainer's feet)
    infer_model = CILRS(num_commands=4, dropout=0.0)
    infer_model.load_state_dict({k: v.detach().cpu() for k, v in model.state_dict().items()}, strict=True)
    infer_model = infer_model.to(dev).eval()
    line = {
        "metric": METRIC, "value": value, "unit": "frames/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_total / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "bf16",
        "data": "synthetic",
        "config": {"workload": WORKLOAD, "per_gpu_batch": BATCH, "global_batch": BATCH * world, "parallelism": "dp%d" % world,
                   "cuda_graph": trainer.graph is not None, "cuda_graph_error": trainer.graph_error, "settle_steps": settle,
                   "initial_weights": "%s under torch.manual_seed(0), loaded via load_state_dict (the CPU leg loads the same state_dict)" % init_kind,
                   "allreduce_schedule": trainer.overlap_allreduce if world > 1 else None,
                   "grad_comm": trainer.grad_comm if world > 1 else None,
                   "async_parts": trainer.async_parts if world > 1 else None,
                   "nccl_max_ctas": os.environ.get("NCCL_MAX_CTAS") if world > 1 else None,
                   "l2": "no explicit flush: one step touches ~0.9 GB of activations + 0.6 GB of optimizer state, > 126 MB L2"},
        "e2e": {"value": e2e_value, "unit": "frames/s", "h2d_bytes_per_step": h2d_bytes, "d2h_bytes_per_step": d2h_bytes,
                "ms_per_step": ms_e2e / args.steps, "input": "uint8 [128,88,200,3] frames + speed/command/targets from pinned host memory",
                "pipeline": "K H2D batch copies and K loss reads inside the timed region; batch i+1's copy is issued before step i "
                            "(FusedTrainer.prefetch_batch / load_prefetched), the loss of step i is on the host before step i+1 starts"},
        "gpu_launches": launches_per_step * args.steps,
        "loss_first": loss_first, "loss_last": loss_last,
    }
    # ---- configs[4]: batched inference sharded over the ranks (512 frames per GPU), all ranks ----
    try:
        from cilrs_b200.preprocess import ShardedInference
        PER = 512
        sh = ShardedInference(infer_model, PER * world, src_hw=(88, 200))   # frames already 200x88 (the rollout's resized stream)
        gi = torch.Generator().manual_seed(7 + rank)
        sh.session.h_frames.copy_(torch.randint(0, 256, tuple(sh.session.h_frames.shape), generator=gi, dtype=torch.uint8))
        sh.session.h_speed.copy_(torch.rand(PER, generator=gi))
        sh.session.h_command.copy_(torch.randint(0, 4, (PER,), generator=gi))
        for _ in range(3):
            sh.session.run()
        barrier()
        reps = 20
        t0 = time.perf_counter()
        for _ in range(reps):
            sh.session.run()           # H2D of the shard's frames, K0 normalise, eval forward, D2H of [512,4]
        torch.cuda.synchronize(dev)
        tloc = torch.tensor([(time.perf_counter() - t0) / reps], device=dev, dtype=torch.float64)
        s = sh.session
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        with torch.cuda.stream(s.stream):
            ev0.record()
            for _ in range(reps):
                s.graph.replay() if s.graph is not None else s._body()
            ev1.record()
        s.stream.synchronize()
        tdev = torch.tensor([ev0.elapsed_time(ev1) * 1e-3 / reps], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(tloc, op=dist.ReduceOp.MAX)
            dist.all_reduce(tdev, op=dist.ReduceOp.MAX)
        line["infer_c5"] = {"workload": "batched CILRS inference, %d frames (512 per GPU x %d), mixed commands, bf16 eval" % (PER * world, world),
                            "frames_per_s": PER * world / float(tdev), "ms_per_batch": float(tdev) * 1e3,
                            "conv_TFLOPs_per_gpu": PER * FLOP_FWD / float(tdev) / 1e12,
                            "e2e_frames_per_s": PER * world / float(tloc),
                            "e2e_span": "uint8 frames in pinned host memory -> [512,4] controls on the host, per rank; max over ranks"}
        del sh
    except Exception as ex:  # never lose the training line to the extra measurement
        line["infer_c5"] = {"error": repr(ex)}
    if rank == 0:
        pk = peaks()
        names = ["conv_fprop", "conv_dgrad", "conv_wgrad", "bn_forward_pool", "bn_backward", "heads", "other"]
        breakdown = {n: {"ms": round(float(out_ms[i]), 4), "launches": int(out_n[i])} for i, n in enumerate(names)}
        t_gemm = (out_ms[0] + out_ms[1]) * 1e-3
        n_gemm = out_n[0] + out_n[1]
        achieved = BATCH * (FLOP_FWD + FLOP_DGRAD) / t_gemm / 1e12 if t_gemm > 0 else 0.0
        burst, sustained = pk["bf16_tflops"], pk["bf16_tflops_sustained"]
        traffic, traffic_note = None, None
        for tname in ("r02_conv_flat_traffic.json", "r01_conv_flat_traffic.json"):   # the committed ncu --set full capture
            try:
                tj = json.load(open(os.path.join(ROOT, "profiles", tname)))
                traffic, traffic_note = tj["traffic_bytes_per_launch"], tj["launches"] + " | " + tj["source"]
                break
            except Exception:
                pass
        step_tf = value / world * FLOP_TRAIN / 1e12
        line["roofline"] = {
            "bound": "tensor",
            "kernel": "conv_flat_kernel + conv_gemm_kernel (tcgen05 implicit-GEMM fprop + dgrad with fused BN statistics / ReLU mask / "
                      "BN-backward reductions, %d launches/step)" % n_gemm,
            "achieved": achieved, "peak": burst, "unit": "TFLOP/s", "frac": achieved / burst, "traffic": traffic,
            "traffic_note": traffic_note,
            "peak_source": pk["source"] + " bf16_tflops (burst; primary per SURVEY.md 8d)",
            "frac_sustained": achieved / sustained, "peak_sustained": sustained,
            "avg_launch_ms": (out_ms[0] + out_ms[1]) / max(1, n_gemm),
            "wgrad_kernel": {"achieved": BATCH * FLOP_WGRAD / (out_ms[2] * 1e-3) / 1e12 if out_ms[2] > 0 else 0.0,
                             "frac": (BATCH * FLOP_WGRAD / (out_ms[2] * 1e-3) / 1e12 / burst) if out_ms[2] > 0 else 0.0,
                             "note": "wgrad_flat_kernel + wgrad_reduce_kernel + wgrad_gemm_kernel; in the timed step they run on a side "
                                     "stream concurrently with the dgrad / BatchNorm chain"},
            "step_level": {"achieved": step_tf, "frac": step_tf / burst, "frac_sustained": step_tf / sustained,
                           "note": "whole-step frames/s x 8.305 GFLOP conv work per frame (SURVEY.md 8d), all GEMM + non-GEMM time included"},
            "breakdown_ms": breakdown,
        }
        line["clocks"] = clocks
        # ---- batch-1 inference latency (the other half of BASELINE.json's metric) ----
        try:
            from cilrs_b200.preprocess import InferenceSession
            sess = InferenceSession(infer_model, batch=1)
            sess.h_frames.random_(0, 256)
            sess.h_speed.fill_(0.3)
            sess.h_command.fill_(1)
            lat = []
            for i in range(1020):
                t0 = time.perf_counter()
                sess.run()
                lat.append((time.perf_counter() - t0) * 1e3)
            lat = sorted(lat[20:])
            line["infer_b1"] = {"p50_ms": lat[len(lat) // 2], "p90_ms": lat[int(len(lat) * 0.9)], "iterations": len(lat),
                                "span": "raw uint8 600x800x3 frame in pinned host memory -> (steer, throttle, brake, speed) on the host"}
            del sess
        except Exception as ex:
            line["infer_b1"] = {"error": repr(ex)}
        if not args.no_extras:
            line.update(extra_kernel_lines(torch, infer_model, pk))
        # ---- CPU baseline (the reference's class when oracle/_ref was built, else the oracle port), bounded sample ----
        check_failed = None
        if world == 1 and not args.no_cpu_baseline:
            fps, sec, cores, kind, ref_first = cpu_reference_step_rate(BATCH, 12, 1, state_dict=init_sd,
                                                                       batches=[tuple(t.clone() for t in b) for b in host[:2]])
            line["cpu_baseline"] = {"value": fps, "unit": "frames/s", "cores": cores, "kind": kind,
                                    "sample": "12 timed full train steps (fwd + MSE loss + bwd + torch.optim.Adam) at batch 128 after 1 warm-up, "
                                              "torch fp32 on the host, %d threads" % cores}
            rel = abs(loss_first - ref_first) / abs(ref_first)
            line["loss_check"] = {"ours_first_step": loss_first, "cpu_first_step": ref_first, "rel": rel, "bar": 2e-2, "ok": rel <= 2e-2,
                                  "note": "same initial weights, same batch 0: the CUDA step's loss vs the CPU arm's"}
            if rel > 2e-2:
                check_failed = "bench.py: first-step loss %.6f differs from the CPU reference's %.6f by %.2e (> 2e-2)" % (loss_first, ref_first, rel)
        print(json.dumps(line))
        sys.stdout.flush()
        if check_failed:
            raise SystemExit(check_failed)
    if world > 1:
        # orderly teardown: drop the captured graph (it references the communicator) before the process group goes away. A
        # watchdog keeps a wedged NCCL teardown from turning a finished measurement into a hung job (the line is already out).
        import threading
        threading.Timer(60.0, lambda: os._exit(0)).start()
        trainer.close()
        torch.cuda.synchronize(dev)
        dist.barrier()
        dist.destroy_process_group()
        os._exit(0)


if __name__ == "__main__":
    main()
