#!/usr/bin/env python
"""bench.py — CILRS training-step throughput on N B200s (BASELINE.json: "train frames/s (bf16, bs128/GPU)").

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

N > 1 is launched by torchrun (one rank per GPU, NCCL); rank 0 prints ONE JSON line.

Workload (BASELINE.json configs[1] / [3]): CILRS train step, batch 128 per GPU, 200x88 synthetic frames, MSE(controls) +
0.05 MSE(speed), Adam lr 2e-4 wd 1e-4, train-mode BatchNorm, dropout 0, random-init weights (reference initialisers), bf16
tensor-core convs with fp32 accumulate, fp32 master weights / heads / optimizer. A step = forward + loss + zero_grad + backward
(+ gradient allreduce) + Adam + bf16 operand repack; nothing is skipped or cached.

  value : frames/s, inputs already in HBM (a rotating pool of device batches), CUDA events over K steps, max over ranks
  e2e   : same steps through the public API with HOST (pinned) uint8 frames: H2D copy + on-device normalise (K0) + step +
          D2H read of the loss scalars every step
  roofline     : the implicit-GEMM conv kernels (conv_flat_kernel: 58 of the 77 fprop+dgrad launches, conv_gemm_kernel: stem /
                 stride-2 / 1x1): algorithmic conv FLOPs / time inside those launches (CUDA events around every launch of
                 one eager step; the launches also carry the fused BN statistics / ReLU mask / BN-backward reductions)
  cpu_baseline : the oracle port of the reference step (torch fp32 on the host cores), rank 0 / N=1 only
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


def cpu_reference_step_rate(batch, steps, warmup):
    """The reference's train step restated on the CPU (oracle port: torch fp32, all host threads):
    forward (train-mode BN) + MSE losses + backward + Adam (notebook/notebook.ipynb:545-555 with the BASELINE recipe)."""
    import torch
    from oracle import cilrs_oracle as O
    torch.manual_seed(0)
    sd = O.synthetic_state_dict(0, perturb_bn=False)
    params = {k: v.clone().requires_grad_(True) for k, v in sd.items() if v.is_floating_point() and not k.endswith(("running_mean", "running_var"))}
    state = dict(sd)
    state.update(params)
    opt = torch.optim.Adam(list(params.values()), lr=2e-4, weight_decay=1e-4)
    g = torch.Generator().manual_seed(1)
    image = torch.randn(batch, 3, 88, 200, generator=g)
    speed = torch.rand(batch, generator=g)
    command = torch.randint(0, 4, (batch,), generator=g)
    targets = torch.rand(batch, 3, generator=g)
    times = []
    for i in range(warmup + steps):
        t0 = time.perf_counter()
        upd = {}
        c, p = O.forward(state, image, speed, command, training=True, update=upd)
        loss, _ = O.loss_mse(c, targets, p, speed)
        opt.zero_grad()
        loss.backward()
        opt.step()
        for k, v in upd.items():
            state[k] = v
        if i >= warmup:
            times.append(time.perf_counter() - t0)
    total = sum(times)
    return batch * len(times) / total, total / len(times), torch.get_num_threads()


def run_reference(args, rank):
    """--impl reference: the reference's own CPU implementation of the path. The reference is pure Python on top of torch
    (nothing to compile into oracle/_ref, and /root/reference is absent on the GPU box), so this is the oracle port."""
    if rank != 0:
        return
    sample = 32
    fps, sec, cores = cpu_reference_step_rate(sample, max(1, args.steps), min(args.warmup, 2))
    line = {
        "impl": "reference", "metric": METRIC, "value": fps, "unit": "frames/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": min(args.warmup, 2), "ms_per_step": sec * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD, "per_gpu_batch": BATCH, "global_batch": BATCH * args.gpus, "parallelism": "cpu"},
        "cpu_baseline": {"value": fps, "unit": "frames/s", "cores": cores, "kind": "port",
                         "sample": "each step = one full train step (fwd+loss+bwd+Adam) on a %d-frame batch, torch fp32 CPU" % sample},
        "e2e": {"value": fps, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-graph", action="store_true")
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
        dist.init_process_group("nccl", device_id=dev)
    import cilrs_b200  # noqa: F401
    from cilrs_b200 import _lib
    from cilrs_b200.model import CILRS
    from cilrs_b200.train import FusedTrainer

    lib = _lib.lib()
    lib.cilrs_launch_count.restype = ctypes.c_longlong
    torch.manual_seed(0)
    model = CILRS(num_commands=4, dropout=0.0).to(dev)
    trainer = FusedTrainer(model, BATCH, lr=2e-4, weight_decay=1e-4, loss="mse", speed_w=0.05, frames="u8",
                           use_graph=(not args.no_graph),
                           overlap_allreduce={"tail": False, "all": True}.get(os.environ.get("CILRS_BENCH_ALLREDUCE", "first"), "first"),  # measurement aids
                           async_parts=os.environ.get("CILRS_BENCH_ASYNC_PARTS", "0") == "1")

    # synthetic batches: uint8 200x88 frames (what the reference's dataset stores after prepare_dataset.py), per-rank seed
    g = torch.Generator().manual_seed(100 + rank)
    POOL = 4
    host = []
    for _ in range(POOL):
        host.append((torch.randint(0, 256, (BATCH, 88, 200, 3), generator=g, dtype=torch.uint8).pin_memory(),
                     torch.rand(BATCH, generator=g).pin_memory(), torch.randint(0, 4, (BATCH,), generator=g).pin_memory(),
                     torch.rand(BATCH, 3, generator=g).pin_memory()))
    devb = [tuple(t.to(dev) for t in b) for b in host]
    loss_host = torch.zeros(6).pin_memory()
    h2d_bytes = sum(t.numel() * t.element_size() for t in host[0])
    d2h_bytes = loss_host.numel() * 4

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    # launches per step (counted on one eager step; a graph replays exactly the same kernels)
    trainer.load_batch(*devb[0])
    c0 = lib.cilrs_launch_count()
    trainer._device_step()
    launches_per_step = int(lib.cilrs_launch_count() - c0)
    if trainer.graph is not None:
        trainer.opt._step += 0
    torch.cuda.synchronize(dev)

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
    for i in range(args.steps):
        trainer.load_batch(*devb[i % POOL])
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
    loss_last = trainer.loss6.tolist()

    # ---------------- end-to-end timing: host frames in, loss out, every step ----------------
    for i in range(2):
        trainer.load_batch(*host[i % POOL])
        trainer.step()
    barrier()
    e2, e3 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e2.record()
    for i in range(args.steps):
        trainer.load_batch(*host[i % POOL])          # H2D from pinned memory
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

    line = {
        "metric": METRIC, "value": value, "unit": "frames/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_total / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "bf16",
        "data": "synthetic",
        "config": {"workload": WORKLOAD, "per_gpu_batch": BATCH, "global_batch": BATCH * world, "parallelism": "dp%d" % world,
                   "cuda_graph": trainer.graph is not None, "cuda_graph_error": trainer.graph_error, "settle_steps": settle,
                   "l2": "no explicit flush: one step touches ~0.9 GB of activations + 0.6 GB of optimizer state, > 126 MB L2"},
        "e2e": {"value": e2e_value, "unit": "frames/s", "h2d_bytes_per_step": h2d_bytes, "d2h_bytes_per_step": d2h_bytes,
                "ms_per_step": ms_e2e / args.steps, "input": "uint8 [128,88,200,3] frames + speed/command/targets from pinned host memory"},
        "gpu_launches": launches_per_step * args.steps,
        "loss_last": loss_last[0],
    }
    # ---- one profiled eager step (CUDA events around every launch of the plan). With N > 1 the step contains the gradient
    # allreduce, so EVERY rank runs it (a collective issued by rank 0 alone would hang); only rank 0 collects the timings.
    torch.cuda.synchronize(dev)
    if rank == 0:
        _lib.call("cilrs_model_profile", model._handle, 1)
    trainer.load_batch(*devb[0])
    trainer._device_step()
    torch.cuda.synchronize(dev)
    if rank == 0:
        pk = peaks()
        out_ms = (ctypes.c_float * 7)()
        out_n = (ctypes.c_int * 7)()
        _lib.call("cilrs_model_profile_collect", model._handle, out_ms, out_n)
        _lib.call("cilrs_model_profile", model._handle, 0)
        names = ["conv_fprop", "conv_dgrad", "conv_wgrad", "bn_forward_pool", "bn_backward", "heads", "other"]
        breakdown = {n: {"ms": round(float(out_ms[i]), 4), "launches": int(out_n[i])} for i, n in enumerate(names)}
        t_gemm = (out_ms[0] + out_ms[1]) * 1e-3
        n_gemm = out_n[0] + out_n[1]
        achieved = BATCH * (FLOP_FWD + FLOP_DGRAD) / t_gemm / 1e12 if t_gemm > 0 else 0.0
        peak = pk["bf16_tflops_sustained"]
        traffic, traffic_note = None, None
        try:
            tj = json.load(open(os.path.join(ROOT, "profiles", "r01_conv_flat_traffic.json")))
            traffic, traffic_note = tj["traffic_bytes_per_launch"], tj["launches"] + " | " + tj["source"]
        except Exception:
            pass
        line["roofline"] = {
            "bound": "tensor",
            "kernel": "conv_flat_kernel + conv_gemm_kernel (tcgen05 implicit-GEMM fprop + dgrad with fused BN statistics / ReLU mask / "
                      "BN-backward reductions, %d launches/step)" % n_gemm,
            "achieved": achieved, "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak, "traffic": traffic,
            "traffic_note": traffic_note,
            "peak_source": pk["source"] + " bf16_tflops_sustained (kernel timed inside a long step)",
            "avg_launch_ms": (out_ms[0] + out_ms[1]) / max(1, n_gemm),
            "wgrad_kernel": {"achieved": BATCH * FLOP_WGRAD / (out_ms[2] * 1e-3) / 1e12 if out_ms[2] > 0 else 0.0,
                             "frac": (BATCH * FLOP_WGRAD / (out_ms[2] * 1e-3) / 1e12 / peak) if out_ms[2] > 0 else 0.0,
                             "note": "wgrad_flat_kernel + wgrad_reduce_kernel + wgrad_gemm_kernel; in the timed step they run on a side "
                                     "stream concurrently with the dgrad / BatchNorm chain"},
            "step_level": {"achieved": value / world * FLOP_TRAIN / 1e12, "frac_burst": value / world * FLOP_TRAIN / 1e12 / pk["bf16_tflops"],
                           "frac_sustained": value / world * FLOP_TRAIN / 1e12 / peak,
                           "note": "whole-step frames/s x 8.305 GFLOP conv work per frame (SURVEY.md §8d)"},
            "breakdown_ms": breakdown,
        }
        line["clocks"] = clocks
        # ---- batch-1 inference latency (the other half of BASELINE.json's metric) ----
        try:
            from cilrs_b200.preprocess import InferenceSession
            model.eval()
            sess = InferenceSession(model, batch=1)
            sess.h_frames.random_(0, 256)
            sess.h_speed.fill_(0.3)
            sess.h_command.fill_(1)
            lat = []
            for i in range(320):
                t0 = time.perf_counter()
                sess.run()
                lat.append((time.perf_counter() - t0) * 1e3)
            lat = sorted(lat[20:])
            line["infer_b1"] = {"p50_ms": lat[len(lat) // 2], "p90_ms": lat[int(len(lat) * 0.9)],
                                "span": "raw uint8 600x800x3 frame in pinned host memory -> (steer, throttle, brake, speed) on the host"}
        except Exception as ex:  # never lose the training line to the extra measurement
            line["infer_b1"] = {"error": repr(ex)}
        # ---- CPU baseline (oracle port), bounded sample ----
        if world == 1 and not args.no_cpu_baseline:
            fps, sec, cores = cpu_reference_step_rate(BATCH, 12, 1)   # ~10 s of host work
            line["cpu_baseline"] = {"value": fps, "unit": "frames/s", "cores": cores, "kind": "port",
                                    "sample": "12 timed full train steps (fwd+MSE loss+bwd+Adam) at batch 128 after 1 warm-up, torch fp32 on the host"}
        print(json.dumps(line))
        sys.stdout.flush()
    if world > 1:
        # Leave without tearing NCCL down: with the allreduces captured in a CUDA graph, destroy_process_group() (and a
        # barrier issued while rank 0 is still busy with its single-GPU extras) can block on communicator resources the
        # graph still references. Every rank has finished its collectives at this point; exit code 0.
        torch.cuda.synchronize(dev)
        sys.stderr.flush()
        os._exit(0)


if __name__ == "__main__":
    main()
