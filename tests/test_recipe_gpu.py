"""Parity of the product paths a user actually calls (missing #6-#7): the whole model at the bench
batch, the bf16-emulating oracle, the FusedTrainer trajectory (eager and CUDA graph), the reference's executed recipe on the
graph path (StepLR, clip_grad_norm_, dropout), predict_controls / InferenceSession, optimizer checkpoints, validate()."""
import ctypes
import math
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden")


def _O():
    from oracle import cilrs_oracle as O
    return O


def _rel(a, b):
    a, b = a.double().cpu(), b.double().cpu()
    return float((a - b).abs().max() / (b.abs().max() + 1e-30))


def _rel_l2(a, b):
    a, b = a.double().cpu(), b.double().cpu()
    return float((a - b).norm() / (b.norm() + 1e-300))


def _model(sd, train=False, dropout=0.0):
    from cilrs_b200.model import CILRS
    m = CILRS(num_commands=4, dropout=dropout)
    m.load_state_dict(sd, strict=True)
    m = m.to("cuda")
    m.train(train)
    return m


def _inputs(B, seed, device="cpu"):
    g = torch.Generator().manual_seed(seed)
    coarse = torch.randn(B, 3, 11, 25, generator=g)
    image = torch.nn.functional.interpolate(coarse, size=(88, 200), mode="bicubic", align_corners=False).contiguous()
    speed = torch.rand(B, generator=g)
    command = torch.randint(0, 4, (B,), generator=g)
    command[:4] = torch.tensor([0, 1, 2, 3])
    targets = torch.stack([torch.rand(B, generator=g) * 2 - 1, torch.rand(B, generator=g), torch.rand(B, generator=g)], dim=1)
    return tuple(t.to(device) for t in (image, speed, command, targets))


def _leaf_sd(sd, dtype, device):
    out = {}
    for k, v in sd.items():
        if not v.is_floating_point():
            out[k] = v.to(device)
        elif k.endswith("running_mean") or k.endswith("running_var"):
            out[k] = v.to(device=device, dtype=dtype)
        else:
            out[k] = v.to(device=device, dtype=dtype).clone().requires_grad_(True)
    return out


# ------------------------------------------------------------------------------------------------------------------
# whole model at the bench batch
# ------------------------------------------------------------------------------------------------------------------
def test_whole_model_b128_eval_forward_and_frozen_bn_gradients():
    """B = 128 (the benchmarked tile configuration): eval forward 2e-2 vs the fp64 oracle; frozen-BN gradients
    <= max(2e-2, 1.5 x the reference's own bf16-autocast error) globally."""
    O = _O()
    sd = O.synthetic_state_dict(0)
    image, speed, command, targets = _inputs(128, 5, "cuda")
    m = _model(sd, train=False)
    with torch.no_grad():
        c, p = m(image, speed, command)
    sd64 = _leaf_sd(sd, torch.float64, "cuda")
    c64, p64 = O.forward(sd64, image.double(), speed.double(), command, training=False)
    ec, ep = _rel(c, c64), _rel(p, p64)
    print("B=128 eval forward: controls %.3e, speed %.3e" % (ec, ep))
    assert ec <= 2e-2 and ep <= 2e-2
    # frozen-BN (eval-mode autograd) gradients
    tot64, _ = O.loss_mse(c64, targets.double(), p64, speed.double())
    tot64.backward()
    c, p = m(image, speed, command)
    tot, _ = O.loss_mse(c, targets, p, speed)
    m.zero_grad()
    tot.backward()
    torch.cuda.synchronize()
    names = [n for n, _ in m.named_parameters()]
    got = torch.cat([q.grad.double().reshape(-1) for _, q in m.named_parameters()])
    ref = torch.cat([sd64[n].grad.reshape(-1) for n in names])
    glob = float((got - ref).norm() / ref.norm())
    sdr = _leaf_sd(sd, torch.float32, "cuda")
    with torch.autocast("cuda", dtype=torch.bfloat16):
        cr, pr = O.forward(sdr, image, speed, command, training=False)
    O.loss_mse(cr.float(), targets, pr.float(), speed)[0].backward()
    noise = float((torch.cat([sdr[n].grad.double().reshape(-1) for n in names]) - ref).norm() / ref.norm())
    print("B=128 frozen-BN: loss %.6f vs %.6f, global grad err ours %.3e, reference bf16-autocast %.3e" % (float(tot), float(tot64), glob, noise))
    assert abs(float(tot) - float(tot64)) <= 2e-2 * abs(float(tot64))
    assert glob <= max(2e-2, 1.5 * noise)


@pytest.mark.parametrize("B", [16, 128])
def test_train_mode_gradients_whole_model_report(B):
    """Train-mode (batch-statistics) gradients of the WHOLE model. Measured on B200 (round 2): trunk gradient 0.53-0.58 from the
    fp64 oracle with bf16 rounding emulated at the CUDA path's storage points (oracle.forward_bf16emu), 0.66-0.73 from plain
    fp64, while the reference's own bf16 autocast is 0.5-0.7 from fp64 - at random init the 36 chained train-mode BatchNorm
    backwards amplify ANY perturbation by ~1e5 (the reference's fp32 is already 2e-3..8e-3 off), so no whole-model
    bound on the trunk can have teeth; tests/test_blocks_b128_gpu.py pins every block, every layer group and the 16-block chain
    instead. Asserted here: the loss (2e-2), the head gradients and the global trunk figure against the reference's own bf16
    noise measured in the same test; the emulated-oracle numbers are printed for the record."""
    O = _O()
    sd = O.synthetic_state_dict(0)
    image, speed, command, targets = _inputs(B, 31, "cuda")
    m = _model(sd, train=True)
    c, p = m(image, speed, command)
    tot, _ = O.loss_mse(c, targets, p, speed)
    m.zero_grad()
    tot.backward()
    torch.cuda.synchronize()
    names = [n for n, _ in m.named_parameters()]
    got = {n: q.grad.double() for n, q in m.named_parameters()}

    def run(fwd, sdx, img, dt):
        cx, px = fwd(sdx, img, speed.to(dt), command, training=True)
        t, _ = O.loss_mse(cx.to(dt), targets.to(dt), px.to(dt), speed.to(dt))
        t.backward()
        return float(t)

    sd_emu = O.bf16_weights(_leaf_sd(sd, torch.float64, "cuda"))
    t_emu = run(O.forward_bf16emu, sd_emu, image.double(), torch.float64)
    sd_64 = _leaf_sd(sd, torch.float64, "cuda")
    t_64 = run(O.forward, sd_64, image.double(), torch.float64)
    sd_ac = _leaf_sd(sd, torch.float32, "cuda")
    with torch.autocast("cuda", dtype=torch.bfloat16):
        cr, pr = O.forward(sd_ac, image, speed, command, training=True)
    O.loss_mse(cr.float(), targets, pr.float(), speed)[0].backward()

    def glob(a, ref, sel):
        x = torch.cat([a(n).reshape(-1) for n in names if sel(n)])
        y = torch.cat([ref[n].grad.double().reshape(-1) for n in names if sel(n)])
        return float((x - y).norm() / y.norm())

    trunk = lambda n: n.startswith("visual_encoder")
    heads = lambda n: not n.startswith("visual_encoder")
    ours = lambda n: got[n]
    autoc = lambda n: sd_ac[n].grad.double()
    e_emu_t, e_emu_h = glob(ours, sd_emu, trunk), glob(ours, sd_emu, heads)
    e_64_t, e_64_h = glob(ours, sd_64, trunk), glob(ours, sd_64, heads)
    n_64_t, n_64_h = glob(autoc, sd_64, trunk), glob(autoc, sd_64, heads)
    print("B=%d train-mode: loss ours %.6f, emu %.6f, fp64 %.6f | trunk grad err vs emu %.3e, vs fp64 %.3e (reference bf16-autocast "
          "vs fp64 %.3e) | head grad err vs emu %.3e, vs fp64 %.3e (reference bf16-autocast %.3e)"
          % (B, float(tot), t_emu, t_64, e_emu_t, e_64_t, n_64_t, e_emu_h, e_64_h, n_64_h))
    assert abs(float(tot) - t_emu) <= 2e-2 * abs(t_emu) and abs(float(tot) - t_64) <= 2e-2 * abs(t_64)
    assert e_64_h <= max(2e-2, 1.5 * n_64_h)
    assert e_64_t <= max(2e-2, 1.5 * n_64_t)


# ------------------------------------------------------------------------------------------------------------------
# FusedTrainer: the benchmarked path
# ------------------------------------------------------------------------------------------------------------------
def _oracle_training_run(O, sd, frames_u8, speed, command, targets, steps, lr, loss, clip, wd=1e-4, emulate_bf16=False):
    """The reference's train step (notebook/notebook.ipynb:545-555) in fp32 on the CPU: forward (train-mode BN), loss,
    backward, optional clip_grad_norm_, torch.optim.Adam. emulate_bf16: same run with the bf16 rounding points of the CUDA path
    (oracle.forward_bf16emu) - the reference algorithm's own sensitivity to bf16 storage."""
    image = torch.from_numpy(O.normalise_np(frames_u8.numpy()))
    state = {k: v.clone() for k, v in sd.items()}
    params = {k: v.requires_grad_(True) for k, v in state.items() if v.is_floating_point() and "running" not in k}
    opt = torch.optim.Adam(list(params.values()), lr=lr, weight_decay=wd)
    lossfn = O.loss_mse if loss == "mse" else O.loss_l1
    losses = []
    for _ in range(steps):
        upd = {}
        if emulate_bf16:
            st = O.bf16_weights(state)
            c, p = O.forward_bf16emu(st, image, speed, command, training=True, update=upd)
        else:
            c, p = O.forward(state, image, speed, command, training=True, update=upd)
        tot, _ = lossfn(c, targets, p, speed)
        opt.zero_grad()
        tot.backward()
        if emulate_bf16:
            for k, v in st.items():
                if v.dim() == 4:
                    state[k].grad = v.grad
        if clip > 0:
            torch.nn.utils.clip_grad_norm_(list(params.values()), clip)
        opt.step()
        state.update(upd)
        losses.append(float(tot.detach()))
    return losses, state


@pytest.mark.parametrize("use_graph", [False, True])
@pytest.mark.parametrize("recipe", ["mse", "l1_clip"])
def test_fused_trainer_trajectory_matches_oracle(use_graph, recipe):
    """5 steps of FusedTrainer (uint8 frames in; eager and CUDA graph) against the oracle's fp32 run of the reference step at the
    BASELINE learning rate. 'l1_clip' is the notebook's executed recipe: L1x(5,1,1) + 0.5 MSE, clip_grad_norm_(1.0)
    (notebook.ipynb:504-527,553-554). Bar per step: 2e-2 relative for the first two steps (a forward pass and one update), then
    6e-2 - the loss falls 2x per step here, so a 3 % difference in effective step size is a 3 % loss difference; measured on B200:
    <= 3.3e-2 (mse), <= 5.4e-2 (l1+clip), while the REFERENCE algorithm with its tensors stored in bf16 (the bf16-emulating oracle,
    run in this test) deviates by up to 0.11 - or 2x that deviation where it is larger."""
    from cilrs_b200.train import FusedTrainer
    O = _O()
    B, steps, lr = 16, 5, 2e-4
    sd = O.synthetic_state_dict(0)
    g = torch.Generator().manual_seed(12)
    coarse = torch.rand(B, 3, 11, 25, generator=g) * 255
    frames = torch.nn.functional.interpolate(coarse, size=(88, 200), mode="bicubic", align_corners=False).clamp(0, 255)
    frames = frames.permute(0, 2, 3, 1).round().to(torch.uint8).contiguous()
    _, speed, command, targets = _inputs(B, 13)
    loss, clip = ("mse", 0.0) if recipe == "mse" else ("l1", 1.0)
    ref_losses, ref_state = _oracle_training_run(O, sd, frames, speed, command, targets, steps, lr, loss, clip)
    emu_losses, _ = _oracle_training_run(O, sd, frames, speed, command, targets, steps, lr, loss, clip, emulate_bf16=True)
    m = _model(sd, train=True)
    tr = FusedTrainer(m, B, lr=lr, weight_decay=1e-4, loss=loss, speed_w=0.05 if loss == "mse" else 0.5, grad_clip=clip,
                      use_graph=use_graph, frames="u8")
    assert (tr.graph is not None) == use_graph
    tr.load_batch(frames.cuda(), speed.cuda(), command.cuda(), targets.cuda())
    ours = []
    for _ in range(steps):
        tr.step()
        ours.append(tr.read_loss()["total"])
    print("trajectory %s graph=%s: ours %s | oracle fp32 %s | oracle bf16-emulated %s"
          % (recipe, use_graph, ["%.5f" % v for v in ours], ["%.5f" % v for v in ref_losses], ["%.5f" % v for v in emu_losses]))
    assert ref_losses[-1] < 0.5 * ref_losses[0], "the test must see the loss move"
    for i, (a, b, e) in enumerate(zip(ours, ref_losses, emu_losses)):
        assert abs(a - b) <= max((2e-2 if i < 2 else 6e-2) * abs(b), 2.0 * abs(e - b)), (i, ours, ref_losses, emu_losses)
    # the update direction of the (well-conditioned) head parameters follows the oracle's
    new = m.state_dict()
    num = den_a = den_b = 0.0
    for k, v in sd.items():
        if v.is_floating_point() and not k.startswith("visual_encoder"):
            da = (new[k].cpu().double() - v.double()).reshape(-1)
            db = (ref_state[k].detach().double() - v.double()).reshape(-1)
            num += float(da @ db); den_a += float(da @ da); den_b += float(db @ db)
    cos = num / math.sqrt(den_a * den_b)
    print("   head-parameter update cosine vs oracle: %.4f" % cos)
    assert cos >= 0.8
    # running statistics and the step counter went through the same number of updates
    assert int(new["visual_encoder.1.num_batches_tracked"]) == steps
    # (after five updates of slightly different weights: measured 2.3e-2 mse / 4.6e-2 l1+clip)
    assert _rel(new["visual_encoder.1.running_mean"], ref_state["visual_encoder.1.running_mean"]) <= 1e-1
    assert float(tr.opt.state_dict()["state"][0]["step"]) == steps


def test_graph_path_follows_the_lr_schedule():
    """StepLR on a captured graph (notebook.ipynb:535-536,604): lr lives in device memory, so param_groups[0]['lr'] edits made
    between replays take effect. lr = 0 must freeze the parameters; restoring it must move them again."""
    from cilrs_b200.train import FusedTrainer
    O = _O()
    sd = O.synthetic_state_dict(0)
    image, speed, command, targets = _inputs(8, 3, "cuda")
    m = _model(sd, train=True)
    tr = FusedTrainer(m, 8, lr=1e-3, weight_decay=0.0, use_graph=True)
    sched = torch.optim.lr_scheduler.StepLR(tr.opt, step_size=2, gamma=0.0)   # lr -> 0 after two scheduler steps
    tr.load_batch(image, speed, command, targets)
    snaps = []
    for _ in range(4):
        tr.step()
        sched.step()
        torch.cuda.synchronize()
        snaps.append(m.flat_parameters().clone())
    assert not torch.equal(snaps[0], snaps[1])        # lr = 1e-3
    assert torch.equal(snaps[2], snaps[1]) and torch.equal(snaps[3], snaps[2])   # lr = 0 since the second scheduler step
    tr.opt.param_groups[0]["lr"] = 1e-3
    tr.step()
    torch.cuda.synchronize()
    assert not torch.equal(m.flat_parameters(), snaps[3])


def test_prefetched_batches_equal_directly_loaded_ones():
    """prefetch_batch / load_prefetched (the next batch's H2D copies on a copy stream under the current step, then one D2D copy)
    must feed the step exactly what load_batch does: identical device inputs at every one of four steps on four different
    host batches and the same first loss, graph and eager, uint8 frames and normalised images."""
    from cilrs_b200.train import FusedTrainer
    O = _O()
    sd = O.synthetic_state_dict(0)
    g = torch.Generator().manual_seed(21)
    for frames, use_graph in (("u8", True), ("f32", False)):
        host = []
        for _ in range(4):
            img = (torch.randint(0, 256, (8, 88, 200, 3), generator=g, dtype=torch.uint8) if frames == "u8"
                   else torch.randn(8, 3, 88, 200, generator=g))
            host.append(tuple(t.pin_memory() for t in (img, torch.rand(8, generator=g), torch.randint(0, 4, (8,), generator=g),
                                                       torch.rand(8, 3, generator=g))))
        runs = []
        for prefetch in (False, True):
            m = _model(sd, train=True)
            tr = FusedTrainer(m, 8, lr=1e-3, weight_decay=1e-4, eps=1e-3, use_graph=use_graph, frames=frames)
            losses = []
            if prefetch:
                tr.prefetch_batch(*host[0])
            for i in range(4):
                if prefetch:
                    tr.load_prefetched()
                    if i + 1 < 4:
                        tr.prefetch_batch(*host[i + 1])
                else:
                    tr.load_batch(*host[i])
                staged = (tr.d_frames if frames == "u8" else tr.d_image, tr.d_speed, tr.d_command, tr.d_targets)
                assert all(torch.equal(d.cpu(), h) for d, h in zip(staged, host[i])), (frames, prefetch, i)
                tr.step()
                losses.append(tr.read_loss()["total"])
            torch.cuda.synchronize()
            runs.append((losses,))
        # Same inputs (asserted above, every step) -> same first loss. Later losses only agree loosely, with or without
        # prefetching: the split-K weight gradients add in L2 in arrival order (fp32 rounding noise), and at random init the 36
        # chained train-mode BatchNorms amplify a 1e-7 parameter difference ~650x per step (tools/prefetch_check.py: two
        # identical runs differ by 2e-4 in the second loss and 1e-2 in the fourth).
        assert abs(runs[0][0][0] - runs[1][0][0]) <= 1e-6 * abs(runs[0][0][0]), (frames, runs[0][0], runs[1][0])
        assert max(abs(a - b) / abs(b) for a, b in zip(*[r[0] for r in runs])) < 8e-2, (frames, runs[0][0], runs[1][0])
        assert len(set(round(v, 6) for v in runs[0][0])) == 4          # four different batches did go through
    with pytest.raises(RuntimeError):
        tr.load_prefetched()                      # nothing staged


def test_graph_path_runs_with_dropout_and_draws_new_masks():
    """dropout = 0.5 (notebook.ipynb:480) under use_graph: the mask counter is the device step counter, so every replay draws a
    different mask; keep-rate and 1/(1-p) scaling are checked on the saved activations."""
    from cilrs_b200.train import FusedTrainer
    O = _O()
    sd = O.synthetic_state_dict(0)
    image, speed, command, targets = _inputs(32, 4, "cuda")
    m = _model(sd, train=True, dropout=0.5)
    tr = FusedTrainer(m, 32, lr=0.0, weight_decay=0.0, use_graph=True)   # lr 0: only the masks differ between steps
    tr.load_batch(image, speed, command, targets)
    outs = []
    for _ in range(3):
        tr.step()
        torch.cuda.synchronize()
        outs.append(tr.controls.clone())
    assert not torch.equal(outs[0], outs[1]) and not torch.equal(outs[1], outs[2])
    assert torch.isfinite(tr.loss6).all()


def test_fused_clip_matches_clip_grad_norm():
    """cilrs_grad_sumsq + the clip coefficient folded into the fused Adam == torch.nn.utils.clip_grad_norm_(…, 1.0) followed
    by torch.optim.Adam (notebook.ipynb:553-555), on the model's real arena layout."""
    from cilrs_b200 import _lib
    from cilrs_b200.optim import FusedAdam
    O = _O()
    sd = O.synthetic_state_dict(0)
    m = _model(sd, train=True)
    opt = FusedAdam(m.parameters(), lr=1e-3, weight_decay=1e-4, model=m)
    gflat = m.flat_gradients()
    gen = torch.Generator(device="cuda").manual_seed(8)
    views = m._views(gflat)
    for p, v in zip(m.parameters(), views):
        v.copy_(torch.randn(v.shape, generator=gen, device="cuda") * 1e-3)
        p.grad = v
    # reference: torch's own clip + Adam on copies
    ref_params = [p.detach().clone().requires_grad_(True) for p in m.parameters()]
    for rp, v in zip(ref_params, views):
        rp.grad = v.clone()
    ref_norm = torch.nn.utils.clip_grad_norm_(ref_params, 1.0)
    ropt = torch.optim.Adam(ref_params, lr=1e-3, weight_decay=1e-4)
    ropt.step()
    ws = torch.zeros(1024, dtype=torch.float64, device="cuda")
    cnt = torch.zeros(4, dtype=torch.int32, device="cuda")
    out2 = torch.zeros(2, dtype=torch.float32, device="cuda")
    _lib.call("cilrs_grad_sumsq", gflat, ctypes.c_longlong(gflat.numel()), ws, cnt, ctypes.c_float(1.0), out2, _lib.stream_ptr())
    opt.step(grad_scale_dev=out2[1:], grads_in_arena=True)
    torch.cuda.synchronize()
    assert abs(math.sqrt(float(out2[0])) - float(ref_norm)) <= 1e-5 * float(ref_norm)
    coef = min(1.0, 1.0 / (float(ref_norm) + 1e-6))
    assert float(ref_norm) > 1.0 and abs(float(out2[1]) - coef) <= 1e-6
    worst = max(float((p.detach() - rp.detach()).abs().max()) for p, rp in zip(m.parameters(), ref_params))
    print("fused clip: norm %.6f (torch %.6f), coefficient %.6f, max |p - torch| %.3e" % (math.sqrt(float(out2[0])), float(ref_norm), float(out2[1]), worst))
    assert worst <= 3e-7
    assert int(cnt[0]) == 0   # the reduction left its counter ready for the next launch


def test_fused_adam_checkpoint_interchange_with_torch_adam():
    """optimizer_state_dict (notebook.ipynb:642-646) saved by torch.optim.Adam resumes in FusedAdam (and back) with identical
    moments and bias correction: save after 2 steps, load, one more step, compare with torch's third step."""
    from cilrs_b200.optim import FusedAdam
    O = _O()
    sd = O.synthetic_state_dict(0)
    m = _model(sd, train=True)
    ref_params = [p.detach().clone().requires_grad_(True) for p in m.parameters()]
    ropt = torch.optim.Adam(ref_params, lr=1e-3, weight_decay=1e-4)
    gen = torch.Generator(device="cuda").manual_seed(2)
    grads = [[torch.randn(p.shape, generator=gen, device="cuda") * 1e-2 for p in ref_params] for _ in range(3)]
    for s in range(2):
        for rp, g in zip(ref_params, grads[s]):
            rp.grad = g.clone()
        ropt.step()
    ck = ropt.state_dict()
    with torch.no_grad():
        for p, rp in zip(m.parameters(), ref_params):
            p.copy_(rp)
    opt = FusedAdam(m.parameters(), lr=5e-4, weight_decay=0.0, model=m)
    opt.load_state_dict(ck)
    assert opt._step == 2 and opt.param_groups[0]["lr"] == 1e-3 and opt.param_groups[0]["weight_decay"] == 1e-4
    for rp, g in zip(ref_params, grads[2]):
        rp.grad = g.clone()
    ropt.step()
    for p, g in zip(m.parameters(), grads[2]):
        p.grad = g.clone()
    opt.step()
    torch.cuda.synchronize()
    worst = max(float((p.detach() - rp.detach()).abs().max()) for p, rp in zip(m.parameters(), ref_params))
    print("resume from torch.optim.Adam checkpoint: max |p - torch| after the next step %.3e" % worst)
    assert worst <= 3e-7
    out = opt.state_dict()
    assert float(out["state"][0]["step"]) == 3.0
    ropt2 = torch.optim.Adam(ref_params, lr=1e-3, weight_decay=1e-4)
    ropt2.load_state_dict(out)   # and back into torch
    assert float(ropt2.state_dict()["state"][5]["step"]) == 3.0
    assert _rel(ropt2.state_dict()["state"][5]["exp_avg"], ropt.state_dict()["state"][5]["exp_avg"]) <= 1e-6


def test_torch_optimizer_and_reload_refresh_the_packed_weights():
    """parameters updated THROUGH torch (torch.optim.Adam.step(), a second load_state_dict) must reach the
    packed bf16 operands: an eval forward afterwards equals a fresh module built from m.state_dict()."""
    O = _O()
    sd = O.synthetic_state_dict(0)
    image, speed, command, targets = _inputs(8, 6, "cuda")
    m = _model(sd, train=True)
    optimizer = torch.optim.Adam(m.parameters(), lr=1e-2)
    for _ in range(2):
        c, p = m(image, speed, command)
        loss, _ = O.loss_mse(c, targets, p, speed)
        optimizer.zero_grad()
        loss.backward()
        optimizer.step()
    m.eval()
    with torch.no_grad():
        c1, p1 = m(image, speed, command)
    fresh = _model({k: v.detach().cpu().clone() for k, v in m.state_dict().items()}, train=False)
    with torch.no_grad():
        c2, p2 = fresh(image, speed, command)
    assert torch.equal(c1, c2) and torch.equal(p1, p2)
    m0 = _model(sd, train=False)
    with torch.no_grad():
        c0, _ = m0(image, speed, command)
    assert not torch.equal(c0, c1)            # the conv weights really moved
    m.load_state_dict(sd)                     # reload the initial checkpoint into a module that has already run
    with torch.no_grad():
        c3, p3 = m(image, speed, command)
    assert torch.equal(c3, c0)


# ------------------------------------------------------------------------------------------------------------------
# inference call sites
# ------------------------------------------------------------------------------------------------------------------
def test_predict_controls_and_inference_session_match_oracle():
    """predict_controls (model/autonomous_drive.py:908-920): resize + normalise + eval forward + x90 de-normalisation, speed
    clamp min(kmh/90, 1); InferenceSession (one CUDA graph) on RGB and on CARLA's BGRA frames with channel reversal (:869-873,
    :1551)."""
    from cilrs_b200.preprocess import InferenceSession, predict_controls
    O = _O()
    sd = O.synthetic_state_dict(0)
    frames, _, _, _ = O.synthetic_batch(2, seed=7, smooth=True)
    m = _model(sd, train=False)
    sd64 = {k: (v.double() if v.is_floating_point() else v) for k, v in sd.items()}
    sess = InferenceSession(m, batch=1)
    bgra = np.concatenate([frames[..., ::-1], np.full(frames.shape[:3] + (1,), 255, np.uint8)], axis=3)
    sess_bgra = InferenceSession(m, batch=1, src_c=4, reverse=True)
    for i, (kmh, cmd) in enumerate([(30.0, 1), (120.0, 3)]):     # the second one clamps the speed input to 1.0
        _, img = O.preprocess_c(frames[i:i + 1])
        spd = torch.tensor([O.normalise_speed(kmh)], dtype=torch.float64)
        c64, p64 = O.forward(sd64, torch.from_numpy(img).double(), spd, torch.tensor([cmd]), training=False)
        ref = [float(c64[0, 0]), float(c64[0, 1]), float(c64[0, 2]), float(p64[0]) * 90.0]
        scale = max(abs(v) for v in ref[:3])
        for name, got in (("predict_controls", predict_controls(m, frames[i], kmh, cmd)),
                          ("InferenceSession", sess.predict(frames[i], kmh, cmd)),
                          ("InferenceSession BGRA", sess_bgra.predict(bgra[i], kmh, cmd))):
            err = max(abs(a - b) for a, b in zip(got[:3], ref[:3])) / scale
            errs = abs(got[3] - ref[3]) / abs(ref[3])
            print("%s frame %d: controls rel %.3e, speed rel %.3e" % (name, i, err, errs))
            assert err <= 2e-2 and errs <= 2e-2
    with pytest.raises(IndexError):
        sess.predict(frames[0], 10.0, 7)       # the reference's gather raises on an out-of-range command
    assert len(sess.predict(frames[0], 10.0, 2)) == 4   # and the session keeps working afterwards


def test_sharded_inference_world1_and_batch512():
    """configs[4] per-GPU shard: 512 mixed-command frames through ShardedInference (world size 1 here) equal the module's own
    eval forward on the same preprocessed frames, and match the fp64 oracle on a sample of them."""
    from cilrs_b200 import ops
    from cilrs_b200.preprocess import ShardedInference
    O = _O()
    sd = O.synthetic_state_dict(0)
    m = _model(sd, train=False)
    n = 512
    g = torch.Generator().manual_seed(3)
    base, _, _, _ = O.synthetic_batch(8, seed=11, smooth=True)
    frames = torch.from_numpy(base)[torch.randint(0, 8, (n,), generator=g)].contiguous()
    frames = frames + torch.randint(0, 8, frames.shape, generator=g, dtype=torch.uint8)   # (uint8 wrap-around is fine here)
    kmh = torch.rand(n, generator=g) * 120
    cmd = torch.randint(0, 4, (n,), generator=g)
    sh = ShardedInference(m, n)
    assert sh.shard() == (0, n)
    out = sh.predict_batch(frames, kmh, cmd).clone()
    full = sh.gather(out)
    assert full.shape == (n, 4)
    img = ops.preprocess(frames.cuda(), want_f32=True)["f32"]
    with torch.no_grad():
        c, p = m(img, torch.clamp(kmh / 90, max=1.0).cuda(), cmd.cuda())
    assert _rel(out[:, :3], c) <= 1e-3 and _rel(out[:, 3], p * 90) <= 1e-3
    sd64 = {k: (v.double() if v.is_floating_point() else v) for k, v in sd.items()}
    idx = list(range(0, n, 64))
    c64, p64 = O.forward(sd64, img[idx].double().cpu(), torch.clamp(kmh[idx] / 90, max=1.0).double(), cmd[idx], training=False)
    assert _rel(out[idx, :3], c64) <= 2e-2 and _rel(out[idx, 3], p64 * 90) <= 2e-2


def test_validate_on_device_matches_the_reference_loop():
    """validate() (notebook/notebook.ipynb:563-585): mean of the per-batch loss dicts and per-command steer MAE, restated here
    line by line on the CPU as the checker; ragged last batch; a command that never occurs gives nan."""
    from cilrs_b200.loss import CILRSLoss, validate
    O = _O()
    sd = O.synthetic_state_dict(0)
    m = _model(sd, train=False)
    batches = []
    for i, b in enumerate((16, 16, 5)):
        image, speed, command, targets = _inputs(b, 50 + i)
        command = command.clamp(max=2)       # command 3 never occurs
        batches.append((image, speed, command, targets))
    crit = CILRSLoss()
    losses, cmd_avg = validate(m, batches, crit, "cuda")
    ref = {k: 0.0 for k in ("total", "control", "steer", "throttle", "brake", "speed")}
    errs = {i: [] for i in range(4)}
    with torch.no_grad():
        for image, speed, command, targets in batches:
            c, p = m(image.cuda(), speed.cuda(), command.cuda())
            _, ld = O.loss_l1(c.cpu(), targets, p.cpu(), speed)
            for k in ref:
                ref[k] += float(ld[k])
            serr = (c[:, 0].cpu() - targets[:, 0]).abs()
            for ci in range(4):
                if (command == ci).any():
                    errs[ci].extend(serr[command == ci].numpy().tolist())
    names = {0: "FOLLOW", 1: "LEFT", 2: "RIGHT", 3: "STRAIGHT"}
    for k in ref:
        assert abs(losses[k] - ref[k] / 3) <= 1e-5 * abs(ref[k] / 3), k
    for ci in range(3):
        assert abs(cmd_avg[names[ci]] - float(np.mean(errs[ci]))) <= 1e-5 * float(np.mean(errs[ci]))
    assert math.isnan(cmd_avg["STRAIGHT"])
