"""fp32 compute mode (CILRS(compute_dtype="fp32"), csrc/fp32_path.cu) against the fp64 oracle: BASELINE.json north_star
"fp32 mode: controls, speed prediction, losses and gradients match within 1e-4 relative" (the reference itself runs fp32:
configs/train_config.json:54, model/autonomous_drive.py:495).

  * forward (eval and train-mode BN), running statistics, losses: 1e-4 (measured ~1e-6)
  * gradients with frozen BatchNorm (eval-mode autograd), every tensor: 1e-4 of the tensor's max
  * gradients with train-mode BatchNorm: heads 1e-4; trunk - where the reference's OWN fp32 (torch on the CPU, same inputs) is
    2e-3..8e-3 from fp64 because 36 chained BatchNorm backwards amplify rounding - at most 2x the reference's
    fp32 error, both numbers printed
  * one optimizer step (FusedAdam on the fp32-mode module) vs the oracle's Adam
"""
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
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return float((a - b).abs().max() / (b.abs().max() + 1e-30))


def _setup(B=4, seed=21):
    O = _O()
    frames, speed, command, targets = O.synthetic_batch(B, seed=seed, smooth=True)
    command[:4] = [0, 1, 2, 3]
    _, image = O.preprocess_c(frames)
    sd = O.synthetic_state_dict(0)
    return O, sd, torch.from_numpy(image), torch.from_numpy(speed), torch.from_numpy(command), torch.from_numpy(targets)


def _model(sd, train=False):
    from cilrs_b200.model import CILRS
    m = CILRS(num_commands=4, dropout=0.0, compute_dtype="fp32")
    m.load_state_dict(sd, strict=True)
    m = m.to("cuda")
    m.train(train)
    return m


def _leaf_sd(sd, dtype, device="cpu"):
    out = {}
    for k, v in sd.items():
        if not v.is_floating_point():
            out[k] = v.to(device)
        elif k.endswith("running_mean") or k.endswith("running_var"):
            out[k] = v.to(device=device, dtype=dtype)
        else:
            out[k] = v.to(device=device, dtype=dtype).clone().requires_grad_(True)
    return out


def test_fp32_eval_and_train_forward_match_reference_golden():
    g = np.load(os.path.join(GOLD, "cilrs_ref_b4.npz"))
    O, sd, image, speed, command, targets = _setup()
    m = _model(sd)
    with torch.no_grad():
        c, p = m(image.cuda(), speed.cuda(), command.cuda())
    ec, ep = _rel(c, torch.from_numpy(g["f64_eval_controls"])), _rel(p, torch.from_numpy(g["f64_eval_pred_speed"]))
    print("fp32 mode eval forward vs reference fp64 golden: controls %.3e, speed %.3e" % (ec, ep))
    assert ec <= 1e-4 and ep <= 1e-4
    m.train()
    with torch.no_grad():
        c, p = m(image.cuda(), speed.cuda(), command.cuda())
    ec, ep = _rel(c, torch.from_numpy(g["f64_train_controls"])), _rel(p, torch.from_numpy(g["f64_train_pred_speed"]))
    print("fp32 mode train forward vs reference fp64 golden: controls %.3e, speed %.3e" % (ec, ep))
    assert ec <= 1e-4 and ep <= 1e-4
    nsd = m.state_dict()
    assert int(nsd["visual_encoder.1.num_batches_tracked"]) == 1
    assert _rel(nsd["visual_encoder.1.running_mean"], torch.from_numpy(g["f64_train_bn1_running_mean"])) <= 1e-4
    assert _rel(nsd["visual_encoder.1.running_var"], torch.from_numpy(g["f64_train_bn1_running_var"])) <= 1e-4
    assert _rel(nsd["visual_encoder.7.2.bn2.running_var"], torch.from_numpy(g["f64_train_l4_running_var"])) <= 1e-4


def _grads(m, O, sd, image, speed, command, targets, training, loss):
    lossfn = O.loss_mse if loss == "mse" else O.loss_l1
    sd64 = _leaf_sd(sd, torch.float64, "cuda")
    c64, p64 = O.forward(sd64, image.double().cuda(), speed.double().cuda(), command.cuda(), training=training)
    tot64, _ = lossfn(c64, targets.double().cuda(), p64, speed.double().cuda())
    tot64.backward()
    c, p = m(image.cuda(), speed.cuda(), command.cuda())
    tot, _ = lossfn(c, targets.cuda(), p, speed.cuda())
    m.zero_grad()
    tot.backward()
    torch.cuda.synchronize()
    return float(tot), float(tot64), {n: q.grad.double() for n, q in m.named_parameters()}, {n: sd64[n].grad for n, _ in m.named_parameters()}


@pytest.mark.parametrize("loss", ["mse", "l1"])
def test_fp32_frozen_bn_gradients_every_tensor(loss):
    O, sd, image, speed, command, targets = _setup(B=8, seed=41)
    m = _model(sd, train=False)
    tot, tot64, got, ref = _grads(m, O, sd, image, speed, command, targets, False, loss)
    errs = sorted(((_rel(got[n], ref[n]), n) for n in got if float(ref[n].abs().max()) > 0), reverse=True)
    flat_g = torch.cat([got[n].reshape(-1) for n in got]); flat_r = torch.cat([ref[n].reshape(-1) for n in got])
    glob = float((flat_g - flat_r).norm() / flat_r.norm())
    print("fp32 mode frozen-BN %s: loss %.8f vs fp64 %.8f, global grad err %.3e, worst tensors %s" % (loss, tot, tot64, glob, errs[:3]))
    assert abs(tot - tot64) <= 1e-5 * abs(tot64)
    assert glob <= 1e-4 and errs[0][0] <= 1e-4
    for n in got:   # tensors whose reference gradient is exactly zero (non-selected branches never occur here: all 4 commands present)
        if float(ref[n].abs().max()) == 0:
            assert float(got[n].abs().max()) == 0, n


def test_fp32_train_mode_gradients():
    O, sd, image, speed, command, targets = _setup(B=16, seed=31)
    m = _model(sd, train=True)
    tot, tot64, got, ref = _grads(m, O, sd, image, speed, command, targets, True, "mse")
    # the reference's own fp32 (torch on the CPU: no TF32 involved) on the same inputs
    sdr = _leaf_sd(sd, torch.float32)
    cr, pr = O.forward(sdr, image, speed, command, training=True)
    O.loss_mse(cr, targets, pr, speed)[0].backward()

    def glob(a, sel):
        x = torch.cat([a[n].double().cpu().reshape(-1) for n in got if sel(n)])
        y = torch.cat([ref[n].cpu().reshape(-1) for n in got if sel(n)])
        return float((x - y).norm() / y.norm())

    refg = {n: sdr[n].grad for n in got}
    trunk = lambda n: n.startswith("visual_encoder")
    heads = lambda n: not n.startswith("visual_encoder")
    e_t, e_h, r_t, r_h = glob(got, trunk), glob(got, heads), glob(refg, trunk), glob(refg, heads)
    print("fp32 mode train-mode: loss %.8f vs fp64 %.8f | trunk grad err ours %.3e, reference fp32 %.3e | heads ours %.3e, reference fp32 %.3e"
          % (tot, tot64, e_t, r_t, e_h, r_h))
    assert abs(tot - tot64) <= 1e-5 * abs(tot64)
    assert e_h <= 1e-4
    assert e_t <= max(1e-4, 2.0 * r_t)


def test_fp32_mode_adam_step_and_state_dict():
    from cilrs_b200.optim import FusedAdam
    O, sd, image, speed, command, targets = _setup(B=4)
    m = _model(sd, train=True)
    opt = FusedAdam(m.parameters(), lr=2e-4, weight_decay=1e-4, model=m)
    p0 = m.flat_parameters().clone()
    c, p = m(image.cuda(), speed.cuda(), command.cuda())
    tot, _ = O.loss_mse(c, targets.cuda(), p, speed.cuda())
    opt.zero_grad()
    tot.backward()
    g = m.flat_gradients().clone()
    opt.step()
    pp, _, _ = O.adam_step(p0.double(), g.double(), torch.zeros_like(p0).double(), torch.zeros_like(p0).double(), 1)
    torch.cuda.synchronize()
    assert float((m.flat_parameters().double() - pp).abs().max()) <= 3e-7
    out = m.state_dict()
    assert list(out.keys()) == [k for k, _, _ in O.state_dict_spec()]
    m.eval()
    with torch.no_grad():
        c1, _ = m(image.cuda(), speed.cuda(), command.cuda())
    fresh = _model({k: v.detach().cpu() for k, v in out.items()})
    with torch.no_grad():
        c2, _ = fresh(image.cuda(), speed.cuda(), command.cuda())
    assert torch.equal(c1, c2)


def test_fp32_mode_reference_training_loop_matches_oracle_trajectory():
    """train_one_epoch's body (notebook/notebook.ipynb:545-555) with the drop-in module in fp32 mode, CILRSLoss, clip_grad_norm_
    and torch.optim.Adam, against the oracle's fp32 run: the trajectories agree to 1e-3 over 4 steps (both are fp32)."""
    from cilrs_b200.loss import CILRSLoss
    O, sd, image, speed, command, targets = _setup(B=8, seed=5)
    m = _model(sd, train=True)
    optimizer = torch.optim.Adam(m.parameters(), lr=1e-4, weight_decay=1e-4)
    criterion = CILRSLoss()
    ours = []
    for _ in range(4):
        pred_ctrl, pred_spd = m(image.cuda(), speed.cuda(), command.cuda())
        loss, ld = criterion(pred_ctrl, targets.cuda(), pred_spd, speed.cuda())
        optimizer.zero_grad()
        loss.backward()
        torch.nn.utils.clip_grad_norm_(m.parameters(), 1.0)
        optimizer.step()
        ours.append(ld["total"])
    state = {k: v.clone() for k, v in sd.items()}
    params = [v.requires_grad_(True) for k, v in state.items() if v.is_floating_point() and "running" not in k]
    ropt = torch.optim.Adam(params, lr=1e-4, weight_decay=1e-4)
    ref = []
    for _ in range(4):
        upd = {}
        c, p = O.forward(state, image, speed, command, training=True, update=upd)
        tot, _ = O.loss_l1(c, targets, p, speed)
        ropt.zero_grad()
        tot.backward()
        torch.nn.utils.clip_grad_norm_(params, 1.0)
        ropt.step()
        state.update(upd)
        ref.append(float(tot.detach()))
    print("fp32 mode reference loop: ours %s | oracle fp32 %s" % (["%.6f" % v for v in ours], ["%.6f" % v for v in ref]))
    for a, b in zip(ours, ref):
        assert abs(a - b) <= 1e-3 * abs(b)
