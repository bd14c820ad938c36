"""CPU tests: the oracle against the reference-generated golden vectors (and against the reference itself / OpenCV when
they are available in this container), plus the host-side logic that needs no GPU."""
import hashlib
import os

import numpy as np
import pytest
import torch

from oracle import cilrs_oracle as O

GOLD = os.path.join(os.path.dirname(__file__), "golden")
REF = "/root/reference"


def _frames():
    noise, _, _, _ = O.synthetic_batch(2, seed=11, smooth=False)
    smooth, _, _, _ = O.synthetic_batch(2, seed=12, smooth=True)
    return np.concatenate([noise[:1], smooth[:1]], axis=0)


def test_preprocess_oracles_match_reference_golden():
    g = np.load(os.path.join(GOLD, "preprocess_ref.npz"))
    frames = _frames()
    for fn in (O.preprocess_np, O.preprocess_c):
        small, f32 = fn(frames)
        assert np.array_equal(small, g["small"])
        assert np.array_equal(f32[0], g["f32_frame0"])
        assert np.array_equal(np.frombuffer(hashlib.sha256(f32.tobytes()).digest(), dtype=np.uint8), g["f32_sha256"])
    odd = np.random.default_rng(13).integers(0, 256, size=(1, 123, 321, 3), dtype=np.uint8)
    assert np.array_equal(O.resize_u8_np(odd), g["odd_small"])
    assert np.array_equal(O.preprocess_c(odd)[0], g["odd_small"])


def test_preprocess_oracle_matches_opencv_when_present():
    cv2 = pytest.importorskip("cv2")
    rng = np.random.default_rng(3)
    for shape in ((600, 800), (480, 640), (88, 200), (97, 203), (1080, 1920)):
        img = rng.integers(0, 256, size=(shape[0], shape[1], 3), dtype=np.uint8)
        assert np.array_equal(O.resize_u8_np(img[None])[0], cv2.resize(img, (200, 88)))
    grad = np.tile(np.arange(800, dtype=np.uint8)[None, :, None], (600, 1, 3))
    assert np.array_equal(O.preprocess_c(grad[None])[0][0], cv2.resize(grad, (200, 88)))
    # BGRA + channel reversal (CARLA camera layout, model/autonomous_drive.py:869-873,1551)
    bgra = rng.integers(0, 256, size=(1, 600, 800, 4), dtype=np.uint8)
    rgb = np.ascontiguousarray(bgra[0, :, :, :3][:, :, ::-1])
    assert np.array_equal(O.preprocess_c(bgra, reverse=True)[0][0], cv2.resize(rgb, (200, 88)))


def test_state_dict_spec_is_the_reference_layout():
    spec = O.state_dict_spec()
    assert len(spec) == 250
    n_param = sum(int(np.prod(s)) for _, s, k in spec if k not in ("rm", "rv", "nbt"))
    assert n_param == 22421453                       # notebook/notebook.ipynb:52
    n_buf = sum(int(np.prod(s)) if s else 1 for _, s, k in spec if k in ("rm", "rv", "nbt"))
    assert n_buf == 17060
    assert len([1 for _, _, k in spec if k not in ("rm", "rv", "nbt")]) == 142


def _inputs():
    frames, speed, command, targets = O.synthetic_batch(4, seed=21, smooth=True)
    command[:4] = [0, 1, 2, 3]
    _, image = O.preprocess_c(frames)
    return torch.from_numpy(image), torch.from_numpy(speed), torch.from_numpy(command), torch.from_numpy(targets)


@pytest.mark.parametrize("mode", ["eval", "train"])
def test_model_oracle_matches_reference_golden_fp64(mode):
    """fp64 restatement vs the reference class run in fp64: agreement to round-off"""
    g = np.load(os.path.join(GOLD, "cilrs_ref_b4.npz"))
    image, speed, command, targets = _inputs()
    assert np.array_equal(np.frombuffer(hashlib.sha256(image.numpy().tobytes()).digest(), dtype=np.uint8), g["image_sha256"])
    sd = O.synthetic_state_dict(0)
    sd64 = {}
    for k, v in sd.items():
        if not v.is_floating_point() or k.endswith(("running_mean", "running_var")):
            sd64[k] = v.double() if v.is_floating_point() else v
        else:
            sd64[k] = v.double().requires_grad_(True)
    upd = {}
    c, p = O.forward(sd64, image.double(), speed.double(), command, training=(mode == "train"), update=upd)
    pre = "f64_%s_" % mode
    assert np.allclose(c.detach().numpy(), g[pre + "controls"], rtol=1e-9, atol=1e-11)
    assert np.allclose(p.detach().numpy(), g[pre + "pred_speed"], rtol=1e-9, atol=1e-11)
    tot_l1, d = O.loss_l1(c, targets.double(), p, speed.double())
    got = np.asarray([float(d[k]) for k in ("total", "control", "steer", "throttle", "brake", "speed")])
    assert np.allclose(got, g[pre + "loss_l1"], rtol=1e-9)
    tot_mse, _ = O.loss_mse(c, targets.double(), p, speed.double())
    assert np.allclose(float(tot_mse), g[pre + "loss_mse"][0], rtol=1e-9)
    tot_mse.backward()
    names = [k for k, _ in O.params_in_order(sd64)]
    norms = np.asarray([float(sd64[k].grad.norm()) for k in names])
    assert np.allclose(norms, g[pre + "gradnorm_mse"], rtol=1e-6, atol=1e-12)
    heads = np.stack([np.pad(sd64[k].grad.reshape(-1)[:3].numpy(), (0, max(0, 3 - sd64[k].grad.numel()))) for k in names])
    assert np.allclose(heads, g[pre + "gradhead_mse"], rtol=1e-6, atol=1e-12)
    if mode == "train":
        assert np.allclose(upd["visual_encoder.1.running_mean"].numpy(), g[pre + "bn1_running_mean"], rtol=1e-9)
        assert np.allclose(upd["visual_encoder.1.running_var"].numpy(), g[pre + "bn1_running_var"], rtol=1e-9)
        assert np.allclose(upd["visual_encoder.7.2.bn2.running_var"].numpy(), g[pre + "l4_running_var"], rtol=1e-9)
        assert int(upd["visual_encoder.1.num_batches_tracked"]) == int(g[pre + "nbt"][0])


def test_adam_oracle_matches_reference_golden():
    """two torch.optim.Adam steps of the reference (fp32) vs the restated update rule driven by oracle gradients"""
    g = np.load(os.path.join(GOLD, "cilrs_ref_b4.npz"))
    image, speed, command, targets = _inputs()
    sd = O.synthetic_state_dict(0)
    names = [k for k, _ in O.params_in_order(sd)]
    m = {k: torch.zeros_like(sd[k]) for k in names}
    v = {k: torch.zeros_like(sd[k]) for k in names}
    state = dict(sd)
    for step in (1, 2):
        leaf = {k: state[k].clone().requires_grad_(True) for k in names}
        cur = dict(state)
        cur.update(leaf)
        upd = {}
        c, p = O.forward(cur, image, speed, command, training=True, update=upd)
        O.loss_mse(c, targets, p, speed)[0].backward()
        for k in names:
            pn, mn, vn = O.adam_step(state[k], leaf[k].grad, m[k], v[k], step)
            state[k], m[k], v[k] = pn.detach(), mn, vn
        state.update(upd)
    assert np.allclose(state["visual_encoder.0.weight"].reshape(-1)[:16].numpy(), g["adam2_stem_w_head"], rtol=2e-4, atol=1.5e-6)
    got = torch.cat([state["control_branches.%d.6.bias" % k] for k in range(4)]).numpy()
    assert np.allclose(got, g["adam2_br6_b"], rtol=2e-4, atol=1.5e-6)
    assert np.allclose(state["speed_predictor.5.weight"].reshape(-1)[:16].numpy(), g["adam2_sp5_w_head"], rtol=2e-4, atol=1.5e-6)


@pytest.mark.skipif(not os.path.exists(REF), reason="reference tree only exists in the build container")
def test_oracle_against_the_reference_class_itself():
    import ast
    import warnings
    import torchvision.models as models
    import torch.nn as nn
    warnings.simplefilter("ignore")
    src = open(os.path.join(REF, "model/autonomous_drive.py")).read()
    ns = {"torch": torch, "nn": nn, "models": models}
    for node in ast.parse(src).body:
        if isinstance(node, ast.ClassDef) and node.name == "CILRS":
            exec(ast.get_source_segment(src, node), ns)
    ref = ns["CILRS"](num_commands=4, dropout=0.0)
    sd = O.synthetic_state_dict(5)
    ref.load_state_dict(sd, strict=True)
    image, speed, command, targets = _inputs()
    for training in (False, True):
        ref.load_state_dict(sd)
        ref.train(training)
        c_ref, p_ref = ref(image, speed, command)
        c, p = O.forward(sd, image, speed, command, training=training)
        assert torch.allclose(c, c_ref, rtol=1e-4, atol=1e-5) and torch.allclose(p, p_ref, rtol=1e-4, atol=1e-5)


def test_speed_normalisation():
    assert O.normalise_speed(45.0) == 0.5 and O.normalise_speed(200.0) == 1.0
