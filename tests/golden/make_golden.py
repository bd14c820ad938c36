"""Generates tests/golden/*.npz by running the REFERENCE's own code (AST-extracted from /root/reference, which is only
present in the build container) on seeded synthetic inputs. Run once here; the fixtures travel with the repo.

    python tests/golden/make_golden.py

Reference pieces executed unmodified:
  class CILRS                model/autonomous_drive.py:361-399
  preprocess_image           model/autonomous_drive.py:897-902 (+ constants :481-485, transform :502-504)
  class CILRSLoss            notebook/notebook.ipynb:504-527
  optim.Adam(lr, weight_decay)  notebook/notebook.ipynb:533-534
"""
import ast
import hashlib
import json
import os
import sys
import warnings

import numpy as np
import torch
import torch.nn as nn

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
REF = "/root/reference"

from oracle import cilrs_oracle as O  # noqa: E402


def load_reference():
    import cv2
    import torchvision.models as models
    import torchvision.transforms as transforms
    src = open(os.path.join(REF, "model/autonomous_drive.py")).read()
    tree = ast.parse(src)
    ns = {"torch": torch, "nn": nn, "models": models, "transforms": transforms, "cv2": cv2, "np": np}
    for node in tree.body:
        if isinstance(node, ast.ClassDef) and node.name == "CILRS":
            exec(ast.get_source_segment(src, node), ns)
        if isinstance(node, ast.ClassDef) and node.name == "AutonomousDriver":
            for sub in node.body:
                if isinstance(sub, ast.FunctionDef) and sub.name == "preprocess_image":
                    import textwrap
                    exec(textwrap.dedent(ast.get_source_segment(src, sub)), ns)
    nb = json.load(open(os.path.join(REF, "notebook/notebook.ipynb")))
    cell = "".join(nb["cells"][0]["source"])
    ctree = ast.parse(cell)
    for node in ctree.body:
        if isinstance(node, ast.ClassDef) and node.name == "CILRSLoss":
            exec(ast.get_source_segment(cell, node), ns)

    class Stub:
        IMG_WIDTH, IMG_HEIGHT = 200, 88
        device = torch.device("cpu")
        transform = transforms.Compose([transforms.Normalize(mean=[0.485, 0.456, 0.406], std=[0.229, 0.224, 0.225])])

    return ns["CILRS"], ns["CILRSLoss"], (lambda img: ns["preprocess_image"](Stub, img))


def grads_summary(model):
    norms, heads = [], []
    for _, p in model.named_parameters():
        g = p.grad.detach().double().reshape(-1)
        norms.append(float(g.norm()))
        heads.append(g[:3].numpy() if g.numel() >= 3 else np.pad(g.numpy(), (0, 3 - g.numel())))
    return np.asarray(norms), np.stack(heads)


def main():
    warnings.simplefilter("ignore")
    torch.manual_seed(0)
    torch.set_num_threads(8)
    CILRS, CILRSLoss, preprocess_image = load_reference()

    # ---------------- preprocessing ----------------
    import cv2
    frames_noise, _, _, _ = O.synthetic_batch(2, seed=11, smooth=False)
    frames_smooth, _, _, _ = O.synthetic_batch(2, seed=12, smooth=True)
    frames = np.concatenate([frames_noise[:1], frames_smooth[:1]], axis=0)
    small = np.stack([cv2.resize(f, (200, 88)) for f in frames])
    f32 = torch.cat([preprocess_image(f) for f in frames]).numpy()
    odd = np.random.default_rng(13).integers(0, 256, size=(1, 123, 321, 3), dtype=np.uint8)
    odd_small = np.stack([cv2.resize(f, (200, 88)) for f in odd])
    np.savez_compressed(os.path.join(HERE, "preprocess_ref.npz"), small=small, f32_frame0=f32[0],
                        f32_sha256=np.frombuffer(hashlib.sha256(f32.tobytes()).digest(), dtype=np.uint8), odd_small=odd_small)

    # ---------------- model ----------------
    B = 4
    frames, speed, command, targets = O.synthetic_batch(B, seed=21, smooth=True)
    command[:4] = [0, 1, 2, 3]
    image = torch.cat([preprocess_image(f) for f in frames])
    sd32 = O.synthetic_state_dict(0)
    out = {"speed": speed, "command": command, "targets": targets}
    for tag, dt in (("f64", torch.float64), ("f32", torch.float32)):
        model = CILRS(num_commands=4, dropout=0.0)
        model.load_state_dict(sd32, strict=True)  # also proves the key/shape layout of synthetic_state_dict
        model = model.to(dt)
        img, spd, cmd, tgt = image.to(dt), torch.from_numpy(speed).to(dt), torch.from_numpy(command), torch.from_numpy(targets).to(dt)
        for mode in ("eval", "train"):
            model.load_state_dict({k: (v.to(dt) if v.is_floating_point() else v) for k, v in sd32.items()})
            model.train(mode == "train")
            model.zero_grad()
            ctrl, ps = model(img, spd, cmd)
            crit = CILRSLoss()
            l1_total, l1_dict = crit(ctrl, tgt, ps, spd)
            mse_total = nn.functional.mse_loss(ctrl, tgt) + 0.05 * nn.functional.mse_loss(ps, spd)
            mse_total.backward(retain_graph=True)
            n_mse, h_mse = grads_summary(model)
            model.zero_grad()
            l1_total.backward()
            n_l1, h_l1 = grads_summary(model)
            pre = "%s_%s_" % (tag, mode)
            out[pre + "controls"] = ctrl.detach().double().numpy()
            out[pre + "pred_speed"] = ps.detach().double().numpy()
            out[pre + "loss_l1"] = np.asarray([l1_dict[k] for k in ("total", "control", "steer", "throttle", "brake", "speed")])
            out[pre + "loss_mse"] = np.asarray([float(mse_total)])
            out[pre + "gradnorm_mse"] = n_mse
            out[pre + "gradhead_mse"] = h_mse
            out[pre + "gradnorm_l1"] = n_l1
            out[pre + "gradhead_l1"] = h_l1
            if mode == "train":
                sdn = model.state_dict()
                out[pre + "bn1_running_mean"] = sdn["visual_encoder.1.running_mean"].double().numpy()
                out[pre + "bn1_running_var"] = sdn["visual_encoder.1.running_var"].double().numpy()
                out[pre + "l4_running_var"] = sdn["visual_encoder.7.2.bn2.running_var"].double().numpy()
                out[pre + "nbt"] = np.asarray([int(sdn["visual_encoder.1.num_batches_tracked"])])
        if tag == "f32":
            # one Adam step (BASELINE recipe: lr 2e-4, wd 1e-4) on the train-mode MSE gradients
            model.load_state_dict(sd32)
            model.train()
            opt = torch.optim.Adam(model.parameters(), lr=2e-4, weight_decay=1e-4)
            for _ in range(2):
                opt.zero_grad()
                ctrl, ps = model(img, spd, cmd)
                (nn.functional.mse_loss(ctrl, tgt) + 0.05 * nn.functional.mse_loss(ps, spd)).backward()
                opt.step()
            sdn = model.state_dict()
            out["adam2_stem_w_head"] = sdn["visual_encoder.0.weight"].reshape(-1)[:16].double().numpy()
            out["adam2_br6_b"] = torch.cat([sdn["control_branches.%d.6.bias" % k] for k in range(4)]).double().numpy()
            out["adam2_sp5_w_head"] = sdn["speed_predictor.5.weight"].reshape(-1)[:16].double().numpy()
    out["image_sha256"] = np.frombuffer(hashlib.sha256(image.numpy().tobytes()).digest(), dtype=np.uint8)
    np.savez_compressed(os.path.join(HERE, "cilrs_ref_b4.npz"), **out)
    print("wrote", os.listdir(HERE))


if __name__ == "__main__":
    main()
