"""K0 (resize + normalise) on the GPU against the oracle and the reference-generated golden frames: bit-exact."""
import hashlib
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden")


def _o():
    from oracle import cilrs_oracle as O
    return O


def _run(frames_np, **kw):
    from cilrs_b200 import ops
    out = ops.preprocess(torch.from_numpy(frames_np).cuda(), want_u8=True, want_f32=True, want_s2d=kw.pop("s2d", False), **kw)
    torch.cuda.synchronize()
    return out


def test_matches_reference_golden():
    O = _o()
    g = np.load(os.path.join(GOLD, "preprocess_ref.npz"))
    noise, _, _, _ = O.synthetic_batch(2, seed=11, smooth=False)
    smooth, _, _, _ = O.synthetic_batch(2, seed=12, smooth=True)
    frames = np.concatenate([noise[:1], smooth[:1]], axis=0)
    out = _run(frames)
    assert np.array_equal(out["u8"].cpu().numpy(), g["small"])          # cv2.resize of the reference, bit for bit
    f32 = out["f32"].cpu().numpy()
    assert np.array_equal(f32[0], g["f32_frame0"])                      # preprocess_image of the reference, bit for bit
    assert np.array_equal(np.frombuffer(hashlib.sha256(f32.tobytes()).digest(), dtype=np.uint8), g["f32_sha256"])


@pytest.mark.parametrize("shape,c,reverse", [((600, 800), 3, False), ((600, 800), 4, True), ((123, 321), 3, False),
                                              ((88, 200), 3, False), ((720, 1280), 3, True), ((97, 203), 4, False)])
def test_matches_oracle_bit_exact(shape, c, reverse):
    O = _o()
    rng = np.random.default_rng(5)
    frames = rng.integers(0, 256, size=(3, shape[0], shape[1], c), dtype=np.uint8)
    out = _run(frames, reverse=reverse)
    u8, f32 = O.preprocess_c(frames, reverse=reverse)
    assert np.array_equal(out["u8"].cpu().numpy(), u8)
    assert np.array_equal(out["f32"].cpu().numpy(), f32)
    src = frames[..., :3][..., ::-1] if reverse else frames[..., :3]
    assert np.array_equal(O.resize_u8_np(np.ascontiguousarray(src)), u8)  # numpy and C restatements agree


@pytest.mark.parametrize("c,reverse", [(3, False), (4, True)])
def test_already_resized_frames_fast_path_is_bit_identical(c, reverse):
    """88x200 frames with only the conv1-ready output requested take the no-resize kernel: same bits as the general one"""
    from cilrs_b200 import _lib
    frames = torch.from_numpy(np.random.default_rng(8).integers(0, 256, size=(6, 88, 200, c), dtype=np.uint8)).cuda()
    fast = torch.full((6, 47, 103, 16), 5.0, dtype=torch.bfloat16, device="cuda")
    gen = torch.full((6, 47, 103, 16), 9.0, dtype=torch.bfloat16, device="cuda")
    f32 = torch.empty(6, 3, 88, 200, device="cuda")
    _lib.call("cilrs_preprocess_u8", frames, 6, 88, 200, c, int(reverse), 88, 200, None, None, fast, _lib.stream_ptr())
    _lib.call("cilrs_preprocess_u8", frames, 6, 88, 200, c, int(reverse), 88, 200, None, f32, gen, _lib.stream_ptr())
    torch.cuda.synchronize()
    assert torch.equal(fast.view(torch.int16), gen.view(torch.int16))


def test_golden_odd_size():
    g = np.load(os.path.join(GOLD, "preprocess_ref.npz"))
    odd = np.random.default_rng(13).integers(0, 256, size=(1, 123, 321, 3), dtype=np.uint8)
    assert np.array_equal(_run(odd)["u8"].cpu().numpy(), g["odd_small"])


def test_s2d_layout_matches_image_to_s2d():
    from cilrs_b200 import ops
    frames = np.random.default_rng(6).integers(0, 256, size=(5, 600, 800, 3), dtype=np.uint8)
    out = _run(frames, s2d=True)
    ref = ops.image_to_s2d(out["f32"])
    torch.cuda.synchronize()
    assert torch.equal(out["s2d"].view(torch.int16), ref.view(torch.int16))
    # padding is written as zeros even into a dirty buffer
    from cilrs_b200 import _lib
    dirty = torch.full((5, 47, 103, 16), 7.0, dtype=torch.bfloat16, device="cuda")
    _lib.call("cilrs_preprocess_u8", torch.from_numpy(frames).cuda(), 5, 600, 800, 3, 0, 88, 200, None, None, dirty, _lib.stream_ptr())
    torch.cuda.synchronize()
    assert torch.equal(dirty.view(torch.int16), ref.view(torch.int16))


def test_full_size_properties():
    """BASELINE config 3 size (1024 frames): constant frames map to constants, and the op commutes with batching."""
    b = 1024
    frames = torch.empty(b, 600, 800, 3, dtype=torch.uint8, device="cuda")
    vals = torch.arange(b, device="cuda") % 256
    frames[:] = vals.view(b, 1, 1, 1).to(torch.uint8)
    from cilrs_b200 import ops
    out = ops.preprocess(frames, want_u8=True, want_f32=False)
    torch.cuda.synchronize()
    assert torch.equal(out["u8"], vals.view(b, 1, 1, 1).to(torch.uint8).expand(b, 88, 200, 3))
    g = torch.Generator(device="cuda").manual_seed(3)
    frames = torch.randint(0, 256, (b, 600, 800, 3), generator=g, device="cuda", dtype=torch.uint8)
    whole = ops.preprocess(frames, want_u8=True, want_f32=False)["u8"]
    part = ops.preprocess(frames[517:519].clone(), want_u8=True, want_f32=False)["u8"]
    torch.cuda.synchronize()
    assert torch.equal(whole[517:519], part)
    O = _o()
    assert np.array_equal(part.cpu().numpy(), O.preprocess_c(frames[517:519].cpu().numpy())[0])


def test_rejects_cpu_and_bad_args():
    from cilrs_b200 import ops
    with pytest.raises(RuntimeError):
        ops.preprocess(torch.zeros(1, 600, 800, 3, dtype=torch.uint8))
    with pytest.raises(RuntimeError):
        ops.preprocess(torch.zeros(1, 600, 800, 2, dtype=torch.uint8, device="cuda"))
    empty = ops.preprocess(torch.zeros(0, 600, 800, 3, dtype=torch.uint8, device="cuda"))
    assert empty["f32"].shape == (0, 3, 88, 200)
