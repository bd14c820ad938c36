"""CPU tests of the host side: the C-ABI library loads and exports every declared symbol, the drop-in module mirrors the
reference's state_dict layout, the product path fails loudly without CUDA, and the data-parallel range logic (gloo, world 2)."""
import ctypes
import os
import re
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    from cilrs_b200 import _lib
    lib = _lib.lib()
    hdr = open(os.path.join(ROOT, "include", "cilrs_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    names = sorted(set(re.findall(r"\b(cilrs_[a-z0-9_]+)\s*\(", hdr)))
    assert len(names) >= 35
    for n in names:
        assert hasattr(lib, n), "missing export: " + n
    assert lib.cilrs_abi_version() == 4
    assert lib.cilrs_status_string(2) == b"unsupported shape or configuration"


def test_argument_validation_without_gpu():
    """entry points validate before touching the device: safe to call on the CPU box"""
    from cilrs_b200 import _lib
    from cilrs_b200._lib import ConvDesc
    lib = _lib.lib()
    lib.cilrs_conv_packed_weight_bytes.restype = ctypes.c_size_t
    d = ConvDesc(4, 22, 50, 64, 128, 3, 3, 2, 1)
    assert lib.cilrs_conv_packed_weight_bytes(ctypes.byref(d)) == 9 * 64 * 128 * 2
    bad = ConvDesc(4, 22, 50, 60, 128, 3, 3, 2, 1)
    assert lib.cilrs_conv_packed_weight_bytes(ctypes.byref(bad)) == 0
    assert lib.cilrs_conv_fprop(ctypes.byref(d), None, None, None, None, None, None, None, 0, None) == 1
    assert lib.cilrs_preprocess_u8(None, 1, 600, 800, 5, 0, 88, 200, None, None, None, None) == 1
    assert lib.cilrs_preprocess_u8(None, 0, 600, 800, 3, 0, 88, 200, None, None, None, None) == 0   # empty batch
    assert lib.cilrs_adam_step(None, None, None, None, ctypes.c_longlong(8), ctypes.c_float(1e-3), ctypes.c_float(0.9),
                               ctypes.c_float(0.999), ctypes.c_float(1e-8), ctypes.c_float(0.0), ctypes.c_longlong(1), None,
                               ctypes.c_float(1.0), None, None) == 1


def test_param_layout_matches_reference_order():
    from cilrs_b200 import _lib
    from oracle import cilrs_oracle as O
    lib = _lib.lib()
    n = 256
    off = (ctypes.c_longlong * n)()
    siz = (ctypes.c_longlong * n)()
    tot, buf, nbn = ctypes.c_longlong(), ctypes.c_longlong(), ctypes.c_int()
    cnt = lib.cilrs_model_param_layout(off, siz, n, ctypes.byref(tot), ctypes.byref(buf), ctypes.byref(nbn))
    spec = [(k, s) for k, s, kind in O.state_dict_spec() if kind not in ("rm", "rv", "nbt")]
    assert cnt == len(spec) == 142 and nbn.value == 36 and buf.value == 17024
    assert [siz[i] for i in range(cnt)] == [int(np.prod(s)) for _, s in spec]
    assert all(off[i] % 16 == 0 for i in range(cnt)) and all(off[i + 1] >= off[i] + siz[i] for i in range(cnt - 1))
    assert sum(siz[i] for i in range(cnt)) == 22421453
    firsts = [lib.cilrs_model_backward_part_first_tensor(p) for p in range(5)]
    names = [k for k, _ in spec]
    assert [names[f] for f in firsts] == ["visual_encoder.7.0.conv1.weight", "visual_encoder.6.0.conv1.weight",
                                           "visual_encoder.5.0.conv1.weight", "visual_encoder.4.0.conv1.weight", "visual_encoder.0.weight"]


def test_backward_part_ranges_tile_the_gradient_arena():
    """FusedTrainer's default data-parallel schedule allreduces part 0's range early and [0, lo(part 0)) in one collective at
    the end: the five part ranges must be contiguous, back to front, and cover the arena exactly once."""
    from cilrs_b200.model import CILRS
    from cilrs_b200.ddp import backward_part_ranges
    m = CILRS(num_commands=4, dropout=0.0)
    r = backward_part_ranges(m)
    assert len(r) == 5 and r[0][1] == m._total and r[4][0] == 0
    for p in range(4):
        assert r[p][0] == r[p + 1][1] and r[p][0] < r[p][1]
    assert r[0][1] - r[0][0] > 0.6 * m._total   # heads + layer4: most of the bytes go out first


def test_dropin_module_state_dict_and_cpu_refusal():
    from cilrs_b200.model import CILRS
    from oracle import cilrs_oracle as O
    m = CILRS(num_commands=4, dropout=0.0)
    spec = O.state_dict_spec()
    sd = m.state_dict()
    assert list(sd.keys()) == [k for k, _, _ in spec]
    assert all(tuple(sd[k].shape) == tuple(s) for k, s, _ in spec)
    assert sd["visual_encoder.1.num_batches_tracked"].dtype == torch.int64
    assert sum(p.numel() for p in m.parameters()) == 22421453
    syn = O.synthetic_state_dict(2)
    m.load_state_dict(syn, strict=True)
    for k, v in m.state_dict().items():
        assert torch.equal(v, syn[k]), k
    # parameters are views of one arena, in order, and stay leaves
    flat = m.flat_parameters()
    for p, o in zip(m.parameters(), m._offsets):
        assert p.is_leaf and p.requires_grad and p.data_ptr() == flat.data_ptr() + 4 * o
    # strict loading (model/autonomous_drive.py:497 uses the default strict=True): a missing key, an unexpected key and a wrong
    # shape must all raise, as they do for the reference module
    with pytest.raises(RuntimeError, match="Missing key"):
        m.load_state_dict({k: v for k, v in syn.items() if k != "speed_predictor.5.bias"}, strict=True)
    with pytest.raises(RuntimeError, match="Unexpected key"):
        m.load_state_dict(dict(syn, **{"visual_encoder.8.weight": torch.zeros(1)}), strict=True)
    with pytest.raises(RuntimeError, match="size mismatch"):
        m.load_state_dict(dict(syn, **{"speed_encoder.0.weight": torch.zeros(128, 2)}), strict=True)
    with pytest.raises(RuntimeError, match="CUDA"):
        m(torch.zeros(1, 3, 88, 200), torch.zeros(1), torch.zeros(1, dtype=torch.long))
    with pytest.raises(ValueError):
        CILRS(num_commands=3)
    # kaiming fan_out init of the convs as in torchvision's resnet (std = sqrt(2 / (Cout*k*k)))
    m2 = CILRS()
    w = m2.state_dict()["visual_encoder.6.1.conv1.weight"]
    assert abs(float(w.std()) - (2.0 / (256 * 9)) ** 0.5) < 2e-3


def test_missing_library_fails_loudly(monkeypatch):
    from cilrs_b200 import _lib
    monkeypatch.setattr(_lib, "_lib", None)
    monkeypatch.setattr(_lib, "LIB_PATH", "/nonexistent/libcilrs_b200.so")
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        _lib.lib()


def _ddp_worker(rank, world, port, q):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    sys.path.insert(0, ROOT)
    from cilrs_b200.ddp import SCHEDULES, allreduce_ranges, backward_part_ranges, schedule_ranges
    from cilrs_b200.model import CILRS
    m = CILRS()
    ranges = backward_part_ranges(m)
    total = m._total
    assert ranges[0][1] == total and ranges[-1][0] == 0
    assert all(ranges[i][0] == ranges[i + 1][1] for i in range(4))      # the five ranges tile the arena back to front
    g = torch.full((total,), float(rank + 1))
    works = [allreduce_ranges(g, [r], None) for r in ranges]
    for w in works:
        for x in w:
            x.wait()
    g.mul_(1.0 / world)
    # FusedTrainer's default schedule: part 0's range early, everything below it in one collective at the end
    h = torch.full((total,), float(rank + 1))
    for x in allreduce_ranges(h, [ranges[0]], None) + allreduce_ranges(h, [(0, ranges[0][0])], None):
        x.wait()
    h.mul_(1.0 / world)
    assert torch.equal(g, h)
    # every schedule FusedTrainer offers exchanges each gradient exactly once (fp32 and the bf16 exchange buffer alike)
    for name, sched in SCHEDULES.items():
        rs = schedule_ranges(ranges, sched)
        assert rs[0][1] == total and rs[-1][0] == 0 and all(rs[i][0] == rs[i + 1][1] for i in range(len(rs) - 1)), name
        for dtype in (torch.float32, torch.bfloat16):
            t = torch.full((total,), float(rank + 1), dtype=dtype)
            for x in allreduce_ranges(t, rs, None):
                x.wait()
            assert float(t.float().min()) == float(t.float().max()) == 3.0, (name, dtype)
    q.put((rank, float(g.min()), float(g.max())))
    dist.destroy_process_group()


def test_gradient_allreduce_ranges_gloo_world2():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_ddp_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(2)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for _, lo, hi in res:
        assert lo == hi == 1.5      # mean of ranks' gradients (1 and 2)
