"""N3 on the CPU: the JPEG oracle against the reference loader's own decode (cv2.imread, notebook/notebook.ipynb:404-405) - the
committed golden arrays and OpenCV run live - and the host half of the C-ABI (marker parsing, derived Huffman tables, file reads)."""
import ctypes
import glob
import os

import numpy as np
import pytest

from oracle import jpeg_oracle as J

HERE = os.path.dirname(os.path.abspath(__file__))
JPEGS = sorted(glob.glob(os.path.join(HERE, "golden", "jpeg", "*.jpg")))
GOLDEN = np.load(os.path.join(HERE, "golden", "jpeg_golden.npz"))


def _name(p):
    return os.path.splitext(os.path.basename(p))[0]


@pytest.mark.parametrize("path", JPEGS, ids=_name)
def test_oracle_equals_the_reference_loader_golden(path):
    got = J.decode_rgb(open(path, "rb").read())
    assert got.dtype == np.uint8 and np.array_equal(got, GOLDEN[_name(path)])   # bit for bit


def test_fixture_coverage():
    modes = set()
    for p in JPEGS:
        h = J.parse(open(p, "rb").read())
        modes.add((len(h.comps), h.comps[0][1], h.comps[0][2]))
    assert {(3, 2, 2), (3, 1, 1), (1, 1, 1)} <= modes   # 4:2:0 (the collector's), 4:4:4, grey


def test_oracle_equals_opencv_live_on_fresh_encodes():
    cv2 = pytest.importorskip("cv2")
    rng = np.random.default_rng(5)
    for k, (h, w, q) in enumerate([(88, 200, 95), (88, 200, 95), (88, 200, 30), (16, 16, 95), (9, 250, 95), (123, 57, 80)]):
        img = rng.integers(0, 256, (h, w, 3), dtype=np.uint8)
        if k % 2 == 0:   # smooth content exercises long zero runs, noise the long codes
            img = cv2.GaussianBlur(img, (9, 9), 3)
        ok, buf = cv2.imencode(".jpg", img, [cv2.IMWRITE_JPEG_QUALITY, q])
        ref = cv2.cvtColor(cv2.imdecode(buf, cv2.IMREAD_COLOR), cv2.COLOR_BGR2RGB)
        assert np.array_equal(J.decode_rgb(buf.tobytes()), ref)


def test_oracle_rejects_what_the_collector_never_writes():
    cv2 = pytest.importorskip("cv2")
    img = np.zeros((32, 32, 3), np.uint8)
    ok, buf = cv2.imencode(".jpg", img, [cv2.IMWRITE_JPEG_PROGRESSIVE, 1])
    with pytest.raises(ValueError):
        J.decode_rgb(buf.tobytes())
    with pytest.raises(ValueError):
        J.decode_rgb(b"\x89PNG\r\n\x1a\n" + b"\0" * 64)


# ---- host half of the C-ABI ----------------------------------------------------------------------------------------------------
def _prepare(streams):
    from cilrs_b200 import _lib
    lib = _lib.lib()
    db, sb = lib.cilrs_jpeg_desc_bytes(), lib.cilrs_jpeg_table_set_bytes()
    offs = np.zeros(len(streams) + 1, dtype=np.int64)
    chunks = []
    for i, s in enumerate(streams):
        pad = (-len(s)) % 16
        chunks.append(s + b"\0" * pad)
        offs[i + 1] = offs[i] + len(s) + pad
    blob = np.frombuffer(b"".join(chunks) or b"\0", dtype=np.uint8).copy()
    descs = np.zeros(max(1, len(streams)) * db, dtype=np.uint8)
    sets = np.zeros(4 * sb, dtype=np.uint8)
    ns = ctypes.c_int(-1)
    st = lib.cilrs_jpeg_prepare(blob.ctypes.data_as(ctypes.c_void_p), offs.ctypes.data_as(ctypes.c_void_p), len(streams),
                                descs.ctypes.data_as(ctypes.c_void_p), sets.ctypes.data_as(ctypes.c_void_p), 4, ctypes.byref(ns))
    assert st == 0
    return descs.reshape(-1, db), sets.reshape(4, sb), ns.value, offs


def _desc_fields(d):
    u32 = d[:48].view("<u4")
    u16 = d[:48].view("<u2")
    return dict(data_off=int(u32[0]), data_len=int(u32[1]), scan_off=int(u32[2]), width=int(u16[6]), height=int(u16[7]), mode=int(d[16]),
                huff_set=int(d[17]), mcus_x=int(u16[14]), mcus_y=int(u16[15]), status=int(u32[8]), qt=d[48:].view("<u2").reshape(2, 64))


def test_prepare_parses_like_the_oracle():
    streams = [open(p, "rb").read() for p in JPEGS]
    descs, sets, ns, offs = _prepare(streams)
    assert ns == 2   # OpenCV's standard tables for colour frames; the grey frame carries only the luminance pair
    for i, s in enumerate(streams):
        f = _desc_fields(descs[i])
        h = J.parse(s)
        assert f["status"] == 0 and f["data_off"] == offs[i]
        assert (f["width"], f["height"], f["scan_off"]) == (h.width, h.height, h.scan_offset)
        want_mode = 2 if len(h.comps) == 1 else (0 if h.comps[0][1] == 2 else 1)
        assert f["mode"] == want_mode
        mcu = 16 if want_mode == 0 else 8
        assert (f["mcus_x"], f["mcus_y"]) == ((h.width + mcu - 1) // mcu, (h.height + mcu - 1) // mcu)
        assert np.array_equal(f["qt"][0], h.qt[h.comps[0][3]])
        if len(h.comps) == 3:
            assert np.array_equal(f["qt"][1], h.qt[h.comps[1][3]])


def test_prepare_builds_the_huffman_lookup_tables():
    s = open(JPEGS[0], "rb").read()
    descs, sets, ns, _ = _prepare([s])
    h = J.parse(s)
    tsz = 512 * 2 + 18 * 4 + 18 * 4 + 256
    for (tc, th), (counts, symbols) in h.huff.items():
        t = sets[0][(tc * 2 + th) * tsz:(tc * 2 + th + 1) * tsz]
        look = t[:1024].view("<u2")
        maxcode = t[1024:1024 + 72].view("<i4")
        valoff = t[1096:1096 + 72].view("<i4")
        huffval = t[1168:]
        code, k = 0, 0
        for length in range(1, 17):
            for _ in range(counts[length - 1]):
                if length <= 9:   # every 9-bit prefix extension of the code maps to (length, symbol)
                    lo = code << (9 - length)
                    assert all(look[lo + j] == ((length << 8) | symbols[k]) for j in range(1 << (9 - length)))
                else:
                    assert code <= maxcode[length] and huffval[valoff[length] + code] == symbols[k]
                code += 1
                k += 1
            assert maxcode[length] == (code - 1 if counts[length - 1] else -1)
            code <<= 1


def test_prepare_flags_bad_streams_without_failing_the_batch():
    cv2 = pytest.importorskip("cv2")
    good = open(JPEGS[0], "rb").read()
    ok, prog = cv2.imencode(".jpg", np.zeros((32, 32, 3), np.uint8), [cv2.IMWRITE_JPEG_PROGRESSIVE, 1])
    ok, rst = cv2.imencode(".jpg", np.zeros((32, 32, 3), np.uint8), [cv2.IMWRITE_JPEG_RST_INTERVAL, 2])
    descs, _, ns, _ = _prepare([good, b"not a jpeg at all", prog.tobytes(), good[:200], rst.tobytes(), b""])
    st = [_desc_fields(d)["status"] for d in descs]
    assert st[0] == 0 and st[1] == 1 and st[2] == 2 and st[3] == 3 and st[4] == 2 and st[5] == 1
    assert ns == 1


def test_read_files_packs_aligned_streams(tmp_path):
    from cilrs_b200 import _lib
    lib = _lib.lib()
    paths = JPEGS[:5]
    n = len(paths)
    arr = (ctypes.c_char_p * n)(*[os.fsencode(p) for p in paths])
    dst = np.zeros(1 << 20, dtype=np.uint8)
    offs = np.zeros(2 * n + 1, dtype=np.int64)
    assert lib.cilrs_jpeg_read_files(arr, n, dst.ctypes.data_as(ctypes.c_void_p), ctypes.c_longlong(dst.size),
                                     offs.ctypes.data_as(ctypes.c_void_p), 3) == 0
    for i, p in enumerate(paths):
        raw = open(p, "rb").read()
        assert offs[i] % 16 == 0 and offs[n + 1 + i] - offs[i] == len(raw)
        assert dst[offs[i]:offs[n + 1 + i]].tobytes() == raw
    # too small a buffer, and a missing file
    assert lib.cilrs_jpeg_read_files(arr, n, dst.ctypes.data_as(ctypes.c_void_p), ctypes.c_longlong(1000), offs.ctypes.data_as(ctypes.c_void_p), 2) == 3
    arr2 = (ctypes.c_char_p * 2)(os.fsencode(paths[0]), os.fsencode(str(tmp_path / "missing.jpg")))
    assert lib.cilrs_jpeg_read_files(arr2, 2, dst.ctypes.data_as(ctypes.c_void_p), ctypes.c_longlong(dst.size), offs.ctypes.data_as(ctypes.c_void_p), 2) == 1
    assert offs[2 + 1 + 1] == offs[1]


def test_load_sessions_and_class_weights(tmp_path):
    from cilrs_b200 import data
    names = ["LANEFOLLOW", "LEFT", "RIGHT", "STRAIGHT"]
    for s, rows in (("session1_town01", 7), ("session2_town02", 5), ("notes", 0)):
        d = tmp_path / s
        (d / "images").mkdir(parents=True)
        if rows == 0:
            continue
        with open(d / "measurements.csv", "w") as f:   # the collector's 14 columns (model/collect_data.py:549-564)
            f.write("frame,image_filename,steer,throttle,brake,speed_kmh,speed_normalized,high_level_command,command_name,"
                    "position_x,position_y,position_z,yaw,timestamp\n")
            for i in range(rows):
                f.write("%d,frame_%08d.jpg,%.6f,%.6f,%.6f,%.2f,%.6f,%d,%s,1.0,2.0,0.1,90.0,%.3f\n"
                        % (i, i, 0.01 * i, 0.5, 0.0, 3.6 * i, 3.6 * i / 90.0, i % 4, names[i % 4 if i < 4 else 0], 0.05 * i))
    (tmp_path / "notes" / "x").mkdir()
    t = data.load_sessions(str(tmp_path))
    assert len(t["image_path"]) == 12 and t["image_path"][7].endswith(os.path.join("session2_town02", "images", "frame_00000000.jpg"))
    assert t["command_idx"].tolist()[:7] == [0, 1, 2, 3, 0, 0, 0] and t["speed_normalized"].dtype == np.float32
    cw, w = data.class_weights(t["command_idx"])
    counts = np.bincount(t["command_idx"], minlength=4)
    assert all(abs(cw[c] - 12 / (4 * counts[c])) < 1e-12 for c in range(4))
    assert np.allclose(w, [cw[int(c)] for c in t["command_idx"]])
