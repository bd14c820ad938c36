"""TEST INFRASTRUCTURE - CPU restatement of the JPEG decode the reference's loader performs.

The reference reads every frame with `cv2.imread(path)` + `cv2.cvtColor(BGR2RGB)` (notebook/notebook.ipynb:404-405); the files are
the collector's `cv2.imwrite(..., [IMWRITE_JPEG_QUALITY, 95])` frames (model/collect_data.py:685-716): baseline sequential
Huffman JPEG, 8 bit, YCbCr 4:2:0 (2x2 luma sampling), no restart markers. OpenCV decodes them with its bundled libjpeg-turbo
(3.1.2 in this image; a third-party dependency that is not in /root/reference) at that library's defaults, whose published
algorithm this file restates:

  * Huffman entropy decoding (ITU T.81 F.2.2), DC prediction per component;
  * dequantisation + `jpeg_idct_islow` (jidctint.c: 13-bit fixed-point LL&M inverse DCT, CONST_BITS 13, PASS1_BITS 2);
  * `h2v2_fancy_upsample` (jdsample.c: triangle filter, 9/3/3/1 sixteenths, rounding constants 8 / 7 alternating by column,
    edge rows / columns replicated);
  * `ycc_rgb_convert` (jdcolor.c: 16-bit fixed-point tables, FIX(1.40200) ...).

Pinned: tests/test_jpeg_cpu.py checks `decode_rgb` bit for bit against `cv2.imdecode` (run here and on the GPU box - cv2 is part
of the image) on the committed fixtures under tests/golden/jpeg/ and on freshly encoded frames.
Only tests/, __graft_entry__.smoke() and bench.py's CPU legs may import this module.
"""
import numpy as np

ZIGZAG = np.array([0, 1, 8, 16, 9, 2, 3, 10, 17, 24, 32, 25, 18, 11, 4, 5, 12, 19, 26, 33, 40, 48, 41, 34, 27, 20, 13, 6, 7, 14, 21, 28,
                   35, 42, 49, 56, 57, 50, 43, 36, 29, 22, 15, 23, 30, 37, 44, 51, 58, 59, 52, 45, 38, 31, 39, 46, 53, 60, 61, 54, 47, 55, 62,
                   63], dtype=np.int32)


class JpegHeader:
    pass


def parse(data):
    """Marker segments up to the scan data. Returns a JpegHeader (width, height, comps [(id, h, v, tq)], scan table ids,
    qt [4][64] natural order, huff {(class, id): (counts[16], symbols)}, scan_offset)."""
    b = data
    if b[0] != 0xFF or b[1] != 0xD8:
        raise ValueError("not a JPEG stream")
    h = JpegHeader()
    h.qt = {}
    h.huff = {}
    h.restart_interval = 0
    i = 2
    while True:
        if b[i] != 0xFF:
            raise ValueError("marker expected at %d" % i)
        m = b[i + 1]
        if m == 0xFF:
            i += 1
            continue
        L = (b[i + 2] << 8) | b[i + 3]
        seg = b[i + 4:i + 2 + L]
        if m == 0xDB:
            j = 0
            while j < len(seg):
                pq, tq = seg[j] >> 4, seg[j] & 15
                if pq != 0:
                    raise ValueError("16-bit quantisation tables are not baseline")
                t = np.zeros(64, dtype=np.int32)
                t[ZIGZAG] = np.frombuffer(bytes(seg[j + 1:j + 65]), dtype=np.uint8)
                h.qt[tq] = t
                j += 65
        elif m == 0xC0 or m == 0xC1:
            if seg[0] != 8:
                raise ValueError("8-bit samples only")
            h.height = (seg[1] << 8) | seg[2]
            h.width = (seg[3] << 8) | seg[4]
            n = seg[5]
            h.comps = [(seg[6 + 3 * k], seg[7 + 3 * k] >> 4, seg[7 + 3 * k] & 15, seg[8 + 3 * k]) for k in range(n)]
        elif m in (0xC2, 0xC3, 0xC5, 0xC6, 0xC7, 0xC9, 0xCA, 0xCB, 0xCD, 0xCE, 0xCF):
            raise ValueError("only baseline sequential Huffman JPEG (SOF0) is supported, got SOF%d" % (m - 0xC0))
        elif m == 0xC4:
            j = 0
            while j < len(seg):
                tc, th = seg[j] >> 4, seg[j] & 15
                counts = list(seg[j + 1:j + 17])
                ns = sum(counts)
                h.huff[(tc, th)] = (counts, list(seg[j + 17:j + 17 + ns]))
                j += 17 + ns
        elif m == 0xDD:
            h.restart_interval = (seg[0] << 8) | seg[1]
        elif m == 0xDA:
            n = seg[0]
            h.scan = [(seg[1 + 2 * k], seg[2 + 2 * k] >> 4, seg[2 + 2 * k] & 15) for k in range(n)]
            h.scan_offset = i + 2 + L
            return h
        i += 2 + L


def _build_decode_table(counts, symbols):
    """code -> symbol maps per length (T.81 Annex C)."""
    table = {}
    code = 0
    k = 0
    for length in range(1, 17):
        for _ in range(counts[length - 1]):
            table[(length, code)] = symbols[k]
            code += 1
            k += 1
        code <<= 1
    return table


class _Bits:
    def __init__(self, data, pos):
        self.d = data
        self.p = pos
        self.acc = 0
        self.n = 0

    def bit(self):
        if self.n == 0:
            c = self.d[self.p]
            self.p += 1
            if c == 0xFF:
                c2 = self.d[self.p]
                if c2 == 0:
                    self.p += 1
                else:           # a marker (EOI): feed zeros like libjpeg's "insufficient data" path
                    self.p -= 1
                    c = 0
            self.acc = c
            self.n = 8
        self.n -= 1
        return (self.acc >> self.n) & 1

    def bits(self, k):
        v = 0
        for _ in range(k):
            v = (v << 1) | self.bit()
        return v


def _extend(v, t):
    return v if v >= (1 << (t - 1)) else v - (1 << t) + 1


def decode_coefficients(data):
    """Entropy decoding: returns (header, [per component int32 array [blocks_h, blocks_w, 64] of QUANTISED coefficients, natural order])."""
    h = parse(data)
    if h.restart_interval:
        raise ValueError("restart intervals are not produced by the collector (not supported)")
    hmax = max(c[1] for c in h.comps)
    vmax = max(c[2] for c in h.comps)
    mcu_w, mcu_h = 8 * hmax, 8 * vmax
    mcus_x = (h.width + mcu_w - 1) // mcu_w
    mcus_y = (h.height + mcu_h - 1) // mcu_h
    coefs = [np.zeros((mcus_y * c[2], mcus_x * c[1], 64), dtype=np.int32) for c in h.comps]
    tabs = {k: _build_decode_table(*v) for k, v in h.huff.items()}
    br = _Bits(data, h.scan_offset)
    pred = [0] * len(h.comps)
    scan_tab = {cid: (td, ta) for cid, td, ta in h.scan}

    def sym(tab):
        code = 0
        for length in range(1, 17):
            code = (code << 1) | br.bit()
            s = tab.get((length, code))
            if s is not None:
                return s
        raise ValueError("bad Huffman code")

    for my in range(mcus_y):
        for mx in range(mcus_x):
            for ci, (cid, ch, cv, tq) in enumerate(h.comps):
                td, ta = scan_tab[cid]
                for by in range(cv):
                    for bx in range(ch):
                        blk = coefs[ci][my * cv + by, mx * ch + bx]
                        t = sym(tabs[(0, td)])
                        diff = _extend(br.bits(t), t) if t else 0
                        pred[ci] += diff
                        blk[0] = pred[ci]
                        k = 1
                        while k < 64:
                            rs = sym(tabs[(1, ta)])
                            r, s = rs >> 4, rs & 15
                            if s == 0:
                                if r != 15:
                                    break
                                k += 16
                                continue
                            k += r
                            blk[ZIGZAG[k]] = _extend(br.bits(s), s)
                            k += 1
    h.mcus_x, h.mcus_y, h.hmax, h.vmax = mcus_x, mcus_y, hmax, vmax
    return h, coefs


# ---- jidctint.c: jpeg_idct_islow --------------------------------------------------------------------------------------------
_C = dict(F_0_298=2446, F_0_390=3196, F_0_541=4433, F_0_765=6270, F_0_899=7373, F_1_175=9633, F_1_501=12299, F_1_847=15137,
          F_1_961=16069, F_2_053=16819, F_2_562=20995, F_3_072=25172)


def _idct_1d(x0, x1, x2, x3, x4, x5, x6, x7, shift, pre_dc_shift):
    """One LL&M pass over int64 vectors (even part / odd part exactly as jidctint.c); returns 8 outputs DESCALEd by `shift`."""
    c = _C
    z2, z3 = x2, x6
    z1 = (z2 + z3) * c["F_0_541"]
    tmp2 = z1 + z3 * (-c["F_1_847"])
    tmp3 = z1 + z2 * c["F_0_765"]
    tmp0 = (x0 + x4) << 13
    tmp1 = (x0 - x4) << 13
    tmp10, tmp13 = tmp0 + tmp3, tmp0 - tmp3
    tmp11, tmp12 = tmp1 + tmp2, tmp1 - tmp2
    t0, t1, t2, t3 = x7, x5, x3, x1
    z1 = t0 + t3
    z2 = t1 + t2
    z3 = t0 + t2
    z4 = t1 + t3
    z5 = (z3 + z4) * c["F_1_175"]
    t0 = t0 * c["F_0_298"]
    t1 = t1 * c["F_2_053"]
    t2 = t2 * c["F_3_072"]
    t3 = t3 * c["F_1_501"]
    z1 = z1 * (-c["F_0_899"])
    z2 = z2 * (-c["F_2_562"])
    z3 = z3 * (-c["F_1_961"]) + z5
    z4 = z4 * (-c["F_0_390"]) + z5
    t0 = t0 + z1 + z3
    t1 = t1 + z2 + z4
    t2 = t2 + z2 + z3
    t3 = t3 + z1 + z4
    r = 1 << (shift - 1)
    return [(tmp10 + t3 + r) >> shift, (tmp11 + t2 + r) >> shift, (tmp12 + t1 + r) >> shift, (tmp13 + t0 + r) >> shift,
            (tmp13 - t0 + r) >> shift, (tmp12 - t1 + r) >> shift, (tmp11 - t2 + r) >> shift, (tmp10 - t3 + r) >> shift]


def idct_islow(coef, qt):
    """coef [..., 64] quantised coefficients (natural order), qt [64] -> uint8 [..., 8, 8]."""
    x = (coef.astype(np.int64) * qt.astype(np.int64)).reshape(coef.shape[:-1] + (8, 8))
    # pass 1: columns, results scaled up by PASS1_BITS = 2
    cols = _idct_1d(*[x[..., r, :] for r in range(8)], shift=13 - 2, pre_dc_shift=0)
    ws = np.stack(cols, axis=-2)                       # [..., 8 rows, 8 cols]
    # pass 2: rows; descale by CONST_BITS + PASS1_BITS + 3, +128 level shift, range limit
    rows = _idct_1d(*[ws[..., :, c] for c in range(8)], shift=13 + 2 + 3, pre_dc_shift=0)
    out = np.stack(rows, axis=-1) + 128
    return np.clip(out, 0, 255).astype(np.uint8)


def planes(data):
    """Decoded component planes at their native (down-sampled) resolution, padded to whole blocks."""
    h, coefs = decode_coefficients(data)
    out = []
    for (cid, ch, cv, tq), cf in zip(h.comps, coefs):
        px = idct_islow(cf, h.qt[tq])                                   # [bh, bw, 8, 8]
        out.append(px.transpose(0, 2, 1, 3).reshape(cf.shape[0] * 8, cf.shape[1] * 8))
    return h, out


def h2v2_fancy_upsample(plane, out_h, out_w):
    """jdsample.c h2v2_fancy_upsample on the real (un-padded) chroma samples: plane [ceil(out_h/2), ceil(out_w/2)] -> [out_h, out_w]."""
    ch, cw = (out_h + 1) // 2, (out_w + 1) // 2
    p = plane[:ch, :cw].astype(np.int32)
    up = np.concatenate([p[:1], p[:-1]], axis=0)        # the row above (edge replicated by the main controller's context rows)
    dn = np.concatenate([p[1:], p[-1:]], axis=0)
    res = np.zeros((2 * ch, 2 * cw), dtype=np.int32)
    for v, other in ((0, up), (1, dn)):
        colsum = 3 * p + other                            # thiscolsum per input column
        last = np.concatenate([colsum[:, :1], colsum[:, :-1]], axis=1)
        nxt = np.concatenate([colsum[:, 1:], colsum[:, -1:]], axis=1)
        even = (3 * colsum + last + 8) >> 4
        odd = (3 * colsum + nxt + 7) >> 4
        # special cases of the first and last columns
        even[:, 0] = (4 * colsum[:, 0] + 8) >> 4
        odd[:, -1] = (4 * colsum[:, -1] + 7) >> 4
        res[v::2, 0::2] = even
        res[v::2, 1::2] = odd
    return res[:out_h, :out_w].astype(np.uint8)


def _fix(x):
    return int(x * 65536 + 0.5)


_X = np.arange(256, dtype=np.int64) - 128
CR_R = ((_fix(1.40200) * _X + 32768) >> 16).astype(np.int32)
CB_B = ((_fix(1.77200) * _X + 32768) >> 16).astype(np.int32)
CR_G = (-_fix(0.71414) * _X).astype(np.int64)
CB_G = (-_fix(0.34414) * _X + 32768).astype(np.int64)


def ycc_to_rgb(y, cb, cr):
    y = y.astype(np.int32)
    r = y + CR_R[cr]
    g = y + ((CB_G[cb] + CR_G[cr]) >> 16).astype(np.int32)
    b = y + CB_B[cb]
    return np.clip(np.stack([r, g, b], axis=-1), 0, 255).astype(np.uint8)


def decode_rgb(data):
    """bytes of a baseline JPEG -> uint8 [H, W, 3] RGB, == cv2.cvtColor(cv2.imdecode(...), BGR2RGB)."""
    h, pl = planes(data)
    H, W = h.height, h.width
    if len(pl) == 1:
        g = pl[0][:H, :W]
        return np.stack([g, g, g], axis=-1)
    (_, yh, yv, _), (_, ch_, cv_, _) = h.comps[0], h.comps[1]
    y = pl[0][:H, :W]
    if (yh, yv) == (2, 2) and (ch_, cv_) == (1, 1):
        cb = h2v2_fancy_upsample(pl[1], H, W)
        cr = h2v2_fancy_upsample(pl[2], H, W)
    elif (yh, yv) == (1, 1):
        cb, cr = pl[1][:H, :W], pl[2][:H, :W]
    else:
        raise ValueError("sampling %dx%d is not produced by the collector (4:2:0 and 4:4:4 are supported)" % (yh, yv))
    return ycc_to_rgb(y, cb, cr)
