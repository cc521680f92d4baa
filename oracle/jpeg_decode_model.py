"""Integer model of the baseline JPEG decoder behind ``cv2.imread(path)`` for 8-bit YCbCr files (ref
``app/panorama_to_plane-pitch.py:244``).  TEST INFRASTRUCTURE (see ``oracle/__init__.py``).

The arithmetic lives in libjpeg-turbo (bundled in ``opencv-python``; not under ``/root/reference``) at its decoder
defaults, which OpenCV does not change: ``dct_method = JDCT_ISLOW``, ``do_fancy_upsampling = TRUE``.  Restated here from the
published sources, operation by operation:

* ``jdhuff.c``    Huffman decoding of one interleaved sequential scan (restart markers supported)
* ``jidctint.c``  jpeg_idct_islow (CONST_BITS 13, PASS1_BITS 2) on dequantised coefficients, + 128, clamp to 0..255
* ``jdsample.c``  h2v2_fancy_upsample / h2v1_fancy_upsample (triangle filters with their alternating rounding), edge rows
                  replicated (``jdmainct.c`` context rows)
* ``jdcolor.c``   ycc_rgb_convert: 16-bit fixed-point BT.601 tables, clamp

``decode`` is pinned against ``cv2.imdecode`` in ``tests/test_jpeg_oracle.py``.  Files this model does not cover
(progressive, non-interleaved scans, CMYK / grayscale, 12-bit, arithmetic coding) raise ``Unsupported``: the product
falls back to ``cv2.imread`` for those, as it does for an EXIF orientation other than 1.
"""
from __future__ import annotations

import numpy as np

from .jpeg_model import ZIGZAG


class Unsupported(ValueError):
    pass


def parse(data: bytes):
    """Markers up to the start of the (single) scan.  Returns a dict with the frame, tables and the entropy-coded bytes."""
    if data[:2] != b"\xff\xd8":
        raise Unsupported("not a JPEG")
    i = 2
    qt, huff = {}, {}
    frame, scan, dri = None, None, 0
    jfif, adobe = False, None
    progressive = False
    while i < len(data):
        if data[i] != 0xFF:
            raise Unsupported("marker expected")
        while data[i + 1] == 0xFF:
            i += 1
        m = data[i + 1]
        i += 2
        if m == 0xD9:
            break
        if 0xD0 <= m <= 0xD7 or m == 0x01:
            continue
        L = (data[i] << 8) | data[i + 1]
        seg = data[i + 2:i + L]
        i += L
        if m == 0xDB:
            j = 0
            while j < len(seg):
                pq, tq = seg[j] >> 4, seg[j] & 15
                j += 1
                if pq:
                    vals = [(seg[j + 2 * k] << 8) | seg[j + 2 * k + 1] for k in range(64)]
                    j += 128
                else:
                    vals = list(seg[j:j + 64])
                    j += 64
                q = np.zeros(64, np.int64)
                q[ZIGZAG] = vals            # stored in zigzag order -> natural order
                qt[tq] = q
        elif m == 0xC4:
            j = 0
            while j < len(seg):
                tc_th = seg[j]
                bits = list(seg[j + 1:j + 17])
                n = sum(bits)
                vals = list(seg[j + 17:j + 17 + n])
                huff[tc_th] = (bits, vals)
                j += 17 + n
        elif m in (0xC0, 0xC1, 0xC2):
            progressive = (m == 0xC2)
            if seg[0] != 8:
                raise Unsupported("sample precision")
            H, W, nc = (seg[1] << 8) | seg[2], (seg[3] << 8) | seg[4], seg[5]
            comps = [(seg[6 + 3 * k], seg[7 + 3 * k] >> 4, seg[7 + 3 * k] & 15, seg[8 + 3 * k]) for k in range(nc)]
            frame = dict(W=W, H=H, comps=comps)
        elif 0xC3 <= m <= 0xCF and m not in (0xC4, 0xC8, 0xCC):
            raise Unsupported("not a baseline / extended sequential / progressive Huffman JPEG")
        elif m == 0xDD:
            dri = (seg[0] << 8) | seg[1]
        elif m == 0xE0 and len(seg) >= 14 and seg[:5] == b"JFIF\x00":
            jfif = True
        elif m == 0xEE and len(seg) >= 12 and seg[:5] == b"Adobe":
            adobe = seg[11]
        elif m == 0xDA:
            ns = seg[0]
            scan = [(seg[1 + 2 * k], seg[2 + 2 * k] >> 4, seg[2 + 2 * k] & 15) for k in range(ns)]
            if frame is None or (ns != len(frame["comps"]) and not progressive):
                raise Unsupported("non-interleaved scans")
            # jdapimin.c default_decompress_parms, 3 components: JFIF -> YCbCr; else Adobe transform 0 -> RGB, other ->
            # YCbCr; else component ids 'R','G','B' -> RGB, anything else YCbCr
            ids = bytes(c[0] for c in frame["comps"])
            if len(ids) == 3 and not jfif and (adobe == 0 if adobe is not None else ids == b"RGB"):
                raise Unsupported("RGB-coded file")
            # (a progressive file: only the frame and the quantisation tables are of use here - the pixel pipeline behind the
            # coefficients, reconstruct(), is the same; its scans are not restated in this model)
            return dict(frame=frame, qt=qt, huff=huff, scan=scan, dri=dri, ecs=data[i:], progressive=progressive)
    raise Unsupported("no scan")


def build_decode_table(bits, vals):
    """code length / value lookup: dict (length, code) -> symbol"""
    table, code, k = {}, 0, 0
    for length in range(1, 17):
        for _ in range(bits[length - 1]):
            table[(length, code)] = vals[k]
            code += 1
            k += 1
        code <<= 1
    return table


class BitReader:
    def __init__(self, data: bytes):
        self.d, self.i, self.acc, self.n = data, 0, 0, 0

    def _fill(self):
        b = self.d[self.i] if self.i < len(self.d) else 0
        if b == 0xFF:
            nxt = self.d[self.i + 1] if self.i + 1 < len(self.d) else 0xD9
            if nxt == 0:
                self.i += 2
            else:
                b = 0  # a marker: feed zeros (jdhuff.c does the same at the end of the data)
        else:
            self.i += 1
        self.acc = (self.acc << 8) | b
        self.n += 8

    def bit(self) -> int:
        if self.n == 0:
            self._fill()
        self.n -= 1
        return (self.acc >> self.n) & 1

    def bits(self, k: int) -> int:
        v = 0
        for _ in range(k):
            v = (v << 1) | self.bit()
        return v

    def restart(self):
        self.acc, self.n = 0, 0
        # skip to the RSTn marker
        while not (self.d[self.i] == 0xFF and 0xD0 <= self.d[self.i + 1] <= 0xD7):
            self.i += 1
        self.i += 2


def decode_symbol(br, table):
    code = 0
    for length in range(1, 17):
        code = (code << 1) | br.bit()
        s = table.get((length, code))
        if s is not None:
            return s
    raise ValueError("bad Huffman code")


def extend(v, s):
    return v if v >= (1 << (s - 1)) else v - (1 << s) + 1


def entropy_decode(p):
    """Coefficient planes per component: [blocks_y, blocks_x, 64] natural order, on the MCU-padded block grid."""
    if p.get("progressive"):
        raise Unsupported("progressive scans are not restated in this model (see parse)")
    f = p["frame"]
    comps = f["comps"]
    if len(comps) == 1:          # a single-component scan is not interleaved: one block per MCU whatever the sampling factors
        comps = [(comps[0][0], 1, 1, comps[0][3])]
    hmax, vmax = max(c[1] for c in comps), max(c[2] for c in comps)
    mcux = -(-f["W"] // (8 * hmax))
    mcuy = -(-f["H"] // (8 * vmax))
    planes = [np.zeros((mcuy * c[2], mcux * c[1], 64), np.int64) for c in comps]
    sel = {cid: (td, ta) for cid, td, ta in p["scan"]}
    dct = {k: build_decode_table(*v) for k, v in p["huff"].items()}
    br = BitReader(p["ecs"])
    pred = [0] * len(comps)
    count = 0
    for my in range(mcuy):
        for mx in range(mcux):
            if p["dri"] and count and count % p["dri"] == 0:
                br.restart()
                pred = [0] * len(comps)
            count += 1
            for ci, (cid, h, v, tq) in enumerate(comps):
                td, ta = sel[cid]
                for yi in range(v):
                    for xi in range(h):
                        blk = planes[ci][my * v + yi, mx * h + xi]
                        s = decode_symbol(br, dct[td])
                        diff = extend(br.bits(s), s) if s else 0
                        pred[ci] += diff
                        blk[0] = pred[ci]
                        k = 1
                        while k < 64:
                            rs = decode_symbol(br, dct[0x10 | ta])
                            r, s = rs >> 4, rs & 15
                            if s == 0:
                                if r != 15:
                                    break
                                k += 16
                                continue
                            k += r
                            blk[ZIGZAG[k]] = extend(br.bits(s), s)
                            k += 1
    return planes, (mcux, mcuy, hmax, vmax)


def idct_islow(coef: np.ndarray, q: np.ndarray) -> np.ndarray:
    """jidctint.c jpeg_idct_islow: [..., 64] natural-order coefficients -> [..., 8, 8] samples 0..255."""
    C = dict(c0_298=2446, c0_390=3196, c0_541=4433, c0_765=6270, c0_899=7373, c1_175=9633, c1_501=12299, c1_847=15137,
             c1_961=16069, c2_053=16819, c2_562=20995, c3_072=25172)

    def descale(x, n):
        return (x + (1 << (n - 1))) >> n

    def pass_(d, n):
        # d[..., k] = input k of the 1-D transform
        z2, z3 = d[..., 2], d[..., 6]
        z1 = (z2 + z3) * C["c0_541"]
        tmp2 = z1 + z3 * (-C["c1_847"])
        tmp3 = z1 + z2 * C["c0_765"]
        z2, z3 = d[..., 0], d[..., 4]
        tmp0, tmp1 = (z2 + z3) << 13, (z2 - z3) << 13
        tmp10, tmp13, tmp11, tmp12 = tmp0 + tmp3, tmp0 - tmp3, tmp1 + tmp2, tmp1 - tmp2
        tmp0, tmp1, tmp2, tmp3 = d[..., 7], d[..., 5], d[..., 3], d[..., 1]
        z1, z2, z3, z4 = tmp0 + tmp3, tmp1 + tmp2, tmp0 + tmp2, tmp1 + tmp3
        z5 = (z3 + z4) * C["c1_175"]
        tmp0, tmp1, tmp2, tmp3 = tmp0 * C["c0_298"], tmp1 * C["c2_053"], tmp2 * C["c3_072"], tmp3 * C["c1_501"]
        z1, z2 = z1 * (-C["c0_899"]), z2 * (-C["c2_562"])
        z3, z4 = z3 * (-C["c1_961"]) + z5, z4 * (-C["c0_390"]) + z5
        tmp0, tmp1, tmp2, tmp3 = tmp0 + z1 + z3, tmp1 + z2 + z4, tmp2 + z2 + z3, tmp3 + z1 + z4
        out = np.empty_like(d)
        out[..., 0], out[..., 7] = descale(tmp10 + tmp3, n), descale(tmp10 - tmp3, n)
        out[..., 1], out[..., 6] = descale(tmp11 + tmp2, n), descale(tmp11 - tmp2, n)
        out[..., 2], out[..., 5] = descale(tmp12 + tmp1, n), descale(tmp12 - tmp1, n)
        out[..., 3], out[..., 4] = descale(tmp13 + tmp0, n), descale(tmp13 - tmp0, n)
        return out

    blk = (coef * q).reshape(coef.shape[:-1] + (8, 8))          # [row, col]
    ws = np.swapaxes(pass_(np.swapaxes(blk, -1, -2), 13 - 2), -1, -2)   # pass 1: columns
    out = pass_(ws, 13 + 2 + 3)                                  # pass 2: rows
    return np.clip(out + 128, 0, 255)


def plane_from_blocks(b: np.ndarray) -> np.ndarray:
    by, bx = b.shape[:2]
    return b.swapaxes(1, 2).reshape(by * 8, bx * 8)


def h2v1_fancy(c: np.ndarray) -> np.ndarray:
    """jdsample.c h2v1_fancy_upsample on [rows, w] -> [rows, 2 w]"""
    c = c.astype(np.int64)
    w = c.shape[1]
    out = np.empty((c.shape[0], 2 * w), np.int64)
    if w == 1:
        out[:, 0] = out[:, 1] = c[:, 0]
        return out
    left = np.concatenate([c[:, :1], c[:, :-1]], axis=1)
    right = np.concatenate([c[:, 1:], c[:, -1:]], axis=1)
    out[:, 0::2] = (3 * c + left + 1) >> 2
    out[:, 1::2] = (3 * c + right + 2) >> 2
    out[:, 0] = c[:, 0]
    out[:, -1] = c[:, -1]
    return out


def h2v2_fancy(c: np.ndarray) -> np.ndarray:
    """jdsample.c h2v2_fancy_upsample on [h, w] -> [2 h, 2 w]; context rows replicate the edge rows"""
    c = c.astype(np.int64)
    h, w = c.shape
    above = np.concatenate([c[:1], c[:-1]], axis=0)
    below = np.concatenate([c[1:], c[-1:]], axis=0)
    out = np.empty((2 * h, 2 * w), np.int64)
    for v, far in ((0, above), (1, below)):
        colsum = 3 * c + far                                  # thiscolsum per column
        if w == 1:
            out[v::2, 0] = (colsum[:, 0] * 4 + 8) >> 4
            out[v::2, 1] = (colsum[:, 0] * 4 + 7) >> 4
            continue
        last = np.concatenate([colsum[:, :1], colsum[:, :-1]], axis=1)
        nxt = np.concatenate([colsum[:, 1:], colsum[:, -1:]], axis=1)
        even = (colsum * 3 + last + 8) >> 4
        odd = (colsum * 3 + nxt + 7) >> 4
        even[:, 0] = (colsum[:, 0] * 4 + 8) >> 4
        odd[:, -1] = (colsum[:, -1] * 4 + 7) >> 4
        out[v::2, 0::2] = even
        out[v::2, 1::2] = odd
    return out


def ycc_to_bgr(y, cb, cr):
    """jdcolor.c ycc_rgb_convert (build_ycc_rgb_table), BGR order"""
    y, cb, cr = y.astype(np.int64), cb.astype(np.int64) - 128, cr.astype(np.int64) - 128
    half = 1 << 15
    r = y + ((91881 * cr + half) >> 16)
    b = y + ((116130 * cb + half) >> 16)
    g = y + ((-22554 * cb + half - 46802 * cr) >> 16)
    return np.clip(np.stack([b, g, r], -1), 0, 255).astype(np.uint8)


def decode(data: bytes) -> np.ndarray:
    """The array ``cv2.imdecode(data, cv2.IMREAD_COLOR)`` returns for a supported file."""
    p = parse(data)
    f = p["frame"]
    comps = f["comps"]
    if len(comps) not in (1, 3):
        raise Unsupported("not a grayscale or 3-component YCbCr file")
    planes, _ = entropy_decode(p)
    return reconstruct(p, planes)


def reconstruct(p, planes) -> np.ndarray:
    """Pixels from the coefficient planes of ``entropy_decode`` (dequantisation, IDCT, upsampling, colour)."""
    f = p["frame"]
    comps = f["comps"]
    W, H = f["W"], f["H"]
    if len(comps) == 1:          # grayscale: jdcolor.c gray_rgb_convert, B = G = R = Y
        y = plane_from_blocks(idct_islow(planes[0], p["qt"][comps[0][3]]))[:H, :W]
        return np.repeat(y[..., None], 3, axis=2).astype(np.uint8)
    hmax, vmax = max(c[1] for c in comps), max(c[2] for c in comps)
    if (comps[0][1], comps[0][2]) != (hmax, vmax) or any((c[1], c[2]) != (1, 1) for c in comps[1:]):
        raise Unsupported("sampling factors")
    if (hmax, vmax) not in ((1, 1), (2, 1), (2, 2)):
        raise Unsupported("sampling factors")
    samples = [plane_from_blocks(idct_islow(pl, p["qt"][c[3]])) for pl, c in zip(planes, comps)]
    y = samples[0][:H, :W]
    cw, ch = -(-W // hmax), -(-H // vmax)          # downsampled_width / height of the chroma components
    ups = []
    for s in samples[1:]:
        s = s[:ch, :cw]
        fancy = cw > 2      # jdsample.c jinit_upsampler: fancy upsampling only if downsampled_width > 2, else replication
        if (hmax, vmax) == (2, 2):
            s = h2v2_fancy(s) if fancy else np.repeat(np.repeat(s, 2, axis=0), 2, axis=1)
        elif (hmax, vmax) == (2, 1):
            s = h2v1_fancy(s) if fancy else np.repeat(s, 2, axis=1)
        ups.append(s[:H, :W])
    return ycc_to_bgr(y, ups[0], ups[1])
