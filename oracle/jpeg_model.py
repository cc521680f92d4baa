"""Integer model of the baseline JPEG encoder behind ``cv2.imwrite(path.jpg, bgr)`` (ref
``app/panorama_to_plane-pitch.py:277`` with ``--output_format jpg|jpeg``, ref :400-405).
TEST INFRASTRUCTURE (see ``oracle/__init__.py``).

The arithmetic lives in libjpeg-turbo (bundled in ``opencv-python``, pinned 4.10.0.84 by the reference's
``requirements.txt:13``; 4.13.0 / libjpeg-turbo 3.1.2 in this image), not under ``/root/reference``.  OpenCV calls it with
its defaults: quality 95, baseline sequential, 4:2:0 (luma 2x2), standard Huffman tables (no optimisation), no
restart markers, JFIF 1.01 header with 1:1 density.  Every stage of that path is exact integer arithmetic, restated
here operation by operation from the published sources:

* ``jccolor.c``   rgb_ycc_convert: 16-bit fixed-point BT.601 with the (1 << 15) / (1 << 15) - 1 rounding offsets
* ``jcsample.c``  h2v2_downsample (bias 1, 2, 1, 2 ...) after expand_right_edge; ``jcprepct.c`` bottom-edge replication
* ``jfdctint.c``  jpeg_fdct_islow (CONST_BITS 13, PASS1_BITS 2; output scaled by 8)
* ``jcdctmgr.c``  quantisation: round half away from zero of coef / (8 q)   (the reciprocal form is exact for 16 bits)
* ``jccoefct.c``  dummy blocks at the right / bottom edge: zero AC, DC of the previous block of the MCU buffer
* ``jchuff.c``    Huffman coding with the Annex K tables, 0xFF byte stuffing, final padding with 1-bits
* ``jcmarker.c``  SOI, APP0 (JFIF), DQT x 2, SOF0, DHT x 4, SOS ... EOI

``encode`` is pinned byte for byte against ``cv2.imencode('.jpg', img)`` in ``tests/test_jpeg_oracle.py``.
"""
from __future__ import annotations

import numpy as np

ZIGZAG = np.array([
    0, 1, 8, 16, 9, 2, 3, 10, 17, 24, 32, 25, 18, 11, 4, 5, 12, 19, 26, 33, 40, 48, 41, 34, 27, 20, 13, 6, 7, 14, 21, 28,
    35, 42, 49, 56, 57, 50, 43, 36, 29, 22, 15, 23, 30, 37, 44, 51, 58, 59, 52, 45, 38, 31, 39, 46, 53, 60, 61, 54, 47, 55,
    62, 63], dtype=np.int64)

STD_LUMA_Q = np.array([
    16, 11, 10, 16, 24, 40, 51, 61, 12, 12, 14, 19, 26, 58, 60, 55, 14, 13, 16, 24, 40, 57, 69, 56, 14, 17, 22, 29, 51, 87,
    80, 62, 18, 22, 37, 56, 68, 109, 103, 77, 24, 35, 55, 64, 81, 104, 113, 92, 49, 64, 78, 87, 103, 121, 120, 101, 72, 92,
    95, 98, 112, 100, 103, 99], dtype=np.int64)
STD_CHROMA_Q = np.array([
    17, 18, 24, 47, 99, 99, 99, 99, 18, 21, 26, 66, 99, 99, 99, 99, 24, 26, 56, 99, 99, 99, 99, 99, 47, 66, 99, 99, 99, 99,
    99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99,
    99, 99, 99, 99], dtype=np.int64)

# Annex K Huffman tables: (bits[1..16], values)
DC_LUMA = ([0, 1, 5, 1, 1, 1, 1, 1, 1, 0, 0, 0, 0, 0, 0, 0], list(range(12)))
DC_CHROMA = ([0, 3, 1, 1, 1, 1, 1, 1, 1, 1, 1, 0, 0, 0, 0, 0], list(range(12)))
AC_LUMA = ([0, 2, 1, 3, 3, 2, 4, 3, 5, 5, 4, 4, 0, 0, 1, 0x7d], [
    0x01, 0x02, 0x03, 0x00, 0x04, 0x11, 0x05, 0x12, 0x21, 0x31, 0x41, 0x06, 0x13, 0x51, 0x61, 0x07, 0x22, 0x71, 0x14, 0x32,
    0x81, 0x91, 0xa1, 0x08, 0x23, 0x42, 0xb1, 0xc1, 0x15, 0x52, 0xd1, 0xf0, 0x24, 0x33, 0x62, 0x72, 0x82, 0x09, 0x0a, 0x16,
    0x17, 0x18, 0x19, 0x1a, 0x25, 0x26, 0x27, 0x28, 0x29, 0x2a, 0x34, 0x35, 0x36, 0x37, 0x38, 0x39, 0x3a, 0x43, 0x44, 0x45,
    0x46, 0x47, 0x48, 0x49, 0x4a, 0x53, 0x54, 0x55, 0x56, 0x57, 0x58, 0x59, 0x5a, 0x63, 0x64, 0x65, 0x66, 0x67, 0x68, 0x69,
    0x6a, 0x73, 0x74, 0x75, 0x76, 0x77, 0x78, 0x79, 0x7a, 0x83, 0x84, 0x85, 0x86, 0x87, 0x88, 0x89, 0x8a, 0x92, 0x93, 0x94,
    0x95, 0x96, 0x97, 0x98, 0x99, 0x9a, 0xa2, 0xa3, 0xa4, 0xa5, 0xa6, 0xa7, 0xa8, 0xa9, 0xaa, 0xb2, 0xb3, 0xb4, 0xb5, 0xb6,
    0xb7, 0xb8, 0xb9, 0xba, 0xc2, 0xc3, 0xc4, 0xc5, 0xc6, 0xc7, 0xc8, 0xc9, 0xca, 0xd2, 0xd3, 0xd4, 0xd5, 0xd6, 0xd7, 0xd8,
    0xd9, 0xda, 0xe1, 0xe2, 0xe3, 0xe4, 0xe5, 0xe6, 0xe7, 0xe8, 0xe9, 0xea, 0xf1, 0xf2, 0xf3, 0xf4, 0xf5, 0xf6, 0xf7, 0xf8,
    0xf9, 0xfa])
AC_CHROMA = ([0, 2, 1, 2, 4, 4, 3, 4, 7, 5, 4, 4, 0, 1, 2, 0x77], [
    0x00, 0x01, 0x02, 0x03, 0x11, 0x04, 0x05, 0x21, 0x31, 0x06, 0x12, 0x41, 0x51, 0x07, 0x61, 0x71, 0x13, 0x22, 0x32, 0x81,
    0x08, 0x14, 0x42, 0x91, 0xa1, 0xb1, 0xc1, 0x09, 0x23, 0x33, 0x52, 0xf0, 0x15, 0x62, 0x72, 0xd1, 0x0a, 0x16, 0x24, 0x34,
    0xe1, 0x25, 0xf1, 0x17, 0x18, 0x19, 0x1a, 0x26, 0x27, 0x28, 0x29, 0x2a, 0x35, 0x36, 0x37, 0x38, 0x39, 0x3a, 0x43, 0x44,
    0x45, 0x46, 0x47, 0x48, 0x49, 0x4a, 0x53, 0x54, 0x55, 0x56, 0x57, 0x58, 0x59, 0x5a, 0x63, 0x64, 0x65, 0x66, 0x67, 0x68,
    0x69, 0x6a, 0x73, 0x74, 0x75, 0x76, 0x77, 0x78, 0x79, 0x7a, 0x82, 0x83, 0x84, 0x85, 0x86, 0x87, 0x88, 0x89, 0x8a, 0x92,
    0x93, 0x94, 0x95, 0x96, 0x97, 0x98, 0x99, 0x9a, 0xa2, 0xa3, 0xa4, 0xa5, 0xa6, 0xa7, 0xa8, 0xa9, 0xaa, 0xb2, 0xb3, 0xb4,
    0xb5, 0xb6, 0xb7, 0xb8, 0xb9, 0xba, 0xc2, 0xc3, 0xc4, 0xc5, 0xc6, 0xc7, 0xc8, 0xc9, 0xca, 0xd2, 0xd3, 0xd4, 0xd5, 0xd6,
    0xd7, 0xd8, 0xd9, 0xda, 0xe2, 0xe3, 0xe4, 0xe5, 0xe6, 0xe7, 0xe8, 0xe9, 0xea, 0xf2, 0xf3, 0xf4, 0xf5, 0xf6, 0xf7, 0xf8,
    0xf9, 0xfa])


def quant_table(base: np.ndarray, quality: int) -> np.ndarray:
    """jpeg_quality_scaling + jpeg_add_quant_table(force_baseline = TRUE), natural order."""
    quality = max(1, min(100, int(quality)))
    scale = 5000 // quality if quality < 50 else 200 - quality * 2
    q = (base * scale + 50) // 100
    return np.clip(q, 1, 255)


def derive_codes(bits, vals):
    """jchuff.c jpeg_make_c_derived_tbl: symbol -> (code, length)."""
    code, k = 0, 0
    ehufco, ehufsi = np.zeros(256, np.int64), np.zeros(256, np.int64)
    for length in range(1, 17):
        for _ in range(bits[length - 1]):
            ehufco[vals[k]] = code
            ehufsi[vals[k]] = length
            code += 1
            k += 1
        code <<= 1
    return ehufco, ehufsi


def color_convert(bgr: np.ndarray):
    """jccolor.c (JCS_EXT_BGR -> YCbCr), exact 16-bit fixed point."""
    b = bgr[..., 0].astype(np.int64)
    g = bgr[..., 1].astype(np.int64)
    r = bgr[..., 2].astype(np.int64)
    half, off = 1 << 15, 128 << 16
    y = (19595 * r + 38470 * g + 7471 * b + half) >> 16
    cb = (-11059 * r - 21709 * g + 32768 * b + off + half - 1) >> 16
    cr = (32768 * r - 27439 * g - 5329 * b + off + half - 1) >> 16
    return y, cb, cr


def pad_edge(a: np.ndarray, rows: int, cols: int) -> np.ndarray:
    """replicate the last column / row (expand_right_edge, expand_bottom_edge)"""
    return np.pad(a, ((0, rows - a.shape[0]), (0, cols - a.shape[1])), mode="edge")


def h2v2_downsample(c: np.ndarray) -> np.ndarray:
    """jcsample.c h2v2_downsample: (a + b + c + d + bias) >> 2 with bias 1, 2, 1, 2 ... along the row"""
    s = c[0::2, 0::2] + c[0::2, 1::2] + c[1::2, 0::2] + c[1::2, 1::2]
    bias = np.where(np.arange(s.shape[1]) % 2 == 0, 1, 2)[None, :]
    return (s + bias) >> 2


def fdct_islow(blocks: np.ndarray) -> np.ndarray:
    """jfdctint.c jpeg_fdct_islow on [..., 8, 8] level-shifted samples; exact integers (output scaled by 8)."""
    CONST_BITS, PASS1_BITS = 13, 2
    F = dict(c0_298=2446, c0_390=3196, c0_541=4433, c0_765=6270, c0_899=7373, c1_175=9633, c1_501=12299,
             c1_847=15137, c1_961=16069, c2_053=16819, c2_562=20995, c3_072=25172)

    def descale(x, n):
        return (x + (1 << (n - 1))) >> n

    def pass_(d, first):
        t0, t7 = d[..., 0] + d[..., 7], d[..., 0] - d[..., 7]
        t1, t6 = d[..., 1] + d[..., 6], d[..., 1] - d[..., 6]
        t2, t5 = d[..., 2] + d[..., 5], d[..., 2] - d[..., 5]
        t3, t4 = d[..., 3] + d[..., 4], d[..., 3] - d[..., 4]
        t10, t13, t11, t12 = t0 + t3, t0 - t3, t1 + t2, t1 - t2
        out = np.empty_like(d)
        if first:
            out[..., 0] = (t10 + t11) << PASS1_BITS
            out[..., 4] = (t10 - t11) << PASS1_BITS
            n = CONST_BITS - PASS1_BITS
        else:
            out[..., 0] = descale(t10 + t11, PASS1_BITS)
            out[..., 4] = descale(t10 - t11, PASS1_BITS)
            n = CONST_BITS + PASS1_BITS
        z1 = (t12 + t13) * F["c0_541"]
        out[..., 2] = descale(z1 + t13 * F["c0_765"], n)
        out[..., 6] = descale(z1 + t12 * (-F["c1_847"]), n)
        z1, z2, z3, z4 = t4 + t7, t5 + t6, t4 + t6, t5 + t7
        z5 = (z3 + z4) * F["c1_175"]
        t4, t5, t6, t7 = t4 * F["c0_298"], t5 * F["c2_053"], t6 * F["c3_072"], t7 * F["c1_501"]
        z1, z2 = z1 * (-F["c0_899"]), z2 * (-F["c2_562"])
        z3, z4 = z3 * (-F["c1_961"]) + z5, z4 * (-F["c0_390"]) + z5
        out[..., 7] = descale(t4 + z1 + z3, n)
        out[..., 5] = descale(t5 + z2 + z4, n)
        out[..., 3] = descale(t6 + z2 + z3, n)
        out[..., 1] = descale(t7 + z1 + z4, n)
        return out

    d = pass_(blocks.astype(np.int64), True)            # rows
    d = pass_(np.swapaxes(d, -1, -2), False)            # columns
    return np.swapaxes(d, -1, -2)


def quantise(coef: np.ndarray, q: np.ndarray) -> np.ndarray:
    """jcdctmgr.c: sign(x) * floor((|x| + d / 2) / d), d = 8 q (islow output is scaled by 8)"""
    d = (q.reshape(8, 8) * 8).astype(np.int64)
    a = (np.abs(coef) + (d >> 1)) // d
    return np.where(coef < 0, -a, a)


def to_blocks(plane: np.ndarray) -> np.ndarray:
    h, w = plane.shape
    return plane.reshape(h // 8, 8, w // 8, 8).swapaxes(1, 2)  # [by, bx, 8, 8]


def component_coefficients(bgr: np.ndarray, quality: int = 95):
    """Quantised coefficients in natural order per component, laid out on the MCU-padded block grid, with the dummy
    blocks of jccoefct.c still unset (flag arrays say which blocks are real)."""
    H, W, _ = bgr.shape
    y, cb, cr = color_convert(bgr)
    mcux, mcuy = (W + 15) // 16, (H + 15) // 16
    qy, qc = quant_table(STD_LUMA_Q, quality), quant_table(STD_CHROMA_Q, quality)
    # real block counts (jcmaster.c): ceil(ceil(dim * samp / 2) / 8)
    ybw, ybh = (W + 7) // 8, (H + 7) // 8
    cw, ch = (W + 1) // 2, (H + 1) // 2
    cbw, cbh = (cw + 7) // 8, (ch + 7) // 8
    # luma: right edge to its real block width, bottom edge to the full iMCU height
    yp = pad_edge(y, mcuy * 16, ybw * 8)
    # chroma: the full-resolution planes are padded to 2 * (real block width * 8) columns and an even number of
    # rows (jcprepct row groups) before h2v2, the downsampled plane then to the iMCU height
    cbp = pad_edge(h2v2_downsample(pad_edge(cb, 2 * ch, 2 * cbw * 8)), mcuy * 8, cbw * 8)
    crp = pad_edge(h2v2_downsample(pad_edge(cr, 2 * ch, 2 * cbw * 8)), mcuy * 8, cbw * 8)
    out = []
    for plane, q, (bw, bh), (gw, gh) in ((yp, qy, (ybw, ybh), (mcux * 2, mcuy * 2)), (cbp, qc, (cbw, cbh), (mcux, mcuy)),
                                         (crp, qc, (cbw, cbh), (mcux, mcuy))):
        blocks = to_blocks(plane - 128)
        coef = quantise(fdct_islow(blocks), q)
        grid = np.zeros((gh, gw, 8, 8), np.int64)
        real = np.zeros((gh, gw), bool)
        grid[:coef.shape[0], :bw] = coef[:, :bw]
        real[:bh, :bw] = True
        out.append((grid, real))
    return out, (mcux, mcuy), (qy, qc)


def mcu_block_sequence(comps, mcux, mcuy):
    """Blocks in scan order ([n, 64] natural order, component id per block) with jccoefct.c's dummy blocks filled in:
    zero AC, DC = DC of the previous block in the MCU buffer."""
    (yg, yr), (cbg, cbr), (crg, crr) = comps
    seq, comp = [], []
    for my in range(mcuy):
        for mx in range(mcux):
            buf = []
            for ci, (g, r, hs, vs) in enumerate(((yg, yr, 2, 2), (cbg, cbr, 1, 1), (crg, crr, 1, 1))):
                for yi in range(vs):
                    for xi in range(hs):
                        by, bx = my * vs + yi, mx * hs + xi
                        if r[by, bx]:
                            blk = g[by, bx].reshape(64).copy()
                        else:
                            blk = np.zeros(64, np.int64)
                            blk[0] = buf[-1][0]
                        buf.append(blk)
                        comp.append(ci)
            seq.extend(buf)
    return np.array(seq), np.array(comp)


class BitWriter:
    def __init__(self):
        self.acc, self.n, self.out = 0, 0, bytearray()

    def put(self, code: int, size: int):
        self.acc = (self.acc << size) | (code & ((1 << size) - 1))
        self.n += size
        while self.n >= 8:
            byte = (self.acc >> (self.n - 8)) & 0xFF
            self.out.append(byte)
            if byte == 0xFF:
                self.out.append(0)
            self.n -= 8
        self.acc &= (1 << self.n) - 1

    def flush(self):
        if self.n:
            self.put(0x7F, 7)  # pad with 1-bits (jchuff.c flush_bits)
            self.acc, self.n = 0, 0


def nbits(v: int) -> int:
    return int(abs(int(v))).bit_length()


def entropy_encode(seq: np.ndarray, comp: np.ndarray) -> bytes:
    dc = [derive_codes(*DC_LUMA), derive_codes(*DC_CHROMA)]
    ac = [derive_codes(*AC_LUMA), derive_codes(*AC_CHROMA)]
    bw = BitWriter()
    last = [0, 0, 0]
    for blk, ci in zip(seq, comp):
        t = 0 if ci == 0 else 1
        zz = blk[ZIGZAG]
        diff = int(zz[0]) - last[ci]
        last[ci] = int(zz[0])
        s = nbits(diff)
        bw.put(int(dc[t][0][s]), int(dc[t][1][s]))
        if s:
            bw.put(diff if diff >= 0 else diff - 1, s)
        run = 0
        for k in range(1, 64):
            v = int(zz[k])
            if v == 0:
                run += 1
                continue
            while run > 15:
                bw.put(int(ac[t][0][0xF0]), int(ac[t][1][0xF0]))
                run -= 16
            s = nbits(v)
            sym = (run << 4) | s
            bw.put(int(ac[t][0][sym]), int(ac[t][1][sym]))
            bw.put(v if v >= 0 else v - 1, s)
            run = 0
        if run:
            bw.put(int(ac[t][0][0]), int(ac[t][1][0]))
    bw.flush()
    return bytes(bw.out)


def header(W: int, H: int, qy: np.ndarray, qc: np.ndarray) -> bytes:
    def seg(marker, payload):
        return bytes([0xFF, marker]) + (len(payload) + 2).to_bytes(2, "big") + payload

    out = b"\xff\xd8"
    out += seg(0xE0, b"JFIF\x00\x01\x01\x00\x00\x01\x00\x01\x00\x00")
    out += seg(0xDB, bytes([0]) + bytes(int(x) for x in qy[ZIGZAG]))
    out += seg(0xDB, bytes([1]) + bytes(int(x) for x in qc[ZIGZAG]))
    out += seg(0xC0, bytes([8]) + H.to_bytes(2, "big") + W.to_bytes(2, "big") + bytes([3, 1, 0x22, 0, 2, 0x11, 1, 3, 0x11, 1]))
    for tc_th, (bits, vals) in ((0x00, DC_LUMA), (0x10, AC_LUMA), (0x01, DC_CHROMA), (0x11, AC_CHROMA)):
        out += seg(0xC4, bytes([tc_th]) + bytes(bits) + bytes(vals))
    out += seg(0xDA, bytes([3, 1, 0x00, 2, 0x11, 3, 0x11, 0, 63, 0]))
    return out


def encode(bgr: np.ndarray, quality: int = 95) -> bytes:
    """The bytes ``cv2.imencode('.jpg', bgr)`` produces (default parameters)."""
    H, W, _ = bgr.shape
    comps, (mcux, mcuy), (qy, qc) = component_coefficients(bgr, quality)
    seq, comp = mcu_block_sequence(comps, mcux, mcuy)
    return header(W, H, qy, qc) + entropy_encode(seq, comp) + b"\xff\xd9"
