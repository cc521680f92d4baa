"""CPU oracle of the PNG *decoder* (TEST INFRASTRUCTURE ONLY - never imported by the product package).

The reference reads its inputs with ``cv2.imread(path)`` (ref ``/root/reference/app/panorama_to_plane-pitch.py:244``); for a
``.png`` file that is libpng + zlib behind OpenCV (opencv-python pinned 4.10.0.84 by the reference, 4.13 here; neither is
vendored in ``/root/reference``).  This module restates the published algorithms the device decoder implements
(``csrc/p2p_pngdec.cuh``) in NumPy + the standard library's ``zlib``:

* chunk walk and IHDR rules of the PNG specification (sections 5, 11.2.2), the subset the device decoder accepts;
* ``zlib.decompress`` for the deflate stream (RFC 1950 / 1951) - the decoder's own inflate is checked against it;
* the five scanline filters (PNG specification 9.2; libpng ``png_read_filter_row``);
* what ``cv2.imread(path)`` (flag ``IMREAD_COLOR``) makes of the colour types: gray replicated, alpha dropped, RGB -> BGR.

Pinned against ``cv2.imdecode`` itself in ``tests/test_png_decode_oracle.py`` (every colour type, every filter, many
writers) - parity is pinned to the real library, not to this restatement.

It also holds the writers that generate test files no OpenCV call produces: adaptive filters, every colour type, arbitrary
IDAT splits, every zlib level / strategy / window size, flushes in the middle of the stream, ancillary chunks, and the
damage patterns (``tests/``, ``tools/``).
"""
from __future__ import annotations

import struct
import zlib

import numpy as np

SIG = b"\x89PNG\r\n\x1a\n"
BPP = {0: 1, 2: 3, 4: 2, 6: 4}


def chunk(typ: bytes, data: bytes) -> bytes:
    return struct.pack(">I", len(data)) + typ + data + struct.pack(">I", zlib.crc32(typ + data) & 0xFFFFFFFF)


def chunks(data: bytes):
    """[(type, data, stored_crc, crc_ok)] of a PNG file; raises ValueError on a malformed chunk structure."""
    if data[:8] != SIG:
        raise ValueError("not a PNG file")
    at, out = 8, []
    while at + 12 <= len(data):
        (n,) = struct.unpack(">I", data[at:at + 4])
        if at + 12 + n > len(data):
            raise ValueError("truncated chunk")
        typ, body = data[at + 4:at + 8], data[at + 8:at + 8 + n]
        (crc,) = struct.unpack(">I", data[at + 8 + n:at + 12 + n])
        out.append((typ, body, crc, (zlib.crc32(typ + body) & 0xFFFFFFFF) == crc))
        at += 12 + n
        if typ == b"IEND":
            break
    return out


def in_subset(data: bytes):
    """(W, H, colour type) if the device decoder accepts the file's structure (csrc/p2p_pngdec.cuh parse_png), else None."""
    try:
        ch = chunks(data)
    except ValueError:
        return None
    if not ch or ch[0][0] != b"IHDR" or len(ch[0][1]) != 13 or ch[-1][0] != b"IEND":
        return None
    if not all(ok for _, _, _, ok in ch):
        return None
    W, H, depth, ctype, comp, filt, lace = struct.unpack(">IIBBBBB", ch[0][1])
    if depth != 8 or ctype not in BPP or comp or filt or lace or not (0 < W < 32767 and 0 < H < 32767):
        return None
    names = [c[0] for c in ch]
    if b"IDAT" not in names:
        return None
    first = names.index(b"IDAT")
    last = len(names) - 1 - names[::-1].index(b"IDAT")
    if any(n != b"IDAT" for n in names[first:last + 1]):
        return None
    for n in names[1:-1]:
        if n == b"IDAT":
            continue
        if n in (b"acTL", b"fcTL", b"fdAT", b"tRNS", b"eXIf", b"IHDR") or not (n[0] & 0x20) or not n.isalpha():
            return None
    return W, H, ctype


def unfilter(raw: np.ndarray, H: int, row_bytes: int, bpp: int) -> np.ndarray:
    """[H, row_bytes] reconstructed bytes from the inflated image (filter type byte + filtered bytes per row)."""
    rows = raw.reshape(H, 1 + row_bytes)
    out = np.zeros((H, row_bytes), np.uint8)
    prev = np.zeros(row_bytes, np.int32)
    for y in range(H):
        ft = int(rows[y, 0])
        f = rows[y, 1:].astype(np.int32)
        if ft == 0:
            cur = f
        elif ft == 2:
            cur = (f + prev) & 255
        elif ft == 1:
            cur = f.copy()
            for k in range(bpp):  # a prefix sum per channel
                cur[k::bpp] = np.cumsum(f[k::bpp]) & 255
        elif ft in (3, 4):
            cur = np.zeros(row_bytes, np.int32)
            fl, pl = f.tolist(), prev.tolist()
            c = [0] * row_bytes
            for i in range(row_bytes):
                a = c[i - bpp] if i >= bpp else 0
                b = pl[i]
                if ft == 3:
                    pred = (a + b) >> 1
                else:
                    cc = pl[i - bpp] if i >= bpp else 0
                    p = a + b - cc
                    pa, pb, pc = abs(p - a), abs(p - b), abs(p - cc)
                    pred = a if (pa <= pb and pa <= pc) else (b if pb <= pc else cc)
                c[i] = (fl[i] + pred) & 255
            cur = np.asarray(c, np.int32)
        else:
            raise ValueError("invalid filter type")
        out[y] = cur
        prev = cur
    return out


def to_bgr(px: np.ndarray, W: int, ctype: int) -> np.ndarray:
    """What cv2.imread(path) (IMREAD_COLOR) returns for reconstructed rows of a colour type."""
    H = px.shape[0]
    p = px.reshape(H, W, BPP[ctype])
    if ctype in (0, 4):
        return np.repeat(p[:, :, :1], 3, axis=2).copy()
    return p[:, :, 2::-1].copy()


def decode(data: bytes) -> np.ndarray:
    """BGR u8 [H, W, 3] of a file inside the subset (ValueError otherwise; zlib.error for a damaged stream)."""
    sub = in_subset(data)
    if sub is None:
        raise ValueError("outside the subset")
    W, H, ctype = sub
    z = b"".join(body for typ, body, _, _ in chunks(data) if typ == b"IDAT")
    raw = np.frombuffer(zlib.decompress(z), np.uint8)
    bpp = BPP[ctype]
    if raw.size != H * (1 + W * bpp):
        raise ValueError("wrong amount of image data")
    return to_bgr(unfilter(raw, H, W * bpp, bpp), W, ctype)


# ---- writers (test files) ---------------------------------------------------------------------------------------
def filter_rows(px: np.ndarray, bpp: int, filters) -> bytes:
    """Filtered scanlines of reconstructed rows px [H, row_bytes]; filters: one type per row (0 .. 4) or 'adaptive'
    (libpng's minimum-sum-of-absolute-differences heuristic)."""
    H, rb = px.shape
    p = px.astype(np.int32)
    zero = np.zeros(rb, np.int32)
    out = bytearray()
    for y in range(H):
        cur = p[y]
        up = p[y - 1] if y else zero
        a = np.concatenate([np.zeros(bpp, np.int32), cur[:-bpp]]) if rb > bpp else np.zeros(rb, np.int32)
        c = np.concatenate([np.zeros(bpp, np.int32), up[:-bpp]]) if rb > bpp else np.zeros(rb, np.int32)
        pp = a + up - c
        pa, pb, pc = np.abs(pp - a), np.abs(pp - up), np.abs(pp - c)
        paeth = np.where((pa <= pb) & (pa <= pc), a, np.where(pb <= pc, up, c))
        cands = [cur, cur - a, cur - up, cur - ((a + up) >> 1), cur - paeth]
        if isinstance(filters, str):
            costs = [int(np.abs(((v & 255) ^ 128) - 128).sum()) for v in cands]
            ft = int(np.argmin(costs))
        else:
            ft = int(filters[y])
        out.append(ft)
        out += (cands[ft] & 255).astype(np.uint8).tobytes()
    return bytes(out)


def write_png(img: np.ndarray, ctype: int = 2, filters="adaptive", level: int = 6, strategy: int = zlib.Z_DEFAULT_STRATEGY,
              wbits: int = 15, idat: int | list = 8192, flush_every: int = 0, flush_mode: int = zlib.Z_SYNC_FLUSH,
              extra_chunks=(), mem_level: int = 8) -> bytes:
    """A PNG file of img (u8, [H, W] for gray, [H, W, C] in PNG channel order RGB / GA / RGBA otherwise).
    idat: chunk size, or a list of sizes used cyclically.  flush_every: bytes of filtered data between zlib flushes."""
    img = np.asarray(img, np.uint8)
    H, W = img.shape[:2]
    bpp = BPP[ctype]
    px = img.reshape(H, W * bpp)
    raw = filter_rows(px, bpp, filters)
    co = zlib.compressobj(level, zlib.DEFLATED, wbits, mem_level, strategy)
    if flush_every:
        z = b""
        for at in range(0, len(raw), flush_every):
            z += co.compress(raw[at:at + flush_every])
            z += co.flush(flush_mode)
        z += co.flush()
    else:
        z = co.compress(raw) + co.flush()
    out = SIG + chunk(b"IHDR", struct.pack(">IIBBBBB", W, H, 8, ctype, 0, 0, 0))
    for typ, body in extra_chunks:
        out += chunk(typ, body)
    sizes = idat if isinstance(idat, list) else [idat]
    at, k = 0, 0
    while at < len(z):
        n = sizes[k % len(sizes)]
        out += chunk(b"IDAT", z[at:at + n])
        at += n
        k += 1
    return out + chunk(b"IEND", b"")


def test_image(H: int, W: int, channels: int, seed: int, kind: str = "mixed") -> np.ndarray:
    """Synthetic image content: 'smooth' (long matches), 'noise' (literals), 'mixed' (smooth + noise + flat areas)."""
    rng = np.random.default_rng(seed)
    y, x = np.mgrid[0:H, 0:W]
    base = np.stack([(x * (3 + k) + y * (5 - k)) // 4 + 40 * k for k in range(channels)], axis=2)
    if kind == "noise":
        img = rng.integers(0, 256, (H, W, channels))
    elif kind == "smooth":
        img = base
    else:
        img = base + rng.integers(-6, 7, (H, W, channels))
        img[H // 3:H // 2, W // 4:W // 2] = 200                      # a flat area: long runs
        img[:, 3 * W // 4:] = rng.integers(0, 256, (H, W - 3 * W // 4, channels))
    img = (img & 255).astype(np.uint8)
    return img[:, :, 0] if channels == 1 else img


def fixed_huffman_png(n_lit: int, dist_code: int, dist_extra_bits: int = 0, dist_extra: int = 0, cmf: int = 0x78,
                      stored_prefix: bytes = b"") -> bytes:
    """A hand-made one-row gray PNG whose zlib stream is [an optional stored block holding stored_prefix,] then ONE
    fixed-Huffman block: n_lit literals 'A', a match of length 3 with the given distance code (+ extra bits), end of block.
    The Adler-32 is left zero: the files are for the decoders' validity checks (distances in front of the data, beyond the
    window of the zlib header), which must decline them before the checksum matters."""

    def huff(code, n):  # Huffman codes are packed starting with their most significant bit
        return [(code >> (n - 1 - k)) & 1 for k in range(n)]

    bits = []
    if stored_prefix:
        bits += [0, 0, 0] + [0] * 5   # not final, stored, padding to the byte boundary
        n = len(stored_prefix)
        for v in (n & 255, n >> 8, (n ^ 0xFFFF) & 255, (n ^ 0xFFFF) >> 8) + tuple(stored_prefix):
            bits += [(v >> k) & 1 for k in range(8)]
    bits += [1, 1, 0]             # final block, fixed Huffman
    for _ in range(n_lit):
        bits += huff(0x30 + 65, 8)  # literal 'A'
    bits += huff(0b0000001, 7)      # length code 257: length 3
    bits += huff(dist_code, 5)
    bits += [(dist_extra >> k) & 1 for k in range(dist_extra_bits)]
    bits += huff(0, 7)              # end of block
    deflate = bytearray()
    for i in range(0, len(bits), 8):
        deflate.append(sum(b << k for k, b in enumerate(bits[i:i + 8])))
    flg = (31 - (cmf * 256) % 31) % 31
    body = bytes([cmf, flg]) + bytes(deflate) + bytes(4)
    W = len(stored_prefix) + n_lit + 3 - 1
    return SIG + chunk(b"IHDR", struct.pack(">IIBBBBB", W, 1, 8, 0, 0, 0, 0)) + chunk(b"IDAT", body) + chunk(b"IEND", b"")
