"""PNG decoder (csrc/p2p_pngdec.cuh, SURVEY 8f-2: the decode side of ``cv2.imread(path)``, ref :244) on the CPU:

* the oracle ``oracle/png_decode_model.py`` pinned against ``cv2.imdecode`` itself (every colour type, every filter);
* the library's HOST MODEL of the device decoder (``p2p_png_decode_host``: the same ``__host__ __device__`` routines as the
  kernels - block-start search, chain walk, inflate with symbolic history, resolution, Adler-32 / CRC-32, unfilter - run
  serially) against ``cv2.imdecode`` on files of many writers, and against the rule for damaged files: decoded like libpng
  or declined, never differently.  No GPU needed: the compute runs on the host.
"""
import ctypes as C
import struct
import zlib

import cv2
import numpy as np
import pytest

from oracle import png_decode_model as M

COLOUR = [(0, 1), (2, 3), (4, 2), (6, 4)]


def cv2_decode(data: bytes):
    return cv2.imdecode(np.frombuffer(data, np.uint8), cv2.IMREAD_COLOR)


@pytest.fixture(scope="module")
def host_decode(pkg):
    lib = pkg._lib.load()

    def run(data: bytes):
        w, h = C.c_int(), C.c_int()
        rc = lib.p2p_png_probe(data, len(data), C.byref(w), C.byref(h))
        if rc:
            return rc, None, None
        out = np.zeros((h.value, w.value, 3), np.uint8)
        stats = (C.c_uint64 * 4)()
        rc = lib.p2p_png_decode_host(data, len(data), out.ctypes.data, out.strides[0], h.value, stats)
        return rc, (out if rc == 0 else None), list(stats)

    return run


@pytest.mark.parametrize("ctype,ch", COLOUR)
def test_oracle_equals_cv2_for_every_colour_type_and_filter(ctype, ch):
    img = M.test_image(40, 70, ch, 10 + ctype)
    rows = list(np.random.default_rng(3).integers(0, 5, 40))
    for filt in ["adaptive", [0] * 40, [1] * 40, [2] * 40, [3] * 40, [4] * 40, rows]:
        f = M.write_png(img, ctype, filters=filt)
        assert np.array_equal(M.decode(f), cv2_decode(f))


def test_oracle_subset_rules():
    img = M.test_image(8, 8, 3, 0)
    f = M.write_png(img, 2)
    assert M.in_subset(f) == (8, 8, 2)
    assert M.in_subset(f[:-1]) is None
    assert M.in_subset(M.write_png(img, 2, extra_chunks=[(b"tRNS", bytes(6))])) is None
    assert M.in_subset(M.write_png(img, 2, extra_chunks=[(b"gAMA", struct.pack(">I", 45455))])) == (8, 8, 2)
    f16 = cv2.imencode(".png", np.zeros((4, 4, 3), np.uint16))[1].tobytes()
    assert M.in_subset(f16) is None


@pytest.mark.parametrize("ctype,ch", COLOUR)
def test_host_model_colour_types_and_filters(host_decode, ctype, ch):
    img = M.test_image(90, 130, ch, ctype)
    rows = list(np.random.default_rng(5).integers(0, 5, 90))
    for filt in ["adaptive", [0] * 90, [1] * 90, [2] * 90, [3] * 90, [4] * 90, rows]:
        f = M.write_png(img, ctype, filters=filt)
        rc, out, _ = host_decode(f)
        assert rc == 0 and np.array_equal(out, cv2_decode(f))


WRITERS = [
    dict(), dict(level=1), dict(level=9), dict(level=0), dict(strategy=zlib.Z_RLE), dict(strategy=zlib.Z_FIXED),
    dict(strategy=zlib.Z_HUFFMAN_ONLY), dict(strategy=zlib.Z_FILTERED), dict(idat=1), dict(idat=[7, 1000, 3]), dict(idat=1 << 30),
    dict(wbits=9), dict(wbits=12, level=9), dict(mem_level=1), dict(mem_level=9),
    dict(flush_every=5000), dict(flush_every=7777, flush_mode=zlib.Z_FULL_FLUSH), dict(flush_every=300, level=1),
    dict(extra_chunks=[(b"gAMA", struct.pack(">I", 45455)), (b"tEXt", b"Comment\0x"), (b"pHYs", bytes(9)), (b"bKGD", bytes(6))]),
]


@pytest.mark.parametrize("kw", WRITERS, ids=[",".join(f"{k}={v if not isinstance(v, list) else 'list'}" for k, v in kw.items()) or "default" for kw in WRITERS])
def test_host_model_many_writers(host_decode, kw):
    for kind, seed in [("mixed", 1), ("smooth", 2), ("noise", 3)]:
        img = M.test_image(120, 200, 3, seed, kind)
        f = M.write_png(img, 2, **kw)
        ref = cv2_decode(f)
        assert ref is not None
        rc, out, stats = host_decode(f)
        assert rc == 0 and np.array_equal(out, ref), (kind, stats)


def test_host_model_cv2_written_files(host_decode):
    """The reference's own outputs (cv2.imwrite defaults: filter Sub, Z_RLE, level 1) and the other OpenCV settings."""
    img = M.test_image(300, 500, 3, 7)
    P = cv2
    for params in [[], [P.IMWRITE_PNG_COMPRESSION, 0], [P.IMWRITE_PNG_COMPRESSION, 9],
                   [P.IMWRITE_PNG_STRATEGY, P.IMWRITE_PNG_STRATEGY_DEFAULT, P.IMWRITE_PNG_COMPRESSION, 6],
                   [P.IMWRITE_PNG_STRATEGY, P.IMWRITE_PNG_STRATEGY_HUFFMAN_ONLY], [P.IMWRITE_PNG_STRATEGY, P.IMWRITE_PNG_STRATEGY_FIXED]]:
        f = cv2.imencode(".png", img, params)[1].tobytes()
        rc, out, stats = host_decode(f)
        assert rc == 0 and np.array_equal(out, cv2_decode(f)), (params, stats)
    for shape in [(1, 1, 3), (1, 300, 3), (300, 1, 3), (33, 65), (17, 40, 4)]:
        img = np.random.default_rng(sum(shape)).integers(0, 256, shape, dtype=np.uint8)
        f = cv2.imencode(".png", img)[1].tobytes()
        rc, out, _ = host_decode(f)
        assert rc == 0 and np.array_equal(out, cv2_decode(f)), shape


def test_host_model_many_blocks_and_search_statistics(host_decode):
    """A stream of ~80 deflate blocks: every dynamic block is found by the search (none measured by the chain walk), about
    one bit position in a thousand passes the first header test, and no false candidate survives the full test."""
    img = M.test_image(700, 1000, 3, 9)
    f = M.write_png(img, 2, level=6)
    rc, out, stats = host_decode(f)
    assert rc == 0 and np.array_equal(out, cv2_decode(f))
    n_quick, n_cand, n_blocks, n_host = stats
    assert n_blocks > 50 and n_host == 0 and n_cand == n_blocks
    assert n_quick < 8 * len(f) // 200
    # full flushes put empty stored blocks between the dynamic ones: the walk measures those itself
    f = M.write_png(img, 2, level=6, flush_every=50000, flush_mode=zlib.Z_FULL_FLUSH)
    rc, out, stats = host_decode(f)
    assert rc == 0 and np.array_equal(out, cv2_decode(f)) and stats[3] > 0


def test_unsupported_files_are_declined(host_decode):
    img8 = M.test_image(20, 30, 3, 0)
    cases = {
        "16 bit": cv2.imencode(".png", np.zeros((5, 5, 3), np.uint16))[1].tobytes(),
        "tRNS": M.write_png(img8, 2, extra_chunks=[(b"tRNS", bytes(6))]),
        "APNG": M.write_png(img8, 2, extra_chunks=[(b"acTL", struct.pack(">II", 1, 0))]),
        "eXIf": M.write_png(img8, 2, extra_chunks=[(b"eXIf", b"MM\0*\0\0\0\x08\0\0")]),
        "unknown critical chunk": M.write_png(img8, 2, extra_chunks=[(b"ABCD", b"x")]),
        "not a PNG": b"\xff\xd8\xff\xe0" + bytes(100),
        "empty": b"",
    }
    f = bytearray(M.write_png(img8, 2))
    f[25] ^= 4  # IHDR colour type 2 -> 6: CRC mismatch
    cases["IHDR damaged"] = bytes(f)
    # interlaced: IHDR flag set (CRC recomputed)
    ihdr = struct.pack(">IIBBBBB", 30, 20, 8, 2, 0, 0, 1)
    good = M.write_png(img8, 2)
    cases["interlaced"] = good[:8] + M.chunk(b"IHDR", ihdr) + good[33:]
    for name, data in cases.items():
        rc, out, _ = host_decode(data)
        assert rc == -6 and out is None, name


def _rechunk(data: bytes, z: bytes) -> bytes:
    """The file with its zlib stream replaced by z (one IDAT chunk, correct CRC)."""
    ch = M.chunks(data)
    out = M.SIG
    done = False
    for typ, body, _, _ in ch:
        if typ == b"IDAT":
            if not done:
                out += M.chunk(b"IDAT", z)
                done = True
            continue
        out += M.chunk(typ, body)
    return out


def test_damaged_files_decode_like_libpng_or_not_at_all(host_decode):
    rng = np.random.default_rng(11)
    img = M.test_image(150, 260, 3, 4)
    good = M.write_png(img, 2, level=6, idat=[4000])
    z = b"".join(body for typ, body, _, _ in M.chunks(good) if typ == b"IDAT")
    declined = decoded = 0

    def check(data, must_decline=False):
        nonlocal declined, decoded
        rc, out, _ = host_decode(data)
        if rc == 0:
            ref = cv2_decode(data)
            assert not must_decline and ref is not None and np.array_equal(out, ref)
            decoded += 1
        else:
            assert rc == -6
            declined += 1

    # (a) a byte of the file flipped, chunk CRCs left alone -> CRC mismatch somewhere
    for _ in range(40):
        d = bytearray(good)
        at = int(rng.integers(8, len(d) - 12))
        d[at] ^= 1 << int(rng.integers(0, 8))
        check(bytes(d), must_decline=True)
    # (b) the deflate data damaged, CRCs recomputed -> invalid codes / distances / lengths, or an Adler-32 mismatch
    for _ in range(60):
        zz = bytearray(z)
        at = int(rng.integers(2, len(zz)))
        zz[at] ^= 1 << int(rng.integers(0, 8))
        check(_rechunk(good, bytes(zz)), must_decline=True)
    # (c) truncated streams, trailing garbage, wrong Adler-32, wrong amount of data
    for cut in (1, 4, 5, 100, len(z) // 2):
        check(_rechunk(good, z[:-cut]), must_decline=True)
    check(_rechunk(good, z + b"\0"), must_decline=True)
    check(_rechunk(good, z[:-4] + bytes(4)), must_decline=True)
    raw = zlib.decompress(z)
    check(_rechunk(good, zlib.compress(raw + b"\0")), must_decline=True)
    check(_rechunk(good, zlib.compress(raw[:-1])), must_decline=True)
    # a filter type above 4 (valid zlib stream)
    bad = bytearray(raw)
    bad[(1 + 260 * 3) * 7] = 5
    check(_rechunk(good, zlib.compress(bytes(bad))), must_decline=True)
    # a preset dictionary flag / a wrong header checksum
    check(_rechunk(good, bytes([z[0], z[1] ^ 0x20]) + z[2:]), must_decline=True)
    # (d) missing IEND / IDAT chunks not consecutive
    check(good[:-12], must_decline=True)
    ch = M.chunks(good)
    parts = [M.chunk(t, b) for t, b, _, _ in ch]
    check(M.SIG + b"".join(parts[:2]) + M.chunk(b"tEXt", b"a\0b") + b"".join(parts[2:]), must_decline=True)
    assert declined > 100 and decoded == 0
    check(good)
    assert decoded == 1


def test_distance_beyond_the_window_or_the_data_is_declined(host_decode):
    """Hand-made fixed-Huffman streams whose match reaches in front of the first byte ("invalid distance too far back"; in
    the first block and in a second block behind a stored one), and one whose distance exceeds the window the zlib header
    declares.  The same stream with a 32 KiB window is valid except for its Adler-32: still declined, never decoded differently."""
    cases = {
        "in front of the data, first block": M.fixed_huffman_png(2, 3),                                   # 2 literals, distance 4
        "in front of the data, second block": M.fixed_huffman_png(2, 5, 1, 0, stored_prefix=b"\0abc"),   # 6 bytes so far, distance 7
        "beyond the declared window": M.fixed_huffman_png(300, 16, 7, 0, cmf=0x08),                        # window 256, distance 257
        "wrong Adler-32 only": M.fixed_huffman_png(300, 16, 7, 0),
    }
    for name, f in cases.items():
        assert M.in_subset(f) is not None, name
        assert cv2_decode(f) is None or name == "wrong Adler-32 only", name
        rc, out, _ = host_decode(f)
        assert rc == -6 and out is None, name


def test_host_model_under_address_sanitizer(tmp_path):
    """The decoder's routines (chunk walk, bit reader, header parser, table builder, block decoder with symbolic history,
    tail / resolution passes, filters), built with ASan + UBSan (tools/asan_png_host.cu), on seeded damaged files incl.
    garbage runs and truncations: no out-of-bounds access, no undefined behaviour."""
    import shutil
    import subprocess
    from pathlib import Path

    root = Path(__file__).resolve().parents[1]
    if shutil.which("nvcc") is None:
        pytest.skip("nvcc not on PATH")
    exe = tmp_path / "asan_png_host"
    build = subprocess.run(["nvcc", "-O1", "-g", "-Xcompiler", "-fsanitize=address,-fsanitize=undefined,-fno-omit-frame-pointer",
                            "-gencode", "arch=compute_100a,code=sm_100a", "-o", str(exe), str(root / "tools" / "asan_png_host.cu"),
                            "-lasan", "-lubsan"], capture_output=True, text=True)
    if build.returncode != 0:
        pytest.skip("sanitizer runtime not available: " + build.stderr[-300:])
    rng = np.random.default_rng(19)
    bases = [M.write_png(M.test_image(60, 90, ch, ct), ct, level=lv, idat=[700], flush_every=fl)
             for (ct, ch), lv, fl in zip(COLOUR, (6, 1, 9, 0), (0, 900, 0, 0))]
    bases.append(M.fixed_huffman_png(2, 5, 1, 0, stored_prefix=b"\0abc"))
    names = []
    for k in range(400):
        base = bases[k % len(bases)]
        d = bytearray(base)
        if k >= len(bases):
            if k % 2:   # inside the deflate data, chunk CRC recomputed: only the inflate's own checks stand in the way
                ch = M.chunks(base)
                z = bytearray(b"".join(b for t, b, _, _ in ch if t == b"IDAT"))
                for _ in range(int(rng.integers(1, 4))):
                    a, n = int(rng.integers(2, len(z))), int(rng.integers(1, 12))
                    z[a:a + n] = bytes(rng.integers(0, 256, n, dtype=np.uint8))
                if rng.integers(0, 4) == 0:
                    z = z[:int(rng.integers(2, len(z)))]
                d = bytearray(M.SIG + M.chunk(b"IHDR", ch[0][1]) + M.chunk(b"IDAT", bytes(z)) + M.chunk(b"IEND", b""))
            else:
                for _ in range(int(rng.integers(1, 4))):
                    a, n = int(rng.integers(8, len(d))), int(rng.integers(1, 12))
                    d[a:a + n] = bytes(rng.integers(0, 256, n, dtype=np.uint8))
                if rng.integers(0, 4) == 0:
                    d = d[:int(rng.integers(8, len(d)))]
        f = tmp_path / f"{k:04d}.png"
        f.write_bytes(bytes(d))
        names.append(str(f))
    run = subprocess.run([str(exe)] + names, capture_output=True, text=True, env={"ASAN_OPTIONS": "detect_leaks=0"})
    assert run.returncode == 0 and "runtime error" not in run.stderr and "ERROR" not in run.stderr, run.stderr[-2000:]
    assert run.stdout.startswith("decoded ")
