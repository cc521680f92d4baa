"""GPU PNG decoder (csrc/p2p_pngdec.cuh) through the C ABI: the pixels must equal what the reference's
``cv2.imread(path)`` (ref :244) returns for a ``.png`` panorama - checked against the real ``cv2.imdecode`` on files of many
writers - and a file libpng would refuse must be declined (the front end then calls cv2.imread as before)."""
import struct
import zlib

import cv2
import numpy as np
import pytest

from oracle import png_decode_model as M
from oracle import synth

pytestmark = pytest.mark.gpu


def cv2_decode(data: bytes):
    return cv2.imdecode(np.frombuffer(data, np.uint8), cv2.IMREAD_COLOR)


@pytest.mark.parametrize("ctype,ch", [(0, 1), (2, 3), (4, 2), (6, 4)])
def test_colour_types_and_filters(proj, ctype, ch):
    """Every colour type of the subset with every filter: single-filter files (all rows Up / Average / Paeth form ONE run of
    90 rows = three bands of the wavefront), random filters per row, libpng's adaptive choice."""
    img = M.test_image(90, 130, ch, ctype)
    rows = list(np.random.default_rng(5).integers(0, 5, 90))
    for filt in ["adaptive", [0] * 90, [1] * 90, [2] * 90, [3] * 90, [4] * 90, rows]:
        f = M.write_png(img, ctype, filters=filt)
        assert np.array_equal(proj.decode_png(f), cv2_decode(f)), (ctype, filt if isinstance(filt, str) else filt[:4])


WRITERS = [
    dict(), dict(level=1), dict(level=9), dict(level=0), dict(strategy=zlib.Z_RLE), dict(strategy=zlib.Z_FIXED),
    dict(strategy=zlib.Z_HUFFMAN_ONLY), dict(idat=1), dict(idat=[7, 1000, 3]), dict(idat=1 << 30), dict(wbits=9), dict(mem_level=1),
    dict(flush_every=5000), dict(flush_every=7777, flush_mode=zlib.Z_FULL_FLUSH), dict(flush_every=300, level=1),
    dict(extra_chunks=[(b"gAMA", struct.pack(">I", 45455)), (b"tEXt", b"Comment\0x"), (b"bKGD", bytes(6))]),
]


@pytest.mark.parametrize("kw", WRITERS, ids=[",".join(f"{k}={v if not isinstance(v, list) else 'list'}" for k, v in kw.items()) or "default" for kw in WRITERS])
def test_many_writers(proj, kw):
    for kind, seed in [("mixed", 1), ("smooth", 2), ("noise", 3)]:
        f = M.write_png(M.test_image(120, 200, 3, seed, kind), 2, **kw)
        assert np.array_equal(proj.decode_png(f), cv2_decode(f)), kind


def test_cv2_written_files_and_odd_shapes(proj):
    """The reference's own outputs (cv2.imwrite defaults: filter Sub on every row = one run per row, Z_RLE) and OpenCV's
    other settings; one-pixel-wide / one-row images."""
    img = M.test_image(300, 500, 3, 7)
    P = cv2
    for params in [[], [P.IMWRITE_PNG_COMPRESSION, 0], [P.IMWRITE_PNG_COMPRESSION, 9],
                   [P.IMWRITE_PNG_STRATEGY, P.IMWRITE_PNG_STRATEGY_DEFAULT, P.IMWRITE_PNG_COMPRESSION, 6],
                   [P.IMWRITE_PNG_STRATEGY, P.IMWRITE_PNG_STRATEGY_HUFFMAN_ONLY], [P.IMWRITE_PNG_STRATEGY, P.IMWRITE_PNG_STRATEGY_FIXED]]:
        f = cv2.imencode(".png", img, params)[1].tobytes()
        assert np.array_equal(proj.decode_png(f), cv2_decode(f)), params
    for shape in [(1, 1, 3), (1, 300, 3), (300, 1, 3), (33, 65), (17, 40, 4), (2, 40000 // 3, 3)]:
        im = np.random.default_rng(sum(shape)).integers(0, 256, shape, dtype=np.uint8)
        f = cv2.imencode(".png", im)[1].tobytes()
        assert np.array_equal(proj.decode_png(f), cv2_decode(f)), shape


def test_history_chains_across_many_blocks(proj):
    """Streams whose blocks consist of back-references into earlier blocks: a constant image cut into blocks by sync flushes
    (every block is one long chain into the previous one), a horizontally periodic image, and an ordinary image of ~80 blocks."""
    flat = np.full((400, 500, 3), 77, np.uint8)
    for kw in [dict(filters=[0] * 400, flush_every=20000), dict(filters=[0] * 400, flush_every=3000, level=1), dict(filters=[2] * 400, flush_every=40000)]:
        f = M.write_png(flat, 2, **kw)
        assert np.array_equal(proj.decode_png(f), cv2_decode(f)), kw
    tile = np.random.default_rng(2).integers(0, 256, (600, 64, 3), dtype=np.uint8)
    f = M.write_png(np.tile(tile, (1, 16, 1)), 2, filters=[0] * 600, level=6)
    assert np.array_equal(proj.decode_png(f), cv2_decode(f))
    f = M.write_png(np.tile(tile, (1, 16, 1)), 2, filters=[0] * 600, level=6, flush_every=10000)
    assert np.array_equal(proj.decode_png(f), cv2_decode(f))
    f = M.write_png(M.test_image(700, 1000, 3, 9), 2, level=6)
    assert np.array_equal(proj.decode_png(f), cv2_decode(f))


def test_panorama_sized_files_and_upload(pkg, proj, tmp_path):
    """2048 x 1024 panoramas written by OpenCV (defaults) and with adaptive filters at zlib level 6: decode_png, upload_png +
    the packed panorama, and the views against the cv2.imread path."""
    Wp, Hp, W, H, fov = 2048, 1024, 240, 136, 120
    pano = np.clip(synth.smooth(Wp, Hp, 7).astype(int) + np.random.default_rng(1).integers(-9, 10, (Hp, Wp, 3)), 0, 255).astype(np.uint8)
    files = {"cv2": cv2.imencode(".png", pano)[1].tobytes(), "adaptive": M.write_png(pano[:, :, ::-1], 2, level=6)}
    shifts = [pkg.yaw_table(Wp, y)[2] for y in (0, 90, 180, 270)]
    consts = [pkg.pitch_constants(W, fov, p) for p in (30, 60, 90)]
    for name, data in files.items():
        assert proj.png_probe(data) == (Wp, Hp)
        assert np.array_equal(cv2_decode(data), pano)
        assert np.array_equal(proj.decode_png(data), pano), name
        with proj.slots(1) as (s,):
            assert proj.upload_png(s, data) == (Wp, Hp)
            got = proj.project(s, shifts, consts, W, H)
            proj.sync(s)
            assert np.array_equal(proj.download_pano(s, Wp, Hp), pano), name
            proj.upload(s, pano)
            want = proj.project(s, shifts, consts, W, H)
            proj.sync(s)
        assert np.array_equal(got, want), name
    with pytest.raises(pkg.P2PError) as ei:   # outside the decoder's subset: the caller falls back to cv2.imread
        proj.decode_png(cv2.imencode(".png", pano[:64, :64].astype(np.uint16) * 257)[1].tobytes())
    assert ei.value.code == -6


def _rechunk(data: bytes, z: bytes) -> bytes:
    out, done = M.SIG, False
    for typ, body, _, _ in M.chunks(data):
        if typ == b"IDAT":
            if not done:
                out += M.chunk(b"IDAT", z)
                done = True
            continue
        out += M.chunk(typ, body)
    return out


def test_damaged_files_are_declined(pkg, proj):
    """CRC-32 mismatches (found by the device CRC), damaged deflate data with correct CRCs (invalid codes / distances / a
    wrong Adler-32, found by the device inflate), truncation, a bad filter type: always -6, and the next good file decodes."""
    rng = np.random.default_rng(11)
    img = M.test_image(150, 260, 3, 4)
    good = M.write_png(img, 2, level=6, idat=[4000])
    z = b"".join(body for typ, body, _, _ in M.chunks(good) if typ == b"IDAT")
    bad = []
    for _ in range(12):
        d = bytearray(good)
        d[int(rng.integers(60, len(d) - 12))] ^= 1 << int(rng.integers(0, 8))
        bad.append(bytes(d))
    for _ in range(24):
        zz = bytearray(z)
        zz[int(rng.integers(2, len(zz)))] ^= 1 << int(rng.integers(0, 8))
        bad.append(_rechunk(good, bytes(zz)))
    bad += [_rechunk(good, z[:-5]), _rechunk(good, z[:len(z) // 2]), _rechunk(good, z[:-4] + bytes(4)), _rechunk(good, z + b"\0")]
    raw = bytearray(zlib.decompress(z))
    raw[(1 + 260 * 3) * 7] = 5
    bad.append(_rechunk(good, zlib.compress(bytes(raw))))
    # matches that reach in front of the data (the decoding pass gives up in the middle of a block: what is left of the
    # block's slice of the match list is stale and must not be executed), a distance beyond the declared window
    bad += [M.fixed_huffman_png(2, 3), M.fixed_huffman_png(2, 5, 1, 0, stored_prefix=b"\0abc"), M.fixed_huffman_png(300, 16, 7, 0, cmf=0x08),
            M.fixed_huffman_png(300, 16, 7, 0)]
    for i, d in enumerate(bad):
        with pytest.raises(pkg.P2PError) as ei:
            proj.decode_png(d)
        assert ei.value.code == -6, i
        if i % 8 == 0:   # a match-heavy good file in between leaves a full match list behind for the next damaged one
            assert np.array_equal(proj.decode_png(good), cv2_decode(good))
    assert np.array_equal(proj.decode_png(good), cv2_decode(good))


def test_front_end_png_input_equals_imread_path(pkg, tmp_path):
    """.png panoramas through every front door (directory pipeline to png and jpg files, ``panorama_to_plane``, a fractional
    yaw): same files / pixels as the cv2.imread path.  A 16-bit file and a damaged file take the cv2 fallback, a file cv2
    cannot read either is skipped."""
    src = tmp_path / "in"
    src.mkdir()
    pano = synth.smooth(1024, 512, 21)
    cv2.imwrite(str(src / "a.png"), pano)
    (src / "b.png").write_bytes(M.write_png(pano[::-1, :, ::-1], 2, level=9))
    (src / "c.png").write_bytes(M.write_png(np.dstack([pano[:, ::-1, ::-1], pano[:, :, :1]]), 6, level=6))   # RGBA
    cv2.imwrite(str(src / "d.png"), pano.astype(np.uint16) * 257)                                            # 16 bit: cv2 path
    good = (src / "b.png").read_bytes()
    dmg = bytearray(good)
    dmg[len(dmg) // 2] ^= 0x10
    (src / "e.png").write_bytes(bytes(dmg))                                                                  # CRC mismatch
    W, H, fov, yaws, pitches = 200, 120, 100, [0, 90, 30], [60, 120]    # yaw 30 is fractional on Wp = 1024
    out_png, out_jpg = tmp_path / "png", tmp_path / "jpg"
    pkg.main(str(src), str(out_png), yaws, pitches, W, H, num_workers=3, output_format="png", fov_deg=fov)
    pkg.main(str(src), str(out_jpg), yaws, pitches, W, H, num_workers=3, output_format="jpg", fov_deg=fov)
    for stem in "abcde":
        img = cv2.imread(str(src / f"{stem}.png"))
        for y in yaws:
            for p in pitches:
                fp, fj = out_png / f"{stem}_{W}x{H}_yaw_{y}_pitch_{p}.png", out_jpg / f"{stem}_{W}x{H}_yaw_{y}_pitch_{p}.jpg"
                if img is None:
                    assert not fp.exists() and not fj.exists(), stem
                    continue
                view = pkg.process_yaw_and_pitchs(img, y, [p], W, H, fov)[0]
                assert np.array_equal(cv2.imread(str(fp)), view), (stem, y, p)
                assert fj.read_bytes() == cv2.imencode(".jpg", view)[1].tobytes(), (stem, y, p)
    one = pkg.panorama_to_plane(src / "a.png", fov, (W, H), 90, 60)
    assert np.array_equal(one, pkg.process_yaw_and_pitchs(cv2.imread(str(src / "a.png")), 90, [60], W, H, fov)[0])


def test_eight_threads_decode_png_files_at_once(pkg, proj):
    """One context, eight host threads, each decoding its own files on its own slot (the directory pipeline's pattern)."""
    from concurrent.futures import ThreadPoolExecutor

    files = [M.write_png(M.test_image(100 + 7 * k, 180 + 5 * k, 3, k), 2, level=1 + k % 9, filters="adaptive" if k % 2 else [4] * (100 + 7 * k))
             for k in range(8)]
    refs = [cv2_decode(f) for f in files]
    p8 = pkg.Projector(0, n_slots=8)
    try:
        def work(k):
            for _ in range(3):
                if not np.array_equal(p8.decode_png(files[k]), refs[k]):
                    return False
            return True

        with ThreadPoolExecutor(8) as ex:
            assert all(ex.map(work, range(8)))
    finally:
        p8.close()
