"""CPU tests of the host side: scalars, yaw tables, C helpers, CLI surface, ABI exports."""
import ctypes as C
import logging
import re
from pathlib import Path

import numpy as np
import pytest

from oracle import fixedpoint as fp
from oracle import ref_port

ROOT = Path(__file__).resolve().parent.parent


def test_library_exports_every_declared_symbol(pkg):
    header = (ROOT / "include" / "p2p.h").read_text()
    declared = set(re.findall(r"\b(p2p_[a-z0-9_]+)\s*\(", header))
    declared -= {"p2p_ctx"}
    assert len(declared) >= 25
    lib = pkg._lib.load()
    for name in sorted(declared):
        assert hasattr(lib, name), f"{name} declared in include/p2p.h but not exported"
    assert declared == set(pkg._lib.SIGNATURES), "ctypes table and header disagree"
    assert lib.p2p_abi_version() == 1
    assert lib.p2p_status_string(0) == b"ok"


def test_create_fails_loudly_without_device(pkg):
    lib = pkg._lib.load()
    if lib.p2p_device_count() > 0:
        pytest.skip("a CUDA device is present")
    ctx = C.c_void_p()
    assert lib.p2p_create(0, 2, C.byref(ctx)) == -2
    assert not ctx.value
    with pytest.raises(pkg.P2PError):
        pkg.Projector(0)
    with pytest.raises(pkg.P2PError):  # the drop-in entry point has no CPU path either
        pkg.process_yaw_and_pitchs(np.zeros((8, 16, 3), np.uint8), 0, [90], 4, 4, 90)


def test_pitch_constants_match_reference_expressions(pkg):
    lib = pkg._lib.load()
    mism = 0
    for W in (640, 1920, 800, 3840, 2048, 333):
        for fov in (30, 60, 90, 100, 120, 150, 179):
            for pitch in range(0, 181):
                want = ref_port.pitch_scalars(W, fov, pitch)
                got = pkg.pitch_constants(W, fov, pitch)
                assert tuple(np.float32(x) for x in got) == tuple(want)
                pc = pkg.PitchConsts()
                assert lib.p2p_pitch_constants(float(fov), float(pitch), W, C.byref(pc)) == 0
                mism += (np.float32(pc.f), np.float32(pc.c), np.float32(pc.s)) != tuple(want)
    # the libm export is for non-Python callers; it must agree with NumPy on this grid
    assert mism == 0


def test_yaw_table_host_and_c_helper_match_oracle(pkg):
    lib = pkg._lib.load()
    for Wp in (1000, 1024, 2048, 4096, 8192, 16384, 360, 1023):
        for yaw in (0, 90, 180, 270, 360, 30, 1, 359, 77, 45, -90, 450, 123):
            ix_o, fx_o = fp.yaw_column_table(Wp, yaw)
            ix, fx, shift = pkg.yaw_table(Wp, yaw)
            assert np.array_equal(ix, ix_o) and np.array_equal(fx, fx_o)
            assert shift == fp.yaw_table_is_roll(ix_o, fx_o)
            cix = np.empty(Wp, np.int32)
            cfx = np.empty(Wp, np.int32)
            cs = C.c_int32(-7)
            rc = lib.p2p_yaw_table(Wp, float(yaw), cix.ctypes.data_as(C.POINTER(C.c_int32)),
                                   cfx.ctypes.data_as(C.POINTER(C.c_int32)), C.byref(cs))
            assert rc == 0
            assert np.array_equal(cix, ix_o) and np.array_equal(cfx, fx_o), (Wp, yaw)
            assert cs.value == (-1 if shift is None else shift)


def test_baseline_yaws_are_integer_rolls(pkg):
    for Wp in (2048, 8192, 16384):
        for yaw, frac in ((0, 0), (90, 1), (180, 2), (270, 3)):
            assert pkg.yaw_table(Wp, yaw)[2] == frac * Wp // 4


def test_cli_flags_and_defaults_match_reference(pkg):
    ap = pkg.panorama_to_plane_pitch.build_parser()
    a = ap.parse_args(["--input_path", "x"])
    assert (a.output_path, a.output_format, a.FOV, a.output_width, a.output_height) == ("output_images", "png", 90, 800, 800)
    assert a.pitch_angles == [30, 60, 90, 120, 150] and a.yaw_angles == [0, 90, 180, 270]
    assert a.num_workers is None and a.enable_file_logging is False
    a = ap.parse_args("--input_path p --FOV 120 --output_width 1920 --output_height 1080 "
                      "--pitch_angles 30 60 90 --yaw_angles 0 90 180 270 --output_format jpg --num_workers 3".split())
    assert (a.FOV, a.output_width, a.output_height, a.output_format, a.num_workers) == (120, 1920, 1080, "jpg", 3)
    with pytest.raises(SystemExit):
        ap.parse_args(["--input_path", "x", "--pitch_angles", "0"])
    with pytest.raises(SystemExit):
        ap.parse_args(["--input_path", "x", "--pitch_angles", "180"])
    with pytest.raises(SystemExit):
        ap.parse_args(["--input_path", "x", "--output_format", "bmp"])
    assert pkg.check_pitch("1") == 1 and pkg.check_pitch("179") == 179
    import argparse

    with pytest.raises(argparse.ArgumentTypeError):
        pkg.check_pitch("abc")
    assert pkg.get_version() == "0.3.2"


def test_unreadable_image_is_logged_and_skipped(pkg, tmp_path, caplog):
    bad = tmp_path / "not_an_image.png"
    bad.write_bytes(b"nope")
    with caplog.at_level(logging.ERROR):
        assert pkg.process_single_image(bad, tmp_path, [0], [90], 8, 8) is None
    assert any("Failed to read image" in r.message for r in caplog.records)
    with pytest.raises(FileNotFoundError):
        pkg.panorama_to_plane(bad, 90, (8, 8), 0, 90)


def test_empty_directory_warns(pkg, tmp_path, caplog):
    with caplog.at_level(logging.WARNING):
        pkg.main(str(tmp_path), str(tmp_path / "out"), [0], [90], 8, 8, num_workers=1)
    assert any("No images found" in r.message for r in caplog.records)
    assert (tmp_path / "out").is_dir()


def test_image_argument_validation(pkg):
    from p2p_b200 import engine

    with pytest.raises(ValueError):
        engine._as_u8_image(np.zeros((4, 4), np.uint8))
    with pytest.raises(ValueError):
        engine._as_u8_image(np.zeros((4, 4, 3), np.float32))
    with pytest.raises(ValueError):
        engine._as_u8_image(np.zeros((4, 4, 4), np.uint8))
    a = np.zeros((4, 8, 3), np.uint8)[:, ::2]  # non-unit pixel stride -> copied
    assert engine._as_u8_image(a).flags.c_contiguous
    b = np.zeros((4, 8, 3), np.uint8)[:, :5]   # row-strided view is accepted as is
    assert engine._as_u8_image(b) is b or engine._as_u8_image(b).base is b.base


def test_sharding(pkg):
    from p2p_b200 import shard

    for n, world in ((256, 8), (5, 2), (3, 4), (0, 2)):
        seen = sorted(i for r in range(world) for i in shard.shard_images(n, r, world))
        assert seen == list(range(n))
    for (ny, npitch, world) in ((4, 3, 1), (4, 3, 2), (4, 3, 4), (4, 3, 8), (1, 1, 2)):
        allv = [v for r in range(world) for v in shard.shard_views(ny, npitch, r, world)]
        assert sorted(allv) == sorted((k, j) for k in range(ny) for j in range(npitch))
        sizes = [len(shard.shard_views(ny, npitch, r, world)) for r in range(world)]
        assert max(sizes) - min(sizes) <= 1
    assert shard.group_by_pitch([(0, 1), (1, 1), (0, 2)]) == {1: [0, 1], 2: [0]}
    # row bands of one image split over GPUs: disjoint, in order, cover [0, H), tile-aligned except the last edge
    for H in (1, 7, 8, 9, 270, 1080, 2160):
        for world in (1, 2, 3, 4, 8, 16):
            bands = [shard.shard_rows(H, r, world) for r in range(world)]
            assert bands[0][0] == 0 and bands[-1][1] == H
            for (a0, a1), (b0, b1) in zip(bands, bands[1:]):
                assert a1 == b0 and a0 <= a1
            assert all(lo % 8 == 0 for lo, _ in bands if lo < H)
            sizes = [hi - lo for lo, hi in bands]
            assert max(sizes) - min(sizes) <= 8 + 7
    with pytest.raises(ValueError):
        shard.shard_rows(8, 2, 2)
    with pytest.raises(ValueError):
        shard.shard_images(4, 2, 2)


def test_header_is_plain_c_and_links(pkg, tmp_path):
    """include/p2p.h must be consumable from C (C99, -pedantic) and the library must link from C."""
    import shutil
    import subprocess

    gcc = shutil.which("gcc")
    if gcc is None:
        pytest.skip("gcc not available")
    src = ROOT / "tests" / "c_abi_smoke.c"
    exe = tmp_path / "c_abi_smoke"
    libdir = pkg._lib.LIB_PATH.parent
    cmd = [gcc, "-std=c99", "-Wall", "-Wextra", "-pedantic", "-Werror", f"-I{ROOT / 'include'}", str(src), "-o", str(exe),
           f"-L{libdir}", "-lp2p_b200", f"-Wl,-rpath,{libdir}"]
    res = subprocess.run(cmd, capture_output=True, text=True)
    assert res.returncode == 0, res.stderr
    if pkg._lib.load().p2p_device_count() == 0:
        # no GPU here: the program must fail loudly at p2p_create, not crash
        run = subprocess.run([str(exe), "64", "32", "16", "8", "90", str(tmp_path / "o.bin")], capture_output=True, text=True)
        assert run.returncode == 2 and "p2p_create" in run.stderr


def test_slot_lease_keeps_the_slot_until_the_files_are_written(pkg):
    """``_slot``: without a lease the slot goes back when the ``with`` block ends; with a lease (an ExitStack) it stays taken
    until the caller closes the stack - also when the work in between raises - so files can be written straight from the
    slot's page-locked buffer.  (Pure host logic, checked with a stand-in for the projector's slot pool.)"""
    import contextlib

    front = pkg.panorama_to_plane_pitch

    class Pool:
        def __init__(self):
            self.free, self.log = [2, 1, 0], []

        @contextlib.contextmanager
        def slots(self, n):
            got = tuple(self.free.pop() for _ in range(n))
            self.log.append(("take", got))
            try:
                yield got
            finally:
                self.free.extend(reversed(got))
                self.log.append(("give", got))

    pool = Pool()
    with front._slot(pool, None) as (s,):
        assert s == 0 and pool.free == [2, 1]
    assert pool.free == [2, 1, 0]
    with contextlib.ExitStack() as lease:
        with front._slot(pool, lease) as (s,):
            assert s == 0
        assert pool.free == [2, 1], "the leased slot must outlive the with block"
        with front._slot(pool, lease) as (s2,):        # a second image under the same lease takes another slot
            assert s2 == 1
    assert pool.free == [2, 1, 0] or sorted(pool.free) == [0, 1, 2]
    try:
        with contextlib.ExitStack() as lease:
            with front._slot(pool, lease) as (s,):
                raise RuntimeError("projection failed")
    except RuntimeError:
        pass
    assert sorted(pool.free) == [0, 1, 2], "an exception must not leak the slot"


def test_host_assumptions_report(pkg):
    """The product's own host check: agrees with the oracle's SVML model on this host and reports versions."""
    from oracle import svml_model

    rep = pkg.host_assumptions()
    assert rep["numpy_svml"] == svml_model.host_numpy_uses_svml()
    assert rep["numpy_nep50"] is True            # NumPy >= 2 in this image
    assert set(rep["versions"]) == {"numpy", "cv2"} and set(rep["versions_match_pinned"]) == {"numpy", "cv2"}
    assert pkg.warn_if_host_differs() == rep


def test_input_files_are_routed_by_content(pkg, tmp_path):
    """``_open_image`` (the front end's ``cv2.imread``, ref :244) without a GPU: a JPEG / PNG file inside its device decoder's
    subset stays as file bytes with the probed size, whatever its suffix says (cv2.imread looks at the signature too); files
    outside the subsets come back as the array cv2.imread returns; junk is None."""
    import cv2
    import numpy as np

    from oracle import png_decode_model as M
    from p2p_b200 import engine
    from p2p_b200 import panorama_to_plane_pitch as front

    img = M.test_image(40, 64, 3, 1)
    files = {
        "a.png": cv2.imencode(".png", img)[1].tobytes(),
        "b.jpg": cv2.imencode(".jpg", img)[1].tobytes(),
        "c_is_a_png.jpg": M.write_png(img[:, :, ::-1], 2, level=9),
        "d_16bit.png": cv2.imencode(".png", img.astype(np.uint16) * 257)[1].tobytes(),
        "e_rgba.png": M.write_png(np.dstack([img[:, :, ::-1], img[:, :, :1]]), 6),
        "f_junk.png": b"\x89PNG\r\n\x1a\n" + bytes(40),
    }
    for name, data in files.items():
        (tmp_path / name).write_bytes(data)
    for name in ("a.png", "b.jpg", "c_is_a_png.jpg", "e_rgba.png"):
        src = front._open_image(tmp_path / name)
        assert isinstance(src, front._JpegSource) and (src.Wp, src.Hp) == (64, 40) and src.data == files[name], name
        assert engine.probe_encoded(files[name]) == (64, 40)
    assert engine.png_probe(files["a.png"]) == (64, 40) and engine.png_probe(files["b.jpg"]) is None
    assert engine.jpeg_probe(files["b.jpg"]) == (64, 40) and engine.jpeg_probe(files["a.png"]) is None
    got = front._open_image(tmp_path / "d_16bit.png")            # outside the PNG decoder's subset: cv2's array
    assert isinstance(got, np.ndarray) and np.array_equal(got, cv2.imread(str(tmp_path / "d_16bit.png")))
    assert front._open_image(tmp_path / "f_junk.png") is None and engine.probe_encoded(files["f_junk.png"]) is None
