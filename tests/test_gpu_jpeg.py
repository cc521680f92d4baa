"""GPU JPEG encoder (csrc/p2p_jpeg.cuh) through the C ABI: the files must be byte-identical to what the reference's
``cv2.imwrite(<name>.jpg, view)`` (ref :277) writes - checked against the real ``cv2.imencode`` and against the oracle."""
import cv2
import numpy as np
import pytest

from oracle import jpeg_model as jm
from oracle import synth

pytestmark = pytest.mark.gpu


def ref_jpeg(img, quality=None):
    params = [] if quality is None else [cv2.IMWRITE_JPEG_QUALITY, quality]
    return cv2.imencode(".jpg", img, params)[1].tobytes()


@pytest.mark.parametrize("w,h", [(1, 1), (8, 8), (16, 16), (17, 33), (53, 37), (56, 40), (240, 136), (640, 480), (1000, 333)])
@pytest.mark.parametrize("kind", ["noise", "smooth"])
def test_encode_equals_cv2(proj, w, h, kind):
    rng = np.random.default_rng(w * 131 + h)
    if kind == "noise":
        img = rng.integers(0, 256, (h, w, 3), dtype=np.uint8)
    else:
        yy, xx = np.mgrid[0:h, 0:w]
        img = np.stack([(127 + 120 * np.sin(xx / 9.0 + c) * np.cos(yy / 6.0)).astype(np.uint8) for c in range(3)], -1)
    got = proj.encode_jpeg(img)[0]
    assert got == ref_jpeg(img)
    if w * h <= 240 * 136:
        assert got == jm.encode(img)


def test_encode_batch_and_extremes(proj):
    rng = np.random.default_rng(5)
    imgs = np.stack([rng.integers(0, 256, (72, 104, 3), dtype=np.uint8), np.full((72, 104, 3), 255, np.uint8),
                     np.zeros((72, 104, 3), np.uint8), synth.smooth(104, 72, 4),
                     (rng.integers(0, 2, (72, 104, 3)) * 255).astype(np.uint8)])   # saturated noise: long codes, many 0xFF
    files = proj.encode_jpeg(imgs)
    for f, img in zip(files, imgs):
        assert f == ref_jpeg(img)


@pytest.mark.parametrize("quality", [1, 50, 75, 100])
def test_encode_quality(proj, quality):
    img = synth.smooth(200, 120, 6)
    assert proj.encode_jpeg(img, quality=quality)[0] == ref_jpeg(img, quality)


def test_full_size_view_equals_cv2(proj):
    img = synth.smooth(1920, 1080, 3)
    assert proj.encode_jpeg(img)[0] == ref_jpeg(img)
    img = synth.noise(1920, 1080, 3)
    assert proj.encode_jpeg(img)[0] == ref_jpeg(img)


def test_project_views_jpeg_equals_imwrite_of_the_views(pkg, proj, tmp_path):
    Wp, Hp, W, H, fov = 1024, 512, 240, 136, 120
    yaws, pitches = [0, 90, 180, 270], [30, 60, 90]
    pano = synth.smooth(Wp, Hp, 0)
    shifts = [pkg.yaw_table(Wp, y)[2] for y in yaws]
    consts = [pkg.pitch_constants(W, fov, p) for p in pitches]
    with proj.slots(1) as (s,):
        proj.upload(s, pano)
        views = proj.project(s, shifts, consts, W, H)
        proj.sync(s)
        files = proj.project_jpeg(s, shifts, consts, W, H)
    assert len(files) == len(yaws) * len(pitches)
    for k in range(len(yaws)):
        for j in range(len(pitches)):
            path = tmp_path / f"v_{k}_{j}.jpg"
            cv2.imwrite(str(path), views[k, j])          # what the reference does with the view (ref :277)
            assert files[k * len(pitches) + j] == path.read_bytes()


def test_front_end_jpg_files_equal_imwrite_of_the_views(pkg, tmp_path):
    """``--output_format jpg``: single-image path, directory pipeline and CLI write the bytes ``cv2.imwrite`` would write
    for the very views the png path produces (integer-roll and fractional yaws)."""
    src = tmp_path / "in"
    src.mkdir()
    sizes = [(1024, 512), (1000, 500), (1024, 512)]
    panos = {}
    for i, (Wp, Hp) in enumerate(sizes):
        panos[f"p{i}"] = synth.smooth(Wp, Hp, 10 + i)
        assert cv2.imwrite(str(src / f"p{i}.png"), panos[f"p{i}"])
    W, H, fov, yaws, pitches = 200, 120, 100, [0, 90, 30], [60, 120]    # yaw 30 is fractional on Wp = 1000 / 1024
    out_dir, out_single = tmp_path / "dir", tmp_path / "single"
    pkg.main(str(src), str(out_dir), yaws, pitches, W, H, num_workers=3, output_format="jpg", fov_deg=fov)
    out_single.mkdir()
    pkg.process_single_image(src / "p1.png", out_single, yaws, pitches, W, H, num_workers=2, output_format="jpeg", fov_deg=fov)
    names = sorted(p.name for p in out_dir.iterdir())
    assert len(names) == len(sizes) * len(yaws) * len(pitches)
    for base, pano in panos.items():
        for y in yaws:
            views = pkg.process_yaw_and_pitchs(pano, y, pitches, W, H, fov)
            for p, view in zip(pitches, views):
                want = cv2.imencode(".jpg", view)[1].tobytes()
                assert (out_dir / f"{base}_{W}x{H}_yaw_{y}_pitch_{p}.jpg").read_bytes() == want, (base, y, p)
                if base == "p1":
                    assert (out_single / f"{base}_{W}x{H}_yaw_{y}_pitch_{p}.jpeg").read_bytes() == want
    out_cli = tmp_path / "cli"
    pkg.cli(["--input_path", str(src / "p0.png"), "--output_path", str(out_cli), "--output_format", "jpg", "--FOV", str(fov),
             "--output_width", str(W), "--output_height", str(H), "--yaw_angles", "90", "--pitch_angles", "60"])
    assert (out_cli / f"p0_{W}x{H}_yaw_90_pitch_60.jpg").read_bytes() == (out_dir / f"p0_{W}x{H}_yaw_90_pitch_60.jpg").read_bytes()


# ---- decode side ----------------------------------------------------------------------------------------------------
SAMPLING = {"420": cv2.IMWRITE_JPEG_SAMPLING_FACTOR_420, "422": cv2.IMWRITE_JPEG_SAMPLING_FACTOR_422,
            "444": cv2.IMWRITE_JPEG_SAMPLING_FACTOR_444}


@pytest.mark.parametrize("w,h", [(1, 1), (2, 3), (3, 4), (17, 33), (53, 37), (240, 136), (1000, 333), (1024, 512)])
@pytest.mark.parametrize("sname", ["420", "422", "444"])
def test_decode_equals_cv2_imdecode(proj, w, h, sname):
    rng = np.random.default_rng(w * 7 + h)
    for q, rst, kind in ((95, 0, "noise"), (100, 0, "smooth"), (60, 5, "noise"), (20, 1, "sat")):
        if kind == "noise":
            img = rng.integers(0, 256, (h, w, 3), dtype=np.uint8)
        elif kind == "sat":
            img = (rng.integers(0, 2, (h, w, 3)) * 255).astype(np.uint8)
        else:
            img = synth.smooth(w, h, 3) if w >= 8 and h >= 8 else rng.integers(0, 256, (h, w, 3), dtype=np.uint8)
        data = cv2.imencode(".jpg", img, [cv2.IMWRITE_JPEG_QUALITY, q, cv2.IMWRITE_JPEG_RST_INTERVAL, rst,
                                          cv2.IMWRITE_JPEG_SAMPLING_FACTOR, SAMPLING[sname]])[1].tobytes()
        ref = cv2.imdecode(np.frombuffer(data, np.uint8), cv2.IMREAD_COLOR)
        assert np.array_equal(proj.decode_jpeg(data), ref), (q, rst, kind)


def test_upload_jpeg_equals_imread_path(pkg, proj, tmp_path):
    """A JPEG panorama decoded on the device gives the views of the cv2.imread path, bit for bit."""
    Wp, Hp, W, H, fov = 2048, 1024, 240, 136, 120
    pano = synth.smooth(Wp, Hp, 7)
    path = tmp_path / "pano.jpg"
    cv2.imwrite(str(path), pano)
    data = path.read_bytes()
    assert proj.jpeg_probe(data) == (Wp, Hp)
    shifts = [pkg.yaw_table(Wp, y)[2] for y in (0, 90, 180, 270)]
    consts = [pkg.pitch_constants(W, fov, p) for p in (30, 60, 90)]
    with proj.slots(1) as (s,):
        assert proj.upload_jpeg(s, data) == (Wp, Hp)
        got = proj.project(s, shifts, consts, W, H)
        proj.sync(s)
        assert np.array_equal(proj.download_pano(s, Wp, Hp), cv2.imread(str(path)))
        proj.upload(s, cv2.imread(str(path)))
        want = proj.project(s, shifts, consts, W, H)
        proj.sync(s)
    assert np.array_equal(got, want)
    with pytest.raises(pkg.P2PError) as ei:   # outside the decoder's subset: the caller falls back to cv2.imread
        proj.decode_jpeg(cv2.imencode(".png", pano[:64, :64])[1].tobytes())
    assert ei.value.code == -6


@pytest.mark.parametrize("sname", ["420", "422", "444"])
def test_progressive_files_decode_like_cv2(pkg, proj, tmp_path, sname):
    """Progressive JPEG files (SOF2): the scans are decoded by the library on the calling thread, IDCT / upsampling /
    colour run on the device; same pixels as cv2.imdecode, through decode_jpeg, upload_jpeg and the front door."""
    rng = np.random.default_rng(5)
    for (w, h, q, rst) in [(1, 1, 90, 0), (17, 33, 95, 0), (240, 136, 75, 7), (1000, 333, 50, 0), (2048, 1024, 92, 0)]:
        img = synth.smooth(w, h, 3) if w >= 8 and w != 240 else rng.integers(0, 256, (h, w, 3), dtype=np.uint8)
        data = cv2.imencode(".jpg", img, [cv2.IMWRITE_JPEG_QUALITY, q, cv2.IMWRITE_JPEG_RST_INTERVAL, rst, cv2.IMWRITE_JPEG_PROGRESSIVE, 1,
                                          cv2.IMWRITE_JPEG_SAMPLING_FACTOR, SAMPLING[sname]])[1].tobytes()
        assert b"\xff\xc2" in data[:800]
        ref = cv2.imdecode(np.frombuffer(data, np.uint8), cv2.IMREAD_COLOR)
        assert np.array_equal(proj.decode_jpeg(data), ref), (w, h, q, rst)
    Wp, Hp, W, H, fov = 2048, 1024, 240, 136, 120
    with proj.slots(1) as (s,):
        assert proj.upload_jpeg(s, data) == (Wp, Hp)
        assert np.array_equal(proj.download_pano(s, Wp, Hp), ref)
    path = tmp_path / "progressive.jpg"
    path.write_bytes(data)
    view = pkg.panorama_to_plane(path, fov, (W, H), 90, 60)
    assert np.array_equal(view, pkg.process_yaw_and_pitchs(cv2.imread(str(path)), 90, [60], W, H, fov)[0])
    g = cv2.imencode(".jpg", img[..., 1], [cv2.IMWRITE_JPEG_PROGRESSIVE, 1])[1].tobytes()   # grayscale, progressive
    assert np.array_equal(proj.decode_jpeg(g), cv2.imdecode(np.frombuffer(g, np.uint8), cv2.IMREAD_COLOR))


def test_front_end_jpeg_input_equals_imread_path(pkg, tmp_path):
    """A .jpg panorama goes through the device decoder in every front door (single image, directory pipeline,
    ``panorama_to_plane``); the results equal what the cv2.imread pixels give.  Files outside the decoder's subset
    (progressive) and fractional yaws take the fallbacks."""
    src = tmp_path / "in"
    src.mkdir()
    pano = synth.smooth(1024, 512, 21)
    cv2.imwrite(str(src / "a.jpg"), pano)
    cv2.imwrite(str(src / "b.jpeg"), pano[::-1].copy(), [cv2.IMWRITE_JPEG_SAMPLING_FACTOR, cv2.IMWRITE_JPEG_SAMPLING_FACTOR_444])
    cv2.imwrite(str(src / "c.jpg"), pano[:, ::-1].copy(), [cv2.IMWRITE_JPEG_PROGRESSIVE, 1])
    W, H, fov, yaws, pitches = 200, 120, 100, [0, 90, 30], [60, 120]    # yaw 30 is fractional on Wp = 1024
    out_png, out_jpg = tmp_path / "png", tmp_path / "jpg"
    pkg.main(str(src), str(out_png), yaws, pitches, W, H, num_workers=3, output_format="png", fov_deg=fov)
    pkg.main(str(src), str(out_jpg), yaws, pitches, W, H, num_workers=3, output_format="jpg", fov_deg=fov)
    for name in ("a.jpg", "b.jpeg", "c.jpg"):
        img = cv2.imread(str(src / name))
        stem = name.split(".")[0]
        for y in yaws:
            views = pkg.process_yaw_and_pitchs(img, y, pitches, W, H, fov)
            for p, view in zip(pitches, views):
                assert np.array_equal(cv2.imread(str(out_png / f"{stem}_{W}x{H}_yaw_{y}_pitch_{p}.png")), view), (name, y, p)
                assert (out_jpg / f"{stem}_{W}x{H}_yaw_{y}_pitch_{p}.jpg").read_bytes() == \
                    cv2.imencode(".jpg", view)[1].tobytes(), (name, y, p)
    one = pkg.panorama_to_plane(src / "a.jpg", fov, (W, H), 90, 60)
    assert np.array_equal(one, pkg.process_yaw_and_pitchs(cv2.imread(str(src / "a.jpg")), 90, [60], W, H, fov)[0])


def test_front_end_with_damaged_jpeg_files_equals_imread_path(pkg, tmp_path, capfd):
    """Damaged .jpg panoramas in the directory pipeline: the device decoder declines them (restart marker out of
    sequence, a block of out-of-range coefficients, a truncated scan) or decodes them like libjpeg (a flipped value bit);
    either way the files equal what the reference's cv2.imread pixels give, and a file cv2 cannot read is skipped."""
    src = tmp_path / "in"
    src.mkdir()
    pano = synth.smooth(1024, 512, 33)
    good = cv2.imencode(".jpg", pano, [cv2.IMWRITE_JPEG_RST_INTERVAL, 8])[1].tobytes()
    sos = good.find(b"\xff\xda")
    rst = good.find(b"\xff\xd2", sos)
    files = {"seq": bytearray(good), "cut": bytearray(good[:len(good) * 2 // 3]), "junk": bytearray(b"\xff\xd8" + bytes(200))}
    files["seq"][rst + 1] = 0xD5
    rng = np.random.default_rng(3)
    for k in range(6):                       # random single-bit damage in the scan: some declined, some decodable
        d = bytearray(good)
        d[int(rng.integers(sos + 14, len(good) - 2))] ^= 1 << int(rng.integers(0, 8))
        files[f"bit{k}"] = d
    for name, d in files.items():
        (src / f"{name}.jpg").write_bytes(bytes(d))
    W, H, fov, yaws, pitches = 160, 96, 100, [0, 180], [60, 90]
    out = tmp_path / "out"
    pkg.main(str(src), str(out), yaws, pitches, W, H, num_workers=3, output_format="png", fov_deg=fov)
    for name in files:
        img = cv2.imread(str(src / f"{name}.jpg"))
        for y in yaws:
            for p in pitches:
                f = out / f"{name}_{W}x{H}_yaw_{y}_pitch_{p}.png"
                if img is None:
                    assert not f.exists(), name
                    continue
                view = pkg.process_yaw_and_pitchs(img, y, [p], W, H, fov)[0]
                assert np.array_equal(cv2.imread(str(f)), view), (name, y, p)
    assert cv2.imread(str(src / "junk.jpg")) is None and cv2.imread(str(src / "seq.jpg")) is not None
    capfd.readouterr()


def test_adobe_marked_files_follow_libjpegs_colour_space_rule(pkg, proj):
    """Photoshop / Lightroom style files (Adobe APP14 marker, no JFIF) are decoded on the device when libjpeg takes them
    for YCbCr, declined (cv2 fallback) when it takes them for RGB."""
    from jpeg_damage import colour_space_variants

    img = synth.smooth(640, 320, 3)
    for label, data, ycc in colour_space_variants(img):
        ref = cv2.imdecode(np.frombuffer(data, np.uint8), cv2.IMREAD_COLOR)
        if ycc:
            assert np.array_equal(proj.decode_jpeg(data), ref), label
        else:
            with pytest.raises(pkg.P2PError) as e:
                proj.decode_jpeg(data)
            assert e.value.code == -6, label


def test_grayscale_files_decode_like_cv2_imread(pkg, proj, tmp_path):
    """Grayscale JPEG panoramas: B = G = R = Y as cv2.imread returns them, with the Huffman stage on the device and on the
    host, with restart intervals, and through the front end."""
    L = pkg._lib
    rng = np.random.default_rng(8)
    for k, (w, h) in enumerate([(1, 1), (9, 7), (200, 64), (1024, 512), (2048, 1024), (1001, 333)]):
        g = synth.smooth(w, h, k)[..., 1] if k % 2 else np.clip(synth.smooth(w, h, k)[..., 0].astype(int) + rng.integers(-9, 10, (h, w)), 0, 255).astype(np.uint8)
        for params in ([cv2.IMWRITE_JPEG_QUALITY, 95], [cv2.IMWRITE_JPEG_QUALITY, 70, cv2.IMWRITE_JPEG_RST_INTERVAL, 3]):
            data = cv2.imencode(".jpg", g, params)[1].tobytes()
            ref = cv2.imdecode(np.frombuffer(data, np.uint8), cv2.IMREAD_COLOR)
            n0 = proj.get_option(L.OPT_GPU_HUFFMAN_COUNT)
            assert np.array_equal(proj.decode_jpeg(data), ref), (w, h, params)
            assert proj.get_option(L.OPT_GPU_HUFFMAN_COUNT) == n0 + 1, "the device Huffman stage did not run"
            proj.set_option(L.OPT_GPU_HUFFMAN, 0)
            try:
                assert np.array_equal(proj.decode_jpeg(data), ref), (w, h, params, "host stage")
            finally:
                proj.set_option(L.OPT_GPU_HUFFMAN, 1)
    pano = synth.smooth(1024, 512, 5)[..., 0].copy()
    path = tmp_path / "gray.jpg"
    cv2.imwrite(str(path), pano)
    view = pkg.panorama_to_plane(path, 100, (200, 120), 90, 60)
    assert np.array_equal(view, pkg.process_yaw_and_pitchs(cv2.imread(str(path)), 90, [60], 200, 120, 100)[0])


def test_device_huffman_stage_is_used_and_equals_host_stage(pkg, proj):
    """Files without restart markers are Huffman-decoded on the device (self-synchronising subsequences); the pixels
    must equal the host decoder's (= cv2's), the fallback must work, and the option must switch the stage off."""
    L = pkg._lib
    rng = np.random.default_rng(99)
    smooth = synth.smooth(2048, 1024, 9)
    textured = np.clip(smooth.astype(np.int16) + rng.integers(-12, 13, smooth.shape, dtype=np.int16), 0, 255).astype(np.uint8)
    for img, q, samp in ((smooth, 95, "420"), (textured, 90, "420"), (textured, 85, "444"), (smooth[:333, :1000], 75, "422")):
        data = cv2.imencode(".jpg", img, [cv2.IMWRITE_JPEG_QUALITY, q, cv2.IMWRITE_JPEG_SAMPLING_FACTOR, SAMPLING[samp]])[1].tobytes()
        ref = cv2.imdecode(np.frombuffer(data, np.uint8), cv2.IMREAD_COLOR)
        n0 = proj.get_option(L.OPT_GPU_HUFFMAN_COUNT)
        got = proj.decode_jpeg(data)
        assert proj.get_option(L.OPT_GPU_HUFFMAN_COUNT) == n0 + 1, "the device Huffman stage did not run"
        assert np.array_equal(got, ref), (q, samp)
        proj.set_option(L.OPT_GPU_HUFFMAN, 0)
        try:
            host = proj.decode_jpeg(data)
            assert proj.get_option(L.OPT_GPU_HUFFMAN_COUNT) == n0 + 1
            # mode 2: the device stage with its plain rounds / write pass (tables in global memory), the yardstick of the
            # shared-memory kernels that mode 1 runs
            proj.set_option(L.OPT_GPU_HUFFMAN, 2)
            plain = proj.decode_jpeg(data)
            assert proj.get_option(L.OPT_GPU_HUFFMAN_COUNT) == n0 + 2
        finally:
            proj.set_option(L.OPT_GPU_HUFFMAN, 1)
        assert np.array_equal(host, ref) and np.array_equal(plain, ref)
    # white noise needs ~80 synchronisation rounds; restart intervals are independent scans with known start states
    noise = synth.noise(1024, 512, 1)
    for params in ([cv2.IMWRITE_JPEG_QUALITY, 95], [cv2.IMWRITE_JPEG_QUALITY, 90, cv2.IMWRITE_JPEG_RST_INTERVAL, 4],
                   [cv2.IMWRITE_JPEG_QUALITY, 90, cv2.IMWRITE_JPEG_RST_INTERVAL, 1],
                   [cv2.IMWRITE_JPEG_QUALITY, 80, cv2.IMWRITE_JPEG_RST_INTERVAL, 1000,
                    cv2.IMWRITE_JPEG_SAMPLING_FACTOR, SAMPLING["444"]]):
        img = noise if len(params) == 2 else textured
        data = cv2.imencode(".jpg", img, params)[1].tobytes()
        n0 = proj.get_option(L.OPT_GPU_HUFFMAN_COUNT)
        ref = cv2.imdecode(np.frombuffer(data, np.uint8), cv2.IMREAD_COLOR)
        assert np.array_equal(proj.decode_jpeg(data), ref), params
        assert proj.get_option(L.OPT_GPU_HUFFMAN_COUNT) == n0 + 1, params
        proj.set_option(L.OPT_GPU_HUFFMAN, 2)
        try:
            assert np.array_equal(proj.decode_jpeg(data), ref), params
        finally:
            proj.set_option(L.OPT_GPU_HUFFMAN, 1)
