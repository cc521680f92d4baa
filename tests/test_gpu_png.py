"""GPU PNG encoder (csrc/p2p_png.cuh) through the C ABI: the files must be byte-identical to what the reference's
``cv2.imwrite(<name>.png, view)`` (ref :277, the default output format) writes."""
import cv2
import numpy as np
import pytest

from oracle import synth

pytestmark = pytest.mark.gpu


def ref_png(img):
    return cv2.imencode(".png", img)[1].tobytes()


def textured(w, h, seed, amp=6):
    rng = np.random.default_rng(seed)
    return np.clip(synth.smooth(w, h, seed).astype(int) + rng.integers(-amp, amp + 1, (h, w, 3)), 0, 255).astype(np.uint8)


@pytest.mark.parametrize("w,h", [(100, 64), (240, 136), (333, 200), (640, 480), (1000, 333), (1920, 1080)])
@pytest.mark.parametrize("kind", ["smooth", "textured", "flat", "stripes"])
def test_encode_png_equals_cv2(proj, w, h, kind):
    if kind == "smooth":
        img = synth.smooth(w, h, 3)
    elif kind == "textured":
        img = textured(w, h, 4)
    elif kind == "flat":
        img = np.full((h, w, 3), 200, np.uint8)
        img[h // 3:, w // 2:] = (10, 20, 30)
    else:
        rng = np.random.default_rng(w + h)
        img = np.repeat(rng.integers(0, 256, (h, (w + 9) // 10, 3), dtype=np.uint8), 10, axis=1)[:, :w].copy()
    got = proj.encode_png(img)[0]
    assert got is not None, "the device encoder declined an image it should handle"
    assert got == ref_png(img)


def test_encode_png_batch_with_noise(proj):
    imgs = np.stack([synth.smooth(320, 200, 1), textured(320, 200, 2), np.zeros((200, 320, 3), np.uint8),
                     synth.noise(320, 200, 3)])
    files = proj.encode_png(imgs)
    for f, img in zip(files, imgs):    # white noise: every deflate block is one zlib stores uncompressed
        assert f == ref_png(img)


@pytest.mark.parametrize("w,h,kind", [(1, 1, "noise"), (1, 40, "noise"), (1, 300, "smooth"), (50, 1, "noise"), (2, 2, "smooth"),
                                      (5, 3, "noise"), (10, 10, "smooth"), (40, 30, "smooth"), (60, 50, "textured"),
                                      (73, 74, "noise"), (80, 68, "smooth"), (100, 54, "textured"), (128, 42, "flat"),
                                      (102, 80, "noise"), (640, 480, "noise"), (400, 300, "mixed"), (500, 300, "noise+runs"),
                                      (85, 1, "smooth"), (85, 2, "smooth"), (85, 8, "noise"), (85, 32, "smooth"), (85, 64, "smooth"),
                                      (85, 65, "smooth"), (86, 63, "textured"), (21, 1, "noise")])
def test_small_images_stored_blocks_and_chunk_boundaries(proj, w, h, kind):
    """libpng's small-image cases (zlib window bits in the stream header for <= 16384 bytes of data, filter type 0 for
    width 1), blocks zlib stores uncompressed (noise, also mixed with compressible blocks and with runs inside), and a
    stream that ends exactly on an IDAT chunk boundary (102 x 80 noise: 24576 bytes = 3 full chunks, no empty one)."""
    if kind == "noise":
        img = synth.noise(w, h, w + h)
    elif kind == "smooth":
        img = synth.smooth(w, h, w)
    elif kind == "textured":
        img = textured(w, h, h)
    elif kind == "flat":
        img = np.full((h, w, 3), 99, np.uint8)
    elif kind == "mixed":
        img = synth.smooth(w, h, 3)
        img[:h // 2] = synth.noise(w, h // 2, 4)
    else:
        img = synth.noise(w, h, 5)
        img[:, ::7] = 128
        img[100:110] = 7
    got = proj.encode_png(img)[0]
    assert got is not None
    assert got == ref_png(img)


def test_png_fuzz(proj):
    rng = np.random.default_rng(77)
    for i in range(80):
        if i % 4 == 3:
            w, h = int(rng.integers(1, 90)), int(rng.integers(1, 70))       # small: <= 16384 bytes of filtered data
        else:
            w, h = int(rng.integers(80, 400)), int(rng.integers(70, 300))
        kind = int(rng.integers(0, 6))
        if kind == 0:
            img = synth.smooth(w, h, i)
        elif kind == 1:
            img = textured(w, h, i, amp=int(rng.integers(1, 30)))
        elif kind == 2:
            img = np.repeat(np.repeat(rng.integers(0, 256, ((h + 7) // 8, (w + 7) // 8, 3), dtype=np.uint8), 8, axis=0), 8, axis=1)[:h, :w].copy()
        elif kind == 3:
            img = np.full((h, w, 3), rng.integers(0, 256, 3), np.uint8)
            img[rng.integers(0, h):, rng.integers(0, w):] = rng.integers(0, 256, 3)
        elif kind == 4:
            img = synth.noise(w, h, i)
        else:                                                               # noise with compressible rows in between
            img = synth.noise(w, h, i)
            img[rng.integers(0, h)::3] = rng.integers(0, 256, 3)
        got = proj.encode_png(img)[0]
        assert got is not None, (i, w, h, kind)
        assert got == ref_png(img), (i, w, h, kind)


def test_front_end_png_files_equal_imwrite(pkg, tmp_path):
    """The default output format through every front door: the .png files are the bytes cv2.imwrite writes for the same
    views (device-encoded where the encoder handles the view, cv2 otherwise), for png and jpg inputs, fractional yaws too."""
    src = tmp_path / "in"
    src.mkdir()
    pano = synth.smooth(1024, 512, 31)
    cv2.imwrite(str(src / "a.png"), pano)
    cv2.imwrite(str(src / "b.jpg"), pano[::-1].copy())
    cv2.imwrite(str(src / "n.png"), synth.noise(1024, 512, 2))          # noise views: the device encoder declines them
    W, H, fov, yaws, pitches = 200, 120, 100, [0, 90, 30], [60, 120]       # yaw 30 is fractional on Wp = 1024
    out_dir, out_single = tmp_path / "dir", tmp_path / "single"
    pkg.main(str(src), str(out_dir), yaws, pitches, W, H, num_workers=3, fov_deg=fov)           # png is the default
    out_single.mkdir()
    pkg.process_single_image(src / "b.jpg", out_single, yaws, pitches, W, H, num_workers=2, fov_deg=fov)
    assert len(list(out_dir.iterdir())) == 3 * len(yaws) * len(pitches)
    for name in ("a.png", "b.jpg", "n.png"):
        img = cv2.imread(str(src / name))
        stem = name.split(".")[0]
        for y in yaws:
            views = pkg.process_yaw_and_pitchs(img, y, pitches, W, H, fov)
            for p, view in zip(pitches, views):
                want = cv2.imencode(".png", view)[1].tobytes()
                assert (out_dir / f"{stem}_{W}x{H}_yaw_{y}_pitch_{p}.png").read_bytes() == want, (name, y, p)
                if stem == "b":
                    assert (out_single / f"{stem}_{W}x{H}_yaw_{y}_pitch_{p}.png").read_bytes() == want
