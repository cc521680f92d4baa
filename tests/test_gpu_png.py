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


def test_encode_png_batch_and_declined_images(proj):
    imgs = np.stack([synth.smooth(320, 200, 1), textured(320, 200, 2), np.zeros((200, 320, 3), np.uint8),
                     synth.noise(320, 200, 3)])
    files = proj.encode_png(imgs)
    for f, img in zip(files[:3], imgs[:3]):
        assert f == ref_png(img)
    assert files[3] is None            # white noise: zlib stores such blocks uncompressed, left to cv2
    assert proj.encode_png(synth.smooth(40, 30, 1))[0] is None   # tiny image: libpng shrinks the zlib window


def test_png_fuzz(proj):
    rng = np.random.default_rng(77)
    handled = 0
    for i in range(40):
        w, h = int(rng.integers(80, 400)), int(rng.integers(70, 300))
        kind = int(rng.integers(0, 4))
        if kind == 0:
            img = synth.smooth(w, h, i)
        elif kind == 1:
            img = textured(w, h, i, amp=int(rng.integers(1, 30)))
        elif kind == 2:
            img = np.repeat(np.repeat(rng.integers(0, 256, ((h + 7) // 8, (w + 7) // 8, 3), dtype=np.uint8), 8, axis=0), 8, axis=1)[:h, :w].copy()
        else:
            img = np.full((h, w, 3), rng.integers(0, 256, 3), np.uint8)
            img[rng.integers(0, h):, rng.integers(0, w):] = rng.integers(0, 256, 3)
        got = proj.encode_png(img)[0]
        if got is not None:
            handled += 1
            assert got == ref_png(img), (i, w, h, kind)
    assert handled >= 30
