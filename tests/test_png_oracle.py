"""The PNG oracle (oracle/png_model.py: step-by-step restatement of zlib's deflate_rle + trees.c and libpng's framing at
OpenCV's settings) pinned against zlib itself and byte for byte against ``cv2.imencode('.png')`` (ref :277)."""
import zlib

import cv2
import numpy as np
import pytest

from oracle import png_model as pm
from oracle import synth


def images():
    rng = np.random.default_rng(5)
    yield "smooth", synth.smooth(300, 200, 1)
    yield "flat", np.full((64, 100, 3), 9, np.uint8)
    yield "textured", np.clip(synth.smooth(256, 128, 2).astype(int) + rng.integers(-5, 6, (128, 256, 3)), 0, 255).astype(np.uint8)
    yield "stripes", np.repeat(rng.integers(0, 256, (90, 30, 3), dtype=np.uint8), 10, axis=1)
    yield "noise", synth.noise(96, 64, 1)
    yield "long_runs", np.concatenate([np.zeros((40, 500, 3), np.uint8), synth.smooth(500, 30, 3)], axis=0)


@pytest.mark.parametrize("name,img", list(images()), ids=[n for n, _ in images()])
def test_deflate_model_equals_zlib(name, img):
    raw = pm.sub_filter(img)
    co = zlib.compressobj(1, zlib.DEFLATED, 15, 8, zlib.Z_RLE)
    assert pm.deflate_rle(raw)[0] == co.compress(raw) + co.flush()


@pytest.mark.parametrize("name,img", list(images()), ids=[n for n, _ in images()])
def test_png_oracle_equals_cv2_imencode(name, img):
    assert pm.encode_png(img) == cv2.imencode(".png", img)[1].tobytes()


def test_small_images_stored_blocks_and_chunk_boundaries():
    """libpng's small-image cases (window bits in the zlib header for <= 16384 bytes of data, filter type 0 for width 1),
    blocks zlib stores uncompressed, and a stream ending exactly on an IDAT boundary (102 x 80 noise = 3 full chunks)."""
    rng = np.random.default_rng(0)
    for t in range(150):
        w, h = int(rng.integers(1, 120)), int(rng.integers(1, 90))
        if t % 10 == 0:
            w = 1
        if t % 10 == 1:
            h = 1
        k = t % 3
        img = synth.noise(w, h, t) if k == 0 else synth.smooth(w, h, t) if k == 1 else np.full((h, w, 3), int(rng.integers(0, 256)), np.uint8)
        assert pm.encode_png(img) == cv2.imencode(".png", img)[1].tobytes(), (w, h, k)
    boundary = synth.noise(102, 80, 1)
    png = pm.encode_png(boundary)
    assert png == cv2.imencode(".png", boundary)[1].tobytes()
    assert png.count(b"IDAT") == 3 and len(png) == 8 + 25 + 3 * (8192 + 12) + 12
    mixed = synth.smooth(400, 300, 3)
    mixed[:150] = synth.noise(400, 150, 4)
    _, kinds = pm.deflate_rle(pm.sub_filter(mixed))
    assert "stored" in kinds and "dynamic" in kinds
    assert pm.encode_png(mixed) == cv2.imencode(".png", mixed)[1].tobytes()


@pytest.mark.parametrize("w,h", [(85, 1), (85, 2), (85, 4), (85, 8), (85, 16), (85, 32), (85, 64), (85, 65), (86, 63), (86, 1), (21, 1), (21, 2)])
def test_window_bits_at_the_power_of_two_boundaries(w, h):
    """85 pixels make a 256-byte row: data sizes of exactly 256 ... 16384 bytes and their neighbours hit every step of
    libpng's window-bits rule (optimize_cmf)."""
    img = synth.smooth(w, h, w + h)
    png = pm.encode_png(img)
    assert png == cv2.imencode(".png", img)[1].tobytes()
    n = h * (1 + 3 * w)
    assert pm.zlib_header(n) == png[41:43]          # signature 8 + IHDR 25 + IDAT length / type 8
    assert (pm.zlib_header(n) == b"\x78\x01") == (n > 16384)
