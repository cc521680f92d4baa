"""Seeded random differential tests of the device JPEG encoder / decoder against cv2 (libjpeg-turbo), plus the
largest BASELINE shapes (4K views, a 16K panorama)."""
import cv2
import numpy as np
import pytest

from oracle import synth

pytestmark = pytest.mark.gpu

SAMPLING = [cv2.IMWRITE_JPEG_SAMPLING_FACTOR_420, cv2.IMWRITE_JPEG_SAMPLING_FACTOR_422, cv2.IMWRITE_JPEG_SAMPLING_FACTOR_444]


def random_image(rng, w, h):
    kind = rng.integers(0, 5)
    if kind == 0:
        return rng.integers(0, 256, (h, w, 3), dtype=np.uint8)
    if kind == 1:
        return (rng.integers(0, 2, (h, w, 3)) * 255).astype(np.uint8)
    if kind == 2:
        return np.full((h, w, 3), rng.integers(0, 256, 3), np.uint8)
    yy, xx = np.mgrid[0:h, 0:w]
    img = np.stack([(127 + 120 * np.sin(xx / rng.uniform(2, 30) + c) * np.cos(yy / rng.uniform(2, 30))) for c in range(3)], -1)
    if kind == 4:
        img = img + rng.integers(-20, 21, img.shape)
    return np.clip(img, 0, 255).astype(np.uint8)


def test_encoder_fuzz(proj):
    rng = np.random.default_rng(2024)
    for i in range(48):
        w, h = int(rng.integers(1, 320)), int(rng.integers(1, 240))
        q = int(rng.integers(1, 101))
        img = random_image(rng, w, h)
        ref = cv2.imencode(".jpg", img, [cv2.IMWRITE_JPEG_QUALITY, q])[1].tobytes()
        assert proj.encode_jpeg(img, quality=q)[0] == ref, (i, w, h, q)


def test_decoder_fuzz(proj):
    rng = np.random.default_rng(4202)
    for i in range(48):
        w, h = int(rng.integers(1, 320)), int(rng.integers(1, 240))
        q = int(rng.integers(5, 101))
        rst = int(rng.integers(0, 6))
        samp = SAMPLING[int(rng.integers(0, 3))]
        img = random_image(rng, w, h)
        data = cv2.imencode(".jpg", img, [cv2.IMWRITE_JPEG_QUALITY, q, cv2.IMWRITE_JPEG_RST_INTERVAL, rst,
                                          cv2.IMWRITE_JPEG_SAMPLING_FACTOR, samp])[1].tobytes()
        ref = cv2.imdecode(np.frombuffer(data, np.uint8), cv2.IMREAD_COLOR)
        assert np.array_equal(proj.decode_jpeg(data), ref), (i, w, h, q, rst, samp)


def test_largest_shapes(proj):
    view = synth.smooth(3840, 2160, 5)                                  # a C4 view
    assert proj.encode_jpeg(view)[0] == cv2.imencode(".jpg", view)[1].tobytes()
    assert proj.encode_png(view)[0] == cv2.imencode(".png", view)[1].tobytes()
    pano = synth.smooth(16384, 8192, 6)                                 # the C4 panorama as a JPEG file
    data = cv2.imencode(".jpg", pano, [cv2.IMWRITE_JPEG_QUALITY, 90])[1].tobytes()
    ref = cv2.imdecode(np.frombuffer(data, np.uint8), cv2.IMREAD_COLOR)
    got = proj.decode_jpeg(data)
    assert np.array_equal(got, ref)
