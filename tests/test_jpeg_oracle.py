"""The JPEG oracle (oracle/jpeg_model.py: integer restatement of libjpeg-turbo at OpenCV's defaults) pinned byte for
byte against the real encoder behind the reference's ``cv2.imwrite`` (ref :277)."""
import cv2
import numpy as np
import pytest

from oracle import jpeg_model as jm
from oracle import synth


def images():
    rng = np.random.default_rng(7)
    for (h, w) in [(1, 1), (8, 8), (16, 16), (17, 33), (37, 53), (40, 56), (136, 240)]:
        yield f"noise{w}x{h}", rng.integers(0, 256, (h, w, 3), dtype=np.uint8)
        yy, xx = np.mgrid[0:h, 0:w]
        yield f"smooth{w}x{h}", np.stack([(127 + 120 * np.sin(xx / 7.0 + c) * np.cos(yy / 5.0)).astype(np.uint8)
                                          for c in range(3)], -1)
    yield "flat", np.full((24, 40, 3), 255, np.uint8)
    yield "black", np.zeros((9, 70, 3), np.uint8)
    yield "pano", synth.smooth(256, 128, 1)


@pytest.mark.parametrize("name,img", list(images()), ids=[n for n, _ in images()])
def test_oracle_equals_cv2_imencode(name, img):
    ref = cv2.imencode(".jpg", img)[1].tobytes()
    assert jm.encode(img) == ref


@pytest.mark.parametrize("quality", [1, 30, 50, 75, 95, 100])
def test_oracle_quality_scaling(quality):
    img = synth.smooth(96, 64, 2)
    ref = cv2.imencode(".jpg", img, [cv2.IMWRITE_JPEG_QUALITY, quality])[1].tobytes()
    assert jm.encode(img, quality) == ref


def test_oracle_file_equals_imwrite(tmp_path):
    img = synth.noise(64, 48, 3)
    path = tmp_path / "x.jpg"
    cv2.imwrite(str(path), img)   # the call at ref :277
    assert path.read_bytes() == jm.encode(img)
