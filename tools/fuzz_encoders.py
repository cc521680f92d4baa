"""Differential campaign of the device PNG / JPEG encoders against cv2.imencode on seeded random images
(sizes 1 .. max, smooth / textured / blocky / flat / noise / noise with runs / binary content):
    python tools/fuzz_encoders.py [--count 3000] [--max-w 700] [--max-h 500] [--seed 1]
Prints one JSON line per codec: images, byte-identical files, declined (sizes[i] = 0), mismatches."""
import argparse
import importlib
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from tools import synth_inputs as synth  # noqa: E402


def random_image(rng, i, max_w, max_h):
    if i % 5 == 4:
        w, h = int(rng.integers(1, 90)), int(rng.integers(1, 70))
    else:
        w, h = int(rng.integers(1, max_w + 1)), int(rng.integers(1, max_h + 1))
    kind = int(rng.integers(0, 8))
    if kind == 0:
        img = synth.smooth(w, h, i)
    elif kind == 1:
        amp = int(rng.integers(1, 40))
        img = np.clip(synth.smooth(w, h, i).astype(int) + rng.integers(-amp, amp + 1, (h, w, 3)), 0, 255).astype(np.uint8)
    elif kind == 2:
        b = int(rng.integers(2, 17))
        img = np.repeat(np.repeat(rng.integers(0, 256, ((h + b - 1) // b, (w + b - 1) // b, 3), dtype=np.uint8), b, axis=0), b, axis=1)[:h, :w].copy()
    elif kind == 3:
        img = np.full((h, w, 3), rng.integers(0, 256, 3), np.uint8)
        img[rng.integers(0, h):, rng.integers(0, w):] = rng.integers(0, 256, 3)
    elif kind == 4:
        img = synth.noise(w, h, i)
    elif kind == 5:
        img = synth.noise(w, h, i)
        img[rng.integers(0, h)::int(rng.integers(2, 6))] = rng.integers(0, 256, 3)
    elif kind == 6:
        img = (rng.integers(0, 2, (h, w, 3)) * 255).astype(np.uint8)
    else:
        img = synth.smooth(w, h, i)
        img[:h // 2] = synth.noise(w, h // 2, i + 1) if h >= 2 else img[:h // 2]
    return kind, np.ascontiguousarray(img)


def main():
    import cv2

    ap = argparse.ArgumentParser()
    ap.add_argument("--count", type=int, default=3000)
    ap.add_argument("--max-w", type=int, default=700)
    ap.add_argument("--max-h", type=int, default=500)
    ap.add_argument("--seed", type=int, default=1)
    a = ap.parse_args()
    pkg = importlib.import_module("360-to-planer-images_b200")
    proj = pkg.Projector(0, n_slots=2)
    for codec in ("png", "jpg"):
        rng = np.random.default_rng(a.seed)
        st = dict(codec=codec, images=0, identical=0, declined=0, mismatches=[])
        t0 = time.time()
        for i in range(a.count):
            kind, img = random_image(rng, i, a.max_w, a.max_h)
            st["images"] += 1
            if codec == "png":
                ref = cv2.imencode(".png", img)[1].tobytes()
                got = proj.encode_png(img)[0]
            else:
                q = int(rng.integers(1, 101))
                ref = cv2.imencode(".jpg", img, [cv2.IMWRITE_JPEG_QUALITY, q])[1].tobytes()
                got = proj.encode_jpeg(img, quality=q)[0]
            if got is None:
                st["declined"] += 1
            elif got == ref:
                st["identical"] += 1
            else:
                st["mismatches"].append((i, kind, img.shape[1], img.shape[0]))
        st["seconds"] = round(time.time() - t0, 1)
        print(json.dumps(st), flush=True)
    proj.close()


if __name__ == "__main__":
    main()
