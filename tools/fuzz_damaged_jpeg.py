"""Differential campaign on damaged JPEG files (tests/jpeg_damage.py) against cv2.imdecode on the GPU box:
every file must be declined (-6) or decode to cv2's pixels; files cv2 cannot read must be declined.
    python tools/fuzz_damaged_jpeg.py [--seeds 1 2 3] [--count 2000] [--max-w 700] [--max-h 500]"""
import argparse
import importlib
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    import cv2
    from jpeg_damage import damaged_files

    ap = argparse.ArgumentParser()
    ap.add_argument("--seeds", type=int, nargs="+", default=[1, 2, 3])
    ap.add_argument("--count", type=int, default=2000)
    ap.add_argument("--max-w", type=int, default=700)
    ap.add_argument("--max-h", type=int, default=500)
    ap.add_argument("--gray-every", type=int, default=0, help="every k-th file is a grayscale JPEG")
    ap.add_argument("--progressive-every", type=int, default=0, help="every k-th file is a progressive JPEG")
    a = ap.parse_args()
    pkg = importlib.import_module("360-to-planer-images_b200")
    L = pkg._lib
    proj = pkg.Projector(0, n_slots=2)
    devnull = os.open(os.devnull, os.O_WRONLY)
    os.dup2(devnull, 2)                      # libjpeg's warnings
    for stage in (1, 0):
        proj.set_option(L.OPT_GPU_HUFFMAN, stage)
        st = dict(stage="device" if stage else "host", files=0, declined=0, same=0, cv2_unreadable=0, differ=[], accepted_unreadable=[])
        n0 = proj.get_option(L.OPT_GPU_HUFFMAN_COUNT)
        t0 = time.time()
        for seed in a.seeds:
            for label, data in damaged_files(seed, a.count, max_wh=(a.max_w, a.max_h), gray_every=a.gray_every, progressive_every=a.progressive_every):
                st["files"] += 1
                ref = cv2.imdecode(np.frombuffer(data, np.uint8), cv2.IMREAD_COLOR)
                st["cv2_unreadable"] += ref is None
                try:
                    got = proj.decode_jpeg(data)
                except pkg.P2PError as e:
                    assert e.code == -6, (label, e)
                    st["declined"] += 1
                    continue
                if ref is None:
                    st["accepted_unreadable"].append((seed, label))
                elif np.array_equal(got, ref):
                    st["same"] += 1
                else:
                    st["differ"].append((seed, label))
        st["device_huffman_runs"] = proj.get_option(L.OPT_GPU_HUFFMAN_COUNT) - n0
        st["seconds"] = round(time.time() - t0, 1)
        print(json.dumps(st), flush=True)
    proj.close()


if __name__ == "__main__":
    try:
        main()
    except Exception:                        # stderr is muted above
        import traceback

        print(traceback.format_exc(), flush=True)
        sys.exit(1)
