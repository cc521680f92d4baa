"""Wall time of the directory front end on a folder of 8192x4096 JPEG panoramas (12 x 1920x1080 views each), GPU flow only:
    python tools/time_folder.py [--files 8] [--repeat 3]      (tools/cpu_baselines.py times the reference's flow beside it)"""
import argparse
import importlib
import json
import os
import shutil
import sys
import tempfile
import time
from pathlib import Path

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from tools import synth_inputs as synth  # noqa: E402


def main():
    import cv2

    ap = argparse.ArgumentParser()
    ap.add_argument("--files", type=int, default=8)
    ap.add_argument("--repeat", type=int, default=3)
    ap.add_argument("--workers", type=int, default=max(1, int((os.cpu_count() or 1) * 0.9)))
    a = ap.parse_args()
    pkg = importlib.import_module("360-to-planer-images_b200")
    td = Path(tempfile.mkdtemp(dir="/dev/shm" if os.path.isdir("/dev/shm") else None))
    try:
        folder = td / "in"
        folder.mkdir()
        for i in range(a.files):
            cv2.imwrite(str(folder / f"p{i}.jpg"), synth.smooth(8192, 4096, 100 + i))
        for fmt in ("jpg", "png"):
            times = []
            for r in range(a.repeat + 1):        # first pass warms slots and buffers
                out = td / f"out_{fmt}_{r}"
                t0 = time.perf_counter()
                pkg.main(str(folder), str(out), [0, 90, 180, 270], [30, 60, 90], 1920, 1080, num_workers=a.workers,
                         output_format=fmt, fov_deg=120)
                times.append(time.perf_counter() - t0)
                nbytes = sum(f.stat().st_size for f in out.iterdir())
                shutil.rmtree(out)
            print(json.dumps({"folder": f"{a.files} x 8192x4096 jpg -> {a.files * 12} x 1920x1080 {fmt}", "workers": a.workers,
                              "warmup_s": round(times[0], 3), "seconds": [round(t, 4) for t in times[1:]],
                              "ms_per_image": round(min(times[1:]) / a.files * 1e3, 2), "output_bytes": nbytes}), flush=True)
    finally:
        shutil.rmtree(td, ignore_errors=True)


if __name__ == "__main__":
    main()
