"""Wall time of the directory front end on a folder of 8192x4096 JPEG panoramas (12 x 1920x1080 views each), GPU flow only:
    python tools/time_folder.py [--files 8] [--repeat 3] [--devices 1 2 4 8]
(tools/cpu_baselines.py times the reference's flow beside it).  ``--devices``: GPU counts to run the folder on through
``main(devices=[0 .. n-1])`` - the product's own multi-GPU front end, files sharded round-robin, one pipeline per device
(ref :320-341 walks the files serially); the output files of every run are hashed and must equal the first run's."""
import argparse
import hashlib
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
    ap.add_argument("--devices", type=int, nargs="+", default=[1])
    ap.add_argument("--formats", nargs="+", default=["jpg", "png"])
    a = ap.parse_args()
    pkg = importlib.import_module("360-to-planer-images_b200")
    td = Path(tempfile.mkdtemp(dir="/dev/shm" if os.path.isdir("/dev/shm") else None))
    try:
        folder = td / "in"
        folder.mkdir()
        for i in range(a.files):
            cv2.imwrite(str(folder / f"p{i}.jpg"), synth.smooth(8192, 4096, 100 + i))
        n_avail = pkg._lib.load().p2p_device_count()
        for fmt in a.formats:
            first_hash = None
            for n_dev in [n for n in a.devices if n <= n_avail]:
                devs = list(range(n_dev)) if n_dev > 1 else None
                times = []
                for r in range(a.repeat + 1):        # first pass warms slots and buffers
                    out = td / f"out_{fmt}_{n_dev}_{r}"
                    t0 = time.perf_counter()
                    pkg.main(str(folder), str(out), [0, 90, 180, 270], [30, 60, 90], 1920, 1080, num_workers=a.workers,
                             output_format=fmt, fov_deg=120, devices=devs)
                    times.append(time.perf_counter() - t0)
                    h = hashlib.sha256()
                    nbytes = 0
                    for f in sorted(out.iterdir()):
                        data = f.read_bytes()
                        h.update(f.name.encode())
                        h.update(data)
                        nbytes += len(data)
                    shutil.rmtree(out)
                first_hash = first_hash or h.hexdigest()
                best = min(times[1:])
                print(json.dumps({"folder": f"{a.files} x 8192x4096 jpg -> {a.files * 12} x 1920x1080 {fmt}", "n_gpus": n_dev,
                                  "workers": a.workers, "warmup_s": round(times[0], 3),
                                  "seconds": [round(t, 4) for t in times[1:]], "ms_per_image": round(best / a.files * 1e3, 2),
                                  "mpix_s": round(a.files * 12 * 1920 * 1080 / best / 1e6, 1), "output_bytes": nbytes,
                                  "files_identical_to_first_run": h.hexdigest() == first_hash}), flush=True)
    finally:
        shutil.rmtree(td, ignore_errors=True)


if __name__ == "__main__":
    main()
