"""Where the time of the "JPEG panorama -> 12 PNG files" flow goes (bench.py extras.png_files): single-thread step times and
throughput against the number of images in flight.   python tools/profile_png_flow.py"""
import importlib
import json
import os
import sys
import time
from concurrent.futures import ThreadPoolExecutor

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from tools import synth_inputs as synth  # noqa: E402


def main():
    import cv2

    pkg = importlib.import_module("360-to-planer-images_b200")
    W, H, FOV = 1920, 1080, 120
    yaws, pitches = [0, 90, 180, 270], [30, 60, 90]
    pano = synth.smooth(8192, 4096, 1)
    data = cv2.imencode(".jpg", pano)[1].tobytes()
    proj = pkg.Projector(0, n_slots=16)
    consts = [pkg.pitch_constants(W, FOV, p) for p in pitches]
    shifts = np.array([pkg.yaw_table(8192, y)[2] for y in yaws], np.int32)

    def one(_=None, keep=True):
        with proj.slots(1) as (s,):
            proj.upload_jpeg(s, data)
            return proj.process_image_png(s, None, shifts, consts, W, H, want_pixels=False)[0]

    for _ in range(3):
        one()
    # single-thread step times
    t = {}
    with proj.slots(1) as (s,):
        for _ in range(3):
            t0 = time.perf_counter(); proj.upload_jpeg(s, data); t1 = time.perf_counter()
            files = proj.process_image_png(s, None, shifts, consts, W, H, want_pixels=False)[0]; t2 = time.perf_counter()
        t["upload_jpeg_ms"] = (t1 - t0) * 1e3
        t["process_image_png_ms (incl. tobytes)"] = (t2 - t1) * 1e3
        buf = proj._file_buffer(s, 12, W * H * 4 + 4096)
        sizes = [len(f) for f in files]
        t0 = time.perf_counter()
        for _ in range(5):
            out = [buf[i, :sizes[i]].tobytes() for i in range(12)]
        t["tobytes_12_files_ms"] = (time.perf_counter() - t0) / 5 * 1e3
        t["file_bytes"] = sum(sizes)
    print(json.dumps(t), flush=True)
    for n_thr in (1, 2, 4, 8, 12):
        with ThreadPoolExecutor(n_thr) as ex:
            list(ex.map(one, range(n_thr)))
            t0 = time.perf_counter()
            list(ex.map(one, range(24)))
            dt = (time.perf_counter() - t0) / 24
        print(json.dumps({"in_flight": n_thr, "ms_per_image": round(dt * 1e3, 2)}), flush=True)
    proj.close()


if __name__ == "__main__":
    main()
