"""PNG encode throughput: the GPU encoder (device-resident views -> files on the host) against cv2.imencode('.png') on the
box's host cores, same views, byte identity checked (run on the GPU box).

    python tools/bench_png.py > gpurun_out/png_bench.jsonl
"""
import json
import os
import sys
import time
from concurrent.futures import ThreadPoolExecutor
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import __graft_entry__ as g  # noqa: E402
from tools import synth_inputs as synth  # noqa: E402

import bench  # noqa: E402


def main():
    import cv2

    g.build()
    pkg = g.load_package()
    proj = pkg.Projector(0, n_slots=4)
    W, H = bench.W, bench.H
    consts = [pkg.pitch_constants(W, bench.FOV, p) for p in bench.PITCHES]
    shifts = [pkg.yaw_table(bench.WP, y)[2] for y in bench.YAWS]
    rng = np.random.default_rng(0)
    smooth = synth.smooth(bench.WP, bench.HP, 0)
    textured = np.clip(smooth.astype(np.int16) + rng.integers(-12, 13, smooth.shape, dtype=np.int16), 0, 255).astype(np.uint8)
    for kind, pano in (("smooth", smooth), ("textured (smooth + noise of amplitude 12)", textured)):
        with proj.slots(1) as (s,):
            proj.upload(s, pano)
            views = proj.project(s, shifts, consts, W, H)
            proj.sync(s)
            files, _ = proj.process_image_png(s, None, shifts, consts, W, H, want_pixels=False)   # warm-up
            reps = 5
            t0 = time.perf_counter()
            for _ in range(reps):
                files, _ = proj.process_image_png(s, None, shifts, consts, W, H, want_pixels=False)
            gpu_s = (time.perf_counter() - t0) / reps
        flat = views.reshape(-1, H, W, 3)
        handled = [f is not None for f in files]
        same = all(f == cv2.imencode(".png", v)[1].tobytes() for f, v in zip(files, flat) if f is not None)
        workers = max(1, int((os.cpu_count() or 1) * 0.9))
        with ThreadPoolExecutor(workers) as ex:
            list(ex.map(lambda v: cv2.imencode(".png", v)[1], flat))
            t0 = time.perf_counter()
            for _ in range(3):
                list(ex.map(lambda v: cv2.imencode(".png", v)[1], flat))
            cpu_s = (time.perf_counter() - t0) / 3
        t0 = time.perf_counter()
        cv2.imencode(".png", flat[0])
        cpu1_s = time.perf_counter() - t0
        mpix = flat.shape[0] * W * H / 1e6
        print(json.dumps({
            "panorama": kind, "views": int(flat.shape[0]), "size": [W, H], "handled_on_device": int(sum(handled)),
            "byte_identical_to_cv2": bool(same), "file_bytes_total": int(sum(len(f) for f in files if f)),
            "raw_bytes_total": int(flat.nbytes), "gpu_project_encode_readback_ms": gpu_s * 1e3, "gpu_mpix_s": mpix / gpu_s,
            "cpu_imencode_threads": workers, "cpu_imencode_ms": cpu_s * 1e3, "cpu_mpix_s": mpix / cpu_s,
            "cpu_single_thread_ms_per_view": cpu1_s * 1e3}), flush=True)
    proj.close()


if __name__ == "__main__":
    main()
