"""JPEG encode throughput: the GPU encoder (device-resident views -> files on the host) against cv2.imencode on the
box's host cores, same views (run on the GPU box).

    python tools/bench_jpeg.py > gpurun_out/jpeg_bench.jsonl
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
    assert proj.n_slots == 4
    W, H = bench.W, bench.H
    consts = [pkg.pitch_constants(W, bench.FOV, p) for p in bench.PITCHES]
    shifts = [pkg.yaw_table(bench.WP, y)[2] for y in bench.YAWS]
    for kind in ("smooth", "noise"):
        pano = synth.make(kind, bench.WP, bench.HP, 0)
        with proj.slots(1) as (s,):
            proj.upload(s, pano)
            views = proj.project(s, shifts, consts, W, H)
            proj.sync(s)
            files = proj.project_jpeg(s, shifts, consts, W, H)          # warm-up (allocations, tables)
            reps = 5
            t0 = time.perf_counter()
            for _ in range(reps):
                files = proj.project_jpeg(s, shifts, consts, W, H)
            gpu_s = (time.perf_counter() - t0) / reps
            t0 = time.perf_counter()
            for _ in range(reps):
                proj.project(s, shifts, consts, W, H, out=views)
                proj.sync(s)
            pix_s = (time.perf_counter() - t0) / reps
        flat = views.reshape(-1, H, W, 3)
        same = all(f == cv2.imencode(".jpg", v)[1].tobytes() for f, v in zip(files, flat))
        workers = max(1, int((os.cpu_count() or 1) * 0.9))
        with ThreadPoolExecutor(workers) as ex:
            list(ex.map(lambda v: cv2.imencode(".jpg", v)[1], flat))
            t0 = time.perf_counter()
            for _ in range(3):
                list(ex.map(lambda v: cv2.imencode(".jpg", v)[1], flat))
            cpu_s = (time.perf_counter() - t0) / 3
        t0 = time.perf_counter()
        cv2.imencode(".jpg", flat[0])
        cpu1_s = time.perf_counter() - t0
        mpix = flat.shape[0] * W * H / 1e6
        print(json.dumps({
            "panorama": kind, "views": int(flat.shape[0]), "size": [W, H], "byte_identical_to_cv2": bool(same),
            "file_bytes_total": int(sum(len(f) for f in files)), "raw_bytes_total": int(flat.nbytes),
            "gpu_project_encode_readback_ms": gpu_s * 1e3, "gpu_project_readback_pixels_ms": pix_s * 1e3,
            "gpu_mpix_s": mpix / gpu_s,
            "cpu_imencode_threads": workers, "cpu_imencode_ms": cpu_s * 1e3, "cpu_mpix_s": mpix / cpu_s,
            "cpu_single_thread_ms_per_view": cpu1_s * 1e3,
        }), flush=True)
    # ---- files end to end: host panorama in -> 12 JPEG files (bytes) out, pipelined over 4 slots ----
    n_img, n_slots = 16, 4
    pano = synth.smooth(bench.WP, bench.HP, 1)
    pin = pkg.PinnedBuffer((bench.HP, bench.WP, 3))
    pin.array[...] = pano
    workers = max(1, int((os.cpu_count() or 1) * 0.9))

    def gpu_files(i):
        with proj.slots(1) as (s,):
            return proj.process_image_jpeg(s, pin.array, shifts, consts, W, H)

    with ThreadPoolExecutor(n_slots) as ex:
        list(ex.map(gpu_files, range(n_slots)))                          # warm-up
        t0 = time.perf_counter()
        res = list(ex.map(gpu_files, range(n_img)))
        gpu_s = time.perf_counter() - t0
    # the same files with the pixels read back and encoded by cv2 on the host cores (what the png path does)
    outs = [pkg.PinnedBuffer((len(bench.YAWS), len(bench.PITCHES), H, W, 3)) for _ in range(n_slots)]

    def cpu_files(i):
        with proj.slots(1) as (s,):
            o = outs[s % n_slots].array
            proj.process_image(s, pin.array, shifts, consts, W, H, o)
            proj.sync(s)
            return list(enc.map(lambda v: cv2.imencode(".jpg", v)[1].tobytes(), o.reshape(-1, H, W, 3)))

    with ThreadPoolExecutor(workers) as enc, ThreadPoolExecutor(n_slots) as ex:
        list(ex.map(cpu_files, range(n_slots)))
        t0 = time.perf_counter()
        res2 = list(ex.map(cpu_files, range(n_img)))
        cpu_s = time.perf_counter() - t0
    mpix = n_img * bench.PX_PER_IMAGE / 1e6
    print(json.dumps({
        "files_end_to_end": f"{n_img} panoramas 8192x4096 (pinned host) -> 12 JPEG files each, {n_slots} slots",
        "same_files": bool(res[0] == res2[0]),
        "gpu_encoder_ms_per_image": gpu_s / n_img * 1e3, "gpu_encoder_mpix_s": mpix / gpu_s,
        "host_encoder_ms_per_image": cpu_s / n_img * 1e3, "host_encoder_mpix_s": mpix / cpu_s,
        "host_encoder_threads": workers,
    }), flush=True)
    proj.close()


if __name__ == "__main__":
    main()
