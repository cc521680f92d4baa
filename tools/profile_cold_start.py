"""Cold-start breakdown of the JPEG -> files flow (what a one-shot CLI run pays before the steady state):
    python tools/profile_cold_start.py"""
import importlib
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

t_start = time.perf_counter()
import numpy as np  # noqa: E402
import cv2  # noqa: E402

t_imports = time.perf_counter()
from tools import synth_inputs as synth  # noqa: E402


def main():
    out = {"import_numpy_cv2_s": round(t_imports - t_start, 3)}
    t0 = time.perf_counter()
    pkg = importlib.import_module("360-to-planer-images_b200")
    out["import_package_s"] = round(time.perf_counter() - t0, 3)
    W, H, FOV = 1920, 1080, 120
    yaws, pitches = [0, 90, 180, 270], [30, 60, 90]
    pano = synth.smooth(8192, 4096, 1)
    data = cv2.imencode(".jpg", pano)[1].tobytes()
    t0 = time.perf_counter()
    proj = pkg.Projector(0, n_slots=16)
    out["create_context_s"] = round(time.perf_counter() - t0, 3)
    consts = [pkg.pitch_constants(W, FOV, p) for p in pitches]
    shifts = np.array([pkg.yaw_table(8192, y)[2] for y in yaws], np.int32)
    steps = []
    for k in range(3):       # three different slots, then the first one again
        with proj.slots(1) as (s,):
            s = k            # force a fresh slot
            t0 = time.perf_counter(); proj.upload_jpeg(s, data); t1 = time.perf_counter()
            proj.project_jpeg(s, shifts, consts, W, H, copy=False); t2 = time.perf_counter()
            proj.process_image_png(s, None, shifts, consts, W, H, want_pixels=False, copy=False); t3 = time.perf_counter()
            proj.upload_jpeg(s, data); t4 = time.perf_counter()
            proj.project_jpeg(s, shifts, consts, W, H, copy=False); t5 = time.perf_counter()
            proj.process_image_png(s, None, shifts, consts, W, H, want_pixels=False, copy=False); t6 = time.perf_counter()
        steps.append({"slot": k, "first_upload_jpeg_ms": round((t1 - t0) * 1e3, 1), "first_project_jpeg_ms": round((t2 - t1) * 1e3, 1),
                      "first_png_ms": round((t3 - t2) * 1e3, 1), "warm_upload_jpeg_ms": round((t4 - t3) * 1e3, 1),
                      "warm_project_jpeg_ms": round((t5 - t4) * 1e3, 1), "warm_png_ms": round((t6 - t5) * 1e3, 1)})
    out["slots"] = steps
    for mb in (4, 32, 100):
        t0 = time.perf_counter()
        pb = pkg.PinnedBuffer((mb << 20,))
        t1 = time.perf_counter()
        pb.free()
        out[f"pinned_alloc_{mb}MB_ms"] = round((t1 - t0) * 1e3, 1)
    print(json.dumps(out), flush=True)
    proj.close()


if __name__ == "__main__":
    main()
