"""JPEG panorama decode: cv2.imdecode (libjpeg-turbo, what cv2.imread does at ref :244) + upload against the device
decoder (host Huffman stage + IDCT / upsampling / colour kernels), same files, same pixels (run on the GPU box).

    python tools/bench_jpegdec.py > gpurun_out/jpegdec_bench.jsonl
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


def main():
    import cv2

    g.build()
    pkg = g.load_package()
    proj = pkg.Projector(0, n_slots=8)
    Wp, Hp = 8192, 4096
    rng = np.random.default_rng(0)
    smooth = synth.smooth(Wp, Hp, 0)
    textured = np.clip(smooth.astype(np.int16) + rng.integers(-12, 13, smooth.shape, dtype=np.int16), 0, 255).astype(np.uint8)
    cases = {"smooth": (smooth, 95), "textured (smooth + noise of amplitude 12)": (textured, 92), "white noise": (synth.noise(Wp, Hp, 0), 95)}
    cases["smooth, progressive (scans decoded on host threads, IDCT / colour on the device)"] = (smooth, -95)
    for name, (img, q) in cases.items():
        data = cv2.imencode(".jpg", img, [cv2.IMWRITE_JPEG_QUALITY, abs(q), cv2.IMWRITE_JPEG_PROGRESSIVE, int(q < 0)])[1].tobytes()
        arr = np.frombuffer(data, np.uint8)
        ref = cv2.imdecode(arr, cv2.IMREAD_COLOR)
        with proj.slots(1) as (s,):
            proj.upload_jpeg(s, data)
            proj.sync(s)
            same = bool(np.array_equal(proj.download_pano(s, Wp, Hp), ref))
            t_dev = []
            for _ in range(3):
                t0 = time.perf_counter()
                proj.upload_jpeg(s, data)
                proj.sync(s)
                t_dev.append(time.perf_counter() - t0)
            t_cv = []
            for _ in range(3):
                t0 = time.perf_counter()
                pix = cv2.imdecode(arr, cv2.IMREAD_COLOR)
                proj.upload(s, pix)
                proj.sync(s)
                t_cv.append(time.perf_counter() - t0)
        # throughput with several host threads (one image per thread and slot)
        n_img, n_thr = 16, min(7, max(1, (os.cpu_count() or 1) // 2))

        def dev_one(_):
            with proj.slots(1) as (s,):
                proj.upload_jpeg(s, data)
                proj.sync(s)

        def cv_one(_):
            pix = cv2.imdecode(arr, cv2.IMREAD_COLOR)
            with proj.slots(1) as (s,):
                proj.upload(s, pix)
                proj.sync(s)

        with ThreadPoolExecutor(n_thr) as ex:
            list(ex.map(dev_one, range(n_thr)))
            t0 = time.perf_counter()
            list(ex.map(dev_one, range(n_img)))
            thr_dev = (time.perf_counter() - t0) / n_img
            list(ex.map(cv_one, range(n_thr)))
            t0 = time.perf_counter()
            list(ex.map(cv_one, range(n_img)))
            thr_cv = (time.perf_counter() - t0) / n_img
        print(json.dumps({"panorama": name, "quality": q, "file_MB": len(data) / 1e6, "same_pixels_as_cv2": same,
                          "device_decoder_ms": min(t_dev) * 1e3, "cv2_imdecode_plus_upload_ms": min(t_cv) * 1e3,
                          "speedup_one_thread": min(t_cv) / min(t_dev),
                          "threads": n_thr, "device_decoder_ms_per_image_threads": thr_dev * 1e3,
                          "cv2_ms_per_image_threads": thr_cv * 1e3}), flush=True)
    proj.close()


if __name__ == "__main__":
    main()
