"""PNG panorama decode: cv2.imdecode (libpng + zlib, what cv2.imread does at ref :244) + upload against the device
decoder (parallel inflate + unfilter kernels, csrc/p2p_pngdec.cuh), same files, same pixels (run on the GPU box).

    python tools/bench_pngdec.py [--small] > gpurun_out/pngdec_bench.jsonl
"""
import json
import sys
import time
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
    proj = pkg.Projector(0, n_slots=2)
    Wp, Hp = (2048, 1024) if "--small" in sys.argv else (8192, 4096)
    reps = 1 if "--once" in sys.argv else 4   # --once: a single decode per file (under ncu)
    rng = np.random.default_rng(0)
    smooth = synth.smooth(Wp, Hp, 0)
    textured = np.clip(smooth.astype(np.int16) + rng.integers(-12, 13, smooth.shape, dtype=np.int16), 0, 255).astype(np.uint8)
    P = cv2
    cases = [
        ("textured, cv2.imwrite defaults (the reference's own output format: filter Sub, Z_RLE, level 1)", textured, []),
        ("textured, zlib level 6 default strategy", textured, [P.IMWRITE_PNG_STRATEGY, P.IMWRITE_PNG_STRATEGY_DEFAULT, P.IMWRITE_PNG_COMPRESSION, 6]),
        ("smooth, cv2.imwrite defaults", smooth, []),
    ]
    for name, img, params in cases:
        data = cv2.imencode(".png", img, params)[1].tobytes()
        arr = np.frombuffer(data, np.uint8)
        ref = cv2.imdecode(arr, cv2.IMREAD_COLOR)
        with proj.slots(1) as (s,):
            try:
                proj.upload_png(s, data)
            except pkg.P2PError as e:
                print(json.dumps({"panorama": name, "file_MB": len(data) / 1e6, "declined": str(e)}), flush=True)
                continue
            proj.sync(s)
            same = bool(np.array_equal(proj.download_pano(s, Wp, Hp), ref))
            t_dev = []
            for _ in range(reps):
                t0 = time.perf_counter()
                proj.upload_png(s, data)
                proj.sync(s)
                t_dev.append(time.perf_counter() - t0)
            t_cv = []
            for _ in range(1 if reps == 1 else 2):
                t0 = time.perf_counter()
                pix = cv2.imdecode(arr, cv2.IMREAD_COLOR)
                proj.upload(s, pix)
                proj.sync(s)
                t_cv.append(time.perf_counter() - t0)
        print(json.dumps({"panorama": name, "size": [Wp, Hp], "file_MB": len(data) / 1e6, "same_pixels_as_cv2": same,
                          "device_decoder_ms": min(t_dev) * 1e3, "device_decoder_ms_all": [round(t * 1e3, 2) for t in t_dev],
                          "cv2_imdecode_plus_upload_ms": min(t_cv) * 1e3, "speedup": min(t_cv) / min(t_dev)}), flush=True)
    proj.close()


if __name__ == "__main__":
    main()
