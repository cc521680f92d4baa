"""The bench's files flow (8192x4096 JPEG bytes -> 12 JPEG / PNG files, decode + projection + encode on the GPU) on ONE GPU,
swept over the number of images in flight and the host cores the process may use.  `--cores 4` emulates the share of the
box's 32 vCPUs one of eight ranks gets (sched_setaffinity before any thread starts), so the N = 8 host pressure can be
studied on a one-GPU box.

    python tools/time_files_flow.py --cores 0 4 --threads 4 8 16 > gpurun_out/files_flow.jsonl
"""
import argparse
import json
import os
import sys
import time
from concurrent.futures import ThreadPoolExecutor
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import __graft_entry__ as g  # noqa: E402
from tools import synth_inputs as synth  # noqa: E402

W, H, FOV = 1920, 1080, 100
YAWS, PITCHES = [0, 90, 180, 270], [30, 60, 90]
WP, HP = 8192, 4096


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--cores", type=int, nargs="+", default=[0], help="host cores the process may use (0 = all)")
    ap.add_argument("--threads", type=int, nargs="+", default=[4, 8])
    ap.add_argument("--images", type=int, default=64)
    ap.add_argument("--formats", nargs="+", default=["jpg"])
    ap.add_argument("--pano", choices=["smooth", "noise"], default="smooth")
    ap.add_argument("--huffman", type=int, nargs="+", default=[1], help="P2P_OPT_GPU_HUFFMAN values to sweep")
    ap.add_argument("--wait", type=int, nargs="+", default=[0], help="P2P_OPT_HOST_WAIT values to sweep (1 = sleeping waits)")
    ap.add_argument("--t0", type=float, default=0.0, help="epoch seconds at which configuration 0 starts (several processes, one "
                    "per GPU, started together: every configuration then runs on all of them at the same time)")
    ap.add_argument("--slot-seconds", type=float, default=4.0)
    args = ap.parse_args()
    import cv2

    g.build()
    pkg = g.load_package()
    L = pkg._lib
    all_cores = sorted(os.sched_getaffinity(0))
    proj = pkg.Projector(0, n_slots=max(args.threads) + 1)
    consts = [pkg.pitch_constants(W, FOV, p) for p in PITCHES]
    shifts = [pkg.yaw_table(WP, y)[2] for y in YAWS]
    pano = synth.smooth(WP, HP, 0) if args.pano == "smooth" else synth.noise(WP, HP, 0)
    data = cv2.imencode(".jpg", pano)[1].tobytes()
    k_cfg = 0
    for fmt in args.formats:

        def one(_):
            with proj.slots(1) as (s,):
                proj.upload_jpeg(s, data)
                if fmt == "jpg":
                    files = proj.project_jpeg(s, shifts, consts, W, H, copy=False)
                else:
                    files = proj.process_image_png(s, None, shifts, consts, W, H, want_pixels=False, copy=False)[0]
                return sum(len(f) for f in files)

        for huff in args.huffman:
            proj.set_option(L.OPT_GPU_HUFFMAN, huff)
            for wait, nc in [(w, c) for w in args.wait for c in args.cores]:
                proj.set_option(L.OPT_HOST_WAIT, wait)
                cores = all_cores if nc <= 0 else all_cores[:nc]
                os.sched_setaffinity(0, cores)
                for n_thr in args.threads:
                    with ThreadPoolExecutor(n_thr) as ex:
                        list(ex.map(one, range(n_thr)))
                        if args.t0:
                            time.sleep(max(0.0, args.t0 + k_cfg * args.slot_seconds - time.time()))
                        k_cfg += 1
                        c0 = time.process_time()
                        t0 = time.perf_counter()
                        list(ex.map(one, range(args.images)))
                        sec = time.perf_counter() - t0
                        cpu = time.process_time() - c0
                    print(json.dumps({"format": f"jpg -> {fmt}", "pano": args.pano, "file_MB": len(data) / 1e6, "gpu_huffman": huff, "host_wait": wait,
                                      "cores": len(cores), "threads": n_thr, "images": args.images,
                                      "ms_per_image": sec / args.images * 1e3,
                                      "cpu_ms_per_image": cpu / args.images * 1e3,
                                      "mpix_s": args.images * len(YAWS) * len(PITCHES) * W * H / sec / 1e6}), flush=True)
                os.sched_setaffinity(0, all_cores)


if __name__ == "__main__":
    main()
