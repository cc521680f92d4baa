"""Resident-throughput sweep over the kernel variants (run on the GPU box).

    python tools/sweep_variants.py [--batch 16] [--steps 5] > gpurun_out/sweep.jsonl
"""
import argparse
import json
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import __graft_entry__ as g  # noqa: E402
from tools import synth_inputs as synth  # noqa: E402

import bench  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=16)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--samplers", type=int, nargs="+", default=[0, 1])
    ap.add_argument("--warp-ws", type=int, nargs="+", default=[32, 8])
    ap.add_argument("--nys", type=int, nargs="+", default=[4])
    ap.add_argument("--nbs", type=int, nargs="+", default=[1, 2, 4])
    ap.add_argument("--mirrors", type=int, nargs="+", default=[0, 1, 2])
    ap.add_argument("--seg-chunks", type=int, nargs="+", default=[4], help="row-segment kernel (mirror 2): chunks per warp")
    ap.add_argument("--streams", type=int, default=1, help="launching streams the resident slots alternate over (1 = the "
                    "bench's single stream; > 1: consecutive images overlap their tails, timed by wall clock around a device sync)")
    ap.add_argument("--tag", default="", help="free-form label copied into every line (e.g. the build variant)")
    ap.add_argument("--fov", type=int, default=None, help="override the FOV (locality experiments)")
    args = ap.parse_args()
    import torch

    if args.fov is not None:
        bench.FOV = args.fov
    g.build()
    pkg = g.load_package()
    L = pkg._lib
    dev = torch.device("cuda", 0)
    proj = pkg.Projector(0, n_slots=args.batch)
    slots = list(range(args.batch))
    for i in slots:  # slot i launches on the stream of slot i % streams
        if i >= args.streams:
            proj.set_stream(i, proj.get_stream(i % args.streams))
    base = synth.noise(bench.WP, bench.HP, 0)
    d_stage = torch.from_numpy(base).to(dev)
    for i in slots:
        st = torch.roll(d_stage, shifts=131 * i, dims=1).contiguous()
        torch.cuda.synchronize()
        proj.upload_device(i, st.data_ptr(), bench.WP, bench.HP, bench.WP * 3)
        proj.sync(i)
    d_out = torch.empty((args.batch, bench.N_VIEWS, bench.H, bench.W, 3), dtype=torch.uint8, device=dev)
    outs = [d_out[i].data_ptr() for i in slots]
    consts = [pkg.pitch_constants(bench.W, bench.FOV, p) for p in bench.PITCHES]
    shifts = [pkg.yaw_table(bench.WP, y)[2] for y in bench.YAWS]
    step = proj.batch_call(slots, shifts, consts, bench.W, bench.H, outs)
    ev0, ev1 = proj.event(), proj.event()
    ref = None
    for sampler in args.samplers:
        for ww in args.warp_ws:
          for ny in args.nys:
           for nb in args.nbs:
            for mirror, segc in [(m, sc) for m in args.mirrors for sc in (args.seg_chunks if m == 2 else [0])]:
                if mirror and (sampler != 1 or nb != 1):
                    continue
                proj.set_option(L.OPT_MIRROR, mirror)
                if mirror == 2:
                    proj.set_option(L.OPT_SEG_CHUNKS, segc)
                proj.set_option(L.OPT_SAMPLER, sampler)
                proj.set_option(L.OPT_WARP_W, ww)
                proj.set_option(L.OPT_YAWS_PER_THREAD, ny)
                proj.set_option(L.OPT_IMAGES_PER_LAUNCH, nb)
                for _ in range(3):
                    step()
                proj.sync(-1)
                if args.streams == 1:
                    proj.record(ev0, 0)
                    for _ in range(args.steps):
                        step()
                    proj.record(ev1, 0)
                    proj.sync(0)
                    ms = proj.elapsed_ms(ev0, ev1) / (args.steps * args.batch)
                else:
                    import time
                    t0 = time.perf_counter()
                    for _ in range(args.steps * 4):
                        step()
                    proj.sync(-1)
                    ms = (time.perf_counter() - t0) * 1e3 / (args.steps * 4 * args.batch)
                torch.cuda.synchronize()
                cur = d_out[0].clone()
                same = True if ref is None else bool(torch.equal(cur, ref))
                ref = cur if ref is None else ref  # every variant must produce the same bytes (bit-exact default mode)
                print(json.dumps({"tag": args.tag, "sampler": sampler, "warp_w": ww, "ny": ny, "nb": nb, "mirror": mirror,
                                  "seg_chunks": segc, "streams": args.streams, "image_us": ms * 1e3,
                                  "gpix_s": bench.PX_PER_IMAGE / (ms * 1e-3) / 1e9,
                                  "roofline_frac": bench.B_ALG_PER_IMAGE / (ms * 1e-3) / 1e9 / bench.read_peaks()[0],
                                  "same_output": same}), flush=True)
    proj.close()


if __name__ == "__main__":
    main()
