#!/usr/bin/env python
"""Strong scaling of ONE image split over the GPUs of a box: BASELINE configs C2, C4, C5.

    python tools/time_split.py [--configs c2 c4 c5] [--devices 1 2 4 8] [--reps 10] > gpurun_out/r2_split.jsonl

One process drives one context per device.  Per repetition (all calls asynchronous, page-locked host buffers):
    the panorama reaches every device - mode "replicate": uploaded to device 0 over PCIe, then cudaMemcpyPeerAsync to the
    others; mode "scatter": every device uploads 1 / n of the rows over its own PCIe link and fetches the other pieces
    from its peers (an NVLink all-gather of peer copies) -
    -> every device renders its band of output rows of ALL views (p2p_project_view_list) -> bands land in one host array.
Three timings per device count: ``e2e_ms`` the whole sequence, ``resident_ms`` panoramas already replicated (project +
readback), ``device_ms`` outputs left in HBM (kernel only, wall clock around sync of all devices).  The N = 1 result is the
reference every split result is compared with (``identical``).
"""
from __future__ import annotations

import argparse
import json
import sys
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import __graft_entry__ as g  # noqa: E402
from tools import synth_inputs as synth  # noqa: E402

CONFIGS = {
    # name: (Wp, Hp, W, H, fov, [(yaw, pitch), ...])
    "c2": (8192, 4096, 1920, 1080, 120, [(y, p) for y in (0, 90, 180, 270) for p in (30, 60, 90)]),
    "c4": (16384, 8192, 3840, 2160, 100, [(y, p) for y in (0, 90, 180, 270) for p in (30, 60, 90)]),
    "c5": (8192, 4096, 2048, 2048, 90, [(0, 90), (90, 90), (180, 90), (270, 90), (0, 0), (0, 180)]),
}


def median_ms(fn, reps):
    ts = []
    for _ in range(reps):
        t0 = time.perf_counter()
        fn()
        ts.append((time.perf_counter() - t0) * 1e3)
    return float(np.median(ts)), float(min(ts))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--configs", nargs="+", default=["c2", "c4", "c5"])
    ap.add_argument("--devices", type=int, nargs="+", default=[1, 2, 4, 8])
    ap.add_argument("--reps", type=int, default=10)
    args = ap.parse_args()
    import torch

    g.build()
    pkg = g.load_package()
    from p2p_b200 import shard

    n_dev = pkg._lib.load().p2p_device_count()
    counts = [n for n in args.devices if n <= n_dev]
    projs = [pkg.Projector(d, n_slots=2) for d in range(max(counts))]
    for name in args.configs:
        Wp, Hp, W, H, fov, views = CONFIGS[name]
        pano = synth.noise(Wp, Hp, 0)
        pin_in = pkg.PinnedBuffer(pano.shape)
        pin_in.array[...] = pano
        pin_out = pkg.PinnedBuffer((len(views), H, W, 3))
        shifts = [pkg.yaw_table(Wp, y)[2] for y, _ in views]
        consts = [pkg.pitch_constants(W, fov, p) for _, p in views]
        ref = None
        for n in counts:
            ps = projs[:n]
            bands = [shard.shard_rows(H, r, n) for r in range(n)]
            d_outs = [torch.empty((len(views), H, W, 3), dtype=torch.uint8, device=f"cuda:{r}") for r in range(n)]

            def replicate(mode="scatter"):
                if mode == "scatter" and n > 1:
                    pkg.scatter_upload(ps, [0] * n, pin_in.array)
                else:
                    ps[0].upload(0, pin_in.array)
                    for p in ps[1:]:
                        p.copy_pano_from(0, ps[0], 0)

            host_calls = [p.project_list_call(0, shifts, consts, W, H, rows=bands[r], out=pin_out.array)
                          for r, p in enumerate(ps) if bands[r][0] < bands[r][1]]
            dev_calls = [p.project_list_call(0, shifts, consts, W, H, rows=bands[r], out_device_ptr=d_outs[r].data_ptr())
                         for r, p in enumerate(ps) if bands[r][0] < bands[r][1]]

            def project(host=True):
                for f in (host_calls if host else dev_calls):
                    f()

            def sync():
                for p in ps:
                    p.sync(0)

            def e2e():
                replicate()
                project()
                sync()

            def resident():
                project()
                sync()

            def device_only():
                project(host=False)
                sync()

            pin_out.array[...] = 0
            for _ in range(3):
                e2e()
            got = pin_out.array.copy()
            ref = got if ref is None else ref
            e2e_ms = median_ms(e2e, args.reps)
            res_ms = median_ms(resident, args.reps)
            dev_ms = median_ms(device_only, args.reps)
            rep_ms = median_ms(lambda: (replicate("replicate"), sync()), args.reps)
            sca_ms = median_ms(lambda: (replicate("scatter"), sync()), args.reps)
            px = len(views) * W * H
            print(json.dumps({
                "config": name, "pano": [Wp, Hp], "out": [W, H], "views": len(views), "n_gpus": n,
                "bands": bands, "identical": bool(np.array_equal(got, ref)),
                "e2e_ms": e2e_ms[0], "e2e_ms_min": e2e_ms[1], "resident_ms": res_ms[0], "device_ms": dev_ms[0],
                "device_ms_min": dev_ms[1], "upload_replicate_ms": rep_ms[0], "upload_scatter_allgather_ms": sca_ms[0],
                "e2e_mpix_s": px / (e2e_ms[0] * 1e-3) / 1e6, "device_mpix_s": px / (dev_ms[0] * 1e-3) / 1e6,
            }), flush=True)
            del d_outs
        pin_in.free()
        pin_out.free()
    for p in projs:
        p.close()


if __name__ == "__main__":
    main()
