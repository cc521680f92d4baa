"""Device-resident timing of the BASELINE parity configs (C1, C2, C4, C5) with CUDA events.
    python tools/time_configs.py > gpurun_out/configs.jsonl
"""
import json
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import __graft_entry__ as g  # noqa: E402
from tools import synth_inputs as synth  # noqa: E402

CONFIGS = {
    "C1": dict(Wp=2048, Hp=1024, W=640, H=480, fov=90, views=[([0], [90])]),
    "C2": dict(Wp=8192, Hp=4096, W=1920, H=1080, fov=120, views=[([0, 90, 180, 270], [30, 60, 90])]),
    "C4": dict(Wp=16384, Hp=8192, W=3840, H=2160, fov=100, views=[([0, 90, 180, 270], [30, 60, 90])]),
    "C5": dict(Wp=8192, Hp=4096, W=2048, H=2048, fov=90,
               views=[([0, 90, 180, 270], [90]), ([0], [0]), ([0], [180])]),
}


def main():
    import torch

    g.build()
    pkg = g.load_package()
    proj = pkg.Projector(0, n_slots=2)
    ev0, ev1 = proj.event(), proj.event()
    for name, c in CONFIGS.items():
        pano = synth.noise(c["Wp"], c["Hp"], 0)
        proj.upload(0, pano)
        proj.sync(0)
        n_views = sum(len(y) * len(p) for y, p in c["views"])
        d_out = torch.empty((n_views, c["H"], c["W"], 3), dtype=torch.uint8, device="cuda:0")
        calls = []
        off = 0
        for yaws, pitches in c["views"]:
            shifts = [pkg.yaw_table(c["Wp"], y)[2] for y in yaws]
            consts = [pkg.pitch_constants(c["W"], c["fov"], p) for p in pitches]
            calls.append(proj.batch_call([0], shifts, consts, c["W"], c["H"], [d_out[off].data_ptr()]))
            off += len(yaws) * len(pitches)
        reps = 20
        for _ in range(3):
            for f in calls:
                f()
        proj.sync(0)
        proj.record(ev0, 0)
        for _ in range(reps):
            for f in calls:
                f()
        proj.record(ev1, 0)
        proj.sync(0)
        us = proj.elapsed_ms(ev0, ev1) / reps * 1e3
        px = n_views * c["W"] * c["H"]
        print(json.dumps({"config": name, "views": n_views, "launches_per_image": len(calls), "image_us": us,
                          "gpix_s": px / us / 1e3,
                          "note": "same panorama every repetition (C1/C5 L2-warm; C4 panorama is 537 MB > L2)"}), flush=True)
    proj.close()


if __name__ == "__main__":
    main()
