#!/usr/bin/env python
"""Headline benchmark: output Mpix/s of the panorama -> plane hot path (8K equirect -> 1920x1080).

    python bench.py [--gpus N] [--steps K] [--warmup W]            # this framework (CUDA)
    python bench.py --impl reference [--steps K] [--warmup W]      # CPU reference arm
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
        --master-port P bench.py --gpus N --steps K --warmup W     # N > 1, one rank per GPU

Workload (BASELINE.json configs[1]/[2]): every image is the README example - one synthetic
8192x4096 panorama, FOV 120, 1920x1080, yaw 0/90/180/270 x pitch 30/60/90 = 12 views.  One
"step" processes a batch of 32 such images per GPU (at 8 GPUs that is exactly configs[2], the
256-image batch sharded by image); per-GPU work is fixed, so scaling is weak and needs no
collective - torch.distributed only provides the barrier and the max-over-ranks of the timing.

value    device-resident throughput: the packed panoramas already sit in HBM, outputs stay in HBM,
         one CUDA-event pair on the launching stream around exactly K steps, max over ranks.
e2e      the same metric through the public API with HOST buffers: every image is uploaded from
         pinned memory (the rows its views can touch: 75 of 100.7 MB), projected and read back
         (74.6 MB) inside the timed region, pipelined over the context's slots (streams).
roofline HBM: algorithmic bytes per launch (3 * sum(W*H) + 3 * N_T = 141,009,384 B for
         one 12-view image) / average launch duration, against MEASURED_PEAKS.json hbm_gbs.
cpu_baseline  oracle/ref_port.py (NumPy + cv2.remap restatement of the reference, same thread
         fan-out) timed on this box's host cores on a bounded sample, N = 1 only.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

METRIC = "output Mpix/s (8K equirect -> 1920x1080)"
UNIT = "Mpix/s"

# the README example (BASELINE configs[1]) ----------------------------------------------------
WP, HP = 8192, 4096
W, H, FOV = 1920, 1080, 120
YAWS, PITCHES = [0, 90, 180, 270], [30, 60, 90]
N_VIEWS = len(YAWS) * len(PITCHES)
PX_PER_IMAGE = N_VIEWS * W * H
N_T_C2 = 22_119_928  # distinct panorama texels with non-zero weight, oracle.fixedpoint.touched_texels
B_ALG_PER_IMAGE = 3 * PX_PER_IMAGE + 3 * N_T_C2  # 141,009,384 bytes = 5.667 B / output px
BATCH = 32  # images per GPU per step (configs[2]: 256 images over 8 GPUs)
H2D_PER_IMAGE = WP * HP * 3
D2H_PER_IMAGE = PX_PER_IMAGE * 3


def workload_config(extra=None):
    cfg = {
        "workload": "configs[2] share per GPU: batch of 32 synthetic 8192x4096 panoramas, each the README "
                    "example (configs[1]): FOV 120, 1920x1080, yaw 0/90/180/270 x pitch 30/60/90 = 12 views",
        "pano": [WP, HP], "out": [W, H], "fov": FOV, "yaws": YAWS, "pitches": PITCHES,
        "batch_images_per_gpu": BATCH, "views_per_image": N_VIEWS, "parallelism": "shard by image, no collective",
        "l2": "inputs larger than L2: each step reads 32 different packed panoramas (4.3 GB) and writes 2.4 GB",
    }
    if extra:
        cfg.update(extra)
    return cfg


def read_peaks():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        try:
            d = json.loads(p.read_text())
            return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def read_traffic():
    """dram bytes per launch of the projection kernel from the committed ncu summary, or None."""
    p = ROOT / "profiles" / "traffic.json"
    if p.exists():
        try:
            return json.loads(p.read_text()).get("project_kernel_dram_bytes_per_launch")
        except Exception:
            return None
    return None


# ---------------------------------------------------------------------------------------------
# clocks sampling (nvidia-smi in a child process, killed by PID)
# ---------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,power.draw,utilization.gpu,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.rows = []
        self.proc = None
        self.thread = None
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "50", "-i", str(index)],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            parts = [x.strip() for x in line.split(",")]
            if len(parts) >= 8:
                self.rows.append(parts)

    def stop(self):
        if self.proc is not None:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=5)
            except Exception:
                self.proc.kill()
        if self.thread is not None:
            self.thread.join(timeout=2)
        sm, smax, reasons, power = [], [], set(), []
        busy = []
        for r in self.rows:
            try:
                c, m = float(r[0]), float(r[1])
            except ValueError:
                continue
            sm.append(c)
            smax.append(m)
            try:
                power.append(float(r[2]))
            except ValueError:
                pass
            try:
                if float(r[3]) > 0:
                    busy.append(c)
            except ValueError:
                pass
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[4:8]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        use = busy if busy else sm
        return {"sm_mhz": statistics.median(use), "sm_max_mhz": max(smax), "reasons": sorted(reasons),
                "samples": len(sm), "samples_busy": len(busy), "power_w_max": max(power) if power else None}


# ---------------------------------------------------------------------------------------------
# CPU reference arm / baseline (oracle port of the reference, timed on the host cores)
# ---------------------------------------------------------------------------------------------
def cpu_reference_run(steps: int, warmup: int, budget_s: float | None = None):
    """Times oracle.ref_port (NumPy + cv2.remap, ThreadPoolExecutor per yaw, like the reference)
    on one README-example image per step with warm map caches (steady state over a directory of
    same-sized images, ref :42-73).  Returns (Mpix/s, ms_per_step, steps_done, info)."""
    import cv2

    from oracle import ref_port
    from tools import synth_inputs as synth

    pano = synth.noise(WP, HP, 0)
    ref_port.clear_caches()
    workers = ref_port.default_workers()
    t0 = time.perf_counter()
    ref_port.process_image_views(pano, YAWS, PITCHES, W, H, FOV, num_workers=workers)  # cold: builds the maps
    cold_s = time.perf_counter() - t0
    for _ in range(max(0, warmup - 1)):
        ref_port.process_image_views(pano, YAWS, PITCHES, W, H, FOV, num_workers=workers)
    done = 0
    t0 = time.perf_counter()
    for _ in range(steps):
        ref_port.process_image_views(pano, YAWS, PITCHES, W, H, FOV, num_workers=workers)
        done += 1
        if budget_s is not None and time.perf_counter() - t0 > budget_s:
            break
    el = time.perf_counter() - t0
    info = {
        "cores": os.cpu_count() or 1,
        "kind": "port",
        "sample": f"{done} README-example images (12 views 8192x4096 -> 1920x1080) with warm map caches after 1 cold "
                  f"image ({cold_s:.2f} s incl. map precompute = {PX_PER_IMAGE / cold_s / 1e6:.1f} Mpix/s cold); "
                  f"ThreadPoolExecutor({workers}) one task per yaw, cv2 threads {cv2.getNumThreads()}, "
                  f"numpy {np.__version__}, cv2 {cv2.__version__}",
        "cold_mpix_s": PX_PER_IMAGE / cold_s / 1e6,
    }
    return done * PX_PER_IMAGE / el / 1e6, el / done * 1e3, done, info


def run_reference_arm(args, rank: int):
    if rank != 0:
        return
    val, ms, done, info = cpu_reference_run(args.steps, args.warmup)
    cb = {"value": val, "unit": UNIT, "cores": info["cores"], "kind": info["kind"], "sample": info["sample"]}
    line = {
        "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": done,
        "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "u8", "data": "synthetic",
        "config": workload_config({"step": "one image (12 views) per step: bounded sample of the 32-image batch"}),
        "cpu_baseline": cb,
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)


# ---------------------------------------------------------------------------------------------
# the CUDA arm
# ---------------------------------------------------------------------------------------------
def run_b200_arm(args, rank: int, local_rank: int, world: int):
    import torch

    import __graft_entry__ as g
    from tools import synth_inputs as synth  # synthetic input generator (no oracle code)

    g.build()
    pkg = g.load_package()
    from p2p_b200 import distrib

    dev = torch.device("cuda", local_rank)
    torch.cuda.set_device(dev)
    ranks = distrib.Ranks("nccl", device=dev)  # no-op when WORLD_SIZE == 1
    barrier, max_over_ranks = ranks.barrier, ranks.max

    n_e2e_slots = 4
    proj = pkg.Projector(local_rank, n_slots=BATCH + n_e2e_slots)
    L = pkg._lib
    if args.sampler is not None:
        proj.set_option(L.OPT_SAMPLER, args.sampler)
    if args.warp_w is not None:
        proj.set_option(L.OPT_WARP_W, args.warp_w)
    if args.ny is not None:
        proj.set_option(L.OPT_YAWS_PER_THREAD, args.ny)
    if args.nb is not None:
        proj.set_option(L.OPT_IMAGES_PER_LAUNCH, args.nb)

    consts = [pkg.pitch_constants(W, FOV, p) for p in PITCHES]
    shifts = [pkg.yaw_table(WP, y)[2] for y in YAWS]
    assert all(s is not None for s in shifts)

    # ---- resident inputs: BATCH packed panoramas per GPU (seeds follow configs[2]: rank-major) ----
    res_slots = list(range(BATCH))
    proj.share_stream(res_slots, 0)  # one launching stream so a single event pair brackets the steps
    n_distinct = min(BATCH, args.distinct)
    host = [synth.noise(WP, HP, rank * BATCH + i) for i in range(n_distinct)]
    d_stage = torch.empty((HP, WP, 3), dtype=torch.uint8, device=dev)
    for i in res_slots:
        # panorama i is noise(seed) rolled by a different column offset when fewer seeds than slots
        src = host[i % n_distinct]
        d_stage.copy_(torch.from_numpy(src))
        if i >= n_distinct:
            d_stage.copy_(torch.roll(d_stage, shifts=37 * i, dims=1))
        torch.cuda.synchronize()
        proj.upload_device(i, d_stage.data_ptr(), WP, HP, WP * 3)
        proj.sync(i)
    d_out = torch.empty((BATCH, N_VIEWS, H, W, 3), dtype=torch.uint8, device=dev)
    out_ptrs = [d_out[i].data_ptr() for i in range(BATCH)]

    resident_step = proj.batch_call(res_slots, shifts, consts, W, H, out_ptrs, on_device=True)

    for _ in range(max(3, args.warmup)):
        resident_step()
    proj.sync(0)
    ev0, ev1 = proj.event(), proj.event()

    sampler = ClockSampler(local_rank) if rank == 0 else None
    launches0 = proj.launches
    barrier()
    torch.cuda.synchronize()
    proj.record(ev0, 0)
    for _ in range(args.steps):
        resident_step()
    proj.record(ev1, 0)
    proj.sync(0)
    torch.cuda.synchronize()
    barrier()
    ms_total = max_over_ranks(proj.elapsed_ms(ev0, ev1))
    launches = proj.launches - launches0
    ms_per_step = ms_total / args.steps
    value = world * args.steps * BATCH * PX_PER_IMAGE / (ms_total * 1e-3) / 1e6
    launch_ms = ms_total / (args.steps * BATCH)

    # ---- end to end through the public API with host buffers -------------------------------
    e2e_slots = list(range(BATCH, BATCH + n_e2e_slots))
    n_host = 3
    pin_in = [pkg.PinnedBuffer((HP, WP, 3)) for _ in range(n_host)]
    for k, b in enumerate(pin_in):
        b.array[...] = host[k % n_distinct]
    pin_out = [pkg.PinnedBuffer((len(YAWS), len(PITCHES), H, W, 3)) for _ in range(n_e2e_slots)]

    def e2e_step():
        for i in range(BATCH):
            s = e2e_slots[i % n_e2e_slots]
            proj.sync(s)  # the slot's previous image (and its readback) is complete
            proj.process_image(s, pin_in[i % n_host].array, shifts, consts, W, H, pin_out[i % n_e2e_slots].array)

    # the library transfers only the panorama rows these views can touch (p2p_view_row_range)
    row_first, row_last = proj.view_row_range(consts, W, H, WP, HP)
    h2d_per_image = (row_last - row_first + 1) * WP * 3 if proj.get_option(L.OPT_PARTIAL_UPLOAD) else H2D_PER_IMAGE
    e2e_steps = max(1, min(args.steps, args.e2e_steps))
    for _ in range(2):
        e2e_step()
    proj.sync(-1)
    barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        e2e_step()
    proj.sync(-1)
    torch.cuda.synchronize()
    e2e_s = max_over_ranks(time.perf_counter() - t0)
    barrier()
    e2e_value = world * e2e_steps * BATCH * PX_PER_IMAGE / e2e_s / 1e6
    clocks = sampler.stop() if sampler is not None else None

    # light self-check of the e2e result buffer (not timed): last image equals a device-resident render
    check = proj.project_image(pin_in[(BATCH - 1) % n_host].array, YAWS, PITCHES, W, H, FOV)
    assert np.array_equal(check, pin_out[(BATCH - 1) % n_e2e_slots].array), "e2e readback differs from resident render"

    extras = None
    if rank == 0 and world == 1 and not args.no_extras:
        extras = files_extras(pkg, proj, synth.smooth(WP, HP, 0), shifts, consts)

    cpu_baseline = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        v, _ms, done, info = cpu_reference_run(10_000, 1, budget_s=args.cpu_seconds)
        cpu_baseline = {"value": v, "unit": UNIT, "cores": info["cores"], "kind": info["kind"], "sample": info["sample"]}
    elif rank == 0:
        cpu_baseline = {"value": None, "unit": UNIT, "cores": os.cpu_count() or 1, "kind": "port",
                        "sample": "not run (N > 1 or --no-cpu-baseline); see the N = 1 line"}

    if rank == 0:
        peak, peak_src = read_peaks()
        achieved = B_ALG_PER_IMAGE / (launch_ms * 1e-3) / 1e9
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": max(3, args.warmup), "ms_per_step": ms_per_step, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
            "config": workload_config({
                "sampler": proj.get_option(L.OPT_SAMPLER), "warp_w": proj.get_option(L.OPT_WARP_W),
                "yaws_per_thread": proj.get_option(L.OPT_YAWS_PER_THREAD),
                "images_per_launch": proj.get_option(L.OPT_IMAGES_PER_LAUNCH),
                "mirror_pairs": proj.get_option(L.OPT_MIRROR),
                "distinct_seeds_per_gpu": n_distinct,
            }),
            "roofline": {
                "bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": read_traffic(), "peak_source": peak_src,
                "algorithmic_bytes_per_launch": B_ALG_PER_IMAGE, "launch_ms": launch_ms,
                "kernel": ("p2p::project_mirror_kernel<4>" if (proj.get_option(L.OPT_MIRROR) and proj.get_option(L.OPT_SAMPLER) == 1
                                                              and proj.get_option(L.OPT_IMAGES_PER_LAUNCH) == 1)
                           else "p2p::project_kernel") + " (one launch = one image = 12 views)",
            },
            "cpu_baseline": cpu_baseline,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": BATCH * h2d_per_image,
                    "d2h_bytes_per_step": BATCH * D2H_PER_IMAGE, "steps": e2e_steps,
                    "ms_per_step": e2e_s / e2e_steps * 1e3, "pipeline_slots": n_e2e_slots,
                    "h2d_rows": [row_first, row_last],
                    "h2d_note": f"rows {row_first}..{row_last} of {HP}: the only panorama rows the 12 views read "
                                f"({h2d_per_image} of {H2D_PER_IMAGE} bytes per image)"},
            "gpu_launches": launches * world,
            "clocks": clocks,
        }
        if extras is not None:
            line["extras"] = extras
        emit(line)
    for b in pin_in + pin_out:
        b.free()
    proj.close()
    ranks.close()


def files_extras(pkg, proj, pano, shifts, consts):
    """Side measurement (not the headline metric): the codec rows either side of the path.  One 8192x4096
    JPEG file in host memory -> the 12 views as jpg / png files in (page-locked) host memory, (a) decoded, projected and
    encoded on the GPU, 4 images in flight on 4 host threads; (b) the reference's flow on the host cores: cv2.imdecode -> oracle port of
    the projection -> cv2.imencode per view.  Same bytes out (checked)."""
    import cv2
    from concurrent.futures import ThreadPoolExecutor

    from oracle import ref_port

    data = cv2.imencode(".jpg", pano)[1].tobytes()
    n_img, n_thr = 8, 4
    out = {}
    for fmt in ("jpg", "png"):
        def gpu_one(keep):
            # the files land in the slot's page-locked host buffer; the front end writes them to disk from there
            # (copy=False), the identity check below takes a copy
            with proj.slots(1) as (s,):
                proj.upload_jpeg(s, data)
                if fmt == "jpg":
                    files = proj.project_jpeg(s, shifts, consts, W, H, copy=False)
                else:
                    files = proj.process_image_png(s, None, shifts, consts, W, H, want_pixels=False, copy=False)[0]
                return [bytes(f) for f in files] if keep else sum(len(f) for f in files)

        with ThreadPoolExecutor(n_thr) as ex:
            first = list(ex.map(gpu_one, [True] * n_thr))[0]
            t0 = time.perf_counter()
            list(ex.map(gpu_one, [False] * n_img))
            gpu_s = (time.perf_counter() - t0) / n_img
        ref_port.clear_caches()
        ref_files = None
        cpu_s = []
        for _ in range(2):  # second pass = warm map caches
            t0 = time.perf_counter()
            img = cv2.imdecode(np.frombuffer(data, np.uint8), cv2.IMREAD_COLOR)
            views = ref_port.process_image_views(img, YAWS, PITCHES, W, H, FOV)
            ref_files = [cv2.imencode("." + fmt, v)[1].tobytes() for per_yaw in views for v in per_yaw]
            cpu_s.append(time.perf_counter() - t0)
        ref_port.clear_caches()
        out[fmt + "_files"] = {
            "what": f"8192x4096 JPEG bytes (host) -> 12 x 1920x1080 {fmt} files (host); decode (Huffman stage included), "
                    f"projection and encode on the GPU, 4 images in flight; CPU = cv2.imdecode + oracle port + cv2.imencode, "
                    f"warm maps",
            "byte_identical_to_cpu_flow": bool(first == ref_files),
            "gpu_ms_per_image": gpu_s * 1e3, "gpu_mpix_s": PX_PER_IMAGE / gpu_s / 1e6,
            "cpu_ms_per_image": min(cpu_s) * 1e3, "cpu_mpix_s": PX_PER_IMAGE / min(cpu_s) / 1e6,
            "input_file_bytes": len(data), "output_file_bytes": sum(len(f) for f in first)}
    return out


_REAL_STDOUT = None


def emit(line: dict):
    """Print the one JSON line on the real stdout (library chatter such as NCCL's version banner has
    been redirected to stderr by ``main``)."""
    data = (json.dumps(line) + "\n").encode()
    if _REAL_STDOUT is None:
        sys.stdout.write(data.decode())
        sys.stdout.flush()
    else:
        os.write(_REAL_STDOUT, data)


def main():
    global _REAL_STDOUT
    # stdout carries exactly one JSON line: everything else any library prints to fd 1 goes to stderr
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", choices=["b200", "reference"], default="b200")
    ap.add_argument("--sampler", type=int, default=None)
    ap.add_argument("--warp-w", type=int, default=None)
    ap.add_argument("--ny", type=int, default=None)
    ap.add_argument("--nb", type=int, default=None, help="resident panoramas per launch (1, 2, 4)")
    ap.add_argument("--distinct", type=int, default=8, help="distinct noise seeds per GPU (others are rolled copies)")
    ap.add_argument("--e2e-steps", type=int, default=5, help="cap on end-to-end steps (each moves 5.6 GB over PCIe)")
    ap.add_argument("--cpu-seconds", type=float, default=12.0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="skip the files-to-files (JPEG in, JPEG out) side measurement")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        run_reference_arm(args, rank)
        return
    if world != args.gpus and world == 1 and args.gpus > 1:
        # launched without torchrun: re-exec under torch.distributed.run, one rank per GPU
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={args.gpus}",
               "--master-addr", "127.0.0.1", "--master-port", str(29500 + os.getpid() % 1000), __file__, *sys.argv[1:]]
        os.dup2(_REAL_STDOUT, 1)  # the torchrun children write their own single line
        raise SystemExit(subprocess.call(cmd))
    run_b200_arm(args, rank, local_rank, world)


if __name__ == "__main__":
    main()
