#!/usr/bin/env python
"""Host <-> device transfer ceiling of the end-to-end step, per rank count.

    python tools/pcie_ceiling.py                                   # N = 1
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        tools/pcie_ceiling.py                                      # N ranks, one per GPU

Bare ``cudaMemcpyAsync`` from / to page-locked host memory with the byte counts of one ``bench.py`` end-to-end step
(32 images per rank: 75.1 MB up - the panorama rows the 12 README views touch - and 74.6 MB down per image), no
kernels, no library code of this repo.  Every rank runs the same loop at the same time (barrier before, max over
ranks after), so the figures are what the box's PCIe / host-memory fabric gives N GPUs at once.  Variants:

  h2d / d2h   one direction alone                       both   upload || readback on two streams (the e2e pattern)
  wc          the input buffer allocated write-combined  pin    the rank's host threads pinned to a slice of the cores

Prints one JSON line per variant on rank 0.  ``ceiling_ms_per_step`` of variant "both" is the denominator of
``e2e.frac_of_transfer_ceiling`` in the bench line (bench.py measures its own copy of it in the same process).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import time

import torch
import torch.distributed as dist

WP, HP, W, H, N_VIEWS, BATCH = 8192, 4096, 1920, 1080, 12, 32
ROWS_TOUCHED = 3055                      # rows 0 .. 3054: p2p_view_row_range of the README example
H2D_IMG = ROWS_TOUCHED * WP * 3          # 75,087,360
D2H_IMG = N_VIEWS * W * H * 3            # 74,649,600


def host_buffer(nbytes: int, write_combined: bool):
    """Page-locked host tensor; write-combined through cudaHostAlloc when asked (torch has no flag for it)."""
    if not write_combined:
        return torch.empty(nbytes, dtype=torch.uint8, pin_memory=True), None
    import ctypes as C

    rt = C.CDLL("libcudart.so.12")
    ptr = C.c_void_p()
    flags = 0x01 | 0x04                   # cudaHostAllocPortable | cudaHostAllocWriteCombined
    rc = rt.cudaHostAlloc(C.byref(ptr), C.c_size_t(nbytes), C.c_uint(flags))
    if rc != 0:
        raise RuntimeError(f"cudaHostAlloc(write-combined) failed: {rc}")
    return ptr, rt


def run_variant(name, dev, world, steps, direction, wc=False):
    n_in, n_out = 3, 4                    # distinct pinned buffers, as the bench's e2e loop rotates them
    ins, outs, raw = [], [], []
    for _ in range(n_in):
        t, rt = host_buffer(H2D_IMG, wc)
        if rt is None:
            t.random_(0, 255)
        raw.append((t, rt))
        ins.append(t)
    for _ in range(n_out):
        outs.append(torch.empty(D2H_IMG, dtype=torch.uint8, pin_memory=True))
    d_in = [torch.empty(H2D_IMG, dtype=torch.uint8, device=dev) for _ in range(4)]
    d_out = [torch.empty(D2H_IMG, dtype=torch.uint8, device=dev) for _ in range(4)]
    s_up, s_dn = torch.cuda.Stream(dev), torch.cuda.Stream(dev)
    rtl = None
    if wc:
        import ctypes as C

        rtl = raw[0][1]
        rtl.cudaMemcpyAsync.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_int, C.c_void_p]

    def step():
        for i in range(BATCH):
            if direction in ("h2d", "both"):
                with torch.cuda.stream(s_up):
                    if wc:
                        rtl.cudaMemcpyAsync(d_in[i % 4].data_ptr(), ins[i % n_in], H2D_IMG, 1, s_up.cuda_stream)
                    else:
                        d_in[i % 4].copy_(ins[i % n_in], non_blocking=True)
            if direction in ("d2h", "both"):
                with torch.cuda.stream(s_dn):
                    outs[i % n_out].copy_(d_out[i % 4], non_blocking=True)

    for _ in range(2):
        step()
    torch.cuda.synchronize(dev)
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    for _ in range(steps):
        step()
    torch.cuda.synchronize(dev)
    el = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(el, op=dist.ReduceOp.MAX)
        dist.barrier()
    sec = float(el.item())
    up = BATCH * H2D_IMG * steps if direction in ("h2d", "both") else 0
    dn = BATCH * D2H_IMG * steps if direction in ("d2h", "both") else 0
    if wc:
        for ptr, rt in raw:
            rt.cudaFreeHost(ptr)
    return {"variant": name, "n_gpus": world, "steps": steps, "ceiling_ms_per_step": sec / steps * 1e3,
            "h2d_gbs_per_gpu": up / sec / 1e9, "d2h_gbs_per_gpu": dn / sec / 1e9,
            "aggregate_gbs": world * (up + dn) / sec / 1e9,
            "mpix_s_equivalent": world * steps * BATCH * N_VIEWS * W * H / sec / 1e6}


def topology():
    info = {"cpu_count": os.cpu_count(), "affinity": len(os.sched_getaffinity(0))}
    for key, cmd in (("nvidia_smi_topo", ["nvidia-smi", "topo", "-m"]), ("lscpu_numa", ["bash", "-c", "lscpu | grep -i 'numa\\|model name\\|socket'"])):
        try:
            info[key] = subprocess.run(cmd, capture_output=True, text=True, timeout=20).stdout.strip().splitlines()
        except Exception as e:  # noqa: BLE001
            info[key] = [f"unavailable: {e}"]
    return info


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--variants", nargs="+", default=["h2d", "d2h", "both", "both_wc", "both_pin"])
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    dev = torch.device("cuda", local)
    torch.cuda.set_device(dev)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    if rank == 0:
        print(json.dumps({"topology": topology()}), flush=True)
    all_cpus = sorted(os.sched_getaffinity(0))
    for v in args.variants:
        if v == "both_pin":   # this rank's threads on its own slice of the cores (the driver's copy threads inherit it)
            per = max(1, len(all_cpus) // world)
            os.sched_setaffinity(0, set(all_cpus[local * per:(local + 1) * per]) or set(all_cpus))
        res = run_variant(v, dev, world, args.steps, "both" if v.startswith("both") else v, wc=(v == "both_wc"))
        if v == "both_pin":
            os.sched_setaffinity(0, set(all_cpus))
        if rank == 0:
            print(json.dumps(res), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    sys.exit(main())
