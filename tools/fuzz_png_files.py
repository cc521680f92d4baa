"""Random VALID PNG files against the library's host model of the device decoder (the kernels' own __host__ __device__
routines, run serially - no GPU needed): random sizes, colour types, filter choices, zlib levels / strategies / window
sizes / memory levels, IDAT splits and flushes; every file must decode to the pixels cv2.imdecode returns.

    python tools/fuzz_png_files.py [n_files] > profiles/r2_fuzz_png_files.jsonl
"""
import ctypes as C
import json
import sys
import time
import zlib
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import __graft_entry__ as g  # noqa: E402
from oracle import png_decode_model as M  # noqa: E402


def main():
    import cv2

    n_files = int(sys.argv[1]) if len(sys.argv) > 1 else 1500
    g.build()
    lib = g.load_package()._lib.load()
    rng = np.random.default_rng(4242)
    stats = {"files": 0, "identical": 0, "declined": 0, "different": 0, "blocks": 0, "blocks_found_by_search": 0, "bytes": 0}
    t0 = time.time()
    for i in range(n_files):
        ctype, ch = [(0, 1), (2, 3), (4, 2), (6, 4)][int(rng.integers(0, 4))]
        H, W = int(rng.integers(1, 160)), int(rng.integers(1, 260))
        kind = ["mixed", "smooth", "noise"][int(rng.integers(0, 3))]
        img = M.test_image(H, W, ch, int(rng.integers(0, 1 << 30)), kind)
        fsel = int(rng.integers(0, 4))
        filters = "adaptive" if fsel == 0 else ([int(rng.integers(0, 5))] * H if fsel == 1 else list(rng.integers(0, 5, H)))
        kw = dict(
            filters=filters, level=int(rng.integers(0, 10)),
            strategy=[zlib.Z_DEFAULT_STRATEGY, zlib.Z_FILTERED, zlib.Z_HUFFMAN_ONLY, zlib.Z_RLE, zlib.Z_FIXED][int(rng.integers(0, 5))],
            wbits=int(rng.integers(9, 16)), mem_level=int(rng.integers(1, 10)),
            idat=[int(v) for v in rng.integers(1, 5000, int(rng.integers(1, 4)))],
            flush_every=int(rng.integers(50, 20000)) if rng.integers(0, 3) == 0 else 0,
            flush_mode=[zlib.Z_SYNC_FLUSH, zlib.Z_FULL_FLUSH][int(rng.integers(0, 2))],
        )
        data = M.write_png(img, ctype, **kw)
        ref = cv2.imdecode(np.frombuffer(data, np.uint8), cv2.IMREAD_COLOR)
        w, h = C.c_int(), C.c_int()
        stats["files"] += 1
        stats["bytes"] += len(data)
        if ref is None or lib.p2p_png_probe(data, len(data), C.byref(w), C.byref(h)):
            stats["declined"] += 1
            continue
        out = np.zeros((h.value, w.value, 3), np.uint8)
        st = (C.c_uint64 * 4)()
        rc = lib.p2p_png_decode_host(data, len(data), out.ctypes.data, out.strides[0], h.value, st)
        if rc:
            stats["declined"] += 1
            print(json.dumps({"declined_valid_file": i, "kw": {k: (v if not isinstance(v, list) else "list") for k, v in kw.items()}}), flush=True)
        elif np.array_equal(out, ref):
            stats["identical"] += 1
            stats["blocks"] += st[2]
            stats["blocks_found_by_search"] += st[2] - st[3]
        else:
            stats["different"] += 1
            print(json.dumps({"different": i}), flush=True)
    stats["seconds"] = round(time.time() - t0, 1)
    stats["what"] = ("random valid PNG files (4 colour types, sizes 1..159 x 1..259, filters per file / per row / adaptive, zlib levels 0-9, "
                     "5 strategies, windows 512 B - 32 KiB, memLevel 1-9, random IDAT splits, sync / full flushes) through "
                     "p2p_png_decode_host vs cv2.imdecode; different and declined must be 0")
    print(json.dumps(stats))
    return 1 if stats["different"] or stats["declined"] else 0


if __name__ == "__main__":
    sys.exit(main())
