"""Build ``libp2p_b200.so`` in-tree with nvcc for sm_100a (no JIT cache, no cmake).

    python 360-to-planer-images_b200/build.py [--force]
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys
from pathlib import Path

PKG_DIR = Path(__file__).resolve().parent
ROOT = PKG_DIR.parent
SOURCES = [PKG_DIR / "csrc" / "p2p_api.cu"]
HEADERS = sorted(p for p in (PKG_DIR / "csrc").iterdir() if p.suffix in (".cuh", ".inl", ".inc")) + [ROOT / "include" / "p2p.h"]
OUT = PKG_DIR / "libp2p_b200.so"

NVCC_FLAGS = [
    "-shared", "-Xcompiler", "-fPIC", "-std=c++17", "-O3", "-lineinfo",
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-I", str(ROOT / "include"),
]


def find_nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and Path(cand).exists():
            return cand
    raise RuntimeError("nvcc not found (set NVCC=/path/to/nvcc)")


def up_to_date() -> bool:
    if not OUT.exists():
        return False
    t = OUT.stat().st_mtime
    return all(p.stat().st_mtime <= t for p in SOURCES + HEADERS + [Path(__file__)])


def build(force: bool = False, verbose: bool = False) -> Path:
    if not force and up_to_date():
        return OUT
    cmd = [find_nvcc(), *NVCC_FLAGS, "-o", str(OUT), *map(str, SOURCES)]
    if verbose:
        cmd.insert(1, "-Xptxas=-v")
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + " ".join(cmd) + "\n" + res.stdout + res.stderr)
    if verbose:
        print(res.stderr)
    return OUT


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
