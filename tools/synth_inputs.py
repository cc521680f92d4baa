"""Deterministic synthetic equirectangular panoramas: the inputs of the benchmarks and tests.
Plain NumPy generators - no projection arithmetic, not part of the oracle (``oracle/synth.py`` re-exports them for tests).

``noise``   uniform u8 noise: every coordinate flip shows up, used for bit-exact sampler gates
``smooth``  band-limited, x-periodic, gradient-bounded image: the <=1 LSB end-to-end gate
``coords``  B = x & 255, G = x >> 8, R = y & 255: a mapping debugger
"""
from __future__ import annotations

import numpy as np

# (k, m, phase) per channel; the C2 example of SURVEY 8d, scaled down for small panoramas
_SMOOTH_C2 = ((37, 11, 0.3), (53, 7, 1.1), (29, 17, 2.0))


def noise(Wp: int, Hp: int, seed: int = 0) -> np.ndarray:
    return np.random.default_rng(seed).integers(0, 256, (Hp, Wp, 3), dtype=np.uint8)


def smooth(Wp: int, Hp: int, seed: int = 0) -> np.ndarray:
    """127.5 + 127.5 sin(k 2 pi x / Wp + ph) cos(m pi y / Hp), k <= Wp/160, m <= Hp/240."""
    rng = np.random.default_rng(seed)
    x = np.arange(Wp, dtype=np.float64)[None, :]
    y = np.arange(Hp, dtype=np.float64)[:, None]
    out = np.empty((Hp, Wp, 3), dtype=np.uint8)
    kmax = max(1, Wp // 160)
    mmax = max(1, Hp // 240)
    for ch, (k0, m0, ph0) in enumerate(_SMOOTH_C2):
        k = max(1, min(kmax, (k0 * Wp) // 8192 + int(rng.integers(0, 2))))
        m = max(1, min(mmax, (m0 * Hp) // 4096 + int(rng.integers(0, 2))))
        ph = ph0 + float(rng.random())
        val = 127.5 + 127.5 * np.sin(k * 2 * np.pi * x / Wp + ph) * np.cos(m * np.pi * y / Hp)
        out[..., ch] = np.clip(np.rint(val), 0, 255).astype(np.uint8)
    return out


def coords(Wp: int, Hp: int) -> np.ndarray:
    x = np.arange(Wp, dtype=np.int64)[None, :]
    y = np.arange(Hp, dtype=np.int64)[:, None]
    out = np.empty((Hp, Wp, 3), dtype=np.uint8)
    out[..., 0] = np.broadcast_to(x & 255, (Hp, Wp))
    out[..., 1] = np.broadcast_to((x >> 8) & 255, (Hp, Wp))
    out[..., 2] = np.broadcast_to(y & 255, (Hp, Wp))
    return out


def make(kind: str, Wp: int, Hp: int, seed: int = 0) -> np.ndarray:
    if kind == "noise":
        return noise(Wp, Hp, seed)
    if kind == "smooth":
        return smooth(Wp, Hp, seed)
    if kind == "coords":
        return coords(Wp, Hp)
    raise ValueError(kind)
