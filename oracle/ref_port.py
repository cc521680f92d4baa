"""Two-pass NumPy + cv2.remap restatement of the reference hot path.  TEST INFRASTRUCTURE.

This module re-states, in our own words, exactly what
``/root/reference/app/panorama_to_plane-pitch.py`` executes for one panorama: a yaw map over
the whole panorama, a ``cv2.remap`` with it, then one pitch map + ``cv2.remap`` per pitch.
It calls the same third-party routines the reference calls (NumPy ufuncs / matmul,
``cv2.remap(INTER_LINEAR, BORDER_CONSTANT)``) so it is bit-identical to the reference on the
same host; ``tests/test_oracle_golden.py`` pins it against stored reference outputs.

It is also the timed CPU baseline of ``bench.py`` (``cpu_baseline.kind = "port"`` and
``--impl reference``): same arithmetic, same libraries, same thread fan-out as the reference.

Never imported by the product package.
"""
from __future__ import annotations

import os
from concurrent.futures import ThreadPoolExecutor

import numpy as np

try:  # cv2 is only needed by the functions that sample; the map functions are NumPy-only
    import cv2
except Exception:  # pragma: no cover - cv2 is present in this image
    cv2 = None

# module-level memo tables, same keys as the reference (ref :17-18, :42-73)
_yaw_maps: dict = {}
_pitch_maps: dict = {}


def clear_caches() -> None:
    _yaw_maps.clear()
    _pitch_maps.clear()


def yaw_row(pano_width: int, yaw_deg) -> np.ndarray:
    """One row of the reference's yaw U map (all rows are identical).

    Follows ref :85-105.  dtype trail: ``phi`` is f32 (:95), ``phi + np.radians(yaw)`` is f64
    because ``np.radians`` of a Python int returns an ``np.float64`` scalar (:85, :98, NEP 50),
    the modulo / scale / clip run in f64 and the result is cast to f32 (:101-105).
    """
    yaw_radians = np.radians(yaw_deg)
    u = np.arange(pano_width, dtype=np.float32)
    phi = (2 * np.pi * u / pano_width).astype(np.float32)
    phi_rotated = (phi + yaw_radians) % (2 * np.pi)
    U = (phi_rotated * pano_width) / (2 * np.pi)
    return np.clip(U, 0, pano_width - 1).astype(np.float32)


def yaw_mapping(pano_width: int, pano_height: int, yaw_deg):
    """Full (U, V) yaw map, ref ``precompute_yaw_mapping`` :79-108."""
    row = yaw_row(pano_width, yaw_deg)
    U = np.broadcast_to(row, (pano_height, pano_width)).copy()
    V = np.broadcast_to(
        np.arange(pano_height, dtype=np.float32)[:, None], (pano_height, pano_width)
    ).copy()
    return U, V


def pitch_scalars(W: int, fov_deg, pitch_deg):
    """Host scalars of the pitch map, exactly as the reference forms them.

    ``f`` (ref :119, cast to f32 by ``np.full_like(..., dtype=float32)`` :131) and the f32
    entries ``c = cos(p)``, ``s = sin(p)`` of ``R_pitch`` (:142-149).  Angles go through
    ``np.radians`` (:64, :68).
    """
    fov_rad = np.radians(fov_deg)
    p = np.radians(pitch_deg)
    f = np.float32((0.5 * W) / np.tan(fov_rad / 2))
    return f, np.float32(np.cos(p)), np.float32(np.sin(p))


def pitch_mapping(W: int, H: int, fov_deg, pitch_deg, pano_width: int, pano_height: int, seam_wrap: bool = False):
    """(U, V) pitch map, ref ``precompute_pitch_mapping`` :114-175 via ``get_pitch_mapping`` :55-73.
    ``seam_wrap`` (NOT the reference: the oracle of the exact-bilinear mode's wrap option) limits U to [0, Wp) instead of
    the reference's [0, Wp - 1]."""
    fov_rad = np.radians(fov_deg)
    p = np.radians(pitch_deg)
    focal = (0.5 * W) / np.tan(fov_rad / 2)
    u, v = np.meshgrid(
        np.arange(W, dtype=np.float32), np.arange(H, dtype=np.float32), indexing="xy"
    )
    x = u - (W / 2.0)
    y = (H / 2.0) - v
    z = np.full_like(x, focal, dtype=np.float32)
    norm = np.sqrt(x**2 + y**2 + z**2)
    xn, yn, zn = x / norm, y / norm, z / norm
    R = np.array(
        [[1, 0, 0], [0, np.cos(p), -np.sin(p)], [0, np.sin(p), np.cos(p)]], dtype=np.float32
    )
    rot = R @ np.stack((xn, yn, zn), axis=0).reshape(3, -1)
    xr, yr, zr = rot.reshape(3, H, W)
    with np.errstate(invalid="ignore"):
        theta = np.arccos(zr).astype(np.float32)
    phi = (np.arctan2(yr, xr) % (2 * np.pi)).astype(np.float32)
    U = (phi * pano_width) / (2 * np.pi)
    V = (theta * pano_height) / np.pi
    u_max = np.nextafter(np.float32(pano_width), np.float32(0)) if seam_wrap else pano_width - 1
    U = np.clip(U, 0, u_max).astype(np.float32)
    V = np.clip(V, 0, pano_height - 1).astype(np.float32)
    return U, V


def get_yaw_mapping(pano_width, pano_height, yaw_deg):
    key = (pano_width, pano_height, yaw_deg)
    if key not in _yaw_maps:
        _yaw_maps[key] = yaw_mapping(pano_width, pano_height, yaw_deg)
    return _yaw_maps[key]


def get_pitch_mapping(W, H, pitch_deg, pano_width, pano_height, fov_deg=90):
    key = (W, H, pitch_deg, pano_width, pano_height, fov_deg)
    if key not in _pitch_maps:
        _pitch_maps[key] = pitch_mapping(W, H, fov_deg, pitch_deg, pano_width, pano_height)
    return _pitch_maps[key]


def _remap(src, U, V):
    return cv2.remap(src, U, V, interpolation=cv2.INTER_LINEAR, borderMode=cv2.BORDER_CONSTANT)


def process_yaw_and_pitchs(pano_image, yaw_angle, pitch_angles, output_width, output_height, fov_deg=90):
    """Ref ``process_yaw_and_pitchs`` :181-221: yaw remap once, then one pitch remap per pitch."""
    Hp, Wp, _ = pano_image.shape
    Uy, Vy = get_yaw_mapping(Wp, Hp, yaw_angle)
    rotated = _remap(pano_image, Uy, Vy)
    out = []
    for pitch in pitch_angles:
        Up, Vp = get_pitch_mapping(output_width, output_height, pitch, Wp, Hp, fov_deg)
        out.append(_remap(rotated, Up, Vp))
    return out


def default_workers() -> int:
    """Ref ``main`` :304-306."""
    return max(1, int((os.cpu_count() or 1) * 0.9))


def process_image_views(pano_image, yaw_angles, pitch_angles, W, H, fov_deg=90, num_workers=None):
    """Compute-only part of ref ``process_single_image`` :251-272 (no imread / imwrite):
    one thread-pool task per yaw, results consumed in submit order."""
    if num_workers is None:
        num_workers = default_workers()
    with ThreadPoolExecutor(max_workers=num_workers) as ex:
        futs = [
            ex.submit(process_yaw_and_pitchs, pano_image, yaw, pitch_angles, W, H, fov_deg)
            for yaw in yaw_angles
        ]
        return [f.result() for f in futs]
