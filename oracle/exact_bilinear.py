"""Oracle of the optional "exact bilinear" mode: ``scipy.ndimage.map_coordinates``
with ``order=1`` per channel on the (rolled) panorama - the call the north-star names.  The
reference itself never calls scipy, so this mode is pinned against scipy, not against
the reference.  TEST INFRASTRUCTURE.
"""
from __future__ import annotations

import numpy as np

from . import ref_port


def sample_view_exact(pano: np.ndarray, U: np.ndarray, V: np.ndarray, yaw_shift: int = 0, seam_wrap: bool = False) -> np.ndarray:
    """u8 [H, W, 3]: order-1 map_coordinates of ``np.roll(pano, -yaw_shift, axis=1)`` at (V, U).
    NaN coordinates give 0 (cval).  For the in-range coordinates of this path the boundary mode is
    irrelevant (the out-of-range neighbour always has weight 0) - except with ``seam_wrap``, where U runs over [0, Wp)
    and scipy's ``mode='grid-wrap'`` supplies column 0 as the right-hand neighbour of column Wp - 1."""
    from scipy.ndimage import map_coordinates

    rolled = np.roll(pano, -int(yaw_shift), axis=1)
    U = np.asarray(U, np.float32)
    V = np.asarray(V, np.float32)
    dead = np.isnan(U) | np.isnan(V)
    coords = np.stack([np.where(dead, 0, V).astype(np.float64), np.where(dead, 0, U).astype(np.float64)])
    out = np.empty(U.shape + (3,), np.uint8)
    for c in range(3):
        out[..., c] = map_coordinates(rolled[..., c], coords, order=1, mode="grid-wrap" if seam_wrap else "nearest",
                                      prefilter=False)
    out[dead] = 0
    return out


def project_view_exact(pano, yaw_deg, pitch_deg, W, H, fov_deg=90, seam_wrap: bool = False):
    Hp, Wp, _ = pano.shape
    U, V = ref_port.pitch_mapping(W, H, fov_deg, pitch_deg, Wp, Hp, seam_wrap=seam_wrap)
    shift = yaw_deg * Wp / 360.0
    assert float(shift).is_integer(), "exact mode is defined for integer column rolls"
    return sample_view_exact(pano, U, V, int(shift) % Wp, seam_wrap=seam_wrap)
