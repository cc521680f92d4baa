"""Work partitioning across the GPUs of one box.  No collective is involved: a work item is
(image, yaw, pitch), outputs are disjoint and the only shared input is the read-only panorama
(SURVEY.md 8e).  Batches shard by image so every panorama is uploaded to exactly one GPU; a
single image replicates the panorama and shards the flat view list (yaw-major) in contiguous
runs so each rank still amortises one coordinate evaluation over its yaws.
"""
from __future__ import annotations


def shard_images(n_images: int, rank: int, world: int) -> list:
    """Image indices of ``rank``: round-robin ``i % world == rank``."""
    if world <= 0 or not (0 <= rank < world):
        raise ValueError(f"bad rank/world {rank}/{world}")
    return list(range(rank, n_images, world))


def shard_views(n_yaw: int, n_pitch: int, rank: int, world: int) -> list:
    """(yaw_index, pitch_index) pairs of ``rank`` for a single replicated panorama.

    Views are split by pitch first (a pitch group shares its coordinate evaluation across all
    yaws), then by yaw when there are more ranks than pitches; counts differ by at most one.
    """
    if world <= 0 or not (0 <= rank < world):
        raise ValueError(f"bad rank/world {rank}/{world}")
    flat = [(k, j) for j in range(n_pitch) for k in range(n_yaw)]  # pitch-major
    n = len(flat)
    lo = (n * rank) // world
    hi = (n * (rank + 1)) // world
    return flat[lo:hi]


def shard_rows(height: int, rank: int, world: int, align: int = 8) -> tuple:
    """(row_begin, row_end) of ``rank`` when ONE image is split over ``world`` GPUs by output row bands: every GPU renders
    the same rows of every view, so twelve views balance over eight GPUs (by whole views they would split 2,2,2,2,1,1,1,1,
    SURVEY 8e).  Bands are multiples of ``align`` rows (the kernel's CTA tile height) except the last one; they cover
    [0, height) exactly once and may be empty when there are more ranks than tiles."""
    if world <= 0 or not (0 <= rank < world):
        raise ValueError(f"bad rank/world {rank}/{world}")
    tiles = (height + align - 1) // align
    lo = min(height, ((tiles * rank) // world) * align)
    hi = min(height, ((tiles * (rank + 1)) // world) * align)
    return lo, hi


def group_by_pitch(views: list) -> dict:
    """{pitch_index: [yaw_index, ...]} preserving order: one launch group per pitch."""
    out: dict = {}
    for k, j in views:
        out.setdefault(j, []).append(k)
    return out
