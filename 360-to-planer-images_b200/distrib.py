"""One process per GPU: the only cross-rank traffic of this path is a barrier and the maximum of a
timing / the sum of a counter (the projection shards by image or view with no data exchange,
SURVEY.md 8e).  ``torch.distributed`` is plumbing: ``nccl`` on the GPU box, ``gloo`` in the CPU
tests.  With WORLD_SIZE == 1 nothing is initialised.
"""
from __future__ import annotations

import os


class Ranks:
    def __init__(self, backend: str | None = None, device=None):
        self.rank = int(os.environ.get("RANK", "0"))
        self.local_rank = int(os.environ.get("LOCAL_RANK", "0"))
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.dist = None
        self.device = device
        if self.world > 1:
            import torch.distributed as dist

            if not dist.is_initialized():
                # keep stdout clean for the one JSON line: NCCL's banner / debug output goes to stderr
                os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
                kw = {}
                if backend == "nccl" and device is not None:
                    kw["device_id"] = device
                dist.init_process_group(backend or "gloo", **kw)
            self.dist = dist

    def barrier(self):
        if self.dist is not None:
            self.dist.barrier()

    def _reduce(self, x: float, op_name: str) -> float:
        if self.dist is None:
            return float(x)
        import torch

        t = torch.tensor([float(x)], dtype=torch.float64, device=self.device if self.device is not None else "cpu")
        self.dist.all_reduce(t, op=getattr(self.dist.ReduceOp, op_name))
        return float(t.item())

    def max(self, x: float) -> float:
        return self._reduce(x, "MAX")

    def sum(self, x: float) -> float:
        return self._reduce(x, "SUM")

    def close(self):
        if self.dist is not None and self.dist.is_initialized():
            self.dist.destroy_process_group()


def aggregate_throughput(units_per_rank: float, seconds_this_rank: float, ranks: Ranks) -> float:
    """Whole-job throughput: all ranks' units / the slowest rank's time."""
    total = ranks.sum(units_per_rank)
    slowest = ranks.max(seconds_this_rank)
    return total / slowest
