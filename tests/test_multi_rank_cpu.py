"""World-size-2 gloo test of the N > 1 plumbing: sharding covers the work exactly once, the
timing is the max over ranks and the throughput the whole-job aggregate (no data-path collective)."""
import os
import socket
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out_dir):
    os.environ.update(RANK=str(rank), LOCAL_RANK=str(rank), WORLD_SIZE=str(world),
                      MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    sys.path.insert(0, str(ROOT))
    import importlib.util

    pkg_dir = ROOT / "360-to-planer-images_b200"

    def load(name):
        spec = importlib.util.spec_from_file_location(f"p2p_cpu_{name}", pkg_dir / f"{name}.py")
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
        return mod

    distrib, shard = load("distrib"), load("shard")
    r = distrib.Ranks("gloo")
    assert (r.rank, r.world) == (rank, world)
    images = shard.shard_images(9, r.rank, r.world)
    views = shard.shard_views(4, 3, r.rank, r.world)
    r.barrier()
    # rank 1 pretends to be slower: the job time must be its time on every rank
    seconds = 1.0 + rank
    slowest = r.max(seconds)
    n_images = r.sum(len(images))
    n_views = r.sum(len(views))
    thr = distrib.aggregate_throughput(len(images) * 100.0, seconds, r)
    (Path(out_dir) / f"rank{rank}.txt").write_text(
        f"{slowest} {n_images} {n_views} {thr} {images} {views}")
    r.close()


def test_two_ranks_gloo(tmp_path):
    torch = pytest.importorskip("torch")
    import torch.multiprocessing as mp

    port = _free_port()
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    seen_images, seen_views = [], []
    for rank in range(2):
        slowest, n_images, n_views, thr, rest = (tmp_path / f"rank{rank}.txt").read_text().split(" ", 4)
        assert float(slowest) == 2.0          # max over ranks
        assert float(n_images) == 9.0 and float(n_views) == 12.0
        assert float(thr) == pytest.approx(900.0 / 2.0)  # all units / slowest rank
        images, views = eval(rest.replace("] [", "]|[").split("|")[0]), eval(rest.replace("] [", "]|[").split("|")[1])
        seen_images += images
        seen_views += views
    assert sorted(seen_images) == list(range(9))
    assert sorted(seen_views) == sorted((k, j) for k in range(4) for j in range(3))
