"""world_size-2 gloo tests (CPU) of the multi-GPU host logic in mapping_private_b200/shard.py."""
import os
import pathlib
import sys

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

# spawned workers re-import this module without conftest.py: make the package importable here
sys.path.insert(0, str(pathlib.Path(__file__).resolve().parent.parent))
import pkgpath  # noqa: E402

pkgpath.load()
from mapping_private_b200 import shard  # noqa: E402


def test_split_range_covers_everything():
    for n in (0, 1, 7, 637_000):
        for w in (1, 2, 3, 8):
            r = shard.split_range(n, w)
            assert r[0][0] == 0 and r[-1][1] == n
            assert all(r[i][1] == r[i + 1][0] for i in range(w - 1))
            assert max(e - b for b, e in r) - min(e - b for b, e in r) <= 1


def test_lpt_assignment():
    sizes = [15000, 1300, 9000, 9000, 2000, 14000, 1300, 7000]
    parts = shard.assign_clusters_lpt(sizes, 3)
    assert sorted(sum(parts, [])) == list(range(8))
    loads = [sum(sizes[i] for i in p) for p in parts]
    assert max(loads) - min(loads) <= max(sizes)
    assert parts == shard.assign_clusters_lpt(sizes, 3)  # deterministic


def _worker(rank, world, port, tmp):
    import pathlib
    import sys

    root = pathlib.Path(__file__).resolve().parent.parent
    sys.path.insert(0, str(root))
    import pkgpath

    pkgpath.load()
    from mapping_private_b200 import shard as sh

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        # query-sharded results: each rank fills its slice, exchange concatenates them in place
        n = 1001
        ranges = [(0, 400), (400, 1001)]
        full = torch.arange(n * 4, dtype=torch.float32).reshape(n, 4)
        buf = torch.full((n, 4), -1.0)
        b, e = ranges[rank]
        buf[b:e] = full[b:e]
        sh.exchange_slices(buf, ranges)
        assert torch.equal(buf, full)
        # slice-wise upload + all-gather of the cloud (ragged and even splits)
        for npts in (1001, 1000):
            cloud = torch.arange(npts * 3, dtype=torch.float32).reshape(npts, 3)
            lo, hi = sh.split_range(npts, world)[rank]
            got = sh.gather_cloud(torch.full((npts, 3), -1.0), cloud[lo:hi].clone(), rank, world)
            assert torch.equal(got, cloud)
        # cluster-per-rank GRSD: integer histograms summed with one all-reduce
        sizes = [5, 9, 2, 7, 7, 1]
        parts = sh.assign_clusters_lpt(sizes, world)
        truth = torch.arange(6 * 21, dtype=torch.int32).reshape(6, 21)
        hist = torch.zeros((6, 21), dtype=torch.int32)
        for i in parts[rank]:
            hist[i] = truth[i]
        sh.allreduce_histograms(hist)
        assert torch.equal(hist, truth)
        # max-over-ranks timing reduction used by bench.py
        t = torch.tensor([1.0 + rank], dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        assert t.item() == float(world)
        (tmp / f"ok{rank}").write_text("ok")
    finally:
        dist.destroy_process_group()


def test_exchange_and_allreduce_world2(tmp_path):
    port = 29600 + os.getpid() % 300
    mp.spawn(_worker, args=(2, port, tmp_path), nprocs=2, join=True)
    assert (tmp_path / "ok0").exists() and (tmp_path / "ok1").exists()
