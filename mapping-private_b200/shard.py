"""Host-side multi-GPU plumbing: one process per GPU, torch.distributed for the exchange steps.

The path shards by queries (contiguous ranges of the sorted order) and by clusters:
  * normals + RSD of one cloud: every rank computes its packet range; the only exchange is the
    concatenation of the ranks' slices (normals between the two passes, radii at the end),
    done in place with one broadcast per rank over NCCL / NVLink;
  * GRSD of a batch of clusters: clusters are assigned to ranks longest-first; the integer
    histograms are summed with one small all-reduce (bit-exact in any order).
Everything here works on CPU tensors with the gloo backend, which is how it is tested.
"""
from __future__ import annotations

from typing import List, Sequence, Tuple


def split_range(n_items: int, world: int) -> List[Tuple[int, int]]:
    """Same split as cab_set_shard / packet_range in csrc/cab_grid.cu: [n*r/w, n*(r+1)/w)."""
    return [(n_items * r // world, n_items * (r + 1) // world) for r in range(world)]


def exchange_slices(buf, ranges: Sequence[Tuple[int, int]], group=None):
    """buf: tensor whose rows [b_g, e_g) were produced by rank g.  After the call every rank holds
    all rows.  In place: rank g broadcasts a view of its own slice."""
    import torch.distributed as dist

    for g, (b, e) in enumerate(ranges):
        if e > b:
            dist.broadcast(buf[b:e], src=g, group=group)
    return buf


def gather_cloud(full, my_slice, rank: int, world: int, group=None):
    """Replicates a cloud whose rows were uploaded slice-wise: rank g copied rows split_range(n, world)[g] of the
    host cloud into `my_slice` (one H2D of n/world points per GPU); after the call `full` (n x 3, on the same
    device) holds all rows on every rank.  One all-gather over NVLink instead of `world` full uploads over PCIe."""
    import torch
    import torch.distributed as dist

    n = full.shape[0]
    ranges = split_range(n, world)
    b, e = ranges[rank]
    assert my_slice.shape[0] == e - b
    width = max(hi - lo for lo, hi in ranges)
    if all(hi - lo == width for lo, hi in ranges) and full.is_contiguous():
        dist.all_gather_into_tensor(full, my_slice.contiguous(), group=group)
        return full
    # ragged split: pad every slice to the widest one
    pad = torch.zeros((width,) + tuple(full.shape[1:]), dtype=full.dtype, device=full.device)
    pad[: e - b] = my_slice
    parts = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(parts, pad, group=group)
    for g, (lo, hi) in enumerate(ranges):
        full[lo:hi] = parts[g][: hi - lo]
    return full


def assign_clusters_lpt(sizes: Sequence[int], world: int) -> List[List[int]]:
    """Longest-processing-time-first assignment of clusters to ranks (deterministic)."""
    order = sorted(range(len(sizes)), key=lambda i: (-int(sizes[i]), i))
    load = [0] * world
    out: List[List[int]] = [[] for _ in range(world)]
    for i in order:
        r = min(range(world), key=lambda k: (load[k], k))
        out[r].append(i)
        load[r] += int(sizes[i])
    return [sorted(v) for v in out]


def allreduce_histograms(hist_full, group=None):
    """hist_full: int32 tensor (n_clusters, 21), zero except for this rank's clusters.  Sum over ranks."""
    import torch.distributed as dist

    dist.all_reduce(hist_full, op=dist.ReduceOp.SUM, group=group)
    return hist_full
