"""GRSD of ONE large cloud (last stage): cab_grsd_cloud on one GPU equals cab_grsd_batch on the cloud as a single
cluster (whose voxels, labels and histograms the oracle pins in tests/test_gpu_parity.py), and the sharded form -- every
rank labels the voxels of its own rows, labels merged, transition counts summed with one int32 all-reduce -- equals the
single-GPU histograms bit for bit, with and without subdivisions (grsd_colorCHLAC_tools.hpp:140-161, 230-260)."""
import threading

import numpy as np
import pytest

from mapping_private_b200 import cab, synth

pytestmark = pytest.mark.gpu
LEAF = 0.025


def _reference(pts):
    """Single GPU, the batch path with one cluster."""
    c = cab.Context(0, exact=True)
    off = np.array([0, pts.shape[0]], np.int32)
    hist = c.grsd_batch(pts, off, LEAF)[0]
    vox = c.grsd_voxels(1)
    sub = c.grsd_signatures(1, cab.SIG_GRSD21, subdivision_size=10)
    s325 = c.grsd_signatures(1, cab.SIG_GRSD325)
    c.close()
    return hist, vox, sub, s325


@pytest.fixture(scope="module")
def cloud():
    pts = synth.room(1_500_000)
    pts[17] = np.nan
    return pts


@pytest.fixture(scope="module")
def reference(cloud):
    return _reference(cloud)


def test_one_gpu_equals_the_batch_path(cloud, reference):
    hist, vox, sub, s325 = reference
    assert hist.sum() > 0 and (hist > 0).sum() >= 6
    c = cab.Context(0, exact=True)
    c.upload(cloud)
    got = c.grsd_cloud(LEAF)
    assert np.array_equal(got, hist)
    v = c.grsd_voxels(1)
    assert np.array_equal(v["labels"], vox["labels"]) and np.array_equal(v["centroids"].view(np.uint32), vox["centroids"].view(np.uint32))
    s = c.grsd_signatures(1, cab.SIG_GRSD21, subdivision_size=10)
    assert np.array_equal(s["hist"], sub["hist"]) and np.array_equal(s["subdiv_b"], sub["subdiv_b"])
    c.close()


def test_sharded_by_hand_without_a_group(cloud, reference):
    """cab_set_shard only: the application merges the labels and sums the histograms itself."""
    hist, vox, sub, _ = reference
    world = 3
    ctxs, labels = [], []
    for r in range(world):
        c = cab.Context(0, exact=True)
        c.upload(cloud)
        c.set_shard(r, world)
        labels.append(c.grsd_cloud_labels(LEAF))
        ctxs.append(c)
    owners = sum((l > 0).astype(np.int32) for l in labels)
    assert np.array_equal(owners, np.ones_like(owners))  # every voxel labelled by exactly one rank
    merged = sum(labels)
    assert np.array_equal(merged - 1, vox["labels"])
    parts, parts_sub = [], []
    for c in ctxs:
        c.grsd_cloud_set_labels(merged)
        parts.append(c.grsd_signatures(1, cab.SIG_GRSD21)["hist"][0])
        parts_sub.append(c.grsd_signatures(1, cab.SIG_GRSD21, subdivision_size=10)["hist"])
        with pytest.raises(cab.CabError, match="PlusGRSD"):
            c.grsd_signatures(1, cab.SIG_PLUSGRSD110)
        with pytest.raises(cab.CabError, match="outside a group"):
            c.grsd_cloud(LEAF)
        c.close()
    assert all(p.sum() > 0 for p in parts)  # every rank had work
    assert np.array_equal(sum(parts), hist)
    assert np.array_equal(sum(parts_sub), sub["hist"])


@pytest.mark.parametrize("world,points", [(3, 0), (2, 5_000_000)])
def test_group_allreduce(cloud, reference, world, points):
    """A group (contexts of one process, one thread each): cab_grsd_cloud leaves the whole cloud's GRSD-21 on every rank;
    the signatures with subdivisions and GRSD-325 follow the same way.  One case on a 5 M-point cloud."""
    if points:
        pts = synth.room(points)
        hist, vox, sub, s325 = _reference(pts)
    else:
        pts = cloud
        hist, vox, sub, s325 = reference
    ctxs = [cab.Context(0, exact=True) for _ in range(world)]
    cab.comm_init_local(ctxs)
    out, errs = [None] * world, []

    def work(r):
        try:
            c = ctxs[r]
            c.comm_upload_cloud(pts)
            h = c.grsd_cloud(LEAF)
            s = c.grsd_signatures(1, cab.SIG_GRSD21, subdivision_size=10)["hist"]
            t = c.grsd_signatures(1, cab.SIG_GRSD325)["hist"]
            out[r] = (h, s, t, c.grsd_voxels(1)["labels"], c.profile()["n_sorted"])
        except Exception as e:  # noqa: BLE001
            errs.append((r, repr(e)))

    ts = [threading.Thread(target=work, args=(r,)) for r in range(world)]
    for t in ts:
        t.start()
    for t in ts:
        t.join(timeout=300)
    assert not errs, errs
    for r in range(world):
        h, s, t, lab, n_sorted = out[r]
        assert np.array_equal(h, hist), f"rank {r}"
        assert np.array_equal(s, sub["hist"]) and np.array_equal(t, s325["hist"])
        assert np.array_equal(lab, vox["labels"])  # the merged labels
        assert n_sorted < 0.75 * pts.shape[0]  # the rank sorted its slab, not the cloud
    for c in ctxs:
        c.close()
