"""Multi-GPU data plane behind the C ABI (csrc/cab_comm.cu, cab_step_normals_rsd): slab-sharded ranks whose RSD kernel
stores its results straight into every rank's copy of the concatenated arrays.  The single-GPU result is the checker:
every rank of a group must end up with exactly those bits.

* contexts of one process (cab_comm_init_local), three ranks on cuda:0, one host thread each;
* one process per rank, CUDA IPC (cab_comm_reserve / cab_comm_connect), two ranks on cuda:0;
* NCCL bootstrap (cab_comm_get_id / cab_comm_init) + the integer all-reduce, two ranks on two GPUs (skipped on one).
"""
import multiprocessing as mp
import threading

import numpy as np
import pytest

from mapping_private_b200 import cab, synth

import comm_workers  # tests/comm_workers.py: the rank processes (importable without conftest)

pytestmark = pytest.mark.gpu

R = 0.02


def _single_gpu(pts, exact=False, max_nn=0):
    c = cab.Context(0, exact=exact)
    c.upload(pts)
    c.build_grid(R)
    n4 = c.normals(R, max_nn=max_nn)
    rmin, rmax = c.rsd(R, max_nn=max_nn)
    k = c.profile()["neighbour_sum"]
    c.close()
    return n4, rmin, rmax, k


def _same(a, b):
    return np.array_equal(np.ascontiguousarray(a).view(np.uint32), np.ascontiguousarray(b).view(np.uint32))


def test_step_equals_the_three_stages():
    """cab_step_normals_rsd (one synchronisation per step) leaves the bits of build + normals + RSD."""
    pts = synth.tabletop(40_000, noise_sigma=0.0003)
    pts[11] = np.nan
    n4, rmin, rmax, k = _single_gpu(pts)
    c = cab.Context(0)
    c.upload(pts)
    for _ in range(2):
        c.step_normals_rsd(R, R)
        s4, smin, smax = c.download()
        assert _same(s4, n4) and _same(smin, rmin) and _same(smax, rmax)
        p = c.profile()
        assert p["neighbour_sum"] == k and p["step_ms"] > 0
    dif = c.download_rdif()
    assert np.abs(dif - (rmax - rmin)).max() <= np.spacing(np.float32(0.1))
    c.close()


class _DevArray:
    """A device pointer as a __cuda_array_interface__ object (torch.as_tensor reads it without a copy)."""

    def __init__(self, ptr, shape, typestr):
        self.__cuda_array_interface__ = {"shape": shape, "typestr": typestr, "data": (ptr, False), "version": 3}


@pytest.mark.parametrize("max_nn", [0, 60])
def test_step_input_order_arrays(max_nn):
    """CAB_STEP_INPUT_ORDER: the RSD kernel of a single context's step also scatters normals and radii to their input
    index (what a group's step leaves rank by rank); same bits as cab_download, non-finite points included."""
    import torch

    pts = synth.tabletop(40_000, noise_sigma=0.0003)
    pts[[7, 11, 39_999]] = np.nan
    n = pts.shape[0]
    c = cab.Context(0)
    c.upload(pts)
    assert c.device_ptr(cab.BUF_NRM_INPUT_RANGE) == 0  # nothing yet
    for _ in range(2):
        c.step_normals_rsd(R, R, max_nn_rsd=max_nn, flags=cab.STEP_INPUT_ORDER)
        f4, fmin, fmax = c.download()
        in4 = torch.as_tensor(_DevArray(c.device_ptr(cab.BUF_NRM_INPUT_RANGE), (n, 4), "<f4"), device="cuda:0").cpu().numpy()
        in2 = torch.as_tensor(_DevArray(c.device_ptr(cab.BUF_RSD_INPUT_RANGE), (n, 2), "<f4"), device="cuda:0").cpu().numpy()
        assert _same(in4, f4) and _same(in2[:, 0], fmin) and _same(in2[:, 1], fmax)
    c.step_normals_rsd(R, R)  # without the flag the arrays are not offered
    assert c.device_ptr(cab.BUF_NRM_INPUT_RANGE) == 0
    c.close()


@pytest.mark.parametrize("max_nn", [60, 150])
def test_step_with_truncated_rsd(max_nn):
    """Normals unlimited, RSD truncated at max_nn (the plugin defaults, radius_estimation.h:82): in one call the normals
    traversal takes the d2 histogram of the truncation along; same bits as the three stages, whose RSD pass builds the
    histogram in a traversal of its own.  Pairs of stray points (two neighbours: no normal) sit next to the table: they
    count towards max_nn but form no pair."""
    pts = synth.analytic_shape("plane", 8_000, side=0.2)  # ~250 neighbours per query: the truncation is active everywhere
    rng = np.random.default_rng(5)
    # a stray point 19.99 mm off the plane sees the point below it and hardly anything else (0.6 mm around its foot)
    stray = pts[rng.choice(pts.shape[0], 40, replace=False)] + np.float32([0.0, 0.0, 0.01999])
    pts = np.concatenate([pts, stray]).astype(np.float32)
    pts[11] = np.nan
    c = cab.Context(0)
    c.upload(pts)
    c.build_grid(R)
    n4 = c.normals(R)
    rmin, rmax = c.rsd(R, max_nn=max_nn)
    k = c.profile()["neighbour_sum"]
    assert np.isnan(n4[-40:, 0]).sum() > 0  # some stray pairs really have no normal
    for _ in range(2):
        c.step_normals_rsd(R, R, max_nn_rsd=max_nn)
        s4, smin, smax = c.download()
        assert _same(s4, n4) and _same(smin, rmin) and _same(smax, rmax)
        assert c.profile()["neighbour_sum"] == k
    c.close()


def test_local_group_on_a_sparse_cloud_with_wide_cells():
    """A cloud sparse enough for cells that are wide along x (mean 16 neighbours; cab_grid.cu, Domain.xwide): the slab
    build of a group's ranks dimensions its table the same way, and every rank ends up with the single-GPU results."""
    pts = synth.density_patches(300_000, 16.0, R)
    n = pts.shape[0]
    n4, rmin, rmax, k = _single_gpu(pts)
    world = 2
    ctxs = [cab.Context(0) for _ in range(world)]
    cab.comm_init_local(ctxs)
    out, errs = [None] * world, []

    def work(r):
        try:
            c = ctxs[r]
            c.comm_upload_cloud(pts)
            c.step_normals_rsd(R, R)
            out[r] = (c.comm_download_range(0, n), c.profile())
        except Exception as e:  # noqa: BLE001
            errs.append((r, repr(e)))

    ts = [threading.Thread(target=work, args=(r,)) for r in range(world)]
    for t in ts:
        t.start()
    for t in ts:
        t.join(timeout=120)
    assert not errs, errs
    ext = pts.max(axis=0).astype(np.float64) - pts.min(axis=0).astype(np.float64)
    cubic = np.prod(np.floor(ext / (R * (1 + 1 / 1024))) + 2)
    for r in range(world):
        (f4, fmin, fmax), prof = out[r]
        assert _same(f4, n4) and _same(fmin, rmin) and _same(fmax, rmax), f"rank {r}"
        assert prof["n_cells"] < cubic / 4  # a slab of a wide table
    assert sum(out[r][1]["neighbour_sum"] for r in range(world)) == k
    for c in ctxs:
        c.close()


@pytest.mark.parametrize("world,max_nn", [(3, 0), (2, 60)])
def test_local_group_every_rank_holds_all_results(world, max_nn):
    """Three contexts of this process on one GPU: after the step EVERY rank's concatenated arrays hold the single-GPU
    results of all points (input order through cab_comm_download_range), pushed there by the ranks' RSD kernels."""
    pts = synth.tabletop(90_000, noise_sigma=0.0002)
    pts[5] = np.nan  # nobody's query
    n = pts.shape[0]
    n4, rmin, rmax, k = _single_gpu(pts, max_nn=max_nn)
    ctxs = [cab.Context(0) for _ in range(world)]
    cab.comm_init_local(ctxs)
    for c in ctxs:
        # ranks that share one GPU time each other's kernels, not their own work: with the feedback on, the cuts of the
        # later steps follow that noise and a slab may come out thinner than two layers (correct, but shard_mode 1)
        c.comm_set_feedback(False)
    out, errs = [None] * world, []
    gate = threading.Barrier(world)

    def work(r):
        try:
            c = ctxs[r]
            for it in range(3):  # the later steps reuse the connected buffers
                c.comm_upload_cloud(pts)
                c.step_normals_rsd(R, R, max_nn_normals=max_nn, max_nn_rsd=max_nn)
                f4, fmin, fmax = c.comm_download_range(0, n)  # the very first step of a fresh group included
                assert _same(f4, n4) and _same(fmin, rmin) and _same(fmax, rmax), f"rank {r}, step {it}"
                gate.wait(timeout=60)  # nobody starts pushing the next step's results into arrays still being read
            full = c.comm_download_range(0, n)
            lo, hi = n * r // world, n * (r + 1) // world
            part = c.comm_download_range(lo, hi)
            out[r] = (full, part, c.profile(), c.shard_range())
        except Exception as e:  # noqa: BLE001
            errs.append((r, repr(e)))

    ts = [threading.Thread(target=work, args=(r,)) for r in range(world)]
    for t in ts:
        t.start()
    for t in ts:
        t.join(timeout=120)
    assert not errs, errs
    total_k = 0
    for r in range(world):
        (f4, fmin, fmax), (p4, pmin, pmax), prof, (b, e) = out[r]
        assert _same(f4, n4) and _same(fmin, rmin) and _same(fmax, rmax), f"rank {r}"
        lo, hi = n * r // world, n * (r + 1) // world
        assert _same(p4, n4[lo:hi]) and _same(pmin, rmin[lo:hi]) and _same(pmax, rmax[lo:hi])
        total_k += prof["neighbour_sum"]
        assert e > b
        assert prof["shard_mode"] == 2  # thick slabs: the halo rows' normals were exchanged, not recomputed
    if max_nn == 0:
        assert total_k == k  # every query answered by exactly one rank
    for c in ctxs:
        c.close()


@pytest.mark.parametrize("world,max_nn", [(3, 0), (4, 60)])
def test_local_group_input_range_layout(world, max_nn):
    """CAB_COMM_LAYOUT_INPUT_RANGES: every result crosses to ONE rank, the owner of the point's input index; after the
    step rank g holds input indices [n g / world, n (g + 1) / world) in input order -- the ranks' ranges, rank after rank,
    are the single-GPU channels bit for bit, non-finite points included."""
    pts = synth.tabletop(90_000, noise_sigma=0.0002)
    n = pts.shape[0]
    for j in (5, n // world, n // 2 + 1, n - 1):  # nobody's query: their owners fill in the defaults
        pts[j] = np.nan
    n4, rmin, rmax, k = _single_gpu(pts, max_nn=max_nn)
    ctxs = [cab.Context(0) for _ in range(world)]
    cab.comm_init_local(ctxs)
    out, errs = [None] * world, []
    gate = threading.Barrier(world)

    def work(r):
        try:
            c = ctxs[r]
            c.comm_set_layout(cab.COMM_LAYOUT_INPUT_RANGES)
            lo, hi = n * r // world, n * (r + 1) // world
            for it in range(3):
                c.comm_upload_cloud(pts)
                c.step_normals_rsd(R, R, max_nn_normals=max_nn, max_nn_rsd=max_nn)
                p4, pmin, pmax = c.comm_download_range(lo, hi)
                assert _same(p4, n4[lo:hi]) and _same(pmin, rmin[lo:hi]) and _same(pmax, rmax[lo:hi]), f"rank {r}, step {it}"
                gate.wait(timeout=60)
            part = c.comm_download_range(lo, hi)
            inner = c.comm_download_range(lo + 7, hi - 3, normals=False)
            ptr, cnt = c.comm_device_ptr(cab.BUF_NRM_INPUT_RANGE)
            assert ptr and cnt == hi - lo and c.comm_device_ptr(cab.BUF_PERM)[0] == 0
            with pytest.raises(cab.CabError, match="input range"):
                c.comm_download_range(max(lo - 1, 0), hi) if r > 0 else c.comm_download_range(lo, hi + 1)
            out[r] = (part, inner, c.profile())
        except Exception as e:  # noqa: BLE001
            errs.append((r, repr(e)))

    ts = [threading.Thread(target=work, args=(r,)) for r in range(world)]
    for t in ts:
        t.start()
    for t in ts:
        t.join(timeout=120)
    assert not errs, errs
    g4 = np.concatenate([o[0][0] for o in out])
    gmin = np.concatenate([o[0][1] for o in out])
    gmax = np.concatenate([o[0][2] for o in out])
    assert _same(g4, n4) and _same(gmin, rmin) and _same(gmax, rmax)
    for r in range(world):
        lo, hi = n * r // world, n * (r + 1) // world
        assert _same(out[r][1][1], rmin[lo + 7:hi - 3]) and _same(out[r][1][2], rmax[lo + 7:hi - 3])
    if max_nn == 0:
        assert sum(o[2]["neighbour_sum"] for o in out) == k
    for c in ctxs:
        c.close()


def test_results_do_not_depend_on_the_cuts():
    """cab_comm_set_shares moves the cuts (what the measured-time feedback does on clouds large enough to be measured):
    the ranks answer different numbers of queries, the results keep their bits."""
    pts = synth.tabletop(90_000, noise_sigma=0.0002)
    n = pts.shape[0]
    n4, rmin, rmax, k = _single_gpu(pts)
    world = 3
    ctxs = [cab.Context(0) for _ in range(world)]
    cab.comm_init_local(ctxs)
    gate = threading.Barrier(world)
    counts, errs = {}, []

    def work(r):
        try:
            c = ctxs[r]
            c.comm_set_layout(cab.COMM_LAYOUT_INPUT_RANGES)
            lo, hi = n * r // world, n * (r + 1) // world
            for shares in ((1, 1, 1), (0.5, 0.2, 0.3), (0.15, 0.7, 0.15)):
                c.comm_set_shares(shares)
                c.comm_upload_cloud(pts)
                c.step_normals_rsd(R, R)
                p4, pmin, pmax = c.comm_download_range(lo, hi)
                assert _same(p4, n4[lo:hi]) and _same(pmin, rmin[lo:hi]) and _same(pmax, rmax[lo:hi]), f"rank {r}, shares {shares}"
                b, e = c.shard_range()
                counts[(r, shares)] = e - b
                gate.wait(timeout=60)
            with pytest.raises(cab.CabError):
                c.comm_set_shares((1, 2))
        except Exception as e:  # noqa: BLE001
            errs.append((r, repr(e)))

    ts = [threading.Thread(target=work, args=(r,)) for r in range(world)]
    for t in ts:
        t.start()
    for t in ts:
        t.join(timeout=120)
    assert not errs, errs
    assert counts[(0, (0.5, 0.2, 0.3))] > 1.5 * counts[(1, (0.5, 0.2, 0.3))]
    assert counts[(1, (0.15, 0.7, 0.15))] > 2 * counts[(0, (0.15, 0.7, 0.15))]
    for c in ctxs:
        c.close()


def test_local_group_thin_slabs_recompute_their_halos():
    """Eight ranks on a small flat cloud: the slabs are thinner than two layers of rows, so the halo exchange is off and
    every rank recomputes the normals of the rows around its own (shard_mode 1); same results."""
    pts = synth.analytic_shape("plane", 6_000, side=0.3)
    n = pts.shape[0]
    n4, rmin, rmax, k = _single_gpu(pts)
    world = 8
    ctxs = [cab.Context(0) for _ in range(world)]
    cab.comm_init_local(ctxs)
    out, errs = [None] * world, []

    def work(r):
        try:
            c = ctxs[r]
            c.comm_upload_cloud(pts)
            c.step_normals_rsd(R, R)
            out[r] = (c.comm_download_range(0, n), c.profile())
        except Exception as e:  # noqa: BLE001
            errs.append((r, repr(e)))

    ts = [threading.Thread(target=work, args=(r,)) for r in range(world)]
    for t in ts:
        t.start()
    for t in ts:
        t.join(timeout=120)
    assert not errs, errs
    for r in range(world):
        (f4, fmin, fmax), prof = out[r]
        assert _same(f4, n4) and _same(fmin, rmin) and _same(fmax, rmax), f"rank {r}"
        assert prof["shard_mode"] == 1
    assert sum(o[1]["neighbour_sum"] for o in out) == k
    for c in ctxs:
        c.close()


def test_local_group_grsd_allreduce():
    """Cluster-per-rank GRSD: the ranks' integer histograms summed through cab_comm_allreduce_i32."""
    xyz, off = synth.clusters(6, 1300, 2500)
    one = cab.Context(0, exact=True)
    want = one.grsd_batch(xyz, off, 0.025)
    one.close()
    world = 2
    ctxs = [cab.Context(0, exact=True) for _ in range(world)]
    cab.comm_init_local(ctxs)
    got, errs = [None] * world, []

    def work(r):
        try:
            mine = [c for c in range(6) if c % world == r]
            sub = np.concatenate([xyz[off[c]:off[c + 1]] for c in mine])
            soff = np.concatenate([[0], np.cumsum([off[c + 1] - off[c] for c in mine])]).astype(np.int32)
            hist = np.zeros((6, 21), np.int32)
            hist[mine] = ctxs[r].grsd_batch(sub, soff, 0.025)
            got[r] = ctxs[r].comm_allreduce_i32(hist)
        except Exception as e:  # noqa: BLE001
            errs.append((r, repr(e)))

    ts = [threading.Thread(target=work, args=(r,)) for r in range(world)]
    for t in ts:
        t.start()
    for t in ts:
        t.join(timeout=120)
    assert not errs, errs
    for r in range(world):
        assert np.array_equal(got[r], want)
    for c in ctxs:
        c.close()


def _run_processes(world, devices, pts):
    ctx = mp.get_context("spawn")
    pipes, procs = [], []
    for r in range(world):
        a, b = ctx.Pipe()
        p = ctx.Process(target=comm_workers.ipc_worker, args=(r, world, devices[r], pts, b), daemon=True)
        p.start()
        pipes.append(a)
        procs.append(p)
    try:
        blobs = []
        for a in pipes:
            assert a.poll(180), "worker did not report its buffers"
            blobs.append(a.recv())
        for a in pipes:
            a.send(blobs)
        res = []
        for a in pipes:
            assert a.poll(300), "worker did not finish"
            res.append(a.recv())
        for a in pipes:
            a.send("bye")
        return res
    finally:
        for p in procs:
            p.join(timeout=30)
            if p.is_alive():
                p.kill()


def test_ipc_group_two_processes_one_gpu():
    """One process per rank, buffers shared through CUDA IPC handles the application moved (no NCCL)."""
    pts = synth.tabletop(60_000, noise_sigma=0.0002)
    n = pts.shape[0]
    n4, rmin, rmax, _ = _single_gpu(pts)
    res = _run_processes(2, [0, 0], pts)
    for r, (status, part, full) in enumerate(res):
        assert status == "ok", part
        lo, hi = n * r // 2, n * (r + 1) // 2
        p4, pmin, pmax = part
        assert _same(p4, n4[lo:hi]) and _same(pmin, rmin[lo:hi]) and _same(pmax, rmax[lo:hi])
        assert _same(full[1], rmin) and _same(full[2], rmax)


def test_nccl_group_two_gpus():
    """cab_comm_get_id / cab_comm_init: NCCL moves the IPC handles, the pushes cross NVLink, the all-reduce is NCCL's."""
    import torch

    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    pts = synth.tabletop(80_000, noise_sigma=0.0002)
    n4, rmin, rmax, _ = _single_gpu(pts)
    comm_id = cab.comm_get_id()
    ctx = mp.get_context("spawn")
    pipes, procs = [], []
    for r in range(2):
        a, b = ctx.Pipe()
        p = ctx.Process(target=comm_workers.nccl_worker, args=(r, 2, comm_id, pts, b), daemon=True)
        p.start()
        pipes.append(a)
        procs.append(p)
    try:
        for a in pipes:
            assert a.poll(300), "worker did not finish"
            status, full, h = a.recv()
            assert status == "ok", full
            assert _same(full[0], n4) and _same(full[1], rmin) and _same(full[2], rmax)
            assert (h == 3).all()
        for a in pipes:
            a.send("bye")
    finally:
        for p in procs:
            p.join(timeout=30)
            if p.is_alive():
                p.kill()
