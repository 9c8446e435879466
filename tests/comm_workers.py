"""Rank processes of tests/test_comm.py.  A spawned child imports this module by name, without conftest.py: it finds the
repository through its own path."""
import pathlib
import sys

import numpy as np

ROOT = pathlib.Path(__file__).resolve().parent.parent
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))

R = 0.02


def ipc_worker(rank, world, device, pts, conn):
    try:
        import pkgpath

        pkgpath.load()
        from mapping_private_b200 import cab as cab_

        c = cab_.Context(device)
        conn.send(c.comm_reserve(rank, world, pts.shape[0]))
        c.comm_connect(conn.recv())
        n = pts.shape[0]
        for _ in range(2):
            c.comm_upload_cloud(pts)
            c.step_normals_rsd(R, R)
        lo, hi = n * rank // world, n * (rank + 1) // world
        conn.send(("ok", c.comm_download_range(lo, hi), c.comm_download_range(0, n, normals=False)))
        conn.recv()  # the peers' mappings stay valid until everybody is done
        c.close()
    except Exception as e:  # noqa: BLE001
        conn.send(("error", repr(e), None))


def nccl_worker(rank, world, comm_id, pts, conn):
    try:
        import pkgpath

        pkgpath.load()
        from mapping_private_b200 import cab as cab_

        c = cab_.Context(rank)
        c.comm_init(comm_id, rank, world)
        n = pts.shape[0]
        c.comm_upload_cloud(pts)
        c.step_normals_rsd(R, R)
        h = np.full((4, 21), rank + 1, np.int32)
        c.comm_allreduce_i32(h)
        conn.send(("ok", c.comm_download_range(0, n), h))
        conn.recv()
        c.close()
    except Exception as e:  # noqa: BLE001
        conn.send(("error", repr(e), None))


