"""ctypes binding of libcloudalgos_b200.so (the C ABI in include/cloud_algos_b200.h).

There is no fallback of any kind: if the shared library or a B200 is missing every entry point
raises CabError.  numpy arrays are host buffers; the *_device helpers take raw device pointers.
"""
from __future__ import annotations

import ctypes as C
import pathlib
import subprocess

import numpy as np

_DIR = pathlib.Path(__file__).resolve().parent
LIB_PATH = _DIR / "csrc" / "libcloudalgos_b200.so"
_LIB = None

PFH_USE_DIST, PFH_DIFFERENTIAL, PFH_CHECK_FLIP, PFH_ABS_ANGLES, PFH_AVERAGE, PFH_COMBINE = 1, 2, 4, 8, 16, 32
PFH_DEFAULT = PFH_CHECK_FLIP | PFH_AVERAGE
RSD_SEED_BIN0 = 2
RSD_SCALE_SORT = 4
STEP_INPUT_ORDER = 0x100
SIG_GRSD21, SIG_GRSD325, SIG_PLUSGRSD110 = 0, 1, 2
SIG_DIM = {0: 21, 1: 325, 2: 110}
BUF_POS_SORTED, BUF_NRM_SORTED, BUF_RSD_SORTED, BUF_PERM = 0, 1, 2, 3
BUF_NRM_INPUT_RANGE, BUF_RSD_INPUT_RANGE = 4, 5
COMM_LAYOUT_REPLICATED, COMM_LAYOUT_INPUT_RANGES = 0, 1

EXPORTS = [
    "cab_create", "cab_destroy", "cab_last_error", "cab_upload_cloud", "cab_upload_clusters",
    "cab_set_cloud_device", "cab_build_grid", "cab_set_shard", "cab_shard_range", "cab_normals",
    "cab_set_normals", "cab_rsd", "cab_normals_rsd", "cab_neighbors_debug", "cab_grsd_batch", "cab_grsd_voxels", "cab_grsd_signatures", "cab_color_chlac", "cab_svm_set_model", "cab_svm_set_scaling", "cab_svm_predict", "cab_svm_predict_grsd", "cab_knn_mean_distance", "cab_normals_knn", "cab_statistical_outliers", "cab_euclidean_clusters", "cab_cluster_csr", "cab_pfh",
    "cab_device_ptr", "cab_stream", "cab_download", "cab_download_sorted", "cab_download_rdif", "cab_profile", "cab_version",
    "cab_step_normals_rsd", "cab_comm_get_id", "cab_comm_init", "cab_comm_init_local", "cab_comm_reserve", "cab_comm_connect",
    "cab_comm_free", "cab_comm_upload_cloud", "cab_comm_download_range", "cab_comm_device_ptr", "cab_comm_allreduce_i32",
    "cab_comm_set_layout", "cab_comm_set_feedback", "cab_comm_set_shares", "cab_grsd_cloud", "cab_grsd_cloud_labels", "cab_grsd_cloud_set_labels",
    "cab_fit_plane_msac",
]
COMM_ID_BYTES, COMM_BLOB_BYTES = 128, 512


class CabError(RuntimeError):
    pass


class Config(C.Structure):
    _fields_ = [("device", C.c_int32), ("exact", C.c_int32), ("max_table_cells", C.c_int64)]


class Timings(C.Structure):
    _fields_ = [
        ("build_ms", C.c_float), ("normals_ms", C.c_float), ("rsd_ms", C.c_float), ("grsd_ms", C.c_float),
        ("h2d_ms", C.c_float), ("d2h_ms", C.c_float),
        ("n_points", C.c_int64), ("n_valid", C.c_int64), ("n_packets", C.c_int64), ("n_rows", C.c_int64),
        ("n_cells", C.c_int64), ("neighbour_sum", C.c_int64), ("candidate_sum", C.c_int64),
        ("kernel_launches", C.c_int64), ("n_sorted", C.c_int64), ("knn_ms", C.c_float), ("knn_rounds", C.c_int32), ("pfh_ms", C.c_float), ("cluster_ms", C.c_float),
        ("exchange_ms", C.c_float), ("step_ms", C.c_float), ("shard_mode", C.c_int32),
    ]

    def as_dict(self):
        return {k: getattr(self, k) for k, _ in self._fields_}


def build(force: bool = False) -> pathlib.Path:
    """Compile the library for sm_100a (nvcc cross-compiles without a GPU)."""
    cmd = ["make", "-C", str(_DIR / "csrc"), "-j8", "libcloudalgos_b200.so"]
    if force:
        cmd.insert(1, "-B")
    subprocess.run(cmd, check=True, stdout=subprocess.DEVNULL)
    return LIB_PATH


def lib():
    global _LIB
    if _LIB is None:
        if not LIB_PATH.exists():
            raise CabError(f"{LIB_PATH} is missing: run __graft_entry__.build() (there is no CPU fallback)")
        L = C.CDLL(str(LIB_PATH))
        L.cab_last_error.restype = C.c_char_p
        L.cab_last_error.argtypes = [C.c_void_p]
        L.cab_neighbors_debug.restype = C.c_int64
        L.cab_grsd_voxels.restype = C.c_int64
        L.cab_grsd_signatures.restype = C.c_int64
        L.cab_grsd_cloud_labels.restype = C.c_int64
        L.cab_fit_plane_msac.restype = C.c_int64
        L.cab_color_chlac.restype = C.c_int64
        L.cab_statistical_outliers.restype = C.c_int64
        L.cab_euclidean_clusters.restype = C.c_int64
        L.cab_cluster_csr.restype = C.c_int64
        L.cab_device_ptr.restype = C.c_void_p
        L.cab_device_ptr.argtypes = [C.c_void_p, C.c_int32]
        L.cab_stream.restype = C.c_void_p
        L.cab_stream.argtypes = [C.c_void_p]
        L.cab_destroy.argtypes = [C.c_void_p]
        L.cab_comm_device_ptr.restype = C.c_void_p
        L.cab_comm_device_ptr.argtypes = [C.c_void_p, C.c_int32, C.POINTER(C.c_int64)]
        _LIB = L
    return _LIB


def _fp(a):
    return a.ctypes.data_as(C.POINTER(C.c_float)) if a is not None else None


def _ip(a):
    return a.ctypes.data_as(C.POINTER(C.c_int32)) if a is not None else None


class Context:
    """One cab_ctx: a CUDA stream plus a grow-only device arena on one GPU."""

    def __init__(self, device: int = 0, exact: bool = False, max_table_cells: int = 0):
        self._L = lib()
        self._h = C.c_void_p()
        cfg = Config(device, 1 if exact else 0, max_table_cells)
        rc = self._L.cab_create(C.byref(cfg), C.byref(self._h))
        if rc != 0:
            raise CabError(f"cab_create failed ({rc}): {self._L.cab_last_error(None).decode()}")
        self.n = 0

    def close(self):
        if self._h:
            self._L.cab_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc, what):
        if rc < 0:
            raise CabError(f"{what} failed ({rc}): {self._L.cab_last_error(self._h).decode()}")
        return rc

    # ---- cloud / grid ----------------------------------------------------------------
    def upload(self, xyz: np.ndarray):
        xyz = np.ascontiguousarray(xyz, dtype=np.float32)
        assert xyz.ndim == 2 and xyz.shape[1] >= 3
        self._keep = xyz
        self.n = xyz.shape[0]
        self._check(self._L.cab_upload_cloud(self._h, _fp(xyz), C.c_int64(self.n), C.c_int32(xyz.shape[1])), "cab_upload_cloud")

    def upload_clusters(self, xyz: np.ndarray, offsets: np.ndarray):
        xyz = np.ascontiguousarray(xyz, dtype=np.float32)
        offsets = np.ascontiguousarray(offsets, dtype=np.int32)
        self._keep = (xyz, offsets)
        self.n = xyz.shape[0]
        self._check(self._L.cab_upload_clusters(self._h, _fp(xyz), C.c_int64(self.n), C.c_int32(xyz.shape[1]),
                                                _ip(offsets), C.c_int32(len(offsets) - 1)), "cab_upload_clusters")

    def set_cloud_device(self, dptr: int, n: int, stride: int = 3):
        self.n = n
        self._check(self._L.cab_set_cloud_device(self._h, C.c_void_p(dptr), C.c_int64(n), C.c_int32(stride)),
                    "cab_set_cloud_device")

    def build_grid(self, cell: float):
        self._check(self._L.cab_build_grid(self._h, C.c_float(cell)), "cab_build_grid")

    def set_shard(self, rank: int, world: int):
        self._check(self._L.cab_set_shard(self._h, C.c_int32(rank), C.c_int32(world)), "cab_set_shard")

    def shard_range(self):
        b, e = C.c_int64(), C.c_int64()
        self._check(self._L.cab_shard_range(self._h, C.byref(b), C.byref(e)), "cab_shard_range")
        return b.value, e.value

    # ---- stages ----------------------------------------------------------------------
    def normals(self, r: float, max_nn: int = 0, vp=(0.0, 0.0, 0.0), download: bool = True):
        v = np.asarray(vp, dtype=np.float32)
        out = np.empty((self.n, 4), np.float32) if download else None
        self._check(self._L.cab_normals(self._h, C.c_float(r), C.c_int32(max_nn), _fp(v), _fp(out)), "cab_normals")
        return out

    def set_normals(self, nrm: np.ndarray):
        nrm = np.asarray(nrm, dtype=np.float32)
        nx, ny, nz = (np.ascontiguousarray(nrm[:, i]) for i in range(3))
        self._check(self._L.cab_set_normals(self._h, _fp(nx), _fp(ny), _fp(nz)), "cab_set_normals")

    def rsd(self, r: float, max_nn: int = 0, ndiv: int = 10, plane_radius: float = 0.1, flags: int = 0,
            download: bool = True):
        rmin = np.empty(self.n, np.float32) if download else None
        rmax = np.empty(self.n, np.float32) if download else None
        self._check(self._L.cab_rsd(self._h, C.c_double(r), C.c_int32(max_nn), C.c_int32(ndiv), C.c_double(plane_radius),
                                    C.c_int32(flags), _fp(rmin), _fp(rmax)), "cab_rsd")
        return rmin, rmax

    def normals_rsd(self, r: float, max_nn_normals: int = 0, vp=(0.0, 0.0, 0.0), max_nn_rsd: int = 0, ndiv: int = 10,
                    plane_radius: float = 0.1, flags: int = 0, sorted_shard: bool = False):
        """Both passes in one call (normals leave on the copy stream while RSD runs).
        Input order: (n4, r_min, r_max); sorted_shard: (n4 (m,4), radii (m,2), input_index (m,))."""
        v = np.asarray(vp, dtype=np.float32)
        if sorted_shard:
            b, e = self.shard_range()
            n4 = np.empty((e - b, 4), np.float32)
            a = np.empty((e - b, 2), np.float32)
            bb, idx = None, np.empty(e - b, np.int32)
        else:
            n4 = np.empty((self.n, 4), np.float32)
            a, bb, idx = np.empty(self.n, np.float32), np.empty(self.n, np.float32), None
        self._check(self._L.cab_normals_rsd(self._h, C.c_double(r), C.c_int32(max_nn_normals), _fp(v), C.c_int32(max_nn_rsd),
                                            C.c_int32(ndiv), C.c_double(plane_radius), C.c_int32(flags),
                                            C.c_int32(1 if sorted_shard else 0), _fp(n4), _fp(a), _fp(bb), _ip(idx)),
                    "cab_normals_rsd")
        return (n4, a, idx) if sorted_shard else (n4, a, bb)

    def download(self, normals: bool = True, rsd: bool = True):
        n4 = np.empty((self.n, 4), np.float32) if normals else None
        rmin = np.empty(self.n, np.float32) if rsd else None
        rmax = np.empty(self.n, np.float32) if rsd else None
        self._check(self._L.cab_download(self._h, _fp(n4), _fp(rmin), _fp(rmax)), "cab_download")
        return n4, rmin, rmax

    def download_sorted(self, begin: int, end: int, normals: bool = True, rsd: bool = True):
        """Sorted-order slice [begin, end): (n4 (m,4) | None, radii (m,2) | None, input_index (m,))."""
        m = end - begin
        n4 = np.empty((m, 4), np.float32) if normals else None
        rr = np.empty((m, 2), np.float32) if rsd else None
        idx = np.empty(m, np.int32)
        self._check(self._L.cab_download_sorted(self._h, C.c_int64(begin), C.c_int64(end), _fp(n4), _fp(rr), _ip(idx)),
                    "cab_download_sorted")
        return n4, rr, idx

    def neighbors(self, r: float, q0: int, q1: int, max_nn: int = 0):
        """Neighbour sets of queries [q0, q1): (offsets int64, idx int32, d2 float32), unsorted."""
        off = np.zeros(q1 - q0 + 1, np.int64)
        total = self._check(self._L.cab_neighbors_debug(self._h, C.c_float(r), C.c_int32(max_nn), C.c_int64(q0), C.c_int64(q1),
                                                        off.ctypes.data_as(C.POINTER(C.c_int64)), None, None, C.c_int64(0)),
                            "cab_neighbors_debug")
        idx = np.zeros(max(total, 1), np.int32)
        d2 = np.zeros(max(total, 1), np.float32)
        self._check(self._L.cab_neighbors_debug(self._h, C.c_float(r), C.c_int32(max_nn), C.c_int64(q0), C.c_int64(q1),
                                                off.ctypes.data_as(C.POINTER(C.c_int64)), _ip(idx), _fp(d2), C.c_int64(total)),
                    "cab_neighbors_debug")
        return off, idx[:total], d2[:total]

    def grsd_batch(self, xyz: np.ndarray, offsets: np.ndarray, leaf: float, r_normals: float = 0.02,
                   rsd_radius_min: float = 0.01, rsd_flags: int = 0, vp=(0.0, 0.0, 0.0), normals: np.ndarray | None = None):
        xyz = np.ascontiguousarray(xyz, dtype=np.float32)
        offsets = np.ascontiguousarray(offsets, dtype=np.int32)
        nc = len(offsets) - 1
        self.n = xyz.shape[0]
        hist = np.zeros((nc, 21), np.int32)
        v = np.asarray(vp, dtype=np.float32)
        nx = ny = nz = None
        if normals is not None:
            nrm = np.asarray(normals, dtype=np.float32)
            nx, ny, nz = (np.ascontiguousarray(nrm[:, i]) for i in range(3))
        self._check(self._L.cab_grsd_batch(self._h, _fp(xyz), C.c_int32(xyz.shape[1]), _ip(offsets), C.c_int32(nc),
                                           C.c_float(leaf), C.c_float(r_normals), C.c_double(rsd_radius_min),
                                           C.c_int32(rsd_flags), _fp(v), _fp(nx), _fp(ny), _fp(nz), _ip(hist)), "cab_grsd_batch")
        return hist

    def fit_plane_msac(self, xyz: np.ndarray, triples: np.ndarray, indices: np.ndarray | None = None, threshold: float = 0.03,
                       max_iterations: int = 500, probability: float = 0.99):
        """fitSACPlane: dict(coeff (4,) float64, inliers int32, projected (m,3) float32, iterations, best_iteration)."""
        xyz = np.ascontiguousarray(xyz, dtype=np.float32)
        tri = np.ascontiguousarray(triples, dtype=np.int32).reshape(-1, 3)
        idx = None if indices is None else np.ascontiguousarray(indices, dtype=np.int32)
        m = xyz.shape[0] if idx is None else idx.shape[0]
        coeff = np.zeros(4, np.float64)
        inl = np.zeros(max(m, 1), np.int32)
        proj = np.zeros((max(m, 1), 3), np.float32)
        it, best = C.c_int32(), C.c_int32()
        nin = self._check(self._L.cab_fit_plane_msac(
            self._h, _fp(xyz), C.c_int64(xyz.shape[0]), C.c_int32(xyz.shape[1]), _ip(idx), C.c_int64(0 if idx is None else idx.shape[0]),
            C.c_double(threshold), C.c_int32(max_iterations), C.c_double(probability), _ip(tri), C.c_int64(tri.shape[0]),
            coeff.ctypes.data_as(C.POINTER(C.c_double)), _ip(inl), _fp(proj), C.c_int64(m), C.byref(it), C.byref(best)), "cab_fit_plane_msac")
        return dict(coeff=coeff, inliers=inl[:nin], projected=proj[:nin], iterations=it.value, best_iteration=best.value)

    def grsd_cloud(self, leaf: float, r_normals: float = 0.02, rsd_radius_min: float = 0.01, rsd_flags: int = 0, vp=(0.0, 0.0, 0.0)):
        """GRSD-21 of the one cloud this context holds; in a group: sharded by rows, labels and histogram merged (21 int32)."""
        hist = np.zeros(21, np.int32)
        v = np.asarray(vp, dtype=np.float32)
        self._check(self._L.cab_grsd_cloud(self._h, C.c_float(leaf), C.c_float(r_normals), C.c_double(rsd_radius_min),
                                           C.c_int32(rsd_flags), _fp(v), _ip(hist)), "cab_grsd_cloud")
        return hist

    def grsd_cloud_labels(self, leaf: float, r_normals: float = 0.02, rsd_radius_min: float = 0.01, rsd_flags: int = 0,
                          vp=(0.0, 0.0, 0.0)):
        """Step 1 of grsd_cloud alone: label + 1 of this rank's voxels, 0 for the others."""
        v = np.asarray(vp, dtype=np.float32)
        args = (self._h, C.c_float(leaf), C.c_float(r_normals), C.c_double(rsd_radius_min), C.c_int32(rsd_flags), _fp(v))
        self._check(self._L.cab_grsd_cloud_labels(*args, None, C.c_int64(0)), "cab_grsd_cloud_labels")
        return self.grsd_voxels(1)["labels"] + 1  # the voxels of the other ranks carry label -1

    def grsd_cloud_set_labels(self, labels_plus1: np.ndarray):
        lab = np.ascontiguousarray(labels_plus1, np.int32)
        self._check(self._L.cab_grsd_cloud_set_labels(self._h, _ip(lab), C.c_int64(lab.size)), "cab_grsd_cloud_set_labels")

    def grsd_voxels(self, nclusters: int):
        off = np.zeros(nclusters + 1, np.int64)
        nv = self._check(self._L.cab_grsd_voxels(self._h, off.ctypes.data_as(C.POINTER(C.c_int64)), None, None, None, None,
                                                 C.c_int64(0)), "cab_grsd_voxels")
        cent = np.zeros((max(nv, 1), 3), np.float32)
        rmin = np.zeros(max(nv, 1), np.float32)
        rmax = np.zeros(max(nv, 1), np.float32)
        lab = np.zeros(max(nv, 1), np.int32)
        self._check(self._L.cab_grsd_voxels(self._h, off.ctypes.data_as(C.POINTER(C.c_int64)), _fp(cent), _fp(rmin), _fp(rmax),
                                            _ip(lab), C.c_int64(nv)), "cab_grsd_voxels")
        return dict(offsets=off, centroids=cent[:nv], r_min=rmin[:nv], r_max=rmax[:nv], labels=lab[:nv])

    def grsd_signatures(self, nclusters: int, kind: int = SIG_GRSD21, subdivision_size: int = 0, off=(0, 0, 0)):
        """Signatures of the clusters of the last grsd_batch: dict(offsets (nc+1), subdiv_b (nc,3), hist (total, dim))."""
        offs = np.zeros(nclusters + 1, np.int64)
        sb = np.zeros((nclusters, 3), np.int32)
        args = (self._h, C.c_int32(kind), C.c_int32(subdivision_size), C.c_int32(off[0]), C.c_int32(off[1]), C.c_int32(off[2]),
                offs.ctypes.data_as(C.POINTER(C.c_int64)), _ip(sb))
        total = self._check(self._L.cab_grsd_signatures(*args, None, C.c_int64(0)), "cab_grsd_signatures")
        hist = np.zeros((total, SIG_DIM[kind]), np.int32)
        if total:
            self._check(self._L.cab_grsd_signatures(*args, _ip(hist), C.c_int64(total)), "cab_grsd_signatures")
        return dict(offsets=offs, subdiv_b=sb, hist=hist)

    def color_chlac(self, nclusters: int, rgb: np.ndarray, c3: bool = True, thr=(127, 127, 127), subdivision_size: int = 0,
                    off=(0, 0, 0)):
        """Colour half of VOSCH for the clusters of the last grsd_batch (rgb: packed 0x00RRGGBB per point of that batch):
        dict(offsets (nc+1), subdiv_b (nc,3), hist (total, 117) float32)."""
        col = np.ascontiguousarray(rgb, np.uint32)
        offs = np.zeros(nclusters + 1, np.int64)
        sb = np.zeros((nclusters, 3), np.int32)
        args = (self._h, col.ctypes.data_as(C.POINTER(C.c_uint32)), C.c_int32(1 if c3 else 0), C.c_int32(thr[0]), C.c_int32(thr[1]),
                C.c_int32(thr[2]), C.c_int32(subdivision_size), C.c_int32(off[0]), C.c_int32(off[1]), C.c_int32(off[2]),
                offs.ctypes.data_as(C.POINTER(C.c_int64)), _ip(sb))
        total = self._check(self._L.cab_color_chlac(*args, None, C.c_int64(0)), "cab_color_chlac")
        hist = np.zeros((total, 117), np.float32)
        if total:
            self._check(self._L.cab_color_chlac(*args, _fp(hist), C.c_int64(total)), "cab_color_chlac")
        return dict(offsets=offs, subdiv_b=sb, hist=hist)

    # ---- point feature histograms -----------------------------------------------------
    def pfh(self, radius: float = 0.03, max_nn: int = 100, quantum: int = 9, flags: int = PFH_DEFAULT):
        nf = 4 if flags & PFH_USE_DIST else 3
        nb = quantum ** nf if flags & PFH_COMBINE else quantum * nf
        out = np.zeros((self.n, nb), np.float32)
        self._check(self._L.cab_pfh(self._h, C.c_double(radius), C.c_int32(max_nn), C.c_int32(quantum), C.c_int32(flags), _fp(out)),
                    "cab_pfh")
        return out

    # ---- statistical outlier removal ---------------------------------------------------
    def knn_mean_distance(self, k: int, cell_hint: float = 0.0):
        avg = np.zeros(self.n, np.float64)
        self._check(self._L.cab_knn_mean_distance(self._h, C.c_int32(k), C.c_float(cell_hint),
                                                  avg.ctypes.data_as(C.POINTER(C.c_double))), "cab_knn_mean_distance")
        return avg

    def normals_knn(self, k: int, vp=(0.0, 0.0, 0.0), cell_hint: float = 0.0):
        """k-nearest-neighbour normals (n, 4) in input order: nx, ny, nz, curvature."""
        out = np.zeros((self.n, 4), np.float32)
        v = (C.c_float * 3)(*vp)
        self._check(self._L.cab_normals_knn(self._h, C.c_int32(k), v, C.c_float(cell_hint), _fp(out)), "cab_normals_knn")
        return out

    def statistical_outliers(self, k: int = 10, alpha: float = 3.0, cell_hint: float = 0.0):
        """Returns dict(keep bool (n,), avg float64 (n,), mean, stddev, kept)."""
        keep = np.zeros(self.n, np.uint8)
        avg = np.zeros(self.n, np.float64)
        mean, std = C.c_double(), C.c_double()
        kept = self._check(self._L.cab_statistical_outliers(self._h, C.c_int32(k), C.c_double(alpha), C.c_float(cell_hint),
                                                            keep.ctypes.data_as(C.POINTER(C.c_uint8)),
                                                            avg.ctypes.data_as(C.POINTER(C.c_double)), C.byref(mean), C.byref(std)),
                           "cab_statistical_outliers")
        return dict(keep=keep.astype(bool), avg=avg, mean=mean.value, stddev=std.value, kept=int(kept))

    # ---- Euclidean clustering ------------------------------------------------------------
    def euclidean_clusters(self, tolerance: float, min_pts: int = 1, max_pts: int = 0):
        """Connected components of "d2 <= tolerance^2" on the uploaded cloud.  Returns (labels int32 (n,), n_clusters):
        clusters numbered by their smallest index, -1 = dropped or non-finite."""
        labels = np.full(self.n, -1, np.int32)
        nc = self._check(self._L.cab_euclidean_clusters(self._h, C.c_double(tolerance), C.c_int32(min_pts), C.c_int32(max_pts),
                                                        labels.ctypes.data_as(C.POINTER(C.c_int32))), "cab_euclidean_clusters")
        return labels, int(nc)

    def cluster_csr(self, labels: np.ndarray, n_clusters: int):
        """labels -> (offsets int32 (n_clusters + 1,), indices int32): every cluster's point indices, ascending."""
        labels = np.ascontiguousarray(labels, np.int32)
        offsets = np.zeros(n_clusters + 1, np.int32)
        indices = np.zeros(max(int((labels >= 0).sum()), 1), np.int32)
        total = self._L.cab_cluster_csr(labels.ctypes.data_as(C.POINTER(C.c_int32)), C.c_int64(labels.shape[0]), C.c_int32(n_clusters),
                                        offsets.ctypes.data_as(C.POINTER(C.c_int32)), indices.ctypes.data_as(C.POINTER(C.c_int32)))
        if total < 0:
            raise CabError(f"cab_cluster_csr failed ({total})")
        return offsets, indices[:total]

    # ---- SVM ---------------------------------------------------------------------------
    def svm_set_model(self, model, scale=None):
        """model: svm_model.SvmModel; scale: None or (lower, upper, fmin, fmax) from svm_model.parse_scale."""
        dp = C.POINTER(C.c_double)
        lab = np.ascontiguousarray(model.labels, np.int32)
        nsv = np.ascontiguousarray(model.nr_sv, np.int32)
        rho = np.ascontiguousarray(model.rho, np.float64)
        coef = np.ascontiguousarray(model.sv_coef, np.float64)
        sv = np.ascontiguousarray(model.sv, np.float64)
        self._check(self._L.cab_svm_set_model(self._h, C.c_int32(sv.shape[1]), C.c_int32(lab.shape[0]), C.c_int32(sv.shape[0]),
                                              C.c_double(model.gamma), _ip(lab), _ip(nsv), rho.ctypes.data_as(dp),
                                              coef.ctypes.data_as(dp), sv.ctypes.data_as(dp)), "cab_svm_set_model")
        if scale is None:
            self._check(self._L.cab_svm_set_scaling(self._h, C.c_int32(sv.shape[1]), C.c_double(0), C.c_double(0), None, None),
                        "cab_svm_set_scaling")
        else:
            lower, upper, fmin, fmax = scale
            fmin = np.ascontiguousarray(fmin, np.float64)
            fmax = np.ascontiguousarray(fmax, np.float64)
            self._check(self._L.cab_svm_set_scaling(self._h, C.c_int32(fmin.shape[0]), C.c_double(lower), C.c_double(upper),
                                                    fmin.ctypes.data_as(dp), fmax.ctypes.data_as(dp)), "cab_svm_set_scaling")
        self._svm_pairs = lab.shape[0] * (lab.shape[0] - 1) // 2

    def svm_predict(self, features: np.ndarray, want_dec: bool = False):
        f = np.ascontiguousarray(features, np.float32)
        out = np.zeros(f.shape[0], np.float32)
        dec = np.zeros((f.shape[0], self._svm_pairs), np.float64) if want_dec else None
        self._check(self._L.cab_svm_predict(self._h, _fp(f), C.c_int64(f.shape[0]), C.c_int32(f.shape[1]), _fp(out),
                                            dec.ctypes.data_as(C.POINTER(C.c_double)) if want_dec else None), "cab_svm_predict")
        return (out, dec) if want_dec else out

    def svm_predict_grsd(self, nclusters: int):
        out = np.zeros(nclusters, np.float32)
        self._check(self._L.cab_svm_predict_grsd(self._h, _fp(out)), "cab_svm_predict_grsd")
        return out

    # ---- one step, one call / multi-GPU groups ------------------------------------------
    def step_normals_rsd(self, cell: float, r: float, max_nn_normals: int = 0, vp=(0.0, 0.0, 0.0), max_nn_rsd: int = 0,
                         ndiv: int = 10, plane_radius: float = 0.1, flags: int = 0):
        """Grid build + normals + RSD with one host synchronisation; in a group also the concatenation of the results."""
        v = (C.c_float * 3)(*vp)
        self._check(self._L.cab_step_normals_rsd(self._h, C.c_float(cell), C.c_double(r), C.c_int32(max_nn_normals), v,
                                                 C.c_int32(max_nn_rsd), C.c_int32(ndiv), C.c_double(plane_radius), C.c_int32(flags)),
                    "cab_step_normals_rsd")

    def download_rdif(self):
        out = np.empty(self.n, np.float32)
        self._check(self._L.cab_download_rdif(self._h, _fp(out)), "cab_download_rdif")
        return out

    def comm_init(self, comm_id: bytes, rank: int, world: int):
        assert len(comm_id) == COMM_ID_BYTES
        self._check(self._L.cab_comm_init(self._h, C.c_char_p(comm_id), C.c_int32(rank), C.c_int32(world)), "cab_comm_init")

    def comm_reserve(self, rank: int, world: int, max_points: int) -> bytes:
        blob = C.create_string_buffer(COMM_BLOB_BYTES)
        self._check(self._L.cab_comm_reserve(self._h, C.c_int32(rank), C.c_int32(world), C.c_int64(max_points), blob), "cab_comm_reserve")
        return blob.raw

    def comm_connect(self, blobs):
        raw = b"".join(blobs)
        self._check(self._L.cab_comm_connect(self._h, C.c_char_p(raw)), "cab_comm_connect")

    def comm_free(self):
        self._check(self._L.cab_comm_free(self._h), "cab_comm_free")

    def comm_set_feedback(self, on: bool):
        self._check(self._L.cab_comm_set_feedback(self._h, C.c_int32(1 if on else 0)), "cab_comm_set_feedback")

    def comm_set_shares(self, shares):
        a = np.ascontiguousarray(shares, np.float64)
        self._check(self._L.cab_comm_set_shares(self._h, a.ctypes.data_as(C.POINTER(C.c_double)), C.c_int32(a.size)), "cab_comm_set_shares")

    def comm_set_layout(self, layout: int):
        self._check(self._L.cab_comm_set_layout(self._h, C.c_int32(layout)), "cab_comm_set_layout")

    def comm_upload_cloud(self, xyz: np.ndarray):
        xyz = np.ascontiguousarray(xyz, dtype=np.float32)
        self._keep = xyz
        self.n = xyz.shape[0]
        self._check(self._L.cab_comm_upload_cloud(self._h, _fp(xyz), C.c_int64(self.n), C.c_int32(xyz.shape[1])), "cab_comm_upload_cloud")

    def comm_download_range(self, j0: int, j1: int, normals: bool = True, rsd: bool = True):
        m = j1 - j0
        n4 = np.empty((m, 4), np.float32) if normals else None
        rmin = np.empty(m, np.float32) if rsd else None
        rmax = np.empty(m, np.float32) if rsd else None
        self._check(self._L.cab_comm_download_range(self._h, C.c_int64(j0), C.c_int64(j1), _fp(n4), _fp(rmin), _fp(rmax)),
                    "cab_comm_download_range")
        return n4, rmin, rmax

    def comm_device_ptr(self, which: int):
        cnt = C.c_int64()
        p = self._L.cab_comm_device_ptr(self._h, which, C.byref(cnt)) or 0
        return p, cnt.value

    def comm_allreduce_i32(self, values: np.ndarray):
        assert values.dtype == np.int32 and values.flags.c_contiguous
        self._check(self._L.cab_comm_allreduce_i32(self._h, _ip(values), C.c_int64(values.size)), "cab_comm_allreduce_i32")
        return values

    # ---- plumbing --------------------------------------------------------------------
    def device_ptr(self, which: int) -> int:
        return self._L.cab_device_ptr(self._h, which) or 0

    def stream(self) -> int:
        return self._L.cab_stream(self._h) or 0

    def profile(self) -> dict:
        t = Timings()
        self._check(self._L.cab_profile(self._h, C.byref(t)), "cab_profile")
        return t.as_dict()


def comm_get_id() -> bytes:
    """ncclGetUniqueId through the C ABI: rank 0 calls it, the application hands the 128 bytes to every rank."""
    buf = C.create_string_buffer(COMM_ID_BYTES)
    rc = lib().cab_comm_get_id(buf)
    if rc != 0:
        raise CabError(f"cab_comm_get_id failed ({rc}): {lib().cab_last_error(None).decode()}")
    return buf.raw


def comm_init_local(contexts):
    """Links the contexts of this process into one group (rank = position in the list)."""
    arr = (C.c_void_p * len(contexts))(*[c._h for c in contexts])
    rc = lib().cab_comm_init_local(arr, C.c_int32(len(contexts)))
    if rc != 0:
        raise CabError(f"cab_comm_init_local failed ({rc})")
