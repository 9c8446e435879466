"""ctypes driver of the host-side C++ plugins (mapping-private_b200/host/libcloud_algos.so).

It calls the real C++ classes (cloud_algos::NormalEstimation, LocalRadiusEstimation, GlobalRSD)
through the flat wrapper in host/src/plugin_capi.cpp, in the reference's call order:
init -> pre -> (public fields) -> process -> output -> post.
"""
from __future__ import annotations

import ctypes as C
import pathlib
import subprocess

import numpy as np

_DIR = pathlib.Path(__file__).resolve().parent
LIB_PATH = _DIR / "host" / "libcloud_algos.so"
_LIB = None


def build():
    subprocess.run(["make", "-C", str(_DIR / "host"), "-j8"], check=True, stdout=subprocess.DEVNULL)
    return LIB_PATH


def lib():
    global _LIB
    if _LIB is None:
        if not LIB_PATH.exists():
            raise RuntimeError(f"{LIB_PATH} is missing: run __graft_entry__.build()")
        L = C.CDLL(str(LIB_PATH))
        L.capi_create.restype = C.c_void_p
        L.capi_create.argtypes = [C.c_char_p]
        for f in ("capi_destroy", "capi_pre", "capi_post"):
            getattr(L, f).argtypes = [C.c_void_p]
        L.capi_set_param.argtypes = [C.c_void_p, C.c_char_p, C.c_double]
        L.capi_set_field.argtypes = [C.c_void_p, C.c_char_p, C.c_double]
        L.capi_set_param_str.argtypes = [C.c_void_p, C.c_char_p, C.c_char_p]
        L.capi_set_field_str.argtypes = [C.c_void_p, C.c_char_p, C.c_char_p]
        L.capi_process.restype = C.c_char_p
        L.capi_process.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p]
        L.capi_output_valid.argtypes = [C.c_void_p]
        L.capi_num_published.argtypes = [C.c_void_p]
        L.capi_topic.restype = C.c_char_p
        L.capi_topic.argtypes = [C.c_void_p]
        for f in ("capi_out_size", "capi_out_num_channels"):
            getattr(L, f).argtypes = [C.c_void_p, C.c_int]
        L.capi_out_channel_name.restype = C.c_char_p
        L.capi_out_channel_name.argtypes = [C.c_void_p, C.c_int, C.c_int]
        L.capi_out_channel.restype = C.POINTER(C.c_float)
        L.capi_out_channel.argtypes = [C.c_void_p, C.c_int, C.c_int]
        L.capi_out_points.restype = C.POINTER(C.c_float)
        L.capi_out_points.argtypes = [C.c_void_p, C.c_int]
        L.capi_list_requires.argtypes = [C.c_void_p, C.c_char_p, C.c_int]
        L.capi_pcd_read.argtypes = [C.c_char_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]
        L.capi_pcd_read_rgb.argtypes = [C.c_char_p, C.c_void_p, C.c_int]
        L.capi_write_feature.argtypes = [C.c_char_p, C.c_void_p, C.c_int, C.c_int, C.c_int]
        L.capi_extract_euclidean_clusters.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_double, C.c_int, C.c_uint,
                                                      C.c_void_p, C.c_void_p]
        L.capi_fit_sac_plane.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_double, C.c_int, C.c_uint, C.c_void_p, C.c_void_p]
        _LIB = L
    return _LIB


def fit_sac_plane(xyz: np.ndarray, indices: np.ndarray, threshold: float = 0.03, min_pts: int = 10, seed: int = 1):
    """cloud_tools::fitSACPlane (host/include/cloud_tools/fit_sac_plane.h) through the member's signature.  Returns
    (rc, inliers, coeff (4,), xyz with the inliers projected in place); rc = number of inliers, 0, -1 (too few indices)
    or -2 (library error)."""
    pts = np.array(xyz, np.float32, copy=True, order="C")
    indices = np.ascontiguousarray(indices, np.int32)
    inl = np.zeros(max(len(indices), 1), np.int32)
    coeff = np.zeros(4, np.float64)
    rc = lib().capi_fit_sac_plane(pts.ctypes.data, pts.shape[0], indices.ctypes.data, len(indices), float(threshold), int(min_pts),
                                  int(seed), inl.ctypes.data, coeff.ctypes.data)
    return rc, inl[:max(rc, 0)], coeff, pts


def extract_euclidean_clusters(xyz: np.ndarray, indices: np.ndarray, tolerance: float, min_pts: int = 1, nx_idx: int = -1):
    """cloud_geometry::nearest::extractEuclideanClusters (host/include/point_cloud_mapping/geometry/nearest.h) through its
    reference signature.  Returns the clusters as a list of index arrays, or None if the call reported an error."""
    xyz = np.ascontiguousarray(xyz, np.float32)
    indices = np.ascontiguousarray(indices, np.int32)
    cluster_of = np.full(max(len(indices), 1), -1, np.int32)
    flat = np.zeros(max(len(indices), 1), np.int32)
    nc = lib().capi_extract_euclidean_clusters(xyz.ctypes.data, xyz.shape[0], indices.ctypes.data, len(indices), float(tolerance),
                                               int(nx_idx), int(min_pts), cluster_of.ctypes.data, flat.ctypes.data)
    if nc < 0:
        return None
    sizes = np.bincount(cluster_of[cluster_of >= 0], minlength=nc)
    return np.split(flat[: sizes.sum()], np.cumsum(sizes)[:-1]) if nc else []


def pcd_read(path: str):
    """host/include/cloud_algos/pcd_io.h readPCDXYZ: (xyz (n,3) float32, normals (n,3) | None)."""
    L = lib()
    has = C.c_int(0)
    n = L.capi_pcd_read(str(path).encode(), None, None, 0, C.byref(has))
    if n < 0:
        raise IOError(f"cannot read {path}")
    xyz = np.zeros((max(n, 1), 3), np.float32)
    nrm = np.zeros((max(n, 1), 3), np.float32)
    L.capi_pcd_read(str(path).encode(), xyz.ctypes.data_as(C.POINTER(C.c_float)), nrm.ctypes.data_as(C.POINTER(C.c_float)), n, C.byref(has))
    return xyz[:n], (nrm[:n] if has.value else None)


def pcd_read_rgb(path: str):
    """Packed 0x00RRGGBB colours of a PCD file with an rgb field (uint32 (n,)), None if it has none."""
    L = lib()
    n = L.capi_pcd_read_rgb(str(path).encode(), None, 0)
    if n < 0:
        raise IOError(f"cannot read {path}")
    if n == 0:
        return None
    rgb = np.zeros(n, np.uint32)
    L.capi_pcd_read_rgb(str(path).encode(), rgb.ctypes.data, n)
    return rgb


def write_feature(path: str, feature: np.ndarray, remove_0: bool = True):
    """host/include/cloud_algos/pcd_io.h writeFeature (the reference's `FIELDS vfh` histogram file)."""
    f = np.ascontiguousarray(feature, np.float32)
    if lib().capi_write_feature(str(path).encode(), f.ctypes.data_as(C.POINTER(C.c_float)), f.shape[0], f.shape[1], int(remove_0)) != 0:
        raise IOError(f"cannot write {path}")


class Plugin:
    def __init__(self, lookup_name: str):
        self._L = lib()
        self._h = self._L.capi_create(lookup_name.encode())
        if not self._h:
            raise KeyError(f"pluginlib: no class {lookup_name}")
        self.name = lookup_name

    def close(self):
        if self._h:
            self._L.capi_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_param(self, key: str, value):
        if isinstance(value, str):
            self._L.capi_set_param_str(self._h, key.encode(), value.encode())
        else:
            self._L.capi_set_param(self._h, key.encode(), float(value))

    def set_field(self, field: str, value):
        if isinstance(value, str):
            rc = self._L.capi_set_field_str(self._h, field.encode(), value.encode())
        else:
            rc = self._L.capi_set_field(self._h, field.encode(), float(value))
        if rc != 0:
            raise AttributeError(field)

    def grsd_process_batch(self, xyz: np.ndarray, normals: np.ndarray, offsets: np.ndarray):
        """GlobalRSD::process_batch: all clusters of a frame in one device call.  Returns (result string, hist (nc, 21))."""
        xyz = np.ascontiguousarray(xyz, np.float32)
        nrm = np.ascontiguousarray(normals, np.float32)
        off = np.ascontiguousarray(offsets, np.int32)
        nc = len(off) - 1
        out = np.zeros((max(nc, 1), 21), np.float32)
        self._L.capi_grsd_process_batch.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]
        rc = self._L.capi_grsd_process_batch(self._h, xyz.ctypes.data, nrm.ctypes.data, off.ctypes.data, nc, out.ctypes.data)
        return ("ok" if rc == 0 else "error"), out[:nc]

    def requires_provides(self):
        buf = C.create_string_buffer(1024)
        self._L.capi_list_requires(self._h, buf, 1024)
        req, prov = buf.value.decode().split("|")
        return [x for x in req.split(",") if x], [x for x in prov.split(",") if x]

    def topic(self):
        return self._L.capi_topic(self._h).decode()

    def run(self, xyz: np.ndarray, channels: dict | None = None, fields: dict | None = None):
        """pre(); set fields; process(); output(); post().  Returns (result string, output dict | None)."""
        xyz = np.ascontiguousarray(xyz, dtype=np.float32)
        channels = channels or {}
        names = list(channels)
        vals = [np.ascontiguousarray(channels[k], dtype=np.float32) for k in names]
        n = xyz.shape[0]
        name_arr = (C.c_char_p * max(len(names), 1))(*[k.encode() for k in names])
        val_arr = (C.c_void_p * max(len(names), 1))(*[v.ctypes.data for v in vals])
        self._L.capi_pre(self._h)
        for k, v in (fields or {}).items():
            self.set_field(k, v)
        res = self._L.capi_process(self._h, xyz.ctypes.data, n, len(names), name_arr, val_arr).decode()
        out = self.output(0) if self._L.capi_output_valid(self._h) else None
        self._L.capi_post(self._h)
        return res, out

    def output_valid(self):
        return bool(self._L.capi_output_valid(self._h))

    def num_published(self):
        return self._L.capi_num_published(self._h)

    def output(self, which: int = 0):
        n = self._L.capi_out_size(self._h, which)
        if n < 0:
            return None
        out = {"points": np.ctypeslib.as_array(self._L.capi_out_points(self._h, which), shape=(n, 3)).copy() if n else np.zeros((0, 3), np.float32),
               "channels": {}}
        for c in range(self._L.capi_out_num_channels(self._h, which)):
            name = self._L.capi_out_channel_name(self._h, which, c).decode()
            out["channels"][name] = (np.ctypeslib.as_array(self._L.capi_out_channel(self._h, which, c), shape=(n,)).copy()
                                     if n else np.zeros(0, np.float32))
        return out
