"""ctypes binding of the CPU oracle (oracle/liboracle.so).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline / --impl reference legs -- never by the product path.
"""
from __future__ import annotations

import ctypes as C
import pathlib
import subprocess

import numpy as np

_DIR = pathlib.Path(__file__).resolve().parent
_LIB = None

RSD_REF_IS_NEAREST = 1
RSD_SEED_BIN0 = 2
RSD_SCALE_SORT = 4

_f32p = np.ctypeslib.ndpointer(np.float32, flags="C_CONTIGUOUS")
_i32p = np.ctypeslib.ndpointer(np.int32, flags="C_CONTIGUOUS")
_i64p = np.ctypeslib.ndpointer(np.int64, flags="C_CONTIGUOUS")


def build(force: bool = False) -> pathlib.Path:
    so = _DIR / "liboracle.so"
    src = [_DIR / "oracle.cpp", _DIR / "oracle.h"]
    if force or not so.exists() or any(s.stat().st_mtime > so.stat().st_mtime for s in src):
        subprocess.run(["make", "-C", str(_DIR), "-B" if force else "-s", "liboracle.so"], check=True)
    return so


def lib():
    global _LIB
    if _LIB is None:
        so = _DIR / "liboracle.so"
        if not so.exists():
            build()
        _LIB = C.CDLL(str(so))
        _LIB.orc_radius_search.restype = C.c_int64
        _LIB.orc_radius_search_brute.restype = C.c_int64
        _LIB.orc_d2.restype = C.c_float
    return _LIB


def _xyz(a):
    a = np.ascontiguousarray(a, dtype=np.float32)
    assert a.ndim == 2 and a.shape[1] == 3
    return a


def _ptr(a, ty):
    return a.ctypes.data_as(C.POINTER(ty)) if a is not None else None


def radius_search(surface, queries, r, max_nn=0, brute=False, nthreads=0):
    """Returns (offsets int64 (nq+1), idx int32, d2 float32), sorted by (d2, idx)."""
    L = lib()
    s, q = _xyz(surface), _xyz(queries)
    off = np.zeros(q.shape[0] + 1, dtype=np.int64)
    if brute:
        total = L.orc_radius_search_brute(_ptr(s, C.c_float), s.shape[0], _ptr(q, C.c_float), q.shape[0],
                                          C.c_double(r), int(max_nn), _ptr(off, C.c_int64), None, None, C.c_int64(0))
    else:
        total = L.orc_radius_search(_ptr(s, C.c_float), s.shape[0], _ptr(q, C.c_float), q.shape[0], C.c_double(r),
                                    int(max_nn), _ptr(off, C.c_int64), None, None, C.c_int64(0), int(nthreads))
    idx = np.zeros(max(total, 1), dtype=np.int32)
    d2 = np.zeros(max(total, 1), dtype=np.float32)
    if brute:
        L.orc_radius_search_brute(_ptr(s, C.c_float), s.shape[0], _ptr(q, C.c_float), q.shape[0], C.c_double(r),
                                  int(max_nn), _ptr(off, C.c_int64), _ptr(idx, C.c_int32), _ptr(d2, C.c_float),
                                  C.c_int64(total))
    else:
        L.orc_radius_search(_ptr(s, C.c_float), s.shape[0], _ptr(q, C.c_float), q.shape[0], C.c_double(r),
                            int(max_nn), _ptr(off, C.c_int64), _ptr(idx, C.c_int32), _ptr(d2, C.c_float),
                            C.c_int64(total), int(nthreads))
    return off, idx[:total], d2[:total]


def normals(xyz, r, max_nn=0, vp=(0.0, 0.0, 0.0), nthreads=0):
    """Returns (n4 float32 (n,4) = nx,ny,nz,curvature ; k int32 (n,))."""
    L = lib()
    p = _xyz(xyz)
    out = np.empty((p.shape[0], 4), dtype=np.float32)
    k = np.empty(p.shape[0], dtype=np.int32)
    v = np.asarray(vp, dtype=np.float32)
    L.orc_normals(_ptr(p, C.c_float), p.shape[0], C.c_double(r), int(max_nn), _ptr(v, C.c_float),
                  _ptr(out, C.c_float), _ptr(k, C.c_int32), int(nthreads))
    return out, k


def normals_gap(xyz, r, max_nn=0, vp=(0.0, 0.0, 0.0), nthreads=0):
    """Returns (n4 (n,4), gap (n,)): the normals of normals() and their conditioning (l1 - l0) / trace."""
    L = lib()
    p = _xyz(xyz)
    out = np.empty((p.shape[0], 4), dtype=np.float32)
    gap = np.empty(p.shape[0], dtype=np.float32)
    v = np.asarray(vp, dtype=np.float32)
    L.orc_normals_gap(_ptr(p, C.c_float), p.shape[0], C.c_double(r), int(max_nn), _ptr(v, C.c_float),
                      _ptr(out, C.c_float), _ptr(gap, C.c_float), int(nthreads))
    return out, gap


def rsd(xyz, nrm, r, max_nn=0, ndiv=10, plane_radius=0.1, flags=0, nthreads=0):
    """In-tree LocalRadiusEstimation arithmetic. nrm: (n,3) or (n,4). Returns r_min, r_max, r_dif."""
    L = lib()
    p = _xyz(xyz)
    nn = np.ascontiguousarray(nrm, dtype=np.float32)
    n = p.shape[0]
    rmin = np.empty(n, np.float32)
    rmax = np.empty(n, np.float32)
    rdif = np.empty(n, np.float32)
    rc = L.orc_rsd(_ptr(p, C.c_float), _ptr(nn, C.c_float), int(nn.shape[1]), n, C.c_double(r), int(max_nn),
                   int(ndiv), C.c_double(plane_radius), int(flags), _ptr(rmin, C.c_float), _ptr(rmax, C.c_float),
                   _ptr(rdif, C.c_float), int(nthreads))
    assert rc == 0
    return rmin, rmax, rdif


def rsd_queries(surface, nrm, queries, r, max_nn=0, ndiv=5, plane_radius=0.2, flags=0, nthreads=0):
    L = lib()
    p, q = _xyz(surface), _xyz(queries)
    nn = np.ascontiguousarray(nrm, dtype=np.float32)
    rmin = np.empty(q.shape[0], np.float32)
    rmax = np.empty(q.shape[0], np.float32)
    rc = L.orc_rsd_queries(_ptr(p, C.c_float), _ptr(nn, C.c_float), int(nn.shape[1]), p.shape[0],
                           _ptr(q, C.c_float), q.shape[0], C.c_double(r), int(max_nn), int(ndiv),
                           C.c_double(plane_radius), int(flags), _ptr(rmin, C.c_float), _ptr(rmax, C.c_float),
                           int(nthreads))
    assert rc == 0
    return rmin, rmax


def rsd_ref_faithful(xyz, nrm, r, max_nn=0, ndiv=10, plane_radius=0.1):
    """kd-tree + materialised lists + estimation, 1 thread. Returns r_min, r_max, phase seconds (3,)."""
    L = lib()
    p = _xyz(xyz)
    nn = np.ascontiguousarray(nrm, dtype=np.float32)
    n = p.shape[0]
    rmin = np.empty(n, np.float32)
    rmax = np.empty(n, np.float32)
    ph = np.zeros(3, np.float64)
    L.orc_rsd_ref_faithful(_ptr(p, C.c_float), _ptr(nn, C.c_float), int(nn.shape[1]), n, C.c_double(r),
                           int(max_nn), int(ndiv), C.c_double(plane_radius), _ptr(rmin, C.c_float),
                           _ptr(rmax, C.c_float), _ptr(ph, C.c_double))
    return rmin, rmax, ph


def voxel_grid(xyz, leaf):
    """Returns dict(min_b, div_b, centroids (V,3), layout (cells,), counts (V,))."""
    L = lib()
    p = _xyz(xyz)
    min_b = np.zeros(3, np.int32)
    div_b = np.zeros(3, np.int32)
    nv = L.orc_voxel_grid(_ptr(p, C.c_float), p.shape[0], C.c_float(leaf), _ptr(min_b, C.c_int32),
                          _ptr(div_b, C.c_int32), None, None, None)
    cent = np.zeros((max(nv, 1), 3), np.float32)
    layout = np.zeros(max(int(np.prod(div_b.astype(np.int64))), 1), np.int32)
    counts = np.zeros(max(nv, 1), np.int32)
    L.orc_voxel_grid(_ptr(p, C.c_float), p.shape[0], C.c_float(leaf), _ptr(min_b, C.c_int32),
                     _ptr(div_b, C.c_int32), _ptr(cent, C.c_float), _ptr(layout, C.c_int32), _ptr(counts, C.c_int32))
    return dict(min_b=min_b, div_b=div_b, centroids=cent[:nv], layout=layout, counts=counts[:nv], nvox=nv)


def get_type(rmin, rmax):
    return lib().orc_get_type(C.c_float(rmin), C.c_float(rmax))


def offsets26():
    o = np.zeros((26, 3), np.int32)
    lib().orc_offsets26(_ptr(o, C.c_int32))
    return o


def grsd_transitions(grid, types):
    L = lib()
    t = np.ascontiguousarray(types, dtype=np.int32)
    M = np.zeros((6, 6), np.int32)
    h = np.zeros(21, np.int32)
    cent = np.ascontiguousarray(grid["centroids"], np.float32)
    L.orc_grsd_transitions(_ptr(cent, C.c_float), int(grid["nvox"]), _ptr(t, C.c_int32), C.c_float(grid["leaf"]),
                           _ptr(grid["min_b"], C.c_int32), _ptr(grid["div_b"], C.c_int32),
                           _ptr(grid["layout"], C.c_int32), _ptr(M, C.c_int32), _ptr(h, C.c_int32))
    return M, h


def grsd21_subdiv(grid, types, subdivision_size, off=(0, 0, 0)):
    L = lib()
    t = np.ascontiguousarray(types, dtype=np.int32)
    cent = np.ascontiguousarray(grid["centroids"], np.float32)
    sb = np.zeros(3, np.int32)
    args = (_ptr(cent, C.c_float), int(grid["nvox"]), _ptr(t, C.c_int32), C.c_float(grid["leaf"]),
            _ptr(grid["min_b"], C.c_int32), _ptr(grid["div_b"], C.c_int32), _ptr(grid["layout"], C.c_int32),
            int(subdivision_size), int(off[0]), int(off[1]), int(off[2]), _ptr(sb, C.c_int32))
    hn = L.orc_grsd21_subdiv(*args, None)
    if hn <= 0:
        return hn, sb, np.zeros((0, 21), np.int32)
    h = np.zeros((hn, 21), np.int32)
    L.orc_grsd21_subdiv(*args, _ptr(h, C.c_int32))
    return hn, sb, h


def grsd21(xyz, leaf, r_normals=0.02, rsd_radius_min=0.01, rsd_flags=0, normals_in=None, vp=(0.0, 0.0, 0.0),
           nthreads=0, cap_vox=1 << 20):
    """Whole recipe for one cluster. Returns dict(hist21, labels, radii (V,2), nvox)."""
    L = lib()
    p = _xyz(xyz)
    h = np.zeros(21, np.int32)
    labels = np.zeros(cap_vox, np.int32)
    radii = np.zeros((cap_vox, 2), np.float32)
    nv = C.c_int32(0)
    v = np.asarray(vp, dtype=np.float32)
    nn = None
    stride = 0
    if normals_in is not None:
        nn = np.ascontiguousarray(normals_in, dtype=np.float32)
        stride = int(nn.shape[1])
    rc = L.orc_grsd21(_ptr(p, C.c_float), _ptr(nn, C.c_float), stride, p.shape[0], C.c_float(leaf),
                      C.c_double(r_normals), C.c_double(rsd_radius_min), int(rsd_flags), _ptr(v, C.c_float),
                      _ptr(h, C.c_int32), _ptr(labels, C.c_int32), _ptr(radii, C.c_float), int(cap_vox),
                      C.byref(nv), int(nthreads))
    assert rc == 0
    return dict(hist21=h, labels=labels[: nv.value].copy(), radii=radii[: nv.value].copy(), nvox=nv.value)


SIG_GRSD21, SIG_GRSD325, SIG_PLUSGRSD110 = 0, 1, 2
SIG_DIM = {0: 21, 1: 325, 2: 110}


def voxel_normals(xyz, nrm, leaf):
    """Un-normalised mean normal of every voxel (V,3), voxel order as voxel_grid()."""
    L = lib()
    p = _xyz(xyz)
    nn = np.ascontiguousarray(nrm, dtype=np.float32)
    out = np.zeros((max(p.shape[0], 1), 3), np.float32)
    nv = L.orc_voxel_normals(_ptr(p, C.c_float), _ptr(nn, C.c_float), int(nn.shape[1]), p.shape[0], C.c_float(leaf),
                             _ptr(out, C.c_float))
    return out[:nv].copy()


def grsd_signature(kind, grid, types, cent_normals=None, subdivision_size=0, off=(0, 0, 0)):
    """Returns (hist_num, subdiv_b (3,), hist (hist_num, dim) int32)."""
    L = lib()
    t = np.ascontiguousarray(types, dtype=np.int32)
    cent = np.ascontiguousarray(grid["centroids"], np.float32)
    cn = np.ascontiguousarray(cent_normals, np.float32) if cent_normals is not None else None
    sb = np.zeros(3, np.int32)
    args = (int(kind), _ptr(cent, C.c_float), _ptr(cn, C.c_float), int(grid["nvox"]), _ptr(t, C.c_int32),
            C.c_float(grid["leaf"]), _ptr(grid["min_b"], C.c_int32), _ptr(grid["div_b"], C.c_int32),
            _ptr(grid["layout"], C.c_int32), int(subdivision_size), int(off[0]), int(off[1]), int(off[2]),
            _ptr(sb, C.c_int32))
    hn = L.orc_grsd_signature(*args, None)
    if hn <= 0:
        return hn, sb, np.zeros((0, SIG_DIM[kind]), np.int32)
    h = np.zeros((hn, SIG_DIM[kind]), np.int32)
    assert L.orc_grsd_signature(*args, _ptr(h, C.c_int32)) == hn
    return hn, sb, h


def voxel_colors(xyz, rgb, leaf):
    """pcl::VoxelGrid's colour of every voxel (packed 0x00RRGGBB, voxel order as voxel_grid)."""
    L = lib()
    p = _xyz(xyz)
    c = np.ascontiguousarray(rgb, np.uint32)
    out = np.zeros(max(p.shape[0], 1), np.uint32)
    nv = L.orc_voxel_colors(_ptr(p, C.c_float), _ptr(c, C.c_uint32), p.shape[0], C.c_float(leaf), _ptr(out, C.c_uint32))
    return out[:nv]


def color_chlac117(grid, voxel_rgb, thr=(127, 127, 127), c3=True, subdivision_size=0, off=(0, 0, 0)):
    """Rotation-invariant C3-HLAC (c3=True) / Color-CHLAC, 117 bins.  grid: voxel_grid(...) + "leaf";
    returns (hist_num, subdiv_b (3,), hist (hist_num, 117) float32)."""
    L = lib()
    cent = np.ascontiguousarray(grid["centroids"], np.float32)
    col = np.ascontiguousarray(voxel_rgb, np.uint32)
    sb = np.zeros(3, np.int32)
    args = (1 if c3 else 0, _ptr(cent, C.c_float), _ptr(col, C.c_uint32), int(grid["nvox"]), C.c_float(grid["leaf"]),
            _ptr(grid["min_b"], C.c_int32), _ptr(grid["div_b"], C.c_int32), _ptr(grid["layout"], C.c_int32),
            int(thr[0]), int(thr[1]), int(thr[2]), int(subdivision_size), int(off[0]), int(off[1]), int(off[2]), _ptr(sb, C.c_int32))
    hn = L.orc_color_chlac117(*args, None)
    if hn <= 0:
        return hn, sb, np.zeros((0, 117), np.float32)
    h = np.zeros((hn, 117), np.float32)
    assert L.orc_color_chlac117(*args, _ptr(h, C.c_float)) == hn
    return hn, sb, h


def grsd_cluster(xyz, leaf, kind=SIG_GRSD21, subdivision_size=0, off=(0, 0, 0), r_normals=0.02, rsd_radius_min=0.01,
                 rsd_flags=0, normals_in=None, vp=(0.0, 0.0, 0.0), nthreads=0):
    """Whole recipe for one cluster and any signature: normals -> voxel grid -> voxel RSD -> labels ->
    signature.  Returns dict(hist_num, subdiv_b, hist, labels, grid, cent_normals)."""
    p = _xyz(xyz)
    if normals_in is None:
        n4, _ = normals(p, r_normals, vp=vp, nthreads=nthreads)
        nrm = np.ascontiguousarray(n4[:, :3])
    else:
        nrm = np.ascontiguousarray(np.asarray(normals_in, np.float32)[:, :3])
    base = grsd21(p, leaf, r_normals=r_normals, rsd_radius_min=rsd_radius_min, rsd_flags=rsd_flags, normals_in=nrm, vp=vp,
                  nthreads=nthreads)
    grid = voxel_grid(p, leaf)
    grid["leaf"] = leaf
    cn = voxel_normals(p, nrm, leaf)
    hn, sb, h = grsd_signature(kind, grid, base["labels"], cn, subdivision_size, off)
    return dict(hist_num=hn, subdiv_b=sb, hist=h, labels=base["labels"], grid=grid, cent_normals=cn)


def svm_predict(model, features, scale=None, want_dec=False):
    """model: mapping_private_b200.svm_model.SvmModel (or any object with its fields); scale: None or
    (lower, upper, fmin, fmax).  Returns predicted labels (n,) float32 [and decision values]."""
    L = lib()
    f = np.ascontiguousarray(features, dtype=np.float32)
    n, dim = f.shape
    assert dim == model.sv.shape[1]
    k = int(model.labels.shape[0])
    out = np.zeros(n, np.float32)
    dec = np.zeros((n, k * (k - 1) // 2), np.float64) if want_dec else None
    d64 = np.ctypeslib.as_ctypes_type(np.float64)
    lower, upper, fmin, fmax = (0.0, 0.0, None, None) if scale is None else scale
    lab = np.ascontiguousarray(model.labels, np.int32)
    nsv = np.ascontiguousarray(model.nr_sv, np.int32)
    rho = np.ascontiguousarray(model.rho, np.float64)
    coef = np.ascontiguousarray(model.sv_coef, np.float64)
    sv = np.ascontiguousarray(model.sv, np.float64)
    fmin = None if fmin is None else np.ascontiguousarray(fmin, np.float64)
    fmax = None if fmax is None else np.ascontiguousarray(fmax, np.float64)
    rc = L.orc_svm_predict(_ptr(f, C.c_float), C.c_int64(n), int(dim), k, int(sv.shape[0]), C.c_double(model.gamma),
                           _ptr(lab, C.c_int32), _ptr(nsv, C.c_int32), _ptr(rho, d64), _ptr(coef, d64), _ptr(sv, d64),
                           C.c_double(lower), C.c_double(upper), _ptr(fmin, d64), _ptr(fmax, d64), _ptr(out, C.c_float),
                           _ptr(dec, d64))
    assert rc == 0
    return (out, dec) if want_dec else out


def knn_mean_distance(xyz, k, nthreads=0):
    """StatisticalNoiseRemoval's per-point mean distance to the k-1 nearest neighbours (float64, NaN for
    non-finite points).  Raises ValueError like the plugin's checks (k < 2, fewer than k points)."""
    L = lib()
    p = _xyz(xyz)
    avg = np.zeros(p.shape[0], np.float64)
    rc = L.orc_knn_mean_distance(_ptr(p, C.c_float), p.shape[0], int(k), avg.ctypes.data_as(C.POINTER(C.c_double)), int(nthreads))
    if rc != 0:
        raise ValueError("not enough neighbours requested" if rc == -1 else "not enough points in the cloud")
    return avg


def normals_knn(xyz, k, vp=(0.0, 0.0, 0.0), nthreads=0):
    """k-nearest-neighbour PCA normals (n, 4): nx, ny, nz, curvature."""
    L = lib()
    p = _xyz(xyz)
    out = np.zeros((p.shape[0], 4), np.float32)
    v = np.asarray(vp, np.float32)
    rc = L.orc_normals_knn(_ptr(p, C.c_float), p.shape[0], int(k), _ptr(v, C.c_float), _ptr(out, C.c_float), int(nthreads))
    if rc != 0:
        raise ValueError("k < 3" if rc == -1 else "not enough points in the cloud")
    return out


def euclidean_clusters(xyz, tolerance, min_pts=1, max_pts=0):
    """extractEuclideanClusters without the normal test: (labels int32 (n,), n_clusters); clusters are numbered in
    the order of their smallest index, -1 = dropped (too small / too large) or non-finite."""
    L = lib()
    p = _xyz(xyz)
    labels = np.full(p.shape[0], -1, np.int32)
    nc = L.orc_euclidean_clusters(_ptr(p, C.c_float), p.shape[0], C.c_double(tolerance), int(min_pts), int(max_pts),
                                  _ptr(labels, C.c_int32))
    return labels, int(nc)


def fit_plane_msac(xyz, triples, indices=None, threshold=0.03, max_iterations=500, probability=0.99):
    """fitSACPlane on a given sample sequence: dict(coeff, inliers, projected, iterations, best_iteration)."""
    L = lib()
    L.orc_fit_plane_msac.restype = C.c_int64
    p = _xyz(xyz)
    tri = np.ascontiguousarray(triples, np.int32).reshape(-1, 3)
    idx = None if indices is None else np.ascontiguousarray(indices, np.int32)
    m = p.shape[0] if idx is None else idx.shape[0]
    coeff = np.zeros(4, np.float64)
    inl = np.zeros(max(m, 1), np.int32)
    proj = np.zeros((max(m, 1), 3), np.float32)
    it, best = C.c_int32(), C.c_int32()
    nin = L.orc_fit_plane_msac(_ptr(p, C.c_float), C.c_int64(p.shape[0]), None if idx is None else _ptr(idx, C.c_int32),
                               C.c_int64(0 if idx is None else idx.shape[0]), C.c_double(threshold), C.c_int32(max_iterations),
                               C.c_double(probability), _ptr(tri, C.c_int32), C.c_int64(tri.shape[0]),
                               coeff.ctypes.data_as(C.POINTER(C.c_double)), _ptr(inl, C.c_int32), _ptr(proj, C.c_float),
                               C.byref(it), C.byref(best))
    return dict(coeff=coeff, inliers=inl[:nin], projected=proj[:nin], iterations=it.value, best_iteration=best.value)


def noise_filter(avg, alpha):
    """Returns (keep bool (n,), mean, stddev)."""
    L = lib()
    L.orc_noise_filter.restype = C.c_int64
    a = np.ascontiguousarray(avg, np.float64)
    keep = np.zeros(a.shape[0], np.uint8)
    mean, std = C.c_double(), C.c_double()
    L.orc_noise_filter(a.ctypes.data_as(C.POINTER(C.c_double)), a.shape[0], C.c_double(alpha),
                       keep.ctypes.data_as(C.POINTER(C.c_uint8)), C.byref(mean), C.byref(std))
    return keep.astype(bool), mean.value, std.value


PFH_USE_DIST, PFH_DIFFERENTIAL, PFH_CHECK_FLIP, PFH_ABS_ANGLES, PFH_AVERAGE, PFH_COMBINE = 1, 2, 4, 8, 16, 32
PFH_DEFAULT = PFH_CHECK_FLIP | PFH_AVERAGE  # the plugin's defaults produce FPFHs (pfh.h:83-93)


def pfh(xyz, nrm, radius=0.03, max_nn=100, quantum=9, flags=PFH_DEFAULT, nthreads=0):
    """PointFeatureHistogram (star features, optional FPFH averaging).  Returns (n, quantum * (3 or 4)) float32."""
    L = lib()
    p = _xyz(xyz)
    nn = np.ascontiguousarray(nrm, dtype=np.float32)
    nf = 4 if flags & PFH_USE_DIST else 3
    nb = quantum ** nf if flags & PFH_COMBINE else quantum * nf
    out = np.zeros((p.shape[0], nb), np.float32)
    rc = L.orc_pfh(_ptr(p, C.c_float), _ptr(nn, C.c_float), int(nn.shape[1]), p.shape[0], C.c_double(radius), int(max_nn),
                   int(quantum), int(flags), _ptr(out, C.c_float), int(nthreads))
    assert rc == nb
    return out


def pfh_pair(ps, ns, pt, nt, d2, max_dist, check_flip=True, abs_angles=False):
    L = lib()
    a = [np.ascontiguousarray(v, np.float32) for v in (ps, ns, pt, nt)]
    f = np.zeros(4, np.float64)
    ok = L.orc_pfh_pair(*[_ptr(v, C.c_float) for v in a], C.c_float(d2), C.c_double(max_dist), int(check_flip), int(abs_angles),
                        f.ctypes.data_as(C.POINTER(C.c_double)))
    return bool(ok), f


def num_threads():
    return lib().orc_num_threads()
