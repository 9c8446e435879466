"""The host-side C++ mirror of the CloudAlgo plugin surface: discovery, requires/provides, topic
names, parameters and error behaviour (CPU), and results against the oracle (GPU)."""
import pathlib
import re

import numpy as np
import pytest

from mapping_private_b200 import plugin, synth

ROOT = pathlib.Path(__file__).resolve().parent.parent
HOST = ROOT / "mapping-private_b200" / "host"


@pytest.fixture(scope="module")
def built():
    from mapping_private_b200 import cab

    cab.build()
    plugin.build()
    return plugin.LIB_PATH


def test_discovery_matches_plugins_xml(built):
    xml = (HOST / "plugins.xml").read_text()
    assert '<library path="lib/libcloud_algos">' in xml
    declared = re.findall(r'<class name="([^"]+)" type="([^"]+)" base_class_type="([^"]+)">', xml)
    ref_xml = {"cloud_algos/NormalEstimation", "cloud_algos/PlanarEstimation", "cloud_algos/RotationalEstimation",
               "cloud_algos/CylinderEstimation", "cloud_algos/SVMClassification", "cloud_algos/StatisticalNoiseRemoval",
               "cloud_algos/LocalRadiusEstimation"}  # the reference's cloud_algos/plugins.xml, every entry kept
    assert {d[0] for d in declared} == ref_xml | {"cloud_algos/GlobalRSD", "cloud_algos/PointFeatureHistogram"}
    not_here = {"cloud_algos/PlanarEstimation", "cloud_algos/RotationalEstimation", "cloud_algos/CylinderEstimation"}
    for name, typ, base in declared:
        assert typ == name.replace("/", "::") and base == "cloud_algos::CloudAlgo"
        if name in not_here:  # advertised as in the reference's file, not part of the hot path: pluginlib reports them missing
            with pytest.raises(KeyError):
                plugin.Plugin(name)
            continue
        p = plugin.Plugin(name)  # pluginlib lookup by the reference's names
        p.close()
    with pytest.raises(KeyError):
        plugin.Plugin("cloud_algos/NoSuchPlugin")


def test_requires_provides_and_topics(built):
    rsd = plugin.Plugin("cloud_algos/LocalRadiusEstimation")
    req, prov = rsd.requires_provides()
    assert req == ["x", "y", "z", "nx", "ny", "nz"]  # radius_estimation.cpp:27-39
    assert prov == ["r_min", "r_max", "r_dif", "point_label"]  # :41-50
    assert rsd.topic() == "cloud_radius"
    ne = plugin.Plugin("cloud_algos/NormalEstimation")
    assert ne.requires_provides() == (["x", "y", "z"], ["nx", "ny", "nz", "curvature"])
    assert ne.topic() == "cloud_normals"
    g = plugin.Plugin("cloud_algos/GlobalRSD")
    assert g.requires_provides()[1] == [f"f{i}" for i in range(1, 22)]


def test_standalone_node_binary_builds(built):
    """radius_estimation_node = the plugin source compiled with -DCREATE_NODE (reference CMakeLists.txt:56-57)."""
    assert (HOST / "radius_estimation_node").exists()
    src = (HOST / "src" / "radius_estimation.cpp").read_text()
    assert "standalone_node <cloud_algos::LocalRadiusEstimation>" in src


def test_missing_normals_is_reported(built):
    pts = synth.analytic_shape("plane", 100)
    for name in ("cloud_algos/LocalRadiusEstimation", "cloud_algos/GlobalRSD"):
        p = plugin.Plugin(name)
        res, out = p.run(pts, {"intensity": np.zeros(100, np.float32)})
        assert res == "missing normals" and out is None and not p.output_valid() and p.num_published() == 0


def test_svm_plugin_surface_and_errors(built, tmp_path):
    svm = plugin.Plugin("cloud_algos/SVMClassification")
    assert svm.requires_provides() == (["f1"], ["point_class"])  # svm_classification.cpp:27-40
    assert svm.topic() == "cloud_svm"
    pts = synth.analytic_shape("plane", 10)
    res, out = svm.run(pts, {"nx": np.zeros(10, np.float32)})
    assert res == "missing features" and out is None and not svm.output_valid()
    feats = {f"f{i}": np.zeros(10, np.float32) for i in range(1, 22)}
    svm.set_param("model_file_name", str(tmp_path / "nope.model"))
    res, out = svm.run(pts, feats)
    assert res == "incorrect model file" and out is None
    bad = tmp_path / "bad.model"
    bad.write_text("svm_type nu_svc\nkernel_type rbf\nSV\n")
    res, _ = svm.run(pts, feats, fields={"model_file_name_": str(bad)})
    assert res == "incorrect model file"


def test_noise_removal_plugin_surface_and_errors(built):
    nr = plugin.Plugin("cloud_algos/StatisticalNoiseRemoval")
    assert nr.requires_provides() == (["x", "y", "z"], ["x", "y", "z"])  # noise_removal.cpp:24-43
    assert nr.topic() == "cloud_denoise"
    pts = synth.analytic_shape("plane", 50)
    res, out = nr.run(pts, fields={"neighborhood_size_": 1})
    assert res == "ERROR: Not enough neighbors requested!" and out is None and not nr.output_valid()
    res, out = nr.run(pts, fields={"alpha_": -1.0})
    assert res == "ERROR: Not enough neighbors requested!"
    # public fields persist between calls, exactly like the reference's (pre() only overrides what has a rosparam)
    res, out = nr.run(pts[:5], fields={"alpha_": 3.0, "neighborhood_size_": 10})
    assert res == "ERROR: Not enough points in the cloud (or too many neighbors requested)!" and out is None


def test_pfh_plugin_surface_and_errors(built):
    pf = plugin.Plugin("cloud_algos/PointFeatureHistogram")
    req, prov = pf.requires_provides()
    assert req == ["x", "y", "z", "nx", "ny", "nz"] and prov == [f"f{i}" for i in range(1, 28)]  # pfh.cpp:31-76, 9 x 3 bins
    assert pf.topic() == "cloud_pfh"
    pts = synth.analytic_shape("plane", 50)
    res, out = pf.run(pts, {"intensity": np.zeros(50, np.float32)})
    assert res == "missing normals" and out is None
    nrm = {"nx": np.zeros(50, np.float32), "ny": np.zeros(50, np.float32), "nz": np.ones(50, np.float32)}
    pf.set_field("combine_", 1)
    pf.set_field("quantum_", 4)
    assert pf.requires_provides()[1] == [f"f{i}" for i in range(1, 4 ** 3 + 1)]  # pfh.cpp:49-52: quantum ^ features bins
    pf.set_field("combine_", 0)
    pf.set_field("quantum_", 9)
    pf.set_field("use_dist_", 1)
    pf.set_field("point_label_", 3)
    assert pf.requires_provides()[1] == [f"f{i}" for i in range(1, 37)] + ["point_label"]


def test_sample_pipeline_yaml_keys(built):
    import yaml

    doc = yaml.safe_load((HOST / "sample_pipeline.yaml").read_text())
    for entry in doc.values():  # the key layout of the reference's sample_pipeline.yaml:2-7
        assert {"launch_pkg", "launch_type", "class_name", "input_topic_name", "output_topic_name"} <= set(entry)
    assert set(doc["LocalRadiusEstimation"]) >= {"radius", "max_nn", "plane_radius", "distance_div", "point_label", "rmin2curvature"}


@pytest.mark.gpu
def test_pipeline_normals_then_rsd_against_oracle(built, oracle):
    pts = synth.tabletop(30_000, noise_sigma=0.0003)
    ne = plugin.Plugin("cloud_algos/NormalEstimation")
    ne.set_param("radius", 0.02)
    res, out = ne.run(pts, {"intensity": np.arange(len(pts), dtype=np.float32)})
    assert res == "ok" and list(out["channels"]) == ["intensity", "nx", "ny", "nz", "curvature"]
    o4, _ = oracle.normals(pts, 0.02)
    n3 = np.stack([out["channels"][k] for k in ("nx", "ny", "nz")], 1)
    good = ~np.isnan(o4[:, 0])
    assert np.mean(np.linalg.norm(np.cross(n3[good].astype(np.float64), o4[good, :3].astype(np.float64)), axis=1) > 1e-4) < 1e-3
    assert np.array_equal(out["points"], pts)

    rsd = plugin.Plugin("cloud_algos/LocalRadiusEstimation")
    # the launch-file parameters of the reference (launch/pipeline_tmp.launch:20)
    for k, v in {"rmin2curvature": 1, "radius": 0.02, "max_nn": 75}.items():
        rsd.set_param(k, v)
    chans = {"nx": o4[:, 0], "ny": o4[:, 1], "nz": o4[:, 2], "curvature": o4[:, 3]}
    res, out = rsd.run(pts, chans, fields={"point_label_": 7})
    assert res == "ok" and rsd.output_valid() and rsd.num_published() == 1
    assert list(out["channels"]) == ["nx", "ny", "nz", "curvature", "r_min", "r_max", "r_dif", "point_label"]
    omin, omax, odif = oracle.rsd(pts, o4, 0.02, max_nn=75)
    assert np.max(np.abs(out["channels"]["r_min"] - omin) / omin) < 1e-4
    assert np.max(np.abs(out["channels"]["r_max"] - omax) / omax) < 1e-4
    assert np.allclose(out["channels"]["r_dif"], odif, atol=1e-6)
    assert np.array_equal(out["channels"]["curvature"], out["channels"]["r_min"])  # rmin2curvature
    assert np.all(out["channels"]["point_label"] == 7)
    # defaults of the class (radius 0.03, max_nn 150, radius_estimation.h:81-86) when no rosparam is set
    rsd2 = plugin.Plugin("cloud_algos/LocalRadiusEstimation")
    res, out2 = rsd2.run(pts, chans)
    omin2, omax2, _ = oracle.rsd(pts, o4, 0.03, max_nn=150)
    assert np.max(np.abs(out2["channels"]["r_min"] - omin2) / omin2) < 1e-4
    assert np.all(out2["channels"]["point_label"] == 0)  # channel added but left at zero (:209-214)


@pytest.mark.gpu
def test_cloud_algo_node_subscribes_and_publishes(built, oracle):
    """CloudAlgoNode<LocalRadiusEstimation> (cloud_algos.h:46-104): a message on cloud_pcd is processed and published once;
    without normals nothing is published (output_valid_ false)."""
    import ctypes as C

    L = plugin.lib()
    pts = synth.tabletop(5_000)
    n4, _ = oracle.normals(pts, 0.02)
    fp = lambda a: np.ascontiguousarray(a, np.float32).ctypes.data_as(C.POINTER(C.c_float))
    cols = [np.ascontiguousarray(n4[:, i]) for i in range(3)]
    xyz = np.ascontiguousarray(pts, np.float32)
    L.capi_radius_node_roundtrip.argtypes = [C.POINTER(C.c_float)] * 4 + [C.c_int, C.c_double]
    assert L.capi_radius_node_roundtrip(fp(xyz), fp(cols[0]), fp(cols[1]), fp(cols[2]), xyz.shape[0], 0.02) == 1
    assert L.capi_radius_node_roundtrip(fp(xyz), None, None, None, xyz.shape[0], 0.02) == 0


@pytest.mark.gpu
def test_global_rsd_plugin_against_oracle(built, oracle):
    xyz, off = synth.clusters(2, 2000, 4000, seed_extra=5)
    pts = xyz[off[1]:off[2]]
    o4, _ = oracle.normals(pts, 0.02)
    nrm = np.nan_to_num(o4[:, :3], nan=0.0)
    g = plugin.Plugin("cloud_algos/GlobalRSD")
    g.set_param("width", 0.03)
    res, out = g.run(pts, {"nx": nrm[:, 0], "ny": nrm[:, 1], "nz": nrm[:, 2]},
                     fields={"min_voxel_pts_": 0, "publish_cloud_centroids_": 1, "publish_cloud_vrsd_": 1})
    assert res == "ok" and out["points"].shape == (1, 3)
    hist = np.array([out["channels"][f"f{i}"][0] for i in range(1, 22)])
    o = oracle.grsd21(pts, 0.03, normals_in=nrm)
    assert np.array_equal(hist.astype(np.int64), o["hist21"].astype(np.int64))
    vrsd = g.output(1)
    assert np.array_equal(vrsd["channels"]["point_label"].astype(np.int32), o["labels"])


@pytest.mark.gpu
def test_global_rsd_process_batch_equals_per_cluster_calls(built, oracle):
    """GlobalRSD::process_batch (all clusters of a frame in one device call) against process () cluster by cluster, the
    loop of table_memory_grsd.cpp:913-997."""
    xyz, off = synth.clusters(6, 1500, 3000, seed_extra=7)
    nrm = np.zeros_like(xyz)
    for c in range(6):
        o4, _ = oracle.normals(xyz[off[c]:off[c + 1]], 0.02)
        nrm[off[c]:off[c + 1]] = np.nan_to_num(o4[:, :3], nan=0.0)
    g = plugin.Plugin("cloud_algos/GlobalRSD")
    g.set_param("width", 0.03)
    g._L.capi_pre(g._h)
    g.set_field("min_voxel_pts_", 0)
    res, hist = g.grsd_process_batch(xyz, nrm, off)
    assert res == "ok" and hist.shape == (6, 21)
    for c in range(6):
        pts = xyz[off[c]:off[c + 1]]
        n3 = nrm[off[c]:off[c + 1]]
        r1, out = g.run(pts, {"nx": n3[:, 0], "ny": n3[:, 1], "nz": n3[:, 2]}, fields={"min_voxel_pts_": 0})
        assert r1 == "ok"
        one = np.array([out["channels"][f"f{i}"][0] for i in range(1, 22)])
        assert np.array_equal(one, hist[c]), c


@pytest.mark.gpu
def test_svm_plugin_against_oracle(built, oracle, tmp_path):
    """GlobalRSD -> SVMClassification the way table_memory_grsd.cpp:974-1018 chains them, with the
    reference's grsd_ijrr model and scale ranges written back to files in libsvm's formats."""
    from mapping_private_b200 import svm_model

    z = np.load(ROOT / "tests" / "golden" / "svm_grsd_ijrr.npz")
    m = svm_model.SvmModel(float(z["gamma"]), z["labels"], z["nr_sv"], z["rho"], z["sv_coef"], z["sv"])
    (tmp_path / "grsd.model").write_text(svm_model.format_model(m))
    scp = "x\n-1 1\n" + "".join(f"{i + 1} {z['fmin'][i]:.9g} {z['fmax'][i]:.9g}\n" for i in range(21))
    (tmp_path / "grsd.scp").write_text(scp)
    scale = (-1.0, 1.0, z["fmin"], z["fmax"])
    xyz, off = synth.clusters(3, 2000, 4000, seed_extra=6)
    g = plugin.Plugin("cloud_algos/GlobalRSD")
    svm = plugin.Plugin("cloud_algos/SVMClassification")
    svm.set_param("model_file_name", str(tmp_path / "grsd.model"))
    svm.set_param("scale_file_name", str(tmp_path / "grsd.scp"))
    for c in range(3):
        pts = xyz[off[c]:off[c + 1]]
        nrm = np.nan_to_num(oracle.normals(pts, 0.02)[0][:, :3], nan=0.0)
        res, grsd = g.run(pts, {"nx": nrm[:, 0], "ny": nrm[:, 1], "nz": nrm[:, 2]}, fields={"min_voxel_pts_": 0})
        assert res == "ok"
        res, out = svm.run(grsd["points"], grsd["channels"])
        assert res == "ok" and svm.output_valid() and list(out["channels"])[-1] == "point_class"
        hist = np.array([[grsd["channels"][f"f{i}"][0] for i in range(1, 22)]], np.float32)
        assert out["channels"]["point_class"][0] == oracle.svm_predict(m, hist, scale=scale)[0]
    # many points at once, scale_self, and no scaling at all
    rng = np.random.default_rng(2)
    F = (rng.random((500, 21)) ** 2 * np.maximum(z["fmax"], 1)).astype(np.float32)
    chans = {f"f{i + 1}": F[:, i] for i in range(21)}
    chans["point_label"] = np.full(500, 31, np.float32)
    pts = np.zeros((500, 3), np.float32)
    res, out = svm.run(pts, chans)
    assert np.array_equal(out["channels"]["point_class"], oracle.svm_predict(m, F, scale=scale))
    res, out = svm.run(pts, chans, fields={"scale_file_": 0})
    assert np.array_equal(out["channels"]["point_class"], oracle.svm_predict(m, F))
    # scale_self_: per-channel min / max with the reference's update rule (first value only lowers the minimum)
    lo = np.full(21, np.finfo(np.float64).max)
    hi = np.full(21, -np.finfo(np.float64).max)
    for row in F.astype(np.float64):
        low = lo > row
        lo = np.where(low, row, lo)
        hi = np.where(~low & (hi < row), row, hi)
    res, out = svm.run(pts, chans, fields={"scale_self_": 1})
    assert np.array_equal(out["channels"]["point_class"], oracle.svm_predict(m, F, scale=(-1.0, 1.0, lo, hi)))


@pytest.mark.gpu
def test_noise_removal_plugin_against_oracle(built, oracle):
    rng = np.random.default_rng(4)
    pts = np.concatenate([synth.tabletop(40_000, noise_sigma=0.0005),
                          synth.quantize(rng.uniform([-0.6, -0.4, 0.5], [0.6, 0.4, 1.3], (400, 3)))]).astype(np.float32)
    tag = np.arange(len(pts), dtype=np.float32)
    nr = plugin.Plugin("cloud_algos/StatisticalNoiseRemoval")
    nr.set_param("alpha", 2.0)
    nr.set_param("neighborhood_size", 12)
    res, out = nr.run(pts, {"tag": tag})
    assert res == "ok" and nr.output_valid() and nr.num_published() == 1
    avg = oracle.knn_mean_distance(pts, 12)
    keep, mean, std = oracle.noise_filter(avg, 2.0)
    margin = np.abs(np.abs(avg - mean) - 2.0 * std)
    sure = margin > 1e-9
    got = np.zeros(len(pts), bool)
    got[out["channels"]["tag"].astype(np.int64)] = True
    assert np.array_equal(got[sure], keep[sure]) and sure.mean() > 0.999
    assert np.array_equal(out["points"], pts[got])  # kept points in input order, channels follow
    assert keep[-400:].mean() < 0.2  # the scattered points are what gets removed
    # the size check (noise_removal.cpp:152-158)
    res, out = nr.run(pts, {"tag": tag}, fields={"min_nr_pts_": len(pts)})
    assert res == "output size check failed (see min_nr_pts parameter)" and out is None and not nr.output_valid()


@pytest.mark.gpu
def test_pfh_plugin_against_oracle(built, oracle):
    pts = synth.tabletop(15_000, noise_sigma=0.0003)
    nrm = np.nan_to_num(oracle.normals(pts, 0.02)[0][:, :3], nan=0.0)
    pf = plugin.Plugin("cloud_algos/PointFeatureHistogram")
    res, out = pf.run(pts, {"nx": nrm[:, 0], "ny": nrm[:, 1], "nz": nrm[:, 2]}, fields={"point_label_": 5})
    assert res == "ok" and list(out["channels"])[:3] == ["nx", "ny", "nz"] and list(out["channels"])[-1] == "point_label"
    got = np.stack([out["channels"][f"f{i}"] for i in range(1, 28)], 1)
    want = oracle.pfh(pts, nrm)  # the plugin's defaults: radius 0.03, max_nn 100, quantum 9, check_flip, average
    assert np.allclose(got, want, rtol=2e-4, atol=2e-3)
    assert np.all(out["channels"]["point_label"] == 5)
