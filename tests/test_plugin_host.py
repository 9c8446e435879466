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
    assert {d[0] for d in declared} == {"cloud_algos/NormalEstimation", "cloud_algos/LocalRadiusEstimation", "cloud_algos/GlobalRSD"}
    for name, typ, base in declared:
        assert typ == name.replace("/", "::") and base == "cloud_algos::CloudAlgo"
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


def test_missing_normals_is_reported(built):
    pts = synth.analytic_shape("plane", 100)
    for name in ("cloud_algos/LocalRadiusEstimation", "cloud_algos/GlobalRSD"):
        p = plugin.Plugin(name)
        res, out = p.run(pts, {"intensity": np.zeros(100, np.float32)})
        assert res == "missing normals" and out is None and not p.output_valid() and p.num_published() == 0


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
