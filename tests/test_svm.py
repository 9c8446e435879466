"""SVM classification: the oracle restatement of libsvm's svm_predict is
pinned against libsvm itself (scikit-learn's SVC wraps it), the model / scale-file parsers are
exercised on generated files, and the reference's grsd_ijrr model (tests/golden/svm_grsd_ijrr.npz,
made by tests/golden/make_golden.py) gives known answers.  The GPU path is checked against the oracle."""
import pathlib

import numpy as np
import pytest

from mapping_private_b200 import cab, svm_model, synth

GOLDEN = pathlib.Path(__file__).resolve().parent / "golden" / "svm_grsd_ijrr.npz"


def _golden():
    z = np.load(GOLDEN)
    m = svm_model.SvmModel(float(z["gamma"]), z["labels"], z["nr_sv"], z["rho"], z["sv_coef"], z["sv"])
    return m, (float(z["lower"]), float(z["upper"]), z["fmin"], z["fmax"])


def _sk_model(seed=0, k=6, dim=21, n=900, gamma=0.5):
    from sklearn.svm import SVC

    rng = np.random.default_rng(seed)
    cent = rng.normal(size=(k, dim))
    y = rng.integers(0, k, n)
    X = cent[y] + 0.8 * rng.normal(size=(n, dim))
    clf = SVC(kernel="rbf", gamma=gamma, C=4.0, decision_function_shape="ovo").fit(X, y * 10 + 3)
    m = svm_model.SvmModel(gamma, clf.classes_.astype(np.int32), clf.n_support_.astype(np.int32), -clf.intercept_,
                           clf.dual_coef_.copy(), clf.support_vectors_.copy())
    Xt = (cent[rng.integers(0, k, 1500)] + 1.2 * rng.normal(size=(1500, dim))).astype(np.float32)
    return clf, m, Xt


def test_oracle_matches_libsvm(oracle):
    clf, m, Xt = _sk_model()
    pred, dec = oracle.svm_predict(m, Xt, want_dec=True)
    assert np.array_equal(pred, clf.predict(Xt.astype(np.float64)).astype(np.float32))
    assert np.max(np.abs(dec - clf.decision_function(Xt.astype(np.float64)))) < 1e-13


def test_model_file_round_trip(oracle):
    _, m, Xt = _sk_model(seed=3, k=4, dim=9, n=300)
    m.sv[:, 2] = 0.0  # a column that the sparse format omits entirely
    m2 = svm_model.parse_model(svm_model.format_model(m), dim=9)
    for f in ("labels", "nr_sv", "rho", "sv_coef", "sv"):
        assert np.array_equal(getattr(m, f), getattr(m2, f)), f
    assert m2.gamma == m.gamma and m2.dim == 9
    assert np.array_equal(oracle.svm_predict(m, Xt), oracle.svm_predict(m2, Xt))
    with pytest.raises(ValueError):
        svm_model.parse_model("svm_type nu_svc\nkernel_type rbf\nSV\n")
    with pytest.raises(ValueError):
        svm_model.parse_model("svm_type c_svc\nkernel_type linear\nSV\n")


def test_scale_file_semantics(oracle):
    # parseScaleParameterFile: "x", bounds, then index/min/max through a C float; unknown layout -> None
    txt = "x\n-1 1\n1 10 14356\n2 0 618\n4 0.1 0.30000001\n9 5 6\n"
    lower, upper, fmin, fmax = svm_model.parse_scale(txt, 5)
    assert (lower, upper) == (-1.0, 1.0)
    assert fmin.tolist() == [10.0, 0.0, 0.0, float(np.float32(0.1)), 0.0]
    assert fmax.tolist() == [14356.0, 618.0, 0.0, float(np.float32(0.30000001)), 0.0]
    assert svm_model.parse_scale("y\n-1 1\n", 5) is None
    # scaleFeature: single-valued attribute -> 0, clamping, linear interpolation (svm_classification.h:68-86)
    _, m, _ = _sk_model(seed=5, k=3, dim=5, n=200)
    X = np.array([[5, 700, 3, 0.2, 1], [20000, -4, 3, 0.1, 1], [7183, 309, 0, 0.25, 0]], np.float32)
    scaled = np.array([[-1, 1, 0, 0, 0], [1, -1, 0, -1, 0], [0, 0, 0, 0, 0]], np.float64)
    scaled[0, 3] = -1 + 2 * (np.float64(np.float32(0.2)) - fmin[3]) / (fmax[3] - fmin[3])
    scaled[2, 3] = -1 + 2 * (np.float64(np.float32(0.25)) - fmin[3]) / (fmax[3] - fmin[3])
    scaled[2, 0] = -1 + 2 * (7183.0 - 10.0) / (14356.0 - 10.0)
    scaled[2, 1] = -1 + 2 * (309.0 - 0.0) / 618.0
    _, d1 = oracle.svm_predict(m, X, scale=(lower, upper, fmin, fmax), want_dec=True)
    _, d2 = oracle.svm_predict(m, scaled.astype(np.float32), want_dec=True)  # scaled values are fp32-exact or near
    assert np.allclose(d1, d2, atol=1e-6)


def test_golden_grsd_model_known_answers(oracle):
    m, scale = _golden()
    assert (m.nr_class, m.total_sv, m.dim, m.gamma) == (18, 641, 21, 0.5)
    assert m.labels.tolist() == [10, 11, 12, 13, 20, 21, 22, 30, 31, 40, 41, 42, 50, 51, 52, 60, 61, 62]
    assert scale[0] == -1.0 and scale[1] == 1.0 and scale[2][0] == 10.0 and scale[3][0] == 14356.0
    # the support vectors are training points in scaled space: most are classified as their own class,
    # and every prediction is one of the model's labels
    truth = np.repeat(m.labels, m.nr_sv).astype(np.float32)
    pred = oracle.svm_predict(m, m.sv.astype(np.float32))
    assert np.isin(pred, m.labels).all()
    assert (pred == truth).sum() == 533
    # raw GRSD counts go through the .scp ranges first; an all-zero histogram clamps to `lower`
    zero = oracle.svm_predict(m, np.zeros((1, 21), np.float32), scale=scale)
    assert zero[0] in m.labels


@pytest.mark.gpu
def test_gpu_svm_predict_matches_oracle(oracle):
    ctx = cab.Context(0)
    m, scale = _golden()
    rng = np.random.default_rng(11)
    # synthetic GRSD-like count vectors spanning the scale ranges (some outside, some zero)
    span = np.maximum(scale[3], 1.0)
    F = (rng.random((3000, 21)) ** 3 * span * 1.2).astype(np.float32)
    F[rng.random(F.shape) < 0.3] = 0
    with pytest.raises(cab.CabError, match="no model"):
        ctx._check(ctx._L.cab_svm_predict(ctx._h, None, 0, 21, None, None), "cab_svm_predict")
    for sc in (scale, None):
        ctx.svm_set_model(m, sc)
        got, gdec = ctx.svm_predict(F, want_dec=True)
        want, wdec = oracle.svm_predict(m, F, scale=sc, want_dec=True)
        # fp64 in libsvm's summation order; exp() may differ in its last bit
        assert np.max(np.abs(gdec - wdec)) < 1e-12
        sure = np.min(np.abs(wdec), axis=1) > 1e-9
        assert sure.mean() > 0.99 and np.array_equal(got[sure], want[sure])
    # a generated model with another shape (k = 6, dim = 21 -> 9)
    _, m2, Xt = _sk_model(seed=8, k=5, dim=9, n=400)
    ctx.svm_set_model(m2)
    assert np.array_equal(ctx.svm_predict(Xt), oracle.svm_predict(m2, Xt))
    with pytest.raises(cab.CabError, match="features"):
        ctx.svm_predict(F)
    ctx.close()


@pytest.mark.gpu
def test_gpu_cluster_to_class_stays_on_device(oracle):
    """cab_grsd_batch followed by cab_svm_predict_grsd: the histograms never leave the GPU; the classes
    equal the oracle's SVM applied to the histograms."""
    ctx = cab.Context(0, exact=True)
    m, scale = _golden()
    xyz, off = synth.clusters(24, 1300, 4000, seed_extra=4)
    hist = ctx.grsd_batch(xyz, off, 0.025, r_normals=0.02)
    ctx.svm_set_model(m, scale)
    got = ctx.svm_predict_grsd(len(off) - 1)
    want = oracle.svm_predict(m, hist.astype(np.float32), scale=scale)
    assert np.array_equal(got, want)
    assert np.array_equal(got, ctx.svm_predict(hist.astype(np.float32)))
    ctx.close()
