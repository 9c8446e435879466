"""Host-side parsing of the files cloud_algos::SVMClassification reads (svm_classification.cpp:83-112):
a libsvm C-SVC model (svm_load_model) and an svm-scale range file (parseScaleParameterFile,
svm_classification.h:129-185).  The C++ plugin (host/src/svm_classification.cpp) holds the same
logic; this module feeds the C ABI (cab_svm_set_model / cab_svm_set_scaling) from Python.

libsvm is a third-party dependency of the reference (manifest.xml:28, not vendored): the model file
format is the published one -- header lines `svm_type c_svc`, `kernel_type rbf`, `gamma g`,
`nr_class k`, `total_sv l`, `rho` (k(k-1)/2 values), `label` (k), `nr_sv` (k), then `SV` followed by
l lines of k-1 coefficients and sparse `index:value` pairs (1-based indices, zeros omitted).
"""
from __future__ import annotations

import dataclasses

import numpy as np


@dataclasses.dataclass
class SvmModel:
    gamma: float
    labels: np.ndarray    # (k,) int32
    nr_sv: np.ndarray     # (k,) int32, support vectors per class, classes in `labels` order
    rho: np.ndarray       # (k(k-1)/2,) float64
    sv_coef: np.ndarray   # (k-1, l) float64
    sv: np.ndarray        # (l, dim) float64, dense (omitted entries are zero)

    @property
    def nr_class(self) -> int:
        return int(self.labels.shape[0])

    @property
    def total_sv(self) -> int:
        return int(self.sv.shape[0])

    @property
    def dim(self) -> int:
        return int(self.sv.shape[1])


def parse_model(text: str, dim: int | None = None) -> SvmModel:
    """Parse a libsvm model file.  Only C-SVC with an RBF kernel is accepted (what the reference's
    svm/*.model files contain); anything else raises ValueError."""
    lines = text.splitlines()
    hdr = {}
    i = 0
    while i < len(lines):
        line = lines[i].strip()
        i += 1
        if line == "SV":
            break
        if not line:
            continue
        key, _, rest = line.partition(" ")
        hdr[key] = rest.split()
    else:
        raise ValueError("libsvm model: no SV section")
    if hdr.get("svm_type") != ["c_svc"] or hdr.get("kernel_type") != ["rbf"]:
        raise ValueError(f"libsvm model: only c_svc / rbf is supported, got {hdr.get('svm_type')} / {hdr.get('kernel_type')}")
    k = int(hdr["nr_class"][0])
    l = int(hdr["total_sv"][0])
    gamma = float(hdr["gamma"][0])
    rho = np.array([float(x) for x in hdr["rho"]], np.float64)
    labels = np.array([int(x) for x in hdr["label"]], np.int32)
    nr_sv = np.array([int(x) for x in hdr["nr_sv"]], np.int32)
    if rho.shape[0] != k * (k - 1) // 2 or labels.shape[0] != k or nr_sv.shape[0] != k or int(nr_sv.sum()) != l:
        raise ValueError("libsvm model: inconsistent header")
    coef = np.zeros((k - 1, l), np.float64)
    rows = []
    maxidx = 0
    for s in range(l):
        tok = lines[i + s].split()
        for c in range(k - 1):
            coef[c, s] = float(tok[c])
        pairs = []
        for t in tok[k - 1:]:
            a, _, b = t.partition(":")
            idx = int(a)
            pairs.append((idx, float(b)))
            maxidx = max(maxidx, idx)
        rows.append(pairs)
    d = max(maxidx, dim or 0)
    sv = np.zeros((l, d), np.float64)
    for s, pairs in enumerate(rows):
        for idx, v in pairs:
            sv[s, idx - 1] = v
    return SvmModel(gamma, labels, nr_sv, rho, coef, sv)


def format_model(m: SvmModel) -> str:
    """Inverse of parse_model (used by the tests to exercise the parser on generated models)."""
    out = ["svm_type c_svc", "kernel_type rbf", f"gamma {m.gamma!r}", f"nr_class {m.nr_class}", f"total_sv {m.total_sv}",
           "rho " + " ".join(repr(float(x)) for x in m.rho), "label " + " ".join(str(int(x)) for x in m.labels),
           "nr_sv " + " ".join(str(int(x)) for x in m.nr_sv), "SV"]
    for s in range(m.total_sv):
        co = " ".join(repr(float(m.sv_coef[c, s])) for c in range(m.nr_class - 1))
        sp = " ".join(f"{j + 1}:{float(v)!r}" for j, v in enumerate(m.sv[s]) if v != 0.0)
        out.append(f"{co} {sp} ")
    return "\n".join(out) + "\n"


def parse_scale(text: str, nr_values: int):
    """parseScaleParameterFile (svm_classification.h:129-185): first token must start with "x", then
    lower and upper, then `index min max` triples; min / max are read through a C float
    (`float fmin, fmax`, :168) before being stored as doubles; indices beyond nr_values are ignored and
    features missing from the file keep min == max == 0 (scaleFeature then returns 0, :72-73).
    Returns (lower, upper, fmin (nr_values,), fmax (nr_values,)) or None on failure."""
    tok = text.split()
    if not tok or not tok[0].startswith("x"):
        return None
    lower, upper = float(tok[1]), float(tok[2])
    fmin = np.zeros(nr_values, np.float64)
    fmax = np.zeros(nr_values, np.float64)
    rest = tok[3:]
    for t in range(0, len(rest) - 2, 3):
        idx = int(rest[t])
        if idx <= nr_values:
            fmin[idx - 1] = float(np.float32(float(rest[t + 1])))
            fmax[idx - 1] = float(np.float32(float(rest[t + 2])))
    return lower, upper, fmin, fmax
