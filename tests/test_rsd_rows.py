"""The arithmetic behind the fast RSD kernel's bin rows (csrc/cab_rsd.cu, fast_chunk), restated in numpy: two truncated
products root * scale, with the scale ndiv / radius biased 2^-18 low and high, bracket the reference's bin
floor(ndiv * (double)sqrtf(d2) / radius) (cloud_algos/src/radius_estimation.cpp:165-168) for every d2, whatever the (bounded)
error of the approximate square root; where the two floors agree the kernel uses them without looking at a threshold, so
that agreement must imply the reference's bin.  CPU only: guards the constants, not the device code (tests/test_gpu_parity.py
compares the kernel's radii with the oracle)."""
import numpy as np
import pytest


def _ref_bin(d2, ndiv, radius):
    return np.floor(ndiv * np.sqrt(d2.astype(np.float32)).astype(np.float64) / radius).astype(np.int64)


def _trunc_mul(a32, b32):
    """fma.rz(a, b, 2^23) - 2^23 for 0 <= a * b < 2^22: the floor of the exact product (the float product of two floats is
    exact in double)."""
    return np.floor(a32.astype(np.float64) * np.float64(b32)).astype(np.int64)


@pytest.mark.parametrize("radius,ndiv", [(0.02, 10), (0.03, 10), (0.02, 5), (0.015, 64), (0.1, 1), (0.025, 7)])
def test_two_floors_bracket_the_reference_bin(radius, ndiv):
    rng = np.random.default_rng(7)
    r2 = np.float32(np.float32(radius) * np.float32(radius))
    # d2 spread over [0, 4 r^2] (misses included), plus the values around every bin edge and the radius
    d2 = (rng.random(400_000) * 4.0 * float(r2)).astype(np.float32)
    edges = np.array([(b * radius / ndiv) ** 2 for b in range(ndiv + 2)], np.float64).astype(np.float32)
    near = np.concatenate([e.view(np.uint32).astype(np.int64)[None] + np.arange(-40, 41)[:, None] for e in edges[1:]], axis=1)
    near = near[(near > 0)].astype(np.uint32).view(np.float32)
    d2 = np.concatenate([d2, near, np.array([0.0, float(r2)], np.float32), np.nextafter(r2, np.float32(1)).reshape(1)])
    ref = _ref_bin(d2, ndiv, radius)
    s_lo = np.float32(ndiv / radius * (1.0 - 1.0 / 262144.0))
    s_hi = np.float32(ndiv / radius * (1.0 + 1.0 / 262144.0))
    exact = np.sqrt(d2.astype(np.float64))
    some_disagree = False
    for rel in (-2.0 ** -22, 0.0, 2.0 ** -22):  # sqrt.approx: relative error within 2^-22 either way
        root = (exact * (1.0 + rel)).astype(np.float32)
        lo, hi = _trunc_mul(root, s_lo), _trunc_mul(root, s_hi)
        assert np.all(lo <= ref) and np.all(ref <= hi), "the estimates do not bracket the reference's bin"
        assert np.all(hi - lo <= 1)
        agree = lo == hi
        assert np.all(ref[agree] == lo[agree])
        # a neighbour (d2 <= r2) never gets a low estimate beyond the last bin: its row needs no clamp
        assert np.all(lo[d2 <= r2] <= ndiv - 1)
        # beyond the radius the high estimate reaches the spare row, so agreement there means "spare row"
        assert np.all(hi[d2 > r2] >= ndiv)
        some_disagree |= bool(np.any(~agree))
        # ... and the random part of the sample is almost never ambiguous
        assert np.mean(~agree[:400_000]) < 1e-3
    assert some_disagree  # the edge-adjacent values do exercise the exact-threshold path


def test_row_address_arithmetic_wraps_consistently():
    """bits(2^23 + n) * stride + K (mod 2^32) is the n-th row of the lane's column for K = column - bits(2^23) * stride."""
    magic = int(np.float32(8388608.0).view(np.uint32))
    for stride in (256, 384):
        for col in (0x1234, 0xFFF0, 0x8000):
            k = (col - magic * stride) % (1 << 32)
            for n in (0, 1, 9, 10, 64):
                v = int(np.float32(8388608.0 + n).view(np.uint32))
                assert (v * stride + k) % (1 << 32) == col + n * stride
