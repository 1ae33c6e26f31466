"""Integer conv requantisation (the RQ 3 epilogue of csrc/conv_tc.cu) against the reference arithmetic, by brute force.

reference: src/mars/mxu_conv.c:663-666 -- r = clamp_int8((int32)(sc + (sc >= 0 ? 0.5f : -0.5f))), sc = (float)t * scale.
The library's host-side fit (mars_b200_requant_fit) claims clamp(floor((t * m + c) / 2^(32 + s))) == r for EVERY |t| <= tmax; here
every t of the unsaturated range (and samples of the saturated one) is evaluated both ways on the CPU.  No GPU needed."""
import ctypes as C

import numpy as np
import pytest

from conftest import load_package

lib = load_package().lib()


def ref_np(t, cs):
    """numpy restatement of the reference expression, one float32 rounding per operation"""
    cs = np.float32(cs)
    sc = t.astype(np.float32) * cs
    h = sc + np.where(sc >= 0, np.float32(0.5), np.float32(-0.5)).astype(np.float32)
    return np.clip(np.trunc(h.astype(np.float64)), -128, 127).astype(np.int64)


def fit(cs, tmax):
    m, s, c = C.c_int(), C.c_int(), C.c_longlong()
    ok = lib.mars_b200_requant_fit(C.c_float(cs), C.c_longlong(tmax), C.byref(m), C.byref(s), C.byref(c))
    return (m.value, s.value, c.value) if ok else None


def g_np(t, m, s, c):
    x = t.astype(np.int64) * np.int64(m) + np.int64(c)  # |t * m| < 2^62 for |t| < 2^31, m < 2^31
    return np.clip(x >> np.int64(32 + s), -128, 127)


# the headline model's conv scales (marsfile.py: 28 / (sqrt(K) * 73.6 * 24)) and the accumulator bounds its K values allow
MODEL_K = [108, 32, 64, 128, 256, 288, 512, 576, 1024, 1152, 2304, 4608]


def test_most_model_scales_fit():
    n = sum(fit(float(np.float32(28.0 / (np.sqrt(K) * 73.6 * 24.0))), K * 127 * 128 + 1500) is not None for K in MODEL_K)
    assert n >= 7, n  # the rest keep the float epilogue


@pytest.mark.parametrize("K", MODEL_K)
def test_fit_matches_reference_on_model_scales(K):
    cs = float(np.float32(28.0 / (np.sqrt(K) * 73.6 * 24.0)))
    tmax = K * 127 * 128 + 1500
    r = fit(cs, tmax)
    if r is None:
        pytest.skip("no integer triple for scale %g: the layer keeps a float variant" % cs)
    m, s, c = r
    assert 0 < m < 2 ** 31 and 0 <= s <= 24
    span = int(min(tmax, 140.0 / cs + 16))
    t = np.arange(-span, span + 1, dtype=np.int64)
    assert np.array_equal(g_np(t, m, s, c), ref_np(t, cs))
    far = np.array([-tmax, -tmax + 1, -(span + tmax) // 2, (span + tmax) // 2, tmax - 1, tmax], dtype=np.int64)
    assert np.array_equal(g_np(far, m, s, c), ref_np(far, cs))


def test_fit_random_scales_and_ranges():
    rng = np.random.default_rng(11)
    fitted = 0
    for i in range(60):
        cs = float(np.float32(np.exp(rng.uniform(np.log(2e-5), np.log(0.4)))))
        tmax = int(rng.choice([50, 1000, 40000, 2 ** 22 - 1, 2 ** 26, 2 ** 31 - 1]))
        r = fit(cs, tmax)
        if r is None:
            continue
        fitted += 1
        m, s, c = r
        span = int(min(tmax, 140.0 / cs + 16, 3_000_000))
        t = np.arange(-span, span + 1, dtype=np.int64)
        assert np.array_equal(g_np(t, m, s, c), ref_np(t, cs)), (cs, tmax, r)
        far = np.unique(np.clip(rng.integers(-tmax, tmax + 1, size=2000), -tmax, tmax)).astype(np.int64)
        assert np.array_equal(g_np(far, m, s, c), ref_np(far, cs)), (cs, tmax, r)
    assert fitted >= 40  # scales with full 24-bit mantissas almost always fit


def test_fit_declines_what_it_does_not_handle():
    assert fit(0.0, 1000) is None and fit(-0.01, 1000) is None and fit(0.75, 1000) is None and fit(float("nan"), 1000) is None
    assert fit(0.01, 0) is None and fit(0.01, 2 ** 31) is None


def test_power_of_two_scale_has_exact_ties_on_both_sides():
    # 2.5 -> 3 and -2.5 -> -3 (half away from zero): a rounding addend alone cannot reproduce both, a tilted multiplier can
    for tmax in (400, 1 << 20):
        r = fit(2.0 ** -10, tmax)
        assert r is not None
        t = np.arange(-min(tmax, 140 * 1024), min(tmax, 140 * 1024) + 1, dtype=np.int64)
        assert np.array_equal(g_np(t, *r), ref_np(t, 2.0 ** -10))


def test_numpy_restatement_equals_the_library_reference():
    rng = np.random.default_rng(5)
    for cs in [0.0123, 0.00153, 3.1e-4, 0.2]:
        t = rng.integers(-2 ** 31, 2 ** 31, size=3000).astype(np.int64)
        t[:400] = rng.integers(int(-200 / cs), int(200 / cs), size=400)
        want = np.array([lib.mars_b200_requant_ref(int(v), C.c_float(cs)) for v in t], dtype=np.int64)
        assert np.array_equal(ref_np(t, cs), want)
