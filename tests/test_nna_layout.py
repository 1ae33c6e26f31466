"""NNA-native layouts: NMHWSOIB2 weights and NDHWC32 features.

Checker = oracle/nna_layout.py, a restatement of mars-compiler/src/mars_format.rs:436-531 that is pinned by
tests/golden/nna_layout.npz: packed bytes the reference's own Python unpackers (mgk-decompiler/mgk_decompiler.py:470-540,
scripts/extract_weights_nmhwsoib2.py:52-80) turn back into the weights (make_nna_layout_golden.py asserts it when the fixtures
are made).  CPU tests: oracle against the fixtures and its loop transcription; GPU tests: the C-ABI against the oracle, bit-exact,
on ragged channel counts, odd plane sizes and a feature map at the headline model's largest size.
"""
import ctypes as C
import os

import numpy as np
import pytest

from oracle import nna_layout as nl
from util import GOLDEN_DIR

G = np.load(os.path.join(GOLDEN_DIR, "nna_layout.npz"))
NW = sum(1 for k in G.files if k.startswith("w"))
NX = sum(1 for k in G.files if k.startswith("x"))


def test_oracle_matches_the_pinned_fixtures():
    assert NW >= 6 and NX >= 4
    for k in range(NW):
        w, p = G["w%d" % k], G["p%d" % k]
        assert np.array_equal(nl.pack_nmhwsoib2(w), p)
        assert np.array_equal(nl.unpack_nmhwsoib2(p, *w.shape), w)
        assert p.size == nl.nmhwsoib2_size(*w.shape)
    for k in range(NX):
        x, n = G["x%d" % k], G["n%d" % k]
        assert np.array_equal(nl.pack_ndhwc32(x), n)
        assert np.array_equal(nl.unpack_ndhwc32(n, *x.shape), x)
        assert n.size == nl.ndhwc32_size(*x.shape)


def test_oracle_vectorised_maps_equal_the_loop_transcriptions():
    rng = np.random.default_rng(5)
    for _ in range(25):
        co, ci, kh, kw = (int(rng.integers(1, 70)), int(rng.integers(1, 70)), int(rng.integers(1, 4)), int(rng.integers(1, 4)))
        w = rng.integers(-128, 128, size=(co, ci, kh, kw), dtype=np.int8)
        p = nl.pack_nmhwsoib2(w)
        assert np.array_equal(p, nl.pack_nmhwsoib2_loops(w))
        # padding (absent channels) is zero: mars_format.rs:450 starts from vec![0u8; packed_size]
        assert int(np.count_nonzero(p)) <= int(np.count_nonzero(w))
    for _ in range(25):
        n, c, h, w_ = (int(rng.integers(1, 3)), int(rng.integers(1, 70)), int(rng.integers(1, 9)), int(rng.integers(1, 9)))
        x = rng.integers(0, 256, size=(n, c, h, w_), dtype=np.uint8)
        assert np.array_equal(nl.pack_ndhwc32(x), nl.pack_ndhwc32_loops(x))


def test_size_helpers_need_no_device(pkg):
    L = pkg.lib()
    for a in [(32, 32, 1, 1), (255, 128, 1, 1), (33, 65, 3, 3), (16, 3, 6, 6)]:
        assert L.mars_b200_nmhwsoib2_size(*a) == nl.nmhwsoib2_size(*a)
    for a in [(1, 32, 4, 4), (2, 40, 5, 7), (1024, 255, 80, 80)]:
        assert L.mars_b200_ndhwc32_size(*a) == nl.ndhwc32_size(*a)
    assert L.mars_b200_nmhwsoib2_size(0, 32, 1, 1) == 0 and L.mars_b200_ndhwc32_size(1, -1, 4, 4) == 0


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p)


@pytest.mark.gpu
def test_weight_packer_matches_fixtures_and_oracle(pkg):
    L = pkg.lib()
    rng = np.random.default_rng(6)
    cases = [G["w%d" % k] for k in range(NW)]
    cases += [rng.integers(-128, 128, size=s, dtype=np.int8) for s in [(1, 1, 1, 1), (31, 33, 3, 3), (512, 256, 3, 3), (255, 512, 1, 1), (96, 58, 5, 5)]]
    for w in cases:
        co, ci, kh, kw = w.shape
        w = np.ascontiguousarray(w)
        packed = np.full(nl.nmhwsoib2_size(co, ci, kh, kw), 0xAB, np.uint8)
        assert L.mars_b200_pack_weights_nmhwsoib2(_ptr(w), co, ci, kh, kw, _ptr(packed)) == 0
        assert np.array_equal(packed, nl.pack_nmhwsoib2(w)), w.shape
        back = np.zeros_like(w)
        assert L.mars_b200_unpack_weights_nmhwsoib2(_ptr(packed), co, ci, kh, kw, _ptr(back)) == 0
        assert np.array_equal(back, w)


@pytest.mark.gpu
def test_feature_converter_matches_fixtures_and_oracle(pkg):
    L = pkg.lib()
    rng = np.random.default_rng(7)
    cases = [G["x%d" % k] for k in range(NX)]
    cases += [rng.integers(0, 256, size=s, dtype=np.uint8) for s in [(1, 1, 1, 1), (2, 58, 13, 11), (1, 255, 80, 80), (3, 64, 20, 20), (1, 33, 127, 3), (2, 32, 16, 129)]]
    for x in cases:
        n, c, h, w_ = x.shape
        x = np.ascontiguousarray(x)
        nat = np.full(nl.ndhwc32_size(n, c, h, w_), 0xCD, np.uint8)
        assert L.mars_b200_nchw_to_ndhwc32(_ptr(x), n, c, h, w_, _ptr(nat)) == 0
        assert np.array_equal(nat, nl.pack_ndhwc32(x)), x.shape
        back = np.zeros_like(x)
        assert L.mars_b200_ndhwc32_to_nchw(_ptr(nat), n, c, h, w_, _ptr(back)) == 0
        assert np.array_equal(back, x)


@pytest.mark.gpu
def test_feature_round_trip_at_the_headline_size(pkg):
    """32 x 320 x 320 (the stem's output of the 640^2 model), batch 8: converted on the device, checked by round trip and by
    spot-checking the index map of mars_format.rs:521-522 on sampled elements"""
    L = pkg.lib()
    n, c, h, w_ = 8, 32, 320, 320
    x = np.random.default_rng(8).integers(0, 256, size=(n, c, h, w_), dtype=np.uint8)
    nat = np.zeros(nl.ndhwc32_size(n, c, h, w_), np.uint8)
    assert L.mars_b200_nchw_to_ndhwc32(_ptr(x), n, c, h, w_, _ptr(nat)) == 0
    idx = np.random.default_rng(9).integers(0, x.size, size=4096)
    nn, cc, hh, ww = np.unravel_index(idx, x.shape)
    dst = (((nn * ((c + 31) // 32) + cc // 32) * h + hh) * w_ + ww) * 32 + cc % 32
    assert np.array_equal(nat[dst], x.reshape(-1)[idx])
    back = np.zeros_like(x)
    assert L.mars_b200_ndhwc32_to_nchw(_ptr(nat), n, c, h, w_, _ptr(back)) == 0
    assert np.array_equal(back, x)


@pytest.mark.gpu
def test_bad_arguments_are_refused(pkg):
    L = pkg.lib()
    buf = np.zeros(2048, np.uint8)
    assert L.mars_b200_pack_weights_nmhwsoib2(None, 32, 32, 1, 1, _ptr(buf)) == -1
    assert L.mars_b200_pack_weights_nmhwsoib2(_ptr(buf), 0, 32, 1, 1, _ptr(buf)) == -1
    assert L.mars_b200_nchw_to_ndhwc32(_ptr(buf), 1, 32, 0, 4, _ptr(buf)) == -1
