"""tests/golden/make_nna_layout_golden.py -- golden vectors for the NNA-native layouts, made in the build container
where /root/reference exists:  python tests/golden/make_nna_layout_golden.py  ->  tests/golden/nna_layout.npz

The reference's packers are Rust (mars-compiler/src/mars_format.rs:436-531; no Rust toolchain here).  What CAN run is the
reference's Python unpackers -- mgk-decompiler/mgk_decompiler.py::unpack_nmhwsoib2 (quantize_type 8) and
mgk-decompiler/scripts/extract_weights_nmhwsoib2.py::unpack_nmhwsoib2 -- so each fixture stores the OIHW weights, the packed
bytes of the oracle restatement, and this script asserts that BOTH reference unpackers turn those bytes back into the weights.
"""
import importlib.util
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from oracle import nna_layout as nl  # noqa: E402

REF = "/root/reference/mgk-decompiler"


def load(path, name):
    spec = importlib.util.spec_from_file_location(name, path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def main():
    dec = load(os.path.join(REF, "mgk_decompiler.py"), "ref_mgk_decompiler")
    ext = load(os.path.join(REF, "scripts", "extract_weights_nmhwsoib2.py"), "ref_extract_weights")
    rng = np.random.default_rng(84)
    out = {}
    for k, (co, ci, kh, kw) in enumerate([(32, 32, 1, 1), (64, 32, 3, 3), (48, 40, 3, 3), (255, 128, 1, 1), (16, 3, 6, 6), (33, 65, 1, 3)]):
        w = rng.integers(-128, 128, size=(co, ci, kh, kw), dtype=np.int8)
        packed = nl.pack_nmhwsoib2(w)
        assert packed.size == nl.nmhwsoib2_size(co, ci, kh, kw) == ext.nmhwsoib2_size(co, ci, kh, kw)
        if co * ci * kh * kw <= 20000:
            assert np.array_equal(packed, nl.pack_nmhwsoib2_loops(w))
        a = np.asarray(dec.unpack_nmhwsoib2(packed.tobytes(), co, ci, kh, kw, quantize_type=8))
        b = np.asarray(ext.unpack_nmhwsoib2(packed.tobytes(), co, ci, kh, kw))
        assert np.array_equal(a.astype(np.int64), w.astype(np.int64)), ("mgk_decompiler.unpack_nmhwsoib2", co, ci, kh, kw)
        assert np.array_equal(b.astype(np.int64), w.astype(np.int64)), ("extract_weights.unpack_nmhwsoib2", co, ci, kh, kw)
        out["w%d" % k] = w
        out["p%d" % k] = packed
    for k, (n, c, h, w_) in enumerate([(1, 32, 4, 4), (2, 40, 5, 7), (1, 3, 8, 8), (3, 64, 6, 6)]):
        x = rng.integers(0, 256, size=(n, c, h, w_), dtype=np.uint8)
        nat = nl.pack_ndhwc32(x)
        assert nat.size == nl.ndhwc32_size(n, c, h, w_) and np.array_equal(nat, nl.pack_ndhwc32_loops(x))
        out["x%d" % k] = x
        out["n%d" % k] = nat
    np.savez_compressed(os.path.join(HERE, "nna_layout.npz"), **out)
    print("wrote", os.path.join(HERE, "nna_layout.npz"), sum(v.nbytes for v in out.values()), "bytes before compression")


if __name__ == "__main__":
    main()
