"""CPU: the .mars writer reproduces the compiler's conventions -- the width-0.25 YOLOv5 graph
equals the shipped yolov5n_int8.mars tables one for one; files round-trip; the oracle and the
reference library load generated files identically."""
import numpy as np
import pytest

from conftest import shipped


def test_yolov5n_tables_match_shipped(pkg):
    mf = pkg.marsfile
    ref = mf.MarsFile.load(shipped("yolov5n_int8.mars"))
    gen = mf.build_yolov5(width=0.25)
    assert len(ref.tensors) == len(gen.tensors) == 378 and len(ref.layers) == len(gen.layers) == 230
    for a, b in zip(ref.tensors, gen.tensors):
        assert (a.id, a.name, a.shape, bool(a.data_size)) == (b.id, b.name, b.shape, b.data is not None)
    for a, b in zip(ref.layers, gen.layers):
        assert (a.id, a.type, list(a.inputs), list(a.outputs)) == (b.id, b.type, b.inputs, b.outputs)
        assert a.params.rstrip(b"\0") == b.params.rstrip(b"\0")
    assert ref.inputs == gen.inputs and ref.outputs == gen.outputs


def test_yolov5s_work_counts(pkg):
    m = pkg.marsfile.build_yolov5(width=0.5)
    assert m.conv_macs() == 8216780800  # 8.2168 GMAC
    conv_b, other_b = m.layer_bytes()
    assert conv_b == 68032816


def test_roundtrip_and_determinism(pkg):
    mf = pkg.marsfile
    a = mf.build_yolov5(width=0.125, size=160, seed=9).to_bytes()
    b = mf.build_yolov5(width=0.125, size=160, seed=9).to_bytes()
    assert a == b
    assert mf.MarsFile.from_bytes(a).to_bytes() == a
    hdr = np.frombuffer(a[:76], dtype=np.uint8)
    assert hdr[:4].tobytes() == b"MARS"


@pytest.mark.parametrize("kind", ["conv", "sigmoid", "leaky", "add", "mul", "maxpool", "upsample", "concat", "batchnorm"])
def test_reference_and_oracle_agree_on_generated_micro_models(pkg, ob, rb, kind):
    blob = pkg.marsfile.build_single_layer(kind).to_bytes()
    r = rb.RefRuntime(blob)
    m = ob.OracleModel(blob)
    rng = np.random.default_rng(1)
    a = m.arena()
    fill = rng.integers(0, 256, size=m.arena_bytes - m.weights_size, dtype=np.uint8)
    a[m.weights_size:] = fill
    r.arena()[m.weights_size: m.arena_bytes] = fill
    r.run()
    m.run()
    assert np.array_equal(r.arena()[: m.arena_bytes], a)
    r.close()
    m.close()
