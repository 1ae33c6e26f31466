"""CPU: pin the oracle (oracle/mars_oracle.c) against the committed golden vectors (outputs of
the reference's own C path, tests/golden/make_golden.py) and, where oracle/_ref/libmars_ref.so
is present, against the reference library directly (whole arena, run 1 and run 2)."""
import os

import numpy as np
import pytest

from conftest import shipped
from util import GOLDEN, GOLDEN_DIR, make_input, numel, sha

MODEL_CASES = [k for k, v in GOLDEN.items() if "model" in v]
FAST = [k for k in MODEL_CASES if not k.startswith("yolov5n/f32")]


@pytest.mark.parametrize("case", MODEL_CASES)
def test_oracle_matches_golden(ob, case):
    g = GOLDEN[case]
    m = ob.OracleModel(shipped(g["model"]), arena_bytes=g["arena"])
    d = m.tensor_desc(m.input_index())
    x = make_input(g["pattern"], numel(d))
    for run in (1, 2):
        m.set_input(x)
        m.run()
        r = g["run%d" % run]
        assert sha(m.output_bytes()) == r["output_sha256"], "%s run %d output" % (case, run)
        assert sha(m.arena()[: r["arena_bytes_hashed"]]) == r["arena_sha256"], "%s run %d arena" % (case, run)
    if "dets" in g:
        o = m.output_bytes().view(np.int8)
        od = m.tensor_desc(m.output_index())
        raw = ob.parse_output(o, od.shape[1], od.scale)
        kept = ob.nms(raw)
        gold = np.load(os.path.join(GOLDEN_DIR, g["dets"]["file"]))
        assert len(raw) == g["dets"]["raw"] and len(kept) == g["dets"]["kept"]
        assert kept.tobytes() == gold.tobytes()
    m.close()


@pytest.mark.parametrize("case", ["tiny_160_int8/rng", "yolov5nu/p0", "test_simple/f32"])
def test_oracle_matches_reference_library(ob, rb, case):
    g = GOLDEN[case]
    path = shipped(g["model"])
    r = rb.RefRuntime(path, arena_bytes=g["arena"])
    m = ob.OracleModel(path, arena_bytes=g["arena"])
    x = make_input(g["pattern"], numel(m.tensor_desc(m.input_index())))
    for _ in range(2):
        r.set_input(x)
        r.run()
        m.set_input(x)
        m.run()
        assert np.array_equal(r.arena()[: m.arena_bytes], m.arena())
    # per-layer stepping agrees too
    r.close()
    m.close()


@pytest.mark.parametrize("tag", [k for k in GOLDEN if k.startswith("post_")])
def test_oracle_postprocess_matches_golden(ob, tag):
    g = GOLDEN[tag]
    rng = np.random.default_rng(g["seed"])
    heads = {"post_random": rng.integers(-128, 128, size=(4000, 85), dtype=np.int8)}
    heads["post_ties"] = rng.choice(np.array([-128, 0, 60, 127], dtype=np.int8), size=(3000, 85))
    h = heads[tag.rsplit("_s", 1)[0]]
    assert sha(h) == g["head_sha256"]
    raw = ob.parse_output(h, h.shape[0], g["scale"])
    kept = ob.nms(raw)
    assert (len(raw), len(kept)) == (g["raw"], g["kept"])
    assert kept.tobytes() == np.load(os.path.join(GOLDEN_DIR, g["file"])).tobytes()


def test_oracle_postprocess_matches_reference_statics(ob, rb):
    rng = np.random.default_rng(5)
    for trial in range(6):
        n = int(rng.integers(1, 1200))
        h = rng.integers(-128, 128, size=(n, 85), dtype=np.int8)
        if trial % 2:
            h = (h // 40 * 40).astype(np.int8)  # tie-heavy
        scale = float(rng.choice([0.02, 0.1, 1.0]))
        a, b = rb.ref_parse_output(h, n, scale), ob.parse_output(h, n, scale)
        assert a.tobytes() == b.tobytes()
        assert rb.ref_nms(a).tobytes() == ob.nms(b).tobytes()
    # corner-box NMS vs the reference C++ (tie-free input: std::sort's tie order is unspecified)
    boxes = np.zeros(300, dtype=ob.BOX_DTYPE)
    xy = rng.uniform(0, 600, size=(300, 2)).astype(np.float32)
    wh = rng.uniform(5, 120, size=(300, 2)).astype(np.float32)
    boxes["x0"], boxes["y0"] = xy[:, 0], xy[:, 1]
    boxes["x1"], boxes["y1"] = xy[:, 0] + wh[:, 0], xy[:, 1] + wh[:, 1]
    boxes["confidence"] = rng.permutation(300).astype(np.float32) / 300
    boxes["class_id"] = rng.integers(0, 3, size=300)
    lib = rb.RefRuntime.lib()
    d = boxes.copy()
    n = lib.oracle_ref_cpp_nms(d.ctypes.data, len(d), 0.45)
    assert d[:n].tobytes() == ob.nms_corner(boxes, 0.45).tobytes()


def test_mars_math_known_answers(ob):
    """reference examples/mars_math_test.c:38-82 -- the only known-answer test on this path"""
    L = ob.lib()
    a = np.array([1, 2, 3, 4], dtype=np.float32)
    b = np.array([5, 6, 7, 8], dtype=np.float32)
    out = np.zeros(4, dtype=np.float32)
    L.mo_vec_add_f32(out.ctypes.data, a.ctypes.data, b.ctypes.data, 4)
    assert out.tolist() == [6, 8, 10, 12]
    assert L.mo_vec_dot_f32(a.ctypes.data, b.ctypes.data, 4) == 70.0
    A = np.arange(1, 7, dtype=np.float32)
    B = np.arange(7, 13, dtype=np.float32)
    Cm = np.zeros(4, dtype=np.float32)
    L.mo_matmul_f32(Cm.ctypes.data, A.ctypes.data, B.ctypes.data, 2, 3, 2)
    assert Cm.tolist() == [58, 64, 139, 154]


def test_anchor_grid_decode_matches_the_reference_python_formula(ob):
    """the reference only has a Python statement of the 3-head decode (mgk-decompiler/test_yolo_inference.py:
    136-202); its output on seeded heads is committed (tests/golden/anchor_decode.json, make_anchor_golden.py).  The C
    restatement -- which the CUDA kernel is held bit-exact to in test_gpu_postproc.py -- must give the same candidates:
    same count and classes, confidences within 1e-6, boxes within 1e-3 px (numpy's float32 exp vs libm expf)."""
    import json
    import sys
    gdir = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
    sys.path.insert(0, gdir)
    from make_anchor_golden import heads
    G = json.load(open(os.path.join(gdir, "anchor_decode.json")))
    O = ob.lib()
    for seed, want in G["cases"].items():
        dets = np.zeros(4096, dtype=ob.BOX_DTYPE)
        cnt = 0
        for level, h in enumerate(heads(int(seed))):
            g = h.shape[1]
            cnt = O.mo_decode_anchor_grid(np.ascontiguousarray(h).ctypes.data, g, g, G["scale"], level, 0.25, dets.ctypes.data, cnt, len(dets))
        d = dets[:cnt]
        d = d[np.argsort(-d["confidence"], kind="stable")]  # the reference sorts by confidence (stable, all levels together)
        assert cnt == len(want["class_id"]), (seed, cnt)
        assert d["class_id"].tolist() == want["class_id"]
        assert np.allclose(d["confidence"], np.array(want["confidence"]), rtol=0, atol=1e-6)
        got = np.stack([d["x0"], d["y0"], d["x1"], d["y1"]], 1).astype(np.float64)
        assert np.abs(got - np.array(want["boxes"])).max() < 1e-3


# ---- the restatement against the reference library on GENERATED models --------------------------------------
# The GPU parity tests of the generated graphs (yolov5s-shaped headline model, NanoDet-shaped, f32, micro shapes) compare
# CUDA with the restatement (oracle/mars_oracle.c).  These tests close that chain: restatement == libmars_ref.so, whole
# arena, on the same generated files.
def _nan_words(a):
    """float32 NaN bit patterns among the 4-byte words of a byte array"""
    w = a[: a.size // 4 * 4].view(np.uint32)
    return ((w & 0x7F800000) == 0x7F800000) & ((w & 0x007FFFFF) != 0)


def _same_arena(rb, ob, blob, arena, x, nan_payload_free=False):
    """nan_payload_free (float32 models): words that are NaN on both sides may differ in sign / payload -- which operand's
    payload an x86 mulss / addss returns depends on the operand order the C compiler picked, not on the source"""
    r = rb.RefRuntime(blob, arena_bytes=arena)
    m = ob.OracleModel(blob, arena_bytes=arena)
    for run in range(2):
        xr = x if run == 0 else np.roll(x, 7919)
        r.set_input(xr)
        r.run()
        m.set_input(xr)
        m.run()
        a, b = r.arena()[: m.arena_bytes], m.arena()
        if nan_payload_free:
            W = m.weights_size // 4 * 4  # work buffers start at weights_size (4-byte aligned in generated files)
            wa, wb = a[W:], b[W:]
            both_nan = _nan_words(wa) & _nan_words(wb)
            ne = wa[: wa.size // 4 * 4].view(np.uint32) != wb[: wb.size // 4 * 4].view(np.uint32)
            d = np.nonzero(ne & ~both_nan)[0]
            assert d.size == 0, "run %d: %d words differ beyond NaN payloads, first at byte %d" % (run + 1, d.size, W + 4 * int(d[0]))
            continue
        d = np.nonzero(a != b)[0]
        assert d.size == 0, "run %d: %d arena bytes differ, first at %d" % (run + 1, d.size, int(d[0]))
    out = m.output_bytes().copy()
    r.close()
    m.close()
    return out


def test_restatement_equals_reference_on_headline_model(pkg, ob, rb):
    """BASELINE configs[2]: the yolov5s-shaped int8 graph at 640x640 (writer seed 5), image seed 1000"""
    mf = pkg.marsfile
    blob = mf.build_yolov5(width=0.5, size=640, seed=5).to_bytes()
    x = np.random.default_rng(1000).integers(-128, 128, size=3 * 640 * 640, dtype=np.int8)
    _same_arena(rb, ob, blob, mf.ARENA_YOLOV5S_INT8, x)


def test_restatement_equals_reference_on_nhwc_model(pkg, ob, rb):
    """the --nhwc convention of the compiler (conv2d_int8_nhwc_mxu, src/mars/mxu_conv.c:713-757) at yolov5s width, 320 px"""
    mf = pkg.marsfile
    blob = mf.build_yolov5(width=0.5, size=320, seed=5, nhwc=True).to_bytes()
    x = np.random.default_rng(1001).integers(-128, 128, size=3 * 320 * 320, dtype=np.int8)
    _same_arena(rb, ob, blob, mf.ARENA_YOLOV5S_INT8, x)


def test_restatement_equals_reference_on_nanodet_like(pkg, ob, rb):
    mf = pkg.marsfile
    blob = mf.build_nanodet_like(size=320, seed=7).to_bytes()
    x = np.random.default_rng(7).integers(-128, 128, size=3 * 320 * 320, dtype=np.int8)
    _same_arena(rb, ob, blob, 16 << 20, x)


def test_restatement_equals_reference_on_f32_model(pkg, ob, rb):
    """BASELINE configs[3] shape: yolov5s width, float32 (320 px here: the CPU cost of 640 px is paid once, in the GPU test)"""
    mf = pkg.marsfile
    blob = mf.build_yolov5(width=0.5, size=320, seed=6, f32=True).to_bytes()
    x = np.random.default_rng(6).random(3 * 320 * 320, dtype=np.float32)
    _same_arena(rb, ob, blob, mf.ARENA_YOLOV5S_F32, x.view(np.int8), nan_payload_free=True)


from util import MICRO  # noqa: E402


@pytest.mark.parametrize("kind,kw", [(k, kw) for k, kw in MICRO if k != "depthwise"])
def test_restatement_equals_reference_on_micro_models(pkg, ob, rb, kind, kw):
    """(the depthwise micro-model is excluded: the reference's DEPTHWISE_CONV2D is a no-op, the restated kernel is unpinned)"""
    blob = pkg.marsfile.build_single_layer(kind, **kw).to_bytes()
    r = rb.RefRuntime(blob)
    m = ob.OracleModel(blob)
    rng = np.random.default_rng(1)
    fill = rng.integers(0, 256, size=m.arena_bytes - m.weights_size, dtype=np.uint8)
    m.arena()[m.weights_size:] = fill
    r.arena()[m.weights_size: m.arena_bytes] = fill
    r.run()
    m.run()
    assert np.array_equal(r.arena()[: m.arena_bytes], m.arena())
    r.close()
    m.close()
