"""GPU: decode + NMS kernels and the host-pointer seam against the oracle / golden vectors."""
import os

import numpy as np
import pytest

from util import GOLDEN, GOLDEN_DIR

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("tag", [k for k in GOLDEN if k.startswith("post_")])
def test_postprocess_golden(pkg, tag):
    g = GOLDEN[tag]
    rng = np.random.default_rng(g["seed"])
    heads = {"post_random": rng.integers(-128, 128, size=(4000, 85), dtype=np.int8)}
    heads["post_ties"] = rng.choice(np.array([-128, 0, 60, 127], dtype=np.int8), size=(3000, 85))
    h = heads[tag.rsplit("_s", 1)[0]]
    raw = pkg.capi.parse_output(h, h.shape[0], g["scale"])
    kept = pkg.capi.nms(raw)
    assert (len(raw), len(kept)) == (g["raw"], g["kept"])
    assert kept.tobytes() == np.load(os.path.join(GOLDEN_DIR, g["file"])).tobytes()


def test_postprocess_random_vs_oracle(pkg, ob):
    rng = np.random.default_rng(21)
    for trial in range(12):
        n = int(rng.integers(1, 3000))
        h = rng.integers(-128, 128, size=(n, 85), dtype=np.int8)
        if trial % 3 == 1:
            h = (h // 50 * 50).astype(np.int8)
        if trial % 3 == 2:
            h[:, 4] = rng.choice(np.array([-128, -20, 5, 127], dtype=np.int8), size=n)
        scale = float(rng.choice([0.01, 0.05, 0.3, 1.0]))
        maxd = int(rng.choice([1000, 1000, 17, 1]))
        a, b = ob.parse_output(h, n, scale, maxd), pkg.capi.parse_output(h, n, scale, maxd)
        assert a.tobytes() == b.tobytes(), (trial, len(a), len(b))
        for th in (0.45, 0.1):
            assert ob.nms(a, th).tobytes() == pkg.capi.nms(b, th).tobytes(), trial
    assert len(pkg.capi.nms(np.zeros(0, dtype=pkg.capi.DET_DTYPE))) == 0


def test_nms_arbitrary_class_ids_and_special_scores(pkg, ob):
    """class ids outside the member-bitset range take the all-pairs path; NaN / inf / tied scores keep the C loop's order"""
    rng = np.random.default_rng(77)
    sizes = [1, 2, 31, 32, 33, 64, 65, 128, 129, 200, 256, 257, 384, 385, 500, 512, 513, 768, 769, 896, 897, 1000, 1023, 1024]
    for trial, n in enumerate(sizes):  # every layout of the warp sort (4 / 8 / 16 / 32 positions per lane) and each of its phase boundaries
        d = np.zeros(n, dtype=pkg.capi.DET_DTYPE)
        d["x"] = rng.integers(0, 64, n).astype(np.float32)
        d["y"] = rng.integers(0, 64, n).astype(np.float32)
        d["w"] = rng.integers(1, 40, n).astype(np.float32)
        d["h"] = rng.integers(1, 40, n).astype(np.float32)
        d["conf"] = (rng.integers(0, 12, n) / 12).astype(np.float32)  # heavy ties
        if trial % 2:
            d["cls"] = rng.choice(np.array([-7, -1, 0, 3, 126, 127, 5000], dtype=np.int32), n)
        else:
            d["cls"] = rng.integers(-1, 127, n).astype(np.int32)
        if trial >= 4 and n > 8:
            k = rng.integers(0, n, 6)
            d["conf"][k[:2]] = np.nan
            d["conf"][k[2:4]] = np.inf
            d["conf"][k[4:]] = -np.inf
        for th in (0.45, 0.0):
            a, b = ob.nms(d.copy(), th), pkg.capi.nms(d.copy(), th)
            assert a.tobytes() == b.tobytes(), (trial, n, th, len(a), len(b))


def test_large_batches_take_the_256_thread_suppression_kernel_with_identical_results(pkg):
    """more than 592 images per launch switch the suppression to k_nms_suppress; every slot must match the small-batch path"""
    from conftest import shipped
    path = shipped("yolov5n_int8.mars")
    rng = np.random.default_rng(9)
    xs = rng.integers(-128, 128, size=(3, 3 * 640 * 640), dtype=np.int8)
    small = pkg.MarsModel(path, batch=3)
    small.upload_inputs(0, 3, xs, xs.shape[1])
    small.step_resident(0, 3, 0.45, True)
    d_ref, c_ref = small.download_detections(0, 3)
    del small
    n = 600
    big = pkg.MarsModel(path, batch=n)
    for i in range(n):
        big.upload_inputs(i, 1, xs[i % 3], xs.shape[1])
    big.step_resident(0, n, 0.45, True)
    d, c = big.download_detections(0, n)
    assert c_ref.min() > 0
    for i in range(n):
        assert c[i] == c_ref[i % 3] and d[i][:c[i]].tobytes() == d_ref[i % 3][:c[i]].tobytes(), i


def test_corner_nms_scale_and_anchor_decode(pkg, ob):
    rng = np.random.default_rng(6)
    boxes = np.zeros(500, dtype=pkg.capi.BOX_DTYPE)
    xy = rng.uniform(0, 600, size=(500, 2)).astype(np.float32)
    wh = rng.uniform(5, 150, size=(500, 2)).astype(np.float32)
    boxes["x0"], boxes["y0"], boxes["x1"], boxes["y1"] = xy[:, 0], xy[:, 1], xy[:, 0] + wh[:, 0], xy[:, 1] + wh[:, 1]
    boxes["confidence"] = (rng.integers(0, 50, size=500) / 50).astype(np.float32)  # ties: input order kept
    boxes["class_id"] = rng.integers(0, 4, size=500)
    assert pkg.capi.nms_boxes(boxes, 0.45).tobytes() == ob.nms_corner(boxes, 0.45).tobytes()
    L, O = pkg.lib(), ob.lib()
    a, b = boxes.copy(), boxes.copy()
    L.mars_yolo_scale_detections(a.ctypes.data, len(a), 1920, 1080, 640, 640)
    O.mo_scale_detections(b.ctypes.data, len(b), 1920, 1080, 640, 640)
    assert a.tobytes() == b.tobytes()
    for level, g in ((0, 20), (1, 10), (2, 5)):
        head = rng.integers(-128, 128, size=(3, g, g, 85), dtype=np.int8)
        d1 = np.zeros(300, dtype=pkg.capi.BOX_DTYPE)
        d2 = np.zeros(300, dtype=pkg.capi.BOX_DTYPE)
        n1 = L.mars_yolo_decode_anchor_grid(head.ctypes.data, g, g, 0.05, level, 0.25, d1.ctypes.data, 0, 300)
        n2 = O.mo_decode_anchor_grid(head.ctypes.data, g, g, 0.05, level, 0.25, d2.ctypes.data, 0, 300)
        assert n1 == n2 and d1[:n1].tobytes() == d2[:n2].tobytes()


def test_mars_math_and_mxu_seam(pkg, ob):
    """reference examples/mars_math_test.c:38-82 known answers, then random vs the oracle"""
    L, O = pkg.lib(), ob.lib()
    a = np.array([1, 2, 3, 4], dtype=np.float32)
    b = np.array([5, 6, 7, 8], dtype=np.float32)
    out = np.zeros(4, dtype=np.float32)
    L.mars_vec_add_f32(out.ctypes.data, a.ctypes.data, b.ctypes.data, 4)
    assert out.tolist() == [6, 8, 10, 12]
    assert L.mars_vec_dot_f32(a.ctypes.data, b.ctypes.data, 4) == 70.0
    A, B, Cm = np.arange(1, 7, dtype=np.float32), np.arange(7, 13, dtype=np.float32), np.zeros(4, dtype=np.float32)
    L.mars_matmul_f32(Cm.ctypes.data, A.ctypes.data, B.ctypes.data, 2, 3, 2)
    assert Cm.tolist() == [58, 64, 139, 154]
    rng = np.random.default_rng(0)
    x, y = rng.standard_normal(10007).astype(np.float32), rng.standard_normal(10007).astype(np.float32)
    assert L.mars_vec_dot_f32(x.ctypes.data, y.ctypes.data, x.size) == O.mo_vec_dot_f32(x.ctypes.data, y.ctypes.data, x.size)
    for gf, of in (("mxu_add_f32", "mo_vec_add_f32"), ("mxu_sub_f32", "mo_vec_sub_f32"), ("mxu_mul_f32", "mo_vec_mul_f32")):
        g, o = np.zeros_like(x), np.zeros_like(x)
        getattr(L, gf)(g.ctypes.data, x.ctypes.data, y.ctypes.data, x.size)
        getattr(O, of)(o.ctypes.data, x.ctypes.data, y.ctypes.data, x.size)
        assert g.tobytes() == o.tobytes()
    g, o = np.zeros_like(x), np.zeros_like(x)
    L.mxu_relu_f32(g.ctypes.data, x.ctypes.data, x.size)
    O.mo_vec_relu_f32(o.ctypes.data, x.ctypes.data, x.size)
    assert g.tobytes() == o.tobytes()
    M, K, N = 37, 129, 23
    A, B = rng.standard_normal((M, K)).astype(np.float32), rng.standard_normal((K, N)).astype(np.float32)
    g, o = np.zeros((M, N), dtype=np.float32), np.zeros((M, N), dtype=np.float32)
    L.mars_matmul_f32(g.ctypes.data, A.ctypes.data, B.ctypes.data, M, K, N)
    O.mo_matmul_f32(o.ctypes.data, A.ctypes.data, B.ctypes.data, M, K, N)
    assert g.tobytes() == o.tobytes()


def test_conv_seam_entry_points(pkg, rb):
    """conv2d_*_mxu on host pointers vs the reference's own portable bodies"""
    L, R = pkg.lib(), rb.RefRuntime.lib()
    import ctypes as C
    for f in ("conv2d_int8_mxu", "conv2d_int8_nhwc_mxu", "conv2d_float32_mxu"):
        getattr(R, f).restype = None
        getattr(R, f).argtypes = pkg.capi.SIGNATURES[f][1]
    rng = np.random.default_rng(3)
    ic, ih, iw, oc, k, s, p = 6, 15, 13, 9, 3, 2, 1
    oh, ow = (ih + 2 * p - k) // s + 1, (iw + 2 * p - k) // s + 1
    x = rng.integers(-128, 128, size=ic * ih * iw, dtype=np.int8)
    w = rng.integers(-127, 128, size=oc * ic * k * k, dtype=np.int8)
    bias = rng.integers(-3000, 3000, size=oc).astype(np.int32)
    for f in ("conv2d_int8_mxu", "conv2d_int8_nhwc_mxu"):
        a, b = np.zeros(oc * oh * ow, dtype=np.int8), np.zeros(oc * oh * ow, dtype=np.int8)
        args = lambda o: (x.ctypes.data, ih, iw, ic, w.ctypes.data, oc, k, k, bias.ctypes.data, o.ctypes.data, oh, ow, s, s, p, p,
                          C.c_float(0.05), C.c_float(0.003), C.c_float(0.07))
        getattr(L, f)(*args(a))
        getattr(R, f)(*args(b))
        assert a.tobytes() == b.tobytes(), f
    xf, wf = rng.standard_normal(ic * ih * iw).astype(np.float32), rng.standard_normal(oc * ic * k * k).astype(np.float32)
    bf = rng.standard_normal(oc).astype(np.float32)
    a, b = np.zeros(oc * oh * ow, dtype=np.float32), np.zeros(oc * oh * ow, dtype=np.float32)
    args = lambda o: (xf.ctypes.data, ih, iw, ic, wf.ctypes.data, oc, k, k, bf.ctypes.data, o.ctypes.data, oh, ow, s, s, p, p, None)
    L.conv2d_float32_mxu(*args(a))
    R.conv2d_float32_mxu(*args(b))
    assert a.tobytes() == b.tobytes()
