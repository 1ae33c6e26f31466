"""Pre-processing in front of mars_run: letterbox resize + int8 packing.

Checker = the reference's own load_image() (src/mars/mars_yolo_test.c:40-77 with the vendored stb headers, compiled
unmodified into oracle/_ref) and the golden hashes it produced (tests/golden/preproc.json, make_preproc_golden.py).
CPU tests cover the host half of the product (the tap lists) through a numpy restatement of the two device sums; the GPU
tests run the kernels through the C-ABI.  Bit-exact throughout.
"""
import hashlib
import json
import os
import sys

import numpy as np
import pytest

from util import GOLDEN_DIR

sys.path.insert(0, GOLDEN_DIR)
from make_preproc_golden import frame  # noqa: E402

PRE_ALL = json.load(open(os.path.join(GOLDEN_DIR, "preproc.json")))
PRE = {k: v for k, v in PRE_ALL.items() if "case" in v}      # int8 tensors (src/mars/mars_yolo_test.c load_image)
RGBA = {k: v for k, v in PRE_ALL.items() if "rgba" in v}     # RGBA frames (examples/yolo_detect.cpp load_and_preprocess_image)
SMALL = [k for k, v in PRE.items() if v["case"][0] * v["case"][1] <= 10000]
ALL = sorted(PRE)


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def axis_sum(a, start, src, w):
    """out[o] = sum of a[src[t]] * w[t] over the taps of o, added onto 0.0f in list order (float32 at every step)"""
    out = np.zeros((len(start) - 1,) + a.shape[1:], np.float32)
    for o in range(len(start) - 1):
        acc = np.zeros(a.shape[1:], np.float32)
        for t in range(start[o], start[o + 1]):
            acc = acc + a[src[t]] * np.float32(w[t])
        out[o] = acc
    return out


def letterbox_numpy(capi, rgb, tw, th, nhwc, rgba=False):
    h, w, _ = rgb.shape
    scale = min(np.float32(tw) / np.float32(w), np.float32(th) / np.float32(h))
    nw, nh = int(np.float32(w) * scale), int(np.float32(h) * scale)
    px, py = (tw - nw) // 2, (th - nh) // 2
    dec = rgb.astype(np.float32) / np.float32(255.0)
    H = axis_sum(dec.transpose(1, 0, 2), *capi.resize_taps(w, nw)).transpose(1, 0, 2)
    V = axis_sum(H, *capi.resize_taps(h, nh))
    u = ((np.clip(V, 0, 1) * np.float32(255.0)).astype(np.float64) + 0.5).astype(np.int64).astype(np.uint8)
    if rgba:
        out = np.full((th, tw, 4), 114, np.uint8)
        out[py:py + nh, px:px + nw, :3] = u
        out[py:py + nh, px:px + nw, 3] = 0
        return out.reshape(-1)
    out = np.full((th, tw, 3), -17, np.int8)
    out[py:py + nh, px:px + nw] = (u.astype(np.int16) - 128).astype(np.int8)
    return (out if nhwc else out.transpose(2, 0, 1)).reshape(-1)


@pytest.mark.parametrize("tag", ALL)
def test_reference_load_image_reproduces_the_golden_hashes(rb, tag):
    w, h, tw, th, nhwc, seed = PRE[tag]["case"]
    t = rb.ref_load_image(frame(w, h, seed), tw, th, bool(nhwc))
    assert sha(t) == PRE[tag]["sha256"] and int((t == -17).sum()) == PRE[tag]["border"]


@pytest.mark.parametrize("tag", SMALL)
def test_host_tap_lists_reproduce_the_reference_resize(pkg, tag):
    w, h, tw, th, nhwc, seed = PRE[tag]["case"]
    assert sha(letterbox_numpy(pkg.capi, frame(w, h, seed), tw, th, nhwc)) == PRE[tag]["sha256"]


@pytest.mark.parametrize("tag", sorted(RGBA))
def test_reference_rgba_preprocess_reproduces_the_golden_hashes(rb, tag):
    w, h, seed = RGBA[tag]["rgba"]
    t = rb.ref_preprocess_rgba(frame(w, h, seed))
    assert sha(t) == RGBA[tag]["sha256"] and int((t.reshape(-1, 4)[:, 3] == 114).sum()) == RGBA[tag]["border"]


def test_host_tap_lists_reproduce_the_reference_rgba_frame(pkg):
    tag = "rgba_64x48_s21"
    w, h, seed = RGBA[tag]["rgba"]
    assert sha(letterbox_numpy(pkg.capi, frame(w, h, seed), 640, 640, True, rgba=True)) == RGBA[tag]["sha256"]


def test_host_tap_lists_random_sizes_against_the_reference_function(pkg, rb):
    """enlarging, shrinking, one-pixel-off and mixed axes: the numpy sums over the product's tap lists equal the reference's
    load_image for frame / tensor sizes drawn at random"""
    rng = np.random.default_rng(12)
    done = 0
    while done < 24:
        w, h = int(rng.integers(1, 90)), int(rng.integers(1, 90))
        tw, th = int(rng.integers(8, 72)), int(rng.integers(8, 72))
        scale = min(np.float32(tw) / np.float32(w), np.float32(th) / np.float32(h))
        if int(np.float32(w) * scale) < 1 or int(np.float32(h) * scale) < 1:
            continue
        f = rng.integers(0, 256, (h, w, 3), dtype=np.uint8)
        nhwc = bool(rng.integers(0, 2))
        want = rb.ref_load_image(f, tw, th, nhwc)
        got = letterbox_numpy(pkg.capi, f, tw, th, nhwc)
        assert got.tobytes() == want.tobytes(), (w, h, tw, th, nhwc)
        done += 1


def test_tap_lists_are_well_formed(pkg):
    for (i, o) in [(1, 1), (1, 7), (7, 1), (640, 640), (1920, 640), (37, 64), (1000, 3)]:
        start, src, w = pkg.capi.resize_taps(i, o)
        assert start[0] == 0 and start[-1] == len(src) == len(w) and np.all(np.diff(start) >= 1)
        assert src.min() >= 0 and src.max() < i
        sums = np.add.reduceat(w.astype(np.float64), start[:-1])
        assert np.allclose(sums, 1.0, atol=1e-5)
    with pytest.raises(ValueError):
        pkg.capi.resize_taps(0, 4)


@pytest.mark.gpu
@pytest.mark.parametrize("tag", ALL)
def test_letterbox_kernel_matches_the_reference(pkg, tag):
    w, h, tw, th, nhwc, seed = PRE[tag]["case"]
    got = pkg.capi.letterbox(frame(w, h, seed), tw, th, bool(nhwc))
    assert sha(got) == PRE[tag]["sha256"]


@pytest.mark.gpu
@pytest.mark.parametrize("tag", sorted(RGBA))
def test_rgba_letterbox_kernel_matches_the_reference(pkg, tag):
    w, h, seed = RGBA[tag]["rgba"]
    assert sha(pkg.capi.letterbox_rgba(frame(w, h, seed))) == RGBA[tag]["sha256"]


@pytest.mark.gpu
def test_letterbox_random_frames_against_the_reference_function(pkg, rb):
    rng = np.random.default_rng(5)
    for _ in range(12):
        w, h = int(rng.integers(1, 400)), int(rng.integers(1, 400))
        tw, th = int(rng.choice([32, 96, 160])), int(rng.choice([32, 96, 160]))
        f = rng.integers(0, 256, (h, w, 3), dtype=np.uint8)
        if int(np.float32(w) * min(np.float32(tw) / np.float32(w), np.float32(th) / np.float32(h))) < 1:
            continue
        if int(np.float32(h) * min(np.float32(tw) / np.float32(w), np.float32(th) / np.float32(h))) < 1:
            continue
        nhwc = bool(rng.integers(0, 2))
        assert pkg.capi.letterbox(f, tw, th, nhwc).tobytes() == rb.ref_load_image(f, tw, th, nhwc).tobytes(), (w, h, tw, th, nhwc)


@pytest.mark.gpu
def test_preprocess_batch_fills_the_input_tensors_of_the_slots(pkg, rb):
    """frames -> slots on the device == reference load_image -> upload; then the detections agree too"""
    from conftest import shipped
    gm = pkg.MarsModel(shipped("yolov5n_int8.mars"), batch=4)
    frames = np.stack([frame(500, 375, 100 + i) for i in range(3)])
    want = np.stack([rb.ref_load_image(f, 640, 640, False) for f in frames])
    gm.upload_inputs(0, 3, want, want.shape[1])
    gm.step_resident(0, 3, 0.45, True)
    d_ref, c_ref = gm.download_detections(0, 3)
    off = gm.tensor_offset(gm.m.contents.header.input_tensor_ids[0])
    gm.preprocess(1, frames)  # deliberately not slot 0
    for i in range(3):
        a = gm.arena_download(1 + i)
        assert a[off:off + want.shape[1]].tobytes() == want[i].tobytes(), i
    gm.step_resident(1, 3, 0.45, True)
    d, c = gm.download_detections(1, 3)
    assert c.tolist() == c_ref.tolist()
    for i in range(3):
        assert d[i][:c[i]].tobytes() == d_ref[i][:c_ref[i]].tobytes()
    with pytest.raises(pkg.capi.MarsError):
        gm.preprocess(3, frames)  # slots 3..5 of 4
