"""tests/golden/make_preproc_golden.py -- golden vectors of the pre-processing step.

Runs the REFERENCE's own load_image() (src/mars/mars_yolo_test.c:40-77, compiled unmodified into
oracle/_ref/libmars_ref.so together with the vendored stb headers) on seeded random frames and records the sha256 of the
int8 tensor it produces.  Frames are regenerated from the seed by the tests.
Usage (build container, where /root/reference exists):  python tests/golden/make_preproc_golden.py
"""
import hashlib
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from oracle import refbind as rb  # noqa: E402

# (frame w, frame h, tensor w, tensor h, nhwc, seed): shrinking, enlarging, identity, one-axis-identity, tiny, odd sizes
CASES = [
    (64, 48, 32, 32, 0, 1), (48, 64, 32, 32, 1, 2), (20, 14, 32, 32, 0, 3), (33, 47, 64, 40, 1, 4), (100, 37, 48, 48, 0, 5),
    (32, 32, 32, 32, 0, 6), (31, 33, 32, 32, 0, 7), (97, 61, 160, 160, 0, 8), (7, 5, 64, 64, 1, 9), (65, 64, 64, 64, 0, 10),
    (1920, 1080, 640, 640, 0, 11), (1280, 720, 640, 640, 1, 12), (500, 375, 640, 640, 0, 13), (640, 480, 160, 160, 0, 14),
    (333, 999, 320, 320, 0, 15),
]


# RGBA frames of examples/yolo_detect.cpp:72-130 (always 640 x 640): (frame w, frame h, seed)
RGBA_CASES = [(64, 48, 21), (800, 600, 22), (640, 640, 23), (1280, 720, 24), (333, 999, 25)]


def frame(w, h, seed):
    """smooth ramps + noise, so that both interpolation and clamping matter"""
    rng = np.random.default_rng(seed)
    y, x = np.mgrid[0:h, 0:w]
    base = np.stack([(x * 255) // max(w - 1, 1), (y * 255) // max(h - 1, 1), ((x + y) * 255) // max(w + h - 2, 1)], -1)
    noise = rng.integers(-40, 41, size=(h, w, 3))
    f = np.clip(base + noise, 0, 255).astype(np.uint8)
    f[rng.integers(0, h, 8), rng.integers(0, w, 8)] = rng.choice(np.array([0, 255], dtype=np.uint8), (8, 3))
    return f


if __name__ == "__main__":
    out = {}
    for (w, h, tw, th, nhwc, seed) in CASES:
        t = rb.ref_load_image(frame(w, h, seed), tw, th, bool(nhwc))
        out["%dx%d_to_%dx%d_%s_s%d" % (w, h, tw, th, "nhwc" if nhwc else "nchw", seed)] = {
            "case": [w, h, tw, th, nhwc, seed], "sha256": hashlib.sha256(t.tobytes()).hexdigest(),
            "border": int((t == -17).sum())}
    for (w, h, seed) in RGBA_CASES:
        t = rb.ref_preprocess_rgba(frame(w, h, seed))
        out["rgba_%dx%d_s%d" % (w, h, seed)] = {"rgba": [w, h, seed], "sha256": hashlib.sha256(t.tobytes()).hexdigest(),
                                                "border": int((t.reshape(-1, 4)[:, 3] == 114).sum())}
    json.dump(out, open(os.path.join(HERE, "preproc.json"), "w"), indent=1, sort_keys=True)
    print("wrote %d cases" % len(out))
