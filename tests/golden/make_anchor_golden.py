"""tests/golden/make_anchor_golden.py -- golden vectors of the 3-head anchor-grid decode.

The reference has no C implementation of this decode (examples/yolo_detect.cpp:184-205 is a stub); its Python tool
mgk-decompiler/test_yolo_inference.py:136-202 (parse_yolo_output) is the one place the formula is written down.  This script
imports THAT function from /root/reference (onnxruntime, which the file imports for other purposes, is stubbed; nms is
replaced by the identity so that every decoded candidate comes back, sorted by confidence as the function sorts them) and
records its output for seeded int8 heads de-quantised with the head's scale.  Usage (build container only):
    python tests/golden/make_anchor_golden.py
"""
import json
import os
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get("REF_DIR", "/root/reference")
SCALE = 0.05
GRIDS = (20, 10, 5)  # level 0 / 1 / 2 (strides 8 / 16 / 32 are tied to the level index, not to the grid size)
SEEDS = (31, 32, 33)


def heads(seed):
    """int8 [3, g, g, 85] per level: objectness mostly far below the threshold, a few dozen confident cells"""
    rng = np.random.default_rng(seed)
    out = []
    for g in GRIDS:
        h = rng.integers(-128, 128, size=(3, g, g, 85), dtype=np.int8)
        h[..., 4] = rng.integers(-128, -40, size=(3, g, g), dtype=np.int8)
        pick = rng.random((3, g, g)) < (12.0 / (3 * g * g))
        h[..., 4][pick] = rng.integers(20, 128, size=int(pick.sum()), dtype=np.int8)
        out.append(h)
    return out


def load_reference():
    for name in ("onnxruntime",):
        if name not in sys.modules:
            sys.modules[name] = types.ModuleType(name)
    sys.path.insert(0, os.path.join(REF, "mgk-decompiler"))
    import test_yolo_inference as ref  # noqa: E402
    ref.nms = lambda dets, iou_threshold=0.45: dets
    return ref


if __name__ == "__main__":
    ref = load_reference()
    cases = {}
    for seed in SEEDS:
        hs = heads(seed)
        outputs = [(h.astype(np.float32) * np.float32(SCALE))[None] for h in hs]
        dets = ref.parse_yolo_output(outputs, conf_threshold=0.25)
        assert 0 < len(dets) < 100, len(dets)  # the function cuts its result at 100
        conf = np.array([d["confidence"] for d in dets], dtype=np.float64)
        assert np.all(np.abs(conf - 0.25) > 1e-4)  # no candidate sits on the threshold
        cases[str(seed)] = {"boxes": [[float(v) for v in d["box"]] for d in dets], "confidence": conf.tolist(),
                            "class_id": [int(d["class_id"]) for d in dets]}
    json.dump({"scale": SCALE, "grids": list(GRIDS), "cases": cases}, open(os.path.join(HERE, "anchor_decode.json"), "w"))
    print("wrote", {k: len(v["class_id"]) for k, v in cases.items()})
