"""tests/golden/make_golden.py -- generate the committed golden vectors.

Runs the REFERENCE's own portable-C path (oracle/_ref/libmars_ref.so, built from
/root/reference by oracle/build_ref.sh) on the shipped models and records, per case,
sha256 of output tensor 0 and of the whole arena after run 1 and run 2, and the
parse_output + nms detection list of the YOLO heads (as .npy).  The reference tree holds no
golden vectors for this path, so these outputs of the reference itself are the pin.
Usage (in the build container, where /root/reference exists):  python tests/golden/make_golden.py
"""
import hashlib
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from oracle import refbind as rb  # noqa: E402

CASES = [
    # name, model file, arena bytes, input pattern
    ("tiny_160_int8/p0", "tiny_160_int8.mars", 8 << 20, "p0"),
    ("tiny_160_int8/rng", "tiny_160_int8.mars", 8 << 20, "rng1234"),
    ("test_model/p0", "test_model.mars", 8 << 20, "p0"),
    ("test_simple/f32", "test_simple.mars", 8 << 20, "f32"),
    ("yolov5n_int8/p0", "yolov5n_int8.mars", 8 << 20, "p0"),
    ("yolov5n_int8/rng", "yolov5n_int8.mars", 8 << 20, "rng1234"),
    ("yolov5n_int8/p0/large", "yolov5n_int8.mars", 64 << 20, "p0"),
    ("yolov5nu/p0", "yolov5nu.mars", 8 << 20, "p0"),
    ("tiny_160_f32/f32/large", "tiny_160_f32.mars", 64 << 20, "f32"),
    ("yolov5n/f32/large", "yolov5n.mars", 64 << 20, "f32"),
]


def make_input(pattern, desc):
    n = rb.tensor_numel(desc)
    if pattern == "p0":
        return rb.pattern_p0(n)
    if pattern == "f32":
        return rb.pattern_f32(n)
    if pattern.startswith("rng"):
        return np.random.default_rng(int(pattern[3:])).integers(-128, 128, size=n, dtype=np.int8)
    raise ValueError(pattern)


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def main():
    out = {}
    for name, fn, arena, pattern in CASES:
        r = rb.RefRuntime(os.path.join(rb.REF_MODELS, fn), arena_bytes=arena)
        x = make_input(pattern, r.input().desc)
        rec = {"model": fn, "arena": arena, "pattern": pattern}
        for run in (1, 2):
            r.set_input(x)
            r.run()
            m = r.m.contents
            used = min(arena, m.weights_size + 3 * r.input().alloc_size)
            rec["run%d" % run] = {"output_sha256": sha(r.output_bytes()), "arena_sha256": sha(r.arena()[:used]),
                                  "arena_bytes_hashed": int(used)}
        o = r.output().desc
        if o.ndims == 3 and o.shape[2] == 85 and o.dtype == 3:
            ob = r.output_bytes().view(np.int8)
            raw = rb.ref_parse_output(ob, o.shape[1], o.scale)
            kept = rb.ref_nms(raw)
            tag = name.replace("/", "_")
            np.save(os.path.join(HERE, tag + "_dets.npy"), kept)
            rec["dets"] = {"raw": int(len(raw)), "kept": int(len(kept)), "file": tag + "_dets.npy", "sha256": sha(kept)}
        r.close()
        out[name] = rec
        print(name, rec["run1"]["output_sha256"][:16], rec.get("dets", {}).get("kept"))
    # post-process golden on synthetic heads (tie-heavy and random), straight from the reference statics
    rng = np.random.default_rng(77)
    heads = {"post_random": rng.integers(-128, 128, size=(4000, 85), dtype=np.int8),
             "post_ties": rng.choice(np.array([-128, 0, 60, 127], dtype=np.int8), size=(3000, 85))}
    for k, h in heads.items():
        for scale in (0.05, 1.0):
            raw = rb.ref_parse_output(h, h.shape[0], scale)
            kept = rb.ref_nms(raw)
            tag = "%s_s%g" % (k, scale)
            np.save(os.path.join(HERE, tag + "_dets.npy"), kept)
            out[tag] = {"seed": 77, "rows": int(h.shape[0]), "scale": scale, "raw": int(len(raw)), "kept": int(len(kept)),
                        "file": tag + "_dets.npy", "head_sha256": sha(h)}
            print(tag, len(raw), len(kept))
    with open(os.path.join(HERE, "golden.json"), "w") as f:
        json.dump(out, f, indent=1, sort_keys=True)


if __name__ == "__main__":
    main()
