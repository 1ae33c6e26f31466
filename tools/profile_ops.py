"""tools/profile_ops.py -- per-op CUDA-event profile of the headline workload (debug / tuning aid).
usage: python tools/profile_ops.py [batch] [steps]"""
import sys

import numpy as np

sys.path.insert(0, ".")
from __graft_entry__ import load_package

KIND = ["nop", "conv_i8_nchw", "conv_i8_nhwc", "conv_f32", "dw_i8", "byte_relu", "sigmoid_i8", "sigmoid_f32", "mul_i8", "add_i8",
        "mul_f32", "add_f32", "relu_i8", "relu_f32", "bn_i8", "bn_f32", "maxpool", "concat", "concat_periodic", "upsample", "lut_i8"]
B = int(sys.argv[1]) if len(sys.argv) > 1 else 128
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
pkg = load_package()
mf = pkg.marsfile
import os
if os.environ.get("PROFILE_MODEL") == "nanodet":  # BASELINE config 5 shape
    blob, arena, side = mf.build_nanodet_like(size=320, seed=7).to_bytes(), 16 << 20, 320
else:
    blob, arena, side = mf.build_yolov5(width=0.5, size=640, seed=5, nhwc=os.environ.get("PROFILE_MODEL") == "nhwc").to_bytes(), mf.ARENA_YOLOV5S_INT8, 640
gm = pkg.MarsModel(blob, arena_bytes=arena, batch=B)
x = np.random.default_rng(1000).integers(-128, 128, size=(1, 3 * side * side), dtype=np.int8)
for i in range(B):
    gm.upload_inputs(i, 1, x, x.shape[1])
DETECT = os.environ.get("PROFILE_MODEL") != "nanodet"
for _ in range(2):
    gm.step_resident(0, B, 0.45, DETECT)
gm.set_profile(True)
tot = 0.0
for _ in range(steps):
    tot += gm.step_resident(0, B, 0.45, DETECT)
prof = gm.op_profile()
print("batch %d: %.2f ms/step, %.1f img/s" % (B, tot / steps, B * steps / tot * 1e3))
rows = [p for p in prof if p["calls"]]
rows.sort(key=lambda p: -p["ms"])
acc = sum(p["ms"] for p in rows)
print("sum of op intervals %.2f ms/step" % (acc / steps))
for p in rows[:int(sys.argv[3]) if len(sys.argv) > 3 else 45]:
    macs = p["oc"] * p["oh"] * p["ow"] * p["ic"] * p["kh"] * p["kw"] * B if p["kind"] in (1, 2, 3, 4) else 0
    ms = p["ms"] / p["calls"]
    print("op %3d L%3d %-14s impl=%d mode=%d ic=%4d oc=%4d o=%3dx%3d k=%d fused=%d  %8.3f ms  %5.1f%%  %s" % (
        p["op"], p["layer"], KIND[p["kind"]], p["impl"], p["mode"], p["ic"], p["oc"], p["oh"], p["ow"], p["kh"], p["fused"], ms,
        100 * p["ms"] / acc, ("%.1f TOPS" % (2 * macs / ms / 1e9)) if macs else ("%.0f GB/s" % ((p["n"] or p["oh"] * p["ow"] * p["ic"]) * B * 2 / ms / 1e6))))
by = {}
for p in rows:
    k = KIND[p["kind"]] + ("/tc" if p["impl"] == 1 else "")
    by[k] = by.get(k, 0) + p["ms"] / steps
print({k: round(v, 2) for k, v in sorted(by.items(), key=lambda kv: -kv[1])})
