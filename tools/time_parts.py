"""tools/time_parts.py -- device time of the layer pass and of decode+NMS separately (tuning aid)."""
import sys

import numpy as np

sys.path.insert(0, ".")
from __graft_entry__ import load_package

B = int(sys.argv[1]) if len(sys.argv) > 1 else 128
pkg = load_package()
mf = pkg.marsfile
gm = pkg.MarsModel(mf.build_yolov5(width=0.5, size=640, seed=5).to_bytes(), arena_bytes=mf.ARENA_YOLOV5S_INT8, batch=B)
x = np.stack([np.random.default_rng(1000 + i).integers(-128, 128, size=3 * 640 * 640, dtype=np.int8) for i in range(min(B, 16))])
for i in range(B):
    gm.upload_inputs(i, 1, x[i % len(x)], x.shape[1])
for _ in range(3):
    gm.step_resident(0, B, 0.45, True)
r = [gm.run_resident(0, B) for _ in range(5)]
d = [gm.detect_resident(0, B, 0.45) for _ in range(5)]
dets, counts = gm.download_detections(0, B)
print("batch %d: layers %.3f ms, decode+nms %.3f ms, kept/img %.1f" % (B, sum(r) / 5, sum(d) / 5, counts.mean()))
