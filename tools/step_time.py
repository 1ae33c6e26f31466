"""tools/step_time.py -- device time of the graphed step (layers + decode + NMS) at several per-GPU batch sizes (tuning aid).
usage: python tools/step_time.py 64 128 256 ..."""
import sys

import numpy as np

sys.path.insert(0, ".")
from __graft_entry__ import load_package

pkg = load_package()
mf = pkg.marsfile
blob = mf.build_yolov5(width=0.5, size=640, seed=5).to_bytes()
x = np.stack([np.random.default_rng(1000 + i).integers(-128, 128, size=3 * 640 * 640, dtype=np.int8) for i in range(16)])
for B in [int(a) for a in sys.argv[1:]] or [128]:
    gm = pkg.MarsModel(blob, arena_bytes=mf.ARENA_YOLOV5S_INT8, batch=B)
    for i in range(B):
        gm.upload_inputs(i, 1, x[i % len(x)], x.shape[1])
    for _ in range(4):
        gm.step_resident(0, B, 0.45, True)
    t = [gm.step_resident(0, B, 0.45, True) for _ in range(10)]
    print("batch %d: step %.3f ms (min %.3f), %.0f img/s" % (B, sum(t) / len(t), min(t), B * len(t) / sum(t) * 1e3))
    del gm
