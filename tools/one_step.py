"""tools/one_step.py -- load the headline model, run `steps` resident steps of `batch` images (ncu target)."""
import sys

import numpy as np

sys.path.insert(0, ".")
from __graft_entry__ import load_package

B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
pkg = load_package()
mf = pkg.marsfile
gm = pkg.MarsModel(mf.build_yolov5(width=0.5, size=640, seed=5).to_bytes(), arena_bytes=mf.ARENA_YOLOV5S_INT8, batch=B)
x = np.stack([np.random.default_rng(1000 + i).integers(-128, 128, size=3 * 640 * 640, dtype=np.int8) for i in range(min(B, 4))])
for i in range(B):
    gm.upload_inputs(i, 1, x[i % len(x)], x.shape[1])
ms = [gm.step_resident(0, B, 0.45, True) for _ in range(steps)]
print("batch", B, "ms/step", ms, "launches", gm.launch_count)
