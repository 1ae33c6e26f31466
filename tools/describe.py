"""tools/describe.py -- print the compiled device program of the headline model (planner decisions per op)."""
import sys

sys.path.insert(0, ".")
from __graft_entry__ import load_package

pkg = load_package()
mf = pkg.marsfile
gm = pkg.MarsModel(mf.build_yolov5(width=0.5, size=640, seed=5).to_bytes(), arena_bytes=mf.ARENA_YOLOV5S_INT8, batch=2)
print(gm.describe())
