"""tools/describe.py -- print the compiled device program of the headline model (planner decisions per op).
Host-only (mars_b200_plan_describe): runs without a GPU.  usage: python tools/describe.py [opt_level]"""
import ctypes as C
import sys

sys.path.insert(0, ".")
from __graft_entry__ import load_package

pkg = load_package()
mf = pkg.marsfile
blob = mf.build_yolov5(width=0.5, size=640, seed=5).to_bytes()
buf = C.create_string_buffer(4 << 20)
n = pkg.lib().mars_b200_plan_describe(blob, len(blob), mf.ARENA_YOLOV5S_INT8, int(sys.argv[1]) if len(sys.argv) > 1 else 3, buf, len(buf))
assert n > 0
print(buf.value.decode())
