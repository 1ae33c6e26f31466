"""tools/time_preproc.py -- throughput of the GPU letterbox pre-processing next to the reference's CPU function.
usage: python tools/time_preproc.py [frames] [w] [h]      (needs oracle/_ref for the CPU leg; skipped if absent)"""
import os
import sys
import time

import numpy as np

sys.path.insert(0, ".")
from __graft_entry__ import load_package

N = int(sys.argv[1]) if len(sys.argv) > 1 else 64
W = int(sys.argv[2]) if len(sys.argv) > 2 else 1920
H = int(sys.argv[3]) if len(sys.argv) > 3 else 1080
pkg = load_package()
mf = pkg.marsfile
gm = pkg.MarsModel(mf.build_yolov5(width=0.5, size=640, seed=5).to_bytes(), arena_bytes=mf.ARENA_YOLOV5S_INT8, batch=N)
rng = np.random.default_rng(0)
frames = rng.integers(0, 256, (N, H, W, 3), dtype=np.uint8)
import ctypes as C
L = pkg.capi.lib()
ptr = L.nna_malloc(frames.nbytes)  # pinned, device-visible host memory (the reference's allocator seam)
pinned = np.ctypeslib.as_array((C.c_uint8 * frames.nbytes).from_address(ptr)).reshape(frames.shape)
pinned[...] = frames
for _ in range(2):
    gm.preprocess(0, pinned)
ms = [gm.preprocess(0, pinned) for _ in range(5)]
t = sum(ms) / len(ms)
inb, outb = W * H * 3, 640 * 640 * 3
print("gpu: %d frames %dx%d -> 640x640: %.3f ms (%.0f frames/s, H2D of %.1f MB/frame inside; in+out %.1f GB/s)" % (
    N, W, H, t, N / t * 1e3, inb / 1e6, N * (inb + outb) / t / 1e6))
try:
    from oracle import refbind
    refbind.RefRuntime.lib()
    t0 = time.perf_counter()
    k = min(N, 4)
    for i in range(k):
        refbind.ref_load_image(frames[i], 640, 640, False)
    dt = (time.perf_counter() - t0) / k
    print("cpu reference load_image (1 core, PPM decode included): %.1f ms/frame (%.1f frames/s)" % (dt * 1e3, 1 / dt))
except Exception as e:  # noqa: BLE001
    print("cpu leg skipped:", e)
