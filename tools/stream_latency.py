"""tools/stream_latency.py -- BASELINE config 5: camera streams through a NanoDet-m-shaped int8 model.

S streams per GPU (4096 streams round-robin over 8 GPUs = 512 per GPU), frames micro-batched <= 64 at a time: every
micro-batch goes host -> device -> all layers -> host through mars_b200_run_batch (the call a stream server would make);
per-frame latency = enqueue of its micro-batch -> its output tensor on the host.  Prints p50 / p99 latency and frames/s.
usage: python tools/stream_latency.py [streams=512] [frames_per_stream=20] [micro_batch=64]"""
import ctypes as C
import sys
import time

import numpy as np

sys.path.insert(0, ".")
from __graft_entry__ import load_package

S = int(sys.argv[1]) if len(sys.argv) > 1 else 512
T = int(sys.argv[2]) if len(sys.argv) > 2 else 20
MB = int(sys.argv[3]) if len(sys.argv) > 3 else 64
pkg = load_package()
mf = pkg.marsfile
L = pkg.capi.lib()
gm = pkg.MarsModel(mf.build_nanodet_like(size=320, seed=7).to_bytes(), arena_bytes=16 << 20, batch=2 * MB)
in_bytes, out_bytes = gm.input_bytes, gm.out_bytes


def pinned(nbytes):
    p = L.nna_malloc(nbytes)
    return np.ctypeslib.as_array((C.c_uint8 * nbytes).from_address(p))


# stream s, frame t: default_rng(s * 1_000_003 + t); a pool of distinct frames is cycled (generating 10k frames dominates otherwise)
POOL = 256
frames = pinned(POOL * in_bytes).reshape(POOL, in_bytes)
for i in range(POOL):
    frames[i] = np.random.default_rng((i % S) * 1_000_003 + i // S).integers(-128, 128, size=in_bytes, dtype=np.int8).view(np.uint8)
stage_in = pinned(MB * in_bytes).reshape(MB, in_bytes)
stage_out = pinned(MB * out_bytes).reshape(MB, out_bytes)

lat = []
for warm in (True, False):
    t_run0 = time.perf_counter()
    done = 0
    for t in range(2 if warm else T):
        for s0 in range(0, S, MB):
            n = min(MB, S - s0)
            for k in range(n):  # the server's gather of the newest frame of each stream of the group
                stage_in[k] = frames[((s0 + k) + t * S) % POOL]
            t0 = time.perf_counter()
            gm.run_batch(n, stage_in, in_bytes, stage_out, out_bytes)
            dt = time.perf_counter() - t0
            if not warm:
                lat.extend([dt] * n)
            done += n
    total = time.perf_counter() - t_run0
lat = np.sort(np.array(lat)) * 1e3
print("nanodet-like 320x320 int8: %d streams x %d frames, micro-batch %d: %.0f frames/s (host staging included), "
      "latency p50 %.3f ms, p99 %.3f ms, max %.3f ms; output %d bytes/frame" % (
          S, T, MB, done / total, lat[len(lat) // 2], lat[int(len(lat) * 0.99)], lat[-1], out_bytes))
