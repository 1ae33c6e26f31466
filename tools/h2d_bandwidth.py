"""tools/h2d_bandwidth.py -- host -> device copy bandwidth of every rank at once (what bounds the end-to-end path at N GPUs).

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tools/h2d_bandwidth.py

Every rank copies a pinned 157 MB buffer (the int8 images of its 128-image shard of BASELINE configs[2]) to its GPU 20 times,
all ranks together; prints per-rank and aggregate GB/s, and the same for the device -> host direction."""
import os
import time

import torch
import torch.distributed as dist

rank, local, world = int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
n = 128 * 3 * 640 * 640
h = torch.empty(n, dtype=torch.int8).pin_memory()
d = torch.empty(n, dtype=torch.int8, device="cuda")
res = {}
for name, (src, dst) in {"h2d": (h, d), "d2h": (d, h)}.items():
    for _ in range(3):
        dst.copy_(src, non_blocking=True)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    for _ in range(20):
        dst.copy_(src, non_blocking=True)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    res[name] = 20 * n / dt / 1e9
t = torch.tensor([res["h2d"], res["d2h"]], dtype=torch.float64, device="cuda")
allr = [torch.zeros_like(t) for _ in range(world)]
if world > 1:
    dist.all_gather(allr, t)
else:
    allr = [t]
if rank == 0:
    h2d = [float(x[0]) for x in allr]
    d2h = [float(x[1]) for x in allr]
    print("ranks %d: H2D per rank GB/s %s -> aggregate %.1f GB/s; D2H per rank %s -> aggregate %.1f GB/s" % (
        world, ["%.1f" % v for v in h2d], sum(h2d), ["%.1f" % v for v in d2h], sum(d2h)))
    print("BASELINE configs[2] moves 1.258 GB of pre-processed int8 images per 1024-image step: at the aggregate H2D rate that is %.2f ms per step = %.0f images/s"
          % (1258.3 / sum(h2d), 1024 / (1.2583 / sum(h2d))))
if world > 1:
    dist.destroy_process_group()
