"""N>1 host logic on CPU: two gloo ranks shard a batch and rank 0 gathers the detection records
(the same DetectionGather object bench.py drives over NCCL).  No GPU, no CUDA library call."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
STRIDE, REC = 1024, 6  # records per image, int32 words per record (mars_det_t = 24 bytes)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _records(image):
    """deterministic fake detection records of one image: count depends on the image index"""
    rng = np.random.default_rng(image)
    n = int(rng.integers(0, 40))
    d = np.zeros((STRIDE, REC), dtype=np.int32)
    d[:n] = rng.integers(-2**31, 2**31 - 1, size=(n, REC), dtype=np.int64).astype(np.int32)
    return d.reshape(-1), n


def _worker(rank, world, port, total, out_path):
    sys.path.insert(0, ROOT)
    from __graft_entry__ import load_package
    shard = load_package().shard
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        first, count = shard.shard_range(total, world, rank)
        per = -(-total // world)
        det = torch.zeros((per, STRIDE * REC), dtype=torch.int32)
        cnt = torch.zeros((per,), dtype=torch.int32)
        for i in range(count):
            d, n = _records(first + i)
            det[i] = torch.from_numpy(d)
            cnt[i] = n
        g = shard.DetectionGather(det, cnt, dist)
        for _ in range(2):  # reusable across steps
            g.run()
        dets, counts = g.result()
        if rank == 0:
            np.savez(out_path, dets=dets.numpy(), counts=counts.numpy())
        else:
            assert dets is None and counts is None
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("total", [8, 7])
def test_two_ranks_shard_and_gather(tmp_path, total):
    world, port = 2, _free_port()
    out = str(tmp_path / "gathered.npz")
    mp.spawn(_worker, args=(world, port, total, out), nprocs=world, join=True)
    z = np.load(out)
    per = -(-total // world)
    assert z["dets"].shape == (world * per, STRIDE * REC) and z["counts"].shape == (world * per,)
    for img in range(total):  # image img sits at row img (contiguous blocks, padded tail)
        d, n = _records(img)
        assert z["counts"][img] == n
        assert np.array_equal(z["dets"][img], d)
    assert not z["counts"][total:].any()


def test_shard_range_covers_the_batch_exactly():
    sys.path.insert(0, ROOT)
    from __graft_entry__ import load_package
    shard = load_package().shard
    for total in (0, 1, 7, 128, 1024, 4096):
        for world in (1, 2, 4, 8):
            seen = []
            for r in range(world):
                f, c = shard.shard_range(total, world, r)
                seen += list(range(f, f + c))
            assert seen == list(range(total))
    assert [shard.stream_owner(s, 8) for s in range(10)] == [0, 1, 2, 3, 4, 5, 6, 7, 0, 1]
    with pytest.raises(ValueError):
        shard.shard_range(8, 2, 2)
