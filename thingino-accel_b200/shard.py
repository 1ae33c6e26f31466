"""shard.py -- image/stream sharding over the GPUs of one box.

Images are independent units: rank r of G processes a contiguous block of the batch and nothing
crosses GPUs on the data path.  The only exchange is the collection of the fixed-size detection
records (`int32 count` + `stride` x 24-byte det_t per image, cap from reference
src/mars/mars_yolo_test.c:187) on rank 0.  Backend-agnostic: NCCL on the GPUs (bench.py), gloo on
CPU tensors in tests/test_sharding_gloo.py.
"""


def shard_range(total, world, rank):
    """contiguous block of ceil(total/world) images for `rank`: (first, count); trailing ranks may get fewer / none"""
    if world < 1 or not (0 <= rank < world):
        raise ValueError("rank %d outside world %d" % (rank, world))
    per = -(-total // world)
    first = min(rank * per, total)
    return first, max(0, min(per, total - first))


def stream_owner(stream_id, world):
    """camera streams are assigned round-robin: stream s lives on GPU s mod G"""
    return stream_id % world


class DetectionGather:
    """gathers every rank's [B, stride*6] int32 detection records and [B] counts on rank `dst`.

    Receive buffers are allocated once; run() issues the two collectives (counts, then records) on
    the tensors given at construction, which may alias device memory owned by libmars_b200.so.
    All ranks must pass equally shaped tensors (pad the last shard).  World size 1: no-op.
    """

    def __init__(self, det_t, cnt_t, dist=None, dst=0, stream=None):
        import torch
        self.det_t, self.cnt_t, self.dist, self.dst = det_t, cnt_t, dist, dst
        # stream: a torch.cuda.Stream the collectives are ordered on (e.g. ExternalStream of the library's compute stream,
        # so that the gather sits between the batch that wrote the records and the next one that overwrites them)
        self.stream = stream
        self.active = dist is not None and dist.is_initialized() and dist.get_world_size() > 1
        self.world = dist.get_world_size() if self.active else 1
        self.rank = dist.get_rank() if self.active else 0
        root = self.active and self.rank == dst
        self.gd = [torch.empty_like(det_t) for _ in range(self.world)] if root else None
        self.gc = [torch.empty_like(cnt_t) for _ in range(self.world)] if root else None

    def run(self):
        if not self.active:
            return
        if self.stream is not None:
            import torch
            with torch.cuda.stream(self.stream):
                self.dist.gather(self.cnt_t, self.gc, dst=self.dst)
                self.dist.gather(self.det_t, self.gd, dst=self.dst)
        else:
            self.dist.gather(self.cnt_t, self.gc, dst=self.dst)
            self.dist.gather(self.det_t, self.gd, dst=self.dst)

    def result(self):
        """(dets [world*B, stride*6], counts [world*B]) on dst after run(); (None, None) elsewhere"""
        import torch
        if not self.active:
            return self.det_t, self.cnt_t
        if self.rank != self.dst:
            return None, None
        return torch.cat(self.gd, 0), torch.cat(self.gc, 0)
