"""tools/nna_layout_bench.py -- device throughput of the NNA-native layout converters against the HBM roofline.
Buffers resident in HBM (torch tensors), CUDA events on the launching stream, inputs larger than L2 (126 MB).
usage: python tools/nna_layout_bench.py"""
import json
import sys

import torch

sys.path.insert(0, ".")
from __graft_entry__ import load_package

pkg = load_package()
L = pkg.lib()
peak = json.load(open("MEASURED_PEAKS.json")).get("hbm_gbs", 6535.4) if __import__("os").path.exists("MEASURED_PEAKS.json") else 6535.4
st = torch.cuda.current_stream().cuda_stream


def timed(fn, reps=10):
    for _ in range(3):
        fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


for (n, c, h, w) in [(256, 32, 320, 320), (256, 64, 160, 160), (512, 128, 80, 80), (64, 255, 80, 80)]:
    x = torch.randint(0, 256, (n, c, h, w), dtype=torch.uint8, device="cuda")
    nb = L.mars_b200_ndhwc32_size(n, c, h, w)
    nat = torch.empty(nb, dtype=torch.uint8, device="cuda")
    back = torch.empty_like(x)
    t1 = timed(lambda: L.mars_b200_nchw_to_ndhwc32_device(x.data_ptr(), n, c, h, w, nat.data_ptr(), st))
    t2 = timed(lambda: L.mars_b200_ndhwc32_to_nchw_device(nat.data_ptr(), n, c, h, w, back.data_ptr(), st))
    assert torch.equal(back, x)
    by = x.numel() + nb
    print("NDHWC32 %4dx%3dx%3dx%3d: to native %.3f ms %.0f GB/s (%.2f of %.0f) | back %.3f ms %.0f GB/s (%.2f)" % (
        n, c, h, w, t1, by / t1 / 1e6, by / t1 / 1e6 / peak, peak, t2, by / t2 / 1e6, by / t2 / 1e6 / peak))
for (co, ci, kh, kw) in [(512, 256, 3, 3), (255, 512, 1, 1), (1024, 1024, 3, 3)]:
    wt = torch.randint(-128, 128, (co, ci, kh, kw), dtype=torch.int8, device="cuda")
    nb = L.mars_b200_nmhwsoib2_size(co, ci, kh, kw)
    pk = torch.empty(nb, dtype=torch.uint8, device="cuda")
    back = torch.empty_like(wt)
    t1 = timed(lambda: L.mars_b200_pack_weights_nmhwsoib2_device(wt.data_ptr(), co, ci, kh, kw, pk.data_ptr(), st))
    t2 = timed(lambda: L.mars_b200_unpack_weights_nmhwsoib2_device(pk.data_ptr(), co, ci, kh, kw, back.data_ptr(), st))
    assert torch.equal(back, wt)
    by = wt.numel() + nb
    print("NMHWSOIB2 %4dx%4dx%dx%d: pack %.3f ms %.0f GB/s | unpack %.3f ms %.0f GB/s (load-time, %d KB)" % (co, ci, kh, kw, t1, by / t1 / 1e6, t2, by / t2 / 1e6, nb >> 10))
