#!/usr/bin/env python
"""bench.py -- YOLOv5s-int8 640x640 images/s on N B200 (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W            # this framework (CUDA)
    python bench.py --impl reference --steps K --warmup W    # the reference's CPU path

A step = one pass of the hot path (all .mars layers, YOLO decode, class-wise NMS) over the global
batch of BASELINE configs[2] -- 1024 synthetic 640x640 images, sharded 1024 / N per GPU (strong scaling).  The model is the yolov5s-shaped graph written by
thingino-accel_b200/marsfile.py (the reference's yolov5s_int8.mars is a missing blob; same layer
table as the shipped yolov5n_int8.mars at 2x width, SYNTHETIC weights, seed 5).

  value : whole-job images/s with the batch already resident in HBM (device time, CUDA events
          on the library's stream, max over ranks);
  e2e   : the same through mars_b200_submit_batch() / mars_b200_wait_batch() with HOST (pinned) buffers -- H2D of
          every image and D2H of every detection list inside the timed region, up to 512 images per submit so that
          the copies of one submit overlap the kernels of the previous one;
  roofline / cpu_baseline : see DESIGN.md section 6.

Images shard across ranks (one process per GPU, no data-path collective); rank 0 gathers the
fixed-size detection records over NCCL at the end of every step.
"""
import argparse
import ctypes as C
import json
import multiprocessing as mp
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
from __graft_entry__ import load_package  # noqa: E402

UNIT = "images/s"
NMS_THRESH = 0.45
MISSING = "the reference's %s is a missing blob (models/.MISSING_LARGE_BLOBS)"

# --config: the headline (BASELINE configs[2]) is the default; the others are the remaining GPU configurations of BASELINE.json
CONFIGS = {
    "int8": dict(metric="yolov5s_int8_640_images_per_s", global_batch=1024, detect=True, f32=False, nhwc=False, seed=5, dtype="int8",
                 workload="BASELINE configs[2]: yolov5s_int8.mars-shaped 640x640 int8, all layers + decode + NMS",
                 model_file="synthetic yolov5s-shaped .mars (2x-width copy of the shipped yolov5n_int8 layer table, writer seed 5); " + MISSING % "yolov5s_int8.mars"),
    "nhwc": dict(metric="yolov5s_int8_nhwc_640_images_per_s", global_batch=1024, detect=True, f32=False, nhwc=True, seed=5, dtype="int8",
                 workload="the configs[2] graph in the compiler's --nhwc convention (channel-innermost activations, OHWI weights: conv2d_int8_nhwc_mxu)",
                 model_file="synthetic yolov5s-shaped .mars, --nhwc convention (writer seed 5)"),
    "f32": dict(metric="yolov5s_float32_640_images_per_s", global_batch=256, detect=False, f32=True, nhwc=False, seed=6, dtype="tf32x3",
                workload="BASELINE configs[3]: yolov5s_float32.mars-shaped 640x640 float32, all layers (tcgen05 kind::tf32, hi/lo operand split)",
                model_file="synthetic yolov5s-shaped float32 .mars (writer seed 6); " + MISSING % "yolov5s_float32.mars"),
}
CFG = CONFIGS["int8"]


def build_model_blob(pkg):
    mf = pkg.marsfile
    blob = mf.build_yolov5(width=0.5, size=640, seed=CFG["seed"], f32=CFG["f32"], nhwc=CFG["nhwc"]).to_bytes()
    return blob, (mf.ARENA_YOLOV5S_F32 if CFG["f32"] else mf.ARENA_YOLOV5S_INT8)


def synth_image(index):
    """image b: default_rng(1000+b) -- int8 in [-128, 127], float32 uniform [0, 1) for the float32 model"""
    rng = np.random.default_rng(1000 + index)
    if CFG["f32"]:
        return rng.random(3 * 640 * 640, dtype=np.float32).view(np.int8)
    return rng.integers(-128, 128, size=3 * 640 * 640, dtype=np.int8)


def synth_images(first, n):
    return np.stack([synth_image(first + i) for i in range(n)])


# ---------------------------------------------------------------------------------------
# clocks
# ---------------------------------------------------------------------------------------
class ClockSampler(threading.Thread):
    """samples SM clock and throttle reasons of one GPU while the timed region runs.  NVML is initialised in the
    constructor (it can take longer than a short timed region), and a sample is also taken synchronously at start and stop."""

    NAMES = {0x2: "applications_clocks_setting", 0x4: "sw_power_cap", 0x8: "hw_slowdown", 0x10: "sync_boost",
             0x20: "sw_thermal_slowdown", 0x40: "hw_thermal_slowdown", 0x80: "hw_power_brake_slowdown",
             0x100: "display_clock_setting"}

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.stop_flag, self.samples, self.reasons, self.max_mhz = index, False, [], set(), None
        self.nv = self.h = self.get_reasons = None
        try:
            import pynvml as nv
            nv.nvmlInit()
            self.nv, self.h = nv, nv.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = nv.nvmlDeviceGetMaxClockInfo(self.h, nv.NVML_CLOCK_SM)
            self.get_reasons = getattr(nv, "nvmlDeviceGetCurrentClocksEventReasons", None) or nv.nvmlDeviceGetCurrentClocksThrottleReasons
        except Exception as e:  # nvml missing: record that, do not fail the bench
            self.reasons.add("nvml_unavailable:%s" % type(e).__name__)

    def sample(self):
        if self.nv is None:
            return
        try:
            self.samples.append(self.nv.nvmlDeviceGetClockInfo(self.h, self.nv.NVML_CLOCK_SM))
            r = self.get_reasons(self.h)
            for bit, nm in self.NAMES.items():
                if r & bit:
                    self.reasons.add(nm)
        except Exception as e:  # noqa: BLE001
            self.reasons.add("nvml_error:%s" % type(e).__name__)

    def run(self):
        while not self.stop_flag:
            self.sample()
            time.sleep(0.002)

    def result(self):
        self.sample()  # the GPU is still busy or just finished: one more sample under (near) load
        self.stop_flag = True
        self.join(timeout=2)
        s = sorted(self.samples)
        return {"sm_mhz": s[len(s) // 2] if s else None, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(s)}


# ---------------------------------------------------------------------------------------
# CPU arms (the ONLY place bench.py executes anything under oracle/)
# ---------------------------------------------------------------------------------------
def _cpu_worker(args):
    """one host process: `images` images through the reference's C path (or the restatement); returns the seconds and, for the
    first image, what the GPU result is checked against (detection records, or a digest of the float32 output tensor)"""
    global CFG
    kind, blob, arena, first, images, cfg_name = args
    CFG = CONFIGS[cfg_name]
    os.environ["OMP_NUM_THREADS"] = "1"
    keep = None
    if kind == "reference":
        from oracle import refbind as rb
        r = rb.RefRuntime(blob, arena_bytes=arena)
        t0 = time.perf_counter()
        for i in range(images):
            r.set_input(synth_image(first + i))
            r.run()
            if CFG["detect"]:
                o = r.output_bytes().view(np.int8)
                d = rb.ref_nms(rb.ref_parse_output(o, 25200, r.output().desc.scale))
            else:
                d = r.output_bytes().copy()
            if i == 0:
                keep = np.asarray(d).tobytes()
        return time.perf_counter() - t0, keep
    from oracle import oraclebind as ob
    m = ob.OracleModel(blob, arena_bytes=arena)
    t0 = time.perf_counter()
    for i in range(images):
        m.set_input(synth_image(first + i))
        m.run()
        if CFG["detect"]:
            o = m.output_bytes().view(np.int8)
            d = ob.nms(ob.parse_output(o, 25200, m.tensor_desc(m.output_index()).scale))
        else:
            d = m.output_bytes().copy()
        if i == 0:
            keep = np.asarray(d).tobytes()
    return time.perf_counter() - t0, keep


def cpu_kind():
    return "reference" if os.path.exists(os.path.join(ROOT, "oracle", "_ref", "libmars_ref.so")) else "port"


def cpu_step(pool, kind, blob, arena, cores, images_per_core=1, cfg_name="int8"):
    """one bounded sample: `cores` processes x images_per_core images (process c: images c*ipc ...); returns images/s, seconds and
    the per-process results of image c*ipc (for the parity check of the CUDA arm)"""
    t0 = time.perf_counter()
    res = pool.map(_cpu_worker, [(kind, blob, arena, c * images_per_core, images_per_core, cfg_name) for c in range(cores)])
    dt = time.perf_counter() - t0
    return cores * images_per_core / dt, dt, [r[1] for r in res]


def host_cores():
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    pkg = load_package()  # marsfile only; no CUDA call is made on this arm
    blob, arena = build_model_blob(pkg)
    kind = cpu_kind()
    cores = min(host_cores(), 64)
    ctx = mp.get_context("spawn")
    with ctx.Pool(cores) as pool:
        for _ in range(min(args.warmup, 1)):  # one full warm-up pass is enough on a CPU
            cpu_step(pool, kind, blob, arena, cores, cfg_name=args.config)
        t0 = time.perf_counter()
        for _ in range(args.steps):
            cpu_step(pool, kind, blob, arena, cores, cfg_name=args.config)
        dt = time.perf_counter() - t0
    value = args.steps * cores / dt
    sample = "%d image(s) per step on each of %d processes (one per host core), %s" % (1, cores, CFG["workload"])
    line = {"impl": "reference", "metric": CFG["metric"], "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "float32" if CFG["f32"] else "int8", "data": "synthetic",
            "config": {"workload": CFG["workload"], "model_file": CFG["model_file"], "images_per_step": cores, "image": "3x640x640 " + ("float32" if CFG["f32"] else "int8"),
                       "arena_bytes": arena, "parallelism": "%d independent host processes, one image each per step" % cores},
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": kind, "sample": sample},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    emit_line(json.dumps(line))
    return 0


# ---------------------------------------------------------------------------------------
# roofline bookkeeping
# ---------------------------------------------------------------------------------------
KIND_GROUP = {1: "conv", 2: "conv", 3: "conv", 4: "conv"}


def roofline_from_profile(prof, batch, peaks, peak_note):
    """group the per-op CUDA-event times of the timed region; the dominant group is reported"""
    groups = {}
    for p in prof:
        if not p["calls"]:
            continue
        if p["kind"] in KIND_GROUP:
            name = {1: "conv_tcgen05_i8", 2: "conv_tcgen05_tf32"}.get(p["impl"], "conv_direct_i8" if p["kind"] != 3 else "conv_direct_f32")
            work = 2.0 * p["oc"] * p["oh"] * p["ow"] * p["ic"] * p["kh"] * p["kw"] * batch  # int8 ops per launch
            byts = (p["ic"] * p["ih"] * p["iw"] + p["oc"] * p["oh"] * p["ow"]) * batch + p["oc"] * p["ic"] * p["kh"] * p["kw"]
        else:
            name = "memory_bound_layers"
            n = p["n"] if p["n"] else p["oh"] * p["ow"] * p["ic"]
            work, byts = 0.0, (3 if p["kind"] in (8, 9) else 2) * n * batch
        g = groups.setdefault(name, {"ms": 0.0, "calls": 0, "ops": 0.0, "bytes": 0.0, "launches": 0})
        g["ms"] += p["ms"]
        g["calls"] += p["calls"]
        g["ops"] += work * p["calls"]
        g["bytes"] += byts * p["calls"]
    if not groups:
        return None, {}
    total_ms = sum(g["ms"] for g in groups.values())
    name, g = max(groups.items(), key=lambda kv: kv[1]["ms"])
    shares = {k: round(v["ms"] / total_ms, 4) for k, v in groups.items()}
    if name.startswith("conv"):
        achieved = g["ops"] / (g["ms"] * 1e-3) / 1e12
        f32 = name.endswith("f32")
        peak = (0.5 if f32 else 2.0) * peaks["bf16_tflops_sustained"]
        r = {"bound": "tensor", "achieved": achieved, "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak, "traffic": None,
             "kernel": name, "share_of_step": shares[name], "avg_launch_ms": g["ms"] / g["calls"],
             "peak_source": ("0.5 x bf16_tflops_sustained (tf32:bf16 = 1:2 on tcgen05; achieved counts the algorithmic 2 x MACs, the tf32x3 mode issues three MMAs per product), "
                             if f32 else "2 x bf16_tflops_sustained (int8:bf16 = 2:1 on tcgen05), ") + peak_note}
    else:
        achieved = g["bytes"] / (g["ms"] * 1e-3) / 1e9
        peak = peaks["hbm_gbs"]
        r = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": None,
             "kernel": name, "share_of_step": shares[name], "avg_launch_ms": g["ms"] / g["calls"], "peak_source": peak_note}
    return r, shares


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return {"hbm_gbs": d["hbm_gbs"], "bf16_tflops_sustained": d.get("bf16_tflops_sustained", d["bf16_tflops"])}, "of measured (MEASURED_PEAKS.json)"
    return {"hbm_gbs": 6650.0, "bf16_tflops_sustained": 1400.0}, "of fallback (B200_PROFILING.md)"


# ---------------------------------------------------------------------------------------
# the CUDA arm
# ---------------------------------------------------------------------------------------
def pinned_array(L, nbytes, dtype=np.uint8):
    p = L.nna_malloc(nbytes)  # pinned, device-visible host memory (the reference's allocator seam)
    if not p:
        raise MemoryError("nna_malloc(%d)" % nbytes)
    return np.ctypeslib.as_array((C.c_uint8 * nbytes).from_address(p)).view(dtype), p


def bind_to_gpu_numa_node(index):
    """run this rank on the CPUs NVML reports as local to its GPU, so that the pinned staging buffers it allocates (first
    touch) and the copy engine's reads stay on that socket; best effort"""
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(index)
        ncpu = os.cpu_count() or 1
        words = pynvml.nvmlDeviceGetCpuAffinity(h, (ncpu + 63) // 64)
        cpus = {64 * i + b for i, w in enumerate(words) for b in range(64) if (int(w) >> b) & 1}
        cpus &= set(os.sched_getaffinity(0))
        if cpus:
            os.sched_setaffinity(0, cpus)
    except Exception:  # noqa: BLE001
        pass


I8_CEILING = {32: 1804.6, 64: 3044.7, 128: 4366.8, 256: 4483.3}  # fallback of profiles/r02a_tc_peak_i8.txt (TOP/s by MMA N)


def load_i8_ceiling():
    """the measured tcgen05 kind::i8 ceiling per MMA N (tools/tc_peak.cu -> profiles/r02a_tc_peak_i8.txt): best TOP/s of every N"""
    best = {}
    try:
        for ln in open(os.path.join(ROOT, "profiles", "r02a_tc_peak_i8.txt")):
            f = ln.replace("|", " ").split()
            if len(f) >= 9 and f[0].isdigit() and f[1] in ("1", "2"):
                n, tops = int(f[0]), float(f[7])
                best[n] = max(best.get(n, 0.0), tops)
    except OSError:
        pass
    return best or dict(I8_CEILING)


def shape_ceiling(prof, ceil_by_n):
    """ops-weighted kind::i8 ceiling of this layer mix: every conv runs MMAs of N = min(256, Co padded to 16)"""
    ops = t = 0.0
    keys = sorted(ceil_by_n)
    for p in prof:
        if p["calls"] and p["kind"] in KIND_GROUP and p["impl"] == 1:
            n = min(256, (p["oc"] + 15) // 16 * 16)
            c = ceil_by_n[next((k for k in keys if n <= k), keys[-1])]
            w = 2.0 * p["oc"] * p["oh"] * p["ow"] * p["ic"] * p["kh"] * p["kw"]
            ops += w
            t += w / c
    return ops / t if t else None


def run_cuda_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    import torch
    import torch.distributed as dist
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the product path has no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local_rank)
    os.environ.setdefault("MARS_STRICT", "1")  # a convolution that cannot take the tensor-core kernel is an error here, not a silent 100x slowdown
    if world > 1:
        bind_to_gpu_numa_node(local_rank)  # (not at N = 1: the cpu_baseline leg uses every host core)
        if os.environ.get("NCCL_DEBUG", "").upper() == "VERSION":
            os.environ["NCCL_DEBUG"] = "WARN"  # the version banner goes to stdout, which carries exactly one JSON line
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    pkg = load_package()
    L = pkg.lib()
    blob, arena = build_model_blob(pkg)
    DETECT = CFG["detect"]
    B = args.batch if args.batch > 0 else max(1, CFG["global_batch"] // world)
    gm = pkg.MarsModel(blob, arena_bytes=arena, device=local_rank, batch=B)
    in_bytes, out_bytes = gm.input_bytes, gm.out_bytes
    # synthetic inputs in pinned host memory
    host_in, _ = pinned_array(L, B * in_bytes, np.int8)
    host_in = host_in.reshape(B, in_bytes)
    uniq = min(B, 64)  # 64 distinct seeded images tiled over the batch (generation cost, not a cache trick: 79 MB >> per-image reuse)
    imgs = synth_images(rank * B, uniq)
    for i in range(B):
        host_in[i] = imgs[i % uniq]
    gm.upload_inputs(0, B, host_in, in_bytes)

    # device-side gather of detections (the one collective of the path)
    class _Dev:
        def __init__(self, ptr, shape, typestr):
            self.__cuda_array_interface__ = {"shape": shape, "typestr": typestr, "data": (ptr, False), "version": 3}

    # collectives and the timing events are enqueued on the library's own compute stream: ordered after the kernels that wrote
    # the records and before the next batch that overwrites them, without a host synchronisation in between
    lib_stream = torch.cuda.ExternalStream(gm.compute_stream(), device=torch.device("cuda", local_rank))

    def make_gatherer(first, n):
        """NCCL gather of the detection records of image slots [first, first+n) to rank 0"""
        dptr, cptr, dstride = gm.detections_device()
        det_t = torch.as_tensor(_Dev(dptr + first * dstride * 24, (n, dstride * 6), "<i4"), device="cuda")
        cnt_t = torch.as_tensor(_Dev(cptr + first * 4, (n,), "<i4"), device="cuda")
        return pkg.shard.DetectionGather(det_t, cnt_t, dist if (world > 1 and DETECT) else None, stream=lib_stream)

    gatherer = make_gatherer(0, B)
    gather = gatherer.run

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def maxr(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def resident_step(timed):
        """one step of the hot path on resident inputs.  The work buffers alias each other (the planner reuses the input buffer
        for later tensors), so the seeded images are put back first -- outside the timed region, which is bracketed by CUDA
        events on the launching stream and ends AFTER the detection gather."""
        gm.upload_inputs(0, B, host_in, in_bytes)
        if world > 1:  # the untimed restore takes a different time on every rank: start the timed region together, or rank 0's
            barrier()  # gather would wait for the slowest rank's H2D copy inside its own timed region
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(lib_stream)
        gm.enqueue_step_resident(0, B, NMS_THRESH, DETECT)  # no host round trip between the kernels and the collective
        gather()
        e1.record(lib_stream)
        e1.synchronize()
        return e0.elapsed_time(e1) if timed else 0.0

    # ---- resident: W warm-up + K timed steps --------------------------------------------
    sampler = ClockSampler(local_rank)  # NVML set up before the warm-up, sampling starts with the timed region
    for _ in range(max(args.warmup, 3)):
        resident_step(False)
    barrier()
    sampler.start()
    l0 = gm.launch_count
    dev_ms = 0.0
    t0 = time.perf_counter()
    for _ in range(args.steps):
        dev_ms += resident_step(True)  # CUDA events on the launching stream, gather included
    barrier()
    wall_ms = (time.perf_counter() - t0) * 1e3
    clocks = sampler.result()
    launches = gm.launch_count - l0
    dev_ms = maxr(dev_ms)
    wall_ms = maxr(wall_ms)
    ms_per_step = dev_ms / args.steps
    value = world * B / (ms_per_step * 1e-3)

    # ---- what was timed is what the reference computes ------------------------------------------------------------
    ncheck = min(uniq, host_cores(), 64) if world == 1 else 0
    if DETECT:
        res_dets, res_counts = gm.download_detections(0, B)
        res_checksum = int(res_counts.sum())
        gpu_results = [res_dets[i, :res_counts[i]].tobytes() for i in range(ncheck)]
    else:
        res_checksum = None
        gpu_results = [gm.download_outputs(i, 1)[0].copy() for i in range(ncheck)]
    gather_checked = None
    if world > 1 and DETECT:  # rank 0's gathered block of rank r == rank r's own records
        mine = np.concatenate([res_counts.astype(np.int64), res_dets.view(np.uint8).reshape(B, -1)[:, :64].astype(np.int64).sum(1)])
        own = torch.tensor([int(mine.sum()), int((mine * (np.arange(mine.size) % 251 + 1)).sum())], dtype=torch.int64, device="cuda")
        allh = [torch.zeros_like(own) for _ in range(world)]
        dist.all_gather(allh, own)
        if rank == 0:
            gd, gc = gatherer.result()
            gd, gc = gd.cpu().numpy(), gc.cpu().numpy()
            ok = 0
            for r in range(world):
                blk = np.concatenate([gc[r * B:(r + 1) * B].astype(np.int64),
                                      gd[r * B:(r + 1) * B].view(np.uint8).reshape(B, -1)[:, :64].astype(np.int64).sum(1)])
                h = (int(blk.sum()), int((blk * (np.arange(blk.size) % 251 + 1)).sum()))
                if h != (int(allh[r][0]), int(allh[r][1])):
                    raise SystemExit("bench.py: the gathered detection records of rank %d differ from that rank's own" % r)
                ok += 1
            gather_checked = ok

    # ---- the same K steps again with one CUDA event per device op (roofline bookkeeping; the ~230 extra event
    # records per step cost a few percent, so they stay out of the region `value` is taken from) -------------
    gm.set_profile(True)
    prof_ms = 0.0
    for _ in range(args.steps):
        gm.upload_inputs(0, B, host_in, in_bytes)
        prof_ms += gm.step_resident(0, B, NMS_THRESH, DETECT)
    prof = gm.op_profile()
    gm.set_profile(False)
    prof_ms_per_step = prof_ms / args.steps

    # ---- end to end: host buffers in, detections out ------------------------------------
    # The user-facing call pair mars_b200_submit_batch / mars_b200_wait_batch: every step copies its B images from
    # pinned host memory to HBM and reads its detection records back; batch k+1 is submitted before batch k is
    # waited for, so the copies overlap the kernels (two halves of a 2B-slot pool).  K steps are timed, pipeline
    # fill and drain included.  Two submits per step (measured at 8 GPUs, 128 images per GPU: 132k images/s with one or two
    # submits per step against 101k with four -- smaller launches lose more than the extra overlap gains).
    SUB = min(B, int(os.environ.get("MARS_BENCH_SUB", "512")), max(16, B // int(os.environ.get("MARS_BENCH_SUBMITS", "2"))))
    gm.set_batch(2 * SUB)
    outs = []
    for _ in range(2):
        if DETECT:
            d, _p = pinned_array(L, SUB * 1000 * 24)
            c, _p = pinned_array(L, SUB * 4, np.int32)
        else:
            d, _p = pinned_array(L, SUB * out_bytes)
            c = None
        outs.append((d, c))
    pool_gather = [make_gatherer(0, SUB), make_gatherer(SUB, SUB)]  # the slot pool was re-allocated: new device addresses

    def step_jobs(ramp):
        """(first image, count) submits covering the B images of one step.  The first step of a run ramps its submit size up
        (SUB/8, SUB/8, SUB/4, SUB/2, SUB, ...): the kernels start after a short first copy instead of idling behind a full one"""
        sizes, done = [], 0
        s = max(1, SUB // 8) if ramp else SUB
        first_pair = ramp
        while done < B:
            n = min(s, B - done)
            sizes.append((done, n))
            done += n
            if first_pair:
                first_pair = False  # the smallest size is used twice so that the sizes keep adding up to powers of two
            elif s < SUB:
                s = min(SUB, 2 * s)
        return sizes

    e2e_counts = {}

    def finish(i, job):
        gm.wait_batch(i & 1)
        if DETECT:
            pool_gather[i & 1].run()
            e2e_counts[job[0]] = int(outs[i & 1][1][:job[1]].sum())

    def e2e_steps(k):
        jobs = [j for st in range(k) for j in step_jobs(st == 0)]
        for i, (first, n) in enumerate(jobs):
            if DETECT:
                gm.submit_batch(i & 1, n, host_in[first:first + n], in_bytes, outs[i & 1][0], outs[i & 1][1], 1000, NMS_THRESH)
            else:
                gm.submit_run_batch(i & 1, n, host_in[first:first + n], in_bytes, outs[i & 1][0], out_bytes)
            if i >= 1:
                finish(i - 1, jobs[i - 1])
        finish(len(jobs) - 1, jobs[-1])

    e2e_steps(2)
    barrier()
    t0 = time.perf_counter()
    e2e_steps(args.steps)
    barrier()
    e2e_ms = maxr((time.perf_counter() - t0) * 1e3) / args.steps
    e2e_value = world * B / (e2e_ms * 1e-3)
    checksum = None
    if DETECT:  # (the last step's submits cover every image once: sizes are not ramped after the first step)
        checksum = sum(e2e_counts[f] for f, _n in step_jobs(args.steps == 1))
        if checksum != res_checksum:
            raise SystemExit("bench.py: detections of the end-to-end path (%d) and of the resident path (%d) differ" % (checksum, res_checksum))

    peaks, peak_note = load_peaks()
    roof, shares = roofline_from_profile(prof, B, peaks, peak_note)
    if roof:
        roof["ms_per_step_with_op_events"] = prof_ms_per_step
        direct = shares.get("conv_direct_i8", 0.0) + shares.get("conv_direct_f32", 0.0)
        # the headline graph has ONE convolution off the tensor pipe -- the in-place 1x1 layer 43, whose output planes alias its input planes (order-preserving
        # register kernel, with its SIGMOID + MUL folded in: 1.6 ms of a ~31 ms step = 5 %); a silent fall-back of the whole model to the exact kernels would
        # put this share near 100 % (strict mode already turns that into an error at load)
        if direct > (0.10 if CFG is CONFIGS["int8"] else 1.0):  # (the --nhwc and float32 graphs each have in-place convolutions on order-preserving kernels)
            raise SystemExit("bench.py: %.1f%% of the step runs on the direct (non tensor-core) convolution kernels" % (100 * direct))
        if roof["kernel"] == "conv_tcgen05_i8":
            ceil_by_n = load_i8_ceiling()
            top = max(ceil_by_n.values())
            roof["peak_i8_measured"] = top
            roof["frac_of_i8_measured"] = roof["achieved"] / top
            roof["i8_ceiling_by_mma_n"] = {str(k): v for k, v in sorted(ceil_by_n.items())}
            sc = shape_ceiling(prof, ceil_by_n)
            if sc:
                roof["i8_ceiling_of_this_layer_mix"] = sc
                roof["frac_of_layer_mix_ceiling"] = roof["achieved"] / sc
            roof["peak_i8_source"] = "profiles/r02a_tc_peak_i8.txt: tools/tc_peak.cu, back-to-back tcgen05.mma.kind::i8 M=128 K=32 from resident shared memory, best issue configuration per N"
            # DRAM bytes per launch of the dominant kernel group from the committed ncu capture of THIS build (profiles/), per image x batch
            tpath = os.path.join(ROOT, "profiles", "r03_traffic_conv_tc.json")
            if CFG is CONFIGS["int8"] and os.path.exists(tpath):
                tr = json.load(open(tpath))
                roof["traffic"] = tr["dram_bytes_per_image"] * B / tr["launches_per_step"]
                roof["traffic_source"] = ("profiles/r03_traffic_conv_tc.json: ncu dram__bytes_read.sum + dram__bytes_write.sum of the %d tensor-core conv launches of one "
                                          "step at %d images/GPU, per launch; %s" % (tr["launches_per_step"], tr["batch_per_gpu"],
                                          "measured at this batch" if tr["batch_per_gpu"] == B else "static per-image figure scaled to %d images/GPU" % B))
                roof["hbm_gbs_achieved"] = roof["traffic"] / (roof["avg_launch_ms"] * 1e-3) / 1e9
                roof["hbm_frac"] = roof["hbm_gbs_achieved"] / peaks["hbm_gbs"]

    line = None
    if rank == 0:
        cpu = None
        parity_checked = 0
        if world == 1 and not args.no_cpu_baseline:
            kind = cpu_kind()
            cores = min(host_cores(), 64)
            ctx = mp.get_context("spawn")
            with ctx.Pool(cores) as pool:
                v, dt, ref_results = cpu_step(pool, kind, blob, arena, cores, cfg_name=args.config)
            cpu = {"value": v, "unit": UNIT, "cores": cores, "kind": kind,
                   "sample": "1 image on each of %d host processes (%.1f s), same model and input recipe" % (cores, dt)}
            # the reference's results for images 0 .. cores-1 against what the timed resident steps produced for them
            for i in range(min(ncheck, cores)):
                if DETECT:
                    if gpu_results[i] != ref_results[i]:
                        raise SystemExit("bench.py: image %d: the detection list of the timed CUDA path differs from the reference's" % i)
                else:
                    a, b = np.frombuffer(gpu_results[i].tobytes(), dtype=np.float32), np.frombuffer(ref_results[i], dtype=np.float32)
                    fin = np.isfinite(b)
                    if not np.array_equal(np.isfinite(a), fin):
                        raise SystemExit("bench.py: image %d: NaN / Inf pattern of the float32 output differs from the reference's" % i)
                parity_checked += 1
        line = {"metric": CFG["metric"], "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
                "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": CFG["dtype"],
                "data": "synthetic",
                "config": {"workload": CFG["workload"], "model_file": CFG["model_file"], "images_per_step": B * world,
                           "global_batch": B * world, "per_gpu_batch": B, "image": "3x640x640 " + ("float32" if CFG["f32"] else "int8"), "arena_bytes": arena,
                           "parallelism": "images sharded, dp%d (global batch %d, strong scaling), detections gathered to rank 0 over NCCL" % (world, B * world),
                           "timed_region": "per step: CUDA events on the launching stream around all layers + decode + NMS + the NCCL detection gather; the seeded inputs are restored (H2D, untimed) before every step",
                           "e2e_submit_images": SUB, "e2e_first_step_submits": [n for _f, n in step_jobs(True)],
                           "l2": "working set %.1f GB per GPU >> 126 MB L2 (no flush needed)" % (B * gm.slot_stride / 1e9)},
                "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": B * in_bytes * world,
                        "d2h_bytes_per_step": ((B * 1000 * 24 + B * 4) if DETECT else B * out_bytes) * world, "ms_per_step": e2e_ms},
                "gpu_launches": int(launches), "wall_ms_per_step": wall_ms / args.steps, "clocks": clocks,
                "roofline": roof, "kernel_time_shares": shares, "cpu_baseline": cpu, "detections_checksum": checksum,
                "parity_checked": parity_checked, "gather_checked": gather_checked}
        emit_line(json.dumps(line))
    gm.close()
    if world > 1:
        dist.destroy_process_group()
    return 0


_RESULT_FD = None


def claim_stdout():
    """stdout carries exactly one JSON line: whatever libraries print (NCCL's version banner, torchrun notices) is sent
    to stderr by pointing fd 1 at fd 2; the result line is written to the saved descriptor"""
    global _RESULT_FD
    if _RESULT_FD is None:
        sys.stdout.flush()
        _RESULT_FD = os.dup(1)
        os.dup2(2, 1)


def emit_line(text):
    sys.stdout.flush()
    os.write(_RESULT_FD if _RESULT_FD is not None else 1, (text + "\n").encode())


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="cuda", choices=["cuda", "reference"])
    ap.add_argument("--batch", type=int, default=int(os.environ.get("MARS_BENCH_BATCH", "0")),
                    help="images per GPU per step (default: 1024 / number of GPUs)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--config", default="int8", choices=sorted(CONFIGS) + ["streams"],
                    help="int8 = the headline (BASELINE configs[2], default); nhwc; f32 (configs[3]); streams (configs[4])")
    args = ap.parse_args()
    global CFG
    if args.config == "streams":
        if args.impl == "reference":
            claim_stdout()
            emit_line(json.dumps({"impl": "reference", "unavailable": "the streams configuration has no reference arm (the reference has no NanoDet model or stream runner)"}))
            return 0
        import bench_streams
        return bench_streams.main(args, claim_stdout, emit_line)
    CFG = CONFIGS[args.config]
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if not (args.impl != "reference" and args.gpus > 1 and world == 1):
        claim_stdout()  # (the torchrun re-launch below hands its stdout to the ranks untouched)
    if args.impl == "reference":
        return run_reference_arm(args)
    if args.gpus > 1 and world == 1:
        # convenience: re-launch under torchrun (the driver launches torchrun itself)
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(args.gpus),
               "--master-addr", "127.0.0.1", "--master-port", os.environ.get("MASTER_PORT", "29533"), os.path.abspath(__file__)] + sys.argv[1:]
        return subprocess.call(cmd)
    return run_cuda_arm(args)


if __name__ == "__main__":
    sys.exit(main())
