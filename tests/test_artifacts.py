"""The committed bench lines under profiles/ carry every key of the bench contract (guards the artifacts DESIGN.md quotes)."""
import glob
import json
import os

from conftest import ROOT

CONTRACT = {"metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline",
            "dtype", "data", "config", "e2e", "gpu_launches", "clocks"}


def last_json_line(path):
    lines = [ln for ln in open(path).read().splitlines() if ln.startswith("{")]
    assert lines, path
    return json.loads(lines[-1])


def latest_prefix():
    """rNNx of the most recent 1-GPU bench line (names sort by round, then by build letter)"""
    names = sorted(os.path.basename(f) for f in glob.glob(os.path.join(ROOT, "profiles", "r*_bench_n1.json")))
    assert names
    return names[-1].split("_bench_")[0]


def test_committed_bench_lines_follow_the_contract():
    files = sorted(glob.glob(os.path.join(ROOT, "profiles", latest_prefix() + "_bench_n*.json")))
    assert files
    for f in files:
        d = last_json_line(f)
        assert CONTRACT <= set(d), (f, CONTRACT - set(d))
        assert d["metric"] == "yolov5s_int8_640_images_per_s" and d["unit"] == "images/s" and d["dtype"] == "int8"
        assert d["value"] > 0 and d["gpu_launches"] > 0 and "workload" in d["config"]
        assert {"value", "unit", "h2d_bytes_per_step", "d2h_bytes_per_step"} <= set(d["e2e"]) and d["e2e"]["h2d_bytes_per_step"] > 0
        if d["n_gpus"] == 1:
            assert {"bound", "achieved", "peak", "unit", "frac", "traffic"} <= set(d["roofline"])
            assert {"value", "unit", "cores", "kind", "sample"} <= set(d["cpu_baseline"])


def test_committed_reference_line():
    names = sorted(glob.glob(os.path.join(ROOT, "profiles", "r*_bench_reference_n1.json")))  # the most recent reference-arm line
    assert names
    d = last_json_line(names[-1])
    assert d["impl"] == "reference" and d["value"] > 0 and d["cpu_baseline"]["kind"] in ("reference", "port")
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0
