"""Drop-in proof: the reference's own caller programs, unmodified.

oracle/build_ref.sh compiles reference src/mars/mars_test.c and src/mars/mars_yolo_test.c twice:
`*.b200` against this repository's include/ + libmars_b200.so (the link line of INTEGRATION.md
section 2) and `*.ref` against the reference's own runtime.  Both are run here and every line that
reports a RESULT (shapes, scales, output values, detection counts and the detection list) must be
identical; lines that print addresses, the summary of the runtime in use or per-layer chatter are
not results and are ignored.
"""
import os
import re
import subprocess

import pytest

from conftest import REF_MODELS, ROOT, shipped

BIN = os.path.join(ROOT, "oracle", "_ref", "bin")
RESULT = re.compile(r"^\s*(Name:|Shape:|Data type:|Scale:|First 16|Stats:|Sample @|Input tensor scale|Input: |Output: \[|"
                    r"Raw detections|Found \d+ detections|No detections|\[\s*\d+\] \S)")


def result_lines(text):
    return [ln.rstrip() for ln in text.splitlines() if RESULT.match(ln)]


def run(binary, *args, env=None):
    p = os.path.join(BIN, binary)
    if not os.path.exists(p):
        pytest.skip(binary + " not built (oracle/build_ref.sh needs /root/reference)")
    e = dict(os.environ)
    e.update(env or {})
    r = subprocess.run([p, *args], capture_output=True, text=True, timeout=600, env=e)
    assert r.returncode == 0, "%s failed (%d): %s" % (binary, r.returncode, r.stderr[-2000:])
    return r.stdout


@pytest.mark.gpu
@pytest.mark.parametrize("model", ["tiny_160_int8.mars", "test_model.mars", "test_simple.mars", "yolov5n_int8.mars", "yolov5nu.mars"])
def test_reference_mars_test_runs_unmodified_on_the_b200_library(model):
    path = shipped(model)
    want = result_lines(run("mars_test.ref", path))
    got = result_lines(run("mars_test.b200", path))
    assert want, "the reference binary printed no result lines"
    assert got == want


@pytest.mark.gpu
def test_reference_mars_yolo_test_detections_are_identical():
    path = shipped("yolov5n_int8.mars")
    want = result_lines(run("mars_yolo_test.ref", path))
    got = result_lines(run("mars_yolo_test.b200", path))
    assert any(ln.startswith("Found") or ln.startswith("No detections") for ln in want)
    assert got == want


def test_dropin_binaries_link_only_the_b200_library():
    """CPU check: the .b200 callers resolve every mars_/nna_ symbol from libmars_b200.so alone"""
    p = os.path.join(BIN, "mars_test.b200")
    if not os.path.exists(p):
        pytest.skip("drop-in binaries not built")
    out = subprocess.run(["ldd", p], capture_output=True, text=True).stdout
    assert "libmars_b200.so" in out and "libmars_ref" not in out and "not found" not in out
    undefined = subprocess.run(["nm", "-D", "--undefined-only", p], capture_output=True, text=True).stdout
    need = {ln.split()[-1].split("@")[0] for ln in undefined.splitlines() if " U " in ln}
    need = {s for s in need if s.startswith(("mars_", "nna_"))}
    assert {"mars_load_file", "mars_run", "mars_get_input", "mars_get_output", "mars_free", "nna_init", "nna_deinit"} <= need
    have = subprocess.run(["nm", "-D", "--defined-only", os.path.join(ROOT, "thingino-accel_b200", "lib", "libmars_b200.so")],
                          capture_output=True, text=True).stdout
    have = {ln.split()[-1] for ln in have.splitlines()}
    assert need <= have
    assert os.path.isdir(REF_MODELS)
