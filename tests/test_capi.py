"""CPU: the C-ABI library loads and exports every symbol include/*.h declares; without a GPU
the product path fails loudly (no CPU fallback)."""
import ctypes
import glob
import os
import re

import pytest

from conftest import ROOT, _has_gpu


def declared_symbols():
    names = set()
    for h in glob.glob(os.path.join(ROOT, "include", "*.h")):
        src = re.sub(r"/\*.*?\*/", "", open(h).read(), flags=re.S)
        for m in re.finditer(r"\b([a-z][a-z0-9_]*)\s*\(", src):
            n = m.group(1)
            if n.startswith(("mars_", "nna_", "mxu_", "conv2d_")) and not n.endswith("_t"):
                names.add(n)
    return sorted(names)


def test_every_declared_symbol_is_exported(pkg):
    L = ctypes.CDLL(pkg.capi.LIB_PATH)
    syms = declared_symbols()
    assert len(syms) > 60
    missing = [s for s in syms if not hasattr(L, s)]
    assert not missing, missing
    # and the binding table covers the headers
    assert not [s for s in syms if s not in pkg.capi.SIGNATURES]


def test_struct_abi(pkg):
    c = pkg.capi
    assert ctypes.sizeof(c.Header) == 76 and ctypes.sizeof(c.TensorDesc) == 124 and ctypes.sizeof(c.LayerDesc) == 112
    assert c.RuntimeTensor.vaddr.offset == 128  # reference include/mars_runtime.h:32-38
    assert pkg.lib().mars_get_error_string(-8) == b"Invalid layer"
    assert pkg.lib().mars_get_error_string(-99) == b"Unknown error"


def test_bad_files_are_rejected_before_touching_the_device(pkg):
    L = pkg.lib()
    m = pkg.capi.PM()
    assert L.mars_load_memory(b"\0" * 80, 80, ctypes.byref(m)) == -1  # invalid magic
    assert L.mars_load_memory(b"MARS", 4, ctypes.byref(m)) == -4
    hdr = bytearray(pkg.marsfile.build_tiny(32).to_bytes()[:76])
    hdr[4] = 9
    assert L.mars_load_memory(bytes(hdr), 76, ctypes.byref(m)) == -2  # version


@pytest.mark.skipif(_has_gpu(), reason="checks the no-GPU failure mode")
def test_no_cpu_fallback(pkg):
    blob = pkg.marsfile.build_tiny(32).to_bytes()
    with pytest.raises(pkg.capi.MarsError) as e:
        pkg.MarsModel(blob)
    assert e.value.code == -5  # MARS_ERR_NNA_INIT_FAILED
    assert pkg.lib().nna_init() != 0
