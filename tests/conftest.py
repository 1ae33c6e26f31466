import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from __graft_entry__ import load_package  # noqa: E402

REF_MODELS = os.path.join(ROOT, "oracle", "_ref", "models")
REF_SO = os.path.join(ROOT, "oracle", "_ref", "libmars_ref.so")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def _has_gpu():
    try:
        return load_package().lib().mars_b200_device_count() > 0
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    if _has_gpu():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for it in items:
        if "gpu" in it.keywords:
            it.add_marker(skip)


@pytest.fixture(scope="session")
def pkg():
    return load_package()


@pytest.fixture(scope="session")
def ob():
    from oracle import oraclebind
    oraclebind.lib()
    return oraclebind


@pytest.fixture(scope="session")
def rb():
    from oracle import refbind
    if not os.path.exists(REF_SO):
        pytest.skip("oracle/_ref/libmars_ref.so not built (needs /root/reference)")
    return refbind


def shipped(name):
    p = os.path.join(REF_MODELS, name)
    if not os.path.exists(p):
        pytest.skip(name + " not staged under oracle/_ref/models")
    return p


def p0(n):
    return (np.arange(n, dtype=np.int64) % 127).astype(np.int8)
