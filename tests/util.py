"""shared helpers for the parity tests"""
import hashlib
import json
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
GOLDEN_DIR = os.path.join(HERE, "golden")
GOLDEN = json.load(open(os.path.join(GOLDEN_DIR, "golden.json")))


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def numel(desc):
    n = 1
    for i in range(min(desc.ndims, 6)):
        n *= max(desc.shape[i], 0)
    return n


def make_input(pattern, n):
    if pattern == "p0":  # reference src/mars/mars_test.c:82-85
        return (np.arange(n, dtype=np.int64) % 127).astype(np.int8)
    if pattern == "f32":  # reference src/mars/mars_test.c:76-80
        return ((np.arange(n, dtype=np.int64) % 256).astype(np.float32) / np.float32(255.0)).astype(np.float32)
    if pattern.startswith("rng"):
        return np.random.default_rng(int(pattern[3:])).integers(-128, 128, size=n, dtype=np.int8)
    raise ValueError(pattern)


def first_diff(a, b):
    a, b = np.asarray(a).ravel(), np.asarray(b).ravel()
    n = min(a.size, b.size)
    d = np.nonzero(a[:n] != b[:n])[0]
    return (int(d[0]), int(d.size)) if d.size else (None, 0)
