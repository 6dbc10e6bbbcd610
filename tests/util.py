import hashlib
import json
import os

import numpy as np

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def cloud(seed, shape, lo=-0.5, hi=0.5):
    """Same generator as tests/golden/make_golden.py."""
    rng = np.random.default_rng(seed)
    return (rng.random(shape, dtype=np.float32) * np.float32(hi - lo) + np.float32(lo)).astype(np.float32)


def digests():
    with open(os.path.join(GOLDEN, "digests.json")) as f:
        return json.load(f)


def golden(name):
    return np.load(os.path.join(GOLDEN, name))


def bits_equal(a, b):
    """Bitwise equality of float arrays, NaNs of any payload counted equal."""
    a = np.ascontiguousarray(a)
    b = np.ascontiguousarray(b)
    if a.shape != b.shape:
        return False
    if a.dtype.kind == "f":
        na, nb = np.isnan(a), np.isnan(b)
        if not np.array_equal(na, nb):
            return False
        return np.array_equal(a[~na].view(np.int32), b[~nb].view(np.int32))
    return np.array_equal(a, b)
