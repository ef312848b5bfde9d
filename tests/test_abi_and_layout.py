"""CPU: the C-ABI library loads and exports every symbol include/diffpure_b200.h declares; the product package
never imports the oracle; creating an engine without a GPU fails loudly (no fallback)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _build():
    import __graft_entry__ as g
    g.build()


def test_library_exports_every_declared_symbol():
    _build()
    from diffpure_b200 import lib
    header = open(os.path.join(ROOT, "include", "diffpure_b200.h")).read()
    declared = set(re.findall(r"\b(dp_[a-z_0-9]+)\s*\(", header))
    declared -= {"dp_engine"}
    l = lib.load()
    for name in sorted(declared):
        assert hasattr(l, name), f"{name} declared in the header but not exported"
    assert declared == set(lib.SYMBOLS), (declared ^ set(lib.SYMBOLS))
    assert l.dp_version() >= 100


def test_create_without_gpu_fails_loudly():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    _build()
    from diffpure_b200 import lib
    l = lib.load()
    h = ctypes.c_void_p()
    rc = l.dp_create(ctypes.byref(h), 0)
    assert rc != 0 and not h.value
    assert b"CUDA" in l.dp_last_error(None) or b"device" in l.dp_last_error(None)


def test_product_never_imports_oracle():
    bad = []
    for dirpath, _, files in os.walk(os.path.join(ROOT, "diffpure_b200")):
        for f in files:
            if f.endswith(".py"):
                src = open(os.path.join(dirpath, f)).read()
                if re.search(r"^\s*(from|import)\s+oracle\b", src, re.M):
                    bad.append(os.path.join(dirpath, f))
    assert not bad, bad


def test_counter_based_normal_host():
    """dp_normal_host: deterministic, keyed by (seed, sample, stream, pixel, channel), ~N(0,1)."""
    _build()
    from diffpure_b200 import lib
    import numpy as np
    l = lib.load()
    v = np.array([l.dp_normal_host(7, s, 3, p, c) for s in range(4) for p in range(512) for c in range(3)])
    assert abs(v.mean()) < 0.05 and abs(v.std() - 1.0) < 0.05
    assert l.dp_normal_host(7, 1, 3, 5, 2) == l.dp_normal_host(7, 1, 3, 5, 2)
    assert l.dp_normal_host(7, 1, 3, 5, 2) != l.dp_normal_host(7, 2, 3, 5, 2)
    assert l.dp_normal_host(7, 1, 3, 5, 2) != l.dp_normal_host(8, 1, 3, 5, 2)
