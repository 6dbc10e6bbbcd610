"""CPU tests of the drop-in boundary: the C-ABI library loads, exports every symbol the
header declares, repeats the reference's argument checks, and has no CPU fallback."""
import ctypes
import os
import re
import subprocess

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "ga_b200.h")


def header_symbols():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(ga_[a-z0-9_]+)\s*\(", src)))


def test_header_declares_the_expected_entry_points():
    syms = header_symbols()
    for must in ["ga_nn_distance_fwd", "ga_nn_distance_bwd", "ga_knn", "ga_selection_sort", "ga_group_point",
                 "ga_knn_dists", "ga_chamfer_all_pairs", "ga_nn_distance_fwd_host", "ga_last_error"]:
        assert must in syms


def test_library_exports_every_header_symbol(ga):
    from geometric_adv_b200 import _lib
    lib = ctypes.CDLL(_lib.LIB_PATH)
    for s in header_symbols():
        assert hasattr(lib, s), "libga_b200.so does not export %s" % s
        assert s in _lib.SIGNATURES, "python binding does not list %s" % s
    assert lib.ga_version() >= 100


def test_library_is_built_for_sm_100a_only(ga):
    from geometric_adv_b200 import _lib
    out = subprocess.run(["cuobjdump", "-lelf", _lib.LIB_PATH], capture_output=True, text=True)
    if out.returncode != 0:
        pytest.skip("cuobjdump not available")
    archs = set(re.findall(r"sm_(\d+a?)", out.stdout))
    assert archs == {"100a"}, archs


def test_argument_checks_repeat_the_reference_messages(ga, oracle):
    """tf_nndistance.cpp:51-58 -- same conditions, same text."""
    cases = [((2, 5, 2), (2, 5, 3), "NnDistance only accepts 3d point set xyz1"),
             ((2, 5), (2, 5, 3), r"NnDistance requires xyz1 be of shape \(batch,#points,3\)"),
             ((2, 5, 3), (2, 5, 3, 1), r"NnDistance requires xyz2 be of shape \(batch,#points,3\)"),
             ((2, 5, 3), (2, 5, 4), "NnDistance only accepts 3d point set xyz2"),
             ((2, 5, 3), (3, 5, 3), "NnDistance expects xyz1 and xyz2 have same batch size")]
    for s1, s2, msg in cases:
        with pytest.raises(ValueError, match=msg):
            ga.nn_distance(torch.zeros(s1), torch.zeros(s2))
        if oracle.have_ref() and len(s1) >= 3 and len(s2) >= 3:
            with pytest.raises(ValueError, match=msg):
                oracle.ref_nn_distance_shaped(np.zeros(s1, np.float32), np.zeros(s2, np.float32))
    with pytest.raises(ValueError, match=r"NnDistanceGrad requires grad_dist1 be of shape\(batch,#points\)"):
        ga.nn_distance_grad(torch.zeros(2, 5, 3), torch.zeros(2, 6, 3), torch.zeros(2, 6),
                            torch.zeros(2, 5, dtype=torch.int32), torch.zeros(2, 6),
                            torch.zeros(2, 6, dtype=torch.int32))
    with pytest.raises(ValueError, match="SelectionSort expects positive k"):
        ga.select_top_k(0, torch.zeros(1, 2, 3))
    with pytest.raises(ValueError, match=r"SelectionSort expects \(b,m,n\) dist shape"):
        ga.select_top_k(2, torch.zeros(2, 3))
    with pytest.raises(ValueError, match="GroupPoint expects"):
        ga.group_point(torch.zeros(2, 3), torch.zeros(2, 3, 4, dtype=torch.int32))
    with pytest.raises(TypeError):
        ga.nn_distance(torch.zeros(1, 4, 3, dtype=torch.float64), torch.zeros(1, 4, 3, dtype=torch.float64))


def test_product_never_touches_the_oracle():
    """The oracle is test infrastructure; a product path through it would void parity."""
    pkg = os.path.join(ROOT, "geometric_adv_b200")
    for d, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")) or f == "Makefile":
                txt = open(os.path.join(d, f), errors="ignore").read()
                assert "oracle" not in txt.lower(), "%s mentions the oracle" % os.path.join(d, f)
                assert "/root/reference" not in txt, "%s reads the reference tree" % os.path.join(d, f)


@pytest.mark.skipif(torch.cuda.is_available(), reason="this checks the no-GPU failure mode")
def test_no_silent_cpu_fallback(ga):
    """Without a device the ops must fail loudly, never compute on the CPU."""
    from geometric_adv_b200._lib import GaError
    x = torch.rand(1, 8, 3)
    with pytest.raises((GaError, RuntimeError)):
        ga.nn_distance(x, x)
    with pytest.raises((GaError, RuntimeError)):
        ga.knn_point(2, x, x)


def test_missing_library_fails_loudly(monkeypatch):
    from geometric_adv_b200 import _lib
    monkeypatch.setattr(_lib, "_lib", None)
    monkeypatch.setattr(_lib, "LIB_PATH", "/nonexistent/libga_b200.so")
    with pytest.raises(_lib.GaError, match="no CPU fallback"):
        _lib.load()


def test_ops_are_registered_with_torch_library():
    """The shim is a set of PyTorch custom ops (north_star: "PyTorch custom-op shim"): registered schemas and fake
    kernels that propagate shapes / dtypes without touching a device."""
    import torch
    import geometric_adv_b200  # noqa: F401
    from torch._subclasses.fake_tensor import FakeTensorMode
    ns = torch.ops.geometric_adv_b200
    for name in ("nn_distance", "nn_distance_grad", "knn_point", "knn_dists", "group_point"):
        assert hasattr(ns, name), name
    schema = str(ns.nn_distance.default._schema).replace("SymInt", "int")  # torch prints int arguments as SymInt
    assert "Tensor xyz1, Tensor xyz2, int mode" in schema
    with FakeTensorMode():
        a, b = torch.empty(3, 100, 3), torch.empty(3, 200, 3)
        d1, i1, d2, i2 = ns.nn_distance(a, b, 0)
        assert d1.shape == (3, 100) and i1.dtype == torch.int32 and d2.shape == (3, 200) and i2.shape == (3, 200)
        g1, g2 = ns.nn_distance_grad(a, b, d1, i1, d2, i2)
        assert g1.shape == a.shape and g2.shape == b.shape
        val, idx = ns.knn_point(5, a, b)
        assert val.shape == (3, 200, 5) and idx.dtype == torch.int32
        assert ns.knn_dists(a, 10).shape == (3, 100, 10)
        assert ns.group_point(a, idx).shape == (3, 200, 5, 3)


def test_every_tuning_key_is_documented_in_the_header():
    """ga_set_tuning's keys (core.cu) and the table in include/ga_b200.h must not drift apart."""
    import re
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    core = open(os.path.join(root, "geometric_adv_b200", "csrc", "core.cu")).read()
    hdr = open(os.path.join(root, "include", "ga_b200.h")).read()
    keys = sorted(set(int(k) for k in re.findall(r"key == (\d+)", core)))
    assert keys and keys[0] == 0
    table = hdr[hdr.index("Tuning hooks for benchmarks and tests"):hdr.index("int ga_set_tuning(int key, int value);")]
    missing = [k for k in keys if not re.search(r"(^|[\s*(])%d\s+[a-zA-Z]" % k, table)]
    assert not missing, "tuning keys without an entry in the header's table: %s" % missing
