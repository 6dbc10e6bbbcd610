"""Tensor-core filter of forward variant 20 (csrc/nn_mma.cuh): layout and error bound of the raw
filter values, and bit parity of the whole kernel in every launch configuration."""
import ctypes

import numpy as np
import pytest
import torch

from util import bits_equal, cloud

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
U = 2.0 ** -24
MMA_CFGS = [1, 2, 3, 4, 5, 21, 1021, 7021, 23, 1023, 7023, 50023]
# 21: persistent kernel (one CTA per SM), 23: balanced persistent kernel (jobs split over warps by target
# blocks, published row keys merged); G*1000+21 / +23: the same with G CTAs


def t(a):
    return torch.from_numpy(np.ascontiguousarray(a)).to(DEV)


def mma_filter(lib, q, tg):
    n, m = q.shape[0], tg.shape[0]
    out = torch.full((n, m), float("nan"), device=DEV)
    tq, tt = t(q), t(tg)
    p = ctypes.c_void_p
    rc = lib.ga_debug_mma_filter(n, m, p(tq.data_ptr()), p(tt.data_ptr()), p(out.data_ptr()),
                                 p(torch.cuda.current_stream().cuda_stream))
    assert rc == 0
    torch.cuda.synchronize()
    return out.cpu().numpy()


@pytest.mark.parametrize("n,m,scale", [(300, 2048, 1.0), (64, 100, 1.0), (257, 1999, 37.5), (130, 515, 1e-3)])
def test_filter_values_within_documented_bound(ga, n, m, scale):
    """|h - (|t|^2 - 2 q.t)| <= e2 = 330 u s^2 (nn_mma.cuh), s = max|q_c| + max|t_c|; every (q,t) is written."""
    from geometric_adv_b200 import _lib
    lib = _lib.load()
    q = cloud(7, (n, 3)) * np.float32(scale)
    tg = cloud(8, (m, 3)) * np.float32(scale)
    h = mma_filter(lib, q, tg)
    assert not np.isnan(h).any(), "some (query, target) pairs were never written: column map is wrong"
    q64, t64 = q.astype(np.float64), tg.astype(np.float64)
    g = (t64 * t64).sum(1)[None, :] - 2.0 * q64 @ t64.T
    s = float(np.abs(q).max() + np.abs(tg).max())
    err = np.abs(h.astype(np.float64) - g).max() / (U * s * s)
    print("max |h-g| = %.1f u s^2 (bound 330)" % err)
    assert err <= 330.0


def run_variant(ga, lib, a, b, cfg, mode=0):
    if cfg % 1000 in (21, 23):
        lib.ga_set_tuning(0, cfg % 1000)
        lib.ga_set_tuning(8, cfg // 1000)
    else:
        lib.ga_set_tuning(0, 20)
        lib.ga_set_tuning(7, cfg)
    try:
        return [x.cpu().numpy() for x in ga.nn_distance(t(a), t(b), mode)]
    finally:
        lib.ga_set_tuning(0, 0)
        lib.ga_set_tuning(7, 0)
        lib.ga_set_tuning(8, 0)


def check(ga, oracle, a, b, mode=0, cfgs=MMA_CFGS):
    from geometric_adv_b200 import _lib
    lib = _lib.load()
    want = oracle.nn_distance(a, b, mode)
    for cfg in cfgs:
        if cfg % 1000 in (21, 23) and (a.shape[1] > 2048 or b.shape[1] > 2048 or a.shape[1] == 0 or b.shape[1] == 0):
            continue  # the persistent kernel takes clouds of at most 2048 points
        got = run_variant(ga, lib, a, b, cfg, mode)
        for nme, g, w in zip(["dist1", "idx1", "dist2", "idx2"], got, want):
            assert bits_equal(g, w), "%s differs (shape %s vs %s, mode %d, mma cfg %d): %d mismatches" % (
                nme, a.shape, b.shape, mode, cfg, int(np.sum(g != w)))


@pytest.mark.parametrize("mode", [0, 1])
def test_config1_bit_exact(ga, oracle, mode):
    check(ga, oracle, cloud(0, (1, 2048, 3)), cloud(1, (1, 2048, 3)), mode)


@pytest.mark.parametrize("shape", [(1, 1, 1), (2, 1, 7), (3, 33, 1), (2, 63, 65), (2, 64, 64), (1, 129, 127),
                                   (3, 255, 513), (1, 2500, 2048), (2, 2025, 2048), (1, 3, 5000),
                                   (1, 4097, 31), (1, 6000, 6001)])
def test_ragged_shapes(ga, oracle, shape):
    b, n, m = shape
    check(ga, oracle, cloud(100 + n, (b, n, 3)), cloud(200 + m, (b, m, 3)))


def test_batch_of_8(ga, oracle):
    check(ga, oracle, cloud(21, (8, 2048, 3)), cloud(22, (8, 2048, 3)), cfgs=[1, 4, 21, 3021, 29021, 23, 3023, 29023])
    check(ga, oracle, cloud(23, (9, 1000, 3)), cloud(24, (9, 777, 3)), cfgs=[21, 2021, 5021, 13021, 23, 2023, 5023, 13023])


def test_adversarial_duplicates_and_grids(ga, oracle):
    a = cloud(0, (2, 2048, 3))
    near = (a + np.random.default_rng(2).standard_normal(a.shape).astype(np.float32) * np.float32(1e-3)).astype(
        np.float32)
    check(ga, oracle, a, near)
    check(ga, oracle, a, a.copy())                      # exact duplicates: distance 0, lowest index
    grid = np.stack(np.meshgrid(*[np.arange(12, dtype=np.float32) / 12] * 3, indexing="ij"), -1).reshape(1, -1, 3)
    check(ga, oracle, grid, grid[:, ::-1].copy())       # massive exact ties
    check(ga, oracle, np.zeros((1, 300, 3), np.float32), np.zeros((1, 700, 3), np.float32))


def test_scales_and_offsets(ga, oracle):
    """Large offsets make the window wide (many tiles qualify); tiny scales hit the absolute term."""
    a, b = cloud(31, (1, 1500, 3)), cloud(32, (1, 1800, 3))
    for scale, off in [(1e3, 0.0), (1.0, 100.0), (1e-6, 0.0), (1e-18, 0.0), (1e15, 0.0), (1e-30, 0.0)]:
        check(ga, oracle, (a * np.float32(scale) + np.float32(off)).astype(np.float32),
              (b * np.float32(scale) + np.float32(off)).astype(np.float32), cfgs=[1, 4, 21, 23])


def test_non_finite_inputs(ga, oracle):
    a, b = cloud(41, (2, 700, 3)), cloud(42, (2, 900, 3))
    a[0, 5, 1] = np.nan
    b[0, 0, 0] = np.nan          # NaN seed target (k == 0)
    b[1, 17, 2] = np.inf
    a[1, 3, 0] = -np.inf
    b[1, 100] = 3e38             # overflowing distances
    check(ga, oracle, a, b, cfgs=[1, 4, 21, 23])
    check(ga, oracle, a, b, mode=1, cfgs=[1, 21, 23])


def test_full_size_equals_plain_kernel(ga):
    """B=50, N=M=2048 (BASELINE config 2): identical bits to the fp32-filter kernel."""
    from geometric_adv_b200 import _lib
    lib = _lib.load()
    a, b = cloud(2, (50, 2048, 3)), cloud(3, (50, 2048, 3))
    lib.ga_set_tuning(0, 1)
    try:
        want = [x.cpu().numpy() for x in ga.nn_distance(t(a), t(b), 0)]
    finally:
        lib.ga_set_tuning(0, 0)
    for cfg in MMA_CFGS:
        got = run_variant(ga, lib, a, b, cfg)
        for g, w in zip(got, want):
            assert bits_equal(g, w), "cfg %d" % cfg
