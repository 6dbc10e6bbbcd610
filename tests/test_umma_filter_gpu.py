"""tcgen05 / TMEM filter of forward variant 22 (csrc/nn_distance_fwd_umma.cu): operand layout and
error bound of the raw filter values, and bit parity of the whole persistent kernel for every
grid size (shares of one CTA cut clouds and M-tiles at arbitrary places)."""
import ctypes

import numpy as np
import pytest
import torch

from util import bits_equal, cloud

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
U = 2.0 ** -24
GRIDS = [0, 1, 3, 7, 50, 147, 1000]  # 0: one CTA per SM


def t(a):
    return torch.from_numpy(np.ascontiguousarray(a)).to(DEV)


def umma_filter(lib, q, tg):
    n, m = q.shape[0], tg.shape[0]
    out = torch.full((n, m), float("nan"), device=DEV)
    tq, tt = t(q), t(tg)
    p = ctypes.c_void_p
    rc = lib.ga_debug_umma_filter(n, m, p(tq.data_ptr()), p(tt.data_ptr()), p(out.data_ptr()),
                                  p(torch.cuda.current_stream().cuda_stream))
    assert rc == 0, lib.ga_last_error()
    torch.cuda.synchronize()
    return out.cpu().numpy()


@pytest.mark.parametrize("n,m,scale", [(128, 256, 1.0), (300, 2048, 1.0), (64, 100, 1.0), (257, 1999, 37.5),
                                       (130, 515, 1e-3)])
def test_filter_values_within_documented_bound(ga, n, m, scale):
    """|h - (|t|^2 - 2 q.t)| <= e2 = 330 u s^2 (nn_mma.cuh), s = max|q_c| + max|t_c|; every (q,t) is
    written: TMEM lane = query row, TMEM column = target, both in natural order."""
    from geometric_adv_b200 import _lib
    lib = _lib.load()
    q = cloud(7, (n, 3)) * np.float32(scale)
    tg = cloud(8, (m, 3)) * np.float32(scale)
    h = umma_filter(lib, q, tg)
    assert not np.isnan(h).any(), "some (query, target) pairs were never written"
    q64, t64 = q.astype(np.float64), tg.astype(np.float64)
    g = (t64 * t64).sum(1)[None, :] - 2.0 * q64 @ t64.T
    s = float(np.abs(q).max() + np.abs(tg).max())
    err = np.abs(h.astype(np.float64) - g) / (U * s * s)
    print("max |h-g| = %.1f u s^2 (bound 330); rows with error > bound: %s" % (
        err.max(), np.unique(np.argwhere(err > 330.0)[:, 0])[:8]))
    assert err.max() <= 330.0


def run_variant(ga, lib, a, b, grid, mode=0):
    lib.ga_set_tuning(0, 22)
    lib.ga_set_tuning(12, grid)
    try:
        return [x.cpu().numpy() for x in ga.nn_distance(t(a), t(b), mode)]
    finally:
        lib.ga_set_tuning(0, 0)
        lib.ga_set_tuning(12, 0)


def check(ga, oracle, a, b, mode=0, grids=GRIDS):
    from geometric_adv_b200 import _lib
    lib = _lib.load()
    want = oracle.nn_distance(a, b, mode)
    for grid in grids:
        got = run_variant(ga, lib, a, b, grid, mode)
        for nme, g, w in zip(["dist1", "idx1", "dist2", "idx2"], got, want):
            assert bits_equal(g, w), "%s differs (shape %s vs %s, mode %d, grid %d): %d mismatches" % (
                nme, a.shape, b.shape, mode, grid, int(np.sum(g != w)))


@pytest.mark.parametrize("mode", [0, 1])
def test_config1_bit_exact(ga, oracle, mode):
    check(ga, oracle, cloud(0, (1, 2048, 3)), cloud(1, (1, 2048, 3)), mode)


@pytest.mark.parametrize("shape", [(1, 257, 257), (2, 300, 2048), (3, 513, 1000), (1, 2025, 2048), (2, 2047, 258),
                                   (5, 1280, 1281), (1, 385, 1793)])
def test_ragged_shapes(ga, oracle, shape):
    b, n, m = shape
    check(ga, oracle, cloud(100 + n, (b, n, 3)), cloud(200 + m, (b, m, 3)))


def test_unsupported_sizes_are_refused(ga):
    from geometric_adv_b200 import _lib
    lib = _lib.load()
    lib.ga_set_tuning(0, 22)
    try:
        with pytest.raises(NotImplementedError):
            ga.nn_distance(t(cloud(1, (1, 100, 3))), t(cloud(2, (1, 2048, 3))))
        with pytest.raises(NotImplementedError):
            ga.nn_distance(t(cloud(1, (1, 2048, 3))), t(cloud(2, (1, 2049, 3))))
    finally:
        lib.ga_set_tuning(0, 0)


def test_batches(ga, oracle):
    check(ga, oracle, cloud(21, (8, 2048, 3)), cloud(22, (8, 2048, 3)), grids=[0, 5, 33])
    check(ga, oracle, cloud(23, (9, 1000, 3)), cloud(24, (9, 777, 3)), grids=[0, 2, 13])


def test_adversarial_duplicates_and_grids(ga, oracle):
    a = cloud(0, (2, 2048, 3))
    near = (a + np.random.default_rng(2).standard_normal(a.shape).astype(np.float32) * np.float32(1e-3)).astype(
        np.float32)
    check(ga, oracle, a, near, grids=[0, 3])
    check(ga, oracle, a, a.copy(), grids=[0, 3])        # exact duplicates: distance 0, lowest index
    grid = np.stack(np.meshgrid(*[np.arange(12, dtype=np.float32) / 12] * 3, indexing="ij"), -1).reshape(1, -1, 3)
    check(ga, oracle, grid, grid[:, ::-1].copy(), grids=[0, 3])       # massive exact ties
    check(ga, oracle, np.zeros((1, 300, 3), np.float32), np.zeros((1, 700, 3), np.float32), grids=[0, 2])


def test_scales_and_offsets(ga, oracle):
    a, b = cloud(31, (1, 1500, 3)), cloud(32, (1, 1800, 3))
    for scale, off in [(1e3, 0.0), (1.0, 100.0), (1e-6, 0.0), (1e-18, 0.0), (1e15, 0.0), (1e-30, 0.0)]:
        check(ga, oracle, (a * np.float32(scale) + np.float32(off)).astype(np.float32),
              (b * np.float32(scale) + np.float32(off)).astype(np.float32), grids=[0, 3])


def test_non_finite_inputs(ga, oracle):
    a, b = cloud(41, (2, 700, 3)), cloud(42, (2, 900, 3))
    a[0, 5, 1] = np.nan
    b[0, 0, 0] = np.nan          # NaN seed target (k == 0)
    b[1, 17, 2] = np.inf
    a[1, 3, 0] = -np.inf
    b[1, 100] = 3e38             # overflowing distances
    check(ga, oracle, a, b, grids=[0, 3])
    check(ga, oracle, a, b, mode=1, grids=[0])


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
    for grid in (0, 100):
        got = run_variant(ga, lib, a, b, grid)
        for g, w in zip(got, want):
            assert bits_equal(g, w), "grid %d" % grid
