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
MMA_CFGS = [1, 2, 3, 4, 5, 21, 1021, 7021, 23, 1023, 7023, 50023, 24, 1024, 7024, 50024]
# 21: persistent kernel (one CTA per SM), 23: balanced persistent kernel (jobs split over warps by target
# blocks, published row keys merged), 24: warp-specialised persistent kernel (scan warps / helper warps);
# G*1000+21 / +23 / +24: the same with G CTAs


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
    if cfg % 1000 in (21, 23, 24):
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
        if cfg % 1000 in (21, 23, 24) and (a.shape[1] > 2048 or b.shape[1] > 2048 or a.shape[1] == 0 or b.shape[1] == 0):
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
    check(ga, oracle, cloud(21, (8, 2048, 3)), cloud(22, (8, 2048, 3)), cfgs=[1, 4, 21, 3021, 29021, 23, 3023, 29023, 24, 3024, 29024])
    check(ga, oracle, cloud(23, (9, 1000, 3)), cloud(24, (9, 777, 3)), cfgs=[21, 2021, 5021, 13021, 23, 2023, 5023, 13023, 24, 2024, 5024, 13024])


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
              (b * np.float32(scale) + np.float32(off)).astype(np.float32), cfgs=[1, 4, 21, 23, 24])


def test_non_finite_inputs(ga, oracle):
    a, b = cloud(41, (2, 700, 3)), cloud(42, (2, 900, 3))
    a[0, 5, 1] = np.nan
    b[0, 0, 0] = np.nan          # NaN seed target (k == 0)
    b[1, 17, 2] = np.inf
    a[1, 3, 0] = -np.inf
    b[1, 100] = 3e38             # overflowing distances
    check(ga, oracle, a, b, cfgs=[1, 4, 21, 23, 24])
    check(ga, oracle, a, b, mode=1, cfgs=[1, 21, 23, 24])


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


# ---- adversarial search for the filter bound ------------------------------------
def _force_low_bits(v, pattern, nbits):
    """Replace the low `nbits` mantissa bits of every fp32 in v by `pattern` (bf16-split residual extremes)."""
    i = v.astype(np.float32).view(np.uint32)
    mask = np.uint32((1 << nbits) - 1)
    return ((i & ~mask) | (np.uint32(pattern) & mask)).view(np.float32)


def _adversarial_families(rng, n, m):
    """(name, queries, targets): inputs built to maximise |h - g| / (u s^2).
    The dropped terms Q1 t3 + Q3 t1 are largest when the third bf16 piece of every coordinate is at its
    extreme (low mantissa bits 0x7f/0x80/0xff after two bf16 roundings) and |q_c| = A, |t_c| = Bm on all
    three axes; the accumulation error is largest when the 15 products have the same sign and the
    largest exponent spread; the norm rounding when |t|^2 sits just below a power of two."""
    fams = []
    corners = np.array([[sx, sy, sz] for sx in (-1, 1) for sy in (-1, 1) for sz in (-1, 1)], np.float32)
    for pat, nb in [(0xFF, 8), (0x7F, 8), (0x80, 8), (0x81, 8), (0xFFFF, 16), (0x7FFF, 16), (0x8000, 16),
                    (0x807F, 16), (0x7F80, 16), (0xFF7F, 16), (0x80FF, 16)]:
        for lo in (0.5, 0.75, 0.99):
            q = (rng.uniform(lo, 1.0, (n, 3)).astype(np.float32) * corners[rng.integers(0, 8, n)])
            t_ = (rng.uniform(lo, 1.0, (m, 3)).astype(np.float32) * corners[rng.integers(0, 8, m)])
            fams.append(("bits%x/%d lo%.2f" % (pat, nb, lo), _force_low_bits(q, pat, nb), _force_low_bits(t_, pat, nb)))
    # same-sign products (q = -t direction: every term of -2 q.t and |t|^2 positive), and the cancelling case
    base = rng.uniform(0.5, 1.0, (m, 3)).astype(np.float32)
    fams.append(("all-positive terms", -base[:n] * np.float32(0.999), base))
    fams.append(("cancelling q=t/2", (base[:n] * np.float32(0.5)), base))
    fams.append(("cancelling q=t", base[:n].copy(), base))
    # mixed magnitudes: one axis dominant, others tiny (exponent spread inside the accumulation)
    mix_q = rng.uniform(-1, 1, (n, 3)).astype(np.float32) * np.array([1.0, 2.0 ** -9, 2.0 ** -17], np.float32)
    mix_t = rng.uniform(-1, 1, (m, 3)).astype(np.float32) * np.array([2.0 ** -17, 1.0, 2.0 ** -9], np.float32)
    fams.append(("mixed magnitudes", mix_q, mix_t))
    # |t|^2 just below / above powers of two, coordinates just below powers of two
    near2 = np.nextafter(np.float32(1.0), np.float32(0.0)) * np.ones((m, 3), np.float32) * corners[rng.integers(0, 8, m)]
    fams.append(("just below 1", near2[:n] * np.float32(-1), near2))
    sq = np.float32(np.sqrt(1.0 / 3.0))
    fams.append(("norm near 1", rng.uniform(-1, 1, (n, 3)).astype(np.float32),
                 (np.full((m, 3), sq, np.float32) * corners[rng.integers(0, 8, m)])))
    # adversarial-attack regime: the two clouds nearly coincide, plus an offset from the origin
    cl = rng.uniform(-0.5, 0.5, (m, 3)).astype(np.float32)
    fams.append(("near-duplicate", cl[:n] + rng.normal(0, 1e-4, (n, 3)).astype(np.float32), cl))
    fams.append(("offset 3", cl[:n] + np.float32(3.0), cl + np.float32(3.0)))
    return fams


def test_filter_bound_adversarial_search(ga, oracle):
    """Directed search over inputs that stress each term of the bound in nn_mma.cuh; the measured worst case
    must stay below e2 = 330 u s^2, and the whole kernel must still return the reference's bits on them."""
    from geometric_adv_b200 import _lib
    lib = _lib.load()
    rng = np.random.default_rng(1234)
    n, m = 256, 2048
    worst = (0.0, None)
    for name, q, tg in _adversarial_families(rng, n, m):
        q = np.ascontiguousarray(q, np.float32)
        tg = np.ascontiguousarray(tg, np.float32)
        h = mma_filter(lib, q, tg).astype(np.float64)
        q64, t64 = q.astype(np.float64), tg.astype(np.float64)
        g = (t64 * t64).sum(1)[None, :] - 2.0 * q64 @ t64.T
        s = float(np.abs(q).max() + np.abs(tg).max())
        err = float(np.abs(h - g).max() / (U * s * s))
        if err > worst[0]:
            worst = (err, name)
        assert err <= 330.0, "%s: |h-g| = %.1f u s^2 exceeds the documented bound" % (name, err)
        check(ga, oracle, q[None], tg[None], cfgs=[4, 5])
    print("adversarial worst case: %.1f u s^2 (%s), bound 330" % worst)
    # the worst case of the round-2 search is committed in tests/golden/mma_filter_worst.json; a larger value
    # here means the search found something new -- still within the bound, but worth recording
    import json
    import os
    os.makedirs("gpurun_out", exist_ok=True)
    with open(os.path.join("gpurun_out", "mma_filter_worst.json"), "w") as f:
        json.dump({"err_u_s2": worst[0], "family": worst[1], "bound": 330.0}, f)


def test_filter_bound_hypothesis(ga):
    """Random search (hypothesis): scales, offsets and bit patterns drawn freely; bound must hold."""
    from hypothesis import given, settings, strategies as st
    from geometric_adv_b200 import _lib
    lib = _lib.load()

    @settings(max_examples=40, deadline=None, derandomize=True)
    @given(seed=st.integers(0, 2 ** 31 - 1), eq=st.integers(-20, 20), et=st.integers(-20, 20),
           pat=st.integers(0, 0xFFFF), off=st.floats(-2.0, 2.0))
    def run(seed, eq, et, pat, off):
        rng = np.random.default_rng(seed)
        q = _force_low_bits(rng.uniform(-1, 1, (64, 3)).astype(np.float32) * np.float32(2.0 ** eq) + np.float32(off * 2.0 ** eq), pat, 16)
        tg = _force_low_bits(rng.uniform(-1, 1, (512, 3)).astype(np.float32) * np.float32(2.0 ** et) + np.float32(off * 2.0 ** et), pat >> 3, 16)
        h = mma_filter(lib, q, tg).astype(np.float64)
        g = (tg.astype(np.float64) ** 2).sum(1)[None, :] - 2.0 * q.astype(np.float64) @ tg.astype(np.float64).T
        s = float(np.abs(q).max() + np.abs(tg).max())
        assert np.abs(h - g).max() <= 330.0 * U * s * s

    run()


@pytest.mark.parametrize("off", [(0.5, 0.5, 0.5), (3.0, -2.0, 0.25), (-10.0, 10.0, 10.0), (1000.0, 0.0, -1000.0)])
def test_off_origin_clouds_use_a_centred_frame_and_stay_bit_exact(ga, oracle, off):
    """Clouds away from the origin: the frame kernel (nn_fwd_mma_kernel<..., FRAME = true>, forced with key 25 = 1)
    evaluates its filter in a frame centred on the target cloud (Frame, nn_tiles.cuh); the results are the reference's
    bits, whatever the frame."""
    from geometric_adv_b200 import _lib
    lib = _lib.load()
    o = np.asarray(off, np.float32)
    a = (cloud(51, (3, 2048, 3)) + o).astype(np.float32)
    b = (cloud(52, (3, 2048, 3)) + o).astype(np.float32)
    lib.ga_set_tuning(25, 1)
    try:
        check(ga, oracle, a, b, cfgs=[1, 2, 3, 4, 5])
        assert lib.ga_last_kernel().decode() == "nn_fwd_mma_kernel<frame>"
        check(ga, oracle, a, b, mode=1, cfgs=[5])
        # the query cloud somewhere else than the target cloud; ragged sizes; a chunked cloud (frame kept over chunks)
        check(ga, oracle, (a[:2, :700] - 2 * o).astype(np.float32), b[:2, :1999], cfgs=[1, 5])
        big = (cloud(53, (1, 5000, 3)) * np.float32(2.0) + o).astype(np.float32)
        check(ga, oracle, a[:1, :300], big, cfgs=[1, 4, 5])
        # exact ties inside an offset cloud, and non-finite points (such a cloud keeps c = 0)
        check(ga, oracle, a[:1], a[:1].copy(), cfgs=[5])
        c, d = a[:2, :900].copy(), b[:2, :800].copy()
        c[0, 5, 1] = np.nan
        d[0, 0, 0] = np.nan
        d[1, 17, 2] = np.inf
        c[1, 3, 0] = -np.inf
        check(ga, oracle, c, d, cfgs=[1, 5])
        # a centred cloud and a shell around the origin (candidate by its points, centred by its box) keep c = 0
        check(ga, oracle, cloud(54, (2, 600, 3)), cloud(55, (2, 700, 3)), cfgs=[5])
        sh = cloud(56, (2, 1500, 3))
        sh = (sh / np.abs(sh).max(axis=2, keepdims=True)).astype(np.float32)  # surface of a cube
        check(ga, oracle, sh[:, :800], sh, cfgs=[1, 5])
    finally:
        lib.ga_set_tuning(25, 2)


def test_default_dispatch_switches_to_the_frame_kernel_once_an_off_origin_cloud_was_seen(ga, oracle):
    """Key 25 = 2 (default): the plain kernel reports a cloud away from the origin through a word of mapped host
    memory; launches after that take the frame kernel.  Shells around the origin and centred clouds do not report."""
    from geometric_adv_b200 import _lib
    lib = _lib.load()

    def run(a, b):
        lib.ga_set_tuning(0, 20)
        try:
            got = [x.cpu().numpy() for x in ga.nn_distance(t(a), t(b))]  # .cpu() synchronises: the report has landed
        finally:
            lib.ga_set_tuning(0, 0)
        for g, w in zip(got, oracle.nn_distance(a, b, 0)):
            assert bits_equal(g, w)
        return lib.ga_last_kernel().decode()

    lib.ga_set_tuning(25, 2)  # clears an earlier report
    try:
        cen = cloud(61, (2, 1024, 3))
        sh = (cen / np.abs(cen).max(axis=2, keepdims=True)).astype(np.float32)
        off = (cen + np.float32(7.0)).astype(np.float32)
        assert run(cen, cen[::-1].copy()) == "nn_fwd_mma_kernel"
        assert run(sh, cen) == "nn_fwd_mma_kernel"
        assert run(cen, sh) == "nn_fwd_mma_kernel", "a shell around the origin must not report"
        assert run(off, off[::-1].copy()) == "nn_fwd_mma_kernel", "the reporting launch itself is a plain one"
        assert run(off, off[::-1].copy()) == "nn_fwd_mma_kernel<frame>"
        assert run(cen, cen[::-1].copy()) == "nn_fwd_mma_kernel<frame>", "the report is sticky"
        lib.ga_set_tuning(25, 2)
        assert run(cen, cen[::-1].copy()) == "nn_fwd_mma_kernel", "setting the key clears it"
    finally:
        lib.ga_set_tuning(25, 2)
