"""GPU parity tests (run on the B200 box: pytest -m gpu).  Every comparison goes through
the C ABI (python shim -> libga_b200.so) and is checked against the CPU oracle, the golden
fixtures generated from the reference, and -- when it travelled with the repo -- the
reference's own CUDA kernels (oracle/_ref/libga_ref_gpu.so).

Bar: bit-exact for dist / idx / gradients / kNN values and indices."""
import ctypes
import os

import numpy as np
import pytest
import torch

from util import bits_equal, cloud, digests, golden, sha

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DEV = "cuda:0"


def t(a):
    return torch.from_numpy(np.ascontiguousarray(a)).to(DEV)


def run_fwd(ga, a, b, mode=0):
    return [x.cpu().numpy() for x in ga.nn_distance(t(a), t(b), mode)]


def check_fwd(ga, oracle, a, b, mode=0):
    """Every forward path against the oracle: the plain scan, the Morton-ordered scan with tile
    skipping, and the cluster kernel that splits the targets over 2 / 8 CTAs (forced here; the
    library picks it by itself for small batches)."""
    from geometric_adv_b200 import _lib
    lib = _lib.load()
    want = oracle.nn_distance(a, b, mode)
    names = ["dist1", "idx1", "dist2", "idx2"]
    got = None
    variants = [("plain", False, 0), ("pruned", True, 0)]
    if 0 < a.shape[1] <= 2048 and 0 < b.shape[1] <= 2048:
        variants += [("split2", False, 2), ("split8", False, 8)]
    for name, pruning, split in variants:
        ga.set_pruning(pruning)
        lib.ga_set_tuning(5, split)
        try:
            got = run_fwd(ga, a, b, mode)
        finally:
            ga.set_pruning(False)
            lib.ga_set_tuning(5, -1)
        for nme, g, w in zip(names, got, want):
            assert bits_equal(g, w), "%s differs (shape %s vs %s, mode %d, path %s): %d mismatches" % (
                nme, a.shape, b.shape, mode, name, int(np.sum(g != w)))
    return run_fwd(ga, a, b, mode)


# ------------------------------------------------------------------ forward
@pytest.mark.parametrize("mode", [0, 1])
def test_config1_bit_exact(ga, oracle, mode):
    """BASELINE config 1: two random 2048x3 clouds, batch 1, dist+idx bit-exact."""
    a, b = cloud(0, (1, 2048, 3)), cloud(1, (1, 2048, 3))
    d1, i1, d2, i2 = check_fwd(ga, oracle, a, b, mode)
    if mode == 0:
        c = digests()["cfg1"]  # digests of the reference CPU kernel's own output
        assert (sha(d1), sha(i1), sha(d2), sha(i2)) == (c["dist1"], c["idx1"], c["dist2"], c["idx2"])


def test_config1_variants_against_reference_digests(ga, oracle):
    dg = digests()
    a = cloud(0, (1, 2048, 3))
    b2 = (a + np.random.default_rng(2).standard_normal(a.shape).astype(np.float32) * np.float32(1e-3)).astype(
        np.float32)
    for key, b in [("cfg1_adv", b2), ("cfg1_dup", a)]:
        d1, i1, d2, i2 = run_fwd(ga, a, b)
        c = dg[key]
        assert (sha(d1), sha(i1), sha(d2), sha(i2)) == (c["dist1"], c["idx1"], c["dist2"], c["idx2"]), key


def test_golden_fixtures(ga):
    for name in ["nnd_unit_4x100x200.npz", "nnd_ties_3x257x131.npz"]:
        g = golden(name)
        d1, i1, d2, i2 = run_fwd(ga, g["xyz1"], g["xyz2"])
        assert bits_equal(d1, g["dist1"]) and bits_equal(d2, g["dist2"]), name
        assert np.array_equal(i1, g["idx1"]) and np.array_equal(i2, g["idx2"]), name
        g1, g2 = ga.nn_distance_grad(t(g["xyz1"]), t(g["xyz2"]), t(g["gd1"]), t(g["idx1"]), t(g["gd2"]),
                                     t(g["idx2"]))
        assert bits_equal(g1.cpu().numpy(), g["gxyz1"]) and bits_equal(g2.cpu().numpy(), g["gxyz2"]), name


def test_unit_test_py_contract(ga):
    """unit_test.py:14-35 against chamfer_python (fp64): tolerance 1e-8, indices identical;
    torch return order (dist1, dist2, idx1, idx2)."""
    g = golden("nnd_unit_4x100x200.npz")
    p = golden("chamfer_python_4x100x200.npz")
    cham = ga.chamfer_3DDist()
    p2 = t(g["xyz2"]).requires_grad_(True)
    dist1, dist2, idx1, idx2 = cham(t(g["xyz1"]), p2)
    torch.sum(dist1).backward()
    d1 = (dist1.detach().cpu().numpy() - p["dist1"]) ** 2
    d2 = (dist2.detach().cpu().numpy() - p["dist2"]) ** 2
    assert d1.mean() + d2.mean() < 1e-8
    assert np.array_equal(idx1.cpu().numpy(), p["idx1"]) and np.array_equal(idx2.cpu().numpy(), p["idx2"])
    assert p2.grad is not None and p2.grad.shape == p2.shape


@pytest.mark.parametrize("shape", [(1, 1, 1), (2, 1, 7), (3, 33, 1), (2, 63, 65), (2, 64, 64), (1, 129, 127),
                                   (3, 255, 513), (1, 2500, 2048), (2, 2025, 2048), (1, 3, 5000),
                                   (1, 4097, 31), (1, 6000, 6001)])
@pytest.mark.parametrize("mode", [0, 1])
def test_ragged_shapes(ga, oracle, shape, mode):
    b, n, m = shape
    check_fwd(ga, oracle, cloud(100 + n, (b, n, 3)), cloud(200 + m, (b, m, 3)), mode)


def test_batched_attack_shape_slice(ga, oracle):
    """BASELINE config 2 inputs (seed 2/3), first 4 clouds vs reference digests, 8 clouds vs oracle."""
    a, b = cloud(2, (4, 2048, 3)), cloud(3, (4, 2048, 3))
    d1, i1, d2, i2 = run_fwd(ga, a, b)
    c = digests()["cfg2_b4"]
    assert (sha(d1), sha(i1), sha(d2), sha(i2)) == (c["dist1"], c["idx1"], c["dist2"], c["idx2"])
    check_fwd(ga, oracle, cloud(21, (8, 2048, 3)), cloud(22, (8, 2048, 3)))


def test_full_size_properties(ga):
    """B=50, N=M=2048 (config 2): size-independent properties instead of a CPU rerun:
    the returned distance is the reference arithmetic of (query, target[idx]), no sampled
    target is closer, equal-distance earlier targets do not exist, batch shards == whole."""
    a, b = cloud(2, (50, 2048, 3)), cloud(3, (50, 2048, 3))
    d1, i1, d2, i2 = run_fwd(ga, a, b)
    assert i1.min() >= 0 and i1.max() < 2048 and i2.min() >= 0 and i2.max() < 2048
    for q, tg, d, i in [(a, b, d1, i1), (b, a, d2, i2)]:
        sel = np.take_along_axis(tg, i[..., None].astype(np.int64), axis=1)
        x = sel - q
        dd = (x[..., 0] * x[..., 0] + x[..., 1] * x[..., 1]) + x[..., 2] * x[..., 2]
        assert bits_equal(dd.astype(np.float32), d)
        rng = np.random.default_rng(9)
        for _ in range(8):
            j = rng.integers(0, 2048, size=(50, 2048))
            o = np.take_along_axis(tg, j[..., None], axis=1) - q
            od = ((o[..., 0] * o[..., 0] + o[..., 1] * o[..., 1]) + o[..., 2] * o[..., 2]).astype(np.float32)
            assert np.all((od > d) | ((od == d) & (j >= i)))
    # shard equivalence: the op on a slice of the batch equals the slice of the op
    s1 = run_fwd(ga, a[10:20], b[10:20])
    assert bits_equal(s1[0], d1[10:20]) and np.array_equal(s1[1], i1[10:20])
    assert bits_equal(s1[2], d2[10:20]) and np.array_equal(s1[3], i2[10:20])


@pytest.mark.parametrize("mode", [0, 1])
def test_ties_duplicates_and_grids(ga, oracle, mode):
    rng = np.random.default_rng(5)
    a = (rng.integers(0, 3, (2, 700, 3)).astype(np.float32) * np.float32(0.5) - np.float32(0.5))
    b = (rng.integers(0, 3, (2, 900, 3)).astype(np.float32) * np.float32(0.5) - np.float32(0.5))
    check_fwd(ga, oracle, a, b, mode)          # massive exact ties: lowest index must win
    check_fwd(ga, oracle, a, a.copy(), mode)   # self distance 0 with duplicates
    z = np.zeros((1, 300, 3), np.float32)
    check_fwd(ga, oracle, z, z, mode)          # everything identical
    near = cloud(6, (1, 1500, 3))
    near2 = (near + np.float32(1e-7) * np.random.default_rng(7).standard_normal(near.shape)).astype(np.float32)
    check_fwd(ga, oracle, near2, near, mode)   # attack initialisation: pert sigma 1e-7 (adversary.py:27)


@pytest.mark.parametrize("scale,offset", [(1e-20, 0.0), (1e-30, 0.0), (1e6, 0.0), (1e15, 0.0), (1e19, 0.0),
                                          (1e25, 0.0), (1.0, 100.0), (1.0, 1e4), (1e-3, 1.0)])
def test_dynamic_range(ga, oracle, scale, offset):
    """Denormal results, overflow to inf, and clouds far from the origin: the filter window
    widens (or degenerates to the plain scan) but the result stays exact."""
    a = (cloud(8, (2, 300, 3)).astype(np.float64) * scale + offset).astype(np.float32)
    b = (cloud(9, (2, 400, 3)).astype(np.float64) * scale + offset).astype(np.float32)
    check_fwd(ga, oracle, a, b, 0)
    check_fwd(ga, oracle, a, b, 1)


def test_non_finite_inputs(ga, oracle):
    a, b = cloud(40, (3, 200, 3)), cloud(41, (3, 260, 3))
    b[0, 0, 1] = np.nan      # NaN target 0: reference seeds best with NaN and never replaces it
    b[0, 77, 0] = np.nan     # NaN elsewhere is skipped
    b[1, 5, 0] = np.inf
    a[1, 3, 2] = np.nan      # NaN query
    a[2, 9, 0] = -np.inf     # inf query
    b[2, 100, 2] = np.inf
    check_fwd(ga, oracle, a, b, 0)
    check_fwd(ga, oracle, a, b, 1)


def test_empty_inputs(ga, oracle):
    z = np.zeros((2, 0, 3), np.float32)
    a = cloud(1, (2, 5, 3))
    for x, y in [(a, z), (z, a), (z, z), (np.zeros((0, 4, 3), np.float32), np.zeros((0, 6, 3), np.float32))]:
        got = run_fwd(ga, x, y)
        want = oracle.nn_distance(x, y, 0)
        for g, w in zip(got, want):
            assert g.shape == w.shape and bits_equal(g, w)


def test_matches_reference_cuda_kernel_on_this_gpu(ga):
    """mode 1 == the reference's own NmDistanceKernel compiled for sm_100a, bit for bit."""
    path = os.path.join(ROOT, "oracle", "_ref", "libga_ref_gpu.so")
    if not os.path.exists(path):
        pytest.skip("oracle/_ref/libga_ref_gpu.so did not travel with this snapshot")
    ref = ctypes.CDLL(path)
    for seed, (b, n, m) in enumerate([(4, 2048, 2048), (2, 1000, 777), (1, 5, 3000)]):
        x1, x2 = t(cloud(60 + seed, (b, n, 3))), t(cloud(70 + seed, (b, m, 3)))
        rd1 = torch.empty(b, n, device=DEV); ri1 = torch.empty(b, n, dtype=torch.int32, device=DEV)
        rd2 = torch.empty(b, m, device=DEV); ri2 = torch.empty(b, m, dtype=torch.int32, device=DEV)
        torch.cuda.synchronize()
        p = ctypes.c_void_p
        rc = ref.ga_refgpu_nn_distance(b, n, m, p(x1.data_ptr()), p(x2.data_ptr()), p(rd1.data_ptr()),
                                       p(ri1.data_ptr()), p(rd2.data_ptr()), p(ri2.data_ptr()))
        torch.cuda.synchronize()
        assert rc == 0
        d1, i1, d2, i2 = ga.nn_distance(x1, x2, ga.GA_MODE_GPU_REF)
        assert torch.equal(d1.view(torch.int32), rd1.view(torch.int32)) and torch.equal(i1, ri1)
        assert torch.equal(d2.view(torch.int32), rd2.view(torch.int32)) and torch.equal(i2, ri2)


def test_host_entry_point_equals_device_entry_point(ga):
    a, b = cloud(80, (3, 700, 3)), cloud(81, (3, 650, 3))
    dev = run_fwd(ga, a, b)
    host = [x.numpy() for x in ga.nn_distance(torch.from_numpy(a), torch.from_numpy(b))]
    assert all(bits_equal(x, y) for x, y in zip(dev, host))
    gd1 = np.random.default_rng(1).standard_normal((3, 700)).astype(np.float32)
    gd2 = np.random.default_rng(2).standard_normal((3, 650)).astype(np.float32)
    gdev = ga.nn_distance_grad(t(a), t(b), t(gd1), t(dev[1]), t(gd2), t(dev[3]))
    ghost = ga.nn_distance_grad(torch.from_numpy(a), torch.from_numpy(b), torch.from_numpy(gd1),
                                torch.from_numpy(dev[1]), torch.from_numpy(gd2), torch.from_numpy(dev[3]))
    assert all(bits_equal(x.cpu().numpy(), y.numpy()) for x, y in zip(gdev, ghost))


@pytest.mark.parametrize("pinned,path", [(True, 0), (False, 0), (True, 2)])
def test_fwd_bwd_host_entry_point(ga, pinned, path):
    """ga_nn_distance_fwd_bwd_host (what bench.py's e2e leg calls): results identical to the device
    entry points, for pinned and pageable caller memory (1-, 2- and 4-chunk two-stream copies) and
    for the opt-in zero-copy path (ingest kernel, mirrored forward outputs, gradients written
    straight to the host)."""
    from geometric_adv_b200 import _lib
    lib = _lib.load()
    lib.ga_set_tuning(2, path)
    pin = (lambda x: x.pin_memory()) if pinned else (lambda x: x)
    for b, n, m in ((1, 600, 500), (5, 601, 500), (20, 600, 503), (48, 4096, 4096), (160, 4096, 4000)):
        a, c = cloud(90 + b, (b, n, 3)), cloud(91 + b, (b, m, 3))
        gd1 = np.random.default_rng(b).standard_normal((b, n)).astype(np.float32)
        gd2 = np.random.default_rng(b + 1).standard_normal((b, m)).astype(np.float32)
        h = [pin(torch.from_numpy(x)) for x in (a, c, gd1, gd2)]
        d1 = pin(torch.empty(b, n)); i1 = pin(torch.empty(b, n, dtype=torch.int32))
        d2 = pin(torch.empty(b, m)); i2 = pin(torch.empty(b, m, dtype=torch.int32))
        o1 = pin(torch.empty(b, n, 3)); o2 = pin(torch.empty(b, m, 3))
        p = ctypes.c_void_p
        _lib.check(lib.ga_nn_distance_fwd_bwd_host(b, n, m, p(h[0].data_ptr()), p(h[1].data_ptr()),
                                                   p(h[2].data_ptr()), p(h[3].data_ptr()), p(d1.data_ptr()),
                                                   p(i1.data_ptr()), p(d2.data_ptr()), p(i2.data_ptr()),
                                                   p(o1.data_ptr()), p(o2.data_ptr()), 0))
        dev = ga.nn_distance(t(a), t(c))
        g = ga.nn_distance_grad(t(a), t(c), t(gd1), dev[1], t(gd2), dev[3])
        for x, y in zip((d1, i1, d2, i2, o1, o2), tuple(dev) + tuple(g)):
            assert bits_equal(x.numpy(), y.cpu().numpy()), (b, n, m, pinned, path)
    lib.ga_set_tuning(2, 0)


@pytest.mark.parametrize("chunks", [0, 1, 3, 8])
def test_fwd_bwd_host_graph_replay(ga, chunks):
    """The captured pipeline of ga_nn_distance_fwd_bwd_host: the same pinned buffers come back with
    NEW contents every step (a training loop), the replayed graph must give the fresh results; a
    different buffer set, a different shape and pageable buffers in between must not disturb it."""
    from geometric_adv_b200 import _lib
    lib = _lib.load()
    p = ctypes.c_void_p
    lib.ga_set_tuning(11, chunks)
    launches = []
    try:
        def buffers(b, n, m, pin=True):
            f = (lambda x: x.pin_memory()) if pin else (lambda x: x)
            return [f(torch.empty(b, n, 3)), f(torch.empty(b, m, 3)), f(torch.empty(b, n)), f(torch.empty(b, m)),
                    f(torch.empty(b, n)), f(torch.empty(b, n, dtype=torch.int32)), f(torch.empty(b, m)),
                    f(torch.empty(b, m, dtype=torch.int32)), f(torch.empty(b, n, 3)), f(torch.empty(b, m, 3))]

        def step(buf, b, n, m, seed):
            a, c = cloud(seed, (b, n, 3)), cloud(seed + 1, (b, m, 3))
            gd1 = np.random.default_rng(seed).standard_normal((b, n)).astype(np.float32)
            gd2 = np.random.default_rng(seed + 1).standard_normal((b, m)).astype(np.float32)
            for dst, src in zip(buf[:4], (a, c, gd1, gd2)):
                dst.copy_(torch.from_numpy(src))
            for o in buf[4:]:
                o.zero_()
            l0 = ga.launch_count()
            _lib.check(lib.ga_nn_distance_fwd_bwd_host(b, n, m, *[p(x.data_ptr()) for x in buf], 0))
            launches.append(ga.launch_count() - l0)
            dev = ga.nn_distance(t(a), t(c))
            g = ga.nn_distance_grad(t(a), t(c), t(gd1), dev[1], t(gd2), dev[3])
            for x, y in zip(buf[4:], tuple(dev) + tuple(g)):
                assert bits_equal(x.numpy(), y.cpu().numpy()), (b, n, m, seed, chunks)

        b, n, m = 24, 2048, 2000
        main, other, pageable = buffers(b, n, m), buffers(b, n, m), buffers(b, n, m, pin=False)
        small = buffers(3, 300, 257)
        for it in range(4):          # step 0 direct, step 1 captures, steps 2.. replay
            step(main, b, n, m, 500 + 10 * it)
        assert all(x >= 2 for x in launches), launches   # replays keep counting the kernels they run
        step(other, b, n, m, 600)
        step(small, 3, 300, 257, 610)
        step(pageable, b, n, m, 620)
        step(pageable, b, n, m, 630)
        for it in range(3):
            step(main, b, n, m, 700 + 10 * it)
            step(other, b, n, m, 800 + 10 * it)
            step(small, 3, 300, 257, 900 + 10 * it)
    finally:
        lib.ga_set_tuning(11, 0)


@pytest.mark.parametrize("mirror", [1, 2])
@pytest.mark.parametrize("ctas", [1, 8, 40])
def test_fwd_bwd_host_pulled_ingest(ga, ctas, mirror):
    """Tuning key 27: the clouds are pulled out of the pinned buffers by a few CTAs (ingest_stream_kernel), one
    arrival flag per batch element, the search starts behind them as a programmatic dependent.  Same checks."""
    _streamed_ingest_case(ga, 27, ctas, mirror)


@pytest.mark.parametrize("mirror", [1, 2])
@pytest.mark.parametrize("groups", [1, 3, 7, 32])
def test_fwd_bwd_host_streamed_ingest(ga, groups, mirror):
    _streamed_ingest_case(ga, 26, groups, mirror)


def _streamed_ingest_case(ga, key, groups, mirror):
    """Streamed ingest of the replayed host step (tuning key 26): the search starts with the first H2D copy and
    its CTAs wait per batch element for the arrival flag of their group.  Fresh contents every step, several
    shapes, both ways of returning dist/idx; results must be the bits of the device entry points, and the replay
    must really have been the streamed pipeline (never the give-up path)."""
    from geometric_adv_b200 import _lib
    lib = _lib.load()
    lib.ga_debug_host_streamed.restype = ctypes.c_int
    p = ctypes.c_void_p
    lib.ga_set_tuning(key, groups)
    lib.ga_set_tuning(17, mirror)
    lib.ga_set_tuning(0, 20)   # the HMMA grid kernel at every shape below (the automatic choice may be the tcgen05 kernel)
    try:
        for (b, n, m) in [(50, 2048, 2048), (48, 2048, 1024), (60, 1024, 2048), (40, 2048, 2016)]:
            buf = [torch.empty(b, n, 3).pin_memory(), torch.empty(b, m, 3).pin_memory(), torch.empty(b, n).pin_memory(),
                   torch.empty(b, m).pin_memory(), torch.empty(b, n).pin_memory(),
                   torch.empty(b, n, dtype=torch.int32).pin_memory(), torch.empty(b, m).pin_memory(),
                   torch.empty(b, m, dtype=torch.int32).pin_memory(), torch.empty(b, n, 3).pin_memory(),
                   torch.empty(b, m, 3).pin_memory()]
            streamed = []
            for it in range(5):
                seed = 40 + 7 * it + b
                a, c = cloud(seed, (b, n, 3)), cloud(seed + 1, (b, m, 3))
                gd1 = np.random.default_rng(seed).standard_normal((b, n)).astype(np.float32)
                gd2 = np.random.default_rng(seed + 1).standard_normal((b, m)).astype(np.float32)
                for dst, src in zip(buf[:4], (a, c, gd1, gd2)):
                    dst.copy_(torch.from_numpy(src))
                for o in buf[4:]:
                    o.zero_()
                _lib.check(lib.ga_nn_distance_fwd_bwd_host(b, n, m, *[p(x.data_ptr()) for x in buf], 0))
                streamed.append(lib.ga_debug_host_streamed())
                dev = ga.nn_distance(t(a), t(c))
                g = ga.nn_distance_grad(t(a), t(c), t(gd1), dev[1], t(gd2), dev[3])
                for x, y in zip(buf[4:], tuple(dev) + tuple(g)):
                    assert bits_equal(x.numpy(), y.cpu().numpy()), (b, n, m, it, groups, mirror)
            assert streamed[0] == 0 and all(v == 1 for v in streamed[2:]), (b, n, m, streamed)
    finally:
        lib.ga_set_tuning(key, 0)
        lib.ga_set_tuning(17, 0)
        lib.ga_set_tuning(0, 0)


@pytest.mark.parametrize("shape", [(600, 2048, 2048), (1500, 1024, 1024), (700, 2048, 512)])
def test_fwd_bwd_host_pulled_ingest_large_batches(ga, shape):
    """Default dispatch at batch sizes where the search runs many waves behind the ingest (a waiting CTA's patience
    scales with the transfer): the replay is the pulled pipeline and the results are the device entry points' bits."""
    from geometric_adv_b200 import _lib
    lib = _lib.load()
    lib.ga_debug_host_streamed.restype = ctypes.c_int
    p = ctypes.c_void_p
    b, n, m = shape
    g = torch.Generator().manual_seed(b)
    buf = [(torch.rand(b, n, 3, generator=g) - 0.5).pin_memory(), (torch.rand(b, m, 3, generator=g) - 0.5).pin_memory(),
           torch.randn(b, n, generator=g).pin_memory(), torch.randn(b, m, generator=g).pin_memory(),
           torch.empty(b, n).pin_memory(), torch.empty(b, n, dtype=torch.int32).pin_memory(),
           torch.empty(b, m).pin_memory(), torch.empty(b, m, dtype=torch.int32).pin_memory(),
           torch.empty(b, n, 3).pin_memory(), torch.empty(b, m, 3).pin_memory()]
    states = []
    for it in range(3):
        for o in buf[4:]:
            o.fill_(-3)
        _lib.check(lib.ga_nn_distance_fwd_bwd_host(b, n, m, *[p(x.data_ptr()) for x in buf], 0))
        states.append(lib.ga_debug_host_streamed())
    assert states == [0, 1, 1], states
    a, c = buf[0].cuda(), buf[1].cuda()
    dev = ga.nn_distance(a, c)
    gr = ga.nn_distance_grad(a, c, buf[2].cuda(), dev[1], buf[3].cuda(), dev[3])
    for x, y in zip(buf[4:], tuple(dev) + tuple(gr)):
        assert torch.equal(x, y.cpu()), shape


def test_fwd_bwd_host_pulled_ingest_gives_up_and_redoes_the_step(ga):
    """The safety net of the pulled pipeline: if an arrival flag never comes (test hook, key 29: the ingest kernel
    withholds the last one), the search's CTAs for that batch element leave after their patience, the host entry sees
    the give-up word, drops the graph, redoes THAT step on the direct path -- the caller gets correct results -- and
    keeps the buffer set off the pulled pipeline from then on."""
    from geometric_adv_b200 import _lib
    lib = _lib.load()
    lib.ga_debug_host_streamed.restype = ctypes.c_int
    p = ctypes.c_void_p
    b, n, m = 50, 2048, 2048
    buf = [torch.empty(b, n, 3).pin_memory(), torch.empty(b, m, 3).pin_memory(), torch.empty(b, n).pin_memory(),
           torch.empty(b, m).pin_memory(), torch.empty(b, n).pin_memory(),
           torch.empty(b, n, dtype=torch.int32).pin_memory(), torch.empty(b, m).pin_memory(),
           torch.empty(b, m, dtype=torch.int32).pin_memory(), torch.empty(b, n, 3).pin_memory(),
           torch.empty(b, m, 3).pin_memory()]
    lib.ga_set_tuning(29, 1)
    try:
        states = []
        for it in range(4):
            seed = 900 + 3 * it
            a, c = cloud(seed, (b, n, 3)), cloud(seed + 1, (b, m, 3))
            gd1 = np.random.default_rng(seed).standard_normal((b, n)).astype(np.float32)
            gd2 = np.random.default_rng(seed + 1).standard_normal((b, m)).astype(np.float32)
            for dst, src in zip(buf[:4], (a, c, gd1, gd2)):
                dst.copy_(torch.from_numpy(src))
            for o in buf[4:]:
                o.zero_()
            _lib.check(lib.ga_nn_distance_fwd_bwd_host(b, n, m, *[p(x.data_ptr()) for x in buf], 0))
            states.append(lib.ga_debug_host_streamed())
            dev = ga.nn_distance(t(a), t(c))
            g = ga.nn_distance_grad(t(a), t(c), t(gd1), dev[1], t(gd2), dev[3])
            for x, y in zip(buf[4:], tuple(dev) + tuple(g)):
                assert bits_equal(x.numpy(), y.cpu().numpy()), (it, states)
        # call 0 direct, call 1 captures the pulled pipeline and gives up (-1), later calls: chunked pipeline (0)
        assert states[0] == 0 and states[1] == -1 and all(v == 0 for v in states[2:]), states
    finally:
        lib.ga_set_tuning(29, 0)


def test_fwd_bwd_host_streamed_ingest_not_for_ragged_lines(ga):
    """Clouds that are not whole 128-byte lines (m = 2000) and launches that do not take the HMMA grid kernel
    (B = 10: tcgen05 kernel) stay on the chunked pipeline."""
    from geometric_adv_b200 import _lib
    lib = _lib.load()
    lib.ga_debug_host_streamed.restype = ctypes.c_int
    p = ctypes.c_void_p
    lib.ga_set_tuning(26, 4)
    lib.ga_set_tuning(27, 8)
    for (b, n, m) in [(24, 2048, 2000), (10, 2048, 2048)]:
        buf = [torch.rand(b, n, 3).pin_memory(), torch.rand(b, m, 3).pin_memory(), torch.rand(b, n).pin_memory(),
               torch.rand(b, m).pin_memory(), torch.empty(b, n).pin_memory(),
               torch.empty(b, n, dtype=torch.int32).pin_memory(), torch.empty(b, m).pin_memory(),
               torch.empty(b, m, dtype=torch.int32).pin_memory(), torch.empty(b, n, 3).pin_memory(),
               torch.empty(b, m, 3).pin_memory()]
        for it in range(3):
            _lib.check(lib.ga_nn_distance_fwd_bwd_host(b, n, m, *[p(x.data_ptr()) for x in buf], 0))
            assert lib.ga_debug_host_streamed() == 0
    lib.ga_set_tuning(26, 0)
    lib.ga_set_tuning(27, 0)


@pytest.mark.parametrize("shape", [(50, 2048, 2048), (37, 2048, 2000), (10, 2048, 2048), (3, 300, 257), (12, 1000, 4000)])
def test_fused_device_entry_equals_separate_calls(ga, shape):
    """ga_nn_distance_fwd_bwd: gradient CTAs start per batch element behind the search's completion tickets
    and never park behind the grid dependency; outputs must be the bits of the two separate calls, every
    time (repeated: the overlap is timing dependent)."""
    from geometric_adv_b200 import _lib
    lib = _lib.load()
    p = ctypes.c_void_p
    b, n, m = shape
    a, c = cloud(70 + b, (b, n, 3)), cloud(71 + b, (b, m, 3))
    gd1 = np.random.default_rng(b).standard_normal((b, n)).astype(np.float32)
    gd2 = np.random.default_rng(b + 1).standard_normal((b, m)).astype(np.float32)
    x1, x2, g1, g2 = t(a), t(c), t(gd1), t(gd2)
    want = ga.nn_distance(x1, x2)
    wg = ga.nn_distance_grad(x1, x2, g1, want[1], g2, want[3])
    st = torch.cuda.current_stream().cuda_stream
    for rep in range(6):
        d1 = torch.full((b, n), -1.0, device=DEV); i1 = torch.full((b, n), -1, dtype=torch.int32, device=DEV)
        d2 = torch.full((b, m), -1.0, device=DEV); i2 = torch.full((b, m), -1, dtype=torch.int32, device=DEV)
        o1 = torch.full((b, n, 3), float("nan"), device=DEV); o2 = torch.full((b, m, 3), float("nan"), device=DEV)
        _lib.check(lib.ga_nn_distance_fwd_bwd(b, n, m, p(x1.data_ptr()), p(x2.data_ptr()), p(g1.data_ptr()),
                                              p(g2.data_ptr()), p(d1.data_ptr()), p(i1.data_ptr()), p(d2.data_ptr()),
                                              p(i2.data_ptr()), p(o1.data_ptr()), p(o2.data_ptr()), 0, p(st)))
        for got, w in zip((d1, i1, d2, i2, o1, o2), tuple(want) + tuple(wg)):
            assert bits_equal(got.cpu().numpy(), w.cpu().numpy()), (shape, rep)


# ------------------------------------------------------------------ backward
def check_bwd(ga, oracle, a, b, gd1, i1, gd2, i2):
    """Both launch shapes of the gradient kernel: one CTA per cloud, and output points split over 4 CTAs."""
    from geometric_adv_b200 import _lib
    lib = _lib.load()
    w1, w2 = oracle.nn_distance_grad(a, b, gd1, i1, gd2, i2)
    # kernel 0: shared-memory atomics + per-list sort (1, 2, 4 CTAs per cloud, auto); kernel 1: stable
    # counting sort (one CTA per cloud / 4 CTAs, partner cloud staged in shared memory or gathered)
    combos = [(0, 0, 1), (0, 2, 1), (0, 1, 1), (0, -1, 1), (1, 0, 1), (1, 1, 1), (1, 0, 0), (1, 1, 0)]
    for kernel, split, stage in combos:
        lib.ga_set_tuning(14, kernel)
        lib.ga_set_tuning(9, split)
        lib.ga_set_tuning(13, stage)
        try:
            g1, g2 = ga.nn_distance_grad(t(a), t(b), t(gd1), t(i1), t(gd2), t(i2))
        finally:
            lib.ga_set_tuning(14, 0)
            lib.ga_set_tuning(9, -1)
            lib.ga_set_tuning(13, 1)
        tag = "kernel %d, split %d, stage %d" % (kernel, split, stage)
        assert bits_equal(g1.cpu().numpy(), w1), "grad_xyz1 (%s): %d mismatches" % (tag, int(np.sum(g1.cpu().numpy() != w1)))
        assert bits_equal(g2.cpu().numpy(), w2), "grad_xyz2 (%s): %d mismatches" % (tag, int(np.sum(g2.cpu().numpy() != w2)))


@pytest.mark.parametrize("shape", [(1, 1, 1), (2, 5, 9), (3, 100, 200), (2, 2048, 2048), (1, 2500, 2048),
                                   (1, 4097, 31), (1, 31, 4097), (1, 6000, 5000)])
def test_grad_bit_exact_with_forward_indices(ga, oracle, shape):
    b, n, m = shape
    a, c = cloud(300 + n, (b, n, 3)), cloud(400 + m, (b, m, 3))
    _, i1, _, i2 = oracle.nn_distance(a, c, 0)
    rng = np.random.default_rng(n + m)
    check_bwd(ga, oracle, a, c, rng.standard_normal((b, n)).astype(np.float32), i1,
              rng.standard_normal((b, m)).astype(np.float32), i2)
    mean1 = np.full((b, n), 1.0 / n, np.float32)  # what reduce_mean feeds (adv_ae.py:121)
    mean2 = np.full((b, m), 1.0 / m, np.float32)
    check_bwd(ga, oracle, a, c, mean1, i1, mean2, i2)


def test_grad_heavy_collisions(ga, oracle):
    """Many-to-one index maps: summation order per target is what makes this bit-exact."""
    rng = np.random.default_rng(11)
    b, n, m = 3, 2048, 1500
    a, c = cloud(500, (b, n, 3)), cloud(501, (b, m, 3))
    gd1 = rng.standard_normal((b, n)).astype(np.float32)
    gd2 = rng.standard_normal((b, m)).astype(np.float32)
    i1 = np.zeros((b, n), np.int32)
    i2 = np.full((b, m), n - 1, np.int32)
    check_bwd(ga, oracle, a, c, gd1, i1, gd2, i2)                 # everything collapses onto one point
    i1 = rng.integers(0, 8, (b, n)).astype(np.int32)
    i2 = rng.integers(n - 5, n, (b, m)).astype(np.int32)
    check_bwd(ga, oracle, a, c, gd1, i1, gd2, i2)                 # a handful of hot targets
    i1 = rng.integers(0, m, (b, n)).astype(np.int32)
    i2 = rng.integers(0, n, (b, m)).astype(np.int32)
    check_bwd(ga, oracle, a, c, gd1, i1, gd2, i2)                 # arbitrary (not nearest) indices


def test_config2_grad_digest_and_autograd(ga, oracle):
    a, b = cloud(2, (4, 2048, 3)), cloud(3, (4, 2048, 3))
    gd1 = np.random.default_rng(4).standard_normal((4, 2048)).astype(np.float32)
    gd2 = np.random.default_rng(5).standard_normal((4, 2048)).astype(np.float32)
    x1, x2 = t(a).requires_grad_(True), t(b).requires_grad_(True)
    d1, i1, d2, i2 = ga.nn_distance(x1, x2)
    (torch.sum(d1 * t(gd1)) + torch.sum(d2 * t(gd2))).backward()
    c = digests()["cfg2_b4"]
    assert (sha(x1.grad.cpu().numpy()), sha(x2.grad.cpu().numpy())) == (c["gxyz1"], c["gxyz2"])
    assert not i1.requires_grad and not i2.requires_grad


def test_grad_is_deterministic_run_to_run(ga):
    a, b = t(cloud(1, (16, 2048, 3))), t(cloud(2, (16, 2048, 3)))
    gd = torch.randn(16, 2048, device=DEV)
    _, i1, _, i2 = ga.nn_distance(a, b)
    i1 = (i1 % 7).contiguous()  # force heavy collisions
    first = ga.nn_distance_grad(a, b, gd, i1, gd, i2)
    for _ in range(5):
        again = ga.nn_distance_grad(a, b, gd, i1, gd, i2)
        assert torch.equal(first[0].view(torch.int32), again[0].view(torch.int32))
        assert torch.equal(first[1].view(torch.int32), again[1].view(torch.int32))


def test_cuda_graph_capture(ga):
    """Kernels launch on the caller's stream and allocate nothing: the attack step can be captured."""
    a, b = t(cloud(1, (10, 2048, 3))), t(cloud(2, (10, 2048, 3)))
    gd = torch.full((10, 2048), 1.0 / 2048, device=DEV)
    eager = ga.nn_distance(a, b)
    eg = ga.nn_distance_grad(a, b, gd, eager[1], gd, eager[3])
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        ga.nn_distance(a, b)  # warm-up on the side stream
    torch.cuda.current_stream().wait_stream(s)
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        out = ga.nn_distance(a, b)
        grads = ga.nn_distance_grad(a, b, gd, out[1], gd, out[3])
    for _ in range(3):
        g.replay()
    torch.cuda.synchronize()
    assert all(torch.equal(x, y) for x, y in zip(eager, out))
    assert all(torch.equal(x, y) for x, y in zip(eg, grads))


def test_chamfer_per_cloud(ga, oracle):
    a, b = cloud(1, (7, 500, 3)), cloud(2, (7, 300, 3))
    d1, _, d2, _ = ga.nn_distance(t(a), t(b))
    got = ga.chamfer_per_cloud(d1, d2).cpu().numpy()
    want = oracle.chamfer_per_cloud(d1.cpu().numpy(), d2.cpu().numpy())
    np.testing.assert_allclose(got, want, rtol=2e-5)  # reduce_mean order is unpinned in the reference


def test_custom_ops_pass_opcheck(ga):
    """torch.library.opcheck: schema, fake kernel and autograd registration of the registered ops on real tensors."""
    import torch
    from torch.library import opcheck
    a = torch.from_numpy(cloud(1, (2, 300, 3))).to(DEV).requires_grad_(True)
    b = torch.from_numpy(cloud(2, (2, 257, 3))).to(DEV).requires_grad_(True)
    opcheck(torch.ops.geometric_adv_b200.nn_distance.default, (a, b, 0),
            test_utils=("test_schema", "test_faketensor", "test_autograd_registration"))
    d1, i1, d2, i2 = ga.nn_distance(a, b)
    assert d1.requires_grad and not i1.requires_grad
    (d1.sum() + d2.sum()).backward()
    assert a.grad is not None and b.grad is not None
    opcheck(torch.ops.geometric_adv_b200.knn_point.default, (5, a.detach(), b.detach()),
            test_utils=("test_schema", "test_faketensor"))
    opcheck(torch.ops.geometric_adv_b200.knn_dists.default, (a.detach(), 4), test_utils=("test_schema", "test_faketensor"))
