"""GPU parity of the grouping ops: knn_point, select_top_k, group_point, knn_dists."""
import numpy as np
import pytest
import torch

from util import bits_equal, cloud

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def t(a):
    return torch.from_numpy(np.ascontiguousarray(a)).to(DEV)


def check_knn(ga, oracle, k, xyz1, xyz2):
    val, idx = ga.knn_point(k, t(xyz1), t(xyz2))
    wv, wi = oracle.knn_point(k, xyz1, xyz2)
    val, idx = val.cpu().numpy(), idx.cpu().numpy()
    assert bits_equal(val, wv), "val: %d mismatches" % int(np.sum(val != wv))
    assert np.array_equal(idx, wi), "idx: %d mismatches (k=%d, %s vs %s)" % (
        int(np.sum(idx != wi)), k, xyz1.shape, xyz2.shape)


@pytest.mark.parametrize("k", [1, 2, 8, 11, 16, 17, 32])
@pytest.mark.parametrize("shape", [(2, 512, 128), (1, 2048, 300), (1, 100, 33), (1, 2500, 64), (1, 5000, 40)])
def test_knn_point_bit_exact(ga, oracle, k, shape):
    b, n, m = shape
    check_knn(ga, oracle, k, cloud(600 + n, (b, n, 3)), cloud(700 + m, (b, m, 3)))


def test_knn_self_query_defense_shape(ga, oracle):
    """get_knn_dists_per_point.py:78: knn_point(k+1, pc, pc) with k+1 = 11 on 2048-point clouds."""
    pc = cloud(4, (3, 2048, 3))
    check_knn(ga, oracle, 11, pc, pc)
    check_knn(ga, oracle, 9, pc, pc)  # the script's default num_knn = 8


@pytest.mark.parametrize("k", [1, 3, 11, 16, 20])
def test_knn_tie_semantics_of_the_selection_sort(ga, oracle, k):
    """Coordinates on a coarse grid and duplicated points: exact ties everywhere.  Indices must
    follow the reference's unstable selection sort (tf_grouping_g.cu:100-121), not a stable sort."""
    rng = np.random.default_rng(k)
    data = (rng.integers(0, 3, (2, 400, 3)).astype(np.float32) * np.float32(0.5))
    qry = (rng.integers(0, 3, (2, 90, 3)).astype(np.float32) * np.float32(0.5))
    check_knn(ga, oracle, k, data, qry)
    pc = cloud(5, (1, 600, 3))
    pc[0, 300:] = pc[0, :300]          # every point duplicated once
    check_knn(ga, oracle, k, pc, pc)
    z = np.zeros((1, 64, 3), np.float32)
    check_knn(ga, oracle, k, z, z)     # all distances equal


def test_knn_k_equals_n_and_small_sets(ga, oracle):
    check_knn(ga, oracle, 5, cloud(1, (2, 5, 3)), cloud(2, (2, 9, 3)))
    check_knn(ga, oracle, 1, cloud(3, (1, 1, 3)), cloud(4, (1, 3, 3)))
    check_knn(ga, oracle, 12, cloud(5, (1, 12, 3)), cloud(6, (1, 40, 3)))


def test_knn_large_k_generic_path(ga, oracle):
    """tf_grouping.py:86-90 self-test shape: 512 data points, 128 queries, nsample 64."""
    check_knn(ga, oracle, 64, cloud(7, (2, 512, 3), 0.0, 1.0), cloud(8, (2, 128, 3), 0.0, 1.0))
    check_knn(ga, oracle, 40, cloud(9, (1, 100, 3)), cloud(10, (1, 50, 3)))


def test_knn_argument_errors(ga):
    x = t(cloud(1, (1, 10, 3)))
    with pytest.raises(ValueError, match="positive k"):
        ga.knn_point(0, x, x)
    with pytest.raises(ValueError, match="exceeds the data set size"):
        ga.knn_point(11, x, x)


def test_knn_non_finite(ga, oracle):
    data, qry = cloud(11, (1, 200, 3)), cloud(12, (1, 50, 3))
    data[0, 150, 1] = np.nan   # NaN beyond the first k positions: never selected
    data[0, 160, 0] = np.inf
    check_knn(ga, oracle, 6, data, qry)
    data[0, 2, 2] = np.nan     # NaN inside the first k positions: "selected" by the reference
    check_knn(ga, oracle, 6, data, qry)
    qry[0, 7, 0] = np.nan
    check_knn(ga, oracle, 6, data, qry)


def test_knn_dists_equals_numpy_fallback_path(ga, oracle):
    """(B,N,k) distances, bitwise equal to defender/get_knn_dists_per_point.py:125-137 (numpy path)
    and to the op graph :78-81 restated in the oracle."""
    pc = cloud(4, (4, 2048, 3))
    got = ga.knn_dists(t(pc), 10).cpu().numpy()
    assert bits_equal(got, oracle.knn_dists(pc, 10))
    assert bits_equal(got[:2], oracle.knn_dists_numpy(pc[:2], 10))
    host = ga.knn_dists(torch.from_numpy(pc), 10).numpy()
    assert bits_equal(got, host)
    assert np.all(got >= 0)


def test_knn_dists_full_size_properties(ga):
    """B=500 (config 5): ascending per point, self-distance dropped, shard == whole."""
    pc = cloud(4, (500, 2048, 3))
    out = ga.knn_dists(t(pc), 10)
    assert out.shape == (500, 2048, 10)
    assert bool(torch.all(out[..., 1:] >= out[..., :-1])) and bool(torch.all(out > 0))
    part = ga.knn_dists(t(pc[62:125]), 10)
    assert torch.equal(part, out[62:125])


def test_script_graph_equals_fused_kernel(ga):
    """The reference graph, op by op, through the shim: knn_point -> idx[:,:,1:] -> group_point ->
    deltas -> sqrt(sum(sq)) equals knn_dists."""
    pc = t(cloud(13, (2, 1024, 3)))
    k = 8
    _, idx = ga.knn_point(k + 1, pc, pc)
    grouped = ga.group_point(pc, idx[:, :, 1:].contiguous())
    deltas = grouped - pc.unsqueeze(2)
    sq = deltas * deltas
    d = torch.sqrt((sq[..., 0] + sq[..., 1]) + sq[..., 2])
    assert torch.equal(d, ga.knn_dists(pc, k))


def test_select_top_k_dense_entry(ga, oracle):
    rng = np.random.default_rng(3)
    dist = rng.random((3, 17, 130), dtype=np.float32)
    dist[0, :, ::3] = 0.25  # ties
    dist[1, 2, 5] = np.nan
    dist[1, 3, 0] = np.nan  # NaN head stays selected
    for k in [1, 4, 16, 130]:
        outi, out = ga.select_top_k(k, t(dist))
        wi, wo = oracle.selection_sort(dist, k)
        assert np.array_equal(outi.cpu().numpy(), wi) and bits_equal(out.cpu().numpy(), wo), k


def test_group_point(ga, oracle):
    rng = np.random.default_rng(4)
    pts = rng.random((3, 50, 7), dtype=np.float32)
    idx = rng.integers(0, 50, (3, 20, 6)).astype(np.int32)
    got = ga.group_point(t(pts), t(idx)).cpu().numpy()
    assert np.array_equal(got, oracle.group_point(pts, idx))
    p = t(pts).requires_grad_(True)
    ga.group_point(p, t(idx)).sum().backward()  # tf_grouping_op_test.py checks this gradient
    cnt = np.zeros((3, 50), np.float32)
    for b in range(3):
        np.add.at(cnt[b], idx[b].reshape(-1), 1.0)
    np.testing.assert_allclose(p.grad.cpu().numpy(), np.repeat(cnt[..., None], 7, axis=2), rtol=1e-6)


# ---- knn_mma_kernel (tensor-core scans; default for large batches, forced here with tuning key 1 = 6) ----
@pytest.fixture
def knn_mma(ga):
    from geometric_adv_b200 import _lib
    lib = _lib.load()
    lib.ga_set_tuning(1, 6)
    yield
    lib.ga_set_tuning(1, 0)


@pytest.mark.parametrize("k", [1, 2, 3, 4, 8, 11])
@pytest.mark.parametrize("shape", [(2, 512, 128), (1, 2048, 300), (2, 2048, 2048), (1, 1000, 1100), (3, 513, 64), (1, 1537, 70)])
def test_knn_mma_bit_exact(ga, oracle, knn_mma, k, shape):
    b, n, m = shape
    check_knn(ga, oracle, k, cloud(800 + n, (b, n, 3)), cloud(900 + m, (b, m, 3)))


def test_knn_mma_self_query_and_dists(ga, oracle, knn_mma):
    pc = cloud(4, (3, 2048, 3))
    check_knn(ga, oracle, 11, pc, pc)
    check_knn(ga, oracle, 9, pc, pc)
    got = ga.knn_dists(t(pc), 10).cpu().numpy()
    assert bits_equal(got, oracle.knn_dists(pc, 10))


@pytest.mark.parametrize("k", [1, 3, 11])
def test_knn_mma_ties_duplicates_and_non_finite(ga, oracle, knn_mma, k):
    rng = np.random.default_rng(k)
    data = (rng.integers(0, 3, (2, 700, 3)).astype(np.float32) * np.float32(0.5))   # grid: ties everywhere
    qry = (rng.integers(0, 3, (2, 90, 3)).astype(np.float32) * np.float32(0.5))
    check_knn(ga, oracle, k, data, qry)
    pc = cloud(5, (1, 1200, 3))
    pc[0, 600:] = pc[0, :600]          # every point duplicated once
    check_knn(ga, oracle, k, pc, pc)
    z = np.zeros((1, 512, 3), np.float32)
    check_knn(ga, oracle, k, z, z)     # all distances equal
    data, qry = cloud(11, (1, 800, 3)), cloud(12, (1, 50, 3))
    data[0, 150, 1] = np.nan
    data[0, 160, 0] = np.inf
    check_knn(ga, oracle, k, data, qry)
    data[0, 0, 2] = np.nan             # NaN inside the first k positions
    check_knn(ga, oracle, k, data, qry)
    qry[0, 7, 0] = np.nan
    check_knn(ga, oracle, k, data, qry)


def test_knn_mma_dynamic_range_and_clusters(ga, oracle, knn_mma):
    """Far from the origin the window covers every tile (exact path); tight clusters fill the queues."""
    pc = cloud(21, (1, 1024, 3)) + np.float32(100.0)
    check_knn(ga, oracle, 11, pc, pc)
    rng = np.random.default_rng(22)
    cl = (rng.normal(0, 1e-3, (1, 2048, 3)) + rng.integers(0, 4, (1, 2048, 1))).astype(np.float32)
    check_knn(ga, oracle, 11, cl, cl)
    check_knn(ga, oracle, 11, cloud(23, (1, 2048, 3), -1e-3, 1e-3), cloud(24, (1, 64, 3), -1e-3, 1e-3))


def test_knn_mma_equals_fp32_kernel_at_full_size(ga, knn_mma):
    """B=500 (config 5): the tensor-core kernel against the fp32-filter kernel, distances and indices."""
    from geometric_adv_b200 import _lib
    lib = _lib.load()
    pc = t(cloud(4, (64, 2048, 3)))
    v6, i6 = ga.knn_point(11, pc, pc)
    lib.ga_set_tuning(1, 3)
    v3, i3 = ga.knn_point(11, pc, pc)
    lib.ga_set_tuning(1, 6)
    assert torch.equal(v6, v3) and torch.equal(i6, i3)


# ---- knn_slab_kernel: ga_knn_dists with the scans pruned to an x-slab (values only) -----------------------------
def _slab_clouds():
    rng = np.random.default_rng(11)
    out = {}
    out["uniform_2048"] = cloud(21, (6, 2048, 3))
    out["ragged_2047"] = cloud(22, (3, 2047, 3))
    out["n1000"] = cloud(23, (3, 1000, 3))
    out["n512"] = cloud(24, (2, 512, 3))
    out["n1500_offset"] = (cloud(25, (2, 1500, 3)) + np.float32(7.25)).astype(np.float32)
    blobs = (rng.standard_normal((3, 2048, 3)) * 0.01).astype(np.float32)
    blobs += rng.integers(0, 4, (3, 2048, 1)).astype(np.float32) * np.float32(0.5)      # four tight clusters
    out["clusters"] = blobs
    dup = cloud(26, (2, 2048, 3))
    dup[:, 1024:] = dup[:, :1024]                                                          # every point twice
    out["duplicates"] = dup
    grid = np.stack(np.meshgrid(np.arange(16), np.arange(16), np.arange(8), indexing="ij"), -1).reshape(1, 2048, 3)
    out["lattice_ties"] = (grid * np.float32(0.125)).astype(np.float32)                  # exact ties everywhere
    plane = cloud(27, (2, 2048, 3))
    plane[..., 0] = np.float32(0.3)                                                        # all x equal: one bin
    out["plane_x_const"] = plane
    line = np.zeros((1, 2048, 3), np.float32)
    line[0, :, 0] = np.linspace(-1, 1, 2048, dtype=np.float32)                             # a line along x
    out["line"] = line
    same = np.full((1, 600, 3), 0.25, np.float32)                                          # one point 600 times
    out["all_equal"] = same
    tiny = (cloud(28, (2, 2048, 3)) * np.float32(1e-12)).astype(np.float32)
    out["tiny_coordinates"] = tiny
    huge = (cloud(29, (2, 2048, 3)) * np.float32(1e12)).astype(np.float32)
    out["huge_coordinates"] = huge
    surf = cloud(30, (3, 2048, 3))
    surf /= np.linalg.norm(surf, axis=-1, keepdims=True).astype(np.float32)              # points on a sphere
    out["sphere_surface"] = surf.astype(np.float32)
    return out


@pytest.mark.parametrize("k", [1, 4, 10])
@pytest.mark.parametrize("name", list(_slab_clouds().keys()))
def test_knn_slab_equals_full_scan_kernel(ga, name, k):
    """The slab-pruned kernel against knn_kernel (key 28 = 0): the same bits, whatever the cloud looks like."""
    from geometric_adv_b200 import _lib
    lib = _lib.load()
    pc = _slab_clouds()[name]
    lib.ga_set_tuning(28, 2)
    got = ga.knn_dists(t(pc), k)
    assert _lib.load().ga_last_kernel().decode() == "knn_slab_kernel"
    lib.ga_set_tuning(28, 0)
    try:
        want = ga.knn_dists(t(pc), k)
        assert _lib.load().ga_last_kernel().decode() == "knn_kernel"
    finally:
        lib.ga_set_tuning(28, 1)
    assert bits_equal(got.cpu().numpy(), want.cpu().numpy()), (name, k, int((got != want).sum()))


def test_knn_slab_equals_oracle_and_handles_non_finite(ga, oracle):
    from geometric_adv_b200 import _lib
    _lib.load().ga_set_tuning(28, 2)
    pc = cloud(31, (5, 2048, 3))
    got = ga.knn_dists(t(pc), 10).cpu().numpy()
    assert _lib.load().ga_last_kernel().decode() == "knn_slab_kernel"
    assert bits_equal(got, oracle.knn_dists(pc, 10))
    assert bits_equal(got[:1], oracle.knn_dists_numpy(pc[:1], 10))
    # a cloud with a NaN / inf point is handed to the full-scan body inside the same launch; its neighbours in the
    # batch are not affected
    from geometric_adv_b200 import _lib
    lib = _lib.load()
    bad = pc.copy()
    bad[1, 77, 1] = np.nan
    bad[3, 5, 0] = np.inf
    lib.ga_set_tuning(28, 2)
    got = ga.knn_dists(t(bad), 10).cpu().numpy()
    lib.ga_set_tuning(28, 0)
    try:
        want = ga.knn_dists(t(bad), 10).cpu().numpy()
    finally:
        lib.ga_set_tuning(28, 1)
    assert bits_equal(got, want)
    assert bits_equal(got[[0, 2, 4]], oracle.knn_dists(pc[[0, 2, 4]], 10))


def test_knn_slab_full_size_config5(ga, oracle):
    """B=500 x 2048, k=10: bitwise equal to the full-scan kernel, 32 clouds against the oracle."""
    from geometric_adv_b200 import _lib
    lib = _lib.load()
    pc = cloud(4, (500, 2048, 3))
    got = ga.knn_dists(t(pc), 10)
    assert _lib.load().ga_last_kernel().decode() == "knn_slab_kernel"
    lib.ga_set_tuning(28, 0)
    try:
        want = ga.knn_dists(t(pc), 10)
    finally:
        lib.ga_set_tuning(28, 1)
    assert torch.equal(got, want)
    sel = np.arange(0, 500, 16)
    assert bits_equal(got[sel].cpu().numpy(), oracle.knn_dists(pc[sel], 10))


@pytest.mark.parametrize("seed", list(range(24)))
def test_knn_slab_randomised_against_full_scan(ga, seed):
    """Random cloud sizes, k, anisotropic scales, offsets, cluster mixtures, repeated x values and coincident points:
    knn_slab_kernel == knn_kernel bit for bit (the window bound, the bin mapping and the overflow service all have to
    hold for any geometry)."""
    from geometric_adv_b200 import _lib
    lib = _lib.load()
    rng = np.random.default_rng(1000 + seed)
    b = int(rng.integers(1, 5))
    n = int(rng.integers(512, 2049))
    k = int(rng.integers(1, 11))
    kind = seed % 6
    pc = rng.random((b, n, 3), dtype=np.float32) - np.float32(0.5)
    if kind == 1:    # anisotropic box, far from the origin
        pc = pc * np.array([10.0 ** rng.uniform(-3, 3), 10.0 ** rng.uniform(-3, 3), 10.0 ** rng.uniform(-3, 3)], np.float32)
        pc = pc + np.float32(10.0 ** rng.uniform(-2, 2))
    elif kind == 2:  # mixture of tight clusters of very different sizes
        c = rng.integers(0, 5, (b, n, 1))
        pc = (rng.standard_normal((b, n, 3)) * (10.0 ** (-1.0 - c))).astype(np.float32) + (c * 0.37).astype(np.float32)
    elif kind == 3:  # x quantised to a few hundred values (full bins), y / z free
        pc[..., 0] = np.round(pc[..., 0] * 150.0) / 150.0
    elif kind == 4:  # a quarter of the points coincide with other points
        idx = rng.integers(0, n, (b, n // 4))
        for bi in range(b):
            pc[bi, : n // 4] = pc[bi, idx[bi]]
    elif kind == 5:  # a surface: points on a paraboloid sheet plus a few outliers far away
        pc[..., 2] = pc[..., 0] ** 2 + pc[..., 1] ** 2
        pc[:, :3] += np.float32(40.0)
    pc = np.ascontiguousarray(pc.astype(np.float32))
    lib.ga_set_tuning(28, 2)
    try:
        got = ga.knn_dists(t(pc), k)
        assert lib.ga_last_kernel().decode() == "knn_slab_kernel"
        lib.ga_set_tuning(28, 0)
        want = ga.knn_dists(t(pc), k)
    finally:
        lib.ga_set_tuning(28, 1)
    assert bits_equal(got.cpu().numpy(), want.cpu().numpy()), (seed, kind, b, n, k, int((got != want).sum()))
