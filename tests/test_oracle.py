"""CPU tests: the oracle is pinned to the reference.

 * golden fixtures generated from the reference's own CPU kernels
   (tests/golden/make_golden.py) -- always;
 * the reference's CPU kernels themselves (oracle/_ref/libga_ref.so) -- when built;
 * the known answer of the reference's selection_sort.cpp;
 * numpy's statement of the reference's kNN fallback path.
"""
import numpy as np
import pytest

from util import bits_equal, cloud, digests, golden, sha


def test_nn_distance_matches_golden_unit_shape(oracle):
    g = golden("nnd_unit_4x100x200.npz")
    d1, i1, d2, i2 = oracle.nn_distance(g["xyz1"], g["xyz2"], 0)
    assert bits_equal(d1, g["dist1"]) and bits_equal(d2, g["dist2"])
    assert np.array_equal(i1, g["idx1"]) and np.array_equal(i2, g["idx2"])
    g1, g2 = oracle.nn_distance_grad(g["xyz1"], g["xyz2"], g["gd1"], i1, g["gd2"], i2)
    assert bits_equal(g1, g["gxyz1"]) and bits_equal(g2, g["gxyz2"])


def test_nn_distance_matches_golden_ties(oracle):
    g = golden("nnd_ties_3x257x131.npz")
    d1, i1, d2, i2 = oracle.nn_distance(g["xyz1"], g["xyz2"], 0)
    assert bits_equal(d1, g["dist1"]) and bits_equal(d2, g["dist2"])
    assert np.array_equal(i1, g["idx1"]) and np.array_equal(i2, g["idx2"])
    g1, g2 = oracle.nn_distance_grad(g["xyz1"], g["xyz2"], g["gd1"], i1, g["gd2"], i2)
    assert bits_equal(g1, g["gxyz1"]) and bits_equal(g2, g["gxyz2"])


def test_unit_test_py_tolerances_vs_chamfer_python(oracle):
    """unit_test.py:23-33: mean squared dist error < 1e-8 and identical indices."""
    g = golden("nnd_unit_4x100x200.npz")
    p = golden("chamfer_python_4x100x200.npz")
    d1, i1, d2, i2 = oracle.nn_distance(g["xyz1"], g["xyz2"], 0)
    assert np.mean((d1 - p["dist1"]) ** 2) + np.mean((d2 - p["dist2"]) ** 2) < 1e-8
    assert np.array_equal(i1, p["idx1"]) and np.array_equal(i2, p["idx2"])


def test_config1_digests(oracle):
    dg = digests()
    a, b = cloud(0, (1, 2048, 3)), cloud(1, (1, 2048, 3))
    d1, i1, d2, i2 = oracle.nn_distance(a, b, 0)
    c = dg["cfg1"]
    assert (sha(d1), sha(i1), sha(d2), sha(i2)) == (c["dist1"], c["idx1"], c["dist2"], c["idx2"])
    gd = np.full((1, 2048), 1.0 / 2048, np.float32)
    g1, g2 = oracle.nn_distance_grad(a, b, gd, i1, gd, i2)
    assert (sha(g1), sha(g2)) == (c["gxyz1"], c["gxyz2"])
    d1, i1, d2, i2 = oracle.nn_distance(a, a, 0)
    c = dg["cfg1_dup"]
    assert (sha(d1), sha(i1), sha(d2), sha(i2)) == (c["dist1"], c["idx1"], c["dist2"], c["idx2"])
    assert not d1.any() and np.array_equal(i1[0], np.arange(2048))


def test_selection_sort_known_answer(oracle):
    """external/grouping/test/selection_sort.cpp:65-93, output captured from the reference binary."""
    k = digests()["selection_sort"]
    dist = np.array(k["dist"], np.float32).reshape(k["b"], k["m"], k["n"])
    outi, out = oracle.selection_sort(dist, k["k"])
    assert outi.reshape(-1).tolist() == k["idx"]
    assert out.reshape(-1).tolist() == k["val"]


def test_selection_sort_is_unstable_on_ties(oracle):
    """The displaced head moves behind later equals."""
    dist = np.array([[[5, 5, 1, 7]]], np.float32)
    outi, out = oracle.selection_sort(dist, 2)
    assert outi[0, 0, :2].tolist() == [2, 1] and out[0, 0, :2].tolist() == [1, 5]


def test_knn_dists_equals_numpy_fallback_path(oracle):
    """defender/get_knn_dists_per_point.py:125-137 + src/general_utils.py:94-106, bitwise."""
    pc = cloud(4, (3, 300, 3))
    assert bits_equal(oracle.knn_dists(pc, 10), oracle.knn_dists_numpy(pc, 10))


def test_knn_point_semantics(oracle):
    xyz1, xyz2 = cloud(5, (2, 97, 3)), cloud(6, (2, 33, 3))
    val, idx = oracle.knn_point(7, xyz1, xyz2)
    d = ((xyz1[:, None, :, :] - xyz2[:, :, None, :]) ** 2)
    d = (d[..., 0] + d[..., 1]) + d[..., 2]
    assert bits_equal(val, np.sort(d, axis=-1)[..., :7])
    assert bits_equal(np.take_along_axis(d, idx.astype(np.int64), axis=-1), val)
    g = oracle.group_point(xyz1, idx)
    assert np.array_equal(g, xyz1[np.arange(2)[:, None, None], idx])


def test_mode1_is_the_fma_contraction(oracle):
    a, b = cloud(7, (1, 64, 3)), cloud(8, (1, 80, 3))
    d1 = oracle.nn_distance(a, b, 1)[0]
    x = (b[0][None, :, :] - a[0][:, None, :]).astype(np.float64)
    f32 = np.float32
    yy = (x[..., 1].astype(f32) * x[..., 1].astype(f32)).astype(np.float64)
    inner = (x[..., 0] * x[..., 0] + yy).astype(f32).astype(np.float64)  # fma: one rounding
    d = (x[..., 2] * x[..., 2] + inner).astype(f32)
    assert bits_equal(d1[0], d.min(axis=1))


@pytest.mark.skipif(not __import__("oracle.oracle", fromlist=["x"]).have_ref(), reason="oracle/_ref not built")
class TestAgainstReferenceBuild:
    def test_forward_and_grad_bitwise(self, oracle):
        for seed, shape1, shape2 in [(30, (2, 513, 3), (2, 700, 3)), (31, (1, 1, 3), (1, 1, 3)),
                                     (32, (3, 5, 3), (3, 1000, 3))]:
            a, b = cloud(seed, shape1), cloud(seed + 100, shape2)
            r = oracle.ref_nn_distance(a, b)
            o = oracle.nn_distance(a, b, 0)
            assert all(bits_equal(x, y) for x, y in zip(r, o))
            gd1 = np.random.default_rng(seed).standard_normal(shape1[:2]).astype(np.float32)
            gd2 = np.random.default_rng(seed + 1).standard_normal(shape2[:2]).astype(np.float32)
            rg = oracle.ref_nn_distance_grad(a, b, gd1, r[1], gd2, r[3])
            og = oracle.nn_distance_grad(a, b, gd1, r[1], gd2, r[3])
            assert all(bits_equal(x, y) for x, y in zip(rg, og))

    def test_nan_and_inf_semantics(self, oracle):
        a, b = cloud(40, (1, 16, 3)), cloud(41, (1, 24, 3))
        b[0, 0, 1] = np.nan   # NaN at target 0 seeds best with NaN for every query
        b[0, 5, 0] = np.inf
        a[0, 3, 2] = np.nan
        r = oracle.ref_nn_distance(a, b)
        o = oracle.nn_distance(a, b, 0)
        assert all(bits_equal(x, y) for x, y in zip(r, o))

    def test_reference_argument_messages(self, oracle):
        with pytest.raises(ValueError, match="only accepts 3d point set xyz1"):
            oracle.ref_nn_distance_shaped(np.zeros((2, 5, 2), np.float32), np.zeros((2, 5, 3), np.float32))
        with pytest.raises(ValueError, match="same batch size"):
            oracle.ref_nn_distance_shaped(np.zeros((2, 5, 3), np.float32), np.zeros((3, 5, 3), np.float32))

    def test_threaded_harness_is_identical(self, oracle):
        a, b = cloud(50, (5, 300, 3)), cloud(51, (5, 280, 3))
        assert all(bits_equal(x, y) for x, y in zip(oracle.ref_nn_distance(a, b), oracle.ref_nn_distance(a, b, 4)))
