"""Randomised parity sweep of the DEFAULT dispatch (whatever kernel the library picks for a shape):
forward, gradient and kNN against the CPU oracle on seeded random shapes that cluster around the
dispatch thresholds (255/256/257 points, 511..513, 2047..2049, 4095..4097 and the batch sizes where
the tensor-core forward and the split gradient kernels take over).  Bit-exact, both arithmetics."""
import numpy as np
import pytest
import torch

from util import bits_equal

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
EDGES = [1, 2, 31, 32, 33, 63, 64, 65, 127, 128, 129, 255, 256, 257, 511, 512, 513, 1000, 2025, 2047, 2048, 2049, 4095,
         4096, 4097]


def t(a):
    return torch.from_numpy(np.ascontiguousarray(a)).to(DEV)


def shapes(seed, count, budget):
    """(b, n, m) with b*n*m <= budget; sizes drawn from the edge list or uniformly."""
    rng = np.random.default_rng(seed)
    out = []
    while len(out) < count:
        n = int(rng.choice(EDGES)) if rng.random() < 0.7 else int(rng.integers(1, 3000))
        m = int(rng.choice(EDGES)) if rng.random() < 0.7 else int(rng.integers(1, 3000))
        bmax = max(1, budget // (n * m))
        b = int(min(bmax, rng.choice([1, 2, 3, 9, 10, 18, 19, 37, 38, 50])))
        out.append((b, n, m))
    return out


def clouds(rng, b, n, kind):
    x = rng.random((b, n, 3), dtype=np.float32) - np.float32(0.5)
    if kind == 1:    # far from the origin: wide windows
        x = x + np.float32(37.0)
    elif kind == 2:  # coarse lattice: exact ties everywhere
        x = np.round(x * 6).astype(np.float32) / np.float32(6)
    elif kind == 3:  # tiny scale
        x = x * np.float32(1e-12)
    return np.ascontiguousarray(x.astype(np.float32))


@pytest.mark.parametrize("seed", [11, 12, 13, 14])
def test_forward_and_gradient_random_shapes(ga, oracle, seed):
    rng = np.random.default_rng(1000 + seed)
    for (b, n, m) in shapes(seed, 14, 12_000_000):
        kind = int(rng.integers(0, 4))
        mode = int(rng.integers(0, 2))
        a, c = clouds(rng, b, n, kind), clouds(rng, b, m, kind)
        want = oracle.nn_distance(a, c, mode)
        got = [x.cpu().numpy() for x in ga.nn_distance(t(a), t(c), mode)]
        for nme, g, w in zip(["dist1", "idx1", "dist2", "idx2"], got, want):
            assert bits_equal(g, w), "%s differs: shape %s kind %d mode %d, %d mismatches" % (
                nme, (b, n, m), kind, mode, int(np.sum(g != w)))
        gd1 = rng.standard_normal((b, n)).astype(np.float32)
        gd2 = rng.standard_normal((b, m)).astype(np.float32)
        w1, w2 = oracle.nn_distance_grad(a, c, gd1, want[1], gd2, want[3])
        g1, g2 = ga.nn_distance_grad(t(a), t(c), t(gd1), t(want[1]), t(gd2), t(want[3]))
        assert bits_equal(g1.cpu().numpy(), w1) and bits_equal(g2.cpu().numpy(), w2), "gradient differs: %s kind %d" % (
            (b, n, m), kind)


@pytest.mark.parametrize("seed", [21, 22])
def test_knn_random_shapes(ga, oracle, seed):
    rng = np.random.default_rng(2000 + seed)
    for (b, n, m) in shapes(seed, 10, 3_000_000):
        kind = int(rng.integers(0, 3))
        k = int(min(n, rng.choice([1, 2, 5, 11, 16, 17, 32, 33])))
        data, query = clouds(rng, b, n, kind), clouds(rng, b, m, kind)
        wv, wi = oracle.knn_point(k, data, query)
        gv, gi = ga.knn_point(k, t(data), t(query))
        assert bits_equal(gv.cpu().numpy(), wv), "kNN values differ: %s k=%d kind %d" % ((b, n, m), k, kind)
        assert np.array_equal(gi.cpu().numpy(), wi), "kNN indices differ: %s k=%d kind %d" % ((b, n, m), k, kind)
