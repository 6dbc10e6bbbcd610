"""CPU tests of the multi-GPU host logic: world_size-2 gloo, compute injected from the oracle."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from util import cloud


def test_shard_range_covers_everything():
    from geometric_adv_b200.sharding import shard_range
    for total in [0, 1, 7, 250, 500, 2000]:
        for world in [1, 2, 3, 4, 8]:
            blocks = [shard_range(total, world, r) for r in range(world)]
            assert blocks[0][0] == 0 and blocks[-1][1] == total
            assert all(a[1] == b[0] for a, b in zip(blocks, blocks[1:]))
            sizes = [hi - lo for lo, hi in blocks]
            assert max(sizes) - min(sizes) <= 1
    assert [hi - lo for lo, hi in [shard_range(500, 8, r) for r in range(8)]] == [63, 63, 63, 63, 62, 62, 62, 62]
    with pytest.raises(ValueError):
        shard_range(10, 2, 2)


def _directed_oracle(clouds, row0, rows):
    """D[row0+r -> j] from the CPU oracle (test infrastructure)."""
    from oracle import oracle as O
    c = clouds.numpy()
    s, n, _ = c.shape
    out = np.empty((rows, s), np.float32)
    for r in range(rows):
        src = np.repeat(c[row0 + r][None], s, axis=0)
        d1, _, _, _ = O.nn_distance(src, c, 0)
        out[r] = d1.astype(np.float64).sum(axis=1) / n  # any fixed order; compared with a tolerance
    return torch.from_numpy(out)


def _knn_oracle(pc, k):
    from oracle import oracle as O
    return torch.from_numpy(O.knn_dists(pc.numpy(), k))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from geometric_adv_b200 import sharding
        clouds = torch.from_numpy(cloud(3, (7, 64, 3)))  # odd count: uneven blocks
        cd = sharding.all_pairs_chamfer(clouds, directed_fn=_directed_oracle)
        pc = torch.from_numpy(cloud(4, (5, 48, 3)))
        kd = sharding.knn_dists_sharded(pc, 4, knn_fn=_knn_oracle)
        src, tgt, (lo, hi) = sharding.shard_pairs(torch.arange(9), torch.arange(9) + 100)
        ret[rank] = (cd.numpy(), kd.numpy(), (lo, hi), src.tolist())
    finally:
        dist.destroy_process_group()


def test_two_rank_gloo_matches_single_process():
    world = 2
    port = _free_port()
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(world, port, ret), nprocs=world, join=True)
    from geometric_adv_b200 import sharding
    from oracle import oracle as O
    clouds = cloud(3, (7, 64, 3))
    single = sharding.all_pairs_chamfer(torch.from_numpy(clouds), directed_fn=_directed_oracle).numpy()
    kd_single = O.knn_dists(cloud(4, (5, 48, 3)), 4)
    for r in range(world):
        cd, kd, (lo, hi), src = ret[r]
        assert np.array_equal(cd, single), "rank %d: sharded matrix differs from the single-process one" % r
        assert np.array_equal(kd, kd_single)
        assert src == list(range(lo, hi))
    assert ret[0][2] == (0, 5) and ret[1][2] == (5, 9)
    # matrix semantics: symmetric, zero diagonal, entries = mean(d1)+mean(d2) of the oracle
    assert np.array_equal(single, single.T) and not single.diagonal().any()
    d1, _, d2, _ = O.nn_distance(clouds[[2]], clouds[[5]], 0)  # source 2, target 5 -> [5, 2]
    np.testing.assert_allclose(single[5, 2], O.chamfer_per_cloud(d1, d2)[0], rtol=2e-5)


def test_sort_dist_mat_and_consumer_view():
    from geometric_adv_b200.sharding import nearest_targets, sort_dist_mat
    rng = np.random.default_rng(0)
    slice_idx = [0, 3, 7, 12]
    dm = rng.random((12, 12)).astype(np.float32)
    dm = dm + dm.T
    np.fill_diagonal(dm, 0)
    nn = sort_dist_mat(dm, slice_idx)
    assert nn.dtype == np.int16 and nn.shape == dm.shape
    for i in range(3):
        for j in range(3):
            blk = dm[slice_idx[i]:slice_idx[i + 1], slice_idx[j]:slice_idx[j + 1]]
            assert np.array_equal(nn[slice_idx[i]:slice_idx[i + 1], slice_idx[j]:slice_idx[j + 1]],
                                  np.argsort(blk, axis=1, kind="stable"))
    # same class: the first entry is the shape itself (distance 0) and is dropped by the consumer
    assert nn[slice_idx[1] + 2, slice_idx[1]:slice_idx[2]][0] == 2
    t = nearest_targets(nn, slice_idx, 1, 2, 1, 2)
    assert 2 not in t.tolist() and len(t) == 2
    t = nearest_targets(nn, slice_idx, 0, 1, 2, 3)
    blk = dm[1, slice_idx[2]:slice_idx[3]]
    assert t.tolist() == np.argsort(blk, kind="stable")[:3].tolist()


def test_stage_files_have_the_reference_names_and_dtypes(tmp_path):
    from geometric_adv_b200.sharding import save_chamfer_nn_files
    rng = np.random.default_rng(1)
    dm = rng.random((9, 9)).astype(np.float32)
    dm = dm + dm.T
    np.fill_diagonal(dm, 0)
    p1, p2 = save_chamfer_nn_files(str(tmp_path), torch.from_numpy(dm), [0, 4, 9])
    assert p1.endswith("chamfer_dist_mat_complete_test_set_13l.npy")
    assert p2.endswith("chamfer_nn_idx_complete_test_set_13l.npy")
    a, b = np.load(p1), np.load(p2)
    assert a.dtype == np.float32 and a.shape == (9, 9) and np.array_equal(a, dm)
    assert b.dtype == np.int16 and b.shape == (9, 9) and b.min() >= 0


def test_sel_idx_rand_file(tmp_path):
    """get_rand_idx (prepare_indices_for_attack.py:66-86): seeded permutation prefix per class, -1 padded."""
    from geometric_adv_b200.sharding import save_sel_idx_rand
    slice_idx = [0, 5, 130, 137]
    before = np.random.get_state()[1][:4].copy()
    path = save_sel_idx_rand(str(tmp_path), slice_idx, num_instance_per_class=100)
    assert np.array_equal(np.random.get_state()[1][:4], before), "the caller's numpy RNG state must be left alone"
    assert path.endswith("sel_idx_rand_100_test_set_13l.npy")
    sel = np.load(path)
    assert sel.dtype == np.int16 and sel.shape == (3, 100)
    for i, n in enumerate([5, 125, 7]):
        np.random.seed(55)
        perm = np.arange(n)
        np.random.shuffle(perm)
        k = min(n, 100)
        assert np.array_equal(sel[i, :k], perm[:k]) and np.all(sel[i, k:] == -1)
        assert len(set(sel[i, :k].tolist())) == k


def test_rows_travel_as_bytes_for_types_nccl_refuses():
    """all_gather_rows moves int16 / bool blocks as bytes on CUDA (NCCL has no such element type: the all-pairs leg
    failed under torchrun before); the view round-trips bit for bit, for empty and multi-dimensional blocks too."""
    import torch
    from geometric_adv_b200 import sharding
    g = torch.Generator().manual_seed(0)
    for shape in [(7, 5), (0, 5), (3,), (4, 2, 3)]:
        a = torch.randint(-32768, 32767, shape, generator=g, dtype=torch.int32).to(torch.int16)
        flat = sharding._rows_as_bytes(a)
        assert flat.dtype == torch.uint8 and flat.shape[0] == shape[0] and flat.dim() == 2
        assert torch.equal(sharding._rows_from_bytes(flat, torch.int16, shape[1:]), a)
    b = torch.rand(6, 9, generator=g) > 0.5
    assert torch.equal(sharding._rows_from_bytes(sharding._rows_as_bytes(b), torch.bool, (9,)), b)
    assert torch.int16 in sharding._BYTEWISE
