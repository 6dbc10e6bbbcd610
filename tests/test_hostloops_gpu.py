"""GPU tests of the callers either side of the hot path: the all-pairs Chamfer matrix
(prepare_indices_for_attack.py), the attack iteration (adv_ae.py) and sharded drivers."""
import numpy as np
import pytest
import torch

from util import bits_equal, cloud

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def t(a):
    return torch.from_numpy(np.ascontiguousarray(a)).to(DEV)


def oracle_cd_matrix(oracle, clouds):
    s = clouds.shape[0]
    out = np.empty((s, s), np.float32)
    for i in range(s):  # row i: target = clouds[i], sources = all clouds (prepare_indices_for_attack.py:123-139)
        d1, _, d2, _ = oracle.nn_distance(clouds, np.repeat(clouds[i][None], s, axis=0), 0)
        out[i] = oracle.chamfer_per_cloud(d1, d2)
    return out


@pytest.mark.parametrize("s,n", [(9, 300), (5, 2048), (12, 777), (3, 1)])
def test_all_pairs_matches_oracle(ga, oracle, s, n):
    c = cloud(3, (s, n, 3))
    got = ga.chamfer_all_pairs(t(c)).cpu().numpy()
    want = oracle_cd_matrix(oracle, c)
    # The per-point distances are bit-exact (same search as nn_distance); only the ORDER of the
    # fp32 mean over n points differs from the oracle's (the reference's reduce_mean order is
    # unpinned).  Reordering a sum of n fp32 terms moves it by ~sqrt(n)*2^-24 relative.
    np.testing.assert_allclose(got, want, rtol=2e-5, atol=1e-9)
    assert np.array_equal(got, got.T), "CD matrix must be exactly symmetric"
    assert not got.diagonal().any(), "self distance must be exactly 0"
    for i in range(s):
        assert np.array_equal(np.argsort(got[i], kind="stable"), np.argsort(want[i], kind="stable"))


def test_all_pairs_row_blocks_and_directed(ga, oracle):
    from geometric_adv_b200 import sharding
    c = t(cloud(5, (11, 640, 3)))
    full = ga.chamfer_all_pairs(c)
    blk = ga.chamfer_all_pairs(c, 3, 5)
    assert torch.equal(blk, full[3:8]), "a row block must equal the rows of the full matrix bit for bit"
    d = ga.chamfer_all_pairs(c, directed=True)
    assert torch.equal(d + d.t(), full), "CD = D + D^T exactly"
    assert torch.equal(sharding.all_pairs_chamfer(c), full)
    # the directed term is mean(dist1) of the op itself
    d1, _, _, _ = ga.nn_distance(c[[2]].contiguous(), c[[7]].contiguous())
    np.testing.assert_allclose(d[2, 7].item(), d1.mean().item(), rtol=1e-6)
    with pytest.raises(ValueError):
        ga.chamfer_all_pairs(c, 8, 5)


def test_all_pairs_is_deterministic(ga):
    c = t(cloud(6, (40, 2048, 3)))
    a = ga.chamfer_all_pairs(c)
    b = ga.chamfer_all_pairs(c)
    assert torch.equal(a, b)


def test_attack_iteration_graph_equals_eager():
    from geometric_adv_b200.attack import GeometricAttack, PointNetAE
    torch.manual_seed(0)
    ae = PointNetAE(512)
    src, tgt = t(cloud(1, (4, 512, 3))), t(cloud(2, (4, 512, 3)))
    res = []
    for graph in (False, True):
        atk = GeometricAttack(ae, 4, 512, num_iterations=12, num_iterations_thresh=8, use_cuda_graph=graph)
        res.append(atk.run(src, tgt))
    for a, b in zip(res[0], res[1]):
        assert torch.allclose(a, b, rtol=1e-4, atol=1e-6)
    metrics, adv, recon = res[1]
    assert metrics.shape == (4, 5) and adv.shape == (4, 512, 3)
    assert bool(torch.all(metrics[:, 4] < 1e9)), "best-so-far must have been collected after the threshold"
    assert bool(torch.all((adv - src).abs().max() > 0))


def test_attack_reduces_the_target_reconstruction_error():
    from geometric_adv_b200.attack import GeometricAttack, PointNetAE, chamfer_per_pc
    torch.manual_seed(0)
    ae = PointNetAE(1024).to(DEV).eval()
    src, tgt = t(cloud(3, (5, 1024, 3))), t(cloud(4, (5, 1024, 3)) * np.float32(0.5))
    with torch.no_grad():
        before, _ = chamfer_per_pc(ae(src)[0], tgt)
    atk = GeometricAttack(ae, 5, 1024, num_iterations=60, num_iterations_thresh=40)
    metrics, _, _ = atk.run(src, tgt)
    assert bool(torch.all(metrics[:, 4] < before)), (metrics[:, 4], before)


def test_attack_pairs_shards_equal_the_whole():
    """Pairs are independent (frozen BatchNorm, Adam slots reset, per-pair noise streams): attacking
    two contiguous shards with the same batch shape gives bit-identical results to one run."""
    from geometric_adv_b200.attack import PointNetAE, attack_pair_range, attack_pairs
    torch.manual_seed(0)
    ae = PointNetAE(256)
    src, tgt = torch.from_numpy(cloud(5, (8, 256, 3))), torch.from_numpy(cloud(6, (8, 256, 3)))
    kw = dict(batch_size=2, num_iterations=6, num_iterations_thresh=3)
    m, a = attack_pairs(ae, src, tgt, **kw)
    assert m.shape == (8, 5) and a.shape == (8, 256, 3)
    m0, a0 = attack_pair_range(ae, src, tgt, 0, 4, **kw)   # what rank 0 of 2 would compute
    m1, a1 = attack_pair_range(ae, src, tgt, 4, 8, **kw)   # rank 1
    assert torch.equal(torch.cat([m0, m1]), m) and torch.equal(torch.cat([a0, a1]), a)
    # a different batch size changes the cuBLAS/cuDNN kernels of the AE, not the attack itself
    m3, _ = attack_pairs(ae, src, tgt, batch_size=3, num_iterations=6, num_iterations_thresh=3)
    assert torch.allclose(m3[:, 4], m[:, 4], rtol=1e-3)


def test_outlier_inlier_split_matches_the_reference_loop(ga, oracle):
    """get_outlier_pc_inlier_pc (adversary_utils.py:149-178), batched on the GPU: bit-equal."""
    from geometric_adv_b200 import defense
    rng = np.random.default_rng(0)
    pc = cloud(7, (6, 700, 3))
    score = rng.random((6, 700), dtype=np.float32) * np.float32(0.08)
    score[1] = 1.0            # every point an outlier
    score[2] = 0.0            # none
    score[3, 5] = np.nan      # belongs to neither part
    score[4, :] = 0.04        # exactly at the threshold: inlier (<=)
    got = defense.get_outlier_pc_inlier_pc(t(pc), t(score), 0.04)
    want = oracle.split_by_threshold(pc, score, 0.04)
    assert bits_equal(got[0].cpu().numpy(), want[0]) and bits_equal(got[3].cpu().numpy(), want[3])
    assert np.array_equal(got[1].cpu().numpy(), want[1].astype(np.int16))
    assert np.array_equal(got[2].cpu().numpy(), want[2].astype(np.int16))
    assert got[1].dtype == torch.int16 and got[2].dtype == torch.int16


def test_surface_defense_pipeline(ga, oracle):
    """kNN distances -> mean of the first two -> threshold -> filtered clouds -> batched loss per cloud."""
    from geometric_adv_b200 import defense
    from geometric_adv_b200.attack import PointNetAE
    torch.manual_seed(0)
    ae = PointNetAE(512).to(DEV).eval()
    src = cloud(8, (5, 512, 3))
    adv = src.copy()
    adv[:, :20] += np.float32(0.3)  # a few points pushed off the surface
    score = defense.knn_dists_mean(t(adv), 8, 2)
    want_score = oracle.knn_dists(adv, 8)[:, :, :2]
    np.testing.assert_allclose(score.cpu().numpy(), (want_score[..., 0] + want_score[..., 1]) / 2, rtol=1e-6)
    recon = lambda x: ae(x)[0]
    defended, err, (opc, oidx, onum) = defense.surface_defense(t(adv), recon, t(src), knn_dist_thresh=0.15)
    assert int(onum.min()) >= 1 and defended.shape == (5, 512, 3) and err.shape == (5,)
    # batched get_loss_per_pc == the reference's one-cloud-at-a-time loop
    with torch.no_grad():
        one = torch.stack([defense.get_loss_per_pc(recon, defended[i:i + 1], t(src)[i:i + 1])[0] for i in range(5)])
    assert torch.allclose(err, one, rtol=1e-5)


def test_foldingnet_call_sites(ga):
    """foldingnet.py:209-238 (dense Chamfer, N != M) and prepare_graph.py:45-73 (KDTree k=16 + cov)."""
    from geometric_adv_b200 import callsites
    x = t(cloud(9, (3, 2025, 3))).requires_grad_(True)   # FoldingNet output: 45^2 points
    y = t(cloud(10, (3, 2048, 3)))
    loss = callsites.foldingnet_chamfer_distance(x, y)
    loss.backward()
    xd = x.detach()
    d = ((xd[:, None, :, :] - y[:, :, None, :]) ** 2).sum(3)            # the reference's dense form
    want = d.min(1)[0].mean() + d.min(2)[0].mean()
    assert torch.allclose(loss, want, rtol=1e-5)
    assert x.grad is not None and bool(torch.isfinite(x.grad).all())
    pc = cloud(11, (2, 700, 3))
    idx, dist, cov = callsites.foldingnet_knn_graph(t(pc), 16)
    from sklearn.neighbors import KDTree
    for i in range(2):
        nd, ni = KDTree(pc[i], leaf_size=30).query(pc[i], k=17, return_distance=True)
        assert np.array_equal(idx[i].cpu().numpy(), ni[:, 1:])
        np.testing.assert_allclose(dist[i].cpu().numpy(), nd, rtol=1e-5, atol=1e-7)
        c = np.stack([np.cov(pc[i][ni[j, 1:]].T).reshape(-1) for j in range(0, 700, 50)])
        np.testing.assert_allclose(cov[i].cpu().numpy()[::50], c, rtol=1e-4, atol=1e-7)
    edges = callsites.knn_edges(idx)
    assert edges[0].shape[0] == 2 and edges[0].shape[1] >= 700 * 16


def test_sort_dist_mat_gpu_equals_numpy_stable():
    """ga_sort_dist_mat vs np.argsort(kind="stable") per class block, incl. exact ties, -0, inf and NaN."""
    from geometric_adv_b200 import sharding
    rng = np.random.default_rng(0)
    s = 301
    slice_idx = [0, 1, 1, 40, 173, 300, 301]  # an empty class, single-element classes, odd sizes
    dm = rng.random((s, s)).astype(np.float32)
    dm[rng.random((s, s)) < 0.2] = np.float32(0.25)            # many exact ties
    dm[rng.random((s, s)) < 0.01] = np.float32(-0.0)
    dm[rng.random((s, s)) < 0.01] = np.float32(0.0)
    dm[5, 7] = np.inf
    dm[9, 100:120] = np.nan
    want = np.empty((s, s), np.int16)
    for c in range(len(slice_idx) - 1):
        a, b = slice_idx[c], slice_idx[c + 1]
        want[:, a:b] = np.argsort(dm[:, a:b], axis=1, kind="stable")
    got = sharding.sort_dist_mat_gpu(t(dm), slice_idx)
    assert got.dtype == torch.int16 and np.array_equal(got.cpu().numpy(), want)
    blk = sharding.sort_dist_mat_gpu(t(dm[17:60]), slice_idx)
    assert np.array_equal(blk.cpu().numpy(), want[17:60])
    big = rng.random((3, 5000)).astype(np.float32)              # one class of 5,000 shapes
    got = sharding.sort_dist_mat_gpu(t(big), [0, 5000]).cpu().numpy()
    assert np.array_equal(got, np.argsort(big, axis=1, kind="stable").astype(np.int16))


def test_prepare_indices_pipeline(ga, tmp_path):
    """Directed rows -> symmetrise -> per-class sort on the GPU == the host-side stage-file path."""
    from geometric_adv_b200 import sharding
    c = t(cloud(8, (45, 400, 3)))
    slice_idx = [0, 10, 11, 30, 45]
    cd, nn = sharding.prepare_indices(c, slice_idx, timings={})
    full = ga.chamfer_all_pairs(c)
    assert torch.equal(cd, full)
    want = sharding.sort_dist_mat(full, slice_idx)
    assert np.array_equal(nn.cpu().numpy(), want)
    d = ga.chamfer_all_pairs(c, directed=True)
    assert torch.equal(sharding.symmetrize_rows(d, 7, 20), full[7:27])
    top = sharding.nearest_targets_gpu(nn, slice_idx, 5).cpu().numpy()
    assert top.shape == (45, 4, 5)
    for row, (sc, si) in [(3, (0, 3)), (10, (1, 0)), (29, (2, 18)), (44, (3, 14))]:
        for tc in range(4):
            w = sharding.nearest_targets(want, slice_idx, sc, si, tc, 5)
            assert top[row, tc, :len(w)].tolist() == w.tolist() and np.all(top[row, tc, len(w):] == -1)
    p1, p2 = sharding.save_chamfer_nn_files(str(tmp_path), cd, slice_idx)
    assert np.array_equal(np.load(p2), nn.cpu().numpy())


def test_chamfer_loss_terms_fused_op(ga):
    """One search + one reduction launch: cd bit-equal to chamfer_per_cloud(nn_distance), max bit-equal to amax,
    gradient bit-equal to autograd through nn_distance + mean."""
    a = t(cloud(31, (6, 700, 3))).requires_grad_(True)
    b = t(cloud(32, (6, 900, 3))).requires_grad_(True)
    cd, mx, i1, i2 = ga.chamfer_loss_terms(a, b)
    d1, j1, d2, j2 = ga.nn_distance(a, b)
    assert torch.equal(i1, j1) and torch.equal(i2, j2)
    assert torch.equal(cd, ga.chamfer_per_cloud(d1.detach(), d2.detach()))
    assert torch.equal(mx, d1.amax(dim=1))
    ref = d1.mean(dim=1) + d2.mean(dim=1)
    assert torch.allclose(cd, ref, rtol=1e-6)
    w = t(np.linspace(0.5, 2.0, 6).astype(np.float32))
    ga1, gb1 = torch.autograd.grad((cd * w).sum(), (a, b))
    ga2, gb2 = torch.autograd.grad((ref * w).sum(), (a, b))
    assert torch.equal(ga1, ga2) and torch.equal(gb1, gb2)
    nanrow = a.detach().clone()
    nanrow[2, 5, 0] = float("nan")
    _, mx2, _, _ = ga.chamfer_loss_terms(nanrow, b.detach())
    assert bool(torch.isnan(mx2[2])) and not bool(torch.isnan(mx2[[0, 1, 3, 4, 5]]).any())


def test_attack_single_forward_equals_the_reference_order():
    """The metrics of an update are what the next iteration's forward computes: the single-forward iteration
    (one AE forward, two searches) gives the results of the reference's order (two forwards, four searches)."""
    from geometric_adv_b200.attack import GeometricAttack, PointNetAE
    torch.manual_seed(0)
    ae = PointNetAE(512)
    src, tgt = t(cloud(1, (4, 512, 3))), t(cloud(2, (4, 512, 3)))
    res = []
    for single in (False, True):
        atk = GeometricAttack(ae, 4, 512, num_iterations=14, num_iterations_thresh=9, use_cuda_graph=False,
                              single_forward=single)
        res.append(atk.run(src, tgt))
    for a, b in zip(res[0], res[1]):
        assert torch.allclose(a, b, rtol=1e-5, atol=1e-7)
    res2 = []
    for fused in (False, True):
        atk = GeometricAttack(ae, 4, 512, num_iterations=14, num_iterations_thresh=9, fused_loss=fused)
        res2.append(atk.run(src, tgt))
    for a, b in zip(res2[0], res2[1]):
        assert torch.allclose(a, b, rtol=1e-4, atol=1e-6)
