"""The defenses' host loops on top of the hot path (BASELINE config 5).

Reference: ``defender/get_knn_dists_per_point.py`` (kNN distances per point),
``defender/run_defense_surface.py:187-202`` (threshold on the mean of the first kNN distances,
filter the points, re-run the auto-encoder), ``src/adversary_utils.py:149-178``
(``get_outlier_pc_inlier_pc``, a numpy loop over clouds) and ``src/autoencoder.py:150-168``
(``get_loss_per_pc``: one ``sess.run`` per cloud, batch 1).  Here every step is one batched launch.
"""
import torch

from . import _lib, ops


def knn_dists_mean(pc, num_knn=8, num_knn_for_defense=2):
    """mean over the first `num_knn_for_defense` of the `num_knn` nearest-neighbour distances of
    every point (get_knn_dists_per_point.py:78-81 then run_defense_surface.py:188)."""
    d = ops.knn_dists(pc, num_knn)
    return d[:, :, :num_knn_for_defense].mean(dim=-1)


def get_outlier_pc_inlier_pc(point_clouds, knn_dists, knn_dist_thresh):
    """Batched get_outlier_pc_inlier_pc: returns outlier_pc (B,N,3), outlier_idx (B,N) int16,
    outlier_num (B) int16, inlier_pc (B,N,3) with the reference's padding rules."""
    lib = _lib.load()
    pc = ops._prep(point_clouds, torch.float32, "point_clouds")
    sc = ops._prep(knn_dists, torch.float32, "knn_dists")
    if pc.dim() != 3 or pc.shape[2] != 3 or sc.shape != pc.shape[:2]:
        raise ValueError("expected point_clouds (B,N,3) and knn_dists (B,N)")
    if pc.device.type != "cuda" or sc.device != pc.device:
        raise ValueError("get_outlier_pc_inlier_pc expects CUDA tensors on one device")
    b, n, _ = pc.shape
    opc = torch.empty_like(pc)
    ipc = torch.empty_like(pc)
    oidx = torch.empty((b, n), dtype=torch.int32, device=pc.device)
    onum = torch.empty((b,), dtype=torch.int32, device=pc.device)
    with ops._Guard(pc.device):
        _lib.check(lib.ga_split_by_threshold(b, n, pc.data_ptr(), sc.data_ptr(), float(knn_dist_thresh),
                                             opc.data_ptr(), oidx.data_ptr(), onum.data_ptr(), ipc.data_ptr(),
                                             ops._stream(pc)))
    return opc, oidx.to(torch.int16), onum.to(torch.int16), ipc


def get_loss_per_pc(reconstruct, feed_data, orig_data=None):
    """autoencoder.py:150-168 for the Chamfer loss, batched: reconstruct(feed) once, one
    nn_distance call over all clouds, one fused per-cloud reduction (the reference runs the
    session once per cloud)."""
    target = feed_data if orig_data is None else orig_data
    with torch.no_grad():
        recon = reconstruct(feed_data)
        d1, _, d2, _ = ops.nn_distance(recon.contiguous(), target.contiguous())
        return ops.chamfer_per_cloud(d1, d2)


def surface_defense(pc_input, reconstruct, source_pc, num_knn=8, num_knn_for_defense=2, knn_dist_thresh=0.04):
    """run_defense_surface.py:187-202 for one dist-weight slice: filter off-surface points, then
    the defended reconstruction error against the source cloud.  Returns the defended clouds,
    their reconstruction error per cloud and the outlier bookkeeping."""
    score = knn_dists_mean(pc_input, num_knn, num_knn_for_defense)
    outlier_pc, outlier_idx, outlier_num, pc_defended = get_outlier_pc_inlier_pc(pc_input, score, knn_dist_thresh)
    err = get_loss_per_pc(reconstruct, pc_defended, source_pc)
    return pc_defended, err, (outlier_pc, outlier_idx, outlier_num)
