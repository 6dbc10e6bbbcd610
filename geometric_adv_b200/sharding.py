"""Data-parallel sharding of the hot path over 1..8 GPUs (one process per GPU).

Every workload of the reference is embarrassingly parallel over clouds / cloud pairs
(SURVEY.md 8e), so there is no data-path collective; the only communication is one
all-gather of small result shards:

* ``all_pairs_chamfer``  - rank r computes the DIRECTED terms D[i -> j] for its block of
  rows i (all j); one all-gather of the (rows, S) blocks; every rank forms
  CD = D + D^T.  That is half the work of computing both directions per block, paid for
  with the one exchange the problem really has
  (attacker/prepare_indices_for_attack.py:104-152 splits by column blocks over 44 processes
  and exchanges through a .npy file on disk).
* ``knn_dists_sharded``  - clouds split over ranks, result shards all-gathered
  (defender/get_knn_dists_per_point.py:116-119 chunks by 100 clouds in one process).
* ``shard_range``        - the contiguous split used everywhere (also for the attack's
  (source, target) pairs, which need no exchange at all until the final gather).

The compute callables are injectable so the host logic is testable on CPU with gloo.
"""
import numpy as np
import torch
import torch.distributed as dist


def shard_range(total, world, rank):
    """Contiguous block [lo, hi) of `total` items for `rank` of `world`; the first
    total % world ranks get one extra item (500 clouds over 8 ranks -> 63,63,63,63,62,...)."""
    if world <= 0 or not (0 <= rank < world):
        raise ValueError("bad world/rank %r/%r" % (world, rank))
    base, rem = divmod(int(total), int(world))
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def _world(group):
    if dist.is_available() and dist.is_initialized():
        return dist.get_world_size(group), dist.get_rank(group)
    return 1, 0


def all_gather_rows(block, total_rows, group=None):
    """All-gather row blocks produced with shard_range into the full (total_rows, ...) tensor."""
    world, rank = _world(group)
    if world == 1:
        return block
    sizes = [shard_range(total_rows, world, r) for r in range(world)]
    maxrows = max(hi - lo for lo, hi in sizes)
    pad = torch.zeros((maxrows,) + tuple(block.shape[1:]), dtype=block.dtype, device=block.device)
    pad[: block.shape[0]] = block
    parts = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(parts, pad, group=group)
    return torch.cat([p[: hi - lo] for p, (lo, hi) in zip(parts, sizes)], dim=0)


def all_pairs_chamfer(clouds, group=None, mode=None, directed_fn=None):
    """Full (S,S) Chamfer matrix, chamfer_dist_mat[i, j] = CD(source=clouds[j], target=clouds[i])
    (prepare_indices_for_attack.py:139,151), identical on every rank.

    clouds: (S,N,3) replicated on every rank.  directed_fn(clouds, row0, rows) -> (rows,S)
    tensor of D[row0+r -> j]; defaults to the CUDA kernel."""
    if directed_fn is None:
        from . import ops

        def directed_fn(c, row0, rows):
            return ops.chamfer_all_pairs(c, row0, rows, mode=mode, directed=True)
    s = clouds.shape[0]
    world, rank = _world(group)
    lo, hi = shard_range(s, world, rank)
    block = directed_fn(clouds, lo, hi - lo)
    d = all_gather_rows(block, s, group)
    return d + d.t()  # commutative add: the result is exactly symmetric


def sort_dist_mat(dist_mat, slice_idx):
    """Per (source class, target class) block argsort along the row, int16
    (prepare_indices_for_attack.py:167-180).  numpy's default argsort leaves the order of
    exact ties unspecified; this one is stable."""
    dm = dist_mat.detach().cpu().numpy() if isinstance(dist_mat, torch.Tensor) else np.asarray(dist_mat)
    nn_idx = -1 * np.ones(dm.shape, dtype=np.int16)
    nc = len(slice_idx) - 1
    for i in range(nc):
        for j in range(nc):
            blk = dm[slice_idx[i]:slice_idx[i + 1], slice_idx[j]:slice_idx[j + 1]]
            nn_idx[slice_idx[i]:slice_idx[i + 1], slice_idx[j]:slice_idx[j + 1]] = np.argsort(
                blk, axis=1, kind="stable").astype(np.int16)
    assert nn_idx.min() >= 0, "the nn_idx matrix was not filled correctly"
    return nn_idx


def nearest_targets(nn_idx, slice_idx, source_class, source_instance, target_class, num_targets):
    """The consumer's view (src/adversary_utils.py:51-63): the `num_targets` nearest instances of
    `target_class` for one source shape (indices local to the class; same-class lookups drop
    the first entry, the shape itself)."""
    row = slice_idx[source_class] + source_instance
    cols = nn_idx[row, slice_idx[target_class]:slice_idx[target_class + 1]]
    if source_class == target_class:
        cols = cols[1:]
    return cols[:num_targets]


def knn_dists_sharded(pc, k, group=None, knn_fn=None):
    """Per-point kNN distances (B,N,k) for a batch of clouds replicated on every rank:
    each rank computes its contiguous share, shards are all-gathered."""
    if knn_fn is None:
        from . import ops
        knn_fn = ops.knn_dists
    world, rank = _world(group)
    lo, hi = shard_range(pc.shape[0], world, rank)
    out = knn_fn(pc[lo:hi].contiguous(), k)
    return all_gather_rows(out, pc.shape[0], group)


def shard_pairs(sources, targets, group=None):
    """Contiguous share of the attack's (source, target) pairs for this rank
    (25 x 10 = 250 pairs in BASELINE config 3).  Pairs are independent (frozen BatchNorm, only the
    perturbation is trained, src/adv_ae.py:105), so nothing is exchanged until the final gather."""
    world, rank = _world(group)
    lo, hi = shard_range(sources.shape[0], world, rank)
    return sources[lo:hi], targets[lo:hi], (lo, hi)


def save_chamfer_nn_files(data_path, chamfer_dist_mat, slice_idx, file_name_parts=("test", "set", "13l")):
    """Write the two stage files the downstream reference scripts read, with the reference's names,
    dtypes and shapes (prepare_indices_for_attack.py:84-86,144-164; consumer
    src/adversary_utils.py:51-63): chamfer_dist_mat_complete_<parts>.npy float32 (S,S) and
    chamfer_nn_idx_complete_<parts>.npy int16 (S,S)."""
    import os
    dm = chamfer_dist_mat.detach().cpu().numpy() if isinstance(chamfer_dist_mat, torch.Tensor) else np.asarray(
        chamfer_dist_mat)
    dm = dm.astype(np.float32)
    assert np.isfinite(dm).all() and dm.min() >= 0, "the chamfer_dist_mat matrix was not filled correctly"
    tail = "_".join(file_name_parts)
    p1 = os.path.join(data_path, "chamfer_dist_mat_complete_" + tail + ".npy")
    p2 = os.path.join(data_path, "chamfer_nn_idx_complete_" + tail + ".npy")
    np.save(p1, dm)
    np.save(p2, sort_dist_mat(dm, slice_idx))
    return p1, p2
