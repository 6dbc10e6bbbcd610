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


_BYTEWISE = (torch.int16, torch.bool, torch.uint16)  # element types NCCL's all-gather does not take


def _rows_as_bytes(block):
    """(rows, ...) of any element type -> (rows, bytes per row) uint8, no copy for contiguous input."""
    rowbytes = block.element_size()
    for d in block.shape[1:]:
        rowbytes *= int(d)
    return block.contiguous().view(torch.uint8).reshape(block.shape[0], rowbytes)


def _rows_from_bytes(flat, dtype, tail):
    """Inverse of _rows_as_bytes for a (rows, bytes per row) uint8 tensor."""
    return flat.contiguous().view(dtype).reshape((flat.shape[0],) + tuple(tail))


def all_gather_rows(block, total_rows, group=None):
    """All-gather row blocks produced with shard_range into the full (total_rows, ...) tensor."""
    world, rank = _world(group)
    if world == 1:
        return block
    # NCCL moves bytes of a handful of element types; int16 (the nn_idx stage file's type) and bool are not among
    # them: gather such blocks as raw bytes and view the result back
    if block.is_cuda and block.dtype in _BYTEWISE:
        out = all_gather_rows(_rows_as_bytes(block), total_rows, group)
        return _rows_from_bytes(out, block.dtype, tuple(block.shape[1:]))
    sizes = [shard_range(total_rows, world, r) for r in range(world)]
    maxrows = max(hi - lo for lo, hi in sizes)
    if total_rows % world == 0 and block.is_cuda:  # equal blocks: one NCCL all-gather straight into the result
        out = torch.empty((total_rows,) + tuple(block.shape[1:]), dtype=block.dtype, device=block.device)
        dist.all_gather_into_tensor(out, block.contiguous(), group=group)
        return out
    pad = torch.zeros((maxrows,) + tuple(block.shape[1:]), dtype=block.dtype, device=block.device)
    pad[: block.shape[0]] = block
    if block.is_cuda:  # unequal blocks: ONE all-gather of blocks padded to the largest, then the pads drop out
        gathered = torch.empty((world * maxrows,) + tuple(block.shape[1:]), dtype=block.dtype, device=block.device)
        dist.all_gather_into_tensor(gathered, pad, group=group)
        parts = gathered.view((world, maxrows) + tuple(block.shape[1:]))
        return torch.cat([parts[r, : hi - lo] for r, (lo, hi) in enumerate(sizes)], dim=0)
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


def sort_dist_mat_gpu(dist_rows, slice_idx):
    """sort_dist_mat (prepare_indices_for_attack.py:167-180) on the GPU for a block of rows of the
    Chamfer matrix: dist_rows (rows,S) CUDA float32 -> nn_idx (rows,S) int16 (class-local indices,
    stable order).  Bit-identical to ``sort_dist_mat`` above on the same rows."""
    from . import _lib
    lib = _lib.load()
    if dist_rows.device.type != "cuda" or dist_rows.dtype != torch.float32 or dist_rows.dim() != 2:
        raise ValueError("sort_dist_mat_gpu expects a (rows,S) CUDA float32 tensor")
    dist_rows = dist_rows.contiguous()
    rows, s = dist_rows.shape
    sl = [int(x) for x in slice_idx]
    if sl[0] != 0 or sl[-1] != s or any(b < a for a, b in zip(sl[:-1], sl[1:])):
        raise ValueError("slice_idx must run from 0 to S in non-decreasing steps")
    nclass = len(sl) - 1
    max_class = max([b - a for a, b in zip(sl[:-1], sl[1:])] + [0])
    sl_dev = torch.tensor(sl, dtype=torch.int32, device=dist_rows.device)
    nn_idx = torch.empty((rows, s), dtype=torch.int16, device=dist_rows.device)
    with torch.cuda.device(dist_rows.device):
        _lib.check(lib.ga_sort_dist_mat(s, rows, dist_rows.data_ptr(), nclass, sl_dev.data_ptr(), max_class,
                                        nn_idx.data_ptr(), torch.cuda.current_stream().cuda_stream))
    return nn_idx


def symmetrize_rows(directed, row0, rows):
    """Rows [row0, row0+rows) of CD = D + D^T from the gathered (S,S) matrix of directed terms."""
    from . import _lib
    lib = _lib.load()
    directed = directed.contiguous()
    s = directed.shape[0]
    out = torch.empty((rows, s), dtype=torch.float32, device=directed.device)
    with torch.cuda.device(directed.device):
        _lib.check(lib.ga_symmetrize_rows(s, row0, rows, directed.data_ptr(), out.data_ptr(),
                                          torch.cuda.current_stream().cuda_stream))
    return out


def prepare_indices(clouds, slice_idx, group=None, mode=None, timings=None):
    """get_chamfer_nn of attacker/prepare_indices_for_attack.py:104-164 for the whole test set, sharded by row
    blocks: returns (chamfer_dist_mat (S,S) float32, chamfer_nn_idx (S,S) int16), both complete on every rank.

    rank r: directed terms D[i -> j] for its rows i (the tensor-core all-pairs kernel) -> ONE all-gather of
    the row blocks (the only exchange the problem has; 16 MB for 2,000 shapes) -> its rows of CD = D + D^T ->
    per-class stable argsort of its rows (ga_sort_dist_mat) -> all-gather of the finished rows.
    `timings`, if a dict, receives CUDA-event milliseconds per phase."""
    from . import ops
    s = clouds.shape[0]
    world, rank = _world(group)
    lo, hi = shard_range(s, world, rank)
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(5)] if timings is not None else None

    def mark(i):
        if ev is not None:
            ev[i].record()

    mark(0)
    block = ops.chamfer_all_pairs(clouds, lo, hi - lo, mode=mode, directed=True)
    mark(1)
    d = all_gather_rows(block, s, group)
    mark(2)
    cd_rows = symmetrize_rows(d, lo, hi - lo)
    nn_rows = sort_dist_mat_gpu(cd_rows, slice_idx)
    mark(3)
    cd = all_gather_rows(cd_rows, s, group)
    nn_idx = all_gather_rows(nn_rows, s, group)
    mark(4)
    if ev is not None:
        torch.cuda.synchronize(clouds.device)
        for name, i in (("directed_ms", 0), ("gather_directed_ms", 1), ("symmetrize_sort_ms", 2),
                        ("gather_results_ms", 3)):
            timings[name] = ev[i].elapsed_time(ev[i + 1])
        timings["total_ms"] = ev[0].elapsed_time(ev[4])
    return cd, nn_idx


def nearest_targets_gpu(nn_idx, slice_idx, num_targets):
    """The attack's view of nn_idx for EVERY (source shape, target class) at once (src/adversary_utils.py:51-63):
    out (S, nclass, num_targets) int16, the `num_targets` nearest instances of each target class (class-local
    indices); for the shape's own class entry 0 (the shape itself) is skipped; -1 where a class is too small."""
    s = nn_idx.shape[0]
    sl = [int(x) for x in slice_idx]
    nclass = len(sl) - 1
    dev = nn_idx.device
    starts = torch.tensor(sl[:-1], device=dev)
    sizes = torch.tensor([b - a for a, b in zip(sl[:-1], sl[1:])], device=dev)
    row_class = torch.bucketize(torch.arange(s, device=dev), torch.tensor(sl[1:], device=dev), right=True)
    skip = (row_class[:, None] == torch.arange(nclass, device=dev)[None, :]).long()          # (S, nclass)
    k = torch.arange(num_targets, device=dev)[None, None, :] + skip[:, :, None]              # (S, nclass, T)
    ok = k < sizes[None, :, None]
    cols = (starts[None, :, None] + k).clamp_(max=max(s - 1, 0))
    out = torch.gather(nn_idx, 1, cols.reshape(s, -1)).reshape(s, nclass, num_targets)
    return torch.where(ok, out, torch.full_like(out, -1))


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


def save_sel_idx_rand(data_path, slice_idx, num_instance_per_class=100, seed=55, file_name_parts=("test", "set", "13l")):
    """get_rand_idx (prepare_indices_for_attack.py:66-86): per class, the first `num_instance_per_class`
    entries of a seeded permutation of its instances, int16, -1 padded, under the reference's file name
    sel_idx_rand_<n>_<parts>.npy.  Uses numpy's legacy global generator exactly as the reference does
    (np.random.seed(55) before every class), so the file is byte-identical to the reference's."""
    import os
    nclass = len(slice_idx) - 1
    sel_idx = -1 * np.ones([nclass, num_instance_per_class], dtype=np.int16)
    state = np.random.get_state()
    try:
        for i in range(nclass):
            np.random.seed(seed)
            num_examples = int(slice_idx[i + 1] - slice_idx[i])
            perm = np.arange(num_examples)
            np.random.shuffle(perm)
            num_instances = min(num_instance_per_class, num_examples)
            sel_idx[i, :num_instances] = perm[:num_instance_per_class]
    finally:
        np.random.set_state(state)
    path = os.path.join(data_path, "_".join(["sel_idx", "rand", "%d" % num_instance_per_class] + list(file_name_parts)) + ".npy")
    np.save(path, sel_idx)
    return path


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
