"""PyTorch shim over libga_b200.so that keeps the reference's operator signatures.

Reference interfaces mirrored here (paths below the reference tree):

* ``nn_distance(xyz1, xyz2) -> dist1, idx1, dist2, idx2``
  external/structural_losses/tf_nndistance.py:15-26, gradient :35-41
* ``nn_distance_grad(xyz1, xyz2, grad_dist1, idx1, grad_dist2, idx2)``
  the NnDistanceGrad op, tf_nndistance.cpp:10-18
* ``chamfer_3DDist()(a, b) -> dist1, dist2, idx1, idx2``
  transfer/atlasnet/auxiliary/ChamferDistancePytorch/chamfer3D/dist_chamfer_3D.py:26-74
* ``knn_point(k, xyz1, xyz2) -> val, idx``       external/grouping/tf_grouping.py:48-75
* ``select_top_k(k, dist) -> idx, dist_out``      tf_grouping.py:22-31
* ``group_point(points, idx)``                    tf_grouping.py:33-47

CUDA tensors run on the caller's current stream (CUDA-graph capturable); CPU
tensors go through the library's ``*_host`` entry points, which stage through the
GPU -- the same kernels, not a CPU implementation.  torch supplies memory and
streams only; all arithmetic happens in the library.
"""
import torch

from . import _lib
from ._lib import GA_MODE_CPU_EXACT, GA_MODE_GPU_REF

__all__ = [
    "nn_distance", "nn_distance_grad", "chamfer_3DDist", "chamfer_3DFunction", "knn_point", "select_top_k",
    "group_point", "knn_dists", "chamfer_per_cloud", "chamfer_loss_terms", "chamfer_all_pairs", "set_default_mode", "set_pruning",
    "launch_count",
    "GA_MODE_CPU_EXACT", "GA_MODE_GPU_REF",
]

_default_mode = GA_MODE_CPU_EXACT
_pruning = False


def set_pruning(enabled):
    """Opt in to the spatially pruned forward (ga_nn_distance_fwd_ws: Morton-ordered clouds, tiles
    skipped by box bounds; same bits) for CUDA clouds of 256..2048 points.  Off by default: on
    uniformly random 2048-point clouds a warp's 128 queries still need 60 % of the tiles and the
    sort costs 27 us, so it is slower at B=50 (103 vs 78 us) and ~11 % faster at B=512
    (profiles/r01_tune.json); clouds with spatial structure and larger batches are where it pays."""
    global _pruning
    _pruning = bool(enabled)


def set_default_mode(mode):
    """GA_MODE_CPU_EXACT (bit-identical to the reference CPU kernel, default) or
    GA_MODE_GPU_REF (bit-identical to the reference CUDA kernel's FMA contraction)."""
    global _default_mode
    if mode not in (GA_MODE_CPU_EXACT, GA_MODE_GPU_REF):
        raise ValueError("unknown mode %r" % (mode,))
    _default_mode = mode


def launch_count():
    return int(_lib.load().ga_launch_count())


def _stream(t):
    return torch.cuda.current_stream(t.device).cuda_stream


def _prep(t, dtype, name):
    if not isinstance(t, torch.Tensor):
        raise TypeError("%s must be a torch.Tensor" % name)
    if t.dtype != dtype:
        # the reference ops are registered for float32 / int32 only (tf_nndistance.cpp:4-9)
        raise TypeError("%s must be %s, got %s" % (name, dtype, t.dtype))
    return t.contiguous()  # dist_chamfer_3D.py:72-73 does the same


def _same_device(*ts):
    d = ts[0].device
    for t in ts[1:]:
        if t.device != d:
            raise ValueError("all tensors must live on the same device (%s vs %s)" % (d, t.device))
    return d


class _Guard:
    """Make the tensor's device current for the duration of a launch."""

    def __init__(self, dev):
        self.dev = dev
        self.prev = None

    def __enter__(self):
        if self.dev.type == "cuda" and torch.cuda.current_device() != self.dev.index:
            self.prev = torch.cuda.current_device()
            torch.cuda.set_device(self.dev)

    def __exit__(self, *a):
        if self.prev is not None:
            torch.cuda.set_device(self.prev)


# --------------------------------------------------------------------------- nn_distance
def _nn_distance_fwd(xyz1, xyz2, mode):
    lib = _lib.load()
    _lib.check(lib.ga_check_nn_distance(xyz1.dim(), _lib.dims(xyz1.shape), xyz2.dim(), _lib.dims(xyz2.shape)))
    xyz1 = _prep(xyz1, torch.float32, "xyz1")
    xyz2 = _prep(xyz2, torch.float32, "xyz2")
    dev = _same_device(xyz1, xyz2)
    b, n, _ = xyz1.shape
    m = xyz2.shape[1]
    dist1 = torch.empty((b, n), dtype=torch.float32, device=dev)
    idx1 = torch.empty((b, n), dtype=torch.int32, device=dev)
    dist2 = torch.empty((b, m), dtype=torch.float32, device=dev)
    idx2 = torch.empty((b, m), dtype=torch.int32, device=dev)
    if dev.type == "cuda":
        ws_bytes = lib.ga_nn_distance_workspace_bytes(b, n, m) if (_pruning and min(n, m) >= 256) else 0
        with _Guard(dev):
            if ws_bytes:
                ws = torch.empty((ws_bytes,), dtype=torch.uint8, device=dev)  # scratch, no state kept
                rc = lib.ga_nn_distance_fwd_ws(b, n, m, xyz1.data_ptr(), xyz2.data_ptr(), dist1.data_ptr(),
                                               idx1.data_ptr(), dist2.data_ptr(), idx2.data_ptr(), mode,
                                               ws.data_ptr(), ws_bytes, _stream(xyz1))
            else:
                rc = lib.ga_nn_distance_fwd(b, n, m, xyz1.data_ptr(), xyz2.data_ptr(), dist1.data_ptr(),
                                            idx1.data_ptr(), dist2.data_ptr(), idx2.data_ptr(), mode,
                                            _stream(xyz1))
    else:
        rc = lib.ga_nn_distance_fwd_host(b, n, m, xyz1.data_ptr(), xyz2.data_ptr(), dist1.data_ptr(),
                                         idx1.data_ptr(), dist2.data_ptr(), idx2.data_ptr(), mode)
    _lib.check(rc)
    return dist1, idx1, dist2, idx2


def nn_distance_grad(xyz1, xyz2, grad_dist1, idx1, grad_dist2, idx2):
    """NnDistanceGrad (tf_nndistance.cpp:84-166): returns grad_xyz1 (b,n,3), grad_xyz2 (b,m,3).
    Dispatches through the registered custom op ``torch.ops.geometric_adv_b200.nn_distance_grad``."""
    lib = _lib.load()
    _lib.check(lib.ga_check_nn_distance_grad(
        xyz1.dim(), _lib.dims(xyz1.shape), xyz2.dim(), _lib.dims(xyz2.shape),
        grad_dist1.dim(), _lib.dims(grad_dist1.shape), idx1.dim(), _lib.dims(idx1.shape),
        grad_dist2.dim(), _lib.dims(grad_dist2.shape), idx2.dim(), _lib.dims(idx2.shape)))
    args = (_prep(xyz1, torch.float32, "xyz1"), _prep(xyz2, torch.float32, "xyz2"),
            _prep(grad_dist1, torch.float32, "grad_dist1"), _prep(idx1, torch.int32, "idx1"),
            _prep(grad_dist2, torch.float32, "grad_dist2"), _prep(idx2, torch.int32, "idx2"))
    _same_device(*args)
    return _nn_distance_grad_op(*args)


def _nn_distance_grad_impl(xyz1, xyz2, grad_dist1, idx1, grad_dist2, idx2):
    lib = _lib.load()
    _lib.check(lib.ga_check_nn_distance_grad(
        xyz1.dim(), _lib.dims(xyz1.shape), xyz2.dim(), _lib.dims(xyz2.shape),
        grad_dist1.dim(), _lib.dims(grad_dist1.shape), idx1.dim(), _lib.dims(idx1.shape),
        grad_dist2.dim(), _lib.dims(grad_dist2.shape), idx2.dim(), _lib.dims(idx2.shape)))
    xyz1 = _prep(xyz1, torch.float32, "xyz1")
    xyz2 = _prep(xyz2, torch.float32, "xyz2")
    grad_dist1 = _prep(grad_dist1, torch.float32, "grad_dist1")
    grad_dist2 = _prep(grad_dist2, torch.float32, "grad_dist2")
    idx1 = _prep(idx1, torch.int32, "idx1")
    idx2 = _prep(idx2, torch.int32, "idx2")
    dev = _same_device(xyz1, xyz2, grad_dist1, grad_dist2, idx1, idx2)
    b, n, _ = xyz1.shape
    m = xyz2.shape[1]
    g1 = torch.empty((b, n, 3), dtype=torch.float32, device=dev)
    g2 = torch.empty((b, m, 3), dtype=torch.float32, device=dev)
    if dev.type == "cuda":
        with _Guard(dev):
            rc = lib.ga_nn_distance_bwd(b, n, m, xyz1.data_ptr(), xyz2.data_ptr(), grad_dist1.data_ptr(),
                                        idx1.data_ptr(), grad_dist2.data_ptr(), idx2.data_ptr(), g1.data_ptr(),
                                        g2.data_ptr(), _stream(xyz1))
    else:
        rc = lib.ga_nn_distance_bwd_host(b, n, m, xyz1.data_ptr(), xyz2.data_ptr(), grad_dist1.data_ptr(),
                                         idx1.data_ptr(), grad_dist2.data_ptr(), idx2.data_ptr(), g1.data_ptr(),
                                         g2.data_ptr())
    _lib.check(rc)
    return g1, g2


# ---- torch.library registration -------------------------------------------------------------------------------
# The shim is a set of PyTorch custom ops (namespace geometric_adv_b200), not opaque Python: torch.compile / export
# see them as single nodes with fake (meta) kernels for shape propagation, and autograd formulas registered on the
# op itself.  The kernels behind them are the C ABI calls above; nothing is computed by torch.
@torch.library.custom_op("geometric_adv_b200::nn_distance", mutates_args=())
def _nn_distance_op(xyz1: torch.Tensor, xyz2: torch.Tensor, mode: int) -> tuple[torch.Tensor, torch.Tensor, torch.Tensor, torch.Tensor]:
    return _nn_distance_fwd(xyz1, xyz2, mode)


@_nn_distance_op.register_fake
def _(xyz1, xyz2, mode):
    b, n, m = xyz1.shape[0], xyz1.shape[1], xyz2.shape[1]
    return (xyz1.new_empty((b, n)), xyz1.new_empty((b, n), dtype=torch.int32), xyz1.new_empty((b, m)),
            xyz1.new_empty((b, m), dtype=torch.int32))


@torch.library.custom_op("geometric_adv_b200::nn_distance_grad", mutates_args=())
def _nn_distance_grad_op(xyz1: torch.Tensor, xyz2: torch.Tensor, grad_dist1: torch.Tensor, idx1: torch.Tensor,
                         grad_dist2: torch.Tensor, idx2: torch.Tensor) -> tuple[torch.Tensor, torch.Tensor]:
    return _nn_distance_grad_impl(xyz1, xyz2, grad_dist1, idx1, grad_dist2, idx2)


@_nn_distance_grad_op.register_fake
def _(xyz1, xyz2, grad_dist1, idx1, grad_dist2, idx2):
    return torch.empty_like(xyz1), torch.empty_like(xyz2)


def _nn_distance_setup(ctx, inputs, output):
    xyz1, xyz2, _mode = inputs
    _d1, idx1, _d2, idx2 = output
    ctx.save_for_backward(xyz1, xyz2, idx1, idx2)


def _nn_distance_backward(ctx, grad_dist1, grad_idx1, grad_dist2, grad_idx2):
    """Gradient only through dist1 / dist2; idx non-differentiable (tf_nndistance.py:35-41; dist_chamfer_3D.py:49-64)."""
    xyz1, xyz2, idx1, idx2 = ctx.saved_tensors
    if grad_dist1 is None:
        grad_dist1 = torch.zeros(idx1.shape, dtype=torch.float32, device=idx1.device)
    if grad_dist2 is None:
        grad_dist2 = torch.zeros(idx2.shape, dtype=torch.float32, device=idx2.device)
    g1, g2 = _nn_distance_grad_op(xyz1, xyz2, grad_dist1.contiguous(), idx1, grad_dist2.contiguous(), idx2)
    return g1, g2, None


_nn_distance_op.register_autograd(_nn_distance_backward, setup_context=_nn_distance_setup)


def _check_nn_distance_args(xyz1, xyz2):
    lib = _lib.load()
    _lib.check(lib.ga_check_nn_distance(xyz1.dim(), _lib.dims(xyz1.shape), xyz2.dim(), _lib.dims(xyz2.shape)))
    for t, name in ((xyz1, "xyz1"), (xyz2, "xyz2")):
        if not isinstance(t, torch.Tensor):
            raise TypeError("%s must be a torch.Tensor" % name)
        if t.dtype != torch.float32:
            raise TypeError("%s must be %s, got %s" % (name, torch.float32, t.dtype))
    _same_device(xyz1, xyz2)


def nn_distance(xyz1, xyz2, mode=None):
    """Nearest-neighbour distances for a pair of batched clouds (tf_nndistance.py:15-26).

    xyz1 (B,N,3), xyz2 (B,M,3) float32 -> dist1 (B,N) squared distance from each point of
    xyz1 to its nearest point of xyz2, idx1 (B,N) int32 its index, dist2 / idx2 the other
    way.  Lowest index wins ties.  Differentiable through dist1 / dist2.
    Dispatches through the registered custom op ``torch.ops.geometric_adv_b200.nn_distance``."""
    mode = _default_mode if mode is None else mode
    _check_nn_distance_args(xyz1, xyz2)
    return _nn_distance_op(xyz1.contiguous(), xyz2.contiguous(), int(mode))


class chamfer_3DFunction:
    """dist_chamfer_3D.py:26-64 -- same op, torch return order (dist1, dist2, idx1, idx2).  Kept as a class with
    ``apply`` because callers use ``chamfer_3DFunction.apply(a, b)``; the work is the registered custom op."""

    @staticmethod
    def apply(xyz1, xyz2):
        dist1, idx1, dist2, idx2 = nn_distance(xyz1, xyz2)
        return dist1, dist2, idx1, idx2


class chamfer_3DDist(torch.nn.Module):
    """Drop-in for AtlasNet's chamfer_3DDist (dist_chamfer_3D.py:66-74)."""

    def forward(self, input1, input2):
        input1 = input1.contiguous()
        input2 = input2.contiguous()
        return chamfer_3DFunction.apply(input1, input2)


def chamfer_per_cloud(dist1, dist2):
    """mean(dist1,1) + mean(dist2,1) per cloud (src/adv_ae.py:120-121), one fused reduction."""
    lib = _lib.load()
    dist1 = _prep(dist1, torch.float32, "dist1")
    dist2 = _prep(dist2, torch.float32, "dist2")
    dev = _same_device(dist1, dist2)
    if dev.type != "cuda":
        raise ValueError("chamfer_per_cloud expects CUDA tensors")
    b, n = dist1.shape
    m = dist2.shape[1]
    out = torch.empty((b,), dtype=torch.float32, device=dev)
    with _Guard(dev):
        _lib.check(lib.ga_chamfer_per_cloud(b, n, m, dist1.data_ptr(), dist2.data_ptr(), out.data_ptr(),
                                            _stream(dist1)))
    return out


# --------------------------------------------------------------------------- fused Chamfer loss terms
@torch.library.custom_op("geometric_adv_b200::chamfer_loss_terms", mutates_args=())
def _chamfer_loss_terms_op(xyz1: torch.Tensor, xyz2: torch.Tensor, mode: int) -> tuple[torch.Tensor, torch.Tensor, torch.Tensor, torch.Tensor]:
    """(cd, max1, idx1, idx2): the search and ONE reduction launch (ga_chamfer_loss_terms)."""
    lib = _lib.load()
    dist1, idx1, dist2, idx2 = _nn_distance_fwd(xyz1, xyz2, mode)
    b, n = dist1.shape
    m = dist2.shape[1]
    cd = torch.empty((b,), dtype=torch.float32, device=dist1.device)
    mx = torch.empty((b,), dtype=torch.float32, device=dist1.device)
    if dist1.device.type != "cuda":
        raise ValueError("chamfer_loss_terms expects CUDA tensors")
    with _Guard(dist1.device):
        _lib.check(lib.ga_chamfer_loss_terms(b, n, m, dist1.data_ptr(), dist2.data_ptr(), cd.data_ptr(), mx.data_ptr(),
                                             _stream(dist1)))
    return cd, mx, idx1, idx2


@_chamfer_loss_terms_op.register_fake
def _(xyz1, xyz2, mode):
    b, n, m = xyz1.shape[0], xyz1.shape[1], xyz2.shape[1]
    return (xyz1.new_empty((b,)), xyz1.new_empty((b,)), xyz1.new_empty((b, n), dtype=torch.int32),
            xyz1.new_empty((b, m), dtype=torch.int32))


def _chamfer_loss_terms_setup(ctx, inputs, output):
    xyz1, xyz2, _mode = inputs
    _cd, _mx, idx1, idx2 = output
    ctx.save_for_backward(xyz1, xyz2, idx1, idx2)


def _chamfer_loss_terms_backward(ctx, grad_cd, grad_max, grad_idx1, grad_idx2):
    """d cd / d dist1 = 1/N, d cd / d dist2 = 1/M, exactly what autograd derives for mean(dist1,1) + mean(dist2,1):
    grad.unsqueeze(1).expand(...) / N.  The max term is an evaluation metric (adv_ae.py:131-133) and carries no
    gradient, like idx."""
    xyz1, xyz2, idx1, idx2 = ctx.saved_tensors
    if grad_cd is None:
        grad_cd = torch.zeros((idx1.shape[0],), dtype=torch.float32, device=idx1.device)
    n, m = idx1.shape[1], idx2.shape[1]
    gd1 = (grad_cd.unsqueeze(1).expand(-1, n) / n).contiguous()
    gd2 = (grad_cd.unsqueeze(1).expand(-1, m) / m).contiguous()
    g1, g2 = _nn_distance_grad_op(xyz1, xyz2, gd1, idx1, gd2, idx2)
    return g1, g2, None


_chamfer_loss_terms_op.register_autograd(_chamfer_loss_terms_backward, setup_context=_chamfer_loss_terms_setup)


def chamfer_loss_terms(xyz1, xyz2, mode=None):
    """The per-cloud loss terms of the attack graph from ONE search and ONE reduction launch:
    cd = mean(dist1,1) + mean(dist2,1) (adv_ae.py:120-121; bit-equal to chamfer_per_cloud(nn_distance(...))) and
    max1 = max(dist1,1) (:131-133).  Differentiable through cd.  Returns (cd, max1, idx1, idx2)."""
    mode = _default_mode if mode is None else mode
    _check_nn_distance_args(xyz1, xyz2)
    return _chamfer_loss_terms_op(xyz1.contiguous(), xyz2.contiguous(), int(mode))


def chamfer_all_pairs(clouds, row0=0, rows=None, mode=None, directed=False):
    """Rows [row0, row0+rows) of the all-pairs Chamfer matrix of
    attacker/prepare_indices_for_attack.py:104-139: out[r, j] = CD(source=clouds[j],
    target=clouds[row0+r]).  directed=True returns only D[row0+r -> j] (mean over the points
    of clouds[row0+r] of the squared distance to their nearest point of clouds[j]); CD = D + D^T."""
    lib = _lib.load()
    clouds = _prep(clouds, torch.float32, "clouds")
    if clouds.dim() != 3 or clouds.shape[2] != 3:
        raise ValueError("clouds must be (S,N,3)")
    if clouds.device.type != "cuda":
        raise ValueError("chamfer_all_pairs expects a CUDA tensor")
    s, n, _ = clouds.shape
    rows = s - row0 if rows is None else rows
    if row0 < 0 or rows < 0 or row0 + rows > s:
        raise ValueError("row block [%d, %d) outside 0..%d" % (row0, row0 + rows, s))
    out = torch.empty((rows, s), dtype=torch.float32, device=clouds.device)
    mode = _default_mode if mode is None else mode
    fn = lib.ga_chamfer_all_pairs_directed if directed else lib.ga_chamfer_all_pairs
    with _Guard(clouds.device):
        _lib.check(fn(s, n, clouds.data_ptr(), row0, rows, out.data_ptr(), mode, _stream(clouds)))
    return out


# --------------------------------------------------------------------------- grouping
def knn_point(k, xyz1, xyz2):
    """k nearest neighbours (tf_grouping.py:48-75).  xyz1 (B,N,3) data set, xyz2 (B,M,3)
    queries -> val (B,M,k) squared distances ascending, idx (B,M,k) int32 into xyz1, with
    the tie behaviour of the reference's selection sort.  No gradient (ops.NoGradient)."""
    lib = _lib.load()
    xyz1 = _prep(xyz1.detach(), torch.float32, "xyz1")
    xyz2 = _prep(xyz2.detach(), torch.float32, "xyz2")
    if xyz1.dim() != 3 or xyz2.dim() != 3 or xyz1.shape[2] != 3 or xyz2.shape[2] != 3:
        raise ValueError("knn_point expects (batch_size, ndataset, 3) xyz1 and (batch_size, npoint, 3) xyz2")
    if xyz1.shape[0] != xyz2.shape[0]:
        raise ValueError("knn_point expects xyz1 and xyz2 have same batch size")
    _same_device(xyz1, xyz2)
    return _knn_point_op(int(k), xyz1, xyz2)


@torch.library.custom_op("geometric_adv_b200::knn_point", mutates_args=())
def _knn_point_op(k: int, xyz1: torch.Tensor, xyz2: torch.Tensor) -> tuple[torch.Tensor, torch.Tensor]:
    lib = _lib.load()
    dev = xyz1.device
    b, n, _ = xyz1.shape
    m = xyz2.shape[1]
    val = torch.empty((b, m, max(k, 0)), dtype=torch.float32, device=dev)
    idx = torch.empty((b, m, max(k, 0)), dtype=torch.int32, device=dev)
    if dev.type == "cuda":
        with _Guard(dev):
            rc = lib.ga_knn(b, n, m, k, xyz1.data_ptr(), xyz2.data_ptr(), val.data_ptr(), idx.data_ptr(),
                            _stream(xyz1))
    else:
        rc = lib.ga_knn_host(b, n, m, k, xyz1.data_ptr(), xyz2.data_ptr(), val.data_ptr(), idx.data_ptr())
    _lib.check(rc)
    return val, idx


@_knn_point_op.register_fake
def _(k, xyz1, xyz2):
    b, m = xyz2.shape[0], xyz2.shape[1]
    return xyz1.new_empty((b, m, max(k, 0))), xyz1.new_empty((b, m, max(k, 0)), dtype=torch.int32)


def select_top_k(k, dist):
    """Legacy dense entry (tf_grouping.py:22-31): dist (b,m,n) -> idx (b,m,n), dist_out (b,m,n);
    the first k columns hold the k smallest, selection-sort order."""
    lib = _lib.load()
    _lib.check(lib.ga_check_selection_sort(int(k), dist.dim(), _lib.dims(dist.shape)))
    dist = _prep(dist.detach(), torch.float32, "dist")
    if dist.device.type != "cuda":
        raise ValueError("select_top_k expects a CUDA tensor (the reference op has GPU kernels only, "
                         "tf_grouping.cpp:139)")
    b, m, n = dist.shape
    outi = torch.empty((b, m, n), dtype=torch.int32, device=dist.device)
    out = torch.empty((b, m, n), dtype=torch.float32, device=dist.device)
    with _Guard(dist.device):
        _lib.check(lib.ga_selection_sort(b, n, m, int(k), dist.data_ptr(), outi.data_ptr(), out.data_ptr(),
                                         _stream(dist)))
    return outi, out


@torch.library.custom_op("geometric_adv_b200::group_point", mutates_args=())
def _group_point_op(points: torch.Tensor, idx: torch.Tensor) -> torch.Tensor:
    lib = _lib.load()
    b, n, c = points.shape
    _, m, ns = idx.shape
    out = torch.empty((b, m, ns, c), dtype=torch.float32, device=points.device)
    with _Guard(points.device):
        _lib.check(lib.ga_group_point(b, n, c, m, ns, points.data_ptr(), idx.data_ptr(), out.data_ptr(),
                                      _stream(points)))
    return out


@_group_point_op.register_fake
def _(points, idx):
    return points.new_empty((idx.shape[0], idx.shape[1], idx.shape[2], points.shape[2]))


def _group_point_setup(ctx, inputs, output):
    points, idx = inputs
    ctx.save_for_backward(idx)
    ctx.shape = tuple(points.shape)


def _group_point_backward(ctx, grad_out):
    # GroupPointGrad (tf_grouping_g.cu:61-78) is a scatter-add; no script differentiates through
    # group_point, so this off-path gradient uses torch's index_add_.
    (idx,) = ctx.saved_tensors
    b, n, c = ctx.shape
    g = torch.zeros((b * n, c), dtype=grad_out.dtype, device=grad_out.device)
    flat = (idx.long() + torch.arange(b, device=idx.device).view(b, 1, 1) * n).reshape(-1)
    g.index_add_(0, flat, grad_out.reshape(-1, c))
    return g.view(b, n, c), None


_group_point_op.register_autograd(_group_point_backward, setup_context=_group_point_setup)


def group_point(points, idx):
    """out[b,j,s,:] = points[b, idx[b,j,s], :] (tf_grouping.py:33-47)."""
    lib = _lib.load()
    _lib.check(lib.ga_check_group_point(points.dim(), _lib.dims(points.shape), idx.dim(), _lib.dims(idx.shape)))
    points = _prep(points, torch.float32, "points")
    idx = _prep(idx, torch.int32, "idx")
    dev = _same_device(points, idx)
    if dev.type != "cuda":
        raise ValueError("group_point expects CUDA tensors (the reference op has GPU kernels only, "
                         "tf_grouping.cpp:171)")
    return _group_point_op(points, idx)


def knn_dists(pc, k):
    """Per-point distances to the k nearest other points, (B,N,k) float32: the graph of
    defender/get_knn_dists_per_point.py:78-81 (knn_point(k+1) -> drop self -> group_point ->
    subtract centre -> sqrt(sum(sq))) as one kernel."""
    lib = _lib.load()
    pc = _prep(pc.detach(), torch.float32, "pc")
    if pc.dim() != 3 or pc.shape[2] != 3:
        raise ValueError("knn_dists expects (batch_size, num_points, 3)")
    return _knn_dists_op(pc, int(k))


@torch.library.custom_op("geometric_adv_b200::knn_dists", mutates_args=())
def _knn_dists_op(pc: torch.Tensor, k: int) -> torch.Tensor:
    lib = _lib.load()
    b, n, _ = pc.shape
    out = torch.empty((b, n, max(k, 0)), dtype=torch.float32, device=pc.device)
    if pc.device.type == "cuda":
        with _Guard(pc.device):
            rc = lib.ga_knn_dists(b, n, k, pc.data_ptr(), out.data_ptr(), _stream(pc))
    else:
        rc = lib.ga_knn_dists_host(b, n, k, pc.data_ptr(), out.data_ptr())
    _lib.check(rc)
    return out


@_knn_dists_op.register_fake
def _(pc, k):
    return pc.new_empty((pc.shape[0], pc.shape[1], max(k, 0)))
