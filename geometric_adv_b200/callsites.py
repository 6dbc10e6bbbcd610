"""Two more consumers of the same kernels: the FoldingNet transfer model.

* ``foldingnet_chamfer_distance`` replaces ``ChamferDistance`` of
  ``transfer/foldingnet/foldingnet.py:209-238``, which materialises two (B,2048,2025,3) tensors
  (~200 MB per batch element) to take row/column minima.  Same value up to fp32 summation
  order, differentiable, N != M.
* ``foldingnet_knn_graph`` replaces ``knn_search`` of ``transfer/foldingnet/prepare_graph.py:45-73``
  (CPU scikit-learn KDTree, k=16, one cloud at a time in a Pool(2)): neighbour indices and
  distances for a whole batch in one kernel, plus the 3x3 local covariance of the 16 neighbours
  (``np.cov`` semantics: unbiased, neighbours only, the point itself excluded).
"""
import torch

from . import ops


def foldingnet_chamfer_distance(x, y):
    """x (B,N,3) reconstruction, y (B,M,3) input -> scalar
    mean_b(mean_n min_m |x-y|^2) + mean_b(mean_m min_n |x-y|^2)  (foldingnet.py:224-236)."""
    d_x, _, d_y, _ = ops.nn_distance(x.contiguous(), y.contiguous())
    return d_x.mean() + d_y.mean()


class ChamferLoss(torch.nn.Module):
    """Drop-in for foldingnet.py's ChamferLoss module."""

    def forward(self, x, y):
        return foldingnet_chamfer_distance(x, y)


def foldingnet_knn_graph(pc, knn=16):
    """pc (B,N,3) -> idx (B,N,knn) int32 neighbour indices (self excluded), dist (B,N,knn+1)
    Euclidean distances including the point itself in column 0 (as KDTree.query returns them),
    cov (B,N,9) row-major 3x3 covariance of each point's knn neighbours."""
    val, idx = ops.knn_point(knn + 1, pc, pc)
    dist = val.sqrt()
    nb = ops.group_point(pc.contiguous(), idx[:, :, 1:].contiguous())      # (B,N,knn,3)
    centred = nb - nb.mean(dim=2, keepdim=True)
    cov = torch.einsum("bnki,bnkj->bnij", centred, centred) / float(knn - 1)  # np.cov: ddof = 1
    return idx[:, :, 1:].contiguous(), dist, cov.reshape(pc.shape[0], pc.shape[1], 9)


def knn_edges(idx, symmetric=True):
    """Edge list of the kNN graph like prepare_graph.py:60-70 builds it (a set of (i,j) pairs,
    symmetrised), per cloud: returns a list of (2,E) int64 tensors."""
    b, n, k = idx.shape
    src = torch.arange(n, device=idx.device).view(1, n, 1).expand(b, n, k)
    out = []
    for i in range(b):
        e = torch.stack([src[i].reshape(-1), idx[i].reshape(-1).long()])
        if symmetric:
            e = torch.cat([e, e.flip(0)], dim=1)
        out.append(torch.unique(e, dim=1))
    return out
