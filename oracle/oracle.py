"""TEST INFRASTRUCTURE ONLY -- not part of the product.

numpy/ctypes front end of the CPU oracle:

* ``libga_oracle.so``  - the C restatement in ``oracle/ga_oracle.c`` (always
  built by ``make -C oracle``);
* ``_ref/libga_ref.so`` - the reference's own CPU kernels
  (external/structural_losses/tf_nndistance.cpp) compiled unmodified, present
  when the oracle was built in a container that has /root/reference.

Only tests/, ``__graft_entry__.smoke()`` and bench.py's ``cpu_baseline`` /
``--impl reference`` legs may import this module.  The product package
``geometric_adv_b200`` never does (tests/test_boundary.py greps for it).
"""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
_ORACLE_SO = os.path.join(HERE, "libga_oracle.so")
_REF_SO = os.path.join(HERE, "_ref", "libga_ref.so")
_REF_GPU_SO = os.path.join(HERE, "_ref", "libga_ref_gpu.so")

_f32p = C.POINTER(C.c_float)
_i32p = C.POINTER(C.c_int)


def build(quiet=True):
    """(Re)build the oracle with make; builds _ref/ too when /root/reference exists."""
    subprocess.run(["make", "-C", HERE], check=True, stdout=subprocess.DEVNULL if quiet else None)


def _fp(a):
    return a.ctypes.data_as(_f32p)


def _ip(a):
    return a.ctypes.data_as(_i32p)


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def _i32(a):
    return np.ascontiguousarray(a, dtype=np.int32)


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_ORACLE_SO):
            build()
        _lib = C.CDLL(_ORACLE_SO)
    return _lib


_ref = None


def have_ref():
    return os.path.exists(_REF_SO)


def ref():
    global _ref
    if _ref is None:
        _ref = C.CDLL(_REF_SO)
        _ref.ga_ref_last_error.restype = C.c_char_p
    return _ref


# ---------------------------------------------------------------- restatement
def nn_distance(xyz1, xyz2, mode=0):
    """tf_nndistance.cpp:21-83.  Returns dist1, idx1, dist2, idx2 (TF order)."""
    xyz1, xyz2 = _f32(xyz1), _f32(xyz2)
    b, n, _ = xyz1.shape
    m = xyz2.shape[1]
    d1 = np.empty((b, n), np.float32)
    i1 = np.empty((b, n), np.int32)
    d2 = np.empty((b, m), np.float32)
    i2 = np.empty((b, m), np.int32)
    lib().ga_oracle_nn_distance(b, n, m, _fp(xyz1), _fp(xyz2), _fp(d1), _ip(i1), _fp(d2), _ip(i2), int(mode))
    return d1, i1, d2, i2


def nn_distance_grad(xyz1, xyz2, gd1, idx1, gd2, idx2):
    """tf_nndistance.cpp:122-163.  Returns grad_xyz1, grad_xyz2."""
    xyz1, xyz2, gd1, gd2 = _f32(xyz1), _f32(xyz2), _f32(gd1), _f32(gd2)
    idx1, idx2 = _i32(idx1), _i32(idx2)
    b, n, _ = xyz1.shape
    m = xyz2.shape[1]
    g1 = np.empty((b, n, 3), np.float32)
    g2 = np.empty((b, m, 3), np.float32)
    lib().ga_oracle_nn_distance_grad(b, n, m, _fp(xyz1), _fp(xyz2), _fp(gd1), _ip(idx1), _fp(gd2), _ip(idx2),
                                     _fp(g1), _fp(g2))
    return g1, g2


def chamfer_per_cloud(dist1, dist2):
    dist1, dist2 = _f32(dist1), _f32(dist2)
    b, n = dist1.shape
    m = dist2.shape[1]
    out = np.empty((b,), np.float32)
    lib().ga_oracle_chamfer_per_cloud(b, n, m, _fp(dist1), _fp(dist2), _fp(out))
    return out


def selection_sort(dist, k):
    """tf_grouping_g.cu:83-123.  dist (b,m,n) -> outi (b,m,n) int32, out (b,m,n)."""
    dist = _f32(dist)
    b, m, n = dist.shape
    outi = np.empty((b, m, n), np.int32)
    out = np.empty((b, m, n), np.float32)
    lib().ga_oracle_selection_sort(b, n, m, int(k), _fp(dist), _ip(outi), _fp(out))
    return outi, out


def knn_point(k, xyz1, xyz2):
    """tf_grouping.py:48-75.  xyz1 data set (b,n,3), xyz2 queries (b,m,3) -> val, idx (b,m,k)."""
    xyz1, xyz2 = _f32(xyz1), _f32(xyz2)
    b, n, _ = xyz1.shape
    m = xyz2.shape[1]
    val = np.empty((b, m, k), np.float32)
    idx = np.empty((b, m, k), np.int32)
    lib().ga_oracle_knn_point(b, n, m, int(k), _fp(xyz1), _fp(xyz2), _fp(val), _ip(idx))
    return val, idx


def group_point(points, idx):
    """tf_grouping_g.cu:40-57."""
    points, idx = _f32(points), _i32(idx)
    b, n, c = points.shape
    _, m, ns = idx.shape
    out = np.empty((b, m, ns, c), np.float32)
    lib().ga_oracle_group_point(b, n, c, m, ns, _fp(points), _ip(idx), _fp(out))
    return out


def knn_dists(pc, k):
    """defender/get_knn_dists_per_point.py:78-81 -> (b,n,k) distances (sqrt applied)."""
    pc = _f32(pc)
    b, n, _ = pc.shape
    out = np.empty((b, n, k), np.float32)
    lib().ga_oracle_knn_dists(b, n, int(k), _fp(pc), _fp(out))
    return out


def split_by_threshold(pc, score, thresh):
    """src/adversary_utils.py:149-178 -> outlier_pc, outlier_idx, outlier_num, inlier_pc."""
    pc, score = _f32(pc), _f32(score)
    b, n, _ = pc.shape
    opc = np.empty((b, n, 3), np.float32)
    ipc = np.empty((b, n, 3), np.float32)
    oidx = np.empty((b, n), np.int32)
    onum = np.empty((b,), np.int32)
    lib().ga_oracle_split_by_threshold(b, n, _fp(pc), _fp(score), C.c_float(thresh), _fp(opc), _ip(oidx),
                                       _ip(onum), _fp(ipc))
    return opc, oidx, onum, ipc


def knn_dists_numpy(pc, k):
    """The reference's numpy fallback, restated with numpy itself:
    src/general_utils.py:94-106 (get_dist_mat) + get_knn_dists_per_point.py:125-137."""
    pc = _f32(pc)
    out = np.empty(pc.shape[:2] + (k,), np.float32)
    for l in range(pc.shape[0]):
        data = pc[l]
        n = len(data)
        source = np.tile(np.expand_dims(data, axis=0), [n, 1, 1])
        target = np.tile(np.expand_dims(data, axis=1), [1, n, 1])
        dist_mat = np.linalg.norm(source - target, axis=-1)
        out[l] = np.sort(dist_mat, axis=1)[:, 1:k + 1]
    return out


# ------------------------------------------------ the reference's own CPU code
def ref_nn_distance(xyz1, xyz2, threads=1):
    """NnDistanceOp::Compute (tf_nndistance.cpp:45-83), compiled unmodified."""
    xyz1, xyz2 = _f32(xyz1), _f32(xyz2)
    b, n, _ = xyz1.shape
    m = xyz2.shape[1]
    d1 = np.empty((b, n), np.float32)
    i1 = np.empty((b, n), np.int32)
    d2 = np.empty((b, m), np.float32)
    i2 = np.empty((b, m), np.int32)
    if threads == 1:
        rc = ref().ga_ref_nn_distance(b, n, m, _fp(xyz1), _fp(xyz2), _fp(d1), _ip(i1), _fp(d2), _ip(i2))
    else:
        rc = ref().ga_ref_nn_distance_mt(b, n, m, _fp(xyz1), _fp(xyz2), _fp(d1), _ip(i1), _fp(d2), _ip(i2),
                                         int(threads))
    if rc != 0:
        raise ValueError(ref().ga_ref_last_error().decode())
    return d1, i1, d2, i2


def ref_nn_distance_shaped(xyz1, xyz2):
    """Same op, shapes passed as they are so the reference's own argument checks
    (tf_nndistance.cpp:51-58) raise their own messages."""
    xyz1, xyz2 = _f32(xyz1), _f32(xyz2)
    d1s = (C.c_longlong * max(1, xyz1.ndim))(*xyz1.shape)
    d2s = (C.c_longlong * max(1, xyz2.ndim))(*xyz2.shape)
    cap = max(1, xyz1.size, xyz2.size)
    d1 = np.empty(cap, np.float32)
    i1 = np.empty(cap, np.int32)
    d2 = np.empty(cap, np.float32)
    i2 = np.empty(cap, np.int32)
    rc = ref().ga_ref_nn_distance_shaped(_fp(xyz1), xyz1.ndim, d1s, _fp(xyz2), xyz2.ndim, d2s, _fp(d1), _ip(i1),
                                         _fp(d2), _ip(i2), C.c_longlong(cap), C.c_longlong(cap))
    if rc != 0:
        raise ValueError(ref().ga_ref_last_error().decode())
    b, n = xyz1.shape[:2]
    m = xyz2.shape[1]
    return (d1[:b * n].reshape(b, n), i1[:b * n].reshape(b, n), d2[:b * m].reshape(b, m),
            i2[:b * m].reshape(b, m))


def ref_nn_distance_grad(xyz1, xyz2, gd1, idx1, gd2, idx2, threads=1):
    """NnDistanceGradOp::Compute (tf_nndistance.cpp:84-166), compiled unmodified."""
    xyz1, xyz2, gd1, gd2 = _f32(xyz1), _f32(xyz2), _f32(gd1), _f32(gd2)
    idx1, idx2 = _i32(idx1), _i32(idx2)
    b, n, _ = xyz1.shape
    m = xyz2.shape[1]
    g1 = np.empty((b, n, 3), np.float32)
    g2 = np.empty((b, m, 3), np.float32)
    args = (b, n, m, _fp(xyz1), _fp(xyz2), _fp(gd1), _ip(idx1), _fp(gd2), _ip(idx2), _fp(g1), _fp(g2))
    if threads == 1:
        rc = ref().ga_ref_nn_distance_grad(*args)
    else:
        rc = ref().ga_ref_nn_distance_grad_mt(*args, int(threads))
    if rc != 0:
        raise ValueError(ref().ga_ref_last_error().decode())
    return g1, g2


def ref_max_threads():
    return int(ref().ga_ref_max_threads())
