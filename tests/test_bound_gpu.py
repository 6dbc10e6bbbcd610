"""The drop-in boundary, for real: the reference's tf_nndistance.cpp compiled unmodified
together with geometric_adv_b200/bindings/tf_nndistance_b200.cpp and linked against libga_b200.so
(oracle/_ref/libga_ref_bound.so, built by oracle/Makefile).  The reference's own GPU OpKernels
(NnDistanceGpuOp::Compute, NnDistanceGradGpuOp::Compute, tf_nndistance.cpp:169-252) run with device pointers and
must return the bits of its CPU OpKernels (NnDistanceOp / NnDistanceGradOp, oracle/_ref/libga_ref.so)."""
import ctypes as C
import os

import numpy as np
import pytest
import torch

from util import bits_equal, cloud

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
HERE = os.path.dirname(os.path.abspath(__file__))
BOUND = os.path.join(os.path.dirname(HERE), "oracle", "_ref", "libga_ref_bound.so")
p = C.c_void_p


@pytest.fixture(scope="module")
def bound():
    if not os.path.exists(BOUND):
        pytest.skip("oracle/_ref/libga_ref_bound.so was not built (needs /root/reference at build time)")
    lib = C.CDLL(BOUND)
    lib.ga_bound_last_error.restype = C.c_char_p
    lib.ga_b200_tf_set_stream(None)
    return lib


def t(a):
    return torch.from_numpy(np.ascontiguousarray(a)).to(DEV)


def run_fwd(lib, a, b):
    ta, tb = t(a), t(b)
    bsz, n = a.shape[0], a.shape[1]
    m = b.shape[1] if b.ndim > 1 else 0
    cap1, cap2 = max(1, a.size), max(1, b.size)
    d1 = torch.empty(cap1, device=DEV)
    i1 = torch.empty(cap1, dtype=torch.int32, device=DEV)
    d2 = torch.empty(cap2, device=DEV)
    i2 = torch.empty(cap2, dtype=torch.int32, device=DEV)
    s1 = (C.c_longlong * max(1, a.ndim))(*a.shape)
    s2 = (C.c_longlong * max(1, b.ndim))(*b.shape)
    rc = lib.ga_bound_nn_distance_gpu(p(ta.data_ptr()), a.ndim, s1, p(tb.data_ptr()), b.ndim, s2, p(d1.data_ptr()),
                                      p(i1.data_ptr()), p(d2.data_ptr()), p(i2.data_ptr()), C.c_longlong(cap1),
                                      C.c_longlong(cap2))
    torch.cuda.synchronize()
    if rc != 0:
        raise ValueError(lib.ga_bound_last_error().decode())
    return (d1[:bsz * n].view(bsz, n).cpu().numpy(), i1[:bsz * n].view(bsz, n).cpu().numpy(),
            d2[:bsz * m].view(bsz, m).cpu().numpy(), i2[:bsz * m].view(bsz, m).cpu().numpy())


@pytest.mark.parametrize("shape", [(50, 2048, 2048), (3, 100, 200), (1, 2048, 2048), (7, 513, 1000)])
def test_reference_gpu_ops_on_our_kernels_equal_reference_cpu_ops(bound, oracle, shape):
    if not oracle.have_ref():
        pytest.skip("reference CPU build not present")
    b, n, m = shape
    x1, x2 = cloud(11, (b, n, 3)), cloud(12, (b, m, 3))
    bound.ga_b200_tf_set_mode(0)  # GA_MODE_CPU_EXACT: the bits of NnDistanceOp
    got = run_fwd(bound, x1, x2)
    want = oracle.ref_nn_distance(x1, x2, threads=oracle.ref_max_threads())
    for g, w, name in zip(got, want, ("dist1", "idx1", "dist2", "idx2")):
        assert bits_equal(g, w), name
    rng = np.random.default_rng(5)
    g1 = rng.standard_normal((b, n)).astype(np.float32)
    g2 = rng.standard_normal((b, m)).astype(np.float32)
    tx1, tx2, tg1, tg2, ti1, ti2 = t(x1), t(x2), t(g1), t(g2), t(want[1]), t(want[3])
    o1 = torch.empty(b, n, 3, device=DEV)
    o2 = torch.empty(b, m, 3, device=DEV)
    rc = bound.ga_bound_nn_distance_grad_gpu(b, n, m, p(tx1.data_ptr()), p(tx2.data_ptr()), p(tg1.data_ptr()),
                                             p(ti1.data_ptr()), p(tg2.data_ptr()), p(ti2.data_ptr()), p(o1.data_ptr()),
                                             p(o2.data_ptr()))
    torch.cuda.synchronize()
    assert rc == 0, bound.ga_bound_last_error()
    w1, w2 = oracle.ref_nn_distance_grad(x1, x2, g1, want[1], g2, want[3], threads=oracle.ref_max_threads())
    assert bits_equal(o1.cpu().numpy(), w1) and bits_equal(o2.cpu().numpy(), w2)


def test_gpu_ref_mode_matches_the_replaced_cuda_kernel(bound, oracle):
    """Default mode of the binding (GA_MODE_GPU_REF) = the bits of tf_nndistance_g.cu compiled for sm_100a."""
    ref_gpu = os.path.join(os.path.dirname(BOUND), "libga_ref_gpu.so")
    if not os.path.exists(ref_gpu):
        pytest.skip("reference CUDA build not present")
    import geometric_adv_b200 as ga
    x1, x2 = cloud(21, (4, 1000, 3)), cloud(22, (4, 1500, 3))
    bound.ga_b200_tf_set_mode(1)
    try:
        got = run_fwd(bound, x1, x2)
    finally:
        bound.ga_b200_tf_set_mode(0)
    want = [v.cpu().numpy() for v in ga.nn_distance(t(x1), t(x2), mode=ga.GA_MODE_GPU_REF)]
    for g, w in zip(got, want):
        assert bits_equal(g, w)


@pytest.mark.parametrize("a,b,msg", [
    (np.zeros((2, 5), np.float32), np.zeros((2, 5, 3), np.float32), "NnDistance requires xyz1 be of shape (batch,#points,3)"),
    (np.zeros((2, 5, 2), np.float32), np.zeros((2, 5, 3), np.float32), "NnDistance only accepts 3d point set xyz1"),
    (np.zeros((2, 5, 3), np.float32), np.zeros((3, 5, 3), np.float32), "NnDistance expects xyz1 and xyz2 have same batch size"),
])
def test_reference_checks_still_guard_the_gpu_op(bound, a, b, msg):
    with pytest.raises(ValueError) as e:
        run_fwd(bound, a, b)
    assert msg in str(e.value)
