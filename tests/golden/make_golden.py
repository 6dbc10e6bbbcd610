"""Generate the golden fixtures under tests/golden/ FROM THE REFERENCE ITSELF.

Run in the build container only (needs /root/reference):

    make -C oracle && python tests/golden/make_golden.py

Sources of truth used here
  * oracle/_ref/libga_ref.so  -- NnDistanceOp / NnDistanceGradOp CPU kernels of
    external/structural_losses/tf_nndistance.cpp compiled UNMODIFIED;
  * transfer/atlasnet/auxiliary/ChamferDistancePytorch/chamfer_python.py imported
    as is (torch fp64 reference used by the reference's unit_test.py:14-35);
  * oracle/_ref/selection_sort_ref.out -- stdout of the reference's
    external/grouping/test/selection_sort.cpp (known-answer vector).
The fixtures are small .npz files; large cases store SHA-256 digests of the
reference outputs instead of the arrays.
"""
import hashlib
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from oracle import oracle as O  # noqa: E402

REF = "/root/reference"


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def cloud(seed, shape, lo=-0.5, hi=0.5):
    rng = np.random.default_rng(seed)
    return (rng.random(shape, dtype=np.float32) * np.float32(hi - lo) + np.float32(lo)).astype(np.float32)


def main():
    assert O.have_ref(), "build oracle/_ref first (make -C oracle)"
    out = {}

    # (i) unit_test.py:15-33 shape: (4,100,3) vs (4,200,3), uniform [0,1)
    a = cloud(10, (4, 100, 3), 0.0, 1.0)
    b = cloud(11, (4, 200, 3), 0.0, 1.0)
    d1, i1, d2, i2 = O.ref_nn_distance(a, b)
    gd1 = np.random.default_rng(12).standard_normal((4, 100)).astype(np.float32)
    gd2 = np.random.default_rng(13).standard_normal((4, 200)).astype(np.float32)
    g1, g2 = O.ref_nn_distance_grad(a, b, gd1, i1, gd2, i2)
    np.savez_compressed(os.path.join(HERE, "nnd_unit_4x100x200.npz"), xyz1=a, xyz2=b, dist1=d1, idx1=i1, dist2=d2,
                        idx2=i2, gd1=gd1, gd2=gd2, gxyz1=g1, gxyz2=g2)

    # chamfer_python (fp64) on the same inputs
    sys.path.insert(0, os.path.join(REF, "transfer/atlasnet/auxiliary/ChamferDistancePytorch"))
    import torch
    import chamfer_python
    p1, p2, pi1, pi2 = chamfer_python.distChamfer(torch.from_numpy(a), torch.from_numpy(b))
    np.savez_compressed(os.path.join(HERE, "chamfer_python_4x100x200.npz"), dist1=p1.numpy(), dist2=p2.numpy(),
                        idx1=pi1.numpy(), idx2=pi2.numpy())

    # (ii) ties and duplicates: coordinates on a coarse grid so exact ties are common
    rng = np.random.default_rng(20)
    ta = (rng.integers(0, 4, (3, 257, 3)).astype(np.float32) * np.float32(0.25))
    tb = (rng.integers(0, 4, (3, 131, 3)).astype(np.float32) * np.float32(0.25))
    d1, i1, d2, i2 = O.ref_nn_distance(ta, tb)
    gd1 = rng.standard_normal((3, 257)).astype(np.float32)
    gd2 = rng.standard_normal((3, 131)).astype(np.float32)
    g1, g2 = O.ref_nn_distance_grad(ta, tb, gd1, i1, gd2, i2)
    np.savez_compressed(os.path.join(HERE, "nnd_ties_3x257x131.npz"), xyz1=ta, xyz2=tb, dist1=d1, idx1=i1, dist2=d2,
                        idx2=i2, gd1=gd1, gd2=gd2, gxyz1=g1, gxyz2=g2)

    # (iii) config 1 of BASELINE.json: B=1, N=M=2048, U[-0.5,0.5), seeds 0/1 -> digests
    a = cloud(0, (1, 2048, 3))
    b = cloud(1, (1, 2048, 3))
    d1, i1, d2, i2 = O.ref_nn_distance(a, b)
    gd = np.full((1, 2048), 1.0 / 2048, np.float32)
    g1, g2 = O.ref_nn_distance_grad(a, b, gd, i1, gd, i2)
    out["cfg1"] = dict(seed1=0, seed2=1, dist1=sha(d1), idx1=sha(i1), dist2=sha(d2), idx2=sha(i2), gxyz1=sha(g1),
                       gxyz2=sha(g2), dist1_head=d1[0, :4].tolist(), idx1_head=i1[0, :4].tolist())
    # adversarial-like: xyz2 = xyz1 + N(0,1e-3)
    b2 = (a + np.random.default_rng(2).standard_normal(a.shape).astype(np.float32) * np.float32(1e-3)).astype(
        np.float32)
    d1, i1, d2, i2 = O.ref_nn_distance(a, b2)
    out["cfg1_adv"] = dict(dist1=sha(d1), idx1=sha(i1), dist2=sha(d2), idx2=sha(i2))
    # exact duplicate cloud
    d1, i1, d2, i2 = O.ref_nn_distance(a, a)
    out["cfg1_dup"] = dict(dist1=sha(d1), idx1=sha(i1), dist2=sha(d2), idx2=sha(i2))

    # (iv) config 2 shape, a slice of it: B=4, N=M=2048, seed 2
    a = cloud(2, (4, 2048, 3))
    b = cloud(3, (4, 2048, 3))
    d1, i1, d2, i2 = O.ref_nn_distance(a, b)
    gd1 = np.random.default_rng(4).standard_normal((4, 2048)).astype(np.float32)
    gd2 = np.random.default_rng(5).standard_normal((4, 2048)).astype(np.float32)
    g1, g2 = O.ref_nn_distance_grad(a, b, gd1, i1, gd2, i2)
    out["cfg2_b4"] = dict(dist1=sha(d1), idx1=sha(i1), dist2=sha(d2), idx2=sha(i2), gxyz1=sha(g1), gxyz2=sha(g2))

    # (v) selection sort known answer, parsed from the reference program's stdout
    with open(os.path.join(ROOT, "oracle/_ref/selection_sort_ref.out")) as f:
        lines = [l.strip() for l in f if l.strip()]
    idx_line = [int(x) for x in lines[-2].split()]
    val_line = [float(x) for x in lines[-1].split()]
    out["selection_sort"] = dict(b=2, n=4, m=2, k=3, dist=[float(10 - i) for i in range(16)], idx=idx_line,
                                 val=val_line)

    import json
    with open(os.path.join(HERE, "digests.json"), "w") as f:
        json.dump(out, f, indent=1, sort_keys=True)
    print("wrote fixtures to", HERE)


if __name__ == "__main__":
    main()
