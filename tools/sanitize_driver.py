"""Small driver that touches every kernel once (for compute-sanitizer; development tool)."""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import geometric_adv_b200 as ga
from geometric_adv_b200 import defense
dev = "cuda:0"
rng = np.random.default_rng(0)
def cl(b, n): return torch.from_numpy((rng.random((b, n, 3), dtype=np.float32) - 0.5).astype(np.float32)).to(dev)
for (b, n, m) in [(2, 300, 517), (1, 2048, 2048), (1, 2500, 100)]:
    x1, x2 = cl(b, n).requires_grad_(True), cl(b, m).requires_grad_(True)
    for pr in (False, True):
        ga.set_pruning(pr)
        d1, i1, d2, i2 = ga.nn_distance(x1, x2)
    ga.set_pruning(False)
    (d1.mean() + d2.mean()).backward()
g = torch.from_numpy((rng.integers(0, 3, (1, 300, 3)) * 0.5).astype(np.float32)).to(dev)  # ties -> cooperative paths
ga.nn_distance(g, g); ga.knn_point(5, g, g); ga.knn_point(40, g, g)
pc = cl(2, 700)
ga.knn_dists(pc, 10); v, i = ga.knn_point(11, pc, pc); ga.knn_point(17, pc, pc); ga.knn_point(30, cl(1, 2300), cl(1, 50))
ga.group_point(pc, i); ga.select_top_k(3, torch.rand(2, 9, 50, device=dev))
ga.chamfer_all_pairs(cl(5, 300)); ga.chamfer_per_cloud(d1.detach(), d2.detach())
defense.get_outlier_pc_inlier_pc(pc, torch.rand(2, 700, device=dev), 0.5)
h = torch.rand(2, 64, 3); ga.nn_distance(h, h); ga.knn_dists(h, 3)
# round-1 additions: tensor-core forward variants (HMMA grid / persistent, tcgen05), both gradient kernels,
# the graph replay of the host entry point
from geometric_adv_b200 import _lib
import ctypes
lib = _lib.load()
for v in (20, 21, 22, 23):
    lib.ga_set_tuning(0, v)
    for (b, n, m) in [(2, 300, 517), (3, 2048, 2048)]:
        ga.nn_distance(cl(b, n), cl(b, m))
lib.ga_set_tuning(0, 0)
x1, x2 = cl(3, 700).requires_grad_(True), cl(3, 900).requires_grad_(True)
for kern in (0, 1):
    lib.ga_set_tuning(14, kern)
    d1, i1, d2, i2 = ga.nn_distance(x1, x2)
    (d1.mean() + d2.mean()).backward()
lib.ga_set_tuning(14, 0)
b, n = 24, 1024
hb = [(torch.rand(b, n, 3) - 0.5).pin_memory(), (torch.rand(b, n, 3) - 0.5).pin_memory(), torch.rand(b, n).pin_memory(),
      torch.rand(b, n).pin_memory(), torch.empty(b, n).pin_memory(), torch.empty(b, n, dtype=torch.int32).pin_memory(),
      torch.empty(b, n).pin_memory(), torch.empty(b, n, dtype=torch.int32).pin_memory(),
      torch.empty(b, n, 3).pin_memory(), torch.empty(b, n, 3).pin_memory()]
for _ in range(3):
    _lib.check(lib.ga_nn_distance_fwd_bwd_host(b, n, n, *[ctypes.c_void_p(x.data_ptr()) for x in hb], 0))
# one-call forward + gradient (completion tickets, early gradient CTAs), tensor-core all-pairs
b, n = 20, 2048
xa, xb = cl(b, n), cl(b, n)
ga1 = torch.rand(b, n, device=dev); ga2 = torch.rand(b, n, device=dev)
od1 = torch.empty(b, n, device=dev); oi1 = torch.empty(b, n, dtype=torch.int32, device=dev)
od2 = torch.empty(b, n, device=dev); oi2 = torch.empty(b, n, dtype=torch.int32, device=dev)
og1 = torch.empty(b, n, 3, device=dev); og2 = torch.empty(b, n, 3, device=dev)
P = ctypes.c_void_p
for _ in range(2):
    _lib.check(lib.ga_nn_distance_fwd_bwd(b, n, n, P(xa.data_ptr()), P(xb.data_ptr()), P(ga1.data_ptr()), P(ga2.data_ptr()),
                                          P(od1.data_ptr()), P(oi1.data_ptr()), P(od2.data_ptr()), P(oi2.data_ptr()),
                                          P(og1.data_ptr()), P(og2.data_ptr()), 0, P(torch.cuda.current_stream().cuda_stream)))
ga.chamfer_all_pairs(cl(6, 512))
# round-2 additions: gradient kernel 3, frame kernel (clouds away from the origin), warp-specialised HMMA forward,
# tensor-core kNN, slab kNN (plain, clustered -> overflow service, planar -> full-scan body), fused loss terms,
# GPU sort_dist_mat pipeline, streamed and pulled ingest of the host step (incl. the give-up path)
x1, x2 = cl(3, 700).requires_grad_(True), cl(3, 900).requires_grad_(True)
lib.ga_set_tuning(14, 3)
d1, i1, d2, i2 = ga.nn_distance(x1, x2)
(d1.mean() + d2.mean()).backward()
lib.ga_set_tuning(14, 0)
lib.ga_set_tuning(25, 1)
ga.nn_distance(cl(20, 2048) + 7.0, cl(20, 2048) + 7.0)
lib.ga_set_tuning(25, 2)
lib.ga_set_tuning(0, 24)
ga.nn_distance(cl(3, 2048), cl(3, 2048))
lib.ga_set_tuning(0, 0)
lib.ga_set_tuning(1, 6)
ga.knn_point(11, cl(2, 1024), cl(2, 1024)); ga.knn_dists(cl(2, 1024), 10)
lib.ga_set_tuning(1, 0)
lib.ga_set_tuning(28, 2)
ga.knn_dists(cl(3, 2048), 10); ga.knn_dists(cl(2, 1000), 4)
blob = (torch.randn(2, 2048, 3) * 0.01 + torch.randint(0, 3, (2, 2048, 1)).float() * 0.5).to(dev)
ga.knn_dists(blob, 10)
flat = cl(1, 2048); flat[..., 0] = 0.25
ga.knn_dists(flat, 10)
lib.ga_set_tuning(28, 1)
da, _, db, _ = ga.nn_distance(cl(4, 600), cl(4, 500))
ga.chamfer_loss_terms(cl(4, 600), cl(4, 500))
from geometric_adv_b200 import sharding
sharding.prepare_indices(cl(12, 256), [0, 5, 12])
b, n = 40, 1024
hb = [(torch.rand(b, n, 3) - 0.5).pin_memory(), (torch.rand(b, n, 3) - 0.5).pin_memory(), torch.rand(b, n).pin_memory(),
      torch.rand(b, n).pin_memory(), torch.empty(b, n).pin_memory(), torch.empty(b, n, dtype=torch.int32).pin_memory(),
      torch.empty(b, n).pin_memory(), torch.empty(b, n, dtype=torch.int32).pin_memory(),
      torch.empty(b, n, 3).pin_memory(), torch.empty(b, n, 3).pin_memory()]
lib.ga_set_tuning(0, 20)
for key, val in ((26, 3), (27, 4), (29, 1)):
    lib.ga_set_tuning(26, 0); lib.ga_set_tuning(27, 4 if key == 29 else 0); lib.ga_set_tuning(29, 0)
    lib.ga_set_tuning(key, val)
    for _ in range(3):
        _lib.check(lib.ga_nn_distance_fwd_bwd_host(b, n, n, *[ctypes.c_void_p(x.data_ptr()) for x in hb], 0))
    print("host pipeline", key, val, "streamed", lib.ga_debug_host_streamed())
lib.ga_set_tuning(26, 0); lib.ga_set_tuning(27, 0); lib.ga_set_tuning(29, 0); lib.ga_set_tuning(0, 0)
torch.cuda.synchronize(); print("driver ok")
