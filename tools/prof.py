"""Launch each hot kernel a few times (for ncu).  usage: prof.py [fwd|bwd|knn|all] [B]"""
import ctypes
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from geometric_adv_b200 import _lib  # noqa: E402

what = sys.argv[1] if len(sys.argv) > 1 else "all"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 50
N = 2048
lib = _lib.load()
for kv in os.environ.get("GA_TUNE", "").split(","):  # e.g. GA_TUNE=0=20,7=5
    if "=" in kv:
        lib.ga_set_tuning(int(kv.split("=")[0]), int(kv.split("=")[1]))
dev = torch.device("cuda:0")
p = ctypes.c_void_p
st = torch.cuda.current_stream().cuda_stream
g = torch.Generator().manual_seed(0)
x1 = (torch.rand(B, N, 3, generator=g) - 0.5).to(dev)
x2 = (torch.rand(B, N, 3, generator=g) - 0.5).to(dev)
d1 = torch.empty(B, N, device=dev); i1 = torch.empty(B, N, dtype=torch.int32, device=dev)
d2 = torch.empty(B, N, device=dev); i2 = torch.empty(B, N, dtype=torch.int32, device=dev)
g1 = torch.full((B, N), 1.0 / N, device=dev)
o1 = torch.empty(B, N, 3, device=dev); o2 = torch.empty(B, N, 3, device=dev)
kd = torch.empty(B, N, 10, device=dev)
for _ in range(3):
    if what in ("fwd", "bwd", "all"):
        _lib.check(lib.ga_nn_distance_fwd(B, N, N, p(x1.data_ptr()), p(x2.data_ptr()), p(d1.data_ptr()),
                                          p(i1.data_ptr()), p(d2.data_ptr()), p(i2.data_ptr()), 0, p(st)))
    if what in ("bwd", "all"):
        _lib.check(lib.ga_nn_distance_bwd(B, N, N, p(x1.data_ptr()), p(x2.data_ptr()), p(g1.data_ptr()),
                                          p(i1.data_ptr()), p(g1.data_ptr()), p(i2.data_ptr()), p(o1.data_ptr()),
                                          p(o2.data_ptr()), p(st)))
    if what == "sorted":
        wsb = lib.ga_nn_distance_workspace_bytes(B, N, N)
        ws = torch.empty(wsb, dtype=torch.uint8, device=dev)
        _lib.check(lib.ga_nn_distance_fwd_ws(B, N, N, p(x1.data_ptr()), p(x2.data_ptr()), p(d1.data_ptr()),
                                             p(i1.data_ptr()), p(d2.data_ptr()), p(i2.data_ptr()), 0,
                                             p(ws.data_ptr()), wsb, p(st)))
    if what in ("knn", "all"):
        _lib.check(lib.ga_knn_dists(B, N, 10, p(x1.data_ptr()), p(kd.data_ptr()), p(st)))
torch.cuda.synchronize()
print("done", what, B)
