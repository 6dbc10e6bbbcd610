"""kNN kernels at config 5 (development tool): fp32-filter variants vs the tensor-core kernel (tuning key 1)."""
import ctypes, os, sys
import numpy as np
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from geometric_adv_b200 import _lib  # noqa: E402
lib = _lib.load()
dev = torch.device("cuda:0"); p = ctypes.c_void_p; st = torch.cuda.current_stream().cuda_stream
for B in (500, 100, 63):
    g = torch.Generator().manual_seed(4)
    pc = (torch.rand(B, 2048, 3, generator=g) - 0.5).to(dev)
    out = torch.empty(B, 2048, 10, device=dev)
    ref = None
    for name, v in (("default", 0), ("fp32 256x2", 3), ("fp32 128x2", 9), ("mma", 6)):
        lib.ga_set_tuning(1, v)
        for _ in range(2):
            _lib.check(lib.ga_knn_dists(B, 2048, 10, p(pc.data_ptr()), p(out.data_ptr()), p(st)))
        torch.cuda.synchronize()
        ts = []
        for _ in range(10):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); _lib.check(lib.ga_knn_dists(B, 2048, 10, p(pc.data_ptr()), p(out.data_ptr()), p(st))); e1.record()
            torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
        if ref is None:
            ref = out.clone()
        print("B=%d %-12s min %.3f ms med %.3f ms  %s" % (B, name, min(ts), float(np.median(ts)), "same bits" if torch.equal(out, ref) else "DIFFERENT"), flush=True)
lib.ga_set_tuning(1, 0)
