"""A/B of the forward entry between builds of the library, interleaved in one process on one box (development tool).
usage: ab_fwd.py B lib1.so lib2.so ...   (paths relative to the repo root)"""
import ctypes, os, sys
import numpy as np
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
B = int(sys.argv[1]); N = 2048
libs = [(p, ctypes.CDLL(os.path.join(ROOT, p))) for p in sys.argv[2:]]
dev = torch.device("cuda:0"); p = ctypes.c_void_p; st = torch.cuda.current_stream().cuda_stream
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
g = torch.Generator().manual_seed(2)
x1 = (torch.rand(B, N, 3, generator=g) - 0.5).to(dev); x2 = (torch.rand(B, N, 3, generator=g) - 0.5).to(dev)
d1 = torch.empty(B, N, device=dev); i1 = torch.empty(B, N, dtype=torch.int32, device=dev)
d2 = torch.empty(B, N, device=dev); i2 = torch.empty(B, N, dtype=torch.int32, device=dev)
ts = {n: [] for n, _ in libs}
for rep in range(12):
    for n, lib in libs:
        for k in range(3):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            rc = lib.ga_nn_distance_fwd(B, N, N, p(x1.data_ptr()), p(x2.data_ptr()), p(d1.data_ptr()), p(i1.data_ptr()),
                                        p(d2.data_ptr()), p(i2.data_ptr()), 0, p(st))
            e1.record(); torch.cuda.synchronize()
            assert rc == 0
            if rep >= 2:
                ts[n].append(e0.elapsed_time(e1) * 1e3)
for n, _ in libs:
    t = np.array(ts[n])
    print("%-40s min %6.1f  median %6.1f  mean %6.1f us  (%d samples)" % (n, t.min(), np.median(t), t.mean(), len(t)), flush=True)
