"""Forward time against the clouds' distance from the origin (the filter window scales with (max|q| + max|t|)^2).
Development tool.  usage: tune_offset.py [B]"""
import ctypes, os, sys
import numpy as np
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from geometric_adv_b200 import _lib  # noqa: E402
lib = _lib.load()
B = int(sys.argv[1]) if len(sys.argv) > 1 else 50
N = 2048
dev = torch.device("cuda:0"); p = ctypes.c_void_p; st = torch.cuda.current_stream().cuda_stream
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
g = torch.Generator().manual_seed(2)
a = (torch.rand(B, N, 3, generator=g) - 0.5); b = (torch.rand(B, N, 3, generator=g) - 0.5)
d1 = torch.empty(B, N, device=dev); i1 = torch.empty(B, N, dtype=torch.int32, device=dev)
d2 = torch.empty(B, N, device=dev); i2 = torch.empty(B, N, dtype=torch.int32, device=dev)
print("B=%d, N=M=%d, unit cubes shifted by `offset` along every axis; key 25 = 0 plain kernel / 1 frame kernel / 2 default" % (B, N))
for off in (0.0, 0.5, 1.0, 2.0, 5.0, 10.0, 100.0, 1000.0):
    x1 = (a + off).to(dev); x2 = (b + off).to(dev)
    def call():
        _lib.check(lib.ga_nn_distance_fwd(B, N, N, p(x1.data_ptr()), p(x2.data_ptr()), p(d1.data_ptr()), p(i1.data_ptr()),
                                          p(d2.data_ptr()), p(i2.data_ptr()), 0, p(st)))
    line = "  offset %7.1f " % off
    for key, what in ((0, "plain"), (1, "frame"), (2, "auto")):
        lib.ga_set_tuning(25, key)
        for _ in range(3):
            call()
            torch.cuda.synchronize()
        ts = []
        for _ in range(10):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); call(); e1.record(); torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1) * 1e3)
        line += " | %s min %7.1f med %7.1f us" % (what, min(ts), float(np.median(ts)))
        if key == 2:
            line += "  (%s)" % lib.ga_last_kernel().decode()
    print(line, flush=True)
lib.ga_set_tuning(25, 2)
