"""Timing of the gradient kernel's two launch shapes (tuning key 9).  Development tool."""
import ctypes
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from geometric_adv_b200 import _lib  # noqa: E402

lib = _lib.load()
dev = torch.device("cuda:0")
p = ctypes.c_void_p
st = torch.cuda.current_stream().cuda_stream
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
for b in (1, 10, 50, 512, 4096):
    n = m = 2048
    g = torch.Generator(device="cpu").manual_seed(1)
    x1 = (torch.rand(b, n, 3, generator=g) - 0.5).to(dev)
    x2 = (torch.rand(b, m, 3, generator=g) - 0.5).to(dev)
    d1 = torch.empty(b, n, device=dev); i1 = torch.empty(b, n, dtype=torch.int32, device=dev)
    d2 = torch.empty(b, m, device=dev); i2 = torch.empty(b, m, dtype=torch.int32, device=dev)
    lib.ga_nn_distance_fwd(b, n, m, p(x1.data_ptr()), p(x2.data_ptr()), p(d1.data_ptr()), p(i1.data_ptr()),
                           p(d2.data_ptr()), p(i2.data_ptr()), 0, p(st))
    g1 = torch.full((b, n), 1.0 / n, device=dev)
    o1 = torch.empty(b, n, 3, device=dev); o2 = torch.empty(b, m, 3, device=dev)

    def call():
        lib.ga_nn_distance_bwd(b, n, m, p(x1.data_ptr()), p(x2.data_ptr()), p(g1.data_ptr()), p(i1.data_ptr()),
                               p(g1.data_ptr()), p(i2.data_ptr()), p(o1.data_ptr()), p(o2.data_ptr()), p(st))

    ref = None
    for kernel, split, stage in ((0, -1, 1), (3, -1, 1), (2, -1, 1)):
        lib.ga_set_tuning(14, kernel)
        lib.ga_set_tuning(9, split)
        lib.ga_set_tuning(13, stage)
        for _ in range(5):
            call()
        ts = []
        for _ in range(20):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); call(); e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        ts.sort()
        if ref is None:
            ref = (o1.clone(), o2.clone())
        same = torch.equal(o1, ref[0]) and torch.equal(o2, ref[1])
        gbs = 32.0 * b * (n + m) / (ts[0] * 1e-3) / 1e9
        print("bwd B=%d kernel=%d split=%d min %.2f us med %.2f us  %.0f GB/s algorithmic  %s" % (b, kernel, split, ts[0] * 1e3, ts[len(ts) // 2] * 1e3, gbs, "same bits" if same else "DIFFERENT"), flush=True)
    lib.ga_set_tuning(9, -1)
    lib.ga_set_tuning(13, 1)
    lib.ga_set_tuning(14, 0)
