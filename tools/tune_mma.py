"""Timing of the tensor-core forward variant (20) against the fp32-filter kernel.  Development tool."""
import ctypes
import json
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


def timeit(fn, reps=40, do_flush=True):
    for _ in range(5):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        if do_flush:
            flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ts.sort()
    return {"min_ms": ts[0], "med_ms": ts[len(ts) // 2]}


out = {}
for (b, n, m) in [(50, 2048, 2048), (10, 2048, 2048), (512, 2048, 2048), (50, 2025, 2048), (4, 8192, 8192)]:
    g = torch.Generator(device="cpu").manual_seed(1)
    x1 = (torch.rand(b, n, 3, generator=g) - 0.5).to(dev)
    x2 = (torch.rand(b, m, 3, generator=g) - 0.5).to(dev)
    d1 = torch.empty(b, n, device=dev); i1 = torch.empty(b, n, dtype=torch.int32, device=dev)
    d2 = torch.empty(b, m, device=dev); i2 = torch.empty(b, m, dtype=torch.int32, device=dev)
    key = "fwd_b%d_n%d_m%d" % (b, n, m)
    out[key] = {}

    def call():
        lib.ga_nn_distance_fwd(b, n, m, p(x1.data_ptr()), p(x2.data_ptr()), p(d1.data_ptr()), p(i1.data_ptr()),
                               p(d2.data_ptr()), p(i2.data_ptr()), 0, p(st))

    for name, v, cfg in [("default", 0, 0), ("fp32_v1", 1, 0)] + [("mma_cfg%d" % c, 20, c) for c in (1, 4, 5)] + [("mma_persist", 21, 0), ("mma_balanced", 23, 0), ("umma", 22, 0)]:
        if v in (21, 22, 23) and max(n, m) > 2048:
            continue
        lib.ga_set_tuning(0, v)
        lib.ga_set_tuning(7, cfg)
        r = timeit(call)
        r["TFLOPs_8flop"] = 8.0 * b * n * m / (r["min_ms"] * 1e-3) / 1e12
        out[key][name] = r
        print(key, name, r, flush=True)
    lib.ga_set_tuning(0, 0)
    lib.ga_set_tuning(7, 0)
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
with open(os.path.join(ROOT, "gpurun_out", "tune_mma.json"), "w") as f:
    json.dump(out, f, indent=1)
