"""Timing table for kernel variants (development tool; run on the GPU box).
Writes gpurun_out/tune.json.  Not part of the product or the tests."""
import ctypes
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import geometric_adv_b200 as ga  # noqa: E402
from geometric_adv_b200 import _lib  # noqa: E402

lib = _lib.load()
dev = torch.device("cuda:0")
p = ctypes.c_void_p
st = torch.cuda.current_stream().cuda_stream
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)


def timeit(fn, reps=30, do_flush=True):
    for _ in range(3):
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


def clouds(b, n, seed):
    g = torch.Generator(device="cpu").manual_seed(seed)
    return (torch.rand(b, n, 3, generator=g) - 0.5).to(dev)


out = {}
tf, ms = ctypes.c_float(), ctypes.c_float()
lib.ga_probe_fp32_peak(8192, ctypes.byref(tf), ctypes.byref(ms), p(st))
out["fp32_peak_tflops"] = tf.value
lf = ctypes.c_float()
lib.ga_probe_launch_floor(500, ctypes.byref(lf), p(st))
out["launch_floor_us"] = lf.value
print("fp32 peak", tf.value, "TF/s; launch floor", lf.value, "us", flush=True)

ref_gpu = None
rp = os.path.join(ROOT, "oracle", "_ref", "libga_ref_gpu.so")
if os.path.exists(rp):
    ref_gpu = ctypes.CDLL(rp)

for (b, n, m) in [(50, 2048, 2048), (10, 2048, 2048), (1, 2048, 2048), (512, 2048, 2048)]:
    x1, x2 = clouds(b, n, 1), clouds(b, m, 2)
    d1 = torch.empty(b, n, device=dev); i1 = torch.empty(b, n, dtype=torch.int32, device=dev)
    d2 = torch.empty(b, m, device=dev); i2 = torch.empty(b, m, dtype=torch.int32, device=dev)
    g1 = torch.full((b, n), 1.0 / n, device=dev); g2 = torch.full((b, m), 1.0 / m, device=dev)
    o1 = torch.empty(b, n, 3, device=dev); o2 = torch.empty(b, m, 3, device=dev)
    pairs = float(b) * n * m
    key = "fwd_b%d" % b
    out[key] = {}
    for v in range(0, 16):
        lib.ga_set_tuning(0, v)
        r = timeit(lambda: lib.ga_nn_distance_fwd(b, n, m, p(x1.data_ptr()), p(x2.data_ptr()), p(d1.data_ptr()),
                                                  p(i1.data_ptr()), p(d2.data_ptr()), p(i2.data_ptr()), 0, p(st)))
        r["TFLOPs_8flop"] = 8 * pairs / (r["min_ms"] * 1e-3) / 1e12
        out[key]["variant%d" % v] = r
        print(key, "variant", v, r, flush=True)
    lib.ga_set_tuning(0, 0)
    r = timeit(lambda: lib.ga_nn_distance_fwd(b, n, m, p(x1.data_ptr()), p(x2.data_ptr()), p(d1.data_ptr()),
                                              p(i1.data_ptr()), p(d2.data_ptr()), p(i2.data_ptr()), 1, p(st)))
    out[key]["variant0_mode1"] = r
    if b <= 50:
        for qq in (2, 4):
            for S in (1, 2, 4, 8):
                lib.ga_set_tuning(5, S if S > 1 else 0); lib.ga_set_tuning(6, qq)
                r = timeit(lambda: lib.ga_nn_distance_fwd(b, n, m, p(x1.data_ptr()), p(x2.data_ptr()),
                                                          p(d1.data_ptr()), p(i1.data_ptr()), p(d2.data_ptr()),
                                                          p(i2.data_ptr()), 0, p(st)))
                out[key]["split_q%d_S%d" % (qq, S)] = r
                print(key, "split q", qq, "S", S, r, flush=True)
        lib.ga_set_tuning(5, -1); lib.ga_set_tuning(6, 0)
    wsb = lib.ga_nn_distance_workspace_bytes(b, n, m)
    ws = torch.empty(max(wsb, 16), dtype=torch.uint8, device=dev)
    for v in range(5):
        lib.ga_set_tuning(4, v)
        r = timeit(lambda: lib.ga_nn_distance_fwd_ws(b, n, m, p(x1.data_ptr()), p(x2.data_ptr()), p(d1.data_ptr()),
                                                     p(i1.data_ptr()), p(d2.data_ptr()), p(i2.data_ptr()), 0,
                                                     p(ws.data_ptr()), wsb, p(st)))
        r["TFLOPs_8flop"] = 8 * pairs / (r["min_ms"] * 1e-3) / 1e12
        out[key]["sorted_variant%d" % v] = r
        print(key, "sorted (pruned) variant", v, r, flush=True)
    lib.ga_set_tuning(4, 0)
    r = timeit(lambda: lib.ga_nn_distance_bwd(b, n, m, p(x1.data_ptr()), p(x2.data_ptr()), p(g1.data_ptr()),
                                              p(i1.data_ptr()), p(g2.data_ptr()), p(i2.data_ptr()),
                                              p(o1.data_ptr()), p(o2.data_ptr()), p(st)))
    r["GBps_32B_per_point"] = 32.0 * b * (n + m) / (r["min_ms"] * 1e-3) / 1e9
    out["bwd_b%d" % b] = r
    print("bwd b", b, r, flush=True)
    if ref_gpu is not None and b <= 50:
        r = timeit(lambda: ref_gpu.ga_refgpu_nn_distance(b, n, m, p(x1.data_ptr()), p(x2.data_ptr()),
                                                         p(d1.data_ptr()), p(i1.data_ptr()), p(d2.data_ptr()),
                                                         p(i2.data_ptr())))
        out["refgpu_fwd_b%d" % b] = r
        print("reference CUDA kernel fwd b", b, r, flush=True)
        r = timeit(lambda: ref_gpu.ga_refgpu_nn_distance_grad(b, n, m, p(x1.data_ptr()), p(x2.data_ptr()),
                                                              p(g1.data_ptr()), p(i1.data_ptr()), p(g2.data_ptr()),
                                                              p(i2.data_ptr()), p(o1.data_ptr()), p(o2.data_ptr())))
        out["refgpu_bwd_b%d" % b] = r
        print("reference CUDA kernel bwd b", b, r, flush=True)

for (b, n, k) in [(100, 2048, 10), (500, 2048, 10)]:
    pc = clouds(b, n, 4)
    o = torch.empty(b, n, k, device=dev)
    for v in range(6):
        lib.ga_set_tuning(1, v)
        r = timeit(lambda: lib.ga_knn_dists(b, n, k, p(pc.data_ptr()), p(o.data_ptr()), p(st)), reps=10)
        r["pairs_per_s"] = float(b) * n * n / (r["min_ms"] * 1e-3)
        out["knn_dists_b%d_v%d" % (b, v)] = r
        print("knn_dists", b, "variant", v, r, flush=True)
    lib.ga_set_tuning(1, 0)
    val = torch.empty(b, n, k + 1, device=dev); idx = torch.empty(b, n, k + 1, dtype=torch.int32, device=dev)
    r = timeit(lambda: lib.ga_knn(b, n, n, k + 1, p(pc.data_ptr()), p(pc.data_ptr()), p(val.data_ptr()),
                                  p(idx.data_ptr()), p(st)), reps=10)
    out["knn_point_b%d" % b] = r
    print("knn_point", b, r, flush=True)

os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
with open(os.path.join(ROOT, "gpurun_out", "tune.json"), "w") as f:
    json.dump(out, f, indent=1)
