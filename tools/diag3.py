"""Which ingredient of bench.py makes the forward kernel 2x slower?  One factor per process."""
import ctypes, os, subprocess, sys, threading, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from geometric_adv_b200 import _lib
flags = set(sys.argv[1:])
lib = _lib.load(); dev = torch.device("cuda:0"); p = ctypes.c_void_p
torch.cuda.set_device(0)
st = torch.cuda.current_stream().cuda_stream
B, N = 50, 2048
if "numpy" in flags:
    a = (np.random.default_rng(2).random((B, N, 3), dtype=np.float32) - np.float32(0.5)).astype(np.float32)
    b = (np.random.default_rng(3).random((B, N, 3), dtype=np.float32) - np.float32(0.5)).astype(np.float32)
    x1, x2 = torch.from_numpy(a).to(dev), torch.from_numpy(b).to(dev)
else:
    x1 = (torch.rand(B, N, 3) - 0.5).to(dev); x2 = (torch.rand(B, N, 3) - 0.5).to(dev)
if "thread" in flags:
    ev = threading.Event()
    threading.Thread(target=lambda: ev.wait(), daemon=True).start()
if "popen" in flags:
    pr = subprocess.Popen(["sleep", "5"])
if "probe" in flags:
    tf, ms = ctypes.c_float(), ctypes.c_float()
    lib.ga_probe_fp32_peak(8192, ctypes.byref(tf), ctypes.byref(ms), p(st))
    lf = ctypes.c_float(); lib.ga_probe_launch_floor(200, ctypes.byref(lf), p(st))
if "flushbuf" in flags:
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev); flush.zero_()
if "manyallocs" in flags:
    keep = [torch.empty(B, N, 3, device=dev) for _ in range(8)]
d1 = torch.empty(B, N, device=dev); i1 = torch.empty(B, N, dtype=torch.int32, device=dev)
d2 = torch.empty(B, N, device=dev); i2 = torch.empty(B, N, dtype=torch.int32, device=dev)
args = (B, N, N, p(x1.data_ptr()), p(x2.data_ptr()), p(d1.data_ptr()), p(i1.data_ptr()), p(d2.data_ptr()), p(i2.data_ptr()), 0, p(st))
for _ in range(10): lib.ga_nn_distance_fwd(*args)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(2000): lib.ga_nn_distance_fwd(*args)
e1.record(); torch.cuda.synchronize()
ptrs = [x.data_ptr() % 4096 for x in (x1, x2, d1, i1, d2, i2)]
print("%-28s %.4f ms/launch   ptr%%4096=%s  nnz(idx1==0)=%d" % (" ".join(sorted(flags)) or "baseline", e0.elapsed_time(e1) / 2000, ptrs, int((i1 == 0).sum())), flush=True)
