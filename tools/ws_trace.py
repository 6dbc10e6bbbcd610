"""Per-warp time budget of the warp-specialised forward kernel (variant 24): where scan and helper warps spend their
clocks.  usage: ws_trace.py [B] [scan warps]   (development tool)"""
import ctypes, os, sys
import numpy as np
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from geometric_adv_b200 import _lib  # noqa: E402
lib = _lib.load()
B = int(sys.argv[1]) if len(sys.argv) > 1 else 50
NS = int(sys.argv[2]) if len(sys.argv) > 2 else 8
N = 2048
dev = torch.device("cuda:0"); p = ctypes.c_void_p; st = torch.cuda.current_stream().cuda_stream
g = torch.Generator().manual_seed(2)
x1 = (torch.rand(B, N, 3, generator=g) - 0.5).to(dev); x2 = (torch.rand(B, N, 3, generator=g) - 0.5).to(dev)
d1 = torch.empty(B, N, device=dev); i1 = torch.empty(B, N, dtype=torch.int32, device=dev)
d2 = torch.empty(B, N, device=dev); i2 = torch.empty(B, N, dtype=torch.int32, device=dev)
lib.ga_set_tuning(0, 24); lib.ga_set_tuning(22, NS)
DEV = int(sys.argv[3]) if len(sys.argv) > 3 else 0
lib.ga_set_tuning(24, DEV)
def call():
    _lib.check(lib.ga_nn_distance_fwd(B, N, N, p(x1.data_ptr()), p(x2.data_ptr()), p(d1.data_ptr()), p(i1.data_ptr()),
                                      p(d2.data_ptr()), p(i2.data_ptr()), 0, p(st)))
for _ in range(3):
    call()
buf = torch.zeros(148 * 16 * 4, dtype=torch.int64, device=dev)
lib.ga_debug_ws_trace(p(buf.data_ptr()))
call(); torch.cuda.synchronize()
lib.ga_debug_ws_trace(None)
t = buf.cpu().numpy().reshape(148, 16, 4).astype(np.float64)
print("dev", DEV); print("B=%d, %d scan warps; clocks per warp (mean over CTAs / max), after the prologue" % (B, NS))
for name, sl in (("scan warps", slice(0, NS)), ("helper warps", slice(NS, 16))):
    w = t[:, sl, :]
    print("  %-12s wait %8.0f  work %8.0f  stage %8.0f  total %8.0f (max %8.0f)" % (
        name, w[..., 0].mean(), w[..., 1].mean(), w[..., 2].mean(), w[..., 3].mean(), w[..., 3].max()))
print("  prologue (entry -> roles begin): mean %.0f max %.0f clk" % (t[:, :NS, 2].mean(), t[:, :NS, 2].max()))
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); call(); e1.record(); torch.cuda.synchronize()
print("  kernel by events: %.1f us = %.0f clk at 1.965 GHz" % (e0.elapsed_time(e1) * 1e3, e0.elapsed_time(e1) * 1e3 * 1965))
jobs = B * 64 / 148.0
print("  jobs per SM %.1f -> scan work per job %.0f clk, refine work per job %.0f clk" % (
    jobs, t[:, :NS, 1].sum(1).mean() / jobs, t[:, NS:, 1].sum(1).mean() / jobs))
