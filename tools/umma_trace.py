"""Timeline of CTA 0 of the tcgen05 forward kernel (ga_debug_umma_trace).  Development tool."""
import ctypes
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from geometric_adv_b200 import _lib  # noqa: E402

lib = _lib.load()
fn = lib.ga_debug_umma_trace
fn.restype = ctypes.c_int
p = ctypes.c_void_p
B = int(sys.argv[1]) if len(sys.argv) > 1 else 50
N = 2048
dev = torch.device("cuda:0")
g = torch.Generator().manual_seed(0)
x1 = (torch.rand(B, N, 3, generator=g) - 0.5).to(dev)
x2 = (torch.rand(B, N, 3, generator=g) - 0.5).to(dev)
d1 = torch.empty(B, N, device=dev); i1 = torch.empty(B, N, dtype=torch.int32, device=dev)
d2 = torch.empty(B, N, device=dev); i2 = torch.empty(B, N, dtype=torch.int32, device=dev)
tr = torch.zeros(4 * 4096, dtype=torch.int64, device=dev)
st = torch.cuda.current_stream().cuda_stream
for _ in range(3):
    tr.zero_()
    rc = fn(ctypes.c_int(B), ctypes.c_int(N), ctypes.c_int(N), p(x1.data_ptr()), p(x2.data_ptr()), p(d1.data_ptr()),
            p(i1.data_ptr()), p(d2.data_ptr()), p(i2.data_ptr()), p(tr.data_ptr()), p(st))
    assert rc == 0, lib.ga_last_error()
    torch.cuda.synchronize()
t = tr.cpu().numpy().reshape(4, 2048, 2)
ph = t[3]
t0 = ph[0, 0]
print("segments (cycles from kernel start): start, staged | scanned, refined")
k = 0
while ph[2 * k, 0] != 0 or k == 0:
    print("  seg %d: start %7d staged %7d scanned %7d refined %7d" % (
        k, ph[2 * k, 0] - t0, ph[2 * k, 1] - t0, ph[2 * k + 1, 0] - t0, ph[2 * k + 1, 1] - t0))
    k += 1
    if 2 * k + 1 >= 2048:
        break
iss = t[0]
n = int((iss[:, 1] != 0).sum())
print("issuer: %d steps" % n)
for s in list(range(0, min(n, 40))):
    d0 = t[1][s]
    print("  step %3d: empty-wait done %7d  issued %7d (+%d) | drain warp 0 (%d): full-wait start %7d done %7d (waited %d, issue->visible %d)" % (
        s, iss[s, 0] - t0, iss[s, 1] - t0, iss[s, 1] - iss[s, 0], s & 1, d0[0] - t0 if d0[0] else -1,
        d0[1] - t0 if d0[1] else -1, d0[1] - d0[0], d0[1] - iss[s, 1] if d0[1] else -1))
gaps = np.diff(iss[:n, 1])
print("issue-to-issue gap: median %.0f mean %.0f max %d cycles" % (np.median(gaps), gaps.mean(), gaps.max()))
w = []
for s in range(4, n):
    d0 = t[1][s]
    if d0[0] and d0[1]:
        w.append((d0[1] - d0[0], d0[1] - iss[s, 1]))
w = np.array(w)
print("drain full-wait: median %.0f mean %.0f cycles; issue->visible-to-drain: median %.0f min %d" % (
    np.median(w[:, 0]), w[:, 0].mean(), np.median(w[:, 1]), w[:, 1].min()))
