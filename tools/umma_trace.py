"""Timeline of CTA 0 of the tcgen05 forward kernel (ga_debug_umma_trace).  Development tool.
Trace layout: see the comment at the top of nn_fwd_umma_kernel."""
import ctypes, os, sys
import numpy as np
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from geometric_adv_b200 import _lib  # noqa: E402
lib = _lib.load()
dev = torch.device("cuda:0"); p = ctypes.c_void_p; st = torch.cuda.current_stream().cuda_stream
B = int(sys.argv[1]) if len(sys.argv) > 1 else 50
modes = [int(x) for x in sys.argv[2].split(",")] if len(sys.argv) > 2 else [0, 1, 3]
cta = int(sys.argv[3]) if len(sys.argv) > 3 else 0
N = 2048
g = torch.Generator().manual_seed(2)
x1 = (torch.rand(B, N, 3, generator=g) - 0.5).to(dev); x2 = (torch.rand(B, N, 3, generator=g) - 0.5).to(dev)
d1 = torch.empty(B, N, device=dev); i1 = torch.empty(B, N, dtype=torch.int32, device=dev)
d2 = torch.empty(B, N, device=dev); i2 = torch.empty(B, N, dtype=torch.int32, device=dev)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
def rel(a, t0):  # 32-bit clock stamps relative to t0 (wrap-safe)
    return ((a.astype(np.int64) - int(t0)) & 0xffffffff).astype(np.int64)
for dbg in modes:
    lib.ga_set_tuning(20, dbg | (cta << 8))
    tr = torch.zeros(8192, dtype=torch.int64, device=dev)
    for _ in range(2):
        tr.zero_(); flush.zero_()
        _lib.check(lib.ga_debug_umma_trace(B, N, N, p(x1.data_ptr()), p(x2.data_ptr()), p(d1.data_ptr()), p(i1.data_ptr()),
                                           p(d2.data_ptr()), p(i2.data_ptr()), p(tr.data_ptr()), p(st)))
        torch.cuda.synchronize()
    t = tr.cpu().numpy()
    t0 = t[704]
    valid = lambda a: a != 0
    iss = t[0:256].reshape(-1, 2); scan = t[256:512].reshape(-1, 2); hlp = t[512:640].reshape(-1, 2)
    stg = t[640:704].reshape(-1, 2); hb = t[768:1280].reshape(4, 128); fd = t[1280:1408]
    ns = int(valid(iss[:, 1]).sum()); nk = int(valid(hlp[:, 1]).sum())
    iss_r = rel(iss[:ns], t0); scan_r = rel(scan[:ns], t0); hlp_r = rel(hlp[:nk], t0); hb_r = rel(hb[:, :ns], t0); fd_r = rel(fd[:ns], t0)
    print("dbg %d: roles begin +%d, steps %d, jobs %d" % (dbg, rel(t[705:706], t0)[0], ns, nk))
    print("  issuer: first issue +%d, last issue +%d, mean issue period %.0f" % (iss_r[0, 1], iss_r[-1, 1], (iss_r[-1, 1] - iss_r[0, 1]) / max(ns - 1, 1)))
    per = np.diff(iss_r[:, 1])
    print("  issue period by step mod 8:", " ".join("%d" % per[i::8].mean() for i in range(8)))
    wait = iss_r[:, 1] - iss_r[:, 0]
    print("  issuer wait (pre-wait -> issued): median %d" % np.median(wait))
    print("  scan warp 0 per step: full seen -> handed back median %d; handed back -> folded median %d; folded -> next full seen median %d, by step mod 8: %s"
          % (np.median(scan_r[:, 1] - scan_r[:, 0]), np.median(fd_r - scan_r[:, 1]), np.median(scan_r[1:, 0] - fd_r[:-1]),
             " ".join("%d" % (scan_r[1:, 0] - fd_r[:-1])[i::8].mean() for i in range(8))))
    lag = scan_r[:, 0] - iss_r[:, 1]
    print("  issue -> full seen by warp 0: median %d min %d" % (np.median(lag), lag.min()))
    skew = hb_r.max(axis=0) - hb_r.min(axis=0)
    print("  hand-back skew over the four column quarters (warps 0,4,8,12): median %d max %d" % (np.median(skew), skew.max()))
    print("  last hand-back (slowest of the four) -> issue of step s + 2: median %d" % np.median(iss_r[2:, 1] - hb_r.max(axis=0)[:-2]))
    print("  last hand-back +%d" % hb_r.max())
    print("  helper: keys seen -> refined: median %d max %d; last refined +%d" % (np.median(hlp_r[:, 1] - hlp_r[:, 0]), (hlp_r[:, 1] - hlp_r[:, 0]).max(), hlp_r[-1, 1]))
    for c in range(1, 32):
        if stg[c, 1] != 0:
            s_r = rel(stg[c], t0)
            print("  stager cloud %d: +%d .. +%d (%d)" % (c, s_r[0], s_r[1], s_r[1] - s_r[0]))
    print("  first 10 steps: issued / full seen / handed back / folded (rel. to TMEM allocated)")
    for s in range(min(ns, 10)):
        print("   %3d  %7d %7d %7d %7d" % (s, iss_r[s, 1], scan_r[s, 0], scan_r[s, 1], fd_r[s]))
    c = t[4096:4096 + 4 * 148].reshape(148, 4)
    z = c[:, 0].min()
    print("  all CTAs (us from the first entry): entry max %.1f | tmem allocated median %.1f max %.1f | roles begin median %.1f | exit min %.1f median %.1f max %.1f"
          % ((c[:, 0].max() - z) / 1e3, np.median(c[:, 1] - z) / 1e3, (c[:, 1].max() - z) / 1e3, np.median(c[:, 2] - z) / 1e3,
             (c[:, 3].min() - z) / 1e3, np.median(c[:, 3] - z) / 1e3, (c[:, 3].max() - z) / 1e3))
    busy = (c[:, 3] - c[:, 2]) / 1e3
    print("  CTA busy time (roles begin -> exit) us: min %.1f median %.1f max %.1f; traced CTA %.1f" % (busy.min(), np.median(busy), busy.max(), busy[cta]))
    print("  issue period, steps 0..127:", " ".join("%d" % x for x in per))
lib.ga_set_tuning(20, 0)
