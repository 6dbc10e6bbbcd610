"""Wave structure of nn_fwd_mma_kernel: time of B = 18 (144 CTAs: at most one per SM), 37 (296 CTAs: exactly
two per SM), 50 and 74, for the grid configurations 5 (two CTAs per SM) and 2 (one CTA per SM).  Development tool."""
import ctypes, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from geometric_adv_b200 import _lib  # noqa: E402
lib = _lib.load()
dev = torch.device("cuda:0"); p = ctypes.c_void_p; st = torch.cuda.current_stream().cuda_stream
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
N = 2048
for b in (9, 18, 37, 50, 55, 74, 111):
    g = torch.Generator().manual_seed(1)
    x1 = (torch.rand(b, N, 3, generator=g) - 0.5).to(dev); x2 = (torch.rand(b, N, 3, generator=g) - 0.5).to(dev)
    d1 = torch.empty(b, N, device=dev); i1 = torch.empty(b, N, dtype=torch.int32, device=dev)
    d2 = torch.empty(b, N, device=dev); i2 = torch.empty(b, N, dtype=torch.int32, device=dev)
    row = []
    for cfg in (5, 2):
        lib.ga_set_tuning(0, 20); lib.ga_set_tuning(7, cfg)
        call = lambda: lib.ga_nn_distance_fwd(b, N, N, p(x1.data_ptr()), p(x2.data_ptr()), p(d1.data_ptr()), p(i1.data_ptr()), p(d2.data_ptr()), p(i2.data_ptr()), 0, p(st))
        for _ in range(3): call()
        ts = []
        for _ in range(30):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); call(); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1) * 1e3)
        ts.sort(); row.append("cfg%d min %.1f med %.1f us" % (cfg, ts[0], ts[len(ts) // 2]))
    lib.ga_set_tuning(0, 0); lib.ga_set_tuning(7, 0)
    print("B=%3d (%4d CTAs): %s" % (b, b * 8, " | ".join(row)), flush=True)
