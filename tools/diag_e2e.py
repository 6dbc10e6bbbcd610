"""e2e host path breakdown (development tool)."""
import ctypes, os, sys, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from geometric_adv_b200 import _lib
lib = _lib.load(); p = ctypes.c_void_p
B, N = 50, 2048
pin = lambda x: x.pin_memory()
h = [pin(torch.rand(B, N, 3) - 0.5), pin(torch.rand(B, N, 3) - 0.5), pin(torch.full((B, N), 1.0 / N)), pin(torch.full((B, N), 1.0 / N))]
d1 = pin(torch.empty(B, N)); i1 = pin(torch.empty(B, N, dtype=torch.int32)); d2 = pin(torch.empty(B, N)); i2 = pin(torch.empty(B, N, dtype=torch.int32))
o1 = pin(torch.empty(B, N, 3)); o2 = pin(torch.empty(B, N, 3))
args = (B, N, N, p(h[0].data_ptr()), p(h[1].data_ptr()), p(h[2].data_ptr()), p(h[3].data_ptr()), p(d1.data_ptr()), p(i1.data_ptr()), p(d2.data_ptr()), p(i2.data_ptr()), p(o1.data_ptr()), p(o2.data_ptr()), 0)
torch.cuda.init(); torch.zeros(1, device="cuda")
for path in (0, 1, 2, 1, 2):
    lib.ga_set_tuning(2, path)
    for _ in range(5): lib.ga_nn_distance_fwd_bwd_host(*args)
    ts = []
    for _ in range(30):
        t0 = time.perf_counter(); rc = lib.ga_nn_distance_fwd_bwd_host(*args); ts.append(time.perf_counter() - t0)
    ts.sort()
    print("host path %d: rc=%d min %.1f us  med %.1f us" % (path, rc, ts[0] * 1e6, ts[15] * 1e6), flush=True)
lib.ga_set_tuning(2, 1)
for ch in (1, 2, 3, 4, 5, 2, 1):
    lib.ga_set_tuning(3, ch)
    for _ in range(5): lib.ga_nn_distance_fwd_bwd_host(*args)
    ts = []
    for _ in range(30):
        t0 = time.perf_counter(); rc = lib.ga_nn_distance_fwd_bwd_host(*args); ts.append(time.perf_counter() - t0)
    ts.sort()
    print("copy path, %d chunks: min %.1f us  med %.1f us" % (ch, ts[0] * 1e6, ts[15] * 1e6), flush=True)
lib.ga_set_tuning(3, 0)
lib.ga_set_tuning(2, 0)
sys.exit(0)
# ---- device-side stage times of the zero-copy building blocks ----
dev = torch.device("cuda")
st = torch.cuda.current_stream().cuda_stream
dx1 = torch.empty(B, N, 3, device=dev); dx2 = torch.empty(B, N, 3, device=dev)
dd1 = torch.empty(B, N, device=dev); di1 = torch.empty(B, N, dtype=torch.int32, device=dev)
dd2 = torch.empty(B, N, device=dev); di2 = torch.empty(B, N, dtype=torch.int32, device=dev)
dg = torch.full((B, N), 1.0 / N, device=dev); do1 = torch.empty(B, N, 3, device=dev); do2 = torch.empty(B, N, 3, device=dev)
lib.ga_debug_ingest.argtypes = [p, p, ctypes.c_size_t, p]
lib.ga_debug_fwd_mirrored.argtypes = [ctypes.c_int] * 3 + [p] * 11
def timeit(name, fn, reps=20):
    for _ in range(3): fn()
    torch.cuda.synchronize(); ts = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1) * 1e3)
    ts.sort(); print("%-46s min %.1f us med %.1f us" % (name, ts[0], ts[reps // 2]), flush=True)
nb = B * N * 12
timeit("ingest kernel 1.23 MB host->device", lambda: lib.ga_debug_ingest(p(h[0].data_ptr()), p(dx1.data_ptr()), nb, p(st)))
timeit("cudaMemcpyAsync 1.23 MB h2d", lambda: dx1.copy_(h[0], non_blocking=True))
timeit("fwd, device outputs only", lambda: lib.ga_debug_fwd_mirrored(B, N, N, p(dx1.data_ptr()), p(dx2.data_ptr()), p(dd1.data_ptr()), p(di1.data_ptr()), p(dd2.data_ptr()), p(di2.data_ptr()), None, None, None, None, p(st)))
timeit("fwd, outputs mirrored to pinned host", lambda: lib.ga_debug_fwd_mirrored(B, N, N, p(dx1.data_ptr()), p(dx2.data_ptr()), p(dd1.data_ptr()), p(di1.data_ptr()), p(dd2.data_ptr()), p(di2.data_ptr()), p(d1.data_ptr()), p(i1.data_ptr()), p(d2.data_ptr()), p(i2.data_ptr()), p(st)))
timeit("bwd, device outputs", lambda: lib.ga_nn_distance_bwd(B, N, N, p(dx1.data_ptr()), p(dx2.data_ptr()), p(dg.data_ptr()), p(di1.data_ptr()), p(dg.data_ptr()), p(di2.data_ptr()), p(do1.data_ptr()), p(do2.data_ptr()), p(st)))
timeit("bwd, outputs straight to pinned host", lambda: lib.ga_nn_distance_bwd(B, N, N, p(dx1.data_ptr()), p(dx2.data_ptr()), p(dg.data_ptr()), p(di1.data_ptr()), p(dg.data_ptr()), p(di2.data_ptr()), p(o1.data_ptr()), p(o2.data_ptr()), p(st)))
timeit("cudaMemcpyAsync 1.23 MB d2h", lambda: o1.copy_(do1, non_blocking=True))
t0 = time.perf_counter()
for _ in range(200): torch.cuda.synchronize()
print("empty synchronize: %.1f us" % ((time.perf_counter() - t0) / 200 * 1e6))
