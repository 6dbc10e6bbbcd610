"""PCIe copy characteristics on the box (development tool)."""
import torch, time
dev = torch.device("cuda:0")
def t_copy(nbytes, direction, reps=50):
    h = torch.empty(nbytes, dtype=torch.uint8).pin_memory(); d = torch.empty(nbytes, dtype=torch.uint8, device=dev)
    for _ in range(5):
        (d.copy_(h, non_blocking=True) if direction == "h2d" else h.copy_(d, non_blocking=True))
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        (d.copy_(h, non_blocking=True) if direction == "h2d" else h.copy_(d, non_blocking=True))
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps
for nb in (4096, 65536, 409600, 1228800, 4 << 20, 64 << 20):
    a, b = t_copy(nb, "h2d"), t_copy(nb, "d2h")
    print("%9d B  h2d %.4f ms (%.1f GB/s)   d2h %.4f ms (%.1f GB/s)" % (nb, a, nb / a / 1e6, b, nb / b / 1e6))
# zero-copy kernel: device reads pinned host memory directly
h = torch.empty(4 << 20, dtype=torch.uint8).pin_memory()
d = torch.empty(4 << 20, dtype=torch.uint8, device=dev)
import ctypes
hv = torch.empty(1 << 20, dtype=torch.float32).pin_memory()
# bidirectional overlap on two streams
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
h1 = torch.empty(4 << 20, dtype=torch.uint8).pin_memory(); d1 = torch.empty(4 << 20, dtype=torch.uint8, device=dev)
h2 = torch.empty(4 << 20, dtype=torch.uint8).pin_memory(); d2 = torch.empty(4 << 20, dtype=torch.uint8, device=dev)
torch.cuda.synchronize(); t0 = time.perf_counter()
for _ in range(50):
    with torch.cuda.stream(s1): d1.copy_(h1, non_blocking=True)
    with torch.cuda.stream(s2): h2.copy_(d2, non_blocking=True)
torch.cuda.synchronize(); dt = (time.perf_counter() - t0) / 50
print("4 MiB h2d + 4 MiB d2h concurrently: %.4f ms per pair" % (dt * 1e3))
