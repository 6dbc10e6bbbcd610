"""Burst vs sustained timing of the forward kernel with clock/power sampling (development tool)."""
import ctypes, os, subprocess, sys, threading, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from geometric_adv_b200 import _lib
lib = _lib.load(); dev = torch.device("cuda:0"); p = ctypes.c_void_p
st = torch.cuda.current_stream().cuda_stream
B, N = 50, 2048
x1 = (torch.rand(B, N, 3) - 0.5).to(dev); x2 = (torch.rand(B, N, 3) - 0.5).to(dev)
d1 = torch.empty(B, N, device=dev); i1 = torch.empty(B, N, dtype=torch.int32, device=dev)
d2 = torch.empty(B, N, device=dev); i2 = torch.empty(B, N, dtype=torch.int32, device=dev)
args = (B, N, N, p(x1.data_ptr()), p(x2.data_ptr()), p(d1.data_ptr()), p(i1.data_ptr()), p(d2.data_ptr()), p(i2.data_ptr()), 0, p(st))
def fwd(): lib.ga_nn_distance_fwd(*args)
rows = []
def sampler(stop):
    pr = subprocess.Popen(["nvidia-smi", "-i", "0", "--query-gpu=clocks.sm,power.draw,clocks_event_reasons.sw_power_cap,clocks_event_reasons.active", "--format=csv,noheader,nounits", "-lms", "20"], stdout=subprocess.PIPE, text=True)
    while not stop.is_set():
        line = pr.stdout.readline()
        if line: rows.append((time.perf_counter(), line.strip()))
    pr.terminate()
stop = threading.Event(); th = threading.Thread(target=sampler, args=(stop,)); th.start()
time.sleep(0.5)
for _ in range(5): fwd()
torch.cuda.synchronize()
ts = []
for _ in range(20):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); fwd(); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1)); time.sleep(0.01)
print("burst (sync + 10 ms idle between launches): min %.4f med %.4f ms" % (min(ts), sorted(ts)[10]))
t_cpu0 = time.perf_counter()
for reps in (100, 1000, 20000):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    w0 = time.perf_counter(); e0.record()
    for _ in range(reps): fwd()
    e1.record(); w_launch = time.perf_counter() - w0; torch.cuda.synchronize(); w1 = time.perf_counter()
    print("sustained %5d launches: %.4f ms/launch (GPU events), CPU launch loop %.4f ms/launch, window %.3f-%.3f" % (reps, e0.elapsed_time(e1) / reps, w_launch / reps * 1e3, w0, w1))
stop.set(); th.join()
for tm, l in rows[::10]: print("%.3f %s" % (tm, l))
