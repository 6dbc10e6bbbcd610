"""Does clock sampling perturb kernel timing?  none / nvidia-smi -lms / pynvml thread."""
import ctypes, os, subprocess, sys, threading, time
import torch
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
def measure(tag, reps=3000):
    for _ in range(10): fwd()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fwd()
    e1.record(); torch.cuda.synchronize()
    print("%-40s %.4f ms/launch" % (tag, e0.elapsed_time(e1) / reps), flush=True)
measure("no sampler")
measure("no sampler (again)")
pr = subprocess.Popen(["nvidia-smi", "-i", "0", "--query-gpu=clocks.sm,power.draw", "--format=csv,noheader,nounits", "-lms", "100"], stdout=subprocess.DEVNULL)
time.sleep(0.5); measure("nvidia-smi clocks.sm,power -lms 100"); pr.terminate(); pr.wait()
pr = subprocess.Popen(["nvidia-smi", "-i", "0", "--query-gpu=clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap", "--format=csv,noheader,nounits", "-lms", "100"], stdout=subprocess.DEVNULL)
time.sleep(0.5); measure("nvidia-smi full query -lms 100"); pr.terminate(); pr.wait()
pr = subprocess.Popen(["nvidia-smi", "-i", "0", "--query-gpu=clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap", "--format=csv,noheader,nounits", "-lms", "1000"], stdout=subprocess.DEVNULL)
time.sleep(0.5); measure("nvidia-smi full query -lms 1000"); pr.terminate(); pr.wait()
time.sleep(0.5); measure("after sampler stopped")
import pynvml
pynvml.nvmlInit(); h = pynvml.nvmlDeviceGetHandleByIndex(0)
stop = threading.Event(); got = []
def poll():
    while not stop.is_set():
        got.append((pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM), pynvml.nvmlDeviceGetCurrentClocksEventReasons(h) if hasattr(pynvml, "nvmlDeviceGetCurrentClocksEventReasons") else pynvml.nvmlDeviceGetCurrentClocksThrottleReasons(h)))
        time.sleep(0.05)
th = threading.Thread(target=poll); th.start(); measure("pynvml thread, 50 ms"); stop.set(); th.join()
print("pynvml samples", len(got), got[:3])
measure("no sampler (end)")
