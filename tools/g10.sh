mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/pytest_gpu.log
for pdl in 1 0; do
GA_PDL=$pdl timeout 300 python - <<'PY'
import os, ctypes, torch, sys
sys.path.insert(0, os.getcwd())
from geometric_adv_b200 import _lib
lib = _lib.load(); lib.ga_set_tuning(15, int(os.environ["GA_PDL"]))
B, N = 50, 2048
dev = torch.device("cuda:0"); p = ctypes.c_void_p; st = torch.cuda.current_stream().cuda_stream
g = torch.Generator().manual_seed(0)
x1 = (torch.rand(B, N, 3, generator=g) - 0.5).to(dev); x2 = (torch.rand(B, N, 3, generator=g) - 0.5).to(dev)
d1 = torch.empty(B, N, device=dev); i1 = torch.empty(B, N, dtype=torch.int32, device=dev)
d2 = torch.empty(B, N, device=dev); i2 = torch.empty(B, N, dtype=torch.int32, device=dev)
g1 = torch.full((B, N), 1.0 / N, device=dev); o1 = torch.empty(B, N, 3, device=dev); o2 = torch.empty(B, N, 3, device=dev)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
def step():
    lib.ga_nn_distance_fwd(B, N, N, p(x1.data_ptr()), p(x2.data_ptr()), p(d1.data_ptr()), p(i1.data_ptr()), p(d2.data_ptr()), p(i2.data_ptr()), 0, p(st))
    lib.ga_nn_distance_bwd(B, N, N, p(x1.data_ptr()), p(x2.data_ptr()), p(g1.data_ptr()), p(i1.data_ptr()), p(g1.data_ptr()), p(i2.data_ptr()), p(o1.data_ptr()), p(o2.data_ptr()), p(st))
for _ in range(5): step()
ts = []
for _ in range(50):
    flush.zero_()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); step(); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
ts.sort(); print("pdl=%s fwd+bwd step: min %.2f us med %.2f us" % (os.environ["GA_PDL"], ts[0]*1e3, ts[len(ts)//2]*1e3))
PY
done
