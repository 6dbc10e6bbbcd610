"""Forward kernel variants by batch size (development tool): HMMA grid kernel (20) vs the tcgen05 kernel (22, G=2/4)."""
import ctypes, os, sys, json
import numpy as np
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from geometric_adv_b200 import _lib  # noqa: E402
lib = _lib.load()
dev = torch.device("cuda:0"); p = ctypes.c_void_p; st = torch.cuda.current_stream().cuda_stream
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
shapes = [(50, 2048, 2048), (512, 2048, 2048), (10, 2048, 2048), (50, 1024, 1024), (100, 2048, 2048)]
if len(sys.argv) > 1:
    shapes = [tuple(int(x) for x in s.split("x")) for s in sys.argv[1:]]
res = {}
for (B, N, M) in shapes:
    g = torch.Generator().manual_seed(2)
    x1 = (torch.rand(B, N, 3, generator=g) - 0.5).to(dev); x2 = (torch.rand(B, M, 3, generator=g) - 0.5).to(dev)
    d1 = torch.empty(B, N, device=dev); i1 = torch.empty(B, N, dtype=torch.int32, device=dev)
    d2 = torch.empty(B, M, device=dev); i2 = torch.empty(B, M, dtype=torch.int32, device=dev)
    ref = None
    only = os.environ.get("TUNE_ONLY")
    ROT = int(os.environ.get("ROTATE", "0"))  # >0: ROT input/output sets walked round-robin (total > L2) instead of the flush
    if ROT:
        xs1 = [x1.clone() for _ in range(ROT)]; xs2 = [x2.clone() for _ in range(ROT)]
        ds1 = [torch.empty_like(d1) for _ in range(ROT)]; is1 = [torch.empty_like(i1) for _ in range(ROT)]
        ds2 = [torch.empty_like(d2) for _ in range(ROT)]; is2 = [torch.empty_like(i2) for _ in range(ROT)]
        rot = [0]
    for name, keys in (("default", {}), ("hmma", {0: 20}), ("hmma_frame", {0: 20, 25: 1}), ("tc", {0: 22}), ("ws", {0: 24}), ("ws12", {0: 24, 22: 12}), ("ws_s2", {0: 24, 23: 2000}), ("ws_s4", {0: 24, 23: 4000}), ("ws_s6", {0: 24, 23: 6000}), ("ws_s8", {0: 24, 23: 8000}), ("tc_dev", {0: 22, 20: 16}), ("tc_norefine", {0: 22, 20: 1}), ("tc_nohelper", {0: 22, 20: 9}), ("tc_neither", {0: 22, 20: 3}), ("tc_nohelper_nodrain", {0: 22, 20: 11})):
        if only and name not in only.split(','):
            continue
        if (keys.get(20, 0) & 8) and B > 100:
            continue  # no helper: the stager would wait for ever on a third cloud
        for k, v in keys.items():
            lib.ga_set_tuning(k, v)
        def call():
            if ROT:
                r = rot[0] = (rot[0] + 1) % ROT
                _lib.check(lib.ga_nn_distance_fwd(B, N, M, p(xs1[r].data_ptr()), p(xs2[r].data_ptr()), p(ds1[r].data_ptr()), p(is1[r].data_ptr()),
                                                  p(ds2[r].data_ptr()), p(is2[r].data_ptr()), 0, p(st)))
                return
            _lib.check(lib.ga_nn_distance_fwd(B, N, M, p(x1.data_ptr()), p(x2.data_ptr()), p(d1.data_ptr()), p(i1.data_ptr()),
                                              p(d2.data_ptr()), p(i2.data_ptr()), 0, p(st)))
        try:
            for _ in range(3):
                call()
            torch.cuda.synchronize()
            ts = []
            for _ in range(20 if not ROT else 3 * ROT):
                if not ROT:
                    flush.zero_()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(); call(); e1.record(); torch.cuda.synchronize()
                ts.append(e0.elapsed_time(e1) * 1e3)
            out = (d1.clone(), i1.clone(), d2.clone(), i2.clone()) if not ROT else (ds1[rot[0]].clone(), is1[rot[0]].clone(), ds2[rot[0]].clone(), is2[rot[0]].clone())
            if ref is None:
                ref = out
            same = all(torch.equal(a, b) for a, b in zip(out, ref))
            res["%dx%dx%d %s" % (B, N, M, name)] = {"min_us": min(ts), "med_us": float(np.median(ts)), "equal_to_hmma": same}
            print(B, N, M, name, "min %.1f med %.1f us" % (min(ts), float(np.median(ts))), "same bits" if same else "DIFFERENT", flush=True)
        except Exception as e:  # noqa: BLE001
            print(B, N, M, name, "failed:", e, flush=True)
        for k in keys:
            lib.ga_set_tuning(k, 2 if k == 25 else 0)
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(res, open(os.path.join(ROOT, "gpurun_out", "tune_fwd.json"), "w"), indent=1)
