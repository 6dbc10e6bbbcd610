"""All-pairs directed Chamfer: tensor-core scan vs fp32 filter scan (ga_set_tuning key 16).  Development tool."""
import ctypes, os, sys, json
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from geometric_adv_b200 import _lib  # noqa: E402
lib = _lib.load()
dev = torch.device("cuda:0"); p = ctypes.c_void_p; st = torch.cuda.current_stream().cuda_stream
out = {}
for (S, rows, n) in [(512, 128, 2048), (2000, 250, 2048), (512, 128, 1000)]:
    g = torch.Generator().manual_seed(3)
    clouds = (torch.rand(S, n, 3, generator=g) - 0.5).to(dev)
    res = {}
    mats = {}
    for name, key in (("mma", 0), ("fp32", 1)):
        lib.ga_set_tuning(16, key)
        o = torch.empty(rows, S, device=dev)
        call = lambda: _lib.check(lib.ga_chamfer_all_pairs_directed(S, n, p(clouds.data_ptr()), 0, rows, p(o.data_ptr()), 0, p(st)))
        call(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); call(); e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        res[name] = {"ms": ms, "evals_per_s": rows * S * float(n) * n / (ms * 1e-3)}
        mats[name] = o.clone()
        assert bool(torch.isfinite(o).all()), "non-finite all-pairs output (%s, S=%d)" % (name, S)
    lib.ga_set_tuning(16, 0)
    ok = mats["fp32"] > 0  # the diagonal (a cloud against itself) is exactly 0 in both
    res["max_rel_diff_between_kernels"] = float(((mats["mma"] - mats["fp32"]).abs()[ok] / mats["fp32"][ok]).max())
    res["nonpositive_entries"] = int((~ok).sum())
    print("torch", torch.__version__, "ok count", int(ok.sum()), "of", ok.numel())
    out["S%d_rows%d_n%d" % (S, rows, n)] = res
    print(S, rows, n, res, flush=True)
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(out, open(os.path.join(ROOT, "gpurun_out", "tune_pairs.json"), "w"), indent=1)
