"""Wall-clock timing of ga_nn_distance_fwd_bwd_host (the e2e leg of bench.py): direct two-lane path
vs the replayed CUDA-graph pipeline with 1..8 chunks.  Development tool; writes gpurun_out/tune_e2e.json."""
import ctypes
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from geometric_adv_b200 import _lib  # noqa: E402

lib = _lib.load()
p = ctypes.c_void_p
out = {}
for (B, N, M) in [(50, 2048, 2048), (10, 2048, 2048), (200, 2048, 2048)]:
    rng = np.random.default_rng(0)
    mk = lambda *s: torch.from_numpy((rng.random(s, dtype=np.float32) - 0.5).astype(np.float32)).pin_memory()
    bufs = [mk(B, N, 3), mk(B, M, 3), mk(B, N), mk(B, M), torch.empty(B, N).pin_memory(),
            torch.empty(B, N, dtype=torch.int32).pin_memory(), torch.empty(B, M).pin_memory(),
            torch.empty(B, M, dtype=torch.int32).pin_memory(), torch.empty(B, N, 3).pin_memory(),
            torch.empty(B, M, 3).pin_memory()]
    args = [p(x.data_ptr()) for x in bufs]

    def step():
        _lib.check(lib.ga_nn_distance_fwd_bwd_host(B, N, M, *args, 0))

    key = "b%d" % B
    out[key] = {}
    for name, k10, k11, k17, k26, k27 in ([("direct", 1, 0, 0, 0, 0)] +
                                          [("graph_c%d_mirror_dist_idx" % c, 0, c, 2, 0, 0) for c in (1, 2)] +
                                          [("graph_c%d_copy_dist_idx" % c, 0, c, 1, 0, 0) for c in (1, 2)] +
                                          [("streamed_g%d_mirror_dist_idx" % g, 0, 0, 2, g, 0) for g in (1, 2, 4)] +
                                          [("pulled_c%d_mirror_dist_idx" % c, 0, 0, 2, 0, c) for c in (2, 4, 8, 12, 16, 24, 32, 50)] +
                                          [("pulled_c%d_copy_dist_idx" % c, 0, 0, 1, 0, c) for c in (8, 16, 32)] +
                                          [("graph_auto", 0, 0, 0, 0, 0)]):
        lib.ga_set_tuning(10, k10)
        lib.ga_set_tuning(11, k11)
        lib.ga_set_tuning(17, k17)
        lib.ga_set_tuning(26, k26)
        lib.ga_set_tuning(27, k27)
        for _ in range(5):
            step()
        ts = []
        for _ in range(60):
            t0 = time.perf_counter()
            step()
            ts.append(time.perf_counter() - t0)
        ts.sort()
        out[key][name] = {"min_us": ts[0] * 1e6, "med_us": ts[len(ts) // 2] * 1e6}
        out[key][name]["streamed"] = lib.ga_debug_host_streamed()
        print(key, name, "min %.1f us  med %.1f us  streamed %d" % (ts[0] * 1e6, ts[len(ts) // 2] * 1e6,
                                                                     out[key][name]["streamed"]), flush=True)
    lib.ga_set_tuning(10, 0)
    lib.ga_set_tuning(11, 0)
    lib.ga_set_tuning(17, 0)
    lib.ga_set_tuning(26, 0)
    lib.ga_set_tuning(27, 0)
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(out, open(os.path.join(ROOT, "gpurun_out", "tune_e2e.json"), "w"), indent=1)
