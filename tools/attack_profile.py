"""Run a few eager attack iterations (for an ncu launch list).  usage: attack_profile.py [B] [mode]"""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from geometric_adv_b200.attack import GeometricAttack, PointNetAE  # noqa: E402
B = int(sys.argv[1]) if len(sys.argv) > 1 else 10
kw = {}
if len(sys.argv) > 2:
    kw["single_forward"] = sys.argv[2] == "single"
torch.manual_seed(0)
atk = GeometricAttack(PointNetAE(2048), B, 2048, use_cuda_graph=False, **kw)
g = torch.Generator().manual_seed(0)
atk.x.copy_((torch.rand(B, 2048, 3, generator=g) - 0.5).cuda())
atk.gt.copy_((torch.rand(B, 2048, 3, generator=g) - 0.5).cuda())
atk.init_pert()
for _ in range(4):
    atk.step()
torch.cuda.synchronize()
# cudaProfilerStart/Stop around ONE iteration (ncu --profile-from-start off): an NVTX range would miss the kernels
# the autograd engine launches from its own thread
torch.cuda.profiler.start()
atk.step()
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("done")
