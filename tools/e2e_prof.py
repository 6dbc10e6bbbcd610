"""Six host steps for an ncu launch list of the replayed pipeline (development tool; profiles/r02_e2e_kernel_durations.txt)."""
import sys, ctypes, torch
sys.path.insert(0,'.')
from geometric_adv_b200 import _lib
lib=_lib.load(); p=ctypes.c_void_p
b,n,m=50,2048,2048
buf=[(torch.rand(b,n,3)-0.5).pin_memory(),(torch.rand(b,m,3)-0.5).pin_memory(),torch.rand(b,n).pin_memory(),torch.rand(b,m).pin_memory(),
     torch.empty(b,n).pin_memory(),torch.empty(b,n,dtype=torch.int32).pin_memory(),torch.empty(b,m).pin_memory(),
     torch.empty(b,m,dtype=torch.int32).pin_memory(),torch.empty(b,n,3).pin_memory(),torch.empty(b,m,3).pin_memory()]
for _ in range(6):
    _lib.check(lib.ga_nn_distance_fwd_bwd_host(b,n,m,*[p(x.data_ptr()) for x in buf],0))
print("done")
