"""Soak of the two default paths that synchronise through flags / atomics: 1500 host steps with fresh contents through the
pulled-ingest pipeline (every 25th checked against the device entry points) and 200 repeats of the slab kNN kernel on
one input.  Development tool; output recorded in profiles/r02_stress.txt."""
import sys, ctypes, numpy as np, torch
sys.path.insert(0,'.')
import geometric_adv_b200 as ga
from geometric_adv_b200 import _lib
lib=_lib.load(); lib.ga_debug_host_streamed.restype=ctypes.c_int
p=ctypes.c_void_p
b,n,m=50,2048,2048
buf=[torch.empty(b,n,3).pin_memory(),torch.empty(b,m,3).pin_memory(),torch.empty(b,n).pin_memory(),torch.empty(b,m).pin_memory(),
     torch.empty(b,n).pin_memory(),torch.empty(b,n,dtype=torch.int32).pin_memory(),torch.empty(b,m).pin_memory(),
     torch.empty(b,m,dtype=torch.int32).pin_memory(),torch.empty(b,n,3).pin_memory(),torch.empty(b,m,3).pin_memory()]
g=torch.Generator().manual_seed(0)
bad=0; states={}
for it in range(1500):
    for x in buf[:2]: x.copy_(torch.rand(x.shape,generator=g)-0.5)
    for x in buf[2:4]: x.copy_(torch.randn(x.shape,generator=g))
    for o in buf[4:]: o.fill_(-7)
    _lib.check(lib.ga_nn_distance_fwd_bwd_host(b,n,m,*[p(x.data_ptr()) for x in buf],0))
    s=lib.ga_debug_host_streamed(); states[s]=states.get(s,0)+1
    if it%25==0:
        a,c=buf[0].cuda(),buf[1].cuda()
        dev=ga.nn_distance(a,c); gr=ga.nn_distance_grad(a,c,buf[2].cuda(),dev[1],buf[3].cuda(),dev[3])
        for x,y in zip(buf[4:],tuple(dev)+tuple(gr)):
            if not torch.equal(x,y.cpu()): bad+=1
print("host step stress: iterations 1500, pipelines",states,"mismatching arrays",bad)
# slab kNN determinism: same input 200 times
pc=(torch.rand(64,2048,3,generator=g)-0.5).cuda()
ref=ga.knn_dists(pc,10); diff=0
for _ in range(200):
    if not torch.equal(ga.knn_dists(pc,10),ref): diff+=1
print("knn slab repeat: kernel",lib.ga_last_kernel().decode(),"differing runs",diff)
