"""Timing-only probe: cost-volume backward with the atomics compiled out (debug knob cv_bwd_nored) — how much of the kernel is atomics."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, panogrf_b200 as pg
from panogrf_b200 import _lib
lib = _lib.load()
dev="cuda"
B,S,H,W,C,D=1,2,256,512,32,64
g = torch.Generator(device=dev).manual_seed(0)
images = torch.randn(B,S,H,W,C,device=dev,generator=g,requires_grad=True)
rots = torch.eye(3,device=dev).expand(B,S,3,3).contiguous()
trans = torch.zeros(B,S,3,device=dev); trans[:,0,2]=0.5; trans[:,1,2]=-0.5
depths = torch.linspace(0.1,10,D,device=dev)
args={"dataset_name":"m3d","contain_dnet":False,"mono_uncertainty":False}
out = pg.calculate_cost_volume_erp(args, images, depths, trans, rots)
gout = torch.randn(out.shape, device=dev, generator=g)
flush = torch.empty(256*1024*1024//4, device=dev)
for nored in (0,1):
  for label, variant, L, minb in (("lane",0,0,8),("run8",1,8,4)):
    lib.pgrf_debug_set(b"cv_bwd_variant",variant); lib.pgrf_debug_set(b"cv_bwd_run",L); lib.pgrf_debug_set(b"cv_bwd_minb",minb); lib.pgrf_debug_set(b"cv_bwd_nored",nored)
    f=lambda: torch.autograd.grad(out, images, gout, retain_graph=True)[0]
    for _ in range(2): f()
    ts=[]
    for _ in range(5):
        flush.zero_(); e0,e1=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
        e0.record(); f(); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
    print("nored",nored,label,round(sorted(ts)[2],4))
# cost of the zero-fill + autograd wrapper: time an empty-ish call
lib.pgrf_debug_set(b"cv_bwd_nored",0)
