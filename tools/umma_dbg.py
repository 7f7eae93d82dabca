import sys; sys.path.insert(0,'/root/repo')
import torch
from panogrf_b200 import _lib
lib=_lib.load()
for variant in (0,1):
    for K,N in [(16,16),(32,32),(64,64)]:
        A=torch.randn(128,K,device='cuda'); W=torch.randn(N,K,device='cuda'); out=torch.zeros(128,N,device='cuda')
        rc=lib.pgrf_umma_selftest(_lib.ptr(A),_lib.ptr(W),_lib.ptr(out),K,N,variant,_lib.stream_ptr()); torch.cuda.synchronize()
        ref=A.bfloat16().float()@W.bfloat16().float().t()
        print('variant',variant,'K',K,'N',N,'rc',rc,'maxerr',float((out-ref).abs().max()),'ref max',float(ref.abs().max()))
