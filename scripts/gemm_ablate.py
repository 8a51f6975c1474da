import sys, os, json
sys.path.insert(0, '/root/repo')
import torch
from video_vae_b200 import _ffi, ops
_ffi.require_device()
g = torch.Generator(device="cuda").manual_seed(0)
for (M,N,K,tB) in [(32768,1536,768,False),(32768,768,1536,True),(32768,3072,768,False)]:
    A = torch.randn(M,K,device="cuda",generator=g).bfloat16()
    B = torch.randn((N,K) if tB else (K,N),device="cuda",generator=g).bfloat16()
    out = torch.empty(M,N,device="cuda",dtype=torch.bfloat16)
    for bits,what in ((0,"full"),(1,"no A loads"),(2,"no B loads"),(3,"no loads")):
        _ffi.lib.vvae_debug_set(10,bits)
        ops.gemm(A,B,transB=tB,out=out); torch.cuda.synchronize()
        e0,e1=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5): ops.gemm(A,B,transB=tB,out=out)
        e1.record(); torch.cuda.synchronize()
        ms=e0.elapsed_time(e1)/5
        print(json.dumps({"M":M,"N":N,"K":K,"ablation":what,"ms":round(ms,4),"tflops":round(2.0*M*N*K/ms/1e9,1)}),flush=True)
    _ffi.lib.vvae_debug_set(10,0)
