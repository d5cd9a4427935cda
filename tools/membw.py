import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, numpy as np
dev="cuda:0"
def t(fn, n=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    ev=[torch.cuda.Event(enable_timing=True) for _ in range(n+1)]
    for i in range(n):
        ev[i].record(); fn()
    ev[n].record(); torch.cuda.synchronize()
    return float(np.median([ev[i].elapsed_time(ev[i+1]) for i in range(n)]))
N=249_250_621
a=torch.empty(N,dtype=torch.int32,device=dev); b=torch.empty(N,dtype=torch.int32,device=dev)
r=torch.empty(180_000_000,dtype=torch.int32,device=dev)
ms=t(lambda: a.fill_(1)); print(f"fill 997MB: {ms:.3f} ms  {N*4/ms/1e6:.0f} GB/s")
ms=t(lambda: b.copy_(a)); print(f"copy 997MB: {ms:.3f} ms  {2*N*4/ms/1e6:.0f} GB/s")
ms=t(lambda: r.sum()); print(f"read 720MB (sum): {ms:.3f} ms  {r.numel()*4/ms/1e6:.0f} GB/s")
ms=t(lambda: (r.sum(), a.fill_(1))); print(f"read720+fill997 serial: {ms:.3f} ms")
big=torch.empty(1<<30,dtype=torch.bfloat16,device=dev); big2=torch.empty_like(big)
ms=t(lambda: big2.copy_(big)); print(f"copy 2GiB bf16: {ms:.3f} ms {2*big.numel()*2/ms/1e6:.0f} GB/s")
