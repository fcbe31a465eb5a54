import torch
def t(fn, n=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    a,b=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b)/n
N=512*1024*1024
x=torch.empty(N,device='cuda'); y=torch.empty(N,device='cuda')
ms=t(lambda: x.fill_(1.0)); print("fill  write GB/s", N*4/ms/1e6)
ms=t(lambda: y.copy_(x)); print("copy  r+w  GB/s", 2*N*4/ms/1e6)
ms=t(lambda: x.sum()); print("sum   read GB/s", N*4/ms/1e6)
z=torch.empty(N//4,device='cuda')
ms=t(lambda: torch.add(x[:N//4], y[:N//4], out=z)); print("add 2r1w GB/s", 3*(N//4)*4/ms/1e6)
