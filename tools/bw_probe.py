import torch
x=torch.empty(1<<30, dtype=torch.uint8, device='cuda'); y=torch.empty_like(x)
def t(fn,n=10):
    for _ in range(3): fn()
    torch.cuda.synchronize(); e0=torch.cuda.Event(enable_timing=True); e1=torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize(); return e0.elapsed_time(e1)/n
ms=t(lambda: x.zero_()); print("memset write GB/s", 1.0737/ms*1e3)
ms=t(lambda: y.copy_(x)); print("copy r+w GB/s", 2*1.0737/ms*1e3)
xf=x.view(torch.float32)
ms=t(lambda: xf.sum()); print("read (sum) GB/s", 1.0737/ms*1e3)
