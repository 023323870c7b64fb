import torch
x = torch.empty(2147483648 // 8, dtype=torch.float64, device="cuda")
y = torch.empty_like(x)
def t(fn, n=5):
    fn(); torch.cuda.synchronize()
    ts=[]
    for _ in range(n):
        e0,e1=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
    return min(ts)
ms=t(lambda: x.fill_(1.5)); print(f"fill 2.15 GB: {ms:.3f} ms {2.147/ms*1e3:.0f} GB/s write-only")
ms=t(lambda: x.zero_()); print(f"memset 2.15 GB: {ms:.3f} ms {2.147/ms*1e3:.0f} GB/s write-only")
ms=t(lambda: y.copy_(x)); print(f"copy 2.15 GB: {ms:.3f} ms {2*2.147/ms*1e3:.0f} GB/s read+write")
ms=t(lambda: x.sum()); print(f"sum 2.15 GB: {ms:.3f} ms {2.147/ms*1e3:.0f} GB/s read-only")
