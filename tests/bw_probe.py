import torch, time
torch.cuda.init()
n = 1<<30  # 1 Gi elements bf16 = 2 GiB
a = torch.empty(n, dtype=torch.bfloat16, device='cuda'); b = torch.empty_like(a)
def t(f, reps=10):
    f(); torch.cuda.synchronize()
    e0=torch.cuda.Event(enable_timing=True); e1=torch.cuda.Event(enable_timing=True)
    best=1e9
    for _ in range(reps):
        e0.record(); f(); e1.record(); torch.cuda.synchronize(); best=min(best,e0.elapsed_time(e1))
    return best
ms=t(lambda: b.copy_(a)); print('copy   %.1f GB/s (r+w)'%(2*a.numel()*2/ms/1e6))
ms=t(lambda: a.zero_()); print('memset %.1f GB/s (w)'%(a.numel()*2/ms/1e6))
ms=t(lambda: a.fill_(1.5)); print('fill   %.1f GB/s (w)'%(a.numel()*2/ms/1e6))
ms=t(lambda: a.sum()); print('sum    %.1f GB/s (r)'%(a.numel()*2/ms/1e6))
c = torch.empty(n//2, dtype=torch.bfloat16, device='cuda')
ms=t(lambda: torch.add(a[:n//2], 1.0, out=c)); print('add    %.1f GB/s (r+w)'%(2*c.numel()*2/ms/1e6))
