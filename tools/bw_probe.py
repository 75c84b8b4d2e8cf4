import torch
dev=torch.device('cuda')
def t(fn, n=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0,e1=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1)/n*1e-3
for mb in (84, 280, 1120, 4480):
    x=torch.empty(mb*250_000, dtype=torch.float32, device=dev).normal_()
    s=t(lambda: x.sum())
    y=torch.empty_like(x)
    c=t(lambda: y.copy_(x))
    print(f"{mb} MB: read-only sum {mb*1e6/s/1e12:.2f} TB/s ({s*1e6:.1f} us)   copy {2*mb*1e6/c/1e12:.2f} TB/s")
