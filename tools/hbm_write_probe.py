import torch
dev=torch.device("cuda:0")
n=128*128*82*1376
x=torch.empty(n,dtype=torch.float32,device=dev)
for f in (lambda: x.fill_(1.0), lambda: x.zero_()):
    for _ in range(3): f()
    torch.cuda.synchronize()
    e0,e1=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5): f()
    e1.record(); torch.cuda.synchronize()
    ms=e0.elapsed_time(e1)/5
    print(f"fill {n*4/2**30:.2f} GiB: {ms:.3f} ms  {n*4/ms/1e6:.0f} GB/s")
y=torch.empty_like(x)
for _ in range(3): y.copy_(x)
torch.cuda.synchronize()
e0,e1=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(5): y.copy_(x)
e1.record(); torch.cuda.synchronize()
ms=e0.elapsed_time(e1)/5
print(f"copy: {ms:.3f} ms  {2*n*4/ms/1e6:.0f} GB/s (r+w)")
