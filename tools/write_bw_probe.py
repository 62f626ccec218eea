import torch, time
x = torch.empty(577*1024*1024//4, dtype=torch.float32, device="cuda")
for fn,name in ((lambda: x.zero_(),"memset 577MB"),(lambda: x.fill_(1.5),"fill 577MB")):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0,e1=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20): fn()
    e1.record(); torch.cuda.synchronize()
    ms=e0.elapsed_time(e1)/20
    print(name, ms*1e3, "us", x.numel()*4/ms/1e6, "GB/s")
y = torch.empty_like(x)
for _ in range(3): y.copy_(x)
torch.cuda.synchronize()
e0,e1=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(20): y.copy_(x)
e1.record(); torch.cuda.synchronize()
ms=e0.elapsed_time(e1)/20
print("copy", ms*1e3, "us", 2*x.numel()*4/ms/1e6, "GB/s (r+w)")
