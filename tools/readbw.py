import torch, time
x = torch.randn(13*(1<<22), dtype=torch.float64, device="cuda")
flush = torch.empty(256*1024*1024//4, dtype=torch.float32, device="cuda")
for fn, name, nbytes in ((lambda: x.sum(), "sum", x.numel()*8), (lambda: x.abs().max(), "absmax(2 kernels)", x.numel()*8*3), (lambda: torch.max(x), "max", x.numel()*8)):
    fn(); torch.cuda.synchronize()
    ts=[]
    for _ in range(5):
        flush.zero_()
        e0=torch.cuda.Event(enable_timing=True); e1=torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); e1.synchronize(); ts.append(e0.elapsed_time(e1))
    ms=sorted(ts)[len(ts)//2]
    print(name, ms, "ms", nbytes/ms/1e6, "GB/s")
y = torch.empty_like(x)
ts=[]
for _ in range(5):
    flush.zero_(); e0=torch.cuda.Event(enable_timing=True); e1=torch.cuda.Event(enable_timing=True)
    e0.record(); y.copy_(x); e1.record(); e1.synchronize(); ts.append(e0.elapsed_time(e1))
ms=sorted(ts)[2]; print("copy", ms, 2*x.numel()*8/ms/1e6, "GB/s")
