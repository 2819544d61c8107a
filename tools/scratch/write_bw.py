import torch
x = torch.empty(10*1024**3//4, dtype=torch.float32, device="cuda")
for name, fn in (("zero_", lambda: x.zero_()), ("fill_", lambda: x.fill_(1.0))):
    for _ in range(3): fn()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); s.record()
    for _ in range(10): fn()
    e.record(); torch.cuda.synchronize()
    ms = s.elapsed_time(e) / 10
    print(name, "ms", ms, "GB/s", x.numel()*4/ms/1e6)
