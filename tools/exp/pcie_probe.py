import torch, time
h = torch.empty(1<<28, dtype=torch.float32, pin_memory=True); h.fill_(1.0)
d = torch.empty_like(h, device="cuda")
for n in (1<<28, 1<<26, 1<<23):
    for _ in range(2): d[:n].copy_(h[:n], non_blocking=True)
    torch.cuda.synchronize()
    e0=torch.cuda.Event(enable_timing=True); e1=torch.cuda.Event(enable_timing=True)
    reps = (1<<28)//n
    e0.record()
    for r in range(reps): d[r*n:(r+1)*n].copy_(h[r*n:(r+1)*n], non_blocking=True)
    e1.record(); torch.cuda.synchronize()
    print("H2D chunk", n*4/1e6, "MB:", (1<<30)/e0.elapsed_time(e1)/1e6, "GB/s")
hd = torch.empty(1<<26, dtype=torch.float32, pin_memory=True)
e0=torch.cuda.Event(enable_timing=True); e1=torch.cuda.Event(enable_timing=True)
s2 = torch.cuda.Stream()
torch.cuda.synchronize()
e0.record()
with torch.cuda.stream(s2):
    hd.copy_(d[:1<<26], non_blocking=True)
d.copy_(h, non_blocking=True)
e1.record(); torch.cuda.synchronize()
print("H2D 1GB with concurrent D2H 256MB:", (1<<30)/e0.elapsed_time(e1)/1e6, "GB/s")
