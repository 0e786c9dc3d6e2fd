import torch, time
x = torch.empty(1<<28, dtype=torch.float32, pin_memory=True)   # 1 GiB
d = torch.empty_like(x, device="cuda")
for _ in range(2): d.copy_(x, non_blocking=True); torch.cuda.synchronize()
t=time.perf_counter(); 
for _ in range(5): d.copy_(x, non_blocking=True)
torch.cuda.synchronize(); dt=time.perf_counter()-t
print("H2D GB/s", 5*x.numel()*4/dt/1e9)
t=time.perf_counter(); 
for _ in range(5): x.copy_(d, non_blocking=True)
torch.cuda.synchronize(); dt=time.perf_counter()-t
print("D2H GB/s", 5*x.numel()*4/dt/1e9)
s1=torch.cuda.Stream(); s2=torch.cuda.Stream()
y = torch.empty(1<<28, dtype=torch.float32, pin_memory=True); e = torch.empty_like(y, device="cuda")
torch.cuda.synchronize(); t=time.perf_counter()
for _ in range(5):
    with torch.cuda.stream(s1): d.copy_(x, non_blocking=True)
    with torch.cuda.stream(s2): y.copy_(e, non_blocking=True)
torch.cuda.synchronize(); dt=time.perf_counter()-t
print("bidirectional: each GB/s", 5*x.numel()*4/dt/1e9)
