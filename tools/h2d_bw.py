import torch, time
x = torch.empty(536346624, dtype=torch.uint8).pin_memory()
d = torch.empty_like(x, device="cuda")
for _ in range(2): d.copy_(x, non_blocking=True)
torch.cuda.synchronize(); t=time.perf_counter()
for _ in range(5): d.copy_(x, non_blocking=True)
torch.cuda.synchronize(); dt=(time.perf_counter()-t)/5
print("H2D pinned 536MB: %.2f ms  %.1f GB/s" % (dt*1e3, 536346624/dt/1e9))
y = torch.empty(42893312, dtype=torch.uint8).pin_memory(); e = torch.empty_like(y, device="cuda")
torch.cuda.synchronize(); t=time.perf_counter()
for _ in range(5): y.copy_(e, non_blocking=True)
torch.cuda.synchronize(); dt=(time.perf_counter()-t)/5
print("D2H pinned 43MB: %.2f ms  %.1f GB/s" % (dt*1e3, 42893312/dt/1e9))
