import time, torch
x = torch.empty(64 * 1024 * 1024 // 4, dtype=torch.float32).pin_memory()
d = torch.empty_like(x, device="cuda")
for _ in range(3):
    d.copy_(x, non_blocking=True); torch.cuda.synchronize()
t = time.perf_counter()
for _ in range(10):
    d.copy_(x, non_blocking=True)
torch.cuda.synchronize()
dt = (time.perf_counter() - t) / 10
print(f"H2D pinned 64 MiB: {dt*1e3:.2f} ms -> {x.numel()*4/dt/1e9:.1f} GB/s")
t = time.perf_counter()
for _ in range(10):
    x.copy_(d, non_blocking=True)
torch.cuda.synchronize()
dt = (time.perf_counter() - t) / 10
print(f"D2H pinned 64 MiB: {dt*1e3:.2f} ms -> {x.numel()*4/dt/1e9:.1f} GB/s")
