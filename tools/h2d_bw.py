"""Pinned host -> device copy bandwidth for the bench's per-step input size (the floor of the e2e number)."""
import time, torch
n = 13879296
h = torch.empty(n, dtype=torch.uint8).pin_memory()
d = torch.empty(n, dtype=torch.uint8, device="cuda")
for _ in range(5):
    d.copy_(h, non_blocking=True)
torch.cuda.synchronize()
t = time.perf_counter()
for _ in range(100):
    d.copy_(h, non_blocking=True)
torch.cuda.synchronize()
dt = (time.perf_counter() - t) / 100
print("H2D %.1f MB: %.3f ms, %.1f GB/s" % (n / 1e6, dt * 1e3, n / dt / 1e9))
