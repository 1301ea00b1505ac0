"""Measure plain HBM write and copy bandwidth on this GPU (context for the roofline numbers)."""
import torch
dev = torch.device("cuda", 0)
n = 2_319_777_792 // 4   # floats written per bench step (64 x 36.25 MB)
x = torch.empty(n, dtype=torch.float32, device=dev)
y = torch.empty(n, dtype=torch.float32, device=dev)
def timeit(fn, reps=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        best = min(best, a.elapsed_time(b))
    return best
t = timeit(lambda: x.zero_())
print(f"write-only  zero_ : {n*4/t/1e6:8.1f} GB/s  ({t:.3f} ms for {n*4/1e9:.2f} GB)")
t = timeit(lambda: x.fill_(1.5))
print(f"write-only  fill_ : {n*4/t/1e6:8.1f} GB/s")
t = timeit(lambda: y.copy_(x))
print(f"copy (r+w bytes)  : {2*n*4/t/1e6:8.1f} GB/s")
